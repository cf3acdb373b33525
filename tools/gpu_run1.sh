set -x
nvidia-smi -L > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_v2.log 2>&1; echo "pytest-exit $?" >> gpurun_out/pytest_v2.log
tail -5 gpurun_out/pytest_v2.log
timeout 300 python profiles/phase_profile.py > gpurun_out/phases_v2.txt 2>&1; echo "exit $?" >> gpurun_out/phases_v2.txt
cat gpurun_out/phases_v2.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench-exit $?"
cat gpurun_out/bench_v2.json | head -c 1500
