timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -x -q --timeout 200 2>&1 | tail -2
timeout 300 python profiles/gemm_tc_bench.py 2>&1 | head -5
timeout 300 python profiles/chain_vs_batch.py 4 2>&1 | tail -1
