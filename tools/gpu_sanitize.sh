export MSA_REPS=1
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_infer.py -m gpu -x -q --timeout 800 -k "small_infer or golden or fwd_ta_mask" 2>&1 | tail -15
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_infer.py -m gpu -x -q --timeout 800 -k "test_infer_small or golden" 2>&1 | tail -15
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_pass.py -m gpu -x -q --timeout 800 -k "forward_attention" 2>&1 | tail -10
