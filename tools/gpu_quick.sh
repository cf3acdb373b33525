# quick GPU check after a kernel change: parity of the pass + inference, per-kernel times of one pass
timeout 900 python -m pytest tests/test_gpu_pass.py tests/test_gpu_infer.py tests/test_gpu_properties.py tests/test_gpu_gemm_tc.py -m gpu -x -q --timeout 300 2>&1 | tail -4
timeout 300 python profiles/chain_vs_batch.py 4 2>&1 | tail -1
