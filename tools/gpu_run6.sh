timeout 600 python -m pytest tests/test_gpu_pass.py -m gpu -x -q --timeout 200 -k "other_shapes" 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
