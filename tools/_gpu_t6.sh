timeout 100 python tools/chain_time.py point 2>&1 | head -14
timeout 100 python tools/chain_time.py point 20 2>&1 | head -8
timeout 100 python tools/chain_time.py point 10 2>&1 | head -8
timeout 100 python tools/chain_time.py point 5 2>&1 | head -8
timeout 100 python tools/chain_time.py cheetah 2>&1 | head -6
timeout 100 python tools/chain_time.py cheetah 10 2>&1 | head -6
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_tf_golden.py -x -q -m gpu 2>&1 | tail -3
