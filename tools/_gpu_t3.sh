timeout 200 python tools/chain_time.py point
timeout 200 python tools/chain_time.py cheetah
timeout 200 python tools/chain_time.py point 10
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_trainer_end_to_end" 2>&1 | grep -v "^$" | tail -40
