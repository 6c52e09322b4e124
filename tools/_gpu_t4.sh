timeout 300 python tools/debug_capture.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | grep -E "CAPTURE|CALL|passed|failed" | head
