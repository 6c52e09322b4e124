set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dataflow_chain or tensor_core or meta_gradient or deterministic" 2>&1 | tail -25
