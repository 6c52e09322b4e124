PROMP_B200_LIB=$PWD/promp_b200/libpromp_b200_clk.so timeout 120 python tools/chain_time.py point
PROMP_B200_LIB=$PWD/promp_b200/libpromp_b200_clk.so CHAIN_QS=0,1 timeout 120 python tools/chain_time.py point 10
PROMP_B200_LIB=$PWD/promp_b200/libpromp_b200_clk.so CHAIN_QS=0 timeout 120 python tools/chain_time.py cheetah
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
