PROMP_B200_LIB=$PWD/promp_b200/libpromp_b200_clk.so timeout 120 python tools/rollout_time.py point
PROMP_B200_LIB=$PWD/promp_b200/libpromp_b200_clk.so timeout 120 python tools/rollout_time.py cheetah
