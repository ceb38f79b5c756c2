mkdir -p gpurun_out
SP2_TAIL_PIPE=0 python tools/sc_round_profile.py 20 > gpurun_out/r2j_round_profile.txt 2>&1
SP2_TAIL_PIPE=0 python tools/sc_clocks.py > gpurun_out/r2j_clocks.txt 2>&1
cat gpurun_out/r2j_round_profile.txt gpurun_out/r2j_clocks.txt
