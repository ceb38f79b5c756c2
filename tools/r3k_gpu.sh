SP2_NN_PIPE=1 python tools/nn_snark_time.py 32 2>&1 | grep "host round\|coef kernel" | tail -12 | cut -c1-200
