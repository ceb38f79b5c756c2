SP2_NN_TRACE=1 python tools/nn_snark_time.py 32 2>&1 | tail -18
