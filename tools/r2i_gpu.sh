set -x
mkdir -p gpurun_out
python tools/tail_trace.py 16 > gpurun_out/r2i_trace_pipe1.txt 2>&1
SP2_TAIL_PIPE=0 python tools/tail_trace.py 16 > gpurun_out/r2i_trace_pipe0.txt 2>&1
cat gpurun_out/r2i_trace_pipe1.txt gpurun_out/r2i_trace_pipe0.txt
