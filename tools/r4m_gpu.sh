timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4m_bench.json 2> gpurun_out/r4m_bench.err; tail -3 gpurun_out/r4m_bench.err | cut -c1-600
