# final measurement pass after the re-entry work (1 GPU): tests, bench (both arms), ncu launch list, ncu --set full of the prove kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2zz_gpu.txt 2>&1
nproc >> gpurun_out/r2zz_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2zz_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2zz_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2zz_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2zz_bench_reference.json 2>> gpurun_out/r2zz_bench.err
SP2_NO_GATES=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2zz_launches_full_prove.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2zz_ncu_bench.log 2>&1
SP2_NO_GATES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_cubic_persist|k_cubic_mid_pipe|k_cubic_tail_pipe|k_quad_persist|k_quad_mid_pipe|k_quad_tail_pipe|k_msm_gather|k_msm_final|k_abc|k_spmv3|k_hyrax_bind)' -s 11 -c 16 -o /tmp/r2zz_prove_kernels python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2zz_ncu1.log 2>&1
ncu -i /tmp/r2zz_prove_kernels.ncu-rep --page raw --csv > /tmp/r2zz_prove_kernels_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/r2zz_prove_kernels_raw.csv > gpurun_out/r2zz_ncu_full_prove_kernels_summary.csv
python - <<'PY'
import csv, json
rows = list(csv.reader(open('/tmp/r2zz_prove_kernels_raw.csv')))
h = rows[0]; idx = {k: i for i, k in enumerate(h)}
out = {}
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
    def val(k):
        v = r[idx[k]].replace(',', ''); u = rows[1][idx[k]]
        f = float(v); return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    if name not in out:
        out[name] = {'dram_read_bytes': val('dram__bytes_read.sum'), 'dram_write_bytes': val('dram__bytes_write.sum')}
        out[name]['dram_bytes'] = out[name]['dram_read_bytes'] + out[name]['dram_write_bytes']
json.dump(out, open('gpurun_out/r2zz_ncu_traffic.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
PY
for c in 96; do SP2_NN_SIDE_CTAS=$c python tools/nn_snark_time.py 32 256 2>&1 | tail -2 >> gpurun_out/r2zz_neutronnova.log; done
cat gpurun_out/r2zz_neutronnova.log
python tools/sc_round_profile.py 20 > gpurun_out/r2zz_sc_round_profile_2p20.txt 2>&1
tail -3 gpurun_out/r2zz_tests.log; ls -la gpurun_out | tail -20
SP2_PROVE_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "sp2 prove" | tail -17 > gpurun_out/r2zz_prove_host_trace.txt
