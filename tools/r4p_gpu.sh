# ncu --set full of the cubic kernels of the second prove (the window of r4k_gpu.sh missed them after the launch order changed)
SP2_NO_GATES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_cubic_persist|k_cubic_mid_pipe)' -s 2 -c 2 -o /tmp/r2zz_cubic python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2zz_ncu3.log 2>&1
ncu -i /tmp/r2zz_cubic.ncu-rep --page raw --csv > /tmp/r2zz_cubic_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/r2zz_cubic_raw.csv > gpurun_out/r2zz_ncu_full_cubic_kernels_summary.csv
python - <<'PY'
import csv, json
rows = list(csv.reader(open('/tmp/r2zz_cubic_raw.csv')))
h = rows[0]; idx = {k: i for i, k in enumerate(h)}
out = {}
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
    def val(k):
        v = r[idx[k]].replace(',', ''); u = rows[1][idx[k]]
        f = float(v); return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    out[name] = {'dram_read_bytes': val('dram__bytes_read.sum'), 'dram_write_bytes': val('dram__bytes_write.sum')}
    out[name]['dram_bytes'] = out[name]['dram_read_bytes'] + out[name]['dram_write_bytes']
json.dump(out, open('gpurun_out/r2zz_ncu_traffic_cubic.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
PY
cat gpurun_out/r2zz_ncu_full_cubic_kernels_summary.csv | cut -c1-300
cuobjdump -sass -fun k_cubic_persist spartan2_b200/libspartan2_b200.so 2>/dev/null | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sort | uniq -c | sort -rn | head -25 > gpurun_out/r2zz_sass_histogram_k_cubic_persist.txt; head -12 gpurun_out/r2zz_sass_histogram_k_cubic_persist.txt
