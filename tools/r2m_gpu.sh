mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:tail -c 6 -o /tmp/tail python tools/sc_round_profile.py 16 > /dev/null 2>&1
ncu -i /tmp/tail.ncu-rep --page raw --csv > /tmp/tail_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/tail_raw.csv > gpurun_out/r2m_ncu_tail_summary.csv
cat gpurun_out/r2m_ncu_tail_summary.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/tail_raw.csv')))
h=rows[0]
for r in rows[-1:]:
    for k,v in zip(h,r):
        if 'issue_stalled' in k and 'per_issue_active' not in k and 'not_issued' not in k: print(k,v)
        if 'icc' in k or 'inst_cache' in k or 'l1i' in k.lower() or 'gcc' in k: print(k,v)
PY
