ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/l.csv python tools/sc_round_profile.py 20 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/l.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; st=i; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[st+1:][-8:]:
    print(r[ki].split('(')[0][:40], r[vi])
PY
