ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/l.csv python tools/nn_snark_time.py 32 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/l.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; st=i; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value')
seq=[(r[ki].split('(')[0].replace('void ','').replace('<unnamed>::','')[:34], float(r[vi].replace(',',''))/1000) for r in rows[st+1:] if len(r)>vi]
# last prove: find last occurrence of k_nifs_round0_small
idx=max(i for i,(k,v) in enumerate(seq) if k.startswith('k_nifs_round0'))
for k,v in seq[idx-8:idx+75]: print("%-36s %7.1f" % (k,v))
PY
