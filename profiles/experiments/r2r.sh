#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python profiles/experiments/train_prof.py 16 > gpurun_out/train_prof.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_train.csv')) if len(r)>10 and r[0].isdigit()]
per=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:60]; val=float(r[-1]); unit=r[-2]
    us = val/1000.0 if unit in ('ns','nsecond') else val if unit in ('us','usecond') else val*1000.0
    per.setdefault(name,[]).append(us)
tot=sum(sum(v) for v in per.values())
print('total us', round(tot), 'launches', len(rows))
for k,v in sorted(per.items(), key=lambda kv:-sum(kv[1]))[:16]:
    print(f'{sum(v):10.0f} us {len(v):5d} x {sum(v)/len(v):8.1f}  {k}')
PY
