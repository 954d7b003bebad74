#!/bin/bash
# round 2, call V: k_lum_spans with a flat work list over 16 rows per CTA: GPU suite, cfg3 bench + launch list
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2v; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/pytest_gpu.log
BEVK_BENCH_NO_API=1 timeout 600 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; python -c "
import json;d=json.loads(open('$O/bench_cfg3.json').read().strip().splitlines()[-1]);print('cfg3 ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'same',d['e2e']['matches_device_path'])"
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/launches_cfg3.csv python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu_cfg3.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2v/launches_cfg3.csv')) if len(r)>5]
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[hi]; kn=H.index('Kernel Name'); mv=H.index('Metric Value')
d=collections.OrderedDict()
for r in rows[hi+1:]: d.setdefault(r[kn][:40],[]).append(float(r[mv].replace(',','')))
for k,x in d.items(): print(k, len(x), round(sum(x)/len(x)/1000,2),'us')
PY
