#!/bin/bash
# round 2, call F: final kernel state -- tests, default bench, two alternative configs, launch list + full capture
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2f; mkdir -p $O
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu.log
echo "== default bench"
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f/bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value']),'api',d['e2e_api']['pinned']['ms_per_call'],d['e2e_api']['pageable']['ms_per_call'],'cpu',d['cpu_baseline']['value'])
PY
for cfg in 6144,2,2 4096,3,2; do
  BEVK_TMA_CFG=$cfg BEVK_BENCH_NO_API=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $O/bench_$cfg.json 2> $O/bench_$cfg.err
  python -c "
import json;d=json.loads(open('$O/bench_$cfg.json').read().strip().splitlines()[-1]);print('$cfg','ms/step',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4))"
done
echo "== ncu launch list + full capture"
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu.log 2>&1
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bev_tma -s 5 -c 1 -o $O/prof_tma python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu2.log 2>&1
ls $O
