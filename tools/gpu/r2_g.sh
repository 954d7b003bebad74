#!/bin/bash
# round 2, call G: full GPU suite on the final tree + the measurements that changed after call E
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2g; mkdir -p $O
echo "== gpu tests"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== cfg3 bench + launch list"
BEVK_BENCH_NO_API=1 timeout 600 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err; python -c "
import json;d=json.loads(open('$O/bench_cfg3.json').read().strip().splitlines()[-1]);print('cfg3 ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])"
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/launches_cfg3.csv python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu_cfg3.log 2>&1
echo "== camera-sharded step, halves timed on one GPU"
timeout 300 python tools/gpu/shard_breakdown.py > $O/shard_breakdown.json 2> $O/shard_breakdown.err; cat $O/shard_breakdown.json
echo "== sanitizers on the final build"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/gpu/sanitize_small.py > $O/sanitize_small_$tool.log 2>&1; tail -2 $O/sanitize_small_$tool.log
done
BEVK_TMA=0 timeout 900 compute-sanitizer --tool racecheck python - > $O/sanitize_gather_racecheck.log 2>&1 <<'PY'
import runpy, sys
sys.argv = ["sanitize_small.py"]
src = open("tools/gpu/sanitize_small.py").read().replace('assert eng.last_path() == "tma"', 'assert eng.last_path() == "gather"')
exec(compile(src, "tools/gpu/sanitize_small.py", "exec"))
PY
tail -2 $O/sanitize_gather_racecheck.log
ls $O
