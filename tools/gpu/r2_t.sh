#!/bin/bash
# round 2, call T: compute-sanitizer on the final build (every k_bev_tma variant, BALANCE pre-passes, host pipeline) + smoke
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2t; mkdir -p $O
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/gpu/sanitize_small.py > $O/sanitize_small_$tool.log 2>&1; tail -3 $O/sanitize_small_$tool.log
done
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_memcheck.log 2>&1; tail -3 $O/smoke_memcheck.log
timeout 300 compute-sanitizer --tool initcheck python tools/gpu/sanitize_small.py > $O/sanitize_small_initcheck.log 2>&1; tail -3 $O/sanitize_small_initcheck.log
BEVK_BENCH_NO_API=1 timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 200 --warmup 5 > $O/bench.json 2> $O/bench.err; python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'traffic',d['roofline']['traffic'])"
