#!/bin/bash
# round 2, call U: final tree (warp sync after the barrier-free write-out): GPU suite, racecheck, bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2u; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck python tools/gpu/sanitize_small.py > $O/sanitize_small_racecheck.log 2>&1; tail -3 $O/sanitize_small_racecheck.log
grep -c "Warning" $O/sanitize_small_racecheck.log
BEVK_BENCH_NO_API=1 timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 200 --warmup 5 > $O/bench.json 2> $O/bench.err; python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('ms/step',d['ms_per_step'],'frac',d['roofline']['frac'])"
