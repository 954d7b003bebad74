#!/bin/bash
# round 2, call Q: interior write-out as 4-pixel chunks (12 instructions per 12 bytes): full GPU suite + bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2q; mkdir -p $O
echo "== gpu tests"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $O/pytest_gpu.log
B="timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 200 --warmup 5"
run() { # name, env...
  local name=$1; shift
  env BEVK_BENCH_NO_API=1 "$@" $B > $O/bench_$name.json 2> $O/bench_$name.err
  python - "$O/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    t=d['plan']['tma']
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'],5), 'isolated', round(d['roofline']['kernel_ms_isolated'],5), 'frac', round(d['roofline']['frac'],4), 'same', d['e2e']['matches_device_path'], 'items', t['items'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
PY
}
run repo
run repo_again
BEVK_BENCH_NO_API=1 timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; python -c "import json;d=json.loads(open(\"$O/bench_cfg3.json\").read().strip().splitlines()[-1]);print(\"cfg3 ms/step\",d[\"ms_per_step\"])"
