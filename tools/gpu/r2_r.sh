#!/bin/bash
# round 2, call R: 64-bit-store interior write-out: full GPU suite + bench A/B against the previous commit's library
# (ab/libbevk_prev.so) and a build that posts the boxes before the LUT entries (ab/libbevk_tmafirst.so)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2r; mkdir -p $O
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
run prev BEVK_LIB_PATH=$PWD/ab/libbevk_prev.so
run tmafirst BEVK_LIB_PATH=$PWD/ab/libbevk_tmafirst.so
run repo_again
