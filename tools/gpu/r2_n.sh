#!/bin/bash
# round 2, call N: the repo's k_bev_tma (slimmer consumer path, FS 7936, barrier-free units for rows-only tiles): full GPU
# suite, bench A/B against the same code without the barrier rule (ab/libbevk_base5.so) and the round's first library
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2n; mkdir -p $O
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
run base5_7936 BEVK_LIB_PATH=$PWD/ab/libbevk_base5.so BEVK_TMA_CFG=7936,2,4
run repo_7680 BEVK_TMA_CFG=7680,2,4
run repo_4096 BEVK_TMA_CFG=4096,2,4
run v3 BEVK_LIB_PATH=$PWD/ab/libbevk_v3.so
run repo_again
run trace BEVK_LIB_PATH=$PWD/ab/libbevk_trace6.so BEVK_TRACE_FILE=$PWD/$O/trace.bin
python tools/gpu/trace_slots.py $O/trace.bin
