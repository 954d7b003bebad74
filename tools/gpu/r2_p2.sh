#!/bin/bash
# round 2, call P2: upper bound of moving the tile write-out off the consumer warps (ablation: interior write-out skipped,
# timing only) at the current stage size and at the one a second accumulator set would leave room for
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2p2; mkdir -p $O
B="timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 200 --warmup 5"
run() { # name, env...
  local name=$1; shift
  env BEVK_BENCH_NO_API=1 "$@" $B > $O/bench_$name.json 2> $O/bench_$name.err
  python - "$O/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    t=d['plan']['tma']
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'],5), 'isolated', round(d['roofline']['kernel_ms_isolated'],5), 'same', d['e2e']['matches_device_path'], 'items', t['items'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
PY
}
run repo
run skip_7936 BEVK_LIB_PATH=$PWD/ab/libbevk_skipwrite.so
run skip_5888 BEVK_LIB_PATH=$PWD/ab/libbevk_skipwrite.so BEVK_TMA_CFG=5888,2,4
run real_5888 BEVK_LIB_PATH=$PWD/ab/libbevk_c5888.so BEVK_TMA_CFG=5888,2,4
