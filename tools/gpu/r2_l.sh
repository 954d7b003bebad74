#!/bin/bash
# round 2, call L: stage size sweep of k_bev_tma (2 ring slots, 4 entry groups per slot)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2l; mkdir -p $O
B="timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 200 --warmup 5"
run() { # name, env...
  local name=$1; shift
  env BEVK_BENCH_NO_API=1 "$@" $B > $O/bench_$name.json 2> $O/bench_$name.err
  python - "$O/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    t=d['plan']['tma']
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'],5), 'isolated', round(d['roofline']['kernel_ms_isolated'],5), 'same', d['e2e']['matches_device_path'], 'items', t['items'], 'gather', t['gather_entries'], 'box', t['box_bytes'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
PY
}
L=$PWD/ab/libbevk_base4.so
for fs in 5120 5632 6144 6656 7168 7936; do run f$fs BEVK_LIB_PATH=$L BEVK_TMA_CFG=$fs,2,4; done
run f7936_m2 BEVK_LIB_PATH=$L BEVK_TMA_CFG=7936,2,4 BEVK_TMA_MAXMULT=2
