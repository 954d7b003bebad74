#!/bin/bash
# round 2, call K: plan-level A/B of k_bev_tma: multi-pass factor (4 FS / 2 FS boxes vs GATHER items) and stage sizes
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2k; mkdir -p $O
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
    print(sys.argv[2], 'FAILED', e)
PY
}
L=$PWD/ab/libbevk_base3.so
run m4 BEVK_LIB_PATH=$L
run m2 BEVK_LIB_PATH=$L BEVK_TMA_MAXMULT=2
run m1 BEVK_LIB_PATH=$L BEVK_TMA_MAXMULT=1
run f5_m4 BEVK_LIB_PATH=$L BEVK_TMA_CFG=5120,2,4
run f5_m2 BEVK_LIB_PATH=$L BEVK_TMA_CFG=5120,2,4 BEVK_TMA_MAXMULT=2
run f5_m1 BEVK_LIB_PATH=$L BEVK_TMA_CFG=5120,2,4 BEVK_TMA_MAXMULT=1
run f6_m2 BEVK_LIB_PATH=$L BEVK_TMA_CFG=6144,2,3 BEVK_TMA_MAXMULT=2
run f6_m1 BEVK_LIB_PATH=$L BEVK_TMA_CFG=6144,2,3 BEVK_TMA_MAXMULT=1
run f3s3_m4 BEVK_LIB_PATH=$L BEVK_TMA_CFG=3072,3,4
run f3s3_m2 BEVK_LIB_PATH=$L BEVK_TMA_CFG=3072,3,4 BEVK_TMA_MAXMULT=2
run f532_m2 BEVK_LIB_PATH=$L BEVK_TMA_CFG=5120,3,2 BEVK_TMA_MAXMULT=2
run m4_again BEVK_LIB_PATH=$L
