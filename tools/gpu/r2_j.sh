#!/bin/bash
# round 2, call J: slot timeline of k_bev_tma (clock64 stamps of producer and consumer per ring slot, BEVK_TRACE build) and
# the 3-stage ring with 3 entry groups per slot
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2j; mkdir -p $O
B="timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 200 --warmup 5"
run() { # name, env...
  local name=$1; shift
  env BEVK_BENCH_NO_API=1 "$@" $B > $O/bench_$name.json 2> $O/bench_$name.err
  python - "$O/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'],5), 'isolated', round(d['roofline']['kernel_ms_isolated'],5), 'same', d['e2e']['matches_device_path'], 'items', d['plan']['tma']['items'], 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
L=$PWD/ab/libbevk_base2.so
run base2 BEVK_LIB_PATH=$L
run base2_433 BEVK_LIB_PATH=$L BEVK_TMA_CFG=4096,3,3
run base2_432 BEVK_LIB_PATH=$L BEVK_TMA_CFG=4096,3,2
run base2_533 BEVK_LIB_PATH=$L BEVK_TMA_CFG=5120,3,2
run v3 BEVK_LIB_PATH=$PWD/ab/libbevk_v3.so
T=$PWD/ab/libbevk_trace.so
run trace BEVK_LIB_PATH=$T BEVK_TRACE_FILE=$PWD/$O/trace_422.bin
run trace_433 BEVK_LIB_PATH=$T BEVK_TRACE_FILE=$PWD/$O/trace_433.bin BEVK_TMA_CFG=4096,3,3
ls -la $O | head -30
