#!/bin/bash
# round 2, call I: where does k_bev_tma's time go?  Ablation builds of the same kernel (ab/libbevk_<x>.so, built from a
# scratch copy with -DBEVK_ABL_*: not parity-valid, timing only) + one ncu --set full capture of the unmodified build.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2i; mkdir -p $O
B="timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 200 --warmup 5"
run() { # name, env...
  local name=$1; shift
  env BEVK_BENCH_NO_API=1 "$@" $B > $O/bench_$name.json 2> $O/bench_$name.err
  python - "$O/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'],5), 'isolated', round(d['roofline']['kernel_ms_isolated'],5), 'same', d['e2e']['matches_device_path'], 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
for v in base nolds noalu notma nowrite noldsalu noall; do run $v BEVK_LIB_PATH=$PWD/ab/libbevk_$v.so; done
run base_433 BEVK_LIB_PATH=$PWD/ab/libbevk_base.so BEVK_TMA_CFG=4096,3,3
run noldsalu_433 BEVK_LIB_PATH=$PWD/ab/libbevk_noldsalu.so BEVK_TMA_CFG=4096,3,3
run v3 BEVK_LIB_PATH=$PWD/ab/libbevk_v3.so
echo "== ncu full, base"
BEVK_LIB_PATH=$PWD/ab/libbevk_base.so BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bev_tma -s 5 -c 1 -o $O/k_bev_tma_v4 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ls -la $O | head -30
