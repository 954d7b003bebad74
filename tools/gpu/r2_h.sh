#!/bin/bash
# round 2, call H: first GPU run of the slimmer k_bev_tma consumer path (one-DP2A blend, pre-digested slot descriptor,
# row-wise interior write-out): full GPU suite, then A/B against the previous library (ab/libbevk_v3.so) on the same box
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2h; mkdir -p $O
echo "== gpu tests"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $O/pytest_gpu.log
B="timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1"
run() { # name, env...
  local name=$1; shift
  env BEVK_BENCH_NO_API=1 "$@" $B --steps 200 --warmup 5 > $O/bench_$name.json 2> $O/bench_$name.err
  python - "$O/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'],5), 'isolated', round(d['roofline']['kernel_ms_isolated'],5), 'frac', round(d['roofline']['frac'],4), 'path', d['plan']['path'], 'same', d['e2e']['matches_device_path'], 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
echo "== A/B"
run new
run old BEVK_LIB_PATH=$PWD/ab/libbevk_v3.so
run new_432 BEVK_TMA_CFG=4096,3,2
run old_432 BEVK_LIB_PATH=$PWD/ab/libbevk_v3.so BEVK_TMA_CFG=4096,3,2
run new_again
echo "== short run (driver's K=20, W=5)"
BEVK_BENCH_NO_API=1 $B --steps 20 --warmup 5 > $O/bench_new_k20.json 2>$O/bench_new_k20.err
python -c "
import json;d=json.loads(open('$O/bench_new_k20.json').read().strip().splitlines()[-1]);print('k20 ms/step',d['ms_per_step'],'isolated',d['roofline']['kernel_ms_isolated'])"
ls $O
