#!/bin/bash
# round 2, call D: k_bev_tma v3 (descriptors and LUT entries delivered through the ring): parity, sweep, profile
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2d; mkdir -p $O
echo "== tma + shard tests"
timeout 900 python -m pytest tests/test_gpu_tma.py tests/test_gpu_shard.py tests/test_gpu_jpeg.py -x -q 2>&1 | tail -15 | tee $O/pytest_tma.log
echo "== config sweep (device-resident value only)"
for cfg in 4096,3,2 6144,2,2 4096,2,2:3 4096,2,4 5120,3,2; do
 for bo in 0; do
  c=${cfg%%:*}
  BEVK_TMA_CFG=$c BEVK_TMA_BACKOFF=$bo BEVK_BENCH_NO_API=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $O/bench_${c}_$bo.json 2> $O/bench_${c}_$bo.err
  python - "$O/bench_${c}_$bo.json" "$c backoff=$bo" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "isolated", round(d["roofline"]["kernel_ms_isolated"],4), d["plan"]["tma"]["items"], d["e2e"]["matches_device_path"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
 done
done
echo "== ncu full capture of the default config"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bev_tma -s 5 -c 1 -o $O/prof_tma python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu2.log 2>&1
ls $O | head -40
