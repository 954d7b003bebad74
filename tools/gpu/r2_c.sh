#!/bin/bash
# round 2, call C: producer back-off + 2-frame-set halves; config sweep; new graph / slot / shard tests on one GPU
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c; mkdir -p $O
echo "== config sweep (device-resident value only)"
for cfg in 6144,3 6144,2 4096,3 4096,4 8192,2; do
  BEVK_TMA_CFG=$cfg BEVK_BENCH_NO_API=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $O/bench_$cfg.json 2> $O/bench_$cfg.err
  python - "$O/bench_$cfg.json" "$cfg" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "isolated", round(d["roofline"]["kernel_ms_isolated"],4), d["e2e"]["matches_device_path"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
echo "== cfg3 (balance) bench"
BEVK_BENCH_NO_API=1 timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; python -c "
import json;d=json.loads(open('$O/bench_cfg3.json').read().strip().splitlines()[-1]);print('cfg3 ms/step',d['ms_per_step'],d['plan']['path'])"
echo "== new tests"
timeout 1200 python -m pytest tests/test_gpu_tma.py tests/test_gpu_shard.py tests/test_gpu_parity.py -x -q 2>&1 | tail -15 | tee $O/pytest.log
ls $O
