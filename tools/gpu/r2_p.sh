#!/bin/bash
# round 2: peer-store camera sharding (bevk_bev_run_scattered): bash tools/gpu/r2_p.sh <N>
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out/r2p$N; mkdir -p $O
if [ "$N" = "2" ]; then
  echo "== shard tests (world-2 NCCL, incl. peer stores)"
  timeout 600 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -25 | tee $O/pytest_shard.log
fi
run() {  # name, args...
  name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err
  python - "$O/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "iso", round(d["roofline"]["kernel_ms_isolated"],4), "link_bytes", d.get("link_bytes_per_step"), "matches", d["e2e"]["matches_device_path"], d["config"]["workload"][:40])
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
echo "== cameras-p2p, N=$N, cfg4 batch 32"
BEVK_BENCH_NO_API=1 run p2p_cfg4 --steps 20 --warmup 3 --shard cameras-p2p
echo "== cameras (all-gather), N=$N, cfg4 batch 32"
BEVK_BENCH_NO_API=1 run cameras_cfg4 --steps 20 --warmup 3 --shard cameras
if [ "$N" != "2" ]; then
  echo "== cameras-p2p, N=$N, cfg5"
  BEVK_BENCH_NO_API=1 run p2p_cfg5 --steps 10 --warmup 3 --shard cameras-p2p --workload cfg5
fi
ls $O
