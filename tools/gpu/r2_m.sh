#!/bin/bash
# round 2, multi-GPU call: bash tools/gpu/r2_m.sh <N>   (gpurun --gpus N)
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out/r2m$N; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
if [ "$N" = "2" ]; then
  echo "== shard tests (world-2 NCCL)"
  timeout 900 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -15 | tee $O/pytest_shard.log
fi
run() {  # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err
  python - "$O/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "iso", round(d["roofline"]["kernel_ms_isolated"],4), "link_bytes", d.get("link_bytes_per_step"), d["config"]["workload"][:40])
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
echo "== frames policy, N=$N (default workload)"
run frames_cfg4 --steps 20 --warmup 3
echo "== cameras policy, N=$N, cfg4 batch 32"
BEVK_BENCH_NO_API=1 run cameras_cfg4 --steps 20 --warmup 3 --shard cameras
if [ "$N" = "8" ] || [ "$N" = "4" ]; then
  echo "== cameras policy, N=$N, cfg5 (8 cams 4K -> 2000^2)"
  BEVK_BENCH_NO_API=1 run cameras_cfg5 --steps 10 --warmup 3 --shard cameras --workload cfg5
fi
ls $O
