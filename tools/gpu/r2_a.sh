#!/bin/bash
# round 2, call A: first GPU contact of the TMA-staged kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a/gpu.txt 2>&1
echo "== tma tests" 
timeout 900 python -m pytest tests/test_gpu_tma.py -x -q 2>&1 | tail -25 | tee gpurun_out/r2a/pytest_tma.log
echo "== bench tma"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2a/bench_tma.json 2> gpurun_out/r2a/bench_tma.err; tail -c 3000 gpurun_out/r2a/bench_tma.json
echo "== bench gather (round-1 kernel, pointer table)"
BEVK_BENCH_TABLE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2a/bench_gather.json 2> gpurun_out/r2a/bench_gather.err; tail -c 1500 gpurun_out/r2a/bench_gather.json
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2a/pytest_gpu.log
echo "== ncu launch list + full capture"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/r2a/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2a/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bev_tma -s 5 -c 1 -o gpurun_out/r2a/prof_tma python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2a/b_ncu2.log 2>&1
ls -la gpurun_out/r2a
