#!/bin/bash
# round 2, call O: finer slot timeline (release, barrier, write-out) of the shipped configuration
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2o; mkdir -p $O
BEVK_BENCH_NO_API=1 BEVK_LIB_PATH=$PWD/ab/libbevk_trace7.so BEVK_TRACE_FILE=$PWD/$O/trace.bin timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 50 --warmup 5 > $O/bench_trace.json 2> $O/bench_trace.err
python tools/gpu/trace_slots.py $O/trace.bin
