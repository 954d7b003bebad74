#!/bin/bash
# round 2, call E: evidence -- sanitizers, launch lists, full captures, default bench + reference arm + cfg3
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2e; mkdir -p $O
echo "== sanitizers"
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/gpu/sanitize_small.py > $O/sanitize_small_$tool.log 2>&1; tail -4 $O/sanitize_small_$tool.log
  timeout 900 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tool.log 2>&1; tail -3 $O/smoke_$tool.log
done
echo "== default bench (with cpu baseline) + reference arm"
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 2500 $O/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 800 $O/bench_reference.json
echo "== cfg3 / cfg2 bench"
BEVK_BENCH_NO_API=1 timeout 600 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err; tail -c 600 $O/bench_cfg3.json
BEVK_BENCH_NO_API=1 timeout 600 python bench.py --workload cfg2 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err; tail -c 600 $O/bench_cfg2.json
echo "== Tools/undistort.py directory throughput"
timeout 600 python tools/bench_undistort_dir.py > $O/undistort_dir.json 2> $O/undistort_dir.err; cat $O/undistort_dir.json
echo "== every BASELINE config next to cv2"
timeout 900 python tools/bench_configs.py > $O/configs.json 2> $O/configs.err; tail -c 600 $O/configs.json
echo "== camera-sharded step, halves timed on one GPU"
timeout 300 python tools/gpu/shard_breakdown.py > $O/shard_breakdown.json 2> $O/shard_breakdown.err; cat $O/shard_breakdown.json
echo "== ncu launch lists"
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu.log 2>&1
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/launches_cfg3.csv python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu_cfg3.log 2>&1
echo "== ncu full captures"
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bev_tma -s 5 -c 1 -o $O/prof_tma python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu2.log 2>&1
ls $O
