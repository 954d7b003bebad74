#!/bin/bash
# round 2, call S: evidence for the final k_bev_tma state: full GPU suite, default bench (+ reference arm), cfg3 / cfg2,
# launch lists, one ncu --set full capture, slot timeline
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2s; mkdir -p $O
echo "== gpu tests"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== default bench (with cpu baseline) + reference arm"
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('default ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'api',d['e2e_api']['pinned']['ms_per_call'],d['e2e_api']['pageable']['ms_per_call'])"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_k20.json 2> $O/bench_k20.err; python -c "
import json;d=json.loads(open('$O/bench_k20.json').read().strip().splitlines()[-1]);print('k20 ms/step',d['ms_per_step'],'e2e',d['e2e']['value'])"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 400 $O/bench_reference.json
echo "== cfg3 / cfg2 bench"
BEVK_BENCH_NO_API=1 timeout 600 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err; python -c "
import json;d=json.loads(open('$O/bench_cfg3.json').read().strip().splitlines()[-1]);print('cfg3 ms/step',d['ms_per_step'])"
BEVK_BENCH_NO_API=1 timeout 600 python bench.py --workload cfg2 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err; python -c "
import json;d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1]);print('cfg2 ms/step',d['ms_per_step'])"
echo "== ncu launch lists"
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu.log 2>&1
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/launches_cfg3.csv python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu_cfg3.log 2>&1
echo "== ncu full capture"
BEVK_BENCH_NO_API=1 BEVK_BENCH_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bev_tma -s 5 -c 1 -f -o $O/prof_tma python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/b_ncu2.log 2>&1
echo "== slot timeline"
BEVK_BENCH_NO_API=1 BEVK_LIB_PATH=$PWD/ab/libbevk_trace.so BEVK_TRACE_FILE=$PWD/$O/trace.bin timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 50 --warmup 5 > $O/bench_trace.json 2> $O/bench_trace.err
python tools/gpu/trace_slots.py $O/trace.bin | tee $O/trace_slots.txt
ls -la $O
