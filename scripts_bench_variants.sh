#!/bin/bash
# usage: scripts_bench_variants.sh "<flags1>" "<flags2>" ...   (run on the GPU box)
for fl in "$@"; do
  BEVK_NVCC_FLAGS="$fl" python -m cameracalibration_b200.build --force > /dev/null 2>&1
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('flags [$fl]', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"
done
