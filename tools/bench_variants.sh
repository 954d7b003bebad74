#!/bin/bash
# usage: scripts_bench_variants.sh "<nvcc flags>|<env assignments>" ...   (run on the GPU box)
for spec in "$@"; do
  fl="${spec%%|*}"; envs="${spec#*|}"; [ "$envs" == "$spec" ] && envs=""
  BEVK_NVCC_FLAGS="$fl" python -m cameracalibration_b200.build --force > /dev/null 2>&1
  env $envs python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$spec]', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e ok', d['e2e']['matches_device_path'], d['plan'])"
done
