#!/bin/bash
# GPU call R: how much the split-K partial planes cost in the two-branch step (cap on S).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for s in 16 6 4 3 2; do
  DSM_TC_MAXSPLIT=$s timeout 300 python tools/profile_steps.py --workload lm --steps 5 > gpurun_out/r_lm_s$s.txt 2>&1
  echo "maxsplit=$s $(grep -h alone gpurun_out/r_lm_s$s.txt)"
  DSM_TC_MAXSPLIT=$s timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --latency-steps 0 > gpurun_out/r_bench_s$s.json 2> gpurun_out/r_bench_s$s.err
  python -c "
import json
d=json.loads(open('gpurun_out/r_bench_s$s.json').read().strip().splitlines()[-1]); print('maxsplit=$s step', round(d['ms_per_step'],3), d['clocks'])"
done
