#!/bin/bash
# GPU call S: attention with the first V batch requested before the softmax passes (alternate build).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VP=$PWD/delayed-streams-modeling_b200/libdsm_b200_vp.so
for lib in default vp; do
  if [ $lib = vp ]; then export DSM_B200_LIB=$VP; fi
  timeout 300 python tools/profile_steps.py --workload lm --steps 5 > gpurun_out/s_lm_$lib.txt 2>&1; echo "$lib $(grep -h alone gpurun_out/s_lm_$lib.txt)"
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --latency-steps 0 > gpurun_out/s_bench_$lib.json 2> gpurun_out/s_bench_$lib.err
  python -c "
import json
d=json.loads(open('gpurun_out/s_bench_$lib.json').read().strip().splitlines()[-1]); print('$lib step', round(d['ms_per_step'],3), round(d['roofline']['avg_launch_ms'],4), d['clocks'])"
done
DSM_B200_LIB=$VP timeout 300 python -m pytest tests/test_gpu_lm.py -q -p no:cacheprovider -k "not tcgen05" 2>&1 | tail -1
