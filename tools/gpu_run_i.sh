#!/bin/bash
# GPU call I: codec knob combinations, gate tile at large batch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lm.py -q -p no:cacheprovider > gpurun_out/i_tests.log 2>&1; tail -3 gpurun_out/i_tests.log
for v in "X=0" "DSM_TC6_SPLITK=1" "DSM_TC6_SPLITK=1 DSM_CODEC_PRESPLIT=0"; do
  n=$(echo $v | tr ' ' '_')
  env $v timeout 300 python tools/profile_steps.py --workload mimi --steps 5 > gpurun_out/i_mimi_$n.txt 2>&1; echo "$v $(grep -h alone gpurun_out/i_mimi_$n.txt)"
  env $v timeout 300 python tools/profile_steps.py --workload mimi --steps 5 --streams 64 > gpurun_out/i_mimi64_$n.txt 2>&1; echo "$v 64: $(grep -h alone gpurun_out/i_mimi64_$n.txt)"
  env $v timeout 300 python tools/profile_steps.py --workload decode --steps 5 --streams 64 > gpurun_out/i_dec64_$n.txt 2>&1
done
timeout 300 python tools/profile_steps.py --workload lm --steps 5 --streams 2000 --kv int8 > gpurun_out/i_lm2000.txt 2>&1; grep -h alone gpurun_out/i_lm2000.txt
timeout 300 python tools/profile_steps.py --workload lm --steps 5 --streams 1024 --kv int8 > gpurun_out/i_lm1024.txt 2>&1; grep -h alone gpurun_out/i_lm1024.txt
for v in "X=0" "DSM_TC6_SPLITK=1 DSM_CODEC_PRESPLIT=0"; do
  n=$(echo $v | tr ' ' '_')
  env $v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --latency-steps 100 > gpurun_out/i_bench_256_$n.json 2> gpurun_out/i_bench_256_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/i_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), round(d["value"]), d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
