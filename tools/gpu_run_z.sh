#!/bin/bash
# GPU call Z: SM partitioning through green contexts (DSM_SM_SPLIT=<LM SMs>): does the frame graph capture and replay,
# is parity kept, and what does the step cost per split.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DSM_SM_SPLIT=104 timeout 900 python -m pytest tests/test_gpu_asr.py -q -x -p no:cacheprovider > gpurun_out/z_tests.log 2>&1; tail -5 gpurun_out/z_tests.log
for s in 0 120 104 88; do
  echo "== split=$s streams=256 bf16" >> gpurun_out/z_bench.log
  DSM_SM_SPLIT=$s timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-steps 50 > gpurun_out/z_b256_$s.json 2> gpurun_out/z_b256_$s.err
  python - "$s" >> gpurun_out/z_bench.log <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/z_b256_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"], 3), d["clocks"], d["roofline"]["lm_step_ms"], d["roofline"]["frac"])
except Exception as e:
    print("ERR", e, open(f"gpurun_out/z_b256_{sys.argv[1]}.err").read()[-800:])
PY
done
for s in 0 104 80; do
  echo "== split=$s streams=2000 int8" >> gpurun_out/z_bench.log
  DSM_SM_SPLIT=$s timeout 900 python bench.py --steps 10 --warmup 3 --kv int8 --streams 2000 --no-cpu-baseline --latency-steps 30 > gpurun_out/z_b2000_$s.json 2> gpurun_out/z_b2000_$s.err
  python - "$s" >> gpurun_out/z_bench.log <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/z_b2000_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"], 3), d["clocks"], d["roofline"]["lm_step_ms"], d["roofline"]["frac"])
except Exception as e:
    print("ERR", e, open(f"gpurun_out/z_b2000_{sys.argv[1]}.err").read()[-800:])
PY
done
cat gpurun_out/z_bench.log
