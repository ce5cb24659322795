#!/bin/bash
# GPU call AH: live-stream sweep on the final build (int8 ring; Gram tables and the fused stage 1 changed the footprint).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 10 --warmup 3 --kv int8 --streams 2000 --latency-steps 200 --live-sweep 2400,2800 --no-cpu-baseline > gpurun_out/ah_bench_2000_int8_live_sweep.json 2> gpurun_out/ah_bench_2000_int8_live_sweep.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ah_bench_2000_int8_live_sweep.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 3), d["latency"]["p99_ms"], d.get("live_streams_max"), d.get("live_sweep"))
PY
