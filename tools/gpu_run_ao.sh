#!/bin/bash
# GPU call AO: does the SM partition pay below 8 streams as well?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 1 2 4; do for s in 96 0; do
  DSM_SM_SPLIT=$s timeout 600 python bench.py --steps 50 --warmup 5 --streams $b --latency-steps 300 --no-cpu-baseline > gpurun_out/ao_b${b}_$s.json 2> gpurun_out/ao_b${b}_$s.err
done; done
python - <<'PY'
import json
for b in (1, 2, 4):
    for s in (96, 0):
        d = json.loads(open(f"gpurun_out/ao_b{b}_{s}.json").read().strip().splitlines()[-1])
        l = d["latency"]
        print(b, s, round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "p50", round(l["p50_ms"], 3), "p99", round(l["p99_ms"], 3))
PY
