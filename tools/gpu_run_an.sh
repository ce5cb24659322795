#!/bin/bash
# GPU call AN: small-batch latency lines (1 and 8 streams) of the final build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 1 8; do
  timeout 600 python bench.py --steps 50 --warmup 5 --streams $b --latency-steps 300 --no-cpu-baseline > gpurun_out/an_bench_b$b.json 2> gpurun_out/an_bench_b$b.err
done
python - <<'PY'
import json
for b in (1, 8):
    d = json.loads(open(f"gpurun_out/an_bench_b{b}.json").read().strip().splitlines()[-1])
    l = d["latency"]
    print(b, round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "p50", round(l["p50_ms"], 3), "p99", round(l["p99_ms"], 3), d["config"].get("sm_partition"))
PY
