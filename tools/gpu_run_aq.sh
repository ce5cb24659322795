#!/bin/bash
# GPU call AQ: the 8-bit ring at small batches with and without the partition (the rule partitions there as well).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 1 16; do for s in 96 0; do
  DSM_SM_SPLIT=$s timeout 300 python bench.py --steps 50 --warmup 5 --streams $b --kv int8 --latency-steps 200 --no-cpu-baseline > gpurun_out/aq_b${b}_$s.json 2> gpurun_out/aq_b${b}_$s.err
done; done
python - <<'PY'
import json
for b in (1, 16):
    for s in (96, 0):
        d = json.loads(open(f"gpurun_out/aq_b{b}_{s}.json").read().strip().splitlines()[-1])
        l = d["latency"]
        print(f"int8 streams={b} DSM_SM_SPLIT={s}: {d['ms_per_step']:.3f} ms/step, e2e {d['e2e']['ms_per_step']:.3f}, p50 {l['p50_ms']:.3f}, p99 {l['p99_ms']:.3f}")
PY
