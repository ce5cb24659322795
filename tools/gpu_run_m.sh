#!/bin/bash
# GPU call M: wide LM GEMM tiles at large batch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_lm_kv.py -q -p no:cacheprovider > gpurun_out/m_tests.log 2>&1; tail -3 gpurun_out/m_tests.log
for b in 2000 1024 512; do
  timeout 300 python tools/profile_steps.py --workload lm --steps 5 --streams $b --kv int8 > gpurun_out/m_lm$b.txt 2>&1; echo "$b $(grep -h alone gpurun_out/m_lm$b.txt)"
done
timeout 900 python bench.py --steps 20 --warmup 5 --kv int8 --streams 2000 --latency-steps 200 --no-cpu-baseline > gpurun_out/m_bench_2000_int8.json 2> gpurun_out/m_bench_2000_int8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/m_bench_2000_int8.json").read().strip().splitlines()[-1])
r = d["roofline"]
print(round(d["ms_per_step"], 3), round(d["value"]), d["clocks"], {k: round(r[k], 3) for k in ("frac", "step_frac", "lm_step_frac", "lm_step_ms")}, d["latency"]["p99_ms"])
PY
