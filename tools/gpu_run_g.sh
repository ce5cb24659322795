#!/bin/bash
# GPU call G: LM GEMM occupancy experiment, refreshed bench lines (256 bf16 with lm_step_frac, 2,000 int8, configs 2/4/5).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for o in 0 1; do
  DSM_TC_OCC2=$o timeout 300 python tools/profile_steps.py --workload lm --steps 5 > gpurun_out/g_lm_occ$o.txt 2>&1
  grep -h alone gpurun_out/g_lm_occ$o.txt
done
DSM_TC_OCC2=1 timeout 300 python -m pytest tests/test_gpu_lm.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/profile_steps.py --workload mimi --steps 5 > gpurun_out/g_mimi.txt 2>&1; grep -h alone gpurun_out/g_mimi.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench_256_bf16.json 2> gpurun_out/g_bench_256_bf16.err
DSM_TC_OCC2=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-steps 100 > gpurun_out/g_bench_256_bf16_occ2.json 2> gpurun_out/g_bench_256_bf16_occ2.err
timeout 900 python bench.py --steps 20 --warmup 5 --kv int8 --streams 2000 --latency-steps 200 --no-cpu-baseline > gpurun_out/g_bench_2000_int8.json 2> gpurun_out/g_bench_2000_int8.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload asr-2.6b --no-cpu-baseline --latency-steps 100 > gpurun_out/g_bench_26b.json 2> gpurun_out/g_bench_26b.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload decode --no-cpu-baseline --latency-steps 100 > gpurun_out/g_bench_decode.json 2> gpurun_out/g_bench_decode.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload mimi --no-cpu-baseline --latency-steps 100 > gpurun_out/g_bench_mimi.json 2> gpurun_out/g_bench_mimi.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/g_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["ms_per_step"], 3), round(d["value"]), d["clocks"], {k: round(r[k], 3) for k in ("frac", "step_frac", "lm_step_frac", "lm_step_ms") if k in r})
    except Exception as e:
        print(f, "ERR", e)
PY
