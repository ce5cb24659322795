#!/bin/bash
# GPU call H: variant parity suite, zero-cost codec knobs, clean-state bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_variants.py tests/test_gpu_codec_gemm.py -q -p no:cacheprovider > gpurun_out/h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/h_tests.log; tail -5 gpurun_out/h_tests.log
for v in "X=0" "DSM_TC6_SPLITK=1" "DSM_CODEC_PRESPLIT=0"; do
  env $v timeout 300 python tools/profile_steps.py --workload mimi --steps 5 > gpurun_out/h_mimi_$v.txt 2>&1; echo "$v $(grep -h alone gpurun_out/h_mimi_$v.txt)"
  env $v timeout 300 python tools/profile_steps.py --workload mimi --steps 5 --streams 2000 > gpurun_out/h_mimi2000_$v.txt 2>&1; echo "$v 2000: $(grep -h alone gpurun_out/h_mimi2000_$v.txt)"
done
timeout 300 python tools/profile_steps.py --workload lm --steps 5 --streams 2000 --kv int8 > gpurun_out/h_lm2000.txt 2>&1; grep -h alone gpurun_out/h_lm2000.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench_256_bf16.json 2> gpurun_out/h_bench_256_bf16.err
DSM_TC6_SPLITK=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-steps 100 > gpurun_out/h_bench_256_bf16_splitk.json 2> gpurun_out/h_bench_256_bf16_splitk.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/h_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["ms_per_step"], 3), round(d["value"]), d["clocks"], {k: round(r[k], 3) for k in ("frac", "step_frac", "lm_step_frac", "lm_step_ms") if k in r})
    except Exception as e:
        print(f, "ERR", e)
PY
