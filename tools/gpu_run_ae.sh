#!/bin/bash
# GPU call AE: full parity suite on the final build, smoke, final bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# per-branch device times (eager, CUDA events on each branch's stream) with and without the SM partition
for s in 96 0; do
  echo "== DSM_SM_SPLIT=$s: step i ms (whole, codec branch, LM branch)" >> gpurun_out/ae_branches.log
  DSM_SM_SPLIT=$s timeout 300 python tools/profile_steps.py --workload asr --eager --steps 6 2>&1 | tail -4 >> gpurun_out/ae_branches.log
done
cat gpurun_out/ae_branches.log
timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/ae_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ae_tests.log; tail -5 gpurun_out/ae_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/ae_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/ae_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ae_bench_ref.json 2> gpurun_out/ae_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ae_bench_256_bf16.json 2> gpurun_out/ae_bench_256_bf16.err
timeout 600 python bench.py --steps 20 --warmup 5 --kv int8 --no-cpu-baseline > gpurun_out/ae_bench_256_int8.json 2> gpurun_out/ae_bench_256_int8.err
timeout 900 python bench.py --steps 20 --warmup 5 --kv int8 --streams 2000 --latency-steps 500 --no-cpu-baseline > gpurun_out/ae_bench_2000_int8.json 2> gpurun_out/ae_bench_2000_int8.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload mimi --no-cpu-baseline --latency-steps 100 > gpurun_out/ae_bench_mimi.json 2> gpurun_out/ae_bench_mimi.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload decode --no-cpu-baseline --latency-steps 100 > gpurun_out/ae_bench_decode.json 2> gpurun_out/ae_bench_decode.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload asr-2.6b --no-cpu-baseline --latency-steps 100 > gpurun_out/ae_bench_26b.json 2> gpurun_out/ae_bench_26b.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ae_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["ms_per_step"], 3), round(d["value"]), d.get("clocks"), {k: (round(r[k], 3) if r[k] is not None else None) for k in ("frac", "step_frac", "lm_step_frac", "lm_step_ms", "traffic") if k in r}, (d.get("latency") or {}).get("p99_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
