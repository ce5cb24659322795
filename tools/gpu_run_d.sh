#!/bin/bash
# GPU call D: fused RVQ kernel -- codec parity tests, codec-alone time with and without it, headline bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mimi.py tests/test_gpu_real_audio.py tests/test_gpu_asr.py tests/test_gpu_full_size.py tests/test_gpu_sched.py -q -s -p no:cacheprovider > gpurun_out/d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/d_tests.log
tail -6 gpurun_out/d_tests.log
for mx in 512 0; do
  DSM_RVQ_FUSED_MAX=$mx timeout 300 python tools/profile_steps.py --workload mimi --steps 5 > gpurun_out/d_mimi_fused$mx.txt 2>&1
  grep -h alone gpurun_out/d_mimi_fused$mx.txt
  DSM_RVQ_FUSED_MAX=$mx timeout 300 python tools/profile_steps.py --workload mimi --steps 5 --streams 64 > gpurun_out/d_mimi64_fused$mx.txt 2>&1
  grep -h alone gpurun_out/d_mimi64_fused$mx.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-steps 100 > gpurun_out/d_bench_256_bf16.json 2> gpurun_out/d_bench_256_bf16.err
echo "bench rc=$?"
DSM_RVQ_FUSED_MAX=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-steps 100 > gpurun_out/d_bench_256_bf16_chain.json 2> gpurun_out/d_bench_256_bf16_chain.err
python - <<'PY'
import json
for f in ("d_bench_256_bf16", "d_bench_256_bf16_chain"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
