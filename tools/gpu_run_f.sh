#!/bin/bash
# GPU call F: tc6 alternate-tile epilogue groups; large-batch split fix.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codec_gemm.py tests/test_gpu_mimi.py tests/test_gpu_real_audio.py tests/test_gpu_asr.py -q -s -p no:cacheprovider > gpurun_out/f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/f_tests.log
tail -6 gpurun_out/f_tests.log
DSM_RVQ_FUSED_MAX=0 timeout 300 python tools/profile_steps.py --workload mimi --steps 5 > gpurun_out/f_mimi.txt 2>&1
grep -h alone gpurun_out/f_mimi.txt
DSM_RVQ_FUSED_MAX=0 timeout 300 python tools/profile_steps.py --workload decode --steps 5 > gpurun_out/f_decode.txt 2>&1
DSM_RVQ_FUSED_MAX=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-steps 100 > gpurun_out/f_bench_256_bf16.json 2> gpurun_out/f_bench_256_bf16.err
echo "bench rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --kv int8 --streams 2000 --latency-steps 100 --no-cpu-baseline > gpurun_out/f_bench_2000_int8.json 2> gpurun_out/f_bench_2000_int8.err
echo "bench2000 rc=$?"
DSM_RVQ_FUSED_MAX=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-steps 100 --workload mimi > gpurun_out/f_bench_mimi64.json 2> gpurun_out/f_bench_mimi64.err
python - <<'PY'
import json
for f in ("f_bench_256_bf16", "f_bench_2000_int8", "f_bench_mimi64"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
