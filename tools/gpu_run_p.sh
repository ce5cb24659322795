#!/bin/bash
# GPU call P: short-K codec layers on the CUDA-core GEMM (experiment), encode and decode, 256 and 2,000 streams.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DSM_CODEC_SIMT_MAXK=64 timeout 600 python -m pytest tests/test_gpu_mimi.py tests/test_gpu_real_audio.py -q -p no:cacheprovider > gpurun_out/p_tests.log 2>&1; tail -3 gpurun_out/p_tests.log
for k in 0 32 64 128; do
  for b in 256 2000 64; do
    DSM_CODEC_SIMT_MAXK=$k timeout 300 python tools/profile_steps.py --workload mimi --steps 5 --streams $b > gpurun_out/p_mimi_${k}_$b.txt 2>&1
    echo "simt_maxk=$k streams=$b $(grep -h alone gpurun_out/p_mimi_${k}_$b.txt)"
  done
done
for k in 0 64; do
  DSM_CODEC_SIMT_MAXK=$k timeout 600 python bench.py --steps 20 --warmup 5 --workload decode --no-cpu-baseline --latency-steps 0 > gpurun_out/p_bench_decode_$k.json 2> gpurun_out/p_bench_decode_$k.err
  DSM_CODEC_SIMT_MAXK=$k timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --latency-steps 0 > gpurun_out/p_bench_256_$k.json 2> gpurun_out/p_bench_256_$k.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/p_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), round(d["value"]), d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
