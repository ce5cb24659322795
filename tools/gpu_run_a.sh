#!/bin/bash
# GPU call A of round 2: parity suite, then bench lines for the three KV formats at 256 and 2,000 streams.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
free -g > gpurun_out/a_host.txt; nproc >> gpurun_out/a_host.txt
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
for kv in bf16 int8 int8k; do
  timeout 600 python bench.py --steps 20 --warmup 5 --kv $kv > gpurun_out/a_bench_256_$kv.json 2> gpurun_out/a_bench_256_$kv.err
  echo "bench 256 $kv rc=$?"
done
for kv in int8 int8k; do
  timeout 900 python bench.py --steps 20 --warmup 5 --kv $kv --streams 2000 --latency-steps 300 --no-cpu-baseline \
     > gpurun_out/a_bench_2000_$kv.json 2> gpurun_out/a_bench_2000_$kv.err
  echo "bench 2000 $kv rc=$?"
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
echo "ref rc=$?"
