#!/bin/bash
# GPU call C: parity suite, headline bench after the consumer-kernel work, router overhead, memcheck of smoke().
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c_tests.log
tail -4 gpurun_out/c_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c_bench_256_bf16.json 2> gpurun_out/c_bench_256_bf16.err
echo "bench rc=$?"
timeout 300 python tools/profile_steps.py --workload lm --steps 5 > gpurun_out/c_lm_alone.txt 2>&1
timeout 300 python tools/profile_steps.py --workload lm --steps 5 --kv int8 > gpurun_out/c_lm_alone_int8.txt 2>&1
grep -h alone gpurun_out/c_lm_alone*.txt
timeout 600 python tools/sched_bench.py --devices 1 --slots 256 --frames 150 > gpurun_out/c_sched_bench.json 2> gpurun_out/c_sched_bench.err
echo "sched rc=$?"; cat gpurun_out/c_sched_bench.json
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/c_sanitizer_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_sanitizer_memcheck.out 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/c_sanitizer_memcheck.log
