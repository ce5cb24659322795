#!/bin/bash
# GPU call B: parity suite on the PDL + partial-plane build, A/B of PDL and split-K, branch-alone times, launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b_tests.log
tail -4 gpurun_out/b_tests.log
for pdl in 1 0; do
  DSM_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-steps 100 > gpurun_out/b_bench_pdl$pdl.json 2> gpurun_out/b_bench_pdl$pdl.err
  echo "bench pdl=$pdl rc=$?"
  DSM_PDL=$pdl timeout 300 python tools/profile_steps.py --workload lm --steps 5 > gpurun_out/b_lm_pdl$pdl.txt 2>&1
  DSM_PDL=$pdl timeout 300 python tools/profile_steps.py --workload mimi --steps 5 > gpurun_out/b_mimi_pdl$pdl.txt 2>&1
done
DSM_TC_MAXSPLIT=1 timeout 300 python tools/profile_steps.py --workload lm --steps 5 > gpurun_out/b_lm_nosplit.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-steps 100 --kv int8 > gpurun_out/b_bench_int8.json 2> gpurun_out/b_bench_int8.err
echo "bench int8 rc=$?"
timeout 300 python tools/profile_steps.py --eager --steps 3 > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/b_launches.csv \
    python tools/profile_steps.py --eager --steps 3 > gpurun_out/b_ncu.log 2>&1
echo "ncu rc=$?"
grep -h "alone" gpurun_out/b_*_pdl*.txt gpurun_out/b_lm_nosplit.txt
