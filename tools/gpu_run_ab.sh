#!/bin/bash
# GPU call AB: SM partitioning, third sweep: is the gain at 96 + 52 stable (p99 over 300 steps), does it need the
# partition-sized GEMM grids, and what do the int8 ring and stt-2.6b want.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # tag, split, extra bench args
  echo "== split=$2 $1 budget=${DSM_SM_SPLIT_BUDGET:-1}" >> gpurun_out/ab_bench.log
  DSM_SM_SPLIT=$2 timeout 900 python bench.py ${@:3} --no-cpu-baseline > gpurun_out/ab_$1_$2_${DSM_SM_SPLIT_BUDGET:-1}.json 2> gpurun_out/ab_$1_$2_${DSM_SM_SPLIT_BUDGET:-1}.err
  python - "gpurun_out/ab_$1_$2_${DSM_SM_SPLIT_BUDGET:-1}" >> gpurun_out/ab_bench.log <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
    l = d["latency"]
    print(round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), d["clocks"]["sm_mhz"], "lm", round(d["roofline"]["lm_step_ms"], 3), round(d["roofline"]["frac"], 3), "p50", round(l["p50_ms"], 3), "p99", round(l["p99_ms"], 3), "max", round(l["max_ms"], 3))
except Exception as e:
    print("ERR", e, open(sys.argv[1] + ".err").read()[-800:])
PY
}
for s in 96 0 96 100 92; do run b256 $s --steps 50 --warmup 5 --latency-steps 300; done
DSM_SM_SPLIT_BUDGET=0 run b256 96 --steps 50 --warmup 5 --latency-steps 300
for s in 88 96 80; do run b256i8 $s --steps 20 --warmup 5 --kv int8 --latency-steps 100; done
for s in 0 104 112 96; do run b26 $s --steps 20 --warmup 5 --workload asr-2.6b --latency-steps 100; done
for s in 96 0; do run b128 $s --steps 20 --warmup 5 --streams 128 --latency-steps 100; done
for s in 96 0; do run b512 $s --steps 20 --warmup 5 --streams 512 --latency-steps 100; done
cat gpurun_out/ab_bench.log
