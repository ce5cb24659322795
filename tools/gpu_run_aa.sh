#!/bin/bash
# GPU call AA: SM partitioning, second sweep: soft split (LM confined, codec everywhere) vs hard split.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # tag, split, extra bench args
  echo "== split=$2 $1" >> gpurun_out/aa_bench.log
  DSM_SM_SPLIT=$2 timeout 900 python bench.py ${@:3} --no-cpu-baseline > gpurun_out/aa_$1_$2.json 2> gpurun_out/aa_$1_$2.err
  python - "$1" "$2" >> gpurun_out/aa_bench.log <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/aa_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"], d["roofline"]["lm_step_ms"], round(d["roofline"]["frac"], 3), d["latency"]["p99_ms"])
except Exception as e:
    print("ERR", e, open(f"gpurun_out/aa_{sys.argv[1]}_{sys.argv[2]}.err").read()[-800:])
PY
}
for s in 104s 112s 96s 120s 96 112 0; do run b256 $s --steps 20 --warmup 5 --latency-steps 50; done
for s in 120s 104s 0; do run b2000 $s --steps 10 --warmup 3 --kv int8 --streams 2000 --latency-steps 30; done
for s in 104s 0; do run b256i8 $s --steps 20 --warmup 5 --kv int8 --latency-steps 50; done
cat gpurun_out/aa_bench.log
