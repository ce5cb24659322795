#!/bin/bash
# GPU call AD: SM partition, edges of the rule: other splits at 1,024 streams, 1,536 streams bf16, 32 streams.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # tag, split, extra bench args
  echo "== split=$2 $1" >> gpurun_out/ad_bench.log
  DSM_SM_SPLIT=$2 timeout 900 python bench.py ${@:3} --no-cpu-baseline > gpurun_out/ad_$1_$2.json 2> gpurun_out/ad_$1_$2.err
  python - "gpurun_out/ad_$1_$2" >> gpurun_out/ad_bench.log <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
    l = d["latency"]
    print(round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), d["clocks"]["sm_mhz"], "lm", round(d["roofline"]["lm_step_ms"], 3), round(d["roofline"]["frac"], 3), "p50", round(l["p50_ms"], 3), "p99", round(l["p99_ms"], 3), d["config"].get("sm_partition") and d["config"]["sm_partition"]["lm_sms"])
except Exception as e:
    print("ERR", e, open(sys.argv[1] + ".err").read()[-800:])
PY
}
run b1024 88 --steps 20 --warmup 5 --streams 1024 --latency-steps 50
run b1024 104 --steps 20 --warmup 5 --streams 1024 --latency-steps 50
run b1536 96 --steps 10 --warmup 3 --streams 1536 --latency-steps 50
run b1536 0 --steps 10 --warmup 3 --streams 1536 --latency-steps 50
run b32 96 --steps 20 --warmup 5 --streams 32 --latency-steps 100
run b32 0 --steps 20 --warmup 5 --streams 32 --latency-steps 100
run b8 96 --steps 20 --warmup 5 --streams 8 --latency-steps 100
run b8 0 --steps 20 --warmup 5 --streams 8 --latency-steps 100
run b512i8 96 --steps 20 --warmup 5 --streams 512 --kv int8 --latency-steps 50
run b512i8 0 --steps 20 --warmup 5 --streams 512 --kv int8 --latency-steps 50
cat gpurun_out/ad_bench.log
