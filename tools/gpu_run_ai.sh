#!/bin/bash
# GPU call AI: last A/B runs on the partitioned step: PDL again (the branches no longer share SMs), other splits for
# stt-2.6b and the int8 ring, RVQ groups of 4.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # tag, extra bench args   (environment as set by the caller)
  echo "== $1 split=${DSM_SM_SPLIT:-rule} pdl=${DSM_PDL:-0} rvq_group=${DSM_RVQ_GROUP:-8}" >> gpurun_out/ai_bench.log
  timeout 900 python bench.py ${@:2} --no-cpu-baseline > gpurun_out/ai_$1.json 2> gpurun_out/ai_$1.err
  python - "gpurun_out/ai_$1" >> gpurun_out/ai_bench.log <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
    l = d["latency"]
    print(round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), d["clocks"]["sm_mhz"], "lm", round(d["roofline"]["lm_step_ms"], 3), round(d["roofline"]["frac"], 3), "p50", round(l["p50_ms"], 3), "p99", round(l["p99_ms"], 3), d["config"].get("sm_partition") and d["config"]["sm_partition"]["lm_sms"])
except Exception as e:
    print("ERR", e, open(sys.argv[1] + ".err").read()[-800:])
PY
}
A="--steps 30 --warmup 5 --latency-steps 100"
run base256 $A
DSM_PDL=1 run pdl256 $A
DSM_RVQ_GROUP=4 run rvq4_256 $A
run base256_again $A
DSM_SM_SPLIT=88 run b26_88 $A --workload asr-2.6b
DSM_SM_SPLIT=80 run b26_80 $A --workload asr-2.6b
DSM_SM_SPLIT=104 run i8_104 $A --kv int8
DSM_SM_SPLIT=112 run i8_112 $A --kv int8
cat gpurun_out/ai_bench.log
