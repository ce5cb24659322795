#!/bin/bash
# GPU call AC: SM partition chosen by the ASR state itself (96 + 52 SMs from 128 to 512 streams): parity of everything
# that builds such a state, the default bench line, and the two ends of the rule (64 and 1,024 streams).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_asr.py tests/test_gpu_full_size.py tests/test_gpu_sched.py tests/test_gpu_session.py tests/test_gpu_lm_kv.py -q -x -p no:cacheprovider > gpurun_out/ac_tests.log 2>&1; tail -4 gpurun_out/ac_tests.log
timeout 1800 python -m pytest tests/test_gpu_variants.py -k "SM_SPLIT" -q -x -p no:cacheprovider > gpurun_out/ac_tests2.log 2>&1; tail -4 gpurun_out/ac_tests2.log
run() {  # tag, split ("" = the library's rule), extra bench args
  echo "== split=${2:-auto} $1" >> gpurun_out/ac_bench.log
  if [ -n "$2" ]; then export DSM_SM_SPLIT=$2; else unset DSM_SM_SPLIT; fi
  timeout 900 python bench.py ${@:3} --no-cpu-baseline > gpurun_out/ac_$1_${2:-auto}.json 2> gpurun_out/ac_$1_${2:-auto}.err
  python - "gpurun_out/ac_$1_${2:-auto}" >> gpurun_out/ac_bench.log <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
    l = d["latency"]
    print(round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), d["clocks"]["sm_mhz"], "lm", round(d["roofline"]["lm_step_ms"], 3), round(d["roofline"]["frac"], 3), "p50", round(l["p50_ms"], 3), "p99", round(l["p99_ms"], 3), d["config"].get("sm_partition") and d["config"]["sm_partition"]["lm_sms"])
except Exception as e:
    print("ERR", e, open(sys.argv[1] + ".err").read()[-800:])
PY
}
run b256 "" --steps 50 --warmup 5 --latency-steps 300
run b64 96 --steps 20 --warmup 5 --streams 64 --latency-steps 100
run b64 0 --steps 20 --warmup 5 --streams 64 --latency-steps 100
run b1024 96 --steps 20 --warmup 5 --streams 1024 --latency-steps 100
run b1024 0 --steps 20 --warmup 5 --streams 1024 --latency-steps 100
run b1024i8 96 --steps 20 --warmup 5 --streams 1024 --kv int8 --latency-steps 100
run b1024i8 0 --steps 20 --warmup 5 --streams 1024 --kv int8 --latency-steps 100
run b768 96 --steps 20 --warmup 5 --streams 768 --latency-steps 100
run b768 0 --steps 20 --warmup 5 --streams 768 --latency-steps 100
cat gpurun_out/ac_bench.log
