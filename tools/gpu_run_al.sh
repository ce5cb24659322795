#!/bin/bash
# GPU call AL (8 GPUs): torchrun N=8 bench line of the final build (every rank partitions its own device).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/al_bench_n8.json 2> gpurun_out/al_bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/al_bench_n8.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]), d["latency"]["p99_ms"], d["clocks"], d["config"].get("sm_partition", {}).get("lm_sms"))
PY
tail -3 gpurun_out/al_bench_n8.err
