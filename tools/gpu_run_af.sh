#!/bin/bash
# GPU call AF (2 GPUs): torchrun N=2 bench line and the two-device scheduler test with the SM partition active;
# launch list of three ASR steps of the final build (one GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sched.py -q -x -p no:cacheprovider > gpurun_out/af_tests.log 2>&1; tail -3 gpurun_out/af_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/af_bench_n2.json 2> gpurun_out/af_bench_n2.err; tail -c 600 gpurun_out/af_bench_n2.json
timeout 300 python tools/profile_steps.py --workload asr --steps 3 > gpurun_out/af_plain.log 2>&1 &&
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/af_launches.csv \
    python tools/profile_steps.py --workload asr --steps 3 > gpurun_out/af_ncu.log 2>&1
echo "ncu rc=$?"
