#!/bin/bash
# GPU call O (2 GPUs): two-device router test, router throughput over 2 x 256 slots, bench under torchrun at N=2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/o_gpus.txt
timeout 600 python -m pytest tests/test_gpu_sched.py tests/test_gpu_lm.py -q -p no:cacheprovider > gpurun_out/o_tests.log 2>&1; tail -3 gpurun_out/o_tests.log
timeout 600 python tools/sched_bench.py --devices 2 --slots 256 --frames 150 > gpurun_out/o_sched_bench_2x256.json 2> gpurun_out/o_sched_bench.err; cat gpurun_out/o_sched_bench_2x256.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
   > gpurun_out/o_bench_n2.json 2> gpurun_out/o_bench_n2.err; tail -c 600 gpurun_out/o_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 \
   > gpurun_out/o_bench_ref_n2.json 2> gpurun_out/o_bench_ref_n2.err; tail -c 300 gpurun_out/o_bench_ref_n2.json
