#!/bin/bash
# GPU call AG: the partition-rule test; launch list of three ASR steps of the final build (under ncu the library shares
# the SMs by itself: green-context kernels cannot be instrumented).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_asr.py -q -x -p no:cacheprovider > gpurun_out/ag_tests.log 2>&1; tail -3 gpurun_out/ag_tests.log
timeout 300 python tools/profile_steps.py --workload asr --steps 3 > gpurun_out/ag_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/ag_launches.csv \
    python tools/profile_steps.py --workload asr --steps 3 > gpurun_out/ag_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ag_ncu.log
