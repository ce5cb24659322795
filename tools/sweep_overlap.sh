#!/bin/bash
# A/B of the frame graph's branch overlap: tc6 CTA cap and stream priorities (bench.py, 256 streams).
out=gpurun_out/sweep_overlap.txt; : > $out
run() { echo "== $*" >> $out; env "$@" python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> $out 2>&1; }
run A=0
run DSM_LM_PRIORITY=1
run DSM_MIMI_PRIORITY=1
run DSM_TC6_MAXCTAS=120
run DSM_TC6_MAXCTAS=96
run DSM_TC6_MAXCTAS=74
run DSM_TC6_MAXCTAS=96 DSM_LM_PRIORITY=1
run DSM_TC6_MAXCTAS=120 DSM_MIMI_PRIORITY=1
cat $out
