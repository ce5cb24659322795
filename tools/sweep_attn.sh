#!/bin/bash
# A/B of the persistent attention grid (DSM_ATTN_CTAS): LM branch alone and the full frame graph.
out=gpurun_out/sweep_attn.txt; : > $out
for c in 0 296 444 592 888 1184; do
  echo "== DSM_ATTN_CTAS=$c" >> $out
  DSM_ATTN_CTAS=$c python tools/profile_steps.py --workload lm --steps 8 2>&1 | tail -2 >> $out
  DSM_ATTN_CTAS=$c python tools/profile_steps.py --workload asr --steps 8 2>&1 | tail -2 >> $out
done
cat $out
