#!/usr/bin/env python
"""Summarises an `ncu --set full` capture of the LM attention kernel into profiles/attn_traffic.json, the file
bench.py reads `roofline.traffic` from.  Each record carries the sha of csrc/lm_kernels.cu at the time of the
summary: bench.py reports the traffic only while the kernel source is byte-identical (a stale number is worse than
none).  Run it right after bringing the .ncu-rep back, before touching the kernel:

    python tools/ncu_traffic.py gpurun_out/prof_attn.ncu-rep --streams 256 --kv bf16 [--model 1b]
"""
import argparse
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsm_b200.spec import stt_1b_en_fr, stt_2_6b_en  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--streams", type=int, required=True)
ap.add_argument("--kv", default="bf16")
ap.add_argument("--model", default="1b")
ap.add_argument("--summary-csv", default="", help="also write the selected raw-page columns of every captured launch here")
a = ap.parse_args()
spec = stt_1b_en_fr() if a.model == "1b" else stt_2_6b_en()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
lines = raw.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
kname = "lm_attention_kernel" if a.kv == "bf16" else "lm_attention_q8_kernel"
tot, n = 0.0, 0
for r in body:
    if kname + "<" not in r[ix["Kernel Name"]]:
        continue
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ix[k]].replace(",", "")) * scale[units[ix[k]]]
    n += 1
if not n:
    raise SystemExit(f"no {kname} launch in {a.rep}")
src = os.path.join(ROOT, "delayed-streams-modeling_b200", "csrc", "lm_kernels.cu")
rec = {"kernel": kname, "streams": a.streams, "heads": spec.num_heads, "head_dim": spec.head_dim, "context": spec.context,
       "kv": a.kv, "launches": n, "dram_bytes_per_launch": tot / n,
       "lm_kernels_sha16": hashlib.sha256(open(src, "rb").read()).hexdigest()[:16], "capture": os.path.basename(a.rep)}
out = os.path.join(ROOT, "profiles", "attn_traffic.json")
recs = json.load(open(out)) if os.path.exists(out) else []
recs = [r for r in recs if (r["streams"], r["kv"], r["head_dim"], r["context"]) != (a.streams, a.kv, spec.head_dim, spec.context)]
recs.append(rec)
json.dump(recs, open(out, "w"), indent=1)
print(json.dumps(rec))
if a.summary_csv:
    cols = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
    cols = [c for c in cols if c in ix]
    with open(a.summary_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[ix[c]] for c in cols])
        for r in body:
            w.writerow([r[ix[c]][:70] for c in cols])
