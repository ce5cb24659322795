"""Every A/B switch that is still in the library must give the same results as the default
path: the parity tests of the component it touches are re-run in a subprocess with the switch set (the switches are
read once per process)."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu

VARIANTS = [
    ({"DSM_TC6_SPLITK": "0"}, ["tests/test_gpu_codec_gemm.py", "tests/test_gpu_mimi.py"]),
    ({"DSM_CODEC_PRESPLIT": "0"}, ["tests/test_gpu_mimi.py"]),
    ({"DSM_SEANET_FUSED1": "0"}, ["tests/test_gpu_mimi.py"]),
    ({"DSM_RVQ_GROUP": "1"}, ["tests/test_gpu_mimi.py", "tests/test_gpu_real_audio.py"]),
    ({"DSM_RVQ_GROUP": "3"}, ["tests/test_gpu_mimi.py"]),
    ({"DSM_RVQ_FUSED_MAX": "512"}, ["tests/test_gpu_mimi.py", "tests/test_gpu_real_audio.py"]),
    ({"DSM_PDL": "1"}, ["tests/test_gpu_asr.py", "tests/test_gpu_lm.py"]),
    ({"DSM_TC_MAXSPLIT": "1"}, ["tests/test_gpu_lm.py"]),
    ({"DSM_CODEC_GEMM": "simt", "DSM_LM_GEMM": "simt"}, ["tests/test_gpu_asr.py"]),
    ({"DSM_NO_GRAPH": "1"}, ["tests/test_gpu_asr.py"]),
    ({"DSM_SM_SPLIT": "0"}, ["tests/test_gpu_full_size.py", "tests/test_gpu_asr.py"]),   # ASR states partition the SMs by default
    ({"DSM_SM_SPLIT": "96"}, ["tests/test_gpu_asr.py"]),               # forced on small batches: re-tuned GEMM grids
    ({"DSM_SM_SPLIT": "64"}, ["tests/test_gpu_full_size.py"]),
]


@pytest.mark.parametrize("env,files", VARIANTS, ids=lambda v: ",".join(f"{k}={x}" for k, x in v.items()) if isinstance(v, dict) else None)
def test_switch_keeps_parity(env, files):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "-m", "gpu", *files], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
