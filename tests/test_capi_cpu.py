"""The C-ABI library loads, exports every symbol include/dsm_b200.h declares, and fails loudly (no CPU
fallback) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT, has_gpu, tiny_lm, tiny_mimi


def _lib():
    from dsm_b200 import _lib
    return _lib


def test_library_is_built_in_tree():
    lib = _lib()
    assert os.path.exists(lib.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(lib.LIB_PATH).startswith(ROOT)


def test_every_declared_symbol_is_exported_and_typed():
    lib = _lib()
    header = open(lib.HEADER_PATH).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dsm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 45
    handle = C.CDLL(lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, f"header declares symbols the library does not export: {missing}"
    untyped = sorted(declared - set(lib.SIGNATURES))
    assert not untyped, f"symbols without a ctypes signature: {untyped}"


def test_spec_helpers_match_reference_configs():
    lib = _lib()
    l = lib.load()
    m = lib.MimiSpecC()
    l.dsm_mimi_spec_v0_1(C.byref(m), 32)
    assert (m.n_q, m.bins, m.qdim, m.dim, m.context, m.num_layers, m.num_heads) == (32, 2048, 256, 512, 250, 8, 8)
    assert list(m.ratios) == [8, 6, 5, 4]
    s = lib.LmSpecC()
    l.dsm_lm_spec_stt_1b_en_fr(C.byref(s))      # configs/stt/config-stt-en_fr-hf.toml:27-56
    assert (s.d_model, s.num_heads, s.num_layers, s.context, s.text_in_vocab_size, s.text_out_vocab_size) == \
        (2048, 16, 16, 750, 8001, 8000)
    assert (s.audio_codebooks, s.extra_heads_num, s.extra_heads_dim) == (32, 4, 6)
    l.dsm_lm_spec_stt_2_6b_en(C.byref(s))       # configs/stt/config-stt-en-hf.toml:24-55
    assert (s.num_heads, s.num_layers, s.context, s.text_in_vocab_size, s.extra_heads_num) == (32, 48, 375, 4001, 0)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from dsm_b200 import api
    from dsm_b200._lib import DsmError
    assert _lib().load().dsm_device_count() == 0
    with pytest.raises(DsmError) as e:
        api.Mimi(tiny_mimi(), None, 2)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(DsmError):
        api.LmModel(tiny_lm(), None, 2)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "delayed-streams-modeling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
