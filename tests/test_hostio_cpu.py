"""Weight / configuration I/O (csrc/hostio.cu) on the CPU: safetensors round trip through the C ABI, the reference's
own server TOMLs (configs/stt/*.toml, restated inline: /root/reference is not readable on the GPU box) -> specs."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

from helpers import LmSpec, lm_weights, tiny_lm

from dsm_b200 import _lib
from dsm_b200._lib import check
from dsm_b200.spec import stt_1b_en_fr, stt_2_6b_en
from dsm_b200.weights import to_bf16_bits


def _table_to_dict(tab):
    out = {}
    for i in range(tab.count):
        t = tab.tensors[i]
        shape = tuple(t.shape[j] for j in range(t.ndim))
        n = int(np.prod(shape)) if shape else 1
        dt = {_lib.DSM_F32: np.float32, _lib.DSM_BF16: np.uint16, _lib.DSM_F16: np.float16, _lib.DSM_U32: np.uint32}[t.dtype]
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(t.data)
        out[t.name.decode()] = (t.dtype, np.frombuffer(buf, dtype=dt).reshape(shape).copy())
    return out


def test_safetensors_round_trip(tmp_path):
    lib = _lib.load()
    w = lm_weights(tiny_lm(), seed=0)
    w = {k: v for k, v in list(w.items())[:12]}
    w["bf16.matrix"] = to_bf16_bits(np.arange(24, dtype=np.float32).reshape(4, 6))      # uint16 = bf16 bit patterns
    w["half.vector"] = np.linspace(-1, 1, 7).astype(np.float16)
    path = str(tmp_path / "w.safetensors").encode()
    table = _lib.WeightTable(w)
    check(lib.dsm_weights_save(path, C.byref(table.c)))
    # the file is a valid safetensors file: 8-byte LE header length, JSON header, then the tensors
    raw = open(path, "rb").read()
    hlen = struct.unpack("<Q", raw[:8])[0]
    hdr = json.loads(raw[8:8 + hlen])
    assert set(hdr) == set(w) and hdr["bf16.matrix"]["dtype"] == "BF16" and hdr["bf16.matrix"]["shape"] == [4, 6]
    f = C.c_void_p()
    check(lib.dsm_weights_open(path, C.byref(f)))
    got = _table_to_dict(lib.dsm_weight_file_table(f).contents)
    assert set(got) == set(w)
    for k, v in w.items():
        assert got[k][1].shape == v.shape and np.array_equal(got[k][1], v), k
    assert got["bf16.matrix"][0] == _lib.DSM_BF16 and got["half.vector"][0] == _lib.DSM_F16
    lib.dsm_weight_file_close(f)


def test_safetensors_written_by_numpy_reader_agrees(tmp_path):
    """A file laid out by hand the way the `safetensors` crate does (unordered entries, __metadata__, 5-D tensor)."""
    lib = _lib.load()
    a = np.arange(2 * 3 * 4 * 5 * 6, dtype=np.float32).reshape(2, 3, 4, 5, 6)
    b = np.array([[1.5, -2.0]], np.float32)
    hdr = {"__metadata__": {"format": "pt", "note": "café"},
           "b": {"dtype": "F32", "shape": [1, 2], "data_offsets": [a.nbytes, a.nbytes + b.nbytes]},
           "a": {"dtype": "F32", "shape": list(a.shape), "data_offsets": [0, a.nbytes]}}
    h = json.dumps(hdr).encode()
    path = tmp_path / "x.safetensors"
    path.write_bytes(struct.pack("<Q", len(h)) + h + a.tobytes() + b.tobytes())
    f = C.c_void_p()
    check(lib.dsm_weights_open(str(path).encode(), C.byref(f)))
    got = _table_to_dict(lib.dsm_weight_file_table(f).contents)
    assert np.array_equal(got["a"][1].reshape(a.shape), a) and np.array_equal(got["b"][1], b)
    buf = C.create_string_buffer(64)
    assert lib.dsm_weight_file_metadata(f, b"note", buf, 64) == len("café".encode())
    assert buf.value.decode() == "café"
    assert lib.dsm_weight_file_metadata(f, b"missing", buf, 64) == -1
    lib.dsm_weight_file_close(f)


@pytest.mark.parametrize("bad", [b"", b"\x10\x00\x00\x00\x00\x00\x00\x00{}", struct.pack("<Q", 2) + b"[]",
                                 struct.pack("<Q", 60) + b'{"a":{"dtype":"F32","shape":[4],"data_offsets":[0,16]}}     '])
def test_safetensors_rejects_malformed(tmp_path, bad):
    lib = _lib.load()
    p = tmp_path / "bad.safetensors"
    p.write_bytes(bad)
    f = C.c_void_p()
    assert lib.dsm_weights_open(str(p).encode(), C.byref(f)) != 0
    assert _lib.last_error()


# /root/reference/configs/stt/config-stt-en_fr-hf.toml:7-56 (comments shortened)
TOML_EN_FR = '''
static_dir = "./static/"
log_dir = "logs/moshi-server-rust/stt"
instance_name = "config-stt-en_fr-hf"
authorized_ids = []

[modules.asr]
path = "/api/asr-streaming"
type = "BatchedAsr"
# dtype_override = "f32"
lm_model_file = "hf://kyutai/stt-1b-en_fr-candle/model.safetensors"
text_tokenizer_file = "hf://kyutai/stt-1b-en_fr-candle/tokenizer_en_fr_audio_8000.model"
audio_tokenizer_file = "hf://kyutai/stt-1b-en_fr-candle/mimi-pytorch-e351c8d8@125.safetensors"
asr_delay_in_tokens = 6
batch_size = 64
conditioning_learnt_padding = true
temperature = 0.0

[modules.asr.model]
audio_vocab_size = 2049
text_in_vocab_size = 8001
text_out_vocab_size = 8000
audio_codebooks = 32

[modules.asr.model.transformer]
d_model = 2048
num_heads = 16
num_layers = 16
dim_feedforward = 8192
causal = true
norm_first = true
bias_ff = false
bias_attn = false
context = 750
max_period = 100000
use_conv_block = false
use_conv_bias = true
gating = "silu"
norm = "RmsNorm"
positional_embedding = "Rope"
conv_layout = false
conv_kernel_size = 3
kv_repeat = 1
max_seq_len = 40960

[modules.asr.model.extra_heads]
num_heads = 4
dim = 6
'''


def _cfg(text, module=None):
    lib = _lib.load()
    c = _lib.AsrConfigC()
    check(lib.dsm_asr_config_from_toml_text(text.encode(), module, C.byref(c)))
    return c


def _same_spec(c, spec):
    for k in ("d_model", "num_heads", "num_layers", "dim_feedforward", "context", "text_in_vocab_size", "text_out_vocab_size",
              "audio_vocab_size", "audio_codebooks", "extra_heads_num", "extra_heads_dim"):
        assert getattr(c.lm, k) == getattr(spec, k), k
    assert c.lm.max_period == spec.max_period and abs(c.lm.rms_eps - spec.rms_eps) < 1e-12


def test_reference_toml_stt_1b_en_fr(tmp_path):
    c = _cfg(TOML_EN_FR)
    _same_spec(c, stt_1b_en_fr())
    assert c.asr_delay_in_tokens == 6 and c.batch_size == 64 and c.temperature == 0.0
    assert c.module_type == b"BatchedAsr"
    assert c.lm_model_file == b"hf://kyutai/stt-1b-en_fr-candle/model.safetensors"
    assert c.audio_tokenizer_file.endswith(b"mimi-pytorch-e351c8d8@125.safetensors")
    assert c.text_tokenizer_file.endswith(b"tokenizer_en_fr_audio_8000.model")
    # and through the file entry point
    p = tmp_path / "cfg.toml"
    p.write_text(TOML_EN_FR)
    c2 = _lib.AsrConfigC()
    check(_lib.load().dsm_asr_config_from_toml(str(p).encode(), b"asr", C.byref(c2)))
    _same_spec(c2, stt_1b_en_fr())


def test_reference_toml_stt_2_6b_shape_and_conditioning_delay():
    # configs/stt/config-stt-en-hf.toml:24-55: 48 layers, 32 heads, context 375, vocab 4001/4000, no extra heads
    t = TOML_EN_FR.replace("num_heads = 16", "num_heads = 32").replace("num_layers = 16", "num_layers = 48") \
        .replace("context = 750", "context = 375").replace("8001", "4001").replace("8000\n", "4000\n") \
        .replace("[modules.asr.model.extra_heads]\nnum_heads = 4\ndim = 6\n", "") \
        .replace("asr_delay_in_tokens = 6", "asr_delay_in_tokens = 32")
    c = _cfg(t)
    _same_spec(c, stt_2_6b_en())
    assert c.asr_delay_in_tokens == 32 and not c.has_conditioning_delay
    # batched_asr.rs:769-770: conditioning_delay overrides: (v * 12.5) as usize + 1
    c = _cfg(t.replace("asr_delay_in_tokens = 32", "asr_delay_in_tokens = 32\nconditioning_delay = 2.0"))
    assert c.has_conditioning_delay and c.asr_delay_in_tokens == 26


def test_toml_refuses_what_is_not_implemented():
    lib = _lib.load()
    c = _lib.AsrConfigC()
    for old, new in (('gating = "silu"', 'gating = "none"'), ("kv_repeat = 1", "kv_repeat = 2"), ('norm = "RmsNorm"', 'norm = "LayerNorm"')):
        assert lib.dsm_asr_config_from_toml_text(TOML_EN_FR.replace(old, new).encode(), None, C.byref(c)) == -5    # DSM_ERR_UNSUPPORTED
    assert lib.dsm_asr_config_from_toml_text(TOML_EN_FR.replace("d_model = 2048", "").encode(), None, C.byref(c)) == -1
