"""ctypes binding of ``libdsm_b200.so`` (the C ABI in ``include/dsm_b200.h``).

There is no fallback: if the shared library is missing, ``load()`` raises and every product class
refuses to construct.  Build it with ``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C delayed-streams-modeling_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSM_B200_LIB") or os.path.join(_HERE, "libdsm_b200.so")   # override: A/B builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dsm_b200.h")

DSM_F32, DSM_BF16, DSM_F16, DSM_U32, DSM_I64 = 0, 1, 2, 3, 4
FRAME_SIZE = 1920
MSG_STEP, MSG_WORD, MSG_END_WORD = 0, 1, 2


class DsmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dsm_b200 error {code}: {msg}")
        self.code = code


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int32), ("ndim", C.c_int32), ("shape", C.c_int64 * 4),
                ("data", C.c_void_p)]


class Weights(C.Structure):
    _fields_ = [("tensors", C.POINTER(Tensor)), ("count", C.c_int64)]


class MimiSpecC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_q", "bins", "qdim", "dim", "n_filters")] + [("ratios", C.c_int32 * 4)] + \
               [(n, C.c_int32) for n in ("kernel_size", "residual_kernel_size", "last_kernel_size", "compress",
                                         "num_heads", "num_layers", "dim_feedforward", "context")] + \
               [("max_period", C.c_float), ("ln_eps", C.c_float), ("downsample_stride", C.c_int32),
                ("with_encoder", C.c_int32), ("with_decoder", C.c_int32)]


class LmSpecC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("d_model", "num_heads", "num_layers", "dim_feedforward", "context")] + \
               [("max_period", C.c_float)] + \
               [(n, C.c_int32) for n in ("text_in_vocab_size", "text_out_vocab_size", "audio_vocab_size",
                                         "audio_codebooks", "extra_heads_num", "extra_heads_dim")] + \
               [("rms_eps", C.c_float)]


class AsrMsgC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("batch_idx", C.c_int32), ("step_idx", C.c_int32), ("n_tokens", C.c_int32),
                ("tokens", C.POINTER(C.c_uint32)), ("time", C.c_double), ("prs", C.POINTER(C.c_float))]


class AsrConfigC(C.Structure):
    _fields_ = [("lm", LmSpecC), ("asr_delay_in_tokens", C.c_int32), ("has_conditioning_delay", C.c_int32),
                ("conditioning_delay", C.c_float), ("temperature", C.c_float), ("batch_size", C.c_int32),
                ("module_type", C.c_char * 32), ("lm_model_file", C.c_char * 512), ("audio_tokenizer_file", C.c_char * 512),
                ("text_tokenizer_file", C.c_char * 512)]


class WireInC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("id", C.c_int64), ("n", C.c_int64), ("pcm", C.POINTER(C.c_float)),
                ("data", C.POINTER(C.c_uint8))]


class WireOutC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("text", C.c_void_p), ("text_len", C.c_int32), ("time", C.c_double), ("id", C.c_int64),
                ("step_idx", C.c_int64), ("prs", C.POINTER(C.c_float)), ("n_prs", C.c_int32), ("buffered_pcm", C.c_int64)]


class SchedEventC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("slot", C.c_int32), ("step_idx", C.c_int32), ("n_tokens", C.c_int32), ("n_prs", C.c_int32),
                ("channel_id", C.c_uint64), ("time", C.c_double), ("marker_id", C.c_int64), ("buffered_pcm", C.c_int64),
                ("tokens", C.c_uint32 * 64), ("prs", C.c_float * 8)]


class SchedStatsC(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("ticks", "steps", "idle_sleeps", "frames", "gather_ns", "step_wait_ns", "fanout_ns")]


class LmOptionsC(C.Structure):
    _fields_ = [("kv_dtype", C.c_int32), ("reserved", C.c_int32 * 7)]


KV_BF16, KV_INT8, KV_INT8K = 0, 1, 2


class SchedOutC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("slot", C.c_int32), ("channel_id", C.c_uint64), ("step_idx", C.c_int32),
                ("n_tokens", C.c_int32), ("tokens", C.POINTER(C.c_uint32)), ("time", C.c_double), ("n_prs", C.c_int32),
                ("prs", C.POINTER(C.c_float)), ("marker_id", C.c_int64), ("buffered_pcm", C.c_int64)]


_lib = None

# name -> (restype, argtypes); every symbol the header declares
_P = C.c_void_p
SIGNATURES = {
    "dsm_version": (C.c_char_p, []),
    "dsm_last_error": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "dsm_device_count": (C.c_int32, []),
    "dsm_set_graphs": (None, [C.c_int32]),
    "dsm_launch_count": (C.c_longlong, []),
    "dsm_launch_count_reset": (None, []),
    "dsm_debug_gemm_bf16": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "dsm_debug_gemm_f32": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P,
                                       C.c_int64, _P, _P, _P]),
    "dsm_debug_ring_trace": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "dsm_mimi_spec_v0_1": (None, [C.POINTER(MimiSpecC), C.c_int32]),
    "dsm_lm_spec_stt_1b_en_fr": (None, [C.POINTER(LmSpecC)]),
    "dsm_lm_spec_stt_2_6b_en": (None, [C.POINTER(LmSpecC)]),
    "dsm_mimi_create": (C.c_int32, [C.POINTER(MimiSpecC), C.POINTER(Weights), C.c_int32, C.c_int32, C.POINTER(_P)]),
    "dsm_mimi_destroy": (None, [_P]),
    "dsm_mimi_encode_step": (C.c_int32, [_P, _P, _P, _P, C.POINTER(C.c_int32)]),
    "dsm_mimi_decode_step": (C.c_int32, [_P, _P, _P, _P, C.POINTER(C.c_int32)]),
    "dsm_mimi_last_latent": (C.c_int32, [_P, _P]),
    "dsm_mimi_debug_tensor": (C.c_int32, [_P, C.c_char_p, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "dsm_mimi_reset_state": (C.c_int32, [_P]),
    "dsm_mimi_reset_batch_idx": (C.c_int32, [_P, C.c_int32]),
    "dsm_mimi_synchronize": (C.c_int32, [_P]),
    "dsm_mimi_batch_size": (C.c_int32, [_P]),
    "dsm_mimi_n_q": (C.c_int32, [_P]),
    "dsm_mimi_device_bytes": (C.c_int64, [_P]),
    "dsm_mimi_kv_bytes": (C.c_int64, [_P]),
    "dsm_lm_create": (C.c_int32, [C.POINTER(LmSpecC), C.POINTER(Weights), C.c_int32, C.c_int32, C.POINTER(_P)]),
    "dsm_lm_create_ex": (C.c_int32, [C.POINTER(LmSpecC), C.POINTER(Weights), C.c_int32, C.c_int32, C.POINTER(LmOptionsC),
                                     C.POINTER(_P)]),
    "dsm_lm_kv_dtype": (C.c_int32, [_P]),
    "dsm_lm_destroy": (None, [_P]),
    "dsm_lm_forward": (C.c_int32, [_P, _P, _P, _P, _P, _P, _P]),
    "dsm_lm_reset_state": (C.c_int32, [_P]),
    "dsm_lm_reset_batch_idx": (C.c_int32, [_P, C.c_int32]),
    "dsm_lm_synchronize": (C.c_int32, [_P]),
    "dsm_lm_text_start_token": (C.c_int32, [_P]),
    "dsm_lm_audio_pad_token": (C.c_int32, [_P]),
    "dsm_lm_in_audio_codebooks": (C.c_int32, [_P]),
    "dsm_lm_stream": (C.c_void_p, [_P]),
    "dsm_mimi_stream": (C.c_void_p, [_P]),
    "dsm_lm_profile_attention": (C.c_int32, [_P, C.c_int32]),
    "dsm_lm_profile_read": (C.c_int32, [_P, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "dsm_lm_debug_set_position": (C.c_int32, [_P, C.c_uint32]),
    "dsm_lm_kv_bytes": (C.c_int64, [_P]),
    "dsm_lm_weight_bytes": (C.c_int64, [_P]),
    "dsm_asr_create": (C.c_int32, [_P, _P, C.c_int32, C.c_float, C.POINTER(_P)]),
    "dsm_asr_destroy": (None, [_P]),
    "dsm_asr_step_pcm": (C.c_int32, [_P, _P, _P, _P]),
    "dsm_asr_step_tokens": (C.c_int32, [_P, _P, _P, _P]),
    "dsm_asr_msg_count": (C.c_int32, [_P]),
    "dsm_asr_msg_get": (C.c_int32, [_P, C.c_int32, C.POINTER(AsrMsgC)]),
    "dsm_asr_reset": (C.c_int32, [_P]),
    "dsm_asr_reset_batch_idx": (C.c_int32, [_P, C.c_int32]),
    "dsm_asr_batch_size": (C.c_int32, [_P]),
    "dsm_asr_model_step_idx": (C.c_int32, [_P]),
    "dsm_asr_delay_in_tokens": (C.c_int32, [_P]),
    "dsm_asr_last_logits": (C.c_int32, [_P, _P]),
    "dsm_asr_last_step_ms": (C.c_int32, [_P, C.POINTER(C.c_float * 3)]),
    "dsm_asr_sm_partition": (C.c_int32, [_P, C.POINTER(C.c_int32 * 2)]),
    "dsm_weights_open": (C.c_int32, [C.c_char_p, C.POINTER(_P)]),
    "dsm_weight_file_table": (C.POINTER(Weights), [_P]),
    "dsm_weight_file_metadata": (C.c_int64, [_P, C.c_char_p, C.c_char_p, C.c_size_t]),
    "dsm_weight_file_close": (None, [_P]),
    "dsm_weights_save": (C.c_int32, [C.c_char_p, C.POINTER(Weights)]),
    "dsm_asr_config_from_toml": (C.c_int32, [C.c_char_p, C.c_char_p, C.POINTER(AsrConfigC)]),
    "dsm_asr_config_from_toml_text": (C.c_int32, [C.c_char_p, C.c_char_p, C.POINTER(AsrConfigC)]),
    "dsm_wire_encode_out": (C.c_int32, [C.POINTER(WireOutC), _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dsm_wire_decode_out": (C.c_int32, [_P, C.c_size_t, C.POINTER(WireOutC), _P, C.c_int32]),
    "dsm_wire_encode_in": (C.c_int32, [C.POINTER(WireInC), _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dsm_wire_decode_in": (C.c_int32, [_P, C.c_size_t, C.POINTER(WireInC), _P, C.c_int64, _P, C.c_int64]),
    "dsm_spm_open": (C.c_int32, [C.c_char_p, C.POINTER(_P)]),
    "dsm_spm_close": (None, [_P]),
    "dsm_spm_vocab_size": (C.c_int32, [_P]),
    "dsm_spm_piece": (C.c_char_p, [_P, C.c_int32]),
    "dsm_spm_decode": (C.c_int32, [_P, _P, C.c_int32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dsm_sched_create": (C.c_int32, [_P, C.c_int32, C.POINTER(_P)]),
    "dsm_sched_create_multi": (C.c_int32, [C.POINTER(_P), C.c_int32, C.c_int32, C.POINTER(_P)]),
    "dsm_sched_num_devices": (C.c_int32, [_P]),
    "dsm_sched_slots_per_device": (C.c_int32, [_P]),
    "dsm_sched_push_audio_ch": (C.c_int32, [_P, C.c_int32, C.c_uint64, _P, C.c_int64]),
    "dsm_sched_push_marker_ch": (C.c_int32, [_P, C.c_int32, C.c_uint64, C.c_int64]),
    "dsm_sched_close_ch": (C.c_int32, [_P, C.c_int32, C.c_uint64]),
    "dsm_sched_start": (C.c_int32, [_P]),
    "dsm_sched_stop": (C.c_int32, [_P]),
    "dsm_sched_poll": (C.c_int32, [_P, C.c_int32, C.POINTER(SchedEventC), C.POINTER(C.c_int32)]),
    "dsm_sched_get_stats": (C.c_int32, [_P, C.c_int32, C.POINTER(SchedStatsC)]),
    "dsm_sched_set_dry_step_us": (None, [_P, C.c_int32]),
    "dsm_sched_device_step_idx": (C.c_int64, [_P, C.c_int32]),
    "dsm_sched_push_msg": (C.c_int32, [_P, C.c_int32, C.c_uint64, _P, C.c_size_t]),
    "dsm_sched_poll_msg": (C.c_int32, [_P, C.c_int32, _P, _P, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    "dsm_session_create": (C.c_int32, [_P, _P, C.POINTER(_P)]),
    "dsm_session_destroy": (None, [_P]),
    "dsm_session_push_msg": (C.c_int32, [_P, _P, C.c_size_t]),
    "dsm_session_poll_msg": (C.c_int32, [_P, _P, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    "dsm_session_steps": (C.c_int64, [_P]),
    "dsm_session_save_log": (C.c_int32, [_P, C.c_char_p]),
    "dsm_tts_sink_create": (C.c_int32, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "dsm_tts_sink_destroy": (None, [_P]),
    "dsm_tts_sink_push": (C.c_int32, [_P, _P, C.c_int64, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dsm_tts_sink_decode_all": (C.c_int32, [_P, _P, C.c_int64, _P]),
    "dsm_sched_destroy": (None, [_P]),
    "dsm_sched_open": (C.c_int32, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "dsm_sched_close": (C.c_int32, [_P, C.c_int32]),
    "dsm_sched_push_audio": (C.c_int32, [_P, C.c_int32, _P, C.c_int64]),
    "dsm_sched_push_marker": (C.c_int32, [_P, C.c_int32, C.c_int64]),
    "dsm_sched_tick": (C.c_int32, [_P, C.POINTER(C.c_int32)]),
    "dsm_sched_out_count": (C.c_int32, [_P]),
    "dsm_sched_out_get": (C.c_int32, [_P, C.c_int32, C.POINTER(SchedOutC)]),
    "dsm_sched_free_slots": (C.c_int32, [_P]),
    "dsm_sched_step_idx": (C.c_int64, [_P]),
    "dsm_sched_last_mask": (C.POINTER(C.c_uint8), [_P]),
    "dsm_sched_batch_pcm": (C.POINTER(C.c_float), [_P]),
}


def load():
    """Loads the shared library (once) and types every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it (make -C {os.path.join(_HERE, 'csrc')}); "
                          "dsm_b200 has no CPU or PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError here means the .so is stale w.r.t. the header
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def last_error() -> str:
    lib = load()
    buf = C.create_string_buffer(2048)
    lib.dsm_last_error(buf, 2048)
    return buf.value.decode(errors="replace")


def check(status: int):
    if status != 0:
        raise DsmError(status, last_error())


def ptr(x) -> Optional[int]:
    """Address of a numpy array (host), a torch tensor (host or CUDA), or None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    raise TypeError(type(x))


class WeightTable:
    """Keeps the numpy arrays and the C table alive for the duration of a create call."""

    def __init__(self, weights: Dict[str, np.ndarray]):
        self._keep: List[object] = []
        n = len(weights)
        arr = (Tensor * n)()
        for i, (name, a) in enumerate(weights.items()):
            if hasattr(a, "numpy"):
                a = a.detach().cpu()
                if str(a.dtype) == "torch.bfloat16":      # numpy has no bf16: hand the bit patterns over as DSM_BF16
                    import torch
                    a = a.contiguous().view(torch.int16).numpy().view(np.uint16)
                else:
                    a = a.numpy()
            if a.dtype == np.float32:
                dt = DSM_F32
            elif a.dtype == np.float16:
                dt = DSM_F16
            elif a.dtype == np.uint16:      # bf16 bit patterns
                dt = DSM_BF16
            else:
                a = a.astype(np.float32)
                dt = DSM_F32
            a = np.ascontiguousarray(a)
            bname = name.encode()
            self._keep += [a, bname]
            arr[i].name = bname
            arr[i].dtype = dt
            arr[i].ndim = min(a.ndim, 4)
            shape = list(a.shape) if a.ndim <= 4 else [a.size]
            if a.ndim > 4:
                arr[i].ndim = 1
            for j, s in enumerate(shape):
                arr[i].shape[j] = s
            arr[i].data = a.ctypes.data
        self._arr = arr
        self.c = Weights(arr, n)
