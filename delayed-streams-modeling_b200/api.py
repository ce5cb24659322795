"""Host-side mirror of the reference's ``moshi`` crate surface for the hot path, over the C ABI.

Names and argument meaning follow the Rust methods (paths relative to
/root/reference/server/rust/moshi/moshi-core/src):
  * ``Mimi.encode_step / decode_step / reset_state / reset_batch_idx``   -- mimi.rs:195-245
  * ``LmModel.forward / forward_cond / reset_state / reset_batch_idx``    -- lm.rs:899-1008,1108-1110
  * ``AsrState.step_pcm / step_tokens / reset / reset_batch_idx``         -- asr.rs:104-283
Every call goes through ``libdsm_b200.so``; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import FRAME_SIZE, DsmError, check, ptr
from .spec import LmSpec, MimiSpec


def _mask_arr(mask, b: int) -> Optional[np.ndarray]:
    if mask is None:
        return None
    m = np.ascontiguousarray(np.asarray(mask, dtype=np.uint8))
    if m.shape != (b,):
        raise ValueError(f"mask must have shape ({b},)")
    return m


def mimi_spec_c(spec: MimiSpec, with_encoder=True, with_decoder=True) -> _lib.MimiSpecC:
    lib = _lib.load()
    c = _lib.MimiSpecC()
    lib.dsm_mimi_spec_v0_1(C.byref(c), spec.n_q)
    c.bins, c.qdim, c.dim, c.n_filters = spec.bins, spec.qdim, spec.dim, spec.n_filters
    for i, r in enumerate(spec.ratios):
        c.ratios[i] = r
    c.kernel_size, c.residual_kernel_size, c.last_kernel_size = spec.kernel_size, spec.residual_kernel_size, spec.last_kernel_size
    c.compress, c.num_heads, c.num_layers = spec.compress, spec.num_heads, spec.num_layers
    c.dim_feedforward, c.context, c.max_period, c.ln_eps = spec.dim_feedforward, spec.context, spec.max_period, spec.ln_eps
    c.downsample_stride = spec.downsample_stride
    c.with_encoder, c.with_decoder = int(with_encoder), int(with_decoder)
    return c


def lm_spec_c(spec: LmSpec) -> _lib.LmSpecC:
    c = _lib.LmSpecC()
    c.d_model, c.num_heads, c.num_layers = spec.d_model, spec.num_heads, spec.num_layers
    c.dim_feedforward, c.context, c.max_period = spec.dim_feedforward, spec.context, spec.max_period
    c.text_in_vocab_size, c.text_out_vocab_size = spec.text_in_vocab_size, spec.text_out_vocab_size
    c.audio_vocab_size, c.audio_codebooks = spec.audio_vocab_size, spec.audio_codebooks
    c.extra_heads_num, c.extra_heads_dim, c.rms_eps = spec.extra_heads_num, spec.extra_heads_dim, spec.rms_eps
    return c


class Mimi:
    """``Mimi::batched`` (mimi.rs:113-171).  ``weights=None`` asks the library for a seeded random init
    (benchmarks without checkpoints)."""

    def __init__(self, spec: MimiSpec, weights: Optional[Dict[str, np.ndarray]], batch_size: int, device: int = 0,
                 encoder: bool = True, decoder: bool = True):
        self._lib = _lib.load()
        self.spec, self.batch_size, self.device = spec, batch_size, device
        self._h = C.c_void_p()
        table = _lib.WeightTable(weights) if weights is not None else None
        c = mimi_spec_c(spec, encoder, decoder)
        check(self._lib.dsm_mimi_create(C.byref(c), C.byref(table.c) if table else None, device, batch_size,
                                        C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dsm_mimi_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def encode_step(self, pcm, mask=None, out=None):
        """pcm ``[B, 1920]`` (or ``[B, 1, 1920]``) float32, numpy or CUDA torch tensor -> codes
        ``[B, n_q]`` uint32 (numpy unless ``out`` is a CUDA tensor)."""
        b = self.batch_size
        if isinstance(pcm, np.ndarray):
            pcm = np.ascontiguousarray(pcm.reshape(b, FRAME_SIZE), dtype=np.float32)
        codes = out if out is not None else np.empty((b, self.spec.n_q), np.uint32)
        m = _mask_arr(mask, b)
        produced = C.c_int32(0)
        check(self._lib.dsm_mimi_encode_step(self._h, ptr(pcm), ptr(m), ptr(codes), C.byref(produced)))
        return codes if produced.value else None

    def decode_step(self, codes, mask=None, out=None):
        """codes ``[B, n_q]`` uint32 -> pcm ``[B, 1920]`` float32."""
        b = self.batch_size
        if isinstance(codes, np.ndarray):
            codes = np.ascontiguousarray(codes.reshape(b, self.spec.n_q), dtype=np.uint32)
        pcm = out if out is not None else np.empty((b, FRAME_SIZE), np.float32)
        m = _mask_arr(mask, b)
        produced = C.c_int32(0)
        check(self._lib.dsm_mimi_decode_step(self._h, ptr(codes), ptr(m), ptr(pcm), C.byref(produced)))
        return pcm if produced.value else None

    def last_latent(self) -> np.ndarray:
        out = np.empty((self.batch_size, self.spec.dim), np.float32)
        check(self._lib.dsm_mimi_last_latent(self._h, ptr(out)))
        return out

    def debug_tensor(self, name: str) -> np.ndarray:
        n = C.c_int64(0)
        check(self._lib.dsm_mimi_debug_tensor(self._h, name.encode(), None, 0, C.byref(n)))
        out = np.empty(n.value, np.float32)
        check(self._lib.dsm_mimi_debug_tensor(self._h, name.encode(), ptr(out), n.value, None))
        return out.reshape(self.batch_size, -1, self.spec.dim) if out.size > self.batch_size * self.spec.dim \
            else out.reshape(self.batch_size, self.spec.dim)

    def reset_state(self):
        check(self._lib.dsm_mimi_reset_state(self._h))

    def reset_batch_idx(self, batch_idx: int):
        check(self._lib.dsm_mimi_reset_batch_idx(self._h, batch_idx))

    def synchronize(self):
        check(self._lib.dsm_mimi_synchronize(self._h))

    @property
    def device_bytes(self) -> int:
        return self._lib.dsm_mimi_device_bytes(self._h)

    @property
    def kv_bytes(self) -> int:
        return self._lib.dsm_mimi_kv_bytes(self._h)


class LmModel:
    """``LmModel::batched`` (lm.rs:812-897) without depformer."""

    KV_DTYPES = {"bf16": _lib.KV_BF16, "int8": _lib.KV_INT8, "int8k": _lib.KV_INT8K}

    def __init__(self, spec: LmSpec, weights: Optional[Dict[str, np.ndarray]], batch_size: int, device: int = 0,
                 kv_dtype: str = "bf16"):
        """``kv_dtype``: storage of the ring KV cache, "bf16" (the reference's, default), "int8" (opt-in, half the
        bytes) or "int8k" (opt-in, int8 K + bf16 V); see ``dsm_lm_options`` in the header."""
        self._lib = _lib.load()
        self.spec, self.batch_size, self.device = spec, batch_size, device
        self._h = C.c_void_p()
        table = _lib.WeightTable(weights) if weights is not None else None
        c = lm_spec_c(spec)
        opt = _lib.LmOptionsC()
        opt.kv_dtype = self.KV_DTYPES[kv_dtype]
        self.kv_dtype = kv_dtype
        check(self._lib.dsm_lm_create_ex(C.byref(c), C.byref(table.c) if table else None, device, batch_size,
                                         C.byref(opt), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dsm_lm_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def forward_cond(self, text_ids, audio_ids, cond=None, mask=None, want_ys=True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        """lm.rs:957-1008: text_ids ``[B]``, audio_ids ``[B, n_cb]`` -> (logits ``[B, V]``, ys ``[B, D]``)."""
        b = self.batch_size
        t = np.ascontiguousarray(np.asarray(text_ids, dtype=np.uint32).reshape(b))
        a = np.ascontiguousarray(np.asarray(audio_ids, dtype=np.uint32).reshape(b, self.spec.audio_codebooks))
        c = None if cond is None else np.ascontiguousarray(np.asarray(cond, dtype=np.float32).reshape(self.spec.d_model))
        logits = np.empty((b, self.spec.text_out_vocab_size), np.float32)
        ys = np.empty((b, self.spec.d_model), np.float32) if want_ys else None
        check(self._lib.dsm_lm_forward(self._h, ptr(t), ptr(a), ptr(c), ptr(_mask_arr(mask, b)), ptr(logits), ptr(ys)))
        return logits, ys

    def forward(self, text_ids, audio_ids, mask=None):
        return self.forward_cond(text_ids, audio_ids, None, mask)

    def reset_state(self):
        check(self._lib.dsm_lm_reset_state(self._h))

    def reset_batch_idx(self, batch_idx: int):
        check(self._lib.dsm_lm_reset_batch_idx(self._h, batch_idx))

    def synchronize(self):
        check(self._lib.dsm_lm_synchronize(self._h))

    def text_start_token(self) -> int:
        return self._lib.dsm_lm_text_start_token(self._h)

    def audio_pad_token(self) -> int:
        return self._lib.dsm_lm_audio_pad_token(self._h)

    def in_audio_codebooks(self) -> int:
        return self._lib.dsm_lm_in_audio_codebooks(self._h)

    def stream(self) -> int:
        """Raw cudaStream_t of the handle (wrap with torch.cuda.ExternalStream to record events on it)."""
        return self._lib.dsm_lm_stream(self._h)

    def debug_set_position(self, pos: int):
        check(self._lib.dsm_lm_debug_set_position(self._h, pos))

    def profile_attention(self, enable: bool):
        check(self._lib.dsm_lm_profile_attention(self._h, int(enable)))

    def profile_read(self) -> Tuple[float, int]:
        ms, n = C.c_double(0), C.c_longlong(0)
        check(self._lib.dsm_lm_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def kv_bytes(self) -> int:
        return self._lib.dsm_lm_kv_bytes(self._h)

    @property
    def weight_bytes(self) -> int:
        return self._lib.dsm_lm_weight_bytes(self._h)


class AsrState:
    """``asr::State`` (asr.rs:53-284).  Messages mirror ``AsrMsg`` as tuples:
    ``('step', step_idx, prs[n_heads, B])``, ``('word', tokens, start_time, batch_idx)``,
    ``('end_word', stop_time, batch_idx)``."""

    def __init__(self, mimi: Optional[Mimi], lm: LmModel, asr_delay_in_tokens: Optional[int] = None,
                 temperature: float = 0.0):
        self._lib = _lib.load()
        self.mimi, self.lm = mimi, lm
        delay = lm.spec.asr_delay_in_tokens if asr_delay_in_tokens is None else asr_delay_in_tokens
        self._h = C.c_void_p()
        check(self._lib.dsm_asr_create(mimi._h if mimi is not None else None, lm._h, delay, temperature,
                                       C.byref(self._h)))
        self.last_tokens: Optional[np.ndarray] = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dsm_asr_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def batch_size(self) -> int:
        return self._lib.dsm_asr_batch_size(self._h)

    def model_step_idx(self) -> int:
        return self._lib.dsm_asr_model_step_idx(self._h)

    def asr_delay_in_tokens(self) -> int:
        return self._lib.dsm_asr_delay_in_tokens(self._h)

    def _messages(self) -> List[tuple]:
        out = []
        b = self.batch_size()
        nh = self.lm.spec.extra_heads_num
        m = _lib.AsrMsgC()
        for i in range(self._lib.dsm_asr_msg_count(self._h)):
            check(self._lib.dsm_asr_msg_get(self._h, i, C.byref(m)))
            if m.kind == _lib.MSG_STEP:
                prs = np.ctypeslib.as_array(m.prs, shape=(nh * b,)).reshape(nh, b).copy()
                out.append(("step", m.step_idx, prs))
            elif m.kind == _lib.MSG_WORD:
                out.append(("word", [int(m.tokens[j]) for j in range(m.n_tokens)], m.time, m.batch_idx))
            else:
                out.append(("end_word", m.time, m.batch_idx))
        return out

    def step_pcm(self, pcm, mask=None, want_messages: bool = True) -> List[tuple]:
        """asr.rs:115-131.  pcm ``[B, 1920]`` float32 (numpy / pinned or CUDA torch tensor)."""
        b = self.batch_size()
        if isinstance(pcm, np.ndarray):
            pcm = np.ascontiguousarray(pcm.reshape(b, FRAME_SIZE), dtype=np.float32)
        toks = np.empty(b, np.uint32)
        check(self._lib.dsm_asr_step_pcm(self._h, ptr(pcm), ptr(_mask_arr(mask, b)), ptr(toks)))
        self.last_tokens = toks
        return self._messages() if want_messages else []

    def step_tokens(self, codes, mask=None, want_messages: bool = True) -> List[tuple]:
        """asr.rs:147-255 with one step of codes ``[B, n_cb]`` (``[B, n_cb, 1]`` accepted)."""
        b = self.batch_size()
        n_cb = self.lm.spec.audio_codebooks
        if isinstance(codes, np.ndarray):
            codes = np.ascontiguousarray(codes.reshape(b, n_cb), dtype=np.uint32)
        toks = np.empty(b, np.uint32)
        check(self._lib.dsm_asr_step_tokens(self._h, ptr(codes), ptr(_mask_arr(mask, b)), ptr(toks)))
        self.last_tokens = toks
        return self._messages() if want_messages else []

    def last_logits(self) -> np.ndarray:
        out = np.empty((self.batch_size(), self.lm.spec.text_out_vocab_size), np.float32)
        check(self._lib.dsm_asr_last_logits(self._h, ptr(out)))
        return out

    def last_step_ms(self) -> Tuple[float, float, float]:
        arr = (C.c_float * 3)()
        check(self._lib.dsm_asr_last_step_ms(self._h, C.byref(arr)))
        return arr[0], arr[1], arr[2]

    def sm_partition(self) -> Tuple[int, int]:
        """(SMs of the LM branch, SMs of the codec branch); (0, 0) when the branches share the device."""
        arr = (C.c_int32 * 2)()
        check(self._lib.dsm_asr_sm_partition(self._h, C.byref(arr)))
        return arr[0], arr[1]

    def reset(self):
        check(self._lib.dsm_asr_reset(self._h))

    def reset_batch_idx(self, batch_idx: int):
        check(self._lib.dsm_asr_reset_batch_idx(self._h, batch_idx))


class BatchedAsr:
    """Slot scheduler of moshi-server's ``BatchedAsr`` (batched_asr.rs:216-717) over ``dsm_sched_*``.
    ``asr=None`` builds a dry scheduler (host logic only) with ``batch_size`` slots."""

    KINDS = {0: "step", 1: "word", 2: "end_word", 3: "marker", 4: "ready"}

    def __init__(self, asr, batch_size: int = 0, n_devices: int = 1):
        """``asr``: one ``AsrState``, a list of them (one per GPU: device g owns slots [g*B, (g+1)*B)), or None for a
        dry scheduler over ``n_devices`` x ``batch_size`` slots."""
        self._lib = _lib.load()
        self._h = C.c_void_p()
        asrs = list(asr) if isinstance(asr, (list, tuple)) else ([asr] if asr is not None else [])
        self.asr = asrs[0] if asrs else None
        self._asrs = asrs
        if len(asrs) > 1 or (not asrs and n_devices > 1):
            n = len(asrs) or n_devices
            arr = (C.c_void_p * n)(*[a._h for a in asrs]) if asrs else None
            check(self._lib.dsm_sched_create_multi(arr, n, batch_size, C.byref(self._h)))
        else:
            check(self._lib.dsm_sched_create(asrs[0]._h if asrs else None, batch_size, C.byref(self._h)))
        self.n_devices = self._lib.dsm_sched_num_devices(self._h)
        self.slots_per_device = self._lib.dsm_sched_slots_per_device(self._h)
        self.batch_size = self.slots_per_device            # rows of last_mask() / batch_pcm() (device 0)

    def close_handle(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dsm_sched_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close_handle

    def open(self) -> Tuple[int, int]:
        slot, cid = C.c_int32(-1), C.c_uint64(0)
        check(self._lib.dsm_sched_open(self._h, C.byref(slot), C.byref(cid)))
        return slot.value, cid.value

    def close(self, slot: int, channel_id: int = 0):
        check(self._lib.dsm_sched_close_ch(self._h, slot, channel_id))

    def push_audio(self, slot: int, pcm, channel_id: int = 0):
        a = np.ascontiguousarray(np.asarray(pcm, dtype=np.float32).reshape(-1))
        check(self._lib.dsm_sched_push_audio_ch(self._h, slot, channel_id, ptr(a), a.size))

    def push_marker(self, slot: int, marker_id: int, channel_id: int = 0):
        check(self._lib.dsm_sched_push_marker_ch(self._h, slot, channel_id, marker_id))

    def start(self):
        """One pipelined tick thread per device; drain messages with ``poll``."""
        check(self._lib.dsm_sched_start(self._h))

    def stop(self):
        check(self._lib.dsm_sched_stop(self._h))

    def poll(self, slot: int) -> Optional[dict]:
        ev, got = _lib.SchedEventC(), C.c_int32(0)
        check(self._lib.dsm_sched_poll(self._h, slot, C.byref(ev), C.byref(got)))
        if not got.value:
            return None
        d = {"kind": self.KINDS[ev.kind], "slot": ev.slot, "channel_id": ev.channel_id, "step_idx": ev.step_idx}
        if ev.kind == 1:
            d.update(tokens=[int(ev.tokens[j]) for j in range(ev.n_tokens)], start_time=ev.time)
        elif ev.kind == 2:
            d.update(stop_time=ev.time)
        elif ev.kind == 0:
            d.update(prs=[float(ev.prs[j]) for j in range(ev.n_prs)], buffered_pcm=ev.buffered_pcm)
        elif ev.kind == 3:
            d.update(marker_id=ev.marker_id)
        return d

    def stats(self, device: int = 0) -> dict:
        st = _lib.SchedStatsC()
        check(self._lib.dsm_sched_get_stats(self._h, device, C.byref(st)))
        return {n: getattr(st, n) for n, _ in st._fields_}

    def set_dry_step_us(self, us: int):
        self._lib.dsm_sched_set_dry_step_us(self._h, us)

    def device_step_idx(self, device: int) -> int:
        return self._lib.dsm_sched_device_step_idx(self._h, device)

    def tick(self) -> Tuple[bool, List[dict]]:
        stepped = C.c_int32(0)
        check(self._lib.dsm_sched_tick(self._h, C.byref(stepped)))
        outs = []
        o = _lib.SchedOutC()
        for i in range(self._lib.dsm_sched_out_count(self._h)):
            check(self._lib.dsm_sched_out_get(self._h, i, C.byref(o)))
            d = {"kind": self.KINDS[o.kind], "slot": o.slot, "channel_id": o.channel_id, "step_idx": o.step_idx}
            if o.kind == 1:
                d.update(tokens=[int(o.tokens[j]) for j in range(o.n_tokens)], start_time=o.time)
            elif o.kind == 2:
                d.update(stop_time=o.time)
            elif o.kind == 0:
                d.update(prs=[float(o.prs[j]) for j in range(o.n_prs)], buffered_pcm=o.buffered_pcm)
            elif o.kind == 3:
                d.update(marker_id=o.marker_id)
            outs.append(d)
        return bool(stepped.value), outs

    def free_slots(self) -> int:
        return self._lib.dsm_sched_free_slots(self._h)

    def step_idx(self) -> int:
        return self._lib.dsm_sched_step_idx(self._h)

    def last_mask(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._lib.dsm_sched_last_mask(self._h), shape=(self.batch_size,)).copy()

    def batch_pcm(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._lib.dsm_sched_batch_pcm(self._h), shape=(self.batch_size, FRAME_SIZE)).copy()


def set_graphs(enabled: bool):
    _lib.load().dsm_set_graphs(int(enabled))


def launch_count() -> int:
    return _lib.load().dsm_launch_count()


def launch_count_reset():
    _lib.load().dsm_launch_count_reset()
