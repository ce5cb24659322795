"""Architecture constants of the hot path (no compute here).

Sources (relative to /root/reference):
  * Mimi v0_1:      server/rust/moshi/moshi-core/src/mimi.rs:32-93
  * stt-1b-en_fr:   configs/stt/config-stt-en_fr-hf.toml:27-56
  * stt-2.6b-en:    configs/stt/config-stt-en-hf.toml:24-55
  * gated hidden:   server/rust/moshi/moshi-core/src/batched_transformer.rs:153-157
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import List, Tuple

FRAME_SIZE = 1920          # server/rust/moshi/moshi-server/src/batched_asr.rs:26
SAMPLE_RATE = 24_000
FRAME_RATE = 12.5


@dataclass(frozen=True)
class MimiSpec:
    n_q: int = 32                      # batched_asr.rs:754 (STT/TTS use 32 codebooks)
    bins: int = 2048
    qdim: int = 256
    dim: int = 512                     # seanet dimension == transformer d_model
    n_filters: int = 64
    ratios: Tuple[int, ...] = (8, 6, 5, 4)   # encoder iterates them reversed (seanet.rs:194)
    kernel_size: int = 7
    residual_kernel_size: int = 3
    last_kernel_size: int = 3
    compress: int = 2
    num_heads: int = 8
    num_layers: int = 8
    dim_feedforward: int = 2048
    context: int = 250
    max_period: float = 10000.0
    layer_scale_init: float = 0.01
    ln_eps: float = 1e-5
    downsample_stride: int = 2         # 25 Hz -> 12.5 Hz (mimi.rs:137-140)

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads

    def encoder_convs(self) -> List[dict]:
        """Conv layers of SeaNetEncoder in execution order (seanet.rs:160-245).

        Each entry: name (weight path prefix below ``encoder.model.``), cin, cout, k, stride,
        elu (ELU applied to the input), res (``None`` | 'open' | 'close' for the two convs of a
        resnet block; 'close' adds the block input).
        """
        out = []
        mult = 1
        nf = self.n_filters
        out.append(dict(name="0", cin=1, cout=nf, k=self.kernel_size, stride=1, elu=False, res=None))
        idx = 1
        for ratio in reversed(self.ratios):
            c = mult * nf
            h = c // self.compress
            out.append(dict(name=f"{idx}.block.1", cin=c, cout=h, k=self.residual_kernel_size,
                            stride=1, elu=True, res="open"))
            out.append(dict(name=f"{idx}.block.3", cin=h, cout=c, k=1, stride=1, elu=True, res="close"))
            out.append(dict(name=f"{idx + 2}", cin=c, cout=2 * c, k=2 * ratio, stride=ratio,
                            elu=True, res=None))
            idx += 3
            mult *= 2
        out.append(dict(name=f"{idx + 1}", cin=mult * nf, cout=self.dim, k=self.last_kernel_size,
                        stride=1, elu=True, res=None))
        return out

    def decoder_layers(self) -> List[dict]:
        """Layers of SeaNetDecoder in execution order (seanet.rs:312-399).  ``op`` is 'conv' or
        'convtr'."""
        out = []
        nf = self.n_filters
        mult = 1 << len(self.ratios)
        out.append(dict(op="conv", name="0", cin=self.dim, cout=mult * nf, k=self.kernel_size,
                        stride=1, elu=False, res=None))
        idx = 1
        for ratio in self.ratios:
            c = mult * nf
            out.append(dict(op="convtr", name=f"{idx + 1}", cin=c, cout=c // 2, k=2 * ratio,
                            stride=ratio, elu=True, res=None))
            c2 = c // 2
            h = c2 // self.compress
            out.append(dict(op="conv", name=f"{idx + 2}.block.1", cin=c2, cout=h,
                            k=self.residual_kernel_size, stride=1, elu=True, res="open"))
            out.append(dict(op="conv", name=f"{idx + 2}.block.3", cin=h, cout=c2, k=1, stride=1,
                            elu=True, res="close"))
            idx += 3
            mult //= 2
        out.append(dict(op="conv", name=f"{idx + 1}", cin=nf, cout=1, k=self.last_kernel_size,
                        stride=1, elu=True, res=None))
        return out


@dataclass(frozen=True)
class LmSpec:
    name: str = "stt-1b-en_fr"
    d_model: int = 2048
    num_heads: int = 16
    num_layers: int = 16
    dim_feedforward: int = 8192
    context: int = 750
    max_period: float = 100000.0
    text_in_vocab_size: int = 8001
    text_out_vocab_size: int = 8000
    audio_vocab_size: int = 2049
    audio_codebooks: int = 32
    extra_heads_num: int = 4
    extra_heads_dim: int = 6
    rms_eps: float = 1e-8              # batched_transformer.rs:246-248
    asr_delay_in_tokens: int = 6

    @property
    def head_dim(self) -> int:
        return self.d_model // self.num_heads

    @property
    def hidden(self) -> int:
        # batched_transformer.rs:153-157
        if self.dim_feedforward == 4 * self.d_model:
            return 11 * self.d_model // 4
        return 2 * self.dim_feedforward // 3

    @property
    def text_start_token(self) -> int:   # lm.rs:916-918
        return self.text_in_vocab_size - 1

    @property
    def audio_pad_token(self) -> int:    # lm.rs:912-914
        return self.audio_vocab_size - 1

    def replace(self, **kw) -> "LmSpec":
        return dataclasses.replace(self, **kw)


def stt_1b_en_fr() -> LmSpec:
    return LmSpec()


def stt_2_6b_en() -> LmSpec:
    return LmSpec(name="stt-2.6b-en", d_model=2048, num_heads=32, num_layers=48, dim_feedforward=8192,
                  context=375, text_in_vocab_size=4001, text_out_vocab_size=4000,
                  extra_heads_num=0, extra_heads_dim=0, asr_delay_in_tokens=32)


def tiny_lm(**kw) -> LmSpec:
    """A shrunken LM of the same structure, for CPU-sized parity cases."""
    base = dict(name="tiny", d_model=256, num_heads=2, num_layers=2, dim_feedforward=1024, context=12,
                text_in_vocab_size=257, text_out_vocab_size=256, audio_vocab_size=2049,
                audio_codebooks=32, extra_heads_num=4, extra_heads_dim=6, asr_delay_in_tokens=2)
    base.update(kw)
    return LmSpec(**base)


def tiny_mimi(**kw) -> MimiSpec:
    """Mimi with a short ring so wrap-around is reached in a few frames."""
    base = dict(context=6, num_layers=2)
    base.update(kw)
    return MimiSpec(**base)
