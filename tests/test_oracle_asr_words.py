"""Word / timestamp state machine of asr::State (asr.rs:218-252) on a scripted token sequence."""
import torch

from oracle.lm import AsrState


class _ScriptedLm:
    """Stands in for LmModel: emits logits whose argmax follows a per-stream script."""

    class _Spec:
        text_start_token, audio_pad_token, audio_codebooks, extra_heads_num = 8, 2048, 2, 0

    def __init__(self, script):
        self.spec, self.script, self.t = self._Spec(), script, 0

    def forward_cond(self, text, audio, cond, mask):
        b = len(self.script)
        logits = torch.zeros(b, 1, 9)
        for i in range(b):
            logits[i, 0, self.script[i][self.t]] = 1.0
        self.t += 1
        return logits, torch.zeros(b, 1, 4)

    def extra_heads(self, ys):
        return []

    def reset_batch_idx(self, b):
        pass


def test_word_and_end_word_messages():
    # delay 2: tokens before step_idx >= 2 are ignored; 5,6 form a word flushed by 3 (pad) then 0 ends it
    script = [[7, 5, 6, 3, 0, 4, 0, 0]]
    asr = AsrState(1, 2, None, _ScriptedLm(script))
    codes = torch.zeros(1, 2, 1, dtype=torch.int64)
    msgs = []
    for _ in range(8):
        msgs += asr.step_tokens(codes, None)
    assert msgs == [
        ("word", [5, 6], 0.0, 0),             # flushed by token 3 at step 4
        ("end_word", (5 - 2) / 12.5, 0),      # token 0 at step 5
        ("word", [4], (5 - 2) / 12.5, 0),     # token 0 at step 7 flushes the word ...
        ("end_word", (7 - 2) / 12.5, 0),      # ... and ends it
    ]


def test_inactive_stream_is_untouched():
    script = [[5, 5, 0, 0], [6, 6, 0, 0]]
    asr = AsrState(2, 1, None, _ScriptedLm(script))
    codes = torch.zeros(2, 2, 1, dtype=torch.int64)
    asr.step_tokens(codes, [True, False])
    assert asr.batch[0].step_idx == 1 and asr.batch[1].step_idx == 0
    assert asr.batch[1].text_token == 8           # still the start token (asr.rs:221-223)
