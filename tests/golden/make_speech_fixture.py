"""Converts the reference's bundled speech sample into a small committed fixture.

    python tests/golden/make_speech_fixture.py

/root/reference/audio/tts-input-sample-01.wav (48 kHz stereo int16, 6.33 s) -> mono, polyphase
resample to 24 kHz, int16 -> tests/golden/speech_24k_int16.npy.  (audio/bria.mp3, the file BASELINE
config 1 names, is MPEG-1 Layer III and no decoder is available offline; this WAV is the only bundled
audio that can be read here.)  The GPU box has no /root/reference, hence the committed copy.
"""
import os

import numpy as np
from scipy.io import wavfile
from scipy.signal import resample_poly

SRC = "/root/reference/audio/tts-input-sample-01.wav"
HERE = os.path.dirname(os.path.abspath(__file__))

sr, x = wavfile.read(SRC)
assert sr == 48000 and x.ndim == 2
mono = x.astype(np.float64).mean(axis=1) / 32768.0
y = resample_poly(mono, 1, 2)
y = np.clip(np.round(y * 32767.0), -32768, 32767).astype(np.int16)
np.save(os.path.join(HERE, "speech_24k_int16.npy"), y)
print("wrote", y.shape, "samples =", y.shape[0] / 24000.0, "s")
