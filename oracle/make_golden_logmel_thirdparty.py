"""Golden log-mel vectors from an independent third-party implementation.  TEST INFRASTRUCTURE.

librosa -- the library /root/reference/feature.py:55-59 calls -- is not installed here and not vendored by the
reference, so the oracle (oracle/logmel_ref.py) cannot be pinned against the reference itself.  The closest thing
available offline is `transformers.audio_utils` (transformers 5.5 in this image), whose `spectrogram`,
`window_function` and `mel_filter_bank` are written to reproduce librosa (`norm="slaney", mel_scale="slaney"`,
centre padding, periodic Hann, power spectrogram, natural log).  This script runs THAT code on the oracle's seeded
synthetic clips and stores inputs and outputs in tests/golden/logmel_thirdparty.npz:

    python oracle/make_golden_logmel_thirdparty.py

tests/test_oracle_logmel.py checks the oracle against these vectors on the CPU; tests/test_logmel_gpu.py checks the
CUDA kernel against them on the GPU.  (mel_floor is set to 1e-37 -- transformers clamps the mel energies from
below, librosa / feature.py:59 do not; none of the clips is silent, so the clamp never acts.)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import logmel_ref as L  # noqa: E402  (synthetic clips only)


def third_party_mbe(y, pad_mode):
    from transformers import audio_utils as A
    mel = A.mel_filter_bank(1025, 40, 0.0, 22050.0, 44100, norm="slaney", mel_scale="slaney")
    win = A.window_function(2048, "hann", periodic=True)
    return A.spectrogram(y, win, 2048, 1024, fft_length=2048, power=2.0, center=True, pad_mode=pad_mode,
                         mel_filters=mel, mel_floor=1e-37, log_mel="log").T.astype(np.float32)


CLIPS = (("mix_1s", 21, 44100, "mix"), ("noise_odd", 22, 2 * 44100 + 1, "noise"), ("chirp_7k", 23, 1024 * 7, "chirp"),
         ("short", 24, 1500, "mix"))


def main():
    import transformers
    out = {"transformers_version": np.array(transformers.__version__)}
    for name, seed, n, kind in CLIPS:
        y = L.synth_clip(seed, n, 1, kind)[0]
        out[name + "_pcm"] = y
        for pm in ("constant", "reflect"):
            out[f"{name}_{pm}"] = third_party_mbe(y, pm)
    path = os.path.join(ROOT, "tests", "golden", "logmel_thirdparty.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
