"""Generates tests/golden/*.npz from the reference's OWN sources (oracle/_ref/libflan_ref.so =
Conversions/AudioPV.cpp, phase_vocoder.cpp, WindowFunctions.cpp, FFTHelper.cpp, PV/PVBuffer.cpp
compiled verbatim, FFT backend 0). Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The reference has no golden vectors of its own (tests/flanTest.cpp is a scratch main); these
fixtures are outputs of the reference itself on seeded inputs and travel to the GPU box.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_lib import RefLib, build_oracle  # noqa: E402
from flan_b200.signals import noise_chirp, sine_sweep  # noqa: E402

CASES = {
    # name: (channels, n, sr, W, hop, N, kind)
    "sweep_w256_h16": (1, 3000, 44100, 256, 16, 256, "sweep"),
    "noise_w512_h32_stereo": (2, 2500, 48000, 512, 32, 512, "noise"),
    "noise_w256_h64_pad1024": (1, 4000, 48000, 256, 64, 1024, "noise"),
    "ragged_w512_h128": (1, 1999, 22050, 512, 128, 512, "noise"),
}


def make_input(C, n, sr, kind):
    if kind == "sweep":
        return np.stack([sine_sweep(n, sr) for _ in range(C)])
    return np.stack([noise_chirp(n, sr, 77 + c) for c in range(C)])


def main():
    build_oracle(ref=True)
    ref = RefLib(0)
    for name, (C, n, sr, W, h, N, kind) in CASES.items():
        x = make_input(C, n, sr, kind)
        pv, ar = ref.convert_to_pv(x, sr, W, h, N)
        audio = ref.convert_to_audio(pv, sr, ar, W)
        extra = {}
        if C == 2:
            pv_ms, _ = ref.convert_to_pv(x, sr, W, h, N, ms=True)
            extra["pv_ms"] = pv_ms
            extra["audio_lr"] = ref.convert_to_audio(pv_ms, sr, ar, W, lr=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), audio_in=x, sr=np.float32(sr), W=W, hop=h, N=N,
                            analysis_rate=ar, pv=pv, audio_out=audio, hann=ref.hann(W), **extra)
        print(name, pv.shape, audio.shape)


if __name__ == "__main__":
    main()
