"""Seeded synthetic signals of the shapes BASELINE.json names.

The reference's own noise synthesis is unseeded (Audio/AudioSynthesis.cpp:80-81), so the
bench, the parity tests and the CPU baseline all draw their inputs from these generators.
"""
import numpy as np


def sine_sweep(n, sr, f0=20.0, f1=None, amp=0.8):
    """Linear sweep f0 -> f1 (default 0.95*sr/2); phase from the exact integral in float64."""
    if f1 is None:
        f1 = 0.95 * sr / 2
    t = np.arange(n, dtype=np.float64) / sr
    dur = n / sr
    phase = 2.0 * np.pi * (f0 * t + (f1 - f0) * t * t / (2.0 * dur))
    return (amp * np.sin(phase)).astype(np.float32)


def noise_chirp(n, sr, seed, noise_amp=0.25, chirp_amp=0.5, f0=50.0, f1=None):
    """0.25*U(-1,1) white noise (Philox, one stream per seed) + 0.5*chirp f0 -> 0.45*sr."""
    if f1 is None:
        f1 = 0.45 * sr
    rng = np.random.Generator(np.random.Philox(seed))
    out = np.empty(n, np.float32)
    dur = n / sr
    step = 1 << 22
    for s in range(0, n, step):
        e = min(n, s + step)
        t = np.arange(s, e, dtype=np.float64) / sr
        phase = 2.0 * np.pi * (f0 * t + (f1 - f0) * t * t / (2.0 * dur))
        u = rng.random(e - s, dtype=np.float32) * 2.0 - 1.0
        out[s:e] = (noise_amp * u + chirp_amp * np.sin(phase)).astype(np.float32)
    return out


def make_config(name, seconds=None):
    """Returns (audio[C][n] float32, sr, W, hop, N) for BASELINE.json configs[...].

    `seconds` shortens the signal (same sample rate / window / hop) for parity tests.
    """
    if name == "cfg1":      # mono 44.1 kHz 10 s sine sweep, W=N=2048 h=128
        sr, W, h, N, C = 44100, 2048, 128, 2048, 1
        n = int(sr * (10 if seconds is None else seconds))
        return np.stack([sine_sweep(n, sr)]), sr, W, h, N
    if name == "cfg2":      # stereo 48 kHz 10 min noise+chirp, W=N=4096 h=256
        sr, W, h, N, C = 48000, 4096, 256, 4096, 2
        n = int(sr * (600 if seconds is None else seconds))
        return np.stack([noise_chirp(n, sr, 1234 + c) for c in range(C)]), sr, W, h, N
    if name == "cfg3":      # mono 96 kHz 1 h, W=N=8192 h=512
        sr, W, h, N, C = 96000, 8192, 512, 8192, 1
        n = int(sr * (3600 if seconds is None else seconds))
        return np.stack([noise_chirp(n, sr, 3)]), sr, W, h, N
    if name == "cfg4":      # 8 ch 48 kHz 30 min, W=N=2048 h=128
        sr, W, h, N, C = 48000, 2048, 128, 2048, 8
        n = int(sr * (1800 if seconds is None else seconds))
        return np.stack([noise_chirp(n, sr, 40 + c) for c in range(C)]), sr, W, h, N
    if name == "cfg5":      # one clip of the 256 x 60 s batch, W=N=1024 h=64
        sr, W, h, N, C = 48000, 1024, 64, 1024, 1
        n = int(sr * (60 if seconds is None else seconds))
        return np.stack([noise_chirp(n, sr, 5000)]), sr, W, h, N
    if name == "apidefault":    # Audio::convert_to_PV() with no arguments (Audio.h:158-163): window 2048, hop 128, dft 4096; one cfg4 channel
        sr, W, h, N, C = 48000, 2048, 128, 4096, 1
        n = int(sr * (600 if seconds is None else seconds))
        return np.stack([noise_chirp(n, sr, 40)]), sr, W, h, N
    raise KeyError(name)
