"""Fuzz: the launch policy's kernels against the 8-point kernels on random short / ragged inputs (development aid).
    python tools/fuzz_variants.py [seconds=60]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flan_b200.engine import Engine  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
eng = Engine(0)
os.environ["FLAN_B200_SYNTH_VARIANT"] = "8"
os.environ["FLAN_B200_PT_ANALYSIS"] = "8"
eng8 = Engine(0)
rng = np.random.default_rng(12345)
t0, cases, worst_a, worst_s = time.time(), 0, 0.0, 0.0
while time.time() - t0 < budget:
    N = int(rng.choice([1024, 2048, 4096, 8192]))
    W, h = N, N // 16
    C = int(rng.integers(1, 4))
    n = int(rng.choice([1, 2, h - 1, h, h + 1, W // 2, W - 1, W, W + 1, 3 * W + 7, int(rng.integers(1, 40 * W))]))
    sr = float(rng.choice([44100.0, 48000.0, 96000.0]))
    x = torch.from_numpy((rng.standard_normal((C, n)) * 0.3).astype(np.float32)).cuda()
    pv = eng.convert_to_pv(x, sr, W, h, N)
    pv8 = eng8.convert_to_pv(x, sr, W, h, N)
    assert pv.shape == pv8.shape and torch.isfinite(pv).all(), (N, C, n)
    m, m8 = pv[..., 0], pv8[..., 0]
    rel = ((m - m8).abs() / m8.clamp_min(1e-2 * float(m8.max()) + 1e-30)).max().item()
    worst_a = max(worst_a, rel)
    assert rel < 1e-4, (N, C, n, rel)
    ar = eng.analysis_rate(sr, h)
    y = eng.convert_to_audio(pv8, sr, ar, W)
    y8 = eng8.convert_to_audio(pv8, sr, ar, W)
    d = (y - y8).abs().max().item()
    worst_s = max(worst_s, d)
    assert torch.isfinite(y).all() and d < 2e-6, (N, C, n, d)
    cases += 1
print("fuzz ok: %d cases, worst rel magnitude diff %.2e, worst sample diff %.2e" % (cases, worst_a, worst_s))
