"""Kernel-level timing of the two transforms on one BASELINE shape (development aid; bench.py is the judged line).

    python tools/kbench.py [cfg2|cfg1|cfg3|cfg5|cfg4|apidefault] [seconds]
Prints ms per launch of each kernel kind and the HBM fraction, for the current FLAN_B200_TPS_* environment."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flan_b200.engine import Engine  # noqa: E402
from flan_b200.signals import make_config  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    name = args[0] if args else "cfg2"
    seconds = float(args[1]) if len(args) > 1 else {"cfg2": 600, "cfg1": 10, "cfg3": 600, "cfg5": 60, "cfg4": 120, "apidefault": 600}[name]
    x, sr, W, h, N = make_config(name, seconds)
    if name == "cfg5":
        x = np.repeat(x, 32, axis=0)          # 32 clips per GPU
    eng = Engine(0)
    xd = torch.from_numpy(x).cuda()
    C, n = x.shape
    F, B = eng.num_frames(n, h), N // 2 + 1
    ar = eng.analysis_rate(sr, h)
    pv = torch.empty((C, F, B, 2), device="cuda")
    y = torch.empty((C, F * h), device="cuda")
    hint = "--hint" in sys.argv          # analysis leaves the phase summaries (flan_b200_hint_resynthesis)
    for _ in range(3):
        eng.convert_to_pv(xd, sr, W, h, N, out=pv, for_resynthesis=hint)
        eng.convert_to_audio(pv, sr, ar, W, out=y, unchanged=hint)
    eng.set_timing(True)
    for k in eng.KERNEL_KINDS:
        eng.kernel_time(k)
    reps = 5
    for _ in range(reps):
        eng.convert_to_pv(xd, sr, W, h, N, out=pv, for_resynthesis=hint)
        eng.convert_to_audio(pv, sr, ar, W, out=y, unchanged=hint)
    res = {k: eng.kernel_time(k) for k in eng.KERNEL_KINDS}
    peak = 6553.9
    an = res["analysis"][0] / reps
    sy = res["synthesis"][0] / reps
    seg = res["phase_seg"][0] / reps
    scan = res["phase_scan"][0] / reps
    byts = 4.0 * C * n + 8.0 * C * F * B
    out = {"cfg": name + (" +hint" if hint else ""), "N": N, "W": W, "hop": h, "frames": C * F,
           "analysis_ms": round(an, 4), "analysis_frac": round(byts / an / 1e6 / peak, 4),
           "synthesis_ms": round(sy, 4), "synthesis_frac": round(byts / sy / 1e6 / peak, 4),
           "phase_seg_ms": round(seg, 4), "phase_scan_ms": round(scan, 4),
           "round_trip_Mframes_s": round(C * F / (an + sy + seg + scan) / 1e3, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
