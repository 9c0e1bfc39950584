"""Kernel-level timing of BASELINE config 4's chain on one GPU's share (one channel of the 8-channel signal):
analysis -> PV::repitch( 1.5 ) -> PV::stretch( 2.0 ) -> resynthesis, all device-resident (development aid).

    python tools/chainbench.py [seconds=600] [channels=1]
Prints ms per stage, HBM GB/s against the algorithmic bytes of each stage and the chain's frames/s."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flan_b200.engine import Engine  # noqa: E402
from flan_b200.signals import noise_chirp  # noqa: E402


PLAIN = "--plain" in sys.argv      # the round-1 form: phase summaries from a second read of the stretched rows
if PLAIN:
    sys.argv.remove("--plain")


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    sr, W, h, N = 48000.0, 2048, 128, 2048
    n = int(sr * seconds)
    x = np.stack([noise_chirp(n, sr, 40 + c) for c in range(C)])
    eng = Engine(0)
    xd = torch.from_numpy(x).cuda()
    F, B = eng.num_frames(n, h), N // 2 + 1
    ar = eng.analysis_rate(sr, h)
    pv = torch.empty((C, F, B, 2), device="cuda")
    rp = torch.empty_like(pv)

    def chain():
        eng.convert_to_pv(xd, sr, W, h, N, out=pv)
        eng.repitch(pv, sr, 1.5, 0, out=rp)
        # the stretch kernel leaves the phase summaries of its rows; resynthesis does not read the rows a second time
        st = eng.stretch(rp, sr, ar, 2.0, 0, summary_window=0 if PLAIN else W)
        y = eng.convert_to_audio(st, sr, ar, W, unchanged=not PLAIN)
        return st, y

    for _ in range(2):
        st, y = chain()
        del st, y
    eng.set_timing(True)
    for k in eng.KERNEL_KINDS:
        eng.kernel_time(k)
    reps = 3
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        st, y = chain()
        F2 = st.shape[1]
        del st, y
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / reps
    res = {k: eng.kernel_time(k)[0] / reps for k in eng.KERNEL_KINDS}
    peak = 6553.9
    row = 8.0 * B
    stage_bytes = {
        "analysis": C * (4.0 * n + row * F),
        "repitch": C * (row * F + row * F),
        "stretch": C * (row * F + row * F2),
        "synthesis": C * (row * F2 + 4.0 * F2 * h),
    }
    out = {"seconds": seconds, "channels": C, "frames_in": C * F, "frames_out": C * F2, "chain_ms": round(wall, 3)}
    for k, v in res.items():
        out[k + "_ms"] = round(v, 4)
        if k in stage_bytes and v > 0:
            out[k + "_GBs"] = round(stage_bytes[k] / v / 1e6, 1)
            out[k + "_frac"] = round(stage_bytes[k] / v / 1e6 / peak, 4)
    out["chain_Mframes_in_per_s"] = round(C * F / wall / 1e3, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
