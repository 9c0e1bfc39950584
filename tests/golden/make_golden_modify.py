"""Generates tests/golden/modify_*.npz from the reference's OWN PV/PVModify.cpp (oracle/_ref/libflan_ref_modify.so,
compiled verbatim with its real Function / FunctionSample2d / Interpolator headers). Run in the build container:

    python tests/golden/make_golden_modify.py

Inputs are PV buffers from the golden analysis fixtures next to this file; factor tables are seeded.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_lib import RefModifyLib, build_oracle  # noqa: E402

# name: (source fixture, frames kept, factor kind, interpolator id)
CASES = {
    "modify_const_linear": ("noise_w512_h32_stereo", 40, "const", 0),
    "modify_table_smoothstep": ("noise_w256_h64_pad1024", 48, "table", 5),
    "modify_signed_nearest": ("sweep_w256_h16", 60, "signed", 2),
}


def factor_table(kind, F, B, seed):
    rng = np.random.default_rng(seed)
    if kind == "const":
        return np.full((F, B), 1.5, np.float32), np.full((F, B), 2.0, np.float32)
    if kind == "table":
        return rng.uniform(0.5, 2.0, (F, B)).astype(np.float32), rng.uniform(0.25, 3.0, (F, B)).astype(np.float32)
    return rng.uniform(-1.0, 2.0, (F, B)).astype(np.float32), rng.uniform(-1.0, 2.5, (F, B)).astype(np.float32)


def main():
    build_oracle(ref=True)
    ref = RefModifyLib()
    for i, (name, (src, keep, kind, interp)) in enumerate(CASES.items()):
        g = np.load(os.path.join(HERE, src + ".npz"))
        pv = np.ascontiguousarray(g["pv"][:, :keep])
        sr, ar, W = float(g["sr"]), float(g["analysis_rate"]), int(g["W"])
        C, F, B, _ = pv.shape
        fr, fs = factor_table(kind, F, B, 900 + i)
        rep = ref.repitch(pv, sr, ar, W, fr, interp)
        stz = ref.stretch(pv, sr, ar, W, fs, interp)
        np.savez_compressed(os.path.join(HERE, "modify", name + ".npz"), pv=pv, sr=np.float32(sr), analysis_rate=np.float32(ar), W=W,
                            interp=interp, repitch_factor=fr, stretch_factor=fs, repitch=rep, stretch=stz)
        print(name, pv.shape, rep.shape, stz.shape)


if __name__ == "__main__":
    main()
