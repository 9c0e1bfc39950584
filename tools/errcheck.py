"""Prints the measured parity errors of the CUDA path against the oracle on short versions of the BASELINE shapes
(development aid; the gated assertions live in tests/test_gpu_parity.py)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from flan_b200.engine import Engine  # noqa: E402
from flan_b200.signals import make_config  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from parity import analysis_report  # noqa: E402

eng, oracle = Engine(0), Oracle()
for name, sec in [("cfg1", 3.0), ("cfg2", 2.0), ("cfg3", 2.0), ("cfg5", 2.0)]:
    x, sr, W, h, N = make_config(name, sec)
    ref_pv = oracle.convert_to_pv(x, sr, W, h, N)
    pv = eng.convert_to_pv(torch.from_numpy(x).cuda(), sr, W, h, N).cpu().numpy()
    rep = analysis_report(pv, ref_pv, sr, h, N)
    ar = oracle.analysis_rate(sr, h)
    ref_y = oracle.convert_to_audio(ref_pv, sr, ar, W)
    y = eng.convert_to_audio(torch.from_numpy(ref_pv).cuda(), sr, float(ar), W).cpu().numpy()
    print(json.dumps({"cfg": name, "max_rel_m": rep["max_rel_m"], "max_df_hz": rep["max_df_hz"], "bad_m": rep["bad_m"],
                      "bad_f": rep["bad_f"], "f_bit_exact": round(rep["frac_f_bit_exact"], 4),
                      "synth_max_abs": float(np.abs(y - ref_y).max()), "synth_rms": float(np.sqrt(np.mean((y - ref_y) ** 2)))}))
