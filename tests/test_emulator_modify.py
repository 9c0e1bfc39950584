"""CPU tests of the PV-domain kernel bodies (flan_b200/csrc/pv_modify_body.cuh, the source the sm_100a kernels of
pv_modify.cu compile) run thread by thread in the host emulator, against the oracle: bit for bit, on the parallel
(monotone) and the sequential (reference-order) forms, for full tables and for the strided views a constant or a
time-only / frequency-only Function samples to."""
import glob
import os

import numpy as np
import pytest

from flan_b200.signals import make_config

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "modify", "modify_*.npz")))
EXACT_INTERPS = [0, 1, 2, 3, 4, 5, 6, 9]       # 7 and 8 go through cos / sin (see include/flan_b200.h)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def emu():
    from emu_lib import EmuModify
    return EmuModify()


@pytest.fixture(scope="module")
def pv_case(oracle):
    x, sr, W, h, N = make_config("cfg5", 0.2)
    pv = oracle.convert_to_pv(x, sr, W, h, N)
    pv = np.concatenate([pv, pv[:, ::-1] * np.float32(0.5)], axis=0)      # two channels
    pv[0, 7, 100:140, 0] = 0                                             # silent cells: the totalWeight == 0 exit
    pv[1, 30:34, :, 0] = 0
    return pv, sr, oracle.analysis_rate(sr, h)


def factor_tables(F, B, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return {
        "full": rng.uniform(lo, hi, (F, B)).astype(np.float32),
        "row": rng.uniform(lo, hi, (B,)).astype(np.float32),
        "column": rng.uniform(lo, hi, (F, 1)).astype(np.float32),
        "constant": np.float32(1.5),
    }


def full(table, F, B):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(table, np.float32).reshape(
        (F, 1) if np.shape(table) == (F, 1) else (1, -1) if np.ndim(table) == 1 else np.shape(table) or (1, 1)), (F, B)))


@pytest.mark.parametrize("interp", EXACT_INTERPS)
def test_repitch_body_matches_oracle(oracle, emu, pv_case, interp):
    pv, sr, ar = pv_case
    _, F, B, _ = pv.shape
    for lo, hi, seed in [(0.3, 2.5, 1), (-1.0, 2.0, 2), (-2.0, -0.1, 3)]:      # ascending, mixed (sequential walk), descending
        for kind, t in factor_tables(F, B, lo, hi, seed).items():
            want = oracle.repitch(pv, sr, full(t, F, B), interp)
            got = emu.repitch(pv, sr, t, interp, threads=64 if interp % 2 else 33)
            assert np.array_equal(bits(got), bits(want)), (kind, lo, hi)


@pytest.mark.parametrize("interp", EXACT_INTERPS)
def test_stretch_body_matches_oracle(oracle, emu, pv_case, interp):
    pv, sr, ar = pv_case
    _, F, B, _ = pv.shape
    for lo, hi, seed in [(0.0, 3.0, 4), (-1.0, 2.5, 5)]:
        for kind, t in factor_tables(F, B, lo, hi, seed).items():
            want = oracle.stretch(pv, sr, ar, full(t, F, B), interp)
            got, seq = emu.stretch(pv, sr, ar, t, interp, chunk=32 if interp % 2 else 7)
            assert got.shape == want.shape and np.array_equal(bits(got), bits(want)), (kind, lo, hi)
            if lo >= 0:
                assert not seq
                got2, _ = emu.stretch(pv, sr, ar, t, interp, force_sequential=1)
                assert np.array_equal(bits(got2), bits(want)), (kind, "sequential")


def test_modify_time_body_matches_oracle(oracle, emu, pv_case):
    pv, sr, ar = pv_case
    _, F, B, _ = pv.shape
    rng = np.random.default_rng(6)
    t = np.arange(F, dtype=np.float32)[:, None] / np.float32(ar)
    maps = {
        "offset": (t * np.float32(1.7) + np.float32(0.05)) * np.ones((1, B), np.float32),      # starts late: head is cleared
        "negative_start": (t * np.float32(0.9) - np.float32(0.02)) * np.ones((1, B), np.float32),
        "per_bin": t * rng.uniform(0.5, 2.0, (1, B)).astype(np.float32),
        "reversed": (t[::-1] * np.float32(1.3)) * np.ones((1, B), np.float32),
        "wobble": t + rng.uniform(-0.01, 0.01, (F, B)).astype(np.float32),
    }
    for kind, m in maps.items():
        m = np.ascontiguousarray(m, np.float32)
        want = oracle.modify_time(pv, sr, ar, m, 0)
        got, seq = emu.modify_time(pv, sr, ar, m, 0)
        assert got.shape == want.shape and np.array_equal(bits(got), bits(want)), kind
        assert seq == (kind in ("reversed", "wobble"))


def test_modify_frequency_body_matches_oracle_scatter(oracle, emu, pv_case):
    # modify_frequency_base with caller-supplied in_mod: reuse repitch's tables (the oracle's repitch is exactly
    # modify_frequency_base( mod = integrated factor in Hz, in_mod = lerp ), PVModify.cpp:304).
    pv, sr, ar = pv_case
    C, F, B, _ = pv.shape
    fac = np.random.default_rng(7).uniform(0.4, 2.0, (F, B)).astype(np.float32)
    hz = np.cumsum(fac, axis=1, dtype=np.float32)            # sequential float32 running sum, as PVModify.cpp:278-280
    for b in range(1, B):
        hz[:, b] = fac[:, b] + hz[:, b - 1]
    hz[:, 0] = fac[:, 0]
    dft = np.float32((B - 1) * 2)
    hz = hz * np.float32(sr) / dft
    bw = np.float32(sr) / dft
    fbin = np.clip(pv[..., 1] / bw, np.float32(0), np.float32(B - 1) - np.float32(0.0001)).astype(np.float32)
    lo = np.floor(fbin).astype(np.int64)
    r = fbin - lo.astype(np.float32)
    fr = np.arange(F)[None, :, None]
    in_mod = (hz[fr, lo] * (np.float32(1) - r) + hz[fr, lo + 1] * r).astype(np.float32)
    want = oracle.repitch(pv, sr, fac, 0)
    got = emu.modify_frequency(pv, sr, hz, in_mod, 0)
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_modify_bodies_reproduce_golden(emu, path):
    g = np.load(path)
    sr, ar, interp = float(g["sr"]), float(g["analysis_rate"]), int(g["interp"])
    assert np.array_equal(bits(emu.repitch(g["pv"], sr, g["repitch_factor"], interp)), bits(g["repitch"]))
    st, _ = emu.stretch(g["pv"], sr, ar, g["stretch_factor"], interp)
    assert st.shape == g["stretch"].shape and np.array_equal(bits(st), bits(g["stretch"]))


def test_sine_interpolators_within_an_ulp(oracle, emu, pv_case):
    pv, sr, ar = pv_case
    _, F, B, _ = pv.shape
    fac = np.full((F, B), 1.7, np.float32)
    for interp in (7, 8):
        want = oracle.stretch(pv, sr, ar, fac, interp)
        got, _ = emu.stretch(pv, sr, ar, np.float32(1.7), interp)
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=3e-6, atol=1e-9)
        assert np.mean(bits(got) == bits(want)) > 0.95


def test_constant_running_sum_closed_form_is_exact():
    # constant_prefix_segments / constant_prefix_value against the plain sequential float32 loop (PVModify.cpp:376-378
    # with a constant factor), including exact-tie increments, denormals, overflow to inf and negative constants.
    import ctypes
    from emu_lib import Emu
    L = Emu().L
    L.pv_emu_constant_prefix_mismatches.restype = ctypes.c_int64
    L.pv_emu_constant_prefix_mismatches.argtypes = [ctypes.c_float, ctypes.c_int64, ctypes.POINTER(ctypes.c_int)]
    rng = np.random.default_rng(0)
    vals = [2.0, 1.5, 1.0, 0.1, 1 / 3, 1e-3, 1e-10, 1e10, 0.0, -1.5, -0.1, 1e-40, 3e-39, float("inf"), 1.0000001, 16777216.0, 1e30, 3e38]
    vals += [1 + 2 ** -23, 3 * 2 ** -24, 5 * 2 ** -25, 1.5 * 2 ** -20, (2 ** 24 - 1) * 2.0 ** -24, 0.5 + 2 ** -24]
    vals += list(rng.uniform(0, 4, 20).astype(np.float32)) + list(np.exp(rng.uniform(-30, 30, 20)).astype(np.float32))
    worst = 0
    for c in vals:
        for F in (1, 2, 3, 100, 300007):
            ns = ctypes.c_int(0)
            assert L.pv_emu_constant_prefix_mismatches(c, F, ctypes.byref(ns)) == 0, (c, F)
            worst = max(worst, ns.value)
    assert worst < 512
