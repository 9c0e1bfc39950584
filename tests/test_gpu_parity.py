"""GPU parity tests proper: the CUDA path, called through the C ABI (libflan_b200.so), against the oracle
on the same seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's full sizes --
through size-independent properties."""
import glob
import os

import numpy as np
import pytest

from flan_b200.signals import make_config, noise_chirp, sine_sweep
from flan_b200.sharding import frame_shard
from parity import analysis_report, assert_analysis_parity, assert_synthesis_parity

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from flan_b200.engine import Engine
    return Engine(0)


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- oracle parity on small, seeded inputs ------------------------------------------------------------

@pytest.mark.parametrize("name,sec", [("cfg1", 2.0), ("cfg2", 1.0), ("cfg3", 1.0), ("cfg4", 0.5), ("cfg5", 1.0)])
def test_convert_to_pv_matches_oracle(eng, oracle, name, sec):
    x, sr, W, h, N = make_config(name, sec)
    pv = eng.convert_to_pv(dev(x), sr, W, h, N).cpu().numpy()
    ref = oracle.convert_to_pv(x, sr, W, h, N)
    assert pv.shape == ref.shape
    assert_analysis_parity(pv, ref, sr, h, N)


@pytest.mark.parametrize("name,sec", [("cfg1", 2.0), ("cfg2", 1.0), ("cfg3", 1.0), ("cfg4", 0.5), ("cfg5", 1.0)])
def test_convert_to_audio_matches_oracle(eng, oracle, name, sec):
    x, sr, W, h, N = make_config(name, sec)
    ref_pv = oracle.convert_to_pv(x, sr, W, h, N)
    ar = oracle.analysis_rate(sr, h)
    out, flag = eng.convert_to_audio(dev(ref_pv), sr, float(ar), W, check_nan=True)
    assert not flag
    ref = oracle.convert_to_audio(ref_pv, sr, ar, W)
    assert_synthesis_parity(out.cpu().numpy(), ref)          # <= 1e-5 max abs, stage-wise on the same PV


@pytest.mark.parametrize("W,h,N", [(256, 16, 256), (512, 32, 512), (256, 64, 1024), (1000, 100, 1024), (2048, 128, 4096),
                                   (512, 512, 512), (300, 7, 512), (2048, 2048, 2048), (8192, 64, 8192),
                                   (256, 400, 256)])        # hop > window: gaps between frames stay zero (full output clear)
def test_odd_shapes_match_oracle(eng, oracle, W, h, N):
    sr = 32000.0
    n = 9001
    x = np.stack([noise_chirp(n, sr, 31), sine_sweep(n, sr), np.zeros(n, np.float32)])
    ref_pv = oracle.convert_to_pv(x, sr, W, h, N)
    pv = eng.convert_to_pv(dev(x), sr, W, h, N).cpu().numpy()
    assert_analysis_parity(pv[:2], ref_pv[:2], sr, h, N)
    # all-zero channel: deterministic known answer, must be bit-identical (m = 0, f from the wrap of -expected)
    assert np.array_equal(pv[2].view(np.uint32), ref_pv[2].view(np.uint32))
    ar = oracle.analysis_rate(sr, h)
    out = eng.convert_to_audio(dev(ref_pv), sr, float(ar), W).cpu().numpy()
    assert_synthesis_parity(out, oracle.convert_to_audio(ref_pv, sr, ar, W))


# Every dft size the reference accepts (FFTW plans any size, FFTHelper.cpp:16-26): the run-time-sized transform of
# pv_generic.cu serves what the templated kernels do not -- VERDICT r1 item 6 asked for (3000, 100, 3000),
# (1536, 96, 1536) and (16384, 1024, 16384); plus small powers of two, an odd size (bins and the inverse transform then
# follow get_dft_size() = (num_bins - 1) * 2, PVBuffer.cpp:356-359), a size whose buffers live in global memory, tiny sizes.
@pytest.mark.parametrize("W,h,N", [(3000, 100, 3000), (1536, 96, 1536), (16384, 1024, 16384), (128, 8, 128), (64, 16, 64),
                                   (100, 10, 128), (2048, 128, 32768), (1001, 91, 1001), (250, 25, 375), (5000, 500, 12000),
                                   (6, 2, 6), (2048, 128, 65536)])
def test_any_dft_size_matches_oracle(eng, oracle, W, h, N):
    sr = 32000.0
    n = 40001 if N >= 16384 else (9001 if N >= 64 else 200)
    pow2 = (N & (N - 1)) == 0
    chans = [noise_chirp(n, sr, 31), sine_sweep(n, sr)] + ([np.zeros(n, np.float32)] if N % 2 == 0 and (pow2 or N <= 4096) else [])
    x = np.stack(chans)
    # the oracle's non-power-of-two transform is the O(n^2) definition: bound its work to a window of frames
    F = n // h + 1
    half_window = 20 if N <= 4096 else 2      # the oracle's any-size DFT is O(n^2): seconds per frame at 12000 points
    f0, f1 = (0, F) if (N & (N - 1)) == 0 or N <= 512 else (max(0, F // 2 - half_window), min(F, F // 2 + half_window))
    ref_pv = oracle.convert_to_pv(x, sr, W, h, N, f0, f1)
    pv = eng.convert_to_pv(dev(x), sr, W, h, N).cpu().numpy()
    assert pv.shape == (len(chans), F, N // 2 + 1, 2)
    if N >= 64:
        # the gate needs the PREVIOUS frame's level too (its phase enters the difference): the first frame of a window
        # that does not start at frame 0 has none to look at and is left out
        assert_analysis_parity(pv[:2, f0:f1], ref_pv[:2], sr, h, N, first_frame_has_history=f0 > 0)
    else:
        assert np.allclose(pv[:2, f0:f1, :, 0], ref_pv[:2, :, :, 0], rtol=1e-4, atol=1e-5)
    if len(chans) == 3:      # all-zero channel: deterministic known answer, bit-identical
        assert np.array_equal(pv[2, f0:f1].view(np.uint32), ref_pv[2].view(np.uint32))
    if W > (N // 2) * 2:
        return
    # resynthesis of (a window of) the oracle's frames, stage-wise on the same PV
    ar = oracle.analysis_rate(sr, h)
    out = eng.convert_to_audio(dev(ref_pv), sr, float(ar), W).cpu().numpy()
    assert_synthesis_parity(out, oracle.convert_to_audio(ref_pv, sr, ar, W))


def test_ragged_and_tiny_inputs(eng, oracle):
    sr, W, h, N = 44100.0, 512, 128, 512
    for n in (0, 1, 127, 128, 129, 511, 1000):
        x = np.stack([noise_chirp(max(n, 1), sr, 2)[:n]]) if n else np.zeros((1, 0), np.float32)
        pv = eng.convert_to_pv(dev(x), sr, W, h, N).cpu().numpy()
        ref = oracle.convert_to_pv(x, sr, W, h, N)
        assert pv.shape == ref.shape == (1, n // h + 1, N // 2 + 1, 2)
        if n:
            assert_analysis_parity(pv, ref, sr, h, N)
        ar = oracle.analysis_rate(sr, h)
        out = eng.convert_to_audio(dev(ref), sr, float(ar), W).cpu().numpy()
        assert_synthesis_parity(out, oracle.convert_to_audio(ref, sr, ar, W))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_fixtures(eng, path):
    # outputs of the reference's own sources (tests/golden/make_golden.py)
    g = np.load(path)
    sr, W, h, N = float(g["sr"]), int(g["W"]), int(g["hop"]), int(g["N"])
    pv = eng.convert_to_pv(dev(g["audio_in"]), sr, W, h, N).cpu().numpy()
    assert_analysis_parity(pv, g["pv"], sr, h, N)
    out = eng.convert_to_audio(dev(g["pv"]), sr, float(g["analysis_rate"]), W).cpu().numpy()
    assert_synthesis_parity(out, g["audio_out"])
    if "pv_ms" in g:
        pv_ms = eng.convert_to_pv_host(g["audio_in"], sr, W, h, N, mid_side=True)
        assert_analysis_parity(pv_ms, g["pv_ms"], sr, h, N)
        lr, _ = eng.convert_to_audio_host(g["pv_ms"], sr, float(g["analysis_rate"]), W, left_right=True)
        assert_synthesis_parity(lr, g["audio_lr"])


def test_mid_side_bit_exact(eng, oracle):
    x = np.stack([noise_chirp(10001, 48000, 1), noise_chirp(10001, 48000, 2)])
    ms = eng.mid_side(dev(x)).cpu().numpy()
    assert np.array_equal(ms.view(np.uint32), oracle.mid_side(x).view(np.uint32))


def test_host_forms_equal_device_forms(eng):
    x, sr, W, h, N = make_config("cfg5", 0.5)
    a = eng.convert_to_pv(dev(x), sr, W, h, N).cpu().numpy()
    b = eng.convert_to_pv_host(x, sr, W, h, N)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ar = eng.analysis_rate(sr, h)
    c = eng.convert_to_audio(dev(a), sr, ar, W).cpu().numpy()
    d, flag = eng.convert_to_audio_host(a, sr, ar, W)
    assert not flag
    assert np.array_equal(c.view(np.uint32), d.view(np.uint32))


def test_error_behaviour(eng):
    import torch
    from flan_b200.capi import FlanB200Error, UNSUPPORTED, INVALID
    x = torch.zeros((1, 4096), device="cuda")
    with pytest.raises(FlanB200Error) as e:
        eng.convert_to_pv(x, 48000.0, 1, 1, 1)               # the smallest dft size is 2
    assert e.value.code == UNSUPPORTED
    with pytest.raises(FlanB200Error) as e:
        eng.convert_to_pv(x, 48000.0, 1024, 64, (1 << 20) + 2)   # ... the largest 2^20
    assert e.value.code == UNSUPPORTED
    with pytest.raises(FlanB200Error) as e:
        eng.convert_to_pv(x, 48000.0, 1024, 64, 512)         # window larger than dft
    assert e.value.code == INVALID
    with pytest.raises(FlanB200Error):
        eng.convert_to_pv_host(np.zeros((3, 100), np.float32), 48000.0, 256, 32, 256, mid_side=True)   # AudioPV.cpp:82
    pv = torch.zeros((1, 8, 129, 2), device="cuda")
    pv[0, 3, 5, 0] = float("nan")
    _, flag = eng.convert_to_audio(pv, 48000.0, 3000.0, 256, check_nan=True)       # AudioPV.cpp:88: warn, continue
    assert flag


def test_negative_frequency_phase_representative(eng, oracle):
    sr, W, h, N = 48000.0, 256, 16, 256
    F, B = 3000, N // 2 + 1
    rng = np.random.default_rng(3)
    pv = np.zeros((1, F, B, 2), np.float32)
    pv[..., 0] = rng.random((1, F, B), dtype=np.float32)
    binf = np.arange(B) * sr / N
    pv[..., 1] = (binf[None, None, :] + rng.normal(0, 900, (1, F, B))).astype(np.float32)
    pv[0, :, 0:3, 1] = -np.abs(pv[0, :, 0:3, 1]) - 500.0
    pv[0, 100:, 5, 1] = -2000.0
    ar = oracle.analysis_rate(sr, h)
    out = eng.convert_to_audio(dev(pv), sr, float(ar), W).cpu().numpy()
    assert_synthesis_parity(out, oracle.convert_to_audio(pv, sr, ar, W))


# ---- frame-range shards on one GPU (the multi-GPU path, ranks emulated serially) --------------------------

@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name", ["cfg3", "cfg2", "cfg5"])      # 8-point one-buffer kernel; mirrored kernels (dft 4096, 1024)
def test_frame_range_shards_reproduce_unsharded(eng, world, name):
    import torch
    x, sr, W, h, N = make_config(name, 2.0)
    n = x.shape[1]
    xd = dev(x)
    full_pv = eng.convert_to_pv(xd, sr, W, h, N)
    ar = eng.analysis_rate(sr, h)
    full_audio = eng.convert_to_audio(full_pv, sr, ar, W)
    shards = [frame_shard(n, h, W, world, r) for r in range(world)]
    states, pvs = [], []
    for s in shards:
        local = xd[:, s.audio_lo:s.audio_hi].contiguous()
        pv = eng.convert_to_pv_range(local, s.audio_lo, n, sr, W, h, N, s.f0, s.f1)
        assert torch.equal(pv, full_pv[:, s.f0:s.f1])
        pvs.append(pv)
        states.append(eng.phase_summary(pv, s.f0, sr, ar, W))
    all_states = torch.stack(states)
    total = torch.zeros_like(full_audio)
    for s, pv in zip(shards, pvs):
        carry = eng.phase_carry(all_states, s.rank)
        out = eng.convert_to_audio_range(pv, s.f0, s.frames_total, sr, ar, W, carry, s.span_lo, s.span_hi - s.span_lo)
        total[:, s.span_lo:s.span_hi] += out
    err = (total - full_audio).abs().max().item()
    assert err <= 2e-6, err


# ---- kernel variants agree: the mirrored 16-point kernels and their fall-backs against the 8-point kernels ---------

@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg5"])
def test_kernel_variants_agree(eng, name, monkeypatch):
    import torch
    from flan_b200.engine import Engine
    x, sr, W, h, N = make_config(name, 2.0)
    x = np.concatenate([x, x[:1] * 0.5], axis=0)            # an odd number of rows before the last channel: both row alignments
    xd = dev(x)
    pv = eng.convert_to_pv(xd, sr, W, h, N)
    ar = eng.analysis_rate(sr, h)
    y = eng.convert_to_audio(pv, sr, ar, W)
    # PV buffer that is only 8-byte aligned: bulk row copies are not possible, every thread stages its own bins
    flat = torch.empty(pv.numel() + 2, device="cuda")
    v = flat[2:].view_as(pv)
    v.copy_(pv)
    assert v.data_ptr() % 16 == 8
    assert torch.equal(eng.convert_to_audio(v, sr, ar, W), y)
    # the 8-point kernels: the launch policy is fixed in the release library; the FLAN_B200_DEBUG development build
    # (tools/experiments) reads overrides from the environment when a context is created
    from flan_b200 import build
    monkeypatch.setenv("FLAN_B200_SYNTH_VARIANT", "8")
    monkeypatch.setenv("FLAN_B200_PT_ANALYSIS", "8")
    eng8 = Engine(0, lib_path=build.build_library(debug=True))
    y8 = eng8.convert_to_audio(pv, sr, ar, W)
    assert (y8 - y).abs().max().item() <= 1e-6
    pv8 = eng8.convert_to_pv(xd, sr, W, h, N)
    rep = assert_analysis_parity(pv.cpu().numpy(), pv8.cpu().numpy(), sr, h, N)
    assert rep["max_rel_m"] <= 1e-4


# ---- BASELINE.json full sizes: size-independent properties ------------------------------------------------

def _full_size_checks(eng, x_dev, sr, W, h, N, oracle, probe_frames=64):
    """Full-size run; parity on windows of frames re-derived by the oracle from the same samples; linearity;
    round-trip gain."""
    import torch
    C, n = x_dev.shape
    F = eng.num_frames(n, h)
    pv = eng.convert_to_pv(x_dev, sr, W, h, N)
    assert pv.shape == (C, F, N // 2 + 1, 2)
    assert not torch.isnan(pv).any()
    # (1) oracle parity on frame windows scattered through the signal (the oracle recomputes the carried phase)
    for f0 in sorted({0, F // 3, F // 2 + 7, F - probe_frames}):
        f0 = max(0, min(f0, F - probe_frames))
        f1 = f0 + probe_frames
        lo = max(0, h * (f0 - 1) - W // 2)
        hi = min(n, h * (f1 - 1) + W // 2)
        # oracle on a cropped copy: shift so the absolute frame grid is preserved
        assert lo % h == 0
        crop = x_dev[:, lo:hi].cpu().numpy()
        ref = oracle.convert_to_pv(crop, sr, W, h, N, (f0 - lo // h), (f0 - lo // h) + probe_frames)
        got = pv[:, f0:f1].cpu().numpy()
        if lo > 0:
            # frames whose window touches the crop's artificial left edge are not comparable
            ref, got = ref[:, 1:], got[:, 1:]
        if hi < n:
            ref, got = ref[:, :-1], got[:, :-1]
        assert_analysis_parity(got, ref, sr, h, N)
    # (2) homogeneity: scaling the input by 2 scales m exactly and leaves f bit-identical
    pv2 = eng.convert_to_pv(x_dev * 2.0, sr, W, h, N)
    assert torch.equal(pv2[..., 1], pv[..., 1])
    assert torch.equal(pv2[..., 0], pv[..., 0] * 2.0)
    del pv2
    # (3) round trip: length F*h, RMS gain of the interior ~1.001 (AudioPV.cpp:99's 2.67 constant)
    ar = eng.analysis_rate(sr, h)
    y = eng.convert_to_audio(pv, sr, ar, W)
    assert y.shape == (C, F * h)
    assert not torch.isnan(y).any()
    return pv, y


def test_full_size_cfg1(eng, oracle):
    import torch
    x, sr, W, h, N = make_config("cfg1")
    pv, y = _full_size_checks(eng, dev(x), sr, W, h, N, oracle)
    ref_pv = oracle.convert_to_pv(x, sr, W, h, N)                      # cfg1 runs on the CPU in seconds: full parity
    assert_analysis_parity(pv.cpu().numpy(), ref_pv, sr, h, N)
    ar = oracle.analysis_rate(sr, h)
    y_ref = oracle.convert_to_audio(ref_pv, sr, ar, W)
    y_stage = eng.convert_to_audio(dev(ref_pv), sr, float(ar), W).cpu().numpy()
    assert_synthesis_parity(y_stage, y_ref)
    # end-to-end chain (engine analysis -> engine synthesis vs oracle chain): reported, gated loosely
    e2e = np.abs(y.cpu().numpy() - y_ref).max()
    print("cfg1 end-to-end max abs diff vs oracle chain:", e2e)
    assert e2e < 2e-2
    sl = slice(4096, x.shape[1] - 4096)
    g = float(np.sqrt(np.mean(y.cpu().numpy()[0, sl] ** 2) / np.mean(x[0, sl] ** 2)))
    assert 0.995 < g < 1.006


def test_full_size_cfg2(eng, oracle):
    import torch
    x, sr, W, h, N = make_config("cfg2")            # stereo 48 kHz 10 min: 112 501 frames x 2049 bins x 2 ch
    xd = dev(x)
    pv, y = _full_size_checks(eng, xd, sr, W, h, N, oracle)
    n = x.shape[1]
    sl = slice(8192, n - 8192)
    for c in range(2):
        g = float(torch.sqrt((y[c, sl] ** 2).mean() / (xd[c, sl] ** 2).mean()))
        assert 0.9 < g < 1.1, g


def test_full_size_cfg5_clip(eng, oracle):
    x, sr, W, h, N = make_config("cfg5")            # one 60 s clip of the 256-clip batch
    _full_size_checks(eng, dev(x), sr, W, h, N, oracle)


def test_full_size_cfg3_ten_minutes(eng, oracle):
    x, sr, W, h, N = make_config("cfg3", 600)       # 96 kHz, W=8192: 10 of the 60 minutes (1 h = 22 GB PV, see bench)
    _full_size_checks(eng, dev(x), sr, W, h, N, oracle)


def test_summary_reuse_is_bit_identical(eng):
    import torch
    x, sr, W, h, N = make_config("cfg5", 1.5)
    pv = eng.convert_to_pv(dev(x), sr, W, h, N)
    ar = eng.analysis_rate(sr, h)
    F = pv.shape[1]
    plain = eng.convert_to_audio_range(pv, 0, F, sr, ar, W, None, 0, F * h)
    eng.phase_summary(pv, 0, sr, ar, W)
    reused = eng.convert_to_audio_range(pv, 0, F, sr, ar, W, None, 0, F * h, reuse_summary=True)
    assert torch.equal(plain, reused)
    # a promise that does not match the preceding call is ignored, not trusted
    eng.phase_summary(pv[:, :100].contiguous(), 0, sr, ar, W)
    again = eng.convert_to_audio_range(pv, 0, F, sr, ar, W, None, 0, F * h, reuse_summary=True)
    assert torch.equal(plain, again)
    assert torch.equal(plain, eng.convert_to_audio(pv, sr, ar, W))


# ---- the parity gate, quantified ---------------------------------------------------
# Floors on the tight statistics per BASELINE shape (measured on B200 minus one percentage point): the fraction of gated
# bins whose frequency is within ONE float32 ulp of the oracle's and the fraction that is bit-identical. A regression in
# the epilogue arithmetic shows here long before it reaches the tolerance gate. The report -- including how many bins
# of SURVEY 8c's unmodified gate pass only through each of the three allowances of tests/parity.py -- is written to
# gpurun_out/parity_report.json (a copy is kept under profiles/).
PARITY_FLOORS = {"cfg1": (0.985, 0.97), "cfg2": (0.95, 0.90), "cfg3": (0.95, 0.90), "cfg5": (0.95, 0.90)}


@pytest.mark.parametrize("name,sec", [("cfg1", 10.0), ("cfg2", 4.0), ("cfg3", 6.0), ("cfg5", 4.0)])
def test_parity_gate_statistics(eng, oracle, name, sec):
    import json
    import os
    x, sr, W, h, N = make_config(name, sec)
    ref = oracle.convert_to_pv(x, sr, W, h, N)
    pv = eng.convert_to_pv(dev(x), sr, W, h, N).cpu().numpy()
    r = assert_analysis_parity(pv, ref, sr, h, N)
    within, exact = PARITY_FLOORS[name]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        allr = json.load(open(path)) if os.path.exists(path) else {}
        allr[name] = dict(r, seconds=sec, floors={"frac_f_within_1ulp_gated": within, "frac_f_bit_exact_gated": exact})
        json.dump(allr, open(path, "w"), indent=1)
    except OSError:
        pass
    print(name, r)
    assert r["frac_f_within_1ulp_gated"] >= within, r
    assert r["frac_f_bit_exact_gated"] >= exact, r      # (magnitudes: |z| is max * sqrt(1 + r^2), a few ulp from hypotf: tolerance only)
