"""The C++ side of the drop-in boundary: user code written against flan::Audio / flan::PV (tests/cpp/
flan_api_driver.cpp, the shape of the reference's tests/flanTest.cpp:39-44) compiled against the B200 build's
headers and run through libflan_b200_host.so -> C ABI -> CUDA kernels."""
import ctypes

import numpy as np
import pytest

from flan_b200.signals import noise_chirp, sine_sweep
from parity import assert_analysis_parity, assert_synthesis_parity

_fp = ctypes.POINTER(ctypes.c_float)


def _ptr(a):
    return a.ctypes.data_as(_fp)


@pytest.fixture(scope="module")
def api():
    from flan_b200 import build
    build.build_host()
    L = ctypes.CDLL(build.api_test_path())
    L.api_convert_to_pv.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, _fp, _fp]
    L.api_round_trip.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, _fp]
    L.api_convert_to_audio.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                       ctypes.c_int, _fp]
    return L


def test_host_layer_loads_and_cancellation_returns_null(api):
    # no GPU needed: a raised canceller returns a null PV before any device work (AudioPV.cpp:49)
    assert api.api_cancelled_is_null() == 1


def test_without_gpu_the_api_returns_null_objects(api):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    x = np.zeros((1, 4096), np.float32)
    pv = np.zeros((1, 129, 129, 2), np.float32)
    ar = ctypes.c_float()
    assert api.api_convert_to_pv(_ptr(x), 1, 4096, 48000.0, 256, 32, 256, 0, _ptr(pv), ctypes.byref(ar)) == -1


@pytest.mark.gpu
def test_cpp_api_matches_oracle(api, oracle):
    sr, W, h, N = 44100.0, 2048, 128, 2048
    n = 30000
    x = np.stack([noise_chirp(n, sr, 1), sine_sweep(n, sr)])
    F, B = n // h + 1, N // 2 + 1
    pv = np.zeros((2, F, B, 2), np.float32)
    ar = ctypes.c_float()
    assert api.api_convert_to_pv(_ptr(x), 2, n, sr, W, h, N, 0, _ptr(pv), ctypes.byref(ar)) == F
    ref = oracle.convert_to_pv(x, sr, W, h, N)
    assert ar.value == oracle.analysis_rate(sr, h)
    assert_analysis_parity(pv, ref, sr, h, N)

    # mid/side wrapper
    pv_ms = np.zeros_like(pv)
    assert api.api_convert_to_pv(_ptr(x), 2, n, sr, W, h, N, 1, _ptr(pv_ms), ctypes.byref(ar)) == F
    assert_analysis_parity(pv_ms, oracle.convert_to_pv(oracle.mid_side(x), sr, W, h, N), sr, h, N)
    # ... and null for non-stereo input (AudioPV.cpp:82)
    assert api.api_convert_to_pv(_ptr(x[:1].copy()), 1, n, sr, W, h, N, 1, _ptr(pv_ms), ctypes.byref(ar)) == -1

    # host PV -> audio, stage-wise against the oracle on the same PV
    out = np.zeros((2, F * h), np.float32)
    assert api.api_convert_to_audio(_ptr(ref), 2, F, B, sr, ar.value, W, _ptr(out)) == F * h
    assert_synthesis_parity(out, oracle.convert_to_audio(ref, sr, ar.value, W))


@pytest.mark.gpu
def test_cpp_api_device_resident_chain_and_dirty_tracking(api, oracle):
    sr, W, h, N = 48000.0, 1024, 64, 1024
    n = 20000
    x = np.stack([noise_chirp(n, sr, 4), noise_chirp(n, sr, 5)])
    F = n // h + 1
    out = np.zeros((2, F * h), np.float32)
    assert api.api_round_trip(_ptr(x), 2, n, sr, W, h, N, 0, 0, _ptr(out)) == F * h
    touched = np.zeros_like(out)
    assert api.api_round_trip(_ptr(x), 2, n, sr, W, h, N, 1, 0, _ptr(touched)) == F * h
    # editing one MF through the reference-style accessor must reach the GPU (host copy became the newer one)
    assert np.abs(touched - out).max() > 0
    # the untouched chain equals engine synthesis of engine analysis: compare against the oracle chain loosely
    # (independent FFTs) and check the round-trip gain
    ref = oracle.convert_to_audio(oracle.convert_to_pv(x, sr, W, h, N), sr, oracle.analysis_rate(sr, h), W)
    assert np.abs(out - ref).max() < 2e-2
    lr = np.zeros_like(out)
    assert api.api_round_trip(_ptr(x), 2, n, sr, W, h, N, 0, 1, _ptr(lr)) == F * h
    assert np.array_equal(lr.view(np.uint32), oracle.mid_side(out).view(np.uint32))


def _grid(F, B, sr, ar, N):
    """The grid the reference samples a Function<TF, float> on (PV/PV.h:31-35), in float32."""
    x_scale = np.float32(1.0) / np.float32(ar)
    y_scale = np.float32(1.0) * np.float32(sr) / np.float32(N)
    t = (np.arange(F, dtype=np.float32) * x_scale)[:, None] * np.ones((1, B), np.float32)
    f = np.ones((F, 1), np.float32) * (np.arange(B, dtype=np.float32) * y_scale)[None, :]
    return t, f


@pytest.mark.gpu
@pytest.mark.parametrize("kind,interp", [(0, 0), (0, 5), (1, 0), (1, 2)])
def test_cpp_api_repitch_stretch_chain(api, oracle, kind, interp):
    # audio.convert_to_PV().repitch().stretch().convert_to_audio() as user code writes it; the PV-domain steps are
    # bit-exact against the oracle run on the engine's own analysis output.
    i, f = ctypes.c_int, ctypes.c_float
    api.api_chain.argtypes = [_fp, i, i, f, i, i, i, i, i, _fp, ctypes.POINTER(i), _fp]
    sr, W, h, N = 48000.0, 512, 64, 512
    n = 12000
    x = np.stack([noise_chirp(n, sr, 8), sine_sweep(n, sr)])
    F, B = n // h + 1, N // 2 + 1
    pv = np.zeros((2, F, B, 2), np.float32)
    ar = ctypes.c_float()
    assert api.api_convert_to_pv(_ptr(x), 2, n, sr, W, h, N, 0, _ptr(pv), ctypes.byref(ar)) == F
    t, fr = _grid(F, B, sr, ar.value, N)
    if kind == 0:
        fac_r, fac_s = np.full((F, B), 1.5, np.float32), np.full((F, B), 2.0, np.float32)
    else:
        fac_r = np.float32(0.75) + np.float32(0.5) * t
        fac_s = np.float32(1.0) + fr / np.float32(24000.0)
    want = oracle.stretch(oracle.repitch(pv, sr, fac_r, interp), sr, ar.value, fac_s, interp)
    got = np.zeros_like(want)
    frames = ctypes.c_int(0)
    out = np.zeros((2, want.shape[1] * h), np.float32)
    rc = api.api_chain(_ptr(x), 2, n, sr, W, h, N, kind, interp, _ptr(got), ctypes.byref(frames), _ptr(out))
    assert frames.value == want.shape[1] and rc == want.shape[1] * h
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert_synthesis_parity(out, oracle.convert_to_audio(want, sr, ar.value, W))


@pytest.mark.gpu
def test_cpp_api_modify_time_and_frequency(api, oracle):
    i, f = ctypes.c_int, ctypes.c_float
    api.api_modify_maps.argtypes = [_fp, i, i, f, i, i, i, _fp, ctypes.POINTER(i), _fp]
    sr, W, h, N = 44100.0, 512, 32, 512
    n = 6000
    x = np.stack([noise_chirp(n, sr, 9)])
    F, B = n // h + 1, N // 2 + 1
    pv = np.zeros((1, F, B, 2), np.float32)
    ar = ctypes.c_float()
    assert api.api_convert_to_pv(_ptr(x), 1, n, sr, W, h, N, 0, _ptr(pv), ctypes.byref(ar)) == F
    t, fr = _grid(F, B, sr, ar.value, N)
    want_t = oracle.modify_time(pv, sr, ar.value, t * np.float32(1.25) + np.float32(0.01), 0)
    got_t = np.zeros_like(want_t)
    got_f = np.zeros_like(pv)
    frames = ctypes.c_int(0)
    # returns 1 when the user-callable Interpolator was (correctly) refused with a null PV
    assert api.api_modify_maps(_ptr(x), 1, n, sr, W, h, N, _ptr(got_t), ctypes.byref(frames), _ptr(got_f)) == 1
    assert frames.value == want_t.shape[1]
    assert np.array_equal(got_t.view(np.uint32), want_t.view(np.uint32))
    # modify_frequency: the scatter of modify_frequency_base with mod sampled on the grid and at every MF's frequency.
    # Restated here with the oracle's repitch internals is not possible (different in_mod), so check the invariants the
    # scatter guarantees: output magnitudes are input magnitudes of the same frame (or 0), mapped frequencies follow mod.
    mapped = pv[..., 1] * np.float32(0.8) + np.float32(30.0)
    nz = got_f[..., 0] > 0
    assert nz.any()
    for fi in (0, F // 2, F - 1):
        assert np.isin(got_f[0, fi, nz[0, fi], 0], pv[0, fi, :, 0]).all()
        assert np.isin(got_f[0, fi, nz[0, fi], 1], mapped[0, fi]).all()


@pytest.mark.gpu
def test_example_program_runs_the_chain_from_file_to_file(tmp_path):
    # examples/pv_chain.cpp: WAV in -> convert_to_PV -> repitch -> stretch -> convert_to_audio -> WAV out, all through the
    # reference-shaped C++ API (the reference's own tests/flanTest.cpp:32-47 does the same with Flan + FFTW + libsndfile)
    import os
    import subprocess
    import wave
    from flan_b200 import build
    build.build_host()
    exe = os.path.join(os.path.dirname(build.host_path()), "pv_chain")
    mid, out = str(tmp_path / "mid.wav"), str(tmp_path / "out.wav")
    r = subprocess.run([exe, "--synthetic", "3", mid, "1.0", "1.0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe, mid, out, "1.5", "2.0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    with wave.open(mid, "rb") as w:
        n_mid = w.getnframes()
    with wave.open(out, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (2, 3, 48000)
        n_out = w.getnframes()
        raw = np.frombuffer(w.readframes(n_out), np.uint8).reshape(-1, 3).astype(np.int32)
    assert abs(n_out - 2 * n_mid) <= 2 * 2048                     # stretched by two
    v = ((raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)) << 8) >> 8
    assert np.sqrt(np.mean((v / 8388608.0) ** 2)) > 0.01          # and audible
