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
