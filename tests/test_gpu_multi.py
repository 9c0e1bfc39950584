"""Several GPUs behind one handle (flan_b200_multi_*, pv_capi_multi.cu): a signal cut into frame-range shards must
give the single-device result BIT FOR BIT -- analysis needs no exchange, resynthesis exchanges the per-bin phase state and
the overlap-add halo device to device. On a box with one GPU the same device is listed several times, which runs every
line of the orchestration (separate contexts and streams, peer copies between them); with more GPUs the real ones."""
import ctypes

import numpy as np
import pytest

from flan_b200 import capi
from flan_b200.signals import noise_chirp, sine_sweep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from flan_b200.engine import Engine
    return Engine(0)


class Multi:
    def __init__(self, devices):
        self.lib = capi.load()
        arr = (ctypes.c_int * len(devices))(*devices)
        self.h = ctypes.c_void_p()
        rc = self.lib.flan_b200_multi_create(arr, len(devices), ctypes.byref(self.h))
        assert rc == 0, self.lib.flan_b200_multi_last_error(None)

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        assert rc == 0, (name, self.lib.flan_b200_multi_last_error(self.h))

    def close(self):
        self.lib.flan_b200_multi_destroy(self.h)


def _devices(k):
    import torch
    n = torch.cuda.device_count()
    return list(range(min(n, k))) if n >= 2 else [0] * k


@pytest.mark.parametrize("sr,W,h,N,C,n,k", [
    (48000.0, 1024, 64, 1024, 2, 700_000, 3),          # mirrored kernels, 3 shards
    (44100.0, 2048, 128, 2048, 1, 1_500_000, 4),
    (48000.0, 2048, 128, 4096, 2, 900_000, 2),         # the API-default shape (zero-padded)
    (96000.0, 8192, 512, 8192, 1, 4_000_000, 4),       # cfg3's shape
    (48000.0, 4096, 256, 4096, 1, 48000 * 500, 2),     # > 256 segments per shard: the analysis can leave the phase summaries
    (32000.0, 3000, 100, 3000, 1, 300_000, 3),         # the run-time-sized transform
    (48000.0, 512, 32, 512, 2, 9_000, 8),              # short signal: fewer shards than devices
])
def test_sharded_round_trip_is_bit_identical_to_one_device(eng, sr, W, h, N, C, n, k):
    import torch
    x = np.stack([noise_chirp(n, sr, 50 + c) if c % 2 == 0 else sine_sweep(n, sr) for c in range(C)])
    F, B = n // h + 1, N // 2 + 1
    pv_ref = eng.convert_to_pv(torch.from_numpy(x).cuda(), sr, W, h, N)
    ar = eng.analysis_rate(sr, h)
    y_ref = eng.convert_to_audio(pv_ref, sr, ar, W).cpu().numpy()
    pv_ref = pv_ref.cpu().numpy()

    m = Multi(_devices(k))
    shards = ctypes.c_int()
    fb = (ctypes.c_int64 * (capi.MAX_DEVICES + 1))()
    m.call("flan_b200_multi_plan", C, n, W, h, N, ctypes.byref(shards), fb)
    assert 1 <= shards.value <= k and fb[0] == 0 and fb[shards.value] == F
    reach = 2 * -(-W // h)
    assert all(fb[i + 1] - fb[i] >= min(reach, F) for i in range(shards.value - 1))

    a, pv, out = capi.ShardedAudio(), capi.ShardedPV(), capi.ShardedAudio()
    m.call("flan_b200_multi_scatter_audio", x.ctypes.data, C, n, W, h, N, ctypes.byref(a))
    m.call("flan_b200_multi_convert_to_pv", ctypes.byref(a), sr, W, h, N, ctypes.byref(pv))
    pv_h = np.empty((C, F, B, 2), np.float32)
    m.call("flan_b200_multi_gather_pv", ctypes.byref(pv), pv_h.ctypes.data, 0, None)
    assert np.array_equal(pv_h.view(np.uint32), pv_ref.view(np.uint32))
    m.call("flan_b200_multi_convert_to_audio", ctypes.byref(pv), ctypes.byref(out))
    y = np.empty((C, F * h), np.float32)
    m.call("flan_b200_multi_gather_audio", ctypes.byref(out), y.ctypes.data)
    assert np.array_equal(y.view(np.uint32), y_ref.view(np.uint32))
    # the round trip of a caller that says so (flan_b200_multi_hint_resynthesis / _promise_unchanged): the shards' analysis
    # leaves their phase summaries where that form exists (full window, dft 2048 ... 8192, > 256 segments per shard), the
    # statements are ignored elsewhere; rows and samples keep their bits
    pv3, out3 = capi.ShardedPV(), capi.ShardedAudio()
    m.call("flan_b200_multi_hint_resynthesis")
    m.call("flan_b200_multi_convert_to_pv", ctypes.byref(a), sr, W, h, N, ctypes.byref(pv3))
    m.call("flan_b200_multi_gather_pv", ctypes.byref(pv3), pv_h.ctypes.data, 0, None)
    assert np.array_equal(pv_h.view(np.uint32), pv_ref.view(np.uint32))
    m.call("flan_b200_multi_promise_unchanged", ctypes.byref(pv3))
    m.call("flan_b200_multi_convert_to_audio", ctypes.byref(pv3), ctypes.byref(out3))
    y3 = np.empty_like(y)
    m.call("flan_b200_multi_gather_audio", ctypes.byref(out3), y3.ctypes.data)
    assert np.array_equal(y3.view(np.uint32), y_ref.view(np.uint32))
    m.call("flan_b200_multi_free_audio", ctypes.byref(out3))
    m.call("flan_b200_multi_free_pv", ctypes.byref(pv3))
    # the host-buffer forms, twice (device blocks and scratch come back from the caches)
    for _ in range(2):
        pv2 = capi.ShardedPV()
        m.call("flan_b200_multi_convert_to_pv_host", x.ctypes.data, C, n, sr, W, h, N, ctypes.byref(pv2))
        y2 = np.empty_like(y)
        flag = ctypes.c_int(-1)
        m.call("flan_b200_multi_convert_to_audio_host", ctypes.byref(pv2), y2.ctypes.data, ctypes.byref(flag))
        assert flag.value == 0
        assert np.array_equal(y2.view(np.uint32), y_ref.view(np.uint32))
        # gather onto one device
        d = torch.empty((C, F, B, 2), device="cuda:0")
        m.call("flan_b200_multi_gather_pv", ctypes.byref(pv2), None, 0, ctypes.c_void_p(d.data_ptr()))
        m.call("flan_b200_multi_synchronize")
        assert np.array_equal(d.cpu().numpy().view(np.uint32), pv_ref.view(np.uint32))
        m.call("flan_b200_multi_free_pv", ctypes.byref(pv2))
    m.call("flan_b200_multi_free_audio", ctypes.byref(a))
    m.call("flan_b200_multi_free_audio", ctypes.byref(out))
    m.call("flan_b200_multi_free_pv", ctypes.byref(pv))
    m.close()


def test_cpp_api_uses_every_device_for_long_signals(tmp_path):
    """flan::Audio::convert_to_PV / PV::convert_to_audio through libflan_b200_host.so with several devices behind the
    process-wide engine ($FLAN_B200_DEVICES; on a one-GPU box the device is listed three times): a long signal is
    sharded, the PV stays on the devices, and the samples are the single-device samples bit for bit. A PV-domain step in
    between gathers the shards onto one device. Each engine lives in its own process."""
    import os
    import subprocess
    import sys
    import torch
    script = tmp_path / "run.py"
    script.write_text('''
import ctypes, sys, numpy as np
sys.path.insert(0, %r)
from flan_b200 import build
from flan_b200.signals import noise_chirp
L = ctypes.CDLL(build.api_test_path())
fp = ctypes.POINTER(ctypes.c_float)
sr, W, h, N, n = 48000.0, 2048, 128, 2048, 4_600_000
x = np.stack([noise_chirp(n, sr, 77)])
F = n // h + 1
out = np.zeros((1, F * h), np.float32)
L.api_round_trip.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
assert L.api_round_trip(x.ctypes.data_as(fp), 1, n, sr, W, h, N, 0, 0, out.ctypes.data_as(fp)) == F * h
touched = np.zeros_like(out)
assert L.api_round_trip(x.ctypes.data_as(fp), 1, n, sr, W, h, N, 1, 0, touched.ctypes.data_as(fp)) == F * h
np.save(sys.argv[1], np.stack([out, touched]))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    results = []
    n_gpus = torch.cuda.device_count()
    for env_extra in ({"FLAN_B200_DEVICE": "0"}, {"FLAN_B200_DEVICES": ",".join(str(i) for i in range(min(n_gpus, 4))) if n_gpus > 1 else "0,0,0"}):
        env = dict(os.environ)
        env.pop("FLAN_B200_DEVICE", None)
        env.pop("FLAN_B200_DEVICES", None)
        env.update(env_extra)
        path = tmp_path / ("out_%d.npy" % len(results))
        subprocess.run([sys.executable, str(script), str(path)], check=True, env=env, timeout=600)
        results.append(np.load(path))
    assert np.array_equal(results[0].view(np.uint32), results[1].view(np.uint32))
    assert np.abs(results[0][0]).max() > 0.1
