"""Where does the multi-device round trip spend its time? (development aid)  python tools/multi_debug.py [gpus] [seconds]"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flan_b200 import capi  # noqa: E402
from flan_b200.signals import noise_chirp  # noqa: E402

HINT = "--hint" in sys.argv
if HINT:
    sys.argv.remove("--hint")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 1800.0
lib = capi.load()
sr, w, hop, n_dft = 96000.0, 8192, 512, 8192
n = int(sr * seconds)
chunk = noise_chirp(int(sr * 60), sr, 3)
x = np.ascontiguousarray(np.tile(chunk, int(seconds // 60) + 1)[None, :n])
F = n // hop + 1
for kk in sorted({1, k}):
    devs = (ctypes.c_int * kk)(*range(kk))
    h = ctypes.c_void_p()
    assert lib.flan_b200_multi_create(devs, kk, ctypes.byref(h)) == 0

    def call(name, *a):
        rc = getattr(lib, name)(h, *a)
        assert rc == 0, (name, lib.flan_b200_multi_last_error(h))
    a = capi.ShardedAudio()
    call("flan_b200_multi_scatter_audio", x.ctypes.data, 1, n, w, hop, n_dft, ctypes.byref(a))

    def analysis():
        pv = capi.ShardedPV()
        if HINT:
            call("flan_b200_multi_hint_resynthesis")
        call("flan_b200_multi_convert_to_pv", ctypes.byref(a), sr, w, hop, n_dft, ctypes.byref(pv))
        return pv

    def timed(fn, reps=6):
        for _ in range(2):
            fn()
        ms = ctypes.c_double(0)
        call("flan_b200_multi_time_begin")
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        t_enq = time.perf_counter() - t0
        call("flan_b200_multi_time_end", ctypes.byref(ms))
        return ms.value / reps, t_enq / reps * 1e3

    def only_analysis():
        pv = analysis()
        call("flan_b200_multi_free_pv", ctypes.byref(pv))
    print(kk, "GPUs: analysis only   %.3f ms device, %.3f ms host enqueue" % timed(only_analysis), flush=True)
    pv = analysis()

    def only_synthesis():
        y = capi.ShardedAudio()
        call("flan_b200_multi_convert_to_audio", ctypes.byref(pv), ctypes.byref(y))
        call("flan_b200_multi_free_audio", ctypes.byref(y))
    print(kk, "GPUs: resynthesis only %.3f ms device, %.3f ms host enqueue" % timed(only_synthesis), flush=True)
    call("flan_b200_multi_free_pv", ctypes.byref(pv))

    def both():
        only_analysis_pv = analysis()
        y = capi.ShardedAudio()
        if HINT:
            call("flan_b200_multi_promise_unchanged", ctypes.byref(only_analysis_pv))
        call("flan_b200_multi_convert_to_audio", ctypes.byref(only_analysis_pv), ctypes.byref(y))
        call("flan_b200_multi_free_pv", ctypes.byref(only_analysis_pv))
        call("flan_b200_multi_free_audio", ctypes.byref(y))
    print(kk, "GPUs: round trip       %.3f ms device, %.3f ms host enqueue" % timed(both), flush=True)
    call("flan_b200_multi_free_audio", ctypes.byref(a))
    lib.flan_b200_multi_destroy(h)
