"""e2e through the C++ API (tools/cpp/e2e_bench.cpp) for 1..4 host threads on the cfg2 shape; prints ms per round trip."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flan_b200 import build  # noqa: E402
from flan_b200.signals import noise_chirp  # noqa: E402

SR, W, HOP, N, CH = 48000.0, 4096, 256, 4096, 2
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
L = ctypes.CDLL(build.e2e_bench_path())
L.e2e_round_trips.restype = ctypes.c_double
L.e2e_round_trips.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
n = int(SR * seconds)
audio = np.stack([noise_chirp(n, SR, 1234 + c) for c in range(CH)])
F = n // HOP + 1
chk = ctypes.c_double(0)
for threads in (1, 2, 3, 4):
    dt = L.e2e_round_trips(audio.ctypes.data, CH, n, SR, W, HOP, N, threads, 6, 20, ctypes.byref(chk))
    print("threads %d: %.3f ms per round trip, %.1f M frames/s" % (threads, dt / (20 * threads) * 1e3, threads * 20 * CH * F / dt / 1e6), flush=True)

# timeline of a few round trips with two threads (kernel and copy-slice events on their own streams)
from flan_b200 import capi  # noqa: E402
lib = capi.load()
H = ctypes.CDLL(build.host_path())
H.flan_b200_host_context.restype = ctypes.c_void_p
ctx = ctypes.c_void_p(H.flan_b200_host_context())
NAMES = {0: "analysis", 1: "phase_seg", 2: "phase_scan", 3: "synthesis", 9: "upload", 10: "download"}
for threads in (2,):
    lib.flan_b200_set_timing(ctx, 1)
    L.e2e_round_trips(audio.ctypes.data, CH, n, SR, W, HOP, N, threads, 6, 3, ctypes.byref(chk))
    cap = 4096
    kinds, t0, t1, cnt = (ctypes.c_int * cap)(), (ctypes.c_double * cap)(), (ctypes.c_double * cap)(), ctypes.c_int(0)
    lib.flan_b200_trace(ctx, kinds, t0, t1, cap, ctypes.byref(cnt))
    lib.flan_b200_set_timing(ctx, 0)
    rows = sorted((t0[i], t1[i], kinds[i]) for i in range(cnt.value))
    last = rows[-1][1]
    print("timeline, %d threads, last %d entries (ms):" % (threads, min(len(rows), 90)))
    for a, b, k in rows[-90:]:
        print("  %9.3f - %9.3f  %-10s %.3f" % (a - last, b - last, NAMES.get(k, str(k)), b - a))
