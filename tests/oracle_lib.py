"""ctypes bindings for the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

`Oracle`  -> oracle/libpv_oracle.so : plain-C restatement of the reference's hot path.
`RefLib`  -> oracle/_ref/libflan_ref.so : the reference's own AudioPV.cpp et al. compiled verbatim
             (built in the container where /root/reference exists; travels prebuilt to the GPU box).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_fp = ctypes.POINTER(ctypes.c_float)


def _ptr(a):
    return a.ctypes.data_as(_fp)


def build_oracle(ref=True):
    """Compile the checkers (idempotent). The reference build needs /root/reference."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src/flan"):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)


class Oracle:
    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libpv_oracle.so")
        if not os.path.exists(path):
            build_oracle(ref=False)
        L = ctypes.CDLL(path)
        L.pvo_num_frames.restype = ctypes.c_int64
        L.pvo_num_frames.argtypes = [ctypes.c_int64, ctypes.c_int]
        L.pvo_hann.argtypes = [ctypes.c_int, _fp]
        L.pvo_convert_to_pv.argtypes = [_fp, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, _fp]
        L.pvo_convert_to_audio.argtypes = [_fp, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                           ctypes.c_float, ctypes.c_int, _fp]
        L.pvo_mid_side.argtypes = [_fp, ctypes.c_int64, _fp]
        L.pvo_hop_from_rates.argtypes = [ctypes.c_float, ctypes.c_float]
        L.pvo_repitch.argtypes = [_fp, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_float, _fp, ctypes.c_int, _fp]
        for fn in (L.pvo_stretch, L.pvo_modify_time):
            fn.restype = ctypes.c_int64
            fn.argtypes = [_fp, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_float, _fp,
                           ctypes.c_int, _fp]
        _bp = ctypes.POINTER(ctypes.c_uint8)
        L.pvo_flan_encode.argtypes = [_fp, ctypes.c_int64, ctypes.c_float, ctypes.c_float, _bp]
        L.pvo_flan_decode.argtypes = [_bp, ctypes.c_int64, ctypes.c_float, ctypes.c_float, _fp]
        L.pvo_pcm24_encode.argtypes = [_fp, ctypes.c_int, ctypes.c_int64, _bp]
        L.pvo_pcm24_decode.argtypes = [_bp, ctypes.c_int, ctypes.c_int64, _fp]
        self.L = L

    # file formats either side of the path (PVBuffer.cpp:99-140,216-273; AudioBuffer.cpp:80-192)
    def flan_encode(self, pv, sr):
        pv = np.ascontiguousarray(pv, np.float32)
        C, F, B, _ = pv.shape
        out = np.empty(C * F * B * 6, np.uint8)
        self.L.pvo_flan_encode(_ptr(pv), C * F * B, float((B - 1) * 2), sr, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return out

    def flan_decode(self, data, shape, sr):
        C, F, B = shape
        data = np.ascontiguousarray(data, np.uint8)
        out = np.empty((C, F, B, 2), np.float32)
        self.L.pvo_flan_decode(data.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), C * F * B, float((B - 1) * 2), sr, _ptr(out))
        return out

    def pcm24_encode(self, audio):
        audio = np.ascontiguousarray(audio, np.float32)
        C, n = audio.shape
        out = np.empty(C * n * 3, np.uint8)
        self.L.pvo_pcm24_encode(_ptr(audio), C, n, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return out

    def pcm24_decode(self, data, C, n):
        data = np.ascontiguousarray(data, np.uint8)
        out = np.empty((C, n), np.float32)
        self.L.pvo_pcm24_decode(data.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), C, n, _ptr(out))
        return out

    def num_frames(self, n, hop):
        return int(self.L.pvo_num_frames(n, hop))

    def hann(self, W):
        out = np.empty(W, np.float32)
        self.L.pvo_hann(W, _ptr(out))
        return out

    def analysis_rate(self, sr, hop):
        return np.float32(np.float32(sr) / np.float32(hop))

    def convert_to_pv(self, audio, sr, W, hop, N, frame_begin=0, frame_end=None):
        audio = np.ascontiguousarray(audio, np.float32)
        C, n = audio.shape
        F = self.num_frames(n, hop)
        if frame_end is None:
            frame_end = F
        pv = np.empty((C, frame_end - frame_begin, N // 2 + 1, 2), np.float32)
        rc = self.L.pvo_convert_to_pv(_ptr(audio), C, n, sr, W, hop, N, frame_begin, frame_end, _ptr(pv))
        if rc != 0:
            raise ValueError("pvo_convert_to_pv rejected its arguments")
        return pv

    def convert_to_audio(self, pv, sr, analysis_rate, W):
        pv = np.ascontiguousarray(pv, np.float32)
        C, F, B, _ = pv.shape
        hop = int(self.L.pvo_hop_from_rates(sr, analysis_rate))
        out = np.empty((C, F * hop), np.float32)
        rc = self.L.pvo_convert_to_audio(_ptr(pv), C, F, B, sr, analysis_rate, W, _ptr(out))
        if rc != 0:
            raise ValueError("pvo_convert_to_audio rejected its arguments")
        return out

    def mid_side(self, audio):
        audio = np.ascontiguousarray(audio, np.float32)
        assert audio.shape[0] == 2
        out = np.empty_like(audio)
        self.L.pvo_mid_side(_ptr(audio), audio.shape[1], _ptr(out))
        return out

    # PV-domain chain (PV/PVModify.cpp): factor / mod tables are float[F][B]
    def repitch(self, pv, sr, factor, interp=0):
        pv = np.ascontiguousarray(pv, np.float32)
        factor = np.ascontiguousarray(factor, np.float32)
        C, F, B, _ = pv.shape
        assert factor.shape == (F, B)
        out = np.empty_like(pv)
        assert self.L.pvo_repitch(_ptr(pv), C, F, B, sr, _ptr(factor), interp, _ptr(out)) == 0
        return out

    def _time(self, fn, pv, sr, ar, table, interp):
        pv = np.ascontiguousarray(pv, np.float32)
        table = np.ascontiguousarray(table, np.float32)
        C, F, B, _ = pv.shape
        assert table.shape == (F, B)
        frames = int(fn(_ptr(pv), C, F, B, sr, ar, _ptr(table), interp, None))
        if frames <= 0:
            return np.zeros((C, 0, B, 2), np.float32)
        out = np.empty((C, frames, B, 2), np.float32)
        assert fn(_ptr(pv), C, F, B, sr, ar, _ptr(table), interp, _ptr(out)) == frames
        return out

    def stretch(self, pv, sr, ar, factor, interp=0):
        return self._time(self.L.pvo_stretch, pv, sr, ar, factor, interp)

    def modify_time(self, pv, sr, ar, mod_seconds, interp=0):
        return self._time(self.L.pvo_modify_time, pv, sr, ar, mod_seconds, interp)


class RefLib:
    """The reference's own sources, compiled verbatim. backend 0 = f64 FFT stand-in (parity),
    1 = the reference's vendored pffft (timing)."""

    @staticmethod
    def path():
        return os.path.join(ORACLE_DIR, "_ref", "libflan_ref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.path())

    def __init__(self, backend=0):
        L = ctypes.CDLL(self.path())
        L.flan_ref_convert_to_pv.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, _fp, _fp]
        L.flan_ref_convert_to_audio.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                ctypes.c_float, ctypes.c_int, ctypes.c_int, _fp]
        L.flan_ref_bench.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, _fp]
        L.flan_ref_hann.argtypes = [ctypes.c_int, _fp]
        self.L = L
        self.set_backend(backend)

    def set_backend(self, backend):
        self.L.flan_ref_set_fft_backend(backend)

    def hann(self, W):
        out = np.empty(W, np.float32)
        self.L.flan_ref_hann(W, _ptr(out))
        return out

    def convert_to_pv(self, audio, sr, W, hop, N, ms=False):
        audio = np.ascontiguousarray(audio, np.float32)
        C, n = audio.shape
        F = self.L.flan_ref_num_frames(n, hop)
        pv = np.empty((C, F, N // 2 + 1, 2), np.float32)
        ar = ctypes.c_float()
        rc = self.L.flan_ref_convert_to_pv(_ptr(audio), C, n, sr, W, hop, N, int(ms), _ptr(pv), ctypes.byref(ar))
        if rc < 0:
            return None, None
        return pv, np.float32(ar.value)

    def convert_to_audio(self, pv, sr, analysis_rate, W, lr=False):
        pv = np.ascontiguousarray(pv, np.float32)
        C, F, B, _ = pv.shape
        hop = int(np.float32(sr) / np.float32(analysis_rate))
        out = np.empty((C, F * hop), np.float32)
        rc = self.L.flan_ref_convert_to_audio(_ptr(pv), C, F, B, sr, analysis_rate, W, int(lr), _ptr(out))
        if rc < 0:
            return None
        return out

    def save_flan(self, path, pv, sr, ar, W):
        pv = np.ascontiguousarray(pv, np.float32)
        C, F, B, _ = pv.shape
        self.L.flan_ref_save_flan.argtypes = [ctypes.c_char_p, _fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                              ctypes.c_float, ctypes.c_int]
        assert self.L.flan_ref_save_flan(path.encode(), _ptr(pv), C, F, B, sr, ar, W) == 0

    def load_flan(self, path):
        """The reference's own PVBuffer::load: returns (pv, sample_rate, analysis_rate_as_loaded, window_size)."""
        self.L.flan_ref_load_flan.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), _fp, _fp]
        shape = (ctypes.c_int * 4)()
        rates = (ctypes.c_float * 2)()
        assert self.L.flan_ref_load_flan(path.encode(), shape, rates, None) == 0
        pv = np.empty((shape[0], shape[1], shape[2], 2), np.float32)
        assert self.L.flan_ref_load_flan(path.encode(), shape, rates, _ptr(pv)) == 0
        return pv, rates[0], rates[1], shape[3]

    def bench(self, audio, sr, W, hop, N, mode):
        audio = np.ascontiguousarray(audio, np.float32)
        C, n = audio.shape
        cs = ctypes.c_float()
        return self.L.flan_ref_bench(_ptr(audio), C, n, sr, W, hop, N, mode, ctypes.byref(cs))


class RefModifyLib:
    """The reference's own PV/PVModify.cpp (PV::repitch / stretch / modify_time), compiled verbatim with its real
    Function / FunctionSample / Interpolator headers."""

    @staticmethod
    def path():
        return os.path.join(ORACLE_DIR, "_ref", "libflan_ref_modify.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.path())

    def __init__(self):
        L = ctypes.CDLL(self.path())
        i, f = ctypes.c_int, ctypes.c_float
        L.flan_ref_repitch.argtypes = [_fp, i, i, i, f, f, i, _fp, i, _fp]
        L.flan_ref_stretch.argtypes = [_fp, i, i, i, f, f, i, _fp, i, _fp, i]
        L.flan_ref_modify_time.argtypes = [_fp, i, i, i, f, f, i, _fp, i, _fp, i]
        self.L = L

    def repitch(self, pv, sr, ar, W, factor, interp=0):
        pv = np.ascontiguousarray(pv, np.float32)
        factor = np.ascontiguousarray(factor, np.float32)
        C, F, B, _ = pv.shape
        out = np.empty_like(pv)
        assert self.L.flan_ref_repitch(_ptr(pv), C, F, B, sr, ar, W, _ptr(factor), interp, _ptr(out)) == F
        return out

    def _time(self, fn, pv, sr, ar, W, table, interp):
        pv = np.ascontiguousarray(pv, np.float32)
        table = np.ascontiguousarray(table, np.float32)
        C, F, B, _ = pv.shape
        frames = fn(_ptr(pv), C, F, B, sr, ar, W, _ptr(table), interp, None, 0)
        if frames <= 0:
            return np.zeros((C, 0, B, 2), np.float32)
        out = np.empty((C, frames, B, 2), np.float32)
        assert fn(_ptr(pv), C, F, B, sr, ar, W, _ptr(table), interp, _ptr(out), frames) == frames
        return out

    def stretch(self, pv, sr, ar, W, factor, interp=0):
        return self._time(self.L.flan_ref_stretch, pv, sr, ar, W, factor, interp)

    def modify_time(self, pv, sr, ar, W, mod_seconds, interp=0):
        return self._time(self.L.flan_ref_modify_time, pv, sr, ar, W, mod_seconds, interp)
