"""ctypes binding of the CPU thread emulator (flan_b200/lib/libpv_emu.so). TEST HARNESS ONLY:
it runs the kernels' CTA body source on host threads so the logic can be checked without a GPU."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.c_int64


def _ptr(a):
    return a.ctypes.data_as(_fp)


class Emu:
    def __init__(self):
        path = os.path.join(ROOT, "flan_b200", "lib", "libpv_emu.so")
        if not os.path.exists(path):
            from flan_b200 import build
            build.build_emulator()
        L = ctypes.CDLL(path)
        L.pv_emu_analysis.argtypes = [_fp, _i64, _i64, ctypes.c_int, _i64, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, _i64, _i64, ctypes.c_int, ctypes.c_int, _fp, _i64, ctypes.c_int]
        L.pv_emu_synthesis.argtypes = [_fp, _i64, ctypes.c_int, _i64, _i64, _i64, ctypes.c_int, ctypes.c_float,
                                       ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, _fp, _i64, _i64, _i64, ctypes.POINTER(ctypes.c_int)]
        L.pv_emu_tables.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, _fp, _fp, _fp]
        L.pv_emu_div_const_mismatches.restype = _i64
        L.pv_emu_div_const_mismatches.argtypes = [ctypes.c_float, ctypes.c_uint32, _i64]
        self.L = L

    def analysis(self, audio, sr, W, hop, N, frame_begin=0, frame_end=None, seg_len=0, sms=4,
                 audio_offset=0, n_total=None, points_per_thread=8):
        audio = np.ascontiguousarray(audio, np.float32)
        C, n_local = audio.shape
        if n_total is None:
            n_total = n_local
        F = n_total // hop + 1
        if frame_end is None:
            frame_end = F
        rows = frame_end - frame_begin
        B = N // 2 + 1
        pv = np.full((C, rows, B, 2), np.nan, np.float32)
        rc = self.L.pv_emu_analysis(_ptr(audio), n_local, audio_offset, C, n_total, sr, W, hop, N, frame_begin,
                                    frame_end, seg_len, sms, _ptr(pv), rows * B, points_per_thread)
        assert rc == 0, rc
        return pv

    def synthesis(self, pv, sr, ar, W, frame_begin=0, frames_total=None, seg_len=0, sms=4, carry_in=None,
                  want_carry=False, out_offset=None, out_len=None, synth=True):
        pv = np.ascontiguousarray(pv, np.float32)
        C, rows, B, _ = pv.shape
        hop = int(np.float32(sr) / np.float32(ar))
        frame_end = frame_begin + rows
        if frames_total is None:
            frames_total = frame_end
        if out_offset is None:
            out_offset, out_len = 0, frames_total * hop
        out = np.full((C, out_len), np.nan, np.float32)
        carry_out = np.zeros((C, B, 4), np.float64) if want_carry else None
        flag = ctypes.c_int(0)
        rc = self.L.pv_emu_synthesis(_ptr(pv), rows * B, C, frame_begin, frame_end, frames_total, B, sr, ar, W, seg_len, sms,
                                     None if carry_in is None else carry_in.ctypes.data,
                                     None if carry_out is None else carry_out.ctypes.data,
                                     _ptr(out) if synth else None, out_len, out_offset, out_len, ctypes.byref(flag))
        assert rc == 0, rc
        return out, carry_out, flag.value

    def tables(self, N, W, hop, sr, ar):
        wa, ws, ex = np.empty(W, np.float32), np.empty(W, np.float32), np.empty(N // 2 + 1, np.float32)
        assert self.L.pv_emu_tables(N, W, hop, sr, ar, _ptr(wa), _ptr(ws), _ptr(ex)) == 0
        return wa, ws, ex

    def div_const_mismatches(self, c, first_bits, count):
        return int(self.L.pv_emu_div_const_mismatches(c, first_bits, count))
