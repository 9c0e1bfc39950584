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
                                      ctypes.c_int, _i64, _i64, ctypes.c_int, ctypes.c_int, _fp, _i64, ctypes.c_int, ctypes.c_void_p]
        L.pv_emu_synthesis.argtypes = [_fp, _i64, ctypes.c_int, _i64, _i64, _i64, ctypes.c_int, ctypes.c_float,
                                       ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, _fp, _i64, _i64, _i64, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        L.pv_emu_phase_segments.argtypes = [_fp, ctypes.c_int, _i64, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        L.pv_emu_tables.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, _fp, _fp, _fp]
        L.pv_emu_div_const_mismatches.restype = _i64
        L.pv_emu_div_const_mismatches.argtypes = [ctypes.c_float, ctypes.c_uint32, _i64]
        L.pv_emu_round_mismatches.restype = _i64
        L.pv_emu_round_mismatches.argtypes = [ctypes.c_uint32, _i64]
        self.L = L

    def analysis(self, audio, sr, W, hop, N, frame_begin=0, frame_end=None, seg_len=0, sms=4,
                 audio_offset=0, n_total=None, points_per_thread=8, emit_summary=False):
        """emit_summary (116: 16 points per thread, one buffer; needs seg_len): also returns the phase summaries the kernel
        leaves, float64 [C, segs, B, 4] = (sum.q, sum.r, max.q, max.r); a NaN in sum.q marks an entry left to the scan."""
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
        seg = None
        if emit_summary:
            assert seg_len > 0
            seg = np.full((C, -(-rows // seg_len), B, 4), 7.0, np.float64)
        rc = self.L.pv_emu_analysis(_ptr(audio), n_local, audio_offset, C, n_total, sr, W, hop, N, frame_begin,
                                    frame_end, seg_len, sms, _ptr(pv), rows * B, points_per_thread,
                                    None if seg is None else seg.ctypes.data)
        assert rc == 0, rc
        return (pv, seg) if emit_summary else pv

    def phase_segments(self, pv, sr, ar, W, seg_len):
        """The summaries pv_phase_seg_kernel computes from the rows: float64 [C, segs, B, 4], and the NaN / Inf flag."""
        pv = np.ascontiguousarray(pv, np.float32)
        C, F, B, _ = pv.shape
        out = np.zeros((C, -(-F // seg_len), B, 4), np.float64)
        flag = ctypes.c_int(0)
        assert self.L.pv_emu_phase_segments(_ptr(pv), C, ctypes.c_int64(F), B, ctypes.c_float(sr), ctypes.c_float(ar), W, seg_len,
                                            out.ctypes.data, ctypes.byref(flag)) == 0
        return out, flag.value

    def synthesis(self, pv, sr, ar, W, frame_begin=0, frames_total=None, seg_len=0, sms=4, carry_in=None,
                  want_carry=False, out_offset=None, out_len=None, synth=True, variant=8):
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
                                     _ptr(out) if synth else None, out_len, out_offset, out_len, ctypes.byref(flag), variant)
        assert rc == 0, rc
        return out, carry_out, flag.value

    def choose_seg_len(self, frames, channels, sms, W, hop, max_len, dft, analysis_only=False):
        return int(self.L.pv_emu_choose_seg_len(ctypes.c_int64(frames), channels, sms, W, hop, max_len, int(analysis_only), dft))

    def tables(self, N, W, hop, sr, ar):
        wa, ws, ex = np.empty(W, np.float32), np.empty(W, np.float32), np.empty(N // 2 + 1, np.float32)
        assert self.L.pv_emu_tables(N, W, hop, sr, ar, _ptr(wa), _ptr(ws), _ptr(ex)) == 0
        return wa, ws, ex

    def round_mismatches(self, first_bits, count):
        return int(self.L.pv_emu_round_mismatches(first_bits, count))

    def div_const_mismatches(self, c, first_bits, count):
        return int(self.L.pv_emu_div_const_mismatches(c, first_bits, count))


class EmuModify:
    """The PV-domain kernel bodies (flan_b200/csrc/pv_modify_body.cuh) run thread by thread on the host."""

    def __init__(self):
        L = Emu().L
        i, f = ctypes.c_int, ctypes.c_float
        L.pv_emu_repitch.argtypes = [_fp, i, _i64, i, f, _fp, _i64, i, _fp, _fp, i, i, _fp]
        L.pv_emu_stretch.restype = _i64
        L.pv_emu_stretch.argtypes = [_fp, i, _i64, i, f, f, _fp, _fp, _i64, i, i, i, i, ctypes.POINTER(i), _fp]
        self.L = L

    @staticmethod
    def view(table, F, B):
        """(array, frame_stride, bin_stride) of a table given as [F][B], [B] (shared row), [F,1] (shared column) or scalar."""
        t = np.ascontiguousarray(table, np.float32)
        if t.shape == (F, B):
            return t, B, 1
        if t.shape == (B,):
            return t, 0, 1
        if t.shape == (F, 1):
            return t, 1, 0
        assert t.size == 1
        return t.reshape(1), 0, 0

    def repitch(self, pv, sr, factor, interp=0, threads=64):
        pv = np.ascontiguousarray(pv, np.float32)
        C, F, B, _ = pv.shape
        t, fs, bs = self.view(factor, F, B)
        out = np.full_like(pv, np.nan)
        assert self.L.pv_emu_repitch(_ptr(pv), C, F, B, sr, _ptr(t), fs, bs, None, None, interp, threads, _ptr(out)) == 0
        return out

    def modify_frequency(self, pv, sr, mod_hz, in_mod, interp=0, threads=64):
        pv = np.ascontiguousarray(pv, np.float32)
        C, F, B, _ = pv.shape
        t, fs, bs = self.view(mod_hz, F, B)
        in_mod = np.ascontiguousarray(in_mod, np.float32)
        out = np.full_like(pv, np.nan)
        assert self.L.pv_emu_repitch(_ptr(pv), C, F, B, sr, None, fs, bs, _ptr(t), _ptr(in_mod), interp, threads, _ptr(out)) == 0
        return out

    def _time(self, pv, sr, ar, factor, seconds, interp, chunk, force_sequential):
        pv = np.ascontiguousarray(pv, np.float32)
        C, F, B, _ = pv.shape
        t, fs, bs = self.view(factor if factor is not None else seconds, F, B)
        fa = _ptr(t) if factor is not None else None
        se = _ptr(t) if factor is None else None
        used = ctypes.c_int(0)
        frames = int(self.L.pv_emu_stretch(_ptr(pv), C, F, B, sr, ar, fa, se, fs, bs, interp, chunk, force_sequential,
                                           ctypes.byref(used), None))
        if frames <= 0:
            return np.zeros((C, 0, B, 2), np.float32), bool(used.value)
        out = np.full((C, frames, B, 2), np.nan, np.float32)
        assert self.L.pv_emu_stretch(_ptr(pv), C, F, B, sr, ar, fa, se, fs, bs, interp, chunk, force_sequential,
                                     ctypes.byref(used), _ptr(out)) == frames
        return out, bool(used.value)

    def stretch(self, pv, sr, ar, factor, interp=0, chunk=32, force_sequential=0):
        return self._time(pv, sr, ar, factor, None, interp, chunk, force_sequential)

    def modify_time(self, pv, sr, ar, seconds, interp=0, chunk=32, force_sequential=0):
        return self._time(pv, sr, ar, None, seconds, interp, chunk, force_sequential)
