"""Adapter giving the CPU thread emulator the method names of flan_b200.engine.Engine (CPU torch tensors),
so the multi-rank orchestration in flan_b200/sharding.py can be exercised under gloo without a GPU."""
import numpy as np
import torch

from emu_lib import Emu


class EmuEngine:
    def __init__(self):
        self.emu = Emu()

    def convert_to_pv_range(self, audio_local, audio_offset, n_total, sr, W, hop, N, frame_begin, frame_end):
        pv = self.emu.analysis(audio_local.numpy(), sr, W, hop, N, frame_begin, frame_end,
                               audio_offset=audio_offset, n_total=n_total, sms=2)
        return torch.from_numpy(pv)

    def phase_summary(self, pv_rows, frame_begin, sr, ar, W):
        _, carry, _ = self.emu.synthesis(pv_rows.numpy(), sr, ar, W, frame_begin=frame_begin, want_carry=True,
                                         synth=False, sms=2)
        return torch.from_numpy(carry)

    def phase_carry(self, all_states, rank):
        # combine states of ranks < rank with the emulator's own scan: feed them as carry chain
        C, B = all_states.shape[1], all_states.shape[2]
        P = float(np.float32(np.float32(np.arccos(np.float32(-1))) * np.float32(2)))
        st = np.zeros((C, B, 4), np.float64)
        for r in range(rank):
            seg = all_states[r].numpy()
            cand_q = st[..., 0] + seg[..., 2]
            cand_r = st[..., 1] + seg[..., 3]
            k = np.floor(cand_r / P); cand_r = cand_r - k * P; cand_q = cand_q + k
            better = (st[..., 2] < cand_q) | ((st[..., 2] == cand_q) & (st[..., 3] < cand_r))
            st[..., 2] = np.where(better, cand_q, st[..., 2])
            st[..., 3] = np.where(better, cand_r, st[..., 3])
            sq = st[..., 0] + seg[..., 0]
            sr_ = st[..., 1] + seg[..., 1]
            k = np.floor(sr_ / P); sr_ = sr_ - k * P; sq = sq + k
            st[..., 0], st[..., 1] = sq, sr_
        return torch.from_numpy(st)

    def convert_to_audio_range(self, pv_rows, frame_begin, frames_total, sr, ar, W, carry, out_offset, out_len,
                               reuse_summary=False):
        out, _, _ = self.emu.synthesis(pv_rows.numpy(), sr, ar, W, frame_begin=frame_begin, frames_total=frames_total,
                                       carry_in=None if carry is None else np.ascontiguousarray(carry.numpy()),
                                       out_offset=out_offset, out_len=out_len, sms=2)
        return torch.from_numpy(out)

    def empty_like_audio(self, C, n):
        return torch.empty((C, n), dtype=torch.float32)

    def add_into(self, dst_view, src):
        dst_view += src
