"""World-size-2 (and 3) run of the frame-range sharded path on CPU: torch.distributed with the gloo backend carries
the phase-state all_gather and the overlap-add halo send/recv of flan_b200/sharding.py; the transforms are the
kernels' CTA bodies under the host thread emulator. The assembled result must equal the oracle's unsharded one."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SR, W, HOP, N = 48000.0, 512, 32, 512
NSAMP = 9000


def _signal():
    sys.path.insert(0, ROOT)
    from flan_b200.signals import noise_chirp, sine_sweep
    return np.stack([noise_chirp(NSAMP, SR, 21), sine_sweep(NSAMP, SR)])


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emu_engine import EmuEngine
    from flan_b200.sharding import frame_shard, sharded_resynthesis
    eng = EmuEngine()
    x = _signal()
    sh = frame_shard(NSAMP, HOP, W, world, rank)
    local = torch.from_numpy(np.ascontiguousarray(x[:, sh.audio_lo:sh.audio_hi]))
    pv = eng.convert_to_pv_range(local, sh.audio_lo, NSAMP, SR, W, HOP, N, sh.f0, sh.f1)
    ar = float(np.float32(SR) / np.float32(HOP))

    def allgather(state):
        bufs = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(bufs, state.contiguous())
        return torch.stack(bufs)

    out, lo = sharded_resynthesis(eng, dist, sh, pv, SR, ar, allgather,
                                  lambda t, dst: dist.isend(t, dst), lambda t, src: dist.recv(t, src))
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), pv=pv.numpy(), f0=sh.f0, f1=sh.f1,
             own=out[:, sh.own_lo - lo:sh.own_hi - lo].numpy(), own_lo=sh.own_lo, own_hi=sh.own_hi)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_frame_sharded_round_trip_under_gloo(oracle, world):
    from flan_b200 import build
    build.build_emulator()
    from parity import assert_analysis_parity, assert_synthesis_parity
    x = _signal()
    ref_pv = oracle.convert_to_pv(x, SR, W, HOP, N)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
    F = NSAMP // HOP + 1
    pv = np.concatenate([p["pv"] for p in parts], axis=1)
    assert pv.shape[1] == F
    assert_analysis_parity(pv, ref_pv, SR, HOP, N)
    # resynthesis of the shards' own analysis: the oracle on the same PV is the stage-wise reference
    ar = oracle.analysis_rate(SR, HOP)
    ref_audio = oracle.convert_to_audio(pv, SR, ar, W)
    audio = np.zeros_like(ref_audio)
    covered = np.zeros(ref_audio.shape[1], bool)
    for p in parts:
        audio[:, int(p["own_lo"]):int(p["own_hi"])] = p["own"]
        assert not covered[int(p["own_lo"]):int(p["own_hi"])].any()
        covered[int(p["own_lo"]):int(p["own_hi"])] = True
    assert covered.all()
    assert_synthesis_parity(audio, ref_audio)


@pytest.mark.parametrize("n,hop,W,world", [(3000, 32, 512, 8), (20000, 256, 4096, 8), (9000, 32, 512, 3), (100, 64, 1024, 4)])
def test_short_signals_use_fewer_ranks(n, hop, W, world):
    """with fewer than window/hop - 1 frames per rank a frame's window reached past the adjacent shard and
    its partial sums were never delivered. Every sample a frame touches must be owned by the frame's rank or the one
    before it (the receiver of the head overlap), and the owned ranges must tile the output."""
    from flan_b200.sharding import frame_shard
    shards = [frame_shard(n, hop, W, world, r) for r in range(world)]
    F, total = n // hop + 1, (n // hop + 1) * hop
    assert sum(s.frames for s in shards) == F
    pos = 0
    for s in shards:
        if s.frames:
            assert s.own_lo == pos
            pos = s.own_hi
    assert pos == total
    active = [s for s in shards if s.frames]
    for i, s in enumerate(active):
        for f in (s.f0, s.f1 - 1):
            lo, hi = max(0, hop * f - W // 2), min(total, hop * f - W // 2 + W)
            assert lo >= s.span_lo and hi <= s.span_hi
            assert hi <= s.own_hi
            assert lo >= (active[i - 1].own_lo if i else 0)
