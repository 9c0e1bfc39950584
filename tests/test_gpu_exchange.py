"""One process per shard on the GPU (the torchrun form of SURVEY 8e) with both exchanges of sharded resynthesis carried by
flan_b200_exchange_*: peer copies into CUDA-IPC mailboxes ordered by sequence flags. Several steps in a row (slot parity,
acknowledgements), three ranks (a rank that both sends and receives). The ranks use device rank % device_count, so the test
runs on a one-GPU box too (IPC between two processes of one device); the assembled samples must have the bits of the same
shards computed by one process (and lie within 2e-6 of the unsharded transform, whose segments are cut elsewhere)."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.gpu

STEPS = 5


def _signal(n, sr, step):
    sys.path.insert(0, ROOT)
    from flan_b200.signals import noise_chirp, sine_sweep
    g = np.float32(1.0 - 0.125 * step)
    return np.stack([noise_chirp(n, sr, 21) * g, sine_sweep(n, sr) * g]).astype(np.float32)


def _worker(rank, world, port, outdir, shape):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flan_b200.engine import Engine
    from flan_b200.sharding import PeerExchange, frame_shard, sharded_resynthesis_peer
    sr, W, hop, N, n = shape
    device = rank % torch.cuda.device_count()
    eng = Engine(device)
    sh = frame_shard(n, hop, W, world, rank)
    ar = eng.analysis_rate(sr, hop)
    ex = PeerExchange(eng, dist, rank, world, 2, N // 2 + 1, max(0, W - hop))
    ev = torch.cuda.Event()
    ev.record()
    owned = []
    for step in range(STEPS):
        x = _signal(n, sr, step)
        local = torch.from_numpy(np.ascontiguousarray(x[:, sh.audio_lo:sh.audio_hi])).to(eng.device)
        # odd steps say that the rows will be resynthesised unchanged: the analysis then leaves their phase summaries where
        # that form exists (the first shape: > 256 segments of a full-window dft 4096), and nothing changes elsewhere
        hinted = step % 2 == 1
        pv = eng.convert_to_pv_range(local, sh.audio_lo, n, sr, W, hop, N, sh.f0, sh.f1, for_resynthesis=hinted)
        out, lo = sharded_resynthesis_peer(eng, ex, torch, sh, pv, sr, ar, ev, unchanged=hinted)
        owned.append(out[:, sh.own_lo - lo:sh.own_hi - lo].clone())      # no host synchronisation between the steps
    torch.cuda.synchronize()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), own=torch.stack(owned).cpu().numpy(), own_lo=sh.own_lo, own_hi=sh.own_hi)
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,shape", [(2, (48000.0, 4096, 256, 4096, 48000 * 150)),     # mirrored kernels; shards of > 256 segments
                                         (3, (48000.0, 1024, 64, 1024, 48000 * 75)),       # > 256 segments per shard: the re-walk-only scan
                                         (3, (44100.0, 2048, 128, 4096, 300000))])         # the API default shape
def test_peer_exchange_round_trip_is_bit_identical(world, shape):
    import torch
    import torch.multiprocessing as mp
    from flan_b200.engine import Engine
    sr, W, hop, N, n = shape
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, shape), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
    from flan_b200.sharding import frame_shard
    eng = Engine(0)
    ar = eng.analysis_rate(sr, hop)
    shards = [frame_shard(n, hop, W, world, r) for r in range(world)]
    for step in range(STEPS):
        x = torch.from_numpy(_signal(n, sr, step)).cuda()
        pv = eng.convert_to_pv(x, sr, W, hop, N)
        full = eng.convert_to_audio(pv, sr, ar, W)
        # the same shards in one process: range calls, partial sums added in rank order
        rows = [pv[:, s.f0:s.f1].contiguous() for s in shards]
        states = torch.stack([eng.phase_summary(r, s.f0, sr, ar, W) for r, s in zip(rows, shards)])
        total = torch.zeros_like(full)
        for s, r in zip(shards, rows):
            carry = eng.phase_carry(states, s.rank) if s.rank else None
            total[:, s.span_lo:s.span_hi] += eng.convert_to_audio_range(r, s.f0, s.frames_total, sr, ar, W, carry, s.span_lo, s.span_hi - s.span_lo)
        assert (total - full).abs().max().item() <= 2e-6
        total = total.cpu().numpy()
        pos = 0
        for p in parts:
            lo, hi = int(p["own_lo"]), int(p["own_hi"])
            assert lo == pos
            assert np.array_equal(p["own"][step], total[:, lo:hi]), "step %d, samples [%d, %d)" % (step, lo, hi)
            pos = hi
        assert pos == total.shape[1]
