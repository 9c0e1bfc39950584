"""Throughput of the BASELINE.json configs that bench.py's single line does not cover, at 1..8 GPUs of one node:

    python tools/configbench.py --config cfg3|cfg4|cfg5 [--steps 3]                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/configbench.py --config cfgX                                                # N GPUs

  cfg3  mono 96 kHz 1 h, window 8192 hop 512: round trip, FRAME-sharded across the ranks (strong scaling), resynthesis
        exchanges the per-bin phase state (all_gather) and the window-hop overlap-add halo (send/recv) over NCCL
  cfg4  8 channels 48 kHz 30 min, window 2048 hop 128: analysis -> PV::repitch(1.5) -> PV::stretch(2.0) -> resynthesis,
        CHANNEL-sharded, no communication
  cfg5  256 clips of 60 s mono 48 kHz, window 1024 hop 64: round trip, FILE-sharded, no communication (clips are
        batched 32 at a time as the channels of one call)
Signals are generated on the device (seeded noise + chirp). One JSON line (rank 0): frames/s of the whole job =
input frames of all ranks / max-over-ranks device time (CUDA events), and the fraction of the HBM roofline that the
job's compulsory traffic represents."""
import argparse
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flan_b200.engine import Engine  # noqa: E402
from flan_b200.sharding import PeerExchange, frame_shard, sharded_resynthesis, sharded_resynthesis_peer  # noqa: E402


def signal(n, sr, seed, dev, offset=0, n_total=None):
    """0.25 * U(-1,1) + 0.5 * chirp 50 Hz -> 0.45 sr over the whole signal, samples [offset, offset + n)."""
    n_total = n_total or n
    g = torch.Generator(device=dev).manual_seed(seed)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    step = 1 << 24
    dur = n_total / sr
    for s in range(0, n, step):
        e = min(n, s + step)
        t = (torch.arange(s + offset, e + offset, dtype=torch.float64, device=dev)) / sr
        phase = 2.0 * math.pi * (50.0 * t + (0.45 * sr - 50.0) * t * t / (2.0 * dur))
        u = torch.rand(e - s, generator=g, dtype=torch.float32, device=dev) * 2.0 - 1.0
        out[s:e] = 0.25 * u + 0.5 * torch.sin(phase).float()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--scale", type=float, default=1.0, help="shorten the signals (debug)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local)
    peak = 6553.9
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    if args.config == "cfg3":
        sr, W, hop, N = 96000.0, 8192, 512, 8192
        n_total = int(sr * 3600 * args.scale)
        B = N // 2 + 1
        ar = eng.analysis_rate(sr, hop)
        sh = frame_shard(n_total, hop, W, world, rank)
        x = signal(sh.audio_hi - sh.audio_lo, sr, 3, dev, sh.audio_lo, n_total).unsqueeze(0)
        pv = torch.empty((1, sh.frames, B, 2), dtype=torch.float32, device=dev)

        def allgather(state):
            if world == 1:
                return state.unsqueeze(0)
            bufs = torch.empty((world,) + tuple(state.shape), dtype=state.dtype, device=dev)
            dist.all_gather_into_tensor(bufs, state.contiguous())
            return bufs

        # phase state + halo between the ranks: peer copies into CUDA-IPC mailboxes (flan_b200_exchange_*), NCCL if that
        # cannot be set up
        exchange = None
        if world > 1:
            try:
                exchange = PeerExchange(eng, dist, rank, world, 1, B, max(0, W - hop))
            except RuntimeError as e:
                if rank == 0:
                    print("configbench: %s; using NCCL" % e, file=sys.stderr)
        head_event = torch.cuda.Event()
        head_event.record()

        def step():
            eng.convert_to_pv_range(x, sh.audio_lo, n_total, sr, W, hop, N, sh.f0, sh.f1, out=pv)
            if world == 1:
                return eng.convert_to_audio(pv, sr, ar, W)
            if exchange is not None:
                o, _ = sharded_resynthesis_peer(eng, exchange, torch, sh, pv, sr, ar, head_event)
                return o
            o, _ = sharded_resynthesis(eng, dist, sh, pv, sr, ar, allgather,
                                       lambda t, dst: dist.isend(t, dst), lambda t, src: dist.recv(t, src))
            return o
        frames_rank = sh.frames
        bytes_rank = 2.0 * sh.frames * (4.0 * hop + 8.0 * B)
        what = "mono 96 kHz 1 h, window 8192 hop 512, round trip, frame-sharded (strong scaling)"
    elif args.config == "cfg4":
        sr, W, hop, N = 48000.0, 2048, 128, 2048
        n = int(sr * 1800 * args.scale)
        B = N // 2 + 1
        ar = eng.analysis_rate(sr, hop)
        chans = [c for c in range(8) if c % world == rank]
        xs = [signal(n, sr, 40 + c, dev).unsqueeze(0) for c in chans]
        F = eng.num_frames(n, hop)
        pv = torch.empty((1, F, B, 2), dtype=torch.float32, device=dev)
        rp = torch.empty_like(pv)
        out_frames = [0]

        def step():
            y = None
            for x in xs:
                eng.convert_to_pv(x, sr, W, hop, N, out=pv)
                eng.repitch(pv, sr, 1.5, 0, out=rp)
                st = eng.stretch(rp, sr, ar, 2.0, 0, summary_window=W)      # leaves the phase summaries of its rows
                out_frames[0] = st.shape[1]
                y = eng.convert_to_audio(st, sr, ar, W, unchanged=True)
                del st
            return y
        frames_rank = F * len(chans)
        F2 = 2 * F
        bytes_rank = len(chans) * ((4.0 * hop + 8.0 * B) * F + 16.0 * B * F + (8.0 * B * F + 8.0 * B * F2) + (8.0 * B + 4.0 * hop) * F2)
        what = "8 ch 48 kHz 30 min, window 2048 hop 128, analysis -> repitch(1.5) -> stretch(2.0) -> resynthesis, channel-sharded"
    else:
        sr, W, hop, N = 48000.0, 1024, 64, 1024
        n = int(sr * 60 * args.scale)
        B = N // 2 + 1
        ar = eng.analysis_rate(sr, hop)
        clips = [i for i in range(256) if i % world == rank]
        batch = 32
        F = eng.num_frames(n, hop)
        # one resident batch of 32 distinct clips stands for every batch of this rank (inputs in HBM, as bench.py)
        x = torch.stack([signal(n, sr, 5000 + i, dev) for i in clips[:batch]])
        pv = torch.empty((x.shape[0], F, B, 2), dtype=torch.float32, device=dev)
        y = torch.empty((x.shape[0], F * hop), dtype=torch.float32, device=dev)
        nb = (len(clips) + batch - 1) // batch

        def step():
            for _ in range(nb):
                eng.convert_to_pv(x, sr, W, hop, N, out=pv)
                eng.convert_to_audio(pv, sr, ar, W, out=y)
            return y
        frames_rank = F * x.shape[0] * nb
        bytes_rank = 2.0 * frames_rank * (4.0 * hop + 8.0 * B)
        what = "256 clips x 60 s mono 48 kHz, window 1024 hop 64, round trip, file-sharded (32 clips per call)"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms, float(frames_rank), bytes_rank], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(mx[0])
    frames, byts = float(t[1]), float(t[2])
    if rank == 0:
        print(json.dumps({"config": args.config, "workload": what, "n_gpus": world, "steps": args.steps, "ms_per_step": ms,
                          "frames": frames, "value": frames / (ms * 1e-3), "unit": "input frames/s",
                          "hbm_gbs_all_gpus": byts / (ms * 1e-3) / 1e9,
                          "frac_of_hbm_roofline_per_gpu": byts / (ms * 1e-3) / 1e9 / (peak * world), "scale": args.scale}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
