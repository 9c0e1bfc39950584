"""Frame-range sharding of one signal across ranks (BASELINE.json config 3).

Pure index arithmetic plus the exchange schedule; the transforms themselves are the C ABI's *_range
entry points (or, in the CPU tests, the thread emulator). One process per GPU; torch.distributed carries
  * analysis: nothing, if each rank holds its frame range's samples plus a halo of W/2 + hop on the left
    (the extra hop lets it recompute the phase of frame f0-1) and W/2 on the right;
  * resynthesis: an all_gather of the per-bin phase state (C x B x 32 bytes per rank) and one
    send/recv per shard boundary of the W - hop overlap-add samples that two shards share.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class FrameShard:
    rank: int
    world: int
    n: int              # samples per channel of the whole signal
    hop: int
    W: int
    frames_total: int   # F = n // hop + 1                     (AudioPV.cpp:17)
    f0: int             # first frame of this rank
    f1: int             # one past the last frame
    audio_lo: int       # samples [audio_lo, audio_hi) must be resident for analysis
    audio_hi: int
    span_lo: int        # output samples [span_lo, span_hi) receive contributions from this rank's frames
    span_hi: int
    own_lo: int         # output samples [own_lo, own_hi) are final on this rank after the halo exchange
    own_hi: int

    @property
    def frames(self):
        return self.f1 - self.f0

    @property
    def out_total(self):
        return self.frames_total * self.hop          # AudioPV.cpp:93


def frame_shard(n, hop, W, world, rank):
    F = n // hop + 1
    # A frame's window reaches window - hop samples into the previous shard and the halo exchange is with the adjacent
    # rank only, so a shard must hold at least ceil(window / hop) frames: short signals use fewer ranks (the others get
    # an empty range and take part in the collectives with zero state and no halo).
    per = max((F + world - 1) // world, (W + hop - 1) // hop)
    f0 = min(F, rank * per)
    f1 = min(F, f0 + per)
    half = W // 2
    total = F * hop
    if f1 > f0:
        first = f0 - 1 if f0 > 0 else 0
        audio_lo = max(0, hop * first - half)
        audio_hi = min(n, hop * (f1 - 1) - half + W)
        audio_hi = max(audio_hi, audio_lo)
        span_lo = max(0, hop * f0 - half)
        span_hi = min(total, hop * (f1 - 1) - half + W)
        own_lo = 0 if f0 == 0 else min(total, hop * f0 - half + (W - hop))
        own_hi = total if f1 == F else min(total, hop * f1 - half + (W - hop))
    else:
        audio_lo = audio_hi = span_lo = span_hi = own_lo = own_hi = 0
    return FrameShard(rank, world, n, hop, W, F, f0, f1, audio_lo, audio_hi, span_lo, span_hi, own_lo, own_hi)


def head_overlap(shard):
    """Samples [lo, hi) of this rank's span that belong to the previous rank (partial sums to send left)."""
    if shard.f0 == 0 or shard.frames == 0:
        return (0, 0)
    return (shard.span_lo, min(shard.own_lo, shard.span_hi))


def sharded_resynthesis(engine, dist, shard, pv_rows, sr, ar, allgather, send, recv):
    """Resynthesise this rank's frames and finish the samples it owns.

    engine: flan_b200.engine.Engine (or the emulator adapter of the CPU tests) -- same method names.
    allgather(state) -> stacked [world, C, B, 4]; send(tensor, dst) / recv(tensor, src) move one halo.
    Returns (local_out, lo) where local_out[:, own_lo-lo : own_hi-lo] is final.
    """
    state = engine.phase_summary(pv_rows, shard.f0, sr, ar, shard.W)
    carry = engine.phase_carry(allgather(state), shard.rank)
    lo, hi = shard.span_lo, shard.span_hi
    # the rows are untouched since phase_summary: its per-segment summaries are reused, the PV is read once more only
    out = engine.convert_to_audio_range(pv_rows, shard.f0, shard.frames_total, sr, ar, shard.W, carry, lo, hi - lo,
                                        reuse_summary=True)
    # lower-frame contributions first (AudioPV.cpp:133-134): the owner adds the right neighbour's partial sums
    h_lo, h_hi = head_overlap(shard)
    reqs = []
    if h_hi > h_lo and shard.rank > 0:
        reqs.append(send(out[:, h_lo - lo:h_hi - lo].contiguous(), shard.rank - 1))
    if shard.rank + 1 < shard.world:
        nxt = frame_shard(shard.n, shard.hop, shard.W, shard.world, shard.rank + 1)
        n_lo, n_hi = head_overlap(nxt)
        if n_hi > n_lo:
            buf = engine.empty_like_audio(out.shape[0], n_hi - n_lo)
            recv(buf, shard.rank + 1)
            tail = out[:, n_lo - lo:n_hi - lo]
            engine.add_into(tail, buf)
    for r in reqs:
        if r is not None:
            r.wait()
    return out, lo


def sharded_resynthesis_overlapped(engine, dist, torch, shard, pv_rows, sr, ar, allgather, side_stream, head_event):
    """sharded_resynthesis with the halo exchange off the critical path (GPU ranks; SURVEY 8e: "boundary tiles first,
    then exchange on a side stream"): the frames whose windows reach into the previous rank are launched first
    (flan_b200_convert_to_audio_range_head records head_event after them), their partial sums leave on side_stream in
    one batched NCCL send / recv with the right neighbour's while the remaining frames compute, and the owner adds what
    it received after its own frames (lower-frame contributions first, AudioPV.cpp:133-134).
    head_event: a torch.cuda.Event that has been recorded once (its handle must exist)."""
    state = engine.phase_summary(pv_rows, shard.f0, sr, ar, shard.W)
    carry = engine.phase_carry(allgather(state), shard.rank)
    lo, hi = shard.span_lo, shard.span_hi
    h_lo, h_hi = head_overlap(shard)
    sends = h_hi > h_lo and shard.rank > 0
    out = engine.convert_to_audio_range_head(pv_rows, shard.f0, shard.frames_total, sr, ar, shard.W, carry, lo, hi - lo,
                                             head_event if sends else None, reuse_summary=True)
    main = torch.cuda.current_stream()
    n_lo = n_hi = 0
    if shard.rank + 1 < shard.world:
        n_lo, n_hi = head_overlap(frame_shard(shard.n, shard.hop, shard.W, shard.world, shard.rank + 1))
    buf = None
    with torch.cuda.stream(side_stream):
        ops = []
        if sends:
            side_stream.wait_event(head_event)
            head = out[:, h_lo - lo:h_hi - lo].contiguous()
            ops.append(dist.P2POp(dist.isend, head, shard.rank - 1))
        if n_hi > n_lo:
            buf = torch.empty((out.shape[0], n_hi - n_lo), dtype=out.dtype, device=out.device)
            ops.append(dist.P2POp(dist.irecv, buf, shard.rank + 1))
        for r in (dist.batch_isend_irecv(ops) if ops else []):
            r.wait()
    if buf is not None:
        main.wait_stream(side_stream)
        buf.record_stream(main)
        engine.add_into(out[:, n_lo - lo:n_hi - lo], buf)
    elif sends:
        main.wait_stream(side_stream)      # `out` must outlive the send that reads it
    return out, lo


class PeerExchange:
    """flan_b200_exchange_* (include/flan_b200.h): the phase states and overlap-add halos of frame-range shards travel
    between the per-GPU processes as device-to-device copies into CUDA-IPC mailboxes, ordered by sequence flags; no
    exchange kernel runs on an SM. `dist` is only used once, to all-gather the 64-byte mailbox handles."""

    def __init__(self, engine, dist, rank, world, channels, bins, halo_samples):
        """Collective: every rank calls it. If the mailbox cannot be set up on ANY rank (no peer access, CUDA IPC or
        cuStreamWaitValue32 unavailable) every rank raises RuntimeError, so that callers can fall back together."""
        import ctypes
        self.engine, self.rank, self.world, self.channels, self.bins = engine, rank, world, channels, bins
        self.lib, self.ctx = engine.lib, engine.ctx
        self.h = None
        mine, error = None, None
        try:
            h = ctypes.c_void_p()
            self.ctx.check(self.lib.flan_b200_exchange_create(self.ctx.h, rank, world, channels, bins, halo_samples, ctypes.byref(h)))
            self.h = h
            buf = ctypes.create_string_buffer(64)
            self._call("flan_b200_exchange_handle", buf)
            mine = bytes(buf.raw)
        except Exception as e:  # noqa: BLE001 - reported to every rank below
            error = "rank %d: %r" % (rank, e)
        handles = [None] * world
        dist.all_gather_object(handles, (mine, error))
        errors = [e for _, e in handles if e]
        if not errors:
            try:
                self._call("flan_b200_exchange_connect", ctypes.create_string_buffer(b"".join(h for h, _ in handles), 64 * world))
            except Exception as e:  # noqa: BLE001
                error = "rank %d: %r" % (rank, e)
            connected = [None] * world
            dist.all_gather_object(connected, error)
            errors = [e for e in connected if e]
        if errors:
            self.close()
            raise RuntimeError("peer exchange unavailable (%s)" % "; ".join(errors))

    def _call(self, name, *args):
        self.ctx.check(getattr(self.lib, name)(self.h, *args))

    def close(self):
        if getattr(self, "h", None):
            self.lib.flan_b200_exchange_destroy(self.h)
            self.h = None


def sharded_resynthesis_peer(engine, ex, torch, shard, pv_rows, sr, ar, head_event, unchanged=False):
    """sharded_resynthesis with both exchanges as peer copies (PeerExchange), the halo off the critical path: the frames
    whose windows reach into the previous rank run beside the rest of the shard (flan_b200_convert_to_audio_range_head
    records head_event after them), their partial sums are pushed into the previous rank's mailbox while both ranks
    compute, and the owner adds what it received after its own frames (lower-frame contributions first, AudioPV.cpp:133-134).
    head_event: a torch.cuda.Event that has been recorded once (its handle must exist).
    unchanged: pv_rows are exactly what convert_to_pv_range(..., for_resynthesis=True) wrote: the phase summaries it left
    are used (flan_b200_promise_unchanged) instead of a second read of the rows."""
    import ctypes
    C, rows, B, _ = pv_rows.shape
    engine._bind_stream()
    d_state = ctypes.c_void_p()
    ex._call("flan_b200_exchange_state_slot", ctypes.byref(d_state))
    if unchanged:
        engine.ctx.call("flan_b200_promise_unchanged", engine._chk(pv_rows))
    engine.ctx.call("flan_b200_phase_summary", engine._chk(pv_rows), rows * B, C, shard.f0, shard.f0 + rows, B, sr, ar, shard.W, d_state)
    ex._call("flan_b200_exchange_put_state", d_state)
    carry = None
    if shard.rank > 0:
        d_all = ctypes.c_void_p()
        ex._call("flan_b200_exchange_get_states", ctypes.byref(d_all))
        carry = torch.empty((C, B, 4), dtype=torch.float64, device=pv_rows.device)
        engine.ctx.call("flan_b200_phase_carry", d_all, shard.rank, C, B, engine._chk(carry, torch.float64))
        ex._call("flan_b200_exchange_release_states")
    lo, hi = shard.span_lo, shard.span_hi
    h_lo, h_hi = head_overlap(shard)
    sends = h_hi > h_lo and shard.rank > 0
    out = engine.convert_to_audio_range_head(pv_rows, shard.f0, shard.frames_total, sr, ar, shard.W, carry, lo, hi - lo,
                                             head_event if sends else None, reuse_summary=True)
    if sends:
        ex._call("flan_b200_exchange_put_halo", ctypes.c_void_p(out.data_ptr() + 4 * (h_lo - lo)), hi - lo, C, h_hi - h_lo,
                 ctypes.c_void_p(head_event.cuda_event))
    if shard.rank + 1 < shard.world:
        n_lo, n_hi = head_overlap(frame_shard(shard.n, shard.hop, shard.W, shard.world, shard.rank + 1))
        if n_hi > n_lo:
            ex._call("flan_b200_exchange_add_halo", ctypes.c_void_p(out.data_ptr() + 4 * (n_lo - lo)), hi - lo, C, n_hi - n_lo)
    return out, lo
