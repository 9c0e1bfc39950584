"""The engine's runtime behind the C ABI: concurrent callers, the block cache, the pipelined host-buffer forms.

The reference's conversions are const and re-entrant (its only lock is FFTW's planner mutex, FFTHelper.cpp:9,19), and
they take and return host vectors; these tests hold the B200 build to the same contract."""
import ctypes
import threading

import numpy as np
import pytest

from flan_b200.signals import noise_chirp, sine_sweep
from parity import assert_analysis_parity, assert_synthesis_parity

pytestmark = pytest.mark.gpu
_fp = ctypes.POINTER(ctypes.c_float)


def _ptr(a):
    return a.ctypes.data_as(_fp)


@pytest.fixture(scope="module")
def eng():
    from flan_b200.engine import Engine
    return Engine(0)


@pytest.fixture(scope="module")
def api():
    from flan_b200 import build
    build.build_host()
    L = ctypes.CDLL(build.api_test_path())
    L.api_concurrent_round_trips.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, _fp]
    L.api_repeated_round_trips.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_float, _fp]
    L.api_round_trip.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, _fp]
    return L


def test_four_threads_through_the_cpp_api_match_the_oracle(api, oracle):
    """4 host threads x {convert_to_PV, convert_to_audio} on different objects at the same time."""
    sr, W, h, N = 48000.0, 1024, 64, 1024
    C, n, T = 2, 60000, 4
    x = np.stack([np.stack([noise_chirp(n, sr, 10 * t + c) for c in range(C)]) for t in range(T)])
    F = n // h + 1
    out = np.zeros((T, C, F * h), np.float32)
    assert api.api_concurrent_round_trips(_ptr(x), T, 6, C, n, sr, W, h, N, _ptr(out)) == F * h
    ar = oracle.analysis_rate(sr, h)
    single = np.zeros((C, F * h), np.float32)
    for t in range(T):
        # the same call alone gives the same bits: nothing of another thread's call leaked into this one
        assert api.api_round_trip(_ptr(x[t]), C, n, sr, W, h, N, 0, 0, _ptr(single)) == F * h
        assert np.array_equal(out[t], single)
    # ... and one of them against the oracle, stage-wise on the oracle's own PV is covered elsewhere; here end to end
    ref = oracle.convert_to_audio(oracle.convert_to_pv(x[0], sr, W, h, N), sr, ar, W)
    assert np.abs(out[0] - ref).max() < 5e-3        # independent analysis -> synthesis chains (2.7e-3 .. 4.7e-3)


def test_concurrent_c_abi_calls_on_one_context(eng, oracle):
    """Two Python threads drive the same context through the C ABI (ctypes releases the GIL): different shapes, so the
    workspace sizes differ; every result equals its single-threaded value bit for bit."""
    import torch
    cases = [(48000.0, 1024, 64, 1024, 90000, 3), (44100.0, 2048, 128, 2048, 150000, 4), (48000.0, 512, 32, 512, 40000, 5)]
    inputs = [torch.from_numpy(np.stack([noise_chirp(n, sr, s), noise_chirp(n, sr, s + 1)])).cuda() for sr, W, h, N, n, s in cases]
    expect = []
    for (sr, W, h, N, n, s), x in zip(cases, inputs):
        pv = eng.convert_to_pv(x, sr, W, h, N)
        expect.append((pv.clone(), eng.convert_to_audio(pv, sr, eng.analysis_rate(sr, h), W).clone()))
    torch.cuda.synchronize()
    errors = []

    def work(i):
        try:
            sr, W, h, N, n, s = cases[i]
            for _ in range(20):        # every thread enqueues on the same (default) stream; the context orders whole calls
                pv = eng.convert_to_pv(inputs[i], sr, W, h, N)
                y = eng.convert_to_audio(pv, sr, eng.analysis_rate(sr, h), W)
                torch.cuda.synchronize()
                if not (torch.equal(pv, expect[i][0]) and torch.equal(y, expect[i][1])):
                    errors.append("case %d differs" % i)
                    return
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_block_cache_reuses_allocations(eng):
    ctx, lib = eng.ctx, eng.lib
    a, b = ctypes.c_void_p(), ctypes.c_void_p()
    ctx.call("flan_b200_malloc", 10 << 20, ctypes.byref(a))
    ctx.call("flan_b200_free", a)
    ctx.call("flan_b200_malloc", (10 << 20) - 4096, ctypes.byref(b))        # a similar size: the cached block comes back
    assert a.value == b.value
    ctx.call("flan_b200_free", b)
    ctx.call("flan_b200_trim")
    ctx.call("flan_b200_malloc", 1 << 20, ctypes.byref(a))
    ctx.call("flan_b200_free", a)


@pytest.mark.parametrize("pinned", [False, True])
def test_pipelined_host_forms_are_bit_identical_to_the_device_forms(eng, pinned):
    """flan_b200_convert_to_pv_h2d / _convert_to_audio_d2h slice the work to overlap the copies; slicing must not change
    a bit (pageable source staged through the ring by the copy threads, and page-locked source as one DMA per slice)."""
    import torch
    sr, W, h, N = 48000.0, 1024, 64, 1024
    C, n = 2, 3_000_000                                  # 24 MB of audio: several slices
    x = np.stack([noise_chirp(n, sr, 7), sine_sweep(n, sr)])
    xh = torch.from_numpy(x)
    if pinned:
        xh = xh.pin_memory()
    F, B = eng.num_frames(n, h), N // 2 + 1
    ar = eng.analysis_rate(sr, h)
    pv_ref = eng.convert_to_pv(xh.cuda(), sr, W, h, N)
    y_ref = eng.convert_to_audio(pv_ref, sr, ar, W)
    torch.cuda.synchronize()

    d_audio, d_pv, d_out = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    ctx = eng.ctx
    ctx.call("flan_b200_malloc", 4 * C * n, ctypes.byref(d_audio))
    ctx.call("flan_b200_malloc", 8 * C * F * B, ctypes.byref(d_pv))
    ctx.call("flan_b200_malloc", 4 * C * F * h, ctypes.byref(d_out))
    eng._bind_stream()
    ctx.call("flan_b200_convert_to_pv_h2d", ctypes.c_void_p(xh.data_ptr()), d_audio, C, n, sr, W, h, N, d_pv, None)
    yh = torch.empty((C, F * h), dtype=torch.float32)
    if pinned:
        yh = yh.pin_memory()
    flag = ctypes.c_void_p()
    ctx.call("flan_b200_convert_to_audio_d2h", d_pv, C, F, B, sr, ar, W, d_out, ctypes.c_void_p(yh.data_ptr()), None, ctypes.byref(flag))
    ctx.call("flan_b200_wait", d_out)
    assert ctypes.cast(flag, ctypes.POINTER(ctypes.c_int))[0] == 0
    pv_host = np.empty((C, F, B, 2), np.float32)
    ctx.call("flan_b200_download", ctypes.c_void_p(pv_host.ctypes.data), d_pv, pv_host.nbytes)
    ctx.call("flan_b200_wait", d_pv)
    assert np.array_equal(pv_host, pv_ref.cpu().numpy())
    assert np.array_equal(yh.numpy(), y_ref.cpu().numpy())
    for p in (d_audio, d_pv, d_out):
        ctx.call("flan_b200_free", p)


def test_host_buffer_calls_match_device_calls(eng, oracle):
    """flan_b200_convert_to_pv_host / _convert_to_audio_host (numpy in, numpy out), incl. the NaN/Inf pre-scan flag."""
    import torch
    sr, W, h, N = 44100.0, 2048, 128, 2048
    n = 200000
    x = np.stack([noise_chirp(n, sr, 3), noise_chirp(n, sr, 4)])
    pv = eng.convert_to_pv_host(x, sr, W, h, N)
    pv_dev = eng.convert_to_pv(torch.from_numpy(x).cuda(), sr, W, h, N).cpu().numpy()
    assert np.array_equal(pv, pv_dev)
    ar = eng.analysis_rate(sr, h)
    y, bad = eng.convert_to_audio_host(pv, sr, ar, W)
    assert not bad
    assert np.array_equal(y, eng.convert_to_audio(torch.from_numpy(pv).cuda(), sr, ar, W).cpu().numpy())
    assert_synthesis_parity(y, oracle.convert_to_audio(pv, sr, ar, W))
    pv_bad = pv.copy()
    pv_bad[1, 5, 7, 1] = np.inf
    _, bad = eng.convert_to_audio_host(pv_bad, sr, ar, W)
    assert bad
    # mid/side and left/right variants
    pv_ms = eng.convert_to_pv_host(x, sr, W, h, N, mid_side=True)
    assert_analysis_parity(pv_ms, oracle.convert_to_pv(oracle.mid_side(x), sr, W, h, N), sr, h, N)
    y_lr, _ = eng.convert_to_audio_host(pv, sr, ar, W, left_right=True)
    assert np.array_equal(y_lr, oracle.mid_side(y))


def test_repeated_round_trips_recycle_and_prefetch(api, oracle):
    """A long-lived Audio edited on the host before every pass: from the third pass on its vector is page-locked, the
    result vectors are recycled and the download is prefetched behind the transform. Every pass must still be right."""
    sr, W, h, N = 48000.0, 2048, 128, 2048
    C, n, reps = 2, 400000, 6
    x = np.stack([noise_chirp(n, sr, 21), noise_chirp(n, sr, 22)])
    F = n // h + 1
    out = np.zeros((reps, C, F * h), np.float32)
    step = np.float32(0.125)
    assert api.api_repeated_round_trips(_ptr(x), reps, C, n, sr, W, h, N, step, _ptr(out)) == F * h
    ar = oracle.analysis_rate(sr, h)
    single = np.zeros((C, F * h), np.float32)
    for r in range(reps):
        xr = (x * np.float32(1.0 + 0.125 * r)).astype(np.float32)
        assert api.api_round_trip(_ptr(xr), C, n, sr, W, h, N, 0, 0, _ptr(single)) == F * h
        assert np.array_equal(out[r], single), "pass %d" % r


def test_head_first_range_resynthesis_is_bit_identical(eng, oracle):
    """flan_b200_convert_to_audio_range_head launches the frames that reach into the previous shard first and records the
    caller's event after them (the torchrun path sends its halo from there); the samples are those of the plain call."""
    import torch
    from flan_b200.sharding import frame_shard, head_overlap
    sr, W, h, N = 48000.0, 2048, 128, 2048
    n = 500000
    x = np.stack([noise_chirp(n, sr, 61), sine_sweep(n, sr)])
    pv = eng.convert_to_pv(torch.from_numpy(x).cuda(), sr, W, h, N)
    ar = eng.analysis_rate(sr, h)
    sh = frame_shard(n, h, W, 3, 1)                               # the middle one of three shards
    rows = pv[:, sh.f0:sh.f1].contiguous()
    state = eng.phase_summary(pv[:, :sh.f0].contiguous(), 0, sr, ar, W)
    carry = eng.phase_carry(state.unsqueeze(0), 1)
    lo, hi = sh.span_lo, sh.span_hi
    plain = eng.convert_to_audio_range(rows, sh.f0, sh.frames_total, sr, ar, W, carry, lo, hi - lo)
    ev = torch.cuda.Event()
    ev.record()
    head_first = eng.convert_to_audio_range_head(rows, sh.f0, sh.frames_total, sr, ar, W, carry, lo, hi - lo, ev)
    side = torch.cuda.Stream()
    side.wait_event(ev)
    h_lo, h_hi = head_overlap(sh)
    with torch.cuda.stream(side):
        head = head_first[:, h_lo - lo:h_hi - lo].clone()        # readable as soon as the event has fired
    torch.cuda.synchronize()
    assert torch.equal(plain, head_first)
    assert torch.equal(head, plain[:, h_lo - lo:h_hi - lo])


def test_range_after_summary_only_rewalks_the_scan_bit_identical(eng):
    """A range call that follows flan_b200_phase_summary on the same rows (reuse_summary) skips the phase summary AND the
    first two scan phases: the carry enters at the re-walk (carry (+) carry-free prefix; exact and associative). More
    than 256 segments per shard so that the three-launch scan is the one that runs; head-first form included."""
    import torch
    from flan_b200.sharding import frame_shard
    sr, W, h, N = 48000.0, 1024, 64, 1024
    n = int(sr * 75)
    x = np.stack([noise_chirp(n, sr, 77)])
    xd = torch.from_numpy(x).cuda()
    pv = eng.convert_to_pv(xd, sr, W, h, N)
    ar = eng.analysis_rate(sr, h)
    full = eng.convert_to_audio(pv, sr, ar, W)
    shards = [frame_shard(n, h, W, 3, r) for r in range(3)]
    rows = [pv[:, s.f0:s.f1].contiguous() for s in shards]
    states = torch.stack([eng.phase_summary(r, s.f0, sr, ar, W) for r, s in zip(rows, shards)])
    total = torch.zeros_like(full)
    ev = torch.cuda.Event()
    ev.record()
    for s, r in zip(shards, rows):
        carry = eng.phase_carry(states, s.rank)
        plain = eng.convert_to_audio_range(r, s.f0, s.frames_total, sr, ar, W, carry, s.span_lo, s.span_hi - s.span_lo)
        eng.phase_summary(r, s.f0, sr, ar, W)
        launches = eng.launch_count()
        reused = eng.convert_to_audio_range(r, s.f0, s.frames_total, sr, ar, W, carry, s.span_lo, s.span_hi - s.span_lo,
                                            reuse_summary=True)
        assert eng.launch_count() - launches == 3, "re-walk, output clear, transform"
        assert torch.equal(plain, reused)
        eng.phase_summary(r, s.f0, sr, ar, W)
        head = eng.convert_to_audio_range_head(r, s.f0, s.frames_total, sr, ar, W, carry, s.span_lo, s.span_hi - s.span_lo, ev,
                                               reuse_summary=True)
        assert torch.equal(plain, head)
        total[:, s.span_lo:s.span_hi] += plain
    assert (total - full).abs().max().item() <= 2e-6


@pytest.mark.parametrize("N,hop,seconds", [(4096, 256, 200), (8192, 512, 420), (2048, 128, 100)])
def test_analysis_hint_leaves_the_phase_summaries(eng, N, hop, seconds):
    """flan_b200_hint_resynthesis: the next convert_to_pv also leaves the phase summaries of its rows (one shared-memory word
    per bin; entries it cannot produce are marked and recomputed by the scan), and convert_to_audio(unchanged=True) skips
    pv_phase_seg_kernel. Rows, samples and the NaN / Inf flag are those of the plain calls, bit for bit."""
    import torch
    sr = 48000.0
    n = int(sr * seconds)
    x = np.stack([noise_chirp(n, sr, 3), sine_sweep(n, sr) * np.float32(0.5)])
    xd = torch.from_numpy(x).cuda()
    ar = eng.analysis_rate(sr, hop)
    pv = eng.convert_to_pv(xd, sr, N, hop, N)
    y, flag = eng.convert_to_audio(pv, sr, ar, N, check_nan=True)
    pv_h = eng.convert_to_pv(xd, sr, N, hop, N, for_resynthesis=True)
    assert torch.equal(pv_h, pv)
    n0 = eng.launch_count()
    y_h, flag_h = eng.convert_to_audio(pv_h, sr, ar, N, check_nan=True, unchanged=True)
    used = eng.launch_count() - n0
    assert torch.equal(y_h, y) and flag_h == flag and not flag
    n0 = eng.launch_count()
    eng.convert_to_audio(pv, sr, ar, N)
    assert eng.launch_count() - n0 == used + 1, "the hint saves exactly the phase summary launch"
    # the scan repaired the marked entries in place: a second resynthesis of the same rows reuses them as they are
    pv_h = eng.convert_to_pv(xd, sr, N, hop, N, for_resynthesis=True)
    a = eng.convert_to_audio(pv_h, sr, ar, N, unchanged=True)
    b = eng.convert_to_audio(pv_h, sr, ar, N, unchanged=True)
    assert torch.equal(a, y) and torch.equal(b, y)
    # NaN in the signal: the frames it reaches are marked, recomputed, and raise the flag like the plain path
    xd[1, n // 3] = float("nan")
    pv = eng.convert_to_pv(xd, sr, N, hop, N)
    y, flag = eng.convert_to_audio(pv, sr, ar, N, check_nan=True)
    pv_h = eng.convert_to_pv(xd, sr, N, hop, N, for_resynthesis=True)
    y_h, flag_h = eng.convert_to_audio(pv_h, sr, ar, N, check_nan=True, unchanged=True)
    assert flag and flag_h
    assert torch.equal(y_h.view(torch.int32), y.view(torch.int32))


def test_analysis_hint_is_ignored_where_the_form_does_not_exist(eng):
    import torch
    for (N, W, hop, n) in [(1024, 1024, 64, 48000 * 30), (4096, 2048, 128, 48000 * 60), (4096, 4096, 256, 48000 * 4), (3000, 3000, 100, 200000)]:
        x = torch.from_numpy(np.stack([noise_chirp(n, 48000.0, 8)])).cuda()
        ar = eng.analysis_rate(48000.0, hop)
        pv = eng.convert_to_pv(x, 48000.0, W, hop, N)
        pv_h = eng.convert_to_pv(x, 48000.0, W, hop, N, for_resynthesis=True)
        assert torch.equal(pv, pv_h)
        assert torch.equal(eng.convert_to_audio(pv_h, 48000.0, ar, W, unchanged=True), eng.convert_to_audio(pv, 48000.0, ar, W))
