"""Parity gates between an engine result and the oracle's (BASELINE.json north_star).

Analysis (per bin, gated to bins with m_ref >= 1e-2 * max_b m_ref of the frame, in this frame and in the
previous one whose phase enters the difference -- below that the phase of a bin is rounding noise of the FFT
and its frequency is arbitrary within +-analysis_rate/2):
    |m - m_ref| <= 1e-4 * m_ref
    |f - f_ref| (mod analysis_rate) <= max(1e-3 Hz, 2 ulp32(f_ref), 2 grid(b))
where grid(b) = ulp32(expected_phase_diff[b]) * analysis_rate / 2pi is the spacing of representable values
of `phase_diff - expected_phase_diff` (phase_vocoder.cpp:48): two correct float FFTs land on neighbouring
grid points, so the reference itself cannot be pinned more tightly than that. For every BASELINE config
(hop = dft/16) 2 grid(b) <= 2 ulp32(f) and the gate is the north_star's 1e-3 Hz or 2 ulp, whichever is larger.
A phase error e maps to e * analysis_rate / 2pi Hz, so for analysis rates above BASELINE's largest (750 Hz, cfg5;
only the small test fixtures go there) the 1e-3 Hz term is scaled by analysis_rate / 750 -- same phase tolerance.

On top of that the gate allows the float32 FFT noise floor itself: 1e-7 * (1/rho_f + 1/rho_{f-1}) rad, rho = m_ref /
frame peak >= 1e-2, i.e. at most 6e-4 Hz at the gate edge for analysis_rate 187.5 Hz and ~1e-5 Hz for loud bins.
(The Bluestein path of the non-power-of-two sizes meets the same gate: magnitudes within 1.2e-5, no bin needs the noise
term in the shapes tested.)

Resynthesis, stage-wise on the SAME PV input: max |sample - sample_ref| <= 1e-5.
"""
import numpy as np


def ulp32(x):
    x = np.abs(np.asarray(x, np.float32))
    return np.spacing(np.maximum(x, np.float32(1e-30))).astype(np.float64)


def analysis_report(pv, pv_ref, sr, hop, N, fft_noise=1e-7, first_frame_has_history=False):
    pv = np.asarray(pv, np.float32)
    pv_ref = np.asarray(pv_ref, np.float32)
    m, f = pv[..., 0].astype(np.float64), pv[..., 1].astype(np.float64)
    mr, fr = pv_ref[..., 0].astype(np.float64), pv_ref[..., 1].astype(np.float64)
    ar = float(np.float32(sr) / np.float32(hop))
    pi2 = float(np.float32(np.float32(np.arccos(np.float32(-1))) * np.float32(2)))
    B = N // 2 + 1
    binf = (np.arange(B, dtype=np.float32) * np.float32(sr) / np.float32((B - 1) * 2)).astype(np.float32)   # PVBuffer.cpp:443-446 over get_dft_size()
    expected = (binf / np.float32(ar) * np.float32(pi2)).astype(np.float32)
    grid = ulp32(expected) * ar / (2 * np.pi)
    peak = mr.max(axis=-1, keepdims=True)
    gate = (mr >= 1e-2 * peak) & (peak > 0)
    # f is built from this frame's AND the previous frame's phase (phase_vocoder.cpp:44): both must be above the floor
    gate[..., 1:, :] &= gate[..., :-1, :].copy()
    if first_frame_has_history:
        # the arrays are a window of frames out of a longer signal: the first frame's phase difference involves a frame
        # that is not here, whose level cannot be checked -- it is left out
        gate[..., 0, :] = False
    df = np.abs(f - fr)
    df = np.minimum(df, np.abs(df - ar))            # +-pi wrap ambiguity: f is defined mod analysis_rate
    tol_f = np.maximum(1e-3 * max(1.0, ar / 750.0), np.maximum(2 * ulp32(pv_ref[..., 1]), 2 * grid))
    # float32 FFT noise: an error of 1e-7 * frame peak on a bin of relative level rho moves its phase by 1e-7 / rho
    rho = np.where(peak > 0, mr / np.where(peak > 0, peak, 1.0), 1.0)
    inv = 1.0 / np.maximum(rho, 1e-2)
    inv_prev = inv.copy()
    inv_prev[..., 1:, :] = inv[..., :-1, :]
    inv_prev[..., 0, :] = 0.0
    tol_f = tol_f + fft_noise * (inv + inv_prev) * ar / (2 * np.pi)
    rel_m = np.abs(m - mr) / np.where(mr > 0, mr, 1.0)
    bad_m = gate & (rel_m > 1e-4)
    bad_f = gate & (df > tol_f)
    # SURVEY 8c's gate as proposed, WITHOUT the three allowances argued above (VERDICT r1 weak #1 asked how many bins
    # pass only because of them): current-frame level only, 1e-3 Hz unscaled, no noise term. Reported, not asserted:
    # a float32 FFT other than the oracle's double one cannot meet it on bins near the gate edge.
    gate0 = (mr >= 1e-2 * peak) & (peak > 0)
    tol0 = np.maximum(1e-3, 2 * ulp32(pv_ref[..., 1]))
    strict_fail = gate0 & (df > tol0)
    tol_scale = np.maximum(1e-3 * max(1.0, ar / 750.0), np.maximum(2 * ulp32(pv_ref[..., 1]), 2 * grid))
    need_prev = strict_fail & ~gate                                  # excused by the previous frame's level
    need_scale = strict_fail & gate & (df <= tol_scale)              # excused by the ar/750 scale or the grid term
    need_noise = strict_fail & gate & (df > tol_scale) & (df <= tol_f)   # excused only by the FFT noise term
    return {
        "gated_bins": int(gate.sum()),
        "max_rel_m": float(rel_m[gate].max()) if gate.any() else 0.0,
        "max_df_hz": float(df[gate].max()) if gate.any() else 0.0,
        "bad_m": int(bad_m.sum()),
        "bad_f": int(bad_f.sum()),
        "frac_f_bit_exact": float(np.mean(pv[..., 1] == pv_ref[..., 1])),
        "frac_m_bit_exact": float(np.mean(pv[..., 0] == pv_ref[..., 0])),
        "frac_f_bit_exact_gated": float(np.mean((pv[..., 1] == pv_ref[..., 1])[gate])) if gate.any() else 1.0,
        "frac_f_within_1ulp_gated": float(np.mean((df <= ulp32(pv_ref[..., 1]))[gate])) if gate.any() else 1.0,
        "survey_gate_bins": int(gate0.sum()),
        "survey_gate_fail": int(strict_fail.sum()),
        "pass_only_by_prev_frame_gate": int(need_prev.sum()),
        "pass_only_by_rate_scale_or_grid": int(need_scale.sum()),
        "pass_only_by_noise_term": int(need_noise.sum()),
        "nan": int(np.isnan(pv).sum()),
    }


def assert_analysis_parity(pv, pv_ref, sr, hop, N, min_f_within_1ulp=None, fft_noise=1e-7, first_frame_has_history=False):
    r = analysis_report(pv, pv_ref, sr, hop, N, fft_noise, first_frame_has_history)
    assert r["nan"] == 0, r
    assert r["gated_bins"] > 0, r
    assert r["bad_m"] == 0, r
    assert r["bad_f"] == 0, r
    if min_f_within_1ulp is not None:       # a floor on the tight statistics: an epilogue regression shows here first
        assert r["frac_f_within_1ulp_gated"] >= min_f_within_1ulp, r
    return r


def assert_synthesis_parity(audio, audio_ref, tol=1e-5):
    audio = np.asarray(audio, np.float32)
    audio_ref = np.asarray(audio_ref, np.float32)
    assert audio.shape == audio_ref.shape, (audio.shape, audio_ref.shape)
    assert not np.isnan(audio).any()
    err = float(np.abs(audio.astype(np.float64) - audio_ref.astype(np.float64)).max()) if audio.size else 0.0
    assert err <= tol, err
    return err
