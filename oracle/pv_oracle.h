/* oracle/pv_oracle.h: TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar, one thread) of the reference's phase-vocoder hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library;
 * the product path (flan_b200/csrc) never calls it.
 *
 * Pinning: the reference holds no golden vectors for this path (tests/flanTest.cpp is a scratch
 * main). This restatement is pinned against the reference ITSELF: oracle/_ref/libflan_ref.so is
 * the reference's own Conversions/AudioPV.cpp, phase_vocoder.cpp, WindowFunctions.cpp, FFTHelper.cpp
 * and PV/PVBuffer.cpp compiled verbatim, and tests/test_oracle_vs_ref.py requires bit-identical
 * output from both on the same inputs (same double-precision FFT stand-in for the absent FFTW).
 * Fixtures generated from that build are committed under tests/golden/.
 */
#ifndef PV_ORACLE_H
#define PV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* AudioPV.cpp:17 -- ceil of an already truncated integer quotient, + 1. */
int64_t pvo_num_frames( int64_t n, int hop );

/* WindowFunctions.cpp:10-13 sampled as AudioPV.cpp:30-34. */
void pvo_hann( int window_size, float * out );

/* Audio::convert_to_PV, AudioPV.cpp:12-78 + phase_vocoder.cpp:37-52.
 * audio: planar float[C][n]; pv_out: interleaved (m,f) float pairs [C][F][dft/2+1].
 * Frames [frame_begin, frame_end) of each channel are produced (pass 0, F for all); pv_out is
 * indexed relative to frame_begin ([C][frame_end-frame_begin][B]). The phase carried into
 * frame_begin is recomputed from frame_begin-1, which is what the serial reference loop holds.
 * Returns 0, or -1 on bad arguments. */
int pvo_convert_to_pv( const float * audio, int C, int64_t n, float sample_rate,
                       int window_size, int hop, int dft_size,
                       int64_t frame_begin, int64_t frame_end, float * pv_out );

/* PV::convert_to_audio, AudioPV.cpp:86-139 + phase_vocoder.cpp:55-61.
 * pv: [C][F][B] (m,f) pairs; audio_out: planar float[C][F*hop], zero-filled here. */
int pvo_convert_to_audio( const float * pv, int C, int64_t F, int B, float sample_rate,
                          float analysis_rate, int window_size, float * audio_out );

/* Audio::convert_to_mid_side, AudioConversions.cpp:32-51 (stereo only). in/out planar float[2][n]. */
void pvo_mid_side( const float * in, int64_t n, float * out );

/* PVBuffer::get_hop_size, PVBuffer.cpp:381-384. */
int pvo_hop_from_rates( float sample_rate, float analysis_rate );

/* ---- PV-domain chain: PV::repitch / PV::stretch / PV::modify_time, PV/PVModify.cpp:196-385 ----
 * factor / mod tables are float[F][B], the reference's sample of its Function argument over the frame x bin grid
 * (PV/PV.h:31-35). interp: 0 linear, 1 midpoint, 2 nearest, 3 floor, 4 ceil, 5 smoothstep, 6 smootherstep, 7 sine,
 * 8 sine2, 9 sqrt (Utility/Interpolator.cpp). pv / out: interleaved (m,f) pairs. */
int pvo_repitch( const float * pv, int C, int64_t F, int B, float sample_rate, const float * factor, int interp, float * out );
/* Both return the output frame count; out == NULL computes only that, otherwise out is [C][count][B]. */
int64_t pvo_stretch( const float * pv, int C, int64_t F, int B, float sample_rate, float analysis_rate,
                     const float * factor, int interp, float * out );
int64_t pvo_modify_time( const float * pv, int C, int64_t F, int B, float sample_rate, float analysis_rate,
                         const float * mod_seconds, int interp, float * out );

/* ---- file formats either side of the path ----
 * .flan RIFF-PV sample codec, PV/PVBuffer.cpp:99-127 (save) and :253-268 (load); count = MF elements, bytes = 6 * count. */
void pvo_flan_encode( const float * pv, int64_t count, float dft_size, float sample_rate, uint8_t * bytes );
void pvo_flan_decode( const uint8_t * bytes, int64_t count, float dft_size, float sample_rate, float * pv );
/* WAV PCM-24 as Audio/AudioBuffer.cpp:136-170 writes it through libsndfile (clamp, interleave, lrintf( x * 0x7FFFFF ))
 * and :112-125 reads it (value / 2^23, de-interleave). libsndfile's conversion restated: parity unpinned. */
void pvo_pcm24_encode( const float * audio, int C, int64_t n, uint8_t * bytes );
void pvo_pcm24_decode( const uint8_t * bytes, int C, int64_t n, float * audio );

#ifdef __cplusplus
}
#endif
#endif
