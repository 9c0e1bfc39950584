// flan_b200/csrc/pv_tables.h -- host-side constant tables of one (dft, window, hop, rates) plan.
//
// Window, per-bin constants and window_scale are evaluated on the HOST with the reference's own
// expressions (g++ semantics, -ffp-contract=off) and uploaded, so no libm / compiler divergence can
// leak into device code. Shared by the C ABI (pv_capi.cu) and the CPU
// thread emulator (emu/pv_emu.cpp).
#pragma once

#include <cmath>
#include <cstdint>
#include <vector>

#include "pv_body.cuh"

namespace pvk {

struct HostTables
	{
	int N = 0, W = 0, hop = 0;
	std::vector<float> win_analysis;    // Windows::hann( i / (W-1) )                         AudioPV.cpp:30-34
	std::vector<float> win_synthesis;   // hann * window_scale                               AudioPV.cpp:98-103
	std::vector<float> expected;        // bin_frequency / analysis_rate * pi2, per bin      phase_vocoder.cpp:47
	std::vector<float> binf;            // bin_to_frequency(b) = b * float(sr) / float(dft)   PVBuffer.cpp:443-446
	std::vector<float2> binc;           // (binf, expected) interleaved, as the analysis kernel reads them
	std::vector<float2> post_tw;        // e^{-2 pi i k / N}, k = 0..N/2 (the 8-point-per-thread kernels use k <= N/4 only)
	std::vector<float2> post_rot;       // -i e^{-2 pi i k / N} = (sin, -cos)(-2 pi k / N): the analysis unpack's multiplier
	std::vector<float4> binc4;          // per unpack pair k = 0..N/2: (binf[k], binf[M-k], expected[k], -expected[M-k])
	std::vector<float2> pass_tw;        // per-pass Stockham twiddles, concatenated (8 points per thread plan)
	std::vector<float2> pass_tw16;      // same for the 16-points-per-thread plan (dft >= 512)
	std::vector<float2> pass_tw_rev;    // passes R, 16, 16 [, 2] of synthesis_cta_mirror (dft 1024 ... 8192; empty otherwise)
	PvConsts k{};
	double P = 0.0, rcpP = 0.0;
	};

// WindowFunctions.cpp:8-13: pi = std::acos(-1.0f); 0.5f * (1.0f - cos(2.0f * pi * x)). With g++ the
// unqualified cos is ::cos(double): float product promoted, arithmetic in double, narrowed on return.
inline float hann_reference( float x )
	{
	const float pi = std::acos( -1.0f );
	const float arg = 2.0f * pi * x;
	return (float)( 0.5 * ( 1.0 - ::cos( (double) arg ) ) );
	}

template<int M, int PT> inline void append_pass_twiddles( std::vector<float2> & out )
	{
	using P = FftPlan<M, PT>;
	const long double two_pi = 6.283185307179586476925286766559005768L;
	for( int p = 1; p < P::num_passes; ++p )
		{
		const int R = P::radix( p ), NS = P::ns( p );
		for( int r = 1; r < R; ++r )
			for( int jm = 0; jm < NS; ++jm )
				{
				const long double a = -two_pi * (long double) r * (long double) jm / (long double)( NS * R );
				float2 w; w.x = (float) cosl( a ); w.y = (float) sinl( a );
				out.push_back( w );
				}
		}
	}

// Twiddles of one Stockham pass (radix R, Ns): w^(r*jm), w = e^{-2 pi i / (Ns*R)}, laid out [r-1][jm].
inline void append_one_pass_twiddles( std::vector<float2> & out, int R, int NS )
	{
	const long double two_pi = 6.283185307179586476925286766559005768L;
	for( int r = 1; r < R; ++r )
		for( int jm = 0; jm < NS; ++jm )
			{
			const long double a = -two_pi * (long double) r * (long double) jm / (long double)( NS * R );
			float2 w; w.x = (float) cosl( a ); w.y = (float) sinl( a );
			out.push_back( w );
			}
	}

inline bool build_tables( int N, int W, int hop, float sample_rate, float analysis_rate, HostTables & t )
	{
	if( W < 1 || W > N || hop < 1 ) return false;
	t.N = N; t.W = W; t.hop = hop;
	const int M = N / 2, B = M + 1;

	t.k.sample_rate = sample_rate;
	t.k.analysis_rate = analysis_rate;
	t.k.rcp_analysis_rate = 1.0f / analysis_rate;
	const float pi = std::acos( -1.0f );            // defines.h:44
	t.k.pi2 = pi * 2.0f;                            // defines.h:45
	t.k.rcp_pi2 = 1.0f / t.k.pi2;
	t.k.bin_scale = 1.0f / (float) N;               // exact when N is a power of two (bin_frequency_of; the tables below divide)
	t.k.use_wrapping = analysis_rate < sample_rate; // phase_vocoder.cpp:37
	t.k.wrap_pi2 = t.k.use_wrapping ? t.k.pi2 : 0.0f;
	t.P = (double) t.k.pi2;
	t.rcpP = 1.0 / t.P;

	t.win_analysis.resize( W );
	t.win_synthesis.resize( W );
	// AudioPV.cpp:99: 2.67f / ( get_dft_size() * get_window_size() / get_hop_size() ), int arithmetic inside
	const int denom = N * W / hop;
	const float window_scale = 2.67f / denom;
	for( int i = 0; i < W; ++i )
		{
		const float h = hann_reference( float( i ) / float( W - 1 ) );
		t.win_analysis[i] = h;
		t.win_synthesis[i] = h * window_scale;      // AudioPV.cpp:102
		}

	t.expected.resize( B );
	t.binf.resize( B );
	for( int b = 0; b < B; ++b )
		{
		// PVBuffer.cpp:443-446 divides by get_dft_size() = (num_bins - 1) * 2 (PVBuffer.cpp:356-359), which is N - 1 for an odd dft size
		const float binf = (float) b * sample_rate / (float)( ( B - 1 ) * 2 );
		t.binf[b] = binf;
		t.expected[b] = binf / analysis_rate * t.k.pi2;             // phase_vocoder.cpp:47
		}
	t.binc.resize( B );
	for( int b = 0; b < B; ++b ) { t.binc[b].x = t.binf[b]; t.binc[b].y = t.expected[b]; }

	const long double two_pi = 6.283185307179586476925286766559005768L;
	t.post_tw.resize( M + 1 );
	for( int k = 0; k <= M; ++k )
		{
		const long double a = -two_pi * (long double) k / (long double) N;
		t.post_tw[k].x = (float) cosl( a );
		t.post_tw[k].y = (float) sinl( a );
		}

	t.post_rot.resize( M + 1 );
	t.binc4.resize( M + 1 );
	for( int k = 0; k <= M; ++k )
		{
		t.post_rot[k].x = t.post_tw[k].y;
		t.post_rot[k].y = -t.post_tw[k].x;
		t.binc4[k].x = t.binf[k];     t.binc4[k].y = t.binf[M - k];
		t.binc4[k].z = t.expected[k]; t.binc4[k].w = -t.expected[M - k];   // lane y runs on the conjugate
		}

	t.pass_tw.clear();
	t.pass_tw16.clear();
	switch( M )
		{
		case 128:  append_pass_twiddles<128, 8>( t.pass_tw ); break;
		case 256:  append_pass_twiddles<256, 8>( t.pass_tw );   append_pass_twiddles<256, 16>( t.pass_tw16 ); break;
		case 512:  append_pass_twiddles<512, 8>( t.pass_tw );   append_pass_twiddles<512, 16>( t.pass_tw16 ); break;
		case 1024: append_pass_twiddles<1024, 8>( t.pass_tw );  append_pass_twiddles<1024, 16>( t.pass_tw16 ); break;
		case 2048: append_pass_twiddles<2048, 8>( t.pass_tw );  append_pass_twiddles<2048, 16>( t.pass_tw16 ); break;
		case 4096: append_pass_twiddles<4096, 8>( t.pass_tw );  append_pass_twiddles<4096, 16>( t.pass_tw16 ); break;
		default: break;     // every other size: the run-time-sized transform (pv_generic.h) brings its own tables
		}
	t.pass_tw_rev.clear();
	if( M == 512 || M == 1024 || M == 2048 || M == 4096 )
		{
		const int R = M >= 2048 ? 8 : M / 256;              // passes R (no twiddles), 16 (Ns = R), 16 (Ns = 16 R) [, 2 (Ns = 256 R)]
		append_one_pass_twiddles( t.pass_tw_rev, 16, R );
		append_one_pass_twiddles( t.pass_tw_rev, 16, 16 * R );
		if( M == 512 * R ) append_one_pass_twiddles( t.pass_tw_rev, 2, 256 * R );
		}
	return true;
	}

// Frames per CTA. Long enough that the warm-up FFT (analysis) and the shared overlap regions
// (resynthesis: a sample may be shared by at most two segments, which needs seg_len*hop >= W-hop) are
// amortised; short enough that the grid covers the SMs several times over.
// `analysis_only`: the overlap constraint does not apply; short signals then get segments down to 8 frames (one
// warm-up FFT per 8) so that the grid still covers the SMs.
// Resident CTAs per SM of the resynthesis kernels under the launch policy of pv_capi.cu (0: not known).
inline int synth_ctas_per_sm( int dft ) { return dft == 8192 ? 2 : ( ( dft == 1024 || dft == 2048 || dft == 4096 ) ? 3 : 0 ); }

// `ctas_per_sm` > 0 (resynthesis of long signals): the length is shortened (never below ~2/3 of max_len) so that the
// CTA count is just under a multiple of EIGHT waves of the kernel -- a whole number of waves on 1, 2, 4 or 8 devices.
// cfg3 (675 001 frames, 296 CTAs per wave at dft 8192): 128 frames per CTA gave 5 274 CTAs = 2.23 waves on each of
// eight devices, of which the third is a quarter full; 96 frames give 879 CTAs = 2.97 waves. The choice is a function of
// the WHOLE signal only, so every shard of a signal -- and the unsharded call -- walks the same segments (same bits).
inline int choose_seg_len( int64_t frames, int channels, int sms, int W, int hop, int max_len = 64, bool analysis_only = false, int ctas_per_sm = 0 )
	{
	int64_t min_len = ( W + hop - 1 ) / hop;                 // >= W/hop
	if( analysis_only && min_len > 8 ) min_len = 8;
	int64_t target_ctas = (int64_t) sms * 8;
	int64_t len = ( frames * channels + target_ctas - 1 ) / target_ctas;
	if( len > max_len ) len = max_len;
	if( len < min_len ) len = min_len;
	if( len < 4 ) len = 4;
	if( len > frames ) len = frames > 0 ? frames : 1;
	if( ctas_per_sm > 0 && !analysis_only && channels > 0 )
		{
		const int64_t wave = (int64_t) sms * ctas_per_sm;
		const int64_t ctas = channels * ( ( frames + len - 1 ) / len );
		// long signals: a multiple of eight waves (see above); a few waves: just the waves the launch needs anyway (a 1/8
		// shard of cfg3 handed to a per-GPU process: 660 CTAs = 2.23 waves -> 879 = 2.97)
		const int64_t unit = ( ctas >= 16 * wave ) ? 8 * wave : wave;
		if( ctas > wave )
			{
			const int64_t k = ( ctas + unit - 1 ) / unit;
			const int64_t per_channel = ( unit * k ) / channels;
			const int64_t shorter = per_channel > 0 ? ( frames + per_channel - 1 ) / per_channel : len;
			if( shorter >= min_len && shorter >= 4 && shorter < len ) len = shorter;
			}
		}
	return (int) len;
	}

} // namespace pvk
