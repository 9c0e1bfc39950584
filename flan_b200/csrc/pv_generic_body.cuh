// flan_b200/csrc/pv_generic_body.cuh
//
// CTA bodies for the dft sizes the templated kernels of pv_body.cuh do not cover: anything that is not a power of two
// in [256, 8192] -- small and large powers of two (64, 128, 16384, 32768, ...), even sizes such as 3000 or 1536, odd
// sizes. The reference plans an FFTW transform for ANY size (FFTHelper.cpp:16-26; Audio.h:151-153 only warns that
// non-powers of two are slower), so no size may come back as a null PV.
//
// Same arithmetic around the transform as the fast kernels (window, phase_vocoder() / inverse_phase_vocoder() in the
// reference's float op order, overlap-add in increasing frame order, pv_core.cuh), same segment walk, same Env, and the
// same emulator coverage. What differs is the FFT: a run-time-sized complex transform of L points (L = N/2 for even
// N with the real pack / unpack, L = N for odd N) done as
//   * L a power of two: Stockham radix-4 / radix-2 passes over two ping-pong buffers, any number of threads;
//   * otherwise Bluestein: X[k] = c[k] * sum_n ( u[n] c[n] ) * conj( c[k-n] ), c[n] = e^{-i pi n^2 / L}, the convolution
//     through two power-of-two FFTs of M >= 2L - 1 points and a precomputed FFT of the chirp (host, double precision).
// The buffers live in shared memory when they fit and in a per-CTA slab of global memory (L2-resident) otherwise, so
// there is no upper size limit other than memory. This is the correctness path for unusual sizes, not a tuned one.
#pragma once

#include "pv_body.cuh"

namespace pvk {

struct GenericFft
	{
	int N;                      // dft size
	int even;                   // N even: half-size complex transform plus the real pack / unpack
	int L;                      // complex transform length: N/2 (even) or N (odd)
	int B;                      // N/2 + 1 bins
	int M;                      // power-of-two FFT length actually run: L, or >= 2L - 1 (Bluestein)
	int bluestein;
	const float2 * tw;          // [M]   e^{-2 pi i k / M}
	const float2 * chirp;       // [L]   e^{-i pi n^2 / L}                                   (Bluestein)
	const float2 * chirp_fft;   // [M]   FFT_M of conj( chirp ) wrapped to (-L, L), times 1/M  (Bluestein)
	};

PV_HD float2 gconj( float2 a ) { float2 r; r.x = a.x; r.y = -a.y; return r; }

// In-order complex FFT of M = 2^m points, forward sign, from `in` (result lands in `in` or `other`; returned).
// Stockham autosort: pass with radix R and stride Ns reads j + r*M/R and writes (j/Ns)*Ns*R + j%Ns + r*Ns.
// Ends with a barrier: every thread may read the whole result.
template<class Env>
PV_HD float2 * generic_fft_pow2( Env & env, int t, int T, int M, float2 * in, float2 * other, const float2 * tw )
	{
	float2 * src = in, * dst = other;
	int Ns = 1;
	while( Ns < M )
		{
		if( M / Ns >= 4 )
			{
			const int q = M >> 2, step = M / ( 4 * Ns );
			for( int j = t; j < q; j += T )
				{
				const int k = j & ( Ns - 1 );
				float2 v[4];
				v[0] = src[j]; v[1] = src[j + q]; v[2] = src[j + 2 * q]; v[3] = src[j + 3 * q];
				if( Ns > 1 )
					{
					v[1] = cmul( v[1], env.ldg2( tw + k * step ) );
					v[2] = cmul( v[2], env.ldg2( tw + 2 * k * step ) );
					v[3] = cmul( v[3], env.ldg2( tw + 3 * k * step ) );
					}
				dft4<1>( v );
				float2 * o = dst + ( ( ( j - k ) << 2 ) + k );
				o[0] = v[0]; o[Ns] = v[1]; o[2 * Ns] = v[2]; o[3 * Ns] = v[3];
				}
			Ns <<= 2;
			}
		else
			{
			const int q = M >> 1, step = M / ( 2 * Ns );
			for( int j = t; j < q; j += T )
				{
				const int k = j & ( Ns - 1 );
				const float2 u = src[j];
				float2 w = src[j + q];
				if( Ns > 1 ) w = cmul( w, env.ldg2( tw + k * step ) );
				float2 * o = dst + ( ( ( j - k ) << 1 ) + k );
				o[0] = add2( u, w ); o[Ns] = sub2( u, w );
				}
			Ns <<= 1;
			}
		env.sync();
		float2 * s = src; src = dst; dst = s;
		}
	return src;
	}

// Forward DFT of the L values in a[0, L) (a and b hold M + 2 elements each). Returns where the L results are; the other
// buffer is free afterwards. Ends with a barrier.
template<class Env>
PV_HD float2 * generic_dft( Env & env, int t, int T, const GenericFft & g, float2 * a, float2 * b )
	{
	if( !g.bluestein ) return generic_fft_pow2( env, t, T, g.M, a, b, g.tw );
	for( int n = t; n < g.M; n += T )
		{
		float2 v; v.x = 0.0f; v.y = 0.0f;
		if( n < g.L ) v = cmul( a[n], env.ldg2( g.chirp + n ) );
		a[n] = v;
		}
	env.sync();
	float2 * A = generic_fft_pow2( env, t, T, g.M, a, b, g.tw );
	float2 * spare = ( A == a ) ? b : a;
	// convolution theorem; the inverse transform is conj( FFT( conj( . ) ) ) / M, the 1/M sits in chirp_fft
	for( int k = t; k < g.M; k += T ) A[k] = gconj( cmul( A[k], env.ldg2( g.chirp_fft + k ) ) );
	env.sync();
	float2 * R = generic_fft_pow2( env, t, T, g.M, A, spare, g.tw );
	float2 * out = ( R == a ) ? b : a;
	for( int k = t; k < g.L; k += T ) out[k] = cmul( gconj( R[k] ), env.ldg2( g.chirp + k ) );
	env.sync();
	return out;
	}

// Per-CTA scratch (global memory): analysis keeps the previous phase of every bin, resynthesis the fp64 phase
// accumulators and the overlap-add ring; the two FFT buffers follow when they do not fit shared memory.
PV_HD int64_t generic_align16( int64_t bytes ) { return ( bytes + 15 ) & ~(int64_t) 15; }
PV_HD int64_t generic_fft_bytes( const GenericFft & g ) { return 2 * (int64_t) sizeof( float2 ) * ( g.M + 2 ); }
PV_HD int64_t generic_analysis_state_bytes( const GenericFft & g ) { return generic_align16( (int64_t) sizeof( float ) * g.B ); }
PV_HD int64_t generic_synthesis_state_bytes( const GenericFft & g, int W )
	{ return generic_align16( (int64_t) sizeof( double ) * g.B ) + generic_align16( (int64_t) sizeof( float ) * W ); }

struct GenericAnalysisArgs
	{
	AnalysisArgs a;             // audio, pv rows, frame range, segments, W, hop, win, binc, post_rot, k  (pv_body.cuh)
	GenericFft g;
	unsigned char * scratch;    // [blocks][scratch_stride]
	int64_t scratch_stride;
	int fft_in_smem;
	int64_t total_segments;     // channels * segs_per_channel; the CTAs of the launch share them round-robin
	};

template<class Env>
PV_HD void generic_analysis_cta( const GenericAnalysisArgs & ga, int64_t block, int64_t nblocks, int T, Env & env, float2 * smem )
	{
	const AnalysisArgs & a = ga.a;
	const GenericFft & g = ga.g;
	const int t = env.tid;
	const int W = a.W, hop = a.hop, half = W / 2, B = g.B, L = g.L;
	unsigned char * my = ga.scratch + block * ga.scratch_stride;
	float * prev = reinterpret_cast<float *>( my );
	float2 * buf0 = ga.fft_in_smem ? smem : reinterpret_cast<float2 *>( my + generic_analysis_state_bytes( g ) );
	float2 * buf1 = buf0 + ( g.M + 2 );

	for( int64_t seg_id = block; seg_id < ga.total_segments; seg_id += nblocks )
		{
		const int c = (int)( seg_id / a.segs_per_channel );
		const int seg = (int)( seg_id % a.segs_per_channel );
		const int64_t fa = a.frame_begin + (int64_t) seg * a.seg_len;
		const int64_t fb = ( fa + a.seg_len < a.frame_end ) ? fa + a.seg_len : a.frame_end;
		if( fa >= fb ) continue;
		const float * xch = a.audio + (int64_t) c * a.audio_stride;
		for( int k = t; k < B; k += T ) prev[k] = 0.0f;
		env.sync();
		// the serial reference loop carries frame f-1's phase into frame f (phase_vocoder.cpp:44-45): a segment that does
		// not start at frame 0 recomputes it with one warm-up transform
		const int64_t first = ( fa > 0 ) ? fa - 1 : fa;
		for( int64_t f = first; f < fb; ++f )
			{
			const int64_t start = (int64_t) hop * f - half;                 // AudioPV.cpp:52
			auto sample = [&]( int i ) -> float                              // windowed, zero outside the window / the signal (:54-65)
				{
				const int64_t p = start + i;
				if( i >= W || p < 0 || p >= a.n_total ) return 0.0f;
				return mul_rn( env.ldg( xch + ( p - a.audio_offset ) ), env.ldg( a.win + i ) );
				};
			if( g.even )
				for( int n = t; n < L; n += T )
					{
					// the factor 1/2 of the real-FFT unpack is applied here (exact)
					float2 v; v.x = 0.5f * sample( 2 * n ); v.y = 0.5f * sample( 2 * n + 1 );
					buf0[n] = v;
					}
			else
				for( int n = t; n < L; n += T ) { float2 v; v.x = sample( n ); v.y = 0.0f; buf0[n] = v; }
			env.sync();
			const float2 * Z = generic_dft( env, t, T, g, buf0, buf1 );

			const bool emit = ( f >= fa );
			float2 * row = a.pv + ( (int64_t) c * a.pv_channel_stride + ( f - a.frame_begin ) * (int64_t) B );
			for( int k = t; k < B; k += T )
				{
				float2 X;
				if( g.even )
					{
					// X[k] = (Zk + conj Zm) + (-i w_k)(Zk - conj Zm), Zm = Z[L-k], indices modulo L (k = 0 and k = L pair Z[0] with itself)
					const float2 zk = Z[k == L ? 0 : k];
					const float2 zm = gconj( Z[( k == 0 || k == L ) ? 0 : L - k] );
					const float2 A = add2( zk, zm ), D = sub2( zk, zm );
					X = add2( A, cmul( D, env.ldg2( a.post_rot + k ) ) );
					}
				else X = Z[k];
				const float2 cc = env.ldg2( a.binc + k );
				float pp = prev[k];
				const float2 mf = phase_vocoder_bin( X.x, X.y, pp, cc.x, cc.y, a.k );     // AudioPV.cpp:69-73
				prev[k] = pp;
				if( emit ) row[k] = mf;
				}
			env.sync();
			}
		}
	}

struct GenericSynthArgs
	{
	SynthArgs a;                // pv rows, output span, acc_start, segments, W, hop, win (Hann * window_scale), post_tw, k, P  (pv_body.cuh)
	GenericFft g;
	unsigned char * scratch;
	int64_t scratch_stride;
	int fft_in_smem;
	};

template<class Env>
PV_HD void generic_synthesis_cta( const GenericSynthArgs & ga, int64_t block, int64_t nblocks, int T, Env & env, float2 * smem )
	{
	const SynthArgs & a = ga.a;
	const GenericFft & g = ga.g;
	const int t = env.tid;
	const int W = a.W, hop = a.hop, half = W / 2, B = g.B, L = g.L;
	const int fin = ( hop < W ) ? hop : W;                   // samples finalised per frame
	unsigned char * my = ga.scratch + block * ga.scratch_stride;
	double * acc = reinterpret_cast<double *>( my );
	float * ring = reinterpret_cast<float *>( my + generic_align16( (int64_t) sizeof( double ) * B ) );
	float2 * buf0 = ga.fft_in_smem ? smem : reinterpret_cast<float2 *>( my + generic_synthesis_state_bytes( g, W ) );
	float2 * buf1 = buf0 + ( g.M + 2 );
	const int per_launch = a.seg_count > 0 ? a.seg_count : a.segs_per_channel;
	const int64_t total = (int64_t) a.channels * per_launch;

	for( int64_t seg_id = block; seg_id < total; seg_id += nblocks )
		{
		const int c = (int)( seg_id / per_launch );
		const int seg = a.seg_first + (int)( seg_id % per_launch );
		const int64_t fa = a.frame_begin + (int64_t) seg * a.seg_len;
		const int64_t fb = ( fa + a.seg_len < a.frame_end ) ? fa + a.seg_len : a.frame_end;
		if( fa >= fb ) continue;
		const double * acc0 = a.acc_start + ( (int64_t) c * a.segs_per_channel + seg ) * B;
		for( int k = t; k < B; k += T ) acc[k] = acc0[k];
		for( int i = t; i < W; i += T ) ring[i] = 0.0f;
		env.sync();

		float * och = a.out + (int64_t) c * a.out_stride;
		// samples whose every contributing frame lies in this segment are stored; those shared with the neighbouring
		// segment receive exactly two partial sums, combined with red.add on the pre-zeroed output (pv_body.cuh: synthesis_cta)
		const int64_t interior_lo = (int64_t) hop * fa + half - hop;
		const int64_t interior_hi = (int64_t) hop * fb - half;
		auto flush = [&]( int64_t lo, int64_t hi )
			{
			for( int64_t s = lo + t; s < hi; s += T )
				{
				const int slot = (int)( ( s % W + W ) % W );
				const float val = ring[slot];
				ring[slot] = 0.0f;
				if( s >= a.out_lo && s < a.out_hi )
					{
					float * dst = och + ( s - a.out_offset );
					if( s >= interior_lo && s < interior_hi ) env.st_stream( dst, val );
					else env.red_add( dst, val );
					}
				}
			};

		const float2 * pv_ch = a.pv + (int64_t) c * a.pv_channel_stride;
		for( int64_t f = fa; f < fb; ++f )
			{
			const int64_t start = (int64_t) hop * f - half;                          // AudioPV.cpp:125
			const float2 * row = pv_ch + ( f - a.frame_begin ) * (int64_t) B;
			float2 * X = buf1;
			for( int k = t; k < B; k += T )
				{
				const float2 mf = row[k];
				double ph = acc[k];
				phase_accumulate( ph, phase_increment( mf.y, a.k ), a.P, a.rcpP );      // phase_vocoder.cpp:57-59
				acc[k] = ph;
				float sn, cs;
				sincos_pv( (float) ph, &sn, &cs );
				float2 x; x.x = mul_rn( mf.x, cs ); x.y = mul_rn( mf.x, sn );           // :60 std::polar
				X[k] = x;
				}
			env.sync();
			// inverse transform = conj( forward transform of the conjugate ): the conjugated input goes into buf0. The dft
			// size of a PV is (num_bins - 1) * 2 (PVBuffer.cpp:356-359), always even: the half-size transform with the real pack.
			for( int k = t; k < L; k += T )
					{
					// Z'[k] = (X[k] + conj X[L-k]) + i e^{+2 pi i k/N} (X[k] - conj X[L-k]); a c2r transform ignores Im of bins 0 and N/2
					float2 xk = X[k], xm = X[L - k];
					if( k == 0 ) { xk.y = 0.0f; xm.y = 0.0f; }
					const float2 A = add2( xk, gconj( xm ) ), D = sub2( xk, gconj( xm ) );
					const float2 w = gconj( env.ldg2( a.post_tw + k ) );                 // e^{+2 pi i k/N}
					const float2 q = cmul( D, w );
					float2 z; z.x = A.x - q.y; z.y = A.y + q.x;                          // A + i q
					buf0[k] = gconj( z );
					}
			env.sync();
			const float2 * R = generic_dft( env, t, T, g, buf0, buf1 );
			// z[n] = conj( R[n] ): y[2n] = Re z[n], y[2n+1] = Im z[n]. Windowed overlap-add (AudioPV.cpp:133-134).
			for( int n = t; n < L; n += T )
				{
				const float2 r = R[n];
				const int i0 = 2 * n;
				if( i0 < W )
					{
					float * p0 = ring + (int)( ( ( start + i0 ) % W + W ) % W );
					*p0 = add_rn( *p0, mul_rn( r.x, env.ldg( a.win + i0 ) ) );
					}
				if( i0 + 1 < W )
					{
					float * p1 = ring + (int)( ( ( start + i0 + 1 ) % W + W ) % W );
					*p1 = add_rn( *p1, mul_rn( -r.y, env.ldg( a.win + i0 + 1 ) ) );
					}
				}
			env.sync();
			flush( start, start + fin );
			env.sync();
			}
		const int64_t last_start = (int64_t) hop * ( fb - 1 ) - half;
		flush( last_start + fin, last_start + W );
		env.sync();
		}
	}

} // namespace pvk
