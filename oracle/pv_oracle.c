/* oracle/pv_oracle.c: TEST INFRASTRUCTURE ONLY -- see pv_oracle.h.
 *
 * Every function cites the reference lines (relative to /root/reference/src/flan) it restates.
 * Float operation order is kept exactly as the reference writes it; build with -ffp-contract=off.
 *
 * Third-party arithmetic: the reference calls FFTW3f (fftwf_plan_dft_r2c_1d / c2r_1d,
 * FFTHelper.cpp:21-24,41,47), which is not vendored and not installed. Its published semantics
 * (unnormalised forward e^{-2 pi i kn/N}; unnormalised c2r that ignores Im of bins 0 and N/2) are
 * restated with a radix-2 FFT evaluated in double and rounded once to float -- operation for
 * operation the same transform as oracle/ref_shim/fftw_standin.cpp backend 0, so that this file and
 * the verbatim reference build agree bit for bit.
 */
#define _GNU_SOURCE
#include "pv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------- double-precision FFT stand-in for FFTW: radix-2 for powers of two, the O(n^2) definition in long double
 * for every other size -- operation for operation oracle/ref_shim/fftw_standin.cpp's F64Plan, so that this file and
 * the verbatim reference build agree bit for bit at any dft size (FFTW itself plans any size, FFTHelper.cpp:16-26) ---------- */

typedef struct { double re, im; } cpx_t;

typedef struct
	{
	int n;
	int pow2;
	cpx_t * tw;     /* exp(-2 pi i k / n), k < n/2 */
	int * rev;
	cpx_t * work;
	cpx_t * out;    /* O(n^2) path only */
	} fft_t;

static int fft_init( fft_t * f, int n )
	{
	if( n < 2 ) return -1;
	f->n = n;
	f->pow2 = ( n & ( n - 1 ) ) == 0;
	f->out = NULL;
	f->work = NULL; f->tw = NULL; f->rev = NULL;
	if( !f->pow2 )
		{
		f->work = (cpx_t *) malloc( sizeof( cpx_t ) * n );
		f->out = (cpx_t *) malloc( sizeof( cpx_t ) * n );
		return 0;
		}
	f->tw = (cpx_t *) malloc( sizeof( cpx_t ) * ( n / 2 ) );
	f->rev = (int *) malloc( sizeof( int ) * n );
	f->work = (cpx_t *) malloc( sizeof( cpx_t ) * n );
	const long double two_pi = 6.283185307179586476925286766559005768L;
	for( int k = 0; k < n / 2; ++k )
		{
		const long double a = -two_pi * (long double) k / (long double) n;
		f->tw[k].re = (double) cosl( a );
		f->tw[k].im = (double) sinl( a );
		}
	int bits = 0;
	while( ( 1 << bits ) < n ) ++bits;
	for( int i = 0; i < n; ++i )
		{
		int r = 0;
		for( int b = 0; b < bits; ++b ) if( i & ( 1 << b ) ) r |= 1 << ( bits - 1 - b );
		f->rev[i] = r;
		}
	return 0;
	}

static void fft_free( fft_t * f ) { free( f->tw ); free( f->rev ); free( f->work ); free( f->out ); }

static void fft_run( fft_t * f, int sign )
	{
	const int n = f->n;
	cpx_t * w = f->work;
	if( !f->pow2 )
		{
		/* the angle depends on (k*j) mod n only: one cosl / sinl per residue instead of per term (same values, same bits) */
		const long double two_pi = 6.283185307179586476925286766559005768L;
		long double * ct = (long double *) malloc( sizeof( long double ) * 2 * (size_t) n );
		long double * st = ct + n;
		for( int r = 0; r < n; ++r )
			{
			const long double a = sign * two_pi * (long double) r / n;
			ct[r] = cosl( a ); st[r] = sinl( a );
			}
		for( int k = 0; k < n; ++k )
			{
			long double acc_re = 0, acc_im = 0;
			for( int j = 0; j < n; ++j )
				{
				const long long r = ( (long long) k * j ) % n;
				const long double c = ct[r], s = st[r];
				const long double xr = w[j].re, xi = w[j].im;
				acc_re += xr * c - xi * s;
				acc_im += xr * s + xi * c;
				}
			f->out[k].re = (double) acc_re; f->out[k].im = (double) acc_im;
			}
		free( ct );
		memcpy( w, f->out, sizeof( cpx_t ) * n );
		return;
		}
	for( int i = 0; i < n; ++i )
		if( i < f->rev[i] ) { cpx_t t = w[i]; w[i] = w[f->rev[i]]; w[f->rev[i]] = t; }
	for( int len = 2; len <= n; len <<= 1 )
		{
		const int half = len / 2, step = n / len;
		for( int i = 0; i < n; i += len )
			for( int j = 0; j < half; ++j )
				{
				cpx_t t = f->tw[j * step];
				if( sign > 0 ) t.im = -t.im;
				const cpx_t u = w[i + j];
				const cpx_t x = w[i + j + half];
				cpx_t v;
				v.re = x.re * t.re - x.im * t.im;
				v.im = x.re * t.im + x.im * t.re;
				w[i + j].re = u.re + v.re;           w[i + j].im = u.im + v.im;
				w[i + j + half].re = u.re - v.re;    w[i + j + half].im = u.im - v.im;
				}
		}
	}

/* ---------- constants: defines.h:44-45 ---------- */

static float pvo_pi( void )  { return acosf( -1.0f ); }
static float pvo_pi2( void ) { return pvo_pi() * 2.0f; }

int64_t pvo_num_frames( int64_t n, int hop )
	{
	/* AudioPV.cpp:17: std::ceil( get_num_frames() / hopSize ) + 1 with an INTEGER quotient. */
	return n / hop + 1;
	}

int pvo_hop_from_rates( float sample_rate, float analysis_rate )
	{
	/* PVBuffer.cpp:381-384: Frame get_hop_size() { return get_sample_rate() / get_analysis_rate(); } */
	return (int)( sample_rate / analysis_rate );
	}

void pvo_hann( int window_size, float * out )
	{
	/* WindowFunctions.cpp:8: const float pi = std::acos( -1.0f );
	 * WindowFunctions.cpp:12: return 0.5f * ( 1.0f - cos( 2.0f * pi * x ) );
	 * With g++ the unqualified cos is ::cos(double): the argument is a float product promoted to
	 * double, and the subtraction / multiply run in double before the return narrows to float.
	 * AudioPV.cpp:33: x = float( i ) / float( window_size - 1 ). */
	const float pi = acosf( -1.0f );
	for( int i = 0; i < window_size; ++i )
		{
		const float x = (float) i / (float)( window_size - 1 );
		const float arg = 2.0f * pi * x;
		out[i] = (float)( 0.5 * ( 1.0 - cos( (double) arg ) ) );
		}
	}

void pvo_mid_side( const float * in, int64_t n, float * out )
	{
	/* AudioConversions.cpp:42-49 */
	const float sqrt2 = sqrtf( 2.0f );
	for( int64_t i = 0; i < n; ++i )
		{
		const float l = in[i], r = in[n + i];
		out[i]     = ( l + r ) / sqrt2;
		out[n + i] = ( l - r ) / sqrt2;
		}
	}

/* phase_vocoder.cpp:5-53 */
static void pvo_phase_vocoder( double * phase_buffer, float re, float im, float bin_frequency,
                               float analysis_rate, float sample_rate, float * m_out, float * f_out )
	{
	const float pi2 = pvo_pi2();
	const int use_wrapping = analysis_rate < sample_rate;                 /* :37 */
	const float phase = atan2f( im, re );                                 /* :43 std::arg */
	const float phase_diff = (float)( (double) phase - *phase_buffer );   /* :44 float - double -> float */
	*phase_buffer = phase;                                                /* :45 */
	const float expected_phase_diff = bin_frequency / analysis_rate * pi2;/* :47 */
	const float delta_phase = phase_diff - expected_phase_diff;           /* :48 */
	float wrapped = delta_phase;
	if( use_wrapping )
		{
		const float q = delta_phase / pi2;                                /* :40 x / pi2 */
		const float r = roundf( q );
		const float pr = pi2 * r;
		wrapped = delta_phase - pr;                                       /* :40 */
		}
	const float t = wrapped * analysis_rate;                              /* :50 left to right */
	const float delta_frequency = t / pi2;
	*m_out = hypotf( re, im );                                            /* :52 std::abs */
	*f_out = bin_frequency + delta_frequency;                             /* :52 */
	}

int pvo_convert_to_pv( const float * audio, int C, int64_t n, float sample_rate,
                       int window_size, int hop, int dft_size,
                       int64_t frame_begin, int64_t frame_end, float * pv_out )
	{
	if( C < 0 || n < 0 || hop <= 0 || window_size <= 0 || dft_size < window_size ) return -1;
	const int num_bins = dft_size / 2 + 1;                                /* :15 */
	const int64_t num_hops = pvo_num_frames( n, hop );                    /* :17 */
	if( frame_begin < 0 || frame_end > num_hops || frame_begin > frame_end ) return -1;
	const float analysis_rate = sample_rate / hop;                        /* :25 float / int */

	float * hann = (float *) malloc( sizeof( float ) * window_size );
	pvo_hann( window_size, hann );                                        /* :30-34 */

	fft_t fft;
	if( fft_init( &fft, dft_size ) != 0 ) { free( hann ); return -1; }
	double * phase_buffer = (double *) malloc( sizeof( double ) * num_bins );   /* :37 */
	float * bin_frequency = (float *) malloc( sizeof( float ) * num_bins );
	for( int b = 0; b < num_bins; ++b )
		bin_frequency[b] = (float) b * sample_rate / (float)( ( num_bins - 1 ) * 2 );    /* PVBuffer.cpp:443-446 with get_dft_size() = (num_bins - 1) * 2, :356-359 */

	const int64_t rows = frame_end - frame_begin;
	for( int c = 0; c < C; ++c )                                          /* :41 */
		{
		const float * x = audio + (int64_t) c * n;
		for( int b = 0; b < num_bins; ++b ) phase_buffer[b] = 0.0;        /* :44 */
		/* frame_begin-1 is run only to obtain the phase the serial loop would carry. */
		const int64_t first = frame_begin > 0 ? frame_begin - 1 : 0;
		for( int64_t pv_frame = first; pv_frame < frame_end; ++pv_frame ) /* :47 */
			{
			const int64_t start = (int64_t) hop * pv_frame - window_size / 2;   /* :52 */
			for( int i = 0; i < window_size; ++i )                        /* :61-62 */
				{
				const int64_t s = start + i;
				const float v = ( s < 0 || n <= s ) ? 0.0f : x[s];        /* :54-58 */
				fft.work[i].re = (double)( v * hann[i] );
				fft.work[i].im = 0.0;
				}
			for( int i = window_size; i < dft_size; ++i ) { fft.work[i].re = 0.0; fft.work[i].im = 0.0; }   /* :65 */
			fft_run( &fft, -1 );                                          /* :67 */
			for( int b = 0; b < num_bins; ++b )                           /* :69-73 */
				{
				const float re = (float) fft.work[b].re, im = (float) fft.work[b].im;
				float m, f;
				pvo_phase_vocoder( &phase_buffer[b], re, im, bin_frequency[b], analysis_rate, sample_rate, &m, &f );
				if( pv_frame >= frame_begin )
					{
					float * o = pv_out + 2 * ( ( (int64_t) c * rows + ( pv_frame - frame_begin ) ) * num_bins + b );
					o[0] = m; o[1] = f;
					}
				}
			}
		}
	free( bin_frequency ); free( phase_buffer ); fft_free( &fft ); free( hann );
	return 0;
	}

int pvo_convert_to_audio( const float * pv, int C, int64_t F, int B, float sample_rate,
                          float analysis_rate, int window_size, float * audio_out )
	{
	if( C < 0 || F < 0 || B < 2 || window_size <= 0 ) return -1;
	const float pi2 = pvo_pi2();
	const int dft_size = ( B - 1 ) * 2;                                   /* PVBuffer.cpp:356-359 */
	const int hop = pvo_hop_from_rates( sample_rate, analysis_rate );
	if( hop <= 0 || window_size > dft_size ) return -1;
	const int64_t out_n = F * hop;                                        /* :93 */
	memset( audio_out, 0, sizeof( float ) * (size_t) C * out_n );         /* :95 zero-filled buffer */

	float * hann = (float *) malloc( sizeof( float ) * window_size );
	pvo_hann( window_size, hann );
	const float window_scale = 2.67f / ( dft_size * window_size / hop );  /* :99 int arithmetic in the parenthesis */
	for( int i = 0; i < window_size; ++i ) hann[i] = hann[i] * window_scale;   /* :102 */

	fft_t fft;
	if( fft_init( &fft, dft_size ) != 0 ) { free( hann ); return -1; }
	double * phase_buffer = (double *) malloc( sizeof( double ) * B );    /* :105 */

	for( int c = 0; c < C; ++c )                                          /* :108 */
		{
		float * out = audio_out + (int64_t) c * out_n;
		for( int b = 0; b < B; ++b ) phase_buffer[b] = 0.0;               /* :111 */
		for( int64_t pv_frame = 0; pv_frame < F; ++pv_frame )             /* :113 */
			{
			const float * row = pv + 2 * ( ( (int64_t) c * F + pv_frame ) * B );
			/* :117-120 + phase_vocoder.cpp:55-61; then Hermitian extension as FFTW c2r reads it
			 * (imaginary parts of bins 0 and N/2 ignored). */
			for( int b = 0; b < B; ++b )
				{
				const float m = row[2 * b], f = row[2 * b + 1];
				const float phase_diff = f / analysis_rate * pi2;         /* phase_vocoder.cpp:57 */
				phase_buffer[b] += phase_diff;                            /* :58 */
				if( phase_buffer[b] > pi2 ) phase_buffer[b] = fmod( phase_buffer[b], (double) pi2 );   /* :59 */
				const float theta = (float) phase_buffer[b];
				const float re = m * cosf( theta ), im = m * sinf( theta );   /* :60 std::polar */
				if( b == 0 || b == dft_size / 2 ) { fft.work[b].re = re; fft.work[b].im = 0.0; }
				else
					{
					fft.work[b].re = re;             fft.work[b].im = im;
					fft.work[dft_size - b].re = re;  fft.work[dft_size - b].im = -(double) im;
					}
				}
			fft_run( &fft, +1 );                                          /* :122 */
			const int64_t start = (int64_t) hop * pv_frame - window_size / 2;   /* :125 */
			const int64_t end = start + window_size;
			const int64_t sb = start > 0 ? start : 0;                     /* :127 */
			const int64_t eb = end < out_n ? end : out_n;                 /* :128 */
			for( int64_t i = sb - start; i < eb - start; ++i )            /* :133-134 */
				{
				const float y = (float) fft.work[i].re;
				const float p = y * hann[i];
				out[start + i] = out[start + i] + p;
				}
			}
		}
	free( phase_buffer ); fft_free( &fft ); free( hann );
	return 0;
	}

/* ---------- PV-domain chain between analysis and resynthesis: PV::repitch / PV::stretch ---------- */

/* Utility/Interpolator.cpp:15-105. sine (7) and sine2 (8) go through libm cos / sin. */
static float pvo_interp( int id, float x )
	{
	switch( id )
		{
		case 1: return 0.5f;                                               /* midpoint   :15-21 */
		case 2: return roundf( x );                                        /* nearest    :24-30 */
		case 3: return 0.0f;                                               /* floor      :33-39 */
		case 4: return 1.0f;                                               /* ceil       :42-48 */
		case 5: return x * x * ( 3.0f - 2.0f * x );                        /* smoothstep :60-66 */
		case 6: return x * x * x * ( x * ( x * 6.0f - 15.0f ) + 10.0f );   /* smootherstep :69-75 */
		case 7: { const float pi = acosf( -1.0f ); return ( 1.0f - cosf( pi * x ) ) / 2.0f; }          /* :77-84 */
		case 8: { const float pi = acosf( -1.0f ); return (float)( (double) sqrtf( 2.0f ) * sin( (double)( pi / 4.0f * x ) ) ); }   /* :86-92 */
		case 9: return sqrtf( x );                                         /* sqrt       :95-101 */
		default: return x;                                                 /* linear     :51-57 */
		}
	}

static int clamp_int( int v, int lo, int hi ) { return v < lo ? lo : ( hi < v ? hi : v ); }

/* modify_frequency_base, PV/PVModify.cpp:196-257. mod: [F][B] mapped bin positions in Hz (shared by all channels,
 * indexed frame*B + bin, :217-219); in_mod: [C][F][B] the mapped frequency written into the output MFs. */
static void pvo_modify_frequency_base( const float * pv, int C, int64_t F, int B, float sample_rate,
                                       const float * mod, const float * in_mod, int interp, float * out )
	{
	const int dft = ( B - 1 ) * 2;                                         /* PVBuffer.cpp:356-359 */
	const float bin_width = sample_rate / (float) dft;                     /* PVBuffer.cpp:438-441 */
	memset( out, 0, sizeof( float ) * 2 * (size_t) C * (size_t) F * (size_t) B );      /* :204-205 */
	for( int c = 0; c < C; ++c )                                           /* :207 */
		for( int64_t frame = 0; frame < F; ++frame )                       /* :211 */
			{
			const float * in_row = pv + 2 * ( ( (int64_t) c * F + frame ) * B );
			const float * im_row = in_mod + ( (int64_t) c * F + frame ) * B;
			const float * mod_row = mod + frame * B;
			float * out_row = out + 2 * ( ( (int64_t) c * F + frame ) * B );
			for( int bin = 1; bin < B; ++bin )                             /* :214 */
				{
				const float loBin = mod_row[bin - 1] / bin_width;          /* :218 */
				const float hiBin = mod_row[bin] / bin_width;              /* :219 */
				const int forward = hiBin > loBin;                         /* :220 */
				const int loBinRound = (int)( forward ? ceilf( loBin ) : floorf( loBin ) );   /* :222 */
				const int hiBinRound = (int)( forward ? ceilf( hiBin ) : floorf( hiBin ) );   /* :223 */
				const int start_bin = clamp_int( loBinRound, 0, B - 1 );   /* :224 */
				const int end_bin = clamp_int( hiBinRound, 0, B - 1 );     /* :225 */
				const float lo_m = in_row[2 * ( bin - 1 )], lo_f = im_row[bin - 1];    /* :227 */
				const float hi_m = in_row[2 * bin], hi_f = im_row[bin];                /* :228 */
				for( int y = start_bin; y != end_bin; forward ? ++y : --y )            /* :230 */
					{
					const float mix = pvo_interp( interp, ( (float) y - loBin ) / ( hiBin - loBin ) );  /* :232 */
					const float w0 = ( 1.0f - mix ) * lo_m;                /* :234 */
					const float w1 = mix * hi_m;                           /* :235 */
					const float max_m = w0 < w1 ? lo_m : hi_m;             /* :237 */
					const float max_f = w0 < w1 ? lo_f : hi_f;
					if( max_m > out_row[2 * y] )                           /* :239 */
						{
						out_row[2 * y] = out_row[2 * y] + max_m;           /* :241 */
						out_row[2 * y + 1] = max_f;                        /* :242 */
						}
					}
				}
			}
	}

/* PV::repitch, PV/PVModify.cpp:273-305. factor: [F][B] the sampled factor (PV/PV.h:31-35). */
int pvo_repitch( const float * pv, int C, int64_t F, int B, float sample_rate, const float * factor, int interp, float * out )
	{
	if( C < 1 || F < 1 || B < 2 ) return -1;
	const int dft = ( B - 1 ) * 2;
	const float bin_width = sample_rate / (float) dft;
	float * fs = (float *) malloc( sizeof( float ) * (size_t) F * B );
	float * in_mod = (float *) malloc( sizeof( float ) * (size_t) C * F * B );
	memcpy( fs, factor, sizeof( float ) * (size_t) F * B );
	for( int64_t frame = 0; frame < F; ++frame )                           /* :278-280 */
		for( int bin = 1; bin < B; ++bin )
			fs[frame * B + bin] = fs[frame * B + bin] + fs[frame * B + bin - 1];
	for( int64_t i = 0; i < F * B; ++i )                                   /* :283-284, PVBuffer.cpp:443-446 */
		fs[i] = fs[i] * sample_rate / (float) dft;
	const float top = (float)( B - 1 ) - 0.0001f;                          /* :293 */
	for( int c = 0; c < C; ++c )                                           /* :289-302 */
		for( int64_t frame = 0; frame < F; ++frame )
			for( int bin = 0; bin < B; ++bin )
				{
				const float f = pv[2 * ( ( (int64_t) c * F + frame ) * B + bin ) + 1];
				float fbin = f / bin_width;                                /* PVBuffer.cpp:438-441 */
				fbin = fbin < 0.0f ? 0.0f : ( top < fbin ? top : fbin );   /* std::clamp */
				const int lo = (int) floorf( fbin );                       /* :294 */
				const int hi = lo + 1;                                     /* :295 */
				const float lo_freq = fs[frame * B + lo];                  /* :296 */
				const float hi_freq = fs[frame * B + hi];                  /* :297 */
				const float r = fbin - (float) lo;                         /* :298 */
				in_mod[( (int64_t) c * F + frame ) * B + bin] = lo_freq * ( 1.0f - r ) + hi_freq * r;   /* :299 */
				}
	pvo_modify_frequency_base( pv, C, F, B, sample_rate, fs, in_mod, interp, out );    /* :304 */
	free( fs ); free( in_mod );
	return 0;
	}

/* Output frame count of PV::stretch / PV::modify_time for a mod table in seconds: ceil( time_to_frame( max ) ),
 * PV/PVModify.cpp:312-315, PVBuffer.cpp:428-431. */
static int64_t pvo_time_frames( const float * mod, int64_t count, float sample_rate, int hop )
	{
	float mx = mod[0];
	for( int64_t i = 1; i < count; ++i ) if( mx < mod[i] ) mx = mod[i];    /* std::max_element */
	const float last = ceilf( mx * sample_rate / (float) hop );
	return (int64_t)(int) last;
	}

/* modify_time_base, PV/PVModify.cpp:307-362. mod: [F][B] seconds. out: [C][out_frames][B], zero-filled here. */
static void pvo_modify_time_base( const float * pv, int C, int64_t F, int B, float sample_rate, int hop,
                                  const float * mod, int interp, int64_t out_frames, float * out )
	{
	memset( out, 0, sizeof( float ) * 2 * (size_t) C * (size_t) out_frames * (size_t) B );          /* :317-318 */
	for( int c = 0; c < C; ++c )                                           /* :320 */
		for( int bin = 0; bin < B; ++bin )                                 /* :326 */
			for( int64_t frame = 1; frame < F; ++frame )                   /* :329 */
				{
				const float lFrame = mod[( frame - 1 ) * B + bin] * sample_rate / (float) hop;      /* :331 */
				const float rFrame = mod[frame * B + bin] * sample_rate / (float) hop;              /* :332 */
				const int forward = rFrame > lFrame;                       /* :333 */
				const int start_frame = (int)( forward ? ceilf( lFrame ) : floorf( lFrame ) );      /* :335 */
				const int end_frame = (int)( forward ? ceilf( rFrame ) : floorf( rFrame ) );        /* :336 */
				const float * l = pv + 2 * ( ( (int64_t) c * F + frame - 1 ) * B + bin );           /* :338 */
				const float * r = pv + 2 * ( ( (int64_t) c * F + frame ) * B + bin );               /* :339 */
				for( int x = start_frame; x != end_frame; forward ? ++x : --x )                     /* :341 */
					{
					if( x < 0 || out_frames <= x ) continue;               /* :343 */
					const float mix = pvo_interp( interp, ( (float) x - lFrame ) / ( rFrame - lFrame ) );   /* :345 */
					const float w0 = ( 1.0f - mix ) * l[0];                /* :346 */
					const float w1 = mix * r[0];                           /* :347 */
					const float totalWeight = w0 + w1;                     /* :348 */
					const float weightedFreqSum = w0 * l[1] + w1 * r[1];   /* :349 */
					if( totalWeight == 0.0f ) break;                       /* :351-352: `return` leaves this frame pair */
					float * o = out + 2 * ( ( (int64_t) c * out_frames + x ) * B + bin );
					o[1] = ( o[1] * o[0] + weightedFreqSum ) / ( o[0] + totalWeight );              /* :354 */
					o[0] = o[0] + totalWeight;                             /* :355 */
					}
				}
	}

/* PV::modify_time with a pre-sampled mod table in seconds (PV/PVModify.cpp:364-369). out == NULL: returns the
 * output frame count only; otherwise out must hold [C][that many][B] MFs. */
int64_t pvo_modify_time( const float * pv, int C, int64_t F, int B, float sample_rate, float analysis_rate,
                         const float * mod_seconds, int interp, float * out )
	{
	if( C < 1 || F < 1 || B < 2 ) return -1;
	const int hop = pvo_hop_from_rates( sample_rate, analysis_rate );
	const int64_t out_frames = pvo_time_frames( mod_seconds, F * B, sample_rate, hop );
	if( out && out_frames > 0 ) pvo_modify_time_base( pv, C, F, B, sample_rate, hop, mod_seconds, interp, out_frames, out );
	return out_frames;
	}

/* PV::stretch, PV/PVModify.cpp:371-385. factor: [F][B]. Same out convention as pvo_modify_time. */
int64_t pvo_stretch( const float * pv, int C, int64_t F, int B, float sample_rate, float analysis_rate,
                     const float * factor, int interp, float * out )
	{
	if( C < 1 || F < 1 || B < 2 ) return -1;
	const int hop = pvo_hop_from_rates( sample_rate, analysis_rate );
	float * fs = (float *) malloc( sizeof( float ) * (size_t) F * B );
	memcpy( fs, factor, sizeof( float ) * (size_t) F * B );
	for( int bin = 0; bin < B; ++bin )                                     /* :376-378 */
		for( int64_t frame = 1; frame < F; ++frame )
			fs[frame * B + bin] = fs[frame * B + bin] + fs[( frame - 1 ) * B + bin];
	const float rate = sample_rate / (float) hop;                          /* PVBuffer.cpp:433-436 */
	for( int64_t i = 0; i < F * B; ++i ) fs[i] = fs[i] / rate;             /* :381-382 */
	const int64_t out_frames = pvo_modify_time( pv, C, F, B, sample_rate, analysis_rate, fs, interp, out );   /* :384 */
	free( fs );
	return out_frames;
	}

/* ---------- file formats either side of the path ---------- */

/* PVBuffer::save's sample codec, PV/PVBuffer.cpp:99-127: 24-bit signed little-endian, magnitude scaled by the dft
 * size, frequency by the sample rate, (m, f) interleaved. count = number of MF elements; bytes: 6 * count. */
void pvo_flan_encode( const float * pv, int64_t count, float dft_size, float sample_rate, uint8_t * bytes )
	{
	const double limit = 8388608.0;                                        /* pow( 2, 8 * 3 - 1 ), :102 */
	for( int64_t i = 0; i < 2 * count; ++i )
		{
		const float div = ( i & 1 ) ? sample_rate : dft_size;              /* :103-104 */
		float v = pv[i] / div;
		v = v < -1.0f ? -1.0f : ( 1.0f < v ? 1.0f : v );                   /* std::clamp, :112-113 */
		const double scaled = (double) v * limit;
		const int32_t q = ( scaled != scaled ) ? (int32_t) 0x80000000u : (int32_t) scaled;   /* x86 cvttsd2si of NaN */
		bytes[3 * i + 0] = (uint8_t)( ( q >> 0 ) & 0xFF );                 /* :117-123 */
		bytes[3 * i + 1] = (uint8_t)( ( q >> 8 ) & 0xFF );
		bytes[3 * i + 2] = (uint8_t)( ( q >> 16 ) & 0xFF );
		}
	}

/* PVBuffer::load's sample codec, PV/PVBuffer.cpp:253-268. */
void pvo_flan_decode( const uint8_t * bytes, int64_t count, float dft_size, float sample_rate, float * pv )
	{
	const double limit = 8388608.0;                                        /* :253 */
	for( int64_t i = 0; i < 2 * count; ++i )
		{
		int32_t q = (int32_t)( bytes[3 * i] | ( bytes[3 * i + 1] << 8 ) | ( bytes[3 * i + 2] << 16 ) );   /* :259-260 */
		if( q & 0x800000 ) q |= (int32_t) 0xFF000000u;                     /* :261 */
		pv[i] = (float)( (double) q / limit ) * ( ( i & 1 ) ? sample_rate : dft_size );      /* :262 */
		}
	}

/* AudioBuffer::save (Audio/AudioBuffer.cpp:136-161) hands interleaved, clamped floats to libsndfile with
 * SF_FORMAT_WAV | SF_FORMAT_PCM_24. libsndfile is an external dependency, neither vendored nor installed here
 * (version unpinned; cmake/FindSndFile.cmake accepts any): PARITY UNPINNED for this codec. Its published float -> 24-bit
 * conversion (src/pcm.c, f2let_array with normalisation on, the default) is restated: lrintf( x * 0x7FFFFF ), three
 * little-endian bytes, frames interleaved. audio: planar float[C][n]; bytes: 3 * C * n. */
void pvo_pcm24_encode( const float * audio, int C, int64_t n, uint8_t * bytes )
	{
	for( int c = 0; c < C; ++c )
		for( int64_t i = 0; i < n; ++i )
			{
			float s = audio[(int64_t) c * n + i];
			s = s < -1.0f ? -1.0f : ( 1.0f < s ? 1.0f : s );               /* AudioBuffer.cpp:158-161 */
			const int32_t q = (int32_t) lrintf( s * 8388607.0f );          /* pcm.c f2let_array */
			uint8_t * o = bytes + 3 * ( i * C + c );                       /* AudioBuffer.cpp:153-155 */
			o[0] = (uint8_t)( q & 0xFF ); o[1] = (uint8_t)( ( q >> 8 ) & 0xFF ); o[2] = (uint8_t)( ( q >> 16 ) & 0xFF );
			}
	}

/* AudioBuffer::load (Audio/AudioBuffer.cpp:112-125): sf_readf_float of 24-bit PCM = value / 2^23 (pcm.c let2f_array:
 * ( sample << 8 ) * ( 1 / 0x80000000 ) ), then de-interleave. */
void pvo_pcm24_decode( const uint8_t * bytes, int C, int64_t n, float * audio )
	{
	for( int c = 0; c < C; ++c )
		for( int64_t i = 0; i < n; ++i )
			{
			const uint8_t * b = bytes + 3 * ( i * C + c );
			const int32_t v = (int32_t)( ( (uint32_t) b[0] << 8 ) | ( (uint32_t) b[1] << 16 ) | ( (uint32_t) b[2] << 24 ) );
			audio[(int64_t) c * n + i] = (float) v * ( 1.0f / 2147483648.0f );
			}
	}
