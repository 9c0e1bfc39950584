// oracle/ref_shim: TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Stand-in for the seven libfftw3f entry points the reference's FFTHelper.cpp calls
// (FFTHelper.cpp:21-24, :32-35, :41, :47). FFTW 3 (single precision, version unpinned by the
// reference: README.md:24 suggests 3.3.9; cmake/FindFFTWF.cmake accepts any) is an external
// dependency that is neither vendored under /root/reference nor installed in this image.
// Published semantics restated here:
//   r2c: out[k] = sum_n in[n] * exp(-2*pi*i*k*n/N), k = 0..N/2, unnormalised.
//   c2r: out[n] = sum over the Hermitian-extended spectrum of in[k] * exp(+2*pi*i*k*n/N),
//        unnormalised (N x the true inverse); the imaginary parts of in[0] and in[N/2] are
//        ignored; the input array may be destroyed.
//
// Two backends, chosen when a plan is created (flan_ref_set_fft_backend):
//   0 = "f64": our own radix-2 FFT evaluated in double and rounded once to float. This is the
//       nearest-float32 answer, which any correct float FFT (FFTW included) approximates to
//       ~1e-7 of the frame peak. Used for parity.
//   1 = "pffft": the reference's vendored float SIMD FFT (src/r8brain/pffft.cpp, compiled in
//       place from /root/reference), the closest in-tree analogue of FFTW's speed class.
//       Used for the timed CPU baseline.
#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <atomic>

#include "pffft.h"

namespace {

std::atomic<int> g_backend{ 0 };

struct F64Plan
	{
	int n = 0;
	bool pow2 = false;
	std::vector<std::complex<double>> tw;   // exp(-2 pi i k / n), k < n/2
	std::vector<int> rev;
	std::vector<std::complex<double>> work;

	explicit F64Plan( int n_ ) : n( n_ ), work( n_ )
		{
		pow2 = n > 0 && ( n & ( n - 1 ) ) == 0;
		if( !pow2 ) return;
		tw.resize( n / 2 > 0 ? n / 2 : 1 );
		const long double two_pi = 6.283185307179586476925286766559005768L;
		for( int k = 0; k < n / 2; ++k )
			{
			const long double a = -two_pi * (long double) k / (long double) n;
			tw[k] = { (double) cosl( a ), (double) sinl( a ) };
			}
		rev.resize( n );
		int bits = 0;
		while( ( 1 << bits ) < n ) ++bits;
		for( int i = 0; i < n; ++i )
			{
			int r = 0;
			for( int b = 0; b < bits; ++b ) if( i & ( 1 << b ) ) r |= 1 << ( bits - 1 - b );
			rev[i] = r;
			}
		}

	// In-place complex FFT of work[], sign = -1 forward, +1 backward, unnormalised.
	void transform( int sign )
		{
		if( !pow2 )
			{
			// O(n^2) fallback, exact enough for an oracle.
			std::vector<std::complex<double>> out( n );
			const long double two_pi = 6.283185307179586476925286766559005768L;
			// the angle depends on (k*j) mod n only: one cosl / sinl per residue instead of per term
			std::vector<std::complex<long double>> root( n );
			for( int r = 0; r < n; ++r )
				{
				const long double a = sign * two_pi * (long double) r / n;
				root[r] = std::complex<long double>( cosl( a ), sinl( a ) );
				}
			for( int k = 0; k < n; ++k )
				{
				std::complex<long double> acc = 0;
				for( int j = 0; j < n; ++j )
					acc += std::complex<long double>( work[j] ) * root[( (long long) k * j ) % n];
					{
					}
				out[k] = { (double) acc.real(), (double) acc.imag() };
				}
			work = out;
			return;
			}
		for( int i = 0; i < n; ++i ) if( i < rev[i] ) std::swap( work[i], work[rev[i]] );
		for( int len = 2; len <= n; len <<= 1 )
			{
			const int half = len / 2, step = n / len;
			for( int i = 0; i < n; i += len )
				for( int j = 0; j < half; ++j )
					{
					std::complex<double> w = tw[j * step];
					if( sign > 0 ) w = std::conj( w );
					const std::complex<double> u = work[i + j];
					const std::complex<double> v = work[i + j + half] * w;
					work[i + j] = u + v;
					work[i + j + half] = u - v;
					}
			}
		}
	};

} // namespace

struct fftwf_plan_s
	{
	int n = 0;
	bool r2c = true;
	int backend = 0;
	float * real = nullptr;
	fftwf_complex * cpx = nullptr;
	F64Plan * f64 = nullptr;
	PFFFT_Setup * pf = nullptr;
	float * pf_in = nullptr;
	float * pf_out = nullptr;
	float * pf_work = nullptr;
	};

extern "C" {

void flan_ref_set_fft_backend( int backend ) { g_backend = backend; }
int flan_ref_get_fft_backend() { return g_backend; }

float * fftwf_alloc_real( size_t n ) { return (float *) pffft_aligned_malloc( n * sizeof( float ) ); }
fftwf_complex * fftwf_alloc_complex( size_t n ) { return (fftwf_complex *) pffft_aligned_malloc( n * sizeof( fftwf_complex ) ); }
void fftwf_free( void * p ) { pffft_aligned_free( p ); }

static fftwf_plan make_plan( int n, float * real, fftwf_complex * cpx, bool r2c )
	{
	fftwf_plan p = new fftwf_plan_s;
	p->n = n; p->r2c = r2c; p->real = real; p->cpx = cpx;
	p->backend = g_backend;
	const bool pf_ok = n >= 32 && ( n & ( n - 1 ) ) == 0;   // pffft real: N = 2^a 3^b 5^c, a >= 5 (pffft.h:60-66)
	if( p->backend == 1 && pf_ok )
		{
		p->pf = pffft_new_setup( n, PFFFT_REAL );
		p->pf_in   = (float *) pffft_aligned_malloc( n * sizeof( float ) );
		p->pf_out  = (float *) pffft_aligned_malloc( n * sizeof( float ) );
		p->pf_work = (float *) pffft_aligned_malloc( n * sizeof( float ) );
		}
	else
		{
		p->backend = 0;
		p->f64 = new F64Plan( n );
		}
	return p;
	}

fftwf_plan fftwf_plan_dft_r2c_1d( int n, float * in, fftwf_complex * out, unsigned ) { return make_plan( n, in, out, true ); }
fftwf_plan fftwf_plan_dft_c2r_1d( int n, fftwf_complex * in, float * out, unsigned ) { return make_plan( n, out, in, false ); }

void fftwf_execute( const fftwf_plan p )
	{
	const int n = p->n, h = n / 2;
	if( p->backend == 1 )
		{
		if( p->r2c )
			{
			// pffft ordered real spectrum: [0]=Re X0, [1]=Re X(N/2), then (Re,Im) of X1..X(N/2-1) (pffft.h:111-133)
			pffft_transform_ordered( p->pf, p->real, p->pf_out, p->pf_work, PFFFT_FORWARD );
			p->cpx[0][0] = p->pf_out[0]; p->cpx[0][1] = 0.0f;
			p->cpx[h][0] = p->pf_out[1]; p->cpx[h][1] = 0.0f;
			std::memcpy( &p->cpx[1][0], p->pf_out + 2, sizeof( float ) * 2 * ( h - 1 ) );
			}
		else
			{
			p->pf_in[0] = p->cpx[0][0];
			p->pf_in[1] = p->cpx[h][0];
			std::memcpy( p->pf_in + 2, &p->cpx[1][0], sizeof( float ) * 2 * ( h - 1 ) );
			pffft_transform_ordered( p->pf, p->pf_in, p->real, p->pf_work, PFFFT_BACKWARD );
			}
		return;
		}
	F64Plan & f = *p->f64;
	if( p->r2c )
		{
		for( int i = 0; i < n; ++i ) f.work[i] = { (double) p->real[i], 0.0 };
		f.transform( -1 );
		for( int k = 0; k <= h; ++k )
			{
			p->cpx[k][0] = (float) f.work[k].real();
			p->cpx[k][1] = (float) f.work[k].imag();
			}
		}
	else
		{
		// Hermitian extension; Im of bins 0 and N/2 ignored (FFTW c2r behaviour).
		f.work[0] = { (double) p->cpx[0][0], 0.0 };
		if( n % 2 == 0 ) f.work[h] = { (double) p->cpx[h][0], 0.0 };
		for( int k = 1; k < ( n + 1 ) / 2; ++k )
			{
			const std::complex<double> v( (double) p->cpx[k][0], (double) p->cpx[k][1] );
			f.work[k] = v;
			f.work[n - k] = std::conj( v );
			}
		f.transform( +1 );
		for( int i = 0; i < n; ++i ) p->real[i] = (float) f.work[i].real();
		}
	}

void fftwf_destroy_plan( fftwf_plan p )
	{
	if( !p ) return;
	delete p->f64;
	if( p->pf ) pffft_destroy_setup( p->pf );
	if( p->pf_in ) pffft_aligned_free( p->pf_in );
	if( p->pf_out ) pffft_aligned_free( p->pf_out );
	if( p->pf_work ) pffft_aligned_free( p->pf_work );
	delete p;
	}

}
