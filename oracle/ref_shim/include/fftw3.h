// oracle/ref_shim: TEST INFRASTRUCTURE ONLY.
// Declares the seven FFTW3 single-precision entry points the reference's
// FFTHelper.cpp uses (FFTHelper.cpp:21-24 alloc/plan, :32-35 destroy/free,
// :41,:47 execute). FFTW itself (libfftw3f, version unpinned by the reference)
// is not vendored under /root/reference and not installed in this image; the
// definitions live in ref_shim/fftw_standin.cpp.
#pragma once
#include <cstddef>

extern "C" {
typedef float fftwf_complex[2];
struct fftwf_plan_s;
typedef struct fftwf_plan_s * fftwf_plan;

#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

float *         fftwf_alloc_real( size_t n );
fftwf_complex * fftwf_alloc_complex( size_t n );
void            fftwf_free( void * p );
fftwf_plan      fftwf_plan_dft_r2c_1d( int n, float * in, fftwf_complex * out, unsigned flags );
fftwf_plan      fftwf_plan_dft_c2r_1d( int n, fftwf_complex * in, float * out, unsigned flags );
void            fftwf_execute( const fftwf_plan p );
void            fftwf_destroy_plan( fftwf_plan p );
}
