// Microbenchmark: does fma.rn.f32x2 (sm_100 packed FP32) double FP32 work per issue slot?
#include <cstdio>
#include <cuda_runtime.h>

__global__ void scalar_fma( float * out, float a, float b, int iters )
	{
	float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for( int i = 0; i < iters; ++i )
		{
		x0 = fmaf( x0, a, b ); x1 = fmaf( x1, a, b ); x2 = fmaf( x2, a, b ); x3 = fmaf( x3, a, b );
		x4 = fmaf( x4, a, b ); x5 = fmaf( x5, a, b ); x6 = fmaf( x6, a, b ); x7 = fmaf( x7, a, b );
		}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
	}

__device__ __forceinline__ void fma2( float2 & x, const float2 & a, const float2 & b )
	{
	unsigned long long xx, aa, bb;
	xx = *reinterpret_cast<unsigned long long *>( &x );
	aa = *reinterpret_cast<const unsigned long long *>( &a );
	bb = *reinterpret_cast<const unsigned long long *>( &b );
	asm( "fma.rn.f32x2 %0, %1, %2, %3;" : "=l"( xx ) : "l"( xx ), "l"( aa ), "l"( bb ) );
	x = *reinterpret_cast<float2 *>( &xx );
	}

__global__ void packed_fma( float * out, float a, float b, int iters )
	{
	float2 A = make_float2( a, a ), B = make_float2( b, b );
	float2 x0 = make_float2( threadIdx.x, threadIdx.x + 1 ), x1 = make_float2( threadIdx.x + 2, threadIdx.x + 3 ),
	       x2 = make_float2( threadIdx.x + 4, threadIdx.x + 5 ), x3 = make_float2( threadIdx.x + 6, threadIdx.x + 7 );
	float2 x4 = x0, x5 = x1, x6 = x2, x7 = x3;
	for( int i = 0; i < iters; ++i )
		{
		fma2( x0, A, B ); fma2( x1, A, B ); fma2( x2, A, B ); fma2( x3, A, B );
		fma2( x4, A, B ); fma2( x5, A, B ); fma2( x6, A, B ); fma2( x7, A, B );
		}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y + x4.x + x4.y + x5.x + x5.y + x6.x + x6.y + x7.x + x7.y;
	}

int main()
	{
	float * out; cudaMalloc( &out, 148 * 8 * 256 * sizeof( float ) );
	cudaEvent_t e0, e1; cudaEventCreate( &e0 ); cudaEventCreate( &e1 );
	const int iters = 20000, blocks = 148 * 8, threads = 256;
	for( int rep = 0; rep < 2; ++rep )
		{
		cudaEventRecord( e0 ); scalar_fma<<<blocks, threads>>>( out, 1.0001f, 0.5f, iters ); cudaEventRecord( e1 ); cudaEventSynchronize( e1 );
		float ms; cudaEventElapsedTime( &ms, e0, e1 );
		double fl = 2.0 * 8 * iters * (double) blocks * threads;
		printf( "scalar FFMA : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9 );
		cudaEventRecord( e0 ); packed_fma<<<blocks, threads>>>( out, 1.0001f, 0.5f, iters ); cudaEventRecord( e1 ); cudaEventSynchronize( e1 );
		cudaEventElapsedTime( &ms, e0, e1 );
		fl = 2.0 * 16 * iters * (double) blocks * threads;
		printf( "packed FFMA2: %.3f ms  %.1f TFLOP/s  (%s)\n", ms, fl / ms / 1e9, cudaGetErrorString( cudaGetLastError() ) );
		}
	return 0;
	}
