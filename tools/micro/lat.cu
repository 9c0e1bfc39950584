// Dependent-chain latency of scalar vs packed FP32 (one warp), and issue rate with k independent chains per warp
// at 4 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template<int MODE> __global__ void lat( float * out, long long * cyc, float a, float b )
	{
	float x = threadIdx.x; float2 p = make_float2( x, x + 1 );
	const float2 A = make_float2( a, a ), B = make_float2( b, b );
	long long t0 = clock64();
#pragma unroll
	for( int i = 0; i < 512; ++i )
		{
		if( MODE == 0 ) x = fmaf( x, a, b );
		if( MODE == 1 ) p = __ffma2_rn( p, A, B );
		if( MODE == 2 ) p = __fadd2_rn( p, A );
		if( MODE == 3 ) x = __fadd_rn( x, a );
		if( MODE == 4 ) x = fmaxf( x, a ) * b;
		}
	long long t1 = clock64();
	out[threadIdx.x] = x + p.x + p.y;
	if( threadIdx.x == 0 ) cyc[0] = t1 - t0;
	}
int main()
	{
	float * out; long long * cyc, h;
	cudaMalloc( &out, 4096 ); cudaMalloc( &cyc, 8 );
	const char * names[] = { "FFMA", "FFMA2", "FADD2", "FADD", "FMNMX+FMUL" };
	lat<0><<<1, 32>>>( out, cyc, 1.0001f, 0.5f ); cudaMemcpy( &h, cyc, 8, cudaMemcpyDeviceToHost ); printf( "%s chain: %.2f cyc/op\n", names[0], h / 512.0 );
	lat<1><<<1, 32>>>( out, cyc, 1.0001f, 0.5f ); cudaMemcpy( &h, cyc, 8, cudaMemcpyDeviceToHost ); printf( "%s chain: %.2f cyc/op\n", names[1], h / 512.0 );
	lat<2><<<1, 32>>>( out, cyc, 1.0001f, 0.5f ); cudaMemcpy( &h, cyc, 8, cudaMemcpyDeviceToHost ); printf( "%s chain: %.2f cyc/op\n", names[2], h / 512.0 );
	lat<3><<<1, 32>>>( out, cyc, 1.0001f, 0.5f ); cudaMemcpy( &h, cyc, 8, cudaMemcpyDeviceToHost ); printf( "%s chain: %.2f cyc/op\n", names[3], h / 512.0 );
	lat<4><<<1, 32>>>( out, cyc, 1.0001f, 0.5f ); cudaMemcpy( &h, cyc, 8, cudaMemcpyDeviceToHost ); printf( "%s chain: %.2f cyc/pair\n", names[4], h / 512.0 );
	return 0;
	}
