// Microbenchmark: per-SM issue rates of the FP32 instruction forms the phase-vocoder kernels are made of.
// Prints warp-instructions per clock per SM for each mix (148 SMs x 8 CTAs x 256 threads, 8 independent chains).
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
template<int MODE> __global__ void k( float * out, float a, float b, int iters )
	{
	float x[CHAINS]; float2 p[CHAINS]; int n[CHAINS];
#pragma unroll
	for( int i = 0; i < CHAINS; ++i ) { x[i] = threadIdx.x + i; p[i] = make_float2( x[i], x[i] + 1 ); n[i] = threadIdx.x * i; }
	const float2 A = make_float2( a, a ), B = make_float2( b, b );
	for( int it = 0; it < iters; ++it )
		{
#pragma unroll
		for( int i = 0; i < CHAINS; ++i )
			{
			if( MODE == 0 ) x[i] = fmaf( x[i], a, b );                         // FFMA
			if( MODE == 1 ) x[i] = __fadd_rn( x[i], a );                       // FADD
			if( MODE == 2 ) x[i] = __fmul_rn( x[i], a );                       // FMUL
			if( MODE == 3 ) p[i] = __ffma2_rn( p[i], A, B );                   // FFMA2
			if( MODE == 4 ) p[i] = __fadd2_rn( p[i], A );                      // FADD2
			if( MODE == 5 ) { if( i & 1 ) x[i] = fmaf( x[i], a, b ); else x[i] = __fadd_rn( x[i], a ); }   // FFMA + FADD
			if( MODE == 6 ) { if( i & 1 ) x[i] = fmaf( x[i], a, b ); else n[i] = ( n[i] ^ it ) + 3; }      // FFMA + ALU (LOP3/IADD)
			if( MODE == 7 ) { if( i & 1 ) p[i] = __ffma2_rn( p[i], A, B ); else x[i] = __fadd_rn( x[i], a ); } // FFMA2 + FADD
			if( MODE == 8 ) { if( i & 1 ) p[i] = __fadd2_rn( p[i], A ); else x[i] = fmaxf( x[i] * 0.5f, a ); }       // FADD2 + FMNMX/FMUL
			if( MODE == 9 ) x[i] = fmaxf( x[i], a + it );                     // FMNMX (alu) + FADD
			if( MODE == 10 ) { if( i & 1 ) p[i] = __ffma2_rn( p[i], A, B ); else n[i] = ( n[i] ^ it ) + 3; }   // FFMA2 + ALU
			if( MODE == 11 ) x[i] = fmaf( x[i], 1.0001f, 0.5f );               // FFMA imm
			}
		}
	float s = 0; 
#pragma unroll
	for( int i = 0; i < CHAINS; ++i ) s += x[i] + p[i].x + p[i].y + n[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	}

template<int MODE> void run( const char * name, float * out, double instr_per_iter )
	{
	cudaEvent_t e0, e1; cudaEventCreate( &e0 ); cudaEventCreate( &e1 );
	const int iters = 20000, blocks = 148 * 8, threads = 256;
	float ms = 0;
	for( int rep = 0; rep < 2; ++rep )
		{
		cudaEventRecord( e0 ); k<MODE><<<blocks, threads>>>( out, 1.0001f, 0.5f, iters ); cudaEventRecord( e1 ); cudaEventSynchronize( e1 );
		cudaEventElapsedTime( &ms, e0, e1 );
		}
	const double warp_instr = instr_per_iter * iters * (double) blocks * threads / 32.0;
	printf( "%-28s %.3f ms  %.2f warp-instr/clk/SM (at 1.965 GHz)  err=%s\n", name, ms, warp_instr / ( ms * 1e-3 ) / 1.965e9 / 148.0, cudaGetErrorString( cudaGetLastError() ) );
	}

int main()
	{
	float * out; cudaMalloc( &out, 148 * 8 * 256 * sizeof( float ) );
	run<0>( "FFMA", out, 8 ); run<1>( "FADD", out, 8 ); run<2>( "FMUL", out, 8 ); run<3>( "FFMA2", out, 8 ); run<4>( "FADD2", out, 8 );
	run<5>( "FFMA+FADD", out, 8 ); run<6>( "FFMA+LOP3+IADD (12/iter)", out, 12 ); run<7>( "FFMA2+FADD", out, 8 );
	run<8>( "FADD2+FMUL+FMNMX (12/iter)", out, 12 ); run<9>( "FMNMX+FADD (16/iter)", out, 16 ); run<10>( "FFMA2+LOP3+IADD (12/iter)", out, 12 );
	run<11>( "FFMA imm", out, 8 );
	return 0;
	}
