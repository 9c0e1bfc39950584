// flan_b200/csrc/pv_generic.cu -- kernels and launchers of the any-size transform path (bodies: pv_generic_body.cuh).
#include "pv_generic.h"

namespace pvk {

namespace {

// The subset of pv_kernels.cu's DeviceEnv the generic bodies use.
struct GenericEnv
	{
	int tid;
	__device__ __forceinline__ void sync() { __syncthreads(); }
	__device__ __forceinline__ float ldg( const float * p ) { return __ldg( p ); }
	__device__ __forceinline__ float2 ldg2( const float2 * p ) { return __ldg( p ); }
	__device__ __forceinline__ void st_stream( float * p, float v ) { __stcs( p, v ); }
	__device__ __forceinline__ void red_add( float * p, float v ) { atomicAdd( p, v ); }
	};

__global__ void __launch_bounds__( 256 ) pv_generic_analysis_kernel( const GenericAnalysisArgs a )
	{
	extern __shared__ __align__( 16 ) unsigned char generic_smem[];
	GenericEnv env; env.tid = threadIdx.x;
	generic_analysis_cta( a, (int64_t) blockIdx.x, (int64_t) gridDim.x, (int) blockDim.x, env, reinterpret_cast<float2 *>( generic_smem ) );
	}

__global__ void __launch_bounds__( 256 ) pv_generic_synthesis_kernel( const GenericSynthArgs a )
	{
	extern __shared__ __align__( 16 ) unsigned char generic_smem[];
	GenericEnv env; env.tid = threadIdx.x;
	generic_synthesis_cta( a, (int64_t) blockIdx.x, (int64_t) gridDim.x, (int) blockDim.x, env, reinterpret_cast<float2 *>( generic_smem ) );
	}

} // namespace

GenericGeometry generic_geometry( const GenericFft & g, int64_t state_bytes, int64_t total_segments, int sms )
	{
	GenericGeometry geo{};
	// a thread per radix-4 butterfly of the largest pass, within [32, 256]
	int threads = 32;
	while( threads < 256 && threads * 4 < g.M ) threads <<= 1;
	geo.threads = threads;
	const int64_t fft_bytes = generic_fft_bytes( g );
	geo.fft_in_smem = fft_bytes <= 160 * 1024;
	geo.smem = geo.fft_in_smem ? (size_t) fft_bytes : 0;
	geo.scratch_stride = generic_align16( state_bytes + ( geo.fft_in_smem ? 0 : fft_bytes ) );
	int per_sm = 1;
	if( geo.fft_in_smem ) { per_sm = (int)( ( 200 * 1024 ) / ( fft_bytes + 1024 ) ); if( per_sm > 2048 / threads ) per_sm = 2048 / threads; if( per_sm > 8 ) per_sm = 8; if( per_sm < 1 ) per_sm = 1; }
	int64_t blocks = (int64_t) sms * per_sm;
	const int64_t budget = (int64_t) 1 << 30;                        // scratch for the whole launch
	if( blocks * geo.scratch_stride > budget ) blocks = budget / geo.scratch_stride;
	if( blocks > total_segments ) blocks = total_segments;
	if( blocks < 1 ) blocks = 1;
	geo.blocks = blocks;
	return geo;
	}

cudaError_t launch_generic_analysis( const GenericAnalysisArgs & a, const GenericGeometry & geo, cudaStream_t st )
	{
	cudaError_t e = cudaFuncSetAttribute( pv_generic_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) geo.smem );
	if( e != cudaSuccess ) return e;
	pv_generic_analysis_kernel<<<(unsigned) geo.blocks, geo.threads, geo.smem, st>>>( a );
	return cudaGetLastError();
	}

cudaError_t launch_generic_synthesis( const GenericSynthArgs & a, const GenericGeometry & geo, cudaStream_t st )
	{
	cudaError_t e = cudaFuncSetAttribute( pv_generic_synthesis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) geo.smem );
	if( e != cudaSuccess ) return e;
	pv_generic_synthesis_kernel<<<(unsigned) geo.blocks, geo.threads, geo.smem, st>>>( a );
	return cudaGetLastError();
	}

} // namespace pvk
