// flan_b200/csrc/pv_kernels.cu -- sm_100a kernels of the phase-vocoder engine and their launchers.
//
//   pv_analysis_kernel<N>    Audio::convert_to_PV     (reference Conversions/AudioPV.cpp:12-78)
//   pv_phase_seg_kernel      per-segment phase-sum summaries     (phase_vocoder.cpp:57-59, scan form)
//   pv_phase_scan_kernel     exclusive scan of the summaries over segments -> accumulator per segment
//   pv_synthesis_kernel<N>   PV::convert_to_audio     (AudioPV.cpp:86-139)
//   pv_mid_side_kernel       Audio::convert_to_mid_side (Audio/AudioConversions.cpp:32-51)
//
// The CTA bodies live in pv_body.cuh (shared with the CPU thread emulator); this file supplies the
// device Env (barrier, cp.async staging, cache-hinted loads/stores, red.add) and the launch geometry.
#include "pv_body.cuh"
#include "pv_launch.h"

#include <cuda_runtime.h>

#include <cstdlib>

namespace pvk {

struct DeviceEnv
	{
	int tid;
	__device__ __forceinline__ void syncwarp() { __syncwarp(); }
	__device__ __forceinline__ bool any( bool p ) { return __any_sync( 0xffffffffu, p ); }
	__device__ __forceinline__ void sync()
		{
#ifndef PV_ABL_NOSYNC
		__syncthreads();
#endif
		}
	__device__ __forceinline__ float ldg( const float * p ) { return __ldg( p ); }
	__device__ __forceinline__ float2 ldg2( const float2 * p ) { return __ldg( p ); }
	__device__ __forceinline__ float4 ldg4( const float4 * p ) { return __ldg( p ); }
	__device__ __forceinline__ float2 ldcs2( const float2 * p ) { return __ldcs( p ); }
	// streaming stores that do not allocate in L1 (STG.E.NA): the PV rows and output samples would otherwise push the
	// window samples and the tables out of it (measured against st.global.cs: analysis 1.685 -> 1.660 ms on cfg2)
	__device__ __forceinline__ void st_stream2( float2 * p, float2 v )
		{ asm volatile( "st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" :: "l"( p ), "f"( v.x ), "f"( v.y ) : "memory" ); }
	__device__ __forceinline__ void st_stream( float * p, float v ) { __stcs( p, v ); }
	__device__ __forceinline__ void red_add( float * p, float v ) { atomicAdd( p, v ); }
	// shared-memory word += v without a round trip through registers (the word is private to the calling thread)
	__device__ __forceinline__ void shared_add( int * p, int v ) { asm volatile( "red.shared.add.s32 [%0], %1;" :: "r"( (unsigned) __cvta_generic_to_shared( p ) ), "r"( v ) : "memory" ); }
	// 8-byte asynchronous global->shared copy (LDGSTS), completion tracked per thread by commit / wait groups
	__device__ __forceinline__ void cp_async8( float2 * dst, const float2 * src )
		{
		const unsigned d = (unsigned) __cvta_generic_to_shared( dst );
		asm volatile( "cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"( d ), "l"( src ) : "memory" );
		}
	__device__ __forceinline__ void cp_async_commit() { asm volatile( "cp.async.commit_group;\n" ::: "memory" ); }
	__device__ __forceinline__ void cp_async_wait_all() { asm volatile( "cp.async.wait_group 0;\n" ::: "memory" ); }
	// Bulk (TMA) global->shared copy tracked by an mbarrier: one thread arms the barrier with the byte count and issues
	// the copy; every thread waits on the phase parity. Addresses and size are multiples of 16 bytes.
	typedef unsigned long long BulkBarrier;
	__device__ __forceinline__ void bulk_init( BulkBarrier * bar )
		{
		const unsigned b = (unsigned) __cvta_generic_to_shared( bar );
		asm volatile( "mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"( b ) : "memory" );
		asm volatile( "fence.mbarrier_init.release.cluster;\n" ::: "memory" );
		}
	__device__ __forceinline__ void bulk_load( void * dst, const void * src, unsigned bytes, BulkBarrier * bar )
		{
		const unsigned b = (unsigned) __cvta_generic_to_shared( bar );
		const unsigned d = (unsigned) __cvta_generic_to_shared( dst );
		asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"( b ), "r"( bytes ) : "memory" );
		asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
		              :: "r"( d ), "l"( src ), "r"( bytes ), "r"( b ) : "memory" );
		}
	__device__ __forceinline__ void bulk_wait( BulkBarrier * bar, unsigned parity )
		{
		const unsigned b = (unsigned) __cvta_generic_to_shared( bar );
		asm volatile(
			"{\n"
			".reg .pred p;\n"
			"WAIT_%=:\n"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
			"@p bra DONE_%=;\n"
			"bra WAIT_%=;\n"
			"DONE_%=:\n"
			"}\n" :: "r"( b ), "r"( parity ) : "memory" );
		}
#ifdef PV_PREFETCH_L1
	__device__ __forceinline__ void prefetch( const void * p ) { asm volatile( "prefetch.global.L1 [%0];" :: "l"( p ) ); }
#else
	__device__ __forceinline__ void prefetch( const void * p ) { asm volatile( "prefetch.global.L2 [%0];" :: "l"( p ) ); }
#endif
	};

// TPS = resident threads per SM the register allocation is sized for (512 -> 128, 768 -> 85, 1024 -> 64 registers).
constexpr int min_blocks( int threads, int TPS ) { return TPS / threads > 32 ? 32 : ( TPS / threads > 0 ? TPS / threads : 1 ); }

// PT = complex points per thread: 8 (radix-8 passes, N/16 threads per frame) or 16 (radix-16 passes, N/32 threads).
template<int N, int PT, int TPS, bool ONE, bool PAD = false, bool EMIT = false>
__global__ void __launch_bounds__( N / ( 2 * PT ), min_blocks( N / ( 2 * PT ), TPS ) ) pv_analysis_kernel( const AnalysisArgs a )
	{
	extern __shared__ __align__( 16 ) unsigned char smem_raw[];
	float2 * x0 = reinterpret_cast<float2 *>( smem_raw );
	float2 * x1 = ONE ? x0 : x0 + XBuf<N / 2>::size;
	int * ssum = EMIT ? reinterpret_cast<int *>( x1 + XBuf<N / 2>::size ) : nullptr;      // N/2 + 1 per-bin sums (EMIT only)
	DeviceEnv env; env.tid = threadIdx.x;
	analysis_cta<N, PT, ONE, PAD, EMIT>( a, (int64_t) blockIdx.x, env, x0, x1, ssum );
	}

// Mirrored last pass (pv_body.cuh: analysis_cta_mirror): 16 points per thread, unpack + phase vocoder on the thread's own registers.
template<int N, int TPS>
__global__ void __launch_bounds__( N / 32, min_blocks( N / 32, TPS ) ) pv_analysis_mirror_kernel( const AnalysisArgs a )
	{
	extern __shared__ __align__( 16 ) unsigned char smem_raw[];
	float2 * x0 = reinterpret_cast<float2 *>( smem_raw );
	float2 * x1 = a.one_buffer ? x0 : x0 + XBuf<N / 2>::size;
	float2 * ring = x1 + XBuf<N / 2>::size;            // N/2 pairs: the window's samples, thread-private entries
	float2 * scratch = ring + N / 2;                   // 16 pairs
	DeviceEnv env; env.tid = threadIdx.x;
	analysis_cta_mirror<N>( a, (int64_t) blockIdx.x, env, x0, x1, ring, scratch );
	}

template<int N, int TPS, bool ONE>
__global__ void __launch_bounds__( N / 16, min_blocks( N / 16, TPS ) ) pv_synthesis_kernel( const SynthArgs a )
	{
	extern __shared__ __align__( 16 ) unsigned char smem_raw[];
	float * ola = reinterpret_cast<float *>( smem_raw );
	float2 * x0 = reinterpret_cast<float2 *>( smem_raw + sizeof( float ) * N );
	float2 * x1 = ONE ? x0 : x0 + XBuf<N / 2>::size;
	float2 * rowbuf = x1 + XBuf<N / 2>::size;
	DeviceEnv env; env.tid = threadIdx.x;
	synthesis_cta<N, ONE>( a, (int64_t) blockIdx.x, env, ola, x0, x1, rowbuf );
	}

// Mirrored first pass (pv_body.cuh: synthesis_cta_mirror): 16 points per thread, thread-private row FIFO and overlap-add ring.
template<int N, int TPS, bool ONE, bool GEN>
__global__ void __launch_bounds__( N / 32, min_blocks( N / 32, TPS ) ) pv_synthesis_mirror_kernel( const SynthArgs a )
	{
	extern __shared__ __align__( 16 ) unsigned char smem_raw[];
	float2 * ring = reinterpret_cast<float2 *>( smem_raw );                 // N/2 pairs
	float2 * x0 = ring + N / 2;
	float2 * x1 = ONE ? x0 : x0 + XBuf<N / 2>::size;
	float2 * rowbuf = x1 + XBuf<N / 2>::size;                               // N/2 + 2 pairs
	__shared__ __align__( 8 ) DeviceEnv::BulkBarrier bar;
	DeviceEnv env; env.tid = threadIdx.x;
	synthesis_cta_mirror<N, ONE, GEN>( a, (int64_t) blockIdx.x, env, ring, x0, x1, rowbuf, &bar );
	}

// One thread per (channel, segment, bin): summary of the segment's phase increments.
__global__ void __launch_bounds__( 256 ) pv_phase_seg_kernel( const PhaseSegArgs a )
	{
	const int bchunks = ( a.B + 255 ) / 256;
	const int64_t blk = blockIdx.x;
	const int b = (int)( blk % bchunks ) * 256 + threadIdx.x;
	const int seg = (int)( ( blk / bchunks ) % a.segs_per_channel );
	const int c = (int)( blk / ( (int64_t) bchunks * a.segs_per_channel ) );
	if( b >= a.B ) return;
	const int64_t fa = a.frame_begin + (int64_t) seg * a.seg_len;
	const int64_t fb = ( fa + a.seg_len < a.frame_end ) ? fa + a.seg_len : a.frame_end;
	const float2 * col = a.pv + (int64_t) c * a.pv_channel_stride + ( fa - a.frame_begin ) * (int64_t) a.B + b;
	int flag = 0;
	const PhaseSeg s = phase_segment_summary( col, (int64_t) a.B, fb - fa, a.k, a.P, a.rcpP, flag,
		[]( const float2 * p ) { return __ldg( p ); } );
	PhaseSeg * dst = a.seg_out + ( (int64_t) c * a.segs_per_channel + seg ) * a.B + b;
	*dst = s;
	if( flag ) *a.nan_flag = 1;
	}

// Summaries left by the analysis kernel (analysis_cta<EMIT>): the entries it marked -- always the lowest bins, whose
// expected phase advance is too small for the 32-bit form; every bin of a segment of digital silence, whose frequencies
// are nowhere near their bins' centres; NaN / Inf -- are recomputed from the rows, one thread per (channel, segment, bin),
// all at once: unmarked entries cost a 8-byte read, a fully marked buffer what pv_phase_seg_kernel costs. (The scan's
// group reduction repairs what a caller-limited launch leaves.)
__global__ void __launch_bounds__( 64 ) pv_phase_fix_kernel( const PhaseScanArgs a, int bins )
	{
	const int b = threadIdx.x + blockIdx.z * 64;
	const int s = blockIdx.x, c = blockIdx.y;
	if( b >= bins || b >= a.B ) return;
	PhaseSeg * e = const_cast<PhaseSeg *>( a.seg ) + ( (int64_t) c * a.segs_per_channel + s ) * a.B + b;
	if( !( e->sum.q != e->sum.q ) ) return;
	const int64_t fa = a.fix_frame_begin + (int64_t) s * a.fix_seg_len;
	const int64_t fb = ( fa + a.fix_seg_len < a.fix_frame_end ) ? fa + a.fix_seg_len : a.fix_frame_end;
	const float2 * col = a.fix_pv + (int64_t) c * a.fix_channel_stride + ( fa - a.fix_frame_begin ) * (int64_t) a.B + b;
	int flag = 0;
	*e = phase_segment_summary( col, (int64_t) a.B, fb - fa, a.fix_k, a.P, a.rcpP, flag, []( const float2 * p ) { return __ldg( p ); } );
	if( flag ) *a.fix_nan_flag = 1;
	}

// One thread per (channel, bin): serial walk over the segments (a few thousand at most), writing the
// accumulator value that enters each segment. carry_in/carry_out chain frame-range shards across GPUs.
// The scan over segments runs in three short phases so that its serial depth is group_len + groups + group_len
// instead of segs_per_channel (the combine is associative):
//   mode 0  reduce each group of consecutive segments to one state          thread per (channel, group, bin)
//   mode 1  exclusive scan of the group states (carry_in / carry_out here)  thread per (channel, bin)
//   mode 2  re-walk each group from its prefix, writing acc_start           thread per (channel, group, bin)
__global__ void __launch_bounds__( 128 ) pv_phase_scan_kernel( const PhaseScanArgs a, int mode )
	{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	const int g = blockIdx.y;
	const int c = blockIdx.z;
	if( b >= a.B ) return;
	PhaseSeg st; st.sum.q = 0.0; st.sum.r = 0.0; st.mx.q = 0.0; st.mx.r = 0.0;
	PhaseSeg * grp = a.group + (int64_t) c * a.groups * a.B + b;
	if( mode == 1 )
		{
		if( a.carry_in ) st = a.carry_in[(int64_t) c * a.B + b];
		for( int i = 0; i < a.groups; ++i )
			{
			const PhaseSeg tmp = grp[(int64_t) i * a.B];
			grp[(int64_t) i * a.B] = st;
			phase_state_combine( st, tmp, a.P, a.rcpP );
			}
		if( a.carry_out ) a.carry_out[(int64_t) c * a.B + b] = st;
		return;
		}
	const int s0 = g * a.group_len;
	const int s1 = ( s0 + a.group_len < a.segs_per_channel ) ? s0 + a.group_len : a.segs_per_channel;
	const PhaseSeg * src = a.seg + (int64_t) c * a.segs_per_channel * a.B + b;
	if( mode == 0 )
		{
		for( int s = s0; s < s1; ++s )
			{
			PhaseSeg v = src[(int64_t) s * a.B];
			if( a.fix_pv && is_nan_marker( v ) )
				{
				const int64_t fa = a.fix_frame_begin + (int64_t) s * a.fix_seg_len;
				const int64_t fb = ( fa + a.fix_seg_len < a.fix_frame_end ) ? fa + a.fix_seg_len : a.fix_frame_end;
				const float2 * col = a.fix_pv + (int64_t) c * a.fix_channel_stride + ( fa - a.fix_frame_begin ) * (int64_t) a.B + b;
				int flag = 0;
				v = phase_segment_summary( col, (int64_t) a.B, fb - fa, a.fix_k, a.P, a.rcpP, flag, []( const float2 * p ) { return __ldg( p ); } );
				if( flag ) *a.fix_nan_flag = 1;
				const_cast<PhaseSeg *>( src )[(int64_t) s * a.B] = v;
				}
			phase_state_combine( st, v, a.P, a.rcpP );
			}
		grp[(int64_t) g * a.B] = st;
		return;
		}
	st = grp[(int64_t) g * a.B];
	if( a.expand_only && a.expand_carry )
		{
		PhaseSeg cin = a.expand_carry[(int64_t) c * a.B + b];
		phase_state_combine( cin, st, a.P, a.rcpP );
		st = cin;
		}
	double * dst = a.acc_start + (int64_t) c * a.segs_per_channel * a.B + b;
	for( int s = s0; s < s1; ++s )
		{
		dst[(int64_t) s * a.B] = phase_state_value( st, a.P );
		phase_state_combine( st, src[(int64_t) s * a.B], a.P, a.rcpP );
		}
	}

// The same scan for signals of a few hundred segments per channel, where the three launches above are all latency:
// one launch, a CTA per (channel, 32 bins), 16 thread rows that each reduce a contiguous run of segments, one serial
// scan over the 16 run totals in shared memory, then every row re-walks its run from its prefix. Serial depth
// 2 * ceil(segs / 16) + 16 instead of segs.
__global__ void __launch_bounds__( 512 ) pv_phase_scan_small_kernel( const PhaseScanArgs a )
	{
	constexpr int G = 16;
	__shared__ PhaseSeg tot[G][32];
	const int lane = threadIdx.x, g = threadIdx.y;
	const int b = blockIdx.x * 32 + lane;
	const int c = blockIdx.y;
	const bool live = b < a.B;
	const int per = ( a.segs_per_channel + G - 1 ) / G;
	const int s0 = g * per;
	const int s1 = ( s0 + per < a.segs_per_channel ) ? s0 + per : a.segs_per_channel;
	const PhaseSeg * src = a.seg + (int64_t) c * a.segs_per_channel * a.B + b;
	PhaseSeg st; st.sum.q = 0.0; st.sum.r = 0.0; st.mx.q = 0.0; st.mx.r = 0.0;
	if( live ) for( int s = s0; s < s1; ++s ) phase_state_combine( st, src[(int64_t) s * a.B], a.P, a.rcpP );
	tot[g][lane] = st;
	__syncthreads();
	if( g == 0 && live )
		{
		PhaseSeg run; run.sum.q = 0.0; run.sum.r = 0.0; run.mx.q = 0.0; run.mx.r = 0.0;
		if( a.carry_in ) run = a.carry_in[(int64_t) c * a.B + b];
		for( int i = 0; i < G; ++i )
			{
			const PhaseSeg tmp = tot[i][lane];
			tot[i][lane] = run;
			phase_state_combine( run, tmp, a.P, a.rcpP );
			}
		if( a.carry_out ) a.carry_out[(int64_t) c * a.B + b] = run;
		}
	__syncthreads();
	if( !live || !a.acc_start ) return;
	st = tot[g][lane];
	double * dst = a.acc_start + (int64_t) c * a.segs_per_channel * a.B + b;
	for( int s = s0; s < s1; ++s )
		{
		dst[(int64_t) s * a.B] = phase_state_value( st, a.P );
		phase_state_combine( st, src[(int64_t) s * a.B], a.P, a.rcpP );
		}
	}

// carry[r] = summaries[0] (+) ... (+) summaries[r-1], for this rank r.
__global__ void __launch_bounds__( 128 ) pv_phase_carry_kernel( const PhaseSeg * all, int rank, int64_t per_rank, PhaseSeg * carry, double P, double rcpP )
	{
	const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if( i >= per_rank ) return;
	PhaseSeg st; st.sum.q = 0.0; st.sum.r = 0.0; st.mx.q = 0.0; st.mx.r = 0.0;
	for( int r = 0; r < rank; ++r ) phase_state_combine( st, all[(int64_t) r * per_rank + i], P, rcpP );
	carry[i] = st;
	}

// Audio::convert_to_mid_side (AudioConversions.cpp:42-49): (L +- R) / sqrt(2.0f), IEEE division.
__global__ void __launch_bounds__( 256 ) pv_mid_side_kernel( const float * in, float * out, int64_t n )
	{
	const float sqrt2 = 1.41421356237309504880f;
	for( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x )
		{
		const float l = in[i], r = in[n + i];
		out[i]     = __fdiv_rn( __fadd_rn( l, r ), sqrt2 );
		out[n + i] = __fdiv_rn( __fsub_rn( l, r ), sqrt2 );
		}
	}

// Clears the output samples the resynthesis kernels reach with red.add (those shared by two segments, and the head and
// tail of the local span) instead of the whole output: every other sample is written exactly once by a plain store.
// Region j of a channel lies between the interior of segment j-1 and the interior of segment j (pv_body.cuh:
// interior_lo = hop*fa + W/2 - hop, interior_hi = hop*fb - W/2; an empty interior counts as the point interior_lo).
// Requires hop <= window (no gaps between the frames' windows).
__global__ void __launch_bounds__( 256 ) pv_zero_shared_kernel( float * out, int64_t out_stride, int64_t out_offset, int64_t out_len,
                                                                 int64_t frame_begin, int64_t frame_end, int seg_len, int segs, int W, int hop )
	{
	const int j = blockIdx.x;            // 0 .. segs
	const int c = blockIdx.y;
	const int half = W / 2;
	auto seg_lo = [&]( int s ) { return (int64_t) hop * ( frame_begin + (int64_t) s * seg_len ) + half - hop; };
	auto seg_hi = [&]( int s )
		{
		const int64_t fa = frame_begin + (int64_t) s * seg_len;
		const int64_t fb = ( fa + seg_len < frame_end ) ? fa + seg_len : frame_end;
		const int64_t hi = (int64_t) hop * fb - half, lo = seg_lo( s );
		return hi > lo ? hi : lo;
		};
	int64_t lo = ( j == 0 ) ? out_offset : seg_hi( j - 1 );
	int64_t hi = ( j == segs ) ? out_offset + out_len : seg_lo( j );
	if( lo < out_offset ) lo = out_offset;
	if( hi > out_offset + out_len ) hi = out_offset + out_len;
	float * dst = out + (int64_t) c * out_stride - out_offset;
	for( int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x ) dst[i] = 0.0f;
	}

cudaError_t launch_zero_shared( float * out, int64_t out_stride, int64_t out_offset, int64_t out_len, int C,
                                int64_t frame_begin, int64_t frame_end, int seg_len, int segs, int W, int hop, cudaStream_t st )
	{
	const dim3 grid( (unsigned)( segs + 1 ), (unsigned) C );
	pv_zero_shared_kernel<<<grid, 256, 0, st>>>( out, out_stride, out_offset, out_len, frame_begin, frame_end, seg_len, segs, W, hop );
	return cudaGetLastError();
	}

// out[i] += add[i] (overlap-add halo received from a neighbouring frame-range shard)
__global__ void __launch_bounds__( 256 ) pv_add_kernel( float * out, const float * add, int64_t n )
	{
	for( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x )
		out[i] = __fadd_rn( out[i], add[i] );
	}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
// Development aids (FLAN_B200_DEBUG builds only): FLAN_B200_SMEM_PAD=<bytes> inflates the dynamic shared memory request
// so that fewer CTAs fit per SM; FLAN_B200_CARVEOUT=<percent> sets the preferred shared-memory carveout of the transforms.
#ifdef FLAN_B200_DEBUG
static int carveout()
	{
	static const int v = [] { const char * e = std::getenv( "FLAN_B200_CARVEOUT" ); return e ? std::atoi( e ) : -1; }();
	return v;
	}
template<class K> static void apply_carveout( K kernel )
	{
	if( carveout() >= 0 ) cudaFuncSetAttribute( kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout() );
	}
static size_t smem_pad()
	{
	static const size_t pad = [] { const char * e = std::getenv( "FLAN_B200_SMEM_PAD" ); return e ? (size_t) std::atol( e ) : (size_t) 0; }();
	return pad;
	}
#else
template<class K> static void apply_carveout( K ) {}
static constexpr size_t smem_pad() { return 0; }
#endif

// blocks < 0 asks the launchers for the kernel's resident CTAs per SM instead of a launch (the pipelined host forms cut a
// signal into slices of whole waves); the answer is left here.
static thread_local int g_occupancy = 0;
template<class K> static cudaError_t report_occupancy( K kernel, int threads, size_t smem )
	{
	g_occupancy = 0;
	return cudaOccupancyMaxActiveBlocksPerMultiprocessor( &g_occupancy, kernel, threads, smem );
	}
int last_occupancy() { return g_occupancy; }

template<int N, int PT, int TPS, bool ONE> static cudaError_t launch_analysis_nto( const AnalysisArgs & a, int64_t blocks, cudaStream_t st )
	{
	const size_t smem = ( ONE ? 1 : 2 ) * sizeof( float2 ) * XBuf<N / 2>::size + smem_pad();
	cudaError_t e = cudaFuncSetAttribute( pv_analysis_kernel<N, PT, TPS, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem );
	if( e != cudaSuccess ) return e;
	apply_carveout( pv_analysis_kernel<N, PT, TPS, ONE> );
	// zero-padded windows that fill whole slots (the API default shape): vector loads, 16 points per thread form only
	if constexpr( PT == 16 && ONE )
		{
		if( a.W < N && a.W % ( N / PT ) == 0 && a.aligned2 )
			{
			e = cudaFuncSetAttribute( pv_analysis_kernel<N, PT, TPS, ONE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem );
			if( e != cudaSuccess ) return e;
			if( blocks < 0 ) return report_occupancy( pv_analysis_kernel<N, PT, TPS, ONE, true>, N / ( 2 * PT ), smem );
			pv_analysis_kernel<N, PT, TPS, ONE, true><<<(unsigned) blocks, N / ( 2 * PT ), smem, st>>>( a );
			return cudaGetLastError();
			}
		}
	// the full-window launch that also leaves the phase summaries of its rows (16 points per thread, one buffer, TPS 512: the
	// policy's variant from dft 2048 up)
	if constexpr( PT == 16 && ONE && TPS == 512 && N >= 2048 )
		{
		if( a.seg_out )
			{
			const size_t smem_emit = smem + sizeof( int ) * ( N / 2 + 4 );
			e = cudaFuncSetAttribute( pv_analysis_kernel<N, PT, TPS, ONE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_emit );
			if( e != cudaSuccess ) return e;
			apply_carveout( pv_analysis_kernel<N, PT, TPS, ONE, false, true> );
			if( blocks < 0 ) return report_occupancy( pv_analysis_kernel<N, PT, TPS, ONE, false, true>, N / ( 2 * PT ), smem_emit );
			pv_analysis_kernel<N, PT, TPS, ONE, false, true><<<(unsigned) blocks, N / ( 2 * PT ), smem_emit, st>>>( a );
			return cudaGetLastError();
			}
		}
	if( a.seg_out ) return cudaErrorInvalidValue;          // the caller asks for summaries only where this form exists
	if( blocks < 0 ) return report_occupancy( pv_analysis_kernel<N, PT, TPS, ONE>, N / ( 2 * PT ), smem );
	pv_analysis_kernel<N, PT, TPS, ONE><<<(unsigned) blocks, N / ( 2 * PT ), smem, st>>>( a );
	return cudaGetLastError();
	}
// the one-buffer form is built for 16 points per thread only (the variant the large transforms use)
template<int N, int PT, int TPS> static cudaError_t launch_analysis_nt( const AnalysisArgs & a, int64_t blocks, cudaStream_t st )
	{
#ifdef FLAN_B200_DEBUG
	if constexpr( PT == 16 ) { if( a.one_buffer ) return launch_analysis_nto<N, PT, TPS, true>( a, blocks, st ); }
	return launch_analysis_nto<N, PT, TPS, false>( a, blocks, st );
#else
	return launch_analysis_nto<N, PT, TPS, PT == 16>( a, blocks, st );      // the policy: one buffer exactly when 16 points per thread
#endif
	}
bool analysis_mirror_applies( int N, const AnalysisArgs & a )
	{
	return mirror_supported( N ) && a.W == N && a.hop == N / 16;
	}
template<int N, int TPS> static cudaError_t launch_analysis_mirror_nt( const AnalysisArgs & a, int64_t blocks, cudaStream_t st )
	{
	const size_t smem = ( a.one_buffer ? 1 : 2 ) * sizeof( float2 ) * XBuf<N / 2>::size + sizeof( float2 ) * ( N / 2 + 16 ) + smem_pad();
	cudaError_t e = cudaFuncSetAttribute( pv_analysis_mirror_kernel<N, TPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem );
	if( e != cudaSuccess ) return e;
	apply_carveout( pv_analysis_mirror_kernel<N, TPS> );
	if( blocks < 0 ) return report_occupancy( pv_analysis_mirror_kernel<N, TPS>, N / 32, smem );
	pv_analysis_mirror_kernel<N, TPS><<<(unsigned) blocks, N / 32, smem, st>>>( a );
	return cudaGetLastError();
	}
// Release builds instantiate only what the launch policy (pv_capi.cu: analysis_range) selects -- 8 points per thread at
// 85 registers, 16 points per thread / mirrored at 128 registers; FLAN_B200_DEBUG builds carry every register variant
// for the experiments under tools/experiments.
template<int N> static cudaError_t launch_analysis_n( const AnalysisArgs & a, int64_t blocks, cudaStream_t st, int tps, int pt )
	{
	if constexpr( N == 1024 || N == 2048 || N == 4096 )
		{
		if( pt == PV_PT_MIRROR && analysis_mirror_applies( N, a ) )
			{
#ifdef FLAN_B200_DEBUG
			if( tps >= 768 ) return launch_analysis_mirror_nt<N, 768>( a, blocks, st );
			if( tps >= 640 ) return launch_analysis_mirror_nt<N, 640>( a, blocks, st );
			if( tps < 512 ) return launch_analysis_mirror_nt<N, 384>( a, blocks, st );
#endif
			return launch_analysis_mirror_nt<N, 512>( a, blocks, st );
			}
		}
#ifdef FLAN_B200_DEBUG
	if constexpr( N >= 512 )
#else
	if constexpr( N >= 2048 )       // the policy asks for 16 points per thread from dft 2048 up only
#endif
		{
		if( pt == 16 )
			{
#ifdef FLAN_B200_DEBUG
			if( tps >= 768 ) return launch_analysis_nt<N, 16, 768>( a, blocks, st );
			if( tps >= 640 ) return launch_analysis_nt<N, 16, 640>( a, blocks, st );
			if( tps < 512 ) return launch_analysis_nt<N, 16, 384>( a, blocks, st );
#endif
			return launch_analysis_nt<N, 16, 512>( a, blocks, st );
			}
		}
#ifdef FLAN_B200_DEBUG
	if( tps >= 1024 ) return launch_analysis_nt<N, 8, 1024>( a, blocks, st );
	if( tps < 768 ) return launch_analysis_nt<N, 8, 512>( a, blocks, st );
#endif
	return launch_analysis_nt<N, 8, 768>( a, blocks, st );
	}
template<int N, int TPS, bool ONE> static cudaError_t launch_synthesis_nt( const SynthArgs & a, int64_t blocks, cudaStream_t st )
	{
	const size_t smem = sizeof( float ) * N + ( ONE ? 1 : 2 ) * sizeof( float2 ) * XBuf<N / 2>::size + sizeof( float2 ) * ( N / 2 + 2 ) + smem_pad();
	cudaError_t e = cudaFuncSetAttribute( pv_synthesis_kernel<N, TPS, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem );
	if( e != cudaSuccess ) return e;
	if( blocks < 0 ) return report_occupancy( pv_synthesis_kernel<N, TPS, ONE>, N / 16, smem );
	pv_synthesis_kernel<N, TPS, ONE><<<(unsigned) blocks, N / 16, smem, st>>>( a );
	return cudaGetLastError();
	}
template<int N, int TPS, bool ONE, bool GEN> static cudaError_t launch_synthesis_mirror_ntg( const SynthArgs & a, int64_t blocks, cudaStream_t st )
	{
	const size_t smem = sizeof( float ) * N + ( ONE ? 1 : 2 ) * sizeof( float2 ) * XBuf<N / 2>::size + sizeof( float2 ) * ( N / 2 + 2 ) + smem_pad();
	cudaError_t e = cudaFuncSetAttribute( pv_synthesis_mirror_kernel<N, TPS, ONE, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem );
	if( e != cudaSuccess ) return e;
	apply_carveout( pv_synthesis_mirror_kernel<N, TPS, ONE, GEN> );
	if( blocks < 0 ) return report_occupancy( pv_synthesis_mirror_kernel<N, TPS, ONE, GEN>, N / 32, smem );
	pv_synthesis_mirror_kernel<N, TPS, ONE, GEN><<<(unsigned) blocks, N / 32, smem, st>>>( a );
	return cudaGetLastError();
	}
// the standard shape (window == dft, hop == dft/16: every BASELINE config) runs the compile-time form, any other aligned
// window / hop (the API default 2048 / 128 / 4096 among them) the general one
template<int N, int TPS, bool ONE> static cudaError_t launch_synthesis_mirror_nt( const SynthArgs & a, int64_t blocks, cudaStream_t st )
	{
	if( a.W == N && a.hop == N / 16 ) return launch_synthesis_mirror_ntg<N, TPS, ONE, false>( a, blocks, st );
	return launch_synthesis_mirror_ntg<N, TPS, ONE, true>( a, blocks, st );
	}
bool synthesis_mirror_applies( int N, const SynthArgs & a )
	{
	if( !synth_mirror_supported( N ) ) return false;
	if( a.W == N && a.hop == N / 16 ) return true;
	// general form: the window fills whole slots of the thread layout (a multiple of dft/16 samples), the hop is even and
	// at most one slot (dft/16 samples) and leaves no gaps
	return a.W >= N / 16 && a.W % ( N / 16 ) == 0 && a.W <= N && a.hop >= 2 && a.hop % 2 == 0 && a.hop <= a.W && a.hop <= N / 16;
	}
template<int N> static cudaError_t launch_synthesis_n( const SynthArgs & a, int64_t blocks, cudaStream_t st, int tps, int variant )
	{
	if constexpr( N == 8192 )
		{
		// 256 threads: two CTAs per SM need 128 registers and one exchange buffer (102 KB of shared memory each)
		if( variant == PV_PT_MIRROR && synthesis_mirror_applies( N, a ) )
			{
#ifdef FLAN_B200_DEBUG
			if( !a.one_buffer ) return launch_synthesis_mirror_nt<N, 256, false>( a, blocks, st );
#endif
			return launch_synthesis_mirror_nt<N, 512, true>( a, blocks, st );
			}
		}
	if constexpr( N == 1024 || N == 2048 || N == 4096 )
		{
		if( variant == PV_PT_MIRROR && synthesis_mirror_applies( N, a ) )
			{
#ifdef FLAN_B200_DEBUG
			if( tps >= 512 ) return a.one_buffer ? launch_synthesis_mirror_nt<N, 512, true>( a, blocks, st ) : launch_synthesis_mirror_nt<N, 512, false>( a, blocks, st );
			if( a.one_buffer ) return launch_synthesis_mirror_nt<N, 384, true>( a, blocks, st );
#endif
			return launch_synthesis_mirror_nt<N, 384, false>( a, blocks, st );
			}
		}
	// dft 8192: one exchange buffer and 64 registers per thread let two 512-thread CTAs share an SM
	if constexpr( N == 8192 ) { if( tps >= 1024 ) return launch_synthesis_nt<N, 1024, true>( a, blocks, st ); }
#ifdef FLAN_B200_DEBUG
	if( tps >= 1024 ) return launch_synthesis_nt<N, 1024, false>( a, blocks, st );
	if( tps < 768 ) return launch_synthesis_nt<N, 512, false>( a, blocks, st );
#endif
	return launch_synthesis_nt<N, 768, false>( a, blocks, st );
	}

bool dft_size_supported( int N )
	{
	return N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096 || N == 8192;
	}

cudaError_t launch_analysis( int N, const AnalysisArgs & a, int64_t blocks, cudaStream_t st, int tps, int pt )
	{
	switch( N )
		{
		case 256:  return launch_analysis_n<256>( a, blocks, st, tps, pt );
		case 512:  return launch_analysis_n<512>( a, blocks, st, tps, pt );
		case 1024: return launch_analysis_n<1024>( a, blocks, st, tps, pt );
		case 2048: return launch_analysis_n<2048>( a, blocks, st, tps, pt );
		case 4096: return launch_analysis_n<4096>( a, blocks, st, tps, pt );
		case 8192: return launch_analysis_n<8192>( a, blocks, st, tps, pt );
		default:   return cudaErrorInvalidValue;
		}
	}

cudaError_t launch_synthesis( int N, const SynthArgs & a, int64_t blocks, cudaStream_t st, int tps, int variant )
	{
	switch( N )
		{
		case 256:  return launch_synthesis_n<256>( a, blocks, st, tps, variant );
		case 512:  return launch_synthesis_n<512>( a, blocks, st, tps, variant );
		case 1024: return launch_synthesis_n<1024>( a, blocks, st, tps, variant );
		case 2048: return launch_synthesis_n<2048>( a, blocks, st, tps, variant );
		case 4096: return launch_synthesis_n<4096>( a, blocks, st, tps, variant );
		case 8192: return launch_synthesis_n<8192>( a, blocks, st, tps, variant );
		default:   return cudaErrorInvalidValue;
		}
	}

cudaError_t launch_phase_seg( const PhaseSegArgs & a, int C, cudaStream_t st )
	{
	const int64_t blocks = (int64_t)( ( a.B + 255 ) / 256 ) * a.segs_per_channel * C;
	pv_phase_seg_kernel<<<(unsigned) blocks, 256, 0, st>>>( a );
	return cudaGetLastError();
	}

cudaError_t launch_phase_scan( const PhaseScanArgs & a, int C, cudaStream_t st )
	{
	const dim3 wide( ( a.B + 127 ) / 128, a.groups, C ), narrow( ( a.B + 127 ) / 128, 1, C );
	if( a.segs_per_channel <= 256 && C <= 65535 && !a.expand_only )     // latency-bound sizes: one launch instead of three
		{
		pv_phase_scan_small_kernel<<<dim3( ( a.B + 31 ) / 32, C ), dim3( 32, 16 ), 0, st>>>( a );
		return cudaGetLastError();
		}
	if( !a.expand_only )
		{
		if( a.fix_pv && C <= 65535 && ( a.B + 63 ) / 64 <= 65535 )
			pv_phase_fix_kernel<<<dim3( a.segs_per_channel, C, ( a.B + 63 ) / 64 ), 64, 0, st>>>( a, a.B );
		pv_phase_scan_kernel<<<wide, 128, 0, st>>>( a, 0 );
		pv_phase_scan_kernel<<<narrow, 128, 0, st>>>( a, 1 );
		}
	if( a.acc_start ) pv_phase_scan_kernel<<<wide, 128, 0, st>>>( a, 2 );
	return cudaGetLastError();
	}

cudaError_t launch_phase_carry( const PhaseSeg * all, int rank, int64_t per_rank, PhaseSeg * carry, double P, double rcpP, cudaStream_t st )
	{
	pv_phase_carry_kernel<<<(unsigned)( ( per_rank + 127 ) / 128 ), 128, 0, st>>>( all, rank, per_rank, carry, P, rcpP );
	return cudaGetLastError();
	}

cudaError_t launch_mid_side( const float * in, float * out, int64_t n, int sms, cudaStream_t st )
	{
	int64_t blocks = ( n + 255 ) / 256;
	if( blocks > (int64_t) sms * 16 ) blocks = (int64_t) sms * 16;
	if( blocks < 1 ) blocks = 1;
	pv_mid_side_kernel<<<(unsigned) blocks, 256, 0, st>>>( in, out, n );
	return cudaGetLastError();
	}

cudaError_t launch_add( float * out, const float * add, int64_t n, int sms, cudaStream_t st )
	{
	int64_t blocks = ( n + 255 ) / 256;
	if( blocks > (int64_t) sms * 16 ) blocks = (int64_t) sms * 16;
	if( blocks < 1 ) blocks = 1;
	pv_add_kernel<<<(unsigned) blocks, 256, 0, st>>>( out, add, n );
	return cudaGetLastError();
	}


__global__ void __launch_bounds__( 1024 ) pv_state_push_kernel( const StatePush p )
	{
	uint4 * dst = p.dst[blockIdx.x];
	for( unsigned i = threadIdx.x; i < p.n16; i += blockDim.x ) dst[i] = p.src[i];
	__threadfence_system();
	__syncthreads();
	if( threadIdx.x == 0 && p.count[blockIdx.x] )
		{
		__threadfence_system();
		atomicAdd_system( p.count[blockIdx.x], 1u );      // the destination waits for ONE value: (its rank) x (steps of this parity)
		}
	}

cudaError_t launch_state_push( const StatePush & p, int destinations, cudaStream_t st )
	{
	if( destinations < 1 ) return cudaSuccess;
	if( destinations > 16 ) return cudaErrorInvalidValue;
	pv_state_push_kernel<<<destinations, 1024, 0, st>>>( p );
	return cudaGetLastError();
	}

} // namespace pvk
