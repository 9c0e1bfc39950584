// flan_b200/csrc/pv_capi.cu -- implementation of the C ABI declared in include/flan_b200.h.
//
// Host-side orchestration only: plan (constant table) cache, scratch workspace, launch geometry and
// error mapping. All arithmetic of the path runs in the kernels of pv_kernels.cu; there is no CPU
// fallback here -- every entry point fails with FLAN_B200_CUDA when no device is usable.
#include "../../include/flan_b200.h"

#include "pv_launch.h"
#include "pv_modify.h"
#include "pv_io.h"
#include "pv_tables.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

using namespace pvk;

static_assert( sizeof( flan_b200_phase_state ) == sizeof( PhaseSeg ), "phase state layout" );

namespace {

struct DevicePlan
	{
	HostTables host;
	float * win_analysis = nullptr;
	float * win_synthesis = nullptr;
	float * expected = nullptr;
	float2 * binc = nullptr;
	float2 * post_tw = nullptr;
	float2 * post_rot = nullptr;
	float4 * binc4 = nullptr;
	float2 * pass_tw = nullptr;
	float2 * pass_tw16 = nullptr;
	float2 * pass_tw_rev = nullptr;
	};

thread_local std::string g_create_error;

} // namespace

struct flan_b200_ctx
	{
	int device = 0;
	int sms = 0;
	cudaStream_t stream = nullptr;
	std::string error;
	std::mutex mutex;       // plan cache + workspace; launches on one ctx are serialised by its stream
	std::map<std::tuple<int, int, int, uint32_t, uint32_t>, std::unique_ptr<DevicePlan>> plans;
	void * workspace = nullptr;
	size_t workspace_bytes = 0;
	int * d_flag = nullptr;
	pvm::MapCheck * d_check = nullptr;     // time-map reduction of the PV-domain chain
	int64_t launches = 0;
	bool timing = false;
	// identity of the phase-segment summaries currently held in the workspace (flan_b200_phase_summary -> _range reuse)
	struct SegKey { const void * pv = nullptr; int64_t stride = 0, fb = 0, fe = 0; int C = 0, B = 0, W = 0; uint32_t sr = 0, ar = 0; bool valid = false; } seg_key;
	// Launch policy (overridable for experiments with FLAN_B200_TPS_* / FLAN_B200_PT_ANALYSIS): complex points per
	// thread of the analysis FFT (0 = by size: 16 from dft 4096 up, else 8) and the register-allocation variant
	// (resident threads per SM the kernel is compiled for; 0 = 512 with 16 points per thread, else 768).
	int tps_analysis = 0, tps_synthesis = 768;
	int pt_analysis = 0;
	int max_seg_len = 0;    // frames per CTA at most; 0 = by size: 128 from dft 2048 up (cfg2 3.96 -> 3.91 ms, cfg4 chain 15.1 -> 14.8 ms: mostly the shorter scan), else 64 (FLAN_B200_SEG_LEN)
	int one_buffer = -1;    // analysis exchange buffers alias: -1 = by size (FLAN_B200_ONEBUF)
	int synth_variant = PV_PT_MIRROR;  // PV_PT_MIRROR = mirrored first pass where it applies; 8 = always the 8-point kernel (FLAN_B200_SYNTH_VARIANT)
	int tps_synthesis_mirror = 384;
	int synth_one_buffer = -1;          // mirrored resynthesis with one exchange buffer: -1 = by size (FLAN_B200_SYNTH_ONEBUF)
	bool tps_synthesis_set = false;     // dft 8192 defaults to the 1024-thread (two CTAs per SM, one exchange buffer) variant
	struct Timed { int kind; cudaEvent_t start, stop; };
	std::vector<Timed> timed;
	};

namespace {

int fail( flan_b200_ctx * ctx, int code, const std::string & msg )
	{
	if( ctx ) ctx->error = msg;
	return code;
	}

int cuda_fail( flan_b200_ctx * ctx, cudaError_t e, const char * what )
	{
	return fail( ctx, e == cudaErrorMemoryAllocation ? FLAN_B200_NOMEM : FLAN_B200_CUDA,
	             std::string( what ) + ": " + cudaGetErrorString( e ) );
	}

#define CK( call, what ) do { cudaError_t e_ = ( call ); if( e_ != cudaSuccess ) return cuda_fail( ctx, e_, what ); } while( 0 )

uint32_t fbits( float f ) { uint32_t u; std::memcpy( &u, &f, 4 ); return u; }

template<class T> cudaError_t upload_vec( const std::vector<T> & v, T ** d, cudaStream_t st )
	{
	cudaError_t e = cudaMalloc( (void **) d, sizeof( T ) * ( v.empty() ? 1 : v.size() ) );
	if( e != cudaSuccess ) return e;
	if( v.empty() ) return cudaSuccess;
	// pageable source: the copy is staged before the call returns, so `v` may die afterwards
	return cudaMemcpyAsync( *d, v.data(), sizeof( T ) * v.size(), cudaMemcpyHostToDevice, st );
	}

int get_plan( flan_b200_ctx * ctx, int N, int W, int hop, float sr, float ar, DevicePlan ** out )
	{
	if( !dft_size_supported( N ) )
		return fail( ctx, FLAN_B200_UNSUPPORTED, "dft_size must be a power of two in [256, 8192], got " + std::to_string( N ) );
	if( W < 2 || W > N || hop < 1 || (int64_t) N * W / hop < 1 )
		return fail( ctx, FLAN_B200_INVALID, "need 2 <= window_size <= dft_size and 1 <= hop <= dft_size*window_size" );
	if( !( sr > 0.0f ) || !( ar > 0.0f ) )
		return fail( ctx, FLAN_B200_INVALID, "sample_rate and analysis_rate must be positive" );
	std::lock_guard<std::mutex> lock( ctx->mutex );
	const auto key = std::make_tuple( N, W, hop, fbits( sr ), fbits( ar ) );
	auto it = ctx->plans.find( key );
	if( it != ctx->plans.end() ) { *out = it->second.get(); return FLAN_B200_OK; }
	auto plan = std::make_unique<DevicePlan>();
	if( !build_tables( N, W, hop, sr, ar, plan->host ) )
		return fail( ctx, FLAN_B200_INVALID, "could not build plan tables" );
	CK( upload_vec( plan->host.win_analysis, &plan->win_analysis, ctx->stream ), "plan upload" );
	CK( upload_vec( plan->host.win_synthesis, &plan->win_synthesis, ctx->stream ), "plan upload" );
	CK( upload_vec( plan->host.expected, &plan->expected, ctx->stream ), "plan upload" );
	CK( upload_vec( plan->host.binc, &plan->binc, ctx->stream ), "plan upload" );
	CK( upload_vec( plan->host.post_tw, &plan->post_tw, ctx->stream ), "plan upload" );
	CK( upload_vec( plan->host.post_rot, &plan->post_rot, ctx->stream ), "plan upload" );
	CK( upload_vec( plan->host.binc4, &plan->binc4, ctx->stream ), "plan upload" );
	CK( upload_vec( plan->host.pass_tw, &plan->pass_tw, ctx->stream ), "plan upload" );
	CK( upload_vec( plan->host.pass_tw16, &plan->pass_tw16, ctx->stream ), "plan upload" );
	CK( upload_vec( plan->host.pass_tw_rev, &plan->pass_tw_rev, ctx->stream ), "plan upload" );
	*out = plan.get();
	ctx->plans[key] = std::move( plan );
	return FLAN_B200_OK;
	}

int get_workspace( flan_b200_ctx * ctx, size_t bytes, void ** out )
	{
	std::lock_guard<std::mutex> lock( ctx->mutex );
	if( bytes > ctx->workspace_bytes )
		{
		if( ctx->workspace )
			{
			CK( cudaStreamSynchronize( ctx->stream ), "workspace sync" );
			cudaFree( ctx->workspace );
			ctx->workspace = nullptr; ctx->workspace_bytes = 0;
			}
		CK( cudaMalloc( &ctx->workspace, bytes ), "workspace alloc" );
		ctx->workspace_bytes = bytes;
		}
	*out = ctx->workspace;
	return FLAN_B200_OK;
	}

// Brackets one kernel launch with events when timing is on (bench.py's per-kernel roofline).
struct LaunchTimer
	{
	flan_b200_ctx * ctx; int kind; cudaEvent_t start = nullptr, stop = nullptr;
	LaunchTimer( flan_b200_ctx * c, int k ) : ctx( c ), kind( k )
		{
		if( !ctx->timing ) return;
		if( cudaEventCreate( &start ) != cudaSuccess || cudaEventCreate( &stop ) != cudaSuccess ) { start = stop = nullptr; return; }
		cudaEventRecord( start, ctx->stream );
		}
	~LaunchTimer()
		{
		ctx->launches++;
		if( !start ) return;
		cudaEventRecord( stop, ctx->stream );
		ctx->timed.push_back( { kind, start, stop } );
		}
	};

bool cancelled( const volatile int * cancel ) { return cancel && *cancel; }

size_t align_up( size_t v, size_t a ) { return ( v + a - 1 ) / a * a; }

// Shared by the whole-signal and the frame-range forms of resynthesis.
int synth_range( flan_b200_ctx * ctx, const float * d_pv_rows, int64_t pv_channel_stride, int C,
                 int64_t frame_begin, int64_t frame_end, int64_t frames_total, int B,
                 float sr, float ar, int W, const PhaseSeg * d_carry_in, PhaseSeg * d_carry_out,
                 float * d_out, int64_t out_stride, int64_t out_offset, int64_t out_len,
                 bool summary_only, const volatile int * cancel, bool reuse_summary = false )
	{
	if( C < 1 || B < 2 || frame_begin < 0 || frame_end < frame_begin || frames_total < frame_end )
		return fail( ctx, FLAN_B200_INVALID, "bad channel / bin / frame-range arguments" );
	const int N = ( B - 1 ) * 2;                                        // PVBuffer.cpp:356-359
	const int hop = flan_b200_hop_from_rates( sr, ar );
	DevicePlan * plan = nullptr;
	int rc = get_plan( ctx, N, W, hop, sr, ar, &plan );
	if( rc ) return rc;
	const int64_t frames = frame_end - frame_begin;
	if( frames == 0 ) return FLAN_B200_OK;
	if( cancelled( cancel ) ) return fail( ctx, FLAN_B200_CANCELLED, "cancelled" );

	const int seg_len = choose_seg_len( frames, C, ctx->sms, W, hop, ctx->max_seg_len ? ctx->max_seg_len : ( N >= 2048 ? 128 : 64 ) );
	const int segs = (int)( ( frames + seg_len - 1 ) / seg_len );
	const size_t seg_bytes = align_up( sizeof( PhaseSeg ) * (size_t) C * segs * B, 256 );
	const size_t acc_bytes = align_up( sizeof( double ) * (size_t) C * segs * B, 256 );
	int group_len = 32;
	while( ( segs + group_len - 1 ) / group_len > 65535 ) group_len *= 2;
	const int groups = ( segs + group_len - 1 ) / group_len;
	const size_t grp_bytes = align_up( sizeof( PhaseSeg ) * (size_t) C * groups * B, 256 );
	void * ws = nullptr;
	rc = get_workspace( ctx, seg_bytes + acc_bytes + grp_bytes, &ws );
	if( rc ) return rc;
	PhaseSeg * d_seg = (PhaseSeg *) ws;
	double * d_acc = (double *)( (char *) ws + seg_bytes );
	PhaseSeg * d_grp = (PhaseSeg *)( (char *) ws + seg_bytes + acc_bytes );

	flan_b200_ctx::SegKey key;
	key.pv = d_pv_rows; key.stride = pv_channel_stride; key.fb = frame_begin; key.fe = frame_end;
	key.C = C; key.B = B; key.W = W; key.sr = fbits( sr ); key.ar = fbits( ar ); key.valid = true;
	const flan_b200_ctx::SegKey & old = ctx->seg_key;
	const bool have_summaries = reuse_summary && old.valid && old.pv == key.pv && old.stride == key.stride && old.fb == key.fb
	                         && old.fe == key.fe && old.C == key.C && old.B == key.B && old.W == key.W && old.sr == key.sr && old.ar == key.ar;
	ctx->seg_key = key;

	PhaseSegArgs sa{};
	sa.pv = (const float2 *) d_pv_rows; sa.pv_channel_stride = pv_channel_stride;
	sa.frame_begin = frame_begin; sa.frame_end = frame_end;
	sa.seg_len = seg_len; sa.segs_per_channel = segs; sa.B = B;
	sa.seg_out = d_seg; sa.nan_flag = ctx->d_flag;
	sa.k = plan->host.k; sa.P = plan->host.P; sa.rcpP = plan->host.rcpP;
	if( !have_summaries ) { LaunchTimer lt( ctx, 1 ); CK( launch_phase_seg( sa, C, ctx->stream ), "phase summary launch" ); }

	PhaseScanArgs sc{};
	sc.seg = d_seg; sc.segs_per_channel = segs; sc.B = B;
	sc.group_len = group_len; sc.groups = groups; sc.group = d_grp;
	sc.carry_in = d_carry_in; sc.carry_out = d_carry_out;
	sc.acc_start = summary_only ? nullptr : d_acc;
	sc.P = plan->host.P; sc.rcpP = plan->host.rcpP;
	{ LaunchTimer lt( ctx, 2 ); CK( launch_phase_scan( sc, C, ctx->stream ), "phase scan launch" ); ctx->launches += ( segs <= 256 ) ? 0 : ( summary_only ? 1 : 2 ); }     // launch_phase_scan: one launch for short signals, else 2 or 3
	if( summary_only ) return FLAN_B200_OK;
	if( cancelled( cancel ) ) return fail( ctx, FLAN_B200_CANCELLED, "cancelled" );

	// The kernels store every sample that only one segment reaches and red.add the rest onto zeros: clear just those
	// (a few per cent of the output) when the frames' windows leave no gaps, everything otherwise.
	if( hop <= W && C < 65535 )
		{ CK( launch_zero_shared( d_out, out_stride, out_offset, out_len, C, frame_begin, frame_end, seg_len, segs, W, hop, ctx->stream ), "output clear" ); ctx->launches++; }
	else
		for( int c = 0; c < C; ++c )
			CK( cudaMemsetAsync( d_out + (int64_t) c * out_stride, 0, sizeof( float ) * (size_t) out_len, ctx->stream ), "output clear" );

	SynthArgs a{};
	a.pv = (const float2 *) d_pv_rows; a.pv_channel_stride = pv_channel_stride;
	a.frame_begin = frame_begin; a.frame_end = frame_end;
	a.out = d_out; a.out_stride = out_stride; a.out_offset = out_offset;
	const int64_t total = frames_total * hop;                           // AudioPV.cpp:93
	a.out_lo = out_offset > 0 ? out_offset : 0;
	a.out_hi = ( out_offset + out_len < total ) ? out_offset + out_len : total;
	a.acc_start = d_acc;
	a.seg_len = seg_len; a.segs_per_channel = segs;
	a.W = W; a.hop = hop;
	a.aligned2 = ( hop % 2 == 0 ) && ( ( W / 2 ) % 2 == 0 );
	a.win = plan->win_synthesis; a.post_tw = plan->post_tw; a.pass_tw = plan->pass_tw; a.pass_tw_rev = plan->pass_tw_rev;
	a.out_aligned2 = ( out_stride % 2 == 0 ) && ( out_offset % 2 == 0 ) && ( (uintptr_t) d_out % 8 == 0 );
	a.pv_aligned16 = ( (uintptr_t) d_pv_rows % 16 == 0 ); a.channels = C;
	a.one_buffer = ctx->synth_one_buffer >= 0 ? ctx->synth_one_buffer : ( N == 8192 ? 1 : 0 );    // dft 8192: two 256-thread CTAs per SM
	a.k = plan->host.k; a.P = plan->host.P; a.rcpP = plan->host.rcpP;
	{ LaunchTimer lt( ctx, 3 ); const bool mirror = ctx->synth_variant == PV_PT_MIRROR && synthesis_mirror_applies( N, a );
	  CK( launch_synthesis( N, a, (int64_t) C * segs, ctx->stream, mirror ? ctx->tps_synthesis_mirror : ( ( N == 8192 && !ctx->tps_synthesis_set ) ? 1024 : ctx->tps_synthesis ), ctx->synth_variant ), "synthesis launch" ); }
	return FLAN_B200_OK;
	}

} // namespace

extern "C" {

int flan_b200_device_count( void )
	{
	int n = 0;
	if( cudaGetDeviceCount( &n ) != cudaSuccess ) return 0;
	return n;
	}

int flan_b200_create( int device, flan_b200_ctx ** out )
	{
	if( !out ) return FLAN_B200_INVALID;
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount( &n );
	if( e != cudaSuccess || n < 1 )
		{
		g_create_error = std::string( "no CUDA device: " ) + ( e != cudaSuccess ? cudaGetErrorString( e ) : "device count is 0" )
		               + " (flan_b200 has no CPU fallback)";
		return FLAN_B200_CUDA;
		}
	if( device < 0 || device >= n ) { g_create_error = "device index out of range"; return FLAN_B200_INVALID; }
	e = cudaSetDevice( device );
	if( e != cudaSuccess ) { g_create_error = cudaGetErrorString( e ); return FLAN_B200_CUDA; }
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties( &prop, device );
	if( e != cudaSuccess ) { g_create_error = cudaGetErrorString( e ); return FLAN_B200_CUDA; }
	if( prop.major < 10 )
		{
		g_create_error = "flan_b200 kernels are built for sm_100a only; device is sm_" + std::to_string( prop.major * 10 + prop.minor );
		return FLAN_B200_CUDA;
		}
	auto * ctx = new flan_b200_ctx;
	ctx->device = device;
	if( const char * e = std::getenv( "FLAN_B200_TPS_ANALYSIS" ) ) ctx->tps_analysis = std::atoi( e );
	if( const char * e = std::getenv( "FLAN_B200_TPS_SYNTHESIS" ) ) { ctx->tps_synthesis = std::atoi( e ); ctx->tps_synthesis_mirror = ctx->tps_synthesis; ctx->tps_synthesis_set = true; }
	if( const char * e = std::getenv( "FLAN_B200_PT_ANALYSIS" ) ) ctx->pt_analysis = std::atoi( e );
	if( const char * e = std::getenv( "FLAN_B200_ONEBUF" ) ) ctx->one_buffer = std::atoi( e );
	if( const char * e = std::getenv( "FLAN_B200_SYNTH_ONEBUF" ) ) ctx->synth_one_buffer = std::atoi( e );
	if( const char * e = std::getenv( "FLAN_B200_SEG_LEN" ) ) { const int v = std::atoi( e ); if( v >= 4 ) ctx->max_seg_len = v; }
	if( const char * e = std::getenv( "FLAN_B200_SYNTH_VARIANT" ) ) ctx->synth_variant = std::atoi( e );
	ctx->sms = prop.multiProcessorCount;
	e = cudaMalloc( (void **) &ctx->d_flag, sizeof( int ) );
	if( e == cudaSuccess ) e = cudaMemset( ctx->d_flag, 0, sizeof( int ) );
	if( e == cudaSuccess ) e = cudaMalloc( (void **) &ctx->d_check, sizeof( pvm::MapCheck ) );
	if( e != cudaSuccess ) { g_create_error = cudaGetErrorString( e ); delete ctx; return FLAN_B200_CUDA; }
	*out = ctx;
	return FLAN_B200_OK;
	}

void flan_b200_destroy( flan_b200_ctx * ctx )
	{
	if( !ctx ) return;
	cudaSetDevice( ctx->device );
	cudaStreamSynchronize( ctx->stream );
	for( auto & kv : ctx->plans )
		{
		DevicePlan * p = kv.second.get();
		cudaFree( p->win_analysis ); cudaFree( p->win_synthesis ); cudaFree( p->expected ); cudaFree( p->binc );
		cudaFree( p->post_tw ); cudaFree( p->post_rot ); cudaFree( p->binc4 ); cudaFree( p->pass_tw ); cudaFree( p->pass_tw16 );
		}
	if( ctx->workspace ) cudaFree( ctx->workspace );
	cudaFree( ctx->d_flag );
	cudaFree( ctx->d_check );
	delete ctx;
	}

const char * flan_b200_last_error( const flan_b200_ctx * ctx )
	{
	return ctx ? ctx->error.c_str() : g_create_error.c_str();
	}

int flan_b200_set_stream( flan_b200_ctx * ctx, void * cuda_stream )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	ctx->stream = (cudaStream_t) cuda_stream;
	return FLAN_B200_OK;
	}

int flan_b200_synchronize( flan_b200_ctx * ctx )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	CK( cudaStreamSynchronize( ctx->stream ), "synchronize" );
	return FLAN_B200_OK;
	}

int flan_b200_sm_count( const flan_b200_ctx * ctx ) { return ctx ? ctx->sms : 0; }
int64_t flan_b200_launch_count( const flan_b200_ctx * ctx ) { return ctx ? ctx->launches : 0; }

int flan_b200_set_timing( flan_b200_ctx * ctx, int enabled )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	ctx->timing = enabled != 0;
	return FLAN_B200_OK;
	}

int flan_b200_kernel_time( flan_b200_ctx * ctx, int kind, double * total_ms, int64_t * launches )
	{
	if( !ctx || !total_ms || !launches ) return FLAN_B200_INVALID;
	CK( cudaStreamSynchronize( ctx->stream ), "timing sync" );
	double ms = 0.0; int64_t n = 0;
	std::vector<flan_b200_ctx::Timed> keep;
	for( auto & t : ctx->timed )
		{
		if( t.kind != kind ) { keep.push_back( t ); continue; }
		float f = 0.0f;
		if( cudaEventElapsedTime( &f, t.start, t.stop ) == cudaSuccess ) { ms += f; ++n; }
		cudaEventDestroy( t.start ); cudaEventDestroy( t.stop );
		}
	ctx->timed.swap( keep );
	*total_ms = ms; *launches = n;
	return FLAN_B200_OK;
	}

int flan_b200_malloc( flan_b200_ctx * ctx, size_t bytes, void ** d_out )
	{
	if( !ctx || !d_out ) return FLAN_B200_INVALID;
	CK( cudaSetDevice( ctx->device ), "set device" );
	CK( cudaMalloc( d_out, bytes ? bytes : 1 ), "device alloc" );
	return FLAN_B200_OK;
	}

int flan_b200_free( flan_b200_ctx * ctx, void * d_ptr )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	CK( cudaFree( d_ptr ), "device free" );
	return FLAN_B200_OK;
	}

int flan_b200_upload( flan_b200_ctx * ctx, void * d_dst, const void * h_src, size_t bytes )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	CK( cudaMemcpyAsync( d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream ), "upload" );
	return FLAN_B200_OK;
	}

int flan_b200_download( flan_b200_ctx * ctx, void * h_dst, const void * d_src, size_t bytes )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	CK( cudaMemcpyAsync( h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream ), "download" );
	return FLAN_B200_OK;
	}

int64_t flan_b200_num_frames( int64_t n, int hop )
	{
	if( hop < 1 ) return 0;
	return n / hop + 1;                                 // AudioPV.cpp:17
	}

int flan_b200_hop_from_rates( float sample_rate, float analysis_rate )
	{
	return (int)( sample_rate / analysis_rate );        // PVBuffer.cpp:381-384
	}

float flan_b200_analysis_rate( float sample_rate, int hop )
	{
	return sample_rate / hop;                           // AudioPV.cpp:25
	}

int flan_b200_convert_to_pv_range( flan_b200_ctx * ctx, const float * d_audio_local, int64_t audio_stride,
                                   int64_t audio_offset, int64_t audio_len, int C, int64_t n_total,
                                   float sr, int W, int hop, int N,
                                   int64_t frame_begin, int64_t frame_end,
                                   float * d_pv_rows, int64_t pv_channel_stride )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	if( C < 1 || n_total < 0 || hop < 1 )
		return fail( ctx, FLAN_B200_INVALID, "bad channel count, length or hop" );
	const int64_t F = flan_b200_num_frames( n_total, hop );
	if( frame_begin < 0 || frame_end < frame_begin || frame_end > F )
		return fail( ctx, FLAN_B200_INVALID, "frame range outside [0, n/hop + 1]" );
	DevicePlan * plan = nullptr;
	int rc = get_plan( ctx, N, W, hop, sr, flan_b200_analysis_rate( sr, hop ), &plan );
	if( rc ) return rc;
	const int64_t frames = frame_end - frame_begin;
	if( frames == 0 ) return FLAN_B200_OK;
	// the shard must hold every in-signal sample its frames (and the warm-up frame) read
	int64_t need_lo = (int64_t) hop * ( frame_begin > 0 ? frame_begin - 1 : 0 ) - W / 2;
	int64_t need_hi = (int64_t) hop * ( frame_end - 1 ) - W / 2 + W;
	if( need_lo < 0 ) need_lo = 0;
	if( need_hi > n_total ) need_hi = n_total;
	if( need_hi > need_lo && ( audio_offset > need_lo || audio_offset + audio_len < need_hi ) )
		return fail( ctx, FLAN_B200_INVALID, "local audio does not cover the halo of the requested frame range" );

	const int seg_len = choose_seg_len( frames, C, ctx->sms, W, hop, ctx->max_seg_len ? ctx->max_seg_len : ( N >= 2048 ? 128 : 64 ), true );
	const int segs = (int)( ( frames + seg_len - 1 ) / seg_len );
	AnalysisArgs a{};
	a.audio = d_audio_local; a.audio_stride = audio_stride; a.audio_offset = audio_offset; a.n_total = n_total;
	a.pv = (float2 *) d_pv_rows; a.pv_channel_stride = pv_channel_stride;
	a.frame_begin = frame_begin; a.frame_end = frame_end;
	a.seg_len = seg_len; a.segs_per_channel = segs;
	a.W = W; a.hop = hop;
	a.aligned2 = ( hop % 2 == 0 ) && ( ( W / 2 ) % 2 == 0 ) && ( audio_stride % 2 == 0 ) && ( audio_offset % 2 == 0 )
	          && ( (uintptr_t) d_audio_local % 8 == 0 );
	// measured on B200 (tools/experiments/exp_r1*.sh): 16 points per thread with one exchange buffer from dft 4096 up; the mirrored
	// last pass for dft 1024 with the standard window / hop; 8 points per thread otherwise
	// (dft 2048: 16 points per thread once the grid covers the SMs a few times over -- 2.52 -> 2.31 ms on a cfg4 channel --
	// 8 for short signals, where twice the threads per frame matter more)
	// (decided on the WHOLE signal's frame count, so that every frame-range shard of a signal runs the same arithmetic)
	const bool large = (int64_t) C * ( n_total / hop + 1 ) >= (int64_t) ctx->sms * 128;
	int pt = ctx->pt_analysis ? ctx->pt_analysis : ( N >= 4096 ? 16 : ( N == 2048 ? ( large ? 16 : 8 ) : ( N == 1024 ? PV_PT_MIRROR : 8 ) ) );
	if( pt == PV_PT_MIRROR && !( mirror_supported( N ) && W == N && hop == N / 16 ) ) pt = ( N >= 4096 ) ? 16 : 8;
	if( pt != PV_PT_MIRROR && ( pt != 16 || N < 512 ) ) pt = 8;
	const int tps_a = ctx->tps_analysis ? ctx->tps_analysis : ( pt >= 16 ? 512 : 768 );
	a.win = plan->win_analysis; a.binc = plan->binc; a.binc4 = plan->binc4; a.post_rot = plan->post_rot;
	a.pass_tw = ( pt >= 16 ) ? plan->pass_tw16 : plan->pass_tw;
	a.one_buffer = ctx->one_buffer >= 0 ? ctx->one_buffer : ( ( pt == 16 || pt == PV_PT_MIRROR ) ? 1 : 0 );
	a.k = plan->host.k;
	ctx->seg_key.valid = false;
	{ LaunchTimer lt( ctx, 0 ); CK( launch_analysis( N, a, (int64_t) C * segs, ctx->stream, tps_a, pt ), "analysis launch" ); }
	return FLAN_B200_OK;
	}

int flan_b200_convert_to_pv( flan_b200_ctx * ctx, const float * d_audio, int C, int64_t n,
                             float sr, int W, int hop, int N, float * d_pv, const volatile int * cancel )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	if( cancelled( cancel ) ) return fail( ctx, FLAN_B200_CANCELLED, "cancelled" );
	if( hop < 1 ) return fail( ctx, FLAN_B200_INVALID, "hop must be >= 1" );
	const int64_t F = flan_b200_num_frames( n, hop );
	int rc = flan_b200_convert_to_pv_range( ctx, d_audio, n, 0, n, C, n, sr, W, hop, N, 0, F, d_pv, F * ( N / 2 + 1 ) );
	if( rc ) return rc;
	if( cancelled( cancel ) ) return fail( ctx, FLAN_B200_CANCELLED, "cancelled" );
	return FLAN_B200_OK;
	}

int flan_b200_convert_to_audio( flan_b200_ctx * ctx, const float * d_pv, int C, int64_t F, int B,
                                float sr, float ar, int W, float * d_audio_out,
                                const volatile int * cancel, int * nan_or_inf )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	if( !( ar > 0.0f ) || !( sr > 0.0f ) ) return fail( ctx, FLAN_B200_INVALID, "rates must be positive" );
	const int hop = flan_b200_hop_from_rates( sr, ar );
	if( hop < 1 ) return fail( ctx, FLAN_B200_INVALID, "analysis_rate above sample_rate gives hop 0" );
	const int64_t out_n = F * hop;
	if( nan_or_inf ) CK( cudaMemsetAsync( ctx->d_flag, 0, sizeof( int ), ctx->stream ), "flag clear" );
	int rc = synth_range( ctx, d_pv, F * B, C, 0, F, F, B, sr, ar, W, nullptr, nullptr,
	                      d_audio_out, out_n, 0, out_n, false, cancel );
	if( rc ) return rc;
	if( nan_or_inf )
		{
		CK( cudaMemcpyAsync( nan_or_inf, ctx->d_flag, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream ), "flag read" );
		CK( cudaStreamSynchronize( ctx->stream ), "flag sync" );
		}
	if( cancelled( cancel ) ) return fail( ctx, FLAN_B200_CANCELLED, "cancelled" );
	return FLAN_B200_OK;
	}

int flan_b200_phase_summary( flan_b200_ctx * ctx, const float * d_pv_rows, int64_t pv_channel_stride,
                             int C, int64_t frame_begin, int64_t frame_end, int B,
                             float sr, float ar, int W, flan_b200_phase_state * d_state_out )
	{
	if( !ctx || !d_state_out ) return FLAN_B200_INVALID;
	if( frame_end == frame_begin )
		{
		CK( cudaMemsetAsync( d_state_out, 0, sizeof( PhaseSeg ) * (size_t) C * B, ctx->stream ), "state clear" );
		return FLAN_B200_OK;
		}
	return synth_range( ctx, d_pv_rows, pv_channel_stride, C, frame_begin, frame_end, frame_end, B, sr, ar, W,
	                    nullptr, (PhaseSeg *) d_state_out, nullptr, 0, 0, 0, true, nullptr );
	}

int flan_b200_phase_carry( flan_b200_ctx * ctx, const flan_b200_phase_state * d_all, int rank,
                           int C, int B, flan_b200_phase_state * d_carry_out )
	{
	if( !ctx || rank < 0 || C < 1 || B < 1 ) return FLAN_B200_INVALID;
	const double P = (double)( std::acos( -1.0f ) * 2.0f );
	{ LaunchTimer lt( ctx, 4 ); CK( launch_phase_carry( (const PhaseSeg *) d_all, rank, (int64_t) C * B, (PhaseSeg *) d_carry_out, P, 1.0 / P, ctx->stream ), "phase carry launch" ); }
	return FLAN_B200_OK;
	}

int flan_b200_convert_to_audio_range( flan_b200_ctx * ctx, const float * d_pv_rows, int64_t pv_channel_stride,
                                      int C, int64_t frame_begin, int64_t frame_end, int64_t frames_total,
                                      int B, float sr, float ar, int W,
                                      const flan_b200_phase_state * d_carry_in, int reuse_summary,
                                      float * d_out_local, int64_t out_stride, int64_t out_offset, int64_t out_len )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	if( !( ar > 0.0f ) || !( sr > 0.0f ) || flan_b200_hop_from_rates( sr, ar ) < 1 )
		return fail( ctx, FLAN_B200_INVALID, "bad rates" );
	return synth_range( ctx, d_pv_rows, pv_channel_stride, C, frame_begin, frame_end, frames_total, B, sr, ar, W,
	                    (const PhaseSeg *) d_carry_in, nullptr, d_out_local, out_stride, out_offset, out_len, false, nullptr,
	                    reuse_summary != 0 );
	}

int flan_b200_add( flan_b200_ctx * ctx, float * d_out, const float * d_add, int64_t n )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	if( n <= 0 ) return FLAN_B200_OK;
	{ LaunchTimer lt( ctx, 4 ); CK( launch_add( d_out, d_add, n, ctx->sms, ctx->stream ), "add launch" ); }
	return FLAN_B200_OK;
	}

int flan_b200_mid_side( flan_b200_ctx * ctx, const float * d_in, float * d_out, int64_t n )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	if( n <= 0 ) return FLAN_B200_OK;
	{ LaunchTimer lt( ctx, 4 ); CK( launch_mid_side( d_in, d_out, n, ctx->sms, ctx->stream ), "mid/side launch" ); }
	return FLAN_B200_OK;
	}

int flan_b200_convert_to_pv_host( flan_b200_ctx * ctx, const float * h_audio, int C, int64_t n,
                                  float sr, int W, int hop, int N, int mid_side,
                                  float * h_pv, const volatile int * cancel )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	if( hop < 1 || C < 1 || n < 0 ) return fail( ctx, FLAN_B200_INVALID, "bad shape" );
	if( mid_side && C != 2 ) return fail( ctx, FLAN_B200_INVALID, "mid/side needs exactly two channels (AudioPV.cpp:82)" );
	const int64_t F = flan_b200_num_frames( n, hop );
	const size_t audio_bytes = sizeof( float ) * (size_t) C * n;
	const size_t pv_bytes = sizeof( float ) * 2 * (size_t) C * F * ( N / 2 + 1 );
	float * d_audio = nullptr, * d_ms = nullptr, * d_pv = nullptr;
	int rc = flan_b200_malloc( ctx, audio_bytes, (void **) &d_audio );
	if( !rc && mid_side ) rc = flan_b200_malloc( ctx, audio_bytes, (void **) &d_ms );
	if( !rc ) rc = flan_b200_malloc( ctx, pv_bytes, (void **) &d_pv );
	if( !rc ) rc = flan_b200_upload( ctx, d_audio, h_audio, audio_bytes );
	if( !rc && mid_side ) rc = flan_b200_mid_side( ctx, d_audio, d_ms, n );
	if( !rc ) rc = flan_b200_convert_to_pv( ctx, mid_side ? d_ms : d_audio, C, n, sr, W, hop, N, d_pv, cancel );
	if( !rc ) rc = flan_b200_download( ctx, h_pv, d_pv, pv_bytes );
	cudaError_t e = cudaStreamSynchronize( ctx->stream );
	cudaFree( d_audio ); cudaFree( d_ms ); cudaFree( d_pv );
	if( !rc && e != cudaSuccess ) return cuda_fail( ctx, e, "convert_to_pv_host" );
	return rc;
	}

int flan_b200_convert_to_audio_host( flan_b200_ctx * ctx, const float * h_pv, int C, int64_t F, int B,
                                     float sr, float ar, int W, int left_right,
                                     float * h_audio_out, const volatile int * cancel, int * nan_or_inf )
	{
	if( !ctx ) return FLAN_B200_INVALID;
	if( C < 1 || F < 0 || B < 2 || !( ar > 0.0f ) ) return fail( ctx, FLAN_B200_INVALID, "bad shape" );
	if( left_right && C != 2 ) return fail( ctx, FLAN_B200_INVALID, "left/right needs exactly two channels (AudioPV.cpp:143)" );
	const int hop = flan_b200_hop_from_rates( sr, ar );
	if( hop < 1 ) return fail( ctx, FLAN_B200_INVALID, "hop < 1" );
	const int64_t out_n = F * hop;
	const size_t pv_bytes = sizeof( float ) * 2 * (size_t) C * F * B;
	const size_t audio_bytes = sizeof( float ) * (size_t) C * out_n;
	float * d_pv = nullptr, * d_audio = nullptr, * d_lr = nullptr;
	int rc = flan_b200_malloc( ctx, pv_bytes, (void **) &d_pv );
	if( !rc ) rc = flan_b200_malloc( ctx, audio_bytes, (void **) &d_audio );
	if( !rc && left_right ) rc = flan_b200_malloc( ctx, audio_bytes, (void **) &d_lr );
	if( !rc ) rc = flan_b200_upload( ctx, d_pv, h_pv, pv_bytes );
	if( !rc ) rc = flan_b200_convert_to_audio( ctx, d_pv, C, F, B, sr, ar, W, d_audio, cancel, nan_or_inf );
	if( !rc && left_right ) rc = flan_b200_mid_side( ctx, d_audio, d_lr, out_n );
	if( !rc ) rc = flan_b200_download( ctx, h_audio_out, left_right ? d_lr : d_audio, audio_bytes );
	cudaError_t e = cudaStreamSynchronize( ctx->stream );
	cudaFree( d_pv ); cudaFree( d_audio ); cudaFree( d_lr );
	if( !rc && e != cudaSuccess ) return cuda_fail( ctx, e, "convert_to_audio_host" );
	return rc;
	}

} // extern "C"

// ---- PV-domain chain (PV/PVModify.cpp:196-385) --------------------------------------------------

namespace {

int check_pv_shape( flan_b200_ctx * ctx, int C, int64_t F, int B, float sr, int interp )
	{
	if( C < 1 || F < 1 || B < 2 ) return fail( ctx, FLAN_B200_INVALID, "need channels >= 1, frames >= 1, bins >= 2" );
	if( !( sr > 0.0f ) ) return fail( ctx, FLAN_B200_INVALID, "sample_rate must be positive" );
	if( interp < 0 || interp > 9 ) return fail( ctx, FLAN_B200_INVALID, "interpolator id outside 0..9 (Utility/Interpolator.cpp)" );
	return FLAN_B200_OK;
	}

bool strides_ok( int64_t fs, int bs, int B ) { return ( bs == 0 || bs == 1 ) && ( fs == 0 || fs == ( bs ? B : 1 ) ); }

// mod_hz.frame_stride == 0 && bin_stride == 1 (one row of positions shared by all frames): plan + gather kernel, with
// the general row kernel as the device-side alternative when the plan kernel finds the positions non-monotone.
// plan_ws: 2 * B * 4 + 256 bytes of scratch for the plan, or null.
int repitch_common( flan_b200_ctx * ctx, const float * d_pv, int C, int64_t F, int B, float sr,
                    const pvm::Table & mod_hz, const float * d_in_mod, int interp, float * d_out, void * plan_ws )
	{
	pvm::RepitchArgs a{};
	a.pv = (const float2 *) d_pv; a.out = (float2 *) d_out;
	a.mod = mod_hz; a.in_mod = d_in_mod;
	a.F = F; a.B = B;
	a.bin_width = sr / float( ( B - 1 ) * 2 );                          // PVBuffer.cpp:438-441
	a.interp = interp;
	if( pvm::RepitchRow::bytes( B ) > 200 * 1024 ) return fail( ctx, FLAN_B200_UNSUPPORTED, "too many bins for one shared-memory row" );
	const int64_t rows = (int64_t) C * F;
	const int * skip_if = nullptr;
	if( plan_ws && mod_hz.frame_stride == 0 && mod_hz.bin_stride == 1 && pvm::repitch_shared_supported( B ) )
		{
		pvm::RepitchPlan plan{};
		plan.src = (int *) plan_ws; plan.mix = (float *)( plan.src + B ); plan.ok = (int *)( plan.mix + B );
		{ LaunchTimer lt( ctx, 7 ); CK( pvm::launch_repitch_plan( mod_hz.p, B, a.bin_width, interp, plan, ctx->stream ), "repitch plan launch" ); }
		{ LaunchTimer lt( ctx, 5 ); CK( pvm::launch_repitch_shared( a, plan, mod_hz.p, rows, ctx->sms, ctx->stream ), "repitch launch" ); }
		skip_if = plan.ok;
		}
	{ LaunchTimer lt( ctx, 5 ); CK( pvm::launch_repitch( a, rows, skip_if, ctx->stream ), "repitch launch" ); }
	return FLAN_B200_OK;
	}

size_t repitch_plan_bytes( int B ) { return align_up( sizeof( float ) * 2 * (size_t) B + sizeof( int ), 256 ); }

// Reads the reduction back (synchronises the stream).
int read_map_check( flan_b200_ctx * ctx, float sr, int hop, int64_t * out_frames, bool * descends )
	{
	pvm::MapCheck h{};
	CK( cudaMemcpyAsync( &h, ctx->d_check, sizeof( h ), cudaMemcpyDeviceToHost, ctx->stream ), "map check read" );
	CK( cudaStreamSynchronize( ctx->stream ), "map check sync" );
	const float mx = pvm::key_float( h.max_key );
	const float last = std::ceil( mx * sr / float( hop ) );            // PVModify.cpp:312, PVBuffer.cpp:428-431
	*out_frames = (int64_t) pvm::to_int( last );                        // format.num_frames = last_output_frame (an int)
	*descends = h.descends != 0;
	return FLAN_B200_OK;
	}

} // namespace

extern "C" {

int flan_b200_repitch( flan_b200_ctx * ctx, const float * d_pv, int C, int64_t F, int B, float sr,
                       const float * d_factor, int64_t factor_frame_stride, int factor_bin_stride,
                       int interp, float * d_pv_out )
	{
	if( !ctx || !d_pv || !d_factor || !d_pv_out ) return FLAN_B200_INVALID;
	int rc = check_pv_shape( ctx, C, F, B, sr, interp );
	if( rc ) return rc;
	if( !strides_ok( factor_frame_stride, factor_bin_stride, B ) )
		return fail( ctx, FLAN_B200_INVALID, "table strides must be (B,1), (0,1), (1,0) or (0,0)" );
	const int64_t rows = factor_frame_stride ? F : 1;
	void * ws = nullptr;
	const size_t hz_bytes = align_up( sizeof( float ) * (size_t) rows * B, 256 );
	rc = get_workspace( ctx, hz_bytes + repitch_plan_bytes( B ), &ws );
	if( rc ) return rc;
	ctx->seg_key.valid = false;
	const pvm::Table factor{ d_factor, factor_frame_stride, factor_bin_stride };
	{ LaunchTimer lt( ctx, 7 ); CK( pvm::launch_bin_prefix( factor, rows, B, sr, float( ( B - 1 ) * 2 ), (float *) ws, ctx->stream ), "repitch table launch" ); }
	const pvm::Table mod{ (const float *) ws, factor_frame_stride ? (int64_t) B : 0, 1 };
	return repitch_common( ctx, d_pv, C, F, B, sr, mod, nullptr, interp, d_pv_out, (char *) ws + hz_bytes );
	}

int flan_b200_modify_frequency( flan_b200_ctx * ctx, const float * d_pv, int C, int64_t F, int B, float sr,
                                const float * d_mod_hz, int64_t mod_frame_stride, int mod_bin_stride,
                                const float * d_in_mod, int interp, float * d_pv_out )
	{
	if( !ctx || !d_pv || !d_mod_hz || !d_in_mod || !d_pv_out ) return FLAN_B200_INVALID;
	int rc = check_pv_shape( ctx, C, F, B, sr, interp );
	if( rc ) return rc;
	if( !strides_ok( mod_frame_stride, mod_bin_stride, B ) )
		return fail( ctx, FLAN_B200_INVALID, "table strides must be (B,1), (0,1), (1,0) or (0,0)" );
	const pvm::Table mod{ d_mod_hz, mod_frame_stride, mod_bin_stride };
	void * ws = nullptr;
	rc = get_workspace( ctx, repitch_plan_bytes( B ), &ws );
	if( rc ) return rc;
	ctx->seg_key.valid = false;
	return repitch_common( ctx, d_pv, C, F, B, sr, mod, d_in_mod, interp, d_pv_out, ws );
	}

int flan_b200_stretch_map( flan_b200_ctx * ctx, const float * d_factor, int64_t factor_frame_stride, int factor_bin_stride,
                           int64_t F, int B, float sr, float ar, float * d_map_out )
	{
	if( !ctx || !d_factor || !d_map_out ) return FLAN_B200_INVALID;
	if( F < 1 || B < 2 || !( sr > 0.0f ) || !( ar > 0.0f ) ) return fail( ctx, FLAN_B200_INVALID, "bad shape or rates" );
	if( !strides_ok( factor_frame_stride, factor_bin_stride, B ) )
		return fail( ctx, FLAN_B200_INVALID, "table strides must be (B,1), (0,1), (1,0) or (0,0)" );
	const int hop = flan_b200_hop_from_rates( sr, ar );
	if( hop < 1 ) return fail( ctx, FLAN_B200_INVALID, "hop < 1" );
	const int cols = factor_bin_stride ? B : 1;
	const pvm::Table factor{ d_factor, factor_frame_stride, factor_bin_stride };
	void * ws = nullptr;
	const bool constant = factor_frame_stride == 0 && factor_bin_stride == 0;
	int rc = get_workspace( ctx, constant ? pvm::constant_prefix_scratch_bytes() : sizeof( float ) * (size_t) F * cols, &ws );
	if( rc ) return rc;
	ctx->seg_key.valid = false;
	LaunchTimer lt( ctx, 7 );
	if( constant )      // closed form per binade instead of F dependent additions
		CK( pvm::launch_constant_prefix( d_factor, F, sr / float( hop ), ws, d_map_out, ctx->sms, ctx->stream ), "stretch map launch" );
	else
		CK( pvm::launch_frame_prefix( factor, F, cols, sr / float( hop ), (float *) ws, d_map_out, ctx->sms, ctx->stream ), "stretch map launch" );
	ctx->launches += 1;
	return FLAN_B200_OK;
	}

int flan_b200_modify_time_frames( flan_b200_ctx * ctx, const float * d_map, int64_t map_frame_stride, int map_bin_stride,
                                  int64_t F, int B, float sr, float ar, int64_t * out_frames )
	{
	if( !ctx || !d_map || !out_frames ) return FLAN_B200_INVALID;
	if( F < 1 || B < 2 || !( sr > 0.0f ) || !( ar > 0.0f ) ) return fail( ctx, FLAN_B200_INVALID, "bad shape or rates" );
	if( !strides_ok( map_frame_stride, map_bin_stride, B ) )
		return fail( ctx, FLAN_B200_INVALID, "table strides must be (B,1), (0,1), (1,0) or (0,0)" );
	const int hop = flan_b200_hop_from_rates( sr, ar );
	if( hop < 1 ) return fail( ctx, FLAN_B200_INVALID, "hop < 1" );
	const pvm::Table mod{ d_map, map_frame_stride, map_bin_stride };
	{ LaunchTimer lt( ctx, 7 ); CK( pvm::launch_map_check( mod, map_frame_stride ? F : 1, map_bin_stride ? B : 1, ctx->d_check, ctx->sms, ctx->stream ), "map check launch" ); }
	bool descends = false;
	return read_map_check( ctx, sr, hop, out_frames, &descends );
	}

int flan_b200_modify_time( flan_b200_ctx * ctx, const float * d_pv, int C, int64_t F, int B, float sr, float ar,
                           const float * d_map, int64_t map_frame_stride, int map_bin_stride,
                           int interp, int64_t out_frames, float * d_pv_out )
	{
	if( !ctx || !d_pv || !d_map ) return FLAN_B200_INVALID;
	int rc = check_pv_shape( ctx, C, F, B, sr, interp );
	if( rc ) return rc;
	if( !( ar > 0.0f ) ) return fail( ctx, FLAN_B200_INVALID, "analysis_rate must be positive" );
	if( !strides_ok( map_frame_stride, map_bin_stride, B ) )
		return fail( ctx, FLAN_B200_INVALID, "table strides must be (B,1), (0,1), (1,0) or (0,0)" );
	const int hop = flan_b200_hop_from_rates( sr, ar );
	if( hop < 1 ) return fail( ctx, FLAN_B200_INVALID, "hop < 1" );
	const pvm::Table mod{ d_map, map_frame_stride, map_bin_stride };
	// The frame count and the choice between the parallel and the sequential walk both come from the map itself.
	{ LaunchTimer lt( ctx, 7 ); CK( pvm::launch_map_check( mod, map_frame_stride ? F : 1, map_bin_stride ? B : 1, ctx->d_check, ctx->sms, ctx->stream ), "map check launch" ); }
	int64_t frames = 0; bool descends = false;
	rc = read_map_check( ctx, sr, hop, &frames, &descends );
	if( rc ) return rc;
	if( frames != out_frames )
		return fail( ctx, FLAN_B200_INVALID, "out_frames does not match the map: expected " + std::to_string( frames ) );
	if( out_frames <= 0 ) return FLAN_B200_OK;
	if( !d_pv_out ) return FLAN_B200_INVALID;

	pvm::StretchArgs a{};
	a.pv = (const float2 *) d_pv; a.out = (float2 *) d_pv_out; a.mod = mod;
	a.F = F; a.out_frames = out_frames; a.B = B;
	a.sample_rate = sr; a.hop = float( hop ); a.interp = interp;
	a.chunk = 32;
	a.chunks = ( F - 1 + a.chunk - 1 ) / a.chunk;
	if( a.chunks < 1 ) a.chunks = 1;
	if( !descends && map_bin_stride == 0 && F < 0x7fffffff && out_frames < 0x7fffffff )
		{
		// one geometry for every bin: plan it once, then only the per-bin arithmetic remains
		void * ws = nullptr;
		const size_t xpos_bytes = align_up( sizeof( int ) * (size_t) F, 256 );
		rc = get_workspace( ctx, xpos_bytes + sizeof( float ) * (size_t) out_frames, &ws );
		if( rc ) return rc;
		ctx->seg_key.valid = false;
		pvm::StretchPlan plan{ (int *) ws, (float *)( (char *) ws + xpos_bytes ) };
		{ LaunchTimer lt( ctx, 7 ); CK( pvm::launch_stretch_plan( a, plan, ctx->stream ), "stretch plan launch" ); }
		{ LaunchTimer lt( ctx, 6 ); CK( pvm::launch_stretch_planned( a, plan, C, ctx->stream ), "stretch launch" ); }
		}
	else if( !descends )
		{ LaunchTimer lt( ctx, 6 ); CK( pvm::launch_stretch_parallel( a, C, ctx->stream ), "stretch launch" ); }
	else
		{
		CK( cudaMemsetAsync( d_pv_out, 0, sizeof( float2 ) * (size_t) C * out_frames * B, ctx->stream ), "output clear" );   // PVModify.cpp:317-318
		LaunchTimer lt( ctx, 6 ); CK( pvm::launch_stretch_sequential( a, C, ctx->stream ), "stretch launch" );
		}
	return FLAN_B200_OK;
	}

} // extern "C"

// ---- file formats either side of the path ------------------------------------------

namespace {

void put16( std::vector<uint8_t> & b, uint16_t v ) { b.push_back( v & 0xFF ); b.push_back( v >> 8 ); }
void put32( std::vector<uint8_t> & b, uint32_t v ) { for( int i = 0; i < 4; ++i ) b.push_back( ( v >> ( 8 * i ) ) & 0xFF ); }
void put4c( std::vector<uint8_t> & b, const char * s ) { for( int i = 0; i < 4; ++i ) b.push_back( (uint8_t)( *s ? *s++ : 0 ) ); }
uint16_t get16( const uint8_t * p ) { return (uint16_t)( p[0] | ( p[1] << 8 ) ); }
uint32_t get32( const uint8_t * p ) { return (uint32_t) p[0] | ( (uint32_t) p[1] << 8 ) | ( (uint32_t) p[2] << 16 ) | ( (uint32_t) p[3] << 24 ); }

struct FileCloser { FILE * f; ~FileCloser() { if( f ) std::fclose( f ); } };
struct PinnedBuf { void * p = nullptr; ~PinnedBuf() { if( p ) cudaFreeHost( p ); } };

constexpr int64_t IO_CHUNK_VALUES = int64_t( 1 ) << 26;      // 24-bit values per staging chunk (192 MiB of file bytes), multiple of 4096

// Streams `values` 24-bit samples between a file and the device in chunks through ctx scratch + a pinned buffer.
// encode( first_value, n_values, d_bytes ) / decode( first_value, n_values, d_bytes ) launch the codec for one chunk.
template<class Launch> int stream_file( flan_b200_ctx * ctx, FILE * f, bool writing, int64_t values, Launch launch, int64_t max_chunk = IO_CHUNK_VALUES )
	{
	const int64_t chunk = values < max_chunk ? ( values > 0 ? values : 1 ) : max_chunk;
	void * ws = nullptr;
	int rc = get_workspace( ctx, (size_t) chunk * 3 + 16, &ws );
	if( rc ) return rc;
	ctx->seg_key.valid = false;
	PinnedBuf host;
	CK( cudaMallocHost( &host.p, (size_t) chunk * 3 ), "pinned staging buffer" );
	for( int64_t v0 = 0; v0 < values; v0 += chunk )
		{
		const int64_t nv = values - v0 < chunk ? values - v0 : chunk;
		if( writing )
			{
			CK( launch( v0, nv, (uint8_t *) ws ), "codec launch" );
			CK( cudaMemcpyAsync( host.p, ws, (size_t) nv * 3, cudaMemcpyDeviceToHost, ctx->stream ), "download" );
			CK( cudaStreamSynchronize( ctx->stream ), "download sync" );
			if( std::fwrite( host.p, 1, (size_t) nv * 3, f ) != (size_t) nv * 3 ) return fail( ctx, FLAN_B200_INVALID, "short write" );
			}
		else
			{
			if( std::fread( host.p, 1, (size_t) nv * 3, f ) != (size_t) nv * 3 ) return fail( ctx, FLAN_B200_INVALID, "file is shorter than its header says" );
			CK( cudaMemcpyAsync( ws, host.p, (size_t) nv * 3, cudaMemcpyHostToDevice, ctx->stream ), "upload" );
			CK( launch( v0, nv, (uint8_t *) ws ), "codec launch" );
			CK( cudaStreamSynchronize( ctx->stream ), "upload sync" );
			}
		ctx->launches++;
		}
	return FLAN_B200_OK;
	}

struct FlanHeader { int C = 0; int64_t F = 0; int B = 0; uint32_t sr = 0, hop = 0, window = 0; long data_offset = 0; };

// Reads the chunks the way PVBuffer::load does (PVBuffer.cpp:231-250): fixed order RIFF / fmt / data.
int read_flan_header( flan_b200_ctx * ctx, FILE * f, FlanHeader & h )
	{
	uint8_t b[58];
	if( std::fread( b, 1, 58, f ) != 58 ) return fail( ctx, FLAN_B200_INVALID, "not a PV file: too short" );
	if( std::memcmp( b, "RIFF", 4 ) != 0 ) return fail( ctx, FLAN_B200_INVALID, "isn't a correctly formatted RIFF file" );
	if( std::strncmp( (const char *) b + 8, "PV", 4 ) != 0 ) return fail( ctx, FLAN_B200_INVALID, "isn't a PV file" );
	if( std::memcmp( b + 12, "fmt ", 4 ) != 0 ) return fail( ctx, FLAN_B200_INVALID, "\"fmt \" wasn't at the start of the format chunk" );
	if( get16( b + 20 ) != 1 ) return fail( ctx, FLAN_B200_INVALID, "Formatting must be 1 (signed int)." );
	h.C = get16( b + 22 ); h.F = get32( b + 24 ); h.B = (int) get32( b + 28 );
	h.sr = get32( b + 32 ); h.hop = get32( b + 36 ); h.window = get32( b + 40 );
	if( get32( b + 44 ) != 24 ) return fail( ctx, FLAN_B200_INVALID, "Bit depth must be 24." );
	if( get16( b + 48 ) != 1 ) return fail( ctx, FLAN_B200_INVALID, "PV window must be 1 (hann)." );
	if( std::memcmp( b + 50, "data", 4 ) != 0 ) return fail( ctx, FLAN_B200_INVALID, "\"data\" wasn't at the start of the data chunk" );
	h.data_offset = 58;
	return FLAN_B200_OK;
	}

struct WavHeader { int C = 0; int64_t n = 0; uint32_t sr = 0; long data_offset = 0; };

int read_wav_header( flan_b200_ctx * ctx, FILE * f, WavHeader & h )
	{
	uint8_t b[12];
	if( std::fread( b, 1, 12, f ) != 12 || std::memcmp( b, "RIFF", 4 ) != 0 || std::memcmp( b + 8, "WAVE", 4 ) != 0 )
		return fail( ctx, FLAN_B200_INVALID, "not a RIFF/WAVE file" );
	bool have_fmt = false;
	int bits = 0, tag = 0, block = 0;
	for( ;; )
		{
		uint8_t c[8];
		if( std::fread( c, 1, 8, f ) != 8 ) return fail( ctx, FLAN_B200_INVALID, "WAVE file without a data chunk" );
		const uint32_t size = get32( c + 4 );
		if( std::memcmp( c, "fmt ", 4 ) == 0 )
			{
			uint8_t m[40] = { 0 };
			const uint32_t take = size < 40 ? size : 40;
			if( size < 16 || std::fread( m, 1, take, f ) != take ) return fail( ctx, FLAN_B200_INVALID, "bad fmt chunk" );
			tag = get16( m ); h.C = get16( m + 2 ); h.sr = get32( m + 4 ); block = get16( m + 12 ); bits = get16( m + 14 );
			if( tag == 0xFFFE && size >= 26 ) tag = get16( m + 24 );        // WAVE_FORMAT_EXTENSIBLE: sub-format
			std::fseek( f, (long)( size - take + ( size & 1 ) ), SEEK_CUR );
			have_fmt = true;
			}
		else if( std::memcmp( c, "data", 4 ) == 0 )
			{
			if( !have_fmt ) return fail( ctx, FLAN_B200_INVALID, "data chunk before fmt chunk" );
			if( tag != 1 || bits != 24 || h.C < 1 || block != 3 * h.C )
				return fail( ctx, FLAN_B200_UNSUPPORTED, "only 24-bit PCM WAVE files (the reference's save format, AudioBuffer.cpp:136) are decoded on the device" );
			h.n = (int64_t) size / block;
			h.data_offset = std::ftell( f );
			return FLAN_B200_OK;
			}
		else std::fseek( f, (long)( size + ( size & 1 ) ), SEEK_CUR );
		}
	}

} // namespace

extern "C" {

int flan_b200_flan_encode( flan_b200_ctx * ctx, const float * d_pv, int64_t count, float dft_size, float sr, uint8_t * d_bytes )
	{
	if( !ctx || count < 0 || ( count && ( !d_pv || !d_bytes ) ) ) return FLAN_B200_INVALID;
	if( count == 0 ) return FLAN_B200_OK;
	{ LaunchTimer lt( ctx, 8 ); CK( pvio::launch_flan_encode( d_pv, count, dft_size, sr, d_bytes, ctx->sms, ctx->stream ), "flan encode launch" ); }
	return FLAN_B200_OK;
	}

int flan_b200_flan_decode( flan_b200_ctx * ctx, const uint8_t * d_bytes, int64_t count, float dft_size, float sr, float * d_pv )
	{
	if( !ctx || count < 0 || ( count && ( !d_pv || !d_bytes ) ) ) return FLAN_B200_INVALID;
	if( count == 0 ) return FLAN_B200_OK;
	{ LaunchTimer lt( ctx, 8 ); CK( pvio::launch_flan_decode( d_bytes, count, dft_size, sr, d_pv, ctx->sms, ctx->stream ), "flan decode launch" ); }
	return FLAN_B200_OK;
	}

int flan_b200_pcm24_encode( flan_b200_ctx * ctx, const float * d_audio, int C, int64_t n, uint8_t * d_bytes )
	{
	if( !ctx || C < 1 || n < 0 || ( n && ( !d_audio || !d_bytes ) ) ) return FLAN_B200_INVALID;
	if( n == 0 ) return FLAN_B200_OK;
	{ LaunchTimer lt( ctx, 8 ); CK( pvio::launch_pcm24_encode( d_audio, C, n, n, d_bytes, ctx->sms, ctx->stream ), "pcm24 encode launch" ); }
	return FLAN_B200_OK;
	}

int flan_b200_pcm24_decode( flan_b200_ctx * ctx, const uint8_t * d_bytes, int C, int64_t n, float * d_audio )
	{
	if( !ctx || C < 1 || n < 0 || ( n && ( !d_audio || !d_bytes ) ) ) return FLAN_B200_INVALID;
	if( n == 0 ) return FLAN_B200_OK;
	{ LaunchTimer lt( ctx, 8 ); CK( pvio::launch_pcm24_decode( d_bytes, C, n, n, d_audio, ctx->sms, ctx->stream ), "pcm24 decode launch" ); }
	return FLAN_B200_OK;
	}

int flan_b200_save_flan( flan_b200_ctx * ctx, const char * path, const float * d_pv, int C, int64_t F, int B,
                         float sr, float ar, int window_size )
	{
	if( !ctx || !path || C < 0 || F < 0 || B < 0 ) return FLAN_B200_INVALID;
	const int64_t count = (int64_t) C * F * B;
	if( count && !d_pv ) return FLAN_B200_INVALID;
	FileCloser file{ std::fopen( path, "wb" ) };
	if( !file.f ) return fail( ctx, FLAN_B200_INVALID, std::string( "Error opening " ) + path + " to write RIFF." );
	std::vector<uint8_t> h;                                              // Utility/Bytes.cpp:70-112, PVBuffer.cpp:128-139
	put4c( h, "RIFF" ); put32( h, 4 ); put4c( h, "PV" );
	put4c( h, "fmt " ); put32( h, 30 );
	put16( h, 1 ); put16( h, (uint16_t) C ); put32( h, (uint32_t) F ); put32( h, (uint32_t) B );
	put32( h, (uint32_t) sr ); put32( h, (uint32_t) flan_b200_hop_from_rates( sr, ar ) ); put32( h, (uint32_t) window_size );
	put32( h, 24 ); put16( h, 1 );
	put4c( h, "data" ); put32( h, (uint32_t)( count * 6 ) );
	if( std::fwrite( h.data(), 1, h.size(), file.f ) != h.size() ) return fail( ctx, FLAN_B200_INVALID, "short write" );
	const float dft = float( ( B - 1 ) * 2 );                            // window_size_f = get_dft_size(), PVBuffer.cpp:103
	return stream_file( ctx, file.f, true, 2 * count, [&]( int64_t v0, int64_t nv, uint8_t * d_bytes )
		{ return pvio::launch_flan_encode( d_pv + v0, nv / 2, dft, sr, d_bytes, ctx->sms, ctx->stream ); } );
	}

int flan_b200_flan_info( flan_b200_ctx * ctx, const char * path, int * C, int64_t * F, int * B, float * sr, float * rate_field, int * window_size )
	{
	if( !ctx || !path ) return FLAN_B200_INVALID;
	FileCloser file{ std::fopen( path, "rb" ) };
	if( !file.f ) return fail( ctx, FLAN_B200_INVALID, std::string( "Error opening " ) + path + " to load PV." );
	FlanHeader h;
	int rc = read_flan_header( ctx, file.f, h );
	if( rc ) return rc;
	if( C ) *C = h.C; if( F ) *F = h.F; if( B ) *B = h.B;
	if( sr ) *sr = float( h.sr ); if( rate_field ) *rate_field = float( h.hop ); if( window_size ) *window_size = (int) h.window;
	return FLAN_B200_OK;
	}

int flan_b200_load_flan( flan_b200_ctx * ctx, const char * path, float * d_pv, int64_t capacity )
	{
	if( !ctx || !path ) return FLAN_B200_INVALID;
	FileCloser file{ std::fopen( path, "rb" ) };
	if( !file.f ) return fail( ctx, FLAN_B200_INVALID, std::string( "Error opening " ) + path + " to load PV." );
	FlanHeader h;
	int rc = read_flan_header( ctx, file.f, h );
	if( rc ) return rc;
	const int64_t count = (int64_t) h.C * h.F * h.B;
	if( count > capacity || ( count && !d_pv ) ) return fail( ctx, FLAN_B200_INVALID, "destination holds fewer MF elements than the file" );
	const float dft = float( ( h.B - 1 ) * 2 ), sr = float( h.sr );
	return stream_file( ctx, file.f, false, 2 * count, [&]( int64_t v0, int64_t nv, uint8_t * d_bytes )
		{ return pvio::launch_flan_decode( d_bytes, nv / 2, dft, sr, d_pv + v0, ctx->sms, ctx->stream ); } );
	}

int flan_b200_save_wav( flan_b200_ctx * ctx, const char * path, const float * d_audio, int C, int64_t n, float sr )
	{
	if( !ctx || !path || C < 1 || n < 0 || ( n && !d_audio ) ) return FLAN_B200_INVALID;
	if( (int64_t) C * n * 3 > 0xFFFFFFFFll - 36 ) return fail( ctx, FLAN_B200_UNSUPPORTED, "signal exceeds the 4 GiB RIFF limit" );
	FileCloser file{ std::fopen( path, "wb" ) };
	if( !file.f ) return fail( ctx, FLAN_B200_INVALID, std::string( path ) + " could not be opened for saving." );
	const uint32_t data_bytes = (uint32_t)( (int64_t) C * n * 3 );
	std::vector<uint8_t> h;
	put4c( h, "RIFF" ); put32( h, 36 + data_bytes + ( data_bytes & 1 ) ); put4c( h, "WAVE" );
	put4c( h, "fmt " ); put32( h, 16 ); put16( h, 1 ); put16( h, (uint16_t) C ); put32( h, (uint32_t) sr );
	put32( h, (uint32_t) sr * 3 * C ); put16( h, (uint16_t)( 3 * C ) ); put16( h, 24 );
	put4c( h, "data" ); put32( h, data_bytes );
	if( std::fwrite( h.data(), 1, h.size(), file.f ) != h.size() ) return fail( ctx, FLAN_B200_INVALID, "short write" );
	// the interleaved order makes a chunk of values a range of FRAMES of the planar buffer
	const int64_t frames_per_chunk = ( IO_CHUNK_VALUES / C ) / 4096 * 4096;
	int rc = stream_file( ctx, file.f, true, (int64_t) C * n, [&]( int64_t v0, int64_t nv, uint8_t * d_bytes )
		{ return pvio::launch_pcm24_encode( d_audio + v0 / C, C, n, nv / C, d_bytes, ctx->sms, ctx->stream ); }, frames_per_chunk * C );
	if( rc ) return rc;
	if( data_bytes & 1 ) { const uint8_t pad = 0; std::fwrite( &pad, 1, 1, file.f ); }
	return FLAN_B200_OK;
	}

int flan_b200_wav_info( flan_b200_ctx * ctx, const char * path, int * C, int64_t * n, float * sr )
	{
	if( !ctx || !path ) return FLAN_B200_INVALID;
	FileCloser file{ std::fopen( path, "rb" ) };
	if( !file.f ) return fail( ctx, FLAN_B200_INVALID, std::string( path ) + " could not be opened." );
	WavHeader h;
	int rc = read_wav_header( ctx, file.f, h );
	if( rc ) return rc;
	if( C ) *C = h.C; if( n ) *n = h.n; if( sr ) *sr = float( h.sr );
	return FLAN_B200_OK;
	}

int flan_b200_load_wav( flan_b200_ctx * ctx, const char * path, float * d_audio, int64_t capacity )
	{
	if( !ctx || !path ) return FLAN_B200_INVALID;
	FileCloser file{ std::fopen( path, "rb" ) };
	if( !file.f ) return fail( ctx, FLAN_B200_INVALID, std::string( path ) + " could not be opened." );
	WavHeader h;
	int rc = read_wav_header( ctx, file.f, h );
	if( rc ) return rc;
	const int64_t values = (int64_t) h.C * h.n;
	if( values > capacity || ( values && !d_audio ) ) return fail( ctx, FLAN_B200_INVALID, "destination holds fewer samples than the file" );
	std::fseek( file.f, h.data_offset, SEEK_SET );
	const int64_t frames_per_chunk = ( IO_CHUNK_VALUES / h.C ) / 4096 * 4096;
	return stream_file( ctx, file.f, false, values, [&]( int64_t v0, int64_t nv, uint8_t * d_bytes )
		{ return pvio::launch_pcm24_decode( d_bytes, h.C, h.n, nv / h.C, d_audio + v0 / h.C, ctx->sms, ctx->stream ); }, frames_per_chunk * h.C );
	}

} // extern "C"

