// flan_b200/csrc/pv_capi_exchange.cu -- the two exchanges of frame-range sharded resynthesis between PROCESSES
// (one process per GPU): the phase state of every earlier shard and the overlap-add halo of the next one
// travel as plain device-to-device copies over NVLink into mailboxes that the peers opened through CUDA IPC, ordered by
// 32-bit sequence flags that the receiver's stream waits on (cuStreamWaitValue32). No kernel of the exchange runs on
// an SM -- a NCCL send / recv pair beside the resynthesis kernel spins on several SMs and cost 9 % of that kernel
// (bench.py, 2 GPUs: 1.57 -> 1.72 ms) -- and nothing on the host synchronises.
//
// Protocol. Every rank owns one mailbox allocation:
//     state slots   [2 parities][world]   the phase state (flan_b200_phase_state[C * B]) pushed by each EARLIER rank
//     halo slots    [2 parities]          the window - hop partial sums pushed by the NEXT rank
//     flags         state_count[2] (arrivals per parity: every pusher adds one per step), halo_flag[2] (sequence number of the
//                                                         data in the slot), both written by the pusher
//                                                         AFTER the data, in stream order: a copy completes before the next starts)
//     acks          state_ack[world], halo_ack             last sequence number the receiver has consumed (written by the
//                                                         receiver into the PUSHER's mailbox): a pusher overwrites the slot
//                                                         of sequence s only after s - 2 was consumed
// Sequence numbers count the steps (calls of flan_b200_exchange_put_state) from 1; every rank makes the same calls in the
// same order. The 4-byte flag values are copied from a ring of pinned host words filled at enqueue time.
#include "pv_ctx.h"

#include <cuda.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

using namespace pvrt;

namespace {

typedef CUresult ( *WaitValue32Fn )( CUstream, CUdeviceptr, cuuint32_t, unsigned int );

WaitValue32Fn wait_value32()
	{
	static WaitValue32Fn fn = []() -> WaitValue32Fn
		{
		void * p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if( cudaGetDriverEntryPoint( "cuStreamWaitValue32", &p, cudaEnableDefault, &q ) != cudaSuccess || q != cudaDriverEntryPointSuccess )
			{ cudaGetLastError(); return nullptr; }
		return (WaitValue32Fn) p;
		}();
	return fn;
	}

constexpr int SEQ_RING = 4096;

} // namespace

struct flan_b200_exchange
	{
	flan_b200_ctx * ctx = nullptr;
	int rank = 0, world = 1;
	size_t state_bytes = 0, halo_bytes = 0;          // per slot
	char * box = nullptr;                            // the local mailbox
	size_t box_bytes = 0;
	std::vector<char *> peer;                        // the others' mailboxes (own entry = box)
	std::vector<char> opened;
	uint32_t * seq_ring = nullptr;                   // pinned host words holding the flag values in flight
	int64_t seq_next = 0;
	uint32_t seq = 0;                                // current step
	cudaStream_t copy = nullptr;                     // halo pushes run here, beside the kernels
	cudaEvent_t ev = nullptr, ev_halo = nullptr;
	cudaEvent_t ev_pushed[2] = { nullptr, nullptr }; bool pushed_recorded[2] = { false, false };   // the state pushes of each parity have left the own slot

	size_t off_state( int parity, int src ) const { return ( (size_t) parity * world + src ) * state_bytes; }
	size_t off_halo( int parity ) const { return align_up( 2 * (size_t) world * state_bytes, 256 ) + (size_t) parity * halo_bytes; }
	size_t off_words() const { return off_halo( 0 ) + 2 * halo_bytes; }
	// word indices
	int w_state_flag( int parity, int src ) const { return parity * world + src; }
	int w_halo_flag( int parity ) const { return 2 * world + parity; }
	int w_state_ack( int dst ) const { return 2 * world + 2 + dst; }
	int w_halo_ack() const { return 3 * world + 2; }
	int words() const { return 3 * world + 3; }
	uint32_t * word( char * base, int i ) const { return (uint32_t *)( base + off_words() ) + i; }
	const uint32_t * seq_word( uint32_t v ) { uint32_t * p = seq_ring + ( seq_next++ % SEQ_RING ); *p = v; return p; }
	};

namespace {

int wait_geq( flan_b200_exchange * ex, cudaStream_t st, uint32_t * addr, uint32_t value )
	{
	flan_b200_ctx * ctx = ex->ctx;
	WaitValue32Fn fn = wait_value32();
	if( !fn ) return fail( ctx, FLAN_B200_UNSUPPORTED, "cuStreamWaitValue32 is not available" );
	if( fn( (CUstream) st, (CUdeviceptr) (uintptr_t) addr, value, CU_STREAM_WAIT_VALUE_GEQ ) != CUDA_SUCCESS )
		return fail( ctx, FLAN_B200_CUDA, "cuStreamWaitValue32 failed" );
	return FLAN_B200_OK;
	}

int put_word( flan_b200_exchange * ex, cudaStream_t st, uint32_t * peer_addr, uint32_t value )
	{
	flan_b200_ctx * ctx = ex->ctx;
	CK( cudaMemcpyAsync( peer_addr, ex->seq_word( value ), 4, cudaMemcpyHostToDevice, st ), "flag copy" );
	return FLAN_B200_OK;
	}

} // namespace

extern "C" {

int flan_b200_exchange_create( flan_b200_ctx * ctx, int rank, int world, int channels, int bins, int64_t halo_samples, flan_b200_exchange ** out )
	{
	if( !ctx || !out || world < 1 || world > FLAN_B200_MAX_DEVICES || rank < 0 || rank >= world || channels < 1 || bins < 2 || halo_samples < 0 ) return FLAN_B200_INVALID;
	*out = nullptr;
	CallLock lock( ctx );
	if( !wait_value32() ) return fail( ctx, FLAN_B200_UNSUPPORTED, "cuStreamWaitValue32 is not available" );
	flan_b200_exchange * ex = new flan_b200_exchange;
	ex->ctx = ctx; ex->rank = rank; ex->world = world;
	ex->state_bytes = sizeof( flan_b200_phase_state ) * (size_t) channels * bins;    // exact: the slots of one parity are the contiguous [world][C][B] array flan_b200_phase_carry reads
	ex->halo_bytes = align_up( sizeof( float ) * (size_t) channels * (size_t) std::max<int64_t>( halo_samples, 1 ), 256 );
	ex->box_bytes = ex->off_words() + align_up( sizeof( uint32_t ) * ex->words(), 256 );
	ex->peer.assign( world, nullptr ); ex->opened.assign( world, 0 );
	cudaError_t e = cudaMalloc( (void **) &ex->box, ex->box_bytes );       // its own allocation: the IPC handle names exactly this
	if( e == cudaSuccess ) e = cudaMemset( ex->box, 0, ex->box_bytes );
	if( e == cudaSuccess ) e = cudaHostAlloc( (void **) &ex->seq_ring, sizeof( uint32_t ) * SEQ_RING, cudaHostAllocDefault );
	int lo = 0, hi = 0;
	if( e == cudaSuccess ) e = cudaDeviceGetStreamPriorityRange( &lo, &hi );
	if( e == cudaSuccess ) e = cudaStreamCreateWithPriority( &ex->copy, cudaStreamNonBlocking, hi );
	for( cudaEvent_t * v : { &ex->ev, &ex->ev_halo, &ex->ev_pushed[0], &ex->ev_pushed[1] } )
		if( e == cudaSuccess ) e = cudaEventCreateWithFlags( v, cudaEventDisableTiming );
	if( e == cudaSuccess ) e = cudaDeviceSynchronize();
	if( e != cudaSuccess ) { cudaGetLastError(); flan_b200_exchange_destroy( ex ); return cuda_fail( ctx, e, "exchange create" ); }
	ex->peer[rank] = ex->box;
	*out = ex;
	return FLAN_B200_OK;
	}

void flan_b200_exchange_destroy( flan_b200_exchange * ex )
	{
	if( !ex ) return;
	cudaSetDevice( ex->ctx->device );
	cudaDeviceSynchronize();
	for( int r = 0; r < ex->world; ++r ) if( ex->opened[r] ) cudaIpcCloseMemHandle( ex->peer[r] );
	if( ex->copy ) cudaStreamDestroy( ex->copy );
	for( cudaEvent_t v : { ex->ev, ex->ev_halo, ex->ev_pushed[0], ex->ev_pushed[1] } ) if( v ) cudaEventDestroy( v );
	if( ex->box ) cudaFree( ex->box );
	if( ex->seq_ring ) cudaFreeHost( ex->seq_ring );
	delete ex;
	}

int flan_b200_exchange_handle( flan_b200_exchange * ex, void * handle64 )
	{
	if( !ex || !handle64 ) return FLAN_B200_INVALID;
	flan_b200_ctx * ctx = ex->ctx;
	CallLock lock( ctx );
	static_assert( sizeof( cudaIpcMemHandle_t ) == FLAN_B200_IPC_HANDLE_BYTES, "handle size" );
	cudaIpcMemHandle_t h;
	CK( cudaIpcGetMemHandle( &h, ex->box ), "ipc export" );
	std::memcpy( handle64, &h, sizeof( h ) );
	return FLAN_B200_OK;
	}

int flan_b200_exchange_connect( flan_b200_exchange * ex, const void * handles )
	{
	if( !ex || !handles ) return FLAN_B200_INVALID;
	flan_b200_ctx * ctx = ex->ctx;
	CallLock lock( ctx );
	for( int r = 0; r < ex->world; ++r )
		{
		if( r == ex->rank || ex->opened[r] ) continue;
		// every other rank is a peer: states go to all later ranks, acknowledgements to all earlier ones
		cudaIpcMemHandle_t h;
		std::memcpy( &h, (const char *) handles + (size_t) r * sizeof( h ), sizeof( h ) );
		void * p = nullptr;
		CK( cudaIpcOpenMemHandle( &p, h, cudaIpcMemLazyEnablePeerAccess ), "ipc open (is peer access between the GPUs available?)" );
		ex->peer[r] = (char *) p; ex->opened[r] = 1;
		}
	return FLAN_B200_OK;
	}

// Where the NEXT step's phase state of this rank should be written (flan_b200_phase_summary's output): the rank's own
// slot of its mailbox. flan_b200_exchange_put_state then pushes it from there without a staging copy.
int flan_b200_exchange_state_slot( flan_b200_exchange * ex, flan_b200_phase_state ** d_slot )
	{
	if( !ex || !d_slot ) return FLAN_B200_INVALID;
	flan_b200_ctx * ctx = ex->ctx;
	CallLock lock( ctx );
	const int parity = (int)( ( ex->seq + 1 ) & 1 );
	// the pushes of two steps ago read this slot on the copy stream
	if( ex->pushed_recorded[parity] ) CK( cudaStreamWaitEvent( ctx->compute, ex->ev_pushed[parity], 0 ), "stream wait" );
	*d_slot = (flan_b200_phase_state *)( ex->box + ex->off_state( parity, ex->rank ) );
	return FLAN_B200_OK;
	}

// Step s begins: this rank's phase state goes to every later rank -- on the exchange's copy stream, behind the kernel
// that produced it, so that the context's stream carries on with the scan meanwhile.
int flan_b200_exchange_put_state( flan_b200_exchange * ex, const flan_b200_phase_state * d_state )
	{
	if( !ex || !d_state ) return FLAN_B200_INVALID;
	flan_b200_ctx * ctx = ex->ctx;
	CallLock lock( ctx );
	BlockUse use( ctx, { d_state } );
	const uint32_t s = ++ex->seq;
	const int parity = (int)( s & 1 );
	char * own = ex->box + ex->off_state( parity, ex->rank );
	if( ex->rank + 1 >= ex->world ) return FLAN_B200_OK;
	if( (const char *) d_state != own )
		{
		if( ex->pushed_recorded[parity] ) CK( cudaStreamWaitEvent( ctx->compute, ex->ev_pushed[parity], 0 ), "stream wait" );
		CK( cudaMemcpyAsync( own, d_state, ex->state_bytes, cudaMemcpyDeviceToDevice, ctx->compute ), "state staging" );
		}
	CK( cudaEventRecord( ex->ev, ctx->compute ), "event record" );
	CK( cudaStreamWaitEvent( ex->copy, ex->ev, 0 ), "stream wait" );
	pvk::StatePush push{};
	push.src = (const uint4 *) own; push.n16 = (unsigned)( ex->state_bytes / 16 );      // 32-byte elements: a whole number of 16-byte units
	int n = 0;
	for( int dst = ex->rank + 1; dst < ex->world; ++dst, ++n )
		{
		if( !ex->peer[dst] ) return fail( ctx, FLAN_B200_INVALID, "exchange is not connected" );
		if( s > 2 ) { int rc = wait_geq( ex, ex->copy, ex->word( ex->box, ex->w_state_ack( dst ) ), s - 2 ); if( rc ) return rc; }
		push.dst[n] = (uint4 *)( ex->peer[dst] + ex->off_state( parity, ex->rank ) );
		push.count[n] = ex->word( ex->peer[dst], ex->w_state_flag( parity, 0 ) );
		}
	CK( pvk::launch_state_push( push, n, ex->copy ), "state push launch" );
	ctx->launches++;
	CK( cudaEventRecord( ex->ev_pushed[parity], ex->copy ), "event record" );
	ex->pushed_recorded[parity] = true;
	return FLAN_B200_OK;
	}

// The states of ranks 0 .. rank-1 of the current step, contiguous ([rank][C][B] flan_b200_phase_state): the context's
// stream waits for each of them. The pointer stays valid until the step after next.
int flan_b200_exchange_get_states( flan_b200_exchange * ex, const flan_b200_phase_state ** d_states )
	{
	if( !ex || !d_states ) return FLAN_B200_INVALID;
	flan_b200_ctx * ctx = ex->ctx;
	CallLock lock( ctx );
	const uint32_t s = ex->seq;
	const int parity = (int)( s & 1 );
	// every earlier rank has incremented the counter of this parity once per step of this parity; none can be a step ahead
	// on it (a pusher overwrites a slot only after the acknowledgement of the step two before)
	if( ex->rank > 0 )
		{
		int rc = wait_geq( ex, ctx->compute, ex->word( ex->box, ex->w_state_flag( parity, 0 ) ), (uint32_t) ex->rank * ( ( s + 1 ) / 2 ) );
		if( rc ) return rc;
		}
	*d_states = (const flan_b200_phase_state *)( ex->box + ex->off_state( parity, 0 ) );
	return FLAN_B200_OK;
	}

// ... and once the kernel that reads them is enqueued: tell the pushers.
int flan_b200_exchange_release_states( flan_b200_exchange * ex )
	{
	if( !ex ) return FLAN_B200_INVALID;
	flan_b200_ctx * ctx = ex->ctx;
	CallLock lock( ctx );
	if( ex->rank == 0 ) return FLAN_B200_OK;
	// off the critical path: the acknowledgements leave on the copy stream, behind the kernel that read the states
	CK( cudaEventRecord( ex->ev, ctx->compute ), "event record" );
	CK( cudaStreamWaitEvent( ex->copy, ex->ev, 0 ), "stream wait" );
	for( int src = 0; src < ex->rank; ++src )
		{
		int rc = put_word( ex, ex->copy, ex->word( ex->peer[src], ex->w_state_ack( ex->rank ) ), ex->seq );
		if( rc ) return rc;
		}
	return FLAN_B200_OK;
	}

// The head of this rank's span (channels rows of n samples, pitch elements apart) goes to the previous rank, on the
// exchange's copy stream, after `after_event` (the event flan_b200_convert_to_audio_range_head records).
int flan_b200_exchange_put_halo( flan_b200_exchange * ex, const float * d_head, int64_t pitch, int channels, int64_t n, void * after_event )
	{
	if( !ex || !d_head || ex->rank == 0 || channels < 1 || n < 1 ) return FLAN_B200_INVALID;
	flan_b200_ctx * ctx = ex->ctx;
	if( sizeof( float ) * (size_t) channels * (size_t) n > ex->halo_bytes ) return fail( ctx, FLAN_B200_INVALID, "halo larger than the mailbox slot" );
	CallLock lock( ctx );
	const uint32_t s = ex->seq;
	const int parity = (int)( s & 1 ), dst = ex->rank - 1;
	if( !ex->peer[dst] ) return fail( ctx, FLAN_B200_INVALID, "exchange is not connected" );
	if( after_event ) CK( cudaStreamWaitEvent( ex->copy, (cudaEvent_t) after_event, 0 ), "stream wait" );
	if( s > 2 ) { int rc = wait_geq( ex, ex->copy, ex->word( ex->box, ex->w_halo_ack() ), s - 2 ); if( rc ) return rc; }
	CK( cudaMemcpy2DAsync( ex->peer[dst] + ex->off_halo( parity ), sizeof( float ) * (size_t) n, d_head, sizeof( float ) * (size_t) pitch,
	                       sizeof( float ) * (size_t) n, (size_t) channels, cudaMemcpyDeviceToDevice, ex->copy ), "halo push" );
	int rc = put_word( ex, ex->copy, ex->word( ex->peer[dst], ex->w_halo_flag( parity ) ), s );
	if( rc ) return rc;
	// the span must not be recycled or overwritten before the push has read it: the context's stream joins the copy stream
	CK( cudaEventRecord( ex->ev_halo, ex->copy ), "event record" );
	CK( cudaStreamWaitEvent( ctx->compute, ex->ev_halo, 0 ), "stream wait" );
	return FLAN_B200_OK;
	}

// d_out[c * pitch + i] += halo of the next rank, i < n (lower-frame contributions first, AudioPV.cpp:133-134: call it
// after this rank's own frames): the context's stream waits for the push, adds, and acknowledges.
int flan_b200_exchange_add_halo( flan_b200_exchange * ex, float * d_out, int64_t pitch, int channels, int64_t n )
	{
	if( !ex || !d_out || ex->rank + 1 >= ex->world || channels < 1 || n < 1 ) return FLAN_B200_INVALID;
	flan_b200_ctx * ctx = ex->ctx;
	if( sizeof( float ) * (size_t) channels * (size_t) n > ex->halo_bytes ) return fail( ctx, FLAN_B200_INVALID, "halo larger than the mailbox slot" );
	CallLock lock( ctx );
	BlockUse use( ctx, { d_out } );
	const uint32_t s = ex->seq;
	const int parity = (int)( s & 1 ), src = ex->rank + 1;
	int rc = wait_geq( ex, ctx->compute, ex->word( ex->box, ex->w_halo_flag( parity ) ), s );
	if( rc ) return rc;
	const float * halo = (const float *)( ex->box + ex->off_halo( parity ) );
	for( int c = 0; c < channels; ++c )
		{
		rc = flan_b200_add( ctx, d_out + (int64_t) c * pitch, halo + (int64_t) c * n, n );
		if( rc ) return rc;
		}
	CK( cudaEventRecord( ex->ev, ctx->compute ), "event record" );
	CK( cudaStreamWaitEvent( ex->copy, ex->ev, 0 ), "stream wait" );
	return put_word( ex, ex->copy, ex->word( ex->peer[src], ex->w_halo_ack() ), s );
	}

} // extern "C"
