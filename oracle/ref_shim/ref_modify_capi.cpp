// oracle/ref_shim: TEST INFRASTRUCTURE ONLY (never linked into the product).
// extern "C" driver over the reference's own PV::repitch / PV::stretch / PV::modify_time (PV/PVModify.cpp, compiled
// verbatim from /root/reference). The factor / mod function handed to the reference is a lookup into a caller-supplied
// frame x bin table: the reference samples it at ( frame / analysis_rate, bin * bin_to_frequency(1) ) (PV/PV.h:31-35,
// Function.h:155-171), and the lookup inverts exactly that.
#include "flan/PV/PV.h"

#include <cmath>
#include <cstdint>
#include <cstring>

using namespace flan;

namespace {

Interpolator make_interp( int id )
	{
	switch( id )
		{
		case 1: return Interpolator::midpoint();
		case 2: return Interpolator::nearest();
		case 3: return Interpolator::floor();
		case 4: return Interpolator::ceil();
		case 5: return Interpolator::smoothstep();
		case 6: return Interpolator::smootherstep();
		case 7: return Interpolator::sine();
		case 8: return Interpolator::sine2();
		case 9: return Interpolator::sqrt();
		default: return Interpolator::linear();
		}
	}

PV make_pv( const float * pv, int C, int F, int B, float sr, float ar, int W )
	{
	PVBuffer::Format fmt;
	fmt.num_channels = C; fmt.num_frames = F; fmt.num_bins = B;
	fmt.sample_rate = sr; fmt.analysis_rate = ar; fmt.window_size = W;
	PV p( fmt );
	std::memcpy( p.get_buffer().data(), pv, sizeof( MF ) * (size_t) C * F * B );
	return p;
	}

Function<TF, float> table_function( const PV & p, const float * table, int F, int B )
	{
	const float ar = p.get_analysis_rate();
	const float bf1 = p.bin_to_frequency( 1 );
	return Function<TF, float>( [=]( TF tf )
		{
		long fr = std::lround( tf.t * ar ), b = std::lround( tf.f / bf1 );
		fr = fr < 0 ? 0 : ( fr >= F ? F - 1 : fr );
		b = b < 0 ? 0 : ( b >= B ? B - 1 : b );
		return table[fr * B + b];
		}, ExecutionPolicy::Linear_Sequenced );
	}

// out == nullptr: only the frame count is returned. Otherwise the result must have exactly out_frames frames.
int emit( const PV & o, float * out, int out_frames )
	{
	if( o.get_buffer().empty() ) return o.get_num_frames() > 0 ? -1 : 0;
	if( out )
		{
		if( o.get_num_frames() != out_frames ) return -2;
		std::memcpy( out, o.get_buffer().data(), sizeof( MF ) * o.get_buffer().size() );
		}
	return o.get_num_frames();
	}

}

extern "C" {

int flan_ref_repitch( const float * pv, int C, int F, int B, float sr, float ar, int W, const float * factor, int interp, float * out )
	{
	PV p = make_pv( pv, C, F, B, sr, ar, W );
	return emit( p.repitch( table_function( p, factor, F, B ), make_interp( interp ) ), out, F );
	}

int flan_ref_stretch( const float * pv, int C, int F, int B, float sr, float ar, int W, const float * factor, int interp, float * out, int out_frames )
	{
	PV p = make_pv( pv, C, F, B, sr, ar, W );
	return emit( p.stretch( table_function( p, factor, F, B ), make_interp( interp ) ), out, out_frames );
	}

int flan_ref_modify_time( const float * pv, int C, int F, int B, float sr, float ar, int W, const float * mod_seconds, int interp, float * out, int out_frames )
	{
	PV p = make_pv( pv, C, F, B, sr, ar, W );
	return emit( p.modify_time( table_function( p, mod_seconds, F, B ), make_interp( interp ) ), out, out_frames );
	}

}
