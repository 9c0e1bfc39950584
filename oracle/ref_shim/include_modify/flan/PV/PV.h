// oracle/ref_shim: TEST INFRASTRUCTURE ONLY (never linked into the product).
// Slim stand-in for the reference's flan/PV/PV.h (MSVC-only at PV/PV.h:338) for the build of the reference's own
// PV/PVModify.cpp: the members that file defines, with the reference's signatures and defaults (PV/PV.h:31-35,
// 266-352), over the reference's own PVBuffer, Function, FunctionSample2d and Interpolator headers.
#pragma once

#include <functional>
#include <complex>

#include "flan/PV/PVBuffer.h"
#include "flan/Function.h"
#include "flan/Utility/Interpolator.h"

namespace flan {

class PV : public PVBuffer
{
public:
	template<typename T>
	FunctionSample2d<T> sample_function_over_domain( const Function<TF, T> & f ) const      // PV/PV.h:31-35
		{
		return f.sample( 0, get_num_frames(), 1.0f / get_analysis_rate(), 0, get_num_bins(), bin_to_frequency( 1 ) );
		}

	template<typename T>
	std::vector<T> sample_function_over_time_domain( const Function<Second, T> & f ) const   // PV/PV.h:37-49
		{
		std::vector<T> out( get_num_frames() );
		for( Frame frame = 0; frame < Frame( out.size() ); ++frame ) out[frame] = f( frame_to_time( frame ) );
		return out;
		}

	PV() : PVBuffer( PVBuffer::Format() ) {}
	PV( PVBuffer && other ) : PVBuffer( std::move( other ) ) {}
	PV( const PVBuffer::Format & f ) : PVBuffer( f ) {}
	PV copy() const { return PVBuffer::copy(); }

	PV modify( const Function<TF, TF> & mod, const Interpolator & interp = Interpolator::linear() ) const;
	PV modify_frequency( const Function<TF, Frequency> & mod, const Interpolator & = Interpolator::linear() ) const;
	PV modify_time( const Function<TF, Second> & mod, const Interpolator & = Interpolator::linear() ) const;
	PV repitch( const Function<TF, float> & factor, const Interpolator & = Interpolator::linear() ) const;
	PV stretch( const Function<TF, float> & factor, const Interpolator & = Interpolator::linear() ) const;
	PV stretch_spline( const Function<Second, float> & expansion ) const;
	PV desample( const Function<TF, float> & decimation_ratio, const Interpolator & interp = Interpolator::linear() ) const;
	PV smear_time( const Function<TF, Second> & smear_size, const Function<TF, int> & granularity,
	               const Function<Second, float> & distribution ) const;
	PV time_extrapolate( Second start_time, Second end_time, Second extrapolationTime,
	                     const Interpolator & interpolator = Interpolator::linear() ) const;
};

}
