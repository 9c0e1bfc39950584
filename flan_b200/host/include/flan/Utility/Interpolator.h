// flan::Interpolator of the B200 build (reference src/flan/Utility/Interpolator.h:10-59, Interpolator.cpp:15-101).
// The ten named interpolators carry an id the GPU kernels evaluate themselves (include/flan_b200.h, `interp`); a
// user-supplied callable has no device form, and the PV-domain methods refuse it with a message and a null result.
#pragma once

#include "flan/Function.h"

namespace flan {

struct Interpolator
	{
	Interpolator( const Interpolator & ) = delete;
	Interpolator & operator=( const Interpolator & ) = delete;
	Interpolator( Interpolator && ) = default;
	Interpolator & operator=( Interpolator && ) = default;

	template<typename F> requires std::convertible_to<F, std::function<float( float )>>
	Interpolator( const F & callable ) : f( callable ), device_id( -1 ) {}
	Interpolator( Function<float, float> && fn ) : f( std::move( fn ) ), device_id( -1 ) {}

	float operator()( float x ) const { return f( x ); }

	static Interpolator midpoint();         /** Constantly 0.5. */
	static Interpolator nearest();          /** Returns nearest integer. */
	static Interpolator floor();            /** Constantly 0.0. */
	static Interpolator ceil();             /** Constantly 1.0. */
	static Interpolator linear();           /** Input returning function. */
	static Interpolator smoothstep();
	static Interpolator smootherstep();
	static Interpolator sine();
	static Interpolator sine2();
	static Interpolator sqrt();

	/** Id of a named interpolator for the GPU engine (0 linear ... 9 sqrt), -1 for a user callable. */
	int get_device_id() const { return device_id; }

	Function<float, float> f;

private:
	Interpolator( Function<float, float> && fn, int id ) : f( std::move( fn ) ), device_id( id ) {}
	int device_id;
	};

}
