// flan_b200/csrc/emu/pv_modify_emu.cpp -- CPU emulator of the PV-domain kernel bodies in pv_modify_body.cuh.
//
// TEST HARNESS, not a product path (nothing in libflan_b200.so links or calls this): it runs the same per-thread
// phase functions the kernels of pv_modify.cu run, thread by thread between the barriers, with the host-side
// orchestration of pv_capi.cu restated, so that the scatter logic, the monotonicity vote and the chunking can be
// checked against the oracle in a container without a GPU (tests/test_emulator_modify.py). Threads of a phase are run
// in DESCENDING thread order to make any accidental dependence on the order visible.
#include "../pv_modify_body.cuh"

#include <cmath>
#include <cstring>
#include <vector>

using namespace pvm;

namespace {

void repitch_rows( const RepitchArgs & a, int64_t rows, int nt )
	{
	std::vector<unsigned char> smem( RepitchRow::bytes( a.B ) );
	RepitchRow s( smem.data(), a.B );
	for( int64_t row = 0; row < rows; ++row )
		{
		for( int t = nt - 1; t >= 0; --t ) repitch_load( a, row, t, nt, s );
		int flags = 0;
		for( int t = nt - 1; t >= 0; --t ) flags |= repitch_map( a, t, nt, s );
		for( int t = nt - 1; t >= 0; --t ) repitch_scatter( a, flags, t, nt, s );
		for( int t = nt - 1; t >= 0; --t ) repitch_store( a, row, t, nt, s );
		}
	}

bool strides_ok( int64_t fs, int bs, int B ) { return ( bs == 0 || bs == 1 ) && ( fs == 0 || fs == ( bs ? B : 1 ) ); }

int64_t map_frames( const Table & mod, int64_t rows, int cols, float sr, int hop, bool & descends )
	{
	float mx = 0.0f; bool any = false; descends = false;
	for( int64_t i = rows * cols - 1; i >= 0; --i )
		{
		float sec; bool d;
		time_map_check( mod, i / cols, (int)( i % cols ), sec, d );
		if( !any || mx < sec ) mx = sec;
		any = true; descends |= d;
		}
	mx = key_float( float_key( mx ) );
	return (int64_t) to_int( std::ceil( mx * sr / float( hop ) ) );
	}

} // namespace

extern "C" {

int pv_emu_repitch( const float * pv, int C, int64_t F, int B, float sr, const float * factor, int64_t fs, int bs,
                    const float * mod_hz, const float * in_mod, int interp, int threads, float * out )
	{
	if( !strides_ok( fs, bs, B ) ) return 1;
	const int64_t rows = fs ? F : 1;
	std::vector<float> hz( (size_t) rows * B );
	Table mod{ mod_hz, fs, bs };
	if( !mod_hz )
		{
		const Table fac{ factor, fs, bs };
		for( int64_t r = 0; r < rows; ++r ) bin_prefix_row( fac, r, B, sr, float( ( B - 1 ) * 2 ), hz.data() + r * B );
		mod = Table{ hz.data(), fs ? (int64_t) B : 0, 1 };
		}
	RepitchArgs a{};
	a.pv = (const float2 *) pv; a.out = (float2 *) out; a.mod = mod; a.in_mod = in_mod;
	a.F = F; a.B = B; a.bin_width = sr / float( ( B - 1 ) * 2 ); a.interp = interp;
	if( mod.frame_stride == 0 && mod.bin_stride == 1 )
		{
		// plan + gather, as pv_repitch_plan_kernel / pv_repitch_shared_kernel
		std::vector<int> src( B ); std::vector<float> mix( B ), pos( B ); int ok = 0;
		RepitchPlan plan{ src.data(), mix.data(), &ok };
		const int nt = threads;
		for( int t = nt - 1; t >= 0; --t ) repitch_plan_positions( mod.p, B, a.bin_width, t, nt, pos.data(), plan );
		int flags = 0;
		for( int t = nt - 1; t >= 0; --t ) flags |= repitch_plan_flags( B, t, nt, pos.data() );
		for( int t = nt - 1; t >= 0; --t ) repitch_plan_pairs( B, interp, flags, t, nt, pos.data(), plan );
		if( ok )
			{
			std::vector<float> m( B ), fm( B );
			for( int64_t row = 0; row < (int64_t) C * F; ++row )
				{
				for( int b = B - 1; b >= 0; --b )
					{
					const float2 mf = a.pv[row * B + b];
					m[b] = mf.x;
					fm[b] = in_mod ? in_mod[row * B + b] : repitch_lerp( mod.p, B, a.bin_width, mf.y );
					}
				for( int y = B - 1; y >= 0; --y ) a.out[row * B + y] = repitch_gather( y, src.data(), mix.data(), m.data(), fm.data() );
				}
			return 0;
			}
		}
	repitch_rows( a, (int64_t) C * F, threads );
	return 0;
	}

// factor != null: PV::stretch (the map is built first); else map_seconds is the time map of PV::modify_time.
// out == null: returns the output frame count only. force_sequential exercises the reference-order walk.
int64_t pv_emu_stretch( const float * pv, int C, int64_t F, int B, float sr, float ar, const float * factor,
                        const float * map_seconds, int64_t fs, int bs, int interp, int chunk, int force_sequential,
                        int * used_sequential, float * out )
	{
	if( !strides_ok( fs, bs, B ) ) return -1;
	const int hop = (int)( sr / ar );
	const int cols = bs ? B : 1;
	std::vector<float> built;
	Table mod{ map_seconds, fs, bs };
	if( factor )
		{
		built.resize( (size_t) F * cols );
		const Table fac{ factor, fs, bs };
		if( fs == 0 && bs == 0 )
			{
			std::vector<PrefixSeg> segs( PREFIX_MAX_SEGS );
			const int ns = constant_prefix_segments( *factor, F, segs.data() );
			if( ns > PREFIX_MAX_SEGS ) return -2;
			for( int64_t k = F - 1; k >= 0; --k ) built[k] = constant_prefix_value( segs.data(), ns, *factor, k ) / ( sr / float( hop ) );
			}
		else
			{
			std::vector<float> raw( (size_t) F * cols );
			for( int col = 0; col < cols; ++col ) frame_prefix_column( fac, col, F, cols, raw.data() );
			for( int64_t i = F * cols - 1; i >= 0; --i ) frame_prefix_convert( raw.data(), built.data(), i, sr / float( hop ) );
			}
		mod = Table{ built.data(), (int64_t) cols, bs };
		}
	bool descends = false;
	const int64_t out_frames = map_frames( mod, mod.frame_stride ? F : 1, cols, sr, hop, descends );
	if( used_sequential ) *used_sequential = descends || force_sequential;
	if( !out || out_frames <= 0 ) return out_frames;
	StretchArgs a{};
	a.pv = (const float2 *) pv; a.out = (float2 *) out; a.mod = mod;
	a.F = F; a.out_frames = out_frames; a.B = B; a.sample_rate = sr; a.hop = float( hop ); a.interp = interp;
	a.chunk = chunk; a.chunks = ( F - 1 + chunk - 1 ) / chunk;
	if( a.chunks < 1 ) a.chunks = 1;
	if( !descends && !force_sequential && bs == 0 )
		{
		// as on the device: plan, then the gather by segments of output frames (`chunk` doubles as the segment length here,
		// so the tests cut segments inside pairs, at pair boundaries and beyond the covered range)
		std::vector<int> xpos( F ); std::vector<float> mix( out_frames ); std::vector<int> src( out_frames, -1 );
		StretchPlan plan{ xpos.data(), mix.data(), src.data() };
		for( int64_t f = F - 1; f >= 0; --f ) stretch_plan_frame( a, plan, f );
		const int seg_len = chunk;
		const int64_t segs = ( out_frames + seg_len - 1 ) / seg_len;
		pvk::PvConsts k{};
		for( int c = C - 1; c >= 0; --c )
			for( int64_t sgm = segs - 1; sgm >= 0; --sgm )
				for( int b = B - 1; b >= 0; --b ) stretch_segment_planned<false>( a, plan, c, sgm, seg_len, b, nullptr, k );
		}
	else if( !descends && !force_sequential )
		{
		for( int c = C - 1; c >= 0; --c )
			for( int64_t k = a.chunks - 1; k >= 0; --k )
				for( int b = B - 1; b >= 0; --b ) stretch_chunk( a, c, k, b );
		}
	else
		{
		std::memset( out, 0, sizeof( float2 ) * (size_t) C * out_frames * B );
		for( int c = 0; c < C; ++c )
			for( int b = 0; b < B; ++b ) stretch_column( a, c, b );
		}
	return out_frames;
	}

}

extern "C" {

// Closed-form running sum of a constant (constant_prefix_segments / constant_prefix_value) against the plain loop:
// returns the number of mismatching elements; *segments receives the segment count.
int64_t pv_emu_constant_prefix_mismatches( float c, int64_t F, int * segments )
	{
	std::vector<PrefixSeg> segs( PREFIX_MAX_SEGS );
	const int ns = constant_prefix_segments( c, F, segs.data() );
	if( segments ) *segments = ns;
	if( ns > PREFIX_MAX_SEGS ) return -1;
	int64_t bad = 0;
	float acc = 0.0f;
	for( int64_t k = 0; k < F; ++k )
		{
		acc = k == 0 ? c : c + acc;
		const float got = constant_prefix_value( segs.data(), ns, c, k );
		if( std::memcmp( &got, &acc, 4 ) != 0 && !( got != got && acc != acc ) ) ++bad;
		}
	return bad;
	}

}
