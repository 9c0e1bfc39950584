// oracle/ref_shim: TEST INFRASTRUCTURE ONLY (never linked into the product).
// extern "C" driver over the reference's own Audio::convert_to_PV / PV::convert_to_audio
// (Conversions/AudioPV.cpp, compiled verbatim from /root/reference), so tests and the
// cpu_baseline leg of bench.py can call them through ctypes.
#include "flan/Audio/Audio.h"
#include "flan/PV/PV.h"
#include "flan/WindowFunctions.h"

#include <cstring>
#include <cstdint>

using namespace flan;

extern "C" {

// Frame count the reference would produce (AudioPV.cpp:17).
int flan_ref_num_frames( int n, int hop ) { return (int) std::ceil( n / hop ) + 1; }

// Windows::hann sampled as AudioPV.cpp:30-34 does.
void flan_ref_hann( int window_size, float * out )
	{
	for( int i = 0; i < window_size; ++i )
		out[i] = Windows::hann( float( i ) / float( window_size - 1 ) );
	}

// audio: planar float[C][n]. pv_out: MF[C][F][dft/2+1] as interleaved (m,f) floats.
// Returns F, or -1 if the reference returned a null PV. ms != 0 -> convert_to_ms_PV.
int flan_ref_convert_to_pv( const float * audio, int C, int n, float sample_rate,
	int window_size, int hop, int dft_size, int ms, float * pv_out, float * analysis_rate_out )
	{
	AudioBuffer::Format fmt;
	fmt.num_channels = C; fmt.num_frames = n; fmt.sample_rate = sample_rate;
	Audio a( fmt );
	std::memcpy( a.get_buffer().data(), audio, sizeof( float ) * (size_t) C * n );
	PV pv = ms ? a.convert_to_ms_PV( window_size, hop, dft_size ) : a.convert_to_PV( window_size, hop, dft_size );
	if( pv.get_buffer().empty() ) return -1;
	if( analysis_rate_out ) *analysis_rate_out = pv.get_analysis_rate();
	std::memcpy( pv_out, pv.get_buffer().data(), sizeof( MF ) * pv.get_buffer().size() );
	return pv.get_num_frames();
	}

// pv: MF[C][F][B]. audio_out: float[C][F*hop]. Returns samples per channel, or -1 on null.
// lr != 0 -> convert_to_lr_audio.
int flan_ref_convert_to_audio( const float * pv, int C, int F, int B, float sample_rate,
	float analysis_rate, int window_size, int lr, float * audio_out )
	{
	PVBuffer::Format fmt;
	fmt.num_channels = C; fmt.num_frames = F; fmt.num_bins = B;
	fmt.sample_rate = sample_rate; fmt.analysis_rate = analysis_rate; fmt.window_size = window_size;
	PV p( fmt );
	std::memcpy( p.get_buffer().data(), pv, sizeof( MF ) * (size_t) C * F * B );
	Audio a = lr ? p.convert_to_lr_audio() : p.convert_to_audio();
	if( a.get_buffer().empty() ) return -1;
	std::memcpy( audio_out, a.get_buffer().data(), sizeof( float ) * a.get_buffer().size() );
	return a.get_num_frames();
	}

// Timed round trip for the CPU baseline: no copies of the PV out of the reference objects.
// mode: 0 analysis only, 1 analysis + resynthesis. Returns frames processed per channel.
int flan_ref_bench( const float * audio, int C, int n, float sample_rate,
	int window_size, int hop, int dft_size, int mode, float * checksum_out )
	{
	AudioBuffer::Format fmt;
	fmt.num_channels = C; fmt.num_frames = n; fmt.sample_rate = sample_rate;
	Audio a( fmt );
	std::memcpy( a.get_buffer().data(), audio, sizeof( float ) * (size_t) C * n );
	PV pv = a.convert_to_PV( window_size, hop, dft_size );
	float cs = pv.get_buffer().empty() ? 0.0f : pv.get_buffer()[pv.get_buffer().size() / 2].m;
	if( mode == 1 )
		{
		Audio b = pv.convert_to_audio();
		if( !b.get_buffer().empty() ) cs += b.get_buffer()[b.get_buffer().size() / 2];
		}
	if( checksum_out ) *checksum_out = cs;
	return pv.get_num_frames();
	}


// The reference's own .flan writer / reader (PV/PVBuffer.cpp:99-140, 216-273). save: 0 on success.
int flan_ref_save_flan( const char * path, const float * pv, int C, int F, int B, float sample_rate, float analysis_rate, int window_size )
	{
	PVBuffer::Format fmt;
	fmt.num_channels = C; fmt.num_frames = F; fmt.num_bins = B;
	fmt.sample_rate = sample_rate; fmt.analysis_rate = analysis_rate; fmt.window_size = window_size;
	PV p( fmt );
	std::memcpy( p.get_buffer().data(), pv, sizeof( MF ) * (size_t) C * F * B );
	return p.save( path ) ? 0 : 1;
	}

// shape_out: C, F, B, window; rates_out: sample_rate, analysis_rate (as the reference's load fills them). pv_out may be
// null (shape query). Returns 0 on success.
int flan_ref_load_flan( const char * path, int * shape_out, float * rates_out, float * pv_out )
	{
	PV p;
	if( !p.load( path ) ) return 1;
	shape_out[0] = p.get_num_channels(); shape_out[1] = p.get_num_frames(); shape_out[2] = p.get_num_bins(); shape_out[3] = p.get_window_size();
	rates_out[0] = p.get_sample_rate(); rates_out[1] = p.get_analysis_rate();
	if( pv_out ) std::memcpy( pv_out, p.get_buffer().data(), sizeof( MF ) * p.get_buffer().size() );
	return 0;
	}

}
