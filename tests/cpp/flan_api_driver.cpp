// tests/cpp/flan_api_driver.cpp -- user code written against the reference's C++ API (the shape of the reference's
// own tests/flanTest.cpp:39-44: audio.convert_to_PV(...).<PV op>.convert_to_audio()), compiled against the B200
// build's headers. Exposed through extern "C" so pytest can feed it numpy buffers and compare with the oracle.
#include "flan/Audio/Audio.h"
#include "flan/PV/PV.h"

#include <cstring>

using namespace flan;

extern "C" {

// Returns frames, or -1 if the API returned a null PV. mode: 0 convert_to_PV, 1 convert_to_ms_PV
int api_convert_to_pv( const float * audio, int C, int n, float sr, int W, int hop, int N, int mode, float * pv_out, float * ar_out )
	{
	Audio a = Audio::create_from_buffer( std::vector<float>( audio, audio + size_t( C ) * n ), C, sr );
	PV pv = mode ? a.convert_to_ms_PV( W, hop, N ) : a.convert_to_PV( W, hop, N );
	if( pv.is_null() ) return -1;
	*ar_out = pv.get_analysis_rate();
	std::memcpy( pv_out, pv.get_buffer().data(), sizeof( MF ) * pv.get_buffer().size() );
	return pv.get_num_frames();
	}

// Device-resident chain: analysis, an in-place host edit of one MF through the reference-style accessor (to exercise
// the dirty tracking), resynthesis. Returns output frames or -1.
int api_round_trip( const float * audio, int C, int n, float sr, int W, int hop, int N, int touch, int lr, float * audio_out )
	{
	Audio a = Audio::create_from_buffer( std::vector<float>( audio, audio + size_t( C ) * n ), C, sr );
	PV pv = a.convert_to_PV( W, hop, N );
	if( pv.is_null() ) return -1;
	if( touch ) pv.get_MF( 0, 1, 3 ).m *= 2.0f;
	Audio out = lr ? pv.convert_to_lr_audio() : pv.convert_to_audio();
	if( out.is_null() ) return -1;
	std::memcpy( audio_out, out.get_buffer().data(), sizeof( float ) * out.get_buffer().size() );
	return out.get_num_frames();
	}

// PV given on the host (e.g. loaded or edited by user code) -> audio.
int api_convert_to_audio( const float * pv, int C, int F, int B, float sr, float ar, int W, float * audio_out )
	{
	PVBuffer::Format f;
	f.num_channels = C; f.num_frames = F; f.num_bins = B; f.sample_rate = sr; f.analysis_rate = ar; f.window_size = W;
	PV p( ( PVBuffer( f ) ) );
	std::memcpy( p.get_buffer().data(), pv, sizeof( MF ) * size_t( C ) * F * B );
	Audio out = p.convert_to_audio();
	if( out.is_null() ) return -1;
	std::memcpy( audio_out, out.get_buffer().data(), sizeof( float ) * out.get_buffer().size() );
	return out.get_num_frames();
	}

// Cancellation protocol: a raised canceller yields a null object (AudioPV.cpp:49,115).
int api_cancelled_is_null( void )
	{
	std::atomic<bool> cancel( true );
	AudioBuffer::Format f; f.num_channels = 1; f.num_frames = 4096; f.sample_rate = 48000;
	Audio a = Audio::create_from_format( f );
	PV pv = a.convert_to_PV( 256, 32, 256, cancel );
	return pv.is_null() ? 1 : 0;
	}


// BASELINE config 4's chain the way user code writes it (the reference's tests/flanTest.cpp:39-44):
//     audio.convert_to_PV( W, hop, N ).repitch( ... ).stretch( ... ).convert_to_audio()
// kind 0: constants (repitch 1.5, stretch 2), 1: lambdas of time / frequency. pv_out (may be null) receives the PV after
// both PV-domain steps, audio_out the resynthesis. Returns output samples per channel, or -1 on a null result.
int api_chain( const float * audio, int C, int n, float sr, int W, int hop, int N, int kind, int interp,
               float * pv_out, int * pv_frames, float * audio_out )
	{
	Audio a = Audio::create_from_buffer( std::vector<float>( audio, audio + size_t( C ) * n ), C, sr );
	PV pv = a.convert_to_PV( W, hop, N );
	if( pv.is_null() ) return -1;
	auto make_interp = [interp]
		{
		switch( interp )
			{
			case 5: return Interpolator::smoothstep();
			case 2: return Interpolator::nearest();
			default: return Interpolator::linear();
			}
		};
	PV shaped = kind == 0
		? pv.repitch( 1.5f, make_interp() ).stretch( 2.0f, make_interp() )
		: pv.repitch( []( TF tf ) { return 0.75f + 0.5f * tf.t; }, make_interp() )
		    .stretch( []( TF tf ) { return 1.0f + tf.f / 24000.0f; }, make_interp() );
	if( shaped.is_null() ) return -1;
	*pv_frames = shaped.get_num_frames();
	if( pv_out ) std::memcpy( pv_out, shaped.get_buffer().data(), sizeof( MF ) * shaped.get_buffer().size() );
	Audio out = shaped.convert_to_audio();
	if( out.is_null() ) return -1;
	if( audio_out ) std::memcpy( audio_out, out.get_buffer().data(), sizeof( float ) * out.get_buffer().size() );
	return out.get_num_frames();
	}

// modify_time / modify_frequency with lambdas; a user-callable Interpolator has no device form -> null PV.
int api_modify_maps( const float * audio, int C, int n, float sr, int W, int hop, int N, float * pv_time_out, int * time_frames,
                     float * pv_freq_out )
	{
	Audio a = Audio::create_from_buffer( std::vector<float>( audio, audio + size_t( C ) * n ), C, sr );
	PV pv = a.convert_to_PV( W, hop, N );
	if( pv.is_null() ) return -1;
	PV t = pv.modify_time( []( TF tf ) { return tf.t * 1.25f + 0.01f; } );
	PV f = pv.modify_frequency( []( TF tf ) { return tf.f * 0.8f + 30.0f; }, Interpolator::smoothstep() );
	if( t.is_null() || f.is_null() ) return -1;
	*time_frames = t.get_num_frames();
	std::memcpy( pv_time_out, t.get_buffer().data(), sizeof( MF ) * t.get_buffer().size() );
	std::memcpy( pv_freq_out, f.get_buffer().data(), sizeof( MF ) * f.get_buffer().size() );
	PV custom = pv.repitch( 1.5f, Interpolator( []( float x ) { return x * x; } ) );
	return custom.is_null() ? 1 : 0;
	}


// Files either side of the path, as user code does it (the reference's tests/flanTest.cpp:34-44 loads a WAV, converts and
// saves): audio -> save WAV -> load WAV -> convert_to_PV -> save .flan -> load .flan -> fix the rate -> convert_to_audio.
// Returns output samples per channel or a negative code.
int api_file_round_trip( const float * audio, int C, int n, float sr, int W, int hop, int N, const char * wav_path,
                         const char * flan_path, float * loaded_audio, float * loaded_pv, float * loaded_rate, float * audio_out )
	{
	Audio a = Audio::create_from_buffer( std::vector<float>( audio, audio + size_t( C ) * n ), C, sr );
	if( !a.save( wav_path ) ) return -1;
	Audio b;
	if( !b.load( wav_path ) || b.get_num_channels() != C || b.get_num_frames() != n ) return -2;
	std::memcpy( loaded_audio, b.get_buffer().data(), sizeof( float ) * b.get_buffer().size() );
	PV pv = b.convert_to_PV( W, hop, N );
	if( pv.is_null() || !pv.save( flan_path ) ) return -3;
	PV q;
	if( !q.load( flan_path ) || q.get_num_frames() != pv.get_num_frames() || q.get_num_bins() != pv.get_num_bins() ) return -4;
	*loaded_rate = q.get_analysis_rate();
	std::memcpy( loaded_pv, q.get_buffer().data(), sizeof( MF ) * q.get_buffer().size() );
	// the loaded analysis rate is the hop (reference quirk): rebuild the buffer with the true rate before resynthesis
	PVBuffer::Format f = q.get_format();
	f.analysis_rate = pv.get_analysis_rate();
	PV fixed( ( PVBuffer( f ) ) );
	std::memcpy( fixed.get_buffer().data(), q.get_buffer().data(), sizeof( MF ) * q.get_buffer().size() );
	Audio out = fixed.convert_to_audio();
	if( out.is_null() ) return -5;
	std::memcpy( audio_out, out.get_buffer().data(), sizeof( float ) * out.get_buffer().size() );
	return out.get_num_frames();
	}

}

// ---- concurrency and steady-state reuse (round 2) ----------------------------------------------------------------
#include <thread>
#include <vector>

extern "C" {

// The reference's conversions are const and re-entrant (FFTHelper.cpp:9 is its only lock): `threads` host threads run
// convert_to_PV -> convert_to_audio at the same time, each on its own Audio (thread t reads audio + t * C * n and writes
// audio_out + t * C * out_n), `reps` times each. Returns output samples per channel, or -1 if any call returned null.
int api_concurrent_round_trips( const float * audio, int threads, int reps, int C, int n, float sr, int W, int hop, int N, float * audio_out )
	{
	const int F = n / hop + 1;
	const size_t out_n = size_t( F ) * hop;
	std::vector<int> ok( threads, 1 );
	std::vector<std::thread> pool;
	for( int t = 0; t < threads; ++t )
		pool.emplace_back( [&, t]
			{
			const float * src = audio + size_t( t ) * C * n;
			for( int r = 0; r < reps; ++r )
				{
				Audio a = Audio::create_from_buffer( std::vector<float>( src, src + size_t( C ) * n ), C, sr );
				PV pv = a.convert_to_PV( W, hop, N );
				if( pv.is_null() ) { ok[t] = 0; return; }
				Audio out = pv.convert_to_audio();
				if( out.is_null() || size_t( out.get_num_frames() ) != out_n ) { ok[t] = 0; return; }
				std::memcpy( audio_out + size_t( t ) * C * out_n, out.get_buffer().data(), sizeof( float ) * C * out_n );
				}
			} );
	for( auto & th : pool ) th.join();
	for( int v : ok ) if( !v ) return -1;
	return int( out_n );
	}

// One long-lived Audio whose samples are edited on the host before every pass (get_buffer() drops the device copy), the
// way a program that synthesises or filters on the CPU between conversions behaves. From the third pass on the host
// vectors are recycled and page-locked and the download is prefetched. out[r] receives pass r's result.
int api_repeated_round_trips( const float * audio, int reps, int C, int n, float sr, int W, int hop, int N, float gain_step, float * audio_out )
	{
	const int F = n / hop + 1;
	const size_t out_n = size_t( F ) * hop;
	Audio a = Audio::create_from_buffer( std::vector<float>( audio, audio + size_t( C ) * n ), C, sr );
	for( int r = 0; r < reps; ++r )
		{
		std::vector<float> & h = a.get_buffer();
		const float g = 1.0f + gain_step * r;
		for( size_t i = 0; i < h.size(); ++i ) h[i] = audio[i] * g;
		Audio out = a.convert_to_PV( W, hop, N ).convert_to_audio();
		if( out.is_null() ) return -1;
		std::memcpy( audio_out + size_t( r ) * C * out_n, out.get_buffer().data(), sizeof( float ) * C * out_n );
		}
	return int( out_n );
	}

}
