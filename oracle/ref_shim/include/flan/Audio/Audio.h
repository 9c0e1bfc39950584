// oracle/ref_shim: TEST INFRASTRUCTURE ONLY (never linked into the product).
// Slim g++-compatible stand-in for the reference's flan/Audio/Audio.h and
// flan/Audio/AudioBuffer.h, declaring only what Conversions/AudioPV.cpp touches.
// The full headers are MSVC-only as written (Audio/AudioBuffer.h:120-124 default
// argument; PV/PV.h:338 std::_Pi) and pull in libsndfile / r8brain / WDL.
// Layout and semantics follow Audio/AudioBuffer.h:34-39 (Format),
// AudioBuffer.cpp:26-29 (zero-filled ctor), :376-379,:440-443 (get_sample),
// :479-482 (planar channel-major position, int arithmetic).
#pragma once

#include <vector>
#include <cmath>
#include <algorithm>
#include <atomic>
#include <iostream>

#include "flan/defines.h"
#include "flan/Utility/execution.h"

namespace flan {

class PV;

class AudioBuffer
{
public:
	AudioBuffer( const AudioBuffer & ) = delete;
	AudioBuffer( AudioBuffer && ) = default;
	AudioBuffer& operator=( const AudioBuffer & ) = delete;
	AudioBuffer& operator=( AudioBuffer && ) = default;
	~AudioBuffer() = default;

	struct Format
		{
		Channel num_channels = 0;
		Frame num_frames = 0;
		FrameRate sample_rate = 48000;
		};

	AudioBuffer() : format(), buffer() {}
	AudioBuffer( const Format & other )
		: format( other )
		, buffer( other.num_channels * other.num_frames )
		{}

	bool is_null() const { return buffer.empty() || get_sample_rate() == 0; }

	Sample get_sample( Channel channel, Frame frame ) const { return buffer[get_buffer_pos( channel, frame )]; }
	Sample & get_sample( Channel channel, Frame frame ) { return buffer[get_buffer_pos( channel, frame )]; }
	Format get_format() const { return format; }
	Channel get_num_channels() const { return format.num_channels; }
	Frame get_num_frames() const { return format.num_frames; }
	FrameRate get_sample_rate() const { return format.sample_rate; }
	std::vector<Sample> & get_buffer() { return buffer; }
	const std::vector<Sample> & get_buffer() const { return buffer; }
	size_t get_buffer_pos( Channel channel, Frame sample ) const { return channel * get_num_frames() + sample; }

private:
	Format format;
	std::vector<Sample> buffer;
};

class Audio : public AudioBuffer
{
public:
	Audio() : AudioBuffer() {}
	Audio( AudioBuffer && other ) : AudioBuffer( std::move( other ) ) {}
	Audio( const AudioBuffer::Format & f ) : AudioBuffer( f ) {}

	static Audio create_null();             // AudioConstructors.cpp:19-23
	Audio copy() const;                     // AudioConstructors.cpp:14-17
	Audio convert_to_mid_side() const;      // AudioConversions.cpp:32-51 (restated in audio_shim.cpp)
	Audio convert_to_left_right() const;    // AudioConversions.cpp:53-56

	// Audio/Audio.h:158-176; defined by the reference's own Conversions/AudioPV.cpp
	PV convert_to_PV( Frame window_size = 2048, Frame hop = 128, Frame dft_size = 4096, flan_CANCEL_ARG ) const;
	PV convert_to_ms_PV( Frame window_size = 2048, Frame hop = 128, Frame dft_size = 4096, flan_CANCEL_ARG ) const;
};

}
