// oracle/ref_shim: TEST INFRASTRUCTURE ONLY (never linked into the product).
// The few Audio members AudioPV.cpp calls that live in reference files which cannot be compiled
// here (Audio/AudioConversions.cpp pulls in r8brain/WDL; Audio/AudioConstructors.cpp pulls in
// libsndfile through AudioBuffer.cpp). Restated from the cited lines.
#include "flan/Audio/Audio.h"

using namespace flan;

// Audio/AudioConstructors.cpp:19-23
Audio Audio::create_null()
	{
	std::cout << "Null Audio created";
	return Audio();
	}

// Audio/AudioConstructors.cpp:14-17 + AudioBuffer.cpp:46-52
Audio Audio::copy() const
	{
	Audio out( get_format() );
	out.get_buffer() = get_buffer();
	return out;
	}

// Audio/AudioConversions.cpp:32-51
Audio Audio::convert_to_mid_side() const
	{
	if( is_null() ) return Audio::create_null();

	if( get_num_channels() != 2 )
		{
		std::cout << "Can't transform non-stereo Audio between Mid-Side and Left-Right formats." << std::endl;
		return copy();
		}

	const float sqrt2 = std::sqrt( 2.0f );

	Audio out( get_format() );
	for( Frame frame = 0; frame < get_num_frames(); ++frame )
		{
		out.get_sample( 0, frame ) = ( get_sample( 0, frame ) + get_sample( 1, frame ) ) / sqrt2;
		out.get_sample( 1, frame ) = ( get_sample( 0, frame ) - get_sample( 1, frame ) ) / sqrt2;
		}
	return out;
	}

// Audio/AudioConversions.cpp:53-56
Audio Audio::convert_to_left_right() const
	{
	return convert_to_mid_side();
	}
