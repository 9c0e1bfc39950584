// flan::Audio of the B200 build: the conversion entry points of the reference's Audio (src/flan/Audio/Audio.h:25-1150)
// that lie on the phase-vocoder path, with the reference's signatures and defaults (Audio.h:158-176).
#pragma once

#include "flan/Audio/AudioBuffer.h"

namespace flan {

class PV;

class Audio : public AudioBuffer
{
public:
	Audio();
	Audio( AudioBuffer && other );

	Audio copy() const;
	static Audio create_null();
	static Audio create_from_buffer( std::vector<float> && buffer, Channel num_channels, FrameRate = 48000 );
	static Audio create_from_format( const AudioBuffer::Format & );

	/** Short-time Fourier transform + phase vocoder (reference Conversions/AudioPV.cpp:12-78), on the GPU.
	 *  dft sizes: powers of two in [256, 8192]; window_size <= dft_size. Failure or cancellation prints the
	 *  reason and returns a null PV, as the reference does. */
	PV convert_to_PV( Frame window_size = 2048, Frame hop = 128, Frame dft_size = 4096, flan_CANCEL_ARG ) const;

	/** convert_to_mid_side() then convert_to_PV(); null unless stereo (AudioPV.cpp:80-84). */
	PV convert_to_ms_PV( Frame window_size = 2048, Frame hop = 128, Frame dft_size = 4096, flan_CANCEL_ARG ) const;

	/** (L +- R) / sqrt(2) for stereo input; a copy otherwise (Audio/AudioConversions.cpp:32-51). */
	Audio convert_to_mid_side() const;
	Audio convert_to_left_right() const;
};

}
