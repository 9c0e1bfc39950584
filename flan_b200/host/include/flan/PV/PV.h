// flan::PV of the B200 build: the conversion entry points of the reference's PV (src/flan/PV/PV.h:27-490) on the
// phase-vocoder path, with the reference's signatures (PV.h:88-96).
#pragma once

#include "flan/PV/PVBuffer.h"

namespace flan {

class Audio;

class PV : public PVBuffer
{
public:
	PV() : PVBuffer( PVBuffer::Format() ) {}
	PV( PVBuffer && other ) : PVBuffer( std::move( other ) ) {}

	static PV create_null() { return PVBuffer(); }

	/** Phase accumulation, inverse FFT and windowed overlap-add (reference Conversions/AudioPV.cpp:86-139), on the GPU.
	 *  A NaN/Inf in the data prints the reference's warning and conversion continues (AudioPV.cpp:88-89). */
	Audio convert_to_audio( flan_CANCEL_ARG ) const;

	/** convert_to_audio() then convert_to_left_right(); null unless stereo (AudioPV.cpp:141-145). */
	Audio convert_to_lr_audio( flan_CANCEL_ARG ) const;
};

}
