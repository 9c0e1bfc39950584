// oracle/ref_shim: TEST INFRASTRUCTURE ONLY (never linked into the product).
// Slim stand-in for the reference's flan/PV/PV.h (MSVC-only at PV/PV.h:338).
// PVBuffer itself is the reference's own header and .cpp, compiled verbatim.
#pragma once

#include <complex>

#include "flan/PV/PVBuffer.h"
#include "flan/Utility/execution.h"

namespace flan {

class Audio;

class PV : public PVBuffer
{
public:
	PV() : PVBuffer( PVBuffer::Format() ) {}           // PV/PV.h:52
	PV( PVBuffer && other ) : PVBuffer( std::move( other ) ) {}   // PV/PV.h:56
	PV( const PVBuffer::Format & f ) : PVBuffer( f ) {}

	// PV/PV.h:88-96; defined by the reference's own Conversions/AudioPV.cpp
	Audio convert_to_audio( flan_CANCEL_ARG ) const;
	Audio convert_to_lr_audio( flan_CANCEL_ARG ) const;
};

}
