// flan::PVBuffer of the B200 build: the reference's public surface (src/flan/PV/PVBuffer.h:27-288) for everything on
// the phase-vocoder path, over device-resident storage. Layout contract: channel -> frame -> bin,
// pos = c * F * B + f * B + b (PVBuffer.cpp:526-529) -- computed in 64 bits here (the reference's int32 product
// overflows beyond 2^31 elements, which BASELINE configs 3 and 4 exceed). The .flan RIFF load / save
// (PVBuffer.cpp:99-140,216-273) quantise / dequantise on the GPU and write the reference's bytes.
#pragma once

#include <iosfwd>
#include <string>
#include <vector>

#include "flan/defines.h"
#include "flan/b200_storage.h"

namespace flan {

class PVBuffer
{
public:
	PVBuffer( const PVBuffer & ) = delete;
	PVBuffer( PVBuffer && ) = default;
	PVBuffer & operator=( const PVBuffer & ) = delete;
	PVBuffer & operator=( PVBuffer && ) = default;
	~PVBuffer() = default;

	struct Format
		{
		Channel num_channels = 0;
		Frame num_frames = 0;
		Bin num_bins = 0;
		FrameRate sample_rate = 48000;
		FrameRate analysis_rate = 48000 / 128;
		Frame window_size = 0;
		};

	PVBuffer();
	PVBuffer( const Format & format );

	PVBuffer copy() const;
	bool is_null() const;
	bool is_nan_or_inf() const;
	void print_summary() const;

	MF get_MF( Channel channel, Frame frame, Bin bin ) const;
	PVBuffer get_frame( Frame frame ) const;
	Format get_format() const;
	Channel get_num_channels() const;
	Frame get_num_frames() const;
	Bin get_num_bins() const;
	FrameRate get_sample_rate() const;
	FrameRate get_analysis_rate() const;
	Frame get_hop_size() const;
	Frame get_dft_size() const;
	Frame get_window_size() const;
	Second get_length() const;
	Frequency get_height() const;
	Magnitude get_max_partial_magnitude() const;
	fFrame time_to_frame( Second ) const;
	Second frame_to_time( fFrame ) const;
	fBin frequency_to_bin( Frequency ) const;
	Frequency bin_to_frequency( fBin ) const;
	Frequency get_frequency_offset( Channel c, Frame f, Bin b ) const;
	Channel bound_channel( Channel c ) const;
	Frame bound_frame( Frame c ) const;
	Bin bound_bin( Bin c ) const;

	void set_MF( Channel channel, Frame frame, Bin bin, MF mf );
	MF & get_MF( Channel channel, Frame frame, Bin bin );
	void clear_buffer();
	MF * get_MF_pointer( Channel channel, Frame frame, Bin bin );
	const MF * get_MF_pointer( Channel channel, Frame frame, Bin bin ) const;
	std::vector<MF> & get_buffer();
	const std::vector<MF> & get_buffer() const;
	std::vector<MF>::iterator channel_begin( Channel channel );
	std::vector<MF>::iterator channel_end( Channel channel );
	std::vector<MF>::const_iterator channel_begin( Channel channel ) const;
	std::vector<MF>::const_iterator channel_end( Channel channel ) const;
	size_t get_buffer_pos( Channel, Frame, Bin ) const;

	/** Load a .flan RIFF-PV file (reference PVBuffer::load, PVBuffer.cpp:216-273; format PVBuffer.h:84-115). As in the
	 *  reference the analysis rate comes back as the stored HOP (PVBuffer.cpp:134 vs :245). */
	bool load( const std::string & filename );
	/** Save as .flan: 24-bit magnitude / dft size and frequency / sample rate (reference PVBuffer::save, :99-140). */
	bool save( const std::string & filename ) const;

	// B200 build: device-side view for the conversion entry points (not part of the reference's surface)
	const b200::Mirror<MF> & storage() const { return buffer; }
	static PVBuffer from_device_result( const Format & format, b200::Mirror<MF> && data );

private:
	Format format;
	b200::Mirror<MF> buffer;
};

std::ostream & operator<<( std::ostream & os, const PVBuffer & flan );

}
