// flan::PVBuffer of the B200 build (surface of reference src/flan/PV/PVBuffer.{h,cpp}; own implementation).
#include "flan/PV/PVBuffer.h"

#include <algorithm>
#include <cmath>
#include <iostream>
#include <ostream>

#include "flan_b200.h"

namespace flan {

static size_t element_count( const PVBuffer::Format & f )
	{
	return size_t( std::max( f.num_channels, 0 ) ) * size_t( std::max( f.num_frames, 0 ) ) * size_t( std::max( f.num_bins, 0 ) );
	}

PVBuffer::PVBuffer() : format(), buffer() {}
PVBuffer::PVBuffer( const Format & f ) : format( f ), buffer( element_count( f ) ) {}

PVBuffer PVBuffer::from_device_result( const Format & f, b200::Mirror<MF> && data )
	{
	PVBuffer out;
	out.format = f;
	out.buffer = std::move( data );
	return out;
	}

PVBuffer PVBuffer::copy() const
	{
	PVBuffer out;
	out.format = format;
	out.buffer = buffer.deep_copy();
	return out;
	}

bool PVBuffer::is_null() const { return buffer.empty() || get_sample_rate() == 0; }

bool PVBuffer::is_nan_or_inf() const
	{
	for( const MF & x : buffer.host() )
		if( std::isnan( x.m ) || std::isnan( x.f ) || std::isinf( x.m ) || std::isinf( x.f ) ) return true;
	return false;
	}

void PVBuffer::print_summary() const { std::cout << *this; }

MF PVBuffer::get_MF( Channel c, Frame f, Bin b ) const { return buffer.host()[get_buffer_pos( c, f, b )]; }

PVBuffer PVBuffer::get_frame( Frame frame ) const
	{
	Format f = format;
	f.num_frames = 1;
	PVBuffer out( f );
	for( Channel c = 0; c < get_num_channels(); ++c )
		for( Bin b = 0; b < get_num_bins(); ++b )
			out.set_MF( c, 0, b, get_MF( c, frame, b ) );
	return out;
	}

PVBuffer::Format PVBuffer::get_format() const { return format; }
Channel PVBuffer::get_num_channels() const { return format.num_channels; }
Frame PVBuffer::get_num_frames() const { return format.num_frames; }
Bin PVBuffer::get_num_bins() const { return format.num_bins; }
FrameRate PVBuffer::get_sample_rate() const { return format.sample_rate; }
FrameRate PVBuffer::get_analysis_rate() const { return format.analysis_rate; }
Frame PVBuffer::get_hop_size() const { return Frame( get_sample_rate() / get_analysis_rate() ); }   // PVBuffer.cpp:381-384
Frame PVBuffer::get_dft_size() const { return ( get_num_bins() - 1 ) * 2; }                          // PVBuffer.cpp:356-359
Frame PVBuffer::get_window_size() const { return format.window_size; }
Second PVBuffer::get_length() const { return frame_to_time( get_num_frames() ); }
Frequency PVBuffer::get_height() const { return bin_to_frequency( get_num_bins() ); }

Magnitude PVBuffer::get_max_partial_magnitude() const
	{
	Magnitude best = 0;
	for( const MF & x : buffer.host() ) best = std::max( best, std::abs( x.m ) );
	return best;
	}

fFrame PVBuffer::time_to_frame( Second t ) const { return t * float( get_sample_rate() ) / float( get_hop_size() ); }
Second PVBuffer::frame_to_time( fFrame f ) const { return f / ( float( get_sample_rate() ) / float( get_hop_size() ) ); }
fBin PVBuffer::frequency_to_bin( Frequency f ) const { return f / ( float( get_sample_rate() ) / float( get_dft_size() ) ); }
Frequency PVBuffer::bin_to_frequency( fBin b ) const { return b * float( get_sample_rate() ) / float( get_dft_size() ); }   // PVBuffer.cpp:443-446
Frequency PVBuffer::get_frequency_offset( Channel c, Frame f, Bin b ) const { return get_MF( c, f, b ).f - bin_to_frequency( b ); }
Channel PVBuffer::bound_channel( Channel c ) const { return std::clamp( c, 0, get_num_channels() - 1 ); }
Frame PVBuffer::bound_frame( Frame f ) const { return std::clamp( f, 0, get_num_frames() - 1 ); }
Bin PVBuffer::bound_bin( Bin b ) const { return std::clamp( b, 0, get_num_bins() - 1 ); }

void PVBuffer::set_MF( Channel c, Frame f, Bin b, MF v ) { buffer.host_mut()[get_buffer_pos( c, f, b )] = v; }
MF & PVBuffer::get_MF( Channel c, Frame f, Bin b ) { return buffer.host_mut()[get_buffer_pos( c, f, b )]; }

void PVBuffer::clear_buffer()
	{
	std::vector<MF> & h = buffer.host_mut();
	std::fill( h.begin(), h.end(), MF{ 0, 0 } );
	}

MF * PVBuffer::get_MF_pointer( Channel c, Frame f, Bin b ) { return buffer.host_mut().data() + get_buffer_pos( c, f, b ); }
const MF * PVBuffer::get_MF_pointer( Channel c, Frame f, Bin b ) const { return buffer.host().data() + get_buffer_pos( c, f, b ); }
std::vector<MF> & PVBuffer::get_buffer() { return buffer.host_mut(); }
const std::vector<MF> & PVBuffer::get_buffer() const { return buffer.host(); }
std::vector<MF>::iterator PVBuffer::channel_begin( Channel c ) { return buffer.host_mut().begin() + get_buffer_pos( c, 0, 0 ); }
std::vector<MF>::iterator PVBuffer::channel_end( Channel c ) { return buffer.host_mut().begin() + get_buffer_pos( c + 1, 0, 0 ); }
std::vector<MF>::const_iterator PVBuffer::channel_begin( Channel c ) const { return buffer.host().begin() + get_buffer_pos( c, 0, 0 ); }
std::vector<MF>::const_iterator PVBuffer::channel_end( Channel c ) const { return buffer.host().begin() + get_buffer_pos( c + 1, 0, 0 ); }

size_t PVBuffer::get_buffer_pos( Channel c, Frame f, Bin b ) const
	{
	return ( size_t( c ) * size_t( get_num_frames() ) + size_t( f ) ) * size_t( get_num_bins() ) + size_t( b );
	}

std::ostream & operator<<( std::ostream & os, const PVBuffer & p )
	{
	os << "\n=========================== PVBuffer Info ==========================="
	   << "\nChannels:\t" << p.get_num_channels()
	   << "\nFrames:\t\t" << p.get_num_frames()
	   << "\nBins:\t\t" << p.get_num_bins()
	   << "\nFrames/second:\t" << p.time_to_frame( 1 )
	   << "\nBins/Hz:\t" << p.frequency_to_bin( 1 )
	   << "\nHop size:\t" << p.get_hop_size()
	   << "\nDFT size:\t" << p.get_dft_size()
	   << "\n======================================================================\n\n";
	return os;
	}


// ---- .flan files (reference PVBuffer.cpp:99-140, 216-273): header on the host, samples through the GPU codec --------
bool PVBuffer::save( const std::string & filename ) const
	{
	flan_b200_ctx * ctx = b200::context();
	if( !ctx ) return false;
	const MF * d = buffer.empty() ? nullptr : buffer.device();
	if( !buffer.empty() && !d ) return false;
	const int rc = flan_b200_save_flan( ctx, filename.c_str(), reinterpret_cast<const float *>( d ), get_num_channels(),
		get_num_frames(), get_num_bins(), get_sample_rate(), get_analysis_rate(), get_window_size() );
	if( rc != FLAN_B200_OK ) std::cout << flan_b200_last_error( ctx ) << std::endl;
	return rc == FLAN_B200_OK;
	}

bool PVBuffer::load( const std::string & filename )
	{
	flan_b200_ctx * ctx = b200::context();
	if( !ctx ) return false;
	int C = 0, B = 0, W = 0; int64_t F = 0; float sr = 0, rate = 0;
	if( flan_b200_flan_info( ctx, filename.c_str(), &C, &F, &B, &sr, &rate, &W ) != FLAN_B200_OK )
		{
		std::cout << flan_b200_last_error( ctx ) << std::endl;
		return false;
		}
	Format f;
	f.num_channels = C; f.num_frames = Frame( F ); f.num_bins = B;
	f.sample_rate = sr; f.analysis_rate = rate; f.window_size = W;       // PVBuffer.cpp:241-246
	MF * d = nullptr;
	b200::Mirror<MF> data = b200::Mirror<MF>::device_result( size_t( C ) * size_t( F ) * size_t( B ), &d );
	if( !d ) return false;
	if( flan_b200_load_flan( ctx, filename.c_str(), reinterpret_cast<float *>( d ), int64_t( C ) * F * B ) != FLAN_B200_OK )
		{
		std::cout << flan_b200_last_error( ctx ) << std::endl;
		return false;
		}
	*this = from_device_result( f, std::move( data ) );
	return true;
	}

}
