// examples/pv_chain.cpp -- user code against the reference's C++ API (the shape of the reference's own scratch program,
// tests/flanTest.cpp:32-47: load a WAV, convert_to_PV, a PV-domain edit, convert_to_audio, save), linked against the
// B200 build instead of Flan + FFTW + libsndfile:
//
//   g++ -std=c++20 -O2 examples/pv_chain.cpp -I flan_b200/host/include -I include -L flan_b200/lib
//       -Wl,-rpath,$PWD/flan_b200/lib -lflan_b200_host -lflan_b200 -o pv_chain
//   ./pv_chain in.wav out.wav [pitch_factor=1.5] [stretch_factor=2.0]
//   ./pv_chain --synthetic 60 out.wav            (a 60 s stereo test signal instead of a file)
//
// Every buffer between load and save lives in HBM; only the file bytes cross PCIe.
#include "flan/Audio/Audio.h"
#include "flan/PV/PV.h"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

using namespace flan;

static double now()
	{
	return std::chrono::duration<double>( std::chrono::steady_clock::now().time_since_epoch() ).count();
	}

int main( int argc, char ** argv )
	{
	if( argc < 3 )
		{
		std::cout << "usage: pv_chain in.wav out.wav [pitch=1.5] [stretch=2.0]   |   pv_chain --synthetic seconds out.wav\n";
		return 2;
		}
	const bool synthetic = std::strcmp( argv[1], "--synthetic" ) == 0;
	const std::string out_path = synthetic ? argv[3] : argv[2];
	const int opt = synthetic ? 4 : 3;
	const float pitch = argc > opt ? float( std::atof( argv[opt] ) ) : 1.5f;
	const float stretch = argc > opt + 1 ? float( std::atof( argv[opt + 1] ) ) : 2.0f;

	Audio in;
	if( synthetic )
		{
		const float sr = 48000.0f;
		const size_t n = size_t( sr * std::atof( argv[2] ) );
		std::vector<float> s( 2 * n );
		for( size_t i = 0; i < n; ++i )
			{
			const double t = double( i ) / sr;
			s[i] = float( 0.4 * std::sin( 2.0 * M_PI * ( 110.0 * t + 40.0 * t * t ) ) );
			s[n + i] = float( 0.4 * std::sin( 2.0 * M_PI * 220.0 * t ) * std::exp( -0.2 * std::fmod( t, 2.0 ) ) );
			}
		in = Audio::create_from_buffer( std::move( s ), 2, sr );
		}
	else if( !in.load( argv[1] ) ) return 1;

	const double t0 = now();
	PV pv = in.convert_to_PV( 2048, 128, 2048 );
	PV shaped = pv.repitch( pitch ).stretch( stretch );
	Audio out = shaped.convert_to_audio();
	if( out.is_null() ) return 1;
	const float probe = out.get_sample( 0, out.get_num_frames() / 2 );      // forces completion (lazy download)
	const double t1 = now();
	if( !out.save( out_path ) ) return 1;

	std::cout << "in: " << in.get_num_channels() << " ch x " << in.get_num_frames() << " samples @ " << in.get_sample_rate() << " Hz\n"
	          << "pv: " << pv.get_num_frames() << " frames x " << pv.get_num_bins() << " bins -> " << shaped.get_num_frames() << " frames\n"
	          << "out: " << out.get_num_frames() << " samples, mid sample " << probe << "\n"
	          << "convert_to_PV + repitch + stretch + convert_to_audio (incl. upload / first-use setup): " << ( t1 - t0 ) * 1e3 << " ms\n";
	return 0;
	}
