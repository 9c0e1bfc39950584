// tools/cpp/e2e_bench.cpp -- the end-to-end leg of bench.py, written as a user of the reference's C++ API would write it
// (the shape of the reference's tests/flanTest.cpp:39-44), compiled against the B200 build's headers:
//     Audio (host std::vector) -> convert_to_PV -> convert_to_audio -> get_buffer() (host std::vector)
// Every step moves the step's input from its host vector to the device and the result back into a host vector; nothing
// here touches CUDA or the C ABI directly. `threads` host threads run independent round trips on their own objects (a
// program converting a batch of files), which is what lets one thread's download overlap another's upload.
#include "flan/Audio/Audio.h"
#include "flan/PV/PV.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

using namespace flan;

extern "C" {

// Returns seconds for `steps` round trips per thread (after `warmup` untimed ones), or a negative value on failure.
// checksum_out receives one sample of every result so that no copy can be skipped.
double e2e_round_trips( const float * audio, int C, long long n, float sr, int W, int hop, int N,
                        int threads, int warmup, int steps, double * checksum_out )
	{
	std::vector<Audio> in;
	for( int t = 0; t < threads; ++t )
		in.push_back( Audio::create_from_buffer( std::vector<float>( audio, audio + size_t( C ) * n ), C, sr ) );
	std::atomic<int> failed( 0 );
	std::vector<double> sums( threads, 0.0 );
	auto pass = [&]( int t )
		{
		// the step's input lives in the Audio's host vector: touching it through the mutable accessor (as any host-side
		// edit does) makes the host copy the newest one, so the conversion uploads it again
		std::vector<float> & h = in[t].get_buffer();
		h[0] = audio[0];
		PV pv = in[t].convert_to_PV( W, hop, N );
		if( pv.is_null() ) { failed = 1; return; }
		Audio out = pv.convert_to_audio();
		if( out.is_null() ) { failed = 1; return; }
		const std::vector<float> & y = out.get_buffer();
		sums[t] += y[y.size() / 2] + y[y.size() - 1];
		};
	auto run = [&]( int count )
		{
		std::vector<std::thread> pool;
		for( int t = 0; t < threads; ++t ) pool.emplace_back( [&, t] { for( int i = 0; i < count && !failed; ++i ) pass( t ); } );
		for( auto & th : pool ) th.join();
		};
	run( warmup );
	const auto t0 = std::chrono::steady_clock::now();
	run( steps );
	const double dt = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
	if( failed ) return -1.0;
	double s = 0.0; for( double v : sums ) s += v;
	if( checksum_out ) *checksum_out = s;
	return dt;
	}

}
