// tools/micro/hostcopy.cu -- what does a host std::vector cost to move? (design input for the host-buffer path)
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o hostcopy.bin hostcopy.cu -lpthread
// Prints GB/s (or ms) for: pinned H2D / D2H / both at once, pageable H2D / D2H through the driver, cudaHostRegister +
// unregister of faulted pageable memory, memcpy into pinned memory with 1..16 threads, first-touch of a fresh vector.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>( std::chrono::steady_clock::now().time_since_epoch() ).count(); }
#define CK( x ) do { cudaError_t e = ( x ); if( e != cudaSuccess ) { std::printf( "%s: %s\n", #x, cudaGetErrorString( e ) ); return 1; } } while( 0 )

static void par_copy( char * dst, const char * src, size_t bytes, int threads )
	{
	std::vector<std::thread> th;
	const size_t per = ( bytes / threads + 4095 ) / 4096 * 4096;
	for( int i = 0; i < threads; ++i )
		{
		const size_t lo = per * i, hi = lo + per < bytes ? lo + per : bytes;
		if( lo < hi ) th.emplace_back( [=] { std::memcpy( dst + lo, src + lo, hi - lo ); } );
		}
	for( auto & t : th ) t.join();
	}

int main()
	{
	const size_t bytes = size_t( 230 ) << 20;
	std::printf( "host threads %u, buffer %zu MiB\n", std::thread::hardware_concurrency(), bytes >> 20 );
	void * d0, * d1, * p0, * p1;
	CK( cudaMalloc( &d0, bytes ) ); CK( cudaMalloc( &d1, bytes ) );
	CK( cudaMallocHost( &p0, bytes ) ); CK( cudaMallocHost( &p1, bytes ) );
	std::memset( p0, 1, bytes ); std::memset( p1, 2, bytes );
	cudaStream_t s0, s1; CK( cudaStreamCreate( &s0 ) ); CK( cudaStreamCreate( &s1 ) );
	for( int rep = 0; rep < 2; ++rep )
		{
		double t = now(); CK( cudaMemcpyAsync( d0, p0, bytes, cudaMemcpyHostToDevice, s0 ) ); CK( cudaStreamSynchronize( s0 ) );
		double a = now() - t;
		t = now(); CK( cudaMemcpyAsync( p1, d1, bytes, cudaMemcpyDeviceToHost, s1 ) ); CK( cudaStreamSynchronize( s1 ) );
		double b = now() - t;
		t = now();
		CK( cudaMemcpyAsync( d0, p0, bytes, cudaMemcpyHostToDevice, s0 ) ); CK( cudaMemcpyAsync( p1, d1, bytes, cudaMemcpyDeviceToHost, s1 ) );
		CK( cudaStreamSynchronize( s0 ) ); CK( cudaStreamSynchronize( s1 ) );
		double c = now() - t;
		std::printf( "pinned: H2D %.1f GB/s (%.2f ms)  D2H %.1f GB/s (%.2f ms)  both at once %.2f ms\n", bytes / a / 1e9, a * 1e3, bytes / b / 1e9, b * 1e3, c * 1e3 );
		}
	// pageable through the driver
		{
		std::vector<char> v( bytes, 3 );
		for( int rep = 0; rep < 2; ++rep )
			{
			double t = now(); CK( cudaMemcpy( d0, v.data(), bytes, cudaMemcpyHostToDevice ) ); double a = now() - t;
			t = now(); CK( cudaMemcpy( v.data(), d1, bytes, cudaMemcpyDeviceToHost ) ); double b = now() - t;
			std::printf( "pageable (driver staging): H2D %.1f GB/s (%.2f ms)  D2H %.1f GB/s (%.2f ms)\n", bytes / a / 1e9, a * 1e3, bytes / b / 1e9, b * 1e3 );
			}
		for( int rep = 0; rep < 2; ++rep )
			{
			double t = now(); CK( cudaHostRegister( v.data(), bytes, cudaHostRegisterDefault ) ); double a = now() - t;
			t = now(); CK( cudaMemcpyAsync( d0, v.data(), bytes, cudaMemcpyHostToDevice, s0 ) ); CK( cudaStreamSynchronize( s0 ) ); double c = now() - t;
			t = now(); CK( cudaHostUnregister( v.data() ) ); double b = now() - t;
			std::printf( "cudaHostRegister %.2f ms, H2D from it %.1f GB/s, unregister %.2f ms\n", a * 1e3, bytes / c / 1e9, b * 1e3 );
			}
		for( int th : { 1, 2, 4, 8, 16 } )
			{
			double best = 1e9;
			for( int rep = 0; rep < 3; ++rep ) { double t = now(); par_copy( (char *) p0, v.data(), bytes, th ); double a = now() - t; if( a < best ) best = a; }
			double best2 = 1e9;
			for( int rep = 0; rep < 3; ++rep ) { double t = now(); par_copy( v.data(), (const char *) p1, bytes, th ); double a = now() - t; if( a < best2 ) best2 = a; }
			std::printf( "memcpy %2d threads: pageable->pinned %.1f GB/s, pinned->pageable %.1f GB/s\n", th, bytes / best / 1e9, bytes / best2 / 1e9 );
			}
		}
	// fresh allocation + first touch (what a new result std::vector<float>( n ) costs)
	for( int rep = 0; rep < 2; ++rep )
		{
		double t = now(); std::vector<float> v( bytes / 4 ); double a = now() - t;
		std::printf( "std::vector<float>( %zu M ) value-initialised: %.2f ms (%.1f GB/s)\n", bytes >> 22, a * 1e3, bytes / a / 1e9 );
		t = now(); par_copy( (char *) v.data(), (const char *) p1, bytes, 8 ); a = now() - t;
		std::printf( "  then 8-thread copy into it %.1f GB/s\n", bytes / a / 1e9 );
		}
	// pipelined staged upload: 8 MiB chunks through a 4-deep pinned ring, 4 copy threads
		{
		std::vector<char> v( bytes, 5 );
		const size_t chunk = size_t( 8 ) << 20; const int depth = 4;
		cudaEvent_t ev[depth]; for( auto & evk : ev ) CK( cudaEventCreateWithFlags( &evk, cudaEventDisableTiming ) );
		for( int th : { 1, 2, 4, 8 } )
			{
			double t = now();
			int k = 0;
			for( size_t off = 0; off < bytes; off += chunk, ++k )
				{
				const size_t nb = off + chunk < bytes ? chunk : bytes - off;
				char * stage = (char *) p0 + size_t( k % depth ) * chunk;
				if( k >= depth ) CK( cudaEventSynchronize( ev[k % depth] ) );
				par_copy( stage, v.data() + off, nb, th );
				CK( cudaMemcpyAsync( (char *) d0 + off, stage, nb, cudaMemcpyHostToDevice, s0 ) );
				CK( cudaEventRecord( ev[k % depth], s0 ) );
				}
			CK( cudaStreamSynchronize( s0 ) );
			double a = now() - t;
			std::printf( "staged H2D, 8 MiB chunks, %d copy threads: %.1f GB/s (%.2f ms)\n", th, bytes / a / 1e9, a * 1e3 );
			}
		}
	return 0;
	}
