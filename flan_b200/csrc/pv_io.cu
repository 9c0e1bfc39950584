// flan_b200/csrc/pv_io.cu -- sm_100a kernels for the 24-bit sample codecs of the file formats either side of the path:
//   * .flan RIFF-PV: magnitude / dft size and frequency / sample rate, clamped to [-1,1], times 2^23, truncated, three
//     little-endian bytes each (PVBuffer::save, reference PV/PVBuffer.cpp:99-127; load :253-268);
//   * WAV PCM-24 as the reference writes it through libsndfile (Audio/AudioBuffer.cpp:136-170): clamp, interleave,
//     lrintf( x * 0x7FFFFF ); and reads it (:112-125): value / 2^23, de-interleave.
// Byte work, HBM-bound: a thread converts four values into three 32-bit words (or back); words pass through a
// shared-memory tile so that global traffic on the byte side is full 16-byte vectors.
#include "pv_io.h"

namespace pvio {

constexpr int THREADS = 256;
constexpr int TILE_VALUES = THREADS * 4;            // 24-bit values per tile
constexpr int TILE_WORDS = THREADS * 3;             // = 3072 bytes

__device__ __forceinline__ int flan_quantise( float v, float div )
	{
	float x = v / div;                                                  // PVBuffer.cpp:112-113
	x = x < -1.0f ? -1.0f : ( 1.0f < x ? 1.0f : x );                    // std::clamp (NaN passes through)
	const double scaled = (double) x * 8388608.0;
	return ( scaled != scaled ) ? (int) 0x80000000u : (int) scaled;     // truncation; NaN as the x86 conversion gives it
	}

__device__ __forceinline__ int pcm24_quantise( float s )
	{
	s = s < -1.0f ? -1.0f : ( 1.0f < s ? 1.0f : s );                    // AudioBuffer.cpp:158-161
	return __float2int_rn( s * 8388607.0f );                            // lrintf (round to nearest even), libsndfile pcm.c
	}

__device__ __forceinline__ void pack4( const int q[4], unsigned w[3] )
	{
	const unsigned a = q[0] & 0xFFFFFF, b = q[1] & 0xFFFFFF, c = q[2] & 0xFFFFFF, d = q[3] & 0xFFFFFF;
	w[0] = a | ( b << 24 );
	w[1] = ( b >> 8 ) | ( c << 16 );
	w[2] = ( c >> 16 ) | ( d << 8 );
	}

__device__ __forceinline__ void unpack4( const unsigned w[3], int q[4] )
	{
	const unsigned a = w[0] & 0xFFFFFF, b = ( w[0] >> 24 ) | ( ( w[1] & 0xFFFF ) << 8 ),
	               c = ( w[1] >> 16 ) | ( ( w[2] & 0xFF ) << 16 ), d = w[2] >> 8;
	const unsigned v[4] = { a, b, c, d };
#pragma unroll
	for( int i = 0; i < 4; ++i ) q[i] = (int)( v[i] << 8 ) >> 8;         // sign-extend 24 -> 32 (PVBuffer.cpp:261)
	}

// Tile of words -> global bytes. Full tiles go out as uint4; the last, partial tile byte by byte.
__device__ __forceinline__ void store_tile( const unsigned * tile, uint8_t * bytes, int64_t tile_index, int64_t total_bytes )
	{
	const int64_t base = tile_index * ( TILE_WORDS * 4 );
	if( base + TILE_WORDS * 4 <= total_bytes )
		{
		if( threadIdx.x < TILE_WORDS / 4 )
			__stcs( reinterpret_cast<uint4 *>( bytes + base ) + threadIdx.x, reinterpret_cast<const uint4 *>( tile )[threadIdx.x] );
		}
	else
		for( int64_t b = threadIdx.x; base + b < total_bytes; b += THREADS )
			bytes[base + b] = (uint8_t)( tile[b >> 2] >> ( 8 * ( b & 3 ) ) );
	}

__device__ __forceinline__ void load_tile( unsigned * tile, const uint8_t * bytes, int64_t tile_index, int64_t total_bytes )
	{
	const int64_t base = tile_index * ( TILE_WORDS * 4 );
	if( base + TILE_WORDS * 4 <= total_bytes )
		{
		if( threadIdx.x < TILE_WORDS / 4 )
			reinterpret_cast<uint4 *>( tile )[threadIdx.x] = __ldcs( reinterpret_cast<const uint4 *>( bytes + base ) + threadIdx.x );
		}
	else
		{
		for( int w = threadIdx.x; w < TILE_WORDS; w += THREADS ) tile[w] = 0;
		__syncthreads();
		for( int64_t b = threadIdx.x; base + b < total_bytes; b += THREADS )
			atomicOr( &tile[b >> 2], (unsigned) bytes[base + b] << ( 8 * ( b & 3 ) ) );
		}
	}

// values = 2 * count floats (m, f, m, f, ...); even positions are scaled by the dft size, odd ones by the sample rate.
__global__ void __launch_bounds__( THREADS ) pv_flan_encode_kernel( const float * pv, int64_t values, float dft_size, float sample_rate, uint8_t * bytes )
	{
	__shared__ __align__( 16 ) unsigned tile[TILE_WORDS];
	const int64_t tiles = ( values + TILE_VALUES - 1 ) / TILE_VALUES;
	for( int64_t t = blockIdx.x; t < tiles; t += gridDim.x )
		{
		const int64_t v0 = t * TILE_VALUES + 4 * threadIdx.x;
		int q[4] = { 0, 0, 0, 0 };
		if( v0 + 4 <= values )
			{
			const float4 x = __ldcs( reinterpret_cast<const float4 *>( pv + v0 ) );
			q[0] = flan_quantise( x.x, dft_size ); q[1] = flan_quantise( x.y, sample_rate );
			q[2] = flan_quantise( x.z, dft_size ); q[3] = flan_quantise( x.w, sample_rate );
			}
		else
			for( int i = 0; i < 4; ++i ) if( v0 + i < values ) q[i] = flan_quantise( pv[v0 + i], ( i & 1 ) ? sample_rate : dft_size );
		unsigned w[3];
		pack4( q, w );
		tile[3 * threadIdx.x] = w[0]; tile[3 * threadIdx.x + 1] = w[1]; tile[3 * threadIdx.x + 2] = w[2];
		__syncthreads();
		store_tile( tile, bytes, t, values * 3 );
		__syncthreads();
		}
	}

__global__ void __launch_bounds__( THREADS ) pv_flan_decode_kernel( const uint8_t * bytes, int64_t values, float dft_size, float sample_rate, float * pv )
	{
	__shared__ __align__( 16 ) unsigned tile[TILE_WORDS];
	const int64_t tiles = ( values + TILE_VALUES - 1 ) / TILE_VALUES;
	for( int64_t t = blockIdx.x; t < tiles; t += gridDim.x )
		{
		load_tile( tile, bytes, t, values * 3 );
		__syncthreads();
		const unsigned w[3] = { tile[3 * threadIdx.x], tile[3 * threadIdx.x + 1], tile[3 * threadIdx.x + 2] };
		int q[4];
		unpack4( w, q );
		const int64_t v0 = t * TILE_VALUES + 4 * threadIdx.x;
		float x[4];
#pragma unroll
		for( int i = 0; i < 4; ++i ) x[i] = (float)( (double) q[i] / 8388608.0 ) * ( ( i & 1 ) ? sample_rate : dft_size );   // PVBuffer.cpp:262
		if( v0 + 4 <= values ) __stcs( reinterpret_cast<float4 *>( pv + v0 ), make_float4( x[0], x[1], x[2], x[3] ) );
		else for( int i = 0; i < 4; ++i ) if( v0 + i < values ) pv[v0 + i] = x[i];
		__syncthreads();
		}
	}

// Interleaved value j = frame * C + channel reads planar audio[channel * stride + frame] (n frames of a longer signal).
__global__ void __launch_bounds__( THREADS ) pv_pcm24_encode_kernel( const float * audio, int C, int64_t stride, int64_t n, uint8_t * bytes )
	{
	__shared__ __align__( 16 ) unsigned tile[TILE_WORDS];
	const int64_t values = n * C;
	const int64_t tiles = ( values + TILE_VALUES - 1 ) / TILE_VALUES;
	for( int64_t t = blockIdx.x; t < tiles; t += gridDim.x )
		{
		const int64_t v0 = t * TILE_VALUES + 4 * threadIdx.x;
		int q[4] = { 0, 0, 0, 0 };
#pragma unroll
		for( int i = 0; i < 4; ++i )
			if( v0 + i < values )
				{
				const int64_t j = v0 + i;
				q[i] = pcm24_quantise( __ldcs( audio + ( j % C ) * stride + j / C ) );
				}
		unsigned w[3];
		pack4( q, w );
		tile[3 * threadIdx.x] = w[0]; tile[3 * threadIdx.x + 1] = w[1]; tile[3 * threadIdx.x + 2] = w[2];
		__syncthreads();
		store_tile( tile, bytes, t, values * 3 );
		__syncthreads();
		}
	}

__global__ void __launch_bounds__( THREADS ) pv_pcm24_decode_kernel( const uint8_t * bytes, int C, int64_t stride, int64_t n, float * audio )
	{
	__shared__ __align__( 16 ) unsigned tile[TILE_WORDS];
	const int64_t values = n * C;
	const int64_t tiles = ( values + TILE_VALUES - 1 ) / TILE_VALUES;
	for( int64_t t = blockIdx.x; t < tiles; t += gridDim.x )
		{
		load_tile( tile, bytes, t, values * 3 );
		__syncthreads();
		const unsigned w[3] = { tile[3 * threadIdx.x], tile[3 * threadIdx.x + 1], tile[3 * threadIdx.x + 2] };
		int q[4];
		unpack4( w, q );
		const int64_t v0 = t * TILE_VALUES + 4 * threadIdx.x;
#pragma unroll
		for( int i = 0; i < 4; ++i )
			if( v0 + i < values )
				{
				const int64_t j = v0 + i;
				// libsndfile pcm.c let2f_array: ( sample << 8 ) * ( 1 / 0x80000000 ) = sample / 2^23, exact in float
				audio[( j % C ) * stride + j / C] = (float)( q[i] * 256 ) * ( 1.0f / 2147483648.0f );
				}
		__syncthreads();
		}
	}

static unsigned grid_for( int64_t values, int sms )
	{
	int64_t tiles = ( values + TILE_VALUES - 1 ) / TILE_VALUES;
	if( tiles > (int64_t) sms * 16 ) tiles = (int64_t) sms * 16;
	return (unsigned)( tiles < 1 ? 1 : tiles );
	}

cudaError_t launch_flan_encode( const float * pv, int64_t count, float dft_size, float sample_rate, uint8_t * bytes, int sms, cudaStream_t st )
	{
	pv_flan_encode_kernel<<<grid_for( 2 * count, sms ), THREADS, 0, st>>>( pv, 2 * count, dft_size, sample_rate, bytes );
	return cudaGetLastError();
	}
cudaError_t launch_flan_decode( const uint8_t * bytes, int64_t count, float dft_size, float sample_rate, float * pv, int sms, cudaStream_t st )
	{
	pv_flan_decode_kernel<<<grid_for( 2 * count, sms ), THREADS, 0, st>>>( bytes, 2 * count, dft_size, sample_rate, pv );
	return cudaGetLastError();
	}
cudaError_t launch_pcm24_encode( const float * audio, int C, int64_t stride, int64_t n, uint8_t * bytes, int sms, cudaStream_t st )
	{
	pv_pcm24_encode_kernel<<<grid_for( n * C, sms ), THREADS, 0, st>>>( audio, C, stride, n, bytes );
	return cudaGetLastError();
	}
cudaError_t launch_pcm24_decode( const uint8_t * bytes, int C, int64_t stride, int64_t n, float * audio, int sms, cudaStream_t st )
	{
	pv_pcm24_decode_kernel<<<grid_for( n * C, sms ), THREADS, 0, st>>>( bytes, C, stride, n, audio );
	return cudaGetLastError();
	}

} // namespace pvio
