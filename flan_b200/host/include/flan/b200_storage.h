// Device-resident storage behind flan::AudioBuffer / flan::PVBuffer in the B200 build.
//
// The reference's buffers are std::vector<Sample> / std::vector<MF> with reference-returning accessors
// (AudioBuffer.h:190,212-213; PVBuffer.h:250,272-273). Here every buffer is a host std::vector plus a device
// allocation with validity flags: conversions consume and produce the device copy; any host accessor lazily
// downloads it; any mutable host accessor invalidates the device copy. A chain such as
//     audio.convert_to_PV().convert_to_audio()
// therefore never moves the PV data across PCIe.
#pragma once

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <mutex>
#include <vector>

struct flan_b200_ctx;

namespace flan::b200 {

// Process-wide engine context (device from $FLAN_B200_DEVICE, default 0). Returns nullptr and prints the reason
// when no CUDA device / library is usable: there is no CPU fallback.
flan_b200_ctx * context();

template<typename T>
class Mirror
	{
public:
	Mirror() = default;
	explicit Mirror( size_t count ) : host_( count ), count_( count ) {}
	Mirror( std::vector<T> && v ) : host_( std::move( v ) ), count_( host_.size() ) {}
	Mirror( Mirror && o ) noexcept { take( o ); }
	Mirror & operator=( Mirror && o ) noexcept { if( this != &o ) take( o ); return *this; }
	Mirror( const Mirror & ) = delete;
	Mirror & operator=( const Mirror & ) = delete;

	size_t size() const { return count_; }
	bool empty() const { return count_ == 0; }

	// Host view. The const form downloads if the device copy is newer; the mutable form also drops the device copy.
	// Concurrent const access from several threads is safe, as it is for the reference's plain vectors.
	const std::vector<T> & host() const { if( !host_valid_.load( std::memory_order_acquire ) ) sync_to_host(); return host_; }
	std::vector<T> & host_mut() { if( !host_valid_.load( std::memory_order_acquire ) ) sync_to_host(); device_valid_ = false; return host_; }

	// Device view (uploads if the host copy is newer). nullptr on failure.
	const T * device() const;
	// Fresh device allocation of `count` elements whose contents the caller is about to produce on the GPU.
	static Mirror device_result( size_t count, T ** d_out );

	Mirror deep_copy() const;

private:
	struct DeviceMem;
	void sync_to_host() const;

	mutable std::vector<T> host_;
	size_t count_ = 0;
	void take( Mirror & o )
		{
		host_ = std::move( o.host_ ); count_ = o.count_; dev_ = std::move( o.dev_ );
		host_valid_ = o.host_valid_.load(); device_valid_ = o.device_valid_.load();
		o.count_ = 0; o.host_valid_ = true; o.device_valid_ = false;
		}

	mutable std::atomic<bool> host_valid_{ true };
	mutable std::atomic<bool> device_valid_{ false };
	mutable std::shared_ptr<DeviceMem> dev_;
	mutable std::mutex lazy_;      // serialises the lazy upload / download
	};

}
