// Device mirror + process-wide engine context of the B200 build (see flan/b200_storage.h).
#include "flan/b200_storage.h"
#include "flan/defines.h"

#include <cstdlib>
#include <iostream>
#include <mutex>

#include "flan_b200.h"

namespace flan::b200 {

flan_b200_ctx * context()
	{
	static std::once_flag once;
	static flan_b200_ctx * ctx = nullptr;
	std::call_once( once, []
		{
		int device = 0;
		if( const char * e = std::getenv( "FLAN_B200_DEVICE" ) ) device = std::atoi( e );
		if( flan_b200_create( device, &ctx ) != FLAN_B200_OK )
			{
			std::cout << "flan_b200: cannot create the GPU engine: " << flan_b200_last_error( nullptr ) << std::endl;
			ctx = nullptr;
			}
		} );
	return ctx;
	}

template<typename T>
struct Mirror<T>::DeviceMem
	{
	void * ptr = nullptr;
	~DeviceMem() { if( ptr && context() ) flan_b200_free( context(), ptr ); }
	};

template<typename T>
const T * Mirror<T>::device() const
	{
	flan_b200_ctx * ctx = context();
	if( !ctx ) return nullptr;
	std::lock_guard<std::mutex> lock( lazy_ );
	if( device_valid_ && dev_ ) return static_cast<const T *>( dev_->ptr );
	if( !dev_ )
		{
		auto mem = std::make_shared<DeviceMem>();
		if( flan_b200_malloc( ctx, sizeof( T ) * count_, &mem->ptr ) != FLAN_B200_OK )
			{
			std::cout << "flan_b200: " << flan_b200_last_error( ctx ) << std::endl;
			return nullptr;
			}
		dev_ = mem;
		}
	if( count_ )
		{
		if( flan_b200_upload( ctx, dev_->ptr, host_.data(), sizeof( T ) * count_ ) != FLAN_B200_OK
		 || flan_b200_synchronize( ctx ) != FLAN_B200_OK )       // host_ is pageable and may change afterwards
			{
			std::cout << "flan_b200: " << flan_b200_last_error( ctx ) << std::endl;
			return nullptr;
			}
		}
	device_valid_ = true;
	return static_cast<const T *>( dev_->ptr );
	}

template<typename T>
Mirror<T> Mirror<T>::device_result( size_t count, T ** d_out )
	{
	Mirror<T> m;
	*d_out = nullptr;
	flan_b200_ctx * ctx = context();
	if( !ctx ) return m;
	auto mem = std::make_shared<DeviceMem>();
	if( flan_b200_malloc( ctx, sizeof( T ) * count, &mem->ptr ) != FLAN_B200_OK )
		{
		std::cout << "flan_b200: " << flan_b200_last_error( ctx ) << std::endl;
		return m;
		}
	m.dev_ = mem;
	m.count_ = count;
	m.host_valid_ = false;
	m.device_valid_ = true;
	*d_out = static_cast<T *>( mem->ptr );
	return m;
	}

template<typename T>
void Mirror<T>::sync_to_host() const
	{
	std::lock_guard<std::mutex> lock( lazy_ );
	if( host_valid_ ) return;
	host_.resize( count_ );
	flan_b200_ctx * ctx = context();
	if( ctx && dev_ && count_ )
		{
		if( flan_b200_download( ctx, host_.data(), dev_->ptr, sizeof( T ) * count_ ) != FLAN_B200_OK
		 || flan_b200_synchronize( ctx ) != FLAN_B200_OK )
			std::cout << "flan_b200: " << flan_b200_last_error( ctx ) << std::endl;
		}
	host_valid_.store( true, std::memory_order_release );
	}

template<typename T>
Mirror<T> Mirror<T>::deep_copy() const
	{
	Mirror<T> m;
	m.host_ = host();          // downloads if needed; plain host copy, device copy rebuilt on demand
	m.count_ = count_;
	return m;
	}

template class Mirror<float>;
template class Mirror<MF>;

}
