// flan_b200/csrc/pv_modify.h -- host-visible launch interface of pv_modify.cu (PV-domain kernels).
#pragma once

#include <cuda_runtime.h>
#include "pv_modify_body.cuh"

namespace pvm {

// Result of the time-map reduction, device memory: key = float_key( maximum ), descends != 0 when some column descends.
struct MapCheck { unsigned int max_key; int descends; };

cudaError_t launch_bin_prefix( const Table & factor, int64_t rows, int B, float sample_rate, float dft, float * out, cudaStream_t st );
cudaError_t launch_frame_prefix( const Table & factor, int64_t F, int cols, float rate, float * out, MapCheck * check, cudaStream_t st );
cudaError_t launch_map_check( const Table & mod, int64_t F, int cols, MapCheck * check, int sms, cudaStream_t st );
cudaError_t launch_repitch( const RepitchArgs & a, int64_t rows, cudaStream_t st );
cudaError_t launch_stretch_parallel( const StretchArgs & a, int C, cudaStream_t st );
cudaError_t launch_stretch_sequential( const StretchArgs & a, int C, cudaStream_t st );

} // namespace pvm
