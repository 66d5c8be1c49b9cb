// mb_kernels.h -- host-callable launchers of the frame-path kernels.
#pragma once
#include <cuda_runtime.h>

#include "mb_device.cuh"

// Block-per-frame kernel, any power-of-two bufferSize in [16, 32768].
size_t mb_generic_smem_bytes(int M);
cudaError_t mb_launch_generic(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                              int num_sms, cudaStream_t stream);
