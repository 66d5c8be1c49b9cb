// mb_kernels.h -- host-callable launchers of the frame-path kernels.
#pragma once
#include <cuda_runtime.h>

#include "mb_device.cuh"

// Block-per-frame kernel, any power-of-two bufferSize in [16, 32768].
size_t mb_generic_smem_bytes(int M, bool exact);
// Warp-per-frame kernel, bufferSize 2048, float32 FFT.  `buffer` output rows must be 16-byte aligned; frames
// that are not 16-byte aligned (odd clip offsets / hops) are loaded by the lanes instead of TMA.
size_t mb_warp2048_smem_bytes();
cudaError_t mb_launch_warp2048(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                               int num_sms, cudaStream_t stream);

// Multi-frame warp kernel, bufferSize 256, 512 and 1024 (F = 2048 / N frames per warp at a time), float32 FFT.
size_t mb_warpmf_smem_bytes();
cudaError_t mb_launch_warpmf(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                             int num_sms, cudaStream_t stream);

// bufferSize 32768, float32 FFT: 16 warps per frame (16 x 1024-point register sub-FFTs + radix-16 combine).
// Needs 16-byte aligned frames like the warp kernel.
size_t mb_big32768_smem_bytes();
size_t mb_big_smem_bytes(int N);  // bufferSize 4096 / 8192 / 16384 / 32768
cudaError_t mb_launch_big32768(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                               int num_sms, cudaStream_t stream);

// Exact-FFT mode on a 2-CTA cluster (DSMEM exchange in the last radix-2 stage): bufferSize up to 32768.
size_t mb_exact_cluster_smem_bytes(int N);
cudaError_t mb_launch_exact_cluster(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                                    int num_sms, cudaStream_t stream);

// Exact-FFT mode, one warp per frame (bufferSize 512 / 1024 / 2048): float64 register passes with the float32 stage
// stores emulated in registers.  tw_small: the 15 recurrence twiddles of widths 1, 2, 4, 8 as (re, im) pairs.
size_t mb_exact_warp_smem_bytes(int N);
bool mb_exact_warp_supports(int N);
cudaError_t mb_launch_exact_warp(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                                 int num_sms, cudaStream_t stream, const double *tw_small);

cudaError_t mb_launch_generic(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                              int num_sms, cudaStream_t stream);
