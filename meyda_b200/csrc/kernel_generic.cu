// kernel_generic.cu -- block-per-frame kernel for any power-of-two bufferSize
// in [16, 32768].  One CTA walks frames (grid-stride); the frame lives in
// shared memory, is transformed in place, and every requested Meyda feature is
// produced in the same pass.  Nothing but the requested features is written
// to HBM.  Two FFT modes:
//
//   fast (default)  the frame is packed as N/2 complex values, transformed by a
//                   float32 decimation-in-frequency FFT with radix-8 fused
//                   passes and split into the real spectrum.
//   exact           (MB_FLAG_EXACT_FFT) the reference's own arithmetic: N-point
//                   radix-2 decimation-in-time on a zero-imaginary array,
//                   float64 butterflies with the recurrence twiddles and a
//                   float32 rounding at every stage store
//                   (lib/jsfft/fft.js:123-171), so complexSpectrum /
//                   amplitudeSpectrum / powerSpectrum come out bit for bit.
//
// Reference path being replaced: src/meyda.js:69-91,104-114,158-168,
// lib/jsfft/fft.js:123-208, the extractor files under src/extractors/ (see
// mb_device.cuh for the per-formula citations).
#include <cooperative_groups.h>

#include "mb_adaptive.cuh"
#include "mb_device.cuh"
#include "mb_fft.cuh"
#include "mb_kernels.h"

namespace cg = cooperative_groups;

// The kernels are compiled for every CTA size from one warp to 1024 threads; mb_launch_generic picks by
// bufferSize (32 threads up to 512 ... 1024 threads at 32768, where a frame needs most of an SM's shared memory).
namespace g32 {
#define MB_GENERIC_THREADS 32
#include "kernel_generic_impl.cuh"
#undef MB_GENERIC_THREADS
}  // namespace g32
namespace g64 {
#define MB_GENERIC_THREADS 64
#include "kernel_generic_impl.cuh"
#undef MB_GENERIC_THREADS
}  // namespace g64
namespace g128 {
#define MB_GENERIC_THREADS 128
#include "kernel_generic_impl.cuh"
#undef MB_GENERIC_THREADS
}  // namespace g128
namespace g256 {
#define MB_GENERIC_THREADS 256
#include "kernel_generic_impl.cuh"
#undef MB_GENERIC_THREADS
}  // namespace g256
namespace g512 {  // hosts the bufferSize-32768 kernel (16 warps per frame)
#define MB_GENERIC_THREADS 512
#include "kernel_generic_impl.cuh"
#undef MB_GENERIC_THREADS
}  // namespace g512
namespace g1024 {
#define MB_GENERIC_THREADS 1024
#include "kernel_generic_impl.cuh"
#undef MB_GENERIC_THREADS
}  // namespace g1024


size_t mb_generic_smem_bytes(int M, bool exact) {
    if (exact) return (size_t)2 * (2 * M + (2 * M >> 5) + 1) * sizeof(float) + (size_t)M * sizeof(float);  // padded re, im; amp
    const int padded = M + (M >> 5) + (M >> 10) + 1;
    return (size_t)padded * sizeof(float2) + (size_t)M * sizeof(float);
}

#define MB_DEFINE_LAUNCH_GENERIC(NS, THREADS)                                                                        \
    template <bool EXACT>                                                                                            \
    static cudaError_t launch_generic_##NS(const MbDevPlan &P, const MbClipTable &T, const float *samples,          \
                                           const mb_outputs &O, int num_sms, cudaStream_t stream) {                  \
        const size_t smem = mb_generic_smem_bytes(P.M, EXACT);                                                       \
        cudaError_t e = cudaFuncSetAttribute(NS::mb_generic_kernel<EXACT>,                                           \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                \
        if (e != cudaSuccess) return e;                                                                              \
        int per_sm = 1;                                                                                              \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, NS::mb_generic_kernel<EXACT>, THREADS, smem);         \
        if (per_sm < 1) per_sm = 1;                                                                                  \
        int64_t grid = (int64_t)num_sms * per_sm;                                                                    \
        if (grid > T.total_frames) grid = T.total_frames;                                                            \
        if (grid < 1) return cudaSuccess;                                                                            \
        (void)cudaGetLastError(); /* drop any stale non-sticky error left by other users of the context */          \
        NS::mb_generic_kernel<EXACT><<<(unsigned)grid, THREADS, smem, stream>>>(P, T, samples, O);                   \
        return cudaGetLastError();                                                                                   \
    }
MB_DEFINE_LAUNCH_GENERIC(g32, 32)
MB_DEFINE_LAUNCH_GENERIC(g64, 64)
MB_DEFINE_LAUNCH_GENERIC(g512, 512)
MB_DEFINE_LAUNCH_GENERIC(g128, 128)
MB_DEFINE_LAUNCH_GENERIC(g256, 256)
MB_DEFINE_LAUNCH_GENERIC(g1024, 1024)

// bufferSize 4096 / 8192 / 16384 / 32768: R = bufferSize / 2048 warps per frame, 16 / R persistent CTAs per SM
size_t mb_big_smem_bytes(int N) {
    return N == 32768 ? sizeof(g512::BigSmem) : N == 16384 ? sizeof(g256::BigSmem) : N == 8192 ? sizeof(g128::BigSmem)
                                                                                               : sizeof(g64::BigSmem);
}
size_t mb_big32768_smem_bytes() { return mb_big_smem_bytes(32768); }

template <typename K>
static cudaError_t launch_big(K kernel, int threads, const MbDevPlan &P, const MbClipTable &T, const float *samples,
                              const mb_outputs &O, int num_sms, cudaStream_t stream) {
    const size_t smem = mb_big_smem_bytes(P.N);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)num_sms * per_sm;
    if (grid > T.total_frames) grid = T.total_frames;
    if (grid < 1) return cudaSuccess;
    (void)cudaGetLastError();
    kernel<<<(unsigned)grid, threads, smem, stream>>>(P, T, samples, O);
    return cudaGetLastError();
}

cudaError_t mb_launch_big32768(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                               int num_sms, cudaStream_t stream) {
    if (P.N == 32768) return launch_big(g512::mb_big32768_kernel, 512, P, T, samples, O, num_sms, stream);
    if (P.N == 16384) return launch_big(g256::mb_big32768_kernel, 256, P, T, samples, O, num_sms, stream);
    if (P.N == 8192) return launch_big(g128::mb_big32768_kernel, 128, P, T, samples, O, num_sms, stream);
    return launch_big(g64::mb_big32768_kernel, 64, P, T, samples, O, num_sms, stream);
}

size_t mb_exact_cluster_smem_bytes(int N) {
    const size_t H = (size_t)N / 2;
    return (2 * (H + (H >> 5) + 1) + H) * sizeof(float);  // padded re, im halves; gathered amplitudes
}

cudaError_t mb_launch_exact_cluster(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                                    int num_sms, cudaStream_t stream) {
    const size_t smem = mb_exact_cluster_smem_bytes(P.N);
    const bool big = P.N >= 8192;
    cudaError_t e = cudaFuncSetAttribute(big ? g1024::mb_exact_cluster_kernel : g256::mb_exact_cluster_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t clusters = num_sms / 2;
    if (clusters > T.total_frames) clusters = T.total_frames;
    if (clusters < 1) return cudaSuccess;
    (void)cudaGetLastError();
    if (big) g1024::mb_exact_cluster_kernel<<<(unsigned)(2 * clusters), 1024, smem, stream>>>(P, T, samples, O);
    else g256::mb_exact_cluster_kernel<<<(unsigned)(2 * clusters), 256, smem, stream>>>(P, T, samples, O);
    return cudaGetLastError();
}

cudaError_t mb_launch_generic(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                              int num_sms, cudaStream_t stream) {
    // CTA size by bufferSize, from a sweep on the B200 (round 1, profiles/README.md): one warp per frame up to 512
    // (no block-wide barriers left, every reduction is a shuffle), then just enough threads to keep the N/2
    // complex points busy; large frames are shared-memory bound and want few, big CTAs.
#define MB_GO(NS)                                                                           \
    return P.exact ? launch_generic_##NS<true>(P, T, samples, O, num_sms, stream)           \
                   : launch_generic_##NS<false>(P, T, samples, O, num_sms, stream)
    // (the choice depends on bufferSize alone: a frame's result must not depend on how many frames share its
    // launch, so that the streaming path -- one buffer per push -- stays bit-identical to the batch call)
    if (P.N <= 512) MB_GO(g32);
    if (P.N <= 2048) MB_GO(g64);
    if (P.N == 4096) MB_GO(g128);
    if (P.N == 8192) MB_GO(g256);
    if (P.N == 16384) MB_GO(g512);
    MB_GO(g1024);
#undef MB_GO
}
