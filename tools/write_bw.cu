// write_bw.cu -- what HBM sustains for the frame path's traffic mix: a kernel that
// reads 2 KB and writes 33 KB per "frame" (128-byte rows per warp store, the same
// pattern as the feature kernel), plus a pure copy and a pure write for reference.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void mix_kernel(const float4 *in, float4 *out, long frames, int out_f4, int in_f4) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * (long)blockDim.x) >> 5;
    for (long f = warp; f < frames; f += nwarps) {
        float4 acc = make_float4(0, 0, 0, 0);
        for (int i = lane; i < in_f4; i += 32) { float4 v = in[f * in_f4 + i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        float4 *o = out + f * (long)out_f4;
        for (int i = lane; i < out_f4; i += 32) o[i] = acc;
    }
}
__global__ void write32_kernel(float *out, long frames, int out_f) {  // 4-byte stores, 128 B per warp instruction
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * (long)blockDim.x) >> 5;
    for (long f = warp; f < frames; f += nwarps) {
        float *o = out + f * (long)out_f;
#pragma unroll 8
        for (int i = lane; i < out_f; i += 32) o[i] = (float)i;
    }
}
template <typename F> float best_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    const long frames = 600000; const int out_f4 = 33024 / 16, in_f4 = 2048 / 16;
    float4 *in, *out; cudaMalloc(&in, frames * in_f4 * 16L); cudaMalloc(&out, frames * out_f4 * 16L);
    cudaMemset(in, 0, frames * in_f4 * 16L);
    const size_t obytes = frames * out_f4 * 16L, ibytes = frames * in_f4 * 16L;
    float ms_mix = best_ms([&] { mix_kernel<<<148 * 4, 512>>>(in, out, frames, out_f4, in_f4); });
    float ms_w32 = best_ms([&] { write32_kernel<<<148 * 4, 512>>>((float *)out, frames, out_f4 * 4); });
    float ms_set = best_ms([&] { cudaMemsetAsync(out, 1, obytes); });
    float ms_cpy = best_ms([&] { cudaMemcpyAsync(out, (char *)out + obytes / 2, obytes / 2, cudaMemcpyDeviceToDevice); });
    printf("{\"mix_2k_in_33k_out_gbs\": %.1f, \"write_4B_stores_gbs\": %.1f, \"memset_gbs\": %.1f, \"copy_rw_gbs\": %.1f}\n",
           (obytes + ibytes) / ms_mix / 1e6, obytes / ms_w32 / 1e6, obytes / ms_set / 1e6, obytes / ms_cpy / 1e6);
    return 0;
}
