// write_pattern.cu -- does the ORDER in which warps visit frames change what HBM sustains for the
// frame path's output mix?  Every "frame" reads 8 KB (overlapping its neighbours by 6 KB, so ~2 KB
// is new) and writes 8+8+8+4+4 KB into five separate arrays with 128-byte rows per warp store,
// like mb_warp2048_kernel.  `chunk` consecutive frames go to one warp (the feature kernel uses 32).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
__global__ void __launch_bounds__(512, 1)
pattern_kernel(const float *in, float *buf, float *re, float *im, float *amp, float *pw, long frames, int chunk) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * (long)blockDim.x) >> 5;
    const long nchunks = (frames + chunk - 1) / chunk;
    for (long c = warp; c < nchunks; c += nwarps) {
        for (int j = 0; j < chunk; j++) {
            const long f = c * chunk + j;
            if (f >= frames) break;
            float acc = 0.f;
#pragma unroll 8
            for (int i = lane; i < 2048; i += 32) acc += __ldg(in + f * 512 + i);
            float *b = buf + f * 2048, *r = re + f * 2048, *m = im + f * 2048, *a = amp + f * 1024, *p = pw + f * 1024;
#pragma unroll 8
            for (int i = lane; i < 2048; i += 32) { __stcs(b + i, acc); __stcs(r + i, acc + 1.f); __stcs(m + i, acc + 2.f); }
#pragma unroll 8
            for (int i = lane; i < 1024; i += 32) { __stcs(a + i, acc + 3.f); __stcs(p + i, acc + 4.f); }
        }
    }
}
int main(int argc, char **argv) {
    const long frames = 700000;
    float *in, *buf, *re, *im, *amp, *pw;
    cudaMalloc(&in, (frames * 512 + 2048) * 4); cudaMemset(in, 0, (frames * 512 + 2048) * 4);
    cudaMalloc(&buf, frames * 2048 * 4); cudaMalloc(&re, frames * 2048 * 4); cudaMalloc(&im, frames * 2048 * 4);
    cudaMalloc(&amp, frames * 1024 * 4); cudaMalloc(&pw, frames * 1024 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int chunks[] = {32, 1, 2, 4, 8, 16, 32, 1};
    printf("{");
    for (int t = 0; t < 8; t++) {
        float best = 1e30f;
        for (int r = 0; r < 4; r++) {
            cudaEventRecord(a); pattern_kernel<<<148, 512>>>(in, buf, re, im, amp, pw, frames, chunks[t]); cudaEventRecord(b);
            cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms;
        }
        printf("%s\"chunk%d_%d\": {\"Mframes_s\": %.1f, \"GBs\": %.0f}", t ? ", " : "", chunks[t], t, frames / best / 1e3, frames * 34816.0 / best / 1e6);
    }
    printf("}\n");
    if (cudaGetLastError() != cudaSuccess) return 1;
    return 0;
}
