// microbench.cu -- measures the non-tensor FP32 FFMA and FP64 DFMA peaks and a
// shared-memory bandwidth figure on the box, for the FP32/FP64 roofline
// denominators SURVEY.md 8(d) asks to "measure with an FFMA micro-kernel and record".
#include <cuda_runtime.h>
#include <stdio.h>

template <typename T>
__global__ void fma_kernel(T *out, int iters) {
    T a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T b = (T)1.000001, c = (T)0.5;
    for (int i = 0; i < iters; i++) {
        a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
        a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void smem_kernel(float *out, int iters) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, i, i, i);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int i = 0; i < iters; i++) {
        float4 v = buf[idx & 1023];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        idx += 256;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 1 << 16;
    void *out; cudaMalloc(&out, (size_t)blocks * threads * 8);
    float ms32 = time_ms([&] { fma_kernel<float><<<blocks, threads>>>((float *)out, iters); });
    float ms64 = time_ms([&] { fma_kernel<double><<<blocks, threads>>>((double *)out, iters / 4); });
    float mssm = time_ms([&] { smem_kernel<<<blocks, threads>>>((float *)out, iters); });
    double f32 = 2.0 * 8 * iters * (double)blocks * threads / (ms32 * 1e-3) / 1e12;
    double f64 = 2.0 * 8 * (iters / 4) * (double)blocks * threads / (ms64 * 1e-3) / 1e12;
    double smem = 16.0 * iters * (double)blocks * threads / (mssm * 1e-3) / 1e12;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp32_ffma_tflops\": %.2f, \"fp64_dfma_tflops\": %.2f, "
           "\"smem_lds128_tbs\": %.2f}\n", p.name, sms, f32, f64, smem);
    return 0;
}
