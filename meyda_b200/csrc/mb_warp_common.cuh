// mb_warp_common.cuh -- pieces shared by the warp kernels (kernel_warp.cu, kernel_warp_mf.cu): mbarrier / TMA bulk-copy
// wrappers, L2 cache policies, streaming stores, MUFU approximations, and the 32-frame scalar stash.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbwarp {

constexpr int kChunk = 32;  // frames per work unit of a warp (= columns of its scalar stash)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MB_DONE_%=;\n\t"
        "bra MB_WAIT_%=;\n\t"
        "MB_DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA bulk copies (1-D): global -> shared with mbarrier completion, shared -> global as a bulk group.
// L2 policies: a frame's samples are read again by the next three frames (hop 512 of 2048), while the
// output stream is written once and is 16x larger -- keep the former, let the latter go first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar,
                                          uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes), "l"(policy)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// streaming store: written once, never read back by this kernel
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One MUFU each, <= 1 ulp.  .ftz: the arguments are |Z|^2 and |Z| of a frame whose samples were brought
// into [2^-40, 2^40] (see kscale), so a subnormal argument is 2^-86 below the frame's scale: flushing it
// to zero changes nothing that float32 could have resolved.  0 -> 0 / -inf, NaN -> NaN.
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float log2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}


// x^0.23 (loudness.js:60) and ln x (mfcc.js:63) for the band finish: two MUFUs instead of the 45- / 20-instruction
// library routines.  Subnormals are kept (no .ftz), 0 -> 0 / -inf, inf -> inf, NaN -> NaN; the error (~2e-7
// relative for the power, ~2e-6 absolute for the logarithm of a float32 energy) is that of a float32 rounding.
__device__ __forceinline__ float pow023_approx(float x) {
    float l, r;
    asm("lg2.approx.f32 %0, %1;" : "=f"(l) : "f"(x));
    asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(0.23f * l));
    return r;
}
__device__ __forceinline__ float ln_approx(float x) {
    float l;
    asm("lg2.approx.f32 %0, %1;" : "=f"(l) : "f"(x));
    return l * 0.6931471805599453f;
}

// a double parked as two float rows of the stash
__device__ __forceinline__ void stash_put_d(float (*st)[kChunk], int row, int col, double v) {
    st[row][col] = __int_as_float(__double2hiint(v));
    st[row + 1][col] = __int_as_float(__double2loint(v));
}
__device__ __forceinline__ double stash_get_d(float (*st)[kChunk], int row, int col) {
    return __hiloint2double(__float_as_int(st[row][col]), __float_as_int(st[row + 1][col]));
}

}  // namespace mbwarp
