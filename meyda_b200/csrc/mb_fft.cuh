// mb_fft.cuh -- power-of-two FFTs held entirely in registers (R <= 32 points per thread):
// radix-2 decimation in frequency, forward sign +i as in lib/jsfft/fft.js:145, compile-time
// twiddles with the trivial ones (1, i, (+-1+i)/sqrt2) folded.  Natural order in, X[k] is left in
// v[brev(k)] (bit reversal over log2 R bits); register indices are compile-time, so that costs nothing.
//
// The arithmetic is sm_100's packed float32 pair (add / mul / fma .f32x2 -> FADD2 / FMUL2 / FFMA2): a complex
// value is one (re, im) register pair, a complex add is ONE instruction and a complex multiply TWO, because the
// instructions take a swapped view of an operand pair (.LO_HI), a per-half sign and a broadcast scalar or
// immediate for free.  Each half is an IEEE round-to-nearest float32 operation, as the scalar forms are.
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace mbx2 {

#ifdef __CUDA_ARCH__
using u64 = unsigned long long;
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ u64 pk(float2 v) { return pk(v.x, v.y); }
__device__ __forceinline__ float2 upk(u64 r) {
    float2 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
    return v;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// (a.x + b.x, a.y + b.y) and friends, on float2 values: the packing moves vanish in register allocation when the
// float2 already lives in an aligned pair (loaded by LDS.64 / produced by another packed instruction)
__device__ __forceinline__ float2 add(float2 a, float2 b) { return upk(add2(pk(a), pk(b))); }
__device__ __forceinline__ float2 sub(float2 a, float2 b) { return upk(sub2(pk(a), pk(b))); }
__device__ __forceinline__ float2 mul(float2 a, float2 b) { return upk(mul2(pk(a), pk(b))); }
__device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c) { return upk(fma2(pk(a), pk(b), pk(c))); }
#else  // host: the same functions in scalar form (CPU checks of the index / sign logic; never the product path)
inline float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
inline float2 mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif

__host__ __device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
__host__ __device__ __forceinline__ float2 swap(float2 a) { return make_float2(a.y, a.x); }
// d * (c + i s) = (d.x c - d.y s, d.y c + d.x s): FMUL2 + FFMA2
__host__ __device__ __forceinline__ float2 cmul(float2 d, float c, float s) {
    return fma(swap(d), make_float2(-s, s), mul(d, bc(c)));
}
__host__ __device__ __forceinline__ float2 cmul(float2 d, float2 t) { return cmul(d, t.x, t.y); }

}  // namespace mbx2

namespace mbfft {

constexpr float kCos32[16] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f,
                                         7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f,
                                         0.0f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f,
                                         -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
constexpr float kSin32[16] = {0.000000000e+00f, 1.950903220e-01f, 3.826834324e-01f, 5.555702330e-01f,
                                         7.071067812e-01f, 8.314696123e-01f, 9.238795325e-01f, 9.807852804e-01f,
                                         1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f,
                                         7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f};

template <int E>  // d * exp(+2 pi i E / 32), 0 <= E < 16
__host__ __device__ __forceinline__ float2 mul_w32(float2 d) {
    constexpr float R = 7.071067812e-01f;
    if constexpr (E == 0) return d;
    else if constexpr (E == 8) return make_float2(-d.y, d.x);
    else if constexpr (E == 4) return mbx2::mul(mbx2::add(d, make_float2(-d.y, d.x)), mbx2::bc(R));           // (d.x - d.y, d.y + d.x) R
    else if constexpr (E == 12) return mbx2::mul(mbx2::sub(make_float2(-d.y, d.x), d), mbx2::bc(R));          // (-d.y - d.x, d.x - d.y) R
    else {
        constexpr float c = kCos32[E], s = kSin32[E];
        return mbx2::cmul(d, c, s);
    }
}
// butterfly I of the stage with half-span H (twiddle exp(2 pi i J / 2H) = exp(2 pi i J (16/H) / 32))
template <int R, int H, int I>
__host__ __device__ __forceinline__ void bfly(float2 (&v)[R]) {
    constexpr int B = (I / H) * 2 * H, J = I % H;
    const float2 u = v[B + J], w = v[B + J + H];
    v[B + J] = mbx2::add(u, w);
    v[B + J + H] = mul_w32<J * (16 / H)>(mbx2::sub(u, w));
}
template <int R, int H, int... I>
__host__ __device__ __forceinline__ void stage(float2 (&v)[R], std::integer_sequence<int, I...>) {
    (bfly<R, H, I>(v), ...);
}
template <int R, int H>
__host__ __device__ __forceinline__ void stages_from(float2 (&v)[R]) {
    stage<R, H>(v, std::make_integer_sequence<int, R / 2>{});
    if constexpr (H > 1) stages_from<R, H / 2>(v);
}
template <int R>
__host__ __device__ __forceinline__ void fft_reg(float2 (&v)[R]) {
    static_assert(R >= 2 && R <= 32 && (R & (R - 1)) == 0, "2 <= R <= 32, power of two");
    stages_from<R, R / 2>(v);
}
template <int BITS>
__host__ __device__ constexpr int brev(int k) {
    int r = 0;
    for (int b = 0; b < BITS; b++) r |= ((k >> b) & 1) << (BITS - 1 - b);
    return r;
}

}  // namespace mbfft
