// mb_fft.cuh -- power-of-two FFTs held entirely in registers (R <= 32 points per thread):
// radix-2 decimation in frequency, forward sign +i as in lib/jsfft/fft.js:145, compile-time
// twiddles with the trivial ones (1, i, (+-1+i)/sqrt2) folded.  Natural order in, X[k] is left in
// v[brev(k)] (bit reversal over log2 R bits); register indices are compile-time, so that costs nothing.
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace mbfft {

__device__ constexpr float kCos32[16] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f,
                                         7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f,
                                         0.0f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f,
                                         -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
__device__ constexpr float kSin32[16] = {0.000000000e+00f, 1.950903220e-01f, 3.826834324e-01f, 5.555702330e-01f,
                                         7.071067812e-01f, 8.314696123e-01f, 9.238795325e-01f, 9.807852804e-01f,
                                         1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f,
                                         7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f};

template <int E>  // d * exp(+2 pi i E / 32), 0 <= E < 16
__device__ __forceinline__ float2 mul_w32(float2 d) {
    constexpr float R = 7.071067812e-01f;
    if constexpr (E == 0) return d;
    else if constexpr (E == 8) return make_float2(-d.y, d.x);
    else if constexpr (E == 4) return make_float2((d.x - d.y) * R, (d.x + d.y) * R);
    else if constexpr (E == 12) return make_float2((-d.x - d.y) * R, (d.x - d.y) * R);
    else {
        constexpr float c = kCos32[E], s = kSin32[E];
        return make_float2(d.x * c - d.y * s, d.x * s + d.y * c);
    }
}
// butterfly I of the stage with half-span H (twiddle exp(2 pi i J / 2H) = exp(2 pi i J (16/H) / 32))
template <int R, int H, int I>
__device__ __forceinline__ void bfly(float2 (&v)[R]) {
    constexpr int B = (I / H) * 2 * H, J = I % H;
    const float2 u = v[B + J], w = v[B + J + H];
    v[B + J] = make_float2(u.x + w.x, u.y + w.y);
    v[B + J + H] = mul_w32<J * (16 / H)>(make_float2(u.x - w.x, u.y - w.y));
}
template <int R, int H, int... I>
__device__ __forceinline__ void stage(float2 (&v)[R], std::integer_sequence<int, I...>) {
    (bfly<R, H, I>(v), ...);
}
template <int R, int H>
__device__ __forceinline__ void stages_from(float2 (&v)[R]) {
    stage<R, H>(v, std::make_integer_sequence<int, R / 2>{});
    if constexpr (H > 1) stages_from<R, H / 2>(v);
}
template <int R>
__device__ __forceinline__ void fft_reg(float2 (&v)[R]) {
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
