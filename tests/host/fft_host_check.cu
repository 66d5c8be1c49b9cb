// Host-side check of meyda_b200/csrc/mb_fft.cuh: the register FFTs compiled for the CPU (the packed f32x2 helpers fall
// back to scalar float32 there) against a float64 DFT.  Covers the index, sign and twiddle logic of fft_reg<R>, the
// complex multiply built from a swapped pair and a signed scalar, and the real-FFT split formula of the kernels.
// Test infrastructure only: built and run by tests/test_fft_host.py.
#include <cmath>
#include <cstdio>

#include "../../meyda_b200/csrc/mb_fft.cuh"

template <int R, int BITS>
static double check_fft() {
    float2 v[R];
    double xr[R], xi[R];
    for (int i = 0; i < R; i++) {
        xr[i] = (float)(sin(1.3 * i) + 0.1 * i);
        xi[i] = (float)cos(2.1 * i * i);
        v[i] = make_float2((float)xr[i], (float)xi[i]);
    }
    mbfft::fft_reg<R>(v);
    double err = 0, peak = 0;
    for (int k = 0; k < R; k++) {
        double sr = 0, si = 0;
        for (int n = 0; n < R; n++) {  // forward sign +i (lib/jsfft/fft.js:145)
            const double a = 2 * M_PI * k * n / R;
            sr += xr[n] * cos(a) - xi[n] * sin(a);
            si += xr[n] * sin(a) + xi[n] * cos(a);
        }
        const float2 o = v[mbfft::brev<BITS>(k)];
        err = fmax(err, fmax(fabs(o.x - sr), fabs(o.y - si)));
        peak = fmax(peak, fmax(fabs(sr), fabs(si)));
    }
    return err / peak;
}

int main() {
    printf("fft32 %.3e\nfft16 %.3e\nfft8 %.3e\nfft4 %.3e\nfft2 %.3e\n", check_fft<32, 5>(), check_fft<16, 4>(), check_fft<8, 3>(),
           check_fft<4, 2>(), check_fft<2, 1>());
    // complex multiply
    const float2 d = make_float2(0.3f, -1.7f), t = make_float2(0.6f, 0.8f);
    const float2 p = mbx2::cmul(d, t);
    printf("cmul %.3e\n", fmax(fabs(p.x - (0.3 * 0.6 - (-1.7) * 0.8)), fabs(p.y - (0.3 * 0.8 + (-1.7) * 0.6))));
    // the split of kernel_warp.cu: Z = hsc E + sy w + dx (w.y, -w.x) with E = a + conj b, F = a - conj b, against
    // Z = ((a + conj b) / 2 + exp(i th) (a - conj b) / (2 i)) sc written out in float64 (w = (sc / 2) exp(i th))
    const float2 a = make_float2(1.25f, -0.5f), b = make_float2(-0.75f, 2.0f);
    const double th = 0.37, sc = 0.022;
    const float2 w = make_float2((float)(0.5 * sc * cos(th)), (float)(0.5 * sc * sin(th)));
    const float hsc = (float)(0.5 * sc);
    const float2 cb = make_float2(b.x, -b.y);
    const float2 E = mbx2::add(a, cb), F = mbx2::sub(a, cb);
    const float2 Z = mbx2::fma(E, mbx2::bc(hsc), mbx2::fma(w, mbx2::bc(F.y), mbx2::mul(make_float2(w.y, -w.x), mbx2::bc(F.x))));
    const double er = 0.5 * (a.x + b.x), ei = 0.5 * (a.y - b.y), orr = 0.5 * (a.y + b.y), oi = -0.5 * (a.x - b.x);
    const double zr = (er + (cos(th) * orr - sin(th) * oi)) * sc, zi = (ei + (cos(th) * oi + sin(th) * orr)) * sc;
    printf("split %.3e\n", fmax(fabs(Z.x - zr), fabs(Z.y - zi)) / fmax(fabs(zr), fabs(zi)));
    return 0;
}
