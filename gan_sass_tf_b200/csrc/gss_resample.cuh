// gss_resample.cuh - scipy.signal.resample (Fourier method) as the reference calls it on non-16 kHz files
// (main.py:89-95): X = rfft(x); keep min(n, num)/2 + 1 bins (the shared Nyquist bin folded / split as SciPy does for real
// input); y = irfft(Y, num) * num / n.  n and num are arbitrary (44100 -> 16000 gives lengths with large prime factors),
// so both transforms are Bluestein chirp-z transforms: a length-L DFT becomes a circular convolution of length
// M = 2^ceil(log2(2L - 1)), carried out with a power-of-two Stockham FFT in global memory.  Everything is float64 (SciPy
// promotes the int16 samples to float64) and every phase is reduced exactly in integers (n^2 mod 2L, jk mod 2Ns) before
// sincospi, so the result agrees with SciPy's pocketfft / DUCC to ~1e-13.
//
// An edge op (one clip per demo run): O(M log M) work in ~3 log2 M + 5 small launches per transform; no attempt is made
// to fuse the passes.  The caller provides the workspace (gss_resample_workspace_bytes).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace gss {
namespace rs {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// e^{sign * i * pi * (n^2 mod 2L) / L}
__device__ __forceinline__ double2 chirp(int64_t n, int64_t L, int sign) {
    const unsigned long long r = ((unsigned long long)n * (unsigned long long)n) % (unsigned long long)(2 * L);   // n < 2^31
    double s, c;
    sincospi((double)r / (double)L, &s, &c);
    return make_double2(c, sign > 0 ? s : -s);
}

// a[j] = in[j] * w_s[j] (j < L), 0 (L <= j < M);  b[j] = conj(w_s[j]) for |j| < L (wrapped), 0 elsewhere
template <bool REAL_IN>
__global__ void __launch_bounds__(256) prepare_kernel(const double* __restrict__ in_re, const double2* __restrict__ in_c,
                                                      int64_t L, int64_t M, int sign, double2* __restrict__ a, double2* __restrict__ b) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        double2 av = make_double2(0.0, 0.0), bv = make_double2(0.0, 0.0);
        if (j < L) {
            const double2 w = chirp(j, L, sign);
            av = REAL_IN ? make_double2(in_re[j] * w.x, in_re[j] * w.y) : cmul(in_c[j], w);
            bv = make_double2(w.x, -w.y);
        } else if (M - j < L) {
            const double2 w = chirp(M - j, L, sign);
            bv = make_double2(w.x, -w.y);
        }
        a[j] = av; b[j] = bv;
    }
}

// one radix-2 Stockham pass (decimation in time, autosort): butterflies j < M/2, k = j mod Ns,
// out[(j - k) * 2 + k] = in[j] + w in[j + M/2], out[(j - k) * 2 + k + Ns] = in[j] - w in[j + M/2], w = e^{sign i pi k / Ns}
__global__ void __launch_bounds__(256) fft_pass_kernel(const double2* __restrict__ in, double2* __restrict__ out, int64_t M, int64_t Ns, int sign) {
    const int64_t half = M >> 1;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = j & (Ns - 1);
        double s, c;
        sincospi((double)k / (double)Ns, &s, &c);
        const double2 w = make_double2(c, sign > 0 ? s : -s);
        const double2 u = in[j], v = cmul(in[j + half], w);
        const int64_t o = ((j - k) << 1) + k;
        out[o] = make_double2(u.x + v.x, u.y + v.y);
        out[o + Ns] = make_double2(u.x - v.x, u.y - v.y);
    }
}

__global__ void __launch_bounds__(256) pointwise_kernel(double2* __restrict__ a, const double2* __restrict__ b, int64_t M) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) a[j] = cmul(a[j], b[j]);
}

// DFT bins from the convolution: X[k] = w_s[k] * c[k] / M, k < L
__global__ void __launch_bounds__(256) finish_kernel(const double2* __restrict__ c, int64_t L, int64_t M, int sign, double2* __restrict__ X) {
    const double inv = 1.0 / (double)M;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < L; k += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = cmul(c[k], chirp(k, L, sign));
        X[k] = make_double2(v.x * inv, v.y * inv);
    }
}

// scipy.signal.resample's spectrum surgery for real input, written as the full Hermitian spectrum of length num
__global__ void __launch_bounds__(256) resize_kernel(const double2* __restrict__ X, int64_t n, int64_t num, double2* __restrict__ Y) {
    const int64_t m = n < num ? n : num, nyq = m / 2 + 1;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < num; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t kk = k <= num / 2 ? k : num - k;            // the bin of the half spectrum this entry mirrors
        double2 v = make_double2(0.0, 0.0);
        if (kk < nyq) {
            v = X[kk];
            if (m % 2 == 0 && kk == m / 2) {
                if (num < n) { v.x *= 2.0; v.y = 0.0; }           // down-sampling: both halves of the bin fold onto the new Nyquist (irfft keeps its real part)
                else if (num > n) { v.x *= 0.5; v.y *= 0.5; }     // up-sampling: the old Nyquist bin is split between +f and -f
            }
            if (kk == 0 || (num % 2 == 0 && kk == num / 2)) v.y = 0.0;
            else if (k != kk) v.y = -v.y;                         // negative frequency: conjugate
        }
        Y[k] = v;
    }
}

__global__ void __launch_bounds__(256) real_scale_kernel(const double2* __restrict__ c, int64_t num, double scale, double* __restrict__ y) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < num; k += (int64_t)gridDim.x * blockDim.x) y[k] = c[k].x * scale;
}

inline int64_t pow2_at_least(int64_t v) { int64_t p = 1; while (p < v) p <<= 1; return p; }

}  // namespace rs
}  // namespace gss
