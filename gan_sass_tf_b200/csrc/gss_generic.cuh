// gss_generic.cuh - any-size (power of two, 64..4096) STFT / iSTFT / mask-iSTFT kernels.
//
// The register-resident streaming kernels of gss_stream.cuh exist for the FFT sizes the
// three-pass 8 x M x 8 geometry covers; every other size runs here: one CTA per
// (row, tile of frames), the real transform as an N/2-point complex Stockham FFT
// (radix 4, one radix-2 stage when log2(N/2) is odd) in shared memory, overlap-add and
// the window-square norm accumulated in a shared-memory tile exactly the way
// scipy.signal.istft does it (acc += w*y, nrm += w^2, divide where nrm > 1e-10).
// Same semantics, same packed layout, lower throughput than the streaming path.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "gss_stream.cuh"

namespace gss {
namespace gen {

constexpr int THREADS = 256;

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// tw[k] = W_N^k = e^{-2 pi i k / N}, k in [0, N/2)
__device__ __forceinline__ void build_tw(float2* tw, int N) {
    for (int k = threadIdx.x; k < N / 2; k += THREADS) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)N, &s, &c);
        tw[k] = make_float2(c, s);
    }
}
// W_M^e for e in [0, M), M = N/2, from the W_N table; conjugated for the inverse direction
template <bool INV>
__device__ __forceinline__ float2 tw_m(const float2* tw, int M, int e) {
    float2 w = (e < M / 2) ? tw[2 * e] : make_float2(-tw[2 * e - M].x, -tw[2 * e - M].y);
    return INV ? cconj(w) : w;
}
__device__ __forceinline__ float hann_tw(const float2* tw, int N, int i) {   // 0.5 - 0.5 cos(2 pi i / N)
    float c = (i < N / 2) ? tw[i].x : -tw[i - N / 2].x;
    return 0.5f - 0.5f * c;
}

// M-point complex FFT (unnormalised), Stockham autosort, ping-pong between b0 and b1.
// Returns the buffer holding the result.  Ends with a __syncthreads().
template <bool INV>
__device__ float2* fft_smem(float2* b0, float2* b1, int M, const float2* tw) {
    float2* in = b0; float2* out = b1;
    for (int Ns = 1; Ns < M;) {
        if (M / Ns >= 4) {
            const int q = M / 4;
            for (int j = threadIdx.x; j < q; j += THREADS) {
                const int k = j & (Ns - 1);
                const int e = k * (q / Ns);                // angle index: k * M / (4 Ns)
                float2 v0 = in[j];
                float2 v1 = cmulf(in[j + q], tw_m<INV>(tw, M, e));
                float2 v2 = cmulf(in[j + 2 * q], tw_m<INV>(tw, M, 2 * e));
                float2 v3 = cmulf(in[j + 3 * q], tw_m<INV>(tw, M, 3 * e));
                float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
                float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y), a3 = make_float2(v1.x - v3.x, v1.y - v3.y);
                // forward: a3 * (-i) = (a3.y, -a3.x); inverse: a3 * (+i) = (-a3.y, a3.x)
                float2 r = INV ? make_float2(-a3.y, a3.x) : make_float2(a3.y, -a3.x);
                const int j0 = ((j - k) << 2) + k;
                out[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
                out[j0 + Ns] = make_float2(a1.x + r.x, a1.y + r.y);
                out[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
                out[j0 + 3 * Ns] = make_float2(a1.x - r.x, a1.y - r.y);
            }
            Ns *= 4;
        } else {
            const int h = M / 2;
            for (int j = threadIdx.x; j < h; j += THREADS) {
                const int k = j & (Ns - 1);
                const int e = k * (h / Ns);
                float2 v0 = in[j];
                float2 v1 = cmulf(in[j + h], tw_m<INV>(tw, M, e));
                const int j0 = ((j - k) << 1) + k;
                out[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
                out[j0 + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
            }
            Ns *= 2;
        }
        __syncthreads();
        float2* t = in; in = out; out = t;
    }
    return in;
}

template <typename TIn>
__device__ __forceinline__ float sample_at(const TIn* row, int64_t n, int64_t p) {
    return (p >= 0 && p < n) ? (float)__ldg(row + p) : 0.f;
}

// frame t of the padded signal, windowed, as M = N/2 complex points z[m] = x[2m] + i x[2m+1]
template <typename TIn>
__device__ __forceinline__ void load_frame(const TIn* row, int64_t n, int64_t t, int N, int H, const float2* tw, float2* z, float scale) {
    const int64_t p0 = t * H - N / 2;
    for (int m = threadIdx.x; m < N / 2; m += THREADS) {
        float a = sample_at(row, n, p0 + 2 * m) * hann_tw(tw, N, 2 * m) * scale;
        float b = sample_at(row, n, p0 + 2 * m + 1) * hann_tw(tw, N, 2 * m + 1) * scale;
        z[m] = make_float2(a, b);
    }
    __syncthreads();
}

// bins of a real transform from the half-size complex transform Z: slot k in [0, M) gives the
// packed pair (re, im) = (Re X[k], Im X[k]); slot 0 = (X[0], X[M]) (app/utils.py:22-26)
__device__ __forceinline__ float2 real_bin(const float2* Z, int M, const float2* tw, int k) {
    if (k == 0) return make_float2(Z[0].x + Z[0].y, Z[0].x - Z[0].y);
    float2 a = Z[k], b = cconj(Z[M - k]);
    float2 s = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
    float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y - b.y));
    float2 wd = cmulf(tw[k], d);                       // W_N^k (Z[k] - conj Z[M-k]) / 2
    return make_float2(s.x + wd.y, s.y - wd.x);        // s - i * wd
}
// inverse of real_bin: packed pairs P[k] (slot 0 = (X[0], X[M])) -> Z[k] of the half-size inverse
__device__ __forceinline__ float2 half_bin(const float2* P, int M, const float2* tw, int k) {
    if (k == 0) return make_float2(P[0].x + P[0].y, P[0].x - P[0].y);
    float2 a = P[k], b = cconj(P[M - k]);
    float2 s = make_float2(a.x + b.x, a.y + b.y), d = make_float2(a.x - b.x, a.y - b.y);
    float2 wd = cmulf(cconj(tw[k]), d);                // W_N^{-k} (X[k] - conj X[M-k])
    return make_float2(s.x - wd.y, s.y + wd.x);        // s + i * wd
}

struct StftArgs {
    const void* wave; float* feat;
    int64_t B, n, ld, T;
    int N, H, fpc;           // frames per CTA
    int log; float eps;
};

template <typename TIn>
__global__ void __launch_bounds__(THREADS) stft_kernel(const StftArgs p) {
    extern __shared__ float2 sm[];
    const int N = p.N, M = N / 2;
    float2* tw = sm; float2* b0 = sm + M; float2* b1 = b0 + M;
    build_tw(tw, N);
    __syncthreads();
    const int64_t b = blockIdx.y;
    const TIn* row = reinterpret_cast<const TIn*>(p.wave) + b * p.ld;
    const int64_t t0 = (int64_t)blockIdx.x * p.fpc, t1 = min(t0 + p.fpc, p.T);
    for (int64_t t = t0; t < t1; ++t) {
        load_frame(row, p.n, t, N, p.H, tw, b0, 2.0f / (float)N);      // 1 / sum(w)
        const float2* Z = fft_smem<false>(b0, b1, M, tw);
        float* out = p.feat + (b * p.T + t) * N;
        for (int k = threadIdx.x; k < M; k += THREADS) {
            float2 x = real_bin(Z, M, tw, k);
            if (p.log) { float g = log_gain(x.x, x.y, p.eps); x.x *= g; x.y *= g; }
            out[k] = x.x; out[M + k] = x.y;
        }
        __syncthreads();
    }
}

// overlap-add tile: frames [f0, f1) contribute to padded samples [f0*H, (f1-1)*H + N); the CTA
// owns padded samples [own0, own1) and writes them (shifted by -N/2) after the last frame.
struct OlaArgs {
    const float* feat;       // [rows, T, N] packed features (FROM_WAVE = false)
    const float* wave;       // [B, ld]                      (FROM_WAVE = true)
    const float* mask;       // [B, S, T, N/2]               (FROM_WAVE = true)
    float* out;              // [rows, ld_out]
    int64_t rows, n, ld, T, ld_out;
    int N, H, S, ft;         // ft = output hops per CTA
    int exp; float eps;
};

template <bool FROM_WAVE>
__global__ void __launch_bounds__(THREADS) ola_kernel(const OlaArgs p) {
    extern __shared__ float2 sm[];
    const int N = p.N, M = N / 2, H = p.H, R = N / H;
    float2* tw = sm; float2* b0 = sm + M; float2* b1 = b0 + M;
    const int span = (p.ft + R - 1) * H + N;            // samples any frame of this tile can touch
    float* acc = reinterpret_cast<float*>(b1 + M); float* nrm = acc + span;
    build_tw(tw, N);
    for (int i = threadIdx.x; i < 2 * span; i += THREADS) acc[i] = 0.f;
    __syncthreads();
    const int64_t r = blockIdx.y;                        // output row (b*S + s when FROM_WAVE)
    const int64_t h0 = (int64_t)blockIdx.x * p.ft;       // first owned hop
    const int64_t own0 = h0 * H, own1 = min((h0 + p.ft) * H, (p.T - 1) * (int64_t)H + N);
    const int64_t f0 = max((int64_t)0, h0 - (R - 1)), f1 = min(p.T, h0 + p.ft);
    const int64_t base = f0 * H;                         // padded position of acc[0]
    const int64_t b = FROM_WAVE ? r / p.S : r;
    for (int64_t t = f0; t < f1; ++t) {
        float2* P;                                       // packed pairs of this frame's spectrum
        if (FROM_WAVE) {
            load_frame(p.wave + b * p.ld, p.n, t, N, H, tw, b0, 2.0f / (float)N);
            float2* Z = fft_smem<false>(b0, b1, M, tw);
            P = (Z == b0) ? b1 : b0;
            const float* mk = p.mask + (r * p.T + t) * M;
            for (int k = threadIdx.x; k < M; k += THREADS) {
                float2 x = real_bin(Z, M, tw, k);
                float g = __ldg(mk + k);
                P[k] = make_float2(x.x * g, x.y * g);
            }
        } else {
            P = b0;
            const float* f = p.feat + (r * p.T + t) * N;
            for (int k = threadIdx.x; k < M; k += THREADS) {
                float2 x = make_float2(__ldg(f + k), __ldg(f + M + k));
                if (p.exp) { float g = exp_gain(x.x, x.y, p.eps); x.x *= g; x.y *= g; }
                P[k] = x;
            }
        }
        __syncthreads();
        float2* Zin = (P == b0) ? b1 : b0;
        for (int k = threadIdx.x; k < M; k += THREADS) Zin[k] = half_bin(P, M, tw, k);
        __syncthreads();
        const float2* z = fft_smem<true>(Zin, P, M, tw);
        const int off = (int)(t * H - base);
        for (int m = threadIdx.x; m < M; m += THREADS) {
            // frame = sum(w) * irfft = (N/2)/N * unnormalised inverse
            float w0 = hann_tw(tw, N, 2 * m), w1 = hann_tw(tw, N, 2 * m + 1);
            acc[off + 2 * m] += 0.5f * z[m].x * w0; nrm[off + 2 * m] += w0 * w0;
            acc[off + 2 * m + 1] += 0.5f * z[m].y * w1; nrm[off + 2 * m + 1] += w1 * w1;
        }
        __syncthreads();
    }
    // trimmed output: padded [N/2, N/2 + (T-1)H)
    const int64_t lo = max(own0, (int64_t)N / 2), hi = min(own1, (int64_t)N / 2 + (p.T - 1) * H);
    float* orow = p.out + r * p.ld_out;
    for (int64_t pp = lo + threadIdx.x; pp < hi; pp += THREADS) {
        float nv = nrm[pp - base];
        orow[pp - N / 2] = acc[pp - base] / (nv > 1e-10f ? nv : 1.0f);
    }
}

}  // namespace gen
}  // namespace gss
