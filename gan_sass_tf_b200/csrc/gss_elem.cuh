// gss_elem.cuh - the element-wise and reduction ops that sit beside the transforms:
// to_log_signal / to_exp_signal (app/ops.py:228-251), the mask multiply,
// batch_cross_snr (app/ops.py:191-225), the auto-encoder partial (main.py:353-361)
// and the int16 WAV normalisation (main.py:112-116).  All HBM-bound streaming kernels:
// 128-bit accesses, grid-stride loops sized to the SM count by the host.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "gss_stream.cuh"

namespace gss {

// rows x N packed features; a thread handles bins k..k+3 of both halves of a row
template <bool EXP>
__global__ void __launch_bounds__(256) logexp_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                     int64_t rows, int N, float eps) {
    const int q = N / 8;                              // float4 per half row
    const int64_t total = rows * q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / q; const int k4 = (int)(i - r * q);
        const float4* pr = reinterpret_cast<const float4*>(in + r * N) + k4;
        float4 re = __ldg(pr), im = __ldg(pr + q);
        float g;
#define GSS_G(c) g = EXP ? exp_gain(re.c, im.c, eps) : log_gain(re.c, im.c, eps); re.c *= g; im.c *= g;
        GSS_G(x) GSS_G(y) GSS_G(z) GSS_G(w)
#undef GSS_G
        float4* po = reinterpret_cast<float4*>(out + r * N) + k4;
        po[0] = re; po[q] = im;
    }
}

// mix [B,T,N], mask [B,S,T,N/2] -> out [B*S,T,N]
__global__ void __launch_bounds__(256) apply_mask_kernel(const float* __restrict__ mix, const float* __restrict__ mask,
                                                         float* __restrict__ out, int64_t B, int S, int64_t T, int N) {
    const int q = N / 8;
    const int64_t total = B * S * T * q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t rowo = i / q; const int k4 = (int)(i - rowo * q);   // rowo = (b*S + s)*T + t
        const int64_t bs = rowo / T, t = rowo - bs * T, b = bs / S;
        const float4* pm = reinterpret_cast<const float4*>(mix + (b * T + t) * N) + k4;
        float4 g = __ldg(reinterpret_cast<const float4*>(mask + rowo * (N / 2)) + k4);
        float4 re = __ldg(pm), im = __ldg(pm + q);
        re.x *= g.x; re.y *= g.y; re.z *= g.z; re.w *= g.w;
        im.x *= g.x; im.y *= g.y; im.z *= g.z; im.w *= g.w;
        float4* po = reinterpret_cast<float4*>(out + rowo * N) + k4;
        po[0] = re; po[q] = im;
    }
}

// Overlap-add weight of scipy.signal.istft at sample p of the trimmed output: norm[p] = sum over the frames t in
// [0, T) that cover it of hann[p + N/2 - t*H]^2, with SciPy's guard (norm > 1e-10, else 1).  out = in / norm
// (INV) or in * norm, times `scale`.  Used by the adjoints of the transforms: d iSTFT^T = (N/2) STFT(g / norm),
// d STFT^T = (4/N) norm * iSTFT(g').
template <bool INV>
__global__ void __launch_bounds__(256) ola_norm_scale_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows,
                                                             int64_t len, int64_t ld_in, int64_t ld_out, int64_t T, int N, int H, float scale) {
    const int64_t total = rows * len;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / len, p = i - r * len;
        const int64_t pp = p + N / 2;
        int64_t tlo = pp - (N - 1); tlo = tlo <= 0 ? 0 : (tlo + H - 1) / H;
        int64_t thi = pp / H; if (thi > T - 1) thi = T - 1;
        float nrm = 0.f;
        for (int64_t t = tlo; t <= thi; ++t) {
            const float w = 0.5f - 0.5f * cospif(2.0f * (float)(pp - t * H) / (float)N);
            nrm = fmaf(w, w, nrm);
        }
        nrm = nrm > 1e-10f ? nrm : 1.0f;
        const float v = __ldg(in + r * ld_in + p);
        out[r * ld_out + p] = scale * (INV ? v / nrm : v * nrm);
    }
}

// packed features: every slot times c_all, the DC slot (0) and the Nyquist slot (N/2) times c_edge in addition
__global__ void __launch_bounds__(256) scale_packed_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int N,
                                                           float c_all, float c_edge) {
    const int q = N / 4;
    const int64_t total = rows * q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k4 = (int)(i % q);
        float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
        v.x *= c_all; v.y *= c_all; v.z *= c_all; v.w *= c_all;
        if (k4 == 0 || k4 == q / 2) v.x *= c_edge;
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    v = l < nw ? red[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// Reductions over long rows (metrics of main.py:353-361 / :446-457, min / max of main.py:112-115).  One thread-block
// CLUSTER per row: the CTAs of a cluster split the row, reduce their share, and CTA 0 collects the partials of its
// peers through distributed shared memory - no workspace, no atomics, a deterministic order of additions, and rows x
// cluster-size CTAs in flight (the reference's defaults give 24 rows: one CTA each would leave 5/6 of the GPU idle).
// The cluster size is a launch attribute chosen by the host (1 ... 8).
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// read a float of CTA `rank`'s shared memory at the address `p` has in this CTA
__device__ __forceinline__ float dsmem_ld(const float* p, unsigned rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

constexpr int SNR_KMAX = 8;     // outputs compared per pass over `clear`

// one cluster per (b, i): ONE pass over clear_i computes mean clear_i^2 and mean (clear_i - noisy_k)^2 for every
// output k (groups of SNR_KMAX), snr[b,i,k] = c*(ln(mean clear_i^2 + eps) - ln(mean (clear_i - noisy_k)^2 + eps))
__global__ void __launch_bounds__(256) cross_snr_kernel(const float* __restrict__ clear, const float* __restrict__ noisy,
                                                        int m, int n, int64_t L, float eps, float* __restrict__ snr) {
    __shared__ float red[32];
    __shared__ float part[SNR_KMAX + 1];
    const unsigned cr = cluster_rank(), cs = cluster_size();
    const int64_t id = blockIdx.x / cs;                       // (b, i)
    const int64_t b = id / m;
    const float* pc = clear + id * L;
    const bool v4 = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(clear) | reinterpret_cast<uintptr_t>(noisy)) & 15) == 0;
    for (int k0 = 0; k0 < n; k0 += SNR_KMAX) {
        const int nk = min(SNR_KMAX, n - k0);
        const float* pn = noisy + (b * n + k0) * L;
        float sp = 0.f, np[SNR_KMAX];
#pragma unroll
        for (int k = 0; k < SNR_KMAX; ++k) np[k] = 0.f;
        if (v4) {
            for (int64_t x = (int64_t)cr * blockDim.x + threadIdx.x; x < L / 4; x += (int64_t)cs * blockDim.x) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(pc) + x);
                sp = fmaf(c4.x, c4.x, sp); sp = fmaf(c4.y, c4.y, sp); sp = fmaf(c4.z, c4.z, sp); sp = fmaf(c4.w, c4.w, sp);
#pragma unroll
                for (int k = 0; k < SNR_KMAX; ++k)
                    if (k < nk) {
                        const float4 n4 = __ldg(reinterpret_cast<const float4*>(pn + k * L) + x);
                        float d;
                        d = c4.x - n4.x; np[k] = fmaf(d, d, np[k]); d = c4.y - n4.y; np[k] = fmaf(d, d, np[k]);
                        d = c4.z - n4.z; np[k] = fmaf(d, d, np[k]); d = c4.w - n4.w; np[k] = fmaf(d, d, np[k]);
                    }
            }
        } else {
            for (int64_t x = (int64_t)cr * blockDim.x + threadIdx.x; x < L; x += (int64_t)cs * blockDim.x) {
                const float c = __ldg(pc + x);
                sp = fmaf(c, c, sp);
#pragma unroll
                for (int k = 0; k < SNR_KMAX; ++k)
                    if (k < nk) { const float d = c - __ldg(pn + k * L + x); np[k] = fmaf(d, d, np[k]); }
            }
        }
        sp = block_sum(sp, red);
        if (threadIdx.x == 0) part[SNR_KMAX] = sp;
#pragma unroll
        for (int k = 0; k < SNR_KMAX; ++k)
            if (k < nk) { const float v = block_sum(np[k], red); if (threadIdx.x == 0) part[k] = v; }
        cluster_sync_all();                                    // every CTA's partials are in its shared memory
        if (cr == 0 && threadIdx.x <= nk) {
            const int slot = threadIdx.x == nk ? SNR_KMAX : threadIdx.x;
            float v = 0.f;
            for (unsigned r = 0; r < cs; ++r) v += dsmem_ld(&part[slot], r);
            part[slot] = v;                                    // only CTA 0 reads its own slots again
        }
        cluster_sync_all();                                    // peers may overwrite / exit only after CTA 0 has read
        if (cr == 0 && threadIdx.x < nk)
            snr[id * n + k0 + threadIdx.x] = 4.342944819f * (logf(part[SNR_KMAX] / (float)L + eps) - logf(part[threadIdx.x] / (float)L + eps));   // app/ops.py:188
        __syncthreads();
    }
}

// one cluster per mixture: partial[b] = sum_l (sum_s sep[b,s,l] - mix[b,l])^2
__global__ void __launch_bounds__(512) ae_partial_kernel(const float* __restrict__ sep, const float* __restrict__ mix,
                                                         int S, int64_t L, float* __restrict__ partial) {
    __shared__ float red[32];
    __shared__ float part;
    const unsigned cr = cluster_rank(), cs = cluster_size();
    const int64_t b = blockIdx.x / cs;
    float acc = 0.f;
    for (int64_t x = (int64_t)cr * blockDim.x + threadIdx.x; x < L; x += (int64_t)cs * blockDim.x) {
        float s = -__ldg(mix + b * L + x);
        for (int k = 0; k < S; ++k) s += __ldg(sep + (b * S + k) * L + x);
        acc = fmaf(s, s, acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part = acc;
    cluster_sync_all();
    if (cr == 0 && threadIdx.x == 0) {
        float v = 0.f;
        for (unsigned r = 0; r < cs; ++r) v += dsmem_ld(&part, r);
        partial[b] = v;
    }
    cluster_sync_all();
}

// The per-batch metric vector of the multi-GPU all-reduce (app/parallel.py): vec4 = [sum_b mean_i max_k snr[b,i,k],
// sum_b ae_rows[b] * inv_elems, 0, B] - the batch means of main.py:456-457 and :353-361 are vec4[0] / vec4[3] and
// vec4[1] / vec4[3] after the sum over ranks.  Either input may be null (its slot is written as 0).  One small CTA.
__global__ void __launch_bounds__(256) metric_finalise_kernel(const float* __restrict__ ae_rows, const float* __restrict__ snr,
                                                              int64_t B, int m, int n, float inv_elems, float* __restrict__ vec4) {
    __shared__ float red[32];
    float a = 0.f, q = 0.f;
    for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
        if (ae_rows) a += __ldg(ae_rows + b) * inv_elems;
        if (snr) {
            float sm = 0.f;
            for (int i = 0; i < m; ++i) {
                float best = -INFINITY;
                for (int k = 0; k < n; ++k) best = fmaxf(best, __ldg(snr + (b * m + i) * n + k));
                sm += best;
            }
            q += sm / (float)m;
        }
    }
    a = block_sum(a, red);
    q = block_sum(q, red);
    if (threadIdx.x == 0) { vec4[0] = q; vec4[1] = a; vec4[2] = 0.f; vec4[3] = (float)B; }
}

// per-row min / max -> minmax[2r], minmax[2r+1]; one cluster per row
__global__ void __launch_bounds__(512) minmax_kernel(const float* __restrict__ x, int64_t len, int64_t ld, float* __restrict__ minmax) {
    __shared__ float rlo[16], rhi[16];
    __shared__ float plo, phi;
    const unsigned cr = cluster_rank(), cs = cluster_size();
    const int64_t rowi = blockIdx.x / cs;
    const float* row = x + rowi * ld;
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t i = (int64_t)cr * blockDim.x + threadIdx.x; i < len; i += (int64_t)cs * blockDim.x) { float v = __ldg(row + i); lo = fminf(lo, v); hi = fmaxf(hi, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { rlo[threadIdx.x >> 5] = lo; rhi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        lo = threadIdx.x < (blockDim.x >> 5) ? rlo[threadIdx.x] : INFINITY;
        hi = threadIdx.x < (blockDim.x >> 5) ? rhi[threadIdx.x] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if (threadIdx.x == 0) { plo = lo; phi = hi; }
    }
    cluster_sync_all();
    if (cr == 0 && threadIdx.x == 0) {
        float l2 = INFINITY, h2 = -INFINITY;
        for (unsigned r = 0; r < cs; ++r) { l2 = fminf(l2, dsmem_ld(&plo, r)); h2 = fmaxf(h2, dsmem_ld(&phi, r)); }
        minmax[2 * rowi] = l2; minmax[2 * rowi + 1] = h2;
    }
    cluster_sync_all();
}

// main.py:113-115: d -= min; d *= 32767/(max-min); astype(int16) (truncation), float32 arithmetic
__global__ void __launch_bounds__(256) wav16_kernel(const float* __restrict__ x, int64_t R, int64_t len, int64_t ld,
                                                    const float* __restrict__ minmax, int16_t* __restrict__ pcm) {
    const int64_t total = R * len;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / len, c = i - r * len;
        const float lo = __ldg(minmax + 2 * r), hi = __ldg(minmax + 2 * r + 1);
        const float scale = __fdiv_rn(32767.0f, __fsub_rn(hi, lo));
        const float d = __fmul_rn(__fsub_rn(__ldg(x + r * ld + c), lo), scale);
        pcm[i] = (int16_t)__float2int_rz(d);
    }
}

// ---------------------------------------------------------------------------
// A10 (main.py:328-337): mix[b] = sum_i src[b,i] (+ noise[b]) on packed features, optionally
// followed by to_log_signal (main.py:338) in the same pass.  rows = B*T, a thread handles bins
// k..k+3 of both halves of a row (the to_log gain couples bin k with bin k + N/2).
// ---------------------------------------------------------------------------
template <bool LOG>
__global__ void __launch_bounds__(256) mix_kernel(const float* __restrict__ src, const float* __restrict__ noise,
                                                  float* __restrict__ mix, float* __restrict__ mix_log,
                                                  int64_t B, int n_sig, int64_t T, int N, float eps) {
    const int q = N / 8;
    const int64_t total = B * T * q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / q; const int k4 = (int)(i - r * q);       // r = b*T + t
        const int64_t b = r / T, t = r - b * T;
        float4 re = make_float4(0.f, 0.f, 0.f, 0.f), im = re;
        for (int s = 0; s < n_sig; ++s) {
            const float4* ps = reinterpret_cast<const float4*>(src + ((b * n_sig + s) * T + t) * N) + k4;
            const float4 a = __ldg(ps), c = __ldg(ps + q);
            re.x += a.x; re.y += a.y; re.z += a.z; re.w += a.w;
            im.x += c.x; im.y += c.y; im.z += c.z; im.w += c.w;
        }
        if (noise) {
            const float4* pn = reinterpret_cast<const float4*>(noise + r * N) + k4;
            const float4 a = __ldg(pn), c = __ldg(pn + q);
            re.x += a.x; re.y += a.y; re.z += a.z; re.w += a.w;
            im.x += c.x; im.y += c.y; im.z += c.z; im.w += c.w;
        }
        if (mix) { float4* po = reinterpret_cast<float4*>(mix + r * N) + k4; po[0] = re; po[q] = im; }
        if (LOG) {
            float g;
#define GSS_G(c) g = log_gain(re.c, im.c, eps); re.c *= g; im.c *= g;
            GSS_G(x) GSS_G(y) GSS_G(z) GSS_G(w)
#undef GSS_G
            float4* po = reinterpret_cast<float4*>(mix_log + r * N) + k4; po[0] = re; po[q] = im;
        }
    }
}

// ---------------------------------------------------------------------------
// Backward passes of the element-wise ops the reference back-propagates through
// (to_log_signal / to_exp_signal, app/ops.py:228-251; optimisers main.py:481-484) and of the
// mask multiply.  y = g(a2) * (re, im), a2 = re^2 + im^2:
//   d(re) = g*d(y_re) + 2 g'(a2) re (re d(y_re) + im d(y_im)),   same for im.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void log_gain_d(float a2, float eps, float& g, float& dg) {
    const float l = 0.5f * log1pf(a2), e = a2 + eps, rs = rsqrtf(e);
    g = l * rs;
    dg = 0.5f * rs / (1.0f + a2) - 0.5f * l * rs / e;          // d/da2 [0.5 log1p(a2) (a2+eps)^-1/2]
}
__device__ __forceinline__ void exp_gain_d(float a2, float eps, float& g, float& dg) {
    const float a = sqrtf(a2 + eps), em = expm1f(a);
    g = em / a;
    // d/da [expm1(a)/a] = (a e^a - expm1(a)) / a^2 ;  da/da2 = 1/(2a).  Small a: series (1/2 + a/3 + a^2/8 + a^3/30)
    const float dga = a < 0.05f ? fmaf(a, fmaf(a, fmaf(a, 1.0f / 30.0f, 0.125f), 1.0f / 3.0f), 0.5f)
                                : (a * (em + 1.0f) - em) / (a * a);
    dg = dga / (2.0f * a);
}

template <bool EXP>
__global__ void __launch_bounds__(256) logexp_bwd_kernel(const float* __restrict__ in, const float* __restrict__ gout,
                                                         float* __restrict__ gin, int64_t rows, int N, float eps) {
    const int q = N / 8;
    const int64_t total = rows * q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / q; const int k4 = (int)(i - r * q);
        const float4* pr = reinterpret_cast<const float4*>(in + r * N) + k4;
        const float4* pg = reinterpret_cast<const float4*>(gout + r * N) + k4;
        const float4 re = __ldg(pr), im = __ldg(pr + q), gr = __ldg(pg), gi = __ldg(pg + q);
        float4 dr, di;
#define GSS_B(c) { float g, dg; const float a2 = fmaf(re.c, re.c, im.c * im.c); \
                   if (EXP) exp_gain_d(a2, eps, g, dg); else log_gain_d(a2, eps, g, dg); \
                   const float dot = 2.0f * dg * fmaf(re.c, gr.c, im.c * gi.c); \
                   dr.c = fmaf(g, gr.c, dot * re.c); di.c = fmaf(g, gi.c, dot * im.c); }
        GSS_B(x) GSS_B(y) GSS_B(z) GSS_B(w)
#undef GSS_B
        float4* po = reinterpret_cast<float4*>(gin + r * N) + k4;
        po[0] = dr; po[q] = di;
    }
}

// out[b,s,t] = mask[b,s,t] (x) mix[b,t]:  gmix[b,t] = sum_s mask * gout,  gmask[b,s,t,k] = re*gout_re + im*gout_im
__global__ void __launch_bounds__(256) apply_mask_bwd_kernel(const float* __restrict__ mix, const float* __restrict__ mask,
                                                             const float* __restrict__ gout, float* __restrict__ gmix,
                                                             float* __restrict__ gmask, int64_t B, int S, int64_t T, int N) {
    const int q = N / 8;
    const int64_t total = B * T * q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / q; const int k4 = (int)(i - r * q);       // r = b*T + t
        const int64_t b = r / T, t = r - b * T;
        const float4* pm = reinterpret_cast<const float4*>(mix + r * N) + k4;
        const float4 re = __ldg(pm), im = __ldg(pm + q);
        float4 ar = make_float4(0.f, 0.f, 0.f, 0.f), ai = ar;
        for (int s = 0; s < S; ++s) {
            const int64_t ro = (b * S + s) * T + t;
            const float4* pg = reinterpret_cast<const float4*>(gout + ro * N) + k4;
            const float4 gr = __ldg(pg), gi = __ldg(pg + q);
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask + ro * (N / 2)) + k4);
            ar.x = fmaf(m.x, gr.x, ar.x); ar.y = fmaf(m.y, gr.y, ar.y); ar.z = fmaf(m.z, gr.z, ar.z); ar.w = fmaf(m.w, gr.w, ar.w);
            ai.x = fmaf(m.x, gi.x, ai.x); ai.y = fmaf(m.y, gi.y, ai.y); ai.z = fmaf(m.z, gi.z, ai.z); ai.w = fmaf(m.w, gi.w, ai.w);
            if (gmask) {
                float4 d;
                d.x = fmaf(re.x, gr.x, im.x * gi.x); d.y = fmaf(re.y, gr.y, im.y * gi.y);
                d.z = fmaf(re.z, gr.z, im.z * gi.z); d.w = fmaf(re.w, gr.w, im.w * gi.w);
                reinterpret_cast<float4*>(gmask + ro * (N / 2))[k4] = d;
            }
        }
        if (gmix) { float4* po = reinterpret_cast<float4*>(gmix + r * N) + k4; po[0] = ar; po[q] = ai; }
    }
}

// Batch assembly of the device-resident corpus (app/datasets/wave.py; padding rule of app/datasets/timit.py:47-52): row r of
// out [B, ld] = utterance idx[r] of the flat int16 store (offsets[i] .. offsets[i] + lengths[i]), zero-padded to ld.
__global__ void __launch_bounds__(256) gather_rows_i16_kernel(const int16_t* __restrict__ flat, const int64_t* __restrict__ offsets,
                                                              const int64_t* __restrict__ lengths, const int64_t* __restrict__ idx,
                                                              int64_t B, int64_t ld, int16_t* __restrict__ out) {
    const int64_t total = B * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ld, c = i - r * ld;
        const int64_t u = __ldg(idx + r);
        out[i] = c < __ldg(lengths + u) ? __ldg(flat + __ldg(offsets + u) + c) : (int16_t)0;
    }
}

// int16 PCM -> float32 (same values, no rescale: the reference feeds raw sample values to SciPy, process.py:97)
__global__ void __launch_bounds__(256) i16_to_f32_kernel(const int16_t* __restrict__ in, float* __restrict__ out, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)__ldg(in + i);
}

}  // namespace gss
