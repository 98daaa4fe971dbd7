// gss_stream.cuh - the three streaming kernels of the spectral hot path.
//
// One FFT "team" (Geo<N>::TPF threads, one warp at N = 512) walks a run of
// consecutive frame PAIRS (2q, 2q+1) of one utterance.  Everything that
// overlaps between frames lives in registers:
//
//   analysis  : a ring of raw sample "slots" (slot = L = N/8 samples, v2 per
//               thread); each sample is loaded from global memory once per team
//               although it belongs to N/H frames.
//   synthesis : an overlap-add accumulator of 8+HS slots per output row; the
//               2*HS oldest slots are complete after every pair and are
//               written straight to global memory (no atomics, no shared
//               memory round trip).
//
// Shared memory holds the two register exchanges inside each FFT (gss_fft.cuh),
// private to the team and synchronised with __syncwarp, and - in the fused
// synthesis kernel - a double-buffered stage for the separator's masks that a
// single lane fills with 1-D TMA bulk copies (cp.async.bulk + mbarrier) one
// frame pair ahead of the math.
//
// HS = H / L is the hop in slots: 1 (H = N/8), 2 (H = N/4), 4 (H = N/2, the
// reference's SciPy default, main.py:97).  Positions are "padded" coordinates of
// scipy.signal.stft(boundary='zeros'): padded index pp = p + N/2.
#pragma once
#include <stdint.h>
#include <type_traits>
#include "gss_fft.cuh"

#ifndef GSS_STFT_MAXREG
#define GSS_STFT_MAXREG 168           // 3 CTAs of 4 warps per SM
#endif
#ifndef GSS_ROLL_SOURCES
#define GSS_ROLL_SOURCES 1
#endif

namespace gss {

template <int N_, int HS_>
struct SGeo {
    typedef Geo<N_> G;
    static constexpr int N = N_;
    static constexpr int HS = HS_;
    static constexpr int L = G::L;
    static constexpr int H = HS * L;
    static constexpr int R = 8 / HS;            // frames covering one sample
    static constexpr int RS = 8 + HS;           // slots spanned by a frame pair
    static constexpr int ADV = 2 * HS;          // slots advanced per pair
    static constexpr int KEEP = RS - ADV;       // slots carried to the next pair
    static constexpr int HALO = R / 2;          // pairs to recompute before an owned run (ceil((R-1)/2))
    static constexpr bool CONST_NORM = (R >= 4); // sum_t w^2 is 3R/8 in the interior
    static_assert(HS == 1 || HS == 2 || HS == 4, "hop must be N/8, N/4 or N/2");
};

// periodic Hann (scipy.signal.get_window('hann', N)): 0.5 - 0.5 cos(2 pi i / N)
template <int N>
__device__ __forceinline__ float hann(int i) { return 0.5f - 0.5f * cospif(2.0f * (float)i / (float)N); }

// read-only global load that stays where it is written (prefetches must not be moved down to their use)
__device__ __forceinline__ float ldg_pinned(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------
// sample access: the pair (p, p+1), p even, zero outside [0, n)
// ---------------------------------------------------------------------------
__device__ __forceinline__ v2 load_pair_fast(const float* row, int64_t p, bool al) {
    if (al) return __ldg(reinterpret_cast<const float2*>(row + p));
    return make_float2(__ldg(row + p), __ldg(row + p + 1));
}
__device__ __forceinline__ v2 load_pair_fast(const int16_t* row, int64_t p, bool al) {
    if (al) { short2 s = __ldg(reinterpret_cast<const short2*>(row + p)); return make_float2((float)s.x, (float)s.y); }
    return make_float2((float)__ldg(row + p), (float)__ldg(row + p + 1));
}
template <typename TIn>
__device__ __noinline__ v2 load_pair_edge(const TIn* row, int64_t n, int64_t p) {
    v2 r = make_float2(0.f, 0.f);
    if (p >= 0 && p < n) r.x = (float)__ldg(row + p);
    if (p + 1 >= 0 && p + 1 < n) r.y = (float)__ldg(row + p + 1);
    return r;
}
// slots [sl, sl + CNT) of the padded signal; one range test for the whole group
template <int L, int CNT, typename TIn>
__device__ __forceinline__ void load_slots(const TIn* row, int64_t n, int64_t sl, int j, bool al, v2* dst) {
    const bool inside = sl >= 4 && (sl + CNT - 4) * L <= n;
#pragma unroll
    for (int i = 0; i < CNT; ++i) {
        const int64_t p = (sl + i - 4) * L + 2 * j;
        dst[i] = inside ? load_pair_fast(row, p, al) : load_pair_edge(row, n, p);
    }
}

// ---------------------------------------------------------------------------
// to_log_signal / to_exp_signal gains (app/ops.py:228-251) for one complex bin
// ---------------------------------------------------------------------------
__device__ __forceinline__ float log_gain(float re, float im, float eps) {
    float a2 = fmaf(re, re, im * im);
    return 0.5f * log1pf(a2) * rsqrtf(a2 + eps);
}
__device__ __forceinline__ float exp_gain(float re, float im, float eps) {
    float a = sqrtf(fmaf(re, re, fmaf(im, im, eps)));
    return expm1f(a) / a;
}
// to_exp gain inside the fused iSTFT prologue, two bins at once, valid for a = sqrt(a2e) < 1 (a2e = re^2 +
// im^2 + eps): expm1(a)/a as its Taylor series (next term a^10/11! < 2.6e-8), a through rsqrt.  Chosen per warp
// (one vote); larger magnitudes take the libm path.  The stand-alone to_exp kernel (gss_to_exp) keeps libm.
__device__ __forceinline__ v2 exp_gain2_small(v2 a2e) {
    const v2 a = vmul(a2e, make_float2(rsqrtf(a2e.x), rsqrtf(a2e.y)));
    v2 p = vfma(a, vset(1.0f / 3628800.0f), vset(1.0f / 362880.0f));
    p = vfma(p, a, vset(1.0f / 40320.0f));
    p = vfma(p, a, vset(1.0f / 5040.0f));
    p = vfma(p, a, vset(1.0f / 720.0f));
    p = vfma(p, a, vset(1.0f / 120.0f));
    p = vfma(p, a, vset(1.0f / 24.0f));
    p = vfma(p, a, vset(1.0f / 6.0f));
    p = vfma(p, a, vset(0.5f));
    return vfma(p, a, vset(1.0f));
}
// two bins at once, valid for a2 <= 1 (always true for waveforms in [-1, 1]: |X0| + |X_nyq| <= 1
// and |X_k| <= 1 under scaling='spectrum'):  0.5*log1p(x) = atanh(s), s = x / (2 + x) <= 1/3,
// atanh(s) = s * sum_k s^2k / (2k+1); truncated after k = 6 (next term < 1.5e-8 relative).
__device__ __forceinline__ v2 log_gain2_small(v2 a2, float eps) {
    v2 d = vadd(a2, vset(2.0f));
    v2 s = vmul(a2, make_float2(__frcp_rn(d.x), __frcp_rn(d.y)));
    v2 t = vmul(s, s);
    v2 p = vfma(t, vset(1.0f / 13.0f), vset(1.0f / 11.0f));
    p = vfma(p, t, vset(1.0f / 9.0f));
    p = vfma(p, t, vset(1.0f / 7.0f));
    p = vfma(p, t, vset(1.0f / 5.0f));
    p = vfma(p, t, vset(1.0f / 3.0f));
    p = vfma(p, t, vset(1.0f));
    v2 e = vadd(a2, vset(eps));
    return vmul(vmul(s, p), make_float2(rsqrtf(e.x), rsqrtf(e.y)));
}

// scalar forms for kernels that handle one bin per thread at a time (gss_team.cuh): one warp vote picks the
// polynomial (every lane's a2 <= 1/4, resp. a < 1) or libm; every lane of the warp must call them together
__device__ __forceinline__ float log_gain_warp(float re, float im, float eps) {
    const float a2 = fmaf(re, re, im * im);
    if (__all_sync(0xffffffffu, a2 <= 0.015625f)) {      // ordinary audio levels: three Taylor terms (next term < 1.2e-8 relative)
        float p = fmaf(a2, -0.125f, 1.0f / 6.0f);
        p = fmaf(p, a2, -0.25f);
        p = fmaf(p, a2, 0.5f);
        return a2 * p * rsqrtf(a2 + eps);
    }
    if (__all_sync(0xffffffffu, a2 <= 0.25f)) {
        float p = fmaf(a2, 3.537580770e-02f, -7.272362134e-02f);
        p = fmaf(p, a2, 9.832768570e-02f);
        p = fmaf(p, a2, -1.248603150e-01f);
        p = fmaf(p, a2, 1.666610216e-01f);
        p = fmaf(p, a2, -2.499999137e-01f);
        p = fmaf(p, a2, 4.999999998e-01f);
        return a2 * p * rsqrtf(a2 + eps);
    }
    return 0.5f * log1pf(a2) * rsqrtf(a2 + eps);
}
__device__ __forceinline__ float exp_gain_warp(float re, float im, float eps) {
    const float a2e = fmaf(re, re, fmaf(im, im, eps));
    if (__all_sync(0xffffffffu, a2e < 1.0f)) {
        const float a = a2e * rsqrtf(a2e);
        float p = fmaf(a, 1.0f / 3628800.0f, 1.0f / 362880.0f);
        p = fmaf(p, a, 1.0f / 40320.0f);
        p = fmaf(p, a, 1.0f / 5040.0f);
        p = fmaf(p, a, 1.0f / 720.0f);
        p = fmaf(p, a, 1.0f / 120.0f);
        p = fmaf(p, a, 1.0f / 24.0f);
        p = fmaf(p, a, 1.0f / 6.0f);
        p = fmaf(p, a, 0.5f);
        return fmaf(p, a, 1.0f);
    }
    const float a = sqrtf(a2e);
    return expm1f(a) / a;
}

// per-thread spectrum of a frame pair: bins k = c + L*i, lane x: c = cA, lane y: c = cB
struct PairSpec {
    v2 ar[4], ai[4];   // frame a: "re" slot feat[k], "im" slot feat[N/2 + k]
    v2 br[4], bi[4];   // frame b
};

// Two-for-one split of the forward transform (the window carries the 1/2 and 1/sum(w)).
// Partner of register i (bin k) is bin N-k: lanes crossed in register 7-i, except thread 0
// whose two columns (0 and L/2) pair with themselves (see gss_fft.cuh fft_forward).
template <int N>
__device__ __forceinline__ void split_pair(const cv2 (&a)[8], bool t0, PairSpec& s) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rx = (i == 0) ? 4 : 8 - i;
        v2 pre = a[i].re, pim = a[i].im;
        v2 qre = make_float2(t0 ? a[rx].re.x : a[7 - i].re.y, t0 ? a[7 - i].re.y : a[7 - i].re.x);
        v2 qim = make_float2(t0 ? a[rx].im.x : a[7 - i].im.y, t0 ? a[7 - i].im.y : a[7 - i].im.x);
        s.ar[i] = vadd(pre, qre); s.ai[i] = vsub(pim, qim);
        s.br[i] = vadd(pim, qim); s.bi[i] = vsub(qre, pre);
        if (i == 0) {   // lane x of thread 0 holds (Z[0], Z[N/2]): DC in "re", Nyquist in "im"
            s.ar[0].x = t0 ? 2.f * pre.x : s.ar[0].x; s.ai[0].x = t0 ? 2.f * qre.x : s.ai[0].x;
            s.br[0].x = t0 ? 2.f * pim.x : s.br[0].x; s.bi[0].x = t0 ? 2.f * qim.x : s.bi[0].x;
        }
    }
}

// Hermitian pack of two real-signal spectra into one complex transform input
template <int N>
__device__ __forceinline__ void pack_pair(const v2 (&yar)[4], const v2 (&yai)[4], const v2 (&ybr)[4], const v2 (&ybi)[4],
                                          bool t0, cv2 (&a)[8]) {
    v2 qre[4], qim[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v2 pre = vsub(yar[i], ybi[i]), pim = vadd(yai[i], ybr[i]);
        qre[i] = vadd(yar[i], ybi[i]); qim[i] = vsub(ybr[i], yai[i]);
        if (i == 0) {
            pre.x = t0 ? yar[0].x : pre.x; pim.x = t0 ? ybr[0].x : pim.x;
            qre[0].x = t0 ? yai[0].x : qre[0].x; qim[0].x = t0 ? ybi[0].x : qim[0].x;
        }
        a[i].re = pre; a[i].im = pim;
    }
#pragma unroll
    for (int r = 4; r < 8; ++r) {
        const int ix = (r == 4) ? 0 : 8 - r;      // thread 0, lane x: the pair whose partner lives in register r
        a[r].re = make_float2(t0 ? qre[ix].x : qre[7 - r].y, t0 ? qre[7 - r].y : qre[7 - r].x);
        a[r].im = make_float2(t0 ? qim[ix].x : qim[7 - r].y, t0 ? qim[7 - r].y : qim[7 - r].x);
    }
}

// ---------------------------------------------------------------------------
// overlap-add output: one slot of v2 per thread, trimmed to [N/2, N/2 + (T-1)H)
// ---------------------------------------------------------------------------
template <class SG>
struct OlaOut {
    int64_t T;
    int j;
    bool al;             // rows are 8-byte aligned
    float oscale;        // kScale x (1 / gain the caller's window carries)
    v2 invn[SG::HS];     // oscale / sum_t w^2 per slot residue (only when !CONST_NORM)
    // sum(w)/N = 1/2 and, when it is constant, the interior 1/sum(w^2)
    static constexpr float kScale = SG::CONST_NORM ? 0.5f / (0.375f * SG::R) : 0.5f;

    __device__ __forceinline__ void init(int64_t T_, int j_, bool al_, float win_gain) {
        T = T_; j = j_; al = al_;
        oscale = kScale * win_gain;
        if (!SG::CONST_NORM) {
#pragma unroll
            for (int m = 0; m < SG::HS; ++m) {
                float s0 = 0.f, s1 = 0.f;
                for (int r = 0; r < SG::R; ++r) {
                    float w0 = hann<SG::N>(m * SG::L + 2 * j + r * SG::H);
                    float w1 = hann<SG::N>(m * SG::L + 2 * j + 1 + r * SG::H);
                    s0 += w0 * w0; s1 += w1 * w1;
                }
                invn[m] = make_float2(oscale / s0, oscale / s1);
            }
        }
    }
    // actual sum_t w^2[pp - tH] over the frames that exist (edges of the signal)
    __device__ __noinline__ float norm_at(int64_t sl, int e) const {
        int64_t tlo = sl - 7; tlo = tlo <= 0 ? 0 : (tlo + SG::HS - 1) / SG::HS;
        int64_t thi = sl / SG::HS; if (thi > T - 1) thi = T - 1;
        float s = 0.f;
        for (int64_t t = tlo; t <= thi; ++t) {
            int i = (int)((sl - t * SG::HS) * SG::L) + 2 * j + e;
            float w = hann<SG::N>(i);
            s += w * w;
        }
        return s > 1e-10f ? s : 1.0f;     // scipy.signal.istft: where(norm > 1e-10, norm, 1)
    }
    __device__ __forceinline__ void store(float* row, int64_t sl, v2 v) const {
        float* p = row + (sl - 4) * SG::L + 2 * j;
        if (al) *reinterpret_cast<float2*>(p) = v; else { p[0] = v.x; p[1] = v.y; }
    }
    // all ADV slots starting at `base` are inside the output and covered by R frames
    __device__ __forceinline__ bool interior(int64_t base) const {
        return base >= 8 - SG::HS && base >= 4 && base + SG::ADV <= (T - 1) * SG::HS + (SG::HS < 4 ? SG::HS : 4);
    }
    // v is the raw overlap-add of hann * (unnormalised inverse transform); `scale` = sum(w)/N
    // (= 1/2) and, when constant, the interior 1/sum(w^2)
    __device__ __forceinline__ void write(float* row, int64_t sl, int m /* sl mod HS, static */, v2 v, bool fast) const {
        v = vmul(v, SG::CONST_NORM ? vset(oscale) : invn[m]);
        if (fast) { store(row, sl, v); return; }
        if (sl < 4 || sl >= 4 + (T - 1) * SG::HS) return;
        if (SG::CONST_NORM && (sl <= 7 - SG::HS || sl >= T * SG::HS)) {   // fewer than R frames cover this slot
            const float c = 0.375f * SG::R;
            v.x *= c / norm_at(sl, 0); v.y *= c / norm_at(sl, 1);
        }
        store(row, sl, v);
    }
};

struct ChunkPlan { int ppc; int nchunk; };   // pairs per chunk, chunks per row

// ---------------------------------------------------------------------------
// 1-D TMA bulk copy + mbarrier (one lane issues, the whole team waits)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one lane of the (converged) warp, chosen by the hardware: lets the compiler keep the bulk-copy
// operands in uniform registers instead of looping over the active lanes
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------
// Loop structure shared by the streaming kernels.  A run of pairs [qs, q1) is walked in
// three stretches: slow [qs, qa) - fast [qa, qb) - slow [qb, q1).  The fast body has no
// bounds checks, no edge normalisation and aligned 64-bit global accesses through running
// pointers; the slow body is the fully general one (signal edges, ragged tails, unaligned
// rows, halo pairs).  Both are instances of one generic lambda, so they share all state.
// ---------------------------------------------------------------------------
struct FastTag { static constexpr bool value = true; };
struct SlowTag { static constexpr bool value = false; };

template <int L, int CNT>
__device__ __forceinline__ void load_slots_fast(const float* p, v2* dst) {
#pragma unroll
    for (int i = 0; i < CNT; ++i) dst[i] = __ldg(reinterpret_cast<const float2*>(p + i * L));
}
template <int L, int CNT>
__device__ __forceinline__ void load_slots_fast(const int16_t* p, v2* dst) {
#pragma unroll
    for (int i = 0; i < CNT; ++i) { short2 s = __ldg(reinterpret_cast<const short2*>(p + i * L)); dst[i] = make_float2((float)s.x, (float)s.y); }
}

// last pair index q (inclusive) whose iteration may run the fast body, from the input side:
// iteration q loads the new slots of pair q+1, [2qHS + RS, 2qHS + RS + ADV), all inside [0, n)
template <class SG>
__device__ __forceinline__ int64_t fast_hi_input(int64_t n) {
    const int64_t hi = n / SG::L - SG::RS - SG::ADV + 4;       // 2*q*HS <= hi
    return hi < 0 ? -1 : hi / (2 * SG::HS);
}

// 0.5*log1p(x)/x on [0, 1/4], degree 6 (Chebyshev fit, 4.5e-8 relative in float32 Horner)
__device__ __forceinline__ v2 half_log1p_over_x(v2 x) {
    v2 p = vfma(x, vset(3.537580770e-02f), vset(-7.272362134e-02f));
    p = vfma(p, x, vset(9.832768570e-02f));
    p = vfma(p, x, vset(-1.248603150e-01f));
    p = vfma(p, x, vset(1.666610216e-01f));
    p = vfma(p, x, vset(-2.499999137e-01f));
    p = vfma(p, x, vset(4.999999998e-01f));
    return p;
}
// a2 <= 1/64 (ordinary audio levels: |X|^2 ~ 1e-4): three Taylor terms, next term x^4/10 < 1.2e-8 relative
__device__ __forceinline__ v2 log_gain2_micro(v2 a2, float eps) {
    v2 p = vfma(a2, vset(-0.125f), vset(1.0f / 6.0f));
    p = vfma(p, a2, vset(-0.25f));
    p = vfma(p, a2, vset(0.5f));
    v2 e = vadd(a2, vset(eps));
    return vmul(vmul(a2, p), make_float2(rsqrtf(e.x), rsqrtf(e.y)));
}
// to_log gain for two bins with a2 <= 1/4: a2 * (0.5*log1p(a2)/a2) * rsqrt(a2 + eps)
__device__ __forceinline__ v2 log_gain2_tiny(v2 a2, float eps) {
    v2 e = vadd(a2, vset(eps));
    v2 g = vmul(a2, half_log1p_over_x(a2));
    return vmul(g, make_float2(rsqrtf(e.x), rsqrtf(e.y)));
}

// ---------------------------------------------------------------------------
// STFT: wave [B, ld] -> packed feature [B, T, N] (+ to_log)      A1 + A2 (+ A3)
// ---------------------------------------------------------------------------
template <typename TIn>
struct StftArgs {
    const TIn* wave; float* feat;
    float* feat_lin;                // DUAL kernels (gss_stft_packed_dual): the linear packed spectrum beside the log features
    int64_t B, n, ld, T;
    int npairs, ppc, nchunk;
    int al_in;                      // every row start is aligned for 2-sample vector loads
    float eps;
};

// DUAL (with LOG): the linear spectrum goes to p.feat_lin and its to_log to p.feat - what the separator reads and what
// the feature-fed synthesis (gss_mask_istft_feature) multiplies the masks into, from one transform.
template <int N, int HS, bool LOG, typename TIn, int WARPS, bool DUAL = false>
__global__ void __launch_bounds__(WARPS * 32) __maxnreg__(WARPS <= 4 ? GSS_STFT_MAXREG : ((65536 / (WARPS * 32)) / 8) * 8) stft_kernel(const StftArgs<TIn> p) {
    typedef SGeo<N, HS> SG; typedef Geo<N> G;
    extern __shared__ float4 smem4[];
    float* smf = reinterpret_cast<float*>(smem4);
    constexpr int TEAMS = WARPS * 32 / G::TPF;              // transforms in flight per CTA (two per warp at N = 256)
    const int warp = threadIdx.x / G::TPF, j = threadIdx.x % G::TPF;
    const int64_t item = (int64_t)blockIdx.x * TEAMS + warp;
    if (item >= p.B * p.nchunk) return;
    const int64_t b = item / p.nchunk;
    const int c = (int)(item - b * p.nchunk);
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);
    const bool t0 = j == 0;

    TeamCtx<N> ctx;
    team_init_tab<N>(ctx, j, smf + warp * G::TEAM_FLOATS);
    v2 win[8];
    window_tab<N>(j, 1.0f / (float)N, win);        // 1/sum(w) = 2/N, and the 1/2 of the two-for-one split

    const TIn* row = p.wave + b * p.ld;
    const bool al = p.al_in != 0;
    int64_t base = (int64_t)2 * q0 * HS;            // padded slot of ring[0]

    // fast stretch [q0, qb): loads of the next pair inside the signal, frame b exists, aligned rows
    int qb;
    {
        int64_t qhi = fast_hi_input<SG>(p.n);
        if (qhi > (p.T - 2) / 2) qhi = (p.T - 2) / 2;
        qb = (int)(qhi + 1 < q1 ? qhi + 1 : q1);
        if (!al || qb < q0) qb = q0;
    }

    v2 ring[SG::RS];
    load_slots<SG::L, SG::RS>(row, p.n, base, j, al, ring);
    const TIn* wptr = row + (base + SG::RS - 4) * SG::L + 2 * j;      // slot base + RS: first slot the next pair adds
    float* fptr = p.feat + (b * p.T + 2 * (int64_t)q0) * N;           // feature row of frame 2q
    float* lptr = DUAL ? p.feat_lin + (b * p.T + 2 * (int64_t)q0) * N : nullptr;

    auto step = [&](int q, auto tag) {
        constexpr bool FAST = decltype(tag)::value;
        cv2 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i].re = vmul(ring[i], win[i]); a[i].im = vmul(ring[HS + i], win[i]); }
#pragma unroll
        for (int i = 0; i < SG::KEEP; ++i) ring[i] = ring[i + SG::ADV];
        if (q + 1 < q1) {
            if (FAST) load_slots_fast<SG::L, SG::ADV>(wptr, &ring[SG::KEEP]);
            else load_slots<SG::L, SG::ADV>(row, p.n, base + SG::RS, j, al, &ring[SG::KEEP]);
        }
        fft_forward<N>(ctx, a);
        PairSpec s;
        split_pair<N>(a, t0, s);

        if (DUAL) {
            float* ra = lptr + ctx.cA;
            float* rb = lptr + ctx.cB;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ra[SG::L * i] = s.ar[i].x; ra[N / 2 + SG::L * i] = s.ai[i].x;
                rb[SG::L * i] = s.ar[i].y; rb[N / 2 + SG::L * i] = s.ai[i].y;
            }
            if (FAST || 2 * (int64_t)q + 1 < p.T) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ra[N + SG::L * i] = s.br[i].x; ra[N + N / 2 + SG::L * i] = s.bi[i].x;
                    rb[N + SG::L * i] = s.br[i].y; rb[N + N / 2 + SG::L * i] = s.bi[i].y;
                }
            }
            lptr += 2 * N;
        }
        if (LOG) {
            v2 a2a[4], a2b[4];
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a2a[i] = vfma(s.ar[i], s.ar[i], vmul(s.ai[i], s.ai[i]));
                a2b[i] = vfma(s.br[i], s.br[i], vmul(s.bi[i], s.bi[i]));
                mx = fmaxf(fmaxf(mx, fmaxf(a2a[i].x, a2a[i].y)), fmaxf(a2b[i].x, a2b[i].y));
            }
            if (__all_sync(ctx.mask, mx <= 0.015625f)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v2 ga = log_gain2_micro(a2a[i], p.eps), gb = log_gain2_micro(a2b[i], p.eps);
                    s.ar[i] = vmul(s.ar[i], ga); s.ai[i] = vmul(s.ai[i], ga);
                    s.br[i] = vmul(s.br[i], gb); s.bi[i] = vmul(s.bi[i], gb);
                }
            } else if (__all_sync(ctx.mask, mx <= 0.25f)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v2 ga = log_gain2_tiny(a2a[i], p.eps), gb = log_gain2_tiny(a2b[i], p.eps);
                    s.ar[i] = vmul(s.ar[i], ga); s.ai[i] = vmul(s.ai[i], ga);
                    s.br[i] = vmul(s.br[i], gb); s.bi[i] = vmul(s.bi[i], gb);
                }
            } else if (__all_sync(ctx.mask, mx <= 1.0f)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v2 ga = log_gain2_small(a2a[i], p.eps), gb = log_gain2_small(a2b[i], p.eps);
                    s.ar[i] = vmul(s.ar[i], ga); s.ai[i] = vmul(s.ai[i], ga);
                    s.br[i] = vmul(s.br[i], gb); s.bi[i] = vmul(s.bi[i], gb);
                }
            } else {            // large magnitudes (e.g. int16 PCM): libm-accurate path
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float g;
                    g = log_gain(s.ar[i].x, s.ai[i].x, p.eps); s.ar[i].x *= g; s.ai[i].x *= g;
                    g = log_gain(s.ar[i].y, s.ai[i].y, p.eps); s.ar[i].y *= g; s.ai[i].y *= g;
                    g = log_gain(s.br[i].x, s.bi[i].x, p.eps); s.br[i].x *= g; s.bi[i].x *= g;
                    g = log_gain(s.br[i].y, s.bi[i].y, p.eps); s.br[i].y *= g; s.bi[i].y *= g;
                }
            }
        }
        float* ra = fptr + ctx.cA;
        float* rb = fptr + ctx.cB;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ra[SG::L * i] = s.ar[i].x; ra[N / 2 + SG::L * i] = s.ai[i].x;
            rb[SG::L * i] = s.ar[i].y; rb[N / 2 + SG::L * i] = s.ai[i].y;
        }
        if (FAST || 2 * (int64_t)q + 1 < p.T) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ra[N + SG::L * i] = s.br[i].x; ra[N + N / 2 + SG::L * i] = s.bi[i].x;
                rb[N + SG::L * i] = s.br[i].y; rb[N + N / 2 + SG::L * i] = s.bi[i].y;
            }
        }
        base += SG::ADV;
        wptr += SG::ADV * SG::L;
        fptr += 2 * N;
    };

    int q = q0;
#pragma unroll 1
    for (; q < qb; ++q) step(q, FastTag());
#pragma unroll 1
    for (; q < q1; ++q) step(q, SlowTag());
}

// ---------------------------------------------------------------------------
// iSTFT: packed feature [R, T, N] (+ to_exp) -> wave [R, ld_out]   (A5 +) A6 + A8
// ---------------------------------------------------------------------------
struct IstftArgs {
    const float* feat; float* out;
    int64_t rows, T, ld_out;
    int npairs, ppc, nchunk;
    int al_out;
    float eps;
};

// Without the fused to_exp the kernel is held to 168 registers (3 CTAs of 4 warps per SM): this one IS sensitive to
// occupancy at large batches (B = 1024 x 3 s: 229 -> 197 us); with to_exp the cap would spill (slower), so it is lifted.
template <int N, int HS, bool EXP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) __maxnreg__(WARPS <= 4 && !EXP ? 168 : 255) istft_kernel(const IstftArgs p) {
    typedef SGeo<N, HS> SG; typedef Geo<N> G;
    extern __shared__ float4 smem4[];
    float* smf = reinterpret_cast<float*>(smem4);
    constexpr int TEAMS = WARPS * 32 / G::TPF;
    const int warp = threadIdx.x / G::TPF, j = threadIdx.x % G::TPF;
    const int64_t item = (int64_t)blockIdx.x * TEAMS + warp;
    if (item >= p.rows * p.nchunk) return;
    const int64_t r = item / p.nchunk;
    const int c = (int)(item - r * p.nchunk);
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);
    const int qs = max(q0 - SG::HALO, 0);
    const bool t0 = j == 0;

    TeamCtx<N> ctx;
    team_init_tab<N>(ctx, j, smf + warp * G::TEAM_FLOATS);
    v2 win[8];
    // frame = sum(w) * irfft = (N/2)(1/N) * raw inverse: the 1/2 and 1/sum(w^2) are applied at the store
    window_tab<N>(j, 1.0f, win);
    OlaOut<SG> o;
    o.init(p.T, j, p.al_out != 0, 1.0f);
    float* orow = p.out + r * p.ld_out;

    v2 acc[SG::RS];
#pragma unroll
    for (int i = 0; i < SG::RS; ++i) acc[i] = make_float2(0.f, 0.f);
    int64_t base = (int64_t)2 * qs * HS;
    const float* frow = p.feat + r * p.T * N;

    // features of the next pair are fetched one iteration ahead (volatile loads: the compiler would otherwise sink
    // them past the transform's warp barriers down to their use and expose the whole DRAM latency every pair)
    v2 nar[4], nai[4], nbr[4], nbi[4];
    auto fetch = [&](int q) {
        const int64_t ta = 2 * (int64_t)q;
        const float* ra = frow + ta * N;
        const bool hb = ta + 1 < p.T;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kx = ctx.cA + SG::L * i, ky = ctx.cB + SG::L * i;
            nar[i] = make_float2(ldg_pinned(ra + kx), ldg_pinned(ra + ky));
            nai[i] = make_float2(ldg_pinned(ra + N / 2 + kx), ldg_pinned(ra + N / 2 + ky));
            if (hb) {
                nbr[i] = make_float2(ldg_pinned(ra + N + kx), ldg_pinned(ra + N + ky));
                nbi[i] = make_float2(ldg_pinned(ra + N + N / 2 + kx), ldg_pinned(ra + N + N / 2 + ky));
            } else {
                nbr[i] = make_float2(0.f, 0.f); nbi[i] = make_float2(0.f, 0.f);
            }
        }
    };
    fetch(qs);
    for (int q = qs; q < q1; ++q) {
        v2 yar[4], yai[4], ybr[4], ybi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { yar[i] = nar[i]; yai[i] = nai[i]; ybr[i] = nbr[i]; ybi[i] = nbi[i]; }
        if (q + 1 < q1) fetch(q + 1);
        if (EXP) {
            v2 ea[4], eb[4];
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ea[i] = vfma(yar[i], yar[i], vfma(yai[i], yai[i], vset(p.eps)));
                eb[i] = vfma(ybr[i], ybr[i], vfma(ybi[i], ybi[i], vset(p.eps)));
                mx = fmaxf(fmaxf(mx, fmaxf(ea[i].x, ea[i].y)), fmaxf(eb[i].x, eb[i].y));
            }
            if (__all_sync(ctx.mask, mx < 1.0f)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const v2 ga = exp_gain2_small(ea[i]), gb = exp_gain2_small(eb[i]);
                    yar[i] = vmul(yar[i], ga); yai[i] = vmul(yai[i], ga);
                    ybr[i] = vmul(ybr[i], gb); ybi[i] = vmul(ybi[i], gb);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float g;
                    g = exp_gain(yar[i].x, yai[i].x, p.eps); yar[i].x *= g; yai[i].x *= g;
                    g = exp_gain(yar[i].y, yai[i].y, p.eps); yar[i].y *= g; yai[i].y *= g;
                    g = exp_gain(ybr[i].x, ybi[i].x, p.eps); ybr[i].x *= g; ybi[i].x *= g;
                    g = exp_gain(ybr[i].y, ybi[i].y, p.eps); ybr[i].y *= g; ybi[i].y *= g;
                }
            }
        }
        cv2 a[8];
        pack_pair<N>(yar, yai, ybr, ybi, t0, a);
        fft_inverse<N>(ctx, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] = vfma(a[i].re, win[i], acc[i]);
            acc[HS + i] = vfma(a[i].im, win[i], acc[HS + i]);
        }
        if (q >= q0) {
            const bool fast = o.interior(base);
#pragma unroll
            for (int i = 0; i < SG::ADV; ++i) o.write(orow, base + i, i % HS, acc[i], fast);
        }
#pragma unroll
        for (int i = 0; i < SG::KEEP; ++i) acc[i] = acc[i + SG::ADV];
#pragma unroll
        for (int i = SG::KEEP; i < SG::RS; ++i) acc[i] = make_float2(0.f, 0.f);
        base += SG::ADV;
    }
    if (c == p.nchunk - 1) {
#pragma unroll
        for (int i = 0; i < SG::KEEP; ++i) o.write(orow, base + i, i % HS, acc[i], false);
    }
}

// ---------------------------------------------------------------------------
// fused synthesis: wave [B, ld], mask [B, S, T, N/2] -> out [B*S, ld_out]   A1 + A7 + A8
// The mixture spectrum is recomputed from the waveform (4n bytes) instead of being
// re-read (4TN bytes).  ST sources are carried per pass; S > ST runs as separate
// items (source groups) that each recompute the forward transform.
// Shared memory per team: FFT exchange + 2 stages x ST x 2 frames x N/2 mask gains.
// ---------------------------------------------------------------------------
#ifdef GSS_TIMING
#define GSS_T(k) do { long long t_ = clock64(); if (j == 0) tacc[k] += t_ - tprev; tprev = t_; } while (0)
#else
#define GSS_T(k) do { } while (0)
#endif
struct SynthArgs {
    const float* wave; const float* mask; float* out;
    const float* feat;              // FEAT kernels: linear packed mixture spectrum [B, T, N] instead of `wave`
    int rev;                        // FEAT kernels: walk the work items from the last row to the first
    float* ae_rows;                 // AE kernels: [B] sums of ((sum_s mask_s - 1) * feature)^2 over the row's packed elements
    long long* timing;              // [items][8] cycle accumulators (GSS_TIMING builds only)
    int64_t B, n, ld, T, ld_out;
    int S, ngroups;                 // ngroups = ceil(S / ST)
    int npairs, ppc, nchunk;
    int al_in, al_out;
};

template <int N, int ST, bool FEAT = false>
struct SynthSmem {
    static constexpr int NH = N / 2;
    static constexpr int STAGE_FLOATS = ST * 2 * NH;              // per team, per stage
    static constexpr int FSTAGE_FLOATS = FEAT ? 2 * N : 0;        // FEAT: the pair's two packed feature rows (single stage)
    static constexpr int TEAM_FLOATS = Geo<N>::TEAM_FLOATS + 2 * STAGE_FLOATS + FSTAGE_FLOATS;
    static constexpr int NBAR = 3;                                // two mask stages + the feature stage
    static constexpr size_t bytes(int teams) {
        return sizeof(float) * ((size_t)teams * TEAM_FLOATS) + sizeof(uint64_t) * NBAR * teams;
    }
};

// mask application fused with the Hermitian pack (A7 + the input side of A8): the masked
// spectra  Ya = ga * Xa,  Yb = gb * Xb  of the two frames go straight into the inverse
// transform's input  Ya + i Yb  (registers 0..3: bins k; registers 4..7: bins N-k).
template <int N>
__device__ __forceinline__ void mask_pack_pair(const PairSpec& x, const v2 (&ga)[4], const v2 (&gb)[4], bool t0, cv2 (&a)[8]) {
    v2 qre[4], qim[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const v2 t1 = vmul(x.bi[i], gb[i]), t2 = vmul(x.ai[i], ga[i]);
        v2 pre = vfma(x.ar[i], ga[i], vneg(t1));      // yar - ybi
        qre[i] = vfma(x.ar[i], ga[i], t1);            // yar + ybi
        v2 pim = vfma(x.br[i], gb[i], t2);            // yai + ybr
        qim[i] = vfma(x.br[i], gb[i], vneg(t2));      // ybr - yai
        if (i == 0) {     // thread 0, lane x: (DC, Nyquist) share slot 0 and pair with themselves
            const float yar = x.ar[0].x * ga[0].x, yai = x.ai[0].x * ga[0].x, ybr = x.br[0].x * gb[0].x, ybi = x.bi[0].x * gb[0].x;
            pre.x = t0 ? yar : pre.x; pim.x = t0 ? ybr : pim.x;
            qre[0].x = t0 ? yai : qre[0].x; qim[0].x = t0 ? ybi : qim[0].x;
        }
        a[i].re = pre; a[i].im = pim;
    }
#pragma unroll
    for (int r = 4; r < 8; ++r) {
        const int ix = (r == 4) ? 0 : 8 - r;      // thread 0, lane x: the pair whose partner lives in register r
        a[r].re = make_float2(t0 ? qre[ix].x : qre[7 - r].y, t0 ? qre[7 - r].y : qre[7 - r].x);
        a[r].im = make_float2(t0 ? qim[ix].x : qim[7 - r].y, t0 ? qim[7 - r].y : qim[7 - r].x);
    }
}

// FEAT = true (gss_mask_istft_feature): the mixture's LINEAR packed spectrum [B, T, N] is read back (one 1-D TMA bulk
// copy of the pair's two feature rows per iteration) instead of being recomputed from the waveform: one of the
// 1 + ST transforms per pair goes away (the kernel is bound by issue slots, not by bytes), for 4TN - 4n more bytes.
// AE = true: the auto-encoder loss partial of main.py:353-361 for a mask separator, sum((sum_s separated_s - mixed)^2) =
// sum(((sum_s mask_s - 1) * feature)^2) over the packed elements, accumulated from the registers that already hold the
// pair's spectrum and gains (needs every source in one pass: S <= ST) and added to p.ae_rows[b] once per work item.
template <int N, int HS, int ST, int WARPS, bool FEAT = false, bool AE = false>
__global__ void __launch_bounds__(WARPS * 32) __maxnreg__(WARPS <= 4 ? 255 : ((65536 / (WARPS * 32)) / 8) * 8) mask_istft_kernel(const SynthArgs p) {
    typedef SGeo<N, HS> SG; typedef Geo<N> G; typedef SynthSmem<N, ST, FEAT> SM;
    constexpr int NH = N / 2;
    extern __shared__ float4 smem4[];
    float* smf = reinterpret_cast<float*>(smem4);
    constexpr int TEAMS = WARPS * 32 / G::TPF;
    const int warp = threadIdx.x / G::TPF, j = threadIdx.x % G::TPF;       // team index inside the CTA, lane inside the team
    float* team = smf + warp * SM::TEAM_FLOATS;
    float* stage = team + G::TEAM_FLOATS;                                  // 2 x STAGE_FLOATS, 16-byte aligned
    float* fstage = stage + 2 * SM::STAGE_FLOATS;                          // FEAT: 2 x N floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(smf + TEAMS * SM::TEAM_FLOATS) + SM::NBAR * warp;
    if (j == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_fence_init(); }
    __syncthreads();
    const int64_t per_b = (int64_t)p.ngroups * p.nchunk;
    int64_t item = (int64_t)blockIdx.x * TEAMS + warp;
    if (item >= p.B * per_b) return;
    if (FEAT && p.rev) item = p.B * per_b - 1 - item;
    const int64_t b = item / per_b;
    const int rem = (int)(item - b * per_b);
    const int grp = rem / p.nchunk, c = rem - grp * p.nchunk;
    const int s0 = grp * ST;
    const int ns = min(ST, p.S - s0);                 // sources handled by this team
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);
    const int qs = max(q0 - SG::HALO, 0);
    const bool t0 = j == 0;

    TeamCtx<N> ctx;
    team_init_tab<N>(ctx, j, team);
    v2 win[8];                                     // hann / N: analysis scale; synthesis rescaled at the store
    window_tab<N>(j, FEAT ? 1.0f : 1.0f / (float)N, win);

    OlaOut<SG> o;
    o.init(p.T, j, p.al_out != 0, FEAT ? 1.0f : (float)N);
    float* orow0 = p.out + (b * p.S + s0) * p.ld_out;
    const float* mrow0 = p.mask + ((b * p.S + s0) * p.T) * NH;       // source s: + s*T*NH; frame t: + t*NH
    const int64_t msrc = p.T * NH;                                   // mask stride between sources

    // one lane stages the masks of pair q into stage (q - qs) & 1 (shared-memory addresses kept as
    // 32-bit values so that the elected lane only moves them to uniform registers)
    const uint32_t stage_s = smem_u32(stage), bars_s = smem_u32(bars);
    auto prefetch = [&](int q) {
        // one lane per team: the hardware-elected one when the team is the whole warp, lane 0 of each half otherwise
        if (G::TPF >= 32 ? elect_one() : j == 0) {
            const uint32_t st = (uint32_t)(q - qs) & 1u;
            const int64_t ta = 2 * (int64_t)q;
            const uint32_t bytes = (ta + 1 < p.T ? 2 : 1) * NH * (uint32_t)sizeof(float);
            const uint32_t bar = bars_s + st * 8u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * ns) : "memory");
            const float* src = mrow0 + ta * NH;
            const uint32_t dst = stage_s + st * (uint32_t)(SM::STAGE_FLOATS * sizeof(float));
#pragma unroll
            for (int s = 0; s < ST; ++s)
                if (s < ns)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst + s * (uint32_t)(2 * NH * sizeof(float))), "l"(src + s * msrc), "r"(bytes), "r"(bar) : "memory");
        }
    };

    // FEAT: one lane stages the two packed feature rows of pair q (single stage: refilled as soon as every lane has
    // consumed the previous pair's values, see `source`)
    const float* frow = FEAT ? p.feat + b * p.T * N : nullptr;
    const uint32_t fstage_s = smem_u32(fstage);
    auto prefetch_feat = [&](int q) {
        if (G::TPF >= 32 ? elect_one() : j == 0) {
            const int64_t ta = 2 * (int64_t)q;
            const uint32_t bytes = (ta + 1 < p.T ? 2 : 1) * N * (uint32_t)sizeof(float);
            const uint32_t bar = bars_s + 16u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(fstage_s), "l"(frow + ta * N), "r"(bytes), "r"(bar) : "memory");
        }
    };

    const float* row = FEAT ? nullptr : p.wave + b * p.ld;
    const bool al = FEAT || p.al_in != 0;

    // fast stretch [qa, qb): own pairs whose input slots (incl. the next pair's loads) lie inside the
    // signal, whose ADV output slots are interior (covered by R frames), with both frames present
    int qa, qb;
    {
        int64_t qhi = FEAT ? (int64_t)1 << 40 : fast_hi_input<SG>(p.n);
        const int64_t hi_out = (p.T - 1) * HS + (HS < 4 ? HS : 4) - SG::ADV;     // base + ADV <= (T-1)HS + min(HS,4)
        const int64_t qo = hi_out < 0 ? -1 : hi_out / (2 * HS);
        if (qo < qhi) qhi = qo;
        if (qhi > (p.T - 2) / 2) qhi = (p.T - 2) / 2;
        constexpr int lo_base = (8 - HS) > 4 ? (8 - HS) : 4;                     // base >= max(8 - HS, 4)
        qa = max(q0, (lo_base + 2 * HS - 1) / (2 * HS));
        qb = (int)(qhi + 1 < q1 ? qhi + 1 : q1);
        if (!al || !p.al_out || ns != ST || qb <= qa) { qa = q1; qb = q1; }
    }

    // overlap-add state carried between pairs: KEEP slots per source
    v2 acc[ST][SG::KEEP];
#pragma unroll
    for (int s = 0; s < ST; ++s)
#pragma unroll
        for (int i = 0; i < SG::KEEP; ++i) acc[s][i] = make_float2(0.f, 0.f);

    int64_t base = (int64_t)2 * qs * HS;
    v2 ring[FEAT ? 1 : SG::RS];
    prefetch(qs);
    if constexpr (FEAT) prefetch_feat(qs);
    else load_slots<SG::L, SG::RS>(row, p.n, base, j, al, ring);
    const float* wptr = FEAT ? nullptr : row + (base + SG::RS - 4) * SG::L + 2 * j;      // first slot the next pair adds
    float* optr = orow0 + (base - 4) * SG::L + 2 * j;                   // output slot `base` of source s0
    const float* mA = stage + ctx.cA;
    const float* mB = stage + ctx.cB;

#ifdef GSS_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    v2 ae_acc = make_float2(0.f, 0.f);
    auto step = [&](int q, auto tag) {
        constexpr bool FAST = decltype(tag)::value;
        GSS_T(7);
        PairSpec x;
        const bool hb = FAST || 2 * (int64_t)q + 1 < p.T;
        if constexpr (FEAT) {
            // the other mask stage was last read in iteration q-1; every lane has passed a __syncwarp since
            if (q + 1 < q1) prefetch(q + 1);
            mbar_wait(&bars[2], (uint32_t)(q - qs) & 1u);
            const float* fa = fstage + ctx.cA;
            const float* fb = fstage + ctx.cB;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x.ar[i] = make_float2(fa[SG::L * i], fb[SG::L * i]);
                x.ai[i] = make_float2(fa[NH + SG::L * i], fb[NH + SG::L * i]);
                x.br[i] = hb ? make_float2(fa[N + SG::L * i], fb[N + SG::L * i]) : make_float2(0.f, 0.f);
                x.bi[i] = hb ? make_float2(fa[N + NH + SG::L * i], fb[N + NH + SG::L * i]) : make_float2(0.f, 0.f);
            }
        } else {
            cv2 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].re = vmul(ring[i], win[i]); a[i].im = vmul(ring[HS + i], win[i]); }
#pragma unroll
            for (int i = 0; i < SG::KEEP; ++i) ring[i] = ring[i + SG::ADV];
            if (q + 1 < q1) {
                // the other stage was last read in iteration q-1; every lane has passed a __syncwarp since
                prefetch(q + 1);
                if (FAST) load_slots_fast<SG::L, SG::ADV>(wptr, &ring[SG::KEEP]);
                else load_slots<SG::L, SG::ADV>(row, p.n, base + SG::RS, j, al, &ring[SG::KEEP]);
            }
            fft_forward<N>(ctx, a);
            split_pair<N>(a, t0, x);
        }
        GSS_T(0);
        const int stg = (q - qs) & 1;
        mbar_wait(&bars[stg], ((q - qs) >> 1) & 1);
        GSS_T(1);
        const float* ma = mA + stg * SM::STAGE_FLOATS;
        const float* mb = mB + stg * SM::STAGE_FLOATS;
        const bool own = FAST || q >= q0;
        v2 gsa[AE ? 4 : 1], gsb[AE ? 4 : 1];          // AE: sum of the sources' gains minus one, per bin
        if constexpr (AE) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { gsa[i] = vset(-1.f); gsb[i] = vset(-1.f); }
        }
        // one source: masks -> inverse transform -> overlap-add -> finished slots to global memory.
        // `as` = accumulator set (static), `s` = source index (static when the loop is unrolled)
        auto source = [&](auto as_c, int s) {
            constexpr int as = decltype(as_c)::value;
            v2 ga[4], gb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ga[i] = make_float2(ma[s * 2 * NH + SG::L * i], mb[s * 2 * NH + SG::L * i]);
                gb[i] = hb ? make_float2(ma[s * 2 * NH + NH + SG::L * i], mb[s * 2 * NH + NH + SG::L * i]) : make_float2(0.f, 0.f);
            }
            if constexpr (AE) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { gsa[i] = vadd(gsa[i], ga[i]); gsb[i] = vadd(gsb[i], gb[i]); }
            }
            cv2 a[8];
            mask_pack_pair<N>(x, ga, gb, t0, a);
            GSS_T(2);
            if constexpr (FEAT) {
                // every lane's values of this pair's feature rows have been consumed by the multiplies above
                // (the loads have landed); after the team barrier the single feature stage can be refilled
                if (s == 0 && q + 1 < q1) { team_sync(ctx); prefetch_feat(q + 1); }
            }
            fft_inverse<N>(ctx, a);
            GSS_T(3);
            v2 cur[SG::RS];
#pragma unroll
            for (int i = 0; i < SG::RS; ++i) cur[i] = i < SG::KEEP ? acc[as][i] : make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                cur[i] = vfma(a[i].re, win[i], cur[i]);
                cur[HS + i] = vfma(a[i].im, win[i], cur[HS + i]);
            }
            if (FAST) {
                float* os = optr + s * p.ld_out;
#pragma unroll
                for (int i = 0; i < SG::ADV; ++i)
                    *reinterpret_cast<float2*>(os + i * SG::L) = vmul(cur[i], SG::CONST_NORM ? vset(o.oscale) : o.invn[i % HS]);
            } else if (own) {
                float* orow = orow0 + s * p.ld_out;
#pragma unroll
                for (int i = 0; i < SG::ADV; ++i) o.write(orow, base + i, i % HS, cur[i], false);
            }
#pragma unroll
            for (int i = 0; i < SG::KEEP; ++i) acc[as][i] = cur[i + SG::ADV];
            GSS_T(4);
        };
        if (GSS_ROLL_SOURCES && ST > 1) {
            // rolled: one copy of the inverse transform in the instruction stream (the loop body of
            // three unrolled sources outgrows the 32 KB instruction cache); the accumulator sets
            // rotate through set 0 so that register indexing stays static
#pragma unroll 1
            for (int s = 0; s < ST; ++s) {
                if (FAST || s < ns) source(std::integral_constant<int, 0>(), s);
#pragma unroll
                for (int i = 0; i < SG::KEEP; ++i) {
                    v2 t = acc[0][i];
#pragma unroll
                    for (int k = 0; k + 1 < ST; ++k) acc[k][i] = acc[k + 1][i];
                    acc[ST - 1][i] = t;
                }
            }
        } else {
            if (FAST || 0 < ns) source(std::integral_constant<int, 0>(), 0);
            if (ST > 1 && (FAST || 1 < ns)) source(std::integral_constant<int, (ST > 1 ? 1 : 0)>(), 1);
            if (ST > 2 && (FAST || 2 < ns)) source(std::integral_constant<int, (ST > 2 ? 2 : 0)>(), 2);
            if (ST > 3 && (FAST || 3 < ns)) source(std::integral_constant<int, (ST > 3 ? 3 : 0)>(), 3);
        }
        if constexpr (AE) {
            if (own) {            // halo pairs belong to the neighbouring work item
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const v2 ea = vfma(x.ar[i], x.ar[i], vmul(x.ai[i], x.ai[i]));
                    const v2 da = vmul(gsa[i], gsa[i]);
                    ae_acc = vfma(da, ea, ae_acc);
                    if (hb) {     // an absent frame b has no elements (its gains are not read)
                        const v2 eb = vfma(x.br[i], x.br[i], vmul(x.bi[i], x.bi[i]));
                        const v2 db = vmul(gsb[i], gsb[i]);
                        ae_acc = vfma(db, eb, ae_acc);
                    }
                }
            }
        }
        base += SG::ADV;
        wptr += SG::ADV * SG::L;
        optr += SG::ADV * SG::L;
    };

    int q = qs;
#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
        const int qe = ph == 0 ? qa : q1;
#pragma unroll 1
        for (; q < qe; ++q) step(q, SlowTag());
        if (ph == 0) {
#pragma unroll 1
            for (; q < qb; ++q) step(q, FastTag());
        }
    }
    if (c == p.nchunk - 1) {
#pragma unroll
        for (int s = 0; s < ST; ++s)
            if (s < ns) {
#pragma unroll
                for (int i = 0; i < SG::KEEP; ++i) o.write(orow0 + s * p.ld_out, base + i, i % HS, acc[s][i], false);
            }
    }
    if constexpr (AE) {
        float e = ae_acc.x + ae_acc.y;
#pragma unroll
        for (int d = G::TPF / 2; d >= 1; d >>= 1) e += __shfl_xor_sync(ctx.mask, e, d);
        if (j == 0) atomicAdd(p.ae_rows + b, e);
    }
#ifdef GSS_TIMING
    if (j == 0 && p.timing) {
        tacc[6] = q1 - qs;
        for (int k = 0; k < 8; ++k) p.timing[item * 8 + k] = tacc[k];
    }
#endif
}

}  // namespace gss
