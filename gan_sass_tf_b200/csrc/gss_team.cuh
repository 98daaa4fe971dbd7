// gss_team.cuh - streaming STFT / iSTFT / mask-iSTFT kernels for the FFT sizes the
// register-exchange geometry of gss_fft.cuh does not cover (256, 1024, 2048, 4096; 512 as a
// cross-check path).
//
// Same streaming structure as gss_stream.cuh - a team walks a run of consecutive frame PAIRS of
// one utterance (re = frame 2q, im = frame 2q+1 of one complex N-point transform), the raw-sample
// ring and the overlap-add accumulators live in registers, every sample is read from global
// memory once and every output sample written once, no atomics - but the team is a whole CTA of
// TPT = N/R0 threads and the transform is a three-pass Stockham autosort FFT  N = R0 x RM x R0
// (radix 8 / 16 outside, 4 / 8 / 16 in the middle), natural order in and out, so the two-for-one
// split, the mask multiply and the Hermitian pack address bins k and N-k directly and masks /
// features are read and written fully coalesced.  Thread t owns sample t of every slot of TPT
// samples - exactly the inputs of its first-pass butterfly (t + r*TPT) and the outputs of its
// last-pass butterfly - so the first pass of a forward transform reads the windowed samples
// straight from the register ring and the last pass of an inverse transform feeds the
// overlap-add accumulators in registers: the data makes two shared-memory round trips per
// transform instead of four.  Per-thread twiddles live in registers as w^1, w^2, w^4, w^8 and are
// expanded on the fly; buffers are padded one element per R0 against the first pass's scatter.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "gss_stream.cuh"

namespace gss {
namespace team {

__device__ __forceinline__ float2 cadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul2(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ float2 cconj2(float2 a) { return make_float2(a.x, -a.y); }
// a * (-i) forward, a * (+i) inverse
template <bool INV> __device__ __forceinline__ float2 rot90(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
// a * w (forward) / a * conj(w) (inverse), w a forward twiddle
template <bool INV> __device__ __forceinline__ float2 twmul(float2 a, float2 w) { return cmul2(a, INV ? cconj2(w) : w); }

template <bool INV>
__device__ __forceinline__ void dft2(float2& a, float2& b) { float2 t = a; a = cadd2(t, b); b = csub2(t, b); }

// in-place 4-point DFT, natural order
template <bool INV>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 d0 = cadd2(a0, a2), d1 = csub2(a0, a2), d2 = cadd2(a1, a3), d3 = rot90<INV>(csub2(a1, a3));
    a0 = cadd2(d0, d2); a2 = csub2(d0, d2);
    a1 = cadd2(d1, d3); a3 = csub2(d1, d3);
}

// in-place 8-point DFT, natural order: n = 4*n1 + n2 (n1 < 2, n2 < 4), k = k1 + 2*k2
template <bool INV>
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
    const float s = 0.70710678118654752440f;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft2<INV>(a[n2], a[4 + n2]);          // over n1: y[n2][k1] in a[4*k1 + n2]
    // twiddle W8^(n2*k1), k1 = 1
    a[5] = twmul<INV>(a[5], make_float2(s, -s));
    a[6] = rot90<INV>(a[6]);
    a[7] = twmul<INV>(a[7], make_float2(-s, -s));
    dft4<INV>(a[0], a[1], a[2], a[3]);                                    // k1 = 0: X[2*k2]   in a[k2]
    dft4<INV>(a[4], a[5], a[6], a[7]);                                    // k1 = 1: X[1+2*k2] in a[4+k2]
    float2 b[8];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) { b[2 * k2] = a[k2]; b[2 * k2 + 1] = a[4 + k2]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = b[i];
}

// in-place 16-point DFT, natural order: n = 4*n1 + n2, k = k1 + 4*k2
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&a)[16]) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(a[n2], a[4 + n2], a[8 + n2], a[12 + n2]);   // y[n2][k1] in a[4*k1 + n2]
    // twiddles W16^(n2*k1)
    a[5] = twmul<INV>(a[5], make_float2(c1, -s1));      // 1
    a[6] = twmul<INV>(a[6], make_float2(h, -h));        // 2
    a[7] = twmul<INV>(a[7], make_float2(s1, -c1));      // 3
    a[9] = twmul<INV>(a[9], make_float2(h, -h));        // 2
    a[10] = rot90<INV>(a[10]);                          // 4
    a[11] = twmul<INV>(a[11], make_float2(-h, -h));     // 6
    a[13] = twmul<INV>(a[13], make_float2(s1, -c1));    // 3
    a[14] = twmul<INV>(a[14], make_float2(-h, -h));     // 6
    a[15] = twmul<INV>(a[15], make_float2(-c1, s1));    // 9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);  // X[k1 + 4*k2] in a[4*k1 + k2]
    float2 b[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) b[k1 + 4 * k2] = a[4 * k1 + k2];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = b[i];
}

template <int R, bool INV>
__device__ __forceinline__ void dftR(float2 (&a)[R]) {
    if constexpr (R == 16) dft16<INV>(a);
    else if constexpr (R == 8) dft8<INV>(a);
    else if constexpr (R == 4) dft4<INV>(a[0], a[1], a[2], a[3]);
    else dft2<INV>(a[0], a[1]);
}

// ---------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------
// N = R0 * RM * R0; TPT = N / R0 threads; MINB*: resident CTAs per SM the register allocation is held to, per kernel
// (chosen so that nothing spills: the radix-16 butterflies with their expanded twiddles are register-hungry, and a
// spilling synthesis kernel at 12 warps/SM measured 15 % slower than a clean one at 8)
template <int N_> struct Plan;
// MINB_FEAT / FEAT_ST1: the feature-fed synthesis.  At 1024 / 2048 it runs ONE source per work item (the sources of a mixture
// become separate items that re-read the pair's features, mostly from L2) at 144-168 registers and 6 / 3 resident CTAs per SM:
// S = 3 at B = 1024 x 3 s 909 -> 778 us and 1042 -> 827 us against three sources per item at 4 / 2 CTAs
// (profiles/r2_team_st1*.txt); at 4096 (256 threads per CTA) the register cap that a second CTA needs spills, and one
// source per item at one CTA is slower than three (1069 against 1005 us), so it keeps the grouping of the waveform-fed kernel.
template <> struct Plan<256>  { static constexpr int R0 = 8,  RM = 4,  MINB = 16, MINB_ISTFT = 16, MINB_SYNTH = 16, MINB_FEAT = 16; static constexpr bool FEAT_ST1 = false; };
template <> struct Plan<512>  { static constexpr int R0 = 8,  RM = 8,  MINB = 8,  MINB_ISTFT = 8,  MINB_SYNTH = 8,  MINB_FEAT = 8;  static constexpr bool FEAT_ST1 = false; };
template <> struct Plan<1024> { static constexpr int R0 = 16, RM = 4,  MINB = 8,  MINB_ISTFT = 6,  MINB_SYNTH = 4,  MINB_FEAT = 6;  static constexpr bool FEAT_ST1 = true; };
template <> struct Plan<2048> { static constexpr int R0 = 16, RM = 8,  MINB = 4,  MINB_ISTFT = 3,  MINB_SYNTH = 2,  MINB_FEAT = 3;  static constexpr bool FEAT_ST1 = true; };
template <> struct Plan<4096> { static constexpr int R0 = 16, RM = 16, MINB = 2,  MINB_ISTFT = 1,  MINB_SYNTH = 1,  MINB_FEAT = 1;  static constexpr bool FEAT_ST1 = false; };

template <int N_, int HS_>
struct TGeo {
    typedef Plan<N_> P;
    static constexpr int N = N_, HS = HS_, R0 = P::R0, RM = P::RM;
    static constexpr int TPT = N / R0;             // threads per team = samples per slot
    static constexpr int SLOT = TPT;
    static constexpr int FS = R0;                  // slots per frame
    static constexpr int H = HS * SLOT;
    static constexpr int R = FS / HS;              // frames covering one sample
    static constexpr int RS = FS + HS;             // slots spanned by a frame pair
    static constexpr int ADV = 2 * HS;
    static constexpr int KEEP = RS - ADV;
    static constexpr int HALO = R / 2;             // pairs to recompute before an owned run: ceil((R-1)/2)
    static constexpr bool CONST_NORM = (R >= 4);
    static constexpr int PADN = N + N / R0;        // padded complex elements per buffer
    static constexpr int NMID = N / RM;            // middle-pass butterflies
    static constexpr int BPT = (NMID + TPT - 1) / TPT;   // ... per thread (1, 2 or 4)
    static_assert(N == R0 * RM * R0, "plan must factor N");
    static_assert(FS % HS == 0 && R >= 2, "hop must divide the frame into >= 2 parts");
    static_assert(R0 % RM == 0 || RM % R0 == 0, "middle butterflies of one thread must share their twiddles");
};
template <int R0> __device__ __forceinline__ int pad(int i) { return i + i / R0; }

// read-only global load that stays where it is written: the compiler is free to sink a plain __ldg past
// the CTA barriers down to its use (saving registers, exposing the whole DRAM latency); prefetches must not move
__device__ __forceinline__ float ldg_here(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// powers w^1 .. w^(R-1) from the stored w^1, w^2, w^4, w^8
struct PassTw { float2 w[4]; };
template <int R>
__device__ __forceinline__ void expand(const PassTw& tw, float2 (&w)[R]) {
    w[1] = tw.w[0];
    if (R > 2) { w[2] = tw.w[1]; w[3] = cmul2(w[1], w[2]); }
    if (R > 4) { w[4] = tw.w[2]; w[5] = cmul2(w[1], w[4]); w[6] = cmul2(w[2], w[4]); w[7] = cmul2(w[3], w[4]); }
    if (R > 8) {
        w[8] = tw.w[3];
#pragma unroll
        for (int r = 9; r < R; ++r) w[r] = cmul2(w[r - 8], w[8]);
    }
}
// W_period^(k * 2^i), i = 0..3
__device__ __forceinline__ void init_pass_tw(PassTw& p, int k, int period) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s, c;
        sincospif(-2.0f * (float)(k << i) / (float)period, &s, &c);
        p.w[i] = make_float2(c, s);
    }
}
template <class G>
struct TeamTw { PassTw mid, last; };
template <class G>
__device__ __forceinline__ void init_tw(TeamTw<G>& t, int j) {
    init_pass_tw(t.mid, j % G::R0, G::R0 * G::RM);       // middle pass: NS = R0, k = j mod R0 (same for j + TPT)
    init_pass_tw(t.last, j, G::N);                       // last pass:   NS = N/R0 = TPT, k = j
}

// Stockham passes (DIT):  v[r] = in[j + r N/R] * w^r ;  V = DFT_R(v) ;  out[(j - k) R + k + r NS] = V[r],  k = j mod NS
// first pass (NS = 1): inputs given in registers (the windowed frame pair, or loaded by the caller)
template <class G, bool INV>
__device__ __forceinline__ void first_pass(float2 (&v)[G::R0], float2* __restrict__ out, int j) {
    dftR<G::R0, INV>(v);
#pragma unroll
    for (int r = 0; r < G::R0; ++r) out[pad<G::R0>(j * G::R0 + r)] = v[r];
}
template <class G>
__device__ __forceinline__ void load_first(const float2* __restrict__ in, float2 (&v)[G::R0], int j) {
#pragma unroll
    for (int r = 0; r < G::R0; ++r) v[r] = in[pad<G::R0>(j + r * G::TPT)];
}
// middle pass (NS = R0): NMID butterflies, BPT per thread
template <class G, bool INV>
__device__ __forceinline__ void mid_pass(const float2* __restrict__ in, float2* __restrict__ out, const PassTw& tw, int j) {
    constexpr int R = G::RM;
    float2 w[R];
    expand<R>(tw, w);
    const int k = j % G::R0;
#pragma unroll
    for (int b = 0; b < G::BPT; ++b) {
        const int jb = j + b * G::TPT;
        if (jb < G::NMID) {
            float2 v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = in[pad<G::R0>(jb + r * G::NMID)];
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = twmul<INV>(v[r], w[r]);
            dftR<R, INV>(v);
            const int o = (jb - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) out[pad<G::R0>(o + r * G::R0)] = v[r];
        }
    }
}
// last pass (NS = TPT, k = j): outputs are elements j + r*TPT - left in registers
template <class G, bool INV>
__device__ __forceinline__ void last_pass(const float2* __restrict__ in, float2 (&v)[G::R0], const PassTw& tw, int j) {
    constexpr int R = G::R0;
    float2 w[R];
    expand<R>(tw, w);
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in[pad<R>(j + r * G::TPT)];
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = twmul<INV>(v[r], w[r]);
    dftR<R, INV>(v);
}
template <class G>
__device__ __forceinline__ void store_natural(const float2 (&v)[G::R0], float2* __restrict__ out, int j) {
#pragma unroll
    for (int r = 0; r < G::R0; ++r) out[pad<G::R0>(j + r * G::TPT)] = v[r];
}

// forward transform of the frame pair held in the ring: result (natural order) in A; B is scratch
template <class G>
__device__ __forceinline__ void forward_from_ring(const float (&ring)[G::RS], const float (&win)[G::FS], float2* A, float2* B,
                                                  const TeamTw<G>& tw, int j) {
    float2 v[G::R0];
#pragma unroll
    for (int r = 0; r < G::R0; ++r) v[r] = make_float2(ring[r] * win[r], ring[G::HS + r] * win[r]);
    first_pass<G, false>(v, A, j);
    __syncthreads();
    mid_pass<G, false>(A, B, tw.mid, j);
    __syncthreads();
    last_pass<G, false>(B, v, tw.last, j);
    store_natural<G>(v, A, j);
    __syncthreads();
}
// inverse transform of Y (natural order, in Y; W is scratch): element j + r*TPT = (frame a, frame b) sample in v[r]
template <class G>
__device__ __forceinline__ void inverse_to_regs(float2* Y, float2* W, float2 (&v)[G::R0], const TeamTw<G>& tw, int j) {
    load_first<G>(Y, v, j);
    first_pass<G, true>(v, W, j);
    __syncthreads();
    mid_pass<G, true>(W, Y, tw.mid, j);
    __syncthreads();
    last_pass<G, true>(Y, v, tw.last, j);
}

// ---------------------------------------------------------------------------
// sample access and overlap-add output (thread t owns sample t of every slot)
// ---------------------------------------------------------------------------
template <class G, int CNT, typename TIn>
__device__ __forceinline__ void load_slots(const TIn* row, int64_t n, int64_t sl, int t, float* dst) {
    const bool inside = sl >= G::FS / 2 && (sl + CNT - G::FS / 2) * G::SLOT <= n;
#pragma unroll
    for (int i = 0; i < CNT; ++i) {
        const int64_t p = (sl + i - G::FS / 2) * G::SLOT + t;
        dst[i] = (inside || (p >= 0 && p < n)) ? (float)__ldg(row + p) : 0.f;
    }
}

template <class G>
struct Ola {
    int64_t T;
    int t;
    float oscale;
    float invn[G::HS];
    static constexpr float kScale = G::CONST_NORM ? 0.5f / (0.375f * G::R) : 0.5f;

    __device__ __forceinline__ void init(int64_t T_, int t_, float win_gain) {
        T = T_; t = t_;
        oscale = kScale * win_gain;
        if (!G::CONST_NORM) {
#pragma unroll
            for (int m = 0; m < G::HS; ++m) {
                float s0 = 0.f;
                for (int r = 0; r < G::R; ++r) { float w0 = hann<G::N>(m * G::SLOT + t + r * G::H); s0 += w0 * w0; }
                invn[m] = oscale / s0;
            }
        }
    }
    // actual sum_f w^2 over the frames that exist (edges of the signal)
    __device__ __noinline__ float norm_at(int64_t sl) const {
        int64_t tlo = sl - (G::FS - 1); tlo = tlo <= 0 ? 0 : (tlo + G::HS - 1) / G::HS;
        int64_t thi = sl / G::HS; if (thi > T - 1) thi = T - 1;
        float s = 0.f;
        for (int64_t f = tlo; f <= thi; ++f) {
            float w = hann<G::N>((int)((sl - f * G::HS) * G::SLOT) + t);
            s += w * w;
        }
        return s > 1e-10f ? s : 1.0f;     // scipy.signal.istft: where(norm > 1e-10, norm, 1)
    }
    // all ADV slots starting at `base` are inside the output and covered by R frames: plain scaled stores
    __device__ __forceinline__ bool interior(int64_t base) const {
        return base >= G::FS - G::HS && base >= G::FS / 2 && base + G::ADV <= (T - 1) * G::HS + (G::HS < G::FS / 2 ? G::HS : G::FS / 2);
    }
    __device__ __forceinline__ void write_fast(float* row, int64_t sl, int m, float v) const {
        row[(sl - G::FS / 2) * G::SLOT + t] = v * (G::CONST_NORM ? oscale : invn[m]);
    }
    __device__ __forceinline__ void write(float* row, int64_t sl, int m, float v) const {
        v *= G::CONST_NORM ? oscale : invn[m];
        if (sl < G::FS / 2 || sl >= G::FS / 2 + (T - 1) * G::HS) return;
        if (G::CONST_NORM && (sl <= G::FS - 1 - G::HS || sl >= T * G::HS)) v *= (0.375f * G::R) / norm_at(sl);
        row[(sl - G::FS / 2) * G::SLOT + t] = v;
    }
};

template <class G>
__device__ __forceinline__ void make_window(int t, float scale, float (&w)[G::FS]) {
#pragma unroll
    for (int i = 0; i < G::FS; ++i) w[i] = scale * hann<G::N>(t + G::SLOT * i);
}

// two-for-one: spectra of frames a and b at bin k from Z[k], Z[N-k] (the window carries the 1/2)
__device__ __forceinline__ void split_bin(float2 zk, float2 zn, float2& A, float2& B) {
    A = make_float2(zk.x + zn.x, zk.y - zn.y);            // Z[k] + conj Z[N-k]
    B = make_float2(zk.y + zn.y, zn.x - zk.x);            // (Z[k] - conj Z[N-k]) / i
}
// spectra Ya, Yb of two real frames at bin k (k in [0, N/2]) -> Y[k], Y[N-k] of the complex transform
template <int R0>
__device__ __forceinline__ void pack_bin(float2 Ya, float2 Yb, float2* Y, int k, int N) {
    Y[pad<R0>(k)] = make_float2(Ya.x - Yb.y, Ya.y + Yb.x);                       // Ya + i Yb
    Y[pad<R0>((N - k) & (N - 1))] = make_float2(Ya.x + Yb.y, Yb.x - Ya.y);       // conj(Ya) + i conj(Yb)
}
// inverse-transform outputs of this thread -> overlap-add accumulators
template <class G>
__device__ __forceinline__ void ola_accumulate(const float2 (&v)[G::R0], const float (&win)[G::FS], float (&cur)[G::RS]) {
#pragma unroll
    for (int r = 0; r < G::FS; ++r) {
        cur[r] = fmaf(v[r].x, win[r], cur[r]);
        cur[G::HS + r] = fmaf(v[r].y, win[r], cur[G::HS + r]);
    }
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
template <typename TIn>
struct StftArgs {
    const TIn* wave; float* feat;
    float* feat_lin;                // not null (gss_stft_packed_dual): the linear spectrum goes here, its to_log to `feat`
    int64_t B, n, ld, T;
    int npairs, ppc, nchunk;
    int log; float eps;
};

template <int N, int HS, typename TIn>
__global__ void __launch_bounds__(TGeo<N, HS>::TPT, Plan<N>::MINB) stft_kernel(const StftArgs<TIn> p) {
    typedef TGeo<N, HS> G;
    constexpr int R0 = G::R0;
    extern __shared__ float4 smem4[];
    float2* A = reinterpret_cast<float2*>(smem4);
    float2* Bf = A + G::PADN;
    const int t = threadIdx.x;
    const int64_t item = blockIdx.x;
    const int64_t b = item / p.nchunk;
    const int c = (int)(item - b * p.nchunk);
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);

    TeamTw<G> tw; init_tw<G>(tw, t);
    float win[G::FS];
    make_window<G>(t, 1.0f / (float)N, win);       // 1/sum(w) = 2/N, and the 1/2 of the two-for-one split
    const TIn* row = p.wave + b * p.ld;
    int64_t base = (int64_t)2 * q0 * HS;
    float ring[G::RS];
    load_slots<G, G::RS>(row, p.n, base, t, ring);

    for (int q = q0; q < q1; ++q) {
        forward_from_ring<G>(ring, win, A, Bf, tw, t);
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) ring[i] = ring[i + G::ADV];
        if (q + 1 < q1) load_slots<G, G::ADV>(row, p.n, base + G::RS, t, &ring[G::KEEP]);
        const int64_t ta = 2 * (int64_t)q;
        float* fa = p.feat + (b * p.T + ta) * N;
        const bool hb = ta + 1 < p.T;
        for (int k = t; k < N / 2; k += G::TPT) {
            float2 Sa, Sb;
            split_bin(A[pad<R0>(k)], A[pad<R0>((N - k) & (N - 1))], Sa, Sb);
            if (k == 0) {               // slot 0 carries (DC, Nyquist) (app/utils.py:22-26)
                float2 An, Bn;
                const float2 zh = A[pad<R0>(N / 2)];
                split_bin(zh, zh, An, Bn);
                Sa.y = An.x; Sb.y = Bn.x;
            }
            if (p.feat_lin) {
                float* la = p.feat_lin + (b * p.T + ta) * N;
                la[k] = Sa.x; la[N / 2 + k] = Sa.y;
                if (hb) { la[N + k] = Sb.x; la[N + N / 2 + k] = Sb.y; }
            }
            if (p.log) {
                float g = log_gain_warp(Sa.x, Sa.y, p.eps); Sa.x *= g; Sa.y *= g;
                g = log_gain_warp(Sb.x, Sb.y, p.eps); Sb.x *= g; Sb.y *= g;
            }
            fa[k] = Sa.x; fa[N / 2 + k] = Sa.y;
            if (hb) { fa[N + k] = Sb.x; fa[N + N / 2 + k] = Sb.y; }
        }
        base += G::ADV;
        __syncthreads();                            // A is rewritten by the next pair's first pass
    }
}

struct IstftArgs {
    const float* feat; float* out;
    int64_t rows, T, ld_out;
    int npairs, ppc, nchunk;
    int exp; float eps;
};

template <int N, int HS>
__global__ void __launch_bounds__(TGeo<N, HS>::TPT, Plan<N>::MINB_ISTFT) istft_kernel(const IstftArgs p) {
    typedef TGeo<N, HS> G;
    constexpr int R0 = G::R0;
    extern __shared__ float4 smem4[];
    float2* A = reinterpret_cast<float2*>(smem4);
    float2* Bf = A + G::PADN;
    const int t = threadIdx.x;
    const int64_t item = blockIdx.x;
    const int64_t r = item / p.nchunk;
    const int c = (int)(item - r * p.nchunk);
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);
    const int qs = max(q0 - G::HALO, 0);

    TeamTw<G> tw; init_tw<G>(tw, t);
    float win[G::FS];
    // frame = sum(w) * irfft = (N/2)(1/N) * raw inverse: the 1/2 and 1/sum(w^2) are applied at the store
    make_window<G>(t, 1.0f, win);
    Ola<G> o;
    o.init(p.T, t, 1.0f);
    float* orow = p.out + r * p.ld_out;
    const float* frow = p.feat + r * p.T * N;
    float acc[G::KEEP];
#pragma unroll
    for (int i = 0; i < G::KEEP; ++i) acc[i] = 0.f;
    int64_t base = (int64_t)2 * qs * HS;

    // packed features of this thread's bins k = t + i*TPT (slot 0 also feeds bin N/2), fetched one pair ahead
    constexpr int KI = (N / 2) / G::TPT;
    float2 fya[KI], fyb[KI], nya[KI], nyb[KI];
    auto fetch = [&](int q, float2 (&ya)[KI], float2 (&yb)[KI]) {
        const int64_t ta = 2 * (int64_t)q;
        const float* fa = frow + ta * N + t;
        const bool hb = ta + 1 < p.T;
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            ya[i] = make_float2(ldg_here(fa + i * G::TPT), ldg_here(fa + N / 2 + i * G::TPT));
            yb[i] = hb ? make_float2(ldg_here(fa + N + i * G::TPT), ldg_here(fa + N + N / 2 + i * G::TPT)) : make_float2(0.f, 0.f);
        }
    };
    fetch(qs, fya, fyb);
    for (int q = qs; q < q1; ++q) {
        if (q + 1 < q1) fetch(q + 1, nya, nyb);
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            const int k = t + i * G::TPT;
            float2 Ya = fya[i], Yb = fyb[i];
            if (p.exp) {          // to_exp pairs slot 0 = (DC, Nyquist) like every other bin pair (ops.py:247)
                float g = exp_gain_warp(Ya.x, Ya.y, p.eps); Ya.x *= g; Ya.y *= g;
                g = exp_gain_warp(Yb.x, Yb.y, p.eps); Yb.x *= g; Yb.y *= g;
            }
            if (k == 0) {         // bin 0 -> (f[0], 0), bin N/2 -> (f[N/2], 0)
                pack_bin<R0>(make_float2(Ya.y, 0.f), make_float2(Yb.y, 0.f), A, N / 2, N);
                Ya.y = 0.f; Yb.y = 0.f;
            }
            pack_bin<R0>(Ya, Yb, A, k, N);
        }
        __syncthreads();
        float2 v[R0];
        inverse_to_regs<G>(A, Bf, v, tw, t);
        float cur[G::RS];
#pragma unroll
        for (int i = 0; i < G::RS; ++i) cur[i] = i < G::KEEP ? acc[i] : 0.f;
        ola_accumulate<G>(v, win, cur);
        if (q >= q0) {
            if (o.interior(base)) {
#pragma unroll
                for (int i = 0; i < G::ADV; ++i) o.write_fast(orow, base + i, i % HS, cur[i]);
            } else {
#pragma unroll
                for (int i = 0; i < G::ADV; ++i) o.write(orow, base + i, i % HS, cur[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) acc[i] = cur[i + G::ADV];
#pragma unroll
        for (int i = 0; i < KI; ++i) { fya[i] = nya[i]; fyb[i] = nyb[i]; }      // the next pair's features have had a whole transform to arrive
        base += G::ADV;
        __syncthreads();                            // A (read by the last pass) is rewritten by the next pair's pack
    }
    if (c == p.nchunk - 1) {
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) o.write(orow, base + i, i % HS, acc[i]);
    }
}

struct SynthArgs {
    const float* wave; const float* mask; float* out;
    int64_t B, n, ld, T, ld_out;
    int S, ngroups;
    int npairs, ppc, nchunk;
};

template <int N, int HS, int ST>
__global__ void __launch_bounds__(TGeo<N, HS>::TPT, Plan<N>::MINB_SYNTH) mask_istft_kernel(const SynthArgs p) {
    typedef TGeo<N, HS> G;
    constexpr int NH = N / 2, R0 = G::R0;
    extern __shared__ float4 smem4[];
    float2* X = reinterpret_cast<float2*>(smem4);      // mixture spectrum of the pair, kept for every source
    float2* Y = X + G::PADN;
    float2* W = Y + G::PADN;
    const int t = threadIdx.x;
    const int64_t item = blockIdx.x;
    const int64_t per_b = (int64_t)p.ngroups * p.nchunk;
    const int64_t b = item / per_b;
    const int rem = (int)(item - b * per_b);
    const int grp = rem / p.nchunk, c = rem - grp * p.nchunk;
    const int s0 = grp * ST;
    const int ns = min(ST, p.S - s0);
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);
    const int qs = max(q0 - G::HALO, 0);

    TeamTw<G> tw; init_tw<G>(tw, t);
    float win[G::FS];
    make_window<G>(t, 1.0f / (float)N, win);       // hann / N: analysis scale; synthesis rescaled at the store
    Ola<G> o;
    o.init(p.T, t, (float)N);
    float* orow0 = p.out + (b * p.S + s0) * p.ld_out;
    const float* mrow0 = p.mask + ((b * p.S + s0) * p.T) * NH;
    const int64_t msrc = p.T * NH;
    const float* row = p.wave + b * p.ld;

    float acc[ST][G::KEEP];
#pragma unroll
    for (int s = 0; s < ST; ++s)
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) acc[s][i] = 0.f;
    int64_t base = (int64_t)2 * qs * HS;
    float ring[G::RS];
    load_slots<G, G::RS>(row, p.n, base, t, ring);

    // gains of one source for this thread's bins (k = t + i*TPT, plus k = N/2 which shares gain 0), fetched one source
    // ahead of their use - across pairs too: source 0 of the next pair is requested while the last source of this one
    // is transformed, so no pair starts by waiting for DRAM
    constexpr int KI = NH / G::TPT;
    auto fetch = [&](int q_, int s, float (&a)[KI], float (&bq)[KI]) {
        const int64_t ta_ = 2 * (int64_t)q_;
        const bool hb_ = ta_ + 1 < p.T;
        const float* ma = mrow0 + s * msrc + ta_ * NH + t;
#pragma unroll
        for (int i = 0; i < KI; ++i) { a[i] = ldg_here(ma + i * G::TPT); bq[i] = hb_ ? ldg_here(ma + NH + i * G::TPT) : 0.f; }
    };
    float ga[KI], gb[KI], ga_n[KI], gb_n[KI];
    fetch(qs, 0, ga_n, gb_n);

    for (int q = qs; q < q1; ++q) {
        forward_from_ring<G>(ring, win, X, Y, tw, t);
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) ring[i] = ring[i + G::ADV];
        if (q + 1 < q1) load_slots<G, G::ADV>(row, p.n, base + G::RS, t, &ring[G::KEEP]);
        const bool own = q >= q0;
        const bool fast = o.interior(base);
#pragma unroll
        for (int i = 0; i < KI; ++i) { ga[i] = ga_n[i]; gb[i] = gb_n[i]; }
#pragma unroll
        for (int s = 0; s < ST; ++s) {
            if (s < ns) {
                if (s + 1 < ns) fetch(q, s + 1, ga_n, gb_n);
                else if (q + 1 < q1) fetch(q + 1, 0, ga_n, gb_n);
#pragma unroll
                for (int i = 0; i < KI; ++i) {
                    const int k = t + i * G::TPT;
                    float2 Sa, Sb;
                    split_bin(X[pad<R0>(k)], X[pad<R0>((N - k) & (N - 1))], Sa, Sb);
                    pack_bin<R0>(make_float2(Sa.x * ga[i], Sa.y * ga[i]), make_float2(Sb.x * gb[i], Sb.y * gb[i]), Y, k, N);
                }
                if (t == 0) {                                      // Nyquist: shares gain 0 with DC (ops.py:234-237)
                    float2 Sa, Sb;
                    split_bin(X[pad<R0>(NH)], X[pad<R0>(NH)], Sa, Sb);
                    pack_bin<R0>(make_float2(Sa.x * ga[0], Sa.y * ga[0]), make_float2(Sb.x * gb[0], Sb.y * gb[0]), Y, NH, N);
                }
                __syncthreads();
                float2 v[R0];
                inverse_to_regs<G>(Y, W, v, tw, t);
                float cur[G::RS];
#pragma unroll
                for (int i = 0; i < G::RS; ++i) cur[i] = i < G::KEEP ? acc[s][i] : 0.f;
                ola_accumulate<G>(v, win, cur);
                if (own) {
                    float* orow = orow0 + s * p.ld_out;
                    if (fast) {
#pragma unroll
                        for (int i = 0; i < G::ADV; ++i) o.write_fast(orow, base + i, i % HS, cur[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < G::ADV; ++i) o.write(orow, base + i, i % HS, cur[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < G::KEEP; ++i) acc[s][i] = cur[i + G::ADV];
#pragma unroll
                for (int i = 0; i < KI; ++i) { ga[i] = ga_n[i]; gb[i] = gb_n[i]; }
                __syncthreads();                    // Y (read by the last pass) is rewritten by the next source's pack
            }
        }
        base += G::ADV;
    }
    if (c == p.nchunk - 1) {
#pragma unroll
        for (int s = 0; s < ST; ++s)
            if (s < ns) {
#pragma unroll
                for (int i = 0; i < G::KEEP; ++i) o.write(orow0 + s * p.ld_out, base + i, i % HS, acc[s][i]);
            }
    }
}

// ---------------------------------------------------------------------------
// feature-fed synthesis (gss_mask_istft_feature): the mixture's linear packed spectrum [B, T, N] is read back
// instead of being recomputed from the waveform - ST instead of 1 + ST transforms per frame pair.  The two feature
// rows of a pair (2N contiguous floats) arrive by one 1-D TMA bulk copy into a double-buffered stage, one pair
// ahead; every thread keeps its own bins (k = t + i*TPT) of the pair in registers for all ST sources.
// ---------------------------------------------------------------------------
struct SynthFeatArgs {
    const float* feat; const float* mask; float* out;
    int64_t B, T, ld_out;
    int S, ngroups;
    int npairs, ppc, nchunk;
    int rev;
};
template <int N> constexpr size_t synth_feat_bytes() {
    return sizeof(float2) * 2 * (size_t)(N + N / Plan<N>::R0) + sizeof(float) * 4 * (size_t)N + 2 * sizeof(uint64_t);
}

// resident CTAs per SM the feature-fed kernel is compiled for (tuning macro; default = the waveform-fed kernel's)
#ifndef GSS_TEAM_FEAT_MINB
#define GSS_TEAM_FEAT_MINB(N) Plan<N>::MINB_FEAT
#endif
template <int N, int HS, int ST>
__global__ void __launch_bounds__(TGeo<N, HS>::TPT, GSS_TEAM_FEAT_MINB(N)) mask_istft_feat_kernel(const SynthFeatArgs p) {
    typedef TGeo<N, HS> G;
    constexpr int NH = N / 2, R0 = G::R0;
    extern __shared__ float4 smem4[];
    float2* Y = reinterpret_cast<float2*>(smem4);
    float2* W = Y + G::PADN;
    float* F = reinterpret_cast<float*>(W + G::PADN);            // 2 stages x 2N floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(F + 4 * N);
    const int t = threadIdx.x;
    if (t == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    const int64_t per_b = (int64_t)p.ngroups * p.nchunk;
    int64_t item = blockIdx.x;
    if (p.rev) item = p.B * per_b - 1 - item;
    const int64_t b = item / per_b;
    const int rem = (int)(item - b * per_b);
    const int grp = rem / p.nchunk, c = rem - grp * p.nchunk;
    const int s0 = grp * ST;
    const int ns = min(ST, p.S - s0);
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);
    const int qs = max(q0 - G::HALO, 0);

    TeamTw<G> tw; init_tw<G>(tw, t);
    float win[G::FS];
    make_window<G>(t, 1.0f, win);
    Ola<G> o;
    o.init(p.T, t, 1.0f);
    float* orow0 = p.out + (b * p.S + s0) * p.ld_out;
    const float* mrow0 = p.mask + ((b * p.S + s0) * p.T) * NH;
    const int64_t msrc = p.T * NH;
    const float* frow = p.feat + b * p.T * N;

    auto stage_features = [&](int q_) {
        if (t == 0) {
            const int st = (q_ - qs) & 1;
            const int64_t ta_ = 2 * (int64_t)q_;
            const uint32_t bytes = (ta_ + 1 < p.T ? 2 : 1) * N * (uint32_t)sizeof(float);
            mbar_expect_tx(&bars[st], bytes);
            tma_load_1d(F + st * 2 * N, frow + ta_ * N, bytes, &bars[st]);
        }
    };

    float acc[ST][G::KEEP];
#pragma unroll
    for (int s = 0; s < ST; ++s)
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) acc[s][i] = 0.f;
    int64_t base = (int64_t)2 * qs * HS;

    constexpr int KI = NH / G::TPT;
    auto fetch = [&](int q_, int s, float (&a)[KI], float (&bq)[KI]) {
        const int64_t ta_ = 2 * (int64_t)q_;
        const bool hb_ = ta_ + 1 < p.T;
        const float* ma = mrow0 + s * msrc + ta_ * NH + t;
#pragma unroll
        for (int i = 0; i < KI; ++i) { a[i] = ldg_here(ma + i * G::TPT); bq[i] = hb_ ? ldg_here(ma + NH + i * G::TPT) : 0.f; }
    };
    float ga[KI], gb[KI], ga_n[KI], gb_n[KI];
    stage_features(qs);
    fetch(qs, 0, ga_n, gb_n);

    for (int q = qs; q < q1; ++q) {
        // the other stage was last read at the top of pair q-1; CTA barriers have passed since
        if (q + 1 < q1) stage_features(q + 1);
        const int st = (q - qs) & 1;
        mbar_wait(&bars[st], (uint32_t)((q - qs) >> 1) & 1u);
        const bool hb = 2 * (int64_t)q + 1 < p.T;
        const float* f = F + st * 2 * N + t;
        float2 Sa[KI], Sb[KI];
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            Sa[i] = make_float2(f[i * G::TPT], f[NH + i * G::TPT]);
            Sb[i] = hb ? make_float2(f[N + i * G::TPT], f[N + NH + i * G::TPT]) : make_float2(0.f, 0.f);
        }
        const bool own = q >= q0;
        const bool fast = o.interior(base);
#pragma unroll
        for (int i = 0; i < KI; ++i) { ga[i] = ga_n[i]; gb[i] = gb_n[i]; }
#pragma unroll
        for (int s = 0; s < ST; ++s) {
            if (s < ns) {
                if (s + 1 < ns) fetch(q, s + 1, ga_n, gb_n);
                else if (q + 1 < q1) fetch(q + 1, 0, ga_n, gb_n);
#pragma unroll
                for (int i = 0; i < KI; ++i) {
                    const int k = t + i * G::TPT;
                    float2 Ya = make_float2(Sa[i].x * ga[i], Sa[i].y * ga[i]), Yb = make_float2(Sb[i].x * gb[i], Sb[i].y * gb[i]);
                    if (k == 0) {         // slot 0 = (DC, Nyquist): bin 0 -> (f[0], 0), bin N/2 -> (f[N/2], 0), one shared gain
                        pack_bin<R0>(make_float2(Ya.y, 0.f), make_float2(Yb.y, 0.f), Y, NH, N);
                        Ya.y = 0.f; Yb.y = 0.f;
                    }
                    pack_bin<R0>(Ya, Yb, Y, k, N);
                }
                __syncthreads();
                float2 v[R0];
                inverse_to_regs<G>(Y, W, v, tw, t);
                float cur[G::RS];
#pragma unroll
                for (int i = 0; i < G::RS; ++i) cur[i] = i < G::KEEP ? acc[s][i] : 0.f;
                ola_accumulate<G>(v, win, cur);
                if (own) {
                    float* orow = orow0 + s * p.ld_out;
                    if (fast) {
#pragma unroll
                        for (int i = 0; i < G::ADV; ++i) o.write_fast(orow, base + i, i % HS, cur[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < G::ADV; ++i) o.write(orow, base + i, i % HS, cur[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < G::KEEP; ++i) acc[s][i] = cur[i + G::ADV];
#pragma unroll
                for (int i = 0; i < KI; ++i) { ga[i] = ga_n[i]; gb[i] = gb_n[i]; }
                __syncthreads();                    // Y (read by the last pass) is rewritten by the next source's pack
            }
        }
        base += G::ADV;
    }
    if (c == p.nchunk - 1) {
#pragma unroll
        for (int s = 0; s < ST; ++s)
            if (s < ns) {
#pragma unroll
                for (int i = 0; i < G::KEEP; ++i) o.write(orow0 + s * p.ld_out, base + i, i % HS, acc[s][i]);
            }
    }
}

}  // namespace team
}  // namespace gss
