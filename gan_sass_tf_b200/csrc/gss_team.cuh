// gss_team.cuh - streaming STFT / iSTFT / mask-iSTFT kernels for the FFT sizes the
// register-exchange geometry of gss_fft.cuh does not cover (256, 1024, 2048, 4096; 512 as a
// cross-check path).
//
// Same streaming structure as gss_stream.cuh - a team walks a run of consecutive frame PAIRS of
// one utterance (re = frame 2q, im = frame 2q+1 of one complex N-point transform), the raw-sample
// ring and the overlap-add accumulators live in registers, every sample is read from global
// memory once and every output sample written once, no atomics - but the team is a whole CTA of
// TPT threads and the transform is a Stockham autosort FFT in shared memory: radix-16 / 8 / 4
// passes, one butterfly per thread per pass, natural order in and out (so the two-for-one split,
// the mask multiply and the Hermitian pack address bins k and N-k directly, and masks / features
// are read and written fully coalesced), per-thread twiddles kept in registers as the powers
// w^1, w^2, w^4, w^8 of the thread's base twiddle and expanded on the fly.
//
//   SLOT = 2*TPT samples (thread t owns samples 2t, 2t+1 of every slot); a frame is FS slots,
//   the hop HS slots.  Buffers are padded by one element per 16 to keep the radix-16 scatter of
//   the first pass free of bank conflicts.
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
template <int N_> struct Plan;      // radices of the passes (product N), team size
// MINB: resident CTAs per SM the register allocation is held to (128 registers per thread)
template <> struct Plan<256>  { static constexpr int TPT = 32,  NP = 2, R0 = 16, R1 = 16, R2 = 1,  MINB = 16, MINB_SYNTH = 16; };
template <> struct Plan<512>  { static constexpr int TPT = 64,  NP = 3, R0 = 8,  R1 = 8,  R2 = 8,  MINB = 8, MINB_SYNTH = 8; };
template <> struct Plan<1024> { static constexpr int TPT = 128, NP = 3, R0 = 16, R1 = 8,  R2 = 8,  MINB = 4, MINB_SYNTH = 4; };
template <> struct Plan<2048> { static constexpr int TPT = 256, NP = 3, R0 = 16, R1 = 16, R2 = 8,  MINB = 2, MINB_SYNTH = 2; };
template <> struct Plan<4096> { static constexpr int TPT = 256, NP = 3, R0 = 16, R1 = 16, R2 = 16, MINB = 2, MINB_SYNTH = 1; };

template <int N_, int HS_>
struct TGeo {
    typedef Plan<N_> P;
    static constexpr int N = N_, HS = HS_, TPT = P::TPT;
    static constexpr int SLOT = 2 * TPT;
    static constexpr int FS = N / SLOT;            // slots per frame (4, or 8 at N = 4096)
    static constexpr int H = HS * SLOT;
    static constexpr int R = FS / HS;              // frames covering one sample
    static constexpr int RS = FS + HS;             // slots spanned by a frame pair
    static constexpr int ADV = 2 * HS;
    static constexpr int KEEP = RS - ADV;
    static constexpr int HALO = R / 2;
    static constexpr bool CONST_NORM = (R >= 4);
    static constexpr int PADN = N + N / 16;        // padded complex elements per buffer
    static_assert(FS % HS == 0 && R >= 2, "hop must divide the frame into >= 2 parts");
};
__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }

// read-only global load that stays where it is written: the compiler is free to sink a plain __ldg past
// the CTA barriers down to its use (saving registers, exposing the whole DRAM latency); prefetches must not move
__device__ __forceinline__ float ldg_here(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// base twiddles of one pass for this thread: W_(NS*R)^(k * 2^i), k = j mod NS
struct PassTw { float2 w[4]; };
template <int R, int NS>
__device__ __forceinline__ void init_pass_tw(PassTw& p, int j) {
    const int k = j & (NS - 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s, c;
        sincospif(-2.0f * (float)(k << i) / (float)(NS * R), &s, &c);
        p.w[i] = make_float2(c, s);
    }
}

// one Stockham pass: thread j < N/R, k = j mod NS:
//   v[r] = in[j + r N/R] * w^r ;  V = DFT_R(v) ;  out[(j - k) R + k + r NS] = V[r]
template <int N, int R, int NS, bool INV>
__device__ __forceinline__ void pass(const float2* __restrict__ in, float2* __restrict__ out, const PassTw& tw, int j) {
    if (j < N / R) {
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[pad(j + r * (N / R))];
        if (NS > 1) {
            float2 w[R];
            w[1] = tw.w[0];
            if (R > 2) { w[2] = tw.w[1]; w[3] = cmul2(w[1], w[2]); }
            if (R > 4) { w[4] = tw.w[2]; w[5] = cmul2(w[1], w[4]); w[6] = cmul2(w[2], w[4]); w[7] = cmul2(w[3], w[4]); }
            if (R > 8) {
                w[8] = tw.w[3];
#pragma unroll
                for (int r = 9; r < R; ++r) w[r] = cmul2(w[r - 8], w[8]);
            }
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = twmul<INV>(v[r], w[r]);
        }
        dftR<R, INV>(v);
        const int k = j & (NS - 1);
        const int o = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out[pad(o + r * NS)] = v[r];
    }
}

template <int N>
struct TeamTw { PassTw p1, p2; };

template <int N>
__device__ __forceinline__ void init_tw(TeamTw<N>& t, int j) {
    typedef Plan<N> P;
    init_pass_tw<P::R1, P::R0>(t.p1, j);
    if constexpr (P::NP > 2) init_pass_tw<P::R2, P::R0 * P::R1>(t.p2, j);
}

// N-point complex FFT of b0 (natural order); b1 is scratch.  Returns the buffer holding the result
// (b1 after an odd number of passes, b0 after an even number).  Ends with a CTA barrier.
template <int N, bool INV>
__device__ __forceinline__ float2* fft(float2* b0, float2* b1, const TeamTw<N>& tw, int j) {
    typedef Plan<N> P;
    PassTw none;
    pass<N, P::R0, 1, INV>(b0, b1, none, j);
    __syncthreads();
    pass<N, P::R1, P::R0, INV>(b1, b0, tw.p1, j);
    __syncthreads();
    if constexpr (P::NP > 2) {
        pass<N, P::R2, P::R0 * P::R1, INV>(b0, b1, tw.p2, j);
        __syncthreads();
        return b1;
    } else {
        return b0;
    }
}

// ---------------------------------------------------------------------------
// sample access and overlap-add output (thread t owns samples 2t, 2t+1 of every slot)
// ---------------------------------------------------------------------------
template <int SLOT, int FS, int CNT, typename TIn>
__device__ __forceinline__ void load_slots(const TIn* row, int64_t n, int64_t sl, int t, bool al, v2* dst) {
    const bool inside = sl >= FS / 2 && (sl + CNT - FS / 2) * SLOT <= n;
#pragma unroll
    for (int i = 0; i < CNT; ++i) {
        const int64_t p = (sl + i - FS / 2) * SLOT + 2 * t;
        dst[i] = inside ? load_pair_fast(row, p, al) : load_pair_edge(row, n, p);
    }
}

template <class G>
struct Ola {
    int64_t T;
    int t;
    bool al;
    float oscale;
    v2 invn[G::HS];
    static constexpr float kScale = G::CONST_NORM ? 0.5f / (0.375f * G::R) : 0.5f;

    __device__ __forceinline__ void init(int64_t T_, int t_, bool al_, float win_gain) {
        T = T_; t = t_; al = al_;
        oscale = kScale * win_gain;
        if (!G::CONST_NORM) {
#pragma unroll
            for (int m = 0; m < G::HS; ++m) {
                float s0 = 0.f, s1 = 0.f;
                for (int r = 0; r < G::R; ++r) {
                    float w0 = hann<G::N>(m * G::SLOT + 2 * t + r * G::H);
                    float w1 = hann<G::N>(m * G::SLOT + 2 * t + 1 + r * G::H);
                    s0 += w0 * w0; s1 += w1 * w1;
                }
                invn[m] = make_float2(oscale / s0, oscale / s1);
            }
        }
    }
    __device__ __noinline__ float norm_at(int64_t sl, int e) const {
        int64_t tlo = sl - (G::FS - 1); tlo = tlo <= 0 ? 0 : (tlo + G::HS - 1) / G::HS;
        int64_t thi = sl / G::HS; if (thi > T - 1) thi = T - 1;
        float s = 0.f;
        for (int64_t f = tlo; f <= thi; ++f) {
            int i = (int)((sl - f * G::HS) * G::SLOT) + 2 * t + e;
            float w = hann<G::N>(i);
            s += w * w;
        }
        return s > 1e-10f ? s : 1.0f;
    }
    __device__ __forceinline__ void write(float* row, int64_t sl, int m, v2 v) const {
        v = vmul(v, G::CONST_NORM ? vset(oscale) : invn[m]);
        if (sl < G::FS / 2 || sl >= G::FS / 2 + (T - 1) * G::HS) return;
        if (G::CONST_NORM && (sl <= G::FS - 1 - G::HS || sl >= T * G::HS)) {
            const float c = 0.375f * G::R;
            v.x *= c / norm_at(sl, 0); v.y *= c / norm_at(sl, 1);
        }
        float* p = row + (sl - G::FS / 2) * G::SLOT + 2 * t;
        if (al) *reinterpret_cast<float2*>(p) = v; else { p[0] = v.x; p[1] = v.y; }
    }
};

template <class G>
__device__ __forceinline__ void make_window(int t, float scale, v2 (&w)[G::FS]) {
#pragma unroll
    for (int i = 0; i < G::FS; ++i)
        w[i] = make_float2(scale * hann<G::N>(2 * t + G::SLOT * i), scale * hann<G::N>(2 * t + 1 + G::SLOT * i));
}

// windowed frame pair -> z[n] = (frame a, frame b) in natural order
template <class G>
__device__ __forceinline__ void stage_pair(const v2 (&ring)[G::RS], const v2 (&win)[G::FS], float2* z, int t) {
#pragma unroll
    for (int i = 0; i < G::FS; ++i) {
        const v2 a = vmul(ring[i], win[i]), b = vmul(ring[G::HS + i], win[i]);
        const int p = pad(2 * t + G::SLOT * i);          // 2t is even: 2t and 2t+1 share a 16-group
        z[p] = make_float2(a.x, b.x);
        z[p + 1] = make_float2(a.y, b.y);
    }
}

// two-for-one: spectra of frames a and b at bin k from Z[k], Z[N-k] (the window carries the 1/2)
__device__ __forceinline__ void split_bin(float2 zk, float2 zn, float2& A, float2& B) {
    A = make_float2(zk.x + zn.x, zk.y - zn.y);            // Z[k] + conj Z[N-k]
    B = make_float2(zk.y + zn.y, zn.x - zk.x);            // (Z[k] - conj Z[N-k]) / i
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
template <typename TIn>
struct StftArgs {
    const TIn* wave; float* feat;
    int64_t B, n, ld, T;
    int npairs, ppc, nchunk;
    int al_in; int log; float eps;
};

template <int N, int HS, typename TIn>
__global__ void __launch_bounds__(Plan<N>::TPT, Plan<N>::MINB) stft_kernel(const StftArgs<TIn> p) {
    typedef TGeo<N, HS> G;
    extern __shared__ float4 smem4[];
    float2* b0 = reinterpret_cast<float2*>(smem4);
    float2* b1 = b0 + G::PADN;
    const int t = threadIdx.x;
    const int64_t item = blockIdx.x;
    const int64_t b = item / p.nchunk;
    const int c = (int)(item - b * p.nchunk);
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);

    TeamTw<N> tw; init_tw<N>(tw, t);
    v2 win[G::FS];
    make_window<G>(t, 1.0f / (float)N, win);
    const TIn* row = p.wave + b * p.ld;
    const bool al = p.al_in != 0;
    int64_t base = (int64_t)2 * q0 * HS;
    v2 ring[G::RS];
    load_slots<G::SLOT, G::FS, G::RS>(row, p.n, base, t, al, ring);

    for (int q = q0; q < q1; ++q) {
        stage_pair<G>(ring, win, b0, t);
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) ring[i] = ring[i + G::ADV];
        if (q + 1 < q1) load_slots<G::SLOT, G::FS, G::ADV>(row, p.n, base + G::RS, t, al, &ring[G::KEEP]);
        __syncthreads();
        const float2* Z = fft<N, false>(b0, b1, tw, t);
        const int64_t ta = 2 * (int64_t)q;
        float* fa = p.feat + (b * p.T + ta) * N;
        const bool hb = ta + 1 < p.T;
        for (int k = t; k < N / 2; k += G::TPT) {
            float2 A, B;
            split_bin(Z[pad(k)], Z[pad((N - k) & (N - 1))], A, B);
            if (k == 0) {               // slot 0 carries (DC, Nyquist) (app/utils.py:22-26)
                float2 An, Bn;
                const float2 zh = Z[pad(N / 2)];
                split_bin(zh, zh, An, Bn);
                A.y = An.x; B.y = Bn.x;
            }
            if (p.log) {
                float g = log_gain(A.x, A.y, p.eps); A.x *= g; A.y *= g;
                g = log_gain(B.x, B.y, p.eps); B.x *= g; B.y *= g;
            }
            fa[k] = A.x; fa[N / 2 + k] = A.y;
            if (hb) { fa[N + k] = B.x; fa[N + N / 2 + k] = B.y; }
        }
        base += G::ADV;
        __syncthreads();
    }
}

struct IstftArgs {
    const float* feat; float* out;
    int64_t rows, T, ld_out;
    int npairs, ppc, nchunk;
    int al_out; int exp; float eps;
};

// spectra Ya, Yb of two real frames at bin k (k in [0, N/2]) -> Y[k], Y[N-k] of the complex transform
__device__ __forceinline__ void pack_bin(float2 Ya, float2 Yb, float2* Y, int k, int N) {
    Y[pad(k)] = make_float2(Ya.x - Yb.y, Ya.y + Yb.x);                       // Ya + i Yb
    Y[pad((N - k) & (N - 1))] = make_float2(Ya.x + Yb.y, Yb.x - Ya.y);       // conj(Ya) + i conj(Yb)
}

template <class G>
__device__ __forceinline__ void ola_accumulate(const float2* z, const v2 (&win)[G::FS], v2 (&cur)[G::RS], int t) {
#pragma unroll
    for (int i = 0; i < G::FS; ++i) {
        const int p = pad(2 * t + G::SLOT * i);
        const float2 z0 = z[p], z1 = z[p + 1];
        cur[i] = vfma(make_float2(z0.x, z1.x), win[i], cur[i]);
        cur[G::HS + i] = vfma(make_float2(z0.y, z1.y), win[i], cur[G::HS + i]);
    }
}

template <int N, int HS>
__global__ void __launch_bounds__(Plan<N>::TPT, Plan<N>::MINB) istft_kernel(const IstftArgs p) {
    typedef TGeo<N, HS> G;
    extern __shared__ float4 smem4[];
    float2* b0 = reinterpret_cast<float2*>(smem4);
    float2* b1 = b0 + G::PADN;
    const int t = threadIdx.x;
    const int64_t item = blockIdx.x;
    const int64_t r = item / p.nchunk;
    const int c = (int)(item - r * p.nchunk);
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);
    const int qs = max(q0 - G::HALO, 0);

    TeamTw<N> tw; init_tw<N>(tw, t);
    v2 win[G::FS];
    make_window<G>(t, 1.0f, win);
    Ola<G> o;
    o.init(p.T, t, p.al_out != 0, 1.0f);
    float* orow = p.out + r * p.ld_out;
    const float* frow = p.feat + r * p.T * N;
    v2 acc[G::KEEP];
#pragma unroll
    for (int i = 0; i < G::KEEP; ++i) acc[i] = make_float2(0.f, 0.f);
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
                float g = exp_gain(Ya.x, Ya.y, p.eps); Ya.x *= g; Ya.y *= g;
                g = exp_gain(Yb.x, Yb.y, p.eps); Yb.x *= g; Yb.y *= g;
            }
            if (k == 0) {         // bin 0 -> (f[0], 0), bin N/2 -> (f[N/2], 0)
                pack_bin(make_float2(Ya.y, 0.f), make_float2(Yb.y, 0.f), b0, N / 2, N);
                Ya.y = 0.f; Yb.y = 0.f;
            }
            pack_bin(Ya, Yb, b0, k, N);
        }
        __syncthreads();
        const float2* z = fft<N, true>(b0, b1, tw, t);
        v2 cur[G::RS];
#pragma unroll
        for (int i = 0; i < G::RS; ++i) cur[i] = i < G::KEEP ? acc[i] : make_float2(0.f, 0.f);
        ola_accumulate<G>(z, win, cur, t);
        if (q >= q0) {
#pragma unroll
            for (int i = 0; i < G::ADV; ++i) o.write(orow, base + i, i % HS, cur[i]);
        }
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) acc[i] = cur[i + G::ADV];
#pragma unroll
        for (int i = 0; i < KI; ++i) { fya[i] = nya[i]; fyb[i] = nyb[i]; }      // the next pair's features have had a whole transform to arrive
        base += G::ADV;
        __syncthreads();
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
    int al_in, al_out;
};

template <int N, int HS, int ST>
__global__ void __launch_bounds__(Plan<N>::TPT, Plan<N>::MINB_SYNTH) mask_istft_kernel(const SynthArgs p) {
    typedef TGeo<N, HS> G;
    constexpr int NH = N / 2;
    extern __shared__ float4 smem4[];
    float2* b0 = reinterpret_cast<float2*>(smem4);
    float2* b1 = b0 + G::PADN;
    float2* b2 = b1 + G::PADN;
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

    TeamTw<N> tw; init_tw<N>(tw, t);
    v2 win[G::FS];
    make_window<G>(t, 1.0f / (float)N, win);
    Ola<G> o;
    o.init(p.T, t, p.al_out != 0, (float)N);
    float* orow0 = p.out + (b * p.S + s0) * p.ld_out;
    const float* mrow0 = p.mask + ((b * p.S + s0) * p.T) * NH;
    const int64_t msrc = p.T * NH;
    const float* row = p.wave + b * p.ld;
    const bool al = p.al_in != 0;

    v2 acc[ST][G::KEEP];
#pragma unroll
    for (int s = 0; s < ST; ++s)
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) acc[s][i] = make_float2(0.f, 0.f);
    int64_t base = (int64_t)2 * qs * HS;
    v2 ring[G::RS];
    load_slots<G::SLOT, G::FS, G::RS>(row, p.n, base, t, al, ring);

    for (int q = qs; q < q1; ++q) {
        stage_pair<G>(ring, win, b0, t);
#pragma unroll
        for (int i = 0; i < G::KEEP; ++i) ring[i] = ring[i + G::ADV];
        if (q + 1 < q1) load_slots<G::SLOT, G::FS, G::ADV>(row, p.n, base + G::RS, t, al, &ring[G::KEEP]);
        __syncthreads();
        // mixture spectrum stays in X for every source; Y / W are the two buffers the inverse uses
        float2* X = fft<N, false>(b0, b1, tw, t);
        float2* Y = (X == b0) ? b1 : b0;
        float2* W = b2;
        const int64_t ta = 2 * (int64_t)q;
        const bool hb = ta + 1 < p.T;
        const bool own = q >= q0;
        // gains of source s for this thread's bins, fetched one source ahead of their use
        constexpr int KI = NH / G::TPT;                           // bins k = t + i*TPT, plus k = N/2 (gain 0 again)
        float ga[KI], gb[KI], ga_n[KI], gb_n[KI];
        auto fetch = [&](int s, float (&a)[KI], float (&bq)[KI]) {
            const float* ma = mrow0 + s * msrc + ta * NH + t;
#pragma unroll
            for (int i = 0; i < KI; ++i) { a[i] = ldg_here(ma + i * G::TPT); bq[i] = hb ? ldg_here(ma + NH + i * G::TPT) : 0.f; }
        };
        fetch(0, ga, gb);
#pragma unroll
        for (int s = 0; s < ST; ++s) {
            if (s < ns) {
                if (s + 1 < ns) fetch(s + 1, ga_n, gb_n);
#pragma unroll
                for (int i = 0; i < KI; ++i) {
                    const int k = t + i * G::TPT;
                    float2 A, B;
                    split_bin(X[pad(k)], X[pad((N - k) & (N - 1))], A, B);
                    pack_bin(make_float2(A.x * ga[i], A.y * ga[i]), make_float2(B.x * gb[i], B.y * gb[i]), Y, k, N);
                }
                if (t == 0) {                                      // Nyquist: shares gain 0 with DC (ops.py:234-237)
                    float2 A, B;
                    split_bin(X[pad(NH)], X[pad(NH)], A, B);
                    pack_bin(make_float2(A.x * ga[0], A.y * ga[0]), make_float2(B.x * gb[0], B.y * gb[0]), Y, NH, N);
                }
                __syncthreads();
                const float2* z = fft<N, true>(Y, W, tw, t);
                v2 cur[G::RS];
#pragma unroll
                for (int i = 0; i < G::RS; ++i) cur[i] = i < G::KEEP ? acc[s][i] : make_float2(0.f, 0.f);
                ola_accumulate<G>(z, win, cur, t);
                if (own) {
                    float* orow = orow0 + s * p.ld_out;
#pragma unroll
                    for (int i = 0; i < G::ADV; ++i) o.write(orow, base + i, i % HS, cur[i]);
                }
#pragma unroll
                for (int i = 0; i < G::KEEP; ++i) acc[s][i] = cur[i + G::ADV];
#pragma unroll
                for (int i = 0; i < KI; ++i) { ga[i] = ga_n[i]; gb[i] = gb_n[i]; }
                __syncthreads();
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
