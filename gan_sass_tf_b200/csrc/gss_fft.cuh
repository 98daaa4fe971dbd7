// gss_fft.cuh - per-team complex FFT used by every spectral kernel of libgss.
//
// Geometry (modelled one-to-one, including the shared-memory layouts and their bank-conflict
// counts, in tools/fft_model.py and checked on the host by tests/test_fft_model.py):
//
//   N = 64*M complex points, M in {4, 8} (N = 256, 512; 16 is modelled only);  a "team" of TPF = N/16
//   threads (half a warp / one warp) owns one transform, 16 complex points per thread, three passes
//   radix 8 / M / 8.
//   L = N/8.  All arithmetic is packed FP32x2 (FADD2/FMUL2/FFMA2): a v2 holds the
//   same quantity for the two butterflies a thread owns in a pass.
//
//   pass 0 : thread j, lanes e=0,1 own butterflies n' = 2j+e over x[n' + L*n0]
//   middle : DFT-M over n1 for (k0, n2), n' = 8*n1 + n2
//   last   : lanes (A,B) own butterflies c = {j, L-j} (thread 0: {0, L/2}) and
//            produce Z[c + L*k2]: Z[k] and Z[N-k] end up in the same thread, so
//            the two-for-one real split / Hermitian pack need no data exchange.
//
// Two real sequences ride one complex transform (re = sequence a, im = sequence b).
#pragma once
#include <cuda_runtime.h>

// middle-pass twiddles w^1..w^7 kept expanded in registers (16 more registers) instead of being rebuilt from
// w^1, w^2, w^4 in every transform (16 FP32x2 each): C2 stft_log 64.1 -> 62.4 us, stft_dual 91.1 -> 85.9 us
// (profiles/r2_variants1.txt); the synthesis kernels measure the same either way.  -DGSS_TW1_FULL=0 restores the rebuild.
#ifndef GSS_TW1_FULL
#define GSS_TW1_FULL 1
#endif
#ifndef GSS_E1_V2
#define GSS_E1_V2 0                  // experiment: second exchange of the N = 512 transform with 64-bit accesses on both sides
#endif

namespace gss {

typedef float2 v2;
struct cv2 { v2 re, im; };

__device__ __forceinline__ v2 vadd(v2 a, v2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ v2 vneg(v2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ v2 vsub(v2 a, v2 b) { return __fadd2_rn(a, vneg(b)); }
__device__ __forceinline__ v2 vmul(v2 a, v2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ v2 vfma(v2 a, v2 b, v2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ v2 vset(float s) { return make_float2(s, s); }
__device__ __forceinline__ v2 vswap(v2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ cv2 cadd(cv2 a, cv2 b) { cv2 r; r.re = vadd(a.re, b.re); r.im = vadd(a.im, b.im); return r; }
__device__ __forceinline__ cv2 csub(cv2 a, cv2 b) { cv2 r; r.re = vsub(a.re, b.re); r.im = vsub(a.im, b.im); return r; }

// a * w (forward) or a * conj(w) (inverse)
template <bool INV>
__device__ __forceinline__ cv2 cmul(cv2 a, cv2 w) {
    cv2 r;
    if (!INV) {
        r.re = vfma(a.re, w.re, vneg(vmul(a.im, w.im)));
        r.im = vfma(a.re, w.im, vmul(a.im, w.re));
    } else {
        r.re = vfma(a.re, w.re, vmul(a.im, w.im));
        r.im = vfma(a.im, w.re, vneg(vmul(a.re, w.im)));
    }
    return r;
}

#define GSS_SQRT1_2 0.70710678118654752440f

// in-place 8-point DFT, natural order in and out.  INV=false: e^{-2 pi i nk/8}.
template <bool INV>
__device__ __forceinline__ void dft8(cv2 (&a)[8]) {
    cv2 b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]);
    cv2 b1 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
    cv2 b2 = cadd(a[2], a[6]), b6 = csub(a[2], a[6]);
    cv2 b3 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
    // even outputs: DFT4(b0..b3)
    cv2 d0 = cadd(b0, b2), d1 = csub(b0, b2), d2 = cadd(b1, b3), d3 = csub(b1, b3);
    a[0] = cadd(d0, d2);
    a[4] = csub(d0, d2);
    if (!INV) {   // d3 * (-i) = (d3.im, -d3.re)
        a[2].re = vadd(d1.re, d3.im); a[2].im = vsub(d1.im, d3.re);
        a[6].re = vsub(d1.re, d3.im); a[6].im = vadd(d1.im, d3.re);
    } else {      // d3 * (+i) = (-d3.im, d3.re)
        a[2].re = vsub(d1.re, d3.im); a[2].im = vadd(d1.im, d3.re);
        a[6].re = vadd(d1.re, d3.im); a[6].im = vsub(d1.im, d3.re);
    }
    // odd outputs: DFT4(b4, b5*W, b6*W^2, b7*W^3), W = e^{-+ i pi/4}
    cv2 e0, e1, t5, t7;
    if (!INV) {
        e0.re = vadd(b4.re, b6.im); e0.im = vsub(b4.im, b6.re);     // b4 + b6*(-i)
        e1.re = vsub(b4.re, b6.im); e1.im = vadd(b4.im, b6.re);
        t5.re = vadd(b5.re, b5.im); t5.im = vsub(b5.im, b5.re);     // b5*(1-i)
        t7.re = vsub(b7.im, b7.re); t7.im = vneg(vadd(b7.im, b7.re)); // b7*(-1-i)
    } else {
        e0.re = vsub(b4.re, b6.im); e0.im = vadd(b4.im, b6.re);     // b4 + b6*(+i)
        e1.re = vadd(b4.re, b6.im); e1.im = vsub(b4.im, b6.re);
        t5.re = vsub(b5.re, b5.im); t5.im = vadd(b5.im, b5.re);     // b5*(1+i)
        t7.re = vneg(vadd(b7.re, b7.im)); t7.im = vsub(b7.re, b7.im); // b7*(-1+i)
    }
    cv2 u = cadd(t5, t7), v = csub(t5, t7);      // both still to be scaled by 1/sqrt2
    const v2 s = vset(GSS_SQRT1_2), ms = vset(-GSS_SQRT1_2);
    a[1].re = vfma(u.re, s, e0.re);  a[1].im = vfma(u.im, s, e0.im);
    a[5].re = vfma(u.re, ms, e0.re); a[5].im = vfma(u.im, ms, e0.im);
    if (!INV) {   // e3 = s*v*(-i) = s*(v.im, -v.re)
        a[3].re = vfma(v.im, s, e1.re);  a[3].im = vfma(v.re, ms, e1.im);
        a[7].re = vfma(v.im, ms, e1.re); a[7].im = vfma(v.re, s, e1.im);
    } else {      // e3 = s*v*(+i) = s*(-v.im, v.re)
        a[3].re = vfma(v.im, ms, e1.re); a[3].im = vfma(v.re, s, e1.im);
        a[7].re = vfma(v.im, s, e1.re);  a[7].im = vfma(v.re, ms, e1.im);
    }
}

// in-place 4-point DFT
template <bool INV>
__device__ __forceinline__ void dft4(cv2 (&a)[4]) {
    cv2 d0 = cadd(a[0], a[2]), d1 = csub(a[0], a[2]), d2 = cadd(a[1], a[3]), d3 = csub(a[1], a[3]);
    a[0] = cadd(d0, d2);
    a[2] = csub(d0, d2);
    if (!INV) {
        a[1].re = vadd(d1.re, d3.im); a[1].im = vsub(d1.im, d3.re);
        a[3].re = vsub(d1.re, d3.im); a[3].im = vadd(d1.im, d3.re);
    } else {
        a[1].re = vsub(d1.re, d3.im); a[1].im = vadd(d1.im, d3.re);
        a[3].re = vadd(d1.re, d3.im); a[3].im = vsub(d1.im, d3.re);
    }
}

// ---------------------------------------------------------------------------
template <int N_>
struct Geo {
    static constexpr int N = N_;
    static constexpr int M = N / 64;          // middle radix
    static constexpr int TPF = N / 16;        // threads per transform
    static constexpr int L = N / 8;
    static constexpr int P0 = TPF + 4;        // E0 pitch, v2 units (re plane, im plane at +E0_PLANE)
    static constexpr int P1 = L + 4;          // E1 pitch, floats
    static constexpr int E0_PLANE = 8 * P0;   // v2 per plane
    static constexpr int E0_F4 = 8 * P0;      // float4-equivalents (both planes)
    static constexpr int E1_PLANE = 8 * P1;   // floats per plane
    static constexpr int E1V_PLANE = 4 * P1;  // v2 per plane in the [q][column] pair layout (GSS_E1_V2): same bytes
    static constexpr int TEAM_FLOATS = E0_F4 * 4 + 2 * E1_PLANE;
    static_assert(M == 8 || M == 4, "implemented: N = 512 (M = 8, one warp per transform) and N = 256 (M = 4, half a warp); "
                                    "1024 is wired in fft_model.py only");
};

// per-thread constants of a team member
template <int N>
struct TeamCtx {
    int j;            // thread index inside the team
    int cA, cB;       // last-pass butterflies
    v2* e0;           // team exchange buffer 0 (two planes of v2)
    float* e1;        // team exchange buffer 1 (re plane, im plane at +E1_PLANE)
    cv2 tw0[8];       // pass-0 twiddles W_N^{(2j+e)*k0}, k0 = 1..7 ([0] unused)
    cv2 tw1[3];       // middle twiddles W_L^{(2q+e)*k1} for k1 = 1, 2, 4, q = j % 4
#if GSS_TW1_FULL
    cv2 tw1x[8];      // ... expanded once to k1 = 1..7 (16 more registers, 16 fewer FP32x2 per transform)
#endif
    unsigned mask;    // lanes of this team inside its warp (all 32 at N = 512, one half at N = 256)
};

// the seven twiddles w^1..w^7 from the stored w^1, w^2, w^4 (4 complex products)
__device__ __forceinline__ void expand_tw(const cv2 (&b)[3], cv2 (&w)[8]) {
    w[1] = b[0]; w[2] = b[1]; w[4] = b[2];
    w[3] = cmul<false>(b[0], b[1]);
    w[5] = cmul<false>(b[0], b[2]);
    w[6] = cmul<false>(b[1], b[2]);
    w[7] = cmul<false>(w[3], b[2]);
}
#if GSS_TW1_FULL
#define GSS_MID_TW(c, w) const cv2 (&w)[8] = (c).tw1x
#else
#define GSS_MID_TW(c, w) cv2 w[8]; expand_tw((c).tw1, w)
#endif

// teams never span warps (TPF <= 32): a team-scoped __syncwarp keeps the two half-warp teams of N = 256 independent
template <int N>
__device__ __forceinline__ void team_sync(const TeamCtx<N>& c) { __syncwarp(c.mask); }
template <int N>
__device__ __forceinline__ unsigned team_mask() {
    constexpr int TPF = Geo<N>::TPF;
    if (TPF >= 32) return 0xffffffffu;
    return ((1u << TPF) - 1u) << ((threadIdx.x & 31) / TPF * TPF);
}

template <int N>
__device__ __forceinline__ void team_init(TeamCtx<N>& c, int j, float* team_smem) {
    typedef Geo<N> G;
    c.j = j;
    c.cA = j ? j : 0;
    c.cB = j ? G::L - j : G::L / 2;
    c.e0 = reinterpret_cast<v2*>(team_smem);
    c.e1 = team_smem + G::E0_F4 * 4;
    c.mask = team_mask<N>();
    const int q = j & 3;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const int k = 1 << b;
        float s0, c0, s1, c1;
        sincospif(-2.0f * (float)((2 * q) * k) / (float)G::L, &s0, &c0);
        sincospif(-2.0f * (float)((2 * q + 1) * k) / (float)G::L, &s1, &c1);
        c.tw1[b].re = make_float2(c0, c1);
        c.tw1[b].im = make_float2(s0, s1);
    }
#pragma unroll
    for (int k0 = 1; k0 < 8; ++k0) {
        float s0, c0, s1, c1;
        sincospif(-2.0f * (float)((2 * j) * k0) / (float)N, &s0, &c0);
        sincospif(-2.0f * (float)((2 * j + 1) * k0) / (float)N, &s1, &c1);
        c.tw0[k0].re = make_float2(c0, c1);
        c.tw0[k0].im = make_float2(s0, s1);
    }
}

// Per-device table of the lane constants (pass-0 twiddles, middle twiddles, Hann window at the lane's
// sample positions), filled once by tables_kernel: a team then starts with 28 coalesced 8-byte loads
// instead of ~36 sincospif / cospif evaluations per thread (about two frame pairs' worth of
// instructions, paid by every work item).  Same code computes the values, so results are bit-identical.
constexpr int TAB_TW0 = 0;        // [k0-1][re/im] : 14 rows
constexpr int TAB_TW1 = 14;       // [b][re/im]    : 6 rows
constexpr int TAB_WIN = 20;       // [n0]          : 8 rows (unscaled periodic Hann at samples 2j+e + L*n0)
constexpr int TAB_ROWS = 28;
// one copy per translation unit (the library is built in parts): each part fills and reads its own
static __device__ v2 g_lane_tab512[TAB_ROWS][32];
static __device__ v2 g_lane_tab256[TAB_ROWS][16];
template <int N> __device__ __forceinline__ v2& lane_tab(int row, int j) {
    if constexpr (N == 512) return g_lane_tab512[row][j]; else return g_lane_tab256[row][j];
}

template <int N>
__global__ void tables_kernel() {
    const int j = threadIdx.x;
    TeamCtx<N> c;
    team_init<N>(c, j, nullptr);
#pragma unroll
    for (int k0 = 1; k0 < 8; ++k0) { lane_tab<N>(TAB_TW0 + 2 * (k0 - 1), j) = c.tw0[k0].re; lane_tab<N>(TAB_TW0 + 2 * (k0 - 1) + 1, j) = c.tw0[k0].im; }
#pragma unroll
    for (int b = 0; b < 3; ++b) { lane_tab<N>(TAB_TW1 + 2 * b, j) = c.tw1[b].re; lane_tab<N>(TAB_TW1 + 2 * b + 1, j) = c.tw1[b].im; }
#pragma unroll
    for (int n0 = 0; n0 < 8; ++n0) {
        const int i0 = 2 * j + Geo<N>::L * n0;
        lane_tab<N>(TAB_WIN + n0, j) = make_float2(0.5f - 0.5f * cospif(2.0f * (float)i0 / (float)N),
                                                   0.5f - 0.5f * cospif(2.0f * (float)(i0 + 1) / (float)N));
    }
}

template <int N>
__device__ __forceinline__ void team_init_tab(TeamCtx<N>& c, int j, float* team_smem) {
    typedef Geo<N> G;
    c.j = j;
    c.cA = j ? j : 0;
    c.cB = j ? G::L - j : G::L / 2;
    c.e0 = reinterpret_cast<v2*>(team_smem);
    c.e1 = team_smem + G::E0_F4 * 4;
    c.mask = team_mask<N>();
#pragma unroll
    for (int k0 = 1; k0 < 8; ++k0) { c.tw0[k0].re = lane_tab<N>(TAB_TW0 + 2 * (k0 - 1), j); c.tw0[k0].im = lane_tab<N>(TAB_TW0 + 2 * (k0 - 1) + 1, j); }
#pragma unroll
    for (int b = 0; b < 3; ++b) { c.tw1[b].re = lane_tab<N>(TAB_TW1 + 2 * b, j); c.tw1[b].im = lane_tab<N>(TAB_TW1 + 2 * b + 1, j); }
#if GSS_TW1_FULL
    expand_tw(c.tw1, c.tw1x);
#endif
}
// analysis / synthesis window (scaled) at this lane's sample positions, from the table
template <int N>
__device__ __forceinline__ void window_tab(int j, float scale, v2 (&w)[8]) {
#pragma unroll
    for (int n0 = 0; n0 < 8; ++n0) w[n0] = vmul(lane_tab<N>(TAB_WIN + n0, j), vset(scale));
}

// ---------------------------------------------------------------------------
// forward: a[n0] (lanes e=0,1: samples 2j+e + L*n0; re = sequence a, im = sequence b)
//       -> a[k2] (lanes A,B: Z[cA + L*k2], Z[cB + L*k2]).  Z[k] and Z[N-k] sit in the same
//          thread: (A[i], B[7-i]) and (B[i], A[7-i]); thread 0 (columns 0 and L/2 pair with
//          themselves): (A[i], A[8-i]), (A[0], A[4]) = (DC, Nyquist), (B[i], B[7-i]).
//          gss_stream.cuh split_pair / pack_pair select the partner accordingly.
template <int N>
__device__ __forceinline__ void fft_forward(const TeamCtx<N>& c, cv2 (&a)[8]) {
    typedef Geo<N> G;
    const int j = c.j;
    team_sync(c);   // the previous transform's last read of the exchange buffers
    dft8<false>(a);
    c.e0[j] = a[0].re; c.e0[G::E0_PLANE + j] = a[0].im;
    {
#pragma unroll
        for (int k0 = 1; k0 < 8; ++k0) {
            cv2 t = cmul<false>(a[k0], c.tw0[k0]);
            c.e0[k0 * G::P0 + j] = t.re; c.e0[G::E0_PLANE + k0 * G::P0 + j] = t.im;
        }
    }
    team_sync(c);
    if constexpr (G::M == 8) {   // middle: thread (k0 = j/4, q = j%4), lanes n2 = 2q+e, one packed DFT-8 over n1
        const int k0 = j >> 2, q = j & 3;
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            a[n1].re = c.e0[k0 * G::P0 + 4 * n1 + q];
            a[n1].im = c.e0[G::E0_PLANE + k0 * G::P0 + 4 * n1 + q];
        }
        dft8<false>(a);
        GSS_MID_TW(c, w);
#if GSS_E1_V2
        // E1 as planes of v2 = (n2 = 2q, n2 = 2q + 1) pairs, [q][column] with pitch P1: the middle pass stores its packed
        // values as they are (64-bit), bank pairs (4q + k0 + 8 k1) mod 16 are distinct over a half-warp
        v2* e1v = reinterpret_cast<v2*>(c.e1) + q * G::P1 + k0;
        e1v[0] = a[0].re; e1v[G::E1V_PLANE] = a[0].im;
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) {
            cv2 t = cmul<false>(a[k1], w[k1]);
            e1v[8 * k1] = t.re; e1v[8 * k1 + G::E1V_PLANE] = t.im;
        }
#else
        float* re0 = c.e1 + (2 * q) * G::P1 + k0;
        float* re1 = re0 + G::P1;
        re0[0] = a[0].re.x; re1[0] = a[0].re.y;
        re0[G::E1_PLANE] = a[0].im.x; re1[G::E1_PLANE] = a[0].im.y;
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) {
            cv2 t = cmul<false>(a[k1], w[k1]);
            re0[8 * k1] = t.re.x; re1[8 * k1] = t.re.y;
            re0[8 * k1 + G::E1_PLANE] = t.im.x; re1[8 * k1 + G::E1_PLANE] = t.im.y;
        }
#endif
    } else {                     // M = 4: two packed DFT-4 over n1, for k0 = j/4 and j/4 + 4
        const int kb = j >> 2, q = j & 3;
        const cv2 w3 = cmul<false>(c.tw1[0], c.tw1[1]);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int k0 = kb + 4 * g;
#pragma unroll
            for (int n1 = 0; n1 < 4; ++n1) {
                a[4 * g + n1].re = c.e0[k0 * G::P0 + 4 * n1 + q];
                a[4 * g + n1].im = c.e0[G::E0_PLANE + k0 * G::P0 + 4 * n1 + q];
            }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int k0 = kb + 4 * g;
            cv2 b[4] = {a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]};
            dft4<false>(b);
            b[1] = cmul<false>(b[1], c.tw1[0]); b[2] = cmul<false>(b[2], c.tw1[1]); b[3] = cmul<false>(b[3], w3);
            float* re0 = c.e1 + (2 * q) * G::P1 + k0;
            float* re1 = re0 + G::P1;
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) {
                re0[8 * k1] = b[k1].re.x; re1[8 * k1] = b[k1].re.y;
                re0[8 * k1 + G::E1_PLANE] = b[k1].im.x; re1[8 * k1 + G::E1_PLANE] = b[k1].im.y;
            }
        }
    }
    team_sync(c);
#if GSS_E1_V2
    if constexpr (G::M == 8) {
        const v2* e1v = reinterpret_cast<const v2*>(c.e1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {       // column loads as (n2 = 2q, 2q + 1) pairs, then a 2 x 2 transpose into (column A, column B) lanes
            const v2 ar = e1v[q * G::P1 + c.cA], br = e1v[q * G::P1 + c.cB];
            const v2 ai = e1v[G::E1V_PLANE + q * G::P1 + c.cA], bi = e1v[G::E1V_PLANE + q * G::P1 + c.cB];
            a[2 * q].re = make_float2(ar.x, br.x); a[2 * q + 1].re = make_float2(ar.y, br.y);
            a[2 * q].im = make_float2(ai.x, bi.x); a[2 * q + 1].im = make_float2(ai.y, bi.y);
        }
    } else
#endif
    {
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
            const float* p = c.e1 + n2 * G::P1;
            a[n2].re = make_float2(p[c.cA], p[c.cB]);
            a[n2].im = make_float2(p[c.cA + G::E1_PLANE], p[c.cB + G::E1_PLANE]);
        }
    }
    dft8<false>(a);
}

// inverse (unnormalised, e^{+...}): a[k2] in the last-pass layout above
//       -> a[n0] (lanes e=0,1: samples 2j+e + L*n0; re/im = the two real sequences).
// `before_last` runs after the last exchange has been read and before the final butterflies: the place to start
// loads whose latency the last pass hides (gss_tmem.cuh fetches the overlap-add state from tensor memory there).
struct NoHook { __device__ __forceinline__ void operator()() const {} };
template <int N, class Hook = NoHook>
__device__ __forceinline__ void fft_inverse(const TeamCtx<N>& c, cv2 (&a)[8], Hook before_last = Hook()) {
    typedef Geo<N> G;
    const int j = c.j;
    team_sync(c);   // the previous transform's last read of the exchange buffers
    dft8<true>(a);
#if GSS_E1_V2
    if constexpr (G::M == 8) {
        v2* e1v = reinterpret_cast<v2*>(c.e1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {       // 2 x 2 transposes into (n2 = 2q, 2q + 1) pairs per column, 64-bit stores
            e1v[q * G::P1 + c.cA] = make_float2(a[2 * q].re.x, a[2 * q + 1].re.x);
            e1v[q * G::P1 + c.cB] = make_float2(a[2 * q].re.y, a[2 * q + 1].re.y);
            e1v[G::E1V_PLANE + q * G::P1 + c.cA] = make_float2(a[2 * q].im.x, a[2 * q + 1].im.x);
            e1v[G::E1V_PLANE + q * G::P1 + c.cB] = make_float2(a[2 * q].im.y, a[2 * q + 1].im.y);
        }
    } else
#endif
    {
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
            float* p = c.e1 + n2 * G::P1;
            p[c.cA] = a[n2].re.x; p[c.cB] = a[n2].re.y;
            p[c.cA + G::E1_PLANE] = a[n2].im.x; p[c.cB + G::E1_PLANE] = a[n2].im.y;
        }
    }
    team_sync(c);
    if constexpr (G::M == 8) {
        const int k0 = j >> 2, q = j & 3;
        GSS_MID_TW(c, w);
#if GSS_E1_V2
        const v2* e1v = reinterpret_cast<const v2*>(c.e1) + q * G::P1 + k0;
        a[0].re = e1v[0]; a[0].im = e1v[G::E1V_PLANE];
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) {
            cv2 t;
            t.re = e1v[8 * k1]; t.im = e1v[8 * k1 + G::E1V_PLANE];
            a[k1] = cmul<true>(t, w[k1]);
        }
#else
        const float* re0 = c.e1 + (2 * q) * G::P1 + k0;
        const float* re1 = re0 + G::P1;
        a[0].re = make_float2(re0[0], re1[0]);
        a[0].im = make_float2(re0[G::E1_PLANE], re1[G::E1_PLANE]);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) {
            cv2 t;
            t.re = make_float2(re0[8 * k1], re1[8 * k1]);
            t.im = make_float2(re0[8 * k1 + G::E1_PLANE], re1[8 * k1 + G::E1_PLANE]);
            a[k1] = cmul<true>(t, w[k1]);
        }
#endif
        dft8<true>(a);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1)
        { c.e0[k0 * G::P0 + 4 * n1 + q] = a[n1].re; c.e0[G::E0_PLANE + k0 * G::P0 + 4 * n1 + q] = a[n1].im; }
    } else {                     // M = 4: two packed inverse DFT-4, for k0 = j/4 and j/4 + 4
        const int kb = j >> 2, q = j & 3;
        const cv2 w3 = cmul<false>(c.tw1[0], c.tw1[1]);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int k0 = kb + 4 * g;
            const float* re0 = c.e1 + (2 * q) * G::P1 + k0;
            const float* re1 = re0 + G::P1;
            cv2 b[4];
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) {
                b[k1].re = make_float2(re0[8 * k1], re1[8 * k1]);
                b[k1].im = make_float2(re0[8 * k1 + G::E1_PLANE], re1[8 * k1 + G::E1_PLANE]);
            }
            b[1] = cmul<true>(b[1], c.tw1[0]); b[2] = cmul<true>(b[2], c.tw1[1]); b[3] = cmul<true>(b[3], w3);
            dft4<true>(b);
#pragma unroll
            for (int n1 = 0; n1 < 4; ++n1)
            { c.e0[k0 * G::P0 + 4 * n1 + q] = b[n1].re; c.e0[G::E0_PLANE + k0 * G::P0 + 4 * n1 + q] = b[n1].im; }
        }
    }
    team_sync(c);
    {
        a[0].re = c.e0[j]; a[0].im = c.e0[G::E0_PLANE + j];
#pragma unroll
        for (int k0 = 1; k0 < 8; ++k0) {
            cv2 t; t.re = c.e0[k0 * G::P0 + j]; t.im = c.e0[G::E0_PLANE + k0 * G::P0 + j];
            a[k0] = cmul<true>(t, c.tw0[k0]);
        }
    }
    before_last();
    dft8<true>(a);
}

}  // namespace gss
