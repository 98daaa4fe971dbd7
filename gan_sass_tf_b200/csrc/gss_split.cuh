// gss_split.cuh - role-split fused synthesis kernel for N = 512 (experimental alternative to
// mask_istft_kernel of gss_stream.cuh, selected with gss_set_synth_variant / GSS_SYNTH_SPLIT).
//
// mask_istft_kernel gives one warp the whole pair: forward transform, S masked inverse transforms, S
// overlap-add accumulators - 249 registers per thread, 2 warps per SM sub-partition.  Here a CTA of
// 1 + ST warps walks one run of frame pairs: warp 0 (analysis role) keeps the sample ring, computes the
// mixture spectrum of a pair and hands it over through a double-buffered 4 KB shared-memory slot; warps
// 1..ST (synthesis roles, one source each) pick it up, apply their mask, run the inverse transform and keep
// their own overlap-add accumulator.  Per-thread state shrinks to <= 168 registers (3 warps per
// sub-partition, 12 per SM), a chunk covers all sources of a row, so fewer, longer chunks (halo
// overhead 9 % -> 2-4 %), at the price of the spectrum hand-over through shared memory.
// Producer / consumer hand-over: two mbarrier pairs (full: 32 arrivals, empty: 32*ns arrivals).
#pragma once
#include "gss_stream.cuh"

namespace gss {

template <int N, int ST>
struct SplitSmem {
    static constexpr int NH = N / 2;
    static constexpr int WARPS = 1 + ST;
    static constexpr int X_V2 = 16 * 32;                         // PairSpec of a pair: 16 v2 per lane
    static constexpr int MASK_FLOATS = 2 * NH;                   // one pair, one source
    // [warps x FFT exchange][2 x spectrum slot][ST x 2 x mask stage][barriers]
    static constexpr size_t bytes() {
        return sizeof(float) * ((size_t)WARPS * Geo<N>::TEAM_FLOATS + 2 * 2 * X_V2 + (size_t)ST * 2 * MASK_FLOATS)
               + sizeof(uint64_t) * (4 + 2 * ST);
    }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int N, int HS, int ST>
__global__ void __maxnreg__(168) mask_istft_split_kernel(const SynthArgs p) {
    typedef SGeo<N, HS> SG; typedef Geo<N> G; typedef SplitSmem<N, ST> SM;
    constexpr int NH = N / 2;
    extern __shared__ float4 smem4[];
    float* smf = reinterpret_cast<float*>(smem4);
    const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    float* team = smf + warp * G::TEAM_FLOATS;
    v2* xbuf = reinterpret_cast<v2*>(smf + SM::WARPS * G::TEAM_FLOATS);              // 2 slots x 16 x 32 v2
    float* mstage = smf + SM::WARPS * G::TEAM_FLOATS + 2 * 2 * SM::X_V2;              // [ST][2][2*NH]
    uint64_t* bars = reinterpret_cast<uint64_t*>(mstage + ST * 2 * SM::MASK_FLOATS);
    uint64_t* full = bars;                 // [2]
    uint64_t* empty = bars + 2;            // [2]
    uint64_t* mbars = bars + 4;            // [ST][2]

    const int64_t item = blockIdx.x;
    const int64_t per_b = (int64_t)p.ngroups * p.nchunk;
    const int64_t b = item / per_b;
    const int rem = (int)(item - b * per_b);
    const int grp = rem / p.nchunk, c = rem - grp * p.nchunk;
    const int s0 = grp * ST;
    const int ns = min(ST, p.S - s0);
    const int q0 = c * p.ppc, q1 = min(q0 + p.ppc, p.npairs);
    const int qs = max(q0 - SG::HALO, 0);
    const bool t0 = j == 0;

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 32); mbar_init(&full[1], 32);
        mbar_init(&empty[0], 32 * ns); mbar_init(&empty[1], 32 * ns);
        for (int s = 0; s < ST; ++s) { mbar_init(&mbars[2 * s], 1); mbar_init(&mbars[2 * s + 1], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp > ns) return;                 // synthesis warps without a source (S not a multiple of ST)

    TeamCtx<N> ctx;
    team_init_tab<N>(ctx, j, team);
    v2 win[8];
    window_tab<N>(j, 1.0f / (float)N, win);

    if (warp == 0) {
        // ------------------------------------------------------------------ analysis role
        const float* row = p.wave + b * p.ld;
        const bool al = p.al_in != 0;
        int qb;
        {
            int64_t qhi = fast_hi_input<SG>(p.n);
            qb = (int)(qhi + 1 < q1 ? qhi + 1 : q1);
            if (!al || qb < qs) qb = qs;
        }
        int64_t base = (int64_t)2 * qs * HS;
        v2 ring[SG::RS];
        load_slots<SG::L, SG::RS>(row, p.n, base, j, al, ring);
        const float* wptr = row + (base + SG::RS - 4) * SG::L + 2 * j;

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
            PairSpec x;
            split_pair<N>(a, t0, x);
            const int i = q - qs, buf = i & 1;
            mbar_wait(&empty[buf], (((uint32_t)i >> 1) & 1) ^ 1);       // slot free (passes at once the first time)
            v2* xs = xbuf + buf * SM::X_V2 + j;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                xs[(0 + k) * 32] = x.ar[k]; xs[(4 + k) * 32] = x.ai[k];
                xs[(8 + k) * 32] = x.br[k]; xs[(12 + k) * 32] = x.bi[k];
            }
            mbar_arrive(&full[buf]);
            base += SG::ADV;
            wptr += SG::ADV * SG::L;
        };
        int q = qs;
#pragma unroll 1
        for (; q < qb; ++q) step(q, FastTag());
#pragma unroll 1
        for (; q < q1; ++q) step(q, SlowTag());
        return;
    }

    // ---------------------------------------------------------------------- synthesis role: source s0 + warp - 1
    const int sl = warp - 1;
    OlaOut<SG> o;
    o.init(p.T, j, p.al_out != 0, (float)N);
    float* orow = p.out + (b * p.S + s0 + sl) * p.ld_out;
    const float* mrow = p.mask + ((b * p.S + s0 + sl) * p.T) * NH;
    float* stage = mstage + sl * 2 * SM::MASK_FLOATS;
    uint64_t* mb = mbars + 2 * sl;
    const uint32_t stage_s = smem_u32(stage), mb_s = smem_u32(mb);
    auto prefetch = [&](int q) {
        if (elect_one()) {
            const uint32_t st = (uint32_t)(q - qs) & 1u;
            const int64_t ta = 2 * (int64_t)q;
            const uint32_t bytes = (ta + 1 < p.T ? 2 : 1) * NH * (uint32_t)sizeof(float);
            const uint32_t bar = mb_s + st * 8u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(stage_s + st * (uint32_t)(SM::MASK_FLOATS * sizeof(float))), "l"(mrow + ta * NH), "r"(bytes), "r"(bar) : "memory");
        }
    };
    int qa, qb;
    {
        const int64_t hi_out = (p.T - 1) * HS + (HS < 4 ? HS : 4) - SG::ADV;
        int64_t qhi = hi_out < 0 ? -1 : hi_out / (2 * HS);
        if (qhi > (p.T - 2) / 2) qhi = (p.T - 2) / 2;
        constexpr int lo_base = (8 - HS) > 4 ? (8 - HS) : 4;
        qa = max(q0, (lo_base + 2 * HS - 1) / (2 * HS));
        qb = (int)(qhi + 1 < q1 ? qhi + 1 : q1);
        if (!p.al_out || qb <= qa) { qa = q1; qb = q1; }
    }
    v2 acc[SG::KEEP];
#pragma unroll
    for (int i = 0; i < SG::KEEP; ++i) acc[i] = make_float2(0.f, 0.f);
    int64_t base = (int64_t)2 * qs * HS;
    float* optr = orow + (base - 4) * SG::L + 2 * j;
    const float* mA = stage + ctx.cA;
    const float* mB = stage + ctx.cB;
    prefetch(qs);

    auto step = [&](int q, auto tag) {
        constexpr bool FAST = decltype(tag)::value;
        if (q + 1 < q1) prefetch(q + 1);            // the other stage was last read in iteration q-1
        const int i = q - qs, buf = i & 1;
        PairSpec x;
        mbar_wait(&full[buf], ((uint32_t)i >> 1) & 1);
        {
            const v2* xs = xbuf + buf * SM::X_V2 + j;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                x.ar[k] = xs[(0 + k) * 32]; x.ai[k] = xs[(4 + k) * 32];
                x.br[k] = xs[(8 + k) * 32]; x.bi[k] = xs[(12 + k) * 32];
            }
        }
        const bool hb = FAST || 2 * (int64_t)q + 1 < p.T;
        mbar_wait(&mb[buf], ((uint32_t)i >> 1) & 1);
        const float* ma = mA + buf * SM::MASK_FLOATS;
        const float* mbp = mB + buf * SM::MASK_FLOATS;
        v2 ga[4], gb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ga[k] = make_float2(ma[SG::L * k], mbp[SG::L * k]);
            gb[k] = hb ? make_float2(ma[NH + SG::L * k], mbp[NH + SG::L * k]) : make_float2(0.f, 0.f);
        }
        cv2 a[8];
        mask_pack_pair<N>(x, ga, gb, t0, a);
        mbar_arrive(&empty[buf]);                   // the spectrum slot is free again (x lives in registers now)
        fft_inverse<N>(ctx, a);
        v2 cur[SG::RS];
#pragma unroll
        for (int k = 0; k < SG::RS; ++k) cur[k] = k < SG::KEEP ? acc[k] : make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cur[k] = vfma(a[k].re, win[k], cur[k]);
            cur[HS + k] = vfma(a[k].im, win[k], cur[HS + k]);
        }
        if (FAST) {
#pragma unroll
            for (int k = 0; k < SG::ADV; ++k)
                *reinterpret_cast<float2*>(optr + k * SG::L) = vmul(cur[k], SG::CONST_NORM ? vset(o.oscale) : o.invn[k % HS]);
        } else if (q >= q0) {
#pragma unroll
            for (int k = 0; k < SG::ADV; ++k) o.write(orow, base + k, k % HS, cur[k], false);
        }
#pragma unroll
        for (int k = 0; k < SG::KEEP; ++k) acc[k] = cur[k + SG::ADV];
        base += SG::ADV;
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
        for (int k = 0; k < SG::KEEP; ++k) o.write(orow, base + k, k % HS, acc[k], false);
    }
}

}  // namespace gss
