// gss_tmem.cuh - EXPERIMENTAL fused synthesis kernel (N = 512) that parks its per-thread streaming state in
// TENSOR MEMORY.  Selected with GSS_SYNTH_SPLIT=2 (S a multiple of 3); kept as measured evidence and as a
// cross-check path for the parity tests, not the default: see "outcome" below.
//
// mask_istft_kernel (gss_stream.cuh) keeps everything a warp carries from one frame pair to the next in
// registers - sample ring, Hann window, S overlap-add accumulator sets, the pair's spectrum, twiddles -
// 250 registers per thread, i.e. 2 warps per SM sub-partition.  Blackwell has a second 256 KB on-chip array
// next to the register file: tensor memory, 512 columns x 128 lanes x 32 bit per SM, read and written with
// tcgen05.ld / tcgen05.st.  With the 32x32b shape thread i of a warp addresses lane 32*(warp % 4) + i, so a
// column range is a private, register-like scratch array per thread that does not touch the shared-memory
// pipe.  No tensor-core instruction is issued; TMEM is used purely as a software-managed extension of the
// register file (tools/ubench5.cu: data integrity, ~85-135 cycles per dependent 16-word round trip):
//
//   columns  0..15   carried part of the raw-sample ring (KEEP slots, v2 each)
//   columns 32+16s.. overlap-add accumulator of source s (KEEP slots), s < 4
//   columns 96..127  mixture spectrum of the current frame pair (written once, read once per source)
//
// What the butterflies need all the time (twiddles, window, the transform's working set) stays in registers:
// 168 registers per thread without spills worth mentioning, 3 CTAs x 4 warps per SM instead of 2 x 4.  Masks are
// staged per (pair, source) - 2 KB stages, four deep - so that three CTAs fit the shared memory of an SM.  The
// accumulator load of a source is issued from inside the inverse transform (fft_inverse's before_last hook), so
// the last butterfly pass hides its latency; because accumulator sets are addressed by a run-time column the
// source loop needs neither unrolling nor register rotation.
//
// Outcome (B200, C2, profiles/r1c_tmem_variant.txt): bit-compatible results, 201 us at 9.3 resident warps per SM (ncu)
// against 188 us of the register-resident kernel at 7.2 and 189 us of the role-split kernel at 10.8.  Tensor-memory
// traffic itself costs ~4 %, the finer mask staging ~15 % (bisected at equal occupancy); the extra resident warps buy
// nothing back.  Together with the other occupancy experiments this is what pins the fused synthesis
// on the SM sub-partitions' issue cadence rather than on latency hiding.
#pragma once
#include "gss_stream.cuh"

#ifndef GSS_TM_MINB
#define GSS_TM_MINB 3                 // resident CTAs per SM the register allocation is held to
#endif

namespace gss {

constexpr int TM_COLS = 128;          // columns allocated per CTA (power of two >= 32); 3 CTAs x 128 <= 512
constexpr int TM_COL_RING = 0;
constexpr int TM_COL_ACC = 32;        // + 16 * s, s < 4
constexpr int TM_COL_X = 96;          // the pair's mixture spectrum (PairSpec, 32 words)

__device__ __forceinline__ void tm_st16(uint32_t taddr, const v2 (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr),
                    "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y),
                    "f"(v[4].x), "f"(v[4].y), "f"(v[5].x), "f"(v[5].y), "f"(v[6].x), "f"(v[6].y), "f"(v[7].x), "f"(v[7].y) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, v2 (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
                   "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y)
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const v2 (&u)[4], const v2 (&w)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr),
                    "f"(u[0].x), "f"(u[0].y), "f"(u[1].x), "f"(u[1].y), "f"(u[2].x), "f"(u[2].y), "f"(u[3].x), "f"(u[3].y),
                    "f"(w[0].x), "f"(w[0].y), "f"(w[1].x), "f"(w[1].y), "f"(w[2].x), "f"(w[2].y), "f"(w[3].x), "f"(w[3].y) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, v2 (&u)[4], v2 (&w)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(u[0].x), "=f"(u[0].y), "=f"(u[1].x), "=f"(u[1].y), "=f"(u[2].x), "=f"(u[2].y), "=f"(u[3].x), "=f"(u[3].y),
                   "=f"(w[0].x), "=f"(w[0].y), "=f"(w[1].x), "=f"(w[1].y), "=f"(w[2].x), "=f"(w[2].y), "=f"(w[3].x), "=f"(w[3].y)
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


template <int N>
struct TmSmem {
    static constexpr int NH = N / 2;
    static constexpr int WARPS = 4;
    static constexpr int NSTAGE = 4;                                             // masks are fetched NSTAGE - 1 sources ahead
    static constexpr int STAGE_FLOATS = 2 * NH;                                  // one pair, one source
    static constexpr int TEAM_FLOATS = Geo<N>::TEAM_FLOATS + NSTAGE * STAGE_FLOATS;   // exchange + mask stages
    static constexpr size_t bytes() { return sizeof(float) * ((size_t)WARPS * TEAM_FLOATS) + sizeof(uint64_t) * NSTAGE * WARPS + 16; }
};

template <int N, int HS, int ST>
__global__ void __launch_bounds__(128, GSS_TM_MINB) mask_istft_tm_kernel(const SynthArgs p) {
    typedef SGeo<N, HS> SG; typedef Geo<N> G; typedef TmSmem<N> SM;
    static_assert(G::TPF == 32, "one warp per transform (N = 512): a team is a TMEM lane quarter");
    static_assert(SG::KEEP <= 8 && ST <= 4, "a state array is one 16-column TMEM slot");
    constexpr int NH = N / 2;
    extern __shared__ float4 smem4[];
    float* smf = reinterpret_cast<float*>(smem4);
    const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    float* team = smf + warp * SM::TEAM_FLOATS;
    float* stage = team + G::TEAM_FLOATS;                                  // 2 x STAGE_FLOATS, 16-byte aligned
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(smf + SM::WARPS * SM::TEAM_FLOATS);
    uint64_t* bars = bars_all + SM::NSTAGE * warp;
    uint32_t* tbase_s = reinterpret_cast<uint32_t*>(bars_all + SM::NSTAGE * SM::WARPS);
    if (j == 0) {
#pragma unroll
        for (int k = 0; k < SM::NSTAGE; ++k) mbar_init(&bars[k], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tbase_s)), "n"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tbase_s;
    const uint32_t tm = tbase + ((uint32_t)(warp * 32) << 16);              // this warp's lane quarter

    const int64_t item = (int64_t)blockIdx.x * SM::WARPS + warp;
    const int64_t per_b = (int64_t)p.ngroups * p.nchunk;
    if (item < p.B * per_b) {
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
        v2 win[8];                                         // hann / N: analysis scale; synthesis rescaled at the store
        window_tab<N>(j, 1.0f / (float)N, win);
        {
            v2 z[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int s = 0; s < ST; ++s) tm_st16(tm + TM_COL_ACC + 16 * s, z);
        }

        OlaOut<SG> o;
        o.init(p.T, j, p.al_out != 0, (float)N);
        float* orow0 = p.out + (b * p.S + s0) * p.ld_out;
        const float* mrow0 = p.mask + ((b * p.S + s0) * p.T) * NH;       // source s: + s*T*NH; frame t: + t*NH
        const int64_t msrc = p.T * NH;                                   // mask stride between sources

        // one lane stages the masks of (pair q, source s) into stage cnt % NSTAGE
        const uint32_t stage_s = smem_u32(stage), bars_s = smem_u32(bars);
        auto prefetch = [&](int cnt, int q, int s) {
            if (elect_one()) {
                const uint32_t st = (uint32_t)cnt & (uint32_t)(SM::NSTAGE - 1);
                const int64_t ta = 2 * (int64_t)q;
                const uint32_t bytes = (ta + 1 < p.T ? 2 : 1) * NH * (uint32_t)sizeof(float);
                const uint32_t bar = bars_s + st * 8u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(stage_s + st * (uint32_t)(SM::STAGE_FLOATS * sizeof(float))), "l"(mrow0 + s * msrc + ta * NH), "r"(bytes), "r"(bar) : "memory");
            }
        };

        const float* row = p.wave + b * p.ld;
        const bool al = p.al_in != 0;

        // fast stretch [qa, qb): see mask_istft_kernel
        int qa, qb;
        {
            int64_t qhi = fast_hi_input<SG>(p.n);
            const int64_t hi_out = (p.T - 1) * HS + (HS < 4 ? HS : 4) - SG::ADV;
            const int64_t qo = hi_out < 0 ? -1 : hi_out / (2 * HS);
            if (qo < qhi) qhi = qo;
            if (qhi > (p.T - 2) / 2) qhi = (p.T - 2) / 2;
            constexpr int lo_base = (8 - HS) > 4 ? (8 - HS) : 4;
            qa = max(q0, (lo_base + 2 * HS - 1) / (2 * HS));
            qb = (int)(qhi + 1 < q1 ? qhi + 1 : q1);
            if (!al || !p.al_out || ns != ST || qb <= qa) { qa = q1; qb = q1; }
        }

        int64_t base = (int64_t)2 * qs * HS;
        int cnt = 0;                                       // (pair, source) iterations started so far
        int pcnt = 0, pq = qs, ps = 0;                     // prefetch cursor: iteration pcnt is (pair pq, source ps)
        auto prefetch_next = [&]() {
            if (pq < q1) {
                prefetch(pcnt, pq, ps);
                ++pcnt;
                if (++ps >= ns) { ps = 0; ++pq; }
            }
        };
#pragma unroll 1
        for (int k = 0; k < SM::NSTAGE - 1; ++k) prefetch_next();
        v2 nxt[SG::ADV];                                   // newest ADV slots of the ring (in flight from global memory)
        {
            v2 ring[SG::RS];
            load_slots<SG::L, SG::RS>(row, p.n, base, j, al, ring);
            v2 keep[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) keep[i] = i < SG::KEEP ? ring[i] : make_float2(0.f, 0.f);
            tm_st16(tm + TM_COL_RING, keep);
#pragma unroll
            for (int i = 0; i < SG::ADV; ++i) nxt[i] = ring[SG::KEEP + i];
        }
        const float* wptr = row + (base + SG::RS - 4) * SG::L + 2 * j;      // first slot the next pair adds
        float* optr = orow0 + (base - 4) * SG::L + 2 * j;                   // output slot `base` of source s0
        const float* mA = stage + ctx.cA;
        const float* mB = stage + ctx.cB;

        auto step = [&](int q, auto tag) {
            constexpr bool FAST = decltype(tag)::value;
            {
                PairSpec x;
                cv2 a[8];
                {
                    v2 keep[8];
                    tm_wait_st();
                    tm_ld16(tm + TM_COL_RING, keep);
                    tm_wait_ld();
                    v2 ring[SG::RS];
#pragma unroll
                    for (int i = 0; i < SG::RS; ++i) ring[i] = i < SG::KEEP ? keep[i] : nxt[i - SG::KEEP];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { a[i].re = vmul(ring[i], win[i]); a[i].im = vmul(ring[HS + i], win[i]); }
#pragma unroll
                    for (int i = 0; i < 8; ++i) keep[i] = i < SG::KEEP ? ring[i + SG::ADV] : make_float2(0.f, 0.f);
                    tm_st16(tm + TM_COL_RING, keep);
                }
                if (q + 1 < q1) {
                    if (FAST) load_slots_fast<SG::L, SG::ADV>(wptr, nxt);
                    else load_slots<SG::L, SG::ADV>(row, p.n, base + SG::RS, j, al, nxt);
                }
                fft_forward<N>(ctx, a);
                split_pair<N>(a, t0, x);
                tm_st16(tm + TM_COL_X, x.ar, x.ai);
                tm_st16(tm + TM_COL_X + 16, x.br, x.bi);
            }
            const bool hb = FAST || 2 * (int64_t)q + 1 < p.T;
            const bool own = FAST || q >= q0;
#pragma unroll 1
            for (int s = 0; s < ST; ++s) {
                if (!FAST && s >= ns) break;
                // the stage this overwrites was last read in the previous iteration; every lane has passed a __syncwarp since
                prefetch_next();
                PairSpec x;
                tm_wait_st();
                tm_ld16(tm + TM_COL_X, x.ar, x.ai);
                tm_ld16(tm + TM_COL_X + 16, x.br, x.bi);
                const int stg = cnt & (SM::NSTAGE - 1);
                mbar_wait(&bars[stg], (cnt / SM::NSTAGE) & 1);
                const float* ma = mA + stg * SM::STAGE_FLOATS;
                const float* mb = mB + stg * SM::STAGE_FLOATS;
                cv2 a[8];
                {
                    v2 ga[4], gb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ga[i] = make_float2(ma[SG::L * i], mb[SG::L * i]);
                        gb[i] = hb ? make_float2(ma[NH + SG::L * i], mb[NH + SG::L * i]) : make_float2(0.f, 0.f);
                    }
                    tm_wait_ld();
                    mask_pack_pair<N>(x, ga, gb, t0, a);
                }
                v2 acc[8];
                const uint32_t tacc = tm + TM_COL_ACC + 16 * s;
                fft_inverse<N>(ctx, a, [&]() { tm_wait_st(); tm_ld16(tacc, acc); });
                tm_wait_ld();
                v2 cur[SG::RS];
#pragma unroll
                for (int i = 0; i < SG::RS; ++i) cur[i] = i < SG::KEEP ? acc[i] : make_float2(0.f, 0.f);
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
                for (int i = 0; i < 8; ++i) acc[i] = i < SG::KEEP ? cur[i + SG::ADV] : make_float2(0.f, 0.f);
                tm_st16(tacc, acc);
                ++cnt;
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
            tm_wait_st();
#pragma unroll 1
            for (int s = 0; s < ns; ++s) {
                v2 acc[8];
                tm_ld16(tm + TM_COL_ACC + 16 * s, acc);
                tm_wait_ld();
#pragma unroll
                for (int i = 0; i < SG::KEEP; ++i) o.write(orow0 + s * p.ld_out, base + i, i % HS, acc[i], false);
            }
        }
        tm_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(TM_COLS) : "memory");
}

}  // namespace gss
