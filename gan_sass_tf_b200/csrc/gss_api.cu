// gss_api.cu - the C ABI of libgss (include/gss_api.h): argument checks, launch
// planning and kernel dispatch.  No torch types, no allocation on the device
// entry points, all work is enqueued on the caller's stream.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <cuda_runtime.h>

#include "../../include/gss_api.h"
#include "gss_stream.cuh"
#include "gss_team.cuh"
// GSS_EXPERIMENTAL (lib/libgss_experimental.so, built beside the product library and loaded only by the tests and the
// tuning tools): the two measured-slower N = 512 synthesis variants (role-split CTAs, tensor-memory-parked state) and the
// process-wide switches gss_set_path / gss_set_synth_variant that pick kernel families.  The product library has no
// mutable global state beyond the per-device constant tables and copy-stream sets.
#ifdef GSS_EXPERIMENTAL
#include "gss_split.cuh"
#include "gss_tmem.cuh"
#endif

// The library is one source compiled in several parts (GSS_PART = 0..4, see gan_sass_tf_b200/build.py) so that
// the ~170 kernel instances build in parallel: part 0 = the C ABI, element-wise kernels, copy pipelines and the
// per-frame fallback; 1 / 2 = register-exchange streaming kernels for N = 512 / 256; 3 / 4 = team kernels for
// N <= 1024 / >= 2048.  Without GSS_PART everything is one translation unit (tools/regs.sh).  Helpers below are
// per-part copies (namespace gss_part_<k>); the state they share lives in inline variables.
#ifndef GSS_PART
#define GSS_PART 9                 // everything in one translation unit
#endif
#define GSS_HAS(p) (GSS_PART == 9 || GSS_PART == (p))
// nvcc names an anonymous namespace after the source file, so the per-part helper copies need distinct names
#define GSS_CAT_(a, b) a##b
#define GSS_CAT(a, b) GSS_CAT_(a, b)
#define GSS_NS GSS_CAT(gss_part_, GSS_PART)
#if GSS_HAS(0)
#include "gss_elem.cuh"          // non-template kernels: defined in part 0 only
#include "gss_generic.cuh"
#include "gss_resample.cuh"
#endif

namespace gss_shared {
inline thread_local std::string g_err;
inline std::atomic<int64_t> g_launches{0};
inline std::atomic<int> g_force_generic{0};
inline std::atomic<int> g_synth_variant{-1};        // -1: not chosen yet (GSS_SYNTH_SPLIT decides on first use)

// declared everywhere, each defined in exactly one part
int stream512_stft_f32(int hs, bool lg, gss::StftArgs<float> a, cudaStream_t st);
int stream512_stft_i16(int hs, bool lg, gss::StftArgs<int16_t> a, cudaStream_t st);
int stream512_istft(int hs, bool ex, gss::IstftArgs a, cudaStream_t st);
int stream512_synth(int hs, gss::SynthArgs a, cudaStream_t st);
int stream512_stft_dual(int hs, gss::StftArgs<float> a, cudaStream_t st);
int stream512_synth_feat(int hs, gss::SynthArgs a, cudaStream_t st);
int stream256_stft_dual(int hs, gss::StftArgs<float> a, cudaStream_t st);
int stream256_synth_feat(int hs, gss::SynthArgs a, cudaStream_t st);
int stream256_stft_f32(int hs, bool lg, gss::StftArgs<float> a, cudaStream_t st);
int stream256_stft_i16(int hs, bool lg, gss::StftArgs<int16_t> a, cudaStream_t st);
int stream256_istft(int hs, bool ex, gss::IstftArgs a, cudaStream_t st);
int stream256_synth(int hs, gss::SynthArgs a, cudaStream_t st);
int team_lo_stft_f32(int N, int ths, gss::team::StftArgs<float> t, cudaStream_t st);     // N <= 1024
int team_lo_stft_i16(int N, int ths, gss::team::StftArgs<int16_t> t, cudaStream_t st);
int team_lo_istft(int N, int ths, gss::team::IstftArgs t, cudaStream_t st);
int team_lo_synth(int N, int ths, gss::team::SynthArgs t, cudaStream_t st);
int team_hi_stft_f32(int N, int ths, gss::team::StftArgs<float> t, cudaStream_t st);     // N >= 2048
int team_hi_stft_i16(int N, int ths, gss::team::StftArgs<int16_t> t, cudaStream_t st);
int team_hi_istft(int N, int ths, gss::team::IstftArgs t, cudaStream_t st);
int team_hi_synth(int N, int ths, gss::team::SynthArgs t, cudaStream_t st);
int team_lo_synth_feat(int N, int ths, gss::team::SynthFeatArgs t, cudaStream_t st);
int team_hi_synth_feat(int N, int ths, gss::team::SynthFeatArgs t, cudaStream_t st);

}  // namespace gss_shared

namespace GSS_NS {
using namespace gss_shared;
using gss_shared::g_err;
using gss_shared::g_launches;
using gss_shared::g_force_generic;
using gss_shared::g_synth_variant;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(GSS_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

int after_launch(const char* name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return GSS_OK;
}

#ifdef GSS_TUNE
int tune(const char* name, int dflt) {      // tuning builds only: kernel variants picked by environment variables
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
#endif

int sm_count() {
    static int sms[64]; static std::once_flag once;
    std::call_once(once, [] { for (int& s : sms) s = 0; });
    int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) { int v = 148; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); sms[dev] = v; }
    return sms[dev];
}

// kernel families: 0 = automatic (N = 256, 512: register-exchange streaming kernels of gss_stream.cuh; other sizes:
// the shared-memory-FFT streaming kernels of gss_team.cuh; whatever those do not cover: gss_generic.cuh),
// 1 = never the N = 512 register kernels, 2 = gss_generic.cuh only (cross-check paths for the tests)
bool fast_n(int N) { return (N == 512 || N == 256) && !g_force_generic.load(std::memory_order_relaxed); }   // register-streaming kernels
// hop in slots for the team kernels (slot = N/R0 samples: 32, 64, 64, 128, 256), 0 when (N, H) is not covered
int team_hs(int N, int H) {
    if (g_force_generic.load(std::memory_order_relaxed) >= 2) return 0;
    int slot = 0, fs = 0;
    switch (N) { case 256: slot = 32; fs = 8; break; case 512: slot = 64; fs = 8; break; case 1024: slot = 64; fs = 16; break;
                 case 2048: slot = 128; fs = 16; break; case 4096: slot = 256; fs = 16; break; default: return 0; }
    if (H % slot) return 0;
    const int hs = H / slot;
    if (hs < 1 || fs % hs || fs / hs < 2 || fs / hs > 8) return 0;
    return hs;
}

bool supported_n(int N) { return N >= 64 && N <= 4096 && !(N & (N - 1)); }  // sizes without a streaming kernel: per-frame fallback

int check_nh(int N, int H, int* hs) {
    if (N < 16 || (N & (N - 1))) return fail(GSS_EUNSUPPORTED, "FFT_SIZE %d is not a power of two", N);
    if (!supported_n(N)) return fail(GSS_EUNSUPPORTED, "FFT_SIZE %d outside the supported range 64..4096", N);
    if (H * 2 == N) *hs = 4; else if (H * 4 == N) *hs = 2; else if (H * 8 == N) *hs = 1;
    else return fail(GSS_EUNSUPPORTED, "hop %d must be N/2, N/4 or N/8 (N=%d)", H, N);
    return GSS_OK;
}

int frame_count(int64_t n, int N, int H, int64_t* T, int64_t* nadd) {
    if (n < 1 || N < 2 || H < 1 || H > N) return fail(GSS_EINVAL, "frame_count: bad (n=%lld, N=%d, H=%d)", (long long)n, N, H);
    int64_t a = ((H - n % H) % H) % N;
    if (nadd) *nadd = a;
    if (T) *T = (n + a) / H + 1;
    return GSS_OK;
}

// Choose how many frame pairs one team walks.  `rows` independent streams of
// `npairs` pairs; each chunk recomputes `halo` pairs and pays ~2 pairs of set-up.
// cost = waves over the resident team slots x pairs walked per team.
gss::ChunkPlan plan_chunks(int64_t rows, int npairs, int halo, int64_t slots) {
    // the search is O(npairs) (hundreds of thousands of iterations for hour-long rows): remember the last few answers
    struct Memo { int64_t rows, slots; int npairs, halo; gss::ChunkPlan plan; bool used; };
    static thread_local Memo memo[8];
    static thread_local int next = 0;
    for (const Memo& m : memo)
        if (m.used && m.rows == rows && m.slots == slots && m.npairs == npairs && m.halo == halo) return m.plan;
    gss::ChunkPlan best{npairs, 1};
    double best_cost = 1e300;
    for (int ppc = 1; ppc <= npairs; ++ppc) {
        int nchunk = (npairs + ppc - 1) / ppc;
        int64_t items = rows * nchunk;
        int64_t waves = (items + slots - 1) / slots;
        double cost = (double)waves * (ppc + (nchunk > 1 ? halo : 0) + 2.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = gss::ChunkPlan{ppc, nchunk}; }
    }
    memo[next] = Memo{rows, slots, npairs, halo, best, true};
    next = (next + 1) % 8;
    return best;
}

// per (kernel, device), asked from the runtime once: the dynamic shared-memory opt-in and the resident CTAs per SM
inline std::mutex g_kinfo_mu;
inline std::unordered_map<uint64_t, int> g_kblocks;      // key -> resident CTAs per SM
inline std::unordered_map<uint64_t, size_t> g_ksmem;     // key -> dynamic shared memory opted in for
int kernel_key(const void* kernel, uint64_t* key) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    *key = (uint64_t)(uintptr_t)kernel * 64u + (uint64_t)(dev & 63);
    return GSS_OK;
}
int ensure_smem(const void* kernel, size_t smem) {
    if (smem <= 48 * 1024) return GSS_OK;
    uint64_t key = 0;
    if (int rc = kernel_key(kernel, &key)) return rc;
    {
        std::lock_guard<std::mutex> lk(g_kinfo_mu);
        auto it = g_ksmem.find(key);
        if (it != g_ksmem.end() && it->second >= smem) return GSS_OK;
    }
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::lock_guard<std::mutex> lk(g_kinfo_mu);
    g_ksmem[key] = smem;
    return GSS_OK;
}
int kernel_info(const void* kernel, int threads, size_t smem, int* blocks) {
    uint64_t key = 0;
    if (int rc = kernel_key(kernel, &key)) return rc;
    {
        std::lock_guard<std::mutex> lk(g_kinfo_mu);
        auto it = g_kblocks.find(key);
        if (it != g_kblocks.end()) { *blocks = it->second; return GSS_OK; }
    }
    if (int rc = ensure_smem(kernel, smem)) return rc;
    int nb = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess || nb < 1) { nb = 1; cudaGetLastError(); }
    std::lock_guard<std::mutex> lk(g_kinfo_mu);
    g_kblocks[key] = nb;
    *blocks = nb;
    return GSS_OK;
}

template <typename K>
int64_t team_slots(K kernel, int warps, int teams, size_t smem) {
    int nb = 1;
    if (kernel_info((const void*)kernel, warps * 32, smem, &nb) != GSS_OK) nb = 1;
    return (int64_t)sm_count() * nb * teams;
}

constexpr int WARPS = 4;

template <int N>
size_t team_smem(int warps) { return sizeof(float) * ((size_t)warps * gss::Geo<N>::TEAM_FLOATS); }

// Lane-constant table of the register-exchange kernels (gss_fft.cuh), filled once per device.  The fill
// runs on the stream of the first call, followed by one stream synchronisation, so the table is visible
// to every later launch on any stream; inside a stream capture it is simply recorded into the graph.
// Not counted by gss_launch_count() (set-up, not one of the path's kernels).
template <int N>
int ensure_tables(cudaStream_t st) {
    static std::atomic<int> ready[64];
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(GSS_EUNSUPPORTED, "device ordinal %d out of range", dev);
    if (ready[dev].load(std::memory_order_acquire)) return GSS_OK;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    gss::tables_kernel<N><<<1, gss::Geo<N>::TPF, 0, st>>>();
    CK(cudaGetLastError());
    if (cap != cudaStreamCaptureStatusNone) return GSS_OK;
    CK(cudaStreamSynchronize(st));
    ready[dev].store(1, std::memory_order_release);
    return GSS_OK;
}

template <typename K>
int prep(K kernel, size_t smem) { return ensure_smem((const void*)kernel, smem); }

#if GSS_HAS(0)
// ---- any-size path (gss_generic.cuh) -------------------------------------------
template <typename TIn>
int launch_stft_generic(const TIn* wave, int64_t B, int64_t n, int64_t ld, int64_t T, int N, int H, int flags, float eps,
                        float* feat, cudaStream_t st) {
    gss::gen::StftArgs a{};
    a.wave = wave; a.feat = feat; a.B = B; a.n = n; a.ld = ld; a.T = T; a.N = N; a.H = H;
    a.log = (flags & GSS_FLAG_LOG) ? 1 : 0; a.eps = eps;
    // enough CTAs for ~4 waves, at least 4 frames per CTA to amortise the twiddle table
    int64_t fpc = (B * T + (int64_t)sm_count() * 16 - 1) / ((int64_t)sm_count() * 16);
    a.fpc = (int)(fpc < 4 ? 4 : (fpc > 64 ? 64 : fpc));
    const size_t smem = sizeof(float2) * 3 * (N / 2);
    auto k = gss::gen::stft_kernel<TIn>;
    if (int rc = prep(k, smem)) return rc;
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {                  // grid.y holds 65535 rows: longer batches take several launches
        const int64_t nb = B - b0 < 65535 ? B - b0 : 65535;
        gss::gen::StftArgs c = a;
        c.wave = wave + b0 * ld; c.feat = feat + b0 * T * N; c.B = nb;
        dim3 grid((unsigned)((T + a.fpc - 1) / a.fpc), (unsigned)nb);
        k<<<grid, gss::gen::THREADS, smem, st>>>(c);
        if (int rc = after_launch("gen::stft_kernel")) return rc;
    }
    return GSS_OK;
}

template <bool FROM_WAVE>
int launch_ola_generic(gss::gen::OlaArgs a, cudaStream_t st) {
    const int R = a.N / a.H;
    a.ft = 8;
    const int span = (a.ft + R - 1) * a.H + a.N;
    const size_t smem = sizeof(float2) * 3 * (a.N / 2) + sizeof(float) * 2 * span;
    auto k = gss::gen::ola_kernel<FROM_WAVE>;
    if (int rc = prep(k, smem)) return rc;
    const int64_t hops = (a.T - 1) + R / 2;
    const int S = FROM_WAVE ? a.S : 1;
    const int64_t per = 65535 - 65535 % S;                        // grid.y holds 65535 rows; a chunk keeps whole mixtures
    for (int64_t r0 = 0; r0 < a.rows; r0 += per) {
        gss::gen::OlaArgs c = a;
        c.rows = a.rows - r0 < per ? a.rows - r0 : per;
        c.out = a.out + r0 * a.ld_out;
        if (FROM_WAVE) { c.wave = a.wave + (r0 / S) * a.ld; c.mask = a.mask + r0 * a.T * (a.N / 2); }
        else c.feat = a.feat + r0 * a.T * a.N;
        dim3 grid((unsigned)((hops + a.ft - 1) / a.ft), (unsigned)c.rows);
        k<<<grid, gss::gen::THREADS, smem, st>>>(c);
        if (int rc = after_launch("gen::ola_kernel")) return rc;
    }
    return GSS_OK;
}

#endif  // GSS_HAS(0)

// ---- team path (gss_team.cuh) -------------------------------------------------------
template <typename K>
int64_t cta_slots(K kernel, int threads, size_t smem) {
    int nb = 1;
    if (kernel_info((const void*)kernel, threads, smem, &nb) != GSS_OK) nb = 1;
    return (int64_t)sm_count() * nb;
}
template <int N> size_t team_bytes(int nbuf) { return sizeof(float2) * (size_t)nbuf * (N + N / gss::team::Plan<N>::R0); }
template <int N> constexpr int team_threads() { return N / gss::team::Plan<N>::R0; }

template <int N, int HS, typename TIn>
int team_stft(gss::team::StftArgs<TIn> a, cudaStream_t st) {
    auto k = gss::team::stft_kernel<N, HS, TIn>;
    const size_t smem = team_bytes<N>(2);
    if (int rc = prep(k, smem)) return rc;
    gss::ChunkPlan pl = plan_chunks(a.B, a.npairs, 0, cta_slots(k, team_threads<N>(), smem));
    a.ppc = pl.ppc; a.nchunk = pl.nchunk;
    k<<<(unsigned)(a.B * a.nchunk), team_threads<N>(), smem, st>>>(a);
    return after_launch("team::stft_kernel");
}
template <int N, int HS>
int team_istft(gss::team::IstftArgs a, cudaStream_t st) {
    auto k = gss::team::istft_kernel<N, HS>;
    const size_t smem = team_bytes<N>(2);
    if (int rc = prep(k, smem)) return rc;
    gss::ChunkPlan pl = plan_chunks(a.rows, a.npairs, gss::team::TGeo<N, HS>::HALO, cta_slots(k, team_threads<N>(), smem));
    a.ppc = pl.ppc; a.nchunk = pl.nchunk;
    k<<<(unsigned)(a.rows * a.nchunk), team_threads<N>(), smem, st>>>(a);
    return after_launch("team::istft_kernel");
}
template <int N, int HS, int ST>
int team_synth_st(gss::team::SynthArgs a, cudaStream_t st) {
    auto k = gss::team::mask_istft_kernel<N, HS, ST>;
    const size_t smem = team_bytes<N>(3);
    if (int rc = prep(k, smem)) return rc;
    a.ngroups = (a.S + ST - 1) / ST;
    gss::ChunkPlan pl = plan_chunks(a.B * a.ngroups, a.npairs, gss::team::TGeo<N, HS>::HALO, cta_slots(k, team_threads<N>(), smem));
    a.ppc = pl.ppc; a.nchunk = pl.nchunk;
    k<<<(unsigned)(a.B * a.ngroups * a.nchunk), team_threads<N>(), smem, st>>>(a);
    return after_launch("team::mask_istft_kernel");
}
template <int N, int HS>
int team_synth(gss::team::SynthArgs a, cudaStream_t st) {
    if (a.S % 3 == 0) return team_synth_st<N, HS, 3>(a, st);
    if (a.S % 2 == 0) return team_synth_st<N, HS, 2>(a, st);
    if (a.S == 1) return team_synth_st<N, HS, 1>(a, st);
    return team_synth_st<N, HS, 3>(a, st);
}
template <int N, int HS, int ST>
int team_synth_feat_st(gss::team::SynthFeatArgs a, cudaStream_t st) {
    auto k = gss::team::mask_istft_feat_kernel<N, HS, ST>;
    const size_t smem = gss::team::synth_feat_bytes<N>();
    if (int rc = prep(k, smem)) return rc;
    a.ngroups = (a.S + ST - 1) / ST;
    gss::ChunkPlan pl = plan_chunks(a.B * a.ngroups, a.npairs, gss::team::TGeo<N, HS>::HALO, cta_slots(k, team_threads<N>(), smem));
    a.ppc = pl.ppc; a.nchunk = pl.nchunk;
    k<<<(unsigned)(a.B * a.ngroups * a.nchunk), team_threads<N>(), smem, st>>>(a);
    return after_launch("team::mask_istft_feat_kernel");
}
template <int N, int HS>
int team_synth_feat(gss::team::SynthFeatArgs a, cudaStream_t st) {
    // one source per work item where that buys resident CTAs (gss_team.cuh, Plan<N>::FEAT_ST1)
    if constexpr (gss::team::Plan<N>::FEAT_ST1) {
        return team_synth_feat_st<N, HS, 1>(a, st);
    } else {
        if (a.S % 3 == 0) return team_synth_feat_st<N, HS, 3>(a, st);
        if (a.S % 2 == 0) return team_synth_feat_st<N, HS, 2>(a, st);
        if (a.S == 1) return team_synth_feat_st<N, HS, 1>(a, st);
        return team_synth_feat_st<N, HS, 3>(a, st);
    }
}
// F<N, HS>::run(args...) for the (N, hs) pairs the team kernels cover
#ifdef GSS_EXPERIMENTAL     // team kernels at 256 / 512 are reachable through gss_set_path(1) only: cross-check builds
#define GSS_TEAM_DISPATCH_LO(N, hs, CALL)                                                                           \
    do {                                                                                                            \
        if (N == 256 && hs == 1) { CALL(256, 1) } if (N == 256 && hs == 2) { CALL(256, 2) } if (N == 256 && hs == 4) { CALL(256, 4) }       \
        if (N == 512 && hs == 1) { CALL(512, 1) } if (N == 512 && hs == 2) { CALL(512, 2) } if (N == 512 && hs == 4) { CALL(512, 4) }       \
        if (N == 1024 && hs == 2) { CALL(1024, 2) } if (N == 1024 && hs == 4) { CALL(1024, 4) } if (N == 1024 && hs == 8) { CALL(1024, 8) } \
    } while (0)
#else
#define GSS_TEAM_DISPATCH_LO(N, hs, CALL)                                                                           \
    do {                                                                                                            \
        if (N == 1024 && hs == 2) { CALL(1024, 2) } if (N == 1024 && hs == 4) { CALL(1024, 4) } if (N == 1024 && hs == 8) { CALL(1024, 8) } \
    } while (0)
#endif
#define GSS_TEAM_DISPATCH_HI(N, hs, CALL)                                                                           \
    do {                                                                                                            \
        if (N == 2048 && hs == 2) { CALL(2048, 2) } if (N == 2048 && hs == 4) { CALL(2048, 4) } if (N == 2048 && hs == 8) { CALL(2048, 8) } \
        if (N == 4096 && hs == 2) { CALL(4096, 2) } if (N == 4096 && hs == 4) { CALL(4096, 4) } if (N == 4096 && hs == 8) { CALL(4096, 8) } \
    } while (0)

// ---- STFT ---------------------------------------------------------------
template <int N, int HS, bool LOG, typename TIn, int WARPS = 4, bool DUAL = false>
int launch_stft_w(gss::StftArgs<TIn> a, cudaStream_t st) {
    auto k = gss::stft_kernel<N, HS, LOG, TIn, WARPS, DUAL>;
    constexpr int TEAMS = WARPS * 32 / gss::Geo<N>::TPF;       // transforms in flight per CTA
    const size_t smem = team_smem<N>(TEAMS);
    if (int rc = prep(k, smem)) return rc;
    if (int rc = ensure_tables<N>(st)) return rc;
    gss::ChunkPlan pl = plan_chunks(a.B, a.npairs, 0, team_slots(k, WARPS, TEAMS, smem));
#ifdef GSS_TUNE
    if (int nc = tune("GSS_STFT_NCHUNK", 0)) { pl.nchunk = nc; pl.ppc = (a.npairs + nc - 1) / nc; pl.nchunk = (a.npairs + pl.ppc - 1) / pl.ppc; }
#endif
    a.ppc = pl.ppc; a.nchunk = pl.nchunk;
    int64_t items = a.B * a.nchunk;
    k<<<(unsigned)((items + TEAMS - 1) / TEAMS), WARPS * 32, smem, st>>>(a);
    return after_launch("stft_kernel");
}
template <int N, int HS, bool LOG, typename TIn>
int launch_stft(gss::StftArgs<TIn> a, cudaStream_t st) {
#ifdef GSS_TUNE
    if (N == 512 && HS == 2 && LOG && sizeof(TIn) == 4) {
        switch (tune("GSS_STFT_WARPS", 4)) {
            case 12: return launch_stft_w<512, 2, true, TIn, 12>(a, st);
            case 14: return launch_stft_w<512, 2, true, TIn, 14>(a, st);
            case 16: return launch_stft_w<512, 2, true, TIn, 16>(a, st);
            default: break;
        }
    }
#endif
    return launch_stft_w<N, HS, LOG, TIn, 4>(a, st);
}
#if GSS_HAS(0)
template <typename TIn>
int stft_dispatch(const TIn* wave, int64_t B, int64_t n, int64_t ld, int N, int H, int flags, float eps, float* feat, void* stream) {
    int hs = 0;
    if (int rc = check_nh(N, H, &hs)) return rc;
    if (!wave || !feat) return fail(GSS_EINVAL, "stft: null pointer");
    if (B < 0 || n < 1 || ld < n) return fail(GSS_EINVAL, "stft: bad shape B=%lld n=%lld ld=%lld", (long long)B, (long long)n, (long long)ld);
    if (n < N) return fail(GSS_EUNSUPPORTED, "stft: n=%lld < FFT_SIZE=%d (SciPy would silently shrink nperseg)", (long long)n, N);
    if (flags & ~GSS_FLAG_LOG) return fail(GSS_EINVAL, "stft: unknown flags 0x%x", flags);
    if (B == 0) return GSS_OK;
    gss::StftArgs<TIn> a{};
    a.wave = wave; a.feat = feat; a.B = B; a.n = n; a.ld = ld; a.eps = eps;
    a.al_in = ((uintptr_t)wave % (2 * sizeof(TIn)) == 0) && (ld % 2 == 0 || B == 1);
    if (int rc = frame_count(n, N, H, &a.T, nullptr)) return rc;
    a.npairs = (int)((a.T + 1) / 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (!fast_n(N)) {
        if (const int ths = team_hs(N, H)) {
            gss::team::StftArgs<TIn> t{};
            t.wave = wave; t.feat = feat; t.B = B; t.n = n; t.ld = ld; t.T = a.T; t.npairs = a.npairs;
            t.log = (flags & GSS_FLAG_LOG) ? 1 : 0; t.eps = eps;
            if constexpr (sizeof(TIn) == 4) return N <= 1024 ? team_lo_stft_f32(N, ths, t, st) : team_hi_stft_f32(N, ths, t, st);
            else return N <= 1024 ? team_lo_stft_i16(N, ths, t, st) : team_hi_stft_i16(N, ths, t, st);
        }
        return launch_stft_generic<TIn>(wave, B, n, ld, a.T, N, H, flags, eps, feat, st);
    }
    const bool lg = flags & GSS_FLAG_LOG;
    if constexpr (sizeof(TIn) == 4) return N == 512 ? stream512_stft_f32(hs, lg, a, st) : stream256_stft_f32(hs, lg, a, st);
    else return N == 512 ? stream512_stft_i16(hs, lg, a, st) : stream256_stft_i16(hs, lg, a, st);
}
#endif  // GSS_HAS(0)

// ---- iSTFT --------------------------------------------------------------
template <int N, int HS, bool EXP>
int launch_istft(gss::IstftArgs a, cudaStream_t st) {
    auto k = gss::istft_kernel<N, HS, EXP, WARPS>;
    constexpr int TEAMS = WARPS * 32 / gss::Geo<N>::TPF;
    const size_t smem = team_smem<N>(TEAMS);
    if (int rc = prep(k, smem)) return rc;
    if (int rc = ensure_tables<N>(st)) return rc;
    gss::ChunkPlan pl = plan_chunks(a.rows, a.npairs, gss::SGeo<N, HS>::HALO, team_slots(k, WARPS, TEAMS, smem));
    a.ppc = pl.ppc; a.nchunk = pl.nchunk;
    int64_t items = a.rows * a.nchunk;
    k<<<(unsigned)((items + TEAMS - 1) / TEAMS), WARPS * 32, smem, st>>>(a);
    return after_launch("istft_kernel");
}

// ---- fused synthesis ------------------------------------------------------
template <int N, int HS, int ST, int WARPS = 4, bool FEAT = false, bool AE = false>
int launch_synth_w(gss::SynthArgs a, cudaStream_t st) {
    auto k = gss::mask_istft_kernel<N, HS, ST, WARPS, FEAT, AE>;
    constexpr int TEAMS = WARPS * 32 / gss::Geo<N>::TPF;
    size_t smem = gss::SynthSmem<N, ST, FEAT>::bytes(TEAMS);
#ifdef GSS_TUNE
    smem += (size_t)tune("GSS_EXTRA_SMEM", 0);      // occupancy limiter for single-warp-per-SMSP experiments
#endif
    if (int rc = prep(k, smem)) return rc;
    if (int rc = ensure_tables<N>(st)) return rc;
    a.ngroups = (a.S + ST - 1) / ST;
#ifdef GSS_TIMING
    static long long* tbuf = nullptr;
    if (!tbuf) cudaMalloc(&tbuf, sizeof(long long) * 8 * 65536);
    a.timing = tbuf;
#endif
    gss::ChunkPlan pl = plan_chunks(a.B * a.ngroups, a.npairs, gss::SGeo<N, HS>::HALO, team_slots(k, WARPS, TEAMS, smem));
#ifdef GSS_TUNE
    if (int nc = tune("GSS_SYNTH_NCHUNK", 0)) { pl.nchunk = nc; pl.ppc = (a.npairs + nc - 1) / nc; pl.nchunk = (a.npairs + pl.ppc - 1) / pl.ppc; }
#endif
    a.ppc = pl.ppc; a.nchunk = pl.nchunk;
    int64_t items = a.B * a.ngroups * a.nchunk;
    k<<<(unsigned)((items + TEAMS - 1) / TEAMS), WARPS * 32, smem, st>>>(a);
#ifdef GSS_TIMING
    {
        static int calls = 0;
        if (++calls == 8) {     // after warm-up: dump the per-phase cycle split of this launch
            cudaStreamSynchronize(st);
            int64_t n = items < 65536 ? items : 65536;
            std::string buf(sizeof(long long) * 8 * n, 0);
            cudaMemcpy(&buf[0], tbuf, buf.size(), cudaMemcpyDeviceToHost);
            const long long* t = (const long long*)buf.data();
            double s[8] = {0}; for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 8; ++k) s[k] += t[i * 8 + k];
            fprintf(stderr, "[gss timing] items %lld pairs/item %.1f | cycles per pair: fwd %.0f maskwait %.0f mask+pack %.0f inv %.0f ola+store %.0f looptop %.0f | total %.0f\n",
                    (long long)n, s[6] / n, s[0] / s[6], s[1] / s[6], s[2] / s[6], s[3] / s[6], s[4] / s[6], s[7] / s[6],
                    (s[0] + s[1] + s[2] + s[3] + s[4] + s[7]) / s[6]);
        }
    }
#endif
    return after_launch("mask_istft_kernel");
}
template <int N, int HS, int ST>
int launch_synth(gss::SynthArgs a, cudaStream_t st) {
#ifdef GSS_TUNE
    if (N == 512 && HS == 2 && ST == 3) {
        switch (tune("GSS_SYNTH_WARPS", 4)) {
            case 9: return launch_synth_w<512, 2, 3, 9>(a, st);
            case 10: return launch_synth_w<512, 2, 3, 10>(a, st);
            case 11: return launch_synth_w<512, 2, 3, 11>(a, st);
            case 12: return launch_synth_w<512, 2, 3, 12>(a, st);
            default: break;
        }
    }
#endif
    return launch_synth_w<N, HS, ST, 4>(a, st);
}
#ifdef GSS_EXPERIMENTAL
// role-split variant (gss_split.cuh): one CTA of 1 + ST warps per (row, source group, chunk)
template <int N, int HS, int ST>
int launch_synth_split(gss::SynthArgs a, cudaStream_t st) {
    auto k = gss::mask_istft_split_kernel<N, HS, ST>;
    const size_t smem = gss::SplitSmem<N, ST>::bytes();
    if (int rc = prep(k, smem)) return rc;
    if (int rc = ensure_tables<N>(st)) return rc;
    a.ngroups = (a.S + ST - 1) / ST;
    gss::ChunkPlan pl = plan_chunks(a.B * a.ngroups, a.npairs, gss::SGeo<N, HS>::HALO, cta_slots(k, (1 + ST) * 32, smem));
    a.ppc = pl.ppc; a.nchunk = pl.nchunk;
    k<<<(unsigned)(a.B * a.ngroups * a.nchunk), (1 + ST) * 32, smem, st>>>(a);
    return after_launch("mask_istft_split_kernel");
}
// tensor-memory variant (gss_tmem.cuh): per-thread streaming state parked in TMEM, 3 CTAs x 4 warps per SM
template <int N, int HS, int ST>
int launch_synth_tm(gss::SynthArgs a, cudaStream_t st) {
    auto k = gss::mask_istft_tm_kernel<N, HS, ST>;
    constexpr int WARPS_TM = gss::TmSmem<N>::WARPS;
    const size_t smem = gss::TmSmem<N>::bytes();
    if (int rc = prep(k, smem)) return rc;
    // three CTAs need 206 KB of shared memory per SM: ask for the largest carve-out (the default heuristic left one CTA per SM resident)
    CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    if (int rc = ensure_tables<N>(st)) return rc;
    a.ngroups = (a.S + ST - 1) / ST;
    // resident slots from the launch bounds the kernel is compiled for (the occupancy query under-reports this kernel)
    int64_t slots = team_slots(k, WARPS_TM, WARPS_TM, smem);
    const int64_t by_bounds = (int64_t)sm_count() * GSS_TM_MINB * WARPS_TM;
    if (slots < by_bounds) slots = by_bounds;
    gss::ChunkPlan pl = plan_chunks(a.B * a.ngroups, a.npairs, gss::SGeo<N, HS>::HALO, slots);
    a.ppc = pl.ppc; a.nchunk = pl.nchunk;
    int64_t items = a.B * a.ngroups * a.nchunk;
    k<<<(unsigned)((items + WARPS_TM - 1) / WARPS_TM), WARPS_TM * 32, smem, st>>>(a);
    return after_launch("mask_istft_tm_kernel");
}
int synth_variant() {       // 0 = one warp per pair (gss_stream.cuh), 1 = role-split CTAs (gss_split.cuh), 2 = tensor-memory state (gss_tmem.cuh)
    int v = g_synth_variant.load(std::memory_order_relaxed);
    if (v < 0) { const char* e = getenv("GSS_SYNTH_SPLIT"); v = e ? atoi(e) : 0; if (v < 0 || v > 2) v = 0; g_synth_variant.store(v); }
    return v;
}
#endif  // GSS_EXPERIMENTAL

template <int N, int HS>
int synth_by_s(gss::SynthArgs a, cudaStream_t st) {
#ifdef GSS_EXPERIMENTAL
    if constexpr (N == 512) {
        if (synth_variant() == 1) {
            if (a.S % 3 == 0) return launch_synth_split<N, HS, 3>(a, st);
            if (a.S % 2 == 0) return launch_synth_split<N, HS, 2>(a, st);
            if (a.S == 1) return launch_synth_split<N, HS, 1>(a, st);
            return launch_synth_split<N, HS, 3>(a, st);
        }
        if (synth_variant() == 2 && a.S % 3 == 0) return launch_synth_tm<N, HS, 3>(a, st);   // other S: default kernel
    }
#endif
    // sources carried per pass: 3 when S is a multiple of 3, else 2 (S even) or 1
    if (a.S % 3 == 0) return launch_synth<N, HS, 3>(a, st);
    if (a.S % 4 == 0 && HS == 2) return launch_synth<N, HS, 4>(a, st);
    if (a.S % 2 == 0) return launch_synth<N, HS, 2>(a, st);
    if (a.S == 1) return launch_synth<N, HS, 1>(a, st);
    return launch_synth<N, HS, 3>(a, st);
}

// feature-fed synthesis (gss_mask_istft_feature): same source grouping
template <int N, int HS>
int synth_feat_by_s(gss::SynthArgs a, cudaStream_t st) {
    if (a.ae_rows) {        // fused auto-encoder partial: every source of a mixture in one pass
        CK(cudaMemsetAsync(a.ae_rows, 0, sizeof(float) * a.B, st));
        if (a.S == 3) return launch_synth_w<N, HS, 3, 4, true, true>(a, st);
        if (a.S == 2) return launch_synth_w<N, HS, 2, 4, true, true>(a, st);
        if (a.S == 1) return launch_synth_w<N, HS, 1, 4, true, true>(a, st);
        // S = 4 (the reference's default MAX_N_SIGNAL + 1, main.py:346): all four sources in one pass, in CTAs of two warps so
        // that the 29 KB of stages per warp still leave three CTAs per SM
        if constexpr (HS >= 2) { if (a.S == 4) return launch_synth_w<N, HS, 4, 2, true, true>(a, st); }
        return fail(GSS_EUNSUPPORTED, "mask_istft_feature: the fused auto-encoder partial needs S <= 4 (S <= 3 at hop N/8), got S=%d", a.S);
    }
    if (a.S % 3 == 0) return launch_synth_w<N, HS, 3, 4, true>(a, st);
    // S = 4, 8: two sources per pass (four per pass need 118 KB of shared memory per CTA with the feature stage: one CTA per SM,
    // 268.8 us against 2 x 113 us at C2 size; profiles/r2_s_sweep.txt) - four per pass only for the fused auto-encoder partial above
    if (a.S % 2 == 0) return launch_synth_w<N, HS, 2, 4, true>(a, st);
    if (a.S == 1) return launch_synth_w<N, HS, 1, 4, true>(a, st);
    return launch_synth_w<N, HS, 3, 4, true>(a, st);
}

// ---- bridges between the parts (declared in gss_shared at the top of the file) ----------------------
}  // namespace GSS_NS
namespace gss_shared {
using namespace GSS_NS;
#define GSS_STREAM_PART(NN)                                                                                          \
    int stream##NN##_stft_f32(int hs, bool lg, gss::StftArgs<float> a, cudaStream_t st) {                            \
        if (hs == 1) return lg ? launch_stft<NN, 1, true, float>(a, st) : launch_stft<NN, 1, false, float>(a, st);   \
        if (hs == 2) return lg ? launch_stft<NN, 2, true, float>(a, st) : launch_stft<NN, 2, false, float>(a, st);   \
        return lg ? launch_stft<NN, 4, true, float>(a, st) : launch_stft<NN, 4, false, float>(a, st);                \
    }                                                                                                                \
    int stream##NN##_stft_i16(int hs, bool lg, gss::StftArgs<int16_t> a, cudaStream_t st) {                          \
        if (hs == 1) return lg ? launch_stft<NN, 1, true, int16_t>(a, st) : launch_stft<NN, 1, false, int16_t>(a, st); \
        if (hs == 2) return lg ? launch_stft<NN, 2, true, int16_t>(a, st) : launch_stft<NN, 2, false, int16_t>(a, st); \
        return lg ? launch_stft<NN, 4, true, int16_t>(a, st) : launch_stft<NN, 4, false, int16_t>(a, st);            \
    }                                                                                                                \
    int stream##NN##_istft(int hs, bool ex, gss::IstftArgs a, cudaStream_t st) {                                     \
        if (hs == 1) return ex ? launch_istft<NN, 1, true>(a, st) : launch_istft<NN, 1, false>(a, st);               \
        if (hs == 2) return ex ? launch_istft<NN, 2, true>(a, st) : launch_istft<NN, 2, false>(a, st);               \
        return ex ? launch_istft<NN, 4, true>(a, st) : launch_istft<NN, 4, false>(a, st);                            \
    }                                                                                                                \
    int stream##NN##_synth(int hs, gss::SynthArgs a, cudaStream_t st) {                                              \
        if (hs == 1) return synth_by_s<NN, 1>(a, st);                                                                \
        if (hs == 2) return synth_by_s<NN, 2>(a, st);                                                                \
        return synth_by_s<NN, 4>(a, st);                                                                             \
    }                                                                                                                \
    int stream##NN##_stft_dual(int hs, gss::StftArgs<float> a, cudaStream_t st) {                                    \
        if (hs == 1) return launch_stft_w<NN, 1, true, float, 4, true>(a, st);                                       \
        if (hs == 2) return launch_stft_w<NN, 2, true, float, 4, true>(a, st);                                       \
        return launch_stft_w<NN, 4, true, float, 4, true>(a, st);                                                    \
    }                                                                                                                \
    int stream##NN##_synth_feat(int hs, gss::SynthArgs a, cudaStream_t st) {                                         \
        if (hs == 1) return synth_feat_by_s<NN, 1>(a, st);                                                           \
        if (hs == 2) return synth_feat_by_s<NN, 2>(a, st);                                                           \
        return synth_feat_by_s<NN, 4>(a, st);                                                                        \
    }
#if GSS_HAS(1)
#ifdef GSS_QUICK   // tuning builds only (tools/variant.sh): the C2 kernels alone, seconds instead of minutes to compile
int stream512_stft_f32(int hs, bool lg, gss::StftArgs<float> a, cudaStream_t st) {
    if (hs != 2) return fail(GSS_EUNSUPPORTED, "GSS_QUICK build: hop N/4 only");
    return lg ? launch_stft<512, 2, true, float>(a, st) : launch_stft<512, 2, false, float>(a, st);
}
int stream512_stft_i16(int, bool, gss::StftArgs<int16_t>, cudaStream_t) { return fail(GSS_EUNSUPPORTED, "GSS_QUICK build"); }
int stream512_istft(int hs, bool ex, gss::IstftArgs a, cudaStream_t st) {
    if (hs != 2) return fail(GSS_EUNSUPPORTED, "GSS_QUICK build: hop N/4 only");
    return ex ? launch_istft<512, 2, true>(a, st) : launch_istft<512, 2, false>(a, st);
}
int stream512_synth(int hs, gss::SynthArgs a, cudaStream_t st) {
    if (hs != 2 || a.S != 3) return fail(GSS_EUNSUPPORTED, "GSS_QUICK build: hop N/4, S = 3 only");
#ifdef GSS_EXPERIMENTAL
    if (synth_variant() == 2) return launch_synth_tm<512, 2, 3>(a, st);
#endif
    return launch_synth<512, 2, 3>(a, st);
}
int stream512_stft_dual(int hs, gss::StftArgs<float> a, cudaStream_t st) {
    if (hs != 2) return fail(GSS_EUNSUPPORTED, "GSS_QUICK build: hop N/4 only");
    return launch_stft_w<512, 2, true, float, 4, true>(a, st);
}
int stream512_synth_feat(int hs, gss::SynthArgs a, cudaStream_t st) {
    if (hs != 2 || a.S != 3) return fail(GSS_EUNSUPPORTED, "GSS_QUICK build: hop N/4, S = 3 only");
    if (a.ae_rows) { CK(cudaMemsetAsync(a.ae_rows, 0, sizeof(float) * a.B, st)); return launch_synth_w<512, 2, 3, 4, true, true>(a, st); }
    return launch_synth_w<512, 2, 3, 4, true>(a, st);
}
#else
GSS_STREAM_PART(512)
#endif
#endif
#if GSS_HAS(2)
GSS_STREAM_PART(256)
#endif
#undef GSS_STREAM_PART

#define GSS_TEAM_PART(NAME, DISPATCH)                                                                                \
    int NAME##_stft_f32(int N, int ths, gss::team::StftArgs<float> t, cudaStream_t st) {                             \
        DISPATCH(N, ths, GSS_CALL_STFT_F32); return fail(GSS_EUNSUPPORTED, "team stft: no kernel for N=%d", N); }    \
    int NAME##_stft_i16(int N, int ths, gss::team::StftArgs<int16_t> t, cudaStream_t st) {                           \
        DISPATCH(N, ths, GSS_CALL_STFT_I16); return fail(GSS_EUNSUPPORTED, "team stft: no kernel for N=%d", N); }    \
    int NAME##_istft(int N, int ths, gss::team::IstftArgs t, cudaStream_t st) {                                      \
        DISPATCH(N, ths, GSS_CALL_ISTFT); return fail(GSS_EUNSUPPORTED, "team istft: no kernel for N=%d", N); }      \
    int NAME##_synth(int N, int ths, gss::team::SynthArgs t, cudaStream_t st) {                                      \
        DISPATCH(N, ths, GSS_CALL_SYNTH); return fail(GSS_EUNSUPPORTED, "team mask_istft: no kernel for N=%d", N); }  \
    int NAME##_synth_feat(int N, int ths, gss::team::SynthFeatArgs t, cudaStream_t st) {                             \
        DISPATCH(N, ths, GSS_CALL_SYNTH_FEAT); return fail(GSS_EUNSUPPORTED, "team mask_istft_feature: no kernel for N=%d", N); }
#define GSS_CALL_STFT_F32(NN, HH) return team_stft<NN, HH, float>(t, st);
#define GSS_CALL_STFT_I16(NN, HH) return team_stft<NN, HH, int16_t>(t, st);
#define GSS_CALL_ISTFT(NN, HH) return team_istft<NN, HH>(t, st);
#define GSS_CALL_SYNTH(NN, HH) return team_synth<NN, HH>(t, st);
#define GSS_CALL_SYNTH_FEAT(NN, HH) return team_synth_feat<NN, HH>(t, st);
#if GSS_HAS(3)
GSS_TEAM_PART(team_lo, GSS_TEAM_DISPATCH_LO)
#endif
#if GSS_HAS(4)
GSS_TEAM_PART(team_hi, GSS_TEAM_DISPATCH_HI)
#endif
#undef GSS_TEAM_PART
}  // namespace gss_shared
namespace GSS_NS {

// cluster size for a row reduction: enough CTAs for about two per SM, at most 8 (the portable limit), and a row
// share of at least ~4 K elements per CTA
int cluster_for(int64_t rows, int64_t len) {
    int64_t want = (2 * (int64_t)sm_count() + rows - 1) / (rows < 1 ? 1 : rows);
    int64_t by_len = len / 4096;
    if (want > by_len) want = by_len;
    int cs = 1;
    while (cs * 2 <= want && cs < 8) cs *= 2;
    return cs;
}
template <typename... KArgs, typename... Args>
int launch_cluster(void (*kernel)(KArgs...), int64_t rows, int cs, int block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(rows * cs)); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    return GSS_OK;
}

int grid_for(int64_t work_items, int block) {
    int64_t want = (work_items + block - 1) / block;
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// grow-only device workspace for the *_host entry points
struct Workspace {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return GSS_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { p = nullptr; return fail(GSS_ENOMEM, "workspace cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
        cap = bytes; return GSS_OK;
    }
};
// one workspace pair, one private (non-blocking) stream and one mutex per device ordinal: a call on device 1 never
// touches device 0's buffers, and the *_host entry points do not serialise the device through the legacy stream
struct HostWs { std::mutex mu; Workspace in, out; cudaStream_t st = nullptr; };
HostWs g_host_ws[64];
int host_ws(HostWs** out) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(GSS_EUNSUPPORTED, "device ordinal %d out of range", dev);
    *out = &g_host_ws[dev];
    return GSS_OK;
}

// Copy engines of the host-buffer pipelines.  One H2D and one D2H stream per host thread, so that
// uploads of batch k+1 and downloads of batch k use both directions of the link at once.
// Buffers with a download in flight are remembered (device source, host destination) so that a later
// call can order itself after that copy without a host-side synchronisation.
struct CopyStreams {
    static constexpr int NTAG = 8;
    struct Tag { const void* key = nullptr; cudaEvent_t ev = nullptr; };
    std::mutex mu;                  // held while a call enqueues on this device's copy streams / edits the tags
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t order = nullptr;
    Tag src[NTAG], dst[NTAG];       // out_d buffers being read by a D2H copy / out_h buffers being written
    bool ok = false;
    int init() {
        if (ok) return GSS_OK;
        CK(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
        for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&order, cudaEventDisableTiming));
        for (auto& t : src) CK(cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming));
        for (auto& t : dst) CK(cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming));
        ok = true; return GSS_OK;
    }
    static Tag* find(Tag* tags, const void* key) {
        for (int i = 0; i < NTAG; ++i) if (tags[i].key == key) return &tags[i];
        return nullptr;
    }
    // slot for `key`; when the table is full every pending download is drained first
    int claim(Tag* tags, const void* key, Tag** out) {
        Tag* t = find(tags, key);
        if (!t) t = find(tags, nullptr);
        if (!t) {
            CK(cudaStreamSynchronize(d2h));
            for (int i = 0; i < NTAG; ++i) { src[i].key = nullptr; dst[i].key = nullptr; }
            t = &tags[0];
        }
        t->key = key; *out = t;
        return GSS_OK;
    }
};
// One set per DEVICE (not per thread): streams and events belong to the device that was current when they were
// created, and a download may be waited for by a thread other than the one that enqueued it.
CopyStreams g_cs_dev[64];
int copy_streams(CopyStreams** out) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(GSS_EUNSUPPORTED, "device ordinal %d out of range", dev);
    *out = &g_cs_dev[dev];
    return GSS_OK;
}
#define GSS_CS_LOCKED(cs)                                   \
    CopyStreams* cs##_p = nullptr;                          \
    if (int rc_ = copy_streams(&cs##_p)) return rc_;        \
    CopyStreams& cs = *cs##_p;                              \
    std::lock_guard<std::mutex> cs##_lk(cs.mu);             \
    if (int rc_ = cs.init()) return rc_

}  // namespace GSS_NS

#if GSS_HAS(0)
using namespace GSS_NS;
extern "C" {

int gss_version(void) { return 100; }
const char* gss_last_error(void) { return g_err.c_str(); }
int64_t gss_launch_count(void) { return g_launches.load(); }
#ifdef GSS_EXPERIMENTAL
int gss_set_path(int path) {
    if (path < 0 || path > 2) return fail(GSS_EINVAL, "set_path: 0 = automatic, 1 = no register-exchange kernels, 2 = gss_generic.cuh only");
    g_force_generic.store(path);
    return GSS_OK;
}
int gss_set_synth_variant(int variant) {
    if (variant < 0 || variant > 2) return fail(GSS_EINVAL, "set_synth_variant: 0 = register-resident (default), 1 = role-split CTAs, 2 = tensor-memory state");
    g_synth_variant.store(variant);
    return GSS_OK;
}
#endif  // GSS_EXPERIMENTAL
int gss_supported_fft_sizes(int* sizes, int cap) {
    int k = 0;
    for (int N = 16; N <= 65536; N *= 2)
        if (supported_n(N)) { if (sizes && k < cap) sizes[k] = N; ++k; }
    return k;
}

int gss_frame_count(int64_t n, int N, int H, int64_t* T, int64_t* nadd) { return frame_count(n, N, H, T, nadd); }

int gss_stft_packed(const float* wave, int64_t B, int64_t n, int64_t ld, int N, int H, int flags, float eps, float* feat, void* stream) {
    return stft_dispatch<float>(wave, B, n, ld, N, H, flags, eps, feat, stream);
}
int gss_stft_packed_i16(const int16_t* wave, int64_t B, int64_t n, int64_t ld, int N, int H, int flags, float eps, float* feat, void* stream) {
    return stft_dispatch<int16_t>(wave, B, n, ld, N, H, flags, eps, feat, stream);
}

int gss_istft_packed(const float* feat, int64_t R, int64_t T, int N, int H, int flags, float eps, float* wave_out, int64_t ld_out, void* stream) {
    int hs = 0;
    if (int rc = check_nh(N, H, &hs)) return rc;
    if (!feat || !wave_out) return fail(GSS_EINVAL, "istft: null pointer");
    if (R < 0 || T < 2 || ld_out < (T - 1) * H) return fail(GSS_EINVAL, "istft: bad shape R=%lld T=%lld ld_out=%lld", (long long)R, (long long)T, (long long)ld_out);
    if (flags & ~GSS_FLAG_EXP) return fail(GSS_EINVAL, "istft: unknown flags 0x%x", flags);
    if (R == 0) return GSS_OK;
    gss::IstftArgs a{};
    a.feat = feat; a.out = wave_out; a.rows = R; a.T = T; a.ld_out = ld_out; a.eps = eps;
    a.al_out = ((uintptr_t)wave_out % 8 == 0) && (ld_out % 2 == 0 || R == 1);
    a.npairs = (int)((T + 1) / 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (!fast_n(N)) {
        if (const int ths = team_hs(N, H)) {
            gss::team::IstftArgs t{};
            t.feat = feat; t.out = wave_out; t.rows = R; t.T = T; t.ld_out = ld_out; t.npairs = a.npairs;
            t.exp = (flags & GSS_FLAG_EXP) ? 1 : 0; t.eps = eps;
            return N <= 1024 ? team_lo_istft(N, ths, t, st) : team_hi_istft(N, ths, t, st);
        }
        gss::gen::OlaArgs g{};
        g.feat = feat; g.out = wave_out; g.rows = R; g.T = T; g.ld_out = ld_out; g.N = N; g.H = H; g.S = 1;
        g.exp = (flags & GSS_FLAG_EXP) ? 1 : 0; g.eps = eps;
        return launch_ola_generic<false>(g, st);
    }
    const bool ex = flags & GSS_FLAG_EXP;
    return N == 512 ? stream512_istft(hs, ex, a, st) : stream256_istft(hs, ex, a, st);
}

int gss_mask_istft(const float* wave, const float* mask, int64_t B, int S, int64_t n, int64_t ld, int N, int H,
                   float* out, int64_t ld_out, void* stream) {
    int hs = 0;
    if (int rc = check_nh(N, H, &hs)) return rc;
    if (!wave || !mask || !out) return fail(GSS_EINVAL, "mask_istft: null pointer");
    if (B < 0 || S < 1 || n < 1 || ld < n) return fail(GSS_EINVAL, "mask_istft: bad shape B=%lld S=%d n=%lld ld=%lld", (long long)B, S, (long long)n, (long long)ld);
    if (n < N) return fail(GSS_EUNSUPPORTED, "mask_istft: n=%lld < FFT_SIZE=%d", (long long)n, N);
    gss::SynthArgs a{};
    a.wave = wave; a.mask = mask; a.out = out; a.B = B; a.n = n; a.ld = ld; a.ld_out = ld_out; a.S = S;
    if ((uintptr_t)mask % 16) return fail(GSS_EINVAL, "mask_istft: mask must be 16-byte aligned (TMA bulk copies)");
    a.al_in = ((uintptr_t)wave % 8 == 0) && (ld % 2 == 0 || B == 1);
    a.al_out = ((uintptr_t)out % 8 == 0) && (ld_out % 2 == 0 || B * S == 1);
    if (int rc = frame_count(n, N, H, &a.T, nullptr)) return rc;
    if (ld_out < (a.T - 1) * H) return fail(GSS_EINVAL, "mask_istft: ld_out=%lld < (T-1)*H=%lld", (long long)ld_out, (long long)((a.T - 1) * H));
    if (B == 0) return GSS_OK;
    a.npairs = (int)((a.T + 1) / 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (!fast_n(N)) {
        if (const int ths = team_hs(N, H)) {
            gss::team::SynthArgs t{};
            t.wave = wave; t.mask = mask; t.out = out; t.B = B; t.n = n; t.ld = ld; t.T = a.T; t.ld_out = ld_out; t.S = S;
            t.npairs = a.npairs;
            return N <= 1024 ? team_lo_synth(N, ths, t, st) : team_hi_synth(N, ths, t, st);
        }
        gss::gen::OlaArgs g{};
        g.wave = wave; g.mask = mask; g.out = out; g.rows = B * S; g.n = n; g.ld = ld; g.T = a.T; g.ld_out = ld_out;
        g.N = N; g.H = H; g.S = S;
        return launch_ola_generic<true>(g, st);
    }
    return N == 512 ? stream512_synth(hs, a, st) : stream256_synth(hs, a, st);
}

int gss_stft_packed_dual(const float* wave, int64_t B, int64_t n, int64_t ld, int N, int H, float eps,
                         float* feat_lin, float* feat_log, void* stream) {
    int hs = 0;
    if (int rc = check_nh(N, H, &hs)) return rc;
    if (!wave || !feat_lin || !feat_log) return fail(GSS_EINVAL, "stft_dual: null pointer");
    if (B < 0 || n < 1 || ld < n) return fail(GSS_EINVAL, "stft_dual: bad shape B=%lld n=%lld ld=%lld", (long long)B, (long long)n, (long long)ld);
    if (n < N) return fail(GSS_EUNSUPPORTED, "stft_dual: n=%lld < FFT_SIZE=%d (SciPy would silently shrink nperseg)", (long long)n, N);
    if (B == 0) return GSS_OK;
    if (!fast_n(N)) {
        int64_t T = 0;
        if (int rc = frame_count(n, N, H, &T, nullptr)) return rc;
        if (const int ths = team_hs(N, H)) {
            gss::team::StftArgs<float> t{};
            t.wave = wave; t.feat = feat_log; t.feat_lin = feat_lin; t.B = B; t.n = n; t.ld = ld; t.T = T;
            t.npairs = (int)((T + 1) / 2); t.log = 1; t.eps = eps;
            return N <= 1024 ? team_lo_stft_f32(N, ths, t, (cudaStream_t)stream) : team_hi_stft_f32(N, ths, t, (cudaStream_t)stream);
        }
        // sizes without a streaming kernel: the plain transform, then to_log over its output (two launches)
        if (int rc = gss_stft_packed(wave, B, n, ld, N, H, 0, eps, feat_lin, stream)) return rc;
        return gss_to_log(feat_lin, feat_log, B * T, N, eps, stream);
    }
    gss::StftArgs<float> a{};
    a.wave = wave; a.feat = feat_log; a.feat_lin = feat_lin; a.B = B; a.n = n; a.ld = ld; a.eps = eps;
    a.al_in = ((uintptr_t)wave % 8 == 0) && (ld % 2 == 0 || B == 1);
    if (int rc = frame_count(n, N, H, &a.T, nullptr)) return rc;
    a.npairs = (int)((a.T + 1) / 2);
    return N == 512 ? stream512_stft_dual(hs, a, (cudaStream_t)stream) : stream256_stft_dual(hs, a, (cudaStream_t)stream);
}

int gss_mask_istft_feature(const float* feat, const float* mask, int64_t B, int S, int64_t T, int N, int H, int flags,
                           float* out, int64_t ld_out, void* stream) {
    return gss_mask_istft_feature_ae(feat, mask, B, S, T, N, H, flags, out, ld_out, nullptr, stream);
}

int gss_mask_istft_feature_ae(const float* feat, const float* mask, int64_t B, int S, int64_t T, int N, int H, int flags,
                              float* out, int64_t ld_out, float* ae_rows, void* stream) {
    int hs = 0;
    if (int rc = check_nh(N, H, &hs)) return rc;
    if (!feat || !mask || !out) return fail(GSS_EINVAL, "mask_istft_feature: null pointer");
    if (B < 0 || S < 1 || T < 2) return fail(GSS_EINVAL, "mask_istft_feature: bad shape B=%lld S=%d T=%lld", (long long)B, S, (long long)T);
    if (flags & ~GSS_FLAG_REVERSE) return fail(GSS_EINVAL, "mask_istft_feature: unknown flags 0x%x", flags);
    if (((uintptr_t)mask | (uintptr_t)feat) % 16) return fail(GSS_EINVAL, "mask_istft_feature: features and masks must be 16-byte aligned (TMA bulk copies)");
    if (ld_out < (T - 1) * H) return fail(GSS_EINVAL, "mask_istft_feature: ld_out=%lld < (T-1)*H=%lld", (long long)ld_out, (long long)((T - 1) * H));
    if (B == 0) return GSS_OK;
    if (!fast_n(N)) {
        if (ae_rows) return fail(GSS_EUNSUPPORTED, "mask_istft_feature: the fused auto-encoder partial exists for FFT_SIZE 256 / 512 only (got %d)", N);
        if (const int ths = team_hs(N, H)) {
            gss::team::SynthFeatArgs t{};
            t.feat = feat; t.mask = mask; t.out = out; t.B = B; t.T = T; t.ld_out = ld_out; t.S = S;
            t.npairs = (int)((T + 1) / 2); t.rev = (flags & GSS_FLAG_REVERSE) ? 1 : 0;
            return N <= 1024 ? team_lo_synth_feat(N, ths, t, (cudaStream_t)stream) : team_hi_synth_feat(N, ths, t, (cudaStream_t)stream);
        }
        return fail(GSS_EUNSUPPORTED, "mask_istft_feature: FFT_SIZE %d has no feature-fed synthesis kernel (256 ... 4096 do); use gss_apply_mask + gss_istft_packed", N);
    }
    gss::SynthArgs a{};
    a.feat = feat; a.mask = mask; a.out = out; a.B = B; a.T = T; a.ld_out = ld_out; a.S = S;
    a.rev = (flags & GSS_FLAG_REVERSE) ? 1 : 0;
    a.ae_rows = ae_rows;
    a.al_in = 1;
    a.al_out = ((uintptr_t)out % 8 == 0) && (ld_out % 2 == 0 || B * S == 1);
    a.npairs = (int)((T + 1) / 2);
    return N == 512 ? stream512_synth_feat(hs, a, (cudaStream_t)stream) : stream256_synth_feat(hs, a, (cudaStream_t)stream);
}

int gss_apply_mask(const float* mix, const float* mask, int64_t B, int S, int64_t T, int N, float* out, void* stream) {
    if (!mix || !mask || !out) return fail(GSS_EINVAL, "apply_mask: null pointer");
    if (N < 8 || N % 8) return fail(GSS_EINVAL, "apply_mask: N=%d must be a multiple of 8", N);
    if (B < 0 || S < 1 || T < 0) return fail(GSS_EINVAL, "apply_mask: bad shape");
    if (((uintptr_t)mix | (uintptr_t)mask | (uintptr_t)out) & 15) return fail(GSS_EINVAL, "apply_mask: buffers must be 16-byte aligned");
    int64_t total = B * S * T * (N / 8);
    if (total == 0) return GSS_OK;
    gss::apply_mask_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mix, mask, out, B, S, T, N);
    return after_launch("apply_mask_kernel");
}

int gss_ola_norm_scale(const float* in, float* out, int64_t rows, int64_t len, int64_t ld_in, int64_t ld_out, int64_t T, int N, int H,
                       int inverse, float scale, void* stream) {
    int hs = 0;
    if (int rc = check_nh(N, H, &hs)) return rc;
    if (!in || !out) return fail(GSS_EINVAL, "ola_norm_scale: null pointer");
    if (rows < 0 || len < 0 || ld_in < len || ld_out < len || T < 1) return fail(GSS_EINVAL, "ola_norm_scale: bad shape");
    if (len > (T - 1) * (int64_t)H) return fail(GSS_EINVAL, "ola_norm_scale: len=%lld exceeds the iSTFT length (T-1)*H=%lld", (long long)len, (long long)((T - 1) * (int64_t)H));
    int64_t total = rows * len;
    if (total == 0) return GSS_OK;
    if (inverse) gss::ola_norm_scale_kernel<true><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, len, ld_in, ld_out, T, N, H, scale);
    else gss::ola_norm_scale_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, len, ld_in, ld_out, T, N, H, scale);
    return after_launch("ola_norm_scale_kernel");
}

int gss_scale_packed(const float* in, float* out, int64_t rows, int N, float c_all, float c_edge, void* stream) {
    if (!in || !out) return fail(GSS_EINVAL, "scale_packed: null pointer");
    if (N < 8 || N % 8) return fail(GSS_EINVAL, "scale_packed: N=%d must be a multiple of 8", N);
    if (rows < 0) return fail(GSS_EINVAL, "scale_packed: rows < 0");
    if (((uintptr_t)in | (uintptr_t)out) & 15) return fail(GSS_EINVAL, "scale_packed: buffers must be 16-byte aligned");
    int64_t total = rows * (N / 4);
    if (total == 0) return GSS_OK;
    gss::scale_packed_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, N, c_all, c_edge);
    return after_launch("scale_packed_kernel");
}

static int logexp(bool ex, const float* in, float* out, int64_t rows, int N, float eps, void* stream) {
    if (!in || !out) return fail(GSS_EINVAL, "to_log/to_exp: null pointer");
    if (N < 8 || N % 8) return fail(GSS_EINVAL, "to_log/to_exp: N=%d must be a multiple of 8", N);
    if (rows < 0) return fail(GSS_EINVAL, "to_log/to_exp: rows < 0");
    if (((uintptr_t)in | (uintptr_t)out) & 15) return fail(GSS_EINVAL, "to_log/to_exp: buffers must be 16-byte aligned");
    int64_t total = rows * (N / 8);
    if (total == 0) return GSS_OK;
    if (ex) gss::logexp_kernel<true><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, N, eps);
    else gss::logexp_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, N, eps);
    return after_launch("logexp_kernel");
}
int gss_to_log(const float* in, float* out, int64_t rows, int N, float eps, void* stream) { return logexp(false, in, out, rows, N, eps, stream); }
int gss_to_exp(const float* in, float* out, int64_t rows, int N, float eps, void* stream) { return logexp(true, in, out, rows, N, eps, stream); }

int gss_mix_features(const float* src, const float* noise, int64_t B, int n_sig, int64_t T, int N, int flags, float eps,
                     float* mix, float* mix_log, void* stream) {
    if (!src) return fail(GSS_EINVAL, "mix_features: null pointer");
    if (N < 8 || N % 8) return fail(GSS_EINVAL, "mix_features: N=%d must be a multiple of 8", N);
    if (B < 0 || n_sig < 1 || T < 0) return fail(GSS_EINVAL, "mix_features: bad shape");
    if (flags & ~GSS_FLAG_LOG) return fail(GSS_EINVAL, "mix_features: unknown flags 0x%x", flags);
    const bool lg = flags & GSS_FLAG_LOG;
    if (lg ? !mix_log : !mix) return fail(GSS_EINVAL, "mix_features: output pointer missing");
    if (((uintptr_t)src | (uintptr_t)noise | (uintptr_t)mix | (uintptr_t)mix_log) & 15) return fail(GSS_EINVAL, "mix_features: buffers must be 16-byte aligned");
    const int64_t total = B * T * (N / 8);
    if (total == 0) return GSS_OK;
    if (lg) gss::mix_kernel<true><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, noise, mix, mix_log, B, n_sig, T, N, eps);
    else gss::mix_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, noise, mix, nullptr, B, n_sig, T, N, eps);
    return after_launch("mix_kernel");
}

static int logexp_bwd(bool ex, const float* in, const float* gout, float* gin, int64_t rows, int N, float eps, void* stream) {
    if (!in || !gout || !gin) return fail(GSS_EINVAL, "to_log/to_exp backward: null pointer");
    if (N < 8 || N % 8) return fail(GSS_EINVAL, "to_log/to_exp backward: N=%d must be a multiple of 8", N);
    if (rows < 0) return fail(GSS_EINVAL, "to_log/to_exp backward: rows < 0");
    if (((uintptr_t)in | (uintptr_t)gout | (uintptr_t)gin) & 15) return fail(GSS_EINVAL, "to_log/to_exp backward: buffers must be 16-byte aligned");
    const int64_t total = rows * (N / 8);
    if (total == 0) return GSS_OK;
    if (ex) gss::logexp_bwd_kernel<true><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, gout, gin, rows, N, eps);
    else gss::logexp_bwd_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, gout, gin, rows, N, eps);
    return after_launch("logexp_bwd_kernel");
}
int gss_to_log_bwd(const float* in, const float* gout, float* gin, int64_t rows, int N, float eps, void* stream) { return logexp_bwd(false, in, gout, gin, rows, N, eps, stream); }
int gss_to_exp_bwd(const float* in, const float* gout, float* gin, int64_t rows, int N, float eps, void* stream) { return logexp_bwd(true, in, gout, gin, rows, N, eps, stream); }

int gss_apply_mask_bwd(const float* mix, const float* mask, const float* gout, int64_t B, int S, int64_t T, int N,
                       float* gmix, float* gmask, void* stream) {
    if (!mix || !mask || !gout || (!gmix && !gmask)) return fail(GSS_EINVAL, "apply_mask_bwd: null pointer");
    if (N < 8 || N % 8) return fail(GSS_EINVAL, "apply_mask_bwd: N=%d must be a multiple of 8", N);
    if (B < 0 || S < 1 || T < 0) return fail(GSS_EINVAL, "apply_mask_bwd: bad shape");
    if (((uintptr_t)mix | (uintptr_t)mask | (uintptr_t)gout | (uintptr_t)gmix | (uintptr_t)gmask) & 15) return fail(GSS_EINVAL, "apply_mask_bwd: buffers must be 16-byte aligned");
    const int64_t total = B * T * (N / 8);
    if (total == 0) return GSS_OK;
    gss::apply_mask_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mix, mask, gout, gmix, gmask, B, S, T, N);
    return after_launch("apply_mask_bwd_kernel");
}

int gss_cross_snr(const float* clear, const float* noisy, int64_t B, int m, int n, int64_t L, float eps, float* snr, void* stream) {
    if (!clear || !noisy || !snr) return fail(GSS_EINVAL, "cross_snr: null pointer");
    if (B < 0 || m < 1 || n < 1 || L < 1) return fail(GSS_EINVAL, "cross_snr: bad shape");
    if (B == 0) return GSS_OK;
    if (int rc = launch_cluster(gss::cross_snr_kernel, B * m, cluster_for(B * m, L), 256, (cudaStream_t)stream, clear, noisy, m, n, L, eps, snr)) return rc;
    return after_launch("cross_snr_kernel");
}

int gss_ae_partial(const float* sep, const float* mix, int64_t B, int S, int64_t L, float* partial, void* stream) {
    if (!sep || !mix || !partial) return fail(GSS_EINVAL, "ae_partial: null pointer");
    if (B < 0 || S < 1 || L < 1) return fail(GSS_EINVAL, "ae_partial: bad shape");
    if (B == 0) return GSS_OK;
    if (int rc = launch_cluster(gss::ae_partial_kernel, B, cluster_for(B, L), 512, (cudaStream_t)stream, sep, mix, S, L, partial)) return rc;
    return after_launch("ae_partial_kernel");
}

int gss_metric_finalise(const float* ae_rows, const float* snr, int64_t B, int m, int n, double elems_per_row, float* vec4, void* stream) {
    if (!vec4 || (!ae_rows && !snr)) return fail(GSS_EINVAL, "metric_finalise: null pointer");
    if (B < 0 || (snr && (m < 1 || n < 1)) || elems_per_row <= 0) return fail(GSS_EINVAL, "metric_finalise: bad shape");
    gss::metric_finalise_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ae_rows, snr, B, m, n, (float)(1.0 / elems_per_row), vec4);
    return after_launch("metric_finalise_kernel");
}

int gss_gather_rows_i16(const int16_t* flat, const int64_t* offsets, const int64_t* lengths, const int64_t* idx,
                        int64_t B, int64_t ld, int16_t* out, void* stream) {
    if (!flat || !offsets || !lengths || !idx || !out) return fail(GSS_EINVAL, "gather_rows_i16: null pointer");
    if (B < 0 || ld < 1) return fail(GSS_EINVAL, "gather_rows_i16: bad shape");
    if (B == 0) return GSS_OK;
    gss::gather_rows_i16_kernel<<<grid_for(B * ld, 256), 256, 0, (cudaStream_t)stream>>>(flat, offsets, lengths, idx, B, ld, out);
    return after_launch("gather_rows_i16_kernel");
}

// ---- A14: scipy.signal.resample (main.py:89-95) ------------------------------------------------------
static int64_t resample_m(int64_t n, int64_t num) { return gss::rs::pow2_at_least(2 * (n > num ? n : num) - 1); }
size_t gss_resample_workspace_bytes(int64_t n, int64_t num) {
    if (n < 1 || num < 1 || n > ((int64_t)1 << 30) || num > ((int64_t)1 << 30)) return 0;
    return sizeof(double2) * (size_t)(4 * resample_m(n, num) + n + 2 * num);
}
// power-of-two Stockham FFT, ping-pong between p and q; returns the buffer that holds the result
static double2* rs_fft(double2* p, double2* q, int64_t M, int sign, cudaStream_t st) {
    const int grid = grid_for(M / 2, 256);
    for (int64_t Ns = 1; Ns < M; Ns <<= 1) {
        gss::rs::fft_pass_kernel<<<grid, 256, 0, st>>>(p, q, M, Ns, sign);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        double2* t = p; p = q; q = t;
    }
    return p;
}
// length-L DFT (sign -1) or unnormalised inverse DFT (sign +1) of `in` (real or complex) into out[L], by Bluestein
static int rs_bluestein(const double* in_re, const double2* in_c, int64_t L, int sign, double2* out,
                        double2* a0, double2* a1, double2* b0, double2* b1, cudaStream_t st) {
    const int64_t M = gss::rs::pow2_at_least(2 * L - 1);
    const int gm = grid_for(M, 256);
    if (in_re) gss::rs::prepare_kernel<true><<<gm, 256, 0, st>>>(in_re, nullptr, L, M, sign, a0, b0);
    else gss::rs::prepare_kernel<false><<<gm, 256, 0, st>>>(nullptr, in_c, L, M, sign, a0, b0);
    if (int rc = after_launch("rs::prepare_kernel")) return rc;
    double2* pa = rs_fft(a0, a1, M, -1, st);
    double2* pb = rs_fft(b0, b1, M, -1, st);
    gss::rs::pointwise_kernel<<<gm, 256, 0, st>>>(pa, pb, M);
    if (int rc = after_launch("rs::pointwise_kernel")) return rc;
    double2* pc = rs_fft(pa, pa == a0 ? a1 : a0, M, +1, st);
    gss::rs::finish_kernel<<<grid_for(L, 256), 256, 0, st>>>(pc, L, M, sign, out);
    return after_launch("rs::finish_kernel");
}
int gss_resample_f64(const double* x, int64_t rows, int64_t n, int64_t ld, int64_t num, double* y, int64_t ld_y,
                     void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !y || !workspace) return fail(GSS_EINVAL, "resample: null pointer");
    if (rows < 0 || n < 1 || num < 1 || ld < n || ld_y < num) return fail(GSS_EINVAL, "resample: bad shape rows=%lld n=%lld num=%lld", (long long)rows, (long long)n, (long long)num);
    const size_t need = gss_resample_workspace_bytes(n, num);
    if (!need) return fail(GSS_EUNSUPPORTED, "resample: lengths above 2^30 are not supported");
    if (workspace_bytes < need || ((uintptr_t)workspace & 15)) return fail(GSS_EINVAL, "resample: workspace of %zu bytes (16-byte aligned) needed, got %zu", need, workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M = resample_m(n, num);
    double2* a0 = (double2*)workspace; double2* a1 = a0 + M; double2* b0 = a1 + M; double2* b1 = b0 + M;
    double2* X = b1 + M; double2* Y = X + n; double2* Z = Y + num;
    for (int64_t r = 0; r < rows; ++r) {
        if (int rc = rs_bluestein(x + r * ld, nullptr, n, -1, X, a0, a1, b0, b1, st)) return rc;          // rfft (all n bins)
        gss::rs::resize_kernel<<<grid_for(num, 256), 256, 0, st>>>(X, n, num, Y);
        if (int rc = after_launch("rs::resize_kernel")) return rc;
        if (int rc = rs_bluestein(nullptr, Y, num, +1, Z, a0, a1, b0, b1, st)) return rc;                // irfft (Hermitian spectrum)
        gss::rs::real_scale_kernel<<<grid_for(num, 256), 256, 0, st>>>(Z, num, 1.0 / (double)n, y + r * ld_y);   // (1/num) * (num/n)
        if (int rc = after_launch("rs::real_scale_kernel")) return rc;
    }
    return GSS_OK;
}

int gss_wav16_normalise(const float* x, int64_t R, int64_t len, int64_t ld, float* minmax, int16_t* pcm, void* stream) {
    if (!x || !minmax || !pcm) return fail(GSS_EINVAL, "wav16: null pointer");
    if (R < 0 || len < 1 || ld < len) return fail(GSS_EINVAL, "wav16: bad shape");
    if (R == 0) return GSS_OK;
    if (int rc = launch_cluster(gss::minmax_kernel, R, cluster_for(R, len), 512, (cudaStream_t)stream, x, len, ld, minmax)) return rc;
    if (int rc = after_launch("minmax_kernel")) return rc;
    gss::wav16_kernel<<<grid_for(R * len, 256), 256, 0, (cudaStream_t)stream>>>(x, R, len, ld, minmax, pcm);
    return after_launch("wav16_kernel");
}

// ---- host-buffer entry points ------------------------------------------------
int gss_stft_packed_host(const float* wave_h, int64_t B, int64_t n, int N, int H, int flags, float eps, float* feat_h) {
    int64_t T = 0;
    if (int rc = frame_count(n, N, H, &T, nullptr)) return rc;
    if (!wave_h || !feat_h) return fail(GSS_EINVAL, "stft_host: null pointer");
    if (B <= 0) return B == 0 ? GSS_OK : fail(GSS_EINVAL, "stft_host: B < 0");
    HostWs* ws = nullptr;
    if (int rc = host_ws(&ws)) return rc;
    std::lock_guard<std::mutex> lk(ws->mu);
    if (!ws->st) CK(cudaStreamCreateWithFlags(&ws->st, cudaStreamNonBlocking));
    const size_t bi = sizeof(float) * B * n, bo = sizeof(float) * B * T * N;
    if (int rc = ws->in.reserve(bi)) return rc;
    if (int rc = ws->out.reserve(bo)) return rc;
    CK(cudaMemcpyAsync(ws->in.p, wave_h, bi, cudaMemcpyHostToDevice, ws->st));
    if (int rc = gss_stft_packed((const float*)ws->in.p, B, n, n, N, H, flags, eps, (float*)ws->out.p, ws->st)) return rc;
    CK(cudaMemcpyAsync(feat_h, ws->out.p, bo, cudaMemcpyDeviceToHost, ws->st));
    CK(cudaStreamSynchronize(ws->st));
    return GSS_OK;
}

int gss_istft_packed_host(const float* feat_h, int64_t R, int64_t T, int N, int H, int flags, float eps, float* wave_h) {
    if (!feat_h || !wave_h) return fail(GSS_EINVAL, "istft_host: null pointer");
    if (R <= 0) return R == 0 ? GSS_OK : fail(GSS_EINVAL, "istft_host: R < 0");
    if (T < 2 || H < 1) return fail(GSS_EINVAL, "istft_host: bad shape");
    HostWs* ws = nullptr;
    if (int rc = host_ws(&ws)) return rc;
    std::lock_guard<std::mutex> lk(ws->mu);
    if (!ws->st) CK(cudaStreamCreateWithFlags(&ws->st, cudaStreamNonBlocking));
    const int64_t len = (T - 1) * H;
    const size_t bi = sizeof(float) * R * T * N, bo = sizeof(float) * R * len;
    if (int rc = ws->in.reserve(bi)) return rc;
    if (int rc = ws->out.reserve(bo)) return rc;
    CK(cudaMemcpyAsync(ws->in.p, feat_h, bi, cudaMemcpyHostToDevice, ws->st));
    if (int rc = gss_istft_packed((const float*)ws->in.p, R, T, N, H, flags, eps, (float*)ws->out.p, len, ws->st)) return rc;
    CK(cudaMemcpyAsync(wave_h, ws->out.p, bo, cudaMemcpyDeviceToHost, ws->st));
    CK(cudaStreamSynchronize(ws->st));
    return GSS_OK;
}

// Chunked pipelines: chunk k's copy overlaps chunk k-1's kernel.  The *_async forms only enqueue
// (uploads on the library's H2D stream, kernels on the caller's stream, downloads on the library's
// D2H stream) and return; gss_wait_host() blocks until a download into a host buffer has landed.
int gss_stft_h2d_async(const float* wave_h, float* wave_d, int64_t B, int64_t n, int64_t ld, int N, int H, int flags, float eps,
                       float* feat_d, int chunks, void* stream) {
    int64_t T = 0;
    if (int rc = frame_count(n, N, H, &T, nullptr)) return rc;
    if (!wave_h || !wave_d || !feat_d) return fail(GSS_EINVAL, "stft_h2d: null pointer");
    if (chunks < 1 || ld < n) return fail(GSS_EINVAL, "stft_h2d: bad chunks/ld");
    GSS_CS_LOCKED(cs);
    cudaStream_t st = (cudaStream_t)stream;
    // the upload may overwrite a wave_d that kernels already enqueued on `stream` still read
    CK(cudaEventRecord(cs.order, st));
    CK(cudaStreamWaitEvent(cs.h2d, cs.order, 0));
    const int64_t per = (B + chunks - 1) / chunks;
    int k = 0;
    for (int64_t b0 = 0; b0 < B; b0 += per, ++k) {
        const int64_t nb = (B - b0 < per) ? B - b0 : per;
        if (ld == n) CK(cudaMemcpyAsync(wave_d + b0 * ld, wave_h + b0 * n, sizeof(float) * n * nb, cudaMemcpyHostToDevice, cs.h2d));   // pitch == width: one linear copy
        else CK(cudaMemcpy2DAsync(wave_d + b0 * ld, sizeof(float) * ld, wave_h + b0 * n, sizeof(float) * n, sizeof(float) * n, nb,
                                  cudaMemcpyHostToDevice, cs.h2d));
        cudaEvent_t ev = cs.ev[k % 8];
        CK(cudaEventRecord(ev, cs.h2d));
        CK(cudaStreamWaitEvent(st, ev, 0));
        if (int rc = gss_stft_packed(wave_d + b0 * ld, nb, n, ld, N, H, flags, eps, feat_d + b0 * T * N, stream)) return rc;
    }
    return GSS_OK;
}

int gss_stft_h2d(const float* wave_h, float* wave_d, int64_t B, int64_t n, int64_t ld, int N, int H, int flags, float eps,
                 float* feat_d, int chunks, void* stream) {
    if (int rc = gss_stft_h2d_async(wave_h, wave_d, B, n, ld, N, H, flags, eps, feat_d, chunks, stream)) return rc;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return GSS_OK;
}

int gss_mask_istft_d2h_async(const float* wave_d, const float* mask_d, int64_t B, int S, int64_t n, int64_t ld, int N, int H,
                             float* out_d, float* out_h, int64_t ld_out, int chunks, void* stream) {
    int64_t T = 0;
    if (int rc = frame_count(n, N, H, &T, nullptr)) return rc;
    if (!out_h || !out_d) return fail(GSS_EINVAL, "mask_istft_d2h: null pointer");
    if (chunks < 1) return fail(GSS_EINVAL, "mask_istft_d2h: bad chunks");
    GSS_CS_LOCKED(cs);
    cudaStream_t st = (cudaStream_t)stream;
    // a download that still reads this out_d (an earlier batch through the same workspace) goes first
    if (CopyStreams::Tag* t = CopyStreams::find(cs.src, out_d)) CK(cudaStreamWaitEvent(st, t->ev, 0));
    const int64_t len = (T - 1) * H;
    const int64_t per = (B + chunks - 1) / chunks;
    int k = 0;
    for (int64_t b0 = 0; b0 < B; b0 += per, ++k) {
        const int64_t nb = (B - b0 < per) ? B - b0 : per;
        if (int rc = gss_mask_istft(wave_d + b0 * ld, mask_d + b0 * S * T * (N / 2), nb, S, n, ld, N, H,
                                    out_d + b0 * S * ld_out, ld_out, stream)) return rc;
        cudaEvent_t ev = cs.ev[k % 8];
        CK(cudaEventRecord(ev, st));
        CK(cudaStreamWaitEvent(cs.d2h, ev, 0));
        if (ld_out == len) CK(cudaMemcpyAsync(out_h + b0 * S * len, out_d + b0 * S * ld_out, sizeof(float) * len * nb * S, cudaMemcpyDeviceToHost, cs.d2h));
        else CK(cudaMemcpy2DAsync(out_h + b0 * S * len, sizeof(float) * len, out_d + b0 * S * ld_out, sizeof(float) * ld_out,
                                  sizeof(float) * len, nb * S, cudaMemcpyDeviceToHost, cs.d2h));
    }
    CopyStreams::Tag* ts = nullptr; CopyStreams::Tag* td = nullptr;
    if (int rc = cs.claim(cs.src, out_d, &ts)) return rc;
    if (int rc = cs.claim(cs.dst, out_h, &td)) return rc;
    CK(cudaEventRecord(ts->ev, cs.d2h));
    CK(cudaEventRecord(td->ev, cs.d2h));
    return GSS_OK;
}

// int16 PCM on the host link in both directions (what the WAV files of main.py:83 / :116 hold): half the
// bytes of the float32 forms.  Upload: pcm_h -> pcm_d (int16) -> features straight from the int16 samples
// (gss_stft_packed_i16) and a float32 copy wave_d for the synthesis stage.  Download: separated waveforms ->
// per-row min/max normalisation to int16 (main.py:112-116) -> pcm_h.
int gss_stft_h2d_i16_async(const int16_t* pcm_h, int16_t* pcm_d, float* wave_d, int64_t B, int64_t n, int64_t ld, int N, int H,
                           int flags, float eps, float* feat_d, int chunks, void* stream) {
    int64_t T = 0;
    if (int rc = frame_count(n, N, H, &T, nullptr)) return rc;
    if (!pcm_h || !pcm_d || !wave_d || !feat_d) return fail(GSS_EINVAL, "stft_h2d_i16: null pointer");
    if (chunks < 1 || ld < n) return fail(GSS_EINVAL, "stft_h2d_i16: bad chunks/ld");
    GSS_CS_LOCKED(cs);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaEventRecord(cs.order, st));
    CK(cudaStreamWaitEvent(cs.h2d, cs.order, 0));
    const int64_t per = (B + chunks - 1) / chunks;
    int k = 0;
    for (int64_t b0 = 0; b0 < B; b0 += per, ++k) {
        const int64_t nb = (B - b0 < per) ? B - b0 : per;
        if (ld == n) CK(cudaMemcpyAsync(pcm_d + b0 * ld, pcm_h + b0 * n, sizeof(int16_t) * n * nb, cudaMemcpyHostToDevice, cs.h2d));
        else CK(cudaMemcpy2DAsync(pcm_d + b0 * ld, sizeof(int16_t) * ld, pcm_h + b0 * n, sizeof(int16_t) * n, sizeof(int16_t) * n, nb,
                                  cudaMemcpyHostToDevice, cs.h2d));
        cudaEvent_t ev = cs.ev[k % 8];
        CK(cudaEventRecord(ev, cs.h2d));
        CK(cudaStreamWaitEvent(st, ev, 0));
        if (int rc = gss_stft_packed_i16(pcm_d + b0 * ld, nb, n, ld, N, H, flags, eps, feat_d + b0 * T * N, stream)) return rc;
        gss::i16_to_f32_kernel<<<grid_for(nb * ld, 256), 256, 0, st>>>(pcm_d + b0 * ld, wave_d + b0 * ld, nb * ld);
        if (int rc = after_launch("i16_to_f32_kernel")) return rc;
    }
    return GSS_OK;
}

int gss_mask_istft_d2h_pcm16_async(const float* wave_d, const float* mask_d, int64_t B, int S, int64_t n, int64_t ld, int N, int H,
                                   float* out_d, float* minmax_d, int16_t* pcm_d, int16_t* pcm_h, int64_t ld_out, int chunks,
                                   void* stream) {
    int64_t T = 0;
    if (int rc = frame_count(n, N, H, &T, nullptr)) return rc;
    if (!out_d || !minmax_d || !pcm_d || !pcm_h) return fail(GSS_EINVAL, "mask_istft_d2h_pcm16: null pointer");
    if (chunks < 1) return fail(GSS_EINVAL, "mask_istft_d2h_pcm16: bad chunks");
    GSS_CS_LOCKED(cs);
    cudaStream_t st = (cudaStream_t)stream;
    // a download that still reads this pcm_d (an earlier batch through the same workspace) goes first
    if (CopyStreams::Tag* t = CopyStreams::find(cs.src, pcm_d)) CK(cudaStreamWaitEvent(st, t->ev, 0));
    const int64_t len = (T - 1) * H;
    const int64_t per = (B + chunks - 1) / chunks;
    int k = 0;
    for (int64_t b0 = 0; b0 < B; b0 += per, ++k) {
        const int64_t nb = (B - b0 < per) ? B - b0 : per;
        if (int rc = gss_mask_istft(wave_d + b0 * ld, mask_d + b0 * S * T * (N / 2), nb, S, n, ld, N, H,
                                    out_d + b0 * S * ld_out, ld_out, stream)) return rc;
        if (int rc = gss_wav16_normalise(out_d + b0 * S * ld_out, nb * S, len, ld_out, minmax_d + 2 * b0 * S,
                                         pcm_d + b0 * S * len, stream)) return rc;
        cudaEvent_t ev = cs.ev[k % 8];
        CK(cudaEventRecord(ev, st));
        CK(cudaStreamWaitEvent(cs.d2h, ev, 0));
        CK(cudaMemcpyAsync(pcm_h + b0 * S * len, pcm_d + b0 * S * len, sizeof(int16_t) * nb * S * len, cudaMemcpyDeviceToHost, cs.d2h));
    }
    CopyStreams::Tag* ts = nullptr; CopyStreams::Tag* td = nullptr;
    if (int rc = cs.claim(cs.src, pcm_d, &ts)) return rc;
    if (int rc = cs.claim(cs.dst, pcm_h, &td)) return rc;
    CK(cudaEventRecord(ts->ev, cs.d2h));
    CK(cudaEventRecord(td->ev, cs.d2h));
    return GSS_OK;
}

int gss_wait_host(const void* host_ptr) {
    // any thread may wait: the tags are looked up under the owning device's mutex, the wait itself happens outside it
    if (host_ptr) {
        for (CopyStreams& c : g_cs_dev) {
            cudaEvent_t ev = nullptr;
            {
                std::lock_guard<std::mutex> lk(c.mu);
                if (!c.ok) continue;
                if (CopyStreams::Tag* t = CopyStreams::find(c.dst, host_ptr)) ev = t->ev;
            }
            if (ev) { CK(cudaEventSynchronize(ev)); return GSS_OK; }
        }
    }
    // NULL, or a pointer without a tag (never enqueued, or its tag was recycled after a drain): every pending
    // download of the current device - never a silent return before the data has landed
    CopyStreams* cp = nullptr;
    if (int rc = copy_streams(&cp)) return rc;
    cudaStream_t d2h = nullptr;
    { std::lock_guard<std::mutex> lk(cp->mu); if (cp->ok) d2h = cp->d2h; }
    if (d2h) CK(cudaStreamSynchronize(d2h));
    return GSS_OK;
}

int gss_mask_istft_d2h(const float* wave_d, const float* mask_d, int64_t B, int S, int64_t n, int64_t ld, int N, int H,
                       float* out_d, float* out_h, int64_t ld_out, int chunks, void* stream) {
    if (int rc = gss_mask_istft_d2h_async(wave_d, mask_d, B, S, n, ld, N, H, out_d, out_h, ld_out, chunks, stream)) return rc;
    return gss_wait_host(out_h);
}

}  // extern "C"
#endif  // GSS_HAS(0)
