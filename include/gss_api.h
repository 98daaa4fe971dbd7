/*
 * gss_api.h - C ABI of libgss, the B200 (sm_100a) spectral hot path that replaces
 * the host-side SciPy/NumPy transforms of ahmedassal/GAN_SASS_TF.
 *
 * Plain C: pointers, sizes, an opaque stream handle.  No torch types.  Every
 * entry point cites the reference interface it replaces (paths relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - N  = FFT_SIZE (app/hparams.py:12), power of two in [64, 4096] (gss_supported_fft_sizes);
 *     H  = hop, one of N/2 (the reference's SciPy default), N/4, N/8.
 *   - "packed feature" = the reference's [T, N] float32 layout of app/utils.py:8-26:
 *     feat[t,k]=Re X[k] (0<=k<N/2), feat[t,N/2]=Re X[N/2], feat[t,N/2+k]=Im X[k].
 *   - Device entry points only enqueue work on `stream` (a cudaStream_t cast to
 *     void*; NULL = legacy default stream); they never synchronise or allocate.
 *     Host entry points (suffix _host) copy in/out themselves and return after the
 *     result is in the host buffer.
 *   - All functions return 0 on success or a negative gss_status; the message is
 *     available from gss_last_error() (thread-local).
 */
#ifndef GSS_API_H
#define GSS_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GSS_OK = 0,
    GSS_EINVAL = -1,        /* bad shape / pointer / alignment                      */
    GSS_EUNSUPPORTED = -2,  /* (N, H, S) outside the supported set, or N > n        */
    GSS_ECUDA = -3,         /* a CUDA runtime call failed                           */
    GSS_ENOMEM = -4         /* workspace allocation failed (host entry points only) */
} gss_status;

enum {
    GSS_FLAG_LOG = 1,       /* fuse to_log_signal (app/ops.py:228-238) into the STFT epilogue  */
    GSS_FLAG_EXP = 2,       /* fuse to_exp_signal (app/ops.py:241-251) into the iSTFT prologue */
    GSS_FLAG_REVERSE = 4    /* gss_mask_istft_feature: walk the rows from the last to the first - a hint for callers
                               whose features were written moments earlier (the tail is still in L2); same results */
};

int         gss_version(void);
const char* gss_last_error(void);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t     gss_launch_count(void);

#ifdef GSS_EXPERIMENTAL
/* Only in lib/libgss_experimental.so (built with -DGSS_EXPERIMENTAL; what the cross-check tests and the tuning tools
 * load).  The product library exports neither: it keeps no process-wide mutable switches. */

/* Kernel selection: 0 = automatic (register-exchange streaming kernels for N = 256 / 512, shared-memory team
 * kernels for 1024 / 2048 / 4096, per-frame kernels for 64 / 128), 1 = no register-exchange kernels (team kernels
 * take 256 / 512 too), 2 = per-frame kernels only.  Process-wide; meant for cross-checking the implementations
 * against each other. */
int         gss_set_path(int path);

/* Fused-synthesis kernel for FFT_SIZE 512 (gss_mask_istft): 0 = register-resident streaming kernel (default),
 * 1 = role-split CTAs (analysis warp + one synthesis warp per source), 2 = per-thread state parked in tensor
 * memory (S a multiple of 3; other S take the default).  All three give the same results; 1 and 2 are kept as
 * measured design alternatives and cross-checks.  Process-wide; the environment variable GSS_SYNTH_SPLIT
 * sets the initial value. */
int         gss_set_synth_variant(int variant);
#endif /* GSS_EXPERIMENTAL */

/* FFT sizes this build has kernels for (writes up to `cap` entries, returns the count);
 * hops N/2 (the reference's SciPy default), N/4 and N/8 are supported for each. */
int         gss_supported_fft_sizes(int* sizes, int cap);

/* Frame arithmetic of scipy.signal.stft(boundary='zeros', padded=True) as the
 * reference calls it (main.py:97): nadd = (-n mod H) mod N, T = (n+nadd)/H + 1.
 * Pure host function. */
int gss_frame_count(int64_t n, int N, int H, int64_t* T, int64_t* nadd);

/* A1+A2(+A3): replaces scipy.signal.stft(x, nperseg=N)[2] + utils.spectrum_to_feature
 * (main.py:97-98, app/datasets/TIMIT/process.py:97-98) and optionally
 * ops.to_log_signal (main.py:338).
 *   wave [B, ld] f32 (first n samples of each row used) -> feat [B, T, N] f32. */
int gss_stft_packed(const float* wave, int64_t B, int64_t n, int64_t ld, int N, int H,
                    int flags, float eps, float* feat, void* stream);

/* same, int16 PCM input as process.py:94 reads it (scipy gives complex64 there) */
int gss_stft_packed_i16(const int16_t* wave, int64_t B, int64_t n, int64_t ld, int N, int H,
                        int flags, float eps, float* feat, void* stream);

/* (A5+)A6+A8: replaces utils.feature_to_spectrum + scipy.signal.istft(nperseg=N)
 * (main.py:110-111), optionally preceded by ops.to_exp_signal (main.py:342).
 *   feat [R, T, N] f32 -> wave_out [R, ld_out] f32, (T-1)*H samples per row. */
int gss_istft_packed(const float* feat, int64_t R, int64_t T, int N, int H,
                     int flags, float eps, float* wave_out, int64_t ld_out, void* stream);

/* A1+A7+A8 fused synthesis stage (no reference code for A7):
 * recompute the mixture spectrum from the waveform, multiply by the separator's
 * per-source real gains and inverse-transform with register overlap-add.
 *   wave [B, ld] f32, mask [B, S, T, N/2] f32 -> out [B*S, ld_out] f32, row b*S+s
 *   (the output order of app/modules.py:396-399), (T-1)*H samples per row. */
int gss_mask_istft(const float* wave, const float* mask, int64_t B, int S, int64_t n, int64_t ld,
                   int N, int H, float* out, int64_t ld_out, void* stream);

/* A1+A2+A3 with both outputs: the linear packed spectrum (what main.py:328-337 calls s_mixed_signals and what a
 * mask is applied to) and its to_log_signal (what the separator reads, main.py:338) from ONE transform.
 *   wave [B, ld] f32 -> feat_lin [B, T, N] f32, feat_log [B, T, N] f32. */
int gss_stft_packed_dual(const float* wave, int64_t B, int64_t n, int64_t ld, int N, int H, float eps,
                         float* feat_lin, float* feat_log, void* stream);

/* A7+A8 fused, fed by the mixture's LINEAR packed spectrum (the reference's own data flow: its graph holds the
 * mixture as packed features, main.py:328-337, not as a waveform) instead of recomputing it from the waveform:
 *   feat [B, T, N] f32, mask [B, S, T, N/2] f32 -> out [B*S, ld_out] f32, row b*S+s, (T-1)*H samples per row.
 * Same results as gss_apply_mask + gss_istft_packed and as gss_mask_istft on the waveform the features came from;
 * trades 4TN - 4n more bytes read per mixture for one transform fewer per frame pair.  FFT_SIZE 256 / 512 have the
 * register-exchange kernel, 1024 / 2048 / 4096 the team kernel; other sizes return GSS_EUNSUPPORTED.
 * feat and mask 16-byte aligned (TMA bulk copies).  flags: 0 or GSS_FLAG_REVERSE. */
int gss_mask_istft_feature(const float* feat, const float* mask, int64_t B, int S, int64_t T, int N, int H, int flags,
                           float* out, int64_t ld_out, void* stream);

/* Same, and the auto-encoder loss partial of main.py:353-361 from the registers that already hold the spectrum and the
 * gains: ae_rows[b] = sum over the packed elements of mixture b of ((sum_s mask_s - 1) * feature)^2, i.e.
 * sum((sum_s separated_s - mixed)^2) for separated_s = mask_s * mixed; the loss is sum_b ae_rows[b] / (B*T*N).
 * ae_rows [B] f32 is zeroed by the call (a memset node on `stream`).  FFT_SIZE 256 / 512 and S <= 4 (S <= 3 at hop
 * N/8); other shapes return GSS_EUNSUPPORTED (use gss_apply_mask + gss_ae_partial).  ae_rows = NULL: plain call. */
int gss_mask_istft_feature_ae(const float* feat, const float* mask, int64_t B, int S, int64_t T, int N, int H, int flags,
                              float* out, int64_t ld_out, float* ae_rows, void* stream);

/* The per-batch metric vector that is all-reduced over the GPUs (app/parallel.py; batch means of main.py:353-361 and
 * :446-457): vec4 = [sum_b mean_i max_k snr[b,i,k], sum_b ae_rows[b] / elems_per_row, 0, B].  snr [B,m,n] (from
 * gss_cross_snr) and ae_rows [B] (from gss_mask_istft_feature_ae or gss_ae_partial) may each be NULL.  One tiny launch,
 * so the all-reduce can follow on the same stream without a host round trip. */
int gss_metric_finalise(const float* ae_rows, const float* snr, int64_t B, int m, int n, double elems_per_row,
                        float* vec4, void* stream);

/* A14: scipy.signal.resample(x, num) (Fourier method) as main.py:89-95 calls it for files that are not at 16 kHz:
 * rfft, keep min(n, num)/2 + 1 bins (shared Nyquist bin folded / split as SciPy does for real input), irfft to num
 * samples, scale num/n.  Arbitrary n and num: both transforms are Bluestein chirp-z transforms over a power-of-two
 * Stockham FFT, all in float64 (SciPy promotes int16 samples to float64), phases reduced exactly in integers.
 *   x [rows, ld] f64 -> y [rows, ld_y] f64, num samples per row.  workspace: gss_resample_workspace_bytes(n, num)
 *   bytes of device memory, 16-byte aligned, owned by the caller. */
size_t gss_resample_workspace_bytes(int64_t n, int64_t num);
int gss_resample_f64(const double* x, int64_t rows, int64_t n, int64_t ld, int64_t num, double* y, int64_t ld_y,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Batch assembly for the device-resident corpus that replaces the FFT_SIZE-baked pickles (TIMIT/process.py:80-161,
 * app/datasets/timit.py:47-52): out[r, :] = utterance idx[r] of the flat int16 store, zero-padded to ld samples.
 * offsets / lengths [n_utterances] and idx [B] are int64 device arrays.  One launch per batch. */
int gss_gather_rows_i16(const int16_t* flat, const int64_t* offsets, const int64_t* lengths, const int64_t* idx,
                        int64_t B, int64_t ld, int16_t* out, void* stream);

/* A7 alone on packed features: mix [B,T,N], mask [B,S,T,N/2] -> out [B*S,T,N] */
int gss_apply_mask(const float* mix, const float* mask, int64_t B, int S, int64_t T, int N,
                   float* out, void* stream);

/* Building blocks of the transforms' adjoints (the reference never differentiates through its SciPy
 * calls, main.py:97/111, so these have no reference counterpart).  With norm[p] = sum_t hann^2 over the frames that
 * cover sample p - the overlap-add weight scipy.signal.istft divides by, including its > 1e-10 guard -
 *   out[r, p] = scale * in[r, p] / norm[p]   (inverse != 0)      or      scale * in[r, p] * norm[p],   p < len <= (T-1)*H;
 * gss_scale_packed multiplies packed features by c_all and their DC / Nyquist slots (0 and N/2) by c_edge on top.
 *   d iSTFT^T (g) = scale_packed(STFT(g / norm), N/2, 1/2)       d STFT^T (g) = (4/N) * norm * iSTFT(scale_packed(g, 1/2, 2)) */
int gss_ola_norm_scale(const float* in, float* out, int64_t rows, int64_t len, int64_t ld_in, int64_t ld_out,
                       int64_t T, int N, int H, int inverse, float scale, void* stream);
int gss_scale_packed(const float* in, float* out, int64_t rows, int N, float c_all, float c_edge, void* stream);

/* A3 / A5: ops.to_log_signal / ops.to_exp_signal (app/ops.py:228-251) on `rows` rows of N */
int gss_to_log(const float* in, float* out, int64_t rows, int N, float eps, void* stream);
int gss_to_exp(const float* in, float* out, int64_t rows, int N, float eps, void* stream);

/* A11: ops.batch_cross_snr (app/ops.py:191-225).  clear [B,m,L], noisy [B,n,L] ->
 * snr [B,m,n]; batch_snr (ops.py:162-189) is the m=n=1 diagonal. */
int gss_cross_snr(const float* clear, const float* noisy, int64_t B, int m, int n, int64_t L,
                  float eps, float* snr, void* stream);

/* A10 (+A3): main.py:328-338.  src [B*n_sig, T, N] packed features -> mix [B, T, N] =
 * sum_i src[b*n_sig+i] (+ noise [B,T,N], NULL = none; the reference draws it from
 * tf.random_normal(stddev=0.1), the caller passes the draw so that runs are reproducible).
 * flags = GSS_FLAG_LOG additionally writes to_log_signal(mix) to mix_log in the same pass
 * (then mix may be NULL). */
int gss_mix_features(const float* src, const float* noise, int64_t B, int n_sig, int64_t T, int N,
                     int flags, float eps, float* mix, float* mix_log, void* stream);

/* Backward passes: the reference back-propagates through to_log_signal /
 * to_exp_signal (app/ops.py:228-251 under the optimisers of main.py:481-484); apply_mask is the
 * new op A7.  in / gout / gin [rows, N]; gmix [B,T,N] and gmask [B,S,T,N/2] may each be NULL. */
int gss_to_log_bwd(const float* in, const float* gout, float* gin, int64_t rows, int N, float eps, void* stream);
int gss_to_exp_bwd(const float* in, const float* gout, float* gin, int64_t rows, int N, float eps, void* stream);
int gss_apply_mask_bwd(const float* mix, const float* mask, const float* gout, int64_t B, int S,
                       int64_t T, int N, float* gmix, float* gmask, void* stream);

/* A12: main.py:353-361.  sep [B,S,L], mix [B,L] -> partial[B] = sum_l (sum_s sep - mix)^2
 * (the caller divides by B*L; per-utterance partials are what the ranks all-reduce). */
int gss_ae_partial(const float* sep, const float* mix, int64_t B, int S, int64_t L,
                   float* partial, void* stream);

/* A9: main.py:112-116 WAV normalisation.  x [R, ld] f32 (len samples used) -> pcm [R, len]
 * int16, each row shifted to its min and scaled to 32767 (truncation).
 * minmax is a caller-provided scratch of 2*R floats. */
int gss_wav16_normalise(const float* x, int64_t R, int64_t len, int64_t ld, float* minmax,
                        int16_t* pcm, void* stream);

/* ---- host-buffer entry points (what load_wavfile / save_wavfile bind to) ---- */

/* main.py:67-99 minus file I/O: host wave [B,n] -> host feat [B,T,N] */
int gss_stft_packed_host(const float* wave_h, int64_t B, int64_t n, int N, int H,
                         int flags, float eps, float* feat_h);
/* main.py:102-111 minus file I/O: host feat [R,T,N] -> host wave [R,(T-1)H] */
int gss_istft_packed_host(const float* feat_h, int64_t R, int64_t T, int N, int H,
                          int flags, float eps, float* wave_h);

/* Chunked, double-buffered end-to-end helpers used by bench.py's e2e leg.
 * gss_stft_h2d: pinned host wave -> device wave workspace (kept for the synthesis
 *               stage) -> device feat; copies and kernels overlap chunk by chunk.
 * gss_mask_istft_d2h: device wave + device mask -> device out workspace -> pinned host out.
 * `chunks` >= 1 splits the batch; both return after the last copy has landed. */
int gss_stft_h2d(const float* wave_h, float* wave_d, int64_t B, int64_t n, int64_t ld, int N, int H,
                 int flags, float eps, float* feat_d, int chunks, void* stream);
int gss_mask_istft_d2h(const float* wave_d, const float* mask_d, int64_t B, int S, int64_t n,
                       int64_t ld, int N, int H, float* out_d, float* out_h, int64_t ld_out,
                       int chunks, void* stream);

/* Non-blocking forms for a batch loop (main.py:749-771 run over many clips): they only enqueue -
 * uploads on the library's H2D stream, kernels on `stream`, downloads on the library's D2H
 * stream - so the upload of batch k+1 and the download of batch k share the link in both
 * directions.  Ordering the library guarantees: the upload starts after everything already
 * enqueued on `stream` (wave_d may be reused); the synthesis kernels start after a pending
 * download from the same out_d.  The caller double-buffers wave_d/feat_d/out_d/out_h and calls
 * gss_wait_host(out_h) before reading out_h (NULL: wait for every pending download). */
int gss_stft_h2d_async(const float* wave_h, float* wave_d, int64_t B, int64_t n, int64_t ld, int N,
                       int H, int flags, float eps, float* feat_d, int chunks, void* stream);
int gss_mask_istft_d2h_async(const float* wave_d, const float* mask_d, int64_t B, int S, int64_t n,
                             int64_t ld, int N, int H, float* out_d, float* out_h,
                             int64_t ld_out, int chunks, void* stream);
int gss_wait_host(const void* host_ptr);

/* The same two stages with int16 PCM on the host link (the sample format of the WAV files main.py:83 reads
 * and main.py:116 writes): half the bytes each way.  Upload: pcm_h [B,n] -> pcm_d [B,ld] -> features from the
 * raw sample values (gss_stft_packed_i16, no rescale, like process.py:97) + a float32 copy wave_d [B,ld] that
 * the synthesis stage reads.  Download: separated waveforms out_d -> per-row shift/scale to [0, 32767]
 * (main.py:112-116; minmax_d: 2*B*S floats of scratch) -> pcm_d [B*S,(T-1)H] -> pcm_h.  Same ordering rules
 * as above; gss_wait_host(pcm_h) before reading. */
int gss_stft_h2d_i16_async(const int16_t* pcm_h, int16_t* pcm_d, float* wave_d, int64_t B, int64_t n,
                           int64_t ld, int N, int H, int flags, float eps, float* feat_d, int chunks,
                           void* stream);
int gss_mask_istft_d2h_pcm16_async(const float* wave_d, const float* mask_d, int64_t B, int S, int64_t n,
                                   int64_t ld, int N, int H, float* out_d, float* minmax_d,
                                   int16_t* pcm_d, int16_t* pcm_h, int64_t ld_out, int chunks,
                                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSS_API_H */
