/* Pure-C consumer of the drop-in boundary: includes include/gss_api.h, links libgss.so, and exercises the
 * entry points that need no GPU (version, frame arithmetic of scipy.signal.stft K1, argument errors).
 * Built and run by tests/test_native_abi.py with gcc - proves the ABI is plain C (no torch, no C++ types). */
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include "gss_api.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #c, gss_last_error()); return 1; } } while (0)

int main(void) {
    int64_t T = 0, nadd = 0;
    int sizes[16];
    CHECK(gss_version() >= 100);
    CHECK(gss_frame_count(48000, 512, 128, &T, &nadd) == GSS_OK && T == 376 && nadd == 0);
    CHECK(gss_frame_count(46797, 512, 128, &T, &nadd) == GSS_OK && T == 367 && nadd == 51);
    CHECK(gss_frame_count(0, 512, 128, &T, &nadd) == GSS_EINVAL);
    CHECK(gss_supported_fft_sizes(sizes, 16) >= 5);
    /* rejected before any CUDA call */
    CHECK(gss_stft_packed(NULL, 1, 4000, 4000, 500, 125, 0, 1e-7f, NULL, NULL) == GSS_EUNSUPPORTED);
    CHECK(strstr(gss_last_error(), "power of two") != NULL);
    CHECK(gss_stft_packed(NULL, 1, 4000, 4000, 512, 128, 0, 1e-7f, NULL, NULL) == GSS_EINVAL);
    CHECK(gss_mask_istft(NULL, NULL, 1, 3, 4000, 4000, 512, 128, NULL, 4096, NULL) == GSS_EINVAL);
    CHECK(gss_mix_features(NULL, NULL, 1, 3, 10, 256, 0, 1e-7f, NULL, NULL, NULL) == GSS_EINVAL);
    CHECK(gss_mask_istft_feature(NULL, NULL, 1, 3, 10, 512, 128, 0, NULL, 4096, NULL) == GSS_EINVAL);
    CHECK(gss_metric_finalise(NULL, NULL, 1, 1, 1, 1.0, NULL, NULL) == GSS_EINVAL);
    printf("abi_smoke ok (libgss %d)\n", gss_version());
    return 0;
}
