// In-warp pairing of FP32x2 with shared-memory instructions.  Model from ubench4: per SM sub-partition an FP32x2
// costs 2 issue cycles and an LDS.64/STS.64 3, ADDED across warps (no cross-warp overlap).  ptxas marks an FP32x2 that
// is followed by a non-FMA instruction of the SAME warp with stall 1: can a warp that carries TWO independent
// transform-like streams (A, B) hide the exchange of one under the butterflies of the other?
//   MODE 0: single stream   [math A | STS A | sync | LDS A | sync]
//   MODE 1: dual, lock-step [STS A, STS B | sync | LDS A, LDS B | math A, math B | sync]
//   MODE 2: dual, skewed    [STS A | sync | LDS A + math B (same block) | STS B | sync | LDS B + math A]
//   MODE 3: dual, skewed, hand-interleaved in the source (one LDS after every few packed ops)
// Reports SM cycles per stream-pass at 4 and 8 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench7 tools/ubench7.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 512
#define LV 9
__device__ __forceinline__ void level(float2 (&a)[16], const float2 c, int lvl) {
    float2 b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) b[i] = (i & 1) ? __fadd2_rn(a[i ^ (1 << ((lvl % 3) + 1))], a[i]) : __ffma2_rn(a[i ^ (1 << ((lvl % 3) + 1))], c, a[i]);
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = b[i];
}
__device__ __forceinline__ void math(float2 (&a)[16], const float2 c) {
#pragma unroll
    for (int l = 0; l < LV; ++l) level(a, c, l);
}
__device__ __forceinline__ void sts(float2* b, int lane, const float2 (&a)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) b[i * 33 + lane] = a[i];
}
__device__ __forceinline__ void lds(const float2* b, int lane, int it, float2 (&a)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = b[i * 33 + ((lane + 8 * i + it) & 31)];
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* cyc) {
    extern __shared__ float2 sm2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* bufA = sm2 + warp * (2 * 16 * 33);
    float2* bufB = bufA + 16 * 33;
    float2 a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = make_float2(out[i + lane], out[i + 64]); b[i] = make_float2(out[i + 32 + lane], out[i + 96]); }
    const float2 c = make_float2(out[300], out[301]);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
            math(a, c); sts(bufA, lane, a); __syncwarp(); lds(bufA, lane, it, a); __syncwarp();
        } else if (MODE == 1) {
            sts(bufA, lane, a); sts(bufB, lane, b); __syncwarp();
            lds(bufA, lane, it, a); lds(bufB, lane, it, b);
            math(a, c); math(b, c); __syncwarp();
        } else if (MODE == 2) {
            sts(bufA, lane, a); __syncwarp();
            lds(bufA, lane, it, a); math(b, c);
            sts(bufB, lane, b); __syncwarp();
            lds(bufB, lane, it, b); math(a, c);
        } else {
            // skewed, hand-interleaved: stores of A ride the last level of A, loads of A ride the levels of B, ...
            sts(bufA, lane, a); __syncwarp();
#pragma unroll
            for (int l = 0; l < LV; ++l) {
                if (l < 8) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) a[2 * l + i] = bufA[(2 * l + i) * 33 + ((lane + 8 * (2 * l + i) + it) & 31)];
                }
                level(b, c, l);
            }
            sts(bufB, lane, b); __syncwarp();
#pragma unroll
            for (int l = 0; l < LV; ++l) {
                if (l < 8) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) b[2 * l + i] = bufB[(2 * l + i) * 33 + ((lane + 8 * (2 * l + i) + it) & 31)];
                }
                level(a, c, l);
            }
        }
    }
    long long t1 = clock64();
    float s_ = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s_ += a[i].x + a[i].y + b[i].x + b[i].y;
    out[4096 + blockIdx.x * blockDim.x + threadIdx.x] = s_;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(int nb, float* out, long long* cyc, const char* name) {
    static long long h[1024];
    for (int W : {4, 8}) {
        size_t smem = (size_t)W * 2 * 16 * 33 * sizeof(float2);
        cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<MODE><<<nb, W * 32, smem>>>(out, cyc); cudaDeviceSynchronize();
        k<MODE><<<nb, W * 32, smem>>>(out, cyc); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < nb; ++i) s += h[i];
        const int streams = MODE == 0 ? 1 : 2;
        printf("%-28s W=%d: %7.1f SM cycles per stream-pass of all warps  (FP32x2-only floor %d, +3/LSU model %d)\n", name, W,
               s / nb / ITERS / streams, 16 * LV * 2 * (W / 4), (16 * LV * 2 + 32 * 3) * (W / 4));
    }
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int nb = prop.multiProcessorCount;
    float* out; long long* cyc; cudaMalloc(&out, sizeof(float) * (4096 + nb * 1024)); cudaMemset(out, 0, sizeof(float) * (4096 + nb * 1024)); cudaMalloc(&cyc, sizeof(long long) * nb);
    run<0>(nb, out, cyc, "single");
    run<1>(nb, out, cyc, "dual lock-step");
    run<2>(nb, out, cyc, "dual skewed");
    run<3>(nb, out, cyc, "dual skewed hand-interleaved");
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
