// FFT-pass-shaped loop: [math on 16 v2] -> [16 STS.64] -> syncwarp -> [16 LDS.64 transposed] -> ...
// SETS = independent register sets per warp processed in a software-pipelined fashion
// (SETS=1: plain; SETS=2: math of set B is issued between the stores and the loads of set A).
// Reports cycles per pass per SM sub-partition-warp and the pure-math / pure-memory references.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 512
template <int LV>
__device__ __forceinline__ void math(float2 (&a)[16], const float2 c) {
    // 3 dependent levels of 16 independent packed ops (like a radix-8 butterfly on 2x8 complex)
#pragma unroll
    for (int lvl = 0; lvl < LV; ++lvl) {
        float2 b[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = (i & 1) ? __fadd2_rn(a[i ^ (1 << ((lvl % 3) + 1))], a[i]) : __ffma2_rn(a[i ^ (1 << ((lvl % 3) + 1))], c, a[i]);
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = b[i];
    }
}
template <int SETS, int MODE, int LV>   // MODE 0: both, 1: math only, 2: mem only
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc) {
    extern __shared__ float2 sm2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* buf = sm2 + warp * (SETS * 16 * 33);
    float2 a[SETS][16];
#pragma unroll
    for (int s = 0; s < SETS; ++s)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[s][i] = make_float2(out[i + lane], out[i + 64 + s]);
    const float2 c = make_float2(out[300], out[301]);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int s = 0; s < SETS; ++s) {
            if (MODE != 2) math<LV>(a[s], c);
            if (MODE != 1) {
                float2* b = buf + s * 16 * 33;
#pragma unroll
                for (int i = 0; i < 16; ++i) b[i * 33 + lane] = a[s][i];
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 16; ++i) a[s][i] = b[i * 33 + ((lane + 8 * i + it) & 31)];
                __syncwarp();
            }
        }
    }
    long long t1 = clock64();
    float s_ = 0;
#pragma unroll
    for (int s = 0; s < SETS; ++s)
#pragma unroll
        for (int i = 0; i < 16; ++i) s_ += a[s][i].x + a[s][i].y;
    out[4096 + blockIdx.x * blockDim.x + threadIdx.x] = s_;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int SETS, int LV>
void run(int nb, float* out, long long* cyc) {
    static long long h[1024];
    for (int W : {8, 12, 16}) {
        double r[3];
        size_t smem = (size_t)W * SETS * 16 * 33 * sizeof(float2);
        for (int m = 0; m < 3; ++m) {
            auto launch = [&](int mode) {
                if (mode == 0) { cudaFuncSetAttribute(k<SETS, 0, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<SETS, 0, LV><<<nb, W * 32, smem>>>(out, cyc); }
                if (mode == 1) { cudaFuncSetAttribute(k<SETS, 1, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<SETS, 1, LV><<<nb, W * 32, smem>>>(out, cyc); }
                if (mode == 2) { cudaFuncSetAttribute(k<SETS, 2, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<SETS, 2, LV><<<nb, W * 32, smem>>>(out, cyc); }
            };
            launch(m); cudaDeviceSynchronize(); launch(m); cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < nb; ++i) s += h[i];
            r[m] = s / nb / ITERS / SETS;      // cycles per pass (per warp-set), whole SM running W warps
        }
        printf("LV=%2d SETS=%d W=%2d (%d/SMSP): per pass  both %6.1f  math %6.1f  mem %6.1f  | SM-level per pass-of-all-warps: max %6.1f sum %6.1f\n",
               LV, SETS, W, W / 4, r[0], r[1], r[2], r[1] > r[2] ? r[1] : r[2], r[1] + r[2]);
    }
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int nb = prop.multiProcessorCount;
    float* out; long long* cyc; cudaMalloc(&out, sizeof(float) * (4096 + nb * 1024)); cudaMemset(out, 0, sizeof(float) * (4096 + nb * 1024)); cudaMalloc(&cyc, sizeof(long long) * nb);
    run<1, 6>(nb, out, cyc);
    run<1, 9>(nb, out, cyc);
    run<1, 12>(nb, out, cyc);
    run<2, 9>(nb, out, cyc);
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
