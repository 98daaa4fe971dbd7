// Issue cost of shared-memory instructions by access width, inside an FFT-pass-shaped loop (8 warps per SM):
// the same 32 floats per thread go through shared memory as 32 x 32-bit, 16 x 64-bit or 8 x 128-bit stores + loads
// (conflict-free layouts), between blocks of 144 FP32x2 operations.  cycles = 2 * FP32x2 + cost * LSU instructions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench8 tools/ubench8.cu
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
template <int WIDTH, int DOMATH>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* cyc) {
    extern __shared__ float4 sm4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* buf = reinterpret_cast<float*>(sm4) + warp * 32 * 36;
    float2 a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = make_float2(out[i + lane], out[i + 64]);
    const float2 c = make_float2(out[300], out[301]);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (DOMATH) {
#pragma unroll
            for (int l = 0; l < LV; ++l) level(a, c, l);
        }
        const int rl = (lane + it) & 31;
        if (WIDTH == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { buf[(2 * i) * 32 + lane] = a[i].x; buf[(2 * i + 1) * 32 + lane] = a[i].y; }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 16; ++i) { a[i].x = buf[(2 * i) * 32 + rl]; a[i].y = buf[(2 * i + 1) * 32 + rl]; }
        } else if (WIDTH == 8) {
            float2* b2 = reinterpret_cast<float2*>(buf);
#pragma unroll
            for (int i = 0; i < 16; ++i) b2[i * 32 + lane] = a[i];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = b2[i * 32 + rl];
        } else {
            float4* b4 = reinterpret_cast<float4*>(buf);
#pragma unroll
            for (int i = 0; i < 8; ++i) b4[i * 32 + lane] = make_float4(a[2 * i].x, a[2 * i].y, a[2 * i + 1].x, a[2 * i + 1].y);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) { float4 v = b4[i * 32 + rl]; a[2 * i] = make_float2(v.x, v.y); a[2 * i + 1] = make_float2(v.z, v.w); }
        }
        __syncwarp();
    }
    long long t1 = clock64();
    float s_ = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s_ += a[i].x + a[i].y;
    out[4096 + blockIdx.x * blockDim.x + threadIdx.x] = s_;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int WIDTH, int DOMATH>
double run(int nb, float* out, long long* cyc) {
    static long long h[1024];
    size_t smem = 8 * 32 * 36 * sizeof(float);
    cudaFuncSetAttribute(k<WIDTH, DOMATH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<WIDTH, DOMATH><<<nb, 256, smem>>>(out, cyc); cudaDeviceSynchronize();
    k<WIDTH, DOMATH><<<nb, 256, smem>>>(out, cyc); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < nb; ++i) s += h[i];
    return s / nb / ITERS;
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int nb = prop.multiProcessorCount;
    float* out; long long* cyc; cudaMalloc(&out, sizeof(float) * (4096 + nb * 1024)); cudaMemset(out, 0, sizeof(float) * (4096 + nb * 1024)); cudaMalloc(&cyc, sizeof(long long) * nb);
    const double fp = 16 * LV * 2 * 2;    // two warps per sub-partition
    double r;
    r = run<4, 1>(nb, out, cyc);  printf("32-bit  x64 instr: %6.1f cycles/pass -> %.2f issue cycles per LDS/STS  (mem only %6.1f)\n", r, (r - fp) / (64 * 2), run<4, 0>(nb, out, cyc));
    r = run<8, 1>(nb, out, cyc);  printf("64-bit  x32 instr: %6.1f cycles/pass -> %.2f issue cycles per LDS/STS  (mem only %6.1f)\n", r, (r - fp) / (32 * 2), run<8, 0>(nb, out, cyc));
    r = run<16, 1>(nb, out, cyc); printf("128-bit x16 instr: %6.1f cycles/pass -> %.2f issue cycles per LDS/STS  (mem only %6.1f)\n", r, (r - fp) / (16 * 2), run<16, 0>(nb, out, cyc));
    printf("FP32x2 only floor %.0f; status %s\n", fp, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
