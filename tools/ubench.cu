// Pipe micro-benchmarks used to size the FFT kernels (DESIGN.md "calibration").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
// Each kernel runs one 1024-thread CTA per SM and reports lane-ops (or bytes)
// per SM clock, measured with clock64() inside the kernel.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

template <int MODE>
__global__ void __launch_bounds__(1024, 1) fp_kernel(float* out, long long* cyc, float seed) {
    float a[8], b[8];
    float2 p[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed + threadIdx.x * 1e-3f + i; b[i] = seed * 0.5f + i;
        p[i] = make_float2(a[i], b[i]); q[i] = make_float2(b[i], a[i]);
    }
    const float c0 = seed * 1.0001f, c1 = seed * 0.9999f;
    const float2 cc = make_float2(c0, c1);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], c0, b[i]);              // FFMA 3 regs
            if (MODE == 1) a[i] = a[i] + b[i];                       // FADD
            if (MODE == 2) p[i] = __ffma2_rn(p[i], cc, q[i]);        // FFMA2
            if (MODE == 3) p[i] = __fadd2_rn(p[i], q[i]);            // FADD2
            if (MODE == 4) a[i] = a[i] * c0;                         // FMUL
            if (MODE == 5) { a[i] = a[i] + b[i]; p[i] = __fadd2_rn(p[i], q[i]); }  // FADD + FADD2 mixed
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared memory: MODE 0 LDS.32, 1 LDS.64, 2 LDS.128, 3 STS.32, 4 STS.64, 5 STS.128, 6 SHFL, 7 LDS.128 + FADD2 mix
template <int MODE>
__global__ void __launch_bounds__(1024, 1) smem_kernel(float* out, long long* cyc, int stride) {
    extern __shared__ float4 sm4[];
    float* sm = reinterpret_cast<float*>(sm4);
    for (int i = threadIdx.x; i < 40960; i += blockDim.x) sm[i] = i;
    __syncthreads();
    float acc = threadIdx.x * 0.37f; float2 acc2 = make_float2(0, 0); float4 acc4 = make_float4(0, 0, 0, 0);
    float2 p[4];
    for (int i = 0; i < 4; ++i) p[i] = make_float2(i, threadIdx.x);
    int base = threadIdx.x * stride;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int off = base + i * 1024;
            if (MODE == 0) acc += sm[(off + it)];
            if (MODE == 1) { float2 v = reinterpret_cast<float2*>(sm)[(off + it)]; acc2.x += v.x; acc2.y += v.y; }
            if (MODE == 2) { float4 v = sm4[(off + it)]; acc4.x += v.x; acc4.y += v.y; acc4.z += v.z; acc4.w += v.w; }
            if (MODE == 3) sm[(off + it)] = acc + i;
            if (MODE == 4) reinterpret_cast<float2*>(sm)[(off + it)] = make_float2(acc + i, acc);
            if (MODE == 5) sm4[(off + it)] = make_float4(acc + i, acc, acc, acc);
            if (MODE == 6) acc += __shfl_xor_sync(0xffffffffu, acc + i, 1 + (i & 15));
            if (MODE == 7) {
                float4 v = sm4[(off + it)]; acc4.x += v.x;
#pragma unroll
                for (int r = 0; r < 4; ++r) p[r] = __fadd2_rn(p[r], make_float2(v.y, v.z));
#pragma unroll
                for (int r = 0; r < 4; ++r) p[r] = __fadd2_rn(p[r], make_float2(v.w, v.x));
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + acc2.x + acc2.y + acc4.x + acc4.y + acc4.z + acc4.w + p[0].x + p[1].y + p[2].x + p[3].y;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double avg_cyc(long long* d_cyc, int nb) {
    static long long h[1024];
    cudaMemcpy(h, d_cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < nb; ++i) s += h[i];
    return s / nb;
}

int main() {
    int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    int nb = prop.multiProcessorCount;
    printf("device %s SMs %d clock %d kHz\n", prop.name, nb, prop.clockRate);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * nb * 1024); cudaMalloc(&cyc, sizeof(long long) * nb);
    const char* fpn[] = {"FFMA", "FADD", "FFMA2", "FADD2", "FMUL", "FADD+FADD2"};
    const double lane_ops[] = {1, 1, 2, 2, 1, 3};
#define RUN_FP(M) { fp_kernel<M><<<nb, 1024>>>(out, cyc, 1.0f); cudaDeviceSynchronize(); \
        fp_kernel<M><<<nb, 1024>>>(out, cyc, 1.0f); cudaDeviceSynchronize(); double c = avg_cyc(cyc, nb); \
        printf("%-12s lane-ops/clk/SM %.1f  warp-instr/clk/SM %.2f\n", fpn[M], 1024.0 * 8 * ITERS * lane_ops[M] / c, 32.0 * 8 * ITERS * (M == 5 ? 2 : 1) / c); }
    RUN_FP(0) RUN_FP(1) RUN_FP(2) RUN_FP(3) RUN_FP(4) RUN_FP(5)
    const char* smn[] = {"LDS.32", "LDS.64", "LDS.128", "STS.32", "STS.64", "STS.128", "SHFL", "LDS.128+8xFADD2"};
    const double bytes[] = {4, 8, 16, 4, 8, 16, 4, 16};
#define RUN_SM(M) { cudaFuncSetAttribute(smem_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840); \
        smem_kernel<M><<<nb, 1024, 163840>>>(out, cyc, 1); cudaDeviceSynchronize(); \
        smem_kernel<M><<<nb, 1024, 163840>>>(out, cyc, 1); cudaDeviceSynchronize(); double c = avg_cyc(cyc, nb); \
        printf("%-16s bytes/clk/SM %.1f  warp-instr/clk/SM %.3f\n", smn[M], 1024.0 * 8 * ITERS * bytes[M] / c, 32.0 * 8 * ITERS / c); }
    RUN_SM(0) RUN_SM(1) RUN_SM(2) RUN_SM(3) RUN_SM(4) RUN_SM(5) RUN_SM(6) RUN_SM(7)
    cudaError_t e = cudaGetLastError();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
