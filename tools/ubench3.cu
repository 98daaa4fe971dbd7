// Do shared-memory traffic and packed FP32 math overlap on one SM?  Each warp runs, per iteration,
// NL shared loads (or stores) of width WD floats and NF independent FADD2/FFMA2.  Reported: cycles per
// iteration per SM sub-partition for math only, memory only and both; "both" == max() means perfect
// overlap, == sum means none.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3 tools/ubench3.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 1024
template <int WD, int NL, int NF, bool ST, bool FMA2>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int domem, int domath) {
    extern __shared__ float4 sm4[];
    float* sm = reinterpret_cast<float*>(sm4);
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = out[i & 255];
    // 64 KB: two 32 KB halves alternate per iteration so the loads cannot be hoisted
    float2 p[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = make_float2(out[i], out[i + 32]); q[i] = make_float2(out[i + 64], out[i + 96]); }
    const float2 cc = make_float2(out[200], out[201]);
    float acc = 0.f; int acci = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* base = sm + (warp % 8) * 1024 + lane * WD;     // conflict-free: consecutive lanes, WD floats each
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (domem) {
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                float* a = base + l * 32 * WD + ((it & 1) * 8192);
                if (ST) {
                    acc = __int_as_float(it + l); if (WD == 1) a[0] = acc; if (WD == 2) *reinterpret_cast<float2*>(a) = make_float2(acc, acc);
                    if (WD == 4) *reinterpret_cast<float4*>(a) = make_float4(acc, acc, acc, acc);
                } else {
                    if (WD == 1) acci ^= __float_as_int(a[0]);
                    if (WD == 2) { float2 v = *reinterpret_cast<float2*>(a); acci ^= __float_as_int(v.x) ^ __float_as_int(v.y); }
                    if (WD == 4) { float4 v = *reinterpret_cast<float4*>(a); acci ^= __float_as_int(v.x) ^ __float_as_int(v.y) ^ __float_as_int(v.z) ^ __float_as_int(v.w); }
                }
            }
        }
        if (domath) {
#pragma unroll
            for (int f = 0; f < NF; ++f) p[f & 7] = FMA2 ? __ffma2_rn(p[f & 7], cc, q[f & 7]) : __fadd2_rn(p[f & 7], q[f & 7]);
        }
    }
    long long t1 = clock64();
    __syncthreads();
    float s = acc + __int_as_float(acci & 0x3fffff) + sm[threadIdx.x] + sm[8192 + threadIdx.x];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += p[i].x + p[i].y;
    out[4096 + blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int WD, int NL, int NF, bool ST, bool FMA2>
void run(const char* name, int nb, float* out, long long* cyc) {
    static long long h[1024];
    cudaFuncSetAttribute(k<WD, NL, NF, ST, FMA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int W : {8, 12, 16, 32}) {
        double r[3];
        int cfg[3][2] = {{0, 1}, {1, 0}, {1, 1}};
        for (int c = 0; c < 3; ++c) {
            for (int rep = 0; rep < 2; ++rep) { k<WD, NL, NF, ST, FMA2><<<nb, W * 32, 65536>>>(out, cyc, cfg[c][0], cfg[c][1]); cudaDeviceSynchronize(); }
            cudaMemcpy(h, cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < nb; ++i) s += h[i]; r[c] = s / nb / ITERS;
        }
        // per-SM cycles per (one iteration of every warp)
        printf("%-22s W=%2d  math %6.1f  mem %6.1f  both %6.1f  (max %6.1f sum %6.1f)  overlap %.2f\n", name, W, r[0], r[1], r[2],
               r[0] > r[1] ? r[0] : r[1], r[0] + r[1], (r[0] + r[1] - r[2]) / (r[0] < r[1] ? r[0] : r[1]));
    }
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int nb = prop.multiProcessorCount;
    float* out; long long* cyc; cudaMalloc(&out, sizeof(float) * (4096 + nb * 1024)); cudaMemset(out, 0, sizeof(float) * (4096 + nb * 1024)); cudaMalloc(&cyc, sizeof(long long) * nb);
    run<1, 16, 16, false, false>("LDS.32 x16 + FADD2 x16", nb, out, cyc);
    run<2, 8, 16, false, false>("LDS.64 x8 + FADD2 x16", nb, out, cyc);
    run<4, 4, 16, false, false>("LDS.128 x4 + FADD2 x16", nb, out, cyc);
    run<1, 16, 16, true, false>("STS.32 x16 + FADD2 x16", nb, out, cyc);
    run<2, 8, 16, true, false>("STS.64 x8 + FADD2 x16", nb, out, cyc);
    run<2, 8, 16, false, true>("LDS.64 x8 + FFMA2 x16", nb, out, cyc);
    run<1, 8, 32, false, false>("LDS.32 x8 + FADD2 x32", nb, out, cyc);
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
