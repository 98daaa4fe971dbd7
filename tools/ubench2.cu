// How many resident warps per SM sub-partition does it take to saturate the FP32 pipe with
// packed (FADD2/FFMA2) and scalar ops?  One CTA per SM with W warps (W/4 per SMSP), ILP-way
// independent chains per thread.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench2 tools/ubench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE, int ILP>
__global__ void k(float* out, long long* cyc, float seed) {
    float2 p[ILP], q[ILP]; float a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { p[i] = make_float2(out[i], out[i + 32]); q[i] = make_float2(out[i + 64], out[i + 96]); a[i] = out[i + 128]; b[i] = out[i + 160]; }
    const float2 cc = make_float2(seed * 1.0001f, seed * 0.9999f);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) p[i] = __fadd2_rn(p[i], q[i]);
            if (MODE == 1) p[i] = __ffma2_rn(p[i], cc, q[i]);
            if (MODE == 2) a[i] = a[i] + b[i];
            if (MODE == 3) a[i] = fmaf(a[i], cc.x, b[i]);
            if (MODE == 4) { p[i] = __fadd2_rn(p[i], q[i]); a[i] = (a[i] > b[i]) ? a[i] : b[i] + 1.0f; }   // packed + ALU-ish
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += p[i].x + p[i].y + a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP>
void run(const char* name, int nb, float* out, long long* cyc) {
    static long long h[1024];
    printf("%-10s ILP %2d:", name, ILP);
    for (int W : {4, 8, 12, 16, 32}) {
        k<MODE, ILP><<<nb, W * 32>>>(out, cyc, 1.0f); cudaDeviceSynchronize();
        k<MODE, ILP><<<nb, W * 32>>>(out, cyc, 1.0f); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < nb; ++i) c += h[i]; c /= nb;
        // warp-instructions per clock per SMSP
        printf("  W=%2d %.3f", W, (double)(W / 4) * ILP * ITERS / c);
    }
    printf("   (warp-instr/clk/SMSP)\n");
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int nb = prop.multiProcessorCount;
    float* out; long long* cyc; cudaMalloc(&out, sizeof(float) * nb * 1024); cudaMemset(out, 0, sizeof(float) * nb * 1024); cudaMalloc(&cyc, sizeof(long long) * nb);
    run<0, 1>("FADD2", nb, out, cyc); run<0, 2>("FADD2", nb, out, cyc); run<0, 4>("FADD2", nb, out, cyc); run<0, 8>("FADD2", nb, out, cyc); run<0, 16>("FADD2", nb, out, cyc);
    run<1, 1>("FFMA2", nb, out, cyc); run<1, 4>("FFMA2", nb, out, cyc); run<1, 8>("FFMA2", nb, out, cyc); run<1, 16>("FFMA2", nb, out, cyc);
    run<2, 1>("FADD", nb, out, cyc); run<2, 4>("FADD", nb, out, cyc); run<2, 8>("FADD", nb, out, cyc); run<2, 16>("FADD", nb, out, cyc);
    run<3, 1>("FFMA", nb, out, cyc); run<3, 8>("FFMA", nb, out, cyc); run<3, 16>("FFMA", nb, out, cyc);
    run<4, 8>("FADD2+ALU", nb, out, cyc);
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
