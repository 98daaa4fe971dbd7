// Shared-memory pipe cost per INSTRUCTION vs per WAVEFRONT: conflict-free LDS/STS of 32 / 64 / 128 bits,
// W warps per SM, all SMs busy.  Reports SM cycles per warp-instruction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench6 tools/ubench6.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <typename T, bool STORE>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc) {
    extern __shared__ float4 sm4[];
    T* sm = reinterpret_cast<T*>(sm4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* p = sm + warp * 32 * 8 + lane;       // 8 rows of 32 elements per warp, conflict-free
    T acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = p[i * 32];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (STORE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { asm volatile("" ::: "memory"); p[i * 32] = acc[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned a = (unsigned)__cvta_generic_to_shared(&p[i * 32]);
                float v0, v1, v2, v3;
                if (sizeof(T) == 4) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(a)); }
                else if (sizeof(T) == 8) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v0), "=f"(v1) : "r"(a)); v0 += v1; }
                else { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(a)); v0 += v1 + v2 + v3; }
                reinterpret_cast<float*>(&acc[i])[0] += v0;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += reinterpret_cast<float*>(&acc[i])[0];
    if (s == 1234.5f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <typename T, bool STORE>
void run(int warps, const char* name) {
    long long* cyc; float* out;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 4);
    size_t smem = (size_t)warps * 32 * 8 * sizeof(T);
    cudaFuncSetAttribute(k<T, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<T, STORE><<<148, warps * 32, smem>>>(out, cyc);
    cudaDeviceSynchronize();
    long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-8s %2d warps/SM: %.2f SM-cycles per warp-instruction (%d B/thread)\n", name, warps, (double)hc / ITERS / 8 / warps, (int)sizeof(T));
    cudaFree(cyc); cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<float, false>(w, "LDS.32"); run<float2, false>(w, "LDS.64"); run<float4, false>(w, "LDS.128");
        run<float, true>(w, "STS.32"); run<float2, true>(w, "STS.64"); run<float4, true>(w, "STS.128");
    }
    return 0;
}
