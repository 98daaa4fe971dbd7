// TMEM as a per-thread scratch ("register extension"): tcgen05.st / tcgen05.ld 32x32b round trips.
// Checks data integrity and reports cycles per 16-word store+load round trip per warp at 4 / 8 / 12 / 16 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench5 tools/ubench5.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 256
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int COLS, int MODE>   // MODE 0: st+ld round trip, 1: ld only, 2: st only
__global__ void __launch_bounds__(128) k(unsigned* bad, long long* cyc, int extra) {
    __shared__ uint32_t tbase_s;
    extern __shared__ float dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tbase_s)), "n"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase_s + ((uint32_t)(warp * 32) << 16);       // this warp's lane quarter
    uint32_t r[16], v[16];
    unsigned errs = 0;
    // integrity: fill all columns with a thread/column signature, read back
    for (int c = 0; c < COLS; c += 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = (blockIdx.x << 20) ^ (threadIdx.x << 10) ^ (c + i);
        tm_st16(tb + c, r);
    }
    tm_wait_st();
    for (int c = 0; c < COLS; c += 16) {
        tm_ld16(tb + c, v);
        tm_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) errs += v[i] != ((blockIdx.x << 20) ^ (threadIdx.x << 10) ^ (c + i));
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const uint32_t a = tb + ((it * 16) & (COLS - 1));
        if (MODE != 1) { tm_st16(a, r); tm_wait_st(); }
        if (MODE != 2) { tm_ld16(a, v); tm_wait_ld(); }
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = v[i] + 1u + r[i];
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 0x12345678u) errs += 1u << 30;
    if (errs) atomicAdd(bad, errs & 0xffff ? errs : 0u);
    if (lane == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase_s), "n"(COLS) : "memory");
    if (extra == 12345) dyn[threadIdx.x] = 0.f;
}
template <int COLS, int MODE>
void run(int ctas_per_sm, const char* name) {
    unsigned* bad; long long* cyc;
    cudaMalloc(&bad, 4); cudaMalloc(&cyc, 64); cudaMemset(bad, 0, 4);
    // limit residency with dynamic shared memory
    size_t smem = (size_t)(227 * 1024 / ctas_per_sm) - 2048;
    cudaFuncSetAttribute(k<COLS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<COLS, MODE><<<148 * ctas_per_sm, 128, smem>>>(bad, cyc, 0);
    cudaDeviceSynchronize();
    cudaMemset(bad, 0, 4);
    cudaEventRecord(e0);
    k<COLS, MODE><<<148 * ctas_per_sm, 128, smem>>>(bad, cyc, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    unsigned hb; long long hc[4];
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 32, cudaMemcpyDeviceToHost);
    printf("%-10s cols %3d  %d CTA/SM (%2d warps): %s  mismatches %u  cycles per 16-word op per warp: %.1f  kernel %.1f us\n", name, COLS, ctas_per_sm,
           4 * ctas_per_sm, cudaGetErrorString(e), hb, (double)hc[0] / ITERS, ms * 1e3);
    cudaFree(bad); cudaFree(cyc);
}
int main() {
    for (int c = 1; c <= 3; ++c) { run<128, 0>(c, "st+ld"); run<128, 1>(c, "ld"); run<128, 2>(c, "st"); }
    run<256, 0>(2, "st+ld"); run<128, 0>(4, "st+ld"); run<128, 1>(4, "ld");
    // residency map: kernel time for c CTAs per SM (fixed work per CTA) by allocation size
    for (int c = 1; c <= 6; ++c) { run<32, 0>(c, "map32"); }
    for (int c = 1; c <= 6; ++c) { run<64, 0>(c, "map64"); }
    for (int c = 1; c <= 4; ++c) { run<128, 0>(c, "map128"); }
    for (int c = 1; c <= 3; ++c) { run<256, 0>(c, "map256"); }
    for (int c = 1; c <= 2; ++c) { run<512, 0>(c, "map512"); }
    return 0;
}
