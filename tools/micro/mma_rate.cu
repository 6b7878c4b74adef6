// tcgen05.mma kind::tf32 issue/execute rate probe: one CTA per SM issues REPS x 8 independent-address MMAs
// (M = 128, K = 8, N in {64, 128, 256}) from one thread, then commits and waits.  Operand contents are irrelevant.
//   mode 0: A from shared memory, K-major SW128      mode 1: A from shared memory, MN-major (32-byte atoms)
//   mode 2: A from tensor memory
//   mode 3: A from shared memory, K-major, 32-byte swizzle (one 32-byte row per k-group: the tile of one MMA is 4 KB contiguous)
//   mode 4: A from shared memory, K-major, 64-byte swizzle
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../multimodal_idbn_b200/csrc/tc_ptx.cuh"
using namespace imdbn::ptx;
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(128, 1) k_rate(int mode, int N, int reps, int two_acc, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1e-3f * (float)(i & 255);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t sA = smem_u32(sm), sB = smem_u32(sm + 32 * 1024);
        const uint32_t idesc = idesc_tf32(128, N, mode == 1, false, false);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const uint32_t d = tm + ((two_acc && g >= 4) ? (uint32_t)N : 0u);
                const uint64_t bd = smem_desc(sB + (g / 4) * (N * 128) + (g % 4) * 32, 16, 1024, LAYOUT_SW128);
                if (mode == 2) {
                    mma_tf32_ts(d, tm + 2 * N + (uint32_t)(g * 8), bd, idesc, 1u);
                } else {
                    const uint64_t ad = mode == 1 ? smem_desc(sA + g * 1024, 64 * 128, 512, LAYOUT_SW128_BASE32B)
                                      : mode == 3 ? smem_desc(sA + g * 4096, 16, 256, 6)                  // 8 rows x 32 B atoms
                                      : mode == 4 ? smem_desc(sA + (g / 2) * 8192 + (g % 2) * 32, 16, 512, 4)   // 8 rows x 64 B atoms
                                                  : smem_desc(sA + (g / 4) * (128 * 128) + (g % 4) * 32, 16, 1024, LAYOUT_SW128);
                    mma_tf32(d, ad, bd, idesc, 1u);
                }
            }
        }
        const long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    long long* out; CK(cudaMalloc(&out, 148 * 2 * sizeof(long long)));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int reps = 200;
    for (int two = 0; two < 2; ++two)
    for (int mode = 0; mode < 5; ++mode)
        for (int N : {64, 128, 256}) {
            if (2 * N + 64 > 512 && mode == 2) continue;
            if (two && 2 * N > 512) continue;
            k_rate<<<148, 128, 200 * 1024>>>(mode, N, reps, two, out);
            CK(cudaDeviceSynchronize());
            long long h[296]; CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
            double iss = 0, tot = 0; for (int i = 0; i < 148; ++i) { iss += h[2 * i]; tot += h[2 * i + 1]; }
            printf("mode %d (%s) N=%3d accumulators=%d: issue %.1f clk/MMA, issue+execute %.1f clk/MMA  (ideal %d)\n", mode,
                   mode == 0 ? "A smem K-major SW128" : mode == 1 ? "A smem MN-major" : mode == 2 ? "A tmem" : mode == 3 ? "A smem K-major SW32" : "A smem K-major SW64", N, two + 1, iss / 148 / (reps * 8),
                   tot / 148 / (reps * 8), 128 * N * 8 / 2048);
        }
    return 0;
}
