// TMA store / load rate probe on the update kernel's address pattern ([V,H] fp32, H*4-byte pitch):
// 148 CTAs walk 128x128 tiles round-robin and move them in 16 KB pieces with DEPTH pieces in flight.
//   shape 0: piece = box [128 rows x 32 floats], SWIZZLE_128B       (what k_tc_stats uses)
//   shape 1: piece = box [ 32 rows x 128 floats], no swizzle         (512 B contiguous per row)
//   shape 2: piece = 32 x cp.async.bulk of 512 B (1-D, no tensor map)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/tma_rate tools/micro/tma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../multimodal_idbn_b200/csrc/tc_ptx.cuh"
using namespace imdbn::ptx;
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int N> __device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int SHAPE, int DEPTH, bool LOAD>
__global__ void __launch_bounds__(128, 1) k_rate(const __grid_constant__ CUtensorMap tm, float* W, int V, int H) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[16];
    const int tr = (V + 127) / 128, tc = (H + 127) / 128, nt = tr * tc;
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
    __syncthreads();
    if (SHAPE == 2 ? (threadIdx.x < 32) : (threadIdx.x == 0)) {
        const int lane = threadIdx.x;
        int i = 0;
        for (int t = blockIdx.x; t < nt; t += gridDim.x) {
            const int r0 = (t / tc) * 128, c0 = (t % tc) * 128;
            for (int p = 0; p < 4; ++p, ++i) {
                uint8_t* slot = sm + (i % DEPTH) * 16384;
                uint64_t* bar = &bars[i % DEPTH];
                if (LOAD) {
                    if (i >= DEPTH) mbar_wait(bar, ((i / DEPTH) - 1) & 1);
                    if (SHAPE == 0) { mbar_expect_tx(bar, 16384); tma_load_2d(slot, &tm, c0 + 32 * p, r0, bar); }
                    if (SHAPE == 1) { mbar_expect_tx(bar, 16384); tma_load_2d(slot, &tm, c0, r0 + 32 * p, bar); }
                    if (SHAPE == 2) {
                        const int r = r0 + 32 * p + lane;
                        const int nb = min(128, H - c0) * 4;
                        int rows = min(32, V - (r0 + 32 * p)); if (rows < 0) rows = 0;
                        if (lane == 0) mbar_expect_tx(bar, rows * nb);
                        __syncwarp();
                        if (r < V) bulk_load_1d(slot + lane * 512, W + (size_t)r * H + c0, nb, bar);
                    }
                } else {
                    if (SHAPE == 0) tma_store_2d(&tm, c0 + 32 * p, r0, slot);
                    if (SHAPE == 1) tma_store_2d(&tm, c0, r0 + 32 * p, slot);
                    if (SHAPE == 2) {
                        const int r = r0 + 32 * p + lane;
                        if (r < V) bulk_store_1d(W + (size_t)r * H + c0, slot + lane * 512, min(128, H - c0) * 4);
                    }
                    tma_store_commit();
                    wait_read<DEPTH - 1>();
                }
            }
        }
        if (LOAD) { for (int j = max(0, i - DEPTH); j < i; ++j) mbar_wait(&bars[j % DEPTH], (j / DEPTH) & 1); }
        else tma_store_wait_all();
    }
}

// TMA copy: load piece i into slot i%D, store it back from the slot LAG pieces later (same engine, both directions)
template <int D, int LAG>
__global__ void __launch_bounds__(128, 1) k_copy_tma(const __grid_constant__ CUtensorMap tm, float* W, int V, int H) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[16];
    const int tr = (V + 127) / 128, tc = (H + 127) / 128, nt = tr * tc;
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int my = 0; for (int t = blockIdx.x; t < nt; t += gridDim.x) ++my;
    const int N = my * 4;
    for (int i = 0; i < N + LAG; ++i) {
        if (i < N) {
            if (i >= D) wait_read<D - LAG - 1>();
            const int t = blockIdx.x + (i >> 2) * gridDim.x, p = i & 3;
            mbar_expect_tx(&bars[i % D], 16384);
            tma_load_2d(sm + (i % D) * 16384, &tm, (t % tc) * 128 + 32 * p, (t / tc) * 128, &bars[i % D]);
        }
        const int j = i - LAG;
        if (j >= 0 && j < N) {
            mbar_wait(&bars[j % D], (j / D) & 1);
            const int t = blockIdx.x + (j >> 2) * gridDim.x, p = j & 3;
            tma_store_2d(&tm, (t % tc) * 128 + 32 * p, (t / tc) * 128, sm + (j % D) * 16384);
            tma_store_commit();
        }
    }
    tma_store_wait_all();
}

// TMA copy with row-slices: a 16 KB piece = [32 rows x 128 floats] = four [32 x 32] swizzled boxes issued back to back
// (512 contiguous bytes per weight row in flight at once instead of 128)
template <int D, int LAG>
__global__ void __launch_bounds__(128, 1) k_copy_tma_rows(const __grid_constant__ CUtensorMap tm, float* W, int V, int H) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[16];
    const int tr = (V + 127) / 128, tc = (H + 127) / 128, nt = tr * tc;
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int my = 0; for (int t = blockIdx.x; t < nt; t += gridDim.x) ++my;
    const int N = my * 4;
    for (int i = 0; i < N + LAG; ++i) {
        if (i < N) {
            if (i >= D) wait_read<D - LAG - 1>();
            const int t = blockIdx.x + (i >> 2) * gridDim.x, p = i & 3;
            mbar_expect_tx(&bars[i % D], 16384);
            for (int cb = 0; cb < 4; ++cb)
                tma_load_2d(sm + (i % D) * 16384 + cb * 4096, &tm, (t % tc) * 128 + 32 * cb, (t / tc) * 128 + 32 * p, &bars[i % D]);
        }
        const int j = i - LAG;
        if (j >= 0 && j < N) {
            mbar_wait(&bars[j % D], (j / D) & 1);
            const int t = blockIdx.x + (j >> 2) * gridDim.x, p = j & 3;
            for (int cb = 0; cb < 4; ++cb)
                tma_store_2d(&tm, (t % tc) * 128 + 32 * cb, (t / tc) * 128 + 32 * p, sm + (j % D) * 16384 + cb * 4096);
            tma_store_commit();
        }
    }
    tma_store_wait_all();
}

// mixed copy: TMA loads into a ring of D swizzled 16 KB slots, NT consumer threads read each slot from shared memory
// and write it back with coalesced 16-byte global stores (8 lanes per 128-byte row piece)
template <int D, int NT>
__global__ void __launch_bounds__(NT + 32, 1) k_copy_mixed(const __grid_constant__ CUtensorMap tm, float* W, int V, int H) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[16], empty[16];
    const int tr = (V + 127) / 128, tc = (H + 127) / 128, nt = tr * tc;
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NT / 32); } fence_barrier_init(); }
    __syncthreads();
    int my = 0; for (int t = blockIdx.x; t < nt; t += gridDim.x) ++my;
    const int N = my * 4;
    if (threadIdx.x >= NT) {
        if (threadIdx.x == NT)
            for (int i = 0; i < N; ++i) {
                if (i >= D) mbar_wait(&empty[i % D], ((i / D) - 1) & 1);
                const int t = blockIdx.x + (i >> 2) * gridDim.x, p = i & 3;
                mbar_expect_tx(&full[i % D], 16384);
                tma_load_2d(sm + (i % D) * 16384, &tm, (t % tc) * 128 + 32 * p, (t / tc) * 128, &full[i % D]);
            }
        return;
    }
    const int ch = threadIdx.x & 7, rr = threadIdx.x >> 3;
    constexpr int RPP = NT / 8;
    for (int i = 0; i < N; ++i) {
        const int t = blockIdx.x + (i >> 2) * gridDim.x, p = i & 3;
        const int r0 = (t / tc) * 128, c = (t % tc) * 128 + 32 * p + ch * 4;
        const uint8_t* slot = sm + (i % D) * 16384;
        mbar_wait(&full[i % D], (i / D) & 1);
        float4 v[128 / RPP];
#pragma unroll
        for (int u = 0; u < 128 / RPP; ++u) {
            const int r = rr + RPP * u;
            v[u] = *reinterpret_cast<const float4*>(slot + r * 128 + ((ch ^ (r & 7)) << 4));
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[i % D]);
#pragma unroll
        for (int u = 0; u < 128 / RPP; ++u) {
            const int r = r0 + rr + RPP * u;
            if (r < V && c < H) { v[u].x += 1.0f; *reinterpret_cast<float4*>(W + (size_t)r * H + c) = v[u]; }
        }
    }
}

// LSU path: 128 threads issue 16-byte cp.async (LDGSTS) into the 128B-swizzled slot layout; completion by
// cp.async.mbarrier.arrive.noinc
template <int D>
__global__ void __launch_bounds__(128, 1) k_load_lsu(const __grid_constant__ CUtensorMap tm, float* W, int V, int H) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[16];
    const int tr = (V + 127) / 128, tc = (H + 127) / 128, nt = tr * tc;
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 128); fence_barrier_init(); }
    __syncthreads();
    const int ch = threadIdx.x & 7, rr = threadIdx.x >> 3;          // 16-byte chunk in the row, row in the pass (16 rows/pass)
    int i = 0;
    for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const int r0 = (t / tc) * 128, c0 = (t % tc) * 128;
        for (int p = 0; p < 4; ++p, ++i) {
            uint8_t* slot = sm + (i % D) * 16384;
            uint64_t* bar = &bars[i % D];
            if (i >= D) mbar_wait(bar, ((i / D) - 1) & 1);
            const int c = c0 + 32 * p + ch * 4;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = rr + 16 * u;
                if (r0 + r < V && c < H) {
                    const uint32_t dst = smem_u32(slot + r * 128 + ((ch ^ (r & 7)) << 4));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(W + (size_t)(r0 + r) * H + c) : "memory");
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
    }
    for (int j = max(0, i - D); j < i; ++j) mbar_wait(&bars[j % D], (j / D) & 1);
}

int main() {
    const int V = 10000, H = 1500;
    const size_t n = (size_t)V * H;
    float *W, *flush;
    CK(cudaMalloc(&W, n * 4)); CK(cudaMalloc(&flush, 256u << 20)); CK(cudaMemset(W, 0, n * 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    CUtensorMap m0, m1;
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)V}; cuuint64_t str[1] = {(cuuint64_t)H * 4}; cuuint32_t es[2] = {1, 1};
    cuuint32_t b0[2] = {32, 128}, b1[2] = {128, 32};
    if (enc(&m0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dims, str, b0, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc0 failed\n"); return 1; }
    if (enc(&m1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dims, str, b1, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc1 failed\n"); return 1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int SM = 12 * 16384 + 1024;
    int G = 148;
    auto run = [&](const char* name, auto kern, const CUtensorMap& tm) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
        float sum = 0;
        for (int it = 0; it < 8; ++it) {
            cudaMemsetAsync(flush, it, 256u << 20);
            cudaEventRecord(e0); kern<<<G, 128, SM>>>(tm, W, V, H); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2) sum += ms;
        }
        cudaError_t e = cudaGetLastError();
        printf("%-34s %7.2f us  %6.0f GB/s %s\n", name, sum / 6 * 1e3, n * 4.0 / (sum / 6 * 1e-3) / 1e9, e ? cudaGetErrorString(e) : "");
    };
#define RUN(S, D, L, TM) run(#L " shape" #S " depth" #D, k_rate<S, D, L>, TM)
    RUN(0, 1, false, m0); RUN(0, 2, false, m0); RUN(0, 4, false, m0); RUN(0, 8, false, m0);
    RUN(1, 1, false, m1); RUN(1, 2, false, m1); RUN(1, 4, false, m1); RUN(1, 8, false, m1);
    RUN(2, 1, false, m0); RUN(2, 2, false, m0); RUN(2, 4, false, m0); RUN(2, 8, false, m0);
    RUN(0, 1, true, m0); RUN(0, 2, true, m0); RUN(0, 4, true, m0); RUN(0, 8, true, m0); RUN(0, 12, true, m0);
    RUN(1, 1, true, m1); RUN(1, 2, true, m1); RUN(1, 4, true, m1); RUN(1, 8, true, m1); RUN(1, 12, true, m1);
    RUN(2, 1, true, m0); RUN(2, 2, true, m0); RUN(2, 4, true, m0); RUN(2, 8, true, m0); RUN(2, 12, true, m0);
    run("copy tma D8 lag4", k_copy_tma<8, 4>, m0);
    run("copy tma D12 lag6", k_copy_tma<12, 6>, m0);
    run("copy tma D12 lag8", k_copy_tma<12, 8>, m0);
    {
        CUtensorMap m2; cuuint32_t b2[2] = {32, 32};
        if (enc(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dims, str, b2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc2 failed\n"); return 1; }
        run("copy tma rows32 D8 lag4", k_copy_tma_rows<8, 4>, m2);
        run("copy tma rows32 D12 lag6", k_copy_tma_rows<12, 6>, m2);
        run("copy tma rows32 D12 lag8", k_copy_tma_rows<12, 8>, m2);
    }
    {
        auto runm = [&](const char* name, auto kern, int nthr) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
            float sum = 0;
            for (int it = 0; it < 8; ++it) {
                cudaMemsetAsync(flush, it, 256u << 20);
                cudaEventRecord(e0); kern<<<G, nthr, SM>>>(m0, W, V, H); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2) sum += ms;
            }
            cudaError_t e = cudaGetLastError();
            printf("%-34s %7.2f us  %6.0f GB/s (in+out) %s\n", name, sum / 6 * 1e3, 2 * n * 4.0 / (sum / 6 * 1e-3) / 1e9, e ? cudaGetErrorString(e) : "");
        };
        runm("copy mixed D8 128thr", k_copy_mixed<8, 128>, 160);
        runm("copy mixed D12 128thr", k_copy_mixed<12, 128>, 160);
        runm("copy mixed D8 256thr", k_copy_mixed<8, 256>, 288);
        runm("copy mixed D12 256thr", k_copy_mixed<12, 256>, 288);
        for (int g : {132, 296}) { G = g; char nm[64]; snprintf(nm, 64, "copy mixed D12 256thr g=%d", g); if (g == 296) { /* two CTAs per SM need half the ring */ } runm(nm, k_copy_mixed<12, 256>, 288); }
        G = 148;
    }
    run("load lsu D2", k_load_lsu<2>, m0); run("load lsu D4", k_load_lsu<4>, m0); run("load lsu D8", k_load_lsu<8>, m0);
    run("load lsu D12", k_load_lsu<12>, m0);
    G = 74;
    printf("--- 74 CTAs\n");
    RUN(0, 4, false, m0); RUN(0, 8, true, m0); run("copy tma D12 lag6", k_copy_tma<12, 6>, m0); run("load lsu D8", k_load_lsu<8>, m0);
    CK(cudaDeviceSynchronize());
    return 0;
}
