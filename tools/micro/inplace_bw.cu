// Speed-of-light probe for the update kernel's memory pattern: in-place read-modify-write of W and W_m
// ([V,H] fp32, 2 x 60 MB) with (a) a flat float4 grid-stride loop, (b) a persistent 148-CTA loop over
// 128x128 tiles in quarter-tile (128 rows x 32 cols) pieces, i.e. the exact address stream k_tc_stats issues.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/inplace_bw tools/micro/inplace_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_flat(float4* __restrict__ W, float4* __restrict__ M, size_t n4, float mom, float lr) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 w = W[i], m = M[i];
        m.x = mom * m.x - lr * w.x; m.y = mom * m.y - lr * w.y; m.z = mom * m.z - lr * w.z; m.w = mom * m.w - lr * w.w;
        w.x += m.x; w.y += m.y; w.z += m.z; w.w += m.w;
        W[i] = w; M[i] = m;
    }
}

// tile pattern: 256 threads; a quarter tile = 128 rows x 32 floats (128 B per row): 8 threads per row, 32 rows per pass
template <int UNROLL>
__global__ void k_tiles(float* __restrict__ W, float* __restrict__ M, int V, int H, float mom, float lr) {
    const int tr = (V + 127) / 128, tc = (H + 127) / 128, nt = tr * tc;
    const int r8 = threadIdx.x >> 3, c4 = (threadIdx.x & 7) * 4;
    for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const int r0 = (t / tc) * 128, c0 = (t % tc) * 128;
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * 32 + c4;
            if (c >= H) continue;
            float4 w[UNROLL], m[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int r = r0 + r8 + 32 * u;
                if (r < V) { w[u] = *(float4*)(W + (size_t)r * H + c); m[u] = *(float4*)(M + (size_t)r * H + c); }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int r = r0 + r8 + 32 * u;
                if (r < V) {
                    m[u].x = mom * m[u].x - lr * w[u].x; m[u].y = mom * m[u].y - lr * w[u].y;
                    m[u].z = mom * m[u].z - lr * w[u].z; m[u].w = mom * m[u].w - lr * w[u].w;
                    w[u].x += m[u].x; w[u].y += m[u].y; w[u].z += m[u].z; w[u].w += m[u].w;
                    *(float4*)(W + (size_t)r * H + c) = w[u]; *(float4*)(M + (size_t)r * H + c) = m[u];
                }
            }
        }
    }
}

// generic piece shape: a 128x128 tile is walked in pieces of RQ rows x CQ cols (CQ*4 bytes contiguous per row)
template <int RQ, int CQ>
__global__ void k_pieces(float* __restrict__ W, float* __restrict__ M, int V, int H, float mom, float lr) {
    constexpr int TPR = CQ / 4, RPP = 256 / TPR, U = RQ / RPP;     // threads per row, rows per pass, passes
    const int tr = (V + 127) / 128, tc = (H + 127) / 128, nt = tr * tc;
    const int rr = threadIdx.x / TPR, c4 = (threadIdx.x % TPR) * 4;
    for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const int r0 = (t / tc) * 128, c0 = (t % tc) * 128;
        for (int pr = 0; pr < 128; pr += RQ)
            for (int pc = 0; pc < 128; pc += CQ) {
                const int c = c0 + pc + c4;
                if (c >= H) continue;
                float4 w[U], m[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = r0 + pr + rr + RPP * u;
                    if (r < V) { w[u] = *(float4*)(W + (size_t)r * H + c); m[u] = *(float4*)(M + (size_t)r * H + c); }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = r0 + pr + rr + RPP * u;
                    if (r < V) {
                        m[u].x = mom * m[u].x - lr * w[u].x; m[u].y = mom * m[u].y - lr * w[u].y;
                        m[u].z = mom * m[u].z - lr * w[u].z; m[u].w = mom * m[u].w - lr * w[u].w;
                        w[u].x += m[u].x; w[u].y += m[u].y; w[u].z += m[u].z; w[u].w += m[u].w;
                        *(float4*)(W + (size_t)r * H + c) = w[u]; *(float4*)(M + (size_t)r * H + c) = m[u];
                    }
                }
            }
    }
}

// thread = tile row (the mapping the tensor-memory accumulator imposes on the update kernel's epilogue): 8 warps,
// warp w owns rows 32*(w&3)+lane and the 32-column quarters (w>>2) and (w>>2)+2; PF = quarters loaded ahead
template <int PF>
__global__ void __launch_bounds__(256, 1) k_rows(float* __restrict__ W, float* __restrict__ M, int V, int H, float mom, float lr) {
    const int tr = (V + 127) / 128, tc = (H + 127) / 128, nt = tr * tc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = (warp & 3) * 32 + lane, grp = warp >> 2;
    float4 w[PF + 1][8], m[PF + 1][8];
    auto addr = [&](int i, int& r, int& c) {            // i-th quarter of this thread's sequence
        const int t = blockIdx.x + (i >> 1) * gridDim.x;
        r = (t / tc) * 128 + row; c = (t % tc) * 128 + (grp + 2 * (i & 1)) * 32;
        return t < nt;
    };
    auto load = [&](int i, int b) {
        int r, c;
        if (!addr(i, r, c) || r >= V) return;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c + 4 * j < H) { w[b][j] = *(const float4*)(W + (size_t)r * H + c + 4 * j); m[b][j] = *(const float4*)(M + (size_t)r * H + c + 4 * j); }
    };
#pragma unroll
    for (int i = 0; i < PF; ++i) load(i, i);
    for (int i = 0;; ++i) {
        int r, c;
        if (!addr(i, r, c)) break;
#pragma unroll
        for (int b = 0; b <= PF; ++b) {
            if (i % (PF + 1) != b) continue;
            if (PF > 0) {
#pragma unroll
                for (int b2 = 0; b2 <= PF; ++b2) if ((i + PF) % (PF + 1) == b2) load(i + PF, b2);
            } else load(i, b);
            if (r < V)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (c + 4 * j >= H) continue;
                float4 ww = w[b][j], mm = m[b][j];
                mm.x = mom * mm.x - lr * ww.x; mm.y = mom * mm.y - lr * ww.y; mm.z = mom * mm.z - lr * ww.z; mm.w = mom * mm.w - lr * ww.w;
                ww.x += mm.x; ww.y += mm.y; ww.z += mm.z; ww.w += mm.w;
                *(float4*)(W + (size_t)r * H + c + 4 * j) = ww; *(float4*)(M + (size_t)r * H + c + 4 * j) = mm;
            }
        }
    }
}

__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int main() {
    const int V = 10000, H = 1500;
    const size_t n = (size_t)V * H;
    float *W, *M, *flush;
    CK(cudaMalloc(&W, n * 4)); CK(cudaMalloc(&M, n * 4)); CK(cudaMalloc(&flush, 256u << 20));
    CK(cudaMemset(W, 0, n * 4)); CK(cudaMemset(M, 0, n * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch, double bytes) {
        float best = 1e9f, sum = 0;
        for (int it = 0; it < 12; ++it) {
            cudaMemsetAsync(flush, it, 256u << 20);              // evict W / W_m from L2
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 2) { sum += ms; if (ms < best) best = ms; }
        }
        printf("%-28s avg %7.2f us  best %7.2f us  -> %6.0f GB/s (avg)\n", name, sum / 10 * 1e3, best * 1e3, bytes / (sum / 10 * 1e-3) / 1e9);
        return 0;
    };
    const double B = 4.0 * n * 4;
    for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32})
        for (int th : {256, 512}) {
            char nm[64]; snprintf(nm, 64, "flat g=%d t=%d", g, th);
            run(nm, [&] { k_flat<<<g, th>>>((float4*)W, (float4*)M, n / 4, 0.5f, 1e-3f); }, B);
        }
    for (int g : {148, 296, 592, 948}) {
        char nm[64]; snprintf(nm, 64, "tiles u4 g=%d", g);
        run(nm, [&] { k_tiles<4><<<g, 256>>>(W, M, V, H, 0.5f, 1e-3f); }, B);
    }
    for (int g : {148, 296, 444}) {
        char nm[64];
        snprintf(nm, 64, "pieces 128x32 g=%d", g); run(nm, [&] { k_pieces<128, 32><<<g, 256>>>(W, M, V, H, 0.5f, 1e-3f); }, B);
        snprintf(nm, 64, "pieces 64x64 g=%d", g);  run(nm, [&] { k_pieces<64, 64><<<g, 256>>>(W, M, V, H, 0.5f, 1e-3f); }, B);
        snprintf(nm, 64, "pieces 32x128 g=%d", g); run(nm, [&] { k_pieces<32, 128><<<g, 256>>>(W, M, V, H, 0.5f, 1e-3f); }, B);
        snprintf(nm, 64, "pieces 64x128 g=%d", g); run(nm, [&] { k_pieces<64, 128><<<g, 256>>>(W, M, V, H, 0.5f, 1e-3f); }, B);
        snprintf(nm, 64, "pieces 128x128 g=%d", g); run(nm, [&] { k_pieces<128, 128><<<g, 256>>>(W, M, V, H, 0.5f, 1e-3f); }, B);
    }
    for (int g : {132, 148, 296}) {
        char nm[64];
        snprintf(nm, 64, "rows pf0 g=%d", g); run(nm, [&] { k_rows<0><<<g, 256>>>(W, M, V, H, 0.5f, 1e-3f); }, B);
        snprintf(nm, 64, "rows pf1 g=%d", g); run(nm, [&] { k_rows<1><<<g, 256>>>(W, M, V, H, 0.5f, 1e-3f); }, B);
        snprintf(nm, 64, "rows pf2 g=%d", g); run(nm, [&] { k_rows<2><<<g, 256>>>(W, M, V, H, 0.5f, 1e-3f); }, B);
    }
    for (int g : {148 * 8, 148 * 16}) {
        char nm[64]; snprintf(nm, 64, "copy W->M g=%d", g);
        run(nm, [&] { k_copy<<<g, 512>>>((float4*)W, (float4*)M, n / 4); }, 2.0 * n * 4);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
