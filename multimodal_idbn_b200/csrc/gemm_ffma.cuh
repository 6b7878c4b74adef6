// fp32 FFMA GEMM with arbitrary operand strides: the parity-mode engine behind the up pass
// (v W), the down pass (h W^T) and the CD statistics (v^T h) -- imdbn/models/rbm.py:92,96,200,209.
//
//   C[m,n] = sum_k A(m,k) B(k,n)  (- sum_k A2(m,k) B2(k,n) when a second segment is given)
//
// Tile 64(M) x 128(N) x 16(K), 256 threads, 4x8 register micro-tile, register-staged double
// buffering.  blockIdx.z splits K; every split writes its own partial slab (deterministic,
// summed by the finish kernels).  EPI_UPDATE fuses the momentum / weight-decay update of
// rbm.py:212-213 into the epilogue of the statistics GEMM.
#pragma once
#include "common.cuh"

namespace imdbn {

constexpr int GBM = 64, GBN = 128, GBK = 16, GTHREADS = 256;
constexpr int GAP = GBM + 4;   // padded smem row strides (floats): keep float4 alignment,
constexpr int GBP = GBN + 4;   // spread banks for the transposing stores

enum { EPI_STORE = 0, EPI_UPDATE = 1 };

struct GemmArgs {
    const float* A;  long long sAm, sAk;
    const float* B;  long long sBk, sBn;
    const float* A2; const float* B2; int K2;   // optional subtracted segment (same strides)
    int M, N, K;
    int k_per_split;
    float* C;               // EPI_STORE: [splits][M][N]
    // EPI_UPDATE (m = visible unit, n = hidden unit)
    float* W; float* Wm;
    float lr, mom, wd, bsz;
};

template <bool A_KCONTIG>
__device__ __forceinline__ void load_a(const GemmArgs& g, const float* __restrict__ A, int m0, int k0,
                                       int kend, float (&r)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * GTHREADS;
        int m, k;
        if (A_KCONTIG) { m = idx / GBK; k = idx % GBK; } else { k = idx / GBM; m = idx % GBM; }
        const int gm = m0 + m, gk = k0 + k;
        r[i] = (gm < g.M && gk < kend) ? __ldg(A + gm * g.sAm + gk * g.sAk) : 0.0f;
    }
}

template <bool A_KCONTIG>
__device__ __forceinline__ void store_a(float (*As)[GAP], const float (&r)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * GTHREADS;
        int m, k;
        if (A_KCONTIG) { m = idx / GBK; k = idx % GBK; } else { k = idx / GBM; m = idx % GBM; }
        As[k][m] = r[i];
    }
}

template <bool B_NCONTIG>
__device__ __forceinline__ void load_b(const GemmArgs& g, const float* __restrict__ Bp, int n0, int k0,
                                       int kend, float sign, float (&r)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = threadIdx.x + i * GTHREADS;
        int n, k;
        if (B_NCONTIG) { k = idx / GBN; n = idx % GBN; } else { n = idx / GBK; k = idx % GBK; }
        const int gn = n0 + n, gk = k0 + k;
        r[i] = (gn < g.N && gk < kend) ? sign * __ldg(Bp + gk * g.sBk + gn * g.sBn) : 0.0f;
    }
}

template <bool B_NCONTIG>
__device__ __forceinline__ void store_b(float (*Bs)[GBP], const float (&r)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = threadIdx.x + i * GTHREADS;
        int n, k;
        if (B_NCONTIG) { k = idx / GBN; n = idx % GBN; } else { n = idx / GBK; k = idx % GBK; }
        Bs[k][n] = r[i];
    }
}

template <bool A_KCONTIG, bool B_NCONTIG, int EPI>
__global__ void __launch_bounds__(GTHREADS) k_gemm_ffma(GemmArgs g) {
    __shared__ __align__(16) float As[GBK][GAP];
    __shared__ __align__(16) float Bs[GBK][GBP];
    pdl_trigger();
    pdl_wait();

    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    const int nseg = g.A2 ? 2 : 1;
    for (int seg = 0; seg < nseg; ++seg) {
        const float* A = seg ? g.A2 : g.A;
        const float* Bp = seg ? g.B2 : g.B;
        const float sign = seg ? -1.0f : 1.0f;
        const int Kseg = seg ? g.K2 : g.K;
        int kbeg = 0, kend = Kseg;
        if (nseg == 1) {
            kbeg = blockIdx.z * g.k_per_split;
            kend = min(Kseg, kbeg + g.k_per_split);
        }
        if (kbeg >= kend) continue;

        float ra[4], rb[8];
        load_a<A_KCONTIG>(g, A, m0, kbeg, kend, ra);
        load_b<B_NCONTIG>(g, Bp, n0, kbeg, kend, sign, rb);
        for (int k0 = kbeg; k0 < kend; k0 += GBK) {
            __syncthreads();                 // previous tile fully consumed
            store_a<A_KCONTIG>(As, ra);
            store_b<B_NCONTIG>(Bs, rb);
            __syncthreads();
            if (k0 + GBK < kend) {           // prefetch the next tile while computing this one
                load_a<A_KCONTIG>(g, A, m0, k0 + GBK, kend, ra);
                load_b<B_NCONTIG>(g, Bp, n0, k0 + GBK, kend, sign, rb);
            }
#pragma unroll
            for (int kk = 0; kk < GBK; ++kk) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= g.N) continue;
            const size_t o = (size_t)m * g.N + n;
            if (EPI == EPI_STORE) {
                g.C[(size_t)blockIdx.z * g.M * g.N + o] = acc[i][j];
            } else {
                // W_m <- mom*W_m + lr*((S+ - S-)/bsz - wd*W);  W <- W + W_m   (rbm.py:212-213)
                const float w = g.W[o];
                const float grad = add_rn(div_by(acc[i][j], g.bsz, 1.0f / g.bsz), -mul_rn(g.wd, w));
                const float wm = add_rn(mul_rn(g.Wm[o], g.mom), mul_rn(g.lr, grad));
                g.Wm[o] = wm;
                g.W[o] = add_rn(w, wm);
            }
        }
    }
}

// Pick a K split so a skinny (small-batch) product still fills the machine.
inline int choose_splits(int M, int N, int K, int num_sms) {
    const int tiles = ((M + GBM - 1) / GBM) * ((N + GBN - 1) / GBN);
    if (tiles >= 2 * num_sms) return 1;
    int s = (2 * num_sms + tiles - 1) / tiles;
    const int max_by_k = (K + 8 * GBK - 1) / (8 * GBK);   // keep >= 128 k per split
    if (s > max_by_k) s = max_by_k;
    if (s > 32) s = 32;
    return s < 1 ? 1 : s;
}

}  // namespace imdbn
