// IMG->TXT conditional Gibbs with the image latents clamped (iMDBN._cross_reconstruct, imdbn.py:419-427;
// rbm.py:391-399): when the first Dz visible units are clamped and the remaining K <= 32 units form one
// softmax group, a mean-field sweep only ever changes the K label units:
//     h = sigmoid( c + y W_y ),   c = z W_z + b_h   (constant per chain, one GEMM up front)
//     y = softmax( h W_y^T + b_y )
// so the whole chain is 2*K*H MACs per step instead of 2*V*H.  One WARP owns one chain: lane <-> label,
// the label block W_y (K x H fp32, 32 KB for the joint RBM) and its transpose sit in shared memory for
// all warps of the CTA, the softmax max / sum are warp shuffles, and all n steps run inside the kernel.
#pragma once
#include "common.cuh"

namespace imdbn {

constexpr int LG_WARPS = 8;
constexpr int LG_MAXH = 256;      // hidden units supported (multiple of 32)
constexpr int LG_CHAINS = 4;      // chains per warp

struct LabelGibbsArgs {
    const float* pre;     // [B,H]  c = z W_z + b_h
    const float* Wy;      // [K,H]  rows Dz..Dz+K of W (contiguous)
    const float* vby;     // [K]
    int B, H, K, Dz, n_steps;
    RngKey key; uint32_t draw0;
    float* y_out;         // [B,K] label state after n_steps clamped sweeps
};

// FAST (tf32 mode): __expf / __fdividef sigmoid and softmax, 1e-6 relative, far inside the tf32 tolerance.
template <bool FAST>
__global__ void __launch_bounds__(LG_WARPS * 32) k_label_gibbs(LabelGibbsArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* W = sm;                          // [K][H]
    float* Wt = W + 32 * a.H;               // [H][32]   (columns >= K are zero)
    float* hbuf = Wt + a.H * 32;            // [LG_WARPS][LG_CHAINS][H]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * a.H; i += blockDim.x) {
        const int k = i / a.H, j = i - k * a.H;
        const float w = k < a.K ? a.Wy[(size_t)k * a.H + j] : 0.0f;
        W[i] = w;
        Wt[j * 32 + k] = w;
    }
    __syncthreads();
    const int nj = a.H >> 5;                // hidden units per lane
    constexpr int CH = LG_CHAINS;           // chains advanced together by one warp (amortises the W reads)
    constexpr int NJ = LG_MAXH / 32;
    float* hw = hbuf + warp * CH * a.H;
    const float by = lane < a.K ? a.vby[lane] : 0.0f;
    for (int row0 = (blockIdx.x * LG_WARPS + warp) * CH; row0 < a.B; row0 += gridDim.x * LG_WARPS * CH) {
        // init: the unknown (label) units start uniform in [0,1)              rbm.py:392
        float y[CH], c[CH][NJ];
#pragma unroll
        for (int q = 0; q < CH; ++q) {
            const int row = min(row0 + q, a.B - 1);
            y[q] = lane < a.K ? rf_uniform(a.key, a.draw0, row, a.Dz + lane) : 0.0f;
#pragma unroll
            for (int i = 0; i < NJ; ++i) c[q][i] = i < nj ? a.pre[(size_t)row * a.H + lane + 32 * i] : 0.0f;
        }
        for (int t = 0; t < a.n_steps; ++t) {
            // h_j = sigmoid(c_j + sum_k y_k W[k][j]),  j = lane + 32 i          rbm.py:394
            float acc[CH][NJ];
#pragma unroll
            for (int q = 0; q < CH; ++q)
#pragma unroll
                for (int i = 0; i < NJ; ++i) acc[q][i] = c[q][i];
            for (int k = 0; k < a.K; ++k) {
                float yk[CH];
#pragma unroll
                for (int q = 0; q < CH; ++q) yk[q] = __shfl_sync(0xffffffffu, y[q], k);
                const float* wr = W + k * a.H + lane;
#pragma unroll
                for (int i = 0; i < NJ; ++i) {
                    if (i < nj) {
                        const float w = wr[32 * i];
#pragma unroll
                        for (int q = 0; q < CH; ++q) acc[q][i] = fmaf(yk[q], w, acc[q][i]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < CH; ++q)
#pragma unroll
                for (int i = 0; i < NJ; ++i)
                    if (i < nj)
                        hw[q * a.H + lane + 32 * i] =
                            FAST ? __fdividef(1.0f, 1.0f + __expf(-acc[q][i])) : sigmoidf_ref(acc[q][i]);
            __syncwarp();
            // logits_k = sum_j h_j W[k][j] + b_k  (lane = k), softmax over the warp   rbm.py:396, 113-114
            float lg[CH];
#pragma unroll
            for (int q = 0; q < CH; ++q) lg[q] = by;
            for (int j = 0; j < a.H; j += 4) {
                const float w0 = Wt[(j + 0) * 32 + lane], w1 = Wt[(j + 1) * 32 + lane];
                const float w2 = Wt[(j + 2) * 32 + lane], w3 = Wt[(j + 3) * 32 + lane];
#pragma unroll
                for (int q = 0; q < CH; ++q) {
                    const float4 h4 = *reinterpret_cast<const float4*>(hw + q * a.H + j);
                    lg[q] = fmaf(h4.x, w0, lg[q]);
                    lg[q] = fmaf(h4.y, w1, lg[q]);
                    lg[q] = fmaf(h4.z, w2, lg[q]);
                    lg[q] = fmaf(h4.w, w3, lg[q]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < CH; ++q) {
                float mx = lane < a.K ? lg[q] : -INFINITY;
                for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const float e = lane < a.K ? (FAST ? __expf(lg[q] - mx) : expf(lg[q] - mx)) : 0.0f;
                float s = e;
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                y[q] = FAST ? __fdividef(e, s) : div_by(e, s, 1.0f / s);
            }
        }
#pragma unroll
        for (int q = 0; q < CH; ++q)
            if (lane < a.K && row0 + q < a.B) a.y_out[(size_t)(row0 + q) * a.K + lane] = y[q];
    }
}

// v[b, :] = [ z (from v_known) | y ]
__global__ void k_assemble_zy(const float* __restrict__ vk, const float* __restrict__ y, int B, int V, int Dz,
                              float* __restrict__ v) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= V) return;
    for (int b = blockIdx.y; b < B; b += gridDim.y)
        v[(size_t)b * V + c] = c < Dz ? vk[(size_t)b * V + c] : y[(size_t)b * (V - Dz) + (c - Dz)];
}

// x = sum_s part + bias  (pre-activation, no non-linearity)
__global__ void k_preact(const float* __restrict__ part, int splits, SKPlan sk, int B, int H,
                         const float* __restrict__ hb, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= H) return;
    const int ns = sk.k_iters ? sk_nslabs(sk, j / sk.tile_w) : splits;
    const size_t n = (size_t)B * H;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t i = (size_t)b * H + j;
        float x = 0.0f;
        for (int s = 0; s < ns; ++s) x += part[(size_t)s * n + i];
        out[i] = add_rn(x, hb[j]);
    }
}

}  // namespace imdbn
