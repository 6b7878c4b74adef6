// Batched IMG->TXT diagnostics of utils/energy_utils.py on the label-only structure of the joint RBM
// (the latents z are clamped, so the hidden pre-activation splits into a constant c = z W_z + b_h and the
// label block W_y):
//   * class free energies (energy_utils.py:32-54):  F_k(z) = -(z.b_z + b_y[k]) - sum_j softplus(c_j + W_y[k][j])
//   * the deterministic "mean-field lite" trace (energy_utils.py:61-90, 132-160): per step
//       h = sigmoid(c + y W_y) ;  s = sigmoid(h W_y^T + b_y) ;  y = softmax(s)      (softmax OF the sigmoid outputs)
//     with the label distribution of every step recorded, so the host derives all trace metrics from one copy.
// One warp per sample, lane <-> label (K <= 32), W_y and its transpose in shared memory (as k_label_gibbs).
#pragma once
#include "common.cuh"
#include "label_gibbs.cuh"

namespace imdbn {

__device__ __forceinline__ float softplus_ref(float x) { return x > 20.0f ? x : log1pf(expf(x)); }   // torch, threshold 20

// F[b][k]; pre = c (from k_preact), zbz[b] is reduced here from z and b_z
__global__ void __launch_bounds__(256) k_class_free_energies(const float* __restrict__ pre, const float* __restrict__ z,
                                                             int ldz, const float* __restrict__ Wy,
                                                             const float* __restrict__ vb, int B, int H, int Dz, int K,
                                                             float* __restrict__ F) {
    extern __shared__ float cf_sm[];          // [H] pre-activations of this sample
    __shared__ float red[8];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float zb = 0.f;
    for (int i = threadIdx.x; i < Dz; i += blockDim.x) zb = fmaf(z[(size_t)b * ldz + i], vb[i], zb);
    for (int j = threadIdx.x; j < H; j += blockDim.x) cf_sm[j] = pre[(size_t)b * H + j];
    for (int o = 16; o; o >>= 1) zb += __shfl_xor_sync(0xffffffffu, zb, o);
    if (lane == 0) red[warp] = zb;
    __syncthreads();
    float zbz = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) zbz += red[w];
    for (int k = warp; k < K; k += blockDim.x >> 5) {
        float s = 0.f;
        for (int j = lane; j < H; j += 32) s += softplus_ref(cf_sm[j] + Wy[(size_t)k * H + j]);
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) F[(size_t)b * K + k] = -(zbz + vb[Dz + k]) - s;
    }
}

struct TraceArgs {
    const float* pre;      // [B,H]  c = z W_z + b_h
    const float* Wy;       // [K,H]
    const float* vby;      // [K]
    const float* y_init;   // nullable [B,K]; default = uniform 1/K (energy_utils.py:134)
    int B, H, K, steps;
    float* y_traj;         // [steps, B, K]
};

__global__ void __launch_bounds__(LG_WARPS * 32) k_trace_img2txt(TraceArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* W = sm;                          // [32][H]
    float* Wt = W + 32 * a.H;               // [H][32]
    float* hbuf = Wt + a.H * 32;            // [LG_WARPS][H]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * a.H; i += blockDim.x) {
        const int k = i / a.H, j = i - k * a.H;
        const float w = k < a.K ? a.Wy[(size_t)k * a.H + j] : 0.0f;
        W[i] = w;
        Wt[j * 32 + k] = w;
    }
    __syncthreads();
    float* hw = hbuf + warp * a.H;
    const float by = lane < a.K ? a.vby[lane] : 0.0f;
    for (int row = blockIdx.x * LG_WARPS + warp; row < a.B; row += gridDim.x * LG_WARPS) {
        float y = lane < a.K ? (a.y_init ? a.y_init[(size_t)row * a.K + lane] : 1.0f / (float)a.K) : 0.0f;
        for (int t = 0; t < a.steps; ++t) {
            for (int j = lane; j < a.H; j += 32) {                   // h = sigmoid(c + y W_y)
                float acc = a.pre[(size_t)row * a.H + j];
                for (int k = 0; k < a.K; ++k) acc = fmaf(__shfl_sync(0xffffffffu, y, k), W[k * a.H + j], acc);
                hw[j] = sigmoidf_ref(acc);
            }
            __syncwarp();
            float lg = by;                                           // s = sigmoid(h W_y^T + b_y)
            for (int j = 0; j < a.H; ++j) lg = fmaf(hw[j], Wt[j * 32 + lane], lg);
            __syncwarp();
            const float s = lane < a.K ? sigmoidf_ref(lg) : -INFINITY;
            float mx = s;                                            // y = softmax(s) over the K labels
            for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float e = lane < a.K ? expf(s - mx) : 0.0f;
            float sum = e;
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            y = e / sum;
            if (lane < a.K) a.y_traj[((size_t)t * a.B + row) * a.K + lane] = y;
        }
    }
}

}  // namespace imdbn
