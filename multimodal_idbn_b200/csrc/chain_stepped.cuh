// Large-batch mean-field chains on the tensor-core passes: when thousands of chains run together the
// two matrix products of a chain step are real GEMMs ([B,V]x[V,H] and [B,H]x[H,V]); they go through
// k_tc_stream (tcgen05) and the element-wise part of the step -- temperature, Gaussian logit noise,
// sigmoid / softmax group, mu-pull, re-clamp (imdbn/models/rbm.py:344-365, 394-397) -- is fused into
// the finish kernels below.  Small batches keep the persistent kernel (chain_kernel.cuh).
#pragma once
#include "common.cuh"
#include "rbm_kernels.cuh"

namespace imdbn {

// v = v_known*km + (1-km)*U   (rbm.py:333,392), or a copy of the given start state
__global__ void k_chain_init(const float* __restrict__ vk, const float* __restrict__ km,
                             const float* __restrict__ v_init, int B, int V, RngKey key, uint32_t draw0,
                             float* __restrict__ v_out) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= V) return;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t i = (size_t)b * V + c;
        if (v_init) { v_out[i] = v_init[i]; continue; }
        const float m = km[i];
        v_out[i] = add_rn(mul_rn(vk[i], m), mul_rn(1.0f - m, rf_uniform(key, draw0, b, c)));
    }
}

// h = sigmoid((sum_s part + hb)/T + sigma*N)                                   rbm.py:344-347
__global__ void k_chain_up_finish(const float* __restrict__ part, int splits, SKPlan sk, int B, int H,
                                  const float* __restrict__ hb, float T, float sigma, RngKey key,
                                  uint32_t draw_n, float* __restrict__ h_out) {
    pdl_trigger();
    pdl_wait();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= H) return;
    const int ns = finish_nslabs(sk, splits, j);
    const size_t n = (size_t)B * H;
    const float bj = hb[j], invT = 1.0f / T;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t i = (size_t)b * H + j;
        float x = div_by(add_rn(sum_slabs(part, ns, n, i), bj), T, invT);
        if (sigma > 0.0f) x = add_rn(x, mul_rn(rf_normal(key, draw_n, b, j), sigma));
        h_out[i] = sigmoidf_ref(x);
    }
}

struct ChainPost {
    const float* vk; const float* km;      // re-clamp (ignored when free_sweep)
    const float* mu; int Dz; float eta;    // mu-pull on columns < Dz (mu == nullptr: off)
    int free_sweep;                        // 1: return the un-clamped probabilities (rbm.py:400)
    float* vprob_out;                      // nullable: un-clamped probabilities of this sweep
    Groups gr;
    int clamp_from;                        // >= 0: known_mask is 1 exactly on columns >= clamp_from (caller's promise)
};

// logits = (sum_s part + vb)/T + sigma*N; non-group columns: sigmoid, mu-pull, re-clamp -> v_out;
// group columns: the noisy logits go to logits_out and k_chain_groups finishes them.   rbm.py:350-365
__global__ void k_chain_down_finish(const float* __restrict__ part, int splits, SKPlan sk, int B, int V,
                                    const float* __restrict__ vb, float T, float sigma, RngKey key,
                                    uint32_t draw_n, ChainPost po, float* __restrict__ logits_out,
                                    float* __restrict__ v_out) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= V) return;
    const int ns = finish_nslabs(sk, splits, c);
    const size_t n = (size_t)B * V;
    const float bc = vb[c], invT = 1.0f / T;
    bool in_group = false;
    for (int g = 0; g < po.gr.n; ++g) in_group |= (c >= po.gr.s[g] && c < po.gr.e[g]);
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t i = (size_t)b * V + c;
        float x = div_by(add_rn(sum_slabs(part, ns, n, i), bc), T, invT);
        if (sigma > 0.0f) x = add_rn(x, mul_rn(rf_normal(key, draw_n, b, c), sigma));
        if (in_group) { logits_out[i] = x; continue; }
        float p = sigmoidf_ref(x);
        if (po.mu && c < po.Dz)
            p = add_rn(mul_rn(1.0f - po.eta, p), mul_rn(po.eta, po.mu[(size_t)b * po.Dz + c]));
        if (po.vprob_out) po.vprob_out[i] = p;
        v_out[i] = po.free_sweep ? p : clampmix(p, po.vk[i], po.km[i]);
    }
}

// one warp per (row, group): softmax of the (noisy) logits, mu-pull if the group lies below Dz, re-clamp
__global__ void k_chain_groups(const float* __restrict__ logits, int B, int V, ChainPost po,
                               float* __restrict__ v_out) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= B * po.gr.n) return;
    const int b = warp / po.gr.n, g = warp % po.gr.n;
    const int s = po.gr.s[g], e = po.gr.e[g];
    const size_t row = (size_t)b * V;
    float mx = -INFINITY;
    for (int c = s + lane; c < e; c += 32) mx = fmaxf(mx, logits[row + c]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int c = s + lane; c < e; c += 32) sum += expf(logits[row + c] - mx);
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float rsum = 1.0f / sum;
    for (int c = s + lane; c < e; c += 32) {
        float p = div_by(expf(logits[row + c] - mx), sum, rsum);
        if (po.mu && c < po.Dz)
            p = add_rn(mul_rn(1.0f - po.eta, p), mul_rn(po.eta, po.mu[(size_t)b * po.Dz + c]));
        if (po.vprob_out) po.vprob_out[row + c] = p;
        v_out[row + c] = po.free_sweep ? p : clampmix(p, po.vk[row + c], po.km[row + c]);
    }
}

}  // namespace imdbn
