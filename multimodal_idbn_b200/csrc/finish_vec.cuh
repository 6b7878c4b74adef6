// Vectorised (float4) finishes of the up / down passes and of the stepped chain: the same
// arithmetic per element as the scalar kernels in rbm_kernels.cuh / chain_stepped.cuh, four columns per
// thread, flattened (row, column-quad) indexing so that narrow layers (V = 532) waste no lanes.
// FAST = tf32 mode: reciprocal-multiply temperature, __expf / __fdividef sigmoid (1e-6 relative), which
// is well inside the tf32 tolerance; the fp32 parity mode keeps IEEE division and expf.
#pragma once
#include "common.cuh"
#include "rbm_kernels.cuh"
#include "chain_stepped.cuh"

namespace imdbn {

template <bool FAST>
__device__ __forceinline__ float sigmoid_t(float x) {
    if (FAST) return __fdividef(1.0f, 1.0f + __expf(-x));
    return sigmoidf_ref(x);
}
template <bool FAST>
__device__ __forceinline__ float div_t(float x, float T, float invT) { return FAST ? x * invT : div_by(x, T, invT); }

// The loads of up to 16 slabs are issued together (one L2 round trip instead of one per group of four -- the finish
// sits on the critical path between two passes); the additions stay in slab order.
// (W = loads in flight: 16 for the up pass, whose few output tiles are split over many CTAs; 4 for the down pass, whose
// tiles have at most three slabs and whose finish carries the chain epilogue -- 16 there meant register spills.)
template <int W = 16>
__device__ __forceinline__ float4 sum_slabs4(const float* __restrict__ part, int ns, size_t stride, size_t i) {
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s0 = 0; s0 < ns; s0 += W) {
        float4 v[W];
#pragma unroll
        for (int u = 0; u < W; ++u)
            v[u] = (s0 + u < ns) ? __ldcg(reinterpret_cast<const float4*>(part + (size_t)(s0 + u) * stride + i))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < W; ++u)
            if (s0 + u < ns) { x.x += v[u].x; x.y += v[u].y; x.z += v[u].z; x.w += v[u].w; }
    }
    return x;
}

// ---- up pass finish (rbm.py:92,175,203) -- optional Gaussian logit noise for the chain step (:344-347)
template <bool FAST>
__global__ void __launch_bounds__(256)
k_finish_up4(const float* __restrict__ part, int splits, SKPlan sk, int B, int H, const float* __restrict__ hb,
             float T, float sigma, float* __restrict__ p_out, float* __restrict__ s_out, RngKey key,
             uint32_t draw_u, uint32_t draw_n) {
    pdl_trigger();
    pdl_wait();
    const unsigned q_per_row = (unsigned)H >> 2;
    const unsigned total = (unsigned)B * q_per_row;          // < 2^32 (checked by the host)
    const size_t n = (size_t)B * H;
    const float invT = 1.0f / T;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned b = idx / q_per_row, j = (idx - b * q_per_row) << 2;
        const size_t i = (size_t)b * H + j;
        const int ns = finish_nslabs(sk, splits, j);
        const float4 a4 = sum_slabs4(part, ns, n, i);
        const float4 b4 = *reinterpret_cast<const float4*>(hb + j);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
        float p[4], s[4], nz[4] = {0.f, 0.f, 0.f, 0.f}, un[4] = {0.f, 0.f, 0.f, 0.f};
        if (sigma > 0.0f) {                      // two Philox calls for the four normals of this quad
            const float2 n0 = rf_normal2_t<FAST>(key, draw_n, b, j), n1 = rf_normal2_t<FAST>(key, draw_n, b, j + 2);
            nz[0] = n0.x; nz[1] = n0.y; nz[2] = n1.x; nz[3] = n1.y;
        }
        if (s_out) {                             // one Philox call for the four uniforms
            const float4 u4 = rf_uniform4(key, draw_u, b, j);
            un[0] = u4.x; un[1] = u4.y; un[2] = u4.z; un[3] = u4.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x = div_t<FAST>(add_rn(a[e], bb[e]), T, invT);
            if (sigma > 0.0f) x = add_rn(x, mul_rn(nz[e], sigma));
            p[e] = sigmoid_t<FAST>(x);
            if (s_out) s[e] = (p[e] > un[e]) ? 1.0f : 0.0f;
        }
        if (p_out) *reinterpret_cast<float4*>(p_out + i) = make_float4(p[0], p[1], p[2], p[3]);
        if (s_out) *reinterpret_cast<float4*>(s_out + i) = make_float4(s[0], s[1], s[2], s[3]);
    }
}

// ---- down pass finish (rbm.py:96,110,125) and, with `po.enabled`, the chain-step epilogue (:350-365):
// group columns keep their (noisy) logits for the softmax kernel, the others get sigmoid, mu-pull and
// re-clamp.  Softmax-group boundaries must be multiples of 4 columns for this kernel (checked by the host).
struct ChainPost4 { int enabled; ChainPost po; };

template <bool FAST>
__global__ void __launch_bounds__(256)
k_finish_down4(const float* __restrict__ part, int splits, SKPlan sk, int B, int V, const float* __restrict__ vb,
               float T, float sigma, float* __restrict__ p_out, float* __restrict__ logits_out,
               float* __restrict__ s_out, RngKey key, uint32_t draw_u, uint32_t draw_n, ChainPost4 cp) {
    pdl_trigger();
    pdl_wait();
    const unsigned q_per_row = (unsigned)V >> 2;
    const unsigned total = (unsigned)B * q_per_row;          // < 2^32 (checked by the host)
    const size_t n = (size_t)B * V;
    const float invT = 1.0f / T;
    const ChainPost& po = cp.po;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned b = idx / q_per_row, c = (idx - b * q_per_row) << 2;
        const size_t i = (size_t)b * V + c;
        // block mask promised by the caller: clamped columns only copy the known value (their logits, noise and
        // softmax never reach an output), free columns need neither the mask nor the known values
        const bool block_mask = cp.enabled && po.clamp_from >= 0 && !po.free_sweep && po.vprob_out == nullptr;
        if (block_mask && (int)c >= po.clamp_from) {
            *reinterpret_cast<float4*>(p_out + i) = *reinterpret_cast<const float4*>(po.vk + i);
            continue;
        }
        const int ns = finish_nslabs(sk, splits, c);
        const float4 a4 = sum_slabs4<4>(part, ns, n, i);
        const float4 b4 = *reinterpret_cast<const float4*>(vb + c);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
        float x[4], p[4], s[4], nz[4] = {0.f, 0.f, 0.f, 0.f};
        if (sigma > 0.0f) {
            const float2 n0 = rf_normal2_t<FAST>(key, draw_n, b, c), n1 = rf_normal2_t<FAST>(key, draw_n, b, c + 2);
            nz[0] = n0.x; nz[1] = n0.y; nz[2] = n1.x; nz[3] = n1.y;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            x[e] = div_t<FAST>(add_rn(a[e], bb[e]), T, invT);
            if (sigma > 0.0f) x[e] = add_rn(x[e], mul_rn(nz[e], sigma));
        }
        if (!cp.enabled) {
            float un[4] = {0.f, 0.f, 0.f, 0.f};
            if (s_out) {
                const float4 u4 = rf_uniform4(key, draw_u, b, c);
                un[0] = u4.x; un[1] = u4.y; un[2] = u4.z; un[3] = u4.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                p[e] = sigmoid_t<FAST>(x[e]);
                if (s_out) s[e] = (p[e] > un[e]) ? 1.0f : 0.0f;
            }
            if (logits_out) *reinterpret_cast<float4*>(logits_out + i) = make_float4(x[0], x[1], x[2], x[3]);
            if (p_out) *reinterpret_cast<float4*>(p_out + i) = make_float4(p[0], p[1], p[2], p[3]);
            if (s_out) *reinterpret_cast<float4*>(s_out + i) = make_float4(s[0], s[1], s[2], s[3]);
            continue;
        }
        bool in_group = false;
        for (int g = 0; g < po.gr.n; ++g) in_group |= ((int)c >= po.gr.s[g] && (int)c < po.gr.e[g]);
        if (in_group) {
            *reinterpret_cast<float4*>(logits_out + i) = make_float4(x[0], x[1], x[2], x[3]);
            continue;
        }
        float4 k4 = make_float4(0, 0, 0, 0), m4 = make_float4(0, 0, 0, 0);     // (free column: v = p)
        if (!po.free_sweep && !block_mask) {
            k4 = *reinterpret_cast<const float4*>(po.vk + i);
            m4 = *reinterpret_cast<const float4*>(po.km + i);
        }
        const float kk[4] = {k4.x, k4.y, k4.z, k4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            p[e] = sigmoid_t<FAST>(x[e]);
            if (po.mu && (int)(c + e) < po.Dz)
                p[e] = add_rn(mul_rn(1.0f - po.eta, p[e]), mul_rn(po.eta, po.mu[(size_t)b * po.Dz + c + e]));
            v[e] = (po.free_sweep || block_mask) ? p[e] : clampmix(p[e], kk[e], mm[e]);
        }
        if (po.vprob_out) *reinterpret_cast<float4*>(po.vprob_out + i) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4*>(p_out + i) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// one resident wave: 256-thread blocks of these kernels fit four to an SM (registers); a grid just above that (625
// blocks for the 64 x 10000 finish on 132 SMs) pays a second, nearly empty wave on the critical path of the step
inline int vec_blocks(size_t quads, int num_sms) {
    const size_t b = (quads + 255) / 256;
    return (int)std::max<size_t>(1, std::min<size_t>(b, (size_t)num_sms * 4));
}

}  // namespace imdbn
