// Data-parallel update over peer memory (NVLink 5 / NVSwitch): reduce-scatter of the CD statistics, update of
// the owned slab of W / W_m, all-gather of the new weights -- ONE kernel, no NCCL call and no staging buffer.
//
// Every rank r has written its local statistics S_r = [dS | dh | dv | sum pos_h | sq] into a peer-mapped
// buffer.  Rank r owns the float4 range [q0, q1) of the flattened [V,H] weight matrix: for each owned quad it
// loads the quad of dS from all ranks (P2P loads, issued together), adds them in rank order (so the sum does
// not depend on which rank owns the quad), applies rbm.py:212-213 to its W / W_m and stores the new W quad
// into EVERY rank's copy of W (P2P stores).  Per rank and update the link carries (world-1)/world * 4VH bytes
// in and out, i.e. exactly a reduce-scatter plus an all-gather, overlapped element-wise with the update; the
// momentum matrix is touched only on the owner, so the update's HBM traffic drops by the world size.
// The caller brackets the kernel with cross-rank barriers (statistics complete before, weights complete after).
#pragma once
#include "common.cuh"

namespace imdbn {

struct PeerPtrs {
    const float* stats[IMDBN_MAX_PEERS];
    float* W[IMDBN_MAX_PEERS];
    const float* stats_mc;      // multicast addresses (NVLS) or nullptr
    float* W_mc;
    int world;
};

// NVSwitch in-fabric reduction: one load returns the fp32 sum of the quad over all ranks' buffers
__device__ __forceinline__ float4 multimem_ld_reduce4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
// NVSwitch broadcast: one store lands in every rank's buffer
__device__ __forceinline__ void multimem_st4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 ld_peer4(const float* p) {      // remote data: read once, never cached stale
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <bool MC>
__global__ void __launch_bounds__(256) k_dp_update(PeerPtrs p, int rank, size_t q0, size_t q1, float* __restrict__ Wm,
                                                   float lr, float mom, float wd, float bsz) {
    const float rb = 1.0f / bsz;
    float* Wl = p.W[rank];
    for (size_t q = q0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < q1; q += (size_t)gridDim.x * blockDim.x) {
        float4 t;
        float4 s[IMDBN_MAX_PEERS];
        if (MC) {
            t = multimem_ld_reduce4(p.stats_mc + 4 * q);
        } else {
#pragma unroll
            for (int r = 0; r < IMDBN_MAX_PEERS; ++r)
                if (r < p.world) s[r] = ld_peer4(p.stats[r] + 4 * q);
        }
        const float4 w4 = __ldcg(reinterpret_cast<const float4*>(Wl) + q);
        const float4 m4 = __ldcg(reinterpret_cast<const float4*>(Wm) + q);
        if (!MC) {
            t = s[0];
#pragma unroll
            for (int r = 1; r < IMDBN_MAX_PEERS; ++r)
                if (r < p.world) { t.x += s[r].x; t.y += s[r].y; t.z += s[r].z; t.w += s[r].w; }
        }
        const float ds[4] = {t.x, t.y, t.z, t.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w};
        float nw[4], nm[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float grad = add_rn(div_by(ds[e], bsz, rb), -mul_rn(wd, wv[e]));
            nm[e] = add_rn(mul_rn(mv[e], mom), mul_rn(lr, grad));
            nw[e] = add_rn(wv[e], nm[e]);
        }
        reinterpret_cast<float4*>(Wm)[q] = make_float4(nm[0], nm[1], nm[2], nm[3]);
        const float4 out = make_float4(nw[0], nw[1], nw[2], nw[3]);
        if (MC) {
            multimem_st4(p.W_mc + 4 * q, out);
        } else {
#pragma unroll
            for (int r = 0; r < IMDBN_MAX_PEERS; ++r)
                if (r < p.world) reinterpret_cast<float4*>(p.W[r])[q] = out;
        }
    }
}

// the small tail of the statistics ([dh | dv | sum pos_h | sq], 2H+V+1 floats) summed over ranks in rank order
__global__ void k_dp_small(PeerPtrs p, size_t off, int n, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float s = 0.f;
    for (int r = 0; r < p.world; ++r) {
        float v;
        asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p.stats[r] + off + c));
        s += v;
    }
    out[c] = s;
}

}  // namespace imdbn
