// Device statement of the counter-based random field (normative host version: oracle/philox.py).
//   uniform(col) = (x[col & 3] >> 8) * 2^-24,  x = Philox4x32-10(key=(seed_lo,seed_hi), counter=(col >> 2, global_row, draw, stream))
//   normal(col)  = sqrt(-2 ln((a>>8)+1)*2^-24) * cos(2 pi (b>>8)*2^-24),  (a, b) = words 2(col & 1), 2(col & 1)+1 of
//                  Philox4x32-10(..., counter=(col >> 1, global_row, draw, stream))
// Replaces torch.rand_like / randn_like / Categorical.sample of imdbn/models/rbm.py:125,131,203,
// 208,333,346,352,392,395,462.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace imdbn {

struct RngKey {
    uint32_t k0, k1;   // seed lo / hi
    uint32_t stream;   // API-call number
    uint32_t row0;     // global index of local row 0
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// One Philox call serves FOUR consecutive columns of a uniform draw (column j = word j & 3 of counter j >> 2) and
// TWO consecutive columns of a normal draw (column j = words 2(j & 1), 2(j & 1) + 1 of counter j >> 1): the
// vectorised kernels, which own four consecutive columns per thread, pay one call per float4 of uniforms and two
// per float4 of normals.
__device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }   // 2^-24, exact

__device__ __forceinline__ float4 rf_uniform4(const RngKey& k, uint32_t draw, uint32_t row, uint32_t col4) {
    const uint4 x = philox4x32_10(col4 >> 2, row + k.row0, draw, k.stream, k.k0, k.k1);       // col4 % 4 == 0
    return make_float4(u24(x.x), u24(x.y), u24(x.z), u24(x.w));
}

__device__ __forceinline__ float rf_uniform(const RngKey& k, uint32_t draw, uint32_t row, uint32_t col) {
    const uint4 x = philox4x32_10(col >> 2, row + k.row0, draw, k.stream, k.k0, k.k1);
    const uint32_t w = (col & 2) ? ((col & 1) ? x.w : x.z) : ((col & 1) ? x.y : x.x);
    return u24(w);
}

template <bool FAST>
__device__ __forceinline__ float box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)(a >> 8) + 1.0f) * 5.9604644775390625e-8f;   // (0, 1]
    const float u2 = u24(b);                                             // [0, 1)
    // FAST (tf32 mode): hardware lg2 / cos approximations (absolute error ~1e-6, far below the tf32 operand
    // rounding of the logits the noise is added to)
    return FAST ? sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831855f * u2) : sqrtf(-2.0f * logf(u1)) * cosf(6.2831855f * u2);
}

template <bool FAST>
__device__ __forceinline__ float2 rf_normal2_t(const RngKey& k, uint32_t draw, uint32_t row, uint32_t col2) {
    const uint4 x = philox4x32_10(col2 >> 1, row + k.row0, draw, k.stream, k.k0, k.k1);       // col2 % 2 == 0
    return make_float2(box_muller<FAST>(x.x, x.y), box_muller<FAST>(x.z, x.w));
}

template <bool FAST>
__device__ __forceinline__ float rf_normal_t(const RngKey& k, uint32_t draw, uint32_t row, uint32_t col) {
    const uint4 x = philox4x32_10(col >> 1, row + k.row0, draw, k.stream, k.k0, k.k1);
    return (col & 1) ? box_muller<FAST>(x.z, x.w) : box_muller<FAST>(x.x, x.y);
}

__device__ __forceinline__ float rf_normal(const RngKey& k, uint32_t draw, uint32_t row, uint32_t col) {
    return rf_normal_t<false>(k, draw, row, col);
}
__device__ __forceinline__ float rf_normal_fast(const RngKey& k, uint32_t draw, uint32_t row, uint32_t col) {
    return rf_normal_t<true>(k, draw, row, col);
}

}  // namespace imdbn
