// Device statement of the counter-based random field (normative host version: oracle/philox.py).
//   x = Philox4x32-10(key=(seed_lo,seed_hi), counter=(col, global_row, draw, stream))
//   uniform = (x0 >> 8) * 2^-24 ; normal = sqrt(-2 ln((x0>>8)+1)*2^-24) * cos(2 pi (x1>>8)*2^-24)
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

__device__ __forceinline__ float rf_uniform(const RngKey& k, uint32_t draw, uint32_t row, uint32_t col) {
    const uint4 x = philox4x32_10(col, row + k.row0, draw, k.stream, k.k0, k.k1);
    return (float)(x.x >> 8) * 5.9604644775390625e-8f;  // 2^-24, exact
}

__device__ __forceinline__ float rf_normal(const RngKey& k, uint32_t draw, uint32_t row, uint32_t col) {
    const uint4 x = philox4x32_10(col, row + k.row0, draw, k.stream, k.k0, k.k1);
    const float u1 = ((float)(x.x >> 8) + 1.0f) * 5.9604644775390625e-8f;  // (0, 1]
    const float u2 = (float)(x.y >> 8) * 5.9604644775390625e-8f;           // [0, 1)
    return sqrtf(-2.0f * logf(u1)) * cosf(6.2831855f * u2);
}

// tf32-mode variant: hardware lg2 / cos approximations (absolute error ~1e-6, far below the tf32
// operand rounding of the logits the noise is added to)
__device__ __forceinline__ float rf_normal_fast(const RngKey& k, uint32_t draw, uint32_t row, uint32_t col) {
    const uint4 x = philox4x32_10(col, row + k.row0, draw, k.stream, k.k0, k.k1);
    const float u1 = ((float)(x.x >> 8) + 1.0f) * 5.9604644775390625e-8f;
    const float u2 = (float)(x.y >> 8) * 5.9604644775390625e-8f;
    return sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831855f * u2);
}

}  // namespace imdbn
