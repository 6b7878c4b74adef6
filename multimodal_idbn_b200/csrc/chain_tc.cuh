// TXT->IMG noisy mean-field annealing (rbm.py:300-367 as iMDBN._cross_reconstruct calls it, imdbn.py:430-449: the label
// block clamped, the image latents free) for MANY chains as ONE persistent tcgen05 kernel (included by tc_gemm.cu).
//
//   per step t:  h = sigmoid((v W + b_h) / T_t + sigma_t N_h)
//                z = sigmoid((h W_z^T + b_z) / T_t + sigma_t N_z);  z <- (1 - eta_t) z + eta_t mu      (mu-pull)
//                v = [z | y]   (y never changes: the label columns of the state are written once)
//
// A CTA owns a tile of CT_NC = 48 chains for all n steps.  The chain state IS the MMA operand: v [48 x 544] and
// h [48 x 256] live in shared memory in the K-major 128-byte-swizzled layout the tensor core reads (swap-AB as in
// k_tc_stream: M = 128 output features, N = 48 chains), the accumulators in TMEM.  Only W moves: the TMA producer
// streams its tiles (L2-resident, 545 KB) through a 4-stage ring and runs ahead across steps, since the weight
// tiles do not depend on the state.  Twelve epilogue warps (three per TMEM lane quadrant, 16 chains each) turn an
// accumulator tile into the next operand: bias, temperature, Philox Gaussian noise (one call per pair of chains per
// lane, exchanged with the neighbouring lane, so no call is wasted), sigmoid, mu-pull -- written straight back into
// the swizzled operand; nothing but the final state goes to global memory.
// Single-pass tf32 with the fast intrinsics of the tf32 finishes (north_star's <= 1e-3 mode); the exact mode keeps
// the stepped path.
#pragma once

constexpr int CT_NC = 48;                      // chains per CTA tile (MMA N)
constexpr int CT_BK = 32;                      // k elements per W stage
constexpr int CT_STAGES = 4;
constexpr int CT_A_BYTES = 128 * CT_BK * 4;    // one W tile: 16 KB
constexpr int CT_EPI_WARPS = 12;
constexpr int CT_THREADS = 64 + 32 * CT_EPI_WARPS;     // warp 0 TMA, warp 1 MMA, warps 2-13 epilogue
constexpr int CT_BOX = CT_NC * 128;            // one [48 chains x 32 k] operand box

struct ChainTcArgs {
    int V, H, Dz, B, n_steps;
    int kb_v, kb_h;                            // 32-wide k boxes of the v / h operands
    int mt_h, mt_v;                            // 128-row output tiles of the up (H) / down (Dz) products
    const float* hb; const float* vb;
    const float* v_known;                      // [B,V]: the clamped label block is read from it
    const float* mu;                           // nullable [B,Dz]
    const float* T; const float* sigma; const float* eta;     // device tables [n_steps]
    float* v_out;                              // [B,V]
    RngKey key; uint32_t draw0;
};

__host__ __device__ inline size_t ct_smem_bytes(int kb_v, int kb_h) {
    return (size_t)CT_STAGES * CT_A_BYTES + (size_t)(kb_v + kb_h) * CT_BOX + 256 + 1024;
}

// byte offset of element (chain c, k) inside a K-major operand: 32-wide k boxes of [48 rows x 128 B], 128-byte swizzle
__device__ __forceinline__ uint32_t ct_off(int c, int k) {
    return (uint32_t)(k >> 5) * CT_BOX + (uint32_t)c * 128u + ((((uint32_t)(k & 31) >> 2) ^ (uint32_t)(c & 7)) << 4) +
           ((uint32_t)(k & 3) << 2);
}

__global__ void __launch_bounds__(CT_THREADS, 1)
k_chain_tc(const __grid_constant__ CUtensorMap tmUp, const __grid_constant__ CUtensorMap tmDn, ChainTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint8_t* sV = ring + CT_STAGES * CT_A_BYTES;
    uint8_t* sH = sV + a.kb_v * CT_BOX;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sH + a.kb_h * CT_BOX);
    uint64_t* full = bars;                // [4]
    uint64_t* empty = bars + 4;           // [4]
    uint64_t* acc_full = bars + 8;        // [6]: 2 up tiles, 4 down tiles
    uint64_t* v_ready = bars + 14;
    uint64_t* h_ready = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (a.B + CT_NC - 1) / CT_NC;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmUp); tma_prefetch_desc(&tmDn);
        for (int s = 0; s < CT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 6; ++s) mbar_init(&acc_full[s], 1);
        mbar_init(v_ready, CT_EPI_WARPS); mbar_init(h_ready, CT_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== W producer: runs ahead of the chain, tile after tile, step after step =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int t = 0; t < a.n_steps; ++t) {
                    for (int mt = 0; mt < a.mt_h; ++mt)
                        for (int kb = 0; kb < a.kb_v; ++kb) {       // up: W[32 k rows, 128 hidden units], MN-major
                            uint8_t* sA = ring + stage * CT_A_BYTES;
                            mbar_wait(&empty[stage], phase ^ 1);
                            mbar_expect_tx(&full[stage], CT_A_BYTES);
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb)
                                tma_load_2d(sA + cb * (CT_BK * 128), &tmUp, mt * 128 + cb * 32, kb * CT_BK, &full[stage]);
                            if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                        }
                    for (int mt = 0; mt < a.mt_v; ++mt)
                        for (int kb = 0; kb < a.kb_h; ++kb) {       // down: W[128 visible units, 32 hidden], K-major
                            uint8_t* sA = ring + stage * CT_A_BYTES;
                            mbar_wait(&empty[stage], phase ^ 1);
                            mbar_expect_tx(&full[stage], CT_A_BYTES);
                            tma_load_2d(sA, &tmDn, kb * CT_BK, mt * 128, &full[stage]);
                            if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                        }
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t id_up = idesc_tf32(128, CT_NC, true, false, false);
            const uint32_t id_dn = idesc_tf32(128, CT_NC, false, false, false);
            const uint32_t sVu = smem_u32(sV), sHu = smem_u32(sH);
            int stage = 0; uint32_t phase = 0, v_cnt = 0, h_cnt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int t = 0; t < a.n_steps; ++t) {
                    mbar_wait(v_ready, v_cnt & 1); ++v_cnt;          // v operand written (init or the last down epilogue)
                    tc_fence_after();
                    for (int mt = 0; mt < a.mt_h; ++mt) {
                        const uint32_t d = tmem_base + (uint32_t)(mt * CT_NC);
                        for (int kb = 0; kb < a.kb_v; ++kb) {
                            mbar_wait(&full[stage], phase);
                            tc_fence_after();
                            const uint32_t sA = smem_u32(ring + stage * CT_A_BYTES);
#pragma unroll
                            for (int g = 0; g < CT_BK / 8; ++g)
                                mma_tf32(d, smem_desc(sA + g * 1024, CT_BK * 128, 512, LAYOUT_SW128_BASE32B),
                                         smem_desc(sVu + kb * CT_BOX + g * 32, 16, 1024, LAYOUT_SW128), id_up, (kb | g) != 0);
                            mma_commit(&empty[stage]);
                            if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                        }
                        mma_commit(&acc_full[mt]);
                    }
                    mbar_wait(h_ready, h_cnt & 1); ++h_cnt;
                    tc_fence_after();
                    for (int mt = 0; mt < a.mt_v; ++mt) {
                        const uint32_t d = tmem_base + 128u + (uint32_t)(mt * CT_NC);
                        for (int kb = 0; kb < a.kb_h; ++kb) {
                            mbar_wait(&full[stage], phase);
                            tc_fence_after();
                            const uint32_t sA = smem_u32(ring + stage * CT_A_BYTES);
#pragma unroll
                            for (int g = 0; g < CT_BK / 8; ++g)
                                mma_tf32(d, smem_desc(sA + g * 32, 16, 1024, LAYOUT_SW128),
                                         smem_desc(sHu + kb * CT_BOX + g * 32, 16, 1024, LAYOUT_SW128), id_dn, (kb | g) != 0);
                            mma_commit(&empty[stage]);
                            if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                        }
                        mma_commit(&acc_full[2 + mt]);
                    }
                }
        }
    } else {
        // ===================== epilogue: accumulator tile -> next operand =====================
        const int quad = warp & 3;                       // TMEM lane quadrant of this warp
        const int grp = (warp - 2) >> 2;                 // chains grp * 16 .. + 16
        const int et = threadIdx.x - 64;                 // 0 .. 383
        const int Kpad = a.kb_v * 32;
        uint32_t step_cnt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int row0 = tile * CT_NC;
            // ---- initial state: free units uniform (rbm.py:333), label block from v_known; padding zero
            named_bar_sync(1, 32 * CT_EPI_WARPS);        // the previous tile's last writes are done
            for (int idx = et; idx < CT_NC * Kpad; idx += 32 * CT_EPI_WARPS) {
                const int c = idx / Kpad, k = idx - c * Kpad;
                const int row = row0 + c;
                float val = 0.0f;
                if (row < a.B && k < a.V) {
                    if (k < a.Dz) val = rf_uniform(a.key, a.draw0, row, k);
                    else { val = a.v_known[(size_t)row * a.V + k]; a.v_out[(size_t)row * a.V + k] = val; }
                }
                *reinterpret_cast<float*>(sV + ct_off(c, k)) = val;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(v_ready);
            for (int t = 0; t < a.n_steps; ++t, ++step_cnt) {
                const float invT = 1.0f / a.T[t], sig = a.sigma[t], eta = a.mu ? a.eta[t] : 0.0f;
                const uint32_t d_h = a.draw0 + 1 + 2 * t, d_v = d_h + 1;
                const bool last = (t == a.n_steps - 1);
                // ---- h = sigmoid((vW + b_h)/T + sigma N)                                   rbm.py:344-347
                for (int mt = 0; mt < a.mt_h; ++mt) {
                    mbar_wait(&acc_full[mt], step_cnt & 1);
                    tc_fence_after();
                    float acc[16];
                    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * CT_NC + grp * 16), acc);
                    const int j = mt * 128 + quad * 32 + lane;
                    const float bj = j < a.H ? a.hb[j] : 0.0f;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const int c0 = grp * 16 + i;
                        float n0 = 0.0f, n1 = 0.0f;
                        if (sig > 0.0f) {       // one Philox call serves two columns: lane pairs share two chains
                            const int myrow = row0 + c0 + (lane & 1);
                            const float2 nn = rf_normal2_t<true>(a.key, d_h, myrow, j & ~1);
                            const float mine = (lane & 1) ? nn.y : nn.x, send = (lane & 1) ? nn.x : nn.y;
                            const float other = __shfl_xor_sync(0xffffffffu, send, 1);
                            n0 = (lane & 1) ? other : mine;
                            n1 = (lane & 1) ? mine : other;
                        }
                        const float x0 = fmaf(n0, sig, (acc[i] + bj) * invT), x1 = fmaf(n1, sig, (acc[i + 1] + bj) * invT);
                        if (j < a.H) {
                            *reinterpret_cast<float*>(sH + ct_off(c0, j)) = __fdividef(1.0f, 1.0f + __expf(-x0));
                            *reinterpret_cast<float*>(sH + ct_off(c0 + 1, j)) = __fdividef(1.0f, 1.0f + __expf(-x1));
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(h_ready);
                // ---- z = sigmoid((hW^T + b_v)/T + sigma N), mu-pull; the label block stays            rbm.py:350-365
                for (int mt = 0; mt < a.mt_v; ++mt) {
                    mbar_wait(&acc_full[2 + mt], step_cnt & 1);
                    tc_fence_after();
                    float acc[16];
                    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + 128u + (uint32_t)(mt * CT_NC + grp * 16), acc);
                    const int iv = mt * 128 + quad * 32 + lane;
                    const bool free_unit = iv < a.Dz;
                    const float bi = free_unit ? a.vb[iv] : 0.0f;
                    float mu_r[16];                      // all 16 mu-pull targets in flight before the arithmetic starts
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int r = row0 + grp * 16 + i;
                        mu_r[i] = (a.mu && free_unit && r < a.B) ? __ldg(a.mu + (size_t)r * a.Dz + iv) : 0.0f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const int c0 = grp * 16 + i;
                        const int r0 = row0 + c0, r1 = r0 + 1;
                        float n0 = 0.0f, n1 = 0.0f;
                        if (sig > 0.0f) {
                            const float2 nn = rf_normal2_t<true>(a.key, d_v, (lane & 1) ? r1 : r0, iv & ~1);
                            const float mine = (lane & 1) ? nn.y : nn.x, send = (lane & 1) ? nn.x : nn.y;
                            const float other = __shfl_xor_sync(0xffffffffu, send, 1);
                            n0 = (lane & 1) ? other : mine;
                            n1 = (lane & 1) ? mine : other;
                        }
                        if (free_unit) {
                            float p0 = __fdividef(1.0f, 1.0f + __expf(-fmaf(n0, sig, (acc[i] + bi) * invT)));
                            float p1 = __fdividef(1.0f, 1.0f + __expf(-fmaf(n1, sig, (acc[i + 1] + bi) * invT)));
                            if (a.mu) {
                                p0 = (1.0f - eta) * p0 + eta * mu_r[i];
                                p1 = (1.0f - eta) * p1 + eta * mu_r[i + 1];
                            }
                            *reinterpret_cast<float*>(sV + ct_off(c0, iv)) = p0;
                            *reinterpret_cast<float*>(sV + ct_off(c0 + 1, iv)) = p1;
                            if (last) {
                                if (r0 < a.B) a.v_out[(size_t)r0 * a.V + iv] = p0;
                                if (r1 < a.B) a.v_out[(size_t)r1 * a.V + iv] = p1;
                            }
                        }
                    }
                }
                tc_fence_before();
                if (!last) {                             // (after the last step nobody reads v: the next tile re-initialises it)
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(v_ready);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}
