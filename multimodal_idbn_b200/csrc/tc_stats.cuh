// CD statistics + update kernel (included by tc_gemm.cu).
//
//   dS[v,h] = sum_b vp[b,v] hp[b,h] - sum_b vn[b,v] hn[b,h]              (rbm.py:200,209)
//   W_m <- mom W_m + lr (dS/bsz - wd W) ;  W <- W + W_m                   (rbm.py:212-213)
//
// Output-streaming and HBM-bound at small batch: per 128 x 128 tile the kernel reads and writes
// 2 x 64 KB of W / W_m and needs only a handful of tf32 MMAs.  Roles:
//   * operand producer (warp 0).  Both operands are MN-major (feature index contiguous in memory, batch index = K),
//     16 batch rows per stage.
//       - narrow variants (PACK; 128 x 128 tiles, batch < 512): the operands were packed by k_pack_ops (below) into
//         tile-ready images, a stage is two bulk copies;
//       - wide variant (128 x 256 tiles, tensor-bound large batches): TMA boxes of [16 batch rows x 32 floats]
//         straight from the activation matrices, 3-stage ring.
//     Exact mode (SPLIT): the stage also carries the tf32 remainders h_lo of the hidden operands and the issuer adds
//     v * h_lo; when some v is not exactly representable (never for binary states) every tile gets a second round
//     of chunks  v_lo * h  (see k_tc_stream for the arithmetic).
//   * MMA issuer (warp 1): tcgen05.mma kind::tf32 into one of two TMEM accumulators; the negative phase is
//     subtracted with the a_negate bit;
//   * IO producer (warp 2): the W and W_m slices (128 rows x 32 columns each) into one of the 32 KB slots -- up to
//     128 KB of weight traffic in flight per SM, independent of the math warps;
//   * epilogue (warps 4-11): thread = tile row; accumulator from TMEM, W / W_m from the swizzled slot
//     (bank-conflict-free with one row per lane), update written back in place;
//   * store issuer (warp 3): TMA-stores the slot back to W / W_m (edges are clipped by the tensor map)
//     and releases the slot when the store has drained shared memory.
#pragma once

constexpr int ST_BM = 128;                 // visible units per tile (MMA M)
constexpr int ST_THREADS = 384;            // 12 warps, roles above
constexpr int ST_EPI_WARPS = 8;
constexpr int ST_KC = 16;                  // batch rows per operand stage
constexpr int ST_OP_BOX = ST_KC * 128;                 // one [16 x 32 floats] box = 2 KB
constexpr int ST_SEG_A = (ST_BM / 32) * ST_OP_BOX;     // 8 KB
constexpr int ST_IO_BOX = ST_BM * 128;                 // [128 rows x 32 floats] = 16 KB
constexpr int ST_SLOT_BYTES = 2 * ST_IO_BOX;           // W + W_m column slice [128 x 32] each: 32 KB
template <int BN> __host__ __device__ constexpr int st_seg_b() { return (BN / 32) * ST_OP_BOX; }
// stage = [vp | vn | hp | hn] and, in the exact mode, [| hp_lo | hn_lo]
template <int BN, bool SPLIT> __host__ __device__ constexpr int st_stage_bytes() {
    return 2 * ST_SEG_A + (SPLIT ? 4 : 2) * st_seg_b<BN>();
}
template <int BN, int STAGES, int NSLOT, bool SPLIT> __host__ __device__ constexpr int st_smem() {
    return STAGES * st_stage_bytes<BN, SPLIT>() + NSLOT * ST_SLOT_BYTES + 1024 /*barriers*/ + 1024 /*align*/;
}

struct StatsArgs {
    int V, H, B;
    int m_tiles, n_tiles, k_chunks;
    float lr, mom, wd, bsz;
    // PACK variants: operand images written by k_pack_ops (one contiguous piece per (tile row / column, chunk))
    const uint8_t* pa; const uint8_t* pa_lo; const uint8_t* pb;
    const uint32_t* flags; uint32_t gen;          // flags[0] == gen: some v value is not exactly representable in tf32
    uint64_t w_policy, wm_policy;   // L2 eviction priorities of the W and W_m streams
    int late_wait;
};

// ---- operand packing ------------------------------------------------------------------------------------------
// The statistics operands are MN-major for the tensor core (feature index contiguous, batch index = K): with TMA that
// means boxes of only [rows x 32 floats] -- ~100 two-kilobyte boxes per 128 x 128 tile, whose issue cost rivals the
// 16 big weight boxes of the tile on the SM's one TMA unit.  At small batch (the HBM-bound regime) the activations are
// tiny, so one pass over them writes, per (128-column block, 16-row chunk), the exact shared-memory image the MMA
// wants (32-byte-atom swizzle, Swizzle<2,5,2> on byte offsets: element (k, c) of a [16 x 32] box at
// k * 128 + (((c / 8) ^ (k & 3)) * 32) + (c % 8) * 4), zero-padded at the edges; the statistics kernel then fetches a
// whole stage with two bulk copies.  In the exact mode the same pass writes the tf32 remainders x - trunc19(x): of h
// into the same piece (always multiplied), of v into a twin image that is only multiplied when some remainder is
// non-zero (flags[0] = gen) -- binary states never are.
struct PackArgs {
    const float* vp; const float* vn; const float* hp; const float* hn;
    int B, V, H, m_tiles, n_tiles, k_chunks, split;
    uint8_t* pa; uint8_t* pa_lo; uint8_t* pb;
    uint32_t* flags; uint32_t gen;
    const float* scan; int scan_rows;      // nullable [scan_rows, V]: only scanned; flags[1] = gen if some value is inexact
    int after_colstats;      // the predecessor triggered only after its own wait: dependents may be released at once
};

__global__ void __launch_bounds__(256) k_pack_ops(PackArgs a) {
    if (a.after_colstats) pdl_trigger();
    pdl_wait();
    if (!a.after_colstats) pdl_trigger();
    const int r = blockIdx.y;                                      // batch row (padded to k_chunks * 16)
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;          // float4 column over [A side | B side]
    const int na4 = a.m_tiles * 32, nb4 = a.n_tiles * 32;
    if (c4 >= na4 + nb4) return;
    const int kc = r / ST_KC, k = r % ST_KC;
    const bool is_a = c4 < na4;
    const int q = is_a ? c4 : c4 - na4;                            // float4 column inside its side
    const int tile = q >> 5, cb = (q >> 3) & 3, j = q & 7;
    const uint32_t box_off = (uint32_t)cb * ST_OP_BOX + (uint32_t)k * 128u +
                             ((((uint32_t)j >> 1) ^ (uint32_t)(k & 3)) << 5) + (((uint32_t)j & 1u) << 4);
    const int col = 4 * q, W = is_a ? a.V : a.H;
    const bool valid = r < a.B && col < W;
    const float* s0 = is_a ? a.vp : a.hp;
    const float* s1 = is_a ? a.vn : a.hn;
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (valid) {
        x0 = *reinterpret_cast<const float4*>(s0 + (size_t)r * W + col);
        x1 = *reinterpret_cast<const float4*>(s1 + (size_t)r * W + col);
    }
    if (is_a) {
        const size_t piece = ((size_t)tile * a.k_chunks + kc) * (2 * ST_SEG_A);
        *reinterpret_cast<float4*>(a.pa + piece + box_off) = x0;
        *reinterpret_cast<float4*>(a.pa + piece + ST_SEG_A + box_off) = x1;
        if (a.split) {
            const float4 l0 = tf32_lo4(x0), l1 = tf32_lo4(x1);
            *reinterpret_cast<float4*>(a.pa_lo + piece + box_off) = l0;
            *reinterpret_cast<float4*>(a.pa_lo + piece + ST_SEG_A + box_off) = l1;
            const bool nz = (l0.x != 0.f) | (l0.y != 0.f) | (l0.z != 0.f) | (l0.w != 0.f) |
                            (l1.x != 0.f) | (l1.y != 0.f) | (l1.z != 0.f) | (l1.w != 0.f);
            if (nz) a.flags[0] = a.gen;                            // (every writer stores the same value)
        }
    } else {
        constexpr int SEG_B = st_seg_b<128>();
        const size_t piece = ((size_t)tile * a.k_chunks + kc) * ((a.split ? 4 : 2) * SEG_B);
        *reinterpret_cast<float4*>(a.pb + piece + box_off) = x0;
        *reinterpret_cast<float4*>(a.pb + piece + SEG_B + box_off) = x1;
        if (a.split) {
            *reinterpret_cast<float4*>(a.pb + piece + 2 * SEG_B + box_off) = tf32_lo4(x0);
            *reinterpret_cast<float4*>(a.pb + piece + 3 * SEG_B + box_off) = tf32_lo4(x1);
        }
    }
}

// k_pack_ops fused with the column statistics of the CD update (k_colstats, rbm.py:216-226 / 478-483): the two passes
// read the same four matrices.  One block = one 128-column block of the v side or of the h side, all batch rows;
// thread = (float4 column, row lane): rows lane, lane + 8, ... are packed and summed by one thread, the eight lanes
// are added in lane order (the summation order of k_colstats, so dh / dv / sum pos_h are bit-identical to it); the
// biases of the block's columns are updated in place, the squared error goes through per-block partials and a ticket.
constexpr int PC_ROWS = 8;
__global__ void __launch_bounds__(32 * PC_ROWS) k_pack_colstats(PackArgs a, ColstatsJob cs) {
    __shared__ float4 red[2][PC_ROWS][32];
    __shared__ float red_sq[PC_ROWS][32];
    __shared__ unsigned int s_last;
    // trigger AFTER the wait (see k_colstats): the statistics kernel launched next may start at once -- its weight
    // stream shares nothing with this kernel, its operand path waits for this kernel's completion
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
    if ((int)blockIdx.x >= a.m_tiles + a.n_tiles) {
        // scan blocks: is the NEXT minibatch exactly representable in tf32?  (it joins `vp` in the forward pass that
        // follows the update; that pass then needs no remainder tiles for its activations)
        const int c = (blockIdx.x - a.m_tiles - a.n_tiles) * 128 + 4 * lane;
        bool bad = false;
        if (c < a.V)
            for (int r = ry; r < a.scan_rows; r += PC_ROWS) {
                const float4 l = tf32_lo4(*reinterpret_cast<const float4*>(a.scan + (size_t)r * a.V + c));
                bad |= (l.x != 0.f) | (l.y != 0.f) | (l.z != 0.f) | (l.w != 0.f);
            }
        if (bad) a.flags[1] = a.gen;
        return;
    }
    const bool is_a = (int)blockIdx.x < a.m_tiles;
    const int tile = is_a ? blockIdx.x : blockIdx.x - a.m_tiles;
    const int W = is_a ? a.V : a.H;
    const int col = tile * 128 + 4 * lane;
    const int cb = lane >> 3, j = lane & 7;
    const float* s0 = is_a ? a.vp : a.hp;
    const float* s1 = is_a ? a.vn : a.hn;
    const bool load_ea = is_a && cs.ea != a.vp, load_eb = is_a && cs.eb != a.vn;
    constexpr int SEG_B = st_seg_b<128>();
    const size_t piece_bytes = is_a ? (size_t)(2 * ST_SEG_A) : (size_t)((a.split ? 4 : 2) * SEG_B);
    const uint32_t seg = is_a ? ST_SEG_A : SEG_B;
    uint8_t* dst = is_a ? a.pa : a.pb;
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
    float sq = 0.f;
    bool nz = false;
    // the bias and bias-momentum entries this block updates at the end: fetched now, under the row loop
    float4 pre_m = make_float4(0.f, 0.f, 0.f, 0.f), pre_b = pre_m;
    const bool pre_ok = ry == 0 && cs.ba.apply && col + 3 < W &&
                        (((reinterpret_cast<uintptr_t>(is_a ? cs.ba.vbm : cs.ba.hbm) |
                           reinterpret_cast<uintptr_t>(is_a ? cs.ba.vb : cs.ba.hb)) & 15) == 0);
    if (pre_ok) {
        pre_m = *reinterpret_cast<const float4*>((is_a ? cs.ba.vbm : cs.ba.hbm) + col);
        pre_b = *reinterpret_cast<const float4*>((is_a ? cs.ba.vb : cs.ba.hb) + col);
    }
    const int rows = a.k_chunks * ST_KC;
    constexpr int UN = 8;                          // rows in flight per thread: all their loads are issued together
    for (int r0 = ry; r0 < rows; r0 += PC_ROWS * UN) {
        float4 x0[UN], x1[UN], e1[UN];
        bool valid[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int r = r0 + u * PC_ROWS;
            valid[u] = r < a.B && col < W;
            x0[u] = x1[u] = e1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[u]) {
                const size_t o = (size_t)r * W + col;
                x0[u] = *reinterpret_cast<const float4*>(s0 + o);
                x1[u] = *reinterpret_cast<const float4*>(s1 + o);
                if (load_eb) e1[u] = *reinterpret_cast<const float4*>(cs.eb + o);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int r = r0 + u * PC_ROWS;
            if (r >= rows) break;
            const int kc = r / ST_KC, k = r % ST_KC;
            if (valid[u]) {
                t0.x += x0[u].x; t0.y += x0[u].y; t0.z += x0[u].z; t0.w += x0[u].w;
                t1.x += x1[u].x; t1.y += x1[u].y; t1.z += x1[u].z; t1.w += x1[u].w;
                if (is_a) {
                    const float4 e0 = load_ea ? *reinterpret_cast<const float4*>(cs.ea + (size_t)r * W + col) : x0[u];
                    const float4 ee = load_eb ? e1[u] : x1[u];
                    float d;
                    d = e0.x - ee.x; sq = fmaf(d, d, sq); d = e0.y - ee.y; sq = fmaf(d, d, sq);
                    d = e0.z - ee.z; sq = fmaf(d, d, sq); d = e0.w - ee.w; sq = fmaf(d, d, sq);
                }
            }
            const uint32_t box_off = (uint32_t)cb * ST_OP_BOX + (uint32_t)k * 128u +
                                     ((((uint32_t)j >> 1) ^ (uint32_t)(k & 3)) << 5) + (((uint32_t)j & 1u) << 4);
            const size_t piece = ((size_t)tile * a.k_chunks + kc) * piece_bytes;
            *reinterpret_cast<float4*>(dst + piece + box_off) = x0[u];
            *reinterpret_cast<float4*>(dst + piece + seg + box_off) = x1[u];
            if (a.split) {
                const float4 l0 = tf32_lo4(x0[u]), l1 = tf32_lo4(x1[u]);
                if (is_a) {
                    *reinterpret_cast<float4*>(a.pa_lo + piece + box_off) = l0;
                    *reinterpret_cast<float4*>(a.pa_lo + piece + seg + box_off) = l1;
                    nz |= (l0.x != 0.f) | (l0.y != 0.f) | (l0.z != 0.f) | (l0.w != 0.f) |
                          (l1.x != 0.f) | (l1.y != 0.f) | (l1.z != 0.f) | (l1.w != 0.f);
                } else {
                    *reinterpret_cast<float4*>(dst + piece + 2 * seg + box_off) = l0;
                    *reinterpret_cast<float4*>(dst + piece + 3 * seg + box_off) = l1;
                }
            }
        }
    }
    if (nz) a.flags[0] = a.gen;
    red[0][ry][lane] = t0; red[1][ry][lane] = t1; red_sq[ry][lane] = sq;
    __syncthreads();
    if (ry == 0) {
        float p[4] = {0.f, 0.f, 0.f, 0.f}, n[4] = {0.f, 0.f, 0.f, 0.f};
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < PC_ROWS; ++q) {
            const float4 u = red[0][q][lane], v = red[1][q][lane];
            p[0] += u.x; p[1] += u.y; p[2] += u.z; p[3] += u.w;
            n[0] += v.x; n[1] += v.y; n[2] += v.z; n[3] += v.w;
            s += red_sq[q][lane];
        }
        const BiasArgs& ba = cs.ba;
        const float pm[4] = {pre_m.x, pre_m.y, pre_m.z, pre_m.w}, pb[4] = {pre_b.x, pre_b.y, pre_b.z, pre_b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = col + e;
            if (c >= W) break;
            const float d = p[e] - n[e];
            if (is_a) {
                cs.out[a.H + c] = d;
                if (ba.apply) {                               // rbm.py:223-224 / 480-481
                    const float m = add_rn(mul_rn(pre_ok ? pm[e] : ba.vbm[c], ba.mom), mul_rn(ba.lr, d) / ba.bsz);
                    ba.vbm[c] = m;
                    ba.vb[c] = add_rn(pre_ok ? pb[e] : ba.vb[c], m);
                }
            } else {
                cs.out[c] = d; cs.out[a.H + a.V + c] = p[e];
                if (ba.apply) {                               // rbm.py:216-220 / 478-479
                    float m = add_rn(mul_rn(pre_ok ? pm[e] : ba.hbm[c], ba.mom), mul_rn(ba.lr, d) / ba.bsz);
                    if (ba.sparsity) m = add_rn(m, mul_rn(-ba.lr, add_rn(p[e] / ba.bsz, -ba.sp_target)));
                    ba.hbm[c] = m;
                    ba.hb[c] = add_rn(pre_ok ? pb[e] : ba.hb[c], m);
                }
            }
        }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            cs.sq_part[blockIdx.x] = s;                       // (h-side blocks contribute 0)
            __threadfence();
            s_last = (atomicAdd(cs.ticket, 1u) == (unsigned)(a.m_tiles + a.n_tiles) - 1) ? 1u : 0u;
        }
        __syncwarp();
        if (s_last) {                                         // last block: partials in index order (deterministic)
            __threadfence();
            float v = 0.f;
            for (int i = lane; i < a.m_tiles; i += 32) v += __ldcg(cs.sq_part + i);
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) {
                cs.out[2 * a.H + a.V] = v;
                if (ba.loss_out) *ba.loss_out = v / ba.n_loss;
                *cs.ticket = 0u;
            }
        }
    }
}

template <bool UPDATE, int ST_BN, int ST_STAGES, int ST_NSLOT, bool PACK, bool SPLIT>
__global__ void __launch_bounds__(ST_THREADS, 1)
k_tc_stats(const __grid_constant__ CUtensorMap tmVP, const __grid_constant__ CUtensorMap tmVN,
           const __grid_constant__ CUtensorMap tmHP, const __grid_constant__ CUtensorMap tmHN,
           const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWm, StatsArgs a) {
    static_assert(!SPLIT || PACK, "the exact mode reads packed operands");
    static_assert(!PACK || ST_BN == 128, "operands are packed for 128-column tiles");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int ST_SEG_B = st_seg_b<ST_BN>();
    constexpr int ST_STAGE_BYTES = st_stage_bytes<ST_BN, SPLIT>();
    constexpr int ST_SLICES = ST_BN / 32;                      // 32-column slices of W / W_m per tile
    static_assert(ST_STAGES <= 4 && ST_NSLOT <= 4 && (ST_NSLOT & (ST_NSLOT - 1)) == 0, "ring sizes");
    uint8_t* ops = smem;                                       // operand ring
    uint8_t* slots = smem + ST_STAGES * ST_STAGE_BYTES;        // IO slots
    uint64_t* bars = reinterpret_cast<uint64_t*>(slots + ST_NSLOT * ST_SLOT_BYTES);
    uint64_t* ops_full = bars;            // [ST_STAGES <= 4]
    uint64_t* ops_empty = bars + 4;       // [ST_STAGES]
    uint64_t* acc_full = bars + 8;        // [2]
    uint64_t* acc_empty = bars + 10;      // [2]
    uint64_t* io_full = bars + 12;        // [ST_NSLOT <= 4]
    uint64_t* io_written = bars + 16;     // [ST_NSLOT]
    uint64_t* io_empty = bars + 20;       // [ST_NSLOT]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_total = a.m_tiles * a.n_tiles;
    // Round-robin tile order: at any moment the CTAs of the grid work on CONSECUTIVE tiles, i.e. on
    // adjacent column blocks of the same 128 weight rows, so DRAM sees whole rows streamed together
    // (row-buffer locality) instead of 512-byte fragments.
    const int t_beg = blockIdx.x, t_end = n_tiles_total, t_step = gridDim.x;

    pdl_trigger();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmVP); tma_prefetch_desc(&tmVN); tma_prefetch_desc(&tmHP); tma_prefetch_desc(&tmHN);
        tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmWm);
        for (int s = 0; s < ST_STAGES; ++s) { mbar_init(&ops_full[s], 1); mbar_init(&ops_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], ST_EPI_WARPS); }
        for (int s = 0; s < ST_NSLOT; ++s) {
            mbar_init(&io_full[s], 1); mbar_init(&io_written[s], ST_EPI_WARPS / 2); mbar_init(&io_empty[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * ST_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // late_wait (TMA-operand variant): the predecessor is k_colstats, which fires its trigger only after ITS
    // predecessors have completed and touches nothing this kernel reads or writes: run concurrently with it and
    // restore the completion chain with a wait at the very end.
    // PACK variants: the predecessors are k_colstats (same property) and k_pack_ops, whose output only the operand
    // path reads: the weight stream starts at once, the operand producer and the issuer wait for the pack kernel.
    if (!PACK && !a.late_wait) pdl_wait();

    if (warp == 0) {
        // ===================== operand producer =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            bool ext = false;
            if (PACK) {
                pdl_wait();
                ext = SPLIT && __ldcg(a.flags) == a.gen;
            }
            for (int t = t_beg; t < t_end; t += t_step) {
                const int mt = t / a.n_tiles, nt = t % a.n_tiles;
                const int m0 = mt * ST_BM, n0 = nt * ST_BN;
                const int n_chunks = a.k_chunks * (ext ? 2 : 1);
                for (int c = 0; c < n_chunks; ++c) {
                    const int kc = c % a.k_chunks;
                    const int b0 = kc * ST_KC;
                    uint8_t* s0 = ops + stage * ST_STAGE_BYTES;
                    mbar_wait(&ops_empty[stage], phase ^ 1);
                    if (PACK) {
                        const size_t pa_off = ((size_t)mt * a.k_chunks + kc) * (2 * ST_SEG_A);
                        const size_t pb_off = ((size_t)nt * a.k_chunks + kc) * ((SPLIT ? 4 : 2) * ST_SEG_B);
                        if (c < a.k_chunks) {
                            mbar_expect_tx(&ops_full[stage], ST_STAGE_BYTES);
                            bulk_load(s0, a.pa + pa_off, 2 * ST_SEG_A, &ops_full[stage]);
                            bulk_load(s0 + 2 * ST_SEG_A, a.pb + pb_off, (SPLIT ? 4 : 2) * ST_SEG_B, &ops_full[stage]);
                        } else {          // second round of an inexact v: [v_lo | h]
                            mbar_expect_tx(&ops_full[stage], 2 * ST_SEG_A + 2 * ST_SEG_B);
                            bulk_load(s0, a.pa_lo + pa_off, 2 * ST_SEG_A, &ops_full[stage]);
                            bulk_load(s0 + 2 * ST_SEG_A, a.pb + pb_off, 2 * ST_SEG_B, &ops_full[stage]);
                        }
                    } else {
                        mbar_expect_tx(&ops_full[stage], ST_STAGE_BYTES);
#pragma unroll
                        for (int cb = 0; cb < ST_BM / 32; ++cb) {
                            tma_load_2d(s0 + cb * ST_OP_BOX, &tmVP, m0 + cb * 32, b0, &ops_full[stage]);
                            tma_load_2d(s0 + ST_SEG_A + cb * ST_OP_BOX, &tmVN, m0 + cb * 32, b0, &ops_full[stage]);
                        }
#pragma unroll
                        for (int cb = 0; cb < ST_BN / 32; ++cb) {
                            tma_load_2d(s0 + 2 * ST_SEG_A + cb * ST_OP_BOX, &tmHP, n0 + cb * 32, b0, &ops_full[stage]);
                            tma_load_2d(s0 + 2 * ST_SEG_A + ST_SEG_B + cb * ST_OP_BOX, &tmHN, n0 + cb * 32, b0,
                                        &ops_full[stage]);
                        }
                    }
                    if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t id_pos = idesc_tf32(ST_BM, ST_BN, true, true, false);
            const uint32_t id_neg = idesc_tf32(ST_BM, ST_BN, true, true, true);     // (-A) * B
            int stage = 0; uint32_t phase = 0;
            bool ext = false;
            if (PACK) {
                pdl_wait();
                ext = SPLIT && __ldcg(a.flags) == a.gen;
            }
            int seg = 0;
            for (int t = t_beg; t < t_end; t += t_step, ++seg) {
                const int buf = seg & 1;
                mbar_wait(&acc_empty[buf], ((seg >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ST_BN);
                const int n_chunks = a.k_chunks * (ext ? 2 : 1);
                for (int c = 0; c < n_chunks; ++c) {
                    mbar_wait(&ops_full[stage], phase);
                    tc_fence_after();
                    const uint32_t s0 = smem_u32(ops + stage * ST_STAGE_BYTES);
#pragma unroll
                    for (int g = 0; g < ST_KC / 8; ++g) {
                        const uint64_t ap = smem_desc(s0 + g * 1024, ST_OP_BOX, 512, LAYOUT_SW128_BASE32B);
                        const uint64_t an = smem_desc(s0 + ST_SEG_A + g * 1024, ST_OP_BOX, 512, LAYOUT_SW128_BASE32B);
                        const uint64_t bp = smem_desc(s0 + 2 * ST_SEG_A + g * 1024, ST_OP_BOX, 512, LAYOUT_SW128_BASE32B);
                        const uint64_t bn = smem_desc(s0 + 2 * ST_SEG_A + ST_SEG_B + g * 1024, ST_OP_BOX, 512,
                                                      LAYOUT_SW128_BASE32B);
                        mma_tf32(d_tmem, ap, bp, id_pos, (c | g) != 0);        // (second round: v_lo * h)
                        mma_tf32(d_tmem, an, bn, id_neg, 1u);
                        if (SPLIT && c < a.k_chunks) {                          // v * h_lo
                            const uint64_t bpl = smem_desc(s0 + 2 * ST_SEG_A + 2 * ST_SEG_B + g * 1024, ST_OP_BOX, 512,
                                                           LAYOUT_SW128_BASE32B);
                            const uint64_t bnl = smem_desc(s0 + 2 * ST_SEG_A + 3 * ST_SEG_B + g * 1024, ST_OP_BOX, 512,
                                                           LAYOUT_SW128_BASE32B);
                            mma_tf32(d_tmem, ap, bpl, id_pos, 1u);
                            mma_tf32(d_tmem, an, bnl, id_neg, 1u);
                        }
                    }
                    mma_commit(&ops_empty[stage]);
                    if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
                }
                mma_commit(&acc_full[buf]);
            }
        }
    } else if (warp == 2) {
        // ===================== IO producer: W / W_m quarter-tiles (128 rows x 32 columns) ============
        if (UPDATE && elect_one()) {
            int hh = 0;
            for (int t = t_beg; t < t_end; t += t_step) {
                const int m0 = (t / a.n_tiles) * ST_BM, n0 = (t % a.n_tiles) * ST_BN;
                for (int qt = 0; qt < ST_SLICES; ++qt, ++hh) {
                    const int s = hh & (ST_NSLOT - 1);
                    uint8_t* slot = slots + s * ST_SLOT_BYTES;
                    mbar_wait(&io_empty[s], ((hh / ST_NSLOT) & 1) ^ 1);
                    mbar_expect_tx(&io_full[s], ST_SLOT_BYTES);
                    tma_load_2d_hint(slot, &tmW, n0 + qt * 32, m0, &io_full[s], a.w_policy);
                    tma_load_2d_hint(slot + ST_IO_BOX, &tmWm, n0 + qt * 32, m0, &io_full[s], a.wm_policy);
                }
            }
        }
    } else if (warp == 3) {
        // ===================== store issuer =====================
        if (elect_one()) {
            int hh = 0;
            for (int t = t_beg; t < t_end; t += t_step) {
                const int m0 = (t / a.n_tiles) * ST_BM, n0 = (t % a.n_tiles) * ST_BN;
                for (int qt = 0; qt < ST_SLICES; ++qt, ++hh) {
                    const int s = hh & (ST_NSLOT - 1);
                    const uint8_t* slot = slots + s * ST_SLOT_BYTES;
                    mbar_wait(&io_written[s], (hh / ST_NSLOT) & 1);
                    tma_store_2d_hint(&tmW, n0 + qt * 32, m0, slot, UPDATE ? a.w_policy : L2_EVICT_NORMAL);
                    if (UPDATE) tma_store_2d_hint(&tmWm, n0 + qt * 32, m0, slot + ST_IO_BOX, a.wm_policy);
                    tma_store_commit();
                    tma_store_wait_read();            // shared memory of the slot may be overwritten
                    mbar_arrive(&io_empty[s]);
                }
            }
            tma_store_wait_all();                     // global writes complete before the kernel ends
        }
    } else if (warp >= 4 && warp < 4 + ST_EPI_WARPS) {
        // ===================== epilogue =====================
        // Two groups of four warps; group g owns the quarter-tiles g and g+2 of every tile.  Inside a
        // group, warp <-> TMEM lane quadrant, thread <-> tile row.
        const int quad = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int row = quad * 32 + lane;
        const uint32_t sw = (uint32_t)(row & 7);      // 128-byte-swizzle phase of this row
        int seg = 0;
        for (int t = t_beg; t < t_end; t += t_step, ++seg) {
            const int buf = seg & 1;
            mbar_wait(&acc_full[buf], (seg >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int qi = 0; qi < ST_SLICES / 2; ++qi) {
                const int qt = grp + 2 * qi;
                const int hh = seg * ST_SLICES + qt;
                const int s = hh & (ST_NSLOT - 1);
                uint8_t* wrow = slots + s * ST_SLOT_BYTES + row * 128;
                uint8_t* mrow = wrow + ST_IO_BOX;
                float acc[32];
                {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * ST_BN + qt * 32);
                    float v0[16], v1[16];
                    tmem_ld16(taddr, v0);
                    tmem_ld16(taddr + 16, v1);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { acc[i] = v0[i]; acc[16 + i] = v1[i]; }
                }
                if (qi == ST_SLICES / 2 - 1) {         // this warp has read its share of the accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                if (UPDATE) mbar_wait(&io_full[s], (hh / ST_NSLOT) & 1);
                else                      mbar_wait(&io_empty[s], ((hh / ST_NSLOT) & 1) ^ 1);
                const float inv_bsz = 1.0f / a.bsz;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t off = ((uint32_t)j ^ sw) << 4;
                    float4* wp = reinterpret_cast<float4*>(wrow + off);
                    if (UPDATE) {
                        float4* mp = reinterpret_cast<float4*>(mrow + off);
                        const float4 w4 = *wp, m4 = *mp;
                        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
                        const float mv[4] = {m4.x, m4.y, m4.z, m4.w};
                        float nw[4], nm[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            // W_m <- mom*W_m + lr*((S+ - S-)/bsz - wd*W);  W <- W + W_m
                            const float g = add_rn(div_by(acc[4 * j + e], a.bsz, inv_bsz), -mul_rn(a.wd, wv[e]));
                            nm[e] = add_rn(mul_rn(mv[e], a.mom), mul_rn(a.lr, g));
                            nw[e] = add_rn(wv[e], nm[e]);
                        }
                        *mp = make_float4(nm[0], nm[1], nm[2], nm[3]);
                        *wp = make_float4(nw[0], nw[1], nw[2], nw[3]);
                    } else {
                        *wp = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                    }
                }
                fence_proxy_async();                   // generic-proxy writes -> visible to the TMA store
                __syncwarp();
                if (lane == 0) mbar_arrive(&io_written[s]);
            }
        }
    }

    if (PACK || a.late_wait) pdl_wait();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * ST_BN);
    }
}
