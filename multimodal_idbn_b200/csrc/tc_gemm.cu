// Tensor-core path (sm_100a): tcgen05.mma kind::tf32 with fp32 accumulators in TMEM, operands fed
// by TMA straight from the fp32 master tensors (no shadow copies), warp-specialised
// producer / MMA-issuer / epilogue roles synchronised with mbarriers.
//
//  k_tc_stream<A_MN, SPLIT, LOM, FUSE>  the weight-streaming passes:
//        up   (rbm.py:92)  D[h, b] = sum_v W[v,h] a[b,v]   A = W tile, MN-major (h contiguous)
//        down (rbm.py:96)  D[v, b] = sum_h W[v,h] a[b,h]   A = W tile, K-major
//     "swap-AB": the 128-row MMA M dimension is the OUTPUT FEATURE, the batch (<= 256) is N, so a
//     batch of 64 uses the full datapath.  Work is stream-K partitioned (SKPlan): every CTA streams
//     an equal, contiguous share of W exactly once; per-tile partial sums go to slabs that the finish
//     kernels (bias / sigmoid / Philox sampling) add in a fixed order.
//     SPLIT = exact mode (every operand as tf32 head + remainder); LOM = where the converters put the weight terms
//     (2: tensor memory, both products read A from TMEM; 0: shared-memory ring); FUSE = large batches in the fast
//     mode: persistent over (tile, batch chunk) units, the epilogue finishes the pass itself (no slabs, no finish kernel).
//
//  k_tc_stats<UPDATE> CD statistics + update (rbm.py:200,209,212-213):
//        dS[v,h] = sum_b vp[b,v] hp[b,h] - sum_b vn[b,v] hn[b,h]   (both operands MN-major; the
//        negative phase uses the a_negate bit of the instruction descriptor), 128x128 tiles,
//        double-buffered TMEM accumulators; the epilogue stages the tile through shared memory and
//        streams W / W_m with coalesced accesses:  W_m <- mom W_m + lr(dS/B - wd W);  W <- W + W_m.
#include <cuda.h>

#include <unordered_map>

#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace imdbn {

using namespace ptx;

// ------------------------------------------------------------------------------------------------
// host state: driver entry point + tensor-map cache
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

struct MapKey {
    const void* ptr; int inner, outer, box_inner, box_outer;   // box_inner: 32 = SW128, -32 = SW128 with 32B atoms
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && inner == o.inner && outer == o.outer && box_inner == o.box_inner &&
               box_outer == o.box_outer;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h = h * 1000003u ^ (size_t)k.inner; h = h * 1000003u ^ (size_t)k.outer;
        h = h * 1000003u ^ (size_t)k.box_inner; h = h * 1000003u ^ (size_t)k.box_outer;
        return h;
    }
};

struct TcState {
    EncodeTiledFn encode = nullptr;
    bool attrs_set = false;
    std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
};

static TcState* tc_state(imdbn_ctx* ctx) {
    if (!ctx->tc) {
        TcState* s = new TcState();
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            s->encode = (EncodeTiledFn)fn;
        ctx->tc = s;
    }
    return static_cast<TcState*>(ctx->tc);
}

void tc_destroy(imdbn_ctx* ctx) {
    delete static_cast<TcState*>(ctx->tc);
    ctx->tc = nullptr;
}

// 2-D fp32 row-major [outer, inner] tensor, box [box_outer, box_inner = 32 floats = 128 B], SW128.
// atom32 selects CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (for MN-major tf32 operands).
static const CUtensorMap* get_map(imdbn_ctx* ctx, const float* ptr, int inner, int outer, int box_outer,
                                  bool atom32) {
    TcState* s = tc_state(ctx);
    MapKey k{ptr, inner, outer, atom32 ? -32 : 32, box_outer};
    auto it = s->maps.find(k);
    if (it != s->maps.end()) return &it->second;
    if (s->maps.size() > 4096) s->maps.clear();
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)inner * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = s->encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nullptr;
    return &s->maps.emplace(k, m).first->second;
}

// ------------------------------------------------------------------------------------------------
// weight-streaming pass kernel
// ------------------------------------------------------------------------------------------------
// SPLIT = the exact mode (IMDBN_PREC_TF32X2): tcgen05 kind::tf32 reads only the top 19 bits of an fp32 word, so
// every operand x is used as x = hi + lo with hi = the truncated word the hardware sees and lo = x - hi (exact in
// fp32, <= 13 significant bits, of which the hardware keeps 11: 22+ bits of x in total).  Four converter warps
// turn each W tile that TMA delivered into its lo tile in a second shared-memory buffer (element-wise, so the
// swizzled layout is preserved) and the issuer adds   W_lo * a   to the same TMEM accumulator as   W * a;
// the activation tile gets the same treatment (a_lo, third product W * a_lo) but its product is skipped when the
// whole tile is exactly representable (binary states always are).  W_lo * a_lo (2^-22 relative) is dropped.
constexpr int TS_BM = 128;                 // output features per tile (MMA M)
constexpr int TS_MAX_STAGES = 8;
constexpr int TS_NLO = 4;                  // W_lo buffers (exact mode): a ring, so the converters never wait for products
constexpr int TS_NGRP = 2;                 // converter groups of four warps, group g takes the iterations n % 2 == g
constexpr int TS_BAR_BYTES = 512;
template <bool SPLIT, bool FUSE = false> struct TsCfg {
    static constexpr int BK = SPLIT ? 32 : 64;          // reduction elements per pipeline stage
    // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, 6.. converters (exact mode) or four more epilogue warps (FUSE: the
    // finishing epilogue of a 128 x 256 tile is ~30k clocks for one warp per lane quadrant -- as long as the products of a
    // layer with K = 2048; two warps per quadrant split the batch columns)
    static constexpr int THREADS = SPLIT ? 192 + TS_NGRP * 128 : (FUSE ? 320 : 192);
    static constexpr int A_BYTES = TS_BM * BK * 4;
};

struct StreamArgs {
    int M_total, K_total, B, Npad;
    int B1;                                // > 0: rows [0,B1) come from tmB, rows [B1,B) from tmB2 (two source matrices)
    SKPlan sk;
    int total_iters;
    float* part;                           // [slab][B][M_total]
    uint32_t tmem_cols;
    int nbuf;                              // TMEM accumulator sets (2 unless the batch tile is too wide)
    int stages;
    int lo_tmem;                           // exact mode: 2 = the converters write W and W_lo into tensor memory behind the accumulators and both
                                           // products take their A operand from there (the tensor core reads no weight from shared memory);
                                           // 0 = W from shared memory, W_lo in a shared-memory ring (batch tiles too wide to leave TMEM room)
    int blo;                               // exact mode: activation remainders are computed (0 = the caller knows them to be zero)
    const uint32_t* hint; uint32_t hint_gen;   // nullable: hint[0], hint[1] != hint_gen  =>  the activations of THIS pass were
                                           // found exactly representable by the operand-packing pass: take the blo = 0 layout
    int stages_x;                          // ring depth of that layout
    int bar_off;                           // byte offset of the barrier block (behind the larger of the two layouts)
    int w_stable;                          // W is not written by any kernel still in flight: prefetch it before the wait
    uint64_t w_policy;                     // L2 eviction priority of the W stream
    FusedFinish fin;                       // FUSE instantiation: the epilogue finishes the pass instead of storing partials;
    int pchunks, punits, pband;            // its CTAs are persistent over (tile, batch chunk) units: CTA c takes the units c, c + G,
                                           // c + 2G, ... (the epilogue of one overlaps the products of the next); unit numbers walk
                                           // bands of `pband` tiles chunk by chunk, so the CTAs running at the same time share
                                           // ~pband weight tiles and ~G / pband activation chunks in L2
    unsigned long long* trace;             // nullable (IMDBN_TS_TRACE): [cta][8] globaltimer stamps
};

__host__ __device__ inline int ts_bk(bool split) { return split ? 32 : 64; }
// blo: the stage also holds the remainder tile of the activations (exact mode, activations not known to be exact)
__host__ __device__ inline int ts_stage_bytes(int Npad, bool split, bool blo) {
    return TS_BM * ts_bk(split) * 4 + ((split && blo) ? 2 : 1) * Npad * ts_bk(split) * 4;
}
__host__ __device__ inline size_t ts_smem_bytes(int Npad, int stages, bool split, bool blo) {
    return (size_t)stages * ts_stage_bytes(Npad, split, blo) + (split ? TS_NLO * TS_BM * 32 * 4 : 0) + TS_BAR_BYTES + 1024;
}

// lo part of an fp32 value under tf32 truncation (exact)
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float4 tf32_lo4(float4 x) {
    return make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
}

template <bool A_MN, bool SPLIT, int LOM, bool FUSE>
__global__ void __launch_bounds__(TsCfg<SPLIT, FUSE>::THREADS, 1)
k_tc_stream(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmB2, StreamArgs a) {
    constexpr int TS_BK = TsCfg<SPLIT>::BK;
    constexpr int TS_A_BYTES = TsCfg<SPLIT>::A_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = a.Npad * TS_BK * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.bar_off);
    uint64_t* full = bars;                      // [TS_MAX_STAGES]
    uint64_t* empty = bars + 8;                 // [TS_MAX_STAGES]
    uint64_t* acc_full = bars + 16;             // [2]
    uint64_t* acc_empty = bars + 18;            // [2]
    uint64_t* lo_full = bars + 20;              // [TS_NLO]
    uint64_t* lo_empty = bars + 24;             // [TS_NLO]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
    volatile uint32_t* blo_flags = reinterpret_cast<volatile uint32_t*>(bars + 30);   // [TS_MAX_STAGES][4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int k_iters = a.sk.k_iters;
    // iteration range of this CTA: a stream-K share of one batch chunk (blockIdx.y), or whole (tile, chunk) units
    // (FUSE: LOCAL iteration numbers, unit j of this CTA = global unit cta + j * gridDim.x)
    const int beg = FUSE ? 0 : sk_beg(a.sk, cta);
    const int end = FUSE ? ((a.punits - cta + (int)gridDim.x - 1) / (int)gridDim.x) * k_iters : sk_beg(a.sk, cta + 1);
    const int m_tiles_all = (a.M_total + TS_BM - 1) / TS_BM;
    auto unit_tc = [&](int j, int& tile, int& chunk) {      // band-major, chunk, tile inside the band
        const int u = cta + j * (int)gridDim.x;
        const int per_band = a.pband * a.pchunks;
        const int band = u / per_band, r = u - band * per_band;
        const int tb = min(a.pband, m_tiles_all - band * a.pband);       // tiles in this (possibly last, narrower) band
        chunk = r / tb;
        tile = band * a.pband + (r - chunk * tb);
    };
    auto tile_of = [&](int unit) { if (!FUSE) return unit; int t, c; unit_tc(unit, t, c); return t; };
    auto b0_of = [&](int unit) { if (!FUSE) return (int)blockIdx.y * a.Npad; int t, c; unit_tc(unit, t, c); return c * a.Npad; };
#ifdef IMDBN_TS_TRACE_BUILD       // nvcc -DIMDBN_TS_TRACE_BUILD: per-CTA stage stamps and counts of the waits that found their
                                  // barrier incomplete, for the launches selected by IMDBN_TS_TRACE=<first traced call>
#define TS_MARK(i) do { if (a.trace) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.trace[(blockIdx.x + gridDim.x * blockIdx.y) * 16 + (i)] = t_; } } while (0)
    int miss[3] = {0, 0, 0};
#define TS_WAIT(slot, bar, par) do { if (a.trace && !mbar_test_wait(bar, par)) ++miss[slot]; mbar_wait(bar, par); } while (0)
#define TS_MISS_FLUSH(slot0, n) do { if (a.trace) for (int i_ = 0; i_ < (n); ++i_) a.trace[(blockIdx.x + gridDim.x * blockIdx.y) * 16 + (slot0) + i_] = (unsigned long long)miss[i_]; } while (0)
#else
#define TS_MARK(i) do { } while (0)
#define TS_WAIT(slot, bar, par) mbar_wait(bar, par)
#define TS_MISS_FLUSH(slot0, n) do { } while (0)
#endif

    // A kernel that may read W before its wait must not let ITS successors run ahead of a weight update either:
    // without w_stable the trigger follows the wait (the successor starts once everything before this kernel is done).
    if (a.w_stable) pdl_trigger();
    if (threadIdx.x == 0) TS_MARK(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (a.B1) tma_prefetch_desc(&tmB2);
        for (int s = 0; s < TS_MAX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], FUSE ? 8 : 4); }
        for (int s = 0; s < TS_NLO; ++s) { mbar_init(&lo_full[s], 4); mbar_init(&lo_empty[s], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (!a.w_stable) { pdl_wait(); pdl_trigger(); }    // everything above overlapped the previous kernel's tail
    // layout of the ring: fixed by the host, or (exact mode, hint given) chosen here from the packing pass's verdict
    int blo_on = a.blo, n_stages = a.stages;
    if (SPLIT && a.hint != nullptr && __ldcg(a.hint) != a.hint_gen && __ldcg(a.hint + 1) != a.hint_gen) {
        blo_on = 0; n_stages = a.stages_x;
    }
    const int stage_bytes = ts_stage_bytes(a.Npad, SPLIT, blo_on != 0);
    uint8_t* lo_base = smem + n_stages * stage_bytes;                 // [TS_NLO][TS_A_BYTES] (exact mode, W_lo in shared memory)
    const uint32_t lo_col0 = (uint32_t)(a.nbuf * 2 * a.Npad);         // [TS_NLO][TS_BK columns] W_lo (lo_tmem 1) or [TS_NLO][W | W_lo] (2)
    const uint32_t slot_cols = (uint32_t)(LOM == 2 ? 2 * TS_BK : TS_BK);

    auto blo_any = [&](int st_) -> uint32_t {
        uint32_t f = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) f |= blo_flags[st_ * 4 + i];
        return f;
    };
    auto load_A = [&](int it, int stage) {
        const int unit = it / k_iters, kit = it - unit * k_iters, tile = tile_of(unit);
        const int m0 = tile * TS_BM, k0 = kit * TS_BK;
        uint8_t* sA = smem + stage * stage_bytes;
        if (A_MN) {          // W[k rows, 32 features] boxes: one per 32-feature column block
#pragma unroll
            for (int cb = 0; cb < TS_BM / 32; ++cb)
                tma_load_2d_hint(sA + cb * (TS_BK * 128), &tmA, m0 + cb * 32, k0, &full[stage], a.w_policy);
        } else {             // W[128 feature rows, 32 k] boxes: one per 32-wide k block
#pragma unroll
            for (int j = 0; j < TS_BK / 32; ++j)
                tma_load_2d_hint(sA + j * (TS_BM * 128), &tmA, k0 + j * 32, m0, &full[stage], a.w_policy);
        }
    };
    auto load_B = [&](int it, int stage) {
        const int unit = it / k_iters, kit = it - unit * k_iters, b0 = b0_of(unit);
        const int k0 = kit * TS_BK;
        uint8_t* sB = smem + stage * stage_bytes + TS_A_BYTES;
#pragma unroll
        for (int j = 0; j < TS_BK / 32; ++j) {
            tma_load_2d(sB + j * (a.Npad * 128), &tmB, k0 + j * 32, b0, &full[stage]);
            if (a.B1)      // second source matrix: its rows land behind the first one's (B1 % 8 == 0)
                tma_load_2d(sB + j * (a.Npad * 128) + a.B1 * 128, &tmB2, k0 + j * 32, 0, &full[stage]);
        }
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            const uint32_t tx = (uint32_t)(TS_A_BYTES + b_bytes);
            int pre = 0;
            if (a.w_stable) {            // the weights of the first ring of stages stream in while the predecessor ends
                pre = min(n_stages, end - beg);
                for (int i = 0; i < pre; ++i) { mbar_expect_tx(&full[i], tx); load_A(beg + i, i); }
                pdl_wait();
            }
            TS_MARK(1);
            int stage = 0; uint32_t phase = 0;
            for (int it = beg; it < end; ++it) {
                if (it - beg >= pre) {
                    TS_WAIT(0, &empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], tx);
                    load_A(it, stage);
                }
                load_B(it, stage);
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
            TS_MISS_FLUSH(8, 1);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (elect_one()) {
            const uint32_t idesc = idesc_tf32(TS_BM, a.Npad, A_MN, false, false);
            const uint32_t idesc_ts = idesc_tf32(TS_BM, a.Npad, false, false, false);     // A from tensor memory
            auto a_desc = [&](uint32_t sA, int g) -> uint64_t {
                if (A_MN)   // k-group g = rows 8g..8g+7 (two 4-row swizzle atoms) of every column-block box
                    return smem_desc(sA + g * 1024, TS_BK * 128, 512, LAYOUT_SW128_BASE32B);
                // K-major: 32-wide k block g/4, 32-byte step inside the swizzle row
                return smem_desc(sA + (g / 4) * (TS_BM * 128) + (g % 4) * 32, 16, 1024, LAYOUT_SW128);
            };
            auto b_desc = [&](uint32_t sB, int g) -> uint64_t {
                return smem_desc(sB + (g / 4) * (a.Npad * 128) + (g % 4) * 32, 16, 1024, LAYOUT_SW128);
            };
            int stage = 0; uint32_t phase = 0;
            // exact mode: the lo products of iteration n are issued after the hi products of iteration n+1, so the
            // converters work on a tile while the tensor pipe is busy with its neighbours
            int p_stage = -1, p_n = 0, p_buf = 0; bool p_last = false, p_first = false; uint32_t p_tmem = 0;
            auto issue_lo = [&]() {
                const int l = p_n & (TS_NLO - 1);
                TS_WAIT(1, &lo_full[l], (p_n / TS_NLO) & 1);
                tc_fence_after();
                const uint32_t sA = smem_u32(smem + p_stage * stage_bytes);
                const uint32_t sB = sA + TS_A_BYTES, sBlo = sB + (uint32_t)b_bytes;
                const uint32_t sAlo = smem_u32(lo_base + l * TS_A_BYTES);
                const uint32_t tAlo = tmem_base + lo_col0 + (uint32_t)l * slot_cols;
                const uint32_t blo = blo_on ? (blo_any(p_stage)) : 0u;
                // the remainder products (2^-11 of the main ones) have their own accumulator: added to the large
                // running sum one by one they would each cost it a truncation
                const uint32_t d_lo = p_tmem + (uint32_t)a.Npad;
                {
                    if (LOM != 0) {
#pragma unroll
                        for (int g = 0; g < TS_BK / 8; ++g)
                            mma_tf32_ts(d_lo, tAlo + g * 8, b_desc(sB, g), idesc_ts, (p_first && g == 0) ? 0u : 1u);
                    } else {
#pragma unroll
                        for (int g = 0; g < TS_BK / 8; ++g)
                            mma_tf32(d_lo, a_desc(sAlo, g), b_desc(sB, g), idesc, (p_first && g == 0) ? 0u : 1u);
                    }
                }
                if (blo) {
#pragma unroll
                    for (int g = 0; g < TS_BK / 8; ++g) mma_tf32(d_lo, a_desc(sA, g), b_desc(sBlo, g), idesc, 1u);
                }
                mma_commit(&empty[p_stage]);
                mma_commit(&lo_empty[l]);
                if (p_last) mma_commit(&acc_full[p_buf]);
            };
            int seg = 0, n = 0;
            for (int cur = beg; cur < end; ++seg) {
                const int tile = cur / k_iters, kit0 = cur - tile * k_iters;
                const int n_it = min(end - cur, k_iters - kit0);
                const int buf = seg % a.nbuf;
                // one accumulator set: the previous segment's deferred products must be issued (and its accumulator
                // handed to the epilogue) before this segment can wait for the set to be drained
                if (SPLIT && a.nbuf == 1 && p_stage >= 0) { issue_lo(); p_stage = -1; }
                mbar_wait(&acc_empty[buf], ((seg / a.nbuf) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * (SPLIT ? 2 : 1) * a.Npad);
                for (int i = 0; i < n_it; ++i, ++n) {
                    TS_WAIT(0, &full[stage], phase);
                    if (cur == beg && i == 0) TS_MARK(2);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * stage_bytes);
                    const uint32_t sB = sA + TS_A_BYTES;
                    if (SPLIT && LOM == 2) {
                        // both weight terms come from tensor memory, written by the converters: no deferral
                        const int l = n & (TS_NLO - 1);
                        TS_WAIT(1, &lo_full[l], (n / TS_NLO) & 1);
                        tc_fence_after();
                        const uint32_t tA = tmem_base + lo_col0 + (uint32_t)l * slot_cols;
                        const uint32_t d_lo = d_tmem + (uint32_t)a.Npad;
#pragma unroll
                        for (int g = 0; g < TS_BK / 8; ++g)
                            mma_tf32_ts(d_tmem, tA + g * 8, b_desc(sB, g), idesc_ts, (i | g) != 0);
#pragma unroll
                        for (int g = 0; g < TS_BK / 8; ++g)
                            mma_tf32_ts(d_lo, tA + TS_BK + g * 8, b_desc(sB, g), idesc_ts, (i | g) != 0);
                        if (blo_on && blo_any(stage)) {
                            const uint32_t sBlo = sB + (uint32_t)b_bytes;
#pragma unroll
                            for (int g = 0; g < TS_BK / 8; ++g) mma_tf32_ts(d_lo, tA + g * 8, b_desc(sBlo, g), idesc_ts, 1u);
                        }
                        { mma_commit(&empty[stage]); mma_commit(&lo_empty[l]); }
                        if (i == n_it - 1) mma_commit(&acc_full[buf]);
                        if (++stage == n_stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
#pragma unroll
                    for (int g = 0; g < TS_BK / 8; ++g)         // one MMA per 8 k (tf32 UMMA_K)
                        mma_tf32(d_tmem, a_desc(sA, g), b_desc(sB, g), idesc, (i | g) != 0);
                    if (SPLIT) {
                        if (p_stage >= 0) issue_lo();
                        p_stage = stage; p_n = n; p_buf = buf; p_last = (i == n_it - 1); p_first = (i == 0); p_tmem = d_tmem;
                    } else {
                        mma_commit(&empty[stage]);                  // smem slot free when these MMAs retire
                    }
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
                if (!SPLIT) mma_commit(&acc_full[buf]);
                cur += n_it;
            }
            if (SPLIT && p_stage >= 0) issue_lo();
            TS_MARK(3);
            TS_MISS_FLUSH(9, 2);
        }
    } else if (warp < (FUSE ? 10 : 6)) {
        // ===================== epilogue: TMEM -> partial slab (FUSE: -> finished activations) =====================
        const int quad = warp & 3;                               // TMEM lane quadrant this warp may read
        int seg = 0;
        for (int cur = beg; cur < end; ++seg) {
            const int unit = cur / k_iters, kit0 = cur - unit * k_iters, tile = tile_of(unit), b0 = b0_of(unit);
            const int n_it = min(end - cur, k_iters - kit0);
            const int buf = seg % a.nbuf;
            const int slab = FUSE ? 0 : cta - sk_cta_of(a.sk, tile * k_iters);
            mbar_wait(&acc_full[buf], (seg / a.nbuf) & 1);
            tc_fence_after();
            const int m = tile * TS_BM + quad * 32 + lane;
            float* dst = a.part + ((size_t)slab * a.B + b0) * a.M_total + m;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * (SPLIT ? 2 : 1) * a.Npad);
            if (FUSE) {
                // whole tile accumulated here: bias, temperature, sigmoid, Bernoulli sample -- thread = output feature m,
                // registers = 16 batch rows.  One Philox call yields the uniforms of 4 neighbouring features of a row:
                // lane q of each 4-lane group makes the calls of the rows i % 4 == q and the group exchanges them.
                const FusedFinish& f = a.fin;
                const bool m_ok = m < a.M_total;
                const float bias = m_ok ? __ldg(f.bias + m) : 0.0f;
                const uint32_t colq = (uint32_t)(m & ~3), q = (uint32_t)lane & 3u;
                const int half = (warp - 2) >> 2, c_lo = half * (a.Npad >> 1), c_hi = c_lo + (a.Npad >> 1);    // (Npad = 256)
                for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
                    float4 u[4];
                    if (f.s_out) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) u[j] = rf_uniform4(f.key, f.draw_u, (uint32_t)(b0 + c0 + 4 * j) + q, colq);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p = __fdividef(1.0f, 1.0f + __expf(-(__fadd_rn(v[i], bias) * f.invT)));
                        const size_t o = (size_t)(b0 + c0 + i) * a.M_total + m;
                        const bool ok = m_ok && b0 + c0 + i < a.B;
                        if (f.s_out) {
                            const int src = (lane & ~3) | (i & 3);
                            const float ux = __shfl_sync(0xffffffffu, u[i >> 2].x, src), uy = __shfl_sync(0xffffffffu, u[i >> 2].y, src);
                            const float uz = __shfl_sync(0xffffffffu, u[i >> 2].z, src), uw = __shfl_sync(0xffffffffu, u[i >> 2].w, src);
                            const float un = (q & 2u) ? ((q & 1u) ? uw : uz) : ((q & 1u) ? uy : ux);
                            if (ok) f.s_out[o] = (p > un) ? 1.0f : 0.0f;
                        }
                        if (ok && f.p_out) f.p_out[o] = p;
                    }
                }
            } else
            for (int c0 = 0; c0 < a.Npad; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (SPLIT) {
                    float w[16];
                    tmem_ld16(taddr + a.Npad + c0, w);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += w[i];
                }
                if (m < a.M_total) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (b0 + c0 + i < a.B) dst[(size_t)(c0 + i) * a.M_total] = v[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            cur += n_it;
            if (warp == 2 && lane == 0) TS_MARK(4 + min(seg, 1));
        }
    } else if (SPLIT) {
        // ===================== converters (exact mode): lo tiles of W and of the activations =====================
        // Two groups of four warps; group g converts the iterations n with n % 2 == g into W_lo buffer g, so two
        // tiles are in conversion at any time (the per-tile latency -- shared-memory round trip, proxy fence,
        // barrier hand-over -- is several times the issue time).
        const int ct = (threadIdx.x - 192) & 127, cw = ct >> 5, grp = (threadIdx.x - 192) >> 7;
        const int nB4 = b_bytes / 16;
        for (int it = beg + grp, n = grp; it < end; it += TS_NGRP, n += TS_NGRP) {
            const int l = n & (TS_NLO - 1);
            const int stage = n % n_stages;
            const uint32_t phase = (uint32_t)(n / n_stages) & 1u;
            const float4* sA = reinterpret_cast<const float4*>(smem + stage * stage_bytes);
            const float4* sB = reinterpret_cast<const float4*>(smem + stage * stage_bytes + TS_A_BYTES);
            float4* sBlo = reinterpret_cast<float4*>(smem + stage * stage_bytes + TS_A_BYTES + b_bytes);
            float4* sAlo = reinterpret_cast<float4*>(lo_base + l * TS_A_BYTES);
            TS_WAIT(0, &full[stage], phase);
            TS_WAIT(1, &lo_empty[l], ((n / TS_NLO) & 1) ^ 1);
            bool nz = false;
            {
                if (LOM != 0) {
                    // thread <-> row m of the tile (the tensor-memory lane this warp may write): its 32 k values
                    const int q = warp & 3, m = q * 32 + lane;
                    const uint8_t* tile = smem + stage * stage_bytes;
                    float x[32];
                    if (A_MN) {      // [column block q][k][32 features], 32-byte swizzle atoms
                        const uint8_t* col = tile + q * (TS_BK * 128) + (lane & 7) * 4;
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            x[k] = *reinterpret_cast<const float*>(col + k * 128 + (((lane >> 3) ^ (k & 3)) << 5));
                    } else {         // [row m][32 k], 16-byte swizzle chunks
                        const uint8_t* row = tile + m * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 v = *reinterpret_cast<const float4*>(row + ((j ^ (m & 7)) << 4));
                            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
                        }
                    }
                    tc_fence_after();
                    const uint32_t tslot = tmem_base + ((uint32_t)(q * 32) << 16) + lo_col0 + (uint32_t)l * slot_cols;
                    if (LOM == 2) tmem_st32_nowait(tslot, x);      // the word as it is: the tensor core reads its top 19 bits
#pragma unroll
                    for (int k = 0; k < 32; ++k) x[k] = tf32_lo(x[k]);
                    tmem_st32(tslot + (LOM == 2 ? TS_BK : 0), x);
                    tc_fence_before();
                } else {
#pragma unroll
                    for (int i = 0; i < TS_A_BYTES / 16 / 128; ++i) sAlo[ct + i * 128] = tf32_lo4(sA[ct + i * 128]);
                }
                if (blo_on)
                for (int i = ct; i < nB4; i += 128) {
                    const float4 lo = tf32_lo4(sB[i]);
                    nz |= (lo.x != 0.0f) | (lo.y != 0.0f) | (lo.z != 0.0f) | (lo.w != 0.0f);
                    sBlo[i] = lo;
                }
            }
            if (blo_on || LOM == 0) {
                nz = __any_sync(0xffffffffu, nz);
                if (lane == 0) blo_flags[stage * 4 + cw] = nz ? 1u : 0u;
                fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's smem reads
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&lo_full[l]);
        }
        if (warp == 6 && lane == 0) TS_MISS_FLUSH(12, 2);
    }

    tc_fence_before();
    __syncthreads();

    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

#include "tc_stats.cuh"

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// IMDBN_L2_W / IMDBN_L2_WM = normal | first | last : experiment switch for the L2 policies
static inline uint64_t l2_policy_env(const char* name, uint64_t dflt) {
    const char* e = getenv(name);
    if (!e) return dflt;
    if (!strcmp(e, "normal")) return L2_EVICT_NORMAL;
    if (!strcmp(e, "first")) return L2_EVICT_FIRST;
    if (!strcmp(e, "last")) return L2_EVICT_LAST;
    return dflt;
}
// Weights of a small layer (<= 16 MB for W and W_m together) are re-read by every pass of every step: ask L2 to keep
// them (evict_last) so that they survive the 240 MB a large layer streams through the cache when both train
// concurrently; the big layer's streams can be marked evict_first.  IMDBN_L2_SMALL / IMDBN_L2_BIG override.
static inline uint64_t l2_policy_for(const imdbn_rbm* r) {
    const bool small = (size_t)r->V * r->H * 8 <= ((size_t)16 << 20);
    return small ? l2_policy_env("IMDBN_L2_SMALL", L2_EVICT_NORMAL) : l2_policy_env("IMDBN_L2_BIG", L2_EVICT_NORMAL);
}
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int npad_of(int B) { return std::max(16, (B + 15) / 16 * 16); }
static inline uint32_t pow2_cols(int n) { uint32_t c = 32; while ((int)c < n) c <<= 1; return c; }
static inline bool tc_split(const imdbn_ctx* ctx) { return ctx->precision == IMDBN_PREC_TF32X2; }

bool tc_shape_ok(const imdbn_rbm* r, int B) {
    return r->V % 4 == 0 && r->H % 4 == 0 && aligned16(r->W) && B >= 1;
}

bool tc_up_supported(const imdbn_ctx* ctx, const imdbn_rbm* r, int B) {
    return tc_shape_ok(r, B) && tc_state(const_cast<imdbn_ctx*>(ctx))->encode != nullptr;
}
bool tc_down_supported(const imdbn_ctx* ctx, const imdbn_rbm* r, int B) { return tc_up_supported(ctx, r, B); }
bool tc_stats_supported(const imdbn_ctx* ctx, const imdbn_rbm* r, int B) {
    return tc_shape_ok(r, B) && tc_state(const_cast<imdbn_ctx*>(ctx))->encode != nullptr;
}

SKPlan tc_plan(const imdbn_ctx* ctx, int M_total, int K_total, int B) {
    SKPlan p;
    const int bk = ts_bk(tc_split(ctx));
    const int m_tiles = (M_total + TS_BM - 1) / TS_BM;
    p.k_iters = (K_total + bk - 1) / bk;
    const int total = m_tiles * p.k_iters;
    // at least 256 reduction elements per CTA: fewer, longer ranges for small layers (fewer slabs to add)
    int G = std::max(1, std::min(tc_sms(ctx), total / (256 / bk)));
    // Large batches are cut into 256-row chunks on blockIdx.y: once (chunks x output tiles) fills the chip, splitting
    // K as well only multiplies the partial slabs (12 slabs of 134 MB each per pass at batch 8192 on 10000 -> 4096)
    const int chunks = (B + 255) / 256;
    // (not in the exact mode: the tensor core's accumulator truncates, so the shorter accumulation chains of the
    // split are part of its error budget)
    if (!tc_split(ctx) && B > 256 && 4 * chunks * m_tiles >= 3 * tc_sms(ctx)) G = m_tiles;
    p.q = total / G;
    p.r = total % G;
    p.tile_w = TS_BM;
    return p;
}

int tc_plan_ctas(const SKPlan& p, int M_total) {
    const int total = ((M_total + TS_BM - 1) / TS_BM) * p.k_iters;
    return p.q > 0 ? (total - p.r) / p.q : total;      // total = G*q + r
}

int tc_plan_max_slabs(const SKPlan& p, int M_total) {
    const int m_tiles = (M_total + TS_BM - 1) / TS_BM;
    int mx = 1;
    for (int t = 0; t < m_tiles; ++t) mx = std::max(mx, sk_nslabs(p, t));
    return mx;
}

template <bool A_MN, bool SPLIT, int LOM, bool FUSE = false>
static int launch_stream(imdbn_ctx* ctx, const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmB2,
                         StreamArgs& a, int G, int chunks, cudaStream_t st) {
    const size_t smem = (size_t)a.bar_off + TS_BAR_BYTES + 1024;
    static size_t smem_set = 0;          // the attribute is sticky: raise it only when a larger size is needed
    if (smem > smem_set) {
        IMDBN_CUDA(ctx, cudaFuncSetAttribute(k_tc_stream<A_MN, SPLIT, LOM, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    IMDBN_CUDA(ctx, launch_pdl(k_tc_stream<A_MN, SPLIT, LOM, FUSE>, dim3(G, chunks), dim3(TsCfg<SPLIT, FUSE>::THREADS), smem, st, *tmA, *tmB,
                               *tmB2, a));
    IMDBN_CHECK_LAUNCH(ctx, "k_tc_stream");
    return 0;
}

// act2 != nullptr: the batch is the virtual concatenation [act (B1 rows) ; act2 (B - B1 rows)]
static int stream_pass(imdbn_ctx* ctx, const imdbn_rbm* r, const float* act, int B, float* part, bool up,
                       cudaStream_t st, const float* act2 = nullptr, int B1 = 0, const FusedFinish* fin = nullptr) {
    const int M_total = up ? r->H : r->V, K_total = up ? r->V : r->H;
    if (!aligned16(act)) return fail(ctx, -1, "tc pass: activation pointer must be 16-byte aligned");
    const bool split = tc_split(ctx);
    const int bk = ts_bk(split);
    StreamArgs a{};
    a.M_total = M_total; a.K_total = K_total; a.B = B; a.Npad = B > 256 ? 256 : npad_of(B);
    const int chunks = (B + a.Npad - 1) / a.Npad;
    a.sk = tc_plan(ctx, M_total, K_total, B);
    a.total_iters = ((M_total + TS_BM - 1) / TS_BM) * a.sk.k_iters;
    a.part = part;
    // exact mode: the W_lo ring goes to tensor memory whenever the accumulators leave it TS_NLO x 32 columns
    static const bool lo_smem_env = getenv("IMDBN_LO_SMEM") != nullptr;
    a.lo_tmem = (split && !lo_smem_env && 2 * a.Npad + TS_NLO * 64 <= 512) ? 2 : 0;
    const int ring_cols = a.lo_tmem * TS_NLO * 32;
    a.nbuf = (split ? 4 : 2) * a.Npad + ring_cols <= 512 ? 2 : 1;
    a.tmem_cols = pow2_cols(a.nbuf * (split ? 2 : 1) * a.Npad + ring_cols);
    a.w_policy = l2_policy_for(r);
    a.w_stable = ctx->w_stable ? 1 : 0;
    a.blo = (split && !ctx->act_exact) ? 1 : 0;
    const int lo_ring_bytes = (split && a.lo_tmem == 0) ? TS_NLO * TS_BM * 32 * 4 : 0;
    const int budget = 226 * 1024 - TS_BAR_BYTES - lo_ring_bytes;
    a.stages = std::max(2, std::min(split ? TS_MAX_STAGES : 4, budget / ts_stage_bytes(a.Npad, split, a.blo != 0)));
    { static const int cap = getenv("IMDBN_TS_STAGES") ? atoi(getenv("IMDBN_TS_STAGES")) : 0; if (cap > 0) a.stages = std::max(2, std::min(a.stages, cap)); }
    a.stages_x = a.stages;
    if (a.blo && ctx->act_hint) {
        a.hint = ctx->act_hint; a.hint_gen = ctx->act_hint_gen;
        a.stages_x = std::max(2, std::min(TS_MAX_STAGES, budget / ts_stage_bytes(a.Npad, split, false)));
    }
    a.bar_off = std::max(a.stages * ts_stage_bytes(a.Npad, split, a.blo != 0), a.stages_x * ts_stage_bytes(a.Npad, split, false)) +
                lo_ring_bytes;
    const int G = tc_plan_ctas(a.sk, M_total);
    // W is [V, H] row-major: inner = H.  up: boxes [bk k-rows x 32 h]; down: boxes [128 v-rows x 32 h]
    const CUtensorMap* tmA = get_map(ctx, r->W, r->H, r->V, up ? bk : TS_BM, up);
    const CUtensorMap* tmB = nullptr;
    const CUtensorMap* tmB2 = nullptr;
    if (act2) {
        if (B1 <= 0 || B1 % 8 || B1 >= B || B > 256 || !aligned16(act2))
            return fail(ctx, -1, "tc pass: two-source batch needs B1 % 8 == 0, B <= 256");
        a.B1 = B1;
        tmB = get_map(ctx, act, K_total, B1, B1, false);                  // box = the B1 rows of the first source
        tmB2 = get_map(ctx, act2, K_total, B - B1, a.Npad - B1, false);   // rows past B - B1 are zero-filled
    } else {
        tmB = get_map(ctx, act, K_total, B, a.Npad, false);
        tmB2 = tmB;
    }
    if (!tmA || !tmB || !tmB2) return fail(ctx, -5, "cuTensorMapEncodeTiled failed");
    // IMDBN_TS_TRACE=<n>: per-CTA globaltimer stamps of three large-layer launches from call n on (debug aid)
#ifdef IMDBN_TS_TRACE_BUILD
    static const bool trace_on = getenv("IMDBN_TS_TRACE") != nullptr;
#else
    static const bool trace_on = false;      // (the kernel carries the stamps only when built with -DIMDBN_TS_TRACE_BUILD)
#endif
    static unsigned long long* trace_buf = nullptr;
    static int big_calls = 0;
    static const int trace_from = trace_on ? std::max(40, atoi(getenv("IMDBN_TS_TRACE"))) : 0;     // (value = first traced call)
    const bool traced = trace_on && (size_t)r->V * r->H > (1u << 22) && ++big_calls >= trace_from && big_calls <= trace_from + 2;
    if (traced) {
        if (!trace_buf) cudaMalloc((void**)&trace_buf, 148 * 16 * 8);
        cudaMemsetAsync(trace_buf, 0, 148 * 16 * 8, st);
        a.trace = trace_buf;
    }
    int rc;
    if (fin) {
        if (split || act2 || a.sk.r != 0 || a.sk.q != a.sk.k_iters)
            return fail(ctx, -1, "tc pass: a fused finish needs whole tiles per CTA in the single-pass tf32 mode");
        a.fin = *fin;
        const int units = ((M_total + TS_BM - 1) / TS_BM) * chunks;          // (tile, batch chunk) pairs, whole K each
        const int Gp = std::min(tc_sms(ctx), units);
        a.pchunks = chunks; a.punits = units;
        { static const int band_env = getenv("IMDBN_FUSE_BAND") ? atoi(getenv("IMDBN_FUSE_BAND")) : 16; a.pband = std::max(1, band_env); }
        rc = up ? launch_stream<true, false, 0, true>(ctx, tmA, tmB, tmB2, a, Gp, 1, st)
                : launch_stream<false, false, 0, true>(ctx, tmA, tmB, tmB2, a, Gp, 1, st);
    } else if (split && a.lo_tmem == 2)
               rc = up ? launch_stream<true, true, 2>(ctx, tmA, tmB, tmB2, a, G, chunks, st)
                       : launch_stream<false, true, 2>(ctx, tmA, tmB, tmB2, a, G, chunks, st);
    else if (split)
               rc = up ? launch_stream<true, true, 0>(ctx, tmA, tmB, tmB2, a, G, chunks, st)
                       : launch_stream<false, true, 0>(ctx, tmA, tmB, tmB2, a, G, chunks, st);
    else       rc = up ? launch_stream<true, false, 0>(ctx, tmA, tmB, tmB2, a, G, chunks, st)
                       : launch_stream<false, false, 0>(ctx, tmA, tmB, tmB2, a, G, chunks, st);
    if (traced && rc == 0) {
        static unsigned long long h[148 * 16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int c = 0; c < G; ++c) t0 = std::min(t0, h[c * 16]);
        const char* names[7] = {"entry", "after pdl_wait", "first stage full", "last mma issued", "epilogue seg0", "epilogue seg1", "exit"};
        fprintf(stderr, "k_tc_stream<%s> B=%d G=%d trace (ns after first CTA entry; min / mean / max over CTAs)\n", up ? "up" : "down", B, G);
        for (int i = 0; i < 6; ++i) {
            double mn = 1e18, mx = 0, sum = 0; int n = 0;
            for (int c = 0; c < G; ++c) { if (!h[c * 16 + i]) continue; double v = (double)(h[c * 16 + i] - t0); mn = std::min(mn, v); mx = std::max(mx, v); sum += v; ++n; }
            if (n) fprintf(stderr, "  %-18s %8.0f %8.0f %8.0f  (n=%d)\n", names[i], mn, sum / n, mx, n);
        }
        const char* wn[5] = {"producer: empty", "mma: full", "mma: lo_full", "conv: full", "conv: lo_empty"};
        const int slot_of[5] = {8, 9, 10, 12, 13};
        for (int i = 0; i < 5; ++i) {       // waits that found their barrier incomplete (mean over CTAs; trace build only)
            double sum = 0;
            for (int c = 0; c < G; ++c) sum += (double)h[c * 16 + slot_of[i]];
            fprintf(stderr, "  waits %-16s %6.1f\n", wn[i], sum / G);
        }
    }
    return rc;
}

int tc_gemm_up(imdbn_ctx* ctx, const imdbn_rbm* r, const float* v, int B, float* part, cudaStream_t st,
               const float* v2, int B1, const FusedFinish* fin) {
    return stream_pass(ctx, r, v, B, part, true, st, v2, B1, fin);
}
int tc_gemm_down(imdbn_ctx* ctx, const imdbn_rbm* r, const float* h, int B, float* part, cudaStream_t st,
                 const FusedFinish* fin) {
    return stream_pass(ctx, r, h, B, part, false, st, nullptr, 0, fin);
}
bool tc_pass_fusable(const imdbn_ctx* ctx, const imdbn_rbm* r, int B, bool up) {
    static const bool off = getenv("IMDBN_NO_FUSED_FINISH") != nullptr;
    if (off || !fast_math(ctx) || B <= 256 || !(up ? tc_up_supported(ctx, r, B) : tc_down_supported(ctx, r, B))) return false;
    const SKPlan p = tc_plan(ctx, up ? r->H : r->V, up ? r->V : r->H, B);
    return p.r == 0 && p.q == p.k_iters;
}

// tile shape: HBM-bound small batches stream W / W_m through 128 x 128 tiles with packed operands; from 512 rows the
// kernel is tensor-bound and uses 128 x 256 tiles fed by TMA (the exact mode always takes the narrow shape)
static bool stats_wide(const imdbn_ctx* ctx, const imdbn_rbm* r, int B) {
    static const bool narrow_env = getenv("IMDBN_STATS_NARROW") != nullptr;
    return !tc_split(ctx) && B >= 512 && r->H >= 256 && !narrow_env;
}
struct PackSizes { size_t pa, pa_lo, pb, total; };
static PackSizes pack_sizes(const imdbn_rbm* r, int B, bool split) {
    const size_t mt = (r->V + ST_BM - 1) / ST_BM, nt = (r->H + 127) / 128, kc = (B + ST_KC - 1) / ST_KC;
    PackSizes p;
    p.pa = mt * kc * 2 * ST_SEG_A;
    p.pa_lo = split ? p.pa : 0;
    p.pb = nt * kc * (split ? 4 : 2) * st_seg_b<128>();
    p.total = p.pa + p.pa_lo + p.pb;
    return p;
}
bool tc_stats_packs(const imdbn_ctx* ctx, const imdbn_rbm* r, int B) {
    return uses_tc(ctx) && tc_stats_supported(ctx, r, B) && !stats_wide(ctx, r, B);
}
size_t tc_ws_bytes(const imdbn_ctx* ctx, const imdbn_rbm* r, int B) {
    if (!uses_tc(ctx) || !tc_shape_ok(r, B) || stats_wide(ctx, r, B)) return 0;
    return pack_sizes(r, B, tc_split(ctx)).total + 512;
}

int tc_gemm_stats(imdbn_ctx* ctx, const imdbn_rbm* r, const float* vp, const float* hp, const float* vn,
                  const float* hn, int B, float* dS_out, const imdbn_update* upd, cudaStream_t st) {
    if (!aligned16(vp) || !aligned16(hp) || !aligned16(vn) || !aligned16(hn) || (dS_out && !aligned16(dS_out)) ||
        (!dS_out && !aligned16(r->Wm)))
        return fail(ctx, -1, "tc stats: pointers must be 16-byte aligned");
    // tile shape: HBM-bound small batches stream W / W_m through 128 x 128 tiles and four IO slots; from 512 rows
    // the kernel is tensor-bound and uses 128 x 256 tiles (more MACs per operand byte), three operand stages
    const bool split = tc_split(ctx);
    const bool wide = stats_wide(ctx, r, B);
    const int BN = wide ? 256 : 128;
    StatsArgs a{};
    a.V = r->V; a.H = r->H; a.B = B;
    a.m_tiles = (r->V + ST_BM - 1) / ST_BM;
    a.n_tiles = (r->H + BN - 1) / BN;
    a.k_chunks = (B + ST_KC - 1) / ST_KC;
    if (upd) { a.lr = upd->lr; a.mom = upd->momentum; a.wd = upd->weight_decay; a.bsz = (float)upd->batch_global; }
    const bool after_colstats = ctx->stats_after_colstats || ctx->colstats_job != nullptr;
    a.late_wait = after_colstats ? 1 : 0;
    ctx->stats_after_colstats = false;
    a.w_policy = l2_policy_for(r);
    a.wm_policy = l2_policy_for(r);
    const CUtensorMap* tW = get_map(ctx, dS_out ? dS_out : r->W, r->H, r->V, ST_BM, false);
    const CUtensorMap* tWm = dS_out ? tW : get_map(ctx, r->Wm, r->H, r->V, ST_BM, false);
    // operand maps: only the wide (TMA-fed) variant reads them
    const CUtensorMap* tVP = wide ? get_map(ctx, vp, r->V, B, ST_KC, true) : tW;
    const CUtensorMap* tVN = wide ? get_map(ctx, vn, r->V, B, ST_KC, true) : tW;
    const CUtensorMap* tHP = wide ? get_map(ctx, hp, r->H, B, ST_KC, true) : tW;
    const CUtensorMap* tHN = wide ? get_map(ctx, hn, r->H, B, ST_KC, true) : tW;
    if (!tVP || !tVN || !tHP || !tHN || !tW || !tWm) return fail(ctx, -5, "cuTensorMapEncodeTiled failed");
    if (!wide) {
        // pack the operands into tile-ready images (workspace from the arena of the current API call)
        const PackSizes ps = pack_sizes(r, B, split);
        uint8_t* ws = arena_take<uint8_t>(ctx, ps.total);
        if (ctx->arena.off > ctx->arena.cap) return fail(ctx, -2, "tc stats: workspace not reserved (tc_ws_bytes)");
        if (!ctx->pack_flags) {
            IMDBN_CUDA(ctx, cudaMalloc((void**)&ctx->pack_flags, 256));
            IMDBN_CUDA(ctx, cudaMemsetAsync(ctx->pack_flags, 0, 256, st));
        }
        PackArgs pk{};
        pk.vp = vp; pk.vn = vn; pk.hp = hp; pk.hn = hn;
        pk.B = B; pk.V = r->V; pk.H = r->H; pk.m_tiles = a.m_tiles; pk.n_tiles = a.n_tiles; pk.k_chunks = a.k_chunks;
        pk.split = split ? 1 : 0;
        pk.pa = ws; pk.pa_lo = ws + ps.pa; pk.pb = ws + ps.pa + ps.pa_lo;
        pk.flags = ctx->pack_flags; pk.gen = ++ctx->pack_gen;
        if (pk.gen == 0) pk.gen = ++ctx->pack_gen;                 // 0 = the memset value, never a valid generation
        pk.after_colstats = after_colstats ? 1 : 0;
        const int n4 = (a.m_tiles + a.n_tiles) * 32;
        {
            ProfScope prof(ctx, IMDBN_KERNEL_PACK, r->V, r->H, st);
            if (ctx->colstats_job) {       // the column statistics of this update ride along (one pass over the matrices)
                pk.after_colstats = 1;
                const bool scan = split && ctx->pack_scan != nullptr && aligned16(ctx->pack_scan);
                if (scan) { pk.scan = ctx->pack_scan; pk.scan_rows = ctx->pack_scan_rows; }
                ctx->act_hint = scan ? ctx->pack_flags : nullptr;
                ctx->act_hint_gen = pk.gen;
                IMDBN_CUDA(ctx, launch_pdl(k_pack_colstats, dim3(a.m_tiles + a.n_tiles + (scan ? a.m_tiles : 0)),
                                           dim3(32 * PC_ROWS), 0, st, pk, *ctx->colstats_job));
            } else {
                IMDBN_CUDA(ctx, launch_pdl(k_pack_ops, dim3((n4 + 255) / 256, a.k_chunks * ST_KC), dim3(256), 0, st, pk));
            }
        }
        ctx->colstats_job = nullptr;
        ctx->pack_scan = nullptr;
        IMDBN_CHECK_LAUNCH(ctx, "k_pack");
        a.pa = pk.pa; a.pa_lo = pk.pa_lo; a.pb = pk.pb; a.flags = pk.flags; a.gen = pk.gen;
    }
    const int G = std::min(tc_sms(ctx), a.m_tiles * a.n_tiles);
    auto launch = [&](auto kernel, int smem, bool& attr_set) -> cudaError_t {
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            attr_set = true;
        }
        return launch_pdl(kernel, dim3(G), dim3(ST_THREADS), (size_t)smem, st, *tVP, *tVN, *tHP, *tHN, *tW, *tWm, a);
    };
    static bool set[6] = {false, false, false, false, false, false};
    cudaError_t e;
    ProfScope prof(ctx, IMDBN_KERNEL_STATS, r->V, r->H, st);
    if (split)
        e = dS_out ? launch(k_tc_stats<false, 128, 2, 4, true, true>, st_smem<128, 2, 4, true>(), set[4])
                   : launch(k_tc_stats<true, 128, 2, 4, true, true>, st_smem<128, 2, 4, true>(), set[5]);
    else if (wide)
        e = dS_out ? launch(k_tc_stats<false, 256, 3, 2, false, false>, st_smem<256, 3, 2, false>(), set[0])
                   : launch(k_tc_stats<true, 256, 3, 2, false, false>, st_smem<256, 3, 2, false>(), set[1]);
    else
        e = dS_out ? launch(k_tc_stats<false, 128, 3, 4, true, false>, st_smem<128, 3, 4, false>(), set[2])
                   : launch(k_tc_stats<true, 128, 3, 4, true, false>, st_smem<128, 3, 4, false>(), set[3]);
    IMDBN_CUDA(ctx, e);
    IMDBN_CHECK_LAUNCH(ctx, "k_tc_stats");
    return 0;
}

#include "chain_tc.cuh"

// TXT->IMG noisy mean-field chains as one persistent kernel (chain_tc.cuh): single-pass tf32, label block clamped
bool tc_chain_supported(const imdbn_ctx* ctx, const imdbn_rbm* r, const imdbn_chain* ch, int B) {
    static const bool off = getenv("IMDBN_NO_CHAIN_TC") != nullptr;
    if (off || ctx->precision != IMDBN_PREC_TF32 || !tc_shape_ok(r, B) || B < 512) return false;
    if (tc_state(const_cast<imdbn_ctx*>(ctx))->encode == nullptr) return false;
    if (ch->kind != IMDBN_CHAIN_NOISY_MF || ch->n_steps < 1 || ch->v_init || ch->sample_h || ch->sample_v) return false;
    const int Dz = ch->clamp_suffix;
    if (Dz <= 0 || Dz >= r->V || Dz > 512 || r->H > 256 || (r->H % 32) != 0) return false;
    if (ch->mu && ch->Dz != Dz) return false;
    for (int g = 0; g < r->ngroups; ++g)
        if (r->group_start[g] < Dz) return false;          // softmax groups only inside the clamped block
    return ct_smem_bytes((r->V + 31) / 32, r->H / 32) <= 227 * 1024;
}

// T / sigma / eta: DEVICE tables of n_steps entries each
int tc_chain_t2i(imdbn_ctx* ctx, const imdbn_rbm* r, const imdbn_chain* ch, int B, float* v_out, const RngKey& key,
                 const float* T, const float* sigma, const float* eta, cudaStream_t st) {
    ChainTcArgs a{};
    a.V = r->V; a.H = r->H; a.Dz = ch->clamp_suffix; a.B = B; a.n_steps = ch->n_steps;
    a.kb_v = (r->V + 31) / 32; a.kb_h = r->H / 32;
    a.mt_h = (r->H + 127) / 128; a.mt_v = (a.Dz + 127) / 128;
    a.hb = r->hb; a.vb = r->vb; a.v_known = ch->v_known; a.mu = ch->mu;
    a.T = T; a.sigma = sigma; a.eta = eta;
    a.v_out = v_out; a.key = key; a.draw0 = ch->draw0;
    const CUtensorMap* tUp = get_map(ctx, r->W, r->H, r->V, CT_BK, true);
    const CUtensorMap* tDn = get_map(ctx, r->W, r->H, r->V, 128, false);
    if (!tUp || !tDn) return fail(ctx, -5, "cuTensorMapEncodeTiled failed");
    const size_t smem = ct_smem_bytes(a.kb_v, a.kb_h);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        IMDBN_CUDA(ctx, cudaFuncSetAttribute(k_chain_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const int n_tiles = (B + CT_NC - 1) / CT_NC;
    const int G = std::min(tc_sms(ctx), n_tiles);
    k_chain_tc<<<G, CT_THREADS, smem, st>>>(*tUp, *tDn, a);
    IMDBN_CHECK_LAUNCH(ctx, "k_chain_tc");
    return 0;
}

// ---- SM partition (CUDA green contexts): two streams whose kernels run on disjoint sets of SMs ---------------
namespace {
struct SmPartition { bool tried = false; bool ok = false; CUstream big = nullptr, small_ = nullptr; int n_big = 0, n_small = 0; int want = 0; };
constexpr int kPartitionsPerDevice = 6;     // distinct reserve sizes a process may ask for
SmPartition g_partitions[16][kPartitionsPerDevice];

template <typename Fn>
bool drv(const char* name, Fn* out) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
        return false;
    *out = reinterpret_cast<Fn>(fn);
    return true;
}
}  // namespace

int sm_partition(int device, int small_sms, void** stream_big, void** stream_small, int* n_big, int* n_small) {
    if (device < 0 || device >= 16 || !stream_big || !stream_small || small_sms < 8) return -1;
    SmPartition* slot = nullptr;
    for (int i = 0; i < kPartitionsPerDevice && !slot; ++i)
        if (g_partitions[device][i].tried && g_partitions[device][i].want == small_sms) slot = &g_partitions[device][i];
    for (int i = 0; i < kPartitionsPerDevice && !slot; ++i)
        if (!g_partitions[device][i].tried) slot = &g_partitions[device][i];
    if (!slot) return -10;                  // too many different partition sizes on this device
    SmPartition& P = *slot;
    if (!P.tried) {
        P.want = small_sms;
        P.tried = true;
        typedef CUresult (*FGetDev)(CUdevice*, int);
        typedef CUresult (*FGetRes)(CUdevice, CUdevResource*, CUdevResourceType);
        typedef CUresult (*FSplit)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
        typedef CUresult (*FDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
        typedef CUresult (*FCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
        typedef CUresult (*FStream)(CUstream*, CUgreenCtx, unsigned int, int);
        FGetDev fGetDev; FGetRes fGetRes; FSplit fSplit; FDesc fDesc; FCreate fCreate; FStream fStream;
        if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) return -2;
        if (!drv("cuDeviceGet", &fGetDev) || !drv("cuDeviceGetDevResource", &fGetRes) ||
            !drv("cuDevSmResourceSplitByCount", &fSplit) || !drv("cuDevResourceGenerateDesc", &fDesc) ||
            !drv("cuGreenCtxCreate", &fCreate) || !drv("cuGreenCtxStreamCreate", &fStream))
            return -3;
        CUdevice dev;
        CUdevResource all, grp, rest;
        unsigned int nb = 1;
        CUdevResourceDesc d_small, d_big;
        CUgreenCtx g_small, g_big;
        if (fGetDev(&dev, device) != CUDA_SUCCESS || fGetRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return -4;
        if (fSplit(&grp, &nb, &all, &rest, 0, (unsigned int)small_sms) != CUDA_SUCCESS || nb < 1) return -5;
        if (fDesc(&d_small, &grp, 1) != CUDA_SUCCESS || fDesc(&d_big, &rest, 1) != CUDA_SUCCESS) return -6;
        if (fCreate(&g_small, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS ||
            fCreate(&g_big, d_big, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS)
            return -7;
        if (fStream(&P.small_, g_small, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS ||
            fStream(&P.big, g_big, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS)
            return -8;
        P.n_small = (int)grp.sm.smCount;
        P.n_big = (int)rest.sm.smCount;
        P.ok = true;
    }
    if (!P.ok) return -9;
    *stream_big = P.big; *stream_small = P.small_;
    if (n_big) *n_big = P.n_big;
    if (n_small) *n_small = P.n_small;
    return 0;
}


}  // namespace imdbn
