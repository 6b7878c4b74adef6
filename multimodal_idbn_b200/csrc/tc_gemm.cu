// Tensor-core path (sm_100a): tcgen05.mma kind::tf32 with fp32 accumulators in TMEM, operands fed
// by TMA straight from the fp32 master tensors (no shadow copies), warp-specialised
// producer / MMA-issuer / epilogue roles synchronised with mbarriers.
//
//  k_tc_stream<A_MN>  the weight-streaming passes at small batch (HBM-bound):
//        up   (rbm.py:92)  D[h, b] = sum_v W[v,h] a[b,v]   A = W tile, MN-major (h contiguous)
//        down (rbm.py:96)  D[v, b] = sum_h W[v,h] a[b,h]   A = W tile, K-major
//     "swap-AB": the 128-row MMA M dimension is the OUTPUT FEATURE, the batch (<= 256) is N, so a
//     batch of 64 uses the full datapath.  Work is stream-K partitioned (SKPlan): every CTA streams
//     an equal, contiguous share of W exactly once; per-tile partial sums go to slabs that the finish
//     kernels (bias / sigmoid / Philox sampling) add in a fixed order.
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
constexpr int TS_BM = 128;                 // output features per tile (MMA M)
constexpr int TS_BK = 64;                  // reduction elements per pipeline stage
constexpr int TS_STAGES = 4;                // maximum; fewer when the batch tile is wide (StreamArgs::stages)
constexpr int TS_THREADS = 192;            // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int TS_A_BYTES = TS_BM * TS_BK * 4;

struct StreamArgs {
    int M_total, K_total, B, Npad;
    int B1;                                // > 0: rows [0,B1) come from tmB, rows [B1,B) from tmB2 (two source matrices)
    SKPlan sk;
    int total_iters;
    float* part;                           // [slab][B][M_total]
    uint32_t tmem_cols;
    int stages;
    uint64_t w_policy;                     // L2 eviction priority of the W stream
    unsigned long long* trace;             // nullable (IMDBN_TS_TRACE): [cta][8] globaltimer stamps
};

__host__ __device__ inline int ts_stage_bytes(int Npad) { return TS_A_BYTES + Npad * TS_BK * 4; }

template <bool A_MN>
__global__ void __launch_bounds__(TS_THREADS, 1)
k_tc_stream(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmB2, StreamArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = ts_stage_bytes(a.Npad);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.stages * stage_bytes);
    uint64_t* full = bars;                      // [TS_STAGES]
    uint64_t* empty = bars + TS_STAGES;         // [TS_STAGES]
    uint64_t* acc_full = bars + 2 * TS_STAGES;  // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int b0 = blockIdx.y * a.Npad;          // batches wider than 256 rows: one 256-row chunk per blockIdx.y
    const int beg = sk_beg(a.sk, cta), end = sk_beg(a.sk, cta + 1);
    const int k_iters = a.sk.k_iters;
#define TS_MARK(i) do { if (a.trace) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.trace[(blockIdx.x + gridDim.x * blockIdx.y) * 8 + (i)] = t_; } } while (0)

    pdl_trigger();
    if (threadIdx.x == 0) TS_MARK(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (a.B1) tma_prefetch_desc(&tmB2);
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // everything above overlapped the previous kernel's tail
    if (threadIdx.x == 0) TS_MARK(1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int it = beg; it < end; ++it) {
                const int tile = it / k_iters, kit = it - tile * k_iters;
                const int m0 = tile * TS_BM, k0 = kit * TS_BK;
                uint8_t* sA = smem + stage * stage_bytes;
                uint8_t* sB = sA + TS_A_BYTES;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
                if (A_MN) {          // W[k rows, 32 features] boxes: one per 32-feature column block
#pragma unroll
                    for (int cb = 0; cb < TS_BM / 32; ++cb)
                        tma_load_2d_hint(sA + cb * (TS_BK * 128), &tmA, m0 + cb * 32, k0, &full[stage], a.w_policy);
                } else {             // W[128 feature rows, 32 k] boxes: one per 32-wide k block
#pragma unroll
                    for (int j = 0; j < TS_BK / 32; ++j)
                        tma_load_2d_hint(sA + j * (TS_BM * 128), &tmA, k0 + j * 32, m0, &full[stage], a.w_policy);
                }
#pragma unroll
                for (int j = 0; j < TS_BK / 32; ++j) {
                    tma_load_2d(sB + j * (a.Npad * 128), &tmB, k0 + j * 32, b0, &full[stage]);
                    if (a.B1)      // second source matrix: its rows land behind the first one's (B1 % 8 == 0)
                        tma_load_2d(sB + j * (a.Npad * 128) + a.B1 * 128, &tmB2, k0 + j * 32, 0, &full[stage]);
                }
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (elect_one()) {
            const uint32_t idesc = idesc_tf32(TS_BM, a.Npad, A_MN, false, false);
            int stage = 0; uint32_t phase = 0;
            int seg = 0;
            for (int cur = beg; cur < end; ++seg) {
                const int tile = cur / k_iters, kit0 = cur - tile * k_iters;
                const int n_it = min(end - cur, k_iters - kit0);
                const int buf = seg & 1;
                mbar_wait(&acc_empty[buf], ((seg >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * a.Npad);
                for (int i = 0; i < n_it; ++i) {
                    mbar_wait(&full[stage], phase);
                    if (cur == beg && i == 0) TS_MARK(2);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * stage_bytes);
                    const uint32_t sB = sA + TS_A_BYTES;
#pragma unroll
                    for (int g = 0; g < TS_BK / 8; ++g) {       // one MMA per 8 k (tf32 UMMA_K)
                        uint64_t ad, bd;
                        if (A_MN)   // k-group g = rows 8g..8g+7 (two 4-row swizzle atoms) of every column-block box
                            ad = smem_desc(sA + g * 1024, TS_BK * 128, 512, LAYOUT_SW128_BASE32B);
                        else        // K-major: 32-wide k block g/4, 32-byte step inside the swizzle row
                            ad = smem_desc(sA + (g / 4) * (TS_BM * 128) + (g % 4) * 32, 16, 1024, LAYOUT_SW128);
                        bd = smem_desc(sB + (g / 4) * (a.Npad * 128) + (g % 4) * 32, 16, 1024, LAYOUT_SW128);
                        mma_tf32(d_tmem, ad, bd, idesc, (i | g) != 0);
                    }
                    mma_commit(&empty[stage]);                  // smem slot free when these MMAs retire
                    if (++stage == a.stages) { stage = 0; phase ^= 1; }
                }
                mma_commit(&acc_full[buf]);
                cur += n_it;
            }
            TS_MARK(3);
        }
    } else {
        // ===================== epilogue: TMEM -> partial slab =====================
        const int quad = warp & 3;                               // TMEM lane quadrant this warp may read
        int seg = 0;
        for (int cur = beg; cur < end; ++seg) {
            const int tile = cur / k_iters, kit0 = cur - tile * k_iters;
            const int n_it = min(end - cur, k_iters - kit0);
            const int buf = seg & 1;
            const int slab = cta - sk_cta_of(a.sk, tile * k_iters);
            mbar_wait(&acc_full[buf], (seg >> 1) & 1);
            tc_fence_after();
            const int m = tile * TS_BM + quad * 32 + lane;
            float* dst = a.part + ((size_t)slab * a.B + b0) * a.M_total + m;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * a.Npad);
            for (int c0 = 0; c0 < a.Npad; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
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
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TS_MARK(6);
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
size_t tc_ws_bytes(const imdbn_ctx*, const imdbn_rbm*, int) { return 0; }

SKPlan tc_plan(const imdbn_ctx* ctx, int M_total, int K_total) {
    SKPlan p;
    const int m_tiles = (M_total + TS_BM - 1) / TS_BM;
    p.k_iters = (K_total + TS_BK - 1) / TS_BK;
    const int total = m_tiles * p.k_iters;
    // at least 4 k-iterations per CTA: fewer, longer ranges for small layers (fewer slabs to add)
    const int G = std::max(1, std::min(tc_sms(ctx), total / 4));
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

template <bool A_MN>
static int launch_stream(imdbn_ctx* ctx, const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmB2,
                         StreamArgs& a, int G, int chunks, cudaStream_t st) {
    const size_t smem = (size_t)a.stages * ts_stage_bytes(a.Npad) + 1024 + 256;
    static size_t smem_set = 0;          // the attribute is sticky: raise it only when a larger size is needed
    if (smem > smem_set) {
        IMDBN_CUDA(ctx, cudaFuncSetAttribute(k_tc_stream<A_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    IMDBN_CUDA(ctx, launch_pdl(k_tc_stream<A_MN>, dim3(G, chunks), dim3(TS_THREADS), smem, st, *tmA, *tmB, *tmB2, a));
    IMDBN_CHECK_LAUNCH(ctx, "k_tc_stream");
    return 0;
}

// act2 != nullptr: the batch is the virtual concatenation [act (B1 rows) ; act2 (B - B1 rows)]
static int stream_pass(imdbn_ctx* ctx, const imdbn_rbm* r, const float* act, int B, float* part, bool up,
                       cudaStream_t st, const float* act2 = nullptr, int B1 = 0) {
    const int M_total = up ? r->H : r->V, K_total = up ? r->V : r->H;
    if (!aligned16(act)) return fail(ctx, -1, "tc pass: activation pointer must be 16-byte aligned");
    StreamArgs a{};
    a.M_total = M_total; a.K_total = K_total; a.B = B; a.Npad = B > 256 ? 256 : npad_of(B);
    const int chunks = (B + a.Npad - 1) / a.Npad;
    a.sk = tc_plan(ctx, M_total, K_total);
    a.total_iters = ((M_total + TS_BM - 1) / TS_BM) * a.sk.k_iters;
    a.part = part;
    a.tmem_cols = pow2_cols(2 * a.Npad);
    a.w_policy = l2_policy_for(r);
    a.stages = std::max(2, std::min(TS_STAGES, (200 * 1024) / ts_stage_bytes(a.Npad)));
    const int G = tc_plan_ctas(a.sk, M_total);
    // W is [V, H] row-major: inner = H.  up: boxes [64 k-rows x 32 h]; down: boxes [128 v-rows x 32 h]
    const CUtensorMap* tmA = get_map(ctx, r->W, r->H, r->V, up ? TS_BK : TS_BM, up);
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
    // IMDBN_TS_TRACE=1: per-CTA globaltimer stamps of one large-layer launch (debug aid)
    static const bool trace_on = getenv("IMDBN_TS_TRACE") != nullptr;
    static unsigned long long* trace_buf = nullptr;
    static int big_calls = 0;
    const bool traced = trace_on && (size_t)r->V * r->H > (1u << 22) && ++big_calls >= 40 && big_calls <= 42;
    if (traced) {
        if (!trace_buf) cudaMalloc((void**)&trace_buf, 148 * 8 * 8);
        cudaMemsetAsync(trace_buf, 0, 148 * 8 * 8, st);
        a.trace = trace_buf;
    }
    int rc = up ? launch_stream<true>(ctx, tmA, tmB, tmB2, a, G, chunks, st)
                : launch_stream<false>(ctx, tmA, tmB, tmB2, a, G, chunks, st);
    if (traced && rc == 0) {
        static unsigned long long h[148 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int c = 0; c < G; ++c) t0 = std::min(t0, h[c * 8]);
        const char* names[7] = {"entry", "after pdl_wait", "first stage full", "last mma issued", "epilogue seg0", "epilogue seg1", "exit"};
        fprintf(stderr, "k_tc_stream<%s> B=%d G=%d trace (ns after first CTA entry; min / mean / max over CTAs)\n", up ? "up" : "down", B, G);
        for (int i = 0; i < 7; ++i) {
            double mn = 1e18, mx = 0, sum = 0; int n = 0;
            for (int c = 0; c < G; ++c) { if (!h[c * 8 + i]) continue; double v = (double)(h[c * 8 + i] - t0); mn = std::min(mn, v); mx = std::max(mx, v); sum += v; ++n; }
            if (n) fprintf(stderr, "  %-18s %8.0f %8.0f %8.0f  (n=%d)\n", names[i], mn, sum / n, mx, n);
        }
    }
    return rc;
}

int tc_gemm_up(imdbn_ctx* ctx, const imdbn_rbm* r, const float* v, int B, float* part, cudaStream_t st,
               const float* v2, int B1) {
    return stream_pass(ctx, r, v, B, part, true, st, v2, B1);
}
int tc_gemm_down(imdbn_ctx* ctx, const imdbn_rbm* r, const float* h, int B, float* part, cudaStream_t st) {
    return stream_pass(ctx, r, h, B, part, false, st);
}

int tc_gemm_stats(imdbn_ctx* ctx, const imdbn_rbm* r, const float* vp, const float* hp, const float* vn,
                  const float* hn, int B, float* dS_out, const imdbn_update* upd, cudaStream_t st) {
    if (!aligned16(vp) || !aligned16(hp) || !aligned16(vn) || !aligned16(hn) || (dS_out && !aligned16(dS_out)) ||
        (!dS_out && !aligned16(r->Wm)))
        return fail(ctx, -1, "tc stats: pointers must be 16-byte aligned");
    // tile shape: HBM-bound small batches stream W / W_m through 128 x 128 tiles and four IO slots; from 512 rows
    // the kernel is tensor-bound and uses 128 x 256 tiles (more MACs per operand byte), three operand stages
    const bool wide = B >= 512 && r->H >= 256 && getenv("IMDBN_STATS_NARROW") == nullptr;
    const int BN = wide ? 256 : 128;
    StatsArgs a{};
    a.V = r->V; a.H = r->H; a.B = B;
    a.m_tiles = (r->V + ST_BM - 1) / ST_BM;
    a.n_tiles = (r->H + BN - 1) / BN;
    a.k_chunks = (B + ST_KC - 1) / ST_KC;
    if (upd) { a.lr = upd->lr; a.mom = upd->momentum; a.wd = upd->weight_decay; a.bsz = (float)upd->batch_global; }
    { static const int dbg_env = getenv("IMDBN_DEBUG_STATS") ? atoi(getenv("IMDBN_DEBUG_STATS")) : 0; a.dbg = dbg_env; }
    a.late_wait = ctx->stats_after_colstats ? 1 : 0;
    ctx->stats_after_colstats = false;
    a.w_policy = l2_policy_for(r);
    a.wm_policy = l2_policy_for(r);
    const CUtensorMap* tVP = get_map(ctx, vp, r->V, B, ST_KC, true);
    const CUtensorMap* tVN = get_map(ctx, vn, r->V, B, ST_KC, true);
    const CUtensorMap* tHP = get_map(ctx, hp, r->H, B, ST_KC, true);
    const CUtensorMap* tHN = get_map(ctx, hn, r->H, B, ST_KC, true);
    const CUtensorMap* tW = get_map(ctx, dS_out ? dS_out : r->W, r->H, r->V, ST_BM, false);
    const CUtensorMap* tWm = dS_out ? tW : get_map(ctx, r->Wm, r->H, r->V, ST_BM, false);
    if (!tVP || !tVN || !tHP || !tHN || !tW || !tWm) return fail(ctx, -5, "cuTensorMapEncodeTiled failed");
    const int G = std::min(tc_sms(ctx), a.m_tiles * a.n_tiles);
    auto launch = [&](auto kernel, int smem, bool& attr_set) -> cudaError_t {
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            attr_set = true;
        }
        return launch_pdl(kernel, dim3(G), dim3(ST_THREADS), (size_t)smem, st, *tVP, *tVN, *tHP, *tHN, *tW, *tWm, a);
    };
    static bool set[4] = {false, false, false, false};
    cudaError_t e;
    if (wide)
        e = dS_out ? launch(k_tc_stats<false, 256, 3, 2>, st_smem<256, 3, 2>(), set[0])
                   : launch(k_tc_stats<true, 256, 3, 2>, st_smem<256, 3, 2>(), set[1]);
    else
        e = dS_out ? launch(k_tc_stats<false, 128, 2, 4>, st_smem<128, 2, 4>(), set[2])
                   : launch(k_tc_stats<true, 128, 2, 4>, st_smem<128, 2, 4>(), set[3]);
    IMDBN_CUDA(ctx, e);
    IMDBN_CHECK_LAUNCH(ctx, "k_tc_stats");
    return 0;
}

// ---- SM partition (CUDA green contexts): two streams whose kernels run on disjoint sets of SMs ---------------
namespace {
struct SmPartition { bool tried = false; bool ok = false; CUstream big = nullptr, small_ = nullptr; int n_big = 0, n_small = 0; };
SmPartition g_partitions[16];

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
    SmPartition& P = g_partitions[device];
    if (!P.tried) {
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
