// Context, workspace arena and launch bookkeeping shared by the kernels of libimdbn_b200.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/imdbn_b200.h"
#include "philox.cuh"

namespace imdbn {

constexpr int kNumSMsFallback = 148;  // B200

struct Arena {
    char* base = nullptr;
    size_t cap = 0;
    size_t off = 0;
};

}  // namespace imdbn

namespace imdbn {
struct ProfRec { int kind, V, H; cudaEvent_t a, b; };
struct ColstatsJob;
}

struct imdbn_ctx {
    bool profile = false;
    std::vector<imdbn::ProfRec> prof;
    int device = 0;
    int num_sms = imdbn::kNumSMsFallback;
    int tc_sms = 0;                   // > 0: persistent tensor-core kernels of this context use at most this many SMs
    int precision = IMDBN_PREC_FP32;
    int64_t launches = 0;
    std::string err;
    imdbn::Arena arena;
    // transposed-weight cache for the chain kernels is rebuilt on every call (W changes between
    // updates); it lives in the arena like everything else.
    void* tc = nullptr;  // tensor-core path state (tc_gemm.cu), opaque here
    bool w_stable = false;               // the next tensor-core pass may read W before its grid-dependency wait: no kernel
                                         // still in flight writes W (set by cd_core around the passes of one CD-k update)
    bool act_exact = false;              // the activations of the next tensor-core pass are sampled states (0 / 1): exactly
                                         // representable in tf32, the exact mode needs no remainder tile for them
    const imdbn::ColstatsJob* colstats_job = nullptr;   // set by finish_stats when the next statistics call will pack
    const uint32_t* act_hint = nullptr;  // the next pass's activations were scanned by the packing pass of this update:
    uint32_t act_hint_gen = 0;           // act_hint[0], act_hint[1] != act_hint_gen  <=>  exactly representable in tf32
    const float* pack_scan = nullptr;    // extra matrix the next packing pass scans for exactness ([pack_scan_rows, V])
    int pack_scan_rows = 0;
    bool stats_after_colstats = false;   // next tc statistics kernel directly follows k_colstats (see tc_stats.cuh)
    uint32_t* pack_flags = nullptr;      // device: [0] = generation of the last statistics call whose v operand was inexact in tf32
    uint32_t pack_gen = 0;
    unsigned int* ticket = nullptr;   // device counter of the last-block reductions (self-resetting)
    // imdbn_idbn_train_step with a second context: cross-stream events (created lazily)
    cudaEvent_t ev_ready = nullptr, ev_in = nullptr, ev_out = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr};
};

namespace imdbn {

// SMs the persistent tensor-core kernels may occupy (imdbn_set_sm_limit): leaves the rest of the chip to kernels
// of another stream, e.g. the upper layers of an iDBN running concurrently with the next layer-0 update.
inline int tc_sms(const imdbn_ctx* ctx) {
    return ctx->tc_sms > 0 && ctx->tc_sms < ctx->num_sms ? ctx->tc_sms : ctx->num_sms;
}

// tensor-core passes (tf32 or the split exact mode) / reduced-accuracy intrinsics in the finishes (tf32 only)
inline bool uses_tc(const imdbn_ctx* ctx) { return ctx->precision != IMDBN_PREC_FP32; }
inline bool fast_math(const imdbn_ctx* ctx) { return ctx->precision == IMDBN_PREC_TF32; }

inline int fail(imdbn_ctx* ctx, int code, const char* what) {
    if (ctx) {
        ctx->err = what;
        if (code > 0) {
            ctx->err += ": ";
            ctx->err += cudaGetErrorString((cudaError_t)code);
        }
    }
    return code;
}

#define IMDBN_CUDA(ctx, expr)                                             \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) return imdbn::fail((ctx), (int)_e, #expr); \
    } while (0)

#define IMDBN_CHECK_LAUNCH(ctx, name)                                         \
    do {                                                                      \
        (ctx)->launches++;                                                    \
        cudaError_t _e = cudaGetLastError();                                  \
        if (_e != cudaSuccess) return imdbn::fail((ctx), (int)_e, "launch " name); \
    } while (0)

#define IMDBN_ARG(ctx, cond)                                             \
    do {                                                                 \
        if (!(cond)) return imdbn::fail((ctx), -1, "invalid argument: " #cond); \
    } while (0)

// RAII bracket of one kernel launch with CUDA events on the launching stream (profile mode only).
struct ProfScope {
    imdbn_ctx* ctx; cudaStream_t st; ProfRec rec; bool on;
    ProfScope(imdbn_ctx* c, int kind, int V, int H, cudaStream_t s) : ctx(c), st(s), on(c->profile) {
        if (!on) return;
        rec.kind = kind; rec.V = V; rec.H = H;
        cudaEventCreate(&rec.a); cudaEventCreate(&rec.b);
        cudaEventRecord(rec.a, st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(rec.b, st);
        ctx->prof.push_back(rec);
    }
};

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// Every kernel launched through launch_pdl() begins with pdl_trigger() (its dependents may be
// scheduled as soon as all of its CTAs have started) and executes pdl_wait() before it touches
// global memory written by earlier kernels (blocks until those grids have completed and flushed).
// The on-chip prologue of kernel N+1 (barrier init, TMEM allocation, descriptor prefetch, block
// scheduling) thereby overlaps the tail of kernel N instead of following it.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Early launch of dependents can be switched off for a span of launches (pdl_early() = false): a dependent grid
// that is launched early parks its CTAs on whatever SMs are free, which defeats leaving SMs to another stream
// (imdbn_idbn_train_step with pipelined layers).
inline bool& pdl_early() { static thread_local bool on = true; return on; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_early() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Reserve `total` bytes for the current API call.  Growing synchronises the stream once (sizes
// settle after the first calls); buffers are only valid until the next API call on this context.
inline int arena_begin(imdbn_ctx* ctx, size_t total, cudaStream_t st) {
    total += 4096;
    if (total > ctx->arena.cap) {
        IMDBN_CUDA(ctx, cudaStreamSynchronize(st));
        if (ctx->arena.base) IMDBN_CUDA(ctx, cudaFree(ctx->arena.base));
        ctx->arena.base = nullptr;
        ctx->arena.cap = 0;
        size_t cap = total + total / 4;
        IMDBN_CUDA(ctx, cudaMalloc((void**)&ctx->arena.base, cap));
        ctx->arena.cap = cap;
    }
    ctx->arena.off = 0;
    return 0;
}

template <typename T>
inline T* arena_take(imdbn_ctx* ctx, size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(ctx->arena.base + ctx->arena.off);
    ctx->arena.off += bytes;
    return p;
}

inline size_t pad256(size_t n_floats) { return (n_floats * 4 + 255) & ~size_t(255); }

inline RngKey make_key(const imdbn_rng* r) {
    RngKey k;
    k.k0 = (uint32_t)(r->seed & 0xFFFFFFFFull);
    k.k1 = (uint32_t)(r->seed >> 32);
    k.stream = r->stream;
    k.row0 = r->row0;
    return k;
}

struct Groups {
    int n;
    int s[IMDBN_MAX_GROUPS];
    int e[IMDBN_MAX_GROUPS];
};

inline Groups make_groups(const imdbn_rbm* r) {
    Groups g;
    g.n = r->ngroups;
    for (int i = 0; i < IMDBN_MAX_GROUPS; ++i) {
        g.s[i] = i < r->ngroups ? r->group_start[i] : 0;
        g.e[i] = i < r->ngroups ? r->group_end[i] : 0;
    }
    return g;
}

struct BiasArgs {            // apply != 0: update the biases of the block's columns in the same kernel
    int apply;
    float* hb; float* hbm; float* vb; float* vbm;
    float lr, mom, bsz; int sparsity; float sp_target;
    float n_loss; float* loss_out;
};
// Column statistics of a CD update (k_colstats) handed to the tensor-core statistics path, which computes them in
// its operand-packing pass over the same four matrices (k_pack_colstats) instead of a kernel of their own.
struct ColstatsJob {
    const float* ea; const float* eb;      // squared error sum((ea - eb)^2) over [B,V]
    float* out;                            // [dh (H) | dv (V) | pos_h column sum (H) | squared error (1)]
    float* sq_part;                        // >= m_tiles + n_tiles floats
    unsigned int* ticket;
    BiasArgs ba;
};

// Stream-K partition of the tensor-core passes (tc_gemm.cu): the flattened (output tile, k-iteration)
// space of `total` iterations is cut into G contiguous, equally long ranges, one per CTA.  An output
// tile therefore receives partial sums from a few consecutive CTAs ("slabs"), which the finish
// kernels add in CTA order.  k_iters == 0 means "not stream-K": a uniform number of K splits.
struct SKPlan { int k_iters, q, r, tile_w; };
__host__ __device__ inline int sk_beg(const SKPlan& p, int c) { return c * p.q + (c < p.r ? c : p.r); }
__host__ __device__ inline int sk_cta_of(const SKPlan& p, int idx) {
    const int big = p.r * (p.q + 1);
    return idx < big ? idx / (p.q + 1) : p.r + (idx - big) / p.q;
}
__host__ __device__ inline int sk_nslabs(const SKPlan& p, int tile) {
    return sk_cta_of(p, (tile + 1) * p.k_iters - 1) - sk_cta_of(p, tile * p.k_iters) + 1;
}

// exact (never FMA-contracted) forms of the reference's element-wise expressions
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// rbm.py:19-21 hand-rolled sigmoid and torch.sigmoid agree to 1 ulp; one device form for both.
// x / b for a divisor shared by many elements (batch size, temperature, softmax sum): multiply by the
// correctly rounded reciprocal rb = 1/b and repair the quotient with one exact-remainder step.  Three
// FMA-pipe instructions instead of the ~30-instruction IEEE division sequence; the result is the correctly
// rounded quotient (identical to x / b) outside denormal / overflow corner cases.
__device__ __forceinline__ float div_by(float x, float b, float rb) {
    const float q = __fmul_rn(x, rb);
    const float e = __fmaf_rn(-q, b, x);
    return __fmaf_rn(e, rb, q);
}
__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// v = a*(1-km) + known*km  (rbm.py:291,365,397) without contraction
__device__ __forceinline__ float clampmix(float a, float known, float km) {
    return add_rn(mul_rn(a, 1.0f - km), mul_rn(known, km));
}

}  // namespace imdbn
