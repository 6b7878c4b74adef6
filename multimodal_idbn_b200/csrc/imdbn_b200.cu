// libimdbn_b200.so -- C ABI (include/imdbn_b200.h) over the sm_100a kernels.
// Host-side sequencing of the reference's RBM methods (imdbn/models/rbm.py); every function only
// enqueues kernels on the caller's stream.
#include <algorithm>
#include <ctime>
#include <vector>

#include "common.cuh"
#include "gemm_ffma.cuh"
#include "rbm_kernels.cuh"
#include "chain_kernel.cuh"
#include "chain_stepped.cuh"
#include "finish_vec.cuh"
#include "label_gibbs.cuh"
#include "cd_small.cuh"
#include "dp_update.cuh"
#include "energy_diag.cuh"
#include "tc_gemm.cuh"

using namespace imdbn;

namespace {

void prof_clear(imdbn_ctx* ctx) {
    for (auto& r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    ctx->prof.clear();
}

inline int ew_blocks(size_t n, int num_sms) {
    size_t b = (n + 255) / 256;
    size_t cap = (size_t)num_sms * 8;
    return (int)std::max<size_t>(1, std::min(b, cap));
}

struct PassPlan {
    int splits;          // uniform K splits (FFMA engine) or max slabs per tile (stream-K tensor-core pass)
    int kps;
    SKPlan sk;           // sk.k_iters != 0 -> tensor-core stream-K pass
    size_t part_floats;
};

// plan of the up pass (up = true: [B,V] x [V,H]) or the down pass ([B,H] x [H,V])
PassPlan plan_pass(const imdbn_ctx* ctx, const imdbn_rbm* r, int B, bool up) {
    const int M = B, N = up ? r->H : r->V, K = up ? r->V : r->H;
    PassPlan p{};
    const bool tc = uses_tc(ctx) &&
                    (up ? tc_up_supported(ctx, r, B) : tc_down_supported(ctx, r, B));
    if (tc) {
        p.sk = tc_plan(ctx, N, K, B);
        p.splits = tc_plan_max_slabs(p.sk, N);
        p.kps = 0;
    } else {
        p.splits = choose_splits(M, N, K, ctx->num_sms);
        int kps = (K + p.splits - 1) / p.splits;
        kps = (kps + GBK - 1) / GBK * GBK;
        p.kps = kps;
        p.splits = (K + kps - 1) / kps;
    }
    p.part_floats = (size_t)p.splits * M * N;
    return p;
}

// part[s][B][H] = v[B,V] . W[V,H]   (K split over blockIdx.z)
int gemm_up(imdbn_ctx* ctx, const imdbn_rbm* r, const float* v, int B, const PassPlan& pl,
            float* part, cudaStream_t st, const float* v2 = nullptr, int B1 = 0) {
    ProfScope prof(ctx, IMDBN_KERNEL_UP, r->V, r->H, st);
    if (pl.sk.k_iters) return tc_gemm_up(ctx, r, v, B, part, st, v2, B1);
    if (v2) return fail(ctx, -1, "two-source batches need the tensor-core path");
    GemmArgs g{};
    g.A = v; g.sAm = r->V; g.sAk = 1;
    g.B = r->W; g.sBk = r->H; g.sBn = 1;
    g.M = B; g.N = r->H; g.K = r->V;
    g.k_per_split = pl.kps; g.C = part;
    dim3 grid((g.N + GBN - 1) / GBN, (g.M + GBM - 1) / GBM, pl.splits);
    IMDBN_CUDA(ctx, launch_pdl(k_gemm_ffma<true, true, EPI_STORE>, grid, dim3(GTHREADS), 0, st, g));
    IMDBN_CHECK_LAUNCH(ctx, "k_gemm_ffma(up)");
    return 0;
}

// part[s][B][V] = h[B,H] . W[V,H]^T
int gemm_down(imdbn_ctx* ctx, const imdbn_rbm* r, const float* h, int B, const PassPlan& pl,
              float* part, cudaStream_t st) {
    ProfScope prof(ctx, IMDBN_KERNEL_DOWN, r->V, r->H, st);
    if (pl.sk.k_iters) return tc_gemm_down(ctx, r, h, B, part, st);
    GemmArgs g{};
    g.A = h; g.sAm = r->H; g.sAk = 1;
    g.B = r->W; g.sBk = 1; g.sBn = r->H;
    g.M = B; g.N = r->V; g.K = r->H;
    g.k_per_split = pl.kps; g.C = part;
    dim3 grid((g.N + GBN - 1) / GBN, (g.M + GBM - 1) / GBM, pl.splits);
    IMDBN_CUDA(ctx, launch_pdl(k_gemm_ffma<true, false, EPI_STORE>, grid, dim3(GTHREADS), 0, st, g));
    IMDBN_CHECK_LAUNCH(ctx, "k_gemm_ffma(down)");
    return 0;
}

// dS[V,H] = vp^T hp - vn^T hn, either stored (dS_out) or consumed by the fused update.
int gemm_stats(imdbn_ctx* ctx, const imdbn_rbm* r, const float* vp, const float* hp,
               const float* vn, const float* hn, int B, float* dS_out, const imdbn_update* upd,
               cudaStream_t st) {
    if (uses_tc(ctx) && tc_stats_supported(ctx, r, B))
        return tc_gemm_stats(ctx, r, vp, hp, vn, hn, B, dS_out, upd, st);     // (profiles its two kernels itself)
    ProfScope prof(ctx, IMDBN_KERNEL_STATS, r->V, r->H, st);
    GemmArgs g{};
    g.A = vp; g.sAm = 1; g.sAk = r->V;
    g.B = hp; g.sBk = r->H; g.sBn = 1;
    g.A2 = vn; g.B2 = hn; g.K2 = B;
    g.M = r->V; g.N = r->H; g.K = B;
    g.k_per_split = B;
    dim3 grid((g.N + GBN - 1) / GBN, (g.M + GBM - 1) / GBM, 1);
    if (dS_out) {
        g.C = dS_out;
        IMDBN_CUDA(ctx, launch_pdl(k_gemm_ffma<false, true, EPI_STORE>, grid, dim3(GTHREADS), 0, st, g));
    } else {
        g.W = r->W; g.Wm = r->Wm;
        g.lr = upd->lr; g.mom = upd->momentum; g.wd = upd->weight_decay;
        g.bsz = (float)upd->batch_global;
        IMDBN_CUDA(ctx, launch_pdl(k_gemm_ffma<false, true, EPI_UPDATE>, grid, dim3(GTHREADS), 0, st, g));
    }
    IMDBN_CHECK_LAUNCH(ctx, "k_gemm_ffma(stats)");
    return 0;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// the float4 finishes need N % 4 == 0, 16-byte aligned rows / outputs and a 32-bit quad index
inline bool vec_ok(int B, int N, const float* bias, const float* a, const float* b, const float* c) {
    return (N % 4) == 0 && al16(bias) && al16(a) && al16(b) && al16(c) && (size_t)B * (N / 4) < 0x7fffffffull;
}

int up_pass(imdbn_ctx* ctx, const imdbn_rbm* r, const float* v, int B, float T, float* p_out,
            float* s_out, const RngKey& key, uint32_t draw_u, const PassPlan& pl, float* part,
            cudaStream_t st, const float* v2 = nullptr, int B1 = 0) {
    if (pl.sk.k_iters && !v2 && tc_pass_fusable(ctx, r, B, true)) {       // large batch: the pass kernel finishes itself
        ProfScope prof(ctx, IMDBN_KERNEL_UP, r->V, r->H, st);
        const FusedFinish f{r->hb, 1.0f / fmaxf(1e-6f, T), p_out, s_out, key, draw_u};
        return tc_gemm_up(ctx, r, v, B, part, st, nullptr, 0, &f);
    }
    int rc = gemm_up(ctx, r, v, B, pl, part, st, v2, B1);
    if (rc) return rc;
    if (vec_ok(B, r->H, r->hb, p_out, s_out, nullptr)) {
        const size_t quads = (size_t)B * (r->H / 4);
        if (fast_math(ctx))
            IMDBN_CUDA(ctx, launch_pdl(k_finish_up4<true>, dim3(vec_blocks(quads, tc_sms(ctx))), dim3(256), 0, st,
                                       part, pl.splits, pl.sk, B, r->H, r->hb, fmaxf(1e-6f, T), 0.0f, p_out, s_out,
                                       key, draw_u, 0u));
        else
            IMDBN_CUDA(ctx, launch_pdl(k_finish_up4<false>, dim3(vec_blocks(quads, tc_sms(ctx))), dim3(256), 0, st,
                                       part, pl.splits, pl.sk, B, r->H, r->hb, fmaxf(1e-6f, T), 0.0f, p_out, s_out,
                                       key, draw_u, 0u));
    } else {
        IMDBN_CUDA(ctx, launch_pdl(k_finish_up, dim3((r->H + 255) / 256, std::min(B, 16384)), dim3(256), 0, st,
                                   part, pl.splits, pl.sk, B, r->H, r->hb, fmaxf(1e-6f, T), p_out, s_out, key, draw_u));
    }
    IMDBN_CHECK_LAUNCH(ctx, "k_finish_up");
    return 0;
}

int down_pass(imdbn_ctx* ctx, const imdbn_rbm* r, const float* h, int B, float T, float* p_out,
              float* logits_out, float* s_out, float* logits_tmp, const RngKey& key,
              uint32_t draw_u, uint32_t draw_cat, const PassPlan& pl, float* part,
              cudaStream_t st) {
    const Groups gr = make_groups(r);
    float* lg = logits_out ? logits_out : (gr.n ? logits_tmp : nullptr);
    if (pl.sk.k_iters && !lg && tc_pass_fusable(ctx, r, B, false)) {
        ProfScope prof(ctx, IMDBN_KERNEL_DOWN, r->V, r->H, st);
        const FusedFinish f{r->vb, 1.0f / fmaxf(1e-6f, T), p_out, s_out, key, draw_u};
        return tc_gemm_down(ctx, r, h, B, part, st, &f);
    }
    int rc = gemm_down(ctx, r, h, B, pl, part, st);
    if (rc) return rc;
    if (vec_ok(B, r->V, r->vb, p_out, s_out, lg)) {
        const size_t quads = (size_t)B * (r->V / 4);
        ChainPost4 cp{};
        if (fast_math(ctx))
            IMDBN_CUDA(ctx, launch_pdl(k_finish_down4<true>, dim3(vec_blocks(quads, tc_sms(ctx))), dim3(256), 0, st,
                                       part, pl.splits, pl.sk, B, r->V, r->vb, fmaxf(1e-6f, T), 0.0f, p_out, lg, s_out,
                                       key, draw_u, 0u, cp));
        else
            IMDBN_CUDA(ctx, launch_pdl(k_finish_down4<false>, dim3(vec_blocks(quads, tc_sms(ctx))), dim3(256), 0, st,
                                       part, pl.splits, pl.sk, B, r->V, r->vb, fmaxf(1e-6f, T), 0.0f, p_out, lg, s_out,
                                       key, draw_u, 0u, cp));
    } else {
        IMDBN_CUDA(ctx, launch_pdl(k_finish_down, dim3((r->V + 255) / 256, std::min(B, 16384)), dim3(256), 0, st,
                                   part, pl.splits, pl.sk, B, r->V, r->vb, fmaxf(1e-6f, T), p_out, lg, s_out, key, draw_u));
    }
    IMDBN_CHECK_LAUNCH(ctx, "k_finish_down");
    if (gr.n && (p_out || s_out)) {
        // the groups need a probability buffer even when the caller only wants samples
        float* pbuf = p_out ? p_out : logits_tmp + (size_t)B * r->V;
        const int warps = B * gr.n;
        k_groups<<<(warps * 32 + 127) / 128, 128, 0, st>>>(lg, pbuf, s_out, B, r->V, gr, key, draw_cat);
        IMDBN_CHECK_LAUNCH(ctx, "k_groups");
    }
    return 0;
}

int check_rbm(imdbn_ctx* ctx, const imdbn_rbm* r, bool need_momenta) {
    IMDBN_ARG(ctx, r != nullptr);
    IMDBN_ARG(ctx, r->W && r->hb && r->vb);
    IMDBN_ARG(ctx, r->V > 0 && r->H > 0);
    IMDBN_ARG(ctx, r->ngroups >= 0 && r->ngroups <= IMDBN_MAX_GROUPS);
    for (int g = 0; g < r->ngroups; ++g)
        IMDBN_ARG(ctx, r->group_start[g] >= 0 && r->group_start[g] < r->group_end[g] &&
                           r->group_end[g] <= r->V);
    if (need_momenta) IMDBN_ARG(ctx, r->Wm && r->hbm && r->vbm);
    return 0;
}

inline int colstat_blocks(const imdbn_rbm* r) { return (std::max(r->V, r->H) + CS_COLS - 1) / CS_COLS; }

// Column statistics into st_small = [dh | dv | sum pos_h | sq]; `apply` = also update the biases and
// write the loss (single-GPU path); otherwise only the squared-error total is finalised.
int finish_stats(imdbn_ctx* ctx, const imdbn_rbm* r, const float* hp, const float* hn,
                 const float* vp, const float* vn, const float* ea, const float* eb, int B,
                 float* st_small, float* sq_part, const imdbn_update* u, float* loss_out, bool apply,
                 cudaStream_t st) {
    const int nb = colstat_blocks(r);
    BiasArgs ba{};
    if (apply) {
        ba.apply = 1;
        ba.hb = r->hb; ba.hbm = r->hbm; ba.vb = r->vb; ba.vbm = r->vbm;
        ba.lr = u->lr; ba.mom = u->momentum; ba.bsz = (float)u->batch_global;
        ba.sparsity = u->sparsity; ba.sp_target = u->sparsity_target;
        ba.n_loss = (float)u->batch_global * r->V; ba.loss_out = loss_out;
    }
    static const bool no_fuse = getenv("IMDBN_NO_FUSED_COLSTATS") != nullptr;
    if (!no_fuse && tc_stats_packs(ctx, r, B) && (r->V % 4) == 0 && (r->H % 4) == 0 && al16(ea) && al16(eb)) {
        // the statistics GEMM that follows packs its operands in one pass over these same matrices: the column
        // statistics are computed there (k_pack_colstats); sq_part has colstat_blocks(r) >= tiles entries
        static thread_local ColstatsJob job;
        job.ea = ea; job.eb = eb; job.out = st_small; job.sq_part = sq_part; job.ticket = ctx->ticket; job.ba = ba;
        ctx->colstats_job = &job;
        return 0;
    }
    IMDBN_CUDA(ctx, launch_pdl(k_colstats, dim3(nb), dim3(CS_COLS * CS_ROWS), 0, st, hp, hn, vp, vn, ea, eb, B,
                               r->V, r->H, st_small, sq_part, ctx->ticket, ba));
    IMDBN_CHECK_LAUNCH(ctx, "k_colstats");
    return 0;
}

// bias update + loss from already reduced statistics (data-parallel path)
int bias_update(imdbn_ctx* ctx, const imdbn_rbm* r, float* st_small, const imdbn_update* u,
                float n_loss, float* loss_out, cudaStream_t st) {
    const int cols = std::max(r->V, r->H);
    k_bias_update<<<(cols + 255) / 256, 256, 0, st>>>(st_small, nullptr, 0, r->V, r->H, r->hb, r->hbm, r->vb,
                                                      r->vbm, u->lr, u->momentum,
                                                      (float)u->batch_global, u->sparsity,
                                                      u->sparsity_target, n_loss, loss_out, 1);
    IMDBN_CHECK_LAUNCH(ctx, "k_bias_update");
    return 0;
}

// ---- chain launcher -------------------------------------------------------------------------
template <int R, int NT, bool SLICED>
int launch_chain_r(imdbn_ctx* ctx, const ChainArgs& a, size_t smem, cudaStream_t st) {
    ProfScope prof(ctx, IMDBN_KERNEL_CHAIN, a.V, a.H, st);
    IMDBN_CUDA(ctx, cudaFuncSetAttribute(k_chain<R, NT, SLICED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    k_chain<R, NT, SLICED><<<(a.B + R - 1) / R, NT, smem, st>>>(a);
    IMDBN_CHECK_LAUNCH(ctx, "k_chain");
    return 0;
}

// IMG->TXT fast path (label_gibbs.cuh): mean-field conditional Gibbs, first Dz units clamped (caller's
// promise), the rest = one softmax group of <= 32 units.
bool chain_is_label_only(const imdbn_rbm* r, const imdbn_chain* ch, const float* vprob_out) {
    return ch->kind == IMDBN_CHAIN_COND_GIBBS && !ch->sample_h && !ch->sample_v && !ch->v_init &&
           r->ngroups == 1 && ch->clamp_prefix > 0 && ch->clamp_prefix == r->group_start[0] &&
           r->group_end[0] == r->V && r->V - ch->clamp_prefix <= 32 && (r->H % 32) == 0 && r->H <= LG_MAXH &&
           (ch->final_free_sweep || !vprob_out);
}

// Large mean-field batches in tf32 mode run the chain step by step on the tensor-core passes
// (chain_stepped.cuh); everything else uses the persistent kernel.
bool chain_is_stepped(const imdbn_ctx* ctx, const imdbn_rbm* r, const imdbn_chain* ch, int B) {
    return uses_tc(ctx) && B >= 512 && !ch->sample_h && !ch->sample_v &&
           tc_up_supported(ctx, r, B) && tc_down_supported(ctx, r, B);
}

size_t chain_ws_floats(const imdbn_ctx* ctx, const imdbn_rbm* r, const imdbn_chain* ch, int B) {
    size_t n = (size_t)r->V * r->H + 3 * (size_t)std::max(1, ch->n_steps) + 256;
    if (chain_is_label_only(r, ch, nullptr)) {
        const PassPlan pu = plan_pass(ctx, r, B, true), pd = plan_pass(ctx, r, B, false);
        n += std::max(pu.part_floats, pd.part_floats) + 2 * (size_t)B * r->H + 4 * (size_t)B * r->V + 4096;
    }
    if (chain_is_stepped(ctx, r, ch, B)) {
        const PassPlan pu = plan_pass(ctx, r, B, true), pd = plan_pass(ctx, r, B, false);
        n += std::max(pu.part_floats, pd.part_floats) + (size_t)B * r->H + 2 * (size_t)B * r->V + 1024;
    }
    return n;
}

int run_chain_stepped(imdbn_ctx* ctx, const imdbn_rbm* r, const imdbn_chain* ch, int B, float* v_out,
                      float* vprob_out, const RngKey& key, cudaStream_t st) {
    const int V = r->V, H = r->H, n = ch->n_steps;
    const bool noisy = ch->kind == IMDBN_CHAIN_NOISY_MF;
    IMDBN_ARG(ctx, !noisy || n == 0 || (ch->T && ch->sigma));
    const PassPlan pu = plan_pass(ctx, r, B, true), pd = plan_pass(ctx, r, B, false);
    float* part = arena_take<float>(ctx, std::max(pu.part_floats, pd.part_floats));
    float* h = arena_take<float>(ctx, (size_t)B * H);
    float* lg = arena_take<float>(ctx, (size_t)B * V);
    float* v = v_out;                                   // the state lives in the output buffer
    const dim3 gv((V + 255) / 256, std::min(B, 16384)), gh((H + 255) / 256, std::min(B, 16384));
    IMDBN_CUDA(ctx, launch_pdl(k_chain_init, gv, dim3(256), 0, st, ch->v_known, ch->known_mask, ch->v_init,
                               B, V, key, ch->draw0, v));
    ctx->launches++;
    const Groups gr = make_groups(r);
    bool vec = vec_ok(B, V, r->vb, v, ch->v_known, ch->known_mask) && vec_ok(B, H, r->hb, h, lg, vprob_out) &&
               (ch->Dz % 4) == 0;
    for (int g = 0; g < gr.n; ++g) vec = vec && (gr.s[g] % 4) == 0 && (gr.e[g] % 4) == 0;
    const int total = n + (ch->final_free_sweep ? 1 : 0);
    for (int t = 0; t < total; ++t) {
        const bool free_sweep = (t == n);
        const float T = (noisy && !free_sweep) ? fmaxf(1e-6f, ch->T[t]) : 1.0f;
        const float sig = (noisy && !free_sweep) ? ch->sigma[t] : 0.0f;
        const uint32_t d_h = ch->draw0 + 1 + (noisy ? 2 : 3) * t, d_v = d_h + 1;
        int rc = gemm_up(ctx, r, v, B, pu, part, st);                                 // rbm.py:344 / 394
        if (rc) return rc;
        if (vec && fast_math(ctx))
            IMDBN_CUDA(ctx, launch_pdl(k_finish_up4<true>, dim3(vec_blocks((size_t)B * (H / 4), tc_sms(ctx))), dim3(256),
                                       0, st, part, pu.splits, pu.sk, B, H, r->hb, T, sig, h, (float*)nullptr, key, 0u, d_h));
        else if (vec)
            IMDBN_CUDA(ctx, launch_pdl(k_finish_up4<false>, dim3(vec_blocks((size_t)B * (H / 4), tc_sms(ctx))), dim3(256),
                                       0, st, part, pu.splits, pu.sk, B, H, r->hb, T, sig, h, (float*)nullptr, key, 0u, d_h));
        else
            IMDBN_CUDA(ctx, launch_pdl(k_chain_up_finish, gh, dim3(256), 0, st, part, pu.splits, pu.sk, B, H, r->hb,
                                       T, sig, key, d_h, h));
        ctx->launches++;
        rc = gemm_down(ctx, r, h, B, pd, part, st);                                   // rbm.py:350 / 396
        if (rc) return rc;
        ChainPost po{};
        po.vk = ch->v_known; po.km = ch->known_mask;
        po.mu = (noisy && !free_sweep) ? ch->mu : nullptr; po.Dz = ch->Dz;
        po.eta = (noisy && ch->eta && !free_sweep) ? ch->eta[t] : 0.0f;
        po.free_sweep = free_sweep ? 1 : 0;
        po.vprob_out = (t == total - 1) ? vprob_out : nullptr;
        po.gr = gr;
        // block-mask promise (TXT->IMG): honoured on the vector path when no softmax group straddles the boundary
        po.clamp_from = -1;
        bool groups_clamped = false;
        if (vec && ch->clamp_suffix > 0 && ch->clamp_suffix < V && ch->clamp_suffix % 4 == 0) {
            int n_clamped = 0, n_free = 0, n_straddle = 0;
            for (int g = 0; g < gr.n; ++g) {
                if (gr.s[g] >= ch->clamp_suffix) ++n_clamped;
                else if (gr.e[g] <= ch->clamp_suffix) ++n_free;
                else ++n_straddle;
            }
            // all softmax groups on one side of the boundary (the group kernel must not read logits that the
            // finish skipped)
            if (n_straddle == 0 && (n_clamped == 0 || n_free == 0)) {
                po.clamp_from = ch->clamp_suffix;
                groups_clamped = n_free == 0 && !free_sweep && po.vprob_out == nullptr;
            }
        }
        if (vec) {
            ChainPost4 cp{}; cp.enabled = 1; cp.po = po;
            if (fast_math(ctx))
                IMDBN_CUDA(ctx, launch_pdl(k_finish_down4<true>, dim3(vec_blocks((size_t)B * (V / 4), tc_sms(ctx))), dim3(256),
                                           0, st, part, pd.splits, pd.sk, B, V, r->vb, T, sig, v, lg, (float*)nullptr, key, 0u,
                                           d_v, cp));
            else
                IMDBN_CUDA(ctx, launch_pdl(k_finish_down4<false>, dim3(vec_blocks((size_t)B * (V / 4), tc_sms(ctx))), dim3(256),
                                           0, st, part, pd.splits, pd.sk, B, V, r->vb, T, sig, v, lg, (float*)nullptr, key, 0u,
                                           d_v, cp));
        } else {
            IMDBN_CUDA(ctx, launch_pdl(k_chain_down_finish, gv, dim3(256), 0, st, part, pd.splits, pd.sk, B, V, r->vb,
                                       T, sig, key, d_v, po, lg, v));
        }
        ctx->launches++;
        if (gr.n && !groups_clamped) {             // (softmax groups inside the clamped block are overwritten anyway)
            IMDBN_CUDA(ctx, launch_pdl(k_chain_groups, dim3((B * gr.n * 32 + 127) / 128), dim3(128), 0, st, lg, B,
                                       V, po, v));
            ctx->launches++;
        }
    }
    return 0;
}

int run_chain_label_only(imdbn_ctx* ctx, const imdbn_rbm* r, const imdbn_chain* ch, int B, float* v_out,
                         float* vprob_out, const RngKey& key, cudaStream_t st) {
    const int V = r->V, H = r->H, Dz = ch->clamp_prefix, K = V - Dz;
    const PassPlan pu = plan_pass(ctx, r, B, true), pd = plan_pass(ctx, r, B, false);
    float* part = arena_take<float>(ctx, std::max(pu.part_floats, pd.part_floats));
    float* pre = arena_take<float>(ctx, (size_t)B * H);
    float* hbuf = arena_take<float>(ctx, (size_t)B * H);
    float* y = arena_take<float>(ctx, (size_t)B * K);
    float* vz = arena_take<float>(ctx, (size_t)B * V);
    float* lg_tmp = arena_take<float>(ctx, 2 * (size_t)B * V);
    const dim3 gv((V + 255) / 256, std::min(B, 16384)), gh((H + 255) / 256, std::min(B, 16384));
    // c = z W_z + b_h : one up-pass GEMM over [z | 0]
    IMDBN_CUDA(ctx, cudaMemsetAsync(y, 0, (size_t)B * K * sizeof(float), st));
    k_assemble_zy<<<gv, 256, 0, st>>>(ch->v_known, y, B, V, Dz, vz);
    IMDBN_CHECK_LAUNCH(ctx, "k_assemble_zy");
    int rc = gemm_up(ctx, r, vz, B, pu, part, st);
    if (rc) return rc;
    k_preact<<<gh, 256, 0, st>>>(part, pu.splits, pu.sk, B, H, r->hb, pre);
    IMDBN_CHECK_LAUNCH(ctx, "k_preact");
    LabelGibbsArgs a{};
    a.pre = pre; a.Wy = r->W + (size_t)Dz * H; a.vby = r->vb + Dz;
    a.B = B; a.H = H; a.K = K; a.Dz = Dz; a.n_steps = ch->n_steps;
    a.key = key; a.draw0 = ch->draw0; a.y_out = y;
    const size_t smem = ((size_t)64 * H + (size_t)LG_WARPS * LG_CHAINS * H) * sizeof(float);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        IMDBN_CUDA(ctx, cudaFuncSetAttribute(k_label_gibbs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IMDBN_CUDA(ctx, cudaFuncSetAttribute(k_label_gibbs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const int per_cta = LG_WARPS * LG_CHAINS;
    const int blocks = std::max(1, std::min((B + per_cta - 1) / per_cta, ctx->num_sms * 2));
    {
        ProfScope prof(ctx, IMDBN_KERNEL_CHAIN, V, H, st);
        if (fast_math(ctx)) k_label_gibbs<true><<<blocks, LG_WARPS * 32, smem, st>>>(a);
        else k_label_gibbs<false><<<blocks, LG_WARPS * 32, smem, st>>>(a);
        IMDBN_CHECK_LAUNCH(ctx, "k_label_gibbs");
    }
    if (!ch->final_free_sweep) {
        k_assemble_zy<<<gv, 256, 0, st>>>(ch->v_known, y, B, V, Dz, v_out);
        IMDBN_CHECK_LAUNCH(ctx, "k_assemble_zy");
        return 0;
    }
    // the un-clamped sweep the reference returns: visible_probs(forward([z | y]))        rbm.py:400
    k_assemble_zy<<<gv, 256, 0, st>>>(ch->v_known, y, B, V, Dz, vz);
    IMDBN_CHECK_LAUNCH(ctx, "k_assemble_zy");
    rc = up_pass(ctx, r, vz, B, 1.0f, hbuf, nullptr, key, 0, pu, part, st);
    if (rc) return rc;
    rc = down_pass(ctx, r, hbuf, B, 1.0f, v_out, nullptr, nullptr, lg_tmp, key, 0, 0, pd, part, st);
    if (rc) return rc;
    if (vprob_out)
        IMDBN_CUDA(ctx, cudaMemcpyAsync(vprob_out, v_out, (size_t)B * V * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// Wt / tables must come from the arena of the current call.
int run_chain(imdbn_ctx* ctx, const imdbn_rbm* r, const imdbn_chain* ch, int B, float* v_out,
              float* vprob_out, const RngKey& key, float* Wt, float* tables, cudaStream_t st) {
    IMDBN_ARG(ctx, ch->n_steps >= 0 && ch->n_steps <= CHAIN_MAX_STEPS);
    IMDBN_ARG(ctx, ch->v_known && ch->known_mask);
    if (chain_is_label_only(r, ch, vprob_out)) return run_chain_label_only(ctx, r, ch, B, v_out, vprob_out, key, st);
    if (!vprob_out && tc_chain_supported(ctx, r, ch, B)) {
        // TXT->IMG annealing of many chains: one persistent tensor-core kernel, state in shared memory (chain_tc.cuh)
        IMDBN_ARG(ctx, ch->T && ch->sigma);
        const int n = ch->n_steps;
        std::vector<float> host(3 * (size_t)n, 0.0f);
        for (int t = 0; t < n; ++t) {
            host[t] = fmaxf(1e-6f, ch->T[t]);
            host[n + t] = ch->sigma[t];
            host[2 * n + t] = ch->eta ? ch->eta[t] : 0.0f;
        }
        IMDBN_CUDA(ctx, cudaMemcpyAsync(tables, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        ProfScope prof(ctx, IMDBN_KERNEL_CHAIN, r->V, r->H, st);
        return tc_chain_t2i(ctx, r, ch, B, v_out, key, tables, tables + n, tables + 2 * n, st);
    }
    if (chain_is_stepped(ctx, r, ch, B)) return run_chain_stepped(ctx, r, ch, B, v_out, vprob_out, key, st);
    ChainArgs a{};
    a.W = r->W; a.Wt = Wt; a.hb = r->hb; a.vb = r->vb;
    a.V = r->V; a.H = r->H; a.B = B;
    a.gr = make_groups(r);
    a.kind = ch->kind; a.n_steps = ch->n_steps;
    a.v_known = ch->v_known; a.km = ch->known_mask; a.v_init = ch->v_init;
    a.mu = ch->mu; a.Dz = ch->Dz;
    a.sample_h = ch->sample_h; a.sample_v = ch->sample_v; a.final_free = ch->final_free_sweep;
    a.draw0 = ch->draw0;
    a.v_out = v_out; a.vprob_out = vprob_out;
    a.key = key;
    a.Vp = (r->V + 3) & ~3; a.Hp = (r->H + 3) & ~3;
    const int n = ch->n_steps;
    if (ch->kind == IMDBN_CHAIN_NOISY_MF) {
        IMDBN_ARG(ctx, n == 0 || (ch->T && ch->sigma));
        std::vector<float> host(3 * (size_t)std::max(1, n), 0.0f);
        for (int t = 0; t < n; ++t) {
            host[t] = fmaxf(1e-6f, ch->T[t]);
            host[n + t] = ch->sigma[t];
            host[2 * n + t] = ch->eta ? ch->eta[t] : 0.0f;
        }
        IMDBN_CUDA(ctx, cudaMemcpyAsync(tables, host.data(), host.size() * sizeof(float),
                                        cudaMemcpyHostToDevice, st));
        a.T = tables; a.sigma = tables + n; a.eta = tables + 2 * n;
    }
    dim3 tb(32, 8), tg((r->H + 31) / 32, (r->V + 31) / 32);
    k_transpose<<<tg, tb, 0, st>>>(r->W, r->V, r->H, Wt);
    IMDBN_CHECK_LAUNCH(ctx, "k_transpose");

    const size_t per_row = (size_t)(2 * a.Vp + a.Hp) * sizeof(float);
    int R = 16;
    while (R > 1 && ((B + R - 1) / R < 2 * ctx->num_sms || per_row * R > 100 * 1024)) R >>= 1;
    size_t smem = per_row * R;
    // few chains per CTA: k-sliced float4 products with 512 threads (see chain_matvec_sliced)
    const bool sliced = R <= 2 && (r->V % 4) == 0 && (r->H % 4) == 0 && al16(r->W) && al16(Wt);
    if (sliced) smem += (size_t)9 * R * std::max(a.Vp, a.Hp) * sizeof(float);
    if (smem > 227 * 1024) return fail(ctx, -2, "chain state does not fit in shared memory");
    if (sliced) return R == 2 ? launch_chain_r<2, 512, true>(ctx, a, smem, st)
                              : launch_chain_r<1, 512, true>(ctx, a, smem, st);
    switch (R) {
        case 16: return launch_chain_r<16, 256, false>(ctx, a, smem, st);
        case 8:  return launch_chain_r<8, 256, false>(ctx, a, smem, st);
        case 4:  return launch_chain_r<4, 256, false>(ctx, a, smem, st);
        case 2:  return launch_chain_r<2, 256, false>(ctx, a, smem, st);
        default: return launch_chain_r<1, 256, false>(ctx, a, smem, st);
    }
}

// shared body of cd_train / cd_stats
// Optional tail of a CD update (imdbn_cd_train_fwd): forward(data) with the UPDATED weights
// (idbn.py:203) and, in the same pass over W, the positive phase of the next minibatch.
struct FwdTail {
    const float* pos_h_in;     // nullable: positive probabilities of `data` from the previous call
    const float* next_data;    // nullable [B_next, V]
    int B_next;
    float* fwd_out;            // nullable [B + B_next, H]
};

// Small layers at small batch (the upper layers of an iDBN): the whole CD-k update + forward tail as ONE
// persistent kernel (cd_small.cuh) instead of ~14 launches.
static bool cd_small_eligible(const imdbn_ctx* ctx, const imdbn_rbm* r, const float* data, int B,
                              const FwdTail* tail) {
    // Measured on B200 (tools/bench_small.py, 1500 -> 500, batch 64): 63 us per call against 112 us for the
    // fp32 multi-launch path and 84 us (host-bound) for the tf32 one; inside the C2 step, where the launches
    // of the small layer are queued behind the 60 MB layer, the PDL-chained tcgen05 kernels are faster
    // (about 48 us), so tf32 mode keeps them unless IMDBN_CD_SMALL=1.
    static const bool off = getenv("IMDBN_NO_CD_SMALL") != nullptr;
    static const bool force = getenv("IMDBN_CD_SMALL") != nullptr;
    if (off || (ctx->precision != IMDBN_PREC_FP32 && !force)) return false;
    if (r->ngroups != 0 || B > CDS_T || r->V % 4 || r->H % 4) return false;
    if ((size_t)r->V * r->H > ((size_t)1 << 21)) return false;          // <= 8 MB of weights: L2-resident
    if (!al16(data) || !al16(r->W) || !al16(r->Wm) || !al16(r->hb) || !al16(r->vb)) return false;
    if (tail) {
        if (tail->pos_h_in && !al16(tail->pos_h_in)) return false;
        if (tail->next_data && (tail->B_next > CDS_T || tail->B_next <= 0 || !al16(tail->next_data))) return false;
        if (tail->fwd_out && !al16(tail->fwd_out)) return false;
    }
    return ctx->num_sms >= 8;
}

constexpr int CDS_FALLBACK = -1000;       // cd_small could not launch cooperatively: use the multi-launch path

static int cd_small(imdbn_ctx* ctx, const imdbn_rbm* r, const float* data, int B, int k,
                    const imdbn_update* upd, const imdbn_rng* rng, float* loss_out, cudaStream_t st,
                    const FwdTail* tail) {
    // every CTA must be resident for the grid barriers: one CTA per SM of the partition this context may use
    const int V = r->V, H = r->H, G = tc_sms(ctx);
    CdSmallArgs a{};
    a.W = r->W; a.Wm = r->Wm; a.hb = r->hb; a.hbm = r->hbm; a.vb = r->vb; a.vbm = r->vbm;
    a.V = V; a.H = H; a.data = data; a.B = B; a.k = k;
    if (tail) {
        a.pos_h_in = tail->pos_h_in;
        a.fwd_out = tail->fwd_out;
        if (tail->fwd_out && tail->next_data) { a.next_data = tail->next_data; a.B_next = tail->B_next; }
    }
    a.lr = upd->lr; a.mom = upd->momentum; a.wd = upd->weight_decay; a.bsz = (float)upd->batch_global;
    a.sparsity = upd->sparsity; a.sp_target = upd->sparsity_target;
    a.loss_out = loss_out;
    a.key = make_key(rng);
    a.up = cds_plan(H, V, G); a.dn = cds_plan(V, H, G); a.fw = a.up;
    const size_t nBH = (size_t)B * H, nBV = (size_t)B * V;
    const size_t part_floats = std::max({(size_t)a.up.ksplit * nBH, (size_t)a.dn.ksplit * nBV,
                                         (size_t)a.fw.ksplit * (size_t)(B + a.B_next) * H});
    int rc = arena_begin(ctx, pad256(part_floats) + 3 * pad256(nBH) + 2 * pad256(nBV) + pad256(G), st);
    if (rc) return rc;
    a.part = arena_take<float>(ctx, part_floats);
    a.pos_h = arena_take<float>(ctx, nBH);
    a.h_s = arena_take<float>(ctx, nBH);
    a.h_prob = arena_take<float>(ctx, nBH);
    a.v_prob = arena_take<float>(ctx, nBV);
    a.v_s = arena_take<float>(ctx, nBV);
    a.sq_part = arena_take<float>(ctx, G);
    a.bar = ctx->ticket + 8;
    static const bool trace_on = getenv("IMDBN_CDS_TRACE") != nullptr;
    static unsigned long long* trace_buf = nullptr;
    if (trace_on && !trace_buf) cudaMalloc((void**)&trace_buf, 64 * sizeof(unsigned long long));
    a.trace = trace_on ? trace_buf : nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        IMDBN_CUDA(ctx, cudaFuncSetAttribute(k_cd_small, cudaFuncAttributeMaxDynamicSharedMemorySize, CDS_SMEM_BYTES));
        attr_set = true;
    }
    ProfScope prof(ctx, IMDBN_KERNEL_STATS, V, H, st);
    {
        // The kernel synchronises its CTAs with grid barriers: a COOPERATIVE launch makes the driver guarantee that
        // all G CTAs are resident together (kernels of other streams holding SMs delay the launch instead of
        // dead-locking a partially resident grid); if the grid can never fit (SMs limited by the context, MPS)
        // the launch fails and the caller takes the multi-launch path.
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(G); cfg.blockDim = dim3(CDS_THREADS); cfg.dynamicSmemBytes = (size_t)CDS_SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, k_cd_small, a);
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources || e == cudaErrorNotSupported) {
            (void)cudaGetLastError();
            return CDS_FALLBACK;
        }
        IMDBN_CUDA(ctx, e);
    }
    IMDBN_CHECK_LAUNCH(ctx, "k_cd_small");
    if (trace_on) {
        static int calls = 0;
        if (++calls == 30) {
            unsigned long long h[64];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "cd_small phase trace (ns since start):");
            for (int i = 1; i < 24; ++i) fprintf(stderr, " %lld", (long long)(h[i] - h[0]));
            fprintf(stderr, "\n");
        }
    }
    return 0;
}

int cd_core(imdbn_ctx* ctx, const imdbn_rbm* r, const float* data, int B, int k,
            const imdbn_update* upd, const imdbn_rng* rng, float* loss_out, float* stats_out,
            cudaStream_t st, const FwdTail* tail = nullptr) {
    int rc = check_rbm(ctx, r, stats_out == nullptr);
    if (rc) return rc;
    IMDBN_ARG(ctx, data && B > 0 && k >= 1 && rng);
    if (!stats_out && upd && cd_small_eligible(ctx, r, data, B, tail)) {
        rc = cd_small(ctx, r, data, B, k, upd, rng, loss_out, st, tail);
        if (rc != CDS_FALLBACK) return rc;
    }
    const int V = r->V, H = r->H;
    const PassPlan pu = plan_pass(ctx, r, B, true), pd = plan_pass(ctx, r, B, false);
    const size_t nBH = (size_t)B * H, nBV = (size_t)B * V;
    const int nb_sq = colstat_blocks(r);
    const bool do_fwd = tail && tail->fwd_out;
    if (tail && tail->next_data)
        IMDBN_ARG(ctx, V % 4 == 0 && al16(data) && al16(tail->next_data));
    const int Bt = do_fwd ? B + (tail->next_data ? tail->B_next : 0) : 0;
    PassPlan pf{};
    if (do_fwd) pf = plan_pass(ctx, r, Bt, true);
    size_t bytes = pad256(std::max(pu.part_floats, pd.part_floats)) + 3 * pad256(nBH) +
                   4 * pad256(nBV) + pad256(2 * H + V + 1) + pad256(nb_sq) +
                   tc_ws_bytes(ctx, r, B) + (do_fwd ? pad256(pf.part_floats) + pad256((size_t)Bt * V) : 0);
    rc = arena_begin(ctx, bytes, st);
    if (rc) return rc;
    float* part = arena_take<float>(ctx, std::max(pu.part_floats, pd.part_floats));
    float* pos_h = arena_take<float>(ctx, nBH);
    float* h_s = arena_take<float>(ctx, nBH);
    float* h_prob = arena_take<float>(ctx, nBH);
    float* v_prob = arena_take<float>(ctx, nBV);
    float* v_s = arena_take<float>(ctx, nBV);
    float* lg_tmp = arena_take<float>(ctx, 2 * nBV);
    float* st_small = stats_out ? stats_out + (size_t)V * H : arena_take<float>(ctx, 2 * H + V + 1);
    float* sq_part = arena_take<float>(ctx, nb_sq);
    const RngKey key = make_key(rng);

    if (tail && tail->pos_h_in) {
        // positive phase already computed together with the previous call's forward pass
        pos_h = const_cast<float*>(tail->pos_h_in);
        if (vec_ok(B, H, pos_h, h_s, nullptr, nullptr))                                               // rbm.py:203
            IMDBN_CUDA(ctx, launch_pdl(k_bernoulli4, dim3(vec_blocks((size_t)B * (H / 4), tc_sms(ctx))), dim3(256), 0, st,
                                       (const float*)pos_h, B, H, h_s, key, 0u));
        else
            k_bernoulli<<<ew_blocks(nBH, ctx->num_sms), 256, 0, st>>>(pos_h, B, H, h_s, key, 0);
        IMDBN_CHECK_LAUNCH(ctx, "k_bernoulli");
    } else {
        rc = up_pass(ctx, r, data, B, 1.0f, pos_h, h_s, key, 0, pu, part, st);      // rbm.py:199,203
        if (rc) return rc;
    }
    // From here to the statistics kernel nothing writes W, and the kernel above is either a plain launch or a pass
    // that triggers only after its own wait: the passes of the CD loop may stream W before their predecessor is done.
    static const bool no_prefetch = getenv("IMDBN_NO_W_PREFETCH") != nullptr;
    ctx->w_stable = !no_prefetch;
    ctx->act_exact = true;                 // h_s and v_s are sampled states
    for (int s = 0; s < k; ++s) {
        rc = down_pass(ctx, r, h_s, B, 1.0f, v_prob, nullptr, v_s, lg_tmp, key, 1 + 3 * s,
                       2 + 3 * s, pd, part, st);                                        // :205-206
        if (rc) break;
        rc = up_pass(ctx, r, v_s, B, 1.0f, h_prob, (s + 1 < k) ? h_s : nullptr, key, 3 + 3 * s, pu,
                     part, st);                                                         // :207-208
        if (rc) break;
    }
    ctx->w_stable = false;
    ctx->act_exact = false;
    if (rc) return rc;
    // The bias / loss kernel and the weight-statistics kernel share only read-only inputs (the biases are
    // not read by the statistics GEMM, W is not read by the column statistics), so their order does not
    // matter; the column statistics go first and the tensor-core statistics kernel overlaps them.
    ctx->act_hint = nullptr;
    if (do_fwd && tail->next_data) {       // (exact mode) let the packing pass also scan the next minibatch
        ctx->pack_scan = tail->next_data; ctx->pack_scan_rows = tail->B_next;
    }
    rc = finish_stats(ctx, r, pos_h, h_prob, data, v_s, data, v_prob, B, st_small, sq_part, upd, loss_out,
                      stats_out == nullptr, st);                                        // :216-226
    if (rc) return rc;
    ctx->stats_after_colstats = true;
    rc = gemm_stats(ctx, r, data, pos_h, v_s, h_prob, B, stats_out, upd, st);           // :200,209,212
    ctx->stats_after_colstats = false;
    ctx->pack_scan = nullptr; ctx->colstats_job = nullptr;
    // ctx->act_hint (set by the packing pass when it scanned [data ; next_data]) applies to the forward pass below only
    struct HintGuard { imdbn_ctx* c; ~HintGuard() { c->act_hint = nullptr; } } hint_guard{ctx};
    if (rc || !do_fwd) return rc;
    // one pass over the updated W for [data ; next_data]
    const float* src = data;
    if (Bt > B && pf.sk.k_iters && B % 8 == 0 && Bt <= 256) {
        // tensor-core path: the TMA producer reads [data ; next_data] from the two matrices directly
        float* part_f = arena_take<float>(ctx, pf.part_floats);
        return up_pass(ctx, r, data, Bt, 1.0f, tail->fwd_out, nullptr, key, 0, pf, part_f, st, tail->next_data, B);
    }
    if (Bt > B) {
        float* cat = arena_take<float>(ctx, (size_t)Bt * V);
        const size_t n1 = nBV / 4, n2 = (size_t)tail->B_next * V / 4;      // V % 4 == 0 on this path
        k_concat2<<<ew_blocks(n1 + n2, ctx->num_sms), 256, 0, st>>>(
            reinterpret_cast<const float4*>(data), n1, reinterpret_cast<const float4*>(tail->next_data), n2,
            reinterpret_cast<float4*>(cat));
        IMDBN_CHECK_LAUNCH(ctx, "k_concat2");
        src = cat;
    }
    float* part_f = arena_take<float>(ctx, pf.part_floats);
    return up_pass(ctx, r, src, Bt, 1.0f, tail->fwd_out, nullptr, key, 0, pf, part_f, st);
}

int clamped_core(imdbn_ctx* ctx, const imdbn_rbm* r, const float* v_known, const float* km, int B,
                 const imdbn_clamped_cfg* cfg, const imdbn_update* upd, const imdbn_rng* rng,
                 float* loss_out, float* stats_out, cudaStream_t st) {
    int rc = check_rbm(ctx, r, stats_out == nullptr);
    if (rc) return rc;
    IMDBN_ARG(ctx, v_known && km && B > 0 && cfg && rng && cfg->k >= 0);
    const int V = r->V, H = r->H;
    const int n = cfg->use_noisy_init ? std::max(10, cfg->cond_init_steps) : cfg->cond_init_steps;
    IMDBN_ARG(ctx, n >= 0 && n <= CHAIN_MAX_STEPS);
    const PassPlan pu = plan_pass(ctx, r, B, true), pd = plan_pass(ctx, r, B, false);
    const size_t nBH = (size_t)B * H, nBV = (size_t)B * V;
    const int nb_sq = colstat_blocks(r);
    imdbn_chain chq{};                       // only what the workspace sizing looks at
    chq.n_steps = n;
    chq.clamp_prefix = chq.clamp_suffix = -1;
    chq.sample_h = cfg->use_noisy_init ? 0 : cfg->sample_h;
    chq.sample_v = cfg->use_noisy_init ? 0 : cfg->sample_v;
    size_t bytes = pad256(std::max(pu.part_floats, pd.part_floats)) + 3 * pad256(nBH) +
                   5 * pad256(nBV) + pad256(2 * H + V + 1) + pad256(nb_sq) +
                   pad256(chain_ws_floats(ctx, r, &chq, B)) + 4096 + tc_ws_bytes(ctx, r, B);
    rc = arena_begin(ctx, bytes, st);
    if (rc) return rc;
    float* part = arena_take<float>(ctx, std::max(pu.part_floats, pd.part_floats));
    float* h_plus = arena_take<float>(ctx, nBH);
    float* h_cur = arena_take<float>(ctx, nBH);
    float* h_neg = arena_take<float>(ctx, nBH);
    float* v_plus = arena_take<float>(ctx, nBV);
    float* v_prob = arena_take<float>(ctx, nBV);
    float* v_neg = arena_take<float>(ctx, nBV);
    float* lg_tmp = arena_take<float>(ctx, 2 * nBV);
    float* st_small = stats_out ? stats_out + (size_t)V * H : arena_take<float>(ctx, 2 * H + V + 1);
    float* sq_part = arena_take<float>(ctx, nb_sq);
    float* Wt = arena_take<float>(ctx, (size_t)V * H);
    float* tables = arena_take<float>(ctx, 3 * (size_t)std::max(1, n));
    const RngKey key = make_key(rng);

    // positive phase: conditional inference (rbm.py:443-453)
    imdbn_chain ch{};
    std::vector<float> T, S, E;
    ch.v_known = v_known; ch.known_mask = km; ch.n_steps = n; ch.draw0 = 0;
    ch.clamp_prefix = ch.clamp_suffix = -1;      // arbitrary mask: no block-mask promise (a zero-initialised 0 would
                                                 // read as "every column clamped" in the stepped chain)
    uint32_t base;
    if (cfg->use_noisy_init) {
        // T0=3, T1=1, sigma0=.9, sharpen_last=2, T_cold_plus=.9 (rbm.py:444-448, schedule :229-234,338-341)
        T.resize(n); S.resize(n); E.assign(n, 0.0f);
        for (int t = 0; t < n; ++t) {
            double Tt = 1.0;
            if (n > 1) {
                double al = std::min(std::max((double)t / (n - 1), 0.0), 1.0);
                Tt = 3.0 + (1.0 - 3.0) * al;
            }
            if (n - t <= 2) Tt = 0.9;
            T[t] = (float)std::max(1e-6, Tt);
            S[t] = (float)(0.9 * std::max(0.0, 1.0 - (double)t / std::max(1, n - 1)));
        }
        ch.kind = IMDBN_CHAIN_NOISY_MF; ch.T = T.data(); ch.sigma = S.data(); ch.eta = E.data();
        base = 1 + 2 * n;
    } else {
        ch.kind = IMDBN_CHAIN_COND_GIBBS; ch.sample_h = cfg->sample_h; ch.sample_v = cfg->sample_v;
        ch.final_free_sweep = 1;
        base = 1 + 3 * n;
    }
    rc = run_chain(ctx, r, &ch, B, v_plus, nullptr, key, Wt, tables, st);
    if (rc) return rc;
    rc = up_pass(ctx, r, v_plus, B, 1.0f, h_plus, nullptr, key, 0, pu, part, st);       // :455
    if (rc) return rc;
    const float* vcur = v_plus;
    for (int s = 0; s < cfg->k; ++s) {                                                  // :460-469
        rc = up_pass(ctx, r, vcur, B, 1.0f, cfg->sample_h ? nullptr : h_cur,
                     cfg->sample_h ? h_cur : nullptr, key, base + 3 * s, pu, part, st);
        if (rc) return rc;
        rc = down_pass(ctx, r, h_cur, B, 1.0f, v_prob, nullptr, nullptr, lg_tmp, key, 0, 0, pd,
                       part, st);
        if (rc) return rc;
        const float* vn = v_prob;
        if (cfg->reclamp_negative) {
            k_clampmix<<<ew_blocks(nBV, ctx->num_sms), 256, 0, st>>>(v_prob, v_known, km, nBV, v_neg);
            IMDBN_CHECK_LAUNCH(ctx, "k_clampmix");
            vn = v_neg;
        }
        if (cfg->sample_v) {
            // sample_visible(v_neg) (rbm.py:469): Bernoulli everywhere, categorical per group
            float* dst = (vn == v_neg) ? v_prob : v_neg;
            IMDBN_CUDA(ctx, cudaMemcpyAsync(lg_tmp, vn, nBV * sizeof(float), cudaMemcpyDeviceToDevice, st));
            k_bernoulli<<<ew_blocks(nBV, ctx->num_sms), 256, 0, st>>>(vn, B, V, dst, key, base + 3 * s + 1);
            IMDBN_CHECK_LAUNCH(ctx, "k_bernoulli");
            const Groups gr = make_groups(r);
            if (gr.n) {
                k_groups<<<(B * gr.n * 32 + 127) / 128, 128, 0, st>>>(nullptr, lg_tmp, dst, B, V, gr, key,
                                                                      base + 3 * s + 2);
                IMDBN_CHECK_LAUNCH(ctx, "k_groups");
            }
            vn = dst;
        }
        if (vn != v_neg) {
            IMDBN_CUDA(ctx, cudaMemcpyAsync(v_neg, vn, nBV * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
        vcur = v_neg;
    }
    if (cfg->k == 0)
        IMDBN_CUDA(ctx, cudaMemcpyAsync(v_neg, v_plus, nBV * sizeof(float), cudaMemcpyDeviceToDevice, st));
    rc = up_pass(ctx, r, v_neg, B, 1.0f, h_neg, nullptr, key, 0, pu, part, st);         // :471
    if (rc) return rc;
    imdbn_update u{};
    if (upd) { u = *upd; u.sparsity = 0; }                                              // :478-481
    rc = finish_stats(ctx, r, h_plus, h_neg, v_plus, v_neg, v_plus, v_neg, B, st_small, sq_part, &u,
                      loss_out, stats_out == nullptr, st);
    if (rc) return rc;
    ctx->stats_after_colstats = true;
    rc = gemm_stats(ctx, r, v_plus, h_plus, v_neg, h_neg, B, stats_out, upd, st);       // :456,472,476
    ctx->stats_after_colstats = false;
    return rc;
}

}  // namespace

// =============================================================================================
extern "C" {

int imdbn_abi_version(void) { return IMDBN_ABI_VERSION; }

int imdbn_ctx_create(imdbn_ctx** out, int device) {
    if (!out) return -1;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return (int)e;
    if (device < 0 || device >= ndev) return -1;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    imdbn_ctx* c = new imdbn_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        c->num_sms = prop.multiProcessorCount;
        if (prop.major < 10) {
            delete c;
            return -3;  // sm_100a code only
        }
    }
    if (cudaMalloc((void**)&c->ticket, 256) != cudaSuccess || cudaMemset(c->ticket, 0, 256) != cudaSuccess) {
        delete c;
        return (int)cudaErrorMemoryAllocation;
    }
    *out = c;
    return 0;
}

void imdbn_ctx_destroy(imdbn_ctx* ctx) {
    if (!ctx) return;
    tc_destroy(ctx);
    prof_clear(ctx);
    if (ctx->ticket) cudaFree(ctx->ticket);
    if (ctx->pack_flags) cudaFree(ctx->pack_flags);
    if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
    if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
    if (ctx->ev_out) cudaEventDestroy(ctx->ev_out);
    for (int i = 0; i < 2; ++i)
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    if (ctx->arena.base) cudaFree(ctx->arena.base);
    delete ctx;
}

const char* imdbn_last_error(imdbn_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int imdbn_sm_partition(int device, int small_sms, void** stream_big, void** stream_small, int* n_big, int* n_small) {
    return sm_partition(device, small_sms, stream_big, stream_small, n_big, n_small);
}

int imdbn_set_sm_limit(imdbn_ctx* ctx, int n_sms) {
    IMDBN_ARG(ctx, ctx && n_sms >= 0);
    ctx->tc_sms = n_sms;
    return 0;
}

int imdbn_set_precision(imdbn_ctx* ctx, int prec) {
    IMDBN_ARG(ctx, ctx && (prec == IMDBN_PREC_FP32 || prec == IMDBN_PREC_TF32 || prec == IMDBN_PREC_TF32X2));
    ctx->precision = prec;
    return 0;
}

int64_t imdbn_launch_count(imdbn_ctx* ctx) { return ctx ? ctx->launches : 0; }

int imdbn_copy_async(void* dst, const void* src, size_t bytes, imdbn_stream stream) {
    if (!dst || !src) return -1;
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
}

int imdbn_profile_enable(imdbn_ctx* ctx, int enable) {
    IMDBN_ARG(ctx, ctx != nullptr);
    prof_clear(ctx);
    ctx->profile = enable != 0;
    return 0;
}

int imdbn_profile_read(imdbn_ctx* ctx, int kind, int V, int H, double* ms_sum, int64_t* count) {
    IMDBN_ARG(ctx, ctx && ms_sum && count);
    double sum = 0.0;
    int64_t n = 0;
    for (auto& r : ctx->prof) {
        if (r.kind != kind || r.V != V || r.H != H) continue;
        IMDBN_CUDA(ctx, cudaEventSynchronize(r.b));
        float ms = 0.f;
        IMDBN_CUDA(ctx, cudaEventElapsedTime(&ms, r.a, r.b));
        sum += ms; ++n;
    }
    *ms_sum = sum; *count = n;
    return 0;
}

int imdbn_up(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* v, int B, float T, float* p_out,
             float* s_out, const imdbn_rng* rng, uint32_t draw_u, imdbn_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rbm(ctx, rbm, false);
    if (rc) return rc;
    IMDBN_ARG(ctx, v && B > 0 && (p_out || s_out) && (!s_out || rng));
    const PassPlan pu = plan_pass(ctx, rbm, B, true);
    rc = arena_begin(ctx, pad256(pu.part_floats) + tc_ws_bytes(ctx, rbm, B), st);
    if (rc) return rc;
    float* part = arena_take<float>(ctx, pu.part_floats);
    RngKey key{};
    if (rng) key = make_key(rng);
    return up_pass(ctx, rbm, v, B, T, p_out, s_out, key, draw_u, pu, part, st);
}

int imdbn_down(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* h, int B, float T, float* p_out,
               float* logits_out, float* s_out, const imdbn_rng* rng, uint32_t draw_u,
               uint32_t draw_cat, imdbn_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rbm(ctx, rbm, false);
    if (rc) return rc;
    IMDBN_ARG(ctx, h && B > 0 && (p_out || s_out || logits_out) && (!s_out || rng));
    const PassPlan pd = plan_pass(ctx, rbm, B, false);
    const size_t nBV = (size_t)B * rbm->V;
    rc = arena_begin(ctx, pad256(pd.part_floats) + pad256(2 * nBV) + tc_ws_bytes(ctx, rbm, B), st);
    if (rc) return rc;
    float* part = arena_take<float>(ctx, pd.part_floats);
    float* lg_tmp = arena_take<float>(ctx, 2 * nBV);
    RngKey key{};
    if (rng) key = make_key(rng);
    return down_pass(ctx, rbm, h, B, T, p_out, logits_out, s_out, lg_tmp, key, draw_u, draw_cat, pd,
                     part, st);
}

int imdbn_sample_visible(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* p, int B, float* s_out,
                         const imdbn_rng* rng, uint32_t draw_u, uint32_t draw_cat,
                         imdbn_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rbm(ctx, rbm, false);
    if (rc) return rc;
    IMDBN_ARG(ctx, p && s_out && rng && B > 0 && p != s_out);
    const size_t nBV = (size_t)B * rbm->V;
    const RngKey key = make_key(rng);
    rc = arena_begin(ctx, pad256(nBV), st);
    if (rc) return rc;
    float* ptmp = arena_take<float>(ctx, nBV);
    k_bernoulli<<<ew_blocks(nBV, ctx->num_sms), 256, 0, st>>>(p, B, rbm->V, s_out, key, draw_u);
    IMDBN_CHECK_LAUNCH(ctx, "k_bernoulli");
    const Groups gr = make_groups(rbm);
    if (gr.n) {
        IMDBN_CUDA(ctx, cudaMemcpyAsync(ptmp, p, nBV * sizeof(float), cudaMemcpyDeviceToDevice, st));
        k_groups<<<(B * gr.n * 32 + 127) / 128, 128, 0, st>>>(nullptr, ptmp, s_out, B, rbm->V, gr, key,
                                                              draw_cat);
        IMDBN_CHECK_LAUNCH(ctx, "k_groups");
    }
    return 0;
}

int imdbn_free_energy(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* v, int B, float* F_out,
                      imdbn_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rbm(ctx, rbm, false);
    if (rc) return rc;
    IMDBN_ARG(ctx, v && F_out && B > 0);
    const PassPlan pu = plan_pass(ctx, rbm, B, true);
    rc = arena_begin(ctx, pad256(pu.part_floats) + tc_ws_bytes(ctx, rbm, B), st);
    if (rc) return rc;
    float* part = arena_take<float>(ctx, pu.part_floats);
    rc = gemm_up(ctx, rbm, v, B, pu, part, st);
    if (rc) return rc;
    k_free_energy<<<B, 256, 0, st>>>(part, pu.splits, pu.sk, v, B, rbm->V, rbm->H, rbm->hb, rbm->vb, F_out);
    IMDBN_CHECK_LAUNCH(ctx, "k_free_energy");
    return 0;
}

int imdbn_cd_train(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* data, int B, int k,
                   const imdbn_update* upd, const imdbn_rng* rng, float* loss_out,
                   imdbn_stream stream) {
    IMDBN_ARG(ctx, upd && upd->batch_global > 0);
    return cd_core(ctx, rbm, data, B, k, upd, rng, loss_out, nullptr, (cudaStream_t)stream);
}

int imdbn_cd_train_fwd(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* data, int B, int k,
                       const imdbn_update* upd, const imdbn_rng* rng, float* loss_out, const float* pos_h_in,
                       const float* next_data, int B_next, float* fwd_out, imdbn_stream stream) {
    IMDBN_ARG(ctx, upd && upd->batch_global > 0 && fwd_out && (!next_data || B_next > 0));
    FwdTail t{pos_h_in, next_data, next_data ? B_next : 0, fwd_out};
    return cd_core(ctx, rbm, data, B, k, upd, rng, loss_out, nullptr, (cudaStream_t)stream, &t);
}

// One minibatch of iDBN.train for ALL layers in one call (idbn.py:199-204): per layer CD-k update + forward of
// the updated layer, which is the next layer's input.  With a second context / stream the upper layers run
// there, concurrently with whatever the caller enqueues next on the first stream (the next minibatch's layer 0).
int imdbn_idbn_train_step(imdbn_ctx* ctx0, imdbn_ctx* ctx1, int n_layers, const imdbn_rbm* rbms,
                          const imdbn_update* upds, const imdbn_rng* rngs, const float* data, int B, int k,
                          const float* pos_h_in, const float* next_data, int B_next, float* const* fwd_out,
                          float* const* loss_out, int buffer_set, imdbn_stream stream0, imdbn_stream stream1,
                          imdbn_stream caller_stream, int early_launch) {
    IMDBN_ARG(ctx0, ctx0 && n_layers >= 1 && rbms && upds && rngs && data && fwd_out && loss_out && B > 0);
    IMDBN_ARG(ctx0, buffer_set == 0 || buffer_set == 1);
    cudaStream_t s0 = (cudaStream_t)stream0, s1 = (cudaStream_t)stream1, sc = (cudaStream_t)caller_stream;
    static const bool host_trace = getenv("IMDBN_HOST_TRACE") != nullptr;
    static double acc_t[4] = {0, 0, 0, 0};
    static long acc_n = 0;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    const double t_begin = host_trace ? now() : 0;
    const bool foreign = caller_stream != nullptr && sc != s0;      // layer 0 runs on a stream other than the caller's
    if (foreign) {
        if (!ctx0->ev_in) {
            IMDBN_CUDA(ctx0, cudaEventCreateWithFlags(&ctx0->ev_in, cudaEventDisableTiming));
            IMDBN_CUDA(ctx0, cudaEventCreateWithFlags(&ctx0->ev_out, cudaEventDisableTiming));
        }
        IMDBN_CUDA(ctx0, cudaEventRecord(ctx0->ev_in, sc));          // inputs produced on the caller's stream
        IMDBN_CUDA(ctx0, cudaStreamWaitEvent(s0, ctx0->ev_in, 0));
    }
    const bool piped = ctx1 != nullptr && n_layers > 1;
    const int par = buffer_set;
    if (piped) {
        if (!ctx0->ev_ready) {
            IMDBN_CUDA(ctx0, cudaEventCreateWithFlags(&ctx0->ev_ready, cudaEventDisableTiming));
            for (int i = 0; i < 2; ++i)
                IMDBN_CUDA(ctx0, cudaEventCreateWithFlags(&ctx0->ev_done[i], cudaEventDisableTiming));
        } else {
            // the caller alternates between two sets of fwd_out buffers: wait for the upper layers that last read
            // the set written now (an event that was never recorded is complete)
            IMDBN_CUDA(ctx0, cudaStreamWaitEvent(s0, ctx0->ev_done[par], 0));
        }
    }
    for (int l = 0; l < n_layers; ++l) IMDBN_ARG(ctx0, fwd_out[l] && upds[l].batch_global > 0);
    FwdTail t0{pos_h_in, next_data, next_data ? B_next : 0, fwd_out[0]};
    // Without an SM partition an early-launched dependent grid parks its CTAs on the SMs that were meant for the
    // other stream: the caller then asks for plain stream-ordered launches of layer 0.
    if (piped && !early_launch) pdl_early() = false;
    const double t_l0 = host_trace ? now() : 0;
    int rc = cd_core(ctx0, &rbms[0], data, B, k, &upds[0], &rngs[0], loss_out[0], nullptr, s0, &t0);
    pdl_early() = true;
    if (rc) return rc;
    const double t_l0_end = host_trace ? now() : 0;
    if (foreign) {                                                    // the caller's stream sees layer 0's results
        IMDBN_CUDA(ctx0, cudaEventRecord(ctx0->ev_out, s0));
        IMDBN_CUDA(ctx0, cudaStreamWaitEvent(sc, ctx0->ev_out, 0));
    }
    imdbn_ctx* cx = ctx0;
    cudaStream_t sx = s0;
    if (piped) {
        IMDBN_CUDA(ctx0, cudaEventRecord(ctx0->ev_ready, s0));
        IMDBN_CUDA(ctx0, cudaStreamWaitEvent(s1, ctx0->ev_ready, 0));
        cx = ctx1; sx = s1;
        cx->precision = ctx0->precision;
    }
    for (int l = 1; l < n_layers; ++l) {
        FwdTail t{nullptr, nullptr, 0, fwd_out[l]};
        rc = cd_core(cx, &rbms[l], fwd_out[l - 1], B, k, &upds[l], &rngs[l], loss_out[l], nullptr, sx, &t);
        if (rc) {
            if (cx != ctx0) ctx0->err = cx->err;
            return rc;
        }
    }
    if (piped) IMDBN_CUDA(ctx0, cudaEventRecord(ctx0->ev_done[par], s1));
    if (host_trace) {
        const double t_end = now();
        acc_t[0] += t_l0 - t_begin; acc_t[1] += t_l0_end - t_l0; acc_t[2] += t_end - t_l0_end; acc_t[3] += t_end - t_begin;
        if (++acc_n % 500 == 0) {
            fprintf(stderr, "imdbn_idbn_train_step host us/call: entry+events %.1f, layer 0 %.1f, upper layers+events %.1f, total %.1f (piped=%d)\n",
                    acc_t[0] / 500, acc_t[1] / 500, acc_t[2] / 500, acc_t[3] / 500, (int)piped);
            acc_t[0] = acc_t[1] = acc_t[2] = acc_t[3] = 0;
        }
    }
    return 0;
}

int64_t imdbn_stats_size(const imdbn_rbm* r) {
    return (int64_t)r->V * r->H + 2 * (int64_t)r->H + r->V + 1;
}

int imdbn_cd_stats(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* data, int B, int k,
                   const imdbn_rng* rng, const float* pos_h_in, float* stats_out, imdbn_stream stream) {
    IMDBN_ARG(ctx, stats_out != nullptr);
    FwdTail t{pos_h_in, nullptr, 0, nullptr};
    return cd_core(ctx, rbm, data, B, k, nullptr, rng, nullptr, stats_out, (cudaStream_t)stream, pos_h_in ? &t : nullptr);
}

int imdbn_apply_update(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* stats,
                       const imdbn_update* upd, float* loss_out, imdbn_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rbm(ctx, rbm, true);
    if (rc) return rc;
    IMDBN_ARG(ctx, stats && upd && upd->batch_global > 0);
    const size_t n = (size_t)rbm->V * rbm->H;
    k_weight_update<<<ew_blocks(n, ctx->num_sms), 256, 0, st>>>(stats, n, rbm->W, rbm->Wm, upd->lr,
                                                               upd->momentum, upd->weight_decay,
                                                               (float)upd->batch_global);
    IMDBN_CHECK_LAUNCH(ctx, "k_weight_update");
    return bias_update(ctx, rbm, const_cast<float*>(stats) + n, upd, (float)upd->batch_global * rbm->V, loss_out, st);
}

int imdbn_dp_update(imdbn_ctx* ctx, const imdbn_rbm* rbm, const imdbn_peers* peers,
                    const imdbn_update* upd, float* loss_out, imdbn_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rbm(ctx, rbm, true);
    if (rc) return rc;
    IMDBN_ARG(ctx, peers && upd && upd->batch_global > 0);
    IMDBN_ARG(ctx, peers->world >= 1 && peers->world <= IMDBN_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world);
    const size_t n = (size_t)rbm->V * rbm->H;
    IMDBN_ARG(ctx, n % 4 == 0 && al16(rbm->Wm) && peers->W[peers->rank] == rbm->W);
    PeerPtrs p{};
    p.world = peers->world;
    for (int r = 0; r < peers->world; ++r) {
        IMDBN_ARG(ctx, peers->stats[r] && peers->W[r] && al16(peers->stats[r]) && al16(peers->W[r]));
        p.stats[r] = peers->stats[r];
        p.W[r] = peers->W[r];
    }
    static const int dp_dbg = getenv("IMDBN_DP_DEBUG") ? atoi(getenv("IMDBN_DP_DEBUG")) : 0;   // timing experiments only
    if (dp_dbg == 1) for (int r = 0; r < peers->world; ++r) p.W[r] = peers->W[peers->rank];        // no peer stores
    if (dp_dbg == 2) for (int r = 0; r < peers->world; ++r) p.stats[r] = peers->stats[peers->rank]; // no peer loads
    const int n_small = 2 * rbm->H + rbm->V + 1;
    rc = arena_begin(ctx, pad256(n_small), st);
    if (rc) return rc;
    float* st_small = arena_take<float>(ctx, n_small);
    k_dp_small<<<(n_small + 255) / 256, 256, 0, st>>>(p, n, n_small, st_small);
    IMDBN_CHECK_LAUNCH(ctx, "k_dp_small");
    const size_t nq = n / 4;
    const size_t q0 = nq * (size_t)peers->rank / peers->world, q1 = nq * (size_t)(peers->rank + 1) / peers->world;
    if (q1 > q0) {
        const int blocks = (int)std::min<size_t>((q1 - q0 + 255) / 256, (size_t)ctx->num_sms * 8);
        if (peers->stats_mc && peers->W_mc) {
            IMDBN_ARG(ctx, al16(peers->stats_mc) && al16(peers->W_mc));
            p.stats_mc = peers->stats_mc; p.W_mc = peers->W_mc;
            k_dp_update<true><<<blocks, 256, 0, st>>>(p, peers->rank, q0, q1, rbm->Wm, upd->lr, upd->momentum,
                                                      upd->weight_decay, (float)upd->batch_global);
        } else {
            k_dp_update<false><<<blocks, 256, 0, st>>>(p, peers->rank, q0, q1, rbm->Wm, upd->lr, upd->momentum,
                                                       upd->weight_decay, (float)upd->batch_global);
        }
        IMDBN_CHECK_LAUNCH(ctx, "k_dp_update");
    }
    return bias_update(ctx, rbm, st_small, upd, (float)upd->batch_global * rbm->V, loss_out, st);
}

// ---- IMG->TXT diagnostics on the label-only structure (utils/energy_utils.py) --------------------------
namespace {
__global__ void k_pad_cols(const float* __restrict__ src, int B, int Dz, int V, float* __restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= V) return;
    for (int b = blockIdx.y; b < B; b += gridDim.y) dst[(size_t)b * V + c] = c < Dz ? src[(size_t)b * Dz + c] : 0.0f;
}

// pre = z W_z + b_h for a batch of latents z [B,Dz] (one up-pass GEMM over [z | 0])
int label_preact(imdbn_ctx* ctx, const imdbn_rbm* r, const float* z, int B, int Dz, size_t extra_floats,
                 float** pre_out, cudaStream_t st) {
    const int V = r->V, H = r->H;
    const PassPlan pu = plan_pass(ctx, r, B, true);
    int rc = arena_begin(ctx, pad256(pu.part_floats) + pad256((size_t)B * V) + pad256((size_t)B * H) +
                                  pad256(extra_floats), st);
    if (rc) return rc;
    float* part = arena_take<float>(ctx, pu.part_floats);
    float* vz = arena_take<float>(ctx, (size_t)B * V);
    float* pre = arena_take<float>(ctx, (size_t)B * H);
    k_pad_cols<<<dim3((V + 255) / 256, std::min(B, 16384)), 256, 0, st>>>(z, B, Dz, V, vz);
    IMDBN_CHECK_LAUNCH(ctx, "k_pad_cols");
    rc = gemm_up(ctx, r, vz, B, pu, part, st);
    if (rc) return rc;
    k_preact<<<dim3((H + 255) / 256, std::min(B, 16384)), 256, 0, st>>>(part, pu.splits, pu.sk, B, H, r->hb, pre);
    IMDBN_CHECK_LAUNCH(ctx, "k_preact");
    *pre_out = pre;
    return 0;
}
}  // namespace

int imdbn_class_free_energies(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* z, int B, int Dz,
                              float* F_out, imdbn_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rbm(ctx, rbm, false);
    if (rc) return rc;
    IMDBN_ARG(ctx, z && F_out && B > 0 && Dz > 0 && Dz < rbm->V);
    const int K = rbm->V - Dz, H = rbm->H;
    float* pre = nullptr;
    rc = label_preact(ctx, rbm, z, B, Dz, 0, &pre, st);
    if (rc) return rc;
    k_class_free_energies<<<B, 256, H * sizeof(float), st>>>(pre, z, Dz, rbm->W + (size_t)Dz * H, rbm->vb, B, H, Dz, K,
                                                            F_out);
    IMDBN_CHECK_LAUNCH(ctx, "k_class_free_energies");
    return 0;
}

int imdbn_trace_img2txt(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* z, int B, int Dz,
                        const float* y_init, int steps, float* y_traj, imdbn_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rbm(ctx, rbm, false);
    if (rc) return rc;
    IMDBN_ARG(ctx, z && y_traj && B > 0 && steps > 0 && Dz > 0 && Dz < rbm->V);
    const int K = rbm->V - Dz, H = rbm->H;
    IMDBN_ARG(ctx, K <= 32 && H % 32 == 0);
    const size_t smem = ((size_t)64 * H + (size_t)LG_WARPS * H) * sizeof(float);
    IMDBN_ARG(ctx, smem <= 200 * 1024);
    float* pre = nullptr;
    rc = label_preact(ctx, rbm, z, B, Dz, 0, &pre, st);
    if (rc) return rc;
    TraceArgs a{};
    a.pre = pre; a.Wy = rbm->W + (size_t)Dz * H; a.vby = rbm->vb + Dz; a.y_init = y_init;
    a.B = B; a.H = H; a.K = K; a.steps = steps; a.y_traj = y_traj;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        IMDBN_CUDA(ctx, cudaFuncSetAttribute(k_trace_img2txt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const int blocks = std::max(1, std::min((B + LG_WARPS - 1) / LG_WARPS, ctx->num_sms * 2));
    k_trace_img2txt<<<blocks, LG_WARPS * 32, smem, st>>>(a);
    IMDBN_CHECK_LAUNCH(ctx, "k_trace_img2txt");
    return 0;
}

int imdbn_assoc_stats(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* vp, const float* hp,
                      const float* vn, const float* hn, int B, float* dS_out, imdbn_stream stream) {
    int rc = check_rbm(ctx, rbm, false);
    if (rc) return rc;
    IMDBN_ARG(ctx, vp && hp && vn && hn && dS_out && B > 0);
    rc = arena_begin(ctx, tc_ws_bytes(ctx, rbm, B), (cudaStream_t)stream);
    if (rc) return rc;
    return gemm_stats(ctx, rbm, vp, hp, vn, hn, B, dS_out, nullptr, (cudaStream_t)stream);
}

int imdbn_run_chain(imdbn_ctx* ctx, const imdbn_rbm* rbm, const imdbn_chain* ch, int B,
                    float* v_out, float* vprob_out, const imdbn_rng* rng, imdbn_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rbm(ctx, rbm, false);
    if (rc) return rc;
    IMDBN_ARG(ctx, ch && v_out && rng && B > 0);
    rc = arena_begin(ctx, pad256(chain_ws_floats(ctx, rbm, ch, B)) + 4096, st);
    if (rc) return rc;
    float* Wt = arena_take<float>(ctx, (size_t)rbm->V * rbm->H);
    float* tables = arena_take<float>(ctx, 3 * (size_t)std::max(1, ch->n_steps));
    return run_chain(ctx, rbm, ch, B, v_out, vprob_out, make_key(rng), Wt, tables, st);
}

int imdbn_cd_train_clamped(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* v_known,
                           const float* known_mask, int B, const imdbn_clamped_cfg* cfg,
                           const imdbn_update* upd, const imdbn_rng* rng, float* loss_out,
                           imdbn_stream stream) {
    IMDBN_ARG(ctx, upd && upd->batch_global > 0);
    return clamped_core(ctx, rbm, v_known, known_mask, B, cfg, upd, rng, loss_out, nullptr,
                        (cudaStream_t)stream);
}

int imdbn_cd_clamped_stats(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* v_known,
                           const float* known_mask, int B, const imdbn_clamped_cfg* cfg,
                           const imdbn_rng* rng, float* stats_out, imdbn_stream stream) {
    IMDBN_ARG(ctx, stats_out != nullptr);
    return clamped_core(ctx, rbm, v_known, known_mask, B, cfg, nullptr, rng, nullptr, stats_out,
                        (cudaStream_t)stream);
}

int imdbn_best_of_k(imdbn_ctx* ctx, const float* cand, const float* F, int K, int B, int V,
                    float* out, int32_t* idx_out, imdbn_stream stream) {
    IMDBN_ARG(ctx, cand && F && out && K > 0 && B > 0 && V > 0);
    k_best_of_k<<<B, 128, 0, (cudaStream_t)stream>>>(cand, F, K, B, V, out, idx_out);
    IMDBN_CHECK_LAUNCH(ctx, "k_best_of_k");
    return 0;
}

int imdbn_class_stats(imdbn_ctx* ctx, const float* z, const float* y, int B, int Dz, int K,
                      float* sum_z, float* class_sum, float* class_count, float* label_sum,
                      imdbn_stream stream) {
    IMDBN_ARG(ctx, z && y && sum_z && class_sum && class_count && label_sum && B > 0 && Dz > 0 && K > 0);
    const int cols = std::max(Dz, K);
    k_class_stats<<<(cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(z, y, B, Dz, K, sum_z, class_sum,
                                                                      class_count, label_sum);
    IMDBN_CHECK_LAUNCH(ctx, "k_class_stats");
    return 0;
}

int imdbn_random_field(imdbn_ctx* ctx, const imdbn_rng* rng, uint32_t draw, int kind, int rows,
                       int cols, float* out, imdbn_stream stream) {
    IMDBN_ARG(ctx, rng && out && rows > 0 && cols > 0);
    k_random_field<<<ew_blocks((size_t)rows * cols, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(
        make_key(rng), draw, kind, rows, cols, out);
    IMDBN_CHECK_LAUNCH(ctx, "k_random_field");
    return 0;
}

}  // extern "C"
