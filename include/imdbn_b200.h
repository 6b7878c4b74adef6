/*
 * imdbn_b200.h -- C ABI of libimdbn_b200.so: the sm_100a CUDA implementation of the RBM
 * contrastive-divergence / conditional-Gibbs hot path of francesco-cal98/multimodal-idbn.
 *
 * The reference has no FFI layer: its boundary for this path is the Python class API of
 * imdbn/models/rbm.py (RBM methods) called by idbn.py / imdbn.py / utils/conditional_steps.py.
 * Each entry point below names the reference method (file:line) whose arithmetic it replaces;
 * the Python mirror of those classes (multimodal_idbn_b200/{rbm,idbn,imdbn}.py) binds them with
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major fp32 unless it says "host";
 *     the caller owns all buffers; the library owns only the opaque context (workspace);
 *   - every call enqueues work on the given cudaStream_t and returns without synchronising
 *     (workspace growth is the one exception: it synchronises the stream once);
 *   - return value: 0 ok, <0 invalid argument, >0 a cudaError_t; text via imdbn_last_error();
 *   - nothing throws across the ABI; there is no CPU fallback anywhere.
 *
 * Random numbers: counter-based Philox4x32-10 addressed by (seed, stream, draw, global_row, col)
 * -- see oracle/philox.py for the normative host statement; `row0` is the global index of the
 * first row of the buffer (for batches sharded over ranks).
 */
#ifndef IMDBN_B200_H
#define IMDBN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMDBN_ABI_VERSION 1
#define IMDBN_MAX_GROUPS 4

typedef struct imdbn_ctx imdbn_ctx;
typedef void* imdbn_stream; /* cudaStream_t */

/* One RBM: parameters + momenta (imdbn/models/rbm.py:41-79). */
typedef struct {
    float* W;    /* [V, H] */
    float* hb;   /* [H]    hid_bias */
    float* vb;   /* [V]    vis_bias */
    float* Wm;   /* [V, H] W_m  (may be NULL for inference-only calls) */
    float* hbm;  /* [H]    hb_m */
    float* vbm;  /* [V]    vb_m */
    int32_t V, H;
    int32_t ngroups;                       /* softmax groups of the visible layer */
    int32_t group_start[IMDBN_MAX_GROUPS]; /* [s, e) column ranges (rbm.py:66,113) */
    int32_t group_end[IMDBN_MAX_GROUPS];
} imdbn_rbm;

typedef struct {
    uint64_t seed;
    uint32_t stream; /* one per stochastic API call of an RBM */
    uint32_t row0;   /* global row index of row 0 of the batch */
} imdbn_rng;

/* Hyper-parameters of one update, already resolved by the host
 * (lr = lr/(1+0.01*epoch), mom = momentum|final_momentum: rbm.py:194-195). */
typedef struct {
    float lr;              /* effective learning rate (includes aux_lr_mult for clamped CD) */
    float momentum;
    float weight_decay;
    int32_t sparsity;      /* rbm.py:217-219 */
    float sparsity_target;
    int32_t batch_global;  /* divisor `bsz` (== B unless the batch is sharded over ranks) */
} imdbn_update;

/* Precision of the GEMM-shaped passes. */
enum {
    IMDBN_PREC_FP32 = 0, /* fp32 FFMA accumulation: parity mode (samples bit-exact up to 1e-6) */
    IMDBN_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 accumulate in TMEM (<=1e-3 relative) */
    IMDBN_PREC_TF32X2 = 2 /* exact mode on the tensor cores: every operand enters as hi + lo tf32 terms (22+ significand
                           * bits), fp32 accumulate in TMEM, IEEE finishes: samples bit-exact up to 1e-6 like FP32 */
};

/* ---- context ------------------------------------------------------------------------------ */
int imdbn_abi_version(void);
int imdbn_ctx_create(imdbn_ctx** out, int device);
void imdbn_ctx_destroy(imdbn_ctx* ctx);
const char* imdbn_last_error(imdbn_ctx* ctx);
int imdbn_set_precision(imdbn_ctx* ctx, int prec);
/* Upper bound on the SMs the persistent tensor-core kernels launched through this context occupy (0 = all).
 * A context is bound to one stream: limiting the context that trains the large bottom layer leaves SMs to the
 * context (stream) that trains the small upper layers of an iDBN at the same time (idbn.py:199-204 pipelined
 * over minibatches: layer 0 of batch t+1 does not depend on the upper layers of batch t). */
int imdbn_set_sm_limit(imdbn_ctx* ctx, int n_sms);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t imdbn_launch_count(imdbn_ctx* ctx);
/* Asynchronous copy on `stream` (cudaMemcpyAsync, direction inferred from the pointers).  No reference counterpart:
 * the minibatch staging of iDBN.train (idbn.py:199-200, `.to(device)`) and the per-step read-back of the losses
 * (idbn.py:202, `.item()` in the reference) go through this call, so that the host side of a training step -- within
 * 10 % of the device side at batch 64 -- does not pay for a framework dispatch per copy. */
int imdbn_copy_async(void* dst, const void* src, size_t bytes, imdbn_stream stream);

/* ---- in-library kernel timing (bench.py's roofline leg) --------------------------------------
 * While enabled, the GEMM-shaped kernels and the chain kernel are bracketed by CUDA events recorded
 * on the launching stream.  imdbn_profile_read synchronises those events and returns the summed
 * duration (ms) and launch count of the kernels of `kind` that ran on an RBM of shape (V, H). */
enum { IMDBN_KERNEL_UP = 0, IMDBN_KERNEL_DOWN = 1, IMDBN_KERNEL_STATS = 2, IMDBN_KERNEL_CHAIN = 3,
       IMDBN_KERNEL_PACK = 4 /* operand packing pass in front of the small-batch statistics kernel */ };
int imdbn_profile_enable(imdbn_ctx* ctx, int enable);
int imdbn_profile_read(imdbn_ctx* ctx, int kind, int V, int H, double* ms_sum, int64_t* count);

/* ---- up / down passes ---------------------------------------------------------------------- */
/* RBM.forward (rbm.py:81-92) + the `p > rand_like(p)` sample (rbm.py:175,203,208).
 * p_out [B,H] = 1/(1+exp(-((v W + hb)/max(1e-6,T)))); s_out (nullable) = (p > U) with
 * U = uniform(draw_u). */
int imdbn_up(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* v, int B, float T,
             float* p_out, float* s_out, const imdbn_rng* rng, uint32_t draw_u,
             imdbn_stream stream);

/* RBM.visible_probs (rbm.py:98-116), RBM.backward(return_logits) (rbm.py:137-151) and
 * RBM.sample_visible (rbm.py:118-135).  Any of the three outputs may be NULL.
 * s_out uses uniform(draw_u) for the Bernoulli part and uniform(draw_cat) (col = group) for the
 * one-hot categorical of each softmax group. */
int imdbn_down(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* h, int B, float T,
               float* p_out, float* logits_out, float* s_out, const imdbn_rng* rng,
               uint32_t draw_u, uint32_t draw_cat, imdbn_stream stream);

/* RBM.sample_visible on given probabilities (rbm.py:118-135). */
int imdbn_sample_visible(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* p, int B,
                         float* s_out, const imdbn_rng* rng, uint32_t draw_u, uint32_t draw_cat,
                         imdbn_stream stream);

/* energy_utils.rbm_free_energy (imdbn/utils/energy_utils.py:18-28): F_out [B]. */
int imdbn_free_energy(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* v, int B, float* F_out,
                      imdbn_stream stream);

/* ---- CD-k ---------------------------------------------------------------------------------- */
/* RBM.train_epoch (rbm.py:180-227): CD-k statistics + in-place momentum / weight-decay update of
 * W, hb, vb, Wm, hbm, vbm.  loss_out (device, 1 float) = mean((data - v_prob_last)^2).
 * Draws: 0 = U[B,H]; step s: 1+3s = U[B,V], 2+3s = categorical, 3+3s = U[B,H]. */
int imdbn_cd_train(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* data, int B, int k,
                   const imdbn_update* upd, const imdbn_rng* rng, float* loss_out,
                   imdbn_stream stream);

/* RBM.train_epoch followed by the `v = rbm.forward(v)` of the iDBN training loop (idbn.py:202-203), with
 * one optimisation of the traffic: the forward pass with the UPDATED weights also computes, in the same
 * pass over W, the positive hidden probabilities of the NEXT minibatch (which uses those same weights).
 *   pos_h_in  nullable [B,H]      : positive probabilities of `data` produced by the previous call
 *   next_data nullable [B_next,V] : the next minibatch
 *   fwd_out   [B + B_next, H]     : rows [0,B) = forward(data), rows [B,..) = forward(next_data),
 *                                   both with the updated parameters */
int imdbn_cd_train_fwd(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* data, int B, int k,
                       const imdbn_update* upd, const imdbn_rng* rng, float* loss_out,
                       const float* pos_h_in, const float* next_data, int B_next, float* fwd_out,
                       imdbn_stream stream);

/* One minibatch of iDBN.train for all layers (idbn.py:199-204): layer l trains on fwd_out[l-1] rows [0,B) (layer 0
 * on `data`, with pos_h_in / next_data / B_next as in imdbn_cd_train_fwd) and writes forward(input) under its
 * updated parameters to fwd_out[l] ([B (+B_next for l = 0), H_l]); loss_out[l] = one float (device or mapped pinned
 * host memory).  ctx1 / stream1 non-NULL: layers >= 1 are enqueued on stream1 behind an event, so they overlap
 * whatever follows on stream0 -- the next minibatch's layer 0, which does not depend on them; the caller must then
 * alternate between TWO sets of fwd_out buffers (the call waits for the readers of the set used two calls ago). */
int imdbn_idbn_train_step(imdbn_ctx* ctx0, imdbn_ctx* ctx1, int n_layers, const imdbn_rbm* rbms,
                          const imdbn_update* upds, const imdbn_rng* rngs, const float* data, int B, int k,
                          const float* pos_h_in, const float* next_data, int B_next, float* const* fwd_out,
                          float* const* loss_out, int buffer_set /* 0 | 1: which set fwd_out belongs to */,
                          imdbn_stream stream0, imdbn_stream stream1,
                          imdbn_stream caller_stream /* nullable: stream the inputs were produced on, if not stream0 */,
                          int early_launch /* 0: layer 0 without programmatic dependent launch */);

/* (No reference counterpart: the reference runs the layers of idbn.py:199-204 back to back on one stream.)
 * Two streams whose kernels run on DISJOINT sets of SMs (CUDA green contexts, created once per device): the small
 * partition has >= small_sms SMs (multiples of 8), the big one the rest.  Used to train the upper layers of an
 * iDBN next to the bottom layer without either side's early-launched grids squatting on the other's SMs. */
int imdbn_sm_partition(int device, int small_sms, void** stream_big, void** stream_small, int* n_big, int* n_small);

/* The same statistics without the update (rbm.py:199-209, 216, 223, 226), for batches sharded over ranks: writes the local sums
 *   stats_out = [ dS (V*H) | dh (H) | dv (V) | pos_h column sum (H) | squared error (1) ]
 * which the host all-reduces (NCCL) and hands to imdbn_apply_update on every rank. */
int imdbn_cd_stats(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* data, int B, int k,
                   const imdbn_rng* rng, const float* pos_h_in /* nullable, as in imdbn_cd_train_fwd */,
                   float* stats_out, imdbn_stream stream);
int64_t imdbn_stats_size(const imdbn_rbm* rbm);
/* rbm.py:211-226 on (all-reduced) statistics; n_elem_loss = global B*V for the loss mean. */
int imdbn_apply_update(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* stats,
                       const imdbn_update* upd, float* loss_out, imdbn_stream stream);

/* Data-parallel update over PEER MEMORY (NVLink / NVSwitch) -- replaces "all-reduce + imdbn_apply_update".
 * Every rank has written its local statistics (imdbn_cd_stats / imdbn_cd_clamped_stats layout) into a buffer
 * that is mapped into all ranks (CUDA IPC / symmetric memory), and every rank's W lives in such a buffer too.
 * One kernel per rank: reduce-scatter of dS in rank order, rbm.py:212-213 on the rank's slab of W and W_m
 * (W_m is only valid on its owner afterwards), all-gather of the new W into every rank's copy; the bias update
 * and the loss (rbm.py:216-226) are computed redundantly on every rank from the rank-ordered sum of the small
 * statistics.  The caller provides a cross-rank barrier BEFORE (all statistics written) and AFTER (all weights
 * written, all statistics consumed) the call. */
#define IMDBN_MAX_PEERS 8
typedef struct {
    int32_t world, rank;
    const float* stats[IMDBN_MAX_PEERS]; /* rank i's statistics buffer, as mapped in THIS process */
    float* W[IMDBN_MAX_PEERS];           /* rank i's W [V,H]; W[rank] == rbm->W */
    const float* stats_mc;               /* nullable: NVSwitch multicast address of the statistics buffers --
                                            the sum over ranks is then formed IN THE SWITCH (multimem.ld_reduce) */
    float* W_mc;                         /* nullable: multicast address of the W buffers (multimem.st broadcast) */
} imdbn_peers;
int imdbn_dp_update(imdbn_ctx* ctx, const imdbn_rbm* rbm, const imdbn_peers* peers,
                    const imdbn_update* upd, float* loss_out, imdbn_stream stream);

/* IMG->TXT diagnostics with the image latents z [B,Dz] clamped and the remaining K = V - Dz visible units being
 * the label group (utils/energy_utils.py):
 *   class free energies (energy_utils.py:32-54): F_out[b][k] = F([z_b, e_k]) for every label k;
 *   deterministic mean-field-lite trace (energy_utils.py:61-90 driven by :132-160): per step
 *     h = sigmoid(zW_z + yW_y + b_h), s = sigmoid(hW_y^T + b_y), y = softmax(s); y_traj[t][b][:] = y after step
 *     t+1; y_init nullable (uniform 1/K).  K <= 32, H % 32 == 0. */
int imdbn_class_free_energies(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* z, int B, int Dz,
                              float* F_out, imdbn_stream stream);
int imdbn_trace_img2txt(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* z, int B, int Dz,
                        const float* y_init, int steps, float* y_traj, imdbn_stream stream);

/* The association statistics alone (rbm.py:200,209): dS_out [V,H] = vp^T hp - vn^T hn with
 * vp, vn [B,V] and hp, hn [B,H]. */
int imdbn_assoc_stats(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* vp, const float* hp,
                      const float* vn, const float* hn, int B, float* dS_out, imdbn_stream stream);

/* ---- conditional inference chains ---------------------------------------------------------- */
enum { IMDBN_CHAIN_NOISY_MF = 0, IMDBN_CHAIN_COND_GIBBS = 1 };

typedef struct {
    int32_t kind;        /* IMDBN_CHAIN_* */
    int32_t n_steps;
    const float* v_known;   /* [B,V] */
    const float* known_mask;/* [B,V] 1 = clamped */
    const float* v_init;    /* nullable [B,V]: start state instead of the random init (draw0) */
    /* noisy mean-field (rbm.py:300-367): HOST arrays of n_steps entries each */
    const float* T;         /* host: max(1e-6, T_t) */
    const float* sigma;     /* host: sigma_t */
    const float* eta;       /* host: eta_t (mu-pull, rbm.py:359-363) */
    const float* mu;        /* nullable [B,Dz] */
    int32_t Dz;
    /* conditional Gibbs (rbm.py:369-400) */
    int32_t sample_h, sample_v;
    int32_t final_free_sweep; /* 1: return visible_probs(forward(v)) un-clamped (rbm.py:400) */
    uint32_t draw0;           /* first draw index used by this chain */
    int32_t clamp_prefix;     /* >= 0: the caller promises known_mask == 1 exactly on columns
                               * [0, clamp_prefix) and 0 elsewhere (enables the label-only fast path
                               * of IMG->TXT inference); -1: arbitrary mask */
    int32_t clamp_suffix;     /* >= 0: the caller promises known_mask == 1 exactly on columns [clamp_suffix, V)
                               * and 0 elsewhere (TXT->IMG inference, imdbn.py:431-433): the large-batch chain then
                               * skips the mask / known-value loads of the free columns and all work on the clamped
                               * ones (their noise draws never reach an output); -1: arbitrary mask */
} imdbn_chain;

/* RBM.noisy_meanfield_annealed (rbm.py:300-367) / RBM.conditional_gibbs (rbm.py:369-400) /
 * conditional_steps._gibbs_conditional_step (utils/conditional_steps.py:15-34; n_steps=1 with
 * v_init).  v_out [B,V]; vprob_out (nullable) = un-clamped visible probabilities of the last
 * sweep.  The whole chain runs inside one persistent kernel. */
int imdbn_run_chain(imdbn_ctx* ctx, const imdbn_rbm* rbm, const imdbn_chain* ch, int B,
                    float* v_out, float* vprob_out, const imdbn_rng* rng, imdbn_stream stream);

/* RBM.train_epoch_clamped (rbm.py:402-483).  Flags as in the reference signature. */
typedef struct {
    int32_t k;                /* CD */
    int32_t cond_init_steps;
    int32_t sample_h, sample_v, reclamp_negative, use_noisy_init;
} imdbn_clamped_cfg;

int imdbn_cd_train_clamped(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* v_known,
                           const float* known_mask, int B, const imdbn_clamped_cfg* cfg,
                           const imdbn_update* upd, const imdbn_rng* rng, float* loss_out,
                           imdbn_stream stream);
int imdbn_cd_clamped_stats(imdbn_ctx* ctx, const imdbn_rbm* rbm, const float* v_known,
                           const float* known_mask, int B, const imdbn_clamped_cfg* cfg,
                           const imdbn_rng* rng, float* stats_out, imdbn_stream stream);

/* ---- iMDBN helpers ------------------------------------------------------------------------- */
/* Best-of-K selection of iMDBN._cross_reconstruct (imdbn.py:472-474): cand [K,B,V], F [K,B]
 * -> out[b] = cand[argmin_k F[k,b], b] (first minimum wins, like torch.argmin), idx_out int32 [B]. */
int imdbn_best_of_k(imdbn_ctx* ctx, const float* cand, const float* F, int K, int B, int V,
                    float* out, int32_t* idx_out, imdbn_stream stream);

/* Per-class sums for iMDBN.init_joint_bias_from_data (imdbn.py:244-277):
 * accumulates sum_z [Dz] += sum_b z, class_sum [K,Dz] += onehot^T z, class_count [K] += sum_b y
 * with class index = argmax(y). */
int imdbn_class_stats(imdbn_ctx* ctx, const float* z, const float* y, int B, int Dz, int K,
                      float* sum_z, float* class_sum, float* class_count, float* label_sum,
                      imdbn_stream stream);

/* Raw random field, for tests: out [rows, cols]; kind 0 uniform, 1 normal. */
int imdbn_random_field(imdbn_ctx* ctx, const imdbn_rng* rng, uint32_t draw, int kind, int rows,
                       int cols, float* out, imdbn_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* IMDBN_B200_H */
