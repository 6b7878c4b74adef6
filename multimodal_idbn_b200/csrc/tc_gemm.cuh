// Tensor-core (tcgen05 / TMEM / TMA) path of the GEMM-shaped passes: interface used by
// imdbn_b200.cu.  Implemented in tc_gemm.cu.
#pragma once
#include "common.cuh"

namespace imdbn {

bool tc_up_supported(const imdbn_ctx* ctx, const imdbn_rbm* r, int B);
bool tc_down_supported(const imdbn_ctx* ctx, const imdbn_rbm* r, int B);
bool tc_stats_supported(const imdbn_ctx* ctx, const imdbn_rbm* r, int B);
size_t tc_ws_bytes(const imdbn_ctx* ctx, const imdbn_rbm* r, int B);
// the statistics call for this shape packs its operands (and can take the column statistics along: ColstatsJob)
bool tc_stats_packs(const imdbn_ctx* ctx, const imdbn_rbm* r, int B);

// Finish of a pass done by the pass kernel's own epilogue (large batches, where every output tile is accumulated by
// ONE CTA and the partial-slab round trip plus the finish kernel are pure overhead): bias, temperature, sigmoid and the
// Bernoulli sample of rbm.py:92,110,125,175,203, the arithmetic and the random field of k_finish_up4 / k_finish_down4<true>.
struct FusedFinish {
    const float* bias; float invT; float* p_out; float* s_out; RngKey key; uint32_t draw_u;
};
// the pass of this shape accumulates whole tiles per CTA in the fast tensor-core mode: its finish can be fused
bool tc_pass_fusable(const imdbn_ctx* ctx, const imdbn_rbm* r, int B, bool up);

// v2 != nullptr: the batch is [v (B1 rows) ; v2 (B - B1 rows)] read from two matrices (B1 % 8 == 0, B <= 256)
int tc_gemm_up(imdbn_ctx* ctx, const imdbn_rbm* r, const float* v, int B, float* part, cudaStream_t st,
               const float* v2 = nullptr, int B1 = 0, const FusedFinish* fin = nullptr);
int tc_gemm_down(imdbn_ctx* ctx, const imdbn_rbm* r, const float* h, int B, float* part, cudaStream_t st,
                 const FusedFinish* fin = nullptr);
int tc_gemm_stats(imdbn_ctx* ctx, const imdbn_rbm* r, const float* vp, const float* hp,
                  const float* vn, const float* hn, int B, float* dS_out, const imdbn_update* upd,
                  cudaStream_t st);
// TXT->IMG noisy mean-field chains as one persistent tcgen05 kernel (chain_tc.cuh); T / sigma / eta: device tables
bool tc_chain_supported(const imdbn_ctx* ctx, const imdbn_rbm* r, const imdbn_chain* ch, int B);
int tc_chain_t2i(imdbn_ctx* ctx, const imdbn_rbm* r, const imdbn_chain* ch, int B, float* v_out, const RngKey& key,
                 const float* T, const float* sigma, const float* eta, cudaStream_t st);
void tc_destroy(imdbn_ctx* ctx);
// stream-K plan of a tensor-core pass producing M_total output features from K_total inputs
SKPlan tc_plan(const imdbn_ctx* ctx, int M_total, int K_total, int B);
int tc_plan_max_slabs(const SKPlan& p, int M_total);

// two streams on disjoint SM sets (CUDA green contexts); see imdbn_sm_partition in the public header
int sm_partition(int device, int small_sms, void** stream_big, void** stream_small, int* n_big, int* n_small);

}  // namespace imdbn
