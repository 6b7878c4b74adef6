// Tensor-core (tcgen05 / TMEM / TMA) path -- placeholder until the kernels land: reports
// "unsupported" so every pass runs on the fp32 FFMA engine.
#include "tc_gemm.cuh"

namespace imdbn {
bool tc_up_supported(const imdbn_ctx*, const imdbn_rbm*, int) { return false; }
bool tc_down_supported(const imdbn_ctx*, const imdbn_rbm*, int) { return false; }
bool tc_stats_supported(const imdbn_ctx*, const imdbn_rbm*, int) { return false; }
size_t tc_ws_bytes(const imdbn_ctx*, const imdbn_rbm*, int) { return 0; }
int tc_gemm_up(imdbn_ctx* ctx, const imdbn_rbm*, const float*, int, float*, cudaStream_t) { return fail(ctx, -4, "tc path not built"); }
int tc_gemm_down(imdbn_ctx* ctx, const imdbn_rbm*, const float*, int, float*, cudaStream_t) { return fail(ctx, -4, "tc path not built"); }
int tc_gemm_stats(imdbn_ctx* ctx, const imdbn_rbm*, const float*, const float*, const float*, const float*, int, float*, const imdbn_update*, cudaStream_t) { return fail(ctx, -4, "tc path not built"); }
void tc_destroy(imdbn_ctx*) {}
}  // namespace imdbn
