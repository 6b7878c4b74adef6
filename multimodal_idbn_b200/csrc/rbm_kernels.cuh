// Element-wise / reduction kernels around the GEMMs: the fused bias + temperature + sigmoid +
// Philox-sampling finishes of the up and down passes, softmax groups, column statistics and the
// bias updates of CD-k (imdbn/models/rbm.py:81-135, 211-226).
#pragma once
#include "common.cuh"

namespace imdbn {

// Rows on blockIdx.y (grid-stride), columns on blockIdx.x * blockDim.x + threadIdx.x: no integer
// division on the element index.  `part` holds `splits` uniform K-split slabs (FFMA engine) or the
// stream-K slabs of the tensor-core pass (SKPlan), summed here in a fixed order.
__device__ __forceinline__ int finish_nslabs(const SKPlan& sk, int splits, int col) {
    return sk.k_iters ? sk_nslabs(sk, col / sk.tile_w) : splits;
}
// part[0*stride + i] + part[1*stride + i] + ... in slab order; the loads of 8 slabs are issued
// together (independent), only the additions are sequential.
__device__ __forceinline__ float sum_slabs(const float* __restrict__ part, int ns, size_t stride, size_t i) {
    float x = 0.0f;
    for (int s0 = 0; s0 < ns; s0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (s0 + u < ns) ? part[(size_t)(s0 + u) * stride + i] : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (s0 + u < ns) x += v[u];
    }
    return x;
}

// ---- up pass finish: p = sigmoid((sum_s part + hb)/T), s = (p > U)           rbm.py:92,175,203
__global__ void k_finish_up(const float* __restrict__ part, int splits, SKPlan sk, int B, int H,
                            const float* __restrict__ hb, float T, float* __restrict__ p_out,
                            float* __restrict__ s_out, RngKey key, uint32_t draw_u) {
    pdl_trigger();
    pdl_wait();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= H) return;
    const int ns = finish_nslabs(sk, splits, j);
    const size_t n = (size_t)B * H;
    const float bj = hb[j], invT = 1.0f / T;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t i = (size_t)b * H + j;
        float x = div_by(add_rn(sum_slabs(part, ns, n, i), bj), T, invT);
        const float p = sigmoidf_ref(x);
        if (p_out) p_out[i] = p;
        if (s_out) s_out[i] = (p > rf_uniform(key, draw_u, b, j)) ? 1.0f : 0.0f;
    }
}

// ---- down pass finish: logits = (sum_s part + vb)/T, p = sigmoid(logits), s = (p > U)
//      (softmax groups are overwritten afterwards by k_groups)                 rbm.py:96,110,125
__global__ void k_finish_down(const float* __restrict__ part, int splits, SKPlan sk, int B, int V,
                              const float* __restrict__ vb, float T, float* __restrict__ p_out,
                              float* __restrict__ logits_out, float* __restrict__ s_out,
                              RngKey key, uint32_t draw_u) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= V) return;
    const int ns = finish_nslabs(sk, splits, c);
    const size_t n = (size_t)B * V;
    const float bc = vb[c], invT = 1.0f / T;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t i = (size_t)b * V + c;
        float x = div_by(add_rn(sum_slabs(part, ns, n, i), bc), T, invT);
        if (logits_out) logits_out[i] = x;
        const float p = sigmoidf_ref(x);
        if (p_out) p_out[i] = p;
        if (s_out) s_out[i] = (p > rf_uniform(key, draw_u, b, c)) ? 1.0f : 0.0f;
    }
}

// Bernoulli part of sample_visible on given probabilities                       rbm.py:125
__global__ void k_bernoulli(const float* __restrict__ p, int B, int V, float* __restrict__ s_out,
                            RngKey key, uint32_t draw_u) {
    const size_t n = (size_t)B * V;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / V), c = (int)(i % V);
        s_out[i] = (p[i] > rf_uniform(key, draw_u, b, c)) ? 1.0f : 0.0f;
    }
}

// The same on four columns per thread (one Philox call), launched as a programmatic dependent: the wait comes first
// and the trigger after it, so whatever is launched behind this kernel still starts only once everything before it
// has completed (the CD passes that follow prefetch W before their own wait).  V % 4 == 0, 16-byte aligned rows.
__global__ void __launch_bounds__(256) k_bernoulli4(const float* __restrict__ p, int B, int V, float* __restrict__ s_out,
                                                    RngKey key, uint32_t draw_u) {
    pdl_wait();
    pdl_trigger();
    const unsigned q_per_row = (unsigned)V >> 2, total = (unsigned)B * q_per_row;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned b = idx / q_per_row, c = (idx - b * q_per_row) << 2;
        const float4 x = *reinterpret_cast<const float4*>(p + (size_t)b * V + c);
        const float4 u = rf_uniform4(key, draw_u, b, c);
        *reinterpret_cast<float4*>(s_out + (size_t)b * V + c) =
            make_float4(x.x > u.x ? 1.0f : 0.0f, x.y > u.y ? 1.0f : 0.0f, x.z > u.z ? 1.0f : 0.0f, x.w > u.w ? 1.0f : 0.0f);
    }
}

// One warp per (row, group): softmax of the logits over [s,e) written over p (rbm.py:113-114) and,
// if s_out, the one-hot categorical draw of sample_visible (rbm.py:129-133): q = clamp(p,1e-8,1),
// q /= sum q, idx = #{j: cdf_j <= u} with a sequential fp32 cdf (oracle.categorical_index).
// With logits == nullptr the probabilities already in p are used (sample_visible on a given p).
__global__ void k_groups(const float* __restrict__ logits, float* __restrict__ p, float* __restrict__ s_out,
                         int B, int V, Groups gr, RngKey key, uint32_t draw_cat) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= B * gr.n) return;
    const int b = warp / gr.n, g = warp % gr.n;
    const int s = gr.s[g], e = gr.e[g];
    const size_t row = (size_t)b * V;
    if (logits) {
        float mx = -INFINITY;
        for (int c = s + lane; c < e; c += 32) mx = fmaxf(mx, logits[row + c]);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.0f;
        for (int c = s + lane; c < e; c += 32) sum += expf(logits[row + c] - mx);
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float rsum = 1.0f / sum;
        for (int c = s + lane; c < e; c += 32) p[row + c] = div_by(expf(logits[row + c] - mx), sum, rsum);
        __syncwarp();
    }
    if (s_out && lane == 0) {
        float tot = 0.0f;
        for (int c = s; c < e; ++c) tot += fminf(fmaxf(p[row + c], 1e-8f), 1.0f);
        const float u = rf_uniform(key, draw_cat, b, g);
        float cdf = 0.0f;
        int idx = 0;
        const float rtot = 1.0f / tot;
        for (int c = s; c < e; ++c) {
            cdf += div_by(fminf(fmaxf(p[row + c], 1e-8f), 1.0f), tot, rtot);
            idx += (cdf <= u) ? 1 : 0;
        }
        idx = min(idx, e - s - 1);
        for (int c = s; c < e; ++c) s_out[row + c] = (c - s == idx) ? 1.0f : 0.0f;
    }
}

// ---- column statistics of one CD step (rbm.py:216,223,226 / 478,480,483)
//   out = [ dh (H) | dv (V) | pos_h column sum (H) | squared error (1, written by k_bias_update) ]
//   dh = sum_b hp - sum_b hn ; dv = sum_b vp - sum_b vn ; sq = sum (ea - eb)^2 over [B,V]
// Block = 32 columns x 8 row lanes; every thread sums rows y, y+8, ...; the 8 lane sums are added in
// lane order (deterministic).  One squared-error partial per block.
constexpr int CS_COLS = 32, CS_ROWS = 8;
__global__ void __launch_bounds__(CS_COLS * CS_ROWS)
k_colstats(const float* __restrict__ hp, const float* __restrict__ hn, const float* __restrict__ vp,
           const float* __restrict__ vn, const float* __restrict__ ea, const float* __restrict__ eb, int B,
           int V, int H, float* __restrict__ out, float* __restrict__ sq_part, unsigned int* __restrict__ ticket,
           BiasArgs ba) {
    __shared__ float red[5][CS_ROWS][CS_COLS];
    __shared__ unsigned int s_last;
    // Trigger AFTER the wait: the kernel launched next (the statistics GEMM, which only shares read-only
    // inputs with this one) may then start immediately and run CONCURRENTLY with this kernel -- everything
    // before this kernel has completed by the time the trigger fires.
    pdl_wait();
    pdl_trigger();
    const int x = threadIdx.x % CS_COLS, y = threadIdx.x / CS_COLS;
    const int c = blockIdx.x * CS_COLS + x;
    float ha = 0.f, hb_ = 0.f, va = 0.f, vb_ = 0.f, sq = 0.f;
    // The loads of CS_UN rows are issued together (large batches walk thousands of rows per thread: one memory round
    // trip per row made this kernel 11 % of a batch-8192 update); the additions keep the row order.
    constexpr int CS_UN = 16;
    if (c < H)
        for (int b0 = y; b0 < B; b0 += CS_ROWS * CS_UN) {
            float p[CS_UN], n[CS_UN];
#pragma unroll
            for (int u = 0; u < CS_UN; ++u) {
                const int b = b0 + u * CS_ROWS;
                p[u] = n[u] = 0.f;
                if (b < B) { p[u] = hp[(size_t)b * H + c]; n[u] = hn[(size_t)b * H + c]; }
            }
#pragma unroll
            for (int u = 0; u < CS_UN; ++u)
                if (b0 + u * CS_ROWS < B) { ha += p[u]; hb_ += n[u]; }
        }
    if (c < V) {
        const bool same = (ea == vp) && (eb == vn);
        for (int b0 = y; b0 < B; b0 += CS_ROWS * CS_UN) {
            float p[CS_UN], n[CS_UN], e0[CS_UN], e1[CS_UN];
#pragma unroll
            for (int u = 0; u < CS_UN; ++u) {
                const int b = b0 + u * CS_ROWS;
                p[u] = n[u] = e0[u] = e1[u] = 0.f;
                if (b < B) {
                    const size_t o = (size_t)b * V + c;
                    p[u] = vp[o]; n[u] = vn[o];
                    if (!same) { e0[u] = ea[o]; e1[u] = eb[o]; }
                }
            }
#pragma unroll
            for (int u = 0; u < CS_UN; ++u)
                if (b0 + u * CS_ROWS < B) {
                    va += p[u]; vb_ += n[u];
                    const float d = same ? p[u] - n[u] : e0[u] - e1[u];
                    sq = fmaf(d, d, sq);
                }
        }
    }
    red[0][y][x] = ha; red[1][y][x] = hb_; red[2][y][x] = va; red[3][y][x] = vb_; red[4][y][x] = sq;
    __syncthreads();
    if (y == 0) {
        float t[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            float a = 0.f;
#pragma unroll
            for (int r = 0; r < CS_ROWS; ++r) a += red[q][r][x];
            t[q] = a;
        }
        if (c < H) {
            const float dh = t[0] - t[1];
            out[c] = dh; out[H + V + c] = t[0];
            if (ba.apply) {                               // rbm.py:216-220 / 478-479
                float m = add_rn(mul_rn(ba.hbm[c], ba.mom), mul_rn(ba.lr, dh) / ba.bsz);
                if (ba.sparsity) m = add_rn(m, mul_rn(-ba.lr, add_rn(t[0] / ba.bsz, -ba.sp_target)));
                ba.hbm[c] = m;
                ba.hb[c] = add_rn(ba.hb[c], m);
            }
        }
        if (c < V) {
            const float dv = t[2] - t[3];
            out[H + c] = dv;
            if (ba.apply) {                               // rbm.py:223-224 / 480-481
                const float m = add_rn(mul_rn(ba.vbm[c], ba.mom), mul_rn(ba.lr, dv) / ba.bsz);
                ba.vbm[c] = m;
                ba.vb[c] = add_rn(ba.vb[c], m);
            }
        }
        float s = t[4];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (x == 0) {
            sq_part[blockIdx.x] = s;
            __threadfence();
            s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        __syncwarp();
        // the last block to finish adds the per-block squared errors in index order (deterministic)
        if (s_last) {
            __threadfence();
            float v = 0.f;
            for (int i = x; i < (int)gridDim.x; i += 32) v += __ldcg(sq_part + i);
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (x == 0) {
                out[2 * H + V] = v;
                if (ba.loss_out) *ba.loss_out = v / ba.n_loss;   // rbm.py:226 / 483
                *ticket = 0u;
            }
        }
    }
}

// ---- bias updates (rbm.py:216-224 / 478-481) and the loss (rbm.py:226 / 483).  Block 0 also adds the
// squared-error partials in index order (when sq_part != nullptr) and stores the total in st[2H+V].
__global__ void k_bias_update(float* __restrict__ st, const float* __restrict__ sq_part, int n_part, int V,
                              int H, float* __restrict__ hb, float* __restrict__ hbm, float* __restrict__ vb,
                              float* __restrict__ vbm, float lr, float mom, float bsz, int sparsity,
                              float sp_target, float n_loss, float* __restrict__ loss_out, int apply) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (apply) {
        if (c < H) {
            float m = add_rn(mul_rn(hbm[c], mom), mul_rn(lr, st[c]) / bsz);
            if (sparsity) m = add_rn(m, mul_rn(-lr, add_rn(st[H + V + c] / bsz, -sp_target)));
            hbm[c] = m;
            hb[c] = add_rn(hb[c], m);
        }
        if (c < V) {
            const float m = add_rn(mul_rn(vbm[c], mom), mul_rn(lr, st[H + c]) / bsz);
            vbm[c] = m;
            vb[c] = add_rn(vb[c], m);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        float total;
        if (sq_part) {
            float v = 0.f;
            for (int i = threadIdx.x; i < n_part; i += 32) v += sq_part[i];
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            total = v;
            if (threadIdx.x == 0) st[2 * H + V] = total;
        } else {
            total = st[2 * H + V];
        }
        if (threadIdx.x == 0 && loss_out) *loss_out = total / n_loss;
    }
}

// element-wise weight update from an (all-reduced) dS                          rbm.py:212-213
__global__ void k_weight_update(const float* __restrict__ dS, size_t n, float* __restrict__ W,
                                float* __restrict__ Wm, float lr, float mom, float wd, float bsz) {
    const float rb = 1.0f / bsz;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const float w = W[i];
        const float grad = add_rn(div_by(dS[i], bsz, rb), -mul_rn(wd, w));
        const float wm = add_rn(mul_rn(Wm[i], mom), mul_rn(lr, grad));
        Wm[i] = wm;
        W[i] = add_rn(w, wm);
    }
}

// v = a*(1-km) + known*km                                                      rbm.py:465
__global__ void k_clampmix(const float* __restrict__ a, const float* __restrict__ known,
                           const float* __restrict__ km, size_t n, float* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = clampmix(a[i], known[i], km[i]);
}

// ---- free energy finish (energy_utils.py:25-28): F[b] = -v.b_v - sum_j softplus(pre_j)
__global__ void k_free_energy(const float* __restrict__ part, int splits, SKPlan sk, const float* __restrict__ v,
                              int B, int V, int H, const float* __restrict__ hb,
                              const float* __restrict__ vb, float* __restrict__ F) {
    const int b = blockIdx.x;
    const size_t n = (size_t)B * H;
    float th = 0.0f, tv = 0.0f;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        const int ns = sk.k_iters ? sk_nslabs(sk, j / sk.tile_w) : splits;
        float x = add_rn(sum_slabs(part, ns, n, (size_t)b * H + j), hb[j]);
        th += (x > 20.0f) ? x : log1pf(expf(x));   // torch softplus, threshold 20
    }
    for (int c = threadIdx.x; c < V; c += blockDim.x) tv = fmaf(v[(size_t)b * V + c], vb[c], tv);
    __shared__ float r1[32], r2[32];
    for (int o = 16; o; o >>= 1) {
        th += __shfl_xor_sync(0xffffffffu, th, o);
        tv += __shfl_xor_sync(0xffffffffu, tv, o);
    }
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = th; r2[threadIdx.x >> 5] = tv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.0f, c = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += r1[w]; c += r2[w]; }
        F[b] = -c - a;
    }
}

// ---- best-of-K pick (imdbn.py:472-474): first minimum wins
__global__ void k_best_of_k(const float* __restrict__ cand, const float* __restrict__ F, int K, int B,
                            int V, float* __restrict__ out, int32_t* __restrict__ idx_out) {
    const int b = blockIdx.x;
    int best = 0;
    float fb = F[b];
    for (int k = 1; k < K; ++k) {
        const float f = F[(size_t)k * B + b];
        if (f < fb) { fb = f; best = k; }
    }
    for (int c = threadIdx.x; c < V; c += blockDim.x)
        out[(size_t)b * V + c] = cand[((size_t)best * B + b) * V + c];
    if (threadIdx.x == 0 && idx_out) idx_out[b] = best;
}

// ---- per-class sums (imdbn.py:249-251, 271-277): one thread per latent column, rows in order
__global__ void k_class_stats(const float* __restrict__ z, const float* __restrict__ y, int B, int Dz,
                              int K, float* __restrict__ sum_z, float* __restrict__ class_sum,
                              float* __restrict__ class_count, float* __restrict__ label_sum) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < Dz) {
        float s = 0.0f;
        for (int b = 0; b < B; ++b) {
            const float zv = z[(size_t)b * Dz + d];
            s += zv;
            int arg = 0;
            float best = y[(size_t)b * K];
            for (int k = 1; k < K; ++k) {
                const float yv = y[(size_t)b * K + k];
                if (yv > best) { best = yv; arg = k; }
            }
            class_sum[(size_t)arg * Dz + d] += zv;
        }
        sum_z[d] += s;
    }
    if (d < K) {
        float cnt = 0.0f, ls = 0.0f;
        for (int b = 0; b < B; ++b) {
            int arg = 0;
            float best = y[(size_t)b * K];
            for (int k = 1; k < K; ++k) {
                const float yv = y[(size_t)b * K + k];
                if (yv > best) { best = yv; arg = k; }
            }
            cnt += (arg == d) ? 1.0f : 0.0f;
            ls += y[(size_t)b * K + d];
        }
        class_count[d] += cnt;
        label_sum[d] += ls;
    }
}

// out = [a ; b] (float4 granularity)
__global__ void k_concat2(const float4* __restrict__ a, size_t na, const float4* __restrict__ b, size_t nb,
                          float4* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < na + nb;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = i < na ? a[i] : b[i - na];
}

__global__ void k_random_field(RngKey key, uint32_t draw, int kind, int rows, int cols,
                               float* __restrict__ out) {
    const size_t n = (size_t)rows * cols;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        out[i] = kind ? rf_normal(key, draw, r, c) : rf_uniform(key, draw, r, c);
    }
}

__global__ void k_transpose(const float* __restrict__ in, int R, int C, float* __restrict__ out) {
    __shared__ float t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        t[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * C + c] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) out[(size_t)c * R + r] = t[threadIdx.x][i];
    }
}

}  // namespace imdbn
