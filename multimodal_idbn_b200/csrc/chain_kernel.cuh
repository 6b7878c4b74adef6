// Persistent conditional-inference chains: the step loops of RBM.noisy_meanfield_annealed
// (imdbn/models/rbm.py:337-365) and RBM.conditional_gibbs (rbm.py:393-400) run INSIDE one kernel.
// A CTA owns R chains (rows); their visible / hidden state lives in shared memory for the whole
// chain; W (and a transposed copy, so both half-steps read coalesced) streams from L2.
#pragma once
#include "common.cuh"

namespace imdbn {

constexpr int CHAIN_THREADS = 256;
constexpr int CHAIN_MAX_STEPS = 4096;

struct ChainArgs {
    const float* W;    // [V,H]
    const float* Wt;   // [H,V]
    const float* hb; const float* vb;
    int V, H, B;
    Groups gr;
    int kind, n_steps;
    const float* v_known; const float* km; const float* v_init;
    const float* T; const float* sigma; const float* eta;   // device tables [n_steps]
    const float* mu; int Dz;
    int sample_h, sample_v, final_free;
    uint32_t draw0;
    float* v_out; float* vprob_out;
    RngKey key;
    int Vp, Hp;   // padded (multiple of 4) smem row lengths
};

// acc[r] += sum_k x[r][k] * w[k*ld]  (k sequential: the summation order of a plain dot product).
// 16 weight loads are in flight per thread: at small batch this loop is bound by L2 latency.
template <int R>
__device__ __forceinline__ void chain_dot(const float* __restrict__ wcol, int ld, int K, int Kp,
                                          const float* __restrict__ xs, float (&acc)[R]) {
    constexpr int KU = 16;
    int k = 0;
    for (; k + KU <= Kp; k += KU) {
        float w[KU];
#pragma unroll
        for (int q = 0; q < KU; ++q) w[q] = (k + q < K) ? __ldg(wcol + (size_t)(k + q) * ld) : 0.0f;
#pragma unroll
        for (int q4 = 0; q4 < KU; q4 += 4) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 x = *reinterpret_cast<const float4*>(xs + r * Kp + k + q4);
                acc[r] = fmaf(x.x, w[q4 + 0], acc[r]);
                acc[r] = fmaf(x.y, w[q4 + 1], acc[r]);
                acc[r] = fmaf(x.z, w[q4 + 2], acc[r]);
                acc[r] = fmaf(x.w, w[q4 + 3], acc[r]);
            }
        }
    }
    for (; k < Kp; k += 4) {
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = (k + q < K) ? __ldg(wcol + (size_t)(k + q) * ld) : 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(xs + r * Kp + k);
            acc[r] = fmaf(x.x, w[0], acc[r]);
            acc[r] = fmaf(x.y, w[1], acc[r]);
            acc[r] = fmaf(x.z, w[2], acc[r]);
            acc[r] = fmaf(x.w, w[3], acc[r]);
        }
    }
}

// raw[r][n] = sum_k xs[r][k] * w[k*ld + n] for n < N, K-SLICED: the thread block is arranged as
// (N/4 column quads) x (S k-slices); every thread streams float4 weights for its quad over its slice of
// k with 16 loads in flight, the S partial sums are added in slice order.  Used when few chains share a
// CTA (small batches), where one thread per column would serialise ~K/16 L2 round trips.
template <int R, int NT>
__device__ __forceinline__ void chain_matvec_sliced(const float* __restrict__ w, int ld, int K, int Kp, int N,
                                                    int Np, const float* __restrict__ xs, float* __restrict__ part,
                                                    float* __restrict__ raw) {
    const int tid = threadIdx.x;
    const int ncg = Np >> 2;
    int S = NT / ncg; S = S > 8 ? 8 : (S < 1 ? 1 : S);
    const int kper = (((Kp + S - 1) / S) + 3) & ~3;
    for (int base = 0; base < ncg; base += NT) {            // (ncg > NT only for very wide layers)
        const int t = tid;
        const int cg = base + (S > 1 ? t % ncg : t), sl = S > 1 ? t / ncg : 0;
        if (cg < ncg && sl < S) {
            float acc[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
            const int k0 = sl * kper, k1 = min(Kp, k0 + kper);
            const float* wp = w + 4 * cg;
            const bool col_ok = 4 * cg < N;
            for (int k = k0; k < k1; k += 16) {
                float4 wv[16];
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    wv[q] = (col_ok && k + q < k1 && k + q < K)
                                ? __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + q) * ld))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    if (k + q < k1) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float x = xs[r * Kp + k + q];
                            acc[r][0] = fmaf(x, wv[q].x, acc[r][0]);
                            acc[r][1] = fmaf(x, wv[q].y, acc[r][1]);
                            acc[r][2] = fmaf(x, wv[q].z, acc[r][2]);
                            acc[r][3] = fmaf(x, wv[q].w, acc[r][3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
                *reinterpret_cast<float4*>(part + ((size_t)sl * R + r) * Np + 4 * cg) =
                    make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
    }
    __syncthreads();
    for (int i = tid; i < R * Np; i += NT) {
        float v = 0.f;
        for (int sl = 0; sl < S; ++sl) v += part[(size_t)sl * R * Np + i];
        raw[i] = v;
    }
    __syncthreads();
}

template <int R, int NT, bool SLICED>
__global__ void __launch_bounds__(NT) k_chain(ChainArgs a) {
    constexpr int CHAIN_THREADS = NT;     // shadows the namespace constant inside this kernel
    extern __shared__ __align__(16) float smem[];
    float* vs = smem;                     // [R][Vp] current visible state
    float* hs = vs + R * a.Vp;            // [R][Hp] hidden state
    float* lg = hs + R * a.Hp;            // [R][Vp] visible logits / probabilities
    float* raw = lg + R * a.Vp;           // SLICED only: [R][max(Vp,Hp)] products, then [8][R][max] partials
    float* part = raw + R * max(a.Vp, a.Hp);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_base = blockIdx.x * R;
    const int nrows = min(R, a.B - row_base);
    const bool noisy = a.kind == IMDBN_CHAIN_NOISY_MF;

    // ---- init: v = v_known*km + (1-km)*U   (rbm.py:333,392), or the given start state
    for (int i = tid; i < R * a.Vp; i += CHAIN_THREADS) {
        const int r = i / a.Vp, c = i % a.Vp;
        float x = 0.0f;
        if (r < nrows && c < a.V) {
            const size_t o = (size_t)(row_base + r) * a.V + c;
            if (a.v_init) {
                x = a.v_init[o];
            } else {
                const float km = a.km[o];
                x = add_rn(mul_rn(a.v_known[o], km),
                           mul_rn(1.0f - km, rf_uniform(a.key, a.draw0, row_base + r, c)));
            }
        }
        vs[i] = x;
    }
    for (int i = tid; i < R * a.Hp; i += CHAIN_THREADS) hs[i] = 0.0f;
    __syncthreads();

    const int total = a.n_steps + (a.final_free ? 1 : 0);
    for (int t = 0; t < total; ++t) {
        const bool free_sweep = (t == a.n_steps);
        const bool last = (t == total - 1);
        const float T = (noisy && !free_sweep) ? a.T[t] : 1.0f;
        const float invT = 1.0f / T;
        const float sig = (noisy && !free_sweep) ? a.sigma[t] : 0.0f;
        uint32_t d_h, d_v, d_c;
        if (noisy) { d_h = a.draw0 + 1 + 2 * t; d_v = a.draw0 + 2 + 2 * t; d_c = 0; }
        else       { d_h = a.draw0 + 1 + 3 * t; d_v = a.draw0 + 2 + 3 * t; d_c = a.draw0 + 3 + 3 * t; }

        // ---- h | v : h = sigmoid((vW + hb)/T + sig*N)                          rbm.py:344-347,394
        if (SLICED) chain_matvec_sliced<R, NT>(a.W, a.H, a.V, a.Vp, a.H, a.Hp, vs, part, raw);
        for (int j = tid; j < a.Hp; j += CHAIN_THREADS) {
            if (j >= a.H) continue;
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = SLICED ? raw[r * a.Hp + j] : 0.0f;
            if (!SLICED) chain_dot<R>(a.W + j, a.H, a.V, a.Vp, vs, acc);
            const float bj = a.hb[j];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float x = div_by(add_rn(acc[r], bj), T, invT);
                if (sig > 0.0f) x = add_rn(x, mul_rn(rf_normal(a.key, d_h, row_base + r, j), sig));
                float p = sigmoidf_ref(x);
                if (a.sample_h && !free_sweep)
                    p = (p > rf_uniform(a.key, d_h, row_base + r, j)) ? 1.0f : 0.0f;
                hs[r * a.Hp + j] = p;
            }
        }
        __syncthreads();

        // ---- v | h : logits = (h W^T + vb)/T + sig*N                           rbm.py:350-352,396
        if (SLICED) chain_matvec_sliced<R, NT>(a.Wt, a.V, a.H, a.Hp, a.V, a.Vp, hs, part, raw);
        for (int c = tid; c < a.Vp; c += CHAIN_THREADS) {
            if (c >= a.V) continue;
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = SLICED ? raw[r * a.Vp + c] : 0.0f;
            if (!SLICED) chain_dot<R>(a.Wt + c, a.V, a.H, a.Hp, hs, acc);
            const float bc = a.vb[c];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float x = div_by(add_rn(acc[r], bc), T, invT);
                if (sig > 0.0f) x = add_rn(x, mul_rn(rf_normal(a.key, d_v, row_base + r, c), sig));
                lg[r * a.Vp + c] = x;
            }
        }
        __syncthreads();

        // ---- softmax groups: one warp per (row, group); the group's logits are replaced by
        //      probabilities and flagged so the sigmoid pass below leaves them alone.
        //      (rbm.py:113-114, 355-356)
        for (int w = warp; w < nrows * a.gr.n; w += CHAIN_THREADS / 32) {
            const int r = w / a.gr.n, g = w % a.gr.n;
            const int s = a.gr.s[g], e = a.gr.e[g];
            float* L = lg + r * a.Vp;
            float mx = -INFINITY;
            for (int c = s + lane; c < e; c += 32) mx = fmaxf(mx, L[c]);
            for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.0f;
            for (int c = s + lane; c < e; c += 32) sum += expf(L[c] - mx);
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            __syncwarp();
            const float rsum = 1.0f / sum;
            for (int c = s + lane; c < e; c += 32) L[c] = div_by(expf(L[c] - mx), sum, rsum);
        }
        __syncthreads();

        // ---- probabilities, mu-pull, (sampling,) re-clamp                      rbm.py:354-365,397-399
        const float eta = (noisy && a.mu && !free_sweep) ? a.eta[t] : 0.0f;
        for (int i = tid; i < nrows * a.Vp; i += CHAIN_THREADS) {
            const int r = i / a.Vp, c = i % a.Vp;
            if (c >= a.V) continue;
            bool in_group = false;
            for (int g = 0; g < a.gr.n; ++g) in_group |= (c >= a.gr.s[g] && c < a.gr.e[g]);
            float p = lg[i];
            if (!in_group) p = sigmoidf_ref(p);
            const size_t o = (size_t)(row_base + r) * a.V + c;
            if (noisy && a.mu && !free_sweep && c < a.Dz)
                p = add_rn(mul_rn(1.0f - eta, p), mul_rn(eta, a.mu[(size_t)(row_base + r) * a.Dz + c]));
            lg[i] = p;   // un-clamped probabilities of this sweep
            if (last && a.vprob_out) a.vprob_out[o] = p;
            float x = p;
            if (a.sample_v && !free_sweep && !in_group)
                x = (p > rf_uniform(a.key, d_v, row_base + r, c)) ? 1.0f : 0.0f;
            if (!free_sweep && !(a.sample_v && in_group)) x = clampmix(x, a.v_known[o], a.km[o]);
            vs[i] = x;
        }
        if (a.sample_v && !free_sweep && a.gr.n > 0) {
            __syncthreads();
            // categorical one-hot per (row, group) from the un-clamped probabilities, then re-clamp
            for (int w = tid; w < nrows * a.gr.n; w += CHAIN_THREADS) {
                const int r = w / a.gr.n, g = w % a.gr.n;
                const int s = a.gr.s[g], e = a.gr.e[g];
                const float* P = lg + r * a.Vp;
                float tot = 0.0f;
                for (int c = s; c < e; ++c) tot += fminf(fmaxf(P[c], 1e-8f), 1.0f);
                const float u = rf_uniform(a.key, d_c, row_base + r, g);
                float cdf = 0.0f;
                int idx = 0;
                const float rtot = 1.0f / tot;
                for (int c = s; c < e; ++c) {
                    cdf += div_by(fminf(fmaxf(P[c], 1e-8f), 1.0f), tot, rtot);
                    idx += (cdf <= u) ? 1 : 0;
                }
                idx = min(idx, e - s - 1);
                for (int c = s; c < e; ++c) {
                    const size_t o = (size_t)(row_base + r) * a.V + c;
                    vs[r * a.Vp + c] = clampmix((c - s == idx) ? 1.0f : 0.0f, a.v_known[o], a.km[o]);
                }
            }
        }
        __syncthreads();
    }

    for (int i = tid; i < nrows * a.Vp; i += CHAIN_THREADS) {
        const int r = i / a.Vp, c = i % a.Vp;
        if (c < a.V) a.v_out[(size_t)(row_base + r) * a.V + c] = vs[i];
    }
}

}  // namespace imdbn
