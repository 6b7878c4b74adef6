// One persistent kernel for a whole CD-k update of a SMALL RBM (rbm.py:180-227) plus the post-update
// forward of idbn.py:203 -- the upper layers of an iDBN (1500 -> 500 in the C2 workload: 3 MB of weights,
// 0.5 GFLOP per step).  Launched as ~14 separate kernels such a layer costs ~70 us of pure launch / drain
// latency for ~5 us of work; here the passes are phases of ONE grid separated by grid-wide barriers, with
// the weights, the activations and the split-K partial sums staying in L2.
//
//   F0  h0 = 1[pos_h > U]                                     (or P0/F0: positive phase when not cached)
//   per CD step:  P1 down GEMM -> F1 sigmoid + Bernoulli -> P2 up GEMM -> F2 sigmoid (+ Bernoulli)
//   P3  dS = v+^T h+ - v-^T h-, momentum / weight-decay update of W, W_m; column statistics, bias update
//   P4  forward GEMM over [data ; next_data] with the UPDATED weights -> F4 sigmoid (+ loss)
//
// GEMM phases: work unit = (64-feature output tile, k range <= 96); 256 threads = 4 k-groups x 64 threads,
// 8 x 8 outputs per thread, fp32 FFMA from shared memory (operands are staged with an XOR-swizzled 16-byte
// slot layout, conflict-free for both the staging stores and the k-vectorised reads); the four k-groups
// are added in fixed order and the unit's partial tile is written to its split-K slab; the following
// finish phase adds the slabs in index order.  Everything is deterministic and exact fp32 (no tensor-core
// rounding), so both precision modes use it.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace imdbn {

constexpr int CDS_THREADS = 256;
constexpr int CDS_T = 64;                         // tile edge (batch rows, output features)
constexpr int CDS_KMAX = 96;                      // k range of one unit
constexpr int CDS_SLOTS = CDS_KMAX / 4;           // 16-byte slots per operand row (24: closed under ^7)
constexpr int CDS_MAXROWS = 128;                  // [data ; next_data]
constexpr int CDS_SM_A = CDS_MAXROWS * CDS_KMAX;  // floats
constexpr int CDS_SM_B = CDS_T * CDS_KMAX;
constexpr int CDS_SM_RED = 4 * CDS_T * CDS_T;
constexpr int CDS_SMEM_BYTES = (CDS_SM_A + CDS_SM_B + CDS_SM_RED) * 4;   // 136 KB

struct CdsPlan { int nt, ksplit, krange; };        // units = nt * ksplit

struct CdSmallArgs {
    float *W, *Wm, *hb, *hbm, *vb, *vbm;
    int V, H;
    const float* data; int B;
    const float* pos_h_in;                 // nullable: cached positive phase
    const float* next_data; int B_next;    // nullable
    float* fwd_out;                        // nullable [B + B_next, H]
    int k;
    float lr, mom, wd, bsz; int sparsity; float sp_target;
    float* loss_out;
    RngKey key;
    CdsPlan up, dn, fw;                    // partitions of the up / down / forward passes
    float* part;                           // split-K slabs
    float *pos_h, *h_s, *h_prob, *v_prob, *v_s;
    float* sq_part;                        // per-CTA squared-error partials
    unsigned int* bar;                     // grid barrier: [0] arrivals, [1] generation
    unsigned long long* trace;             // nullable: phase timestamps of CTA 0 (IMDBN_CDS_TRACE)
};

__device__ __forceinline__ void cds_grid_sync(unsigned int* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int gen = *reinterpret_cast<volatile unsigned int*>(bar + 1);
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            atomicExch(bar, 0u);
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile unsigned int*>(bar + 1) == gen) {
                if (clock64() - t0 > 4000000000LL) __trap();      // a protocol bug must fault, never hang
            }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// stage a [rows x kr] slice of a row-major matrix (leading dimension ld) as S[row][slot ^ (row>>3 & 7)].
// All loads of a batch are issued before the first store (six independent 16-byte L2 loads in flight per
// thread instead of one latency-bound load per loop trip).
__device__ __forceinline__ void cds_stage_rows(float* S, const float* __restrict__ G, int ld, int rows, int rows_pad,
                                               int k0, int kr, int kend) {
    const int kr4 = (kr + 3) >> 2, total = rows_pad * kr4;
    for (int base = 0; base < total; base += 6 * CDS_THREADS) {
        float4 v[6];
        int dst[6];
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int idx = base + threadIdx.x + u * CDS_THREADS;
            const int row = idx / kr4, q = idx - row * kr4;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            dst[u] = idx < total ? row * CDS_KMAX + 4 * (q ^ ((row >> 3) & 7)) : -1;
            if (idx < total && row < rows && k0 + 4 * q < kend) v[u] = ldcg4(G + (size_t)row * ld + k0 + 4 * q);
        }
#pragma unroll
        for (int u = 0; u < 6; ++u)
            if (dst[u] >= 0) *reinterpret_cast<float4*>(S + dst[u]) = v[u];
    }
}

// the four k-groups' accumulators -> red, then summed in group order and handed to `emit(row, col4, value)`
template <typename Emit>
__device__ __forceinline__ void cds_reduce_emit(float* red, const float (&acc)[8][8], Emit emit) {
    const int g = threadIdx.x >> 6, t = threadIdx.x & 63, bb = t >> 3, fb = t & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float* r = red + g * (CDS_T * CDS_T) + (bb * 8 + i) * CDS_T + fb * 8;
        *reinterpret_cast<float4*>(r) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(r + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int o4 = threadIdx.x + CDS_THREADS * u, row = o4 >> 4, c4 = o4 & 15;
        const float* r = red + row * CDS_T + c4 * 4;
        const float4 a = *reinterpret_cast<const float4*>(r);
        const float4 b = *reinterpret_cast<const float4*>(r + CDS_T * CDS_T);
        const float4 c = *reinterpret_cast<const float4*>(r + 2 * CDS_T * CDS_T);
        const float4 d = *reinterpret_cast<const float4*>(r + 3 * CDS_T * CDS_T);
        emit(row, c4, a, b, c, d);
    }
    __syncthreads();
}

// ---- up-type GEMM phase: part[s][row][n] = sum_{k in range s} A[row][k] W[k][n]          rbm.py:92
__device__ void cds_gemm_up(const CdSmallArgs& a, const CdsPlan& pl, const float* __restrict__ A, int rows,
                            float* sm) {
    float* As = sm; float* Ws = sm + CDS_SM_A; float* red = Ws + CDS_SM_B;
    const int K = a.V, N = a.H;
    const int g = threadIdx.x >> 6, t = threadIdx.x & 63, bb = t >> 3, fb = t & 7;
    const int halves = (rows + CDS_T - 1) / CDS_T;
    for (int unit = blockIdx.x; unit < pl.nt * pl.ksplit; unit += gridDim.x) {
        const int nt = unit % pl.nt, s = unit / pl.nt;
        const int n0 = nt * CDS_T, k0 = s * pl.krange;
        const int kr = min(pl.krange, K - k0);
        if (kr <= 0) continue;                        // (uniform per CTA; slab stays unwritten and unread)
        cds_stage_rows(As, A, K, rows, halves * CDS_T, k0, kr, K);
        {
            float4 v[6];
#pragma unroll
            for (int u = 0; u < 6; ++u) {                                  // kr * 16 <= 1536 = 6 * 256
                const int idx = threadIdx.x + u * CDS_THREADS, k = idx >> 4, q = idx & 15;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < kr && n0 + 4 * q < N) v[u] = ldcg4(a.W + (size_t)(k0 + k) * N + n0 + 4 * q);
            }
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int idx = threadIdx.x + u * CDS_THREADS;
                if ((idx >> 4) < kr) *reinterpret_cast<float4*>(Ws + (idx >> 4) * CDS_T + 4 * (idx & 15)) = v[u];
            }
        }
        __syncthreads();
        const int kr4 = (kr + 3) >> 2, kq = (kr4 + 3) >> 2;
        const int q_beg = g * kq, q_end = min(kr4, q_beg + kq);
        for (int hf = 0; hf < halves; ++hf) {
            float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
            const int rbase = hf * CDS_T + bb * 8;
            const int sw = ((rbase >> 3) & 7);
            for (int q = q_beg; q < q_end; ++q) {
                float4 av[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    av[i] = *reinterpret_cast<const float4*>(As + (rbase + i) * CDS_KMAX + 4 * (q ^ sw));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = 4 * q + e;
                    if (k < kr) {
                        const float4 w0 = *reinterpret_cast<const float4*>(Ws + k * CDS_T + fb * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(Ws + k * CDS_T + fb * 8 + 4);
                        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float x = e == 0 ? av[i].x : e == 1 ? av[i].y : e == 2 ? av[i].z : av[i].w;
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(x, w[j], acc[i][j]);
                        }
                    }
                }
            }
            float* slab = a.part + (size_t)s * rows * N;
            cds_reduce_emit(red, acc, [&](int row, int c4, float4 p, float4 q4, float4 r4, float4 s4) {
                const int gr = hf * CDS_T + row, col = n0 + 4 * c4;
                if (gr < rows && col < N)
                    *reinterpret_cast<float4*>(slab + (size_t)gr * N + col) =
                        make_float4(((p.x + q4.x) + r4.x) + s4.x, ((p.y + q4.y) + r4.y) + s4.y,
                                    ((p.z + q4.z) + r4.z) + s4.z, ((p.w + q4.w) + r4.w) + s4.w);
            });
        }
    }
}

// ---- down GEMM phase: part[s][row][c] = sum_{j in range s} Hs[row][j] W[c][j]            rbm.py:96
__device__ void cds_gemm_down(const CdSmallArgs& a, const CdsPlan& pl, const float* __restrict__ Hs, int rows,
                              float* sm) {
    float* As = sm; float* Bs = sm + CDS_SM_A; float* red = Bs + CDS_SM_B;
    const int K = a.H, N = a.V;
    const int g = threadIdx.x >> 6, t = threadIdx.x & 63, bb = t >> 3, fb = t & 7;
    for (int unit = blockIdx.x; unit < pl.nt * pl.ksplit; unit += gridDim.x) {
        const int nt = unit % pl.nt, s = unit / pl.nt;
        const int n0 = nt * CDS_T, k0 = s * pl.krange;
        const int kr = min(pl.krange, K - k0);
        if (kr <= 0) continue;
        cds_stage_rows(As, Hs, K, rows, CDS_T, k0, kr, K);
        cds_stage_rows(Bs, a.W + (size_t)n0 * K, K, min(CDS_T, N - n0), CDS_T, k0, kr, K);
        __syncthreads();
        const int kr4 = (kr + 3) >> 2, kq = (kr4 + 3) >> 2;
        const int q_beg = g * kq, q_end = min(kr4, q_beg + kq);
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        const int swa = bb & 7, swb = fb & 7;
        for (int q = q_beg; q < q_end; ++q) {
            float4 av[8], bv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                av[i] = *reinterpret_cast<const float4*>(As + (bb * 8 + i) * CDS_KMAX + 4 * (q ^ swa));
                bv[i] = *reinterpret_cast<const float4*>(Bs + (fb * 8 + i) * CDS_KMAX + 4 * (q ^ swb));
            }
            // (slots past kr are zero-filled by the staging, so no k guard is needed)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
                }
        }
        float* slab = a.part + (size_t)s * rows * N;
        cds_reduce_emit(red, acc, [&](int row, int c4, float4 p, float4 q4, float4 r4, float4 s4) {
            const int col = n0 + 4 * c4;
            if (row < rows && col < N)
                *reinterpret_cast<float4*>(slab + (size_t)row * N + col) =
                    make_float4(((p.x + q4.x) + r4.x) + s4.x, ((p.y + q4.y) + r4.y) + s4.y,
                                ((p.z + q4.z) + r4.z) + s4.z, ((p.w + q4.w) + r4.w) + s4.w);
        });
    }
}

// ---- finish phase: p = sigmoid(sum_s part[s] + bias), optional Bernoulli sample          rbm.py:92,110,125,175
__device__ void cds_finish(const CdSmallArgs& a, int nslab, int rows, int N, const float* __restrict__ bias,
                           float* __restrict__ p_out, float* __restrict__ s_out, uint32_t draw) {
    const int n4 = N >> 2, total = rows * n4;
    const size_t stride = (size_t)rows * N;
    for (int idx = blockIdx.x * CDS_THREADS + threadIdx.x; idx < total; idx += gridDim.x * CDS_THREADS) {
        const int row = idx / n4, c = (idx - row * n4) << 2;
        const size_t o = (size_t)row * N + c;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < nslab; s0 += 6) {                            // six slab loads in flight
            float4 v[6];
#pragma unroll
            for (int u = 0; u < 6; ++u)
                v[u] = s0 + u < nslab ? ldcg4(a.part + (s0 + u) * stride + o) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 6; ++u)
                if (s0 + u < nslab) { x.x += v[u].x; x.y += v[u].y; x.z += v[u].z; x.w += v[u].w; }
        }
        const float4 b = ldcg4(bias + c);
        float4 p;
        p.x = sigmoidf_ref(add_rn(x.x, b.x)); p.y = sigmoidf_ref(add_rn(x.y, b.y));
        p.z = sigmoidf_ref(add_rn(x.z, b.z)); p.w = sigmoidf_ref(add_rn(x.w, b.w));
        if (p_out) *reinterpret_cast<float4*>(p_out + o) = p;
        if (s_out) {
            const float4 u4 = rf_uniform4(a.key, draw, row, c);     // one Philox call per quad
            float4 sv;
            sv.x = p.x > u4.x ? 1.f : 0.f;
            sv.y = p.y > u4.y ? 1.f : 0.f;
            sv.z = p.z > u4.z ? 1.f : 0.f;
            sv.w = p.w > u4.w ? 1.f : 0.f;
            *reinterpret_cast<float4*>(s_out + o) = sv;
        }
    }
}

// h0 = 1[pos_h > U] from cached probabilities                                             rbm.py:203
__device__ void cds_bernoulli(const CdSmallArgs& a, const float* __restrict__ p, int rows, int N,
                              float* __restrict__ s_out, uint32_t draw) {
    const int total = rows * N;
    for (int idx = blockIdx.x * CDS_THREADS + threadIdx.x; idx < total; idx += gridDim.x * CDS_THREADS) {
        const int row = idx / N, c = idx - row * N;
        s_out[idx] = __ldcg(p + idx) > rf_uniform(a.key, draw, row, c) ? 1.f : 0.f;
    }
}

// ---- statistics + update phase: 64 x 64 tiles of W                                       rbm.py:200,209,212-213
__device__ void cds_stats_update(const CdSmallArgs& a, const float* __restrict__ pos_h, float* sm) {
    float* Vp = sm;                       // [64 k][64 m]
    float* Vn = Vp + CDS_T * CDS_T;
    float* Hp = Vn + CDS_T * CDS_T;       // [64 k][64 n]
    float* Hn = Hp + CDS_T * CDS_T;
    float* red = sm + CDS_SM_A + CDS_SM_B;
    const int V = a.V, H = a.H, B = a.B;
    const int mt = (V + CDS_T - 1) / CDS_T, ntl = (H + CDS_T - 1) / CDS_T;
    const int g = threadIdx.x >> 6, t = threadIdx.x & 63, bb = t >> 3, fb = t & 7;
    const float rb = 1.0f / a.bsz;
    for (int tile = blockIdx.x; tile < mt * ntl; tile += gridDim.x) {
        const int m0 = (tile / ntl) * CDS_T, n0 = (tile % ntl) * CDS_T;
        {
            float4 ld[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {                                  // 64 * 16 = 4 * 256 slots
                const int idx = threadIdx.x + u * CDS_THREADS, k = idx >> 4, q = idx & 15;
#pragma unroll
                for (int w = 0; w < 4; ++w) ld[u][w] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < B) {
                    if (m0 + 4 * q < V) {
                        ld[u][0] = ldcg4(a.data + (size_t)k * V + m0 + 4 * q);
                        ld[u][1] = ldcg4(a.v_s + (size_t)k * V + m0 + 4 * q);
                    }
                    if (n0 + 4 * q < H) {
                        ld[u][2] = ldcg4(pos_h + (size_t)k * H + n0 + 4 * q);
                        ld[u][3] = ldcg4(a.h_prob + (size_t)k * H + n0 + 4 * q);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = threadIdx.x + u * CDS_THREADS, o = (idx >> 4) * CDS_T + 4 * (idx & 15);
                *reinterpret_cast<float4*>(Vp + o) = ld[u][0];
                *reinterpret_cast<float4*>(Vn + o) = ld[u][1];
                *reinterpret_cast<float4*>(Hp + o) = ld[u][2];
                *reinterpret_cast<float4*>(Hn + o) = ld[u][3];
            }
        }
        // the W / W_m values this thread will update (positions of cds_reduce_emit) are fetched now and
        // arrive while the products are computed
        float4 w_pre[4], m_pre[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int o4 = threadIdx.x + CDS_THREADS * u, m = m0 + (o4 >> 4), n = n0 + 4 * (o4 & 15);
            w_pre[u] = m_pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < V && n < H) { w_pre[u] = ldcg4(a.W + (size_t)m * H + n); m_pre[u] = ldcg4(a.Wm + (size_t)m * H + n); }
        }
        __syncthreads();
        // k-groups 0,1: positive phase rows 0..31 / 32..63; groups 2,3: negative phase
        const float* Ak = (g < 2 ? Vp : Vn) + (g & 1) * 32 * CDS_T;
        const float* Bk = (g < 2 ? Hp : Hn) + (g & 1) * 32 * CDS_T;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(Ak + k * CDS_T + bb * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(Ak + k * CDS_T + bb * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(Bk + k * CDS_T + fb * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(Bk + k * CDS_T + fb * 8 + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        cds_reduce_emit(red, acc, [&](int row, int c4, float4 p0, float4 p1, float4 q0, float4 q1) {
            const int m = m0 + row, n = n0 + 4 * c4;
            if (m >= V || n >= H) return;
            const size_t o = (size_t)m * H + n;
            const int u = (row * 16 + c4 - (int)threadIdx.x) / CDS_THREADS;
            const float4 w4 = u == 0 ? w_pre[0] : u == 1 ? w_pre[1] : u == 2 ? w_pre[2] : w_pre[3];
            const float4 m4 = u == 0 ? m_pre[0] : u == 1 ? m_pre[1] : u == 2 ? m_pre[2] : m_pre[3];
            const float ds[4] = {(p0.x + p1.x) - (q0.x + q1.x), (p0.y + p1.y) - (q0.y + q1.y),
                                 (p0.z + p1.z) - (q0.z + q1.z), (p0.w + p1.w) - (q0.w + q1.w)};
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w};
            float nw[4], nm[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float grad = add_rn(div_by(ds[e], a.bsz, rb), -mul_rn(a.wd, wv[e]));
                nm[e] = add_rn(mul_rn(mv[e], a.mom), mul_rn(a.lr, grad));
                nw[e] = add_rn(wv[e], nm[e]);
            }
            *reinterpret_cast<float4*>(a.Wm + o) = make_float4(nm[0], nm[1], nm[2], nm[3]);
            *reinterpret_cast<float4*>(a.W + o) = make_float4(nw[0], nw[1], nw[2], nw[3]);
        });
    }
}

// ---- column statistics, bias updates, squared-error partials                             rbm.py:216-226
//      (32 columns x 8 row lanes per CTA pass, lane sums added in lane order)
__device__ void cds_colstats(const CdSmallArgs& a, const float* __restrict__ pos_h, float* sm) {
    float (*red)[8][32] = reinterpret_cast<float (*)[8][32]>(sm);       // [5][8][32]
    const int V = a.V, H = a.H, B = a.B;
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int nblk = (max(V, H) + 31) / 32;
    const float rb = 1.0f / a.bsz;
    float sq_cta = 0.f;
    // blocks are dealt from the END of the grid: the CTAs with one weight tile fewer take them
    for (int blk = gridDim.x - 1 - blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int c = blk * 32 + x;
        float ha = 0.f, hn = 0.f, va = 0.f, vn = 0.f, sq = 0.f;
        if (c < H)
            for (int b = y; b < B; b += 8) { ha += __ldcg(pos_h + (size_t)b * H + c); hn += __ldcg(a.h_prob + (size_t)b * H + c); }
        if (c < V)
            for (int b = y; b < B; b += 8) {
                const size_t o = (size_t)b * V + c;
                const float d0 = __ldcg(a.data + o);
                va += d0; vn += __ldcg(a.v_s + o);
                const float d = d0 - __ldcg(a.v_prob + o);
                sq = fmaf(d, d, sq);
            }
        red[0][y][x] = ha; red[1][y][x] = hn; red[2][y][x] = va; red[3][y][x] = vn; red[4][y][x] = sq;
        __syncthreads();
        if (y == 0) {
            float tt[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < 8; ++r) s += red[q][r][x];
                tt[q] = s;
            }
            if (c < H) {                                                   // rbm.py:216-220
                const float dh = tt[0] - tt[1];
                float m = add_rn(mul_rn(a.hbm[c], a.mom), div_by(mul_rn(a.lr, dh), a.bsz, rb));
                if (a.sparsity) m = add_rn(m, mul_rn(-a.lr, add_rn(div_by(tt[0], a.bsz, rb), -a.sp_target)));
                a.hbm[c] = m;
                a.hb[c] = add_rn(a.hb[c], m);
            }
            if (c < V) {                                                   // rbm.py:223-224
                const float dv = tt[2] - tt[3];
                const float m = add_rn(mul_rn(a.vbm[c], a.mom), div_by(mul_rn(a.lr, dv), a.bsz, rb));
                a.vbm[c] = m;
                a.vb[c] = add_rn(a.vb[c], m);
            }
            float s = tt[4];
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            sq_cta += s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) a.sq_part[blockIdx.x] = sq_cta;
}

__device__ __forceinline__ void cds_mark(const CdSmallArgs& a, int& i) {
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[i] = t;
    }
    ++i;
}

__global__ void __launch_bounds__(CDS_THREADS, 1) k_cd_small(CdSmallArgs a) {
    extern __shared__ __align__(16) float cds_sm[];
    pdl_wait();
    pdl_trigger();
    int ti = 0;
    cds_mark(a, ti);
    const int B = a.B, V = a.V, H = a.H;
    const float* pos_h = a.pos_h_in;
    if (!pos_h) {                                                          // rbm.py:199
        cds_gemm_up(a, a.up, a.data, B, cds_sm);
        cds_grid_sync(a.bar);
        cds_mark(a, ti);
        cds_finish(a, a.up.ksplit, B, H, a.hb, a.pos_h, a.h_s, 0);         // + rbm.py:203
        pos_h = a.pos_h;
    } else {
        cds_bernoulli(a, pos_h, B, H, a.h_s, 0);                           // rbm.py:203
    }
    cds_mark(a, ti);
    cds_grid_sync(a.bar);
    cds_mark(a, ti);
    for (int s = 0; s < a.k; ++s) {
        cds_gemm_down(a, a.dn, a.h_s, B, cds_sm);                          // rbm.py:205
        cds_mark(a, ti);
        cds_grid_sync(a.bar);
        cds_mark(a, ti);
        cds_finish(a, a.dn.ksplit, B, V, a.vb, a.v_prob, a.v_s, 1 + 3 * s);  // rbm.py:205-206
        cds_mark(a, ti);
        cds_grid_sync(a.bar);
        cds_mark(a, ti);
        cds_gemm_up(a, a.up, a.v_s, B, cds_sm);                            // rbm.py:207
        cds_mark(a, ti);
        cds_grid_sync(a.bar);
        cds_mark(a, ti);
        cds_finish(a, a.up.ksplit, B, H, a.hb, a.h_prob, s + 1 < a.k ? a.h_s : nullptr, 3 + 3 * s);   // :207-208
        cds_mark(a, ti);
        cds_grid_sync(a.bar);
        cds_mark(a, ti);
    }
    cds_stats_update(a, pos_h, cds_sm);
    cds_mark(a, ti);
    cds_colstats(a, pos_h, cds_sm);
    cds_mark(a, ti);
    cds_grid_sync(a.bar);
    cds_mark(a, ti);
    if (blockIdx.x == 0 && threadIdx.x < 32) {                             // rbm.py:226
        float v = 0.f;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) v += __ldcg(a.sq_part + i);
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0 && a.loss_out) *a.loss_out = v / (a.bsz * (float)V);
    }
    if (!a.fwd_out) return;
    const int Bt = B + (a.next_data ? a.B_next : 0);
    // [data ; next_data] is staged from two sources by row index: give the GEMM a virtual concatenation
    if (a.next_data) {
        // rows >= B come from next_data: stage through a small indirection by running the pass twice
        // over disjoint slab row ranges (the slabs are laid out [s][Bt][H]).
        CdSmallArgs b = a;
        cds_gemm_up(a, a.fw, a.data, B, cds_sm);               // rows 0..B-1 of every slab (slab stride = B*H)
        b.part = a.part + (size_t)a.fw.ksplit * B * H;
        cds_mark(a, ti);
        cds_gemm_up(b, a.fw, a.next_data, a.B_next, cds_sm);   // second slab set, stride B_next*H
        cds_mark(a, ti);
        cds_grid_sync(a.bar);
        cds_mark(a, ti);
        cds_finish(a, a.fw.ksplit, B, H, a.hb, a.fwd_out, nullptr, 0);
        cds_finish(b, a.fw.ksplit, a.B_next, H, a.hb, a.fwd_out + (size_t)B * H, nullptr, 0);
        cds_mark(a, ti);
    } else {
        cds_gemm_up(a, a.fw, a.data, B, cds_sm);
        cds_grid_sync(a.bar);
        cds_finish(a, a.fw.ksplit, Bt, H, a.hb, a.fwd_out, nullptr, 0);
    }
}

// host: partition of a pass with output width N and contraction length K over G CTAs
inline CdsPlan cds_plan(int N, int K, int G) {
    CdsPlan p;
    p.nt = (N + CDS_T - 1) / CDS_T;
    int ks = std::max(1, G / p.nt);
    ks = std::max(ks, (K + CDS_KMAX - 1) / CDS_KMAX);
    ks = std::min(ks, std::max(1, (K + 15) / 16));
    int kr = (K + ks - 1) / ks;
    kr = (kr + 3) & ~3;
    if (kr > CDS_KMAX) kr = CDS_KMAX;
    p.krange = kr;
    p.ksplit = (K + kr - 1) / kr;
    return p;
}

}  // namespace imdbn
