// Local attention over packed pair tiles, fp32 SIMT engine (g_update = True, "SCANN+").
//
// Restates LocalAttention.call (scann/layers/attention.py:118-216) for v_proj=False,
// kq_proj=True on the tile-padded pair layout built by plan.cu.  Per tile of <=128 pairs:
//
//   pre = xw1[c] + g @ W2 + xw3[j]            (exact split of the 384->128 filter_geo Dense,
//                                              attention.py:142-151; xw1 = x@W1+bf, xw3 = x@W3)
//   g'  = LN_g(swish(pre) + g)                 (attention.py:153)
//   k   = (x[j] * g') @ Wk + bk                (attention.py:157,163)
//   e   = 0.25 <q_h, k_h> ; p = softmax_n(e)   (attention.py:180-189; masked slots are absent)
//   out = LN(sum_n p k + q)                    (attention.py:206-214)
//
// One CTA (256 threads) per tile; each thread owns an 8x8 register block of the two
// 128x128x128 tile GEMMs: rows {ty + 16 i}, columns {4 tx + j} U {64 + 4 tx + j}.
#include "common.cuh"

#define LA_LDS 132                       // padded row stride of the activation tiles in smem
#define LA_THREADS 256

struct LaFwdArgs {
    const int32_t* ntiles; const int32_t* tile_a0; const int32_t* tile_a1;
    const int32_t* cnt; const int32_t* rowptr; const int32_t* pair_c; const int32_t* pair_j;
    const float* x;        // [R,128]  layer input
    const float* proj;     // [R,384]  [x@W1+bf | x@W3 | x@Wq+bq]
    const float* g_in;     // [rows,128]
    const float* W2; const float* Wk; const float* bk;
    const float* gamma_g; const float* beta_g; const float* gamma; const float* beta;
    float* g_out;          // [rows,128]
    float* ctx_pre;        // [R,128] pre-LayerNorm context (nullable; saved for backward)
    float* out;            // [R,128]
    float* attn;           // [rows,8] softmax weights (nullable)
};

__device__ __forceinline__ float f4c(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

// acc[i][*] += sum_k As[ty+16i][k] * B[k][cols]
template <bool B_GLOBAL>
__device__ __forceinline__ void gemm_nn(const float* __restrict__ As, const float* __restrict__ B, float (&acc)[8][8],
                                        int ty, int tx) {
#pragma unroll 1
    for (int k0 = 0; k0 < SCANN_D; k0 += 4) {
        float4 av[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = ld4(As + (ty + 16 * i) * LA_LDS + k0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float* bp = B + (k0 + kk) * SCANN_D + tx * 4;
            float4 b0 = B_GLOBAL ? ldg4(bp) : ld4(bp);
            float4 b1 = B_GLOBAL ? ldg4(bp + 64) : ld4(bp + 64);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float xv = f4c(av[i], kk);
                acc[i][0] = fmaf(xv, b0.x, acc[i][0]); acc[i][1] = fmaf(xv, b0.y, acc[i][1]);
                acc[i][2] = fmaf(xv, b0.z, acc[i][2]); acc[i][3] = fmaf(xv, b0.w, acc[i][3]);
                acc[i][4] = fmaf(xv, b1.x, acc[i][4]); acc[i][5] = fmaf(xv, b1.y, acc[i][5]);
                acc[i][6] = fmaf(xv, b1.z, acc[i][6]); acc[i][7] = fmaf(xv, b1.w, acc[i][7]);
            }
        }
    }
}

// acc[i][*] = sum_r As[r][ty*8+i] * Bs[r][cols]     (both operands in smem, reduction over rows)
__device__ __forceinline__ void gemm_tn(const float* __restrict__ As, const float* __restrict__ Bs,
                                        float (&acc)[8][8], int ty, int tx) {
#pragma unroll 2
    for (int r = 0; r < SCANN_TILE; ++r) {
        float4 a0 = ld4(As + r * LA_LDS + ty * 8), a1 = ld4(As + r * LA_LDS + ty * 8 + 4);
        float4 b0 = ld4(Bs + r * LA_LDS + tx * 4), b1 = ld4(Bs + r * LA_LDS + 64 + tx * 4);
        float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaf(av[i], b0.x, acc[i][0]); acc[i][1] = fmaf(av[i], b0.y, acc[i][1]);
            acc[i][2] = fmaf(av[i], b0.z, acc[i][2]); acc[i][3] = fmaf(av[i], b0.w, acc[i][3]);
            acc[i][4] = fmaf(av[i], b1.x, acc[i][4]); acc[i][5] = fmaf(av[i], b1.y, acc[i][5]);
            acc[i][6] = fmaf(av[i], b1.z, acc[i][6]); acc[i][7] = fmaf(av[i], b1.w, acc[i][7]);
        }
    }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

__device__ __forceinline__ int col_of(int tx, int j) { return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

__device__ __forceinline__ void load_tile(float* S, const float* __restrict__ src, size_t rowbase, int tid) {
    float4 v[16];                       // all 16 loads in flight before the first store
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        int i = tid + it * LA_THREADS;
        v[it] = ld4(src + (rowbase + (i >> 5)) * SCANN_D + (i & 31) * 4);
    }
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        int i = tid + it * LA_THREADS;
        st4(S + (i >> 5) * LA_LDS + (i & 31) * 4, v[it]);
    }
}

__device__ __forceinline__ void load8(const float* p, int tx, float (&v)[8]) {
    float4 a = ld4(p + tx * 4), b = ld4(p + 64 + tx * 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, int tx, const float (&v)[8]) {
    st4(p + tx * 4, make_float4(v[0], v[1], v[2], v[3]));
    st4(p + 64 + tx * 4, make_float4(v[4], v[5], v[6], v[7]));
}

// =============================================================================================
// Forward
// =============================================================================================
__global__ void __launch_bounds__(LA_THREADS, 1) la_fwd_simt_kernel(const LaFwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* S0 = smem;                                   // g tile -> a tile -> k tile
    float* W2s = S0 + SCANN_TILE * LA_LDS;
    float* Wks = W2s + SCANN_D * SCANN_D;
    float* Es = Wks + SCANN_D * SCANN_D;                // [128][8] scores
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, lane = tid & 31, warp = tid >> 5;
    const int nt = *a.ntiles;
    if ((int)blockIdx.x >= nt) return;
    for (int i = tid; i < SCANN_D * SCANN_D / 4; i += LA_THREADS) {
        st4(W2s + i * 4, ldg4(a.W2 + i * 4));
        st4(Wks + i * 4, ldg4(a.Wk + i * 4));
    }
    float acc[8][8];
    for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const size_t rowbase = (size_t)t * SCANN_TILE;
        __syncthreads();
        load_tile(S0, a.g_in, rowbase, tid);
        int pc[8], pj[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            pc[i] = a.pair_c[rowbase + ty + 16 * i];
            pj[i] = pc[i] >= 0 ? a.pair_j[rowbase + ty + 16 * i] : 0;
        }
        __syncthreads();
        zero_acc(acc);
        gemm_nn<false>(S0, W2s, acc, ty, tx);
        {   // epilogue 1: geometry update, a = x[j] * g'
            float gg[8], bg[8];
            load8(a.gamma_g, tx, gg);
            load8(a.beta_g, tx, bg);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = ty + 16 * i, c = pc[i] >= 0 ? pc[i] : 0, j = pj[i];
                float p1[8], p3[8], xj[8], z[8];
                load8(a.proj + (size_t)c * 3 * SCANN_D, tx, p1);
                load8(a.proj + (size_t)j * 3 * SCANN_D + SCANN_D, tx, p3);
                load8(a.x + (size_t)j * SCANN_D, tx, xj);
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float pre = acc[i][q] + p1[q] + p3[q];
                    z[q] = swish_f(pre) + S0[row * LA_LDS + col_of(tx, q)];
                    s += z[q];
                }
                float mean = half_warp_sum(s) * (1.0f / SCANN_D);
                float v = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) { z[q] -= mean; v += z[q] * z[q]; }
                float inv = rsqrtf(half_warp_sum(v) * (1.0f / SCANN_D) + SCANN_LN_EPS);
                float gp[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    gp[q] = pc[i] >= 0 ? z[q] * inv * gg[q] + bg[q] : 0.f;
                    acc[i][q] = xj[q] * gp[q];
                }
                store8(a.g_out + (rowbase + row) * SCANN_D, tx, gp);
            }
        }
        __syncthreads();                 // every thread is done reading g from S0
#pragma unroll
        for (int i = 0; i < 8; ++i) store8(S0 + (ty + 16 * i) * LA_LDS, tx, acc[i]);
        __syncthreads();
        zero_acc(acc);
        gemm_nn<false>(S0, Wks, acc, ty, tx);
        {   // keys and scores
            float bk[8];
            load8(a.bk, tx, bk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = ty + 16 * i, c = pc[i] >= 0 ? pc[i] : 0;
                float qv[8];
                load8(a.proj + (size_t)c * 3 * SCANN_D + 2 * SCANN_D, tx, qv);
                float e0 = 0.f, e1 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[i][q] += bk[q];
                    acc[i][q + 4] += bk[q + 4];
                    e0 = fmaf(acc[i][q], qv[q], e0);
                    e1 = fmaf(acc[i][q + 4], qv[q + 4], e1);
                }
                e0 = quad_sum(e0) * 0.25f;      // hd^-0.5 = 16^-0.5 (attention.py:180-181)
                e1 = quad_sum(e1) * 0.25f;
                if ((tx & 3) == 0) {
                    Es[row * 8 + (tx >> 2)] = e0;
                    Es[row * 8 + 4 + (tx >> 2)] = e1;
                }
            }
        }
        __syncthreads();                 // every thread is done reading a from S0
#pragma unroll
        for (int i = 0; i < 8; ++i) store8(S0 + (ty + 16 * i) * LA_LDS, tx, acc[i]);
        __syncthreads();
        // per-atom softmax over the atom's rows, context, residual, LayerNorm: one warp per atom
        const int a0 = a.tile_a0[t], a1 = a.tile_a1[t];
        const float4 gam = ldg4(a.gamma + lane * 4), bet = ldg4(a.beta + lane * 4);
        for (int atom = a0 + warp; atom < a1; atom += LA_THREADS / 32) {
            const int n = a.cnt[atom];
            if (n == 0) continue;
            const int r0 = a.rowptr[atom] - (int)rowbase;
            const int h = lane >> 2;
            float m = -INFINITY;
            for (int r = 0; r < n; ++r) m = fmaxf(m, Es[(r0 + r) * 8 + h]);
            float s = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            for (int r = 0; r < n; ++r) {
                float p = expf(Es[(r0 + r) * 8 + h] - m);
                float4 kv = ld4(S0 + (r0 + r) * LA_LDS + lane * 4);
                s += p;
                c0 = fmaf(p, kv.x, c0); c1 = fmaf(p, kv.y, c1); c2 = fmaf(p, kv.z, c2); c3 = fmaf(p, kv.w, c3);
            }
            const float is = 1.0f / s;
            if (a.attn && (lane & 3) == 0)
                for (int r = 0; r < n; ++r) a.attn[(rowbase + r0 + r) * 8 + h] = expf(Es[(r0 + r) * 8 + h] - m) * is;
            float4 q = ld4(a.proj + (size_t)atom * 3 * SCANN_D + 2 * SCANN_D + lane * 4);
            c0 = c0 * is + q.x; c1 = c1 * is + q.y; c2 = c2 * is + q.z; c3 = c3 * is + q.w;
            if (a.ctx_pre) st4(a.ctx_pre + (size_t)atom * SCANN_D + lane * 4, make_float4(c0, c1, c2, c3));
            float mean = warp_sum(c0 + c1 + c2 + c3) * (1.0f / SCANN_D);
            c0 -= mean; c1 -= mean; c2 -= mean; c3 -= mean;
            float inv = rsqrtf(warp_sum(c0 * c0 + c1 * c1 + c2 * c2 + c3 * c3) * (1.0f / SCANN_D) + SCANN_LN_EPS);
            st4(a.out + (size_t)atom * SCANN_D + lane * 4,
                make_float4(c0 * inv * gam.x + bet.x, c1 * inv * gam.y + bet.y, c2 * inv * gam.z + bet.z,
                            c3 * inv * gam.w + bet.w));
        }
    }
}

// =============================================================================================
// Backward (forward is recomputed per tile, only g_l is read back)
// =============================================================================================
struct LaBwdArgs {
    const int32_t* ntiles; const int32_t* tile_a0; const int32_t* tile_a1;
    const int32_t* cnt; const int32_t* rowptr; const int32_t* pair_c; const int32_t* pair_j;
    const float* x; const float* proj; const float* g_in;
    const float* W2; const float* Wk; const float* W2T; const float* WkT; const float* bk;
    const float* gamma_g; const float* beta_g;
    const float* d_ctx;    // [R,128] gradient w.r.t. the pre-LayerNorm context
    const float* dg_up;    // [rows,128] gradient w.r.t. g' from the next layer (nullable)
    float* dg_out;         // [rows,128] gradient w.r.t. g
    float* dq;             // [R,128]  <- d_ctx + 0.25 sum_n de k   (atoms with pairs)
    float* s_pre;          // [R,128]  <- sum_n d_pre               (atoms with pairs)
    float* t_scatter;      // [R,128]  += d_pre scattered to neighbour rows (pre-zeroed)
    float* dx_scatter;     // [R,128]  += (d_a * g') scattered to neighbour rows (pre-zeroed)
    float* wpart;          // [grid][2][128][128] per-CTA partial dWk, dW2
    float* dgamma_g; float* dbeta_g; float* dbk;
};

__global__ void __launch_bounds__(LA_THREADS, 1) la_bwd_simt_kernel(const LaBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* S0 = smem;
    float* S1 = S0 + SCANN_TILE * LA_LDS;
    float* S2 = S1 + SCANN_TILE * LA_LDS;
    float* Es = S2 + SCANN_TILE * LA_LDS;               // [128][8] e -> p
    float* Ds = Es + SCANN_TILE * 8;                    // [128][8] dp -> de
    float* Gacc = Ds + SCANN_TILE * 8;                  // [3][128] dgamma_g, dbeta_g, dbk
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, lane = tid & 31, warp = tid >> 5;
    const int nt = *a.ntiles;
    if ((int)blockIdx.x >= nt) return;
    for (int i = tid; i < 3 * SCANN_D; i += LA_THREADS) Gacc[i] = 0.f;
    float* wpk = a.wpart + (size_t)blockIdx.x * 2 * SCANN_D * SCANN_D;
    float* wp2 = wpk + SCANN_D * SCANN_D;
    bool first = true;
    float acc[8][8];
    float gg[8], bg[8];
    load8(a.gamma_g, tx, gg);
    load8(a.beta_g, tx, bg);
    for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const size_t rowbase = (size_t)t * SCANN_TILE;
        __syncthreads();
        load_tile(S0, a.g_in, rowbase, tid);
        int pc[8], pj[8];
        float mean_[8], inv_[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            pc[i] = a.pair_c[rowbase + ty + 16 * i];
            pj[i] = pc[i] >= 0 ? a.pair_j[rowbase + ty + 16 * i] : 0;
        }
        __syncthreads();
        // ---- recompute forward: pre -> S2, a -> S1
        zero_acc(acc);
        gemm_nn<true>(S0, a.W2, acc, ty, tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = ty + 16 * i, c = pc[i] >= 0 ? pc[i] : 0, j = pj[i];
            float p1[8], p3[8], xj[8], z[8], pre[8];
            load8(a.proj + (size_t)c * 3 * SCANN_D, tx, p1);
            load8(a.proj + (size_t)j * 3 * SCANN_D + SCANN_D, tx, p3);
            load8(a.x + (size_t)j * SCANN_D, tx, xj);
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                pre[q] = acc[i][q] + p1[q] + p3[q];
                z[q] = swish_f(pre[q]) + S0[row * LA_LDS + col_of(tx, q)];
                s += z[q];
            }
            float mean = half_warp_sum(s) * (1.0f / SCANN_D);
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) { float d = z[q] - mean; v += d * d; }
            float inv = rsqrtf(half_warp_sum(v) * (1.0f / SCANN_D) + SCANN_LN_EPS);
            mean_[i] = mean;
            inv_[i] = inv;
            float av[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float gp = (z[q] - mean) * inv * gg[q] + bg[q];
                av[q] = pc[i] >= 0 ? xj[q] * gp : 0.f;
                if (pc[i] < 0) pre[q] = 0.f;
            }
            store8(S2 + row * LA_LDS, tx, pre);
            store8(S1 + row * LA_LDS, tx, av);
        }
        __syncthreads();
        // ---- k = a @ Wk + bk ; e and dp per (row, head)
        zero_acc(acc);
        gemm_nn<true>(S1, a.Wk, acc, ty, tx);
        {
            float bk[8];
            load8(a.bk, tx, bk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = ty + 16 * i, c = pc[i] >= 0 ? pc[i] : 0;
                float qv[8], dc[8];
                load8(a.proj + (size_t)c * 3 * SCANN_D + 2 * SCANN_D, tx, qv);
                load8(a.d_ctx + (size_t)c * SCANN_D, tx, dc);
                float e0 = 0.f, e1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[i][q] += bk[q];
                    acc[i][q + 4] += bk[q + 4];
                    e0 = fmaf(acc[i][q], qv[q], e0);
                    e1 = fmaf(acc[i][q + 4], qv[q + 4], e1);
                    d0 = fmaf(acc[i][q], dc[q], d0);
                    d1 = fmaf(acc[i][q + 4], dc[q + 4], d1);
                }
                e0 = quad_sum(e0) * 0.25f; e1 = quad_sum(e1) * 0.25f;
                d0 = quad_sum(d0); d1 = quad_sum(d1);
                if ((tx & 3) == 0) {
                    Es[row * 8 + (tx >> 2)] = e0; Es[row * 8 + 4 + (tx >> 2)] = e1;
                    Ds[row * 8 + (tx >> 2)] = d0; Ds[row * 8 + 4 + (tx >> 2)] = d1;
                }
                store8(S0 + row * LA_LDS, tx, acc[i]);       // k tile (g is reloaded later)
            }
        }
        __syncthreads();
        // ---- per atom: p, de ; dq = d_ctx + 0.25 sum_n de k
        const int a0 = a.tile_a0[t], a1 = a.tile_a1[t];
        for (int atom = a0 + warp; atom < a1; atom += LA_THREADS / 32) {
            const int n = a.cnt[atom];
            if (n == 0) continue;
            const int r0 = a.rowptr[atom] - (int)rowbase;
            {
                const int h = lane & 7, rs = lane >> 3;
                float m = -INFINITY;
                for (int r = rs; r < n; r += 4) m = fmaxf(m, Es[(r0 + r) * 8 + h]);
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
                float s = 0.f, dot = 0.f;
                for (int r = rs; r < n; r += 4) {
                    float p = expf(Es[(r0 + r) * 8 + h] - m);
                    s += p;
                    dot = fmaf(p, Ds[(r0 + r) * 8 + h], dot);
                }
                s += __shfl_xor_sync(0xffffffffu, s, 8);   s += __shfl_xor_sync(0xffffffffu, s, 16);
                dot += __shfl_xor_sync(0xffffffffu, dot, 8); dot += __shfl_xor_sync(0xffffffffu, dot, 16);
                const float is = 1.0f / s;
                dot *= is;
                for (int r = rs; r < n; r += 4) {
                    float p = expf(Es[(r0 + r) * 8 + h] - m) * is;
                    float dp = Ds[(r0 + r) * 8 + h];
                    Es[(r0 + r) * 8 + h] = p;
                    Ds[(r0 + r) * 8 + h] = p * (dp - dot);
                }
            }
            __syncwarp();
            {
                const int h = lane >> 2;
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
                for (int r = 0; r < n; ++r) {
                    float de = Ds[(r0 + r) * 8 + h];
                    float4 kv = ld4(S0 + (r0 + r) * LA_LDS + lane * 4);
                    c0 = fmaf(de, kv.x, c0); c1 = fmaf(de, kv.y, c1); c2 = fmaf(de, kv.z, c2); c3 = fmaf(de, kv.w, c3);
                }
                float4 dc = ld4(a.d_ctx + (size_t)atom * SCANN_D + lane * 4);
                st4(a.dq + (size_t)atom * SCANN_D + lane * 4,
                    make_float4(dc.x + 0.25f * c0, dc.y + 0.25f * c1, dc.z + 0.25f * c2, dc.w + 0.25f * c3));
            }
        }
        __syncthreads();
        // ---- dk = p d_ctx[c] + 0.25 de q[c]  -> S0 ; dbk partial
        {
            float dbk[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = ty + 16 * i, c = pc[i] >= 0 ? pc[i] : 0;
                float qv[8], dc[8], dk[8];
                load8(a.proj + (size_t)c * 3 * SCANN_D + 2 * SCANN_D, tx, qv);
                load8(a.d_ctx + (size_t)c * SCANN_D, tx, dc);
                const float p0 = Es[row * 8 + (tx >> 2)], p1 = Es[row * 8 + 4 + (tx >> 2)];
                const float e0 = 0.25f * Ds[row * 8 + (tx >> 2)], e1 = 0.25f * Ds[row * 8 + 4 + (tx >> 2)];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    dk[q] = pc[i] >= 0 ? fmaf(p0, dc[q], e0 * qv[q]) : 0.f;
                    dk[q + 4] = pc[i] >= 0 ? fmaf(p1, dc[q + 4], e1 * qv[q + 4]) : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) dbk[q] += dk[q];
                store8(S0 + row * LA_LDS, tx, dk);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) atomicAdd(&Gacc[2 * SCANN_D + col_of(tx, q)], dbk[q]);
        }
        __syncthreads();
        // ---- dWk partial = a^T dk
        zero_acc(acc);
        gemm_tn(S1, S0, acc, ty, tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float* dst = wpk + (ty * 8 + i) * SCANN_D;
            if (!first) {
                float old[8];
                load8(dst, tx, old);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[i][q] += old[q];
            }
            store8(dst, tx, acc[i]);
        }
        // ---- d_a = dk @ Wk^T
        zero_acc(acc);
        gemm_nn<true>(S0, a.WkT, acc, ty, tx);
        __syncthreads();                 // S0 (dk) and S1 (a) are dead from here
        load_tile(S0, a.g_in, rowbase, tid);
        __syncthreads();
        // ---- geometry side: d_nbr, LN_g backward, d_pre
        {
            float dgam[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dbet[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = ty + 16 * i, j = pj[i];
                const bool ok = pc[i] >= 0;
                float pre[8], gv[8], xj[8], up[8], xh[8], dxh[8], dpre[8], dn[8];
                load8(S2 + row * LA_LDS, tx, pre);
                load8(S0 + row * LA_LDS, tx, gv);
                load8(a.x + (size_t)j * SCANN_D, tx, xj);
                if (a.dg_up) load8(a.dg_up + (rowbase + row) * SCANN_D, tx, up);
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float z = swish_f(pre[q]) + gv[q];
                    xh[q] = (z - mean_[i]) * inv_[i];
                    float gp = xh[q] * gg[q] + bg[q];
                    float da = acc[i][q];
                    dn[q] = da * gp;
                    float dgt = (a.dg_up ? up[q] : 0.f) + da * xj[q];
                    if (!ok) dgt = 0.f;
                    dgam[q] = fmaf(dgt, xh[q], dgam[q]);
                    dbet[q] += dgt;
                    dxh[q] = dgt * gg[q];
                    s1 += dxh[q];
                    s2 = fmaf(dxh[q], xh[q], s2);
                }
                s1 = half_warp_sum(s1) * (1.0f / SCANN_D);
                s2 = half_warp_sum(s2) * (1.0f / SCANN_D);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float dz = ok ? inv_[i] * (dxh[q] - s1 - xh[q] * s2) : 0.f;
                    dpre[q] = dz * swish_grad_f(pre[q]);
                    acc[i][q] = dz;
                }
                if (ok) {
                    red_add4(a.dx_scatter + (size_t)j * SCANN_D + tx * 4, dn[0], dn[1], dn[2], dn[3]);
                    red_add4(a.dx_scatter + (size_t)j * SCANN_D + 64 + tx * 4, dn[4], dn[5], dn[6], dn[7]);
                    red_add4(a.t_scatter + (size_t)j * SCANN_D + tx * 4, dpre[0], dpre[1], dpre[2], dpre[3]);
                    red_add4(a.t_scatter + (size_t)j * SCANN_D + 64 + tx * 4, dpre[4], dpre[5], dpre[6], dpre[7]);
                }
                store8(S1 + row * LA_LDS, tx, dpre);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                atomicAdd(&Gacc[col_of(tx, q)], dgam[q]);
                atomicAdd(&Gacc[SCANN_D + col_of(tx, q)], dbet[q]);
            }
        }
        __syncthreads();
        // ---- dg = dz + d_pre @ W2^T
        gemm_nn<true>(S1, a.W2T, acc, ty, tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) store8(a.dg_out + (rowbase + ty + 16 * i) * SCANN_D, tx, acc[i]);
        // ---- s_pre[c] = sum_n d_pre
        for (int atom = a0 + warp; atom < a1; atom += LA_THREADS / 32) {
            const int n = a.cnt[atom];
            if (n == 0) continue;
            const int r0 = a.rowptr[atom] - (int)rowbase;
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            for (int r = 0; r < n; ++r) {
                float4 v = ld4(S1 + (r0 + r) * LA_LDS + lane * 4);
                c0 += v.x; c1 += v.y; c2 += v.z; c3 += v.w;
            }
            st4(a.s_pre + (size_t)atom * SCANN_D + lane * 4, make_float4(c0, c1, c2, c3));
        }
        // ---- dW2 partial = g^T d_pre
        zero_acc(acc);
        gemm_tn(S0, S1, acc, ty, tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float* dst = wp2 + (ty * 8 + i) * SCANN_D;
            if (!first) {
                float old[8];
                load8(dst, tx, old);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[i][q] += old[q];
            }
            store8(dst, tx, acc[i]);
        }
        first = false;
    }
    __syncthreads();
    if (tid < SCANN_D) {
        atomicAdd(a.dgamma_g + tid, Gacc[tid]);
        atomicAdd(a.dbeta_g + tid, Gacc[SCANN_D + tid]);
        atomicAdd(a.dbk + tid, Gacc[2 * SCANN_D + tid]);
    }
}

// dWk += sum_cta wpart[cta][0], dW2 += sum_cta wpart[cta][1]
__global__ void __launch_bounds__(256) la_wpart_reduce_kernel(const float* __restrict__ wpart,
                                                              const int32_t* __restrict__ ntiles, int grid, int groups,
                                                              float* __restrict__ dWk, float* __restrict__ dW2) {
    const int nact = min(grid, (*ntiles + groups - 1) / groups);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. 2*128*128
    if (i >= 2 * SCANN_D * SCANN_D) return;
    float s = 0.f;
    for (int c = 0; c < nact; ++c) s += wpart[(size_t)c * 2 * SCANN_D * SCANN_D + i];
    if (i < SCANN_D * SCANN_D) dWk[i] += s;
    else dW2[i - SCANN_D * SCANN_D] += s;
}

#define LA_FWD_SMEM ((SCANN_TILE * LA_LDS + 2 * SCANN_D * SCANN_D + SCANN_TILE * 8) * sizeof(float))
#define LA_BWD_SMEM ((3 * SCANN_TILE * LA_LDS + 2 * SCANN_TILE * 8 + 3 * SCANN_D) * sizeof(float))

extern "C" int scann_la_forward(int grid, const int32_t* ntiles, const int32_t* tile_a0, const int32_t* tile_a1,
                                const int32_t* cnt, const int32_t* rowptr, const int32_t* pair_c,
                                const int32_t* pair_j, const float* x, const float* proj, const float* g_in,
                                const float* W2, const float* Wk, const float* bk, const float* gamma_g,
                                const float* beta_g, const float* gamma, const float* beta, float* g_out,
                                float* ctx_pre, float* out, float* attn, void* stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(la_fwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)LA_FWD_SMEM);
        if (e != cudaSuccess) { scann_set_error("la_forward: smem opt-in failed: %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    if (grid <= 0) return 0;
    LaFwdArgs a{ntiles, tile_a0, tile_a1, cnt, rowptr, pair_c, pair_j, x, proj, g_in, W2, Wk, bk,
                gamma_g, beta_g, gamma, beta, g_out, ctx_pre, out, attn};
    la_fwd_simt_kernel<<<grid, LA_THREADS, LA_FWD_SMEM, (cudaStream_t)stream>>>(a);
    return scann_check_launch("scann_la_forward");
}

extern "C" int scann_la_backward(int grid, const int32_t* ntiles, const int32_t* tile_a0, const int32_t* tile_a1,
                                 const int32_t* cnt, const int32_t* rowptr, const int32_t* pair_c,
                                 const int32_t* pair_j, const float* x, const float* proj, const float* g_in,
                                 const float* W2, const float* Wk, const float* W2T, const float* WkT, const float* bk,
                                 const float* gamma_g, const float* beta_g, const float* d_ctx, const float* dg_up,
                                 float* dg_out, float* dq, float* s_pre, float* t_scatter, float* dx_scatter,
                                 float* wpart, float* dgamma_g, float* dbeta_g, float* dbk, void* stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(la_bwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)LA_BWD_SMEM);
        if (e != cudaSuccess) { scann_set_error("la_backward: smem opt-in failed: %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    if (grid <= 0) return 0;
    LaBwdArgs a{ntiles, tile_a0, tile_a1, cnt, rowptr, pair_c, pair_j, x, proj, g_in, W2, Wk, W2T, WkT, bk,
                gamma_g, beta_g, d_ctx, dg_up, dg_out, dq, s_pre, t_scatter, dx_scatter, wpart,
                dgamma_g, dbeta_g, dbk};
    la_bwd_simt_kernel<<<grid, LA_THREADS, LA_BWD_SMEM, (cudaStream_t)stream>>>(a);
    return scann_check_launch("scann_la_backward");
}

// dWk += sum over CTAs of wpart[cta][0], dW2 += sum of wpart[cta][1]; `grid` as in scann_la_backward.
// tile_stride as in scann_la_wgrad_tc (CTA c produced a partial iff c * (128 / tile_stride) < ntiles).
extern "C" int scann_la_wpart_reduce(const float* wpart, const int32_t* ntiles, int grid, int tile_stride, float* dWk,
                                     float* dW2, void* stream) {
    if (tile_stride != 64 && tile_stride != 128) { scann_set_error("la_wpart_reduce: tile_stride must be 64 or 128"); return 1; }
    la_wpart_reduce_kernel<<<(2 * SCANN_D * SCANN_D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        wpart, ntiles, grid, SCANN_TILE / tile_stride, dWk, dW2);
    return scann_check_launch("scann_la_wpart_reduce");
}
