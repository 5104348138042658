// Local attention backward on the tcgen05 tensor cores, three kernels per layer:
//
//   la_attn_bwd_tc : softmax / context backward, d_k, d_q ; d_a = d_k @ Wk^T (3xTF32) ;
//                    d_nbr = d_a * g' scattered to dx[j] ; dg' += d_a * x[j]
//   la_geom_bwd_tc : LN_g backward, d_pre ; dg = d_z + d_pre @ W2^T (3xTF32) ; s_pre, t scatter
//   la_wgrad_tc    : dWk += (x[j]*g')^T d_k  and  dW2 += g^T d_pre  (MN-major operands, 3xTF32; a single
//                    TF32 product was measured 5e-4 off on the fullerene case: rows of one species are
//                    too correlated for the rounding errors to average out over the pair sum)
//
// The forward (la_tc.cu, training mode) saved the filter_geo pre-activation and the keys; the backward
// overwrites them in place with d_pre and d_k, which la_wgrad_tc consumes.  As in the forward, the
// weight block is the stationary M x K operand in tensor memory and 128-pair tiles stream through
// shared memory as the N x K operand.
#include "common.cuh"
#include "tc_common.cuh"

#define LTC_THREADS 512
#define LTC_WARPS 16
#define LTC_RPW 8

// ---- helpers shared with la_tc.cu (kept static to this translation unit) ----
__device__ __forceinline__ void b_weightT_to_tmem(const float* __restrict__ W, uint32_t t_hi, uint32_t t_lo, int warp,
                                                  int lane) {
    const int n = (warp & 3) * 32 + lane, kbase = (warp >> 2) * 32;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float w[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) w[q] = __ldg(W + (size_t)(kbase + q) * SCANN_D + n);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        float hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) tf32_split(w[g * 16 + q], hi[q], lo[q]);
        tmem_st16(t_hi + lane_base + kbase + g * 16, hi);
        tmem_st16(t_lo + lane_base + kbase + g * 16, lo);
    }
    tmem_st_wait();
}

__device__ __forceinline__ void b_issue_3xtf32(uint32_t t_whi, uint32_t t_wlo, uint32_t xh, uint32_t xl, uint32_t t_dm,
                                               uint32_t t_dc, uint64_t* bar, int nrows) {
    const uint32_t idesc = tc_idesc_tf32(128, nrows, false, false);
    const uint64_t dh = tc_desc_kmajor(xh, 0), dl = tc_desc_kmajor(xl, 0);
    // one accumulator, correction products first (see la_tc.cu)
    (void)t_dc;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_dm, t_wlo + ks * 8, dh + ks * TC_KSTEP_DESC, idesc, ks != 0);
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_dm, t_whi + ks * 8, dl + ks * TC_KSTEP_DESC, idesc, true);
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_dm, t_whi + ks * 8, dh + ks * TC_KSTEP_DESC, idesc, true);
    tc_commit(bar);
}

__device__ __forceinline__ void b_tmem_to_rows(uint32_t t_dm, uint32_t t_dc, uint8_t* S, int warp, int lane, int nrows) {
    const int n = (warp & 3) * 32 + lane, rbase = (warp >> 2) * 32;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (rbase + h * 16 >= nrows) break;              // warp-uniform: rows beyond nrows are never valid
        float m[16];
        (void)t_dc;
        tmem_ld16(t_dm + lane_base + rbase + h * 16, m);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) *reinterpret_cast<float*>(S + tc_off(rbase + h * 16 + q, n)) = m[q];
    }
}

__device__ __forceinline__ void split_store(uint8_t* sHi, uint8_t* sLo, uint32_t off, float4 v) {
    float4 h, l;
    tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(sHi + off) = h;
    *reinterpret_cast<float4*>(sLo + off) = l;
}

// =============================================================================================
// Attention backward
// =============================================================================================
struct LaAttnBwdArgs {
    const int32_t* ntiles; const int32_t* tile_a0; const int32_t* tile_a1;
    const int32_t* cnt; const int32_t* rowptr; const int32_t* pair_c; const int32_t* pair_j;
    const float* x;          // [R,128]
    const float* proj;       // [R,384]; q = columns 256..383
    const float* g_new;      // [rows,128] g'
    float* kbuf;             // [rows,128] in: keys k (saved by the forward) ; out: d_k
    const float* WkT;        // transposed key kernel: WkT[n][ka] = Wk[ka][n]
    const float* d_ctx;      // [R,128]
    float* dg;               // [rows,128] gradient w.r.t. g': accumulate (dg_accum) or overwrite
    int dg_accum;
    float* dq;               // [R,128] <- d_ctx + 0.25 sum_n de k   (atoms with pairs)
    float* dx_scatter;       // [R,128] += d_a * g' at the neighbour rows
    float* dbk;              // [128]   += column sums of d_k
    int mma_rows;            // rows of a tile slot that can hold pairs (multiple of 16)
    const ScannDropCtl* drop;    // attention-probability Dropout of the forward pass (see LaAttnArgs), NULL = off
    int drop_site;
};

template <int NG>
__global__ void __launch_bounds__(LTC_THREADS, 1) la_attn_bwd_tc_kernel(const LaAttnBwdArgs a) {
    using G = LaGroups<NG>;
    constexpr int TR = G::TR, WG = G::WG, GT = G::GT;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[NG];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_dbk[SCANN_D];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp / WG, wg = warp % WG, gtid = tid - grp * GT;
    uint8_t* sHi = smem + (size_t)grp * 3 * G::IMG;
    uint8_t* sLo = sHi + G::IMG;
    uint8_t* sS = sLo + G::IMG;                                              // keys, later d_a
    float* Es = reinterpret_cast<float*>(smem + 3 * TC_TILE_BYTES) + grp * 2 * TR * 8;   // [TR][8] e -> p
    float* Ds = Es + TR * 8;                                                 // [TR][8] dp -> de
    uint64_t* bar = &bars[grp];
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid < NG) mbar_init(&bars[tid], 1);
    if (tid == 0) mbar_fence_init();
    if (tid < SCANN_D) s_dbk[tid] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t t_whi = tmem, t_wlo = tmem + 128, t_dm = tmem + 256 + grp * 2 * TR, t_dc = t_dm + TR;
    // stationary operand A[M = ka][K = n] = Wk[ka][n]  (= transpose of WkT, loaded coalesced)
    b_weightT_to_tmem(a.WkT, t_whi, t_wlo, warp, lane);
    pdl_wait();
    const int nt = *a.ntiles;
    tc_fence_before();
    __syncthreads();                                     // the whole weight is in tensor memory for every group
    tc_fence_after();
    float4 dbk = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t phase = 0;
    const int t_step = gridDim.x * NG;
    for (int t = blockIdx.x * NG + grp; t < nt; t += t_step) {
        const size_t rowbase = (size_t)t * TR;
        int pc[LTC_RPW];
#pragma unroll
        for (int i = 0; i < LTC_RPW; ++i) pc[i] = a.pair_c[rowbase + wg + WG * i];
        // ---- phase A: keys -> S ; e = 0.25 <q_h,k_h>, dp = <dctx_h,k_h> per (row, head)
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            float4 kv[4], qv[4], dc[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = hb * 4 + ii, r = wg + WG * i;
                kv[ii] = make_float4(0.f, 0.f, 0.f, 0.f); qv[ii] = kv[ii]; dc[ii] = kv[ii];
                if (pc[i] >= 0) {
                    kv[ii] = ld4(a.kbuf + (rowbase + r) * SCANN_D + lane * 4);
                    qv[ii] = ld4(a.proj + (size_t)pc[i] * 3 * SCANN_D + 2 * SCANN_D + lane * 4);
                    dc[ii] = ld4(a.d_ctx + (size_t)pc[i] * SCANN_D + lane * 4);
                }
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int r = wg + WG * (hb * 4 + ii);
                if (pc[hb * 4 + ii] < 0) continue;                       // padding row (warp-uniform)
                *reinterpret_cast<float4*>(sS + tc_off4(r, lane)) = kv[ii];
                float e = kv[ii].x * qv[ii].x + kv[ii].y * qv[ii].y + kv[ii].z * qv[ii].z + kv[ii].w * qv[ii].w;
                float d = kv[ii].x * dc[ii].x + kv[ii].y * dc[ii].y + kv[ii].z * dc[ii].z + kv[ii].w * dc[ii].w;
                e = quad_sum(e) * 0.25f;
                d = quad_sum(d);
                if ((lane & 3) == 0) { Es[r * 8 + (lane >> 2)] = e; Ds[r * 8 + (lane >> 2)] = d; }
            }
        }
        group_sync(grp, GT);
        // ---- phase B (warp per atom): p, de ; dq = d_ctx + 0.25 sum_n de k
        const int a0 = a.tile_a0[t], a1 = a.tile_a1[t];
        for (int atom = a0 + wg; atom < a1; atom += WG) {
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
                    float p = __expf(Es[(r0 + r) * 8 + h] - m);
                    s += p;
                    // gradient w.r.t. the softmax output = (gradient w.r.t. the dropped probabilities) * mask
                    const float dm = drop_mult(a.drop, a.drop_site, (uint32_t)(rowbase + r0 + r) * 8u + h);
                    Ds[(r0 + r) * 8 + h] *= dm;
                    dot = fmaf(p, Ds[(r0 + r) * 8 + h], dot);
                }
                s += __shfl_xor_sync(0xffffffffu, s, 8);     s += __shfl_xor_sync(0xffffffffu, s, 16);
                dot += __shfl_xor_sync(0xffffffffu, dot, 8); dot += __shfl_xor_sync(0xffffffffu, dot, 16);
                const float is = 1.0f / s;
                dot *= is;
                for (int r = rs; r < n; r += 4) {
                    float p = __expf(Es[(r0 + r) * 8 + h] - m) * is;
                    float dp = Ds[(r0 + r) * 8 + h];
                    // d_k uses the dropped probabilities, the softmax backward the undropped ones
                    Es[(r0 + r) * 8 + h] = p * drop_mult(a.drop, a.drop_site, (uint32_t)(rowbase + r0 + r) * 8u + h);
                    Ds[(r0 + r) * 8 + h] = p * (dp - dot);
                }
            }
            __syncwarp();
            {
                const int h = lane >> 2;
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
                for (int r = 0; r < n; ++r) {
                    float de = Ds[(r0 + r) * 8 + h];
                    float4 kv = *reinterpret_cast<const float4*>(sS + tc_off4(r0 + r, lane));
                    c0 = fmaf(de, kv.x, c0); c1 = fmaf(de, kv.y, c1); c2 = fmaf(de, kv.z, c2); c3 = fmaf(de, kv.w, c3);
                }
                float4 dc = ld4(a.d_ctx + (size_t)atom * SCANN_D + lane * 4);
                st4(a.dq + (size_t)atom * SCANN_D + lane * 4,
                    make_float4(dc.x + 0.25f * c0, dc.y + 0.25f * c1, dc.z + 0.25f * c2, dc.w + 0.25f * c3));
            }
        }
        group_sync(grp, GT);
        // ---- phase C: dk = p d_ctx[c] + 0.25 de q[c] -> hi/lo images, global (in place over k), dbk
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            float4 qv[4], dc[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = hb * 4 + ii;
                qv[ii] = make_float4(0.f, 0.f, 0.f, 0.f); dc[ii] = qv[ii];
                if (pc[i] >= 0) {
                    qv[ii] = ld4(a.proj + (size_t)pc[i] * 3 * SCANN_D + 2 * SCANN_D + lane * 4);
                    dc[ii] = ld4(a.d_ctx + (size_t)pc[i] * SCANN_D + lane * 4);
                }
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = hb * 4 + ii, r = wg + WG * i;
                if (pc[i] < 0) continue;
                float4 dk = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pc[i] >= 0) {
                    const float p = Es[r * 8 + (lane >> 2)], de = 0.25f * Ds[r * 8 + (lane >> 2)];
                    dk = make_float4(fmaf(p, dc[ii].x, de * qv[ii].x), fmaf(p, dc[ii].y, de * qv[ii].y),
                                     fmaf(p, dc[ii].z, de * qv[ii].z), fmaf(p, dc[ii].w, de * qv[ii].w));
                }
                dbk = make_float4(dbk.x + dk.x, dbk.y + dk.y, dbk.z + dk.z, dbk.w + dk.w);
                split_store(sHi, sLo, tc_off4(r, lane), dk);
                st4(a.kbuf + (rowbase + r) * SCANN_D + lane * 4, dk);
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(grp, GT);
        if (wg == 0 && tc_elect_one()) {
            tc_fence_after();
            b_issue_3xtf32(t_whi, t_wlo, smem_u32(sHi), smem_u32(sLo), t_dm, t_dc, bar, a.mma_rows);     // d_a^T = Wk dk^T
        }
        // ---- phase D: d_nbr = d_a * g' -> dx[j] ; dg' (+)= d_a * x[j].  The loads of its first four rows are issued
        // before the wait for the MMAs, the second four behind the first four's arithmetic
        float4 gp[4], xj[4], dg[4];
        int jj[4];
        auto load_d = [&](int hb) {
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = hb * 4 + ii, r = wg + WG * i;
                gp[ii] = make_float4(0.f, 0.f, 0.f, 0.f); xj[ii] = gp[ii]; dg[ii] = gp[ii];
                jj[ii] = 0;
                if (pc[i] >= 0) {
                    jj[ii] = a.pair_j[rowbase + r];
                    gp[ii] = ld4(a.g_new + (rowbase + r) * SCANN_D + lane * 4);
                    xj[ii] = ld4(a.x + (size_t)jj[ii] * SCANN_D + lane * 4);
                    if (a.dg_accum) dg[ii] = ld4(a.dg + (rowbase + r) * SCANN_D + lane * 4);
                }
            }
        };
        auto apply_d = [&](int hb) {
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = hb * 4 + ii, r = wg + WG * i;
                if (pc[i] < 0) continue;
                const float4 da = *reinterpret_cast<const float4*>(sS + tc_off4(r, lane));
                red_add4(a.dx_scatter + (size_t)jj[ii] * SCANN_D + lane * 4, da.x * gp[ii].x, da.y * gp[ii].y,
                         da.z * gp[ii].z, da.w * gp[ii].w);
                st4(a.dg + (rowbase + r) * SCANN_D + lane * 4,
                    make_float4(fmaf(da.x, xj[ii].x, dg[ii].x), fmaf(da.y, xj[ii].y, dg[ii].y),
                                fmaf(da.z, xj[ii].z, dg[ii].z), fmaf(da.w, xj[ii].w, dg[ii].w)));
            }
        };
        load_d(0);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        if (t + t_step >= nt) pdl_trigger();             // last tile of this group, only its epilogue is left
        b_tmem_to_rows(t_dm, t_dc, sS, wg, lane, a.mma_rows);
        tc_fence_before();
        group_sync(grp, GT);
        apply_d(0);
        load_d(1);
        apply_d(1);
        group_sync(grp, GT);
    }
    pdl_trigger();
    atomicAdd(&s_dbk[lane * 4 + 0], dbk.x); atomicAdd(&s_dbk[lane * 4 + 1], dbk.y);
    atomicAdd(&s_dbk[lane * 4 + 2], dbk.z); atomicAdd(&s_dbk[lane * 4 + 3], dbk.w);
    tc_fence_before();
    __syncthreads();
    if (tid < SCANN_D) atomicAdd(a.dbk + tid, s_dbk[tid]);
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// Geometry backward
// =============================================================================================
struct LaGeomBwdArgs {
    const int32_t* ntiles; const int32_t* tile_a0; const int32_t* tile_a1;
    const int32_t* cnt; const int32_t* rowptr; const int32_t* pair_c; const int32_t* pair_j;
    const float* g_in;       // [rows,128] layer input geometry g
    float* prebuf;           // [rows,128] in: filter_geo pre-activation ; out: d_pre
    const float* dg_tot;     // [rows,128] gradient w.r.t. g' (upstream + attention part)
    const float* W2T;        // transposed block: W2T[n][k] = W2[k][n]
    const float* gamma_g;
    float* dg_out;           // [rows,128] gradient w.r.t. g
    float* s_pre;            // [R,128]  <- sum_n d_pre (atoms with pairs)
    float* t_scatter;        // [R,128]  += d_pre at the neighbour rows
    float* dgamma_g; float* dbeta_g;
    int mma_rows;
};

template <int NG>
__global__ void __launch_bounds__(LTC_THREADS, 1) la_geom_bwd_tc_kernel(const LaGeomBwdArgs a) {
    using G = LaGroups<NG>;
    constexpr int TR = G::TR, WG = G::WG, GT = G::GT;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[NG];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_acc[2 * SCANN_D];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp / WG, wg = warp % WG, gtid = tid - grp * GT;
    uint8_t* sHi = smem + (size_t)grp * 3 * G::IMG;
    uint8_t* sLo = sHi + G::IMG;
    uint8_t* sS = sLo + G::IMG;
    uint64_t* bar = &bars[grp];
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid < NG) mbar_init(&bars[tid], 1);
    if (tid == 0) mbar_fence_init();
    if (tid < 2 * SCANN_D) s_acc[tid] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t t_whi = tmem, t_wlo = tmem + 128, t_dm = tmem + 256 + grp * 2 * TR, t_dc = t_dm + TR;
    // stationary operand A[M = k][K = n] = W2[k][n]: (d_pre W2^T)^T = W2 d_pre^T
    b_weightT_to_tmem(a.W2T, t_whi, t_wlo, warp, lane);
    const float4 gam = ldg4(a.gamma_g + lane * 4);
    pdl_wait();
    const int nt = *a.ntiles;
    tc_fence_before();
    __syncthreads();                                     // the whole weight is in tensor memory for every group
    tc_fence_after();
    float4 dgam = make_float4(0.f, 0.f, 0.f, 0.f), dbet = dgam;
    uint32_t phase = 0;
    const int t_step = gridDim.x * NG;
    for (int t = blockIdx.x * NG + grp; t < nt; t += t_step) {
        const size_t rowbase = (size_t)t * TR;
        int pc[LTC_RPW];
        float4 dz[LTC_RPW];
#pragma unroll
        for (int i = 0; i < LTC_RPW; ++i) pc[i] = a.pair_c[rowbase + wg + WG * i];
        // ---- phase A: recompute z statistics, LN_g backward, d_pre
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            float4 pv[4], gv[4], dv[4];
            int jj[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = hb * 4 + ii, r = wg + WG * i;
                pv[ii] = make_float4(0.f, 0.f, 0.f, 0.f); gv[ii] = pv[ii]; dv[ii] = pv[ii];
                jj[ii] = 0;
                if (pc[i] >= 0) {
                    jj[ii] = a.pair_j[rowbase + r];
                    pv[ii] = ld4(a.prebuf + (rowbase + r) * SCANN_D + lane * 4);
                    gv[ii] = ld4(a.g_in + (rowbase + r) * SCANN_D + lane * 4);
                    dv[ii] = ld4(a.dg_tot + (rowbase + r) * SCANN_D + lane * 4);
                }
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = hb * 4 + ii, r = wg + WG * i;
                if (pc[i] < 0) { dz[i] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }      // padding row
                const float pre[4] = {pv[ii].x, pv[ii].y, pv[ii].z, pv[ii].w};
                const float g[4] = {gv[ii].x, gv[ii].y, gv[ii].z, gv[ii].w};
                const float dgt[4] = {dv[ii].x, dv[ii].y, dv[ii].z, dv[ii].w};
                const float gm[4] = {gam.x, gam.y, gam.z, gam.w};
                float z[4], sg[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { sg[q] = sigmoid_fast(pre[q]); z[q] = pre[q] * sg[q] + g[q]; }
                float s1 = z[0] + z[1] + z[2] + z[3];
                const float sh = __shfl_sync(0xffffffffu, s1, 0) * 0.25f;
                float d[4] = {z[0] - sh, z[1] - sh, z[2] - sh, z[3] - sh};
                float m1 = d[0] + d[1] + d[2] + d[3], m2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3];
                warp_sum2(m1, m2);
                m1 *= (1.0f / SCANN_D);
                const float inv = rsqrtf(fmaxf(m2 * (1.0f / SCANN_D) - m1 * m1, 0.f) + SCANN_LN_EPS);
                float xh[4], dxh[4];
                float t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    xh[q] = (d[q] - m1) * inv;
                    dxh[q] = dgt[q] * gm[q];
                    t1 += dxh[q];
                    t2 = fmaf(dxh[q], xh[q], t2);
                }
                warp_sum2(t1, t2);
                t1 *= (1.0f / SCANN_D);
                t2 *= (1.0f / SCANN_D);
                float dzv[4], dp[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    dzv[q] = pc[i] >= 0 ? inv * (dxh[q] - t1 - xh[q] * t2) : 0.f;
                    dp[q] = dzv[q] * sg[q] * (1.0f + pre[q] * (1.0f - sg[q]));        // swish'(pre)
                }
                dgam = make_float4(fmaf(dgt[0], xh[0], dgam.x), fmaf(dgt[1], xh[1], dgam.y), fmaf(dgt[2], xh[2], dgam.z),
                                   fmaf(dgt[3], xh[3], dgam.w));
                dbet = make_float4(dbet.x + dgt[0], dbet.y + dgt[1], dbet.z + dgt[2], dbet.w + dgt[3]);
                dz[i] = make_float4(dzv[0], dzv[1], dzv[2], dzv[3]);
                const float4 dpre = make_float4(dp[0], dp[1], dp[2], dp[3]);
                split_store(sHi, sLo, tc_off4(r, lane), dpre);
                st4(a.prebuf + (rowbase + r) * SCANN_D + lane * 4, dpre);
                if (pc[i] >= 0) red_add4(a.t_scatter + (size_t)jj[ii] * SCANN_D + lane * 4, dp[0], dp[1], dp[2], dp[3]);
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(grp, GT);
        if (wg == 0 && tc_elect_one()) {
            tc_fence_after();
            b_issue_3xtf32(t_whi, t_wlo, smem_u32(sHi), smem_u32(sLo), t_dm, t_dc, bar, a.mma_rows);     // W2 d_pre^T
        }
        // ---- phase B (warp per atom, overlaps the MMA): s_pre[c] = sum_n d_pre
        const int a0 = a.tile_a0[t], a1 = a.tile_a1[t];
        for (int atom = a0 + wg; atom < a1; atom += WG) {
            const int n = a.cnt[atom];
            if (n == 0) continue;
            const int r0 = a.rowptr[atom] - (int)rowbase;
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            for (int r = 0; r < n; ++r) {
                const uint32_t off = tc_off4(r0 + r, lane);
                float4 h = *reinterpret_cast<const float4*>(sHi + off), l = *reinterpret_cast<const float4*>(sLo + off);
                c0 += h.x + l.x; c1 += h.y + l.y; c2 += h.z + l.z; c3 += h.w + l.w;
            }
            st4(a.s_pre + (size_t)atom * SCANN_D + lane * 4, make_float4(c0, c1, c2, c3));
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        if (t + t_step >= nt) pdl_trigger();             // last tile of this group, only its epilogue is left
        b_tmem_to_rows(t_dm, t_dc, sS, wg, lane, a.mma_rows);
        tc_fence_before();
        group_sync(grp, GT);
        // ---- phase C: dg = d_z + d_pre @ W2^T
#pragma unroll
        for (int i = 0; i < LTC_RPW; ++i) {
            const int r = wg + WG * i;
            if (pc[i] < 0) continue;
            float4 v = *reinterpret_cast<const float4*>(sS + tc_off4(r, lane));
            st4(a.dg_out + (rowbase + r) * SCANN_D + lane * 4,
                make_float4(v.x + dz[i].x, v.y + dz[i].y, v.z + dz[i].z, v.w + dz[i].w));
        }
        group_sync(grp, GT);
    }
    pdl_trigger();
    atomicAdd(&s_acc[lane * 4 + 0], dgam.x); atomicAdd(&s_acc[lane * 4 + 1], dgam.y);
    atomicAdd(&s_acc[lane * 4 + 2], dgam.z); atomicAdd(&s_acc[lane * 4 + 3], dgam.w);
    atomicAdd(&s_acc[SCANN_D + lane * 4 + 0], dbet.x); atomicAdd(&s_acc[SCANN_D + lane * 4 + 1], dbet.y);
    atomicAdd(&s_acc[SCANN_D + lane * 4 + 2], dbet.z); atomicAdd(&s_acc[SCANN_D + lane * 4 + 3], dbet.w);
    tc_fence_before();
    __syncthreads();
    if (tid < SCANN_D) {
        atomicAdd(a.dgamma_g + tid, s_acc[tid]);
        atomicAdd(a.dbeta_g + tid, s_acc[SCANN_D + tid]);
    }
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// Pair weight gradients:  mode 0: part[cta][0] = sum_tiles (x[j]*g')^T d_k ; mode 1: part[cta][1] = g^T d_pre
// =============================================================================================
struct LaWgradArgs {
    const int32_t* ntiles; const int32_t* pair_c; const int32_t* pair_j;
    const float* x;          // [R,128]        (mode 0)
    const float* xsrc;       // [rows,128] g' (mode 0) or g (mode 1)
    const float* ysrc;       // [rows,128] d_k (mode 0) or d_pre (mode 1)
    int mode;
    float* wpart;            // [grid][2][128][128]
};

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// Each warp group accumulates X^T Y of its own tile stream in its own pair of accumulators (main, correction);
// group 0 adds the other groups' sums (handed over through shared memory) and writes wpart[cta][mode].
template <int NG>
__global__ void __launch_bounds__(LTC_THREADS, 1) la_wgrad_tc_kernel(const LaWgradArgs a) {
    using G = LaGroups<NG>;
    constexpr int TR = G::TR, WG = G::WG, GT = G::GT;
    constexpr uint32_t MNB = (uint32_t)TR * 128u;        // one column block: [TR rows x 32 columns], row pitch 128 B
    constexpr uint32_t MNT = 4u * MNB;                   // one MN-major image [TR x 128] fp32
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // the BASE32B swizzle is a function of the absolute shared address: align the images to 1024 bytes
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bars[NG];
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_has[NG];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp / WG, wg = warp % WG, gtid = tid - grp * GT;
    uint8_t* sX = smem + (size_t)grp * 3 * MNT;          // X_hi, then X_lo
    uint8_t* sYh = sX + MNT;
    uint8_t* sYl = sYh + MNT;
    uint64_t* bar = &bars[grp];
    const int nt = *a.ntiles;
    if ((int)blockIdx.x * NG >= nt) return;              // no tile for any group of this CTA
    if (warp == 0) tmem_alloc(&tmem_base_s, 256 * NG);
    if (tid < NG) mbar_init(&bars[tid], 1);
    if (tid == 0) mbar_fence_init();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t_dm = tmem_base_s + grp * 256, t_dc = t_dm + 128;
    const uint32_t idesc = tc_idesc_tf32(128, 128, true, true);
    auto mn_off = [](int r, int c) -> uint32_t {
        return (uint32_t)(c >> 5) * MNB + (uint32_t)r * 128u + (((((uint32_t)c >> 3) & 3u) ^ ((uint32_t)r & 3u)) << 5) +
               ((uint32_t)c & 7u) * 4u;
    };
    auto mn_desc = [](uint32_t saddr) -> uint64_t { return tc_desc(saddr, MNB, 512u) | ((uint64_t)1 << 61); };
    uint32_t phase = 0;
    bool first = true;
    for (int t = blockIdx.x * NG + grp; t < nt; t += gridDim.x * NG) {
        const size_t rowbase = (size_t)t * TR;
        float4 xv[LTC_RPW], yv[LTC_RPW];
#pragma unroll
        for (int i = 0; i < LTC_RPW; ++i) {
            const int r = wg + WG * i;
            const int c = a.pair_c[rowbase + r];
            xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            yv[i] = xv[i];
            if (c >= 0) {
                xv[i] = ld4(a.xsrc + (rowbase + r) * SCANN_D + lane * 4);
                yv[i] = ld4(a.ysrc + (rowbase + r) * SCANN_D + lane * 4);
                if (a.mode == 0) {
                    float4 nb = ld4(a.x + (size_t)a.pair_j[rowbase + r] * SCANN_D + lane * 4);
                    xv[i] = make_float4(xv[i].x * nb.x, xv[i].y * nb.y, xv[i].z * nb.z, xv[i].w * nb.w);
                }
            }
        }
        if (!first) {                         // the previous tile's last MMAs still read the images
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
        }
#pragma unroll
        for (int i = 0; i < LTC_RPW; ++i) {
            const uint32_t off = mn_off(wg + WG * i, lane * 4);
            float4 h, l;
            tf32_split(xv[i].x, h.x, l.x); tf32_split(xv[i].y, h.y, l.y);
            tf32_split(xv[i].z, h.z, l.z); tf32_split(xv[i].w, h.w, l.w);
            *reinterpret_cast<float4*>(sX + off) = h;
            xv[i] = l;
            split_store(sYh, sYl, off, yv[i]);
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(grp, GT);
        if (wg == 0 && tc_elect_one()) {
            tc_fence_after();
            const uint64_t dx = mn_desc(smem_u32(sX)), dyh = mn_desc(smem_u32(sYh)), dyl = mn_desc(smem_u32(sYl));
#pragma unroll
            for (int ks = 0; ks < TR / 8; ++ks)       // K-step = 8 pair rows = 1024 bytes of each image
                tc_mma_ss(t_dm, dx + (uint64_t)(ks * 64), dyh + (uint64_t)(ks * 64), idesc, !(first && ks == 0));
#pragma unroll
            for (int ks = 0; ks < TR / 8; ++ks)
                tc_mma_ss(t_dc, dx + (uint64_t)(ks * 64), dyl + (uint64_t)(ks * 64), idesc, !(first && ks == 0));
            tc_commit(bar);
        }
        mbar_wait(bar, phase);               // X_hi has been consumed: replace it by X_lo
        phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < LTC_RPW; ++i)
            *reinterpret_cast<float4*>(sX + mn_off(wg + WG * i, lane * 4)) = xv[i];
        fence_async_smem();
        tc_fence_before();
        group_sync(grp, GT);
        if (wg == 0 && tc_elect_one()) {
            tc_fence_after();
            const uint64_t dx = mn_desc(smem_u32(sX)), dyh = mn_desc(smem_u32(sYh));
#pragma unroll
            for (int ks = 0; ks < TR / 8; ++ks) tc_mma_ss(t_dc, dx + (uint64_t)(ks * 64), dyh + (uint64_t)(ks * 64), idesc, true);
            tc_commit(bar);
        }
        first = false;
    }
    if (!first) {
        mbar_wait(bar, phase);
        tc_fence_after();
    }
    if (gtid == 0) s_has[grp] = first ? 0 : 1;
    // D[m][n] (lane = m, column = n): warp wg of a group covers lanes 32*(wg%4).. and CPW columns from (wg/4)*CPW
    constexpr int CPW = 512 / WG;
    const int m = (wg & 3) * 32 + lane, nbase = (wg >> 2) * CPW;
    const uint32_t lane_base = (uint32_t)((wg & 3) * 32) << 16;
    float4* xfer = reinterpret_cast<float4*>(smem + (size_t)3 * MNT);     // groups >= 1 own this region (their images)
    if (NG > 1 && grp > 0 && !first) {
#pragma unroll 1
        for (int h = 0; h < CPW / 16; ++h) {
            float v[16], c[16];
            tmem_ld16(t_dm + lane_base + nbase + h * 16, v);
            tmem_ld16(t_dc + lane_base + nbase + h * 16, c);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 16; q += 4)
                xfer[(size_t)((grp - 1) * (CPW / 4) + h * 4 + q / 4) * GT + gtid] =
                    make_float4(v[q] + c[q], v[q + 1] + c[q + 1], v[q + 2] + c[q + 2], v[q + 3] + c[q + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (grp == 0) {
        float* dst = a.wpart + ((size_t)blockIdx.x * 2 + a.mode) * SCANN_D * SCANN_D;
#pragma unroll 1
        for (int h = 0; h < CPW / 16; ++h) {
            float v[16], c[16];
            if (!first) {
                tmem_ld16(t_dm + lane_base + nbase + h * 16, v);
                tmem_ld16(t_dc + lane_base + nbase + h * 16, c);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) { v[q] = 0.f; c[q] = 0.f; }
            }
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
                float4 o = make_float4(v[q] + c[q], v[q + 1] + c[q + 1], v[q + 2] + c[q + 2], v[q + 3] + c[q + 3]);
                for (int og = 1; og < NG; ++og) {
                    if (!s_has[og]) continue;
                    const float4 p = xfer[(size_t)((og - 1) * (CPW / 4) + h * 4 + q / 4) * GT + gtid];
                    o = make_float4(o.x + p.x, o.y + p.y, o.z + p.z, o.w + p.w);
                }
                st4(dst + (size_t)m * SCANN_D + nbase + h * 16 + q, o);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 256 * NG);
}

#define LA_ATTN_BWD_SMEM (3 * TC_TILE_BYTES + 2 * SCANN_TILE * 8 * sizeof(float))
#define LA_GEOM_BWD_SMEM (3 * TC_TILE_BYTES)
#define LA_WGRAD_SMEM (3 * TC_MN_TILE_BYTES + 1024)

// Tensor-core backward of LocalAttention (TF autodiff of attention.py:118-216 inside keras fit).
// kbuf / prebuf: in = keys / filter_geo pre-activation saved by scann_la_forward_tc, out = d_k / d_pre.
// dg: gradient w.r.t. g' from the next layer (dg_has_up != 0) or scratch that is overwritten.
// s_pre is written for atoms with pairs; t_scatter / dx_scatter are accumulated (pre-zero them).
// The pair weight gradients are a separate call (scann_la_wgrad_tc); wpart is unused here.
static int la_bwd_configure() {
    static bool configured = false;
    if (configured) return 0;
    cudaError_t e = cudaFuncSetAttribute(la_attn_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_ATTN_BWD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la_attn_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_ATTN_BWD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la_geom_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_GEOM_BWD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la_geom_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_GEOM_BWD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la_attn_bwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_ATTN_BWD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la_geom_bwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_GEOM_BWD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la_wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_WGRAD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la_wgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_WGRAD_SMEM);
    if (e != cudaSuccess) { scann_set_error("la_backward_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return 1; }
    configured = true;
    return 0;
}

// which: bit 0 attention kernel, bit 1 geometry kernel (scann_la_backward_tc = both; the halves are separate entry
// points so that either can be replaced by its pipelined form, la_pipe_bwd.cu).
extern "C" int scann_la_backward_tc_part(int which, int grid, int tile_stride, int mma_rows, const int32_t* ntiles, const int32_t* tile_a0,
                                    const int32_t* tile_a1, const int32_t* cnt, const int32_t* rowptr,
                                    const int32_t* pair_c, const int32_t* pair_j, const float* x, const float* proj,
                                    const float* g_in, const float* g_new, float* kbuf, float* prebuf, const float* W2T,
                                    const float* WkT, const float* gamma_g, const float* d_ctx, float* dg, int dg_has_up,
                                    float* dg_out, float* dq, float* s_pre, float* t_scatter, float* dx_scatter,
                                    float* wpart, float* dgamma_g, float* dbeta_g, float* dbk, const void* attn_drop,
                                    int drop_site, void* stream) {
    (void)wpart;
    if (tile_stride != 32 && tile_stride != 64 && tile_stride != 128) { scann_set_error("la_backward_tc: tile_stride must be 32, 64 or 128"); return 1; }
    if (mma_rows < 16 || mma_rows > tile_stride || mma_rows % 16) { scann_set_error("la_backward_tc: bad mma_rows"); return 1; }
    if (la_bwd_configure()) return 1;
    if (grid <= 0) return 0;
    LaAttnBwdArgs ab{ntiles, tile_a0, tile_a1, cnt, rowptr, pair_c, pair_j, x, proj, g_new, kbuf, WkT, d_ctx, dg,
                     dg_has_up, dq, dx_scatter, dbk, mma_rows, (const ScannDropCtl*)attn_drop, drop_site};
    LaGeomBwdArgs gb{ntiles, tile_a0, tile_a1, cnt, rowptr, pair_c, pair_j, g_in, prebuf, dg, W2T, gamma_g, dg_out,
                     s_pre, t_scatter, dgamma_g, dbeta_g, mma_rows};
    if (tile_stride == 32) {
        if (which & 1) scann_launch(la_attn_bwd_tc_kernel<4>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_BWD_SMEM, stream, ab);
        if (which & 2) scann_launch(la_geom_bwd_tc_kernel<4>, dim3(grid), dim3(LTC_THREADS), LA_GEOM_BWD_SMEM, stream, gb);
    } else if (tile_stride == 64) {
        if (which & 1) scann_launch(la_attn_bwd_tc_kernel<2>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_BWD_SMEM, stream, ab);
        if (which & 2) scann_launch(la_geom_bwd_tc_kernel<2>, dim3(grid), dim3(LTC_THREADS), LA_GEOM_BWD_SMEM, stream, gb);
    } else {
        if (which & 1) scann_launch(la_attn_bwd_tc_kernel<1>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_BWD_SMEM, stream, ab);
        if (which & 2) scann_launch(la_geom_bwd_tc_kernel<1>, dim3(grid), dim3(LTC_THREADS), LA_GEOM_BWD_SMEM, stream, gb);
    }
    return scann_check_launch("scann_la_backward_tc");
}

extern "C" int scann_la_backward_tc(int grid, int tile_stride, int mma_rows, const int32_t* ntiles, const int32_t* tile_a0,
                                    const int32_t* tile_a1, const int32_t* cnt, const int32_t* rowptr,
                                    const int32_t* pair_c, const int32_t* pair_j, const float* x, const float* proj,
                                    const float* g_in, const float* g_new, float* kbuf, float* prebuf, const float* W2T,
                                    const float* WkT, const float* gamma_g, const float* d_ctx, float* dg, int dg_has_up,
                                    float* dg_out, float* dq, float* s_pre, float* t_scatter, float* dx_scatter,
                                    float* wpart, float* dgamma_g, float* dbeta_g, float* dbk, const void* attn_drop,
                                    int drop_site, void* stream) {
    return scann_la_backward_tc_part(3, grid, tile_stride, mma_rows, ntiles, tile_a0, tile_a1, cnt, rowptr, pair_c, pair_j, x,
                                     proj, g_in, g_new, kbuf, prebuf, W2T, WkT, gamma_g, d_ctx, dg, dg_has_up, dg_out, dq,
                                     s_pre, t_scatter, dx_scatter, wpart, dgamma_g, dbeta_g, dbk, attn_drop, drop_site, stream);
}

// Backward of LocalAttention.call with g_update = False (attention.py:155-216): the attention kernel only
// (softmax / context / key projection backward; d_nbr scattered to dx_scatter; dg <- gradient w.r.t.
// g' = swish(rbf @ Wf + bf) * w, consumed by scann_noupdate_geom_backward).  g_new / kbuf: the g' and keys saved
// by scann_la_forward_noupdate_tc; kbuf is overwritten with d_k (left operand gradient for the key kernel).
extern "C" int scann_la_backward_noupdate_tc(int grid, int tile_stride, int mma_rows, const int32_t* ntiles, const int32_t* tile_a0,
                                             const int32_t* tile_a1, const int32_t* cnt, const int32_t* rowptr,
                                             const int32_t* pair_c, const int32_t* pair_j, const float* x,
                                             const float* proj, const float* g_new, float* kbuf, const float* WkT,
                                             const float* d_ctx, float* dg, float* dq, float* dx_scatter, float* dbk,
                                             const void* attn_drop, int drop_site, void* stream) {
    if (tile_stride != 32 && tile_stride != 64 && tile_stride != 128) { scann_set_error("la_backward_noupdate_tc: tile_stride must be 32, 64 or 128"); return 1; }
    if (mma_rows < 16 || mma_rows > tile_stride || mma_rows % 16) { scann_set_error("la_backward_noupdate_tc: bad mma_rows"); return 1; }
    if (la_bwd_configure()) return 1;
    if (grid <= 0) return 0;
    LaAttnBwdArgs ab{ntiles, tile_a0, tile_a1, cnt, rowptr, pair_c, pair_j, x, proj, g_new, kbuf, WkT, d_ctx, dg,
                     0, dq, dx_scatter, dbk, mma_rows, (const ScannDropCtl*)attn_drop, drop_site};
    if (tile_stride == 32) scann_launch(la_attn_bwd_tc_kernel<4>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_BWD_SMEM, stream, ab);
    else if (tile_stride == 64) scann_launch(la_attn_bwd_tc_kernel<2>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_BWD_SMEM, stream, ab);
    else scann_launch(la_attn_bwd_tc_kernel<1>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_BWD_SMEM, stream, ab);
    return scann_check_launch("scann_la_backward_noupdate_tc");
}

// Pair weight gradients of one LocalAttention layer into wpart[grid][2][128][128] (off the critical path of
// the backward chain: may run on a side stream once scann_la_backward_tc of the layer has finished):
//   wpart[.][0] = sum (x[j]*g')^T d_k (-> key/kernel),  wpart[.][1] = sum g^T d_pre (-> filter_geo rows 128..255)
// CTA c holds a valid partial iff c * (128 / tile_stride) < ntiles (scann_la_wpart_reduce applies the same rule).
extern "C" int scann_la_wgrad_tc(int grid, int tile_stride, const int32_t* ntiles, const int32_t* pair_c,
                                 const int32_t* pair_j, const float* x, const float* g_in, const float* g_new,
                                 const float* dk, const float* dpre, float* wpart, void* stream) {
    if (tile_stride != 64 && tile_stride != 128) { scann_set_error("la_wgrad_tc: tile_stride must be 64 or 128"); return 1; }
    if (la_bwd_configure()) return 1;
    if (grid <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    LaWgradArgs w0{ntiles, pair_c, pair_j, x, g_new, dk, 0, wpart};
    LaWgradArgs w1{ntiles, pair_c, pair_j, x, g_in, dpre, 1, wpart};
    if (tile_stride == 64) {
        la_wgrad_tc_kernel<2><<<grid, LTC_THREADS, LA_WGRAD_SMEM, st>>>(w0);
        la_wgrad_tc_kernel<2><<<grid, LTC_THREADS, LA_WGRAD_SMEM, st>>>(w1);
    } else {
        la_wgrad_tc_kernel<1><<<grid, LTC_THREADS, LA_WGRAD_SMEM, st>>>(w0);
        la_wgrad_tc_kernel<1><<<grid, LTC_THREADS, LA_WGRAD_SMEM, st>>>(w1);
    }
    return scann_check_launch("scann_la_wgrad_tc");
}
