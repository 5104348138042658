// Local attention forward on the tcgen05 tensor cores (g_update = True), two persistent kernels:
//
//   la_geom_fwd_tc : g' = LN_g(swish(xw1[c] + g @ W2 + xw3[j]) + g)          (attention.py:141-153)
//   la_attn_fwd_tc : k = (x[j] * g') @ Wk + bk ; softmax over the atom's pairs ; out = LN(sum p k + q)
//                                                                           (attention.py:157-214)
//
// Each kernel keeps ONE 128x128 weight block stationary in tensor memory as the M x K operand
// (W^T: lane = output feature, column = input feature; hi and lo tf32 parts, 256 columns) and streams
// 128-pair tiles through shared memory as the N x K operand (canonical K-major image, tc_common.cuh).
// D^T = W^T @ X^T lands in tensor memory (lane = feature, column = pair row) in ONE accumulator: the two
// correction products (lo*hi, hi*lo) are issued before the hi*hi product (see issue_3xtf32).  The accumulator is moved to
// shared memory transposed and the row-wise epilogue runs warp-per-row with coalesced gathers that
// were prefetched into registers while the tensor core was busy.
//
// Both weights of a layer (4 x 64 KB as hi/lo tf32) exceed one SM's shared + tensor memory next to the
// tile operands, hence two kernels; g' makes one extra round trip through L2/HBM.
//
// Warp groups.  A tile's phases (global loads -> operand images -> MMA -> TMEM read-back -> row epilogue)
// are a dependent chain that leaves the SM idle most of the time (ncu: 20-27 % issue utilisation, long-
// scoreboard and barrier stalls).  The kernels are therefore templated on NG = number of independent WARP
// GROUPS per CTA: the 16 warps split into NG groups of 16/NG warps, each group owns its own tile stream
// (tiles of TR = 128/NG rows, tile t -> group t % NG of CTA (t / NG) % grid), its own operand images,
// mbarrier and accumulator columns, and synchronises with a named barrier (bar.sync 1+g).  The groups share
// the stationary weight in tensor memory and drift out of phase, so one group's MMA / memory latency hides
// behind the other's epilogue -- two CTAs per SM would do the same but cannot share the 256 weight columns.
// NG = 1 is the original single-stream kernel (needed when an atom has more than 64 neighbours).
#include "common.cuh"
#include "tc_common.cuh"

#define LTC_THREADS 512
#define LTC_WARPS 16
#define LTC_RPW 8                          // rows per warp in the row-wise epilogue (= TR / warps per group)

// phase timestamps of CTA 0 (clock64), read back with scann_debug_clocks: development aid
#ifdef SCANN_DEV_PROBES
__device__ long long g_dbg_clk[32];
#define DBG_CLK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_dbg_clk[i] = clock64(); } while (0)
#else
#define DBG_CLK(i) do { } while (0)
#endif

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// W[k][n] (row-major, ld 128) -> tensor memory as A[M = n][K = k], hi and lo parts.
// warp w: lane quarter w%4 (features 32*(w%4)..), k range 32*(w/4)..+31
__device__ __forceinline__ void weightT_to_tmem(const float* __restrict__ W, uint32_t t_hi, uint32_t t_lo, int warp,
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

// D_main = W_hi X_hi^T ; D_corr = W_lo X_hi^T + W_hi X_lo^T   (one thread).
// Fully unrolled with the descriptors advanced by immediates: a single thread issues all MMAs, so
// every extra instruction per MMA shows up as tensor-pipe idle time (measured: 107 cycles per MMA
// with descriptors rebuilt in the loop vs the 64-cycle math floor of a 128x128x8 tf32 MMA).
// nrows (multiple of 16, <= tile slot): rows of the tile that can hold pairs = N extent of the MMAs
__device__ __forceinline__ void issue_3xtf32(uint32_t t_whi, uint32_t t_wlo, uint32_t xh, uint32_t xl, uint32_t t_dm,
                                             uint32_t t_dc, uint64_t* bar, int nrows) {
    const uint32_t idesc = tc_idesc_tf32(128, nrows, false, false);
    const uint64_t dh = tc_desc_kmajor(xh, 0), dl = tc_desc_kmajor(xl, 0);
    // ONE accumulator, the two correction products first: the small terms are summed among themselves before the
    // large hi*hi terms arrive, which is as accurate as a separate correction accumulator (tc_probe nprod 5:
    // 3.7e-6 vs 3.4e-6; main-first 1.5e-5) and halves the accumulator columns and the read-back
    (void)t_dc;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_dm, t_wlo + ks * 8, dh + ks * TC_KSTEP_DESC, idesc, ks != 0);
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_dm, t_whi + ks * 8, dl + ks * TC_KSTEP_DESC, idesc, true);
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_dm, t_whi + ks * 8, dh + ks * TC_KSTEP_DESC, idesc, true);
    tc_commit(bar);
}

// accumulators (lane = feature n, column = row r) -> S[r][n] (+ bias[n]); warp w of its group: lanes 32*(w%4)..,
// rows 32*(w/4)..  (the group's 16/NG warps cover its 128/NG rows)
__device__ __forceinline__ void tmem_to_rows(uint32_t t_dm, uint32_t t_dc, uint8_t* S, const float* __restrict__ bias,
                                             int warp, int lane, int nrows) {
    const int n = (warp & 3) * 32 + lane, rbase = (warp >> 2) * 32;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float b = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (rbase + h * 16 >= nrows) break;              // warp-uniform: rows beyond nrows are never valid
        float m[16];
        (void)t_dc;
        tmem_ld16(t_dm + lane_base + rbase + h * 16, m);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) *reinterpret_cast<float*>(S + tc_off(rbase + h * 16 + q, n)) = m[q] + b;
    }
}

// =============================================================================================
// Geometry update
// =============================================================================================
struct LaGeomArgs {
    const int32_t* ntiles; const int32_t* pair_c; const int32_t* pair_j;
    const float* proj;       // [R,384] = [x@W1+bf | x@W3 | x@Wq+bq]
    const float* g_in;       // [rows,128]
    const float* W2;         // rows 128..255 of filter_geo/kernel, [k][n]
    const float* gamma_g; const float* beta_g;
    float* g_out;            // [rows,128]
    float* pre_out;          // [rows,128] pre-activation of filter_geo (nullable; saved for backward)
    int mma_rows;            // rows of a tile slot that can hold pairs (multiple of 16): wave-balanced plans fill less
    int skew;                // tc4 kernels: stagger the start of the warp groups
};

template <int NG>
__global__ void __launch_bounds__(LTC_THREADS, 1) la_geom_fwd_tc_kernel(const LaGeomArgs a) {
    using G = LaGroups<NG>;
    constexpr int TR = G::TR, WG = G::WG, GT = G::GT;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[NG];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp / WG, wg = warp % WG, gtid = tid - grp * GT;
    uint8_t* sHi = smem + (size_t)grp * 3 * G::IMG;
    uint8_t* sLo = sHi + G::IMG;
    uint8_t* sS = sLo + G::IMG;
    uint64_t* bar = &bars[grp];
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid < NG) mbar_init(&bars[tid], 1);
    if (tid == 0) mbar_fence_init();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t t_whi = tmem, t_wlo = tmem + 128, t_dm = tmem + 256 + grp * 2 * TR, t_dc = t_dm + TR;
    DBG_CLK(0);
    weightT_to_tmem(a.W2, t_whi, t_wlo, warp, lane);     // parameters only: overlaps the predecessor kernel's tail
    pdl_wait();
    const int nt = *a.ntiles;
    tc_fence_before();
    __syncthreads();                                     // the whole weight is in tensor memory for every group
    tc_fence_after();
    DBG_CLK(1);
    uint32_t phase = 0;
    const int t_first = blockIdx.x * NG + grp, t_step = gridDim.x * NG;
    for (int t = t_first; t < nt; t += t_step) {
        const size_t rowbase = (size_t)t * TR;
        // ---- stage the geometry tile: coalesced loads, hi/lo split, K-major images
        {
            float4 v[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int i = gtid + it * GT;
                v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((i >> 5) < a.mma_rows) v[it] = ld4(a.g_in + (rowbase + (i >> 5)) * SCANN_D + (i & 31) * 4);
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int i = gtid + it * GT;
                float4 h, l;
                tf32_split(v[it].x, h.x, l.x); tf32_split(v[it].y, h.y, l.y);
                tf32_split(v[it].z, h.z, l.z); tf32_split(v[it].w, h.w, l.w);
                const uint32_t off = tc_off4(i >> 5, i & 31);
                *reinterpret_cast<float4*>(sHi + off) = h;
                *reinterpret_cast<float4*>(sLo + off) = l;
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(grp, GT);
        if (t == t_first) DBG_CLK(2);
        if (wg == 0 && tc_elect_one()) {
            tc_fence_after();
            issue_3xtf32(t_whi, t_wlo, smem_u32(sHi), smem_u32(sLo), t_dm, t_dc, bar, a.mma_rows);
        }
        if (t == t_first) DBG_CLK(3);
        // ---- while the tensor core works: indices and gathered projections of this warp's rows
        // (row groups: 2 steps x 4 rows per warp, 8 lanes per row, 16 columns per lane)
        const int l8 = lane & 7, rsub = lane >> 3;
        int pc[2];
        float4 p13[2][4];
#pragma unroll
        for (int sp = 0; sp < 2; ++sp) {
            const int r = sp * (TR / 2) + wg * 4 + rsub;
            pc[sp] = a.pair_c[rowbase + r];
            if (__all_sync(0xffffffffu, pc[sp] < 0)) continue;
            const int j = pc[sp] >= 0 ? a.pair_j[rowbase + r] : 0;
            const int c = pc[sp] >= 0 ? pc[sp] : 0;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int c0 = (l8 + 8 * it) * 4;
                p13[sp][it] = f4add(ld4(a.proj + (size_t)c * 3 * SCANN_D + c0),
                                    ld4(a.proj + (size_t)j * 3 * SCANN_D + SCANN_D + c0));
            }
        }
        if (t == t_first) DBG_CLK(4);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        if (t + t_step >= nt) pdl_trigger();             // last tile of this group, only its epilogue is left
        if (t == t_first) DBG_CLK(5);
        tmem_to_rows(t_dm, t_dc, sS, nullptr, wg, lane, a.mma_rows);
        tc_fence_before();
        group_sync(grp, GT);
        if (t == t_first) DBG_CLK(6);
        // ---- row-wise epilogue: pre -> swish -> + g -> LayerNorm -> g'
#pragma unroll
        for (int sp = 0; sp < 2; ++sp) {
            if (__all_sync(0xffffffffu, pc[sp] < 0)) continue;        // the warp's 4 rows are padding
            const int r = sp * (TR / 2) + wg * 4 + rsub;
            float z[4][4], pre[4][4];
            float s1 = 0.f;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const uint32_t off = tc_off4(r, l8 + 8 * it);
                const float4 acc = *reinterpret_cast<const float4*>(sS + off);
                const float4 gh = *reinterpret_cast<const float4*>(sHi + off), gl = *reinterpret_cast<const float4*>(sLo + off);
                pre[it][0] = acc.x + p13[sp][it].x; pre[it][1] = acc.y + p13[sp][it].y;
                pre[it][2] = acc.z + p13[sp][it].z; pre[it][3] = acc.w + p13[sp][it].w;
                z[it][0] = swish_fast(pre[it][0]) + (gh.x + gl.x); z[it][1] = swish_fast(pre[it][1]) + (gh.y + gl.y);
                z[it][2] = swish_fast(pre[it][2]) + (gh.z + gl.z); z[it][3] = swish_fast(pre[it][3]) + (gh.w + gl.w);
                s1 += z[it][0] + z[it][1] + z[it][2] + z[it][3];
            }
            // one-pass moments around a common shift (the first lane's partial mean)
            const float sh = __shfl_sync(0xffffffffu, s1, lane & 24) * (1.0f / 16.0f);
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int it = 0; it < 4; ++it)
#pragma unroll
                for (int q = 0; q < 4; ++q) { z[it][q] -= sh; m1 += z[it][q]; m2 = fmaf(z[it][q], z[it][q], m2); }
            oct_sum2(m1, m2);
            m1 *= (1.0f / SCANN_D);
            const float inv = rsqrtf(fmaxf(m2 * (1.0f / SCANN_D) - m1 * m1, 0.f) + SCANN_LN_EPS);
            const bool ok = pc[sp] >= 0;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int c0 = (l8 + 8 * it) * 4;
                const float4 gm = ldg4(a.gamma_g + c0), bt = ldg4(a.beta_g + c0);
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f), po = o;
                if (ok) {
                    o = make_float4((z[it][0] - m1) * inv * gm.x + bt.x, (z[it][1] - m1) * inv * gm.y + bt.y,
                                    (z[it][2] - m1) * inv * gm.z + bt.z, (z[it][3] - m1) * inv * gm.w + bt.w);
                    po = make_float4(pre[it][0], pre[it][1], pre[it][2], pre[it][3]);
                }
                st4(a.g_out + (rowbase + r) * SCANN_D + c0, o);
                if (a.pre_out) st4(a.pre_out + (rowbase + r) * SCANN_D + c0, po);
            }
        }
        group_sync(grp, GT);             // images and S are rewritten by the next tile
        if (t == t_first) DBG_CLK(7);
    }
    DBG_CLK(8);
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// Attention
// =============================================================================================
struct LaAttnArgs {
    const int32_t* ntiles; const int32_t* tile_a0; const int32_t* tile_a1;
    const int32_t* cnt; const int32_t* rowptr; const int32_t* pair_c; const int32_t* pair_j;
    const float* x;          // [R,128]
    const float* proj;       // [R,384]; q = columns 256..383
    const float* g_new;      // [rows,128] updated geometry g'
    const float* Wk; const float* bk; const float* gamma; const float* beta;
    float* ctx_pre;          // [R,128] nullable
    float* out;              // [R,128]
    float* attn;             // [rows,8] nullable
    float* k_out;            // [rows,128] nullable (keys, saved for backward)
    // g_update = False (attention.py:155): g' = swish(rbf(d) @ Wf[20,128] + bf) * w, computed on the fly
    const float* pair_d; const float* pair_w; const float* centers; const float* Wf; const float* bf;
    float* g_save;           // [rows,128] nullable: g' of the g_update = False path, saved for the backward pass
    int mma_rows;            // see LaGeomArgs
    // training-mode Dropout(0.05) on the attention probabilities (attention.py:115-116,191-192; use_drop): the
    // multiplier of (pair row, head) is drop_mult(drop, drop_site, row * 8 + head); NULL = off
    const ScannDropCtl* drop;
    int drop_site;
};

template <int NG>
__global__ void __launch_bounds__(LTC_THREADS, 1) la_attn_fwd_tc_kernel(const LaAttnArgs a) {
    using G = LaGroups<NG>;
    constexpr int TR = G::TR, WG = G::WG, GT = G::GT;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[NG];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp / WG, wg = warp % WG, gtid = tid - grp * GT;
    uint8_t* sHi = smem + (size_t)grp * 3 * G::IMG;
    uint8_t* sLo = sHi + G::IMG;
    uint8_t* sS = sLo + G::IMG;
    float* Es = reinterpret_cast<float*>(smem + 3 * TC_TILE_BYTES) + grp * TR * 8;      // [TR][8] per group
    uint64_t* bar = &bars[grp];
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid < NG) mbar_init(&bars[tid], 1);
    if (tid == 0) mbar_fence_init();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t t_whi = tmem, t_wlo = tmem + 128, t_dm = tmem + 256 + grp * 2 * TR, t_dc = t_dm + TR;
    weightT_to_tmem(a.Wk, t_whi, t_wlo, warp, lane);
    const float4 gam = ldg4(a.gamma + lane * 4), bet = ldg4(a.beta + lane * 4);
    pdl_wait();
    const int nt = *a.ntiles;
    tc_fence_before();
    __syncthreads();                                     // the whole weight is in tensor memory for every group
    tc_fence_after();
    uint32_t phase = 0;
    const int t_step = gridDim.x * NG;
    for (int t = blockIdx.x * NG + grp; t < nt; t += t_step) {
        const size_t rowbase = (size_t)t * TR;
        // row groups: 2 steps x 4 rows per warp, 8 lanes per row, 16 columns per lane
        const int l8 = lane & 7, rsub = lane >> 3;
        int pc[2];
        // ---- stage a = x[j] * g' (coalesced row loads and gathers), hi/lo images
#pragma unroll
        for (int sp = 0; sp < 2; ++sp) {
            const int r = sp * (TR / 2) + wg * 4 + rsub;
            pc[sp] = a.pair_c[rowbase + r];
            if (__all_sync(0xffffffffu, pc[sp] < 0)) continue;       // operand rows of padding stay stale: harmless,
                                                                     // column r of D^T depends on operand row r only
            const bool ok = pc[sp] >= 0;
            const int j = ok ? a.pair_j[rowbase + r] : 0;
            float4 gv[4], xv[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int c0 = (l8 + 8 * it) * 4;
                gv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                xv[it] = gv[it];
                if (ok) {
                    if (a.g_new) gv[it] = ld4(a.g_new + (rowbase + r) * SCANN_D + c0);
                    xv[it] = ld4(a.x + (size_t)j * SCANN_D + c0);
                }
            }
            if (!a.g_new) {
                // SCANN without geometry update: g' = swish(rbf(d) @ Wf + bf) * w.  The 20 Gaussians of the
                // row's distance are computed by its 8 lanes (3 each) and broadcast; Wf is read through L1
                float d = 0.f, wgt = 0.f;
                if (ok) { d = a.pair_d[rowbase + r]; wgt = a.pair_w[rowbase + r]; }
                float rb[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int k = l8 + 8 * q;
                    const float df = d - (k < SCANN_RBF ? __ldg(a.centers + k) : 0.f);
                    rb[q] = expf(-(df * df) / 0.25f);
                }
                float4 acc[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) acc[it] = ldg4(a.bf + (l8 + 8 * it) * 4);
#pragma unroll
                for (int k = 0; k < SCANN_RBF; ++k) {
                    const float rk = __shfl_sync(0xffffffffu, rb[k >> 3], (lane & 24) | (k & 7));
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const float4 wv = ldg4(a.Wf + k * SCANN_D + (l8 + 8 * it) * 4);
                        acc[it].x = fmaf(rk, wv.x, acc[it].x); acc[it].y = fmaf(rk, wv.y, acc[it].y);
                        acc[it].z = fmaf(rk, wv.z, acc[it].z); acc[it].w = fmaf(rk, wv.w, acc[it].w);
                    }
                }
#pragma unroll
                for (int it = 0; it < 4; ++it)
                    gv[it] = ok ? make_float4(swish_f(acc[it].x) * wgt, swish_f(acc[it].y) * wgt,
                                              swish_f(acc[it].z) * wgt, swish_f(acc[it].w) * wgt)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.g_save)
#pragma unroll
                    for (int it = 0; it < 4; ++it) st4(a.g_save + (rowbase + r) * SCANN_D + (l8 + 8 * it) * 4, gv[it]);
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                float4 h, l;
                tf32_split(gv[it].x * xv[it].x, h.x, l.x); tf32_split(gv[it].y * xv[it].y, h.y, l.y);
                tf32_split(gv[it].z * xv[it].z, h.z, l.z); tf32_split(gv[it].w * xv[it].w, h.w, l.w);
                const uint32_t off = tc_off4(r, l8 + 8 * it);
                *reinterpret_cast<float4*>(sHi + off) = h;
                *reinterpret_cast<float4*>(sLo + off) = l;
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(grp, GT);
        if (wg == 0 && tc_elect_one()) {
            tc_fence_after();
            issue_3xtf32(t_whi, t_wlo, smem_u32(sHi), smem_u32(sLo), t_dm, t_dc, bar, a.mma_rows);
        }
        // ---- prefetch the queries of this warp's rows while the tensor core works
        float4 qv[2][4];
#pragma unroll
        for (int sp = 0; sp < 2; ++sp)
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                qv[sp][it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pc[sp] >= 0)
                    qv[sp][it] = ld4(a.proj + (size_t)pc[sp] * 3 * SCANN_D + 2 * SCANN_D + (l8 + 8 * it) * 4);
            }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        if (t + t_step >= nt) pdl_trigger();             // last tile of this group, only its epilogue is left
        tmem_to_rows(t_dm, t_dc, sS, a.bk, wg, lane, a.mma_rows);            // keys k = a @ Wk + bk
        tc_fence_before();
        group_sync(grp, GT);
        // ---- scores e[r][h] = 0.25 <q_h, k_h>  (head h = 16 columns = chunks 4h..4h+3 = 4 adjacent lanes)
#pragma unroll
        for (int sp = 0; sp < 2; ++sp) {
            if (__all_sync(0xffffffffu, pc[sp] < 0)) continue;
            const int r = sp * (TR / 2) + wg * 4 + rsub;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const float4 kv = *reinterpret_cast<const float4*>(sS + tc_off4(r, l8 + 8 * it));
                const float4 q = qv[sp][it];
                const float e = quad_sum(kv.x * q.x + kv.y * q.y + kv.z * q.z + kv.w * q.w) * 0.25f;
                if ((l8 & 3) == 0) Es[r * 8 + 2 * it + (l8 >> 2)] = e;
                if (a.k_out)
                    st4(a.k_out + (rowbase + r) * SCANN_D + (l8 + 8 * it) * 4,
                        pc[sp] >= 0 ? kv : make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
        group_sync(grp, GT);
        // ---- per atom: softmax over its rows, context, residual q, LayerNorm (one warp per atom)
        const int a0 = a.tile_a0[t], a1 = a.tile_a1[t];
        for (int atom = a0 + wg; atom < a1; atom += WG) {
            const int n = a.cnt[atom];
            if (n == 0) continue;
            const int r0 = a.rowptr[atom] - (int)rowbase;
            const int h = lane >> 2;
            float4 q = ld4(a.proj + (size_t)atom * 3 * SCANN_D + 2 * SCANN_D + lane * 4);
            float m = -INFINITY;
            for (int r = 0; r < n; ++r) m = fmaxf(m, Es[(r0 + r) * 8 + h]);
            float s = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            for (int r = 0; r < n; ++r) {
                float p = __expf(Es[(r0 + r) * 8 + h] - m);
                float4 kv = *reinterpret_cast<const float4*>(sS + tc_off4(r0 + r, lane));
                s += p;                                             // the softmax is normalised before the dropout
                p *= drop_mult(a.drop, a.drop_site, (uint32_t)(rowbase + r0 + r) * 8u + h);
                c0 = fmaf(p, kv.x, c0); c1 = fmaf(p, kv.y, c1); c2 = fmaf(p, kv.z, c2); c3 = fmaf(p, kv.w, c3);
            }
            const float is = 1.0f / s;
            if (a.attn && (lane & 3) == 0)
                for (int r = 0; r < n; ++r)
                    a.attn[(rowbase + r0 + r) * 8 + h] = __expf(Es[(r0 + r) * 8 + h] - m) * is *
                                                       drop_mult(a.drop, a.drop_site, (uint32_t)(rowbase + r0 + r) * 8u + h);
            c0 = c0 * is + q.x; c1 = c1 * is + q.y; c2 = c2 * is + q.z; c3 = c3 * is + q.w;
            if (a.ctx_pre) st4(a.ctx_pre + (size_t)atom * SCANN_D + lane * 4, make_float4(c0, c1, c2, c3));
            float mean = warp_sum(c0 + c1 + c2 + c3) * (1.0f / SCANN_D);
            c0 -= mean; c1 -= mean; c2 -= mean; c3 -= mean;
            float inv = rsqrtf(warp_sum(c0 * c0 + c1 * c1 + c2 * c2 + c3 * c3) * (1.0f / SCANN_D) + SCANN_LN_EPS);
            st4(a.out + (size_t)atom * SCANN_D + lane * 4,
                make_float4(c0 * inv * gam.x + bet.x, c1 * inv * gam.y + bet.y, c2 * inv * gam.z + bet.z,
                            c3 * inv * gam.w + bet.w));
        }
        group_sync(grp, GT);
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// Four warp groups per CTA ("tc4" kernels)
// =============================================================================================
// A wave-balanced plan fills at most 48 of the 64 rows of a tile slot (QM9 / 128 structures: 592 tiles of ~40
// rows).  With two groups per CTA those tiles take two rounds; the per-tile cost is mostly latency (global loads,
// barriers, the MMA pipeline, TMEM read-back), so the tc4 kernels run FOUR groups of 4 warps per CTA, each on its
// own tile: one round instead of two and twice as many independent instruction streams per SM.  Shared memory:
// a group owns TWO 48-row images instead of three 64-row ones -- the transposed accumulator S is written over
// the hi image once the MMAs have consumed it (and, where the epilogue needs the operand itself, hi + lo = the
// exact fp32 value is written over the lo image in the same pass).  Tensor memory: 256 weight columns + 4 x 48
// accumulator columns.  Rows per warp: 12 = 3 row-group steps of 4 rows (8 lanes per row).
#define L4_WG 4
#define L4_GT 128
#define L4_ROWS 48
#define L4_IMG (6u * TC_RG_STRIDE)
#define L4_SLOT 64                          // rows per tile slot of the plan (tile_stride)

// accumulator (lane = feature n, column = row r) -> S[r][n] (+ bias) written over the hi image; with KEEP the lo
// image receives hi + lo (= the staged fp32 operand, exactly).  Warp wg of its group owns lane quarter wg.
template <bool KEEP>
__device__ __forceinline__ void tmem_to_rows4(uint32_t t_dm, uint8_t* sHi, uint8_t* sLo, const float* __restrict__ bias,
                                              int wg, int lane, int nrows) {
    const int n = wg * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wg * 32) << 16;
    const float b = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        if (h * 16 >= nrows) break;
        float m[16];
        tmem_ld16(t_dm + lane_base + h * 16, m);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const uint32_t off = tc_off(h * 16 + q, n);
            if (KEEP) {
                const float x = *reinterpret_cast<const float*>(sHi + off) + *reinterpret_cast<const float*>(sLo + off);
                *reinterpret_cast<float*>(sLo + off) = x;
            }
            *reinterpret_cast<float*>(sHi + off) = m[q] + b;
        }
    }
}

__global__ void __launch_bounds__(LTC_THREADS, 1) la_geom_fwd_tc4_kernel(const LaGeomArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 2, wg = warp & 3, gtid = tid & (L4_GT - 1);
    uint8_t* sHi = smem + (size_t)grp * 2 * L4_IMG;      // hi image, later S
    uint8_t* sLo = sHi + L4_IMG;                         // lo image, later g
    uint64_t* bar = &bars[grp];
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid < 4) mbar_init(&bars[tid], 1);
    if (tid == 0) mbar_fence_init();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t t_whi = tmem, t_wlo = tmem + 128, t_dm = tmem + 256 + grp * L4_ROWS;
    DBG_CLK(0);
    weightT_to_tmem(a.W2, t_whi, t_wlo, warp, lane);
    pdl_wait();
    const int nt = *a.ntiles;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    DBG_CLK(1);
    uint32_t phase = 0;
    // tile t -> group (t / grid) % 4 of CTA t % grid: a tile count below 4 x grid still spreads over all SMs
    const int t_first = blockIdx.x + grp * gridDim.x, t_step = gridDim.x * 4;
    const int l8 = lane & 7, rsub = lane >> 3;
    float4 gm[4], bt[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) { gm[it] = ldg4(a.gamma_g + (l8 + 8 * it) * 4); bt[it] = ldg4(a.beta_g + (l8 + 8 * it) * 4); }
    for (int t = t_first; t < nt; t += t_step) {
        const size_t rowbase = (size_t)t * L4_SLOT;
        // start-up skew: the groups of a CTA (and, after a kernel boundary, all CTAs of the grid) would otherwise
        // walk load -> MMA -> read-back -> epilogue in lock-step and queue up on one resource after the other;
        // group g starts loading when group g-1 has its tile, and the groups stay out of phase from there
        if (t == t_first && grp > 0 && a.skew) bar_sync(4 + grp, 2 * L4_GT);
        // ---- stage the geometry tile (48 rows x 32 chunks over 128 threads)
        {
            float4 v[12];
#pragma unroll
            for (int it = 0; it < 12; ++it) {
                const int i = gtid + it * L4_GT;
                v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((i >> 5) < a.mma_rows) v[it] = ld4(a.g_in + (rowbase + (i >> 5)) * SCANN_D + (i & 31) * 4);
            }
#pragma unroll
            for (int it = 0; it < 12; ++it) {
                const int i = gtid + it * L4_GT;
                float4 h, l;
                tf32_split(v[it].x, h.x, l.x); tf32_split(v[it].y, h.y, l.y);
                tf32_split(v[it].z, h.z, l.z); tf32_split(v[it].w, h.w, l.w);
                const uint32_t off = tc_off4(i >> 5, i & 31);
                *reinterpret_cast<float4*>(sHi + off) = h;
                *reinterpret_cast<float4*>(sLo + off) = l;
            }
        }
        if (t == t_first && grp < 3 && a.skew) bar_arrive(5 + grp, 2 * L4_GT);
        fence_async_smem();
        tc_fence_before();
        group_sync(grp, L4_GT);
        if (t == t_first) DBG_CLK(2);
        if (wg == 0 && tc_elect_one()) {
            tc_fence_after();
            issue_3xtf32(t_whi, t_wlo, smem_u32(sHi), smem_u32(sLo), t_dm, 0, bar, a.mma_rows);
        }
        if (t == t_first) DBG_CLK(3);
        // ---- while the tensor core works: indices of this warp's 12 rows, gathered projections of the first 8
        int pc[3], pj[3];
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) {
            const int r = sp * 16 + wg * 4 + rsub;
            pc[sp] = a.pair_c[rowbase + r];
            pj[sp] = pc[sp] >= 0 ? a.pair_j[rowbase + r] : 0;
        }
        auto gather = [&](int sp, float4 (&p)[4]) {
            if (__all_sync(0xffffffffu, pc[sp] < 0)) return;
            const int c = pc[sp] >= 0 ? pc[sp] : 0;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int c0 = (l8 + 8 * it) * 4;
                p[it] = f4add(ld4(a.proj + (size_t)c * 3 * SCANN_D + c0),
                              ld4(a.proj + (size_t)pj[sp] * 3 * SCANN_D + SCANN_D + c0));
            }
        };
        float4 pa[4], pb[4];
        gather(0, pa);
        gather(1, pb);
        if (t == t_first) DBG_CLK(4);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        if (t + t_step >= nt) pdl_trigger();
        if (t == t_first) DBG_CLK(5);
        tmem_to_rows4<true>(t_dm, sHi, sLo, nullptr, wg, lane, a.mma_rows);
        tc_fence_before();
        group_sync(grp, L4_GT);
        if (t == t_first) DBG_CLK(6);
        // ---- row-wise epilogue: pre -> swish -> + g -> LayerNorm -> g'
        auto epi = [&](int sp, const float4 (&p13)[4]) {
            if (__all_sync(0xffffffffu, pc[sp] < 0)) return;
            const int r = sp * 16 + wg * 4 + rsub;
            float z[4][4], pre[4][4];
            float s1 = 0.f;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const uint32_t off = tc_off4(r, l8 + 8 * it);
                const float4 acc = *reinterpret_cast<const float4*>(sHi + off);
                const float4 g = *reinterpret_cast<const float4*>(sLo + off);
                pre[it][0] = acc.x + p13[it].x; pre[it][1] = acc.y + p13[it].y;
                pre[it][2] = acc.z + p13[it].z; pre[it][3] = acc.w + p13[it].w;
                z[it][0] = swish_fast(pre[it][0]) + g.x; z[it][1] = swish_fast(pre[it][1]) + g.y;
                z[it][2] = swish_fast(pre[it][2]) + g.z; z[it][3] = swish_fast(pre[it][3]) + g.w;
                s1 += z[it][0] + z[it][1] + z[it][2] + z[it][3];
            }
            const float sh = __shfl_sync(0xffffffffu, s1, lane & 24) * (1.0f / 16.0f);
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int it = 0; it < 4; ++it)
#pragma unroll
                for (int q = 0; q < 4; ++q) { z[it][q] -= sh; m1 += z[it][q]; m2 = fmaf(z[it][q], z[it][q], m2); }
            oct_sum2(m1, m2);
            m1 *= (1.0f / SCANN_D);
            const float inv = rsqrtf(fmaxf(m2 * (1.0f / SCANN_D) - m1 * m1, 0.f) + SCANN_LN_EPS);
            const bool ok = pc[sp] >= 0;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int c0 = (l8 + 8 * it) * 4;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f), po = o;
                if (ok) {
                    o = make_float4((z[it][0] - m1) * inv * gm[it].x + bt[it].x, (z[it][1] - m1) * inv * gm[it].y + bt[it].y,
                                    (z[it][2] - m1) * inv * gm[it].z + bt[it].z, (z[it][3] - m1) * inv * gm[it].w + bt[it].w);
                    po = make_float4(pre[it][0], pre[it][1], pre[it][2], pre[it][3]);
                }
                st4(a.g_out + (rowbase + r) * SCANN_D + c0, o);
                if (a.pre_out) st4(a.pre_out + (rowbase + r) * SCANN_D + c0, po);
            }
        };
        epi(0, pa);
        gather(2, pa);                   // in flight behind the second step
        epi(1, pb);
        epi(2, pa);
        group_sync(grp, L4_GT);          // the images are rewritten by the next tile
        if (t == t_first) DBG_CLK(7);
    }
    DBG_CLK(8);
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

#define LA4_GEOM_SMEM (4 * 2 * L4_IMG)
#define LA4_ATTN_SMEM (4 * 2 * L4_IMG + 4 * L4_ROWS * 8 * sizeof(float))

#define LA_GEOM_SMEM (3 * TC_TILE_BYTES)
#define LA_ATTN_SMEM (3 * TC_TILE_BYTES + SCANN_TILE * 8 * sizeof(float))

static int la_fwd_configure() {
    static bool configured = false;
    if (configured) return 0;
    cudaError_t e = cudaFuncSetAttribute(la_geom_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_GEOM_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(la_geom_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_GEOM_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(la_attn_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_ATTN_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(la_attn_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_ATTN_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(la_geom_fwd_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA4_GEOM_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(la_geom_fwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_GEOM_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(la_attn_fwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_ATTN_SMEM);
    if (e != cudaSuccess) { scann_set_error("la_forward_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return 1; }
    configured = true;
    return 0;
}

// Tensor-core forward of LocalAttention.call (attention.py:118-216); same data contract as
// scann_la_forward plus the optional training saves pre_out / k_out ([rows,128] each).
// tile_stride = rows per tile slot of the pair plan (scann_plan_build): 128 (one tile stream per CTA) or
// 64 (two warp groups per CTA, each streaming 64-row tiles).
extern "C" int scann_la_forward_tc(int grid, int tile_stride, int mma_rows, const int32_t* ntiles, const int32_t* tile_a0,
                                   const int32_t* tile_a1, const int32_t* cnt, const int32_t* rowptr,
                                   const int32_t* pair_c, const int32_t* pair_j, const float* x, const float* proj,
                                   const float* g_in, const float* W2, const float* Wk, const float* bk,
                                   const float* gamma_g, const float* beta_g, const float* gamma, const float* beta,
                                   float* g_out, float* ctx_pre, float* out, float* attn, float* pre_out, float* k_out,
                                   const void* attn_drop, int drop_site, void* stream) {
    if (tile_stride != 32 && tile_stride != 64 && tile_stride != 128) { scann_set_error("la_forward_tc: tile_stride must be 32, 64 or 128"); return 1; }
    if (mma_rows < 16 || mma_rows > tile_stride || mma_rows % 16) { scann_set_error("la_forward_tc: mma_rows must be a multiple of 16 in 16..tile_stride"); return 1; }
    if (la_fwd_configure()) return 1;
    if (grid <= 0) return 0;
    LaGeomArgs ga{ntiles, pair_c, pair_j, proj, g_in, W2, gamma_g, beta_g, g_out, pre_out, mma_rows,
                  (scann_la_tc4_mask() >> 4) & 1};
    LaAttnArgs aa{ntiles, tile_a0, tile_a1, cnt, rowptr, pair_c, pair_j, x, proj, g_out, Wk, bk, gamma, beta,
                  ctx_pre, out, attn, k_out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, mma_rows,
                  (const ScannDropCtl*)attn_drop, drop_site};
    if (tile_stride == 32) {
        // 32-row tile slots (the layout of the pipelined kernels, la_pipe.cu): four independent 4-warp groups
        scann_launch(la_geom_fwd_tc_kernel<4>, dim3(grid), dim3(LTC_THREADS), LA_GEOM_SMEM, stream, ga);
        scann_launch(la_attn_fwd_tc_kernel<4>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_SMEM, stream, aa);
    } else if (tile_stride == 64) {
        // four warp groups per CTA when the plan's tiles hold at most 48 rows (see the tc4 section above)
        const int m4 = mma_rows <= L4_ROWS ? scann_la_tc4_mask() : 0;
        if (m4 & 1) scann_launch(la_geom_fwd_tc4_kernel, dim3(grid), dim3(LTC_THREADS), LA4_GEOM_SMEM, stream, ga);
        else scann_launch(la_geom_fwd_tc_kernel<2>, dim3(grid), dim3(LTC_THREADS), LA_GEOM_SMEM, stream, ga);
        scann_launch(la_attn_fwd_tc_kernel<2>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_SMEM, stream, aa);
    } else {
        scann_launch(la_geom_fwd_tc_kernel<1>, dim3(grid), dim3(LTC_THREADS), LA_GEOM_SMEM, stream, ga);
        scann_launch(la_attn_fwd_tc_kernel<1>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_SMEM, stream, aa);
    }
    return scann_check_launch("scann_la_forward_tc");
}

// LocalAttention.call with g_update = False (attention.py:155-216): neighbor_geometry' =
// swish(rbf(distance) @ Wf + bf) * weight is recomputed per layer from the 8 bytes/pair of raw geometry;
// proj needs only its query block (columns 256..383).  g_save / k_out ([rows,128], nullable): g' and the keys,
// saved for scann_la_backward_noupdate_tc.
extern "C" int scann_la_forward_noupdate_tc(int grid, int tile_stride, int mma_rows, const int32_t* ntiles, const int32_t* tile_a0,
                                            const int32_t* tile_a1, const int32_t* cnt, const int32_t* rowptr,
                                            const int32_t* pair_c, const int32_t* pair_j, const float* x,
                                            const float* proj, const float* pair_d, const float* pair_w,
                                            const float* centers, const float* Wf, const float* bf, const float* Wk,
                                            const float* bk, const float* gamma, const float* beta, float* ctx_pre,
                                            float* out, float* attn, float* g_save, float* k_out, const void* attn_drop,
                                            int drop_site, void* stream) {
    if (tile_stride != 32 && tile_stride != 64 && tile_stride != 128) { scann_set_error("la_forward_noupdate_tc: tile_stride must be 32, 64 or 128"); return 1; }
    if (mma_rows < 16 || mma_rows > tile_stride || mma_rows % 16) { scann_set_error("la_forward_noupdate_tc: bad mma_rows"); return 1; }
    if (la_fwd_configure()) return 1;
    if (grid <= 0) return 0;
    LaAttnArgs aa{ntiles, tile_a0, tile_a1, cnt, rowptr, pair_c, pair_j, x, proj, nullptr, Wk, bk, gamma, beta,
                  ctx_pre, out, attn, k_out, pair_d, pair_w, centers, Wf, bf, g_save, mma_rows,
                  (const ScannDropCtl*)attn_drop, drop_site};
    if (tile_stride == 32) scann_launch(la_attn_fwd_tc_kernel<4>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_SMEM, stream, aa);
    else if (tile_stride == 64) scann_launch(la_attn_fwd_tc_kernel<2>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_SMEM, stream, aa);
    else scann_launch(la_attn_fwd_tc_kernel<1>, dim3(grid), dim3(LTC_THREADS), LA_ATTN_SMEM, stream, aa);
    return scann_check_launch("scann_la_forward_noupdate_tc");
}

#ifdef SCANN_DEV_PROBES
extern "C" int scann_debug_clocks(long long* host_out32) {
    cudaError_t e = cudaMemcpyFromSymbol(host_out32, g_dbg_clk, sizeof(long long) * 32);
    if (e != cudaSuccess) { scann_set_error("debug_clocks: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}
#endif
