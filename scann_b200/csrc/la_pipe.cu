// Local attention forward (g_update = True) as two WARP-SPECIALISED, TMA-fed pipelines per layer:
//
//   la_geom_fwd_pipe : g' = LN_g(swish(xw1[c] + g @ W2 + xw3[j]) + g)          (attention.py:141-153)
//   la_attn_fwd_pipe : k = (x[j] * g') @ Wk + bk ; softmax over the atom's pairs ; out = LN(sum p k + q)
//                                                                            (attention.py:157-214)
//
// Same arithmetic as la_tc.cu (3xTF32 on tcgen05, the weight block stationary in tensor memory as the M x K
// operand, pair tiles as the N x K operand, the accumulator transposed through shared memory for the row-wise
// epilogue) -- what changes is how a tile moves through the SM:
//
//   producer warp   : TMA tensor loads (cp.async.bulk.tensor, 128-byte swizzle) of the tile's geometry rows and
//                     bulk copies of its index rows into stage s of a PF_NS-deep ring; the attention kernel's
//                     producer also gathers the neighbour features x[j] row by row with 16-byte cp.async straight
//                     into the tile image.  It runs PF_NS - PF_NG tiles ahead of the consumers, so no compute warp
//                     ever waits for HBM.
//   MMA warp        : one elected lane issues the 48 tcgen05.mma of a tile as soon as its operand images are
//                     ready, into the stage's own accumulator columns, and commits to the stage's mbarrier.
//   consumer groups : PF_NG groups of four warps; group g owns tiles g, g + PF_NG, ...: split pass (lo = x -
//                     trunc(x); the TMA-landed fp32 tile itself is the "hi" operand because the tensor core
//                     truncates), accumulator read-back, row-wise epilogue IN PLACE in the stage's images, and a
//                     TMA store of the finished tile (g', and pre / k when training saves them).
//
// While group g runs the epilogue of tile t, the tensor core works on tile t+1 of group g+1 and the TMA unit
// loads tiles t+2..; the phases that used to run in lock-step on every SM (profiles/r01i_tc_summary.md) overlap.
// Tiles are PT = 32 pair rows (plan tile_stride 32, at most 32 neighbours per atom); a stage is two 16 KB tile
// images, so six tiles are in flight per SM.  Every mbarrier wait is bounded (pipe_common.cuh).
#include <string.h>

#include "pipe_common.cuh"

#define PF_NS 6                                // ring stages
#define PF_NG 3                                // consumer groups
#define PF_CW (4 * PF_NG)                      // consumer warps
#define PF_THREADS ((PF_CW + 2) * 32)          // + producer warp + MMA warp
#define PF_STAGE (2u * PT_IMG)
// dynamic shared memory: stages | idx[NS][64] | Es[NS][PT*8] | barriers | flags   (+ 1024 for alignment)
#define PF_OFF_IDX (PF_NS * PF_STAGE)
#define PF_OFF_ES (PF_OFF_IDX + PF_NS * 64u * 4u)
#define PF_OFF_BAR (PF_OFF_ES + PF_NS * PT * 8u * 4u)
#define PF_OFF_FLAGS (PF_OFF_BAR + 4u * PF_NS * 8u)
#define PF_SMEM (PF_OFF_FLAGS + 16u + 1024u)

__device__ __forceinline__ float4 pf4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 lds4(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(uint8_t* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

struct PipeGeomArgs {
    CUtensorMap tm_gin, tm_gout, tm_pre;
    const int32_t* ntiles; const int32_t* pair_c; const int32_t* pair_j;
    const float* proj;       // [R,384] = [x@W1+bf | x@W3 | x@Wq+bq]
    const float* W2;         // rows 128..255 of filter_geo/kernel, [k][n]
    const float* gamma_g; const float* beta_g;
    int has_pre;             // training: save the filter_geo pre-activation through tm_pre
    int32_t* status;
};

struct PipeCtx {
    uint8_t* stages; int32_t* idx; float* es; uint64_t *full, *empty, *ready, *accf; uint32_t* tmem_slot; volatile int* dead;
};
__device__ __forceinline__ PipeCtx pipe_carve(uint8_t* smem_raw) {
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    PipeCtx c;
    c.stages = smem;
    c.idx = reinterpret_cast<int32_t*>(smem + PF_OFF_IDX);
    c.es = reinterpret_cast<float*>(smem + PF_OFF_ES);
    c.full = reinterpret_cast<uint64_t*>(smem + PF_OFF_BAR);
    c.empty = c.full + PF_NS; c.ready = c.empty + PF_NS; c.accf = c.ready + PF_NS;
    c.tmem_slot = reinterpret_cast<uint32_t*>(smem + PF_OFF_FLAGS);
    c.dead = reinterpret_cast<volatile int*>(smem + PF_OFF_FLAGS + 4);
    return c;
}

// =============================================================================================
// Geometry update
// =============================================================================================
__global__ void __launch_bounds__(PF_THREADS, 1) la_geom_fwd_pipe_kernel(const __grid_constant__ PipeGeomArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const PipeCtx c = pipe_carve(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == PF_CW + 1) tmem_alloc(c.tmem_slot, 512);
    if (tid == 0) {
        for (int s = 0; s < PF_NS; ++s) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], 1); mbar_init(&c.ready[s], 1); mbar_init(&c.accf[s], 1); }
        *c.dead = 0;
        mbar_fence_init();
    }
    if (warp == PF_CW && lane == 0) {
        tma_prefetch_desc(&a.tm_gin); tma_prefetch_desc(&a.tm_gout);
        if (a.has_pre) tma_prefetch_desc(&a.tm_pre);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *c.tmem_slot;
    const uint32_t t_wraw = tmem, t_wlo = tmem + 128, t_acc0 = tmem + 256;
    if (warp < PF_CW) pipe_weight_to_tmem(a.W2, t_wraw, t_wlo, warp, lane, PF_CW);   // parameters only: before the PDL wait
    pdl_wait();
    const int nt = *a.ntiles;
    tc_fence_before();
    __syncthreads();                                     // the whole weight is in tensor memory
    tc_fence_after();

    if (warp == PF_CW) {
        // ================= producer =================
        if (lane == 0) {
            int i = 0;
            for (int t = blockIdx.x; t < nt; t += gridDim.x, ++i) {
                const int s = i % PF_NS;
                const uint32_t ph = (uint32_t)(i / PF_NS) & 1u;
                pipe_wait(&c.empty[s], ph ^ 1u, c.dead, a.status, 1, t, s);
                uint8_t* A = c.stages + (size_t)s * PF_STAGE;
                mbar_expect_tx(&c.full[s], PT_IMG + 256u);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(A + kb * PT_CB, &a.tm_gin, kb * 32, t * PT, &c.full[s]);
                bulk_load(c.idx + s * 64, a.pair_c + (size_t)t * PT, 128u, &c.full[s]);
                bulk_load(c.idx + s * 64 + 32, a.pair_j + (size_t)t * PT, 128u, &c.full[s]);
            }
        }
        __syncwarp();
    } else if (warp == PF_CW + 1) {
        // ================= MMA issue =================
        int i = 0;
        for (int t = blockIdx.x; t < nt; t += gridDim.x, ++i) {
            const int s = i % PF_NS;
            const uint32_t ph = (uint32_t)(i / PF_NS) & 1u;
            pipe_wait(&c.ready[s], ph, c.dead, a.status, 2, t, s);
            tc_fence_after();
            __syncwarp();                                       // the lanes leave the wait loop at different times
            if (tc_elect_one()) {
                const uint32_t A = smem_u32(c.stages + (size_t)s * PF_STAGE);
                pipe_issue_3xtf32(t_wraw, t_wlo, A, A + PT_IMG, t_acc0 + s * PT, &c.accf[s]);
            }
            __syncwarp();
        }
    } else {
        // ================= consumers =================
        const int grp = warp >> 2, wg = warp & 3, gtid = tid - grp * 128;
        const int l8 = lane & 7, rsub = lane >> 3;
        float4 gm[4], bt[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) { gm[it] = ldg4(a.gamma_g + (l8 + 8 * it) * 4); bt[it] = ldg4(a.beta_g + (l8 + 8 * it) * 4); }
        int i = grp;
        for (int t = blockIdx.x + grp * gridDim.x; t < nt; t += PF_NG * gridDim.x, i += PF_NG) {
            const int s = i % PF_NS;
            const uint32_t ph = (uint32_t)(i / PF_NS) & 1u;
            uint8_t* A = c.stages + (size_t)s * PF_STAGE;       // g (fp32, as landed) -> g'
            uint8_t* Bm = A + PT_IMG;                           // lo image -> transposed accumulator -> pre
            const int32_t* sidx = c.idx + s * 64;
            pipe_wait(&c.full[s], ph, c.dead, a.status, 3, t, s);
            // ---- split pass: the landed tile is the hi operand; lo = g - trunc(g)
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const uint32_t off = pt_off4(wg + 4 * it, lane);
                const float4 v = lds4(A + off);
                sts4(Bm + off, make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w)));
            }
            fence_async_smem();
            group_sync(grp, 128);
            if (gtid == 0) mbar_arrive(&c.ready[s]);
            // ---- while the tensor core works: gathered per-atom projections of this warp's rows
            int pc[2];
            float4 p13[2][4];
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
                const int r = sp * 16 + wg * 4 + rsub;
                pc[sp] = sidx[r];
                const int j = pc[sp] >= 0 ? sidx[32 + r] : 0;
                const int cc = pc[sp] >= 0 ? pc[sp] : 0;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int c0 = (l8 + 8 * it) * 4;
                    p13[sp][it] = pf4add(ld4(a.proj + (size_t)cc * 3 * SCANN_D + c0),
                                         ld4(a.proj + (size_t)j * 3 * SCANN_D + SCANN_D + c0));
                }
            }
            pipe_wait(&c.accf[s], ph, c.dead, a.status, 4, t, s);
            tc_fence_after();
            pipe_acc_to_image(t_acc0 + s * PT, Bm, nullptr, wg, lane);
            tc_fence_before();
            group_sync(grp, 128);
            // ---- row-wise epilogue: pre -> swish -> + g -> LayerNorm -> g' (in place)
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
                if (__all_sync(0xffffffffu, pc[sp] < 0)) continue;
                const int r = sp * 16 + wg * 4 + rsub;
                float z[4][4], pre[4][4];
                float s1 = 0.f;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const uint32_t off = pt_off4(r, l8 + 8 * it);
                    const float4 acc = lds4(Bm + off), g = lds4(A + off);
                    pre[it][0] = acc.x + p13[sp][it].x; pre[it][1] = acc.y + p13[sp][it].y;
                    pre[it][2] = acc.z + p13[sp][it].z; pre[it][3] = acc.w + p13[sp][it].w;
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
                    const uint32_t off = pt_off4(r, l8 + 8 * it);
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f), po = o;
                    if (ok) {
                        o = make_float4((z[it][0] - m1) * inv * gm[it].x + bt[it].x, (z[it][1] - m1) * inv * gm[it].y + bt[it].y,
                                        (z[it][2] - m1) * inv * gm[it].z + bt[it].z, (z[it][3] - m1) * inv * gm[it].w + bt[it].w);
                        po = make_float4(pre[it][0], pre[it][1], pre[it][2], pre[it][3]);
                    }
                    sts4(A + off, o);
                    sts4(Bm + off, po);
                }
            }
            fence_async_smem();
            group_sync(grp, 128);
            if (gtid == 0) {
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) tma_store_2d(&a.tm_gout, A + kb * PT_CB, kb * 32, t * PT);
                if (a.has_pre) {
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) tma_store_2d(&a.tm_pre, Bm + kb * PT_CB, kb * 32, t * PT);
                }
                tma_commit();
                tma_wait_read0();
                mbar_arrive(&c.empty[s]);
            }
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == PF_CW + 1) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// Attention
// =============================================================================================
struct PipeAttnArgs {
    CUtensorMap tm_g, tm_k;
    const int32_t* ntiles; const int32_t* pair_c; const int32_t* pair_j;
    const float* x;          // [R,128]
    const float* proj;       // [R,384]; q = columns 256..383
    const float* Wk; const float* bk; const float* gamma; const float* beta;
    float* ctx_pre;          // [R,128] nullable
    float* out;              // [R,128]
    float* attn;             // [rows,8] nullable
    int has_k;               // training: save the keys through tm_k
    const ScannDropCtl* drop; int drop_site;     // Dropout(0.05) on the attention probabilities (attention.py:191-192)
    int32_t* status;
};

__global__ void __launch_bounds__(PF_THREADS, 1) la_attn_fwd_pipe_kernel(const __grid_constant__ PipeAttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const PipeCtx c = pipe_carve(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == PF_CW + 1) tmem_alloc(c.tmem_slot, 512);
    if (tid == 0) {
        // full: the expect_tx arrive of lane 0 + the 32 cp.async arrives of the gathering lanes
        for (int s = 0; s < PF_NS; ++s) { mbar_init(&c.full[s], 33); mbar_init(&c.empty[s], 1); mbar_init(&c.ready[s], 1); mbar_init(&c.accf[s], 1); }
        *c.dead = 0;
        mbar_fence_init();
    }
    if (warp == PF_CW && lane == 0) {
        tma_prefetch_desc(&a.tm_g);
        if (a.has_k) tma_prefetch_desc(&a.tm_k);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *c.tmem_slot;
    const uint32_t t_wraw = tmem, t_wlo = tmem + 128, t_acc0 = tmem + 256;
    if (warp < PF_CW) pipe_weight_to_tmem(a.Wk, t_wraw, t_wlo, warp, lane, PF_CW);
    pdl_wait();
    const int nt = *a.ntiles;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == PF_CW) {
        // ================= producer: g' tile by TMA, x[j] rows by per-lane cp.async =================
        int i = 0;
        int jn = 0;
        if ((int)blockIdx.x < nt) {
            const int pcv = a.pair_c[(size_t)blockIdx.x * PT + lane];
            jn = pcv >= 0 ? a.pair_j[(size_t)blockIdx.x * PT + lane] : 0;
        }
        for (int t = blockIdx.x; t < nt; t += gridDim.x, ++i) {
            const int s = i % PF_NS;
            const uint32_t ph = (uint32_t)(i / PF_NS) & 1u;
            const int j = jn;
            const int tn = t + (int)gridDim.x;
            if (tn < nt) {                                      // next tile's neighbour rows: in flight behind this tile's issue
                const int pcv = a.pair_c[(size_t)tn * PT + lane];
                jn = pcv >= 0 ? a.pair_j[(size_t)tn * PT + lane] : 0;
            }
            pipe_wait(&c.empty[s], ph ^ 1u, c.dead, a.status, 11, t, s);
            __syncwarp();
            uint8_t* A = c.stages + (size_t)s * PF_STAGE;
            if (lane == 0) {
                mbar_expect_tx(&c.full[s], PT_IMG + 256u);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(A + kb * PT_CB, &a.tm_g, kb * 32, t * PT, &c.full[s]);
                bulk_load(c.idx + s * 64, a.pair_c + (size_t)t * PT, 128u, &c.full[s]);
                bulk_load(c.idx + s * 64 + 32, a.pair_j + (size_t)t * PT, 128u, &c.full[s]);
            }
            const uint32_t X = smem_u32(A + PT_IMG);
#pragma unroll 8
            for (int r = 0; r < PT; ++r) {
                const int jr = __shfl_sync(0xffffffffu, j, r);
                cp_async16(X + pt_off4(r, lane), a.x + (size_t)jr * SCANN_D + lane * 4);
            }
            cp_async_arrive(&c.full[s]);
        }
    } else if (warp == PF_CW + 1) {
        // ================= MMA issue =================
        int i = 0;
        for (int t = blockIdx.x; t < nt; t += gridDim.x, ++i) {
            const int s = i % PF_NS;
            const uint32_t ph = (uint32_t)(i / PF_NS) & 1u;
            pipe_wait(&c.ready[s], ph, c.dead, a.status, 12, t, s);
            tc_fence_after();
            __syncwarp();                                       // the lanes leave the wait loop at different times
            if (tc_elect_one()) {
                const uint32_t A = smem_u32(c.stages + (size_t)s * PF_STAGE);
                pipe_issue_3xtf32(t_wraw, t_wlo, A, A + PT_IMG, t_acc0 + s * PT, &c.accf[s]);
            }
            __syncwarp();
        }
    } else {
        // ================= consumers =================
        const int grp = warp >> 2, wg = warp & 3, gtid = tid - grp * 128;
        const int l8 = lane & 7, rsub = lane >> 3;
        const float4 gam = ldg4(a.gamma + lane * 4), bet = ldg4(a.beta + lane * 4);
        int i = grp;
        for (int t = blockIdx.x + grp * gridDim.x; t < nt; t += PF_NG * gridDim.x, i += PF_NG) {
            const int s = i % PF_NS;
            const uint32_t ph = (uint32_t)(i / PF_NS) & 1u;
            const size_t rowbase = (size_t)t * PT;
            uint8_t* A = c.stages + (size_t)s * PF_STAGE;       // g' -> a = x[j] * g' (hi operand) -> keys
            uint8_t* Bm = A + PT_IMG;                           // x[j] -> lo image
            const int32_t* sidx = c.idx + s * 64;
            float* Es = c.es + s * PT * 8;
            pipe_wait(&c.full[s], ph, c.dead, a.status, 13, t, s);
            int pc[2];
            float4 qv[2][4];
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
                pc[sp] = sidx[sp * 16 + wg * 4 + rsub];
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    qv[sp][it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pc[sp] >= 0) qv[sp][it] = ld4(a.proj + (size_t)pc[sp] * 3 * SCANN_D + 2 * SCANN_D + (l8 + 8 * it) * 4);
                }
            }
            // ---- a = x[j] * g' in place: the fp32 product is the hi operand, lo = a - trunc(a) over x[j]
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
                if (__all_sync(0xffffffffu, pc[sp] < 0)) continue;        // stale operand rows only feed their own columns
                const int r = sp * 16 + wg * 4 + rsub;
                const bool ok = pc[sp] >= 0;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const uint32_t off = pt_off4(r, l8 + 8 * it);
                    const float4 g = lds4(A + off), xv = lds4(Bm + off);
                    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) av = make_float4(g.x * xv.x, g.y * xv.y, g.z * xv.z, g.w * xv.w);
                    sts4(A + off, av);
                    sts4(Bm + off, make_float4(tf32_lo(av.x), tf32_lo(av.y), tf32_lo(av.z), tf32_lo(av.w)));
                }
            }
            fence_async_smem();
            group_sync(grp, 128);
            if (gtid == 0) mbar_arrive(&c.ready[s]);
            pipe_wait(&c.accf[s], ph, c.dead, a.status, 14, t, s);
            tc_fence_after();
            pipe_acc_to_image(t_acc0 + s * PT, A, a.bk, wg, lane);          // keys k = a @ Wk + bk
            tc_fence_before();
            group_sync(grp, 128);
            // ---- scores e[r][h] = 0.25 <q_h, k_h>  (head h = 16 columns = 4 adjacent lanes of the row)
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
                if (__all_sync(0xffffffffu, pc[sp] < 0)) continue;
                const int r = sp * 16 + wg * 4 + rsub;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const float4 kv = lds4(A + pt_off4(r, l8 + 8 * it));
                    const float4 q = qv[sp][it];
                    const float e = quad_sum(kv.x * q.x + kv.y * q.y + kv.z * q.z + kv.w * q.w) * 0.25f;
                    if ((l8 & 3) == 0) Es[r * 8 + 2 * it + (l8 >> 2)] = e;
                }
            }
            group_sync(grp, 128);
            // ---- per atom (one warp each): softmax over its rows, context, residual q, LayerNorm.  The atoms of the
            // tile are read off the centre indices: valid rows are a prefix, an atom's rows are contiguous
            {
                const int myc = sidx[lane];
                const int prevc = __shfl_up_sync(0xffffffffu, myc, 1);
                const uint32_t vmask = __ballot_sync(0xffffffffu, myc >= 0);
                const uint32_t hmask = __ballot_sync(0xffffffffu, myc >= 0 && (lane == 0 || myc != prevc));
                const int nvalid = __popc(vmask), natoms = __popc(hmask);
                uint32_t m = hmask;
                for (int k = 0; k < wg; ++k) m &= m - 1;
                for (int k = wg; k < natoms; k += 4) {
                    const int r0 = __ffs(m) - 1;
                    uint32_t mn = m;
                    mn &= mn - 1;
                    const int n = (mn ? __ffs(mn) - 1 : nvalid) - r0;
                    const int atom = sidx[r0];
                    const int h = lane >> 2;
                    const float4 q = ld4(a.proj + (size_t)atom * 3 * SCANN_D + 2 * SCANN_D + lane * 4);
                    float mx = -INFINITY;
                    for (int r = 0; r < n; ++r) mx = fmaxf(mx, Es[(r0 + r) * 8 + h]);
                    float sm = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
                    for (int r = 0; r < n; ++r) {
                        float p = __expf(Es[(r0 + r) * 8 + h] - mx);
                        const float4 kv = lds4(A + pt_off4(r0 + r, lane));
                        sm += p;                                        // the softmax is normalised before the dropout
                        p *= drop_mult(a.drop, a.drop_site, (uint32_t)(rowbase + r0 + r) * 8u + h);
                        c0 = fmaf(p, kv.x, c0); c1 = fmaf(p, kv.y, c1); c2 = fmaf(p, kv.z, c2); c3 = fmaf(p, kv.w, c3);
                    }
                    const float is = 1.0f / sm;
                    if (a.attn && (lane & 3) == 0)
                        for (int r = 0; r < n; ++r)
                            a.attn[(rowbase + r0 + r) * 8 + h] = __expf(Es[(r0 + r) * 8 + h] - mx) * is *
                                                               drop_mult(a.drop, a.drop_site, (uint32_t)(rowbase + r0 + r) * 8u + h);
                    c0 = c0 * is + q.x; c1 = c1 * is + q.y; c2 = c2 * is + q.z; c3 = c3 * is + q.w;
                    if (a.ctx_pre) st4(a.ctx_pre + (size_t)atom * SCANN_D + lane * 4, make_float4(c0, c1, c2, c3));
                    const float mean = warp_sum(c0 + c1 + c2 + c3) * (1.0f / SCANN_D);
                    c0 -= mean; c1 -= mean; c2 -= mean; c3 -= mean;
                    const float inv = rsqrtf(warp_sum(c0 * c0 + c1 * c1 + c2 * c2 + c3 * c3) * (1.0f / SCANN_D) + SCANN_LN_EPS);
                    st4(a.out + (size_t)atom * SCANN_D + lane * 4,
                        make_float4(c0 * inv * gam.x + bet.x, c1 * inv * gam.y + bet.y, c2 * inv * gam.z + bet.z,
                                    c3 * inv * gam.w + bet.w));
                    for (int k2 = 0; k2 < 4 && m; ++k2) m &= m - 1;     // this warp's next atom
                }
            }
            fence_async_smem();
            group_sync(grp, 128);
            if (gtid == 0) {
                if (a.has_k) {
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) tma_store_2d(&a.tm_k, A + kb * PT_CB, kb * 32, t * PT);
                    tma_commit();
                    tma_wait_read0();
                }
                mbar_arrive(&c.empty[s]);
            }
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == PF_CW + 1) tmem_dealloc(tmem, 512);
}

// ---- host ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int pipe_encode_tmap(CUtensorMap* tm, const float* base, long long rows) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
            scann_set_error("cuTensorMapEncodeTiled is not available: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return 1;
        }
        fn = (PFN_encodeTiled)p;
    }
    if (!base || rows < PT || ((uintptr_t)base & 127)) { scann_set_error("tensor map: base must be 128-byte aligned with >= 32 rows"); return 1; }
    const cuuint64_t dims[2] = {(cuuint64_t)SCANN_D, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)SCANN_D * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)PT};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { scann_set_error("cuTensorMapEncodeTiled failed with code %d", (int)r); return 1; }
    return 0;
}

static int la_pipe_fwd_configure() {
    static bool configured = false;
    if (configured) return 0;
    cudaError_t e = cudaFuncSetAttribute(la_geom_fwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la_attn_fwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_SMEM);
    if (e != cudaSuccess) { scann_set_error("la_forward_pipe: smem opt-in failed: %s", cudaGetErrorString(e)); return 1; }
    configured = true;
    return 0;
}

// Pipelined forward of LocalAttention.call (attention.py:118-216, g_update = True) for pair plans with
// tile_stride 32.  rows = tile_cap * 32 = the row count of every per-pair tensor.  which: bit 0 geometry kernel,
// bit 1 attention kernel (development: run one half with the other half left to scann_la_forward_tc's kernels).
extern "C" int scann_la_forward_pipe(int grid, long long rows, int which, const int32_t* ntiles, const int32_t* pair_c,
                                     const int32_t* pair_j, const float* x, const float* proj, const float* g_in,
                                     const float* W2, const float* Wk, const float* bk, const float* gamma_g,
                                     const float* beta_g, const float* gamma, const float* beta, float* g_out,
                                     float* ctx_pre, float* out, float* attn, float* pre_out, float* k_out,
                                     const void* attn_drop, int drop_site, int32_t* status, void* stream) {
    if (la_pipe_fwd_configure()) return 1;
    if (grid <= 0) return 0;
    if (which & 1) {
        PipeGeomArgs ga;
        memset(&ga, 0, sizeof(ga));
        if (pipe_encode_tmap(&ga.tm_gin, g_in, rows) || pipe_encode_tmap(&ga.tm_gout, g_out, rows)) return 1;
        if (pre_out && pipe_encode_tmap(&ga.tm_pre, pre_out, rows)) return 1;
        ga.ntiles = ntiles; ga.pair_c = pair_c; ga.pair_j = pair_j; ga.proj = proj; ga.W2 = W2;
        ga.gamma_g = gamma_g; ga.beta_g = beta_g; ga.has_pre = pre_out ? 1 : 0; ga.status = status;
        scann_launch(la_geom_fwd_pipe_kernel, dim3(grid), dim3(PF_THREADS), PF_SMEM, stream, ga);
    }
    if (which & 2) {
        PipeAttnArgs aa;
        memset(&aa, 0, sizeof(aa));
        if (pipe_encode_tmap(&aa.tm_g, g_out, rows)) return 1;
        if (k_out && pipe_encode_tmap(&aa.tm_k, k_out, rows)) return 1;
        aa.ntiles = ntiles; aa.pair_c = pair_c; aa.pair_j = pair_j; aa.x = x; aa.proj = proj; aa.Wk = Wk; aa.bk = bk;
        aa.gamma = gamma; aa.beta = beta; aa.ctx_pre = ctx_pre; aa.out = out; aa.attn = attn; aa.has_k = k_out ? 1 : 0;
        aa.drop = (const ScannDropCtl*)attn_drop; aa.drop_site = drop_site; aa.status = status;
        scann_launch(la_attn_fwd_pipe_kernel, dim3(grid), dim3(PF_THREADS), PF_SMEM, stream, aa);
    }
    return scann_check_launch("scann_la_forward_pipe");
}
