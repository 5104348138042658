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
//   MMA warps       : one per consumer group; an elected lane issues the 48 tcgen05.mma of a tile once its images are
//                     ready, into the stage's own accumulator columns, and commits to the stage's mbarrier.
//   consumer groups : PF_NG groups of eight warps; group g owns tiles g, g + PF_NG, ...: split pass (lo = x -
//                     trunc(x); the TMA-landed fp32 tile itself is the "hi" operand because the tensor core
//                     truncates), accumulator read-back, row-wise epilogue IN PLACE in the stage's images, and a
//                     TMA store of the finished tile (g', and pre / k when training saves them).
//
// While group g runs the epilogue of tile t, the tensor core works on tile t+1 of group g+1 and the TMA unit
// loads tiles t+2..; the phases that used to run in lock-step on every SM (profiles/r01i_tc_summary.md) overlap.
// Tiles are PT = 32 pair rows (plan tile_stride 32, at most 32 neighbours per atom); a stage is two 16 KB tile
// images, so six tiles are in flight per SM.  Every mbarrier wait is bounded (pipe_common.cuh).
#include <string.h>

#include "pipe_frame.cuh"

#define PF_NS 6                                // ring stages: two tile images each
typedef PipeFrame<PF_NS, 2> FwdFrame;
#define PF_STAGE (FwdFrame::STAGE)
#define PF_SMEM (FwdFrame::SMEM)

// phase timestamps of CTA 0, consumer group 0 (development builds: -DSCANN_DEV_PROBES): [kernel][tile ordinal][phase]
#ifdef SCANN_DEV_PROBES
__device__ long long g_pipe_clk[2][4][12];
#define PCLK(k, ph) do { if (blockIdx.x == 0 && tid == 0 && (i / PF_NG) < 4) g_pipe_clk[k][i / PF_NG][ph] = clock64(); } while (0)
#define PCLK0(k, ph) do { if (blockIdx.x == 0 && tid == 0) g_pipe_clk[k][0][ph] = clock64(); } while (0)
extern "C" int scann_pipe_clocks(long long* host_out96) {
    cudaError_t e = cudaMemcpyFromSymbol(host_out96, g_pipe_clk, sizeof(long long) * 96);
    if (e != cudaSuccess) { scann_set_error("pipe_clocks: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}
#else
#define PCLK(k, ph) do { } while (0)
#define PCLK0(k, ph) do { } while (0)
#endif

struct PipeGeomArgs {
    CUtensorMap tm_gin, tm_gout, tm_pre;
    const int32_t* ntiles; const int32_t* pair_c; const int32_t* pair_j;
    const float* proj;       // [R,384] = [x@W1+bf | x@W3 | x@Wq+bq]
    const float* W2;         // rows 128..255 of filter_geo/kernel, [k][n]
    const float* gamma_g; const float* beta_g;
    int has_pre;             // training: save the filter_geo pre-activation through tm_pre
    int32_t* status;
};

// =============================================================================================
// Geometry update
// =============================================================================================
__global__ void __launch_bounds__(PF_THREADS, 1) la_geom_fwd_pipe_kernel(const __grid_constant__ PipeGeomArgs a) {
    PF_PROLOGUE(FwdFrame, PF_NS, a.W2, 1)
    if (warp == PF_CW) {
        // ================= producer =================
        if (lane == 0) {
            tma_prefetch_desc(&a.tm_gin);
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
    } else if (warp > PF_CW) {
        if (lane == 0) { tma_prefetch_desc(&a.tm_gout); if (a.has_pre) tma_prefetch_desc(&a.tm_pre); }
        PF_STORE_LOOP(PF_NS, PF_STAGE, 2,
            for (int kb = 0; kb < 4; ++kb) tma_store_2d(&a.tm_gout, A + kb * PT_CB, kb * 32, t * PT);
            if (a.has_pre) { for (int kb = 0; kb < 4; ++kb) tma_store_2d(&a.tm_pre, Bm + kb * PT_CB, kb * 32, t * PT); })
    } else {
        // ================= consumers =================
        const int grp = warp / PF_GW, wgl = warp % PF_GW, q = warp & 3, half = (warp >> 2) & 1, gtid = tid - grp * PF_GT;
        const int l8 = lane & 7, rsub = lane >> 3;
        const int r = wgl * 4 + rsub;                        // this lane's row of every tile
        const uint32_t t_acc = t_acc0 + grp * 3 * PT;        // a group has one tile between MMA issue and read-back
        int i = grp;
        for (int t = blockIdx.x + grp * gridDim.x; t < nt; t += PF_NG * gridDim.x, i += PF_NG) {
            const int s = i % PF_NS;
            const uint32_t ph = (uint32_t)(i / PF_NS) & 1u;
            uint8_t* A = c.stages + (size_t)s * PF_STAGE;       // g (fp32, as landed) -> g'
            uint8_t* Bm = A + PT_IMG;                           // lo image -> transposed accumulator -> pre
            const int32_t* sidx = c.idx + s * 64;
            PCLK(0, 0);
            pipe_wait(&c.full[s], ph, c.dead, a.status, 3, t, s);
            PCLK(0, 1);
            // ---- gathered per-atom projections of this lane's row: issued first, consumed after the MMA
            const int pc = sidx[r];
            float4 pa[4], pb[4];
            {
                const int j = pc >= 0 ? sidx[32 + r] : 0;
                const int cc = pc >= 0 ? pc : 0;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int c0 = (l8 + 8 * it) * 4;
                    pa[it] = ld4(a.proj + (size_t)cc * 3 * SCANN_D + c0);
                    pb[it] = ld4(a.proj + (size_t)j * 3 * SCANN_D + SCANN_D + c0);
                }
            }
            // ---- split pass: the landed tile is the hi operand; lo = g - trunc(g).  Loads first, then stores
            {
                float4 v[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) v[it] = lds4(A + pt_off4(wgl + PF_GW * it, lane));
#pragma unroll
                for (int it = 0; it < 4; ++it) sts4(Bm + pt_off4(wgl + PF_GW * it, lane), tf32_lo4(v[it]));
            }
            fence_async_smem();
            PCLK(0, 9);
            tc_fence_before();
            pf_group_sync(grp);
            if (wgl < 3) {                                      // warps 0..2 of the group: one product chain each
                tc_fence_after();
                if (tc_elect_one()) pf_issue_chain(wgl, t_wraw, t_wlo, smem_u32(A), smem_u32(Bm), t_acc, &c.accf[s]);
                __syncwarp();
            }
            PCLK(0, 2);
            pipe_wait(&c.accf[s], ph, c.dead, a.status, 4, t, s);
            tc_fence_after();
            PCLK(0, 3);
            // the gathered rows are consumed here, not earlier: the loads stay in flight behind the split pass and the MMAs
#pragma unroll
            for (int it = 0; it < 4; ++it)
                asm volatile("" : "+f"(pa[it].x), "+f"(pa[it].y), "+f"(pa[it].z), "+f"(pa[it].w), "+f"(pb[it].x), "+f"(pb[it].y),
                             "+f"(pb[it].z), "+f"(pb[it].w));
            float4 p13[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) p13[it] = pf4add(pa[it], pb[it]);
            PCLK(0, 4);
            pf_acc_to_image(t_acc, Bm, nullptr, q, half, lane);
            tc_fence_before();
            pf_group_sync(grp);
            PCLK(0, 5);
            // ---- row-wise epilogue: pre -> swish -> + g -> LayerNorm -> g' (in place)
            if (!__all_sync(0xffffffffu, pc < 0)) {
                float z[4][4], pre[4][4];
                float s1 = 0.f;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const uint32_t off = pt_off4(r, l8 + 8 * it);
                    const float4 acc = lds4(Bm + off), g = lds4(A + off);
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
                    for (int k = 0; k < 4; ++k) { z[it][k] -= sh; m1 += z[it][k]; m2 = fmaf(z[it][k], z[it][k], m2); }
                oct_sum2(m1, m2);
                m1 *= (1.0f / SCANN_D);
                const float inv = rsqrtf(fmaxf(m2 * (1.0f / SCANN_D) - m1 * m1, 0.f) + SCANN_LN_EPS);
                const bool ok = pc >= 0;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const uint32_t off = pt_off4(r, l8 + 8 * it);
                    const float4 gm = ldg4(a.gamma_g + (l8 + 8 * it) * 4), bt = ldg4(a.beta_g + (l8 + 8 * it) * 4);
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f), po = o;
                    if (ok) {
                        o = make_float4((z[it][0] - m1) * inv * gm.x + bt.x, (z[it][1] - m1) * inv * gm.y + bt.y,
                                        (z[it][2] - m1) * inv * gm.z + bt.z, (z[it][3] - m1) * inv * gm.w + bt.w);
                        po = make_float4(pre[it][0], pre[it][1], pre[it][2], pre[it][3]);
                    }
                    sts4(A + off, o);
                    sts4(Bm + off, po);
                }
            }
            PCLK(0, 6);
            fence_async_smem();
            pf_group_sync(grp);
            PCLK(0, 7);
            if (gtid == 0) mbar_arrive(&c.ready[s]);            // tile finished in place: over to the store warp
            PCLK(0, 8);
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

// full barrier of the attention kernel: the expect_tx arrive of lane 0 + the 32 cp.async arrives of the gathering lanes
__global__ void __launch_bounds__(PF_THREADS, 1) la_attn_fwd_pipe_kernel(const __grid_constant__ PipeAttnArgs a) {
    PF_PROLOGUE(FwdFrame, PF_NS, a.Wk, 33)
    if (warp == PF_CW) {
        // ================= producer: g' tile by TMA, x[j] rows by per-lane cp.async =================
        if (lane == 0) tma_prefetch_desc(&a.tm_g);
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
            for (int rr = 0; rr < PT; ++rr) {
                const int jr = __shfl_sync(0xffffffffu, j, rr);
                cp_async16(X + pt_off4(rr, lane), a.x + (size_t)jr * SCANN_D + lane * 4);
            }
            cp_async_arrive(&c.full[s]);
        }
    } else if (warp > PF_CW) {
        if (a.has_k) {
            if (lane == 0) tma_prefetch_desc(&a.tm_k);
            PF_STORE_LOOP(PF_NS, PF_STAGE, 12, for (int kb = 0; kb < 4; ++kb) tma_store_2d(&a.tm_k, A + kb * PT_CB, kb * 32, t * PT);)
        }
    } else {
        // ================= consumers =================
        const int grp = warp / PF_GW, wgl = warp % PF_GW, q = warp & 3, half = (warp >> 2) & 1, gtid = tid - grp * PF_GT;
        const int l8 = lane & 7, rsub = lane >> 3;
        const int r = wgl * 4 + rsub;
        const float4 gam = ldg4(a.gamma + lane * 4), bet = ldg4(a.beta + lane * 4);
        const uint32_t t_acc = t_acc0 + grp * 3 * PT;
        int i = grp;
        for (int t = blockIdx.x + grp * gridDim.x; t < nt; t += PF_NG * gridDim.x, i += PF_NG) {
            const int s = i % PF_NS;
            const uint32_t ph = (uint32_t)(i / PF_NS) & 1u;
            const size_t rowbase = (size_t)t * PT;
            uint8_t* A = c.stages + (size_t)s * PF_STAGE;       // g' -> a = x[j] * g' (hi operand) -> keys
            uint8_t* Bm = A + PT_IMG;                           // x[j] -> lo image
            const int32_t* sidx = c.idx + s * 64;
            float* Es = c.es + s * 2 * PT * 8;
            PCLK(1, 0);
            pipe_wait(&c.full[s], ph, c.dead, a.status, 13, t, s);
            PCLK(1, 1);
            const int pc = sidx[r];
            const bool ok = pc >= 0;
            float4 qv[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                qv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) qv[it] = ld4(a.proj + (size_t)pc * 3 * SCANN_D + 2 * SCANN_D + (l8 + 8 * it) * 4);
            }
            // ---- a = x[j] * g' in place: the fp32 product is the hi operand, lo = a - trunc(a) over x[j]
            // (rows of a warp whose four rows are all padding keep stale operands: they only feed their own columns)
            if (!__all_sync(0xffffffffu, !ok)) {
                float4 g[4], xv[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const uint32_t off = pt_off4(r, l8 + 8 * it);
                    g[it] = lds4(A + off);
                    xv[it] = lds4(Bm + off);
                }
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const uint32_t off = pt_off4(r, l8 + 8 * it);
                    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) av = make_float4(g[it].x * xv[it].x, g[it].y * xv[it].y, g[it].z * xv[it].z, g[it].w * xv[it].w);
                    sts4(A + off, av);
                    sts4(Bm + off, tf32_lo4(av));
                }
            }
            fence_async_smem();
            tc_fence_before();
            pf_group_sync(grp);
            PCLK(1, 2);
            if (wgl < 3) {
                tc_fence_after();
                if (tc_elect_one()) pf_issue_chain(wgl, t_wraw, t_wlo, smem_u32(A), smem_u32(Bm), t_acc, &c.accf[s]);
                __syncwarp();
            }
            PCLK(1, 3);
            pipe_wait(&c.accf[s], ph, c.dead, a.status, 14, t, s);
            tc_fence_after();
            PCLK(1, 4);
            pf_acc_to_image(t_acc, A, a.bk, q, half, lane);                  // keys k = a @ Wk + bk
            tc_fence_before();
            pf_group_sync(grp);
            PCLK(1, 5);
            // ---- scores e[r][h] = 0.25 <q_h, k_h>  (head h = 16 columns = 4 adjacent lanes of the row)
            if (!__all_sync(0xffffffffu, !ok)) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const float4 kv = lds4(A + pt_off4(r, l8 + 8 * it));
                    const float e = quad_sum(kv.x * qv[it].x + kv.y * qv[it].y + kv.z * qv[it].z + kv.w * qv[it].w) * 0.25f;
                    if ((l8 & 3) == 0) Es[r * 8 + 2 * it + (l8 >> 2)] = e;
                }
            }
            pf_group_sync(grp);
            PCLK(1, 6);
            // ---- softmax over an atom's rows, by HEAD-warps: lane = row, the rows of an atom are a contiguous run of lanes, so
            // max and sum are segmented shuffle reductions -- every warp is busy and the cost does not depend on the
            // neighbour count.  (First form: one warp per atom looped twice over its rows, three or four of the eight warps
            // worked: 3 000 of the 6 600 cycles of a tile, gpurun_out/r02ap_pipe_clocks_fwd.log.)
            const int myc = sidx[lane];
            const int prevc = __shfl_up_sync(0xffffffffu, myc, 1);
            const uint32_t vmask = __ballot_sync(0xffffffffu, myc >= 0);
            const uint32_t hmask = __ballot_sync(0xffffffffu, myc >= 0 && (lane == 0 || myc != prevc));
            const int nvalid = __popc(vmask);
            {
                const bool valid = myc >= 0;
                const uint32_t below = hmask & ((2u << lane) - 1u);
                const uint32_t above = lane < 31 ? hmask & ~((2u << lane) - 1u) : 0u;
                const int lo = valid ? 31 - __clz(below) : lane;
                const int hi = valid ? (above ? __ffs(above) - 2 : nvalid - 1) : lane;
                const int h = wgl;
                const float e = valid ? Es[lane * 8 + h] : -INFINITY;
                float mx = e;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float v = __shfl_down_sync(0xffffffffu, mx, o);
                    if (lane + o <= hi) mx = fmaxf(mx, v);
                }
                mx = __shfl_sync(0xffffffffu, mx, lo);
                const float pe = valid ? __expf(e - mx) : 0.f;
                float sm = pe;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float v = __shfl_down_sync(0xffffffffu, sm, o);
                    if (lane + o <= hi) sm += v;
                }
                sm = __shfl_sync(0xffffffffu, sm, lo);
                if (valid) {
                    // the softmax is normalised before the dropout (attention.py:189-192)
                    const float pn = pe * (1.0f / sm) * drop_mult(a.drop, a.drop_site, (uint32_t)(rowbase + lane) * 8u + h);
                    Es[lane * 8 + h] = pn;
                    if (a.attn) a.attn[(rowbase + lane) * 8 + h] = pn;
                }
            }
            pf_group_sync(grp);
            // ---- context, residual q, LayerNorm: by the warp that owns the atom's first row (rows 4 wgl .. 4 wgl + 3), four
            // rows per trip so that the shared-memory loads of a trip are in flight together
#pragma unroll 1
            for (int rr = 0; rr < 4; ++rr) {
                const int r0 = wgl * 4 + rr;
                if (!((hmask >> r0) & 1u)) continue;                      // (warp-uniform)
                const uint32_t nxt = r0 < 31 ? hmask & ~((2u << r0) - 1u) : 0u;
                const int n = (nxt ? __ffs(nxt) - 1 : nvalid) - r0;
                const int atom = sidx[r0];
                const int h = lane >> 2;
                const float4 qa = ld4(a.proj + (size_t)atom * 3 * SCANN_D + 2 * SCANN_D + lane * 4);
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
                for (int r2 = 0; r2 < n; r2 += 4) {
                    float pw[4];
                    float4 kv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int row = r0 + (r2 + u < n ? r2 + u : 0);
                        pw[u] = r2 + u < n ? Es[row * 8 + h] : 0.f;
                        kv[u] = lds4(A + pt_off4(row, lane));
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        c0 = fmaf(pw[u], kv[u].x, c0); c1 = fmaf(pw[u], kv[u].y, c1);
                        c2 = fmaf(pw[u], kv[u].z, c2); c3 = fmaf(pw[u], kv[u].w, c3);
                    }
                }
                c0 += qa.x; c1 += qa.y; c2 += qa.z; c3 += qa.w;
                if (a.ctx_pre) st4(a.ctx_pre + (size_t)atom * SCANN_D + lane * 4, make_float4(c0, c1, c2, c3));
                const float mean = warp_sum(c0 + c1 + c2 + c3) * (1.0f / SCANN_D);
                c0 -= mean; c1 -= mean; c2 -= mean; c3 -= mean;
                const float inv = rsqrtf(warp_sum(c0 * c0 + c1 * c1 + c2 * c2 + c3 * c3) * (1.0f / SCANN_D) + SCANN_LN_EPS);
                st4(a.out + (size_t)atom * SCANN_D + lane * 4,
                    make_float4(c0 * inv * gam.x + bet.x, c1 * inv * gam.y + bet.y, c2 * inv * gam.z + bet.z,
                                c3 * inv * gam.w + bet.w));
            }
            PCLK(1, 7);
            if (a.has_k) fence_async_smem();
            pf_group_sync(grp);
            PCLK(1, 8);
            // training: the keys leave through the store warp; inference: the stage goes straight back to the producer
            if (gtid == 0) mbar_arrive(a.has_k ? &c.ready[s] : &c.empty[s]);
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
