// All weight-gradient GEMMs of one train step in ONE persistent launch (tcgen05, kind::tf32, 3xTF32).
//
// Every Dense kernel gradient of the graph is a reduction over rows, dW = X^T Y (TF autodiff of the Dense /
// einsum layers in scann/layers/attention.py:118-216 and scann/models/scann_model.py:424-447 inside keras fit):
//   * per pair  : key/kernel      += (x[j] * g')^T d_k        filter_geo rows 128..255 += g^T d_pre
//   * per atom  : filter_geo rows 0..127 / 256..383, query, ResidualNorm dense / dense_1, after_Lc,
//                 GlobalAttention query / key, bf_property  (+ the bias gradients = column sums of Y)
// They only feed the optimiser, so nothing in the backward chain waits for them.  Launched layer by layer on
// a side stream they cost more SM time than they hide (measured: 0.42 of the 1.59 ms step is lost to them,
// mostly per-launch partial-sum traffic: 148 CTAs x 64 KB x 2 per layer and wave quantisation).  Here the
// ~50 problems are concatenated into one list of 32-row units, the list is cut into equal contiguous ranges,
// one per CTA, and a CTA accumulates a problem's units in tensor memory and flushes each accumulator ONCE with
// vector atomics (red.global.add.v4.f32) straight into the gradient arena: no partial buffers, no reduce
// kernels, perfect balance.
//
// Operands are the MN-major SWIZZLE_128B_BASE32B images of tc_common.cuh (K = pair / atom rows).
#include "common.cuh"
#include "tc_common.cuh"

#define WG_THREADS 512
#define WG_WARPS 16
#define WG_MAX_PROBLEMS 96

// Mirrors ScannWgradProblem in include/scann_b200.h.
struct WgProblem {
    const float* X;      // [rows, ldx] left operand rows
    const float* Y;      // [rows, ldy] right operand rows
    const float* xg;     // pair problems: X rows are multiplied by xg[pair_j[row], :] ([R,128]); else NULL
    float* dW;           // [128,128], accumulated
    float* db;           // [128] += column sums of Y, nullable
    int ldx, ldy;
    int rows;            // >= 0: atom problem with this many rows ; < 0: pair problem (rows = ntiles * stride,
                         //       rows whose pair_c < 0 are padding)
    int pad;
};

struct WgArgs {
    const WgProblem* prob;
    int nprob;
    const int32_t* ntiles;
    int stride;
    const int32_t* pair_c;
    const int32_t* pair_j;
    const int32_t* valid_rows;   // compact list of the valid pair rows (scann_plan_build), nullable
    const int32_t* valid_j;      // neighbour atom row of each compact entry
    const int32_t* nvalid;       // its length (device)
};

__device__ __forceinline__ void wg_split_store(uint8_t* sHi, uint8_t* sLo, uint32_t off, float4 v) {
    float4 h, l;
    tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(sHi + off) = h;
    *reinterpret_cast<float4*>(sLo + off) = l;
}

#define WG_UR 32                       // rows per unit (= K extent of one MMA group)
#define WG_NSTAGE 2                    // staging buffers of raw rows (units u, u+1 in flight / being consumed)
#define WG_NIMG 2                      // operand image sets: the split of unit u+1 overlaps the MMAs of unit u
#define WG_RPW (WG_UR / WG_WARPS)      // rows per warp per unit
#define WG_MNB ((uint32_t)WG_UR * 128u)          // one column block of an MN-major image: [UR rows x 32 columns]
#define WG_MNT (4u * WG_MNB)                     // one image [UR x 128] fp32 = 32 KB
#define WG_STAGE ((uint32_t)WG_UR * 512u)        // raw rows of one operand of one unit (a stage holds X then Y)

__device__ __forceinline__ uint32_t wg_mn_off(int r, int c) {
    return (uint32_t)(c >> 5) * WG_MNB + (uint32_t)r * 128u + (((((uint32_t)c >> 3) & 3u) ^ ((uint32_t)r & 3u)) << 5) +
           ((uint32_t)c & 7u) * 4u;
}
__device__ __forceinline__ uint64_t wg_mn_desc(uint32_t saddr) { return tc_desc(saddr, WG_MNB, 512u) | ((uint64_t)1 << 61); }
// 16-byte asynchronous global -> shared copy; nbytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int nbytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Pipeline per CTA (the inputs come from HBM: 13 % L2 hit rate measured, so the loads of the next unit must be in
// flight while the current one is split and multiplied):
//   cp.async raw rows of units u+1..u+3 -> staging ring | staging(u) -> hi/lo split -> 4 MN-major images -> MMAs
// Every thread copies and later reads the SAME 16-byte chunks, so cp.async.wait_group is the only synchronisation
// the staging ring needs; the images are single-buffered (the split of unit u+1 waits for the MMAs of unit u).
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_batch_tc_kernel(const WgArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // the BASE32B swizzle is a function of the absolute shared address: align the images to 1024 bytes
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sRaw = smem + WG_NIMG * 4 * WG_MNT;        // WG_NSTAGE x [X rows | Y rows]
    __shared__ uint64_t bars[WG_NIMG];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_db[SCANN_D];
    __shared__ int s_units[WG_MAX_PROBLEMS + 1];        // exclusive prefix of units per problem
    __shared__ WgProblem s_prob[WG_MAX_PROBLEMS];       // the problem table (a dependent global load per unit otherwise)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // pair problems run over the compact list of valid rows when the plan provides it (independent of how
    // sparsely the tiles are filled), else over all tile slots with the padding rows skipped
    const int pair_rows = a.valid_rows ? *a.nvalid : *a.ntiles * a.stride;
    for (int i = tid; i < a.nprob * (int)(sizeof(WgProblem) / 8); i += WG_THREADS)
        reinterpret_cast<uint64_t*>(s_prob)[i] = reinterpret_cast<const uint64_t*>(a.prob)[i];
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int p = 0; p < a.nprob; ++p) {
            s_units[p] = run;
            const int rows = s_prob[p].rows >= 0 ? s_prob[p].rows : pair_rows;
            run += (rows + WG_UR - 1) / WG_UR;
        }
        s_units[a.nprob] = run;
    }
    if (tid < SCANN_D) s_db[tid] = 0.f;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid < WG_NIMG) mbar_init(&bars[tid], 1);
    if (tid == 0) mbar_fence_init();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int total = s_units[a.nprob];
    const long long G = gridDim.x;
    const int u_lo = (int)((long long)total * blockIdx.x / G), u_hi = (int)((long long)total * (blockIdx.x + 1) / G);
    const uint32_t t_dm = tmem_base_s, t_dc = tmem_base_s + 128;
    const uint32_t idesc = tc_idesc_tf32(128, 128, true, true);
    uint32_t phase[WG_NIMG];
    bool pending[WG_NIMG];          // an MMA group reading image set b is in flight (not yet waited for)
#pragma unroll
    for (int b = 0; b < WG_NIMG; ++b) { phase[b] = 0; pending[b] = false; }
    bool first = true;              // no MMA has been issued into the accumulators of the current problem
    int p = 0, pl = 0;              // problem of the unit being multiplied / being loaded
    float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f);
    int pjq[WG_NSTAGE][WG_RPW];     // neighbour indices of the units in flight (pjq[k] belongs to unit u + k)
#pragma unroll
    for (int k = 0; k < WG_NSTAGE; ++k)
#pragma unroll
        for (int i = 0; i < WG_RPW; ++i) pjq[k][i] = 0;

    // Indices (pair_c, pair_j) of the unit that will be loaded NEXT are fetched one call ahead: otherwise every
    // iteration starts with two dependent L2 round trips before its cp.async can be issued.
    int pi = 0;                     // problem of the unit whose indices are being fetched
    int pcn[WG_RPW], pjx[WG_RPW];
    auto prefetch_idx = [&](int u) {
#pragma unroll
        for (int i = 0; i < WG_RPW; ++i) { pcn[i] = -1; pjx[i] = 0; }
        if (u >= u_hi) return;
        while (u >= s_units[pi + 1]) ++pi;
        const WgProblem& pr = s_prob[pi];
        const int nrows = pr.rows < 0 ? pair_rows : pr.rows;
        const size_t rowbase = (size_t)(u - s_units[pi]) * WG_UR;
#pragma unroll
        for (int i = 0; i < WG_RPW; ++i) {
            const size_t r = rowbase + warp + WG_WARPS * i;
            if (r < (size_t)nrows) {
                pcn[i] = (int)r;                        // source row (>= 0) or -1
                if (pr.rows < 0) {
                    if (a.valid_rows) {
                        pcn[i] = a.valid_rows[r];
                        if (pr.xg) pjx[i] = a.valid_j[r];         // independent of the load above: one round trip
                    } else {
                        if (a.pair_c[r] < 0) pcn[i] = -1;
                        if (pr.xg) pjx[i] = a.pair_j[r];
                    }
                }
            }
        }
    };
    // asynchronous copy of the raw operand rows of unit u into its staging slot (+ neighbour indices into pjn);
    // always commits a group (possibly empty) so that the group count per iteration is constant.  Calls are made
    // for consecutive units; each call leaves the indices of unit u+1 in flight.
    auto issue_load = [&](int u, int (&pjn)[WG_RPW]) {
        if (u >= u_hi) { cp_async_commit(); return; }
        uint8_t* sRawX = sRaw + (size_t)(u % WG_NSTAGE) * 2 * WG_STAGE;
        uint8_t* sRawY = sRawX + WG_STAGE;
        while (u >= s_units[pl + 1]) ++pl;
        const WgProblem& pr = s_prob[pl];
#pragma unroll
        for (int i = 0; i < WG_RPW; ++i) {
            const int rr = warp + WG_WARPS * i;
            const bool ok = pcn[i] >= 0;
            pjn[i] = ok ? pjx[i] : 0;
            const size_t rs = ok ? (size_t)pcn[i] : 0;
            cp_async16(smem_u32(sRawX) + rr * 512 + lane * 16, pr.X + rs * pr.ldx + lane * 4, ok ? 16 : 0);
            cp_async16(smem_u32(sRawY) + rr * 512 + lane * 16, pr.Y + rs * pr.ldy + lane * 4, ok ? 16 : 0);
        }
        cp_async_commit();
        prefetch_idx(u + 1);
    };
    // accumulators -> gradient arena (vector atomics), column sums -> bias gradient
    auto flush = [&](const WgProblem& pr) {
#pragma unroll
        for (int b = 0; b < WG_NIMG; ++b)
            if (pending[b]) { mbar_wait(&bars[b], phase[b]); phase[b] ^= 1; tc_fence_after(); pending[b] = false; }
        if (!first) {
            const int m = (warp & 3) * 32 + lane, nbase = (warp >> 2) * 32;
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[16], c[16];
                tmem_ld16(t_dm + lane_base + nbase + h * 16, v);
                tmem_ld16(t_dc + lane_base + nbase + h * 16, c);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; q += 4)
                    red_add4(pr.dW + (size_t)m * SCANN_D + nbase + h * 16 + q, v[q] + c[q], v[q + 1] + c[q + 1],
                             v[q + 2] + c[q + 2], v[q + 3] + c[q + 3]);
            }
            if (pr.db) {
                atomicAdd(&s_db[lane * 4 + 0], colsum.x); atomicAdd(&s_db[lane * 4 + 1], colsum.y);
                atomicAdd(&s_db[lane * 4 + 2], colsum.z); atomicAdd(&s_db[lane * 4 + 3], colsum.w);
            }
            tc_fence_before();
            __syncthreads();                 // every warp has read the accumulators; s_db is complete
            tc_fence_after();
            if (pr.db && tid < SCANN_D) { atomicAdd(pr.db + tid, s_db[tid]); s_db[tid] = 0.f; }
            __syncthreads();
        }
        first = true;
        colsum = make_float4(0.f, 0.f, 0.f, 0.f);
    };

    prefetch_idx(u_lo);
#pragma unroll
    for (int k = 0; k < WG_NSTAGE - 1; ++k) issue_load(u_lo + k, pjq[k]);
#pragma unroll 1
    for (int u = u_lo; u < u_hi; ++u) {
        while (u >= s_units[p + 1]) {        // next problem: hand over what was accumulated for the previous one
            flush(s_prob[p]);
            ++p;
        }
        const float* xg = s_prob[p].xg;
        // refill the slot that was consumed in the previous iteration (this thread's own chunks)
        issue_load(u + WG_NSTAGE - 1, pjq[WG_NSTAGE - 1]);
        // neighbour rows x[j] (L2 resident) while the asynchronous copies land
        float4 nb[WG_RPW];
        if (xg) {
#pragma unroll
            for (int i = 0; i < WG_RPW; ++i) nb[i] = ld4(xg + (size_t)pjq[0][i] * SCANN_D + lane * 4);
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(WG_NSTAGE - 1) : "memory");      // unit u has landed
        const uint8_t* sRawX = sRaw + (size_t)(u % WG_NSTAGE) * 2 * WG_STAGE;
        const uint8_t* sRawY = sRawX + WG_STAGE;
        float4 xv[WG_RPW], yv[WG_RPW];
#pragma unroll
        for (int i = 0; i < WG_RPW; ++i) {
            const int rr = warp + WG_WARPS * i;
            xv[i] = *reinterpret_cast<const float4*>(sRawX + rr * 512 + lane * 16);
            yv[i] = *reinterpret_cast<const float4*>(sRawY + rr * 512 + lane * 16);
            if (xg) xv[i] = make_float4(xv[i].x * nb[i].x, xv[i].y * nb[i].y, xv[i].z * nb[i].z, xv[i].w * nb[i].w);
            colsum = make_float4(colsum.x + yv[i].x, colsum.y + yv[i].y, colsum.z + yv[i].z, colsum.w + yv[i].w);
        }
#pragma unroll
        for (int k = 0; k + 1 < WG_NSTAGE; ++k)
#pragma unroll
            for (int i = 0; i < WG_RPW; ++i) pjq[k][i] = pjq[k + 1][i];
        const int ib = u & (WG_NIMG - 1);
        uint8_t* sXh = smem + (size_t)ib * 4 * WG_MNT;
        uint8_t* sXl = sXh + WG_MNT;
        uint8_t* sYh = sXh + 2 * WG_MNT;
        uint8_t* sYl = sXh + 3 * WG_MNT;
#pragma unroll
        for (int b = 0; b < WG_NIMG; ++b)
            if (b == ib && pending[b]) {      // the MMAs of unit u - 2 still read this image set
                mbar_wait(&bars[b], phase[b]);
                phase[b] ^= 1;
                tc_fence_after();
                pending[b] = false;
            }
#pragma unroll
        for (int i = 0; i < WG_RPW; ++i) {
            const uint32_t off = wg_mn_off(warp + WG_WARPS * i, lane * 4);
            wg_split_store(sXh, sXl, off, xv[i]);
            wg_split_store(sYh, sYl, off, yv[i]);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0 && tc_elect_one()) {
            tc_fence_after();
            const uint64_t dxh = wg_mn_desc(smem_u32(sXh)), dxl = wg_mn_desc(smem_u32(sXl)), dyh = wg_mn_desc(smem_u32(sYh)),
                           dyl = wg_mn_desc(smem_u32(sYl));
#pragma unroll
            for (int ks = 0; ks < WG_UR / 8; ++ks)     // K-step = 8 rows = 1024 bytes of each image
                tc_mma_ss(t_dm, dxh + (uint64_t)(ks * 64), dyh + (uint64_t)(ks * 64), idesc, !(first && ks == 0));
#pragma unroll
            for (int ks = 0; ks < WG_UR / 8; ++ks)
                tc_mma_ss(t_dc, dxh + (uint64_t)(ks * 64), dyl + (uint64_t)(ks * 64), idesc, !(first && ks == 0));
#pragma unroll
            for (int ks = 0; ks < WG_UR / 8; ++ks)
                tc_mma_ss(t_dc, dxl + (uint64_t)(ks * 64), dyh + (uint64_t)(ks * 64), idesc, true);
            tc_commit(&bars[ib]);
        }
#pragma unroll
        for (int b = 0; b < WG_NIMG; ++b)
            if (b == ib) pending[b] = true;
        first = false;
    }
    if (u_hi > u_lo) flush(s_prob[p]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 256);
}

#define WG_SMEM (WG_NIMG * 4 * WG_MNT + WG_NSTAGE * 2 * WG_STAGE + 1024)

// problems: DEVICE array of nprob ScannWgradProblem (include/scann_b200.h).  Gradients are accumulated
// (atomics) into dW / db, which the caller zeroes at the start of the step.
extern "C" int scann_wgrad_batch_tc(int grid, const void* problems_dev, int nprob, const int32_t* ntiles, int tile_stride,
                                    const int32_t* pair_c, const int32_t* pair_j, const int32_t* valid_rows,
                                    const int32_t* valid_j, const int32_t* nvalid, void* stream) {
    if (nprob < 1 || nprob > WG_MAX_PROBLEMS) { scann_set_error("wgrad_batch_tc: nprob must be in 1..%d", WG_MAX_PROBLEMS); return 1; }
    if (tile_stride != 64 && tile_stride != 128) { scann_set_error("wgrad_batch_tc: tile_stride must be 64 or 128"); return 1; }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_batch_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM);
        if (e != cudaSuccess) { scann_set_error("wgrad_batch_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    if (grid <= 0) return 0;
    WgArgs a{(const WgProblem*)problems_dev, nprob, ntiles, tile_stride, pair_c, pair_j, valid_rows, valid_j,
             valid_rows ? nvalid : nullptr};
    wgrad_batch_tc_kernel<<<grid, WG_THREADS, WG_SMEM, (cudaStream_t)stream>>>(a);
    return scann_check_launch("scann_wgrad_batch_tc");
}
