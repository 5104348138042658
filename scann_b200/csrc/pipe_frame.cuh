// Frame shared by the warp-specialised local-attention kernels (la_pipe.cu forward, la_pipe_bwd.cu backward):
// the shared-memory stage ring, the role layout of a CTA, the consumer-issued 3xTF32 product chains and the store warp.
//
// CTA = PF_NG consumer groups of PF_GW warps (a group owns tiles g, g + PF_NG, ... of the CTA; each warp 4 rows of a
// 32-row tile), one producer warp and one store warp.  A stage of the ring holds IMGS tile images (pipe_common.cuh),
// the tile's centre / neighbour indices and a small per-row scratch (attention scores); NS stages fill the SM's
// shared memory: forward NS = 6 x 2 images, backward NS = 3 x 4 images.
#pragma once
#include "pipe_common.cuh"

#define PF_NG 2                                // consumer groups
#define PF_GW 8                                // warps per consumer group (each warp: 4 rows of the tile)
#define PF_GT (PF_GW * 32)
#define PF_CW (PF_GW * PF_NG)                  // consumer warps
#define PF_THREADS ((PF_CW + 2) * 32)          // + producer warp + store warp

template <int NS, int IMGS>
struct PipeFrame {
    static constexpr uint32_t STAGE = IMGS * PT_IMG;
    static constexpr uint32_t OFF_IDX = NS * STAGE;                       // idx[NS][64]: centre | neighbour atom rows
    static constexpr uint32_t OFF_ES = OFF_IDX + NS * 64u * 4u;           // es[NS][2][PT*8]
    static constexpr uint32_t OFF_BAR = OFF_ES + NS * 2u * PT * 8u * 4u;  // full | empty | done | accf
    static constexpr uint32_t OFF_FLAGS = OFF_BAR + 4u * NS * 8u;
    static constexpr uint32_t SMEM = OFF_FLAGS + 16u + 1024u;             // + alignment slack
    uint8_t* stages; int32_t* idx; float* es; uint64_t *full, *empty, *ready, *accf; uint32_t* tmem_slot; volatile int* dead;
    __device__ __forceinline__ void carve(uint8_t* smem_raw) {
        uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
        stages = smem;
        idx = reinterpret_cast<int32_t*>(smem + OFF_IDX);
        es = reinterpret_cast<float*>(smem + OFF_ES);
        full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
        empty = full + NS; ready = empty + NS; accf = ready + NS;
        tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_FLAGS);
        dead = reinterpret_cast<volatile int*>(smem + OFF_FLAGS + 4);
    }
};

__device__ __forceinline__ float4 pf4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 lds4(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(uint8_t* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float4 tf32_lo4(float4 v) { return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w)); }
// barrier among the PF_GT threads of consumer group g (barrier 0 stays __syncthreads)
__device__ __forceinline__ void pf_group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(PF_GT) : "memory"); }

// accumulator (lane = feature n, column = row r) -> image S[r][n] (+ bias[n]); warp of an 8-warp group: lane quarter
// q = warp % 4, row half h = (warp / 4) % 2
// The three products of 3xTF32 sit in three accumulators of PT columns each (see pf_issue_chain): the two correction
// products are added first, then the main product -- the accurate order of profiles/r01_tcgen05_probe.md.
__device__ __forceinline__ void pf_acc_to_image(uint32_t t_acc, uint8_t* S, const float* __restrict__ bias, int q, int h,
                                                int lane) {
    const int n = q * 32 + lane;
    const float b = bias ? __ldg(bias + n) : 0.f;
    float c0[16], c1[16], m[16];
    const uint32_t base = t_acc + ((uint32_t)(q * 32) << 16) + h * 16;
    tmem_ld16(base, c0);
    tmem_ld16(base + PT, c1);
    tmem_ld16(base + 2 * PT, m);
    tmem_ld_wait();
#pragma unroll
    for (int r = 0; r < 16; ++r) *reinterpret_cast<float*>(S + pt_off(h * 16 + r, n)) = ((c0[r] + c1[r]) + m[r]) + b;
}

// One of the three products of D^T = W^T X^T (X = x_raw + x_lo as two tile images), 16 K-steps into its own
// accumulator; issued by one elected lane of consumer warp `chain` of the tile's group.  A single thread sustains one
// tcgen05.mma per ~45 cycles whatever its N (profiles/r02b_tc_time2.log: 48 MMAs from one thread = 2 160 cycles per
// 32-row tile against a math floor of 768), so the three chains are issued by three warps at the same time.
__device__ __forceinline__ void pf_issue_chain(int chain, uint32_t t_wraw, uint32_t t_wlo, uint32_t img_raw, uint32_t img_lo,
                                               uint32_t t_acc, uint64_t* bar) {
    const uint32_t idesc = tc_idesc_tf32(128, PT, false, false);
    const uint32_t tw = chain == 0 ? t_wlo : t_wraw;
    const uint64_t d = pt_desc(chain == 1 ? img_lo : img_raw);
    const uint32_t acc = t_acc + chain * PT;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(acc, tw + ks * 8, d + PT_KSTEP(ks), idesc, ks != 0);
    tc_commit(bar);
}

// (Round 2, measured: reading the tile count BEFORE griddepcontrol.wait -- legal, the pair plan is never the direct
// predecessor of these kernels -- made the QM9 train step 0.8 % SLOWER (0.976 -> 0.984 ms, two runs each, one box,
// gpurun_out/r02ca_ab.log) and left MP2018 unchanged; the count is read behind the wait.)
// prologue shared by both kernels: tensor-memory allocation, barrier init, the stationary weight
#define PF_PROLOGUE(FRAME, NSTAGES, W_PTR, FULL_COUNT)                                                                   \
    extern __shared__ uint8_t smem_raw[];                                                                                \
    FRAME c;                                                                                                             \
    c.carve(smem_raw);                                                                                                   \
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                                                       \
    if (warp == PF_CW + 1) tmem_alloc(c.tmem_slot, 512);                                                                 \
    if (tid == 0) {                                                                                                      \
        for (int s = 0; s < NSTAGES; ++s) {                                                                                \
            mbar_init(&c.full[s], FULL_COUNT); mbar_init(&c.empty[s], 1); mbar_init(&c.ready[s], 1); mbar_init(&c.accf[s], 3); \
        }                                                                                                                \
        *c.dead = 0;                                                                                                     \
        mbar_fence_init();                                                                                               \
    }                                                                                                                    \
    tc_fence_before();                                                                                                   \
    __syncthreads();                                                                                                     \
    tc_fence_after();                                                                                                    \
    const uint32_t tmem = *c.tmem_slot;                                                                                  \
    const uint32_t t_wraw = tmem, t_wlo = tmem + 128, t_acc0 = tmem + 256; /* + group * 3 * PT + chain * PT */           \
    /* the stationary weight (parameters only: before the PDL wait) is staged by the consumer warps while the producer */ \
    /* warp already streams the first tiles in; only the consumers -- who issue the MMAs -- wait for it */                \
    if (warp < PF_CW) pipe_weight_to_tmem(W_PTR, t_wraw, t_wlo, warp, lane, PF_CW);                                      \
    pdl_wait();                                                                                                          \
    const int nt = *a.ntiles;                                                                                            \
    if (warp < PF_CW) {                                                                                                  \
        tc_fence_before();                                                                                               \
        asm volatile("bar.sync 15, %0;" ::"n"(PF_CW * 32) : "memory");                                                   \
        tc_fence_after();                                                                                                \
    }

// Store warp (warp PF_CW + 1): when a consumer group has finished a tile in place (`done` = the ready[] barrier of the
// stage), one lane issues the TMA stores of its images and hands the stage back to the producer as soon as THOSE stores
// have finished reading shared memory (a few hundred cycles) -- the consumers never wait for the TMA unit.  (Round 2,
// first form: the stage went back one tile later, behind the next tile's store issue; with the backward ring only
// three stages deep that left the producer nothing to prefetch into, and 18 % of the consumers' samples sat in the
// wait for a landed tile, profiles/r02_summary.md.)
#define PF_STORE_LOOP(NSTAGES, STAGE_BYTES, CODE, STORE_STMTS)                                                                                 \
    {                                                                                                                    \
        if (lane == 0) {                                                                                                 \
            int i = 0;                                                                                                   \
            for (int t = blockIdx.x; t < nt; t += gridDim.x, ++i) {                                                      \
                const int s = i % NSTAGES;                                                                               \
                const uint32_t ph = (uint32_t)(i / NSTAGES) & 1u;                                                        \
                pipe_wait(&c.ready[s], ph, c.dead, a.status, CODE, t, s);                                                \
                uint8_t* A = c.stages + (size_t)s * STAGE_BYTES;                                                         \
                uint8_t* Bm = A + PT_IMG;                                                                                \
                (void)Bm;                                                                                                \
                STORE_STMTS                                                                                              \
                tma_commit();                                                                                            \
                tma_wait_read0();                                                                                        \
                mbar_arrive(&c.empty[s]);                                                                                \
            }                                                                                                            \
        }                                                                                                                \
        __syncwarp();                                                                                                    \
    }

