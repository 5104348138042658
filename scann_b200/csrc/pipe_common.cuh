// Building blocks of the warp-specialised local-attention kernels (la_pipe.cu, la_pipe_bwd.cu):
// TMA tensor loads / stores, per-row asynchronous gathers, the mbarrier stage ring with BOUNDED waits, the
// 128-byte-swizzled tile image that is at the same time the TMA box layout and the K-major tcgen05 operand.
//
// Tile image.  A tile is PT = 32 pair rows x 128 fp32.  In shared memory it is stored as four "chunk blocks"
// (32 columns = 128 bytes each) of [32 rows x 128 B]; inside a chunk block the 16-byte chunk i of row r sits at
// r * 128 + ((i ^ (r & 7)) << 4): CU_TENSOR_MAP_SWIZZLE_128B, which is what a TMA box [32 columns x 32 rows]
// writes and reads, and the canonical K-major SWIZZLE_128B layout of a tcgen05 shared-memory descriptor
// (rows = N extent, 8-row groups 1024 B apart, K-step of 8 tf32 = 32 B inside the 128-byte line).  One image
// therefore serves as TMA destination, tensor-core operand, row-wise epilogue scratch and TMA store source.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tc_common.cuh"

#define PT 32                                  // pair rows per tile (= MMA N extent)
#define PT_CB (PT * 128u)                      // bytes of one chunk block  [32 rows x 128 B]
#define PT_IMG (4u * PT_CB)                    // bytes of one tile image   [32 rows x 128 fp32] = 16 KB
#define SCANN_ERR_PIPE_TIMEOUT 16              // an mbarrier wait of a pipelined kernel gave up (status word)

// byte offset of the 16-byte chunk c4 (0..31) of row r inside a tile image
__device__ __forceinline__ uint32_t pt_off4(int r, int c4) {
    return (uint32_t)(c4 >> 3) * PT_CB + (uint32_t)r * 128u + ((((uint32_t)c4 & 7u) ^ ((uint32_t)r & 7u)) << 4);
}
// byte offset of element (r, c)
__device__ __forceinline__ uint32_t pt_off(int r, int c) { return pt_off4(r, c >> 2) + ((uint32_t)c & 3u) * 4u; }

// K-major SWIZZLE_128B descriptor of a tile image, K-step ks (0..15): chunk block ks/4, 32 bytes per step
// inside the 128-byte line (LBO is unused for swizzled K-major operands and set to 16 B, SBO = 8 rows = 1024 B)
__device__ __forceinline__ uint64_t pt_desc(uint32_t img_saddr) {
    return tc_desc(img_saddr, 16u, 1024u) | ((uint64_t)2 << 61);           // layout_type 2 = SWIZZLE_128B
}
#define PT_KSTEP(ks) ((uint64_t)((((uint32_t)(ks) >> 2) * PT_CB + ((uint32_t)(ks) & 3u) * 32u) >> 4))

// ---- TMA ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)tm),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)tm),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
// L2 prefetch of a box (no shared-memory destination, no completion tracking): issued a few tiles ahead of the ring it
// turns the HBM latency of the later tensor load into an L2 hit
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"((uint64_t)tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}
// plain bulk copy global -> shared (bytes % 16 == 0, 16-byte aligned both sides), completes on the mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 16-byte asynchronous copy global -> shared with a per-thread source (the neighbour gather), and the arrive that
// fires when all earlier cp.async of the executing thread have landed (the mbarrier count includes it: .noinc)
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- bounded mbarrier wait ------------------------------------------------------------------------------
// A protocol error in a warp-specialised kernel is a hang, and a hung GPU is not recoverable from the host
// side of this library.  Every wait of the pipelined kernels therefore gives up after ~1 s: it raises the
// CTA-wide `dead` flag (all later waits of the CTA return at once), sets SCANN_ERR_PIPE_TIMEOUT in the engine's
// status word and records where it happened; the kernel then runs to completion with garbage results and the host
// raises on the status bit (Engine.check_status).
// status[1..4] <- {wait site code, CTA, tile, stage} of the first wait that gave up (the engine's status buffer holds
// 8 ints; word 0 is the flag word every kernel shares).
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
static __device__ __noinline__ void pipe_give_up(volatile int* dead, int32_t* status, int code, int tile, int stage) {
    *dead = 1;
    if (status) atomicOr(status, SCANN_ERR_PIPE_TIMEOUT);
    if (status && atomicCAS(status + 1, 0, code) == 0) { status[2] = (int)blockIdx.x; status[3] = tile; status[4] = stage; }
    __threadfence();
}
__device__ __forceinline__ void pipe_wait(uint64_t* bar, uint32_t parity, volatile int* dead, int32_t* status, int code,
                                          int tile, int stage) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (*dead) return;
        if (clock64() - t0 > 2000000000LL) { pipe_give_up(dead, status, code, tile, stage); return; }
    }
}

// ---- weights -> tensor memory, 3xTF32 issue, accumulator read-back ------------------------------------------
// W[k][n] (row-major, ld 128) -> tensor memory as A[M = n][K = k]: the raw fp32 block (the tensor core truncates it
// to tf32 = the "hi" part) at t_raw, lo = w - trunc(w) at t_lo.  nwarps consumer warps (multiple of 4): warp w
// covers lane quarter w % 4 and the k blocks (w / 4), (w / 4) + nwarps / 4, ...
__device__ __forceinline__ void pipe_weight_to_tmem(const float* __restrict__ W, uint32_t t_raw, uint32_t t_lo, int warp,
                                                    int lane, int nwarps) {
    const int n = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int kb = warp >> 2; kb < 4; kb += nwarps >> 2) {
        const int kbase = kb * 32;
        float w[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) w[q] = __ldg(W + (size_t)(kbase + q) * SCANN_D + n);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float hi[16], lo[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) tf32_split(w[g * 16 + q], hi[q], lo[q]);
            tmem_st16(t_raw + lane_base + kbase + g * 16, hi);
            tmem_st16(t_lo + lane_base + kbase + g * 16, lo);
        }
    }
    tmem_st_wait();
}

// D^T[feature][row] = W^T X^T with X = x_raw + x_lo held as two tile images: one accumulator, the two correction
// products first (tc_probe: as accurate as a separate correction accumulator).  One elected thread.
__device__ __forceinline__ void pipe_issue_3xtf32(uint32_t t_wraw, uint32_t t_wlo, uint32_t img_raw, uint32_t img_lo,
                                                  uint32_t t_acc, uint64_t* bar) {
    const uint32_t idesc = tc_idesc_tf32(128, PT, false, false);
    const uint64_t dr = pt_desc(img_raw), dl = pt_desc(img_lo);
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_acc, t_wlo + ks * 8, dr + PT_KSTEP(ks), idesc, ks != 0);
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_acc, t_wraw + ks * 8, dl + PT_KSTEP(ks), idesc, true);
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_acc, t_wraw + ks * 8, dr + PT_KSTEP(ks), idesc, true);
    tc_commit(bar);
}

// accumulator (lane = feature n, column = row r) -> image S[r][n] (+ bias[n]); warp quarter q = warp % 4
__device__ __forceinline__ void pipe_acc_to_image(uint32_t t_acc, uint8_t* S, const float* __restrict__ bias, int q,
                                                  int lane) {
    const int n = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float b = bias ? __ldg(bias + n) : 0.f;
    float m0[16], m1[16];
    tmem_ld16(t_acc + lane_base, m0);
    tmem_ld16(t_acc + lane_base + 16, m1);
    tmem_ld_wait();
#pragma unroll
    for (int r = 0; r < 16; ++r) *reinterpret_cast<float*>(S + pt_off(r, n)) = m0[r] + b;
#pragma unroll
    for (int r = 0; r < 16; ++r) *reinterpret_cast<float*>(S + pt_off(16 + r, n)) = m1[r] + b;
}

// ---- host: tensor maps -------------------------------------------------------------------------------------
// [rows x 128] fp32 row-major tensor, box = [32 columns x PT rows], 128-byte swizzle.  Encoded per call on the
// host (a pure CPU function, ~1 us) and passed by value inside the __grid_constant__ kernel argument.
int pipe_encode_tmap(CUtensorMap* tm, const float* base, long long rows);
