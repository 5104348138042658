// tcgen05 / TMEM / mbarrier building blocks (sm_100a inline PTX).
//
// Operand tiles live in shared memory in ONE canonical image, used for every role.
// A [128 x 128] fp32 tile X[r][c] is stored as 8x4 "core matrices" (8 rows x 16 bytes, the
// no-swizzle / INTERLEAVE canonical form of the UMMA descriptors): core matrix (rg = r/8, cg = c/4)
// is 128 contiguous bytes, row r%8 at +16*(r%8); core matrices of one row group are TC_CG_STRIDE
// apart, row groups TC_RG_STRIDE apart.  The same bytes serve as
//   * K-major  operand (rows = M or N, reduction along c): LBO = TC_CG_STRIDE, SBO = TC_RG_STRIDE
// (MN-major tf32 operands, needed by the weight-gradient GEMMs X^T @ Y, need their own image, below).
// TC_CG_STRIDE = 144 (not 128) keeps both access patterns of the CUDA-core code conflict-free:
// thread-per-row (8 lanes x 16 B contiguous) and lane-per-column-chunk (stride 144 B = 36 banks).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_CG_STRIDE 144u
#define TC_RG_STRIDE (32u * TC_CG_STRIDE)          // 4608
#define TC_TILE_BYTES (16u * TC_RG_STRIDE)         // 73728 bytes per [128 x 128 fp32] tile image

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of the 16-byte chunk holding columns [4*c4, 4*c4+4) of row r
__device__ __forceinline__ uint32_t tc_off4(int r, int c4) {
    return (uint32_t)(r >> 3) * TC_RG_STRIDE + (uint32_t)c4 * TC_CG_STRIDE + ((uint32_t)r & 7u) * 16u;
}
// byte offset of element (r, c)
__device__ __forceinline__ uint32_t tc_off(int r, int c) { return tc_off4(r, c >> 2) + ((uint32_t)c & 3u) * 4u; }

// tf32 split: hi keeps the top 19 bits (what the tensor core consumes whether it truncates or
// rounds), lo = x - hi exactly (fp32); hi*hi + lo*hi + hi*lo reproduces fp32 products to ~2^-21.
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// ---- shared-memory matrix descriptors (cute::UMMA::SmemDescriptor, version 1 = Blackwell) ----
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version; layout_type (bits 61..63) = 0: no swizzle
    return d;
}
// K-major view of a tile image, K-step ks (8 tf32 = two core matrices along the columns)
__device__ __forceinline__ uint64_t tc_desc_kmajor(uint32_t tile_saddr, int ks) {
    return tc_desc(tile_saddr + (uint32_t)ks * 2u * TC_CG_STRIDE, TC_CG_STRIDE, TC_RG_STRIDE);
}
// descriptor of K-step ks = descriptor of K-step 0 + ks * TC_KSTEP_DESC (start-address field, 16-byte units)
#define TC_KSTEP_DESC ((uint64_t)((2u * TC_CG_STRIDE) >> 4))

// ---- MN-major tf32 operands: the only layout the hardware accepts is SWIZZLE_128B_BASE32B
// (cutlass sm100_common.inl:92): column blocks of [128 rows x 128 B], row pitch 128 B, 32-byte
// chunks XOR-swizzled with (r % 4) (Swizzle<2,5,2>), K atom = 4 rows.
#define TC_MN_BLOCK_BYTES 16384u
#define TC_MN_TILE_BYTES 65536u
__device__ __forceinline__ uint32_t tc_mn_off(int r, int c) {
    return (uint32_t)(c >> 5) * TC_MN_BLOCK_BYTES + (uint32_t)r * 128u +
           (((((uint32_t)c >> 3) & 3u) ^ ((uint32_t)r & 3u)) << 5) + ((uint32_t)c & 7u) * 4u;
}
// K-step ks = rows [8 ks, 8 ks + 8) = two 4-row atoms
__device__ __forceinline__ uint64_t tc_desc_mn32(uint32_t tile_saddr, int ks) {
    uint64_t d = tc_desc(tile_saddr + (uint32_t)ks * 1024u, TC_MN_BLOCK_BYTES, 512u);
    return d | ((uint64_t)1 << 61);    // layout_type 1 = SWIZZLE_128B_BASE32B
}

// instruction descriptor, kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t tc_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// One elected lane of a CONVERGED warp (elect.sync).  tcgen05.mma / commit must be issued from such a lane inside
// a warp-uniform branch: under a divergent `if (tid == 0)` the compiler wraps every MMA in an
// ELECT / BRA.U.ANY loop (one MMA per ~50-60 cycles whatever its size -- measured), whereas with elect.sync it
// issues them back to back from the uniform datapath.
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- tcgen05 wrappers ----
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accum)
        : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accum)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp gets row (lane base + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- warp groups of the local-attention kernels (la_tc.cu, la_tc_bwd.cu) ----
// The 16 warps of a CTA split into NG independent groups; each group owns its own tile stream, operand images,
// mbarrier and accumulator columns and synchronises with a named barrier (see the header of la_tc.cu).
#define LTC_GROUP_THREADS 512
#define LTC_GROUP_WARPS 16
template <int NG>
struct LaGroups {
    static constexpr int TR = 128 / NG;                     // pair rows per tile (MMA N extent)
    static constexpr int WG = LTC_GROUP_WARPS / NG;                      // warps per group
    static constexpr int GT = LTC_GROUP_THREADS / NG;                    // threads per group
    static constexpr uint32_t IMG = (uint32_t)(TR / 8) * TC_RG_STRIDE;   // bytes of one K-major image of TR rows
};
// producer / consumer halves of a named barrier (ids 5..7 hand the start-up skew from group g to group g+1)
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// barrier among the GT threads of warp group g (barrier 0 stays __syncthreads)
__device__ __forceinline__ void group_sync(int g, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nthreads) : "memory");
}

