// Self-test of the tcgen05 building blocks in tc_common.cuh: one 128x128x128 tile product on the
// 5th-generation tensor cores, operands staged by the threads in the canonical SW128 tile image.
//   layout 0: D = A @ W      (A K-major, B = K-major image of W^T)           forward GEMMs
//   layout 1: D = A^T @ W    (A, B MN-major, SWIZZLE_128B_BASE32B images)    weight-gradient GEMMs
//   layout 2: D = A @ W^T    (A K-major, B = K-major image of W)             input-gradient GEMMs
//   layout 3: D = A @ W with A read from tensor memory (tcgen05.mma [d], [a], b)
//   nprod 1: single TF32 product of the raw fp32 operands; nprod 3: hi/lo split, 3 products into one
//   accumulator; nprod 4: the same 3 products, the two cross terms in a second accumulator that is
//   added on the CUDA cores (measures what the tensor core's accumulation costs in accuracy).
#ifdef SCANN_DEV_PROBES      // development builds only: SCANN_NVCC_DEFS=-DSCANN_DEV_PROBES python -m scann_b200.build --force
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

// part: 0 raw, 1 hi, 2 lo ; img: 0 K-major image of src, 1 K-major image of src^T, 2 MN-major image of src
__device__ __forceinline__ void stage_tile(uint8_t* dst, const float* __restrict__ src, int part, int img, int tid) {
    for (int i = tid; i < 128 * 32; i += 128) {
        int r = i >> 5, c4 = i & 31;
        float4 v = ld4(src + r * 128 + c4 * 4);
        float x[4] = {v.x, v.y, v.z, v.w};
        if (part) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float hi, lo;
                tf32_split(x[q], hi, lo);
                x[q] = part == 1 ? hi : lo;
            }
        }
        if (img == 0) {
            *reinterpret_cast<float4*>(dst + tc_off4(r, c4)) = make_float4(x[0], x[1], x[2], x[3]);
        } else if (img == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<float*>(dst + tc_off(c4 * 4 + q, r)) = x[q];
        } else {
            *reinterpret_cast<float4*>(dst + tc_mn_off(r, c4 * 4)) = make_float4(x[0], x[1], x[2], x[3]);
        }
    }
}

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                          float* __restrict__ D, int layout, int nprod) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sW = smem + TC_TILE_BYTES;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t d_tmem = tmem, a_tmem = tmem + 128, d2_tmem = tmem + 256;
    const int np = nprod >= 4 ? 3 : nprod;       // nprod 5: one accumulator, the two correction products FIRST
    const bool a_mn = layout == 1, b_mn = layout == 1;
    const uint32_t idesc = tc_idesc_tf32(128, 128, a_mn, b_mn);
    uint32_t phase = 0;
    for (int p = 0; p < np; ++p) {
        int pa = nprod == 1 ? 0 : (p == 1 ? 2 : 1);     // hi, lo, hi
        int pw = nprod == 1 ? 0 : (p == 2 ? 2 : 1);     // hi, hi, lo
        if (nprod == 5) { pa = p == 0 ? 2 : 1; pw = p == 1 ? 2 : 1; }      // (lo,hi), (hi,lo), (hi,hi)
        if (layout == 3) {
            // A -> tensor memory, thread = row, 128 fp32 columns
            const int row = tid;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 16) {
                float v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    float x = A[row * 128 + c0 + q];
                    if (pa) { float hi, lo; tf32_split(x, hi, lo); x = pa == 1 ? hi : lo; }
                    v[q] = x;
                }
                tmem_st16(a_tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            }
            tmem_st_wait();
        } else {
            stage_tile(sA, A, pa, layout == 1 ? 2 : 0, tid);
        }
        stage_tile(sW, W, pw, layout == 1 ? 2 : (layout == 2 ? 0 : 1), tid);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_s = smem_u32(sA), w_s = smem_u32(sW);
            for (int ks = 0; ks < 16; ++ks) {
                uint64_t bd = b_mn ? tc_desc_mn32(w_s, ks) : tc_desc_kmajor(w_s, ks);
                bool acc = (p | ks) != 0;
                uint32_t dt = d_tmem;
                if (nprod == 4 && p > 0) { dt = d2_tmem; acc = !(p == 1 && ks == 0); }
                if (layout == 3) {
                    tc_mma_ts(dt, a_tmem + ks * 8, bd, idesc, acc);
                } else {
                    uint64_t ad = a_mn ? tc_desc_mn32(a_s, ks) : tc_desc_kmajor(a_s, ks);
                    tc_mma_ss(dt, ad, bd, idesc, acc);
                }
            }
            tc_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
        __syncthreads();
    }
    {
        const int row = tid;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 16) {
            float v[16];
            tmem_ld16(d_tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            if (nprod == 4) {
                float v2[16];
                tmem_ld16(d2_tmem + ((uint32_t)(warp * 32) << 16) + c0, v2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] += v2[q];
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) D[row * 128 + c0 + q] = v[q];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// Diagnostic entry point: D[128,128] = op(A)[128,128] @ op(W)[128,128] on the tensor cores.
extern "C" int scann_tc_probe(const float* A, const float* W, float* D, int layout, int nprod, void* stream) {
    if (layout < 0 || layout > 3 || (nprod != 1 && nprod != 3 && nprod != 4 && nprod != 5)) { scann_set_error("tc_probe: bad arguments"); return 1; }
    size_t smem = 2 * TC_TILE_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { scann_set_error("tc_probe: %s", cudaGetErrorString(e)); return 1; }
    tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, W, D, layout, nprod);
    return scann_check_launch("scann_tc_probe");
}

// ---- timing probe: cycles per tcgen05.mma for a chain of `nmma` accumulating MMAs ----
// mode 0: SS, B K-major image with TC_CG_STRIDE; mode 1: TS (A in TMEM); mode 2: SS with cg stride 128;
// mode 3: TS with cg stride 128 ; N = n_cols (multiple of 16)
__global__ void __launch_bounds__(128, 1) tc_time_kernel(float* out, int mode, int nmma, int ncols) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (2 * (int)TC_TILE_BYTES) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.5f;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 0 && ((mode & 8) ? tc_elect_one() : tid == 0)) {      // mode & 8: elect.sync issue (uniform datapath)
        const uint32_t cg = (mode & 2) ? 128u : TC_CG_STRIDE;
        const uint32_t rg = 32u * cg;
        const uint32_t idesc = tc_idesc_tf32(128, ncols, false, false);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + TC_TILE_BYTES);
        long long t0 = clock64();
        if (mode & 16) {          // fast issue, round-robin over nacc = mode >> 8 independent accumulators of ncols columns
            const int nacc = mode >> 8;
            const uint64_t bd0 = tc_desc(sb, cg, rg);
            const uint64_t step = (uint64_t)((2u * cg) >> 4);
            auto run = [&](auto NA) {
                constexpr int NACC = decltype(NA)::value;
                for (int i = 0; i < nmma; i += 16 * NACC) {
#pragma unroll
                    for (int ks = 0; ks < 16; ++ks)
#pragma unroll
                        for (int q = 0; q < NACC; ++q)
                            tc_mma_ts(tmem + 256 + q * ncols, tmem + ks * 8, bd0 + ks * step, idesc, (i | ks) != 0);
                }
            };
            switch (nacc) {
                case 1: run(std::integral_constant<int, 1>{}); break;
                case 2: run(std::integral_constant<int, 2>{}); break;
                case 3: run(std::integral_constant<int, 3>{}); break;
                case 4: run(std::integral_constant<int, 4>{}); break;
                case 6: run(std::integral_constant<int, 6>{}); break;
                default: run(std::integral_constant<int, 8>{}); break;
            }
        } else
        if (mode & 4) {           // fast issue: descriptors advanced by immediates, 16 MMAs per unrolled group
            const uint64_t bd0 = tc_desc(sb, cg, rg), ad0 = tc_desc(sa, cg, rg);
            const uint64_t step = (uint64_t)((2u * cg) >> 4);
            for (int i = 0; i < nmma; i += 16) {
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    if (mode & 1) tc_mma_ts(tmem + 256, tmem + ks * 8, bd0 + ks * step, idesc, (i | ks) != 0);
                    else tc_mma_ss(tmem + 256, ad0 + ks * step, bd0 + ks * step, idesc, (i | ks) != 0);
                }
            }
        } else
        for (int i = 0; i < nmma; ++i) {
            const int ks = i & 15;
            uint64_t bd = tc_desc(sb + (uint32_t)ks * 2u * cg, cg, rg);
            if (mode & 1) tc_mma_ts(tmem + 256, tmem + ks * 8, bd, idesc, i != 0);
            else tc_mma_ss(tmem + 256, tc_desc(sa + (uint32_t)ks * 2u * cg, cg, rg), bd, idesc, i != 0);
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[0] = (float)(t1 - t0) / (float)nmma;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

extern "C" int scann_tc_time(float* out, int mode, int nmma, int ncols, void* stream) {
    size_t smem = 2 * TC_TILE_BYTES;
    cudaError_t e = cudaFuncSetAttribute(tc_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { scann_set_error("tc_time: %s", cudaGetErrorString(e)); return 1; }
    tc_time_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(out, mode, nmma, ncols);
    return scann_check_launch("scann_tc_time");
}
#endif  // SCANN_DEV_PROBES
