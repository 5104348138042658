// Per-atom Dense layers on the 5th-generation tensor cores (tcgen05, kind::tf32, 3xTF32).
//
//   C[r, nb*128 + c] = epi( sum_kb A_kb[r,:] @ W[kb*nblk+nb] + bias[nb] (+ resid) )
//
// Same contract as scann_dense_forward (atom.cu).  One CTA = 128 rows x one 128-column block:
//   * the weight block is the STATIONARY operand: W^T (hi and lo tf32 parts) lives in tensor
//     memory as the M x K "A" operand (lane = output feature n, column = input feature k),
//   * the activation tile is the "B" operand: its canonical K-major image (tc_common.cuh) is staged
//     in shared memory by the threads (coalesced loads, hi/lo split, no transposition needed),
//   * D^T = W^T @ X^T accumulates in tensor memory (lane = n, column = row r), the main term
//     hi*hi and the correction lo*hi + hi*lo in separate accumulators (single-accumulator
//     3xTF32 loses the small terms to the tensor core's truncating adds: 1.5e-5 vs 3.4e-6),
//   * epilogue: TMEM -> registers -> shared (transposing) -> warp-per-row bias / swish /
//     residual / LayerNorm -> coalesced global stores.
#include "common.cuh"
#include "tc_common.cuh"

#define DTC_THREADS 256
extern "C" int scann_device_sm_count(void);

// phase timestamps of CTA (0,0) (clock64), read back with scann_debug_clocks_dense: development aid
#ifdef SCANN_DEV_PROBES
__device__ long long g_dbg_clk_dense[16];
#define DCLK(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_dbg_clk_dense[i] = clock64(); } while (0)
#else
#define DCLK(i) do { } while (0)
#endif

struct DenseTcArgs {
    const float* A[3];
    int lda;
    const float* W[9];
    const float* bias[3];
    int kblk, nblk, R;
    float* C;
    int ldc, mode;
    const float* resid;
    int ldres;
    const float* pre_in;
    float* pre_out;
    const float* gamma;
    const float* beta;
};

// weight block W[k][n] -> tensor memory as A[M = n][K = k] (hi and lo parts): thread = output feature n,
// warps 0-3 take k in [0,64), warps 4-7 take k in [64,128)
__device__ __forceinline__ void dense_weight_to_tmem(const float* __restrict__ W, uint32_t t_whi, uint32_t t_wlo, int warp,
                                                     int lane) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int n = (warp & 3) * 32 + lane, kbase = (warp >> 2) * 64;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float w[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) w[q] = __ldg(W + (size_t)(kbase + h * 32 + q) * SCANN_D + n);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float hi[16], lo[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) tf32_split(w[g * 16 + q], hi[q], lo[q]);
            tmem_st16(t_whi + lane_base + kbase + h * 32 + g * 16, hi);
            tmem_st16(t_wlo + lane_base + kbase + h * 32 + g * 16, lo);
        }
    }
}

// TR = activation rows per CTA (32 / 64 / 128 = the N extent of the MMA).  The per-atom tensors of one batch are
// only a few thousand rows, so the host picks the smallest TR that still fits one wave of CTAs: the kernel's
// latency (staging, MMA, epilogue) scales with TR while the weight staging is hidden behind the predecessor.
template <int TR>
__global__ void __launch_bounds__(DTC_THREADS, 1) dense_tc_kernel(const DenseTcArgs a) {
    constexpr int XIT = TR / 8;                       // LDG.128 per thread for the activation tile
    constexpr uint32_t IMG = (TR / 8) * TC_RG_STRIDE; // bytes of one K-major image of TR rows
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sXhi = smem;
    uint8_t* sXlo = smem + IMG;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * TR, nb = blockIdx.y;
    DCLK(0);
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t t_whi = tmem, t_wlo = tmem + 128, t_dm = tmem + 256, t_dc = tmem + 256 + TR;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc = tc_idesc_tf32(128, TR, false, false);
    uint32_t phase = 0;
    DCLK(1);
    // the first weight block only depends on parameters: stage it before waiting for the predecessor kernel
    dense_weight_to_tmem(a.W[nb], t_whi, t_wlo, warp, lane);
    pdl_wait();

    for (int kb = 0; kb < a.kblk; ++kb) {
        // (b) activation tile: issue all global loads first (TR/8 x LDG.128 in flight per thread)
        const float* A = a.A[kb];
        float4 xv[XIT];
#pragma unroll
        for (int it = 0; it < XIT; ++it) {
            const int i = tid + it * DTC_THREADS, r = i >> 5, c4 = i & 31;
            xv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < a.R) xv[it] = ld4(A + (size_t)(r0 + r) * a.lda + c4 * 4);
        }
        // (a) weight block -> tensor memory (block 0 was staged before the dependency wait)
        if (kb > 0) dense_weight_to_tmem(a.W[kb * a.nblk + nb], t_whi, t_wlo, warp, lane);
        // activation tile -> K-major images (hi, lo)
#pragma unroll
        for (int it = 0; it < XIT; ++it) {
            const int i = tid + it * DTC_THREADS, r = i >> 5, c4 = i & 31;
            float4 v = xv[it], h, l;
            tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y);
            tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
            const uint32_t off = tc_off4(r, c4);
            *reinterpret_cast<float4*>(sXhi + off) = h;
            *reinterpret_cast<float4*>(sXlo + off) = l;
        }
        tmem_st_wait();
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (kb == 0) DCLK(2);
        if (warp == 0 && tc_elect_one()) {
            tc_fence_after();
            const uint64_t dh = tc_desc_kmajor(smem_u32(sXhi), 0), dl = tc_desc_kmajor(smem_u32(sXlo), 0);
            const bool first = kb == 0;
#pragma unroll
            for (int ks = 0; ks < 16; ++ks)
                tc_mma_ts(t_dm, t_whi + ks * 8, dh + ks * TC_KSTEP_DESC, idesc, !(first && ks == 0));
#pragma unroll
            for (int ks = 0; ks < 16; ++ks)
                tc_mma_ts(t_dc, t_wlo + ks * 8, dh + ks * TC_KSTEP_DESC, idesc, !(first && ks == 0));
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) tc_mma_ts(t_dc, t_whi + ks * 8, dl + ks * TC_KSTEP_DESC, idesc, true);
            tc_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
        __syncthreads();
        if (kb == 0) DCLK(3);
    }
    pdl_trigger();          // only the epilogue is left: the next kernel may start its own prologue
    // epilogue 1: D^T (lane = n, column = r) -> shared image S[r][n]  (S aliases the hi image)
    {
        const int n = (warp & 3) * 32 + lane, rbase = (warp >> 2) * (TR / 2);
#pragma unroll 1
        for (int rr = rbase; rr < rbase + TR / 2; rr += 16) {
            float m[16], c[16];
            tmem_ld16(t_dm + lane_base + rr, m);
            tmem_ld16(t_dc + lane_base + rr, c);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 16; ++q) *reinterpret_cast<float*>(sXhi + tc_off(rr + q, n)) = m[q] + c[q];
        }
    }
    tc_fence_before();
    __syncthreads();
    DCLK(4);
    if (warp == 0) tmem_dealloc(tmem, 512);
    DCLK(6);
    // epilogue 2: row groups (4 rows per warp step, 8 lanes per row, 16 columns per lane)
    const int l8 = lane & 7, rsub = lane >> 3;
    float4 bias[4], gam[4], bet[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int c0 = (l8 + 8 * it) * 4;
        bias[it] = a.bias[nb] ? ldg4(a.bias[nb] + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
        gam[it] = a.mode == 3 ? ldg4(a.gamma + c0) : bias[it];
        bet[it] = a.mode == 3 ? ldg4(a.beta + c0) : bias[it];
    }
#pragma unroll 1
    for (int step = 0; step < TR / (DTC_THREADS / 32) / 4; ++step) {
        const int rr = warp * (TR / (DTC_THREADS / 32)) + step * 4 + rsub, r = r0 + rr;
        const bool ok = r < a.R;
        float v[4][4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int c4 = l8 + 8 * it, c0 = c4 * 4;
            float4 acc = *reinterpret_cast<const float4*>(sXhi + tc_off4(rr, c4));
            float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.resid && ok) rv = ld4(a.resid + (size_t)r * a.ldres + nb * SCANN_D + c0);
            v[it][0] = acc.x + bias[it].x + rv.x; v[it][1] = acc.y + bias[it].y + rv.y;
            v[it][2] = acc.z + bias[it].z + rv.z; v[it][3] = acc.w + bias[it].w + rv.w;
        }
        if (a.mode == 1) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                if (a.pre_out && ok)
                    st4(a.pre_out + (size_t)r * a.ldc + nb * SCANN_D + (l8 + 8 * it) * 4,
                        make_float4(v[it][0], v[it][1], v[it][2], v[it][3]));
#pragma unroll
                for (int q = 0; q < 4; ++q) v[it][q] = swish_fast(v[it][q]);
            }
        } else if (a.mode == 2) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) p = ld4(a.pre_in + (size_t)r * a.ldc + nb * SCANN_D + (l8 + 8 * it) * 4);
                v[it][0] *= swish_grad_fast(p.x); v[it][1] *= swish_grad_fast(p.y);
                v[it][2] *= swish_grad_fast(p.z); v[it][3] *= swish_grad_fast(p.w);
            }
        } else if (a.mode == 3) {
            float s1 = 0.f;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                if (a.pre_out && ok)
                    st4(a.pre_out + (size_t)r * a.ldc + (l8 + 8 * it) * 4, make_float4(v[it][0], v[it][1], v[it][2], v[it][3]));
                s1 += v[it][0] + v[it][1] + v[it][2] + v[it][3];
            }
            // shifted one-pass moments: shift = mean of the row's first 16-column slice
            const float sh = __shfl_sync(0xffffffffu, s1, lane & 24) * (1.0f / 16.0f);
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int it = 0; it < 4; ++it)
#pragma unroll
                for (int q = 0; q < 4; ++q) { v[it][q] -= sh; m1 += v[it][q]; m2 = fmaf(v[it][q], v[it][q], m2); }
            oct_sum2(m1, m2);
            m1 *= (1.0f / SCANN_D);
            const float inv = rsqrtf(fmaxf(m2 * (1.0f / SCANN_D) - m1 * m1, 0.f) + SCANN_LN_EPS);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                v[it][0] = (v[it][0] - m1) * inv * gam[it].x + bet[it].x;
                v[it][1] = (v[it][1] - m1) * inv * gam[it].y + bet[it].y;
                v[it][2] = (v[it][2] - m1) * inv * gam[it].z + bet[it].z;
                v[it][3] = (v[it][3] - m1) * inv * gam[it].w + bet[it].w;
            }
        }
        if (ok) {
#pragma unroll
            for (int it = 0; it < 4; ++it)
                st4(a.C + (size_t)r * a.ldc + nb * SCANN_D + (l8 + 8 * it) * 4,
                    make_float4(v[it][0], v[it][1], v[it][2], v[it][3]));
        }
    }
    DCLK(5);
}

#ifdef SCANN_DEV_PROBES
extern "C" int scann_debug_clocks_dense(long long* host_out16) {
    cudaError_t e = cudaMemcpyFromSymbol(host_out16, g_dbg_clk_dense, sizeof(long long) * 16);
    if (e != cudaSuccess) { scann_set_error("debug_clocks_dense: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}
#endif

extern "C" int scann_dense_forward_tc(const float* const* A, int lda, const float* const* W, const float* const* bias,
                                      int kblk, int nblk, int R, float* C, int ldc, int mode, const float* resid,
                                      int ldres, const float* pre_in, float* pre_out, const float* gamma,
                                      const float* beta, void* stream) {
    if (kblk < 1 || kblk > 3 || nblk < 1 || nblk > 3) { scann_set_error("dense_tc: kblk/nblk must be in 1..3"); return 1; }
    if (mode == 3 && nblk != 1) { scann_set_error("dense_tc: LayerNorm epilogue needs nblk == 1"); return 1; }
    if (R <= 0) return 0;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(2 * TC_TILE_BYTES));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(dense_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_TILE_BYTES);
        if (e != cudaSuccess) { scann_set_error("dense_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    // smallest row tile whose grid still fits one wave of SMs (one CTA per SM: each allocates all of tensor memory)
    static int sms = 0;
    if (sms == 0) sms = scann_device_sm_count();
    int tr = 128;
    if (((R + 31) / 32) * nblk <= sms) tr = 32;
    else if (((R + 63) / 64) * nblk <= sms) tr = 64;
    DenseTcArgs a;
    for (int i = 0; i < 3; ++i) { a.A[i] = i < kblk ? A[i] : nullptr; a.bias[i] = (bias && i < nblk) ? bias[i] : nullptr; }
    for (int i = 0; i < 9; ++i) a.W[i] = i < kblk * nblk ? W[i] : nullptr;
    a.lda = lda; a.kblk = kblk; a.nblk = nblk; a.R = R; a.C = C; a.ldc = ldc; a.mode = mode;
    a.resid = resid; a.ldres = ldres; a.pre_in = pre_in; a.pre_out = pre_out; a.gamma = gamma; a.beta = beta;
    dim3 grid((R + tr - 1) / tr, nblk);
    const size_t smem = 2 * (size_t)(tr / 8) * TC_RG_STRIDE;
    if (tr == 32) scann_launch(dense_tc_kernel<32>, grid, dim3(DTC_THREADS), smem, stream, a);
    else if (tr == 64) scann_launch(dense_tc_kernel<64>, grid, dim3(DTC_THREADS), smem, stream, a);
    else scann_launch(dense_tc_kernel<128>, grid, dim3(DTC_THREADS), smem, stream, a);
    return scann_check_launch("scann_dense_forward_tc");
}
