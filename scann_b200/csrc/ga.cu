// Global attention + property head, one CTA per structure.
//
// Restates GlobalAttention.call (scann/layers/attention.py:267-318) for v_proj=False,
// kq_proj=True and the head Dense layers (scann/models/scann_model.py:437-447).
// The reference builds the [M,M] energy matrix, zeroes its diagonal and sums over the
// query axis (:279-292); with m in {0,1} that equals
//     s_i = m_i * k_i . (Q - m_i q_i),   Q = sum_j m_j q_j
// so no M x M matrix is formed (O(M*D) work).  ga = softmax_i(s/||s|| - 1e9 (1-m)) is the
// ga_score output; ctx = sum_i m_i ga_i k_i feeds bf_property -> predict_property.
#include "common.cuh"

#define GA_THREADS 128

__device__ __forceinline__ float block_sum_128(float v, float* s_red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    return s_red[0] + s_red[1] + s_red[2] + s_red[3];
}
__device__ __forceinline__ float block_max_128(float v, float* s_red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    return fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
}

// Shared forward part: fills s_Q[128], s_s[M] (normalised scores), s_ga[M]; returns ||s|| (or 1).
__device__ float ga_scores(const float* __restrict__ qk, const uint8_t* __restrict__ mask, int M, int norm,
                           float* s_Q, float* s_s, float* s_ga, float* s_red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // loads are unconditional and selected afterwards (rows of padded atoms exist and are finite or not, the
    // select drops them): a load guarded by the mask byte would make every iteration a dependent round trip
    float Q = 0.f;
#pragma unroll 8
    for (int i = 0; i < M; ++i) {
        const float q = qk[(size_t)i * 2 * SCANN_D + tid];
        Q += mask[i] ? q : 0.f;
    }
    s_Q[tid] = Q;
    __syncthreads();
    const float4 Q4 = ld4(s_Q + lane * 4);
#pragma unroll 4
    for (int i = warp; i < M; i += GA_THREADS / 32) {
        const float4 q = ld4(qk + (size_t)i * 2 * SCANN_D + lane * 4);
        const float4 k = ld4(qk + (size_t)i * 2 * SCANN_D + SCANN_D + lane * 4);
        float v = k.x * (Q4.x - q.x) + k.y * (Q4.y - q.y) + k.z * (Q4.z - q.z) + k.w * (Q4.w - q.w);
        v = warp_sum(mask[i] ? v : 0.f);
        if (lane == 0) s_s[i] = v;
    }
    __syncthreads();
    float nrm = 1.0f;
    if (norm) {
        float p = 0.f;
        for (int i = tid; i < M; i += GA_THREADS) p += s_s[i] * s_s[i];
        nrm = sqrtf(block_sum_128(p, s_red));
        for (int i = tid; i < M; i += GA_THREADS) s_s[i] = s_s[i] / nrm;   // 0/0 -> NaN as tf.linalg.normalize
        __syncthreads();
    }
    float mx = -INFINITY;
    for (int i = tid; i < M; i += GA_THREADS) {
        float l = s_s[i] + (mask[i] ? 0.f : -1e9f);
        s_ga[i] = l;
        mx = fmaxf(mx, l);
    }
    mx = block_max_128(mx, s_red);
    float se = 0.f;
    for (int i = tid; i < M; i += GA_THREADS) {
        float e = expf(s_ga[i] - mx);
        s_ga[i] = e;
        se += e;
    }
    se = block_sum_128(se, s_red);
    for (int i = tid; i < M; i += GA_THREADS) s_ga[i] = s_ga[i] / se;
    __syncthreads();
    return nrm;
}

__global__ void __launch_bounds__(GA_THREADS) ga_head_fwd_kernel(const float* __restrict__ qk,
                                                                 const uint8_t* __restrict__ atom_mask, int M,
                                                                 int norm, const float* __restrict__ Wb,
                                                                 const float* __restrict__ bb,
                                                                 const float* __restrict__ wp,
                                                                 const float* __restrict__ bp, int mrelu,
                                                                 float* __restrict__ ga, float* __restrict__ y,
                                                                 float* __restrict__ ctx_out,
                                                                 float* __restrict__ tb_out) {
    extern __shared__ __align__(16) float sm[];
    float* s_Q = sm;                 // 128
    float* s_ctx = s_Q + SCANN_D;    // 128
    float* s_red = s_ctx + SCANN_D;  // 4
    float* s_s = s_red + 4;          // M
    float* s_ga = s_s + M;           // M
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* qkb = qk + (size_t)b * M * 2 * SCANN_D;
    const uint8_t* mb = atom_mask + (size_t)b * M;
    pdl_wait();
    pdl_trigger();      // only after the own wait: at most one kernel ahead becomes resident early
    ga_scores(qkb, mb, M, norm, s_Q, s_s, s_ga, s_red);
    for (int i = tid; i < M; i += GA_THREADS) ga[(size_t)b * M + i] = s_ga[i];
    float c = 0.f;
#pragma unroll 8
    for (int i = 0; i < M; ++i) {
        const float k = qkb[(size_t)i * 2 * SCANN_D + SCANN_D + tid];
        c = fmaf(s_ga[i], mb[i] ? k : 0.f, c);
    }
    s_ctx[tid] = c;
    if (ctx_out) ctx_out[(size_t)b * SCANN_D + tid] = c;
    __syncthreads();
    float t = bb[tid];
#pragma unroll 16
    for (int d = 0; d < SCANN_D; ++d) t = fmaf(s_ctx[d], __ldg(Wb + (size_t)d * SCANN_D + tid), t);
    if (tb_out) tb_out[(size_t)b * SCANN_D + tid] = t;
    float yy = block_sum_128(swish_f(t) * wp[tid], s_red) + bp[0];
    if (mrelu) yy = fmaxf(yy, 0.f);      // mrelu forward (custom_layers.py:15)
    if (tid == 0) y[b] = yy;
}

// Backward of head + global attention for one structure.
// dy[b] is the upstream gradient of y[b] (mrelu has an identity gradient, custom_layers.py:12-13).
__global__ void __launch_bounds__(GA_THREADS) ga_head_bwd_kernel(
    const float* __restrict__ qk, const uint8_t* __restrict__ atom_mask, int M, int norm,
    const float* __restrict__ WbT, const float* __restrict__ wp, const float* __restrict__ tb,
    const float* __restrict__ dy, float* __restrict__ d_qk, float* __restrict__ d_tb, float* __restrict__ dwp,
    float* __restrict__ dbp) {
    extern __shared__ __align__(16) float sm[];
    float* s_Q = sm;                  // 128
    float* s_dctx = s_Q + SCANN_D;    // 128
    float* s_dtb = s_dctx + SCANN_D;  // 128
    float* s_red = s_dtb + SCANN_D;   // 4
    float* s_s = s_red + 4;           // M   normalised scores
    float* s_ga = s_s + M;            // M
    float* s_ds = s_ga + M;           // M   d_ga -> d_t -> d_s
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* qkb = qk + (size_t)b * M * 2 * SCANN_D;
    float* dqkb = d_qk + (size_t)b * M * 2 * SCANN_D;
    const uint8_t* mb = atom_mask + (size_t)b * M;
    pdl_wait();
    pdl_trigger();      // only after the own wait: at most one kernel ahead becomes resident early
    const float g = dy[b];
    // head
    {
        float t = tb[(size_t)b * SCANN_D + tid];
        float hid = swish_f(t);
        atomicAdd(dwp + tid, g * hid);
        if (tid == 0) atomicAdd(dbp, g);
        float dt = g * wp[tid] * swish_grad_f(t);
        s_dtb[tid] = dt;
        d_tb[(size_t)b * SCANN_D + tid] = dt;
    }
    __syncthreads();
    {
        float dc = 0.f;
#pragma unroll 16
        for (int n = 0; n < SCANN_D; ++n) dc = fmaf(s_dtb[n], __ldg(WbT + (size_t)n * SCANN_D + tid), dc);
        s_dctx[tid] = dc;
    }
    const float nrm = ga_scores(qkb, mb, M, norm, s_Q, s_s, s_ga, s_red);   // syncs inside, s_dctx visible after
    // d_ga_i = m_i <d_ctx, k_i>
    {
        const float4 dc4 = ld4(s_dctx + lane * 4);
#pragma unroll 4
        for (int i = warp; i < M; i += GA_THREADS / 32) {
            const float4 k = ld4(qkb + (size_t)i * 2 * SCANN_D + SCANN_D + lane * 4);
            float v = k.x * dc4.x + k.y * dc4.y + k.z * dc4.z + k.w * dc4.w;
            v = warp_sum(mb[i] ? v : 0.f);
            if (lane == 0) s_ds[i] = v;
        }
    }
    __syncthreads();
    float p = 0.f;
    for (int i = tid; i < M; i += GA_THREADS) p = fmaf(s_ga[i], s_ds[i], p);
    const float dot = block_sum_128(p, s_red);
    for (int i = tid; i < M; i += GA_THREADS) s_ds[i] = s_ga[i] * (s_ds[i] - dot);     // d_t
    __syncthreads();
    if (norm) {
        float q = 0.f;
        for (int i = tid; i < M; i += GA_THREADS) q = fmaf(s_s[i], s_ds[i], q);
        const float sd = block_sum_128(q, s_red);
        for (int i = tid; i < M; i += GA_THREADS) s_ds[i] = (s_ds[i] - s_s[i] * sd) / nrm;
        __syncthreads();
    }
    for (int i = tid; i < M; i += GA_THREADS)
        if (!mb[i]) s_ds[i] = 0.f;
    __syncthreads();
    // d_Q[d] = sum_i d_s_i m_i k_i[d]
    const float Q = s_Q[tid], dctx = s_dctx[tid];
    float dQ = 0.f;
#pragma unroll 8
    for (int i = 0; i < M; ++i) {
        const float k = qkb[(size_t)i * 2 * SCANN_D + SCANN_D + tid];
        dQ = fmaf(s_ds[i], mb[i] ? k : 0.f, dQ);            // s_ds is 0 for masked atoms, k may be anything
    }
#pragma unroll 4
    for (int i = 0; i < M; ++i) {
        const float q = qkb[(size_t)i * 2 * SCANN_D + tid], k = qkb[(size_t)i * 2 * SCANN_D + SCANN_D + tid];
        float dq = 0.f, dk = 0.f;
        if (mb[i]) {
            dk = s_ga[i] * dctx + s_ds[i] * (Q - q);
            dq = dQ - s_ds[i] * k;
        }
        dqkb[(size_t)i * 2 * SCANN_D + tid] = dq;
        dqkb[(size_t)i * 2 * SCANN_D + SCANN_D + tid] = dk;
    }
}

// err_b = y_b - t_b ; dy_b = err_b (the 1/(B*RMSE) factor of d sqrt(mean err^2) is applied in the
// optimiser AFTER the gradient all-reduce, so data-parallel ranks need one collective only);
// sse[0] += sum err^2, abs_err[0] += sum |err|.
__global__ void __launch_bounds__(256) rmse_prepare_kernel(const float* __restrict__ y, const float* __restrict__ target,
                                                           int B, float* __restrict__ dy, float* __restrict__ sse) {
    __shared__ float s_a[8], s_b[8];
    pdl_wait();
    pdl_trigger();      // only after the own wait: at most one kernel ahead becomes resident early
    float a = 0.f, c = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        float e = y[i] - target[i];
        dy[i] = e;
        a = fmaf(e, e, a);
        c += fabsf(e);
    }
    a = warp_sum(a);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ta = 0.f, tc = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += s_a[w]; tc += s_b[w]; }
        atomicAdd(sse, ta);
        atomicAdd(sse + 1, tc);
    }
}

extern "C" int scann_ga_head_forward(const float* qk, const uint8_t* atom_mask, int B, int M, int norm,
                                     const float* Wb, const float* bb, const float* wp, const float* bp, int mrelu,
                                     float* ga, float* y, float* ctx_out, float* tb_out, void* stream) {
    if (B <= 0) return 0;
    size_t smem = (size_t)(2 * SCANN_D + 4 + 2 * M) * sizeof(float);
    if (smem > 48 * 1024) { scann_set_error("ga_head_forward: M=%d too large", M); return 1; }
    scann_launch(ga_head_fwd_kernel, dim3(B), dim3(GA_THREADS), smem, stream, qk, atom_mask, M, norm, Wb, bb, wp, bp, mrelu,
                 ga, y, ctx_out, tb_out);
    return scann_check_launch("scann_ga_head_forward");
}

extern "C" int scann_ga_head_backward(const float* qk, const uint8_t* atom_mask, int B, int M, int norm,
                                      const float* WbT, const float* wp, const float* tb, const float* dy,
                                      float* d_qk, float* d_tb, float* dwp, float* dbp, void* stream) {
    if (B <= 0) return 0;
    size_t smem = (size_t)(3 * SCANN_D + 4 + 3 * M) * sizeof(float);
    if (smem > 48 * 1024) { scann_set_error("ga_head_backward: M=%d too large", M); return 1; }
    scann_launch(ga_head_bwd_kernel, dim3(B), dim3(GA_THREADS), smem, stream, qk, atom_mask, M, norm, WbT, wp, tb, dy, d_qk,
                 d_tb, dwp, dbp);
    return scann_check_launch("scann_ga_head_backward");
}

extern "C" int scann_rmse_prepare(const float* y, const float* target, int B, float* dy, float* sse, void* stream) {
    if (B <= 0) return 0;
    int grid = (B + 255) / 256;
    if (grid > 64) grid = 64;
    scann_launch(rmse_prepare_kernel, dim3(grid), dim3(256), 0, stream, y, target, B, dy, sse);
    return scann_check_launch("scann_rmse_prepare");
}
