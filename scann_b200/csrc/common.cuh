// Shared device helpers for the SCANN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// numerics experiment (SCANN_NVCC_DEFS=-DSCANN_ACCURATE_MATH): IEEE expf / division instead of ex2.approx / rcp.approx
// in every epilogue -- used once to attribute the full-depth error budget, not a product mode
#ifdef SCANN_ACCURATE_MATH
#define __expf expf
#define __fdividef(a, b) ((a) / (b))
#endif

#define SCANN_D 128          // local_dim = global_dim = dense_out (every shipped config)
#define SCANN_H 8            // attention heads
#define SCANN_HD 16          // head dim
#define SCANN_RBF 20         // Gaussian centres
#define SCANN_TILE 128       // pair rows per tile (= UMMA M)
#define SCANN_LN_EPS 1e-6f   // LayerNormalization(epsilon=1e-6), attention.py:35,111,113

// error plumbing (abi.cu)
void scann_set_error(const char* fmt, ...);
int scann_check_launch(const char* what);

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// The train / inference step is a chain of ~100 dependent kernels that each run for a few microseconds,
// so launch latency and per-kernel prologues (TMEM allocation, weights -> tensor memory) are a large part
// of the step.  Kernels on that chain are launched with programmatic stream serialisation: every CTA
// executes pdl_wait() before it touches anything a predecessor wrote (waits for the whole predecessor grid
// and its memory) and pdl_trigger() once only its last epilogue is left (the next kernel of the stream may
// then become resident on free SMs and run its own prologue).  RULES: (1) code before pdl_wait() may only
// read model parameters (params / transposed params), which are always complete behind a full
// (non-programmatic) dependency; (2) every CTA calls pdl_wait() on every path, so completion of a kernel
// implies completion of all its predecessors; (3) pdl_trigger() comes after pdl_wait(), so at most ONE
// kernel ahead is resident early (triggering first let whole chains of waiting CTAs pile up on the SMs and
// starve the side-stream weight-gradient kernels: measured 1.97 vs 1.83 ms per train step).
// Both instructions are no-ops in a kernel that was launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// host side: scann_set_pdl(1) (abi.cu, per thread) makes scann_launch attach the attribute
bool scann_pdl_enabled();
// which local-attention kernels run as four warp groups per CTA when the plan allows it (scann_set_la_groups4,
// abi.cu, per thread): bit 0 geometry forward, 1 attention forward, 2 attention backward, 3 geometry backward
int scann_la_tc4_mask();
// development switch (bit 5 of scann_set_la_groups4): build the pair plan with the four separate kernels
bool scann_plan_unfused();
template <typename... KArgs, typename... Args>
static inline void scann_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = scann_pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// device-side status word shared by all kernels of one engine (bit flags)
#define SCANN_ERR_TILE_OVERFLOW 1   // plan needed more tiles than the caller allocated
#define SCANN_ERR_TOO_MANY_NBRS 2   // an atom has more than SCANN_TILE valid neighbours
#define SCANN_ERR_BAD_ATOMIC 4      // atomic number outside the embedding table
#define SCANN_ERR_BAD_NEIGHBOR 8    // neighbour index outside [0, M)

// ---- training-mode Dropout (keras.layers.Dropout: scann_model.py:374 rate 0.1, attention.py:29 rate 0.1) -----
// Counter-based: the keep decision of element `idx` of dropout site `site` is a hash of (seed, site, idx), so the
// forward and backward kernels regenerate the same mask without storing it, and a test can rebuild it on the
// host (scann_b200/dropout.py).  Kept elements are scaled by 1/(1-rate) (inverted dropout, as Keras does).
// The control block lives in device memory (refreshed by the host before every step, so that a captured CUDA
// graph sees a new seed on every replay): {seed, threshold = rate * 2^32, float bits of 1/(1-rate), enabled}.
struct ScannDropCtl { uint32_t seed, threshold, scale_bits, enabled; };
__device__ __forceinline__ uint32_t drop_hash(uint32_t seed, uint32_t site, uint32_t idx) {
    uint32_t x = idx * 0x9E3779B1u ^ (seed + site * 0x85EBCA6Bu);
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
// multiplier of element idx: 0 or 1/(1-rate); 1 when ctl is NULL or disabled
__device__ __forceinline__ float drop_mult(const ScannDropCtl* ctl, uint32_t site, uint32_t idx) {
    if (!ctl) return 1.0f;
    const ScannDropCtl c = *ctl;
    if (!c.enabled) return 1.0f;
    return drop_hash(c.seed, site, idx) >= c.threshold ? __uint_as_float(c.scale_bits) : 0.0f;
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swish_f(float x) { return x * sigmoid_f(x); }
// d/dx [x * sigmoid(x)] = s * (1 + x * (1 - s))
__device__ __forceinline__ float swish_grad_f(float x) {
    float s = sigmoid_f(x);
    return s * (1.0f + x * (1.0f - s));
}

// Fast variants for the tensor-core kernels, whose row-wise epilogues are instruction-issue bound:
// ex2.approx + rcp.approx (2 ulp each) instead of the ~25-instruction expf / IEEE division.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float swish_fast(float x) { return x * sigmoid_fast(x); }
__device__ __forceinline__ float swish_grad_fast(float x) {
    float s = sigmoid_fast(x);
    return s * (1.0f + x * (1.0f - s));
}
// sum of two values over the warp in one shuffle sequence
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
}

// Row-group mapping of the tensor-core epilogues: a warp step covers 4 rows, 8 lanes per row, lane l of a
// row owns the four 16-byte chunks c4 = l + 8*it (it = 0..3), i.e. 16 of the 128 columns.  Per-row fixed
// costs (index loads, reductions, addressing) are amortised over 16 elements per lane instead of 4; global
// accesses stay fully coalesced (8 lanes x 16 B = one 128-byte line per row and instruction).
__device__ __forceinline__ void oct_sum2(float& a, float& b) {      // sums over the 8 lanes of a row
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// reduce over the 16 lanes that share (lane & 16)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// reduce over groups of 4 consecutive lanes
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// vectorised fire-and-forget reduction into global memory (sm_90+)
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
