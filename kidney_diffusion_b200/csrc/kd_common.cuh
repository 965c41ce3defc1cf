// Shared host/device helpers for libkidney_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/kidney_b200.h"

// ---------------------------------------------------------------- host-side error plumbing
void kd_set_error(const char* fmt, ...);

#define KD_FAIL(code, ...)            \
  do {                                \
    kd_set_error(__VA_ARGS__);        \
    return (code);                    \
  } while (0)

#define KD_REQUIRE(cond, ...)                          \
  do {                                                 \
    if (!(cond)) KD_FAIL(KD_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define KD_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      KD_FAIL(KD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define KD_LAUNCH_CHECK()                                                                \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      KD_FAIL(KD_ERR_LAUNCH, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static inline int kd_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch: the kernel may be scheduled while its stream predecessor is still draining; it must execute
// kd_pdl_wait() before touching global memory (which waits for the predecessor's completion and memory flush).  In a step of
// ~550-730 small launches the per-launch scheduling gap is what this hides (CUDA graphs keep these edges programmatic).
#ifdef __CUDACC__
extern int g_kd_pdl;  // kd_abi.cu; 0 disables (test hook)
template <typename... KArgs, typename... Args>
static inline cudaError_t kd_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_kd_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
int kd_num_sms();

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// 16-bit activation / weight type of the UNet: IEEE fp16 (11-bit significand) with fp32 accumulation everywhere.
// The reference's GroupNorm-everywhere UNet keeps activations O(1..1e3) (imagen-pytorch itself trains it under fp16
// autocast), so fp16's range suffices, and its 8x finer rounding than fp16 is what keeps the per-step UNet output within
// rel-L2 1e-2 of the fp32 reference with margin (fp16 storage measured 6e-3 .. 1.04e-2).  Conversions saturate at
// +-65504 so an outlier can never turn into inf / NaN.
typedef __half h16;
typedef __half2 h162;

struct __align__(16) h16x8 {
  h162 v[4];
};

__device__ __forceinline__ void kd_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void kd_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float sat16(float x) { return fminf(fmaxf(x, -65504.0f), 65504.0f); }

__device__ __forceinline__ void h16x8_to_float(const h16x8& in, float* out) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(in.v[i]);
    out[2 * i] = f.x;
    out[2 * i + 1] = f.y;
  }
}
// {lo = a, hi = b} rounded to nearest-even and saturated to +-65504 in ONE instruction (F2FP.SATFINITE.F16.F32.PACK_AB)
__device__ __forceinline__ uint32_t pack_h16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ h16x8 float_to_h16x8(const float* in) {
  h16x8 o;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t r = pack_h16x2(in[2 * i], in[2 * i + 1]);
    o.v[i] = *reinterpret_cast<const h162*>(&r);
  }
  return o;
}

// fast activations: the library is built without --use_fast_math so that the sampler update and the time embedding keep
// IEEE semantics, but SiLU runs 4.2 G times per step.  Raw ex2.approx.ftz / rcp.approx.ftz (<= 2 ulp) -- __expf and
// __fdividef carry range fix-ups (FSETP + 3 extra FMUL per element) that made the conv epilogue issue-bound.
// Limits are exact: x -> -inf gives x * rcp(inf) = -0, x -> +inf gives x * rcp(1) = x.
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_fast(1.0f + ex2_fast(-1.4426950408889634f * x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
// SiLU from ONE MUFU op: x * sigmoid(x) = h * tanh(h) + h with h = x / 2 (tanh.approx.f32: abs error ~2^-11 on tanh, i.e. an
// error of at most ~|x| * 2.4e-4 on the result -- the size of the fp16 rounding applied to it right after).  Takes h.
__device__ __forceinline__ float silu_from_half_arg(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == KD_ACT_SILU) return silu_f(x);
  if (act == KD_ACT_GELU) return gelu_f(x);
  if (act == KD_ACT_SIGMOID) return sigmoid_f(x);
  return x;
}

// 8 values at a time under ONE uniform branch per activation kind: a per-element apply_act gets if-converted by the
// compiler (all activations computed speculatively, MUFU included, then selected) even when act == NONE
__device__ __forceinline__ void apply_act8(float* v, int act) {
  if (act == KD_ACT_NONE) return;
  if (act == KD_ACT_SILU) {
    // every caller rounds the result to fp16 right after: the one-MUFU tanh form (abs error <= |x| * 2.4e-4, half an fp16 ulp)
    // instead of ex2 + rcp.  The pixel-shuffle upsample convs are bound by their epilogue's issue rate (ncu r02_shuffle: XU pipe
    // 37 %, 2 epilogue warps per scheduler), and the GroupNorm fallback pass now rounds exactly like the conv's fused prologue.
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = silu_from_half_arg(0.5f * v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], act);
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

// streaming 128-bit global access (read-once / write-once data)
__device__ __forceinline__ int4 ld_stream(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
#endif
