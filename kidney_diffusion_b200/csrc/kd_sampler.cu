// K6 / K7: the DDPM p_sample update with exact dynamic thresholding, RePaint blend / re-noise, final clamp + paste,
// counter-based Gaussian noise, and the overlap-border pack of the patch-grid sampler.  All NCHW fp32, HBM-bound,
// 128-bit vectorised where alignment allows.  Arithmetic mirrors the reference expression order with explicit
// round-to-nearest intrinsics (no FMA contraction) so that, given identical inputs, results match the fp32 oracle.
#include "kd_common.cuh"

namespace {

__device__ __forceinline__ float x0_from_pred(float x, float pred, int objective, float alpha, float sigma) {
  if (objective == KD_PRED_V) return __fsub_rn(__fmul_rn(alpha, x), __fmul_rn(sigma, pred));       // alpha * x_t - sigma * v
  if (objective == KD_PRED_NOISE) return __fdiv_rn(__fsub_rn(x, __fmul_rn(sigma, pred)), fmaxf(alpha, 1e-8f));
  return pred;
}

// ------------------------------------------------------------------------------------------------ K7: radix select
// State per (b, which): {prefix (high bits decided so far), remaining rank}.  4 passes of 8 bits over the fp32 bit
// pattern of |x0| (non-negative floats order like unsigned integers).  Two order statistics (ranks lo and hi) are
// selected in the same passes.
struct SelState {
  uint32_t prefix[2];
  uint32_t rank[2];
};

__global__ void sel_init_kernel(SelState* st, uint32_t* hist, int B, uint32_t rank_lo, uint32_t rank_hi) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    st[i].prefix[0] = st[i].prefix[1] = 0u;
    st[i].rank[0] = rank_lo;
    st[i].rank[1] = rank_hi;
  }
  if (i < B * 512) hist[i] = 0u;
}

__global__ void sel_hist_kernel(const float* __restrict__ x_t, const float* __restrict__ pred, long n_per, int objective,
                                float alpha, float sigma, const SelState* __restrict__ st, uint32_t* __restrict__ hist, int pass) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  __shared__ uint32_t sh[512];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const uint32_t p0 = st[b].prefix[0], p1 = st[b].prefix[1];
  const uint32_t hmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
  const float* xb = x_t + (long)b * n_per;
  const float* pb = pred + (long)b * n_per;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per; i += stride) {
    const float x0 = x0_from_pred(xb[i], pb[i], objective, alpha, sigma);
    const uint32_t key = __float_as_uint(fabsf(x0));
    const uint32_t digit = (key >> shift) & 0xFFu;
    if ((key & hmask) == p0) atomicAdd(&sh[digit], 1u);
    if ((key & hmask) == p1) atomicAdd(&sh[256 + digit], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[(long)b * 512 + i], sh[i]);
}

__global__ void sel_scan_kernel(SelState* st, uint32_t* hist, int pass) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int b = blockIdx.x;
  const int which = threadIdx.x;  // 2 threads
  if (which >= 2) return;
  const int shift = 24 - 8 * pass;
  uint32_t* h = hist + (long)b * 512 + which * 256;
  uint32_t rank = st[b].rank[which];
  uint32_t cum = 0;
  int d = 0;
  for (; d < 256; ++d) {
    const uint32_t c = h[d];
    if (rank < cum + c) break;
    cum += c;
  }
  if (d > 255) d = 255;
  st[b].prefix[which] |= ((uint32_t)d) << shift;
  st[b].rank[which] = rank - cum;
  for (int i = 0; i < 256; ++i) h[i] = 0u;  // ready for the next pass
}

__global__ void sel_final_kernel(const SelState* st, int B, float weight, float* s_out) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float lo = __uint_as_float(st[b].prefix[0]);
  const float hi = __uint_as_float(st[b].prefix[1]);
  // torch.lerp(lo, hi, weight)
  const float diff = __fsub_rn(hi, lo);
  float q = (weight < 0.5f) ? __fadd_rn(lo, __fmul_rn(weight, diff)) : __fsub_rn(hi, __fmul_rn(diff, __fsub_rn(1.0f, weight)));
  s_out[b] = fmaxf(q, 1.0f);  // s.clamp_(min = 1.)
}

// ------------------------------------------------------------------------------------------------ K6: p_sample update
struct StepArgs {
  int objective;
  float alpha, sigma, one_minus_c, c, alpha_next, std;
  float rn_k1, rn_num, rn_alpha;
};

__device__ __forceinline__ float step_one(float x, float pr, float nz, float s, const StepArgs& a, float* x0_clamped) {
  float x0 = x0_from_pred(x, pr, a.objective, a.alpha, a.sigma);
  x0 = __fdiv_rn(fminf(fmaxf(x0, -s), s), s);  // x_start.clamp(-s, s) / s   (s = 1 for static thresholding)
  *x0_clamped = x0;
  const float t1 = __fdiv_rn(__fmul_rn(x, a.one_minus_c), a.alpha);             // x_t * (1 - c) / alpha
  const float mean = __fmul_rn(a.alpha_next, __fadd_rn(t1, __fmul_rn(a.c, x0)));  // alpha_next * (... + c * x_start)
  return __fadd_rn(mean, __fmul_rn(a.std, nz));                                   // + nonzero * exp(0.5 logvar) * noise
}

__global__ void ddpm_step_kernel(const float* __restrict__ x_t, const float* __restrict__ pred, const float* __restrict__ noise,
                                 const float* __restrict__ s_dev, float* __restrict__ out, float* __restrict__ x0_out,
                                 const float* __restrict__ renoise, long n_per, StepArgs a) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int b = blockIdx.y;
  const float s = s_dev ? s_dev[b] : 1.0f;
  const long base = (long)b * n_per;
  const long nv = n_per >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 xv = *reinterpret_cast<const float4*>(x_t + base + i * 4);
    const float4 pv = *reinterpret_cast<const float4*>(pred + base + i * 4);
    const float4 nz = *reinterpret_cast<const float4*>(noise + base + i * 4);
    float4 x0v, o;
    o.x = step_one(xv.x, pv.x, nz.x, s, a, &x0v.x);
    o.y = step_one(xv.y, pv.y, nz.y, s, a, &x0v.y);
    o.z = step_one(xv.z, pv.z, nz.z, s, a, &x0v.z);
    o.w = step_one(xv.w, pv.w, nz.w, s, a, &x0v.w);
    if (renoise != nullptr) {
      // q_sample_from_to: x * (alpha_to / alpha) + noise * (sigma_to * alpha - sigma * alpha_to) / alpha
      const float4 rz = *reinterpret_cast<const float4*>(renoise + base + i * 4);
      o.x = __fadd_rn(__fmul_rn(o.x, a.rn_k1), __fdiv_rn(__fmul_rn(rz.x, a.rn_num), a.rn_alpha));
      o.y = __fadd_rn(__fmul_rn(o.y, a.rn_k1), __fdiv_rn(__fmul_rn(rz.y, a.rn_num), a.rn_alpha));
      o.z = __fadd_rn(__fmul_rn(o.z, a.rn_k1), __fdiv_rn(__fmul_rn(rz.z, a.rn_num), a.rn_alpha));
      o.w = __fadd_rn(__fmul_rn(o.w, a.rn_k1), __fdiv_rn(__fmul_rn(rz.w, a.rn_num), a.rn_alpha));
    }
    *reinterpret_cast<float4*>(out + base + i * 4) = o;
    if (x0_out) *reinterpret_cast<float4*>(x0_out + base + i * 4) = x0v;
  }
}

__global__ void inpaint_blend_kernel(float* __restrict__ img, const float* __restrict__ inpaint, const uint8_t* __restrict__ mask,
                                     const float* __restrict__ noise, float alpha, float sigma, int C, long HW) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int b = blockIdx.y;
  const long n = (long)C * HW;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long hw = i % HW;
    if (mask[(long)b * HW + hw]) {
      const long g = (long)b * n + i;
      float v = inpaint[g];
      if (noise) v = __fadd_rn(__fmul_rn(alpha, v), __fmul_rn(sigma, noise[g]));  // q_sample(inpaint, t)
      img[g] = v;
    }
  }
}

__global__ void finalize_image_kernel(float* __restrict__ img, const float* __restrict__ inpaint, const uint8_t* __restrict__ mask,
                                      int C, long HW) {
  const int b = blockIdx.y;
  const long n = (long)C * HW;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long g = (long)b * n + i;
    float v = fminf(fmaxf(img[g], -1.0f), 1.0f);
    if (inpaint && mask[(long)b * HW + (i % HW)]) v = inpaint[g];
    img[g] = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);  // unnormalize_zero_to_one
  }
}

__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, float alpha, float sigma,
                                float* __restrict__ out, long n) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __fadd_rn(__fmul_rn(alpha, x0[i]), __fmul_rn(sigma, noise[i]));
}

// ------------------------------------------------------------------------------------------------ Philox4x32-10 + Box-Muller
__device__ __forceinline__ void philox_round(uint32_t* c, uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__global__ void randn_kernel(float* __restrict__ out, long n, uint64_t seed, uint64_t key) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += stride) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)key, (uint32_t)(key >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = ((float)(c[2 * h] >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
      const float u2 = (float)(c[2 * h + 1] >> 8) * (1.0f / 16777216.0f);       // [0, 1)
      const float rr = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincosf(6.283185307179586f * u2, &sn, &cs);
      z[2 * h] = rr * cs;
      z[2 * h + 1] = rr * sn;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i * 4 + j < n) out[i * 4 + j] = z[j];
  }
}

// ------------------------------------------------------------------------------------------------ K8: border pack
// sample_ultra_res.py:149-170 -- write order above -> side -> corner; the corner sets no mask bits of its own.
// Each neighbour is given as a STRIP view: element (c, y, x) of the strip = ptr[c * cs + y * rs + x], where
//   above  strip = the neighbour's bottom `ov` rows            [3, ov, S]
//   side   strip = the neighbour's `ov` columns facing us      [3, S, ov]
//   corner strip = the diagonal neighbour's facing ov x ov box [3, ov, ov]
// so the same kernel reads either a full resident patch (cs = S*S, rs = S, pointer offset to the strip origin) or a
// contiguous strip received from another rank over NVLink.
struct Strip {
  const float* p;
  long cs, rs;
};

__global__ void border_pack_kernel(float* __restrict__ inpaint, uint8_t* __restrict__ mask, Strip above, Strip side, Strip corner,
                                   int S, int ov, int orientation) {
  const long n = (long)S * S;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int y = (int)(i / S), x = (int)(i % S);
    const bool in_top = y < ov;
    const bool in_side = orientation == -1 ? (x < ov) : (x >= S - ov);
    const int sx = orientation == -1 ? x : (x - (S - ov));  // column inside the side / corner strip
    uint8_t m = 0;
    float v[3] = {0.f, 0.f, 0.f};
    if (above.p && in_top) {
      m = 1;
      for (int c = 0; c < 3; ++c) v[c] = above.p[c * above.cs + (long)y * above.rs + x];
    }
    if (side.p && in_side) {
      m = 1;
      for (int c = 0; c < 3; ++c) v[c] = side.p[c * side.cs + (long)y * side.rs + sx];
    }
    if (corner.p && in_top && in_side)
      for (int c = 0; c < 3; ++c) v[c] = corner.p[c * corner.cs + (long)y * corner.rs + sx];
    for (int c = 0; c < 3; ++c) inpaint[(long)c * n + i] = v[c];
    mask[i] = m;
  }
}

unsigned ew_blocks(long n) {
  long b = (n + 255) / 256;
  const long cap = (long)kd_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

extern "C" size_t kd_dynthresh_workspace_bytes(int B) { return (size_t)B * (512 * sizeof(uint32_t) + sizeof(SelState)); }

extern "C" int kd_dynthresh(const float* x_t, const float* pred, int B, long n_per, int objective, float alpha, float sigma,
                            long rank_lo, long rank_hi, float weight, void* workspace, size_t ws_bytes, float* s_out,
                            kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x_t && pred && workspace && s_out && B > 0 && n_per > 0, "kd_dynthresh: bad argument");
  KD_REQUIRE(ws_bytes >= kd_dynthresh_workspace_bytes(B), "kd_dynthresh: workspace too small");
  KD_REQUIRE(rank_lo >= 0 && rank_hi >= rank_lo && rank_hi < n_per && n_per < 4294967296L, "kd_dynthresh: bad ranks");
  uint32_t* hist = reinterpret_cast<uint32_t*>(workspace);
  SelState* st = reinterpret_cast<SelState*>(hist + (size_t)B * 512);
  KD_CUDA(kd_launch(sel_init_kernel, dim3(kd_ceil_div((long)B * 512, 256)), dim3(256), 0, stream, st, hist, B, (uint32_t)rank_lo, (uint32_t)rank_hi));
  long blocks = (n_per + 256L * 8 - 1) / (256L * 8);
  const long cap = (long)kd_num_sms() * 8 / B + 1;
  if (blocks > cap) blocks = cap;
  for (int pass = 0; pass < 4; ++pass) {
    KD_CUDA(kd_launch(sel_hist_kernel, dim3((unsigned)blocks, B), dim3(256), 0, stream, x_t, pred, n_per, objective, alpha, sigma, st, hist, pass));
    KD_CUDA(kd_launch(sel_scan_kernel, dim3(B), dim3(32), 0, stream, st, hist, pass));
  }
  KD_CUDA(kd_launch(sel_final_kernel, dim3(kd_ceil_div(B, 64)), dim3(64), 0, stream, st, B, weight, s_out));
  return KD_OK;
}

extern "C" int kd_ddpm_step(const float* x_t, const float* pred, const float* noise, const float* s, float* out, float* x0_out, int B,
                            long n_per, int objective, float alpha, float sigma, float one_minus_c, float c, float alpha_next,
                            float std, const float* renoise, float rn_k1, float rn_num, float rn_alpha, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x_t && pred && noise && out && B > 0 && n_per > 0, "kd_ddpm_step: bad argument");
  KD_REQUIRE(n_per % 4 == 0, "kd_ddpm_step: elements per sample (%ld) must be a multiple of 4", n_per);
  KD_REQUIRE(objective >= 0 && objective <= 2, "kd_ddpm_step: bad objective");
  StepArgs a{objective, alpha, sigma, one_minus_c, c, alpha_next, std, rn_k1, rn_num, rn_alpha};
  long blocks = (n_per / 4 + 255) / 256;
  const long cap = (long)kd_num_sms() * 8 / B + 1;
  if (blocks > cap) blocks = cap;
  KD_CUDA(kd_launch(ddpm_step_kernel, dim3((unsigned)blocks, B), dim3(256), 0, stream, x_t, pred, noise, s, out, x0_out, renoise, n_per, a));
  return KD_OK;
}

extern "C" int kd_inpaint_blend(float* img, const float* inpaint, const uint8_t* mask, const float* noise, float alpha, float sigma,
                                int B, int C, long HW, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(img && inpaint && mask && B > 0 && C > 0 && HW > 0, "kd_inpaint_blend: bad argument");
  KD_CUDA(kd_launch(inpaint_blend_kernel, dim3(ew_blocks((long)C * HW / 2), B), dim3(256), 0, stream, img, inpaint, mask, noise, alpha, sigma, C, HW));
  return KD_OK;
}

extern "C" int kd_finalize_image(float* img, const float* inpaint, const uint8_t* mask, int B, int C, long HW, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(img && B > 0 && C > 0 && HW > 0 && (!inpaint || mask), "kd_finalize_image: bad argument");
  finalize_image_kernel<<<dim3(ew_blocks((long)C * HW / 2), B), 256, 0, stream>>>(img, inpaint, mask, C, HW);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_q_sample(const float* x0, const float* noise, float alpha, float sigma, float* out, long n, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x0 && noise && out && n > 0, "kd_q_sample: bad argument");
  q_sample_kernel<<<ew_blocks(n / 2), 256, 0, stream>>>(x0, noise, alpha, sigma, out, n);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_randn(float* out, long n, uint64_t seed, uint64_t key, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(out && n > 0, "kd_randn: bad argument");
  KD_CUDA(kd_launch(randn_kernel, dim3(ew_blocks(n / 4 + 1)), dim3(256), 0, stream, out, n, seed, key));
  return KD_OK;
}

extern "C" int kd_border_pack(float* inpaint, uint8_t* mask, const float* above, long above_cs, long above_rs, const float* side,
                              long side_cs, long side_rs, const float* corner, long corner_cs, long corner_rs, int S, int overlap_pos,
                              int orientation, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(inpaint && mask && S > 0 && overlap_pos > 0 && overlap_pos <= S && (orientation == 1 || orientation == -1),
             "kd_border_pack: bad argument");
  Strip a{above, above_cs, above_rs}, sd{side, side_cs, side_rs}, c{corner, corner_cs, corner_rs};
  border_pack_kernel<<<ew_blocks((long)S * S), 256, 0, stream>>>(inpaint, mask, a, sd, c, S, overlap_pos, orientation);
  KD_LAUNCH_CHECK();
  return KD_OK;
}
