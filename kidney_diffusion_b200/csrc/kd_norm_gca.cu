// HBM-bound NHWC h16 kernels around the convolutions: GroupNorm (stats / finalize / apply), GlobalContext pooling,
// gate * h + residual, LayerNorm over channels.  All reductions are fixed-order (no float atomics) so results are
// identical run to run and across GPU counts.
//
// Common thread mapping for [B, HW, C] tensors: a pixel's C channels are split into C/8 "octets" (one 16-byte load);
// a block of T = (256 / oct) * oct threads covers T / oct pixels at a time, every thread keeps the same octet for its
// whole pixel loop, so per-channel constants (gamma, beta, scale/shift, gate) are loaded once per thread.
#include "kd_common.cuh"

namespace {

__host__ __device__ inline int threads_for_oct(int oct) { return oct >= 256 ? 256 : (256 / oct) * oct; }

// ------------------------------------------------------------------------------------------------ GroupNorm stats
__global__ void gn_stats_kernel(const h16* __restrict__ x, long HW, int C, int c_offset, int group_size, int G,
                                float* __restrict__ partial, int nblk) {
  extern __shared__ float sm[];  // [T][2]
  const int oct = C >> 3;
  const int T = blockDim.x;
  const int lanes = T / oct;  // pixels processed in parallel
  const int o = threadIdx.x % oct;
  const int pl = threadIdx.x / oct;
  const int b = blockIdx.y;
  const long per = (HW + nblk - 1) / nblk;
  const long p0 = (long)blockIdx.x * per;
  const long p1 = p0 + per < HW ? p0 + per : HW;
  const h16* xb = x + (long)b * HW * C + (long)o * 8;
  float s = 0.f, ss = 0.f;
  long p = p0 + pl;
  // 4 independent 16-byte loads in flight per thread (the kernel is pure streaming: latency hiding is all that matters)
  for (; p + 3L * lanes < p1; p += 4L * lanes) {
    int4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = ld_stream(xb + (p + (long)u * lanes) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v[8];
      h16x8_to_float(*reinterpret_cast<h16x8*>(&raw[u]), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += v[j];
        ss += v[j] * v[j];
      }
    }
  }
  for (; p < p1; p += lanes) {
    int4 raw = ld_stream(xb + p * C);
    float v[8];
    h16x8_to_float(*reinterpret_cast<h16x8*>(&raw), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s += v[j];
      ss += v[j] * v[j];
    }
  }
  sm[threadIdx.x * 2] = s;
  sm[threadIdx.x * 2 + 1] = ss;
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    float gs = 0.f, gss = 0.f;
    // octets of this source whose global channel falls into group g, all pixel lanes, fixed order
    for (int oo = 0; oo < oct; ++oo) {
      if ((c_offset + oo * 8) / group_size != g) continue;
      for (int l = 0; l < lanes; ++l) {
        gs += sm[(l * oct + oo) * 2];
        gss += sm[(l * oct + oo) * 2 + 1];
      }
    }
    float* out = partial + (((long)b * nblk + blockIdx.x) * G + g) * 2;
    out[0] = gs;
    out[1] = gss;
  }
}

// one warp per (b, group): lanes stride over the per-block partials, then a fixed-order shuffle tree (deterministic)
__global__ void gn_finalize_kernel(const float* __restrict__ pa, int nblk_a, float scale_a, const float* __restrict__ pb,
                                   int nblk_b, float scale_b, int G, double count, float eps, float* __restrict__ mean_rstd) {
  const int b = blockIdx.x;
  const int g = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= G) return;
  double s = 0.0, ss = 0.0;
  for (int k = lane; k < nblk_a; k += 32) {
    const float2 q = *reinterpret_cast<const float2*>(pa + (((long)b * nblk_a + k) * G + g) * 2);
    s += (double)q.x * scale_a;
    ss += (double)q.y * scale_a * scale_a;
  }
  if (pb != nullptr) {
    for (int k = lane; k < nblk_b; k += 32) {
      const float2 q = *reinterpret_cast<const float2*>(pb + (((long)b * nblk_b + k) * G + g) * 2);
      s += (double)q.x * scale_b;
      ss += (double)q.y * scale_b * scale_b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (lane == 0) {
    const double mean = s / count;
    double var = ss / count - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_rstd[((long)b * G + g) * 2] = (float)mean;
    mean_rstd[((long)b * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

__global__ void gn_apply_kernel(const h16* __restrict__ x, h16* __restrict__ y, long HW, int C, int c_offset, int group_size,
                                int G, float src_scale, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ scale_shift, long ss_stride, int Ctot,
                                int act, int nblk) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int oct = C >> 3;
  const int lanes = blockDim.x / oct;
  const int o = threadIdx.x % oct;
  const int pl = threadIdx.x / oct;
  const int b = blockIdx.y;
  const int cg = c_offset + o * 8;  // global channel of this thread's first element
  const int g = cg / group_size;
  const float mean = mean_rstd[((long)b * G + g) * 2];
  const float rstd = mean_rstd[((long)b * G + g) * 2 + 1];
  float A[8], Bc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float ga = gamma[cg + j], be = beta[cg + j];
    float a = rstd * ga, c = be - mean * rstd * ga;
    if (scale_shift != nullptr) {
      const float sc = scale_shift[(long)b * ss_stride + cg + j] + 1.0f;
      const float sh = scale_shift[(long)b * ss_stride + Ctot + cg + j];
      a *= sc;
      c = c * sc + sh;
    }
    A[j] = a * src_scale;
    Bc[j] = c;
  }
  const long per = (HW + nblk - 1) / nblk;
  const long p0 = (long)blockIdx.x * per;
  const long p1 = p0 + per < HW ? p0 + per : HW;
  const h16* xb = x + (long)b * HW * C + (long)o * 8;
  h16* yb = y + (long)b * HW * C + (long)o * 8;
  long p = p0 + pl;
  for (; p + 3L * lanes < p1; p += 4L * lanes) {
    int4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = ld_stream(xb + (p + (long)u * lanes) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v[8];
      h16x8_to_float(*reinterpret_cast<h16x8*>(&raw[u]), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(A[j], v[j], Bc[j]);
      apply_act8(v, act);
      h16x8 o8 = float_to_h16x8(v);
      *reinterpret_cast<int4*>(yb + (p + (long)u * lanes) * C) = *reinterpret_cast<int4*>(&o8);
    }
  }
  for (; p < p1; p += lanes) {
    int4 raw = ld_stream(xb + p * C);
    float v[8];
    h16x8_to_float(*reinterpret_cast<h16x8*>(&raw), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(A[j], v[j], Bc[j]);
    apply_act8(v, act);
    h16x8 o8 = float_to_h16x8(v);
    *reinterpret_cast<int4*>(yb + p * C) = *reinterpret_cast<int4*>(&o8);
  }
}

// ------------------------------------------------------------------------------------------------ GlobalContext
// logits[b, n] = sum_c x[b,n,c] * w[c] + bias : one warp per pixel (two pixels per warp when C = 128).
__global__ void rowdot_kernel(const h16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                              float* __restrict__ out, long rows, int C) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int oct = C >> 3;
  const int sub = (oct < 32 && (oct & (oct - 1)) == 0) ? oct : 32;  // lanes cooperating on one pixel (power of two)
  const int ppw = 32 / sub;             // pixels per warp
  const int lane = threadIdx.x & 31;
  const long warp_global = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const int sl = lane % sub, sp = lane / sub;
  const float bb = bias ? bias[0] : 0.f;
  for (long r0 = warp_global * ppw; r0 < rows; r0 += nwarps * ppw) {
    const long r = r0 + sp;
    float acc = 0.f;
    if (r < rows) {
      for (int o = sl; o < oct; o += sub) {
        int4 raw = ld_stream(x + r * C + o * 8);
        float v[8];
        h16x8_to_float(*reinterpret_cast<h16x8*>(&raw), v);
        const float4 w0 = *reinterpret_cast<const float4*>(w + o * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(w + o * 8 + 4);
        acc += v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w + v[4] * w1.x + v[5] * w1.y + v[6] * w1.z + v[7] * w1.w;
      }
    }
    for (int off = sub >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (sl == 0 && r < rows) out[r] = acc + bb;
  }
}

// Softmax-weighted channel pooling, one block per (pixel chunk, b): partial[b][blk][c] = sum_n exp(l_n - m_blk) x[n,c]
// logits may arrive as n_parts partial dot products per pixel ([n_parts][B*HW], from the conv epilogue): summed in fixed order
__device__ __forceinline__ float gca_logit(const float* __restrict__ lg, long p, int n_parts, long part_stride) {
  float v = lg[p];
  for (int k = 1; k < n_parts; ++k) v += lg[p + k * part_stride];
  return v;
}

__global__ void gca_pool_kernel(const h16* __restrict__ x, const float* __restrict__ logits, int n_parts, long part_stride, long HW,
                                int C, int nblk, float* __restrict__ part, float* __restrict__ ml, int e_cache) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  extern __shared__ float sm[];  // max(T, lanes*C) floats
  __shared__ float s_red[32];
  __shared__ float s_m;
  const int oct = C >> 3;
  const int T = blockDim.x;
  const int lanes = T / oct;
  const int o = threadIdx.x % oct;
  const int pl = threadIdx.x / oct;
  const int b = blockIdx.y;
  const long per = (HW + nblk - 1) / nblk;
  const long p0 = (long)blockIdx.x * per;
  const long p1 = p0 + per < HW ? p0 + per : HW;
  const float* lg = logits + (long)b * HW;
  // The softmax weight of a pixel is the same for all of its C / 8 octet threads: with the chunk's logits cached in shared memory
  // (e_cache floats behind the reduction scratch) each pixel's n_parts partial logits are summed and exponentiated ONCE per block
  // instead of once per octet thread (C = 1024: 128 threads x 16 loads + 1 exp per pixel made the kernel instruction-bound).
  float* s_e = sm + (size_t)T * 8;
  const bool cached = (p1 - p0) <= (long)e_cache;
  // block max of the chunk's logits
  float m = -INFINITY;
  for (long p = p0 + threadIdx.x; p < p1; p += T) {
    const float v = gca_logit(lg, p, n_parts, part_stride);
    if (cached) s_e[p - p0] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = -INFINITY;
    for (int i = 0; i < (T + 31) / 32; ++i) mm = fmaxf(mm, s_red[i]);
    s_m = mm;
  }
  __syncthreads();
  m = s_m;
  if (cached) {
    for (long p = p0 + threadIdx.x; p < p1; p += T) s_e[p - p0] = __expf(s_e[p - p0] - m);
    __syncthreads();
  }
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float l = 0.f;
  const h16* xb = x + (long)b * HW * C + (long)o * 8;
  long p = p0 + pl;
  // four pixels per step, loads first (memory-level parallelism); the last step is predicated instead of falling back to one
  // dependent load per pixel (a chunk of the 512^2 map is 27-28 pixels per lane: three serial round trips were ~30 % of the block's
  // time).  Accumulation order per thread is unchanged.
  for (; p < p1; p += 4L * lanes) {
    int4 raw[4];
    float ev[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long pu = p + (long)u * lanes;
      const bool ok = pu < p1;
      raw[u] = ok ? ld_stream(xb + pu * C) : make_int4(0, 0, 0, 0);
      ev[u] = !ok ? 0.f : (cached ? s_e[pu - p0] : __expf(gca_logit(lg, pu, n_parts, part_stride) - m));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (p + (long)u * lanes < p1) {
        const float e = ev[u];
        float v[8];
        h16x8_to_float(*reinterpret_cast<h16x8*>(&raw[u]), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(e, v[j], acc[j]);
        l += e;
      }
    }
  }
  // reduce over pixel lanes in fixed order
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[(pl * oct + o) * 8 + j] = acc[j];
  __syncthreads();
  float* outp = part + ((long)b * nblk + blockIdx.x) * C;
  for (int c = threadIdx.x; c < C; c += T) {
    float s = 0.f;
    for (int q = 0; q < lanes; ++q) s += sm[(q * oct + (c >> 3)) * 8 + (c & 7)];
    outp[c] = s;
  }
  __syncthreads();
  // sum of exp: only octet-0 threads hold distinct pixels; fixed-order tree per warp, then the warps in order
  float lw = warp_sum((o == 0) ? l : 0.f);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = lw;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (T + 31) / 32; ++i) s += s_red[i];
    ml[((long)b * nblk + blockIdx.x) * 2] = m;
    ml[((long)b * nblk + blockIdx.x) * 2 + 1] = s;
  }
}

// block = 64 channels x 4 k-slices of one batch element: the per-chunk rescale factors exp(m_k - M) are computed once
// into smem, each thread sums a fixed quarter of the chunks (fixed order), then the 4 slices are added in fixed order
__global__ void gca_finalize_kernel(const float* __restrict__ part, const float* __restrict__ ml, int nblk, int C,
                                    float* __restrict__ pooled) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  extern __shared__ float f[];  // [nblk] + [4][64]
  __shared__ float s_red[8];
  __shared__ float s_M, s_L;
  float* s_part = f + nblk;
  const int b = blockIdx.y;
  const int cl = threadIdx.x & 63, ks = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* mlb = ml + (long)b * nblk * 2;
  float m = -INFINITY;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) m = fmaxf(m, mlb[k * 2]);
  m = warp_max(m);
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = -INFINITY;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mm = fmaxf(mm, s_red[i]);
    s_M = mm;
  }
  __syncthreads();
  const float M = s_M;
  float l = 0.f;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) {
    const float mk = mlb[k * 2];
    const float fk = (mk == -INFINITY) ? 0.f : __expf(mk - M);
    f[k] = fk;
    l += fk * mlb[k * 2 + 1];
  }
  l = warp_sum(l);
  __syncthreads();
  if (lane == 0) s_red[warp] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float ll = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) ll += s_red[i];
    s_L = ll;
  }
  __syncthreads();
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    const float* pp = part + (long)b * nblk * C + c;
    const int per = (nblk + 3) / 4;
    const int k0 = ks * per, k1 = (k0 + per < nblk) ? k0 + per : nblk;
    int k = k0;
    // eight partials in flight per thread (two per step left one exposed L2 round trip per pair: 15 us for 592 chunks); the
    // assignment of chunks to the two accumulators and the order within each are unchanged
    for (; k + 8 <= k1; k += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(pp + (long)(k + u) * C);
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        s0 = fmaf(f[k + u], v[u], s0);
        s1 = fmaf(f[k + u + 1], v[u + 1], s1);
      }
    }
    for (; k + 2 <= k1; k += 2) {
      s0 = fmaf(f[k], pp[(long)k * C], s0);
      s1 = fmaf(f[k + 1], pp[(long)(k + 1) * C], s1);
    }
    for (; k < k1; ++k) s0 = fmaf(f[k], pp[(long)k * C], s0);
  }
  s_part[ks * 64 + cl] = s0 + s1;
  __syncthreads();
  if (ks == 0 && c < C) pooled[(long)b * C + c] = ((s_part[cl] + s_part[64 + cl]) + (s_part[128 + cl] + s_part[192 + cl])) / s_L;
}

__global__ void gate_residual_kernel(const h16* __restrict__ h, const float* __restrict__ gate, const h16* __restrict__ res,
                                     h16* __restrict__ out, float* __restrict__ oct_partial, long HW, int C, int nblk) {
  extern __shared__ float sm[];  // [T][2] when statistics are requested
  kd_pdl_wait();
  kd_pdl_trigger();
  const int oct = C >> 3;
  const int lanes = blockDim.x / oct;
  const int o = threadIdx.x % oct;
  const int pl = threadIdx.x / oct;
  const int b = blockIdx.y;
  float g[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = gate ? gate[(long)b * C + o * 8 + j] : 1.0f;
  const long per = (HW + nblk - 1) / nblk;
  const long p0 = (long)blockIdx.x * per;
  const long p1 = p0 + per < HW ? p0 + per : HW;
  const long base = (long)b * HW * C + (long)o * 8;
  float s1 = 0.f, s2 = 0.f;
  auto body = [&](long p, const int4& raw, const int4& rr) {
    float v[8];
    h16x8_to_float(*reinterpret_cast<const h16x8*>(&raw), v);
    if (res != nullptr) {
      float r[8];
      h16x8_to_float(*reinterpret_cast<const h16x8*>(&rr), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], g[j], r[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= g[j];
    }
    h16x8 o8 = float_to_h16x8(v);
    st_stream(out + base + p * C, *reinterpret_cast<int4*>(&o8));
    if (oct_partial != nullptr) {  // statistics of the rounded values that were stored
      float f[8];
      h16x8_to_float(o8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1 += f[j];
        s2 = fmaf(f[j], f[j], s2);
      }
    }
  };
  long p = p0 + pl;
  // four pixels per step, all loads first; the per-thread accumulation order is unchanged
  for (; p + 3L * lanes < p1; p += 4L * lanes) {
    int4 raw[4], rr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      raw[u] = ld_stream(h + base + (p + (long)u * lanes) * C);
      rr[u] = res != nullptr ? ld_stream(res + base + (p + (long)u * lanes) * C) : make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) body(p + (long)u * lanes, raw[u], rr[u]);
  }
  for (; p < p1; p += lanes) {
    const int4 raw = ld_stream(h + base + p * C);
    const int4 rr = res != nullptr ? ld_stream(res + base + p * C) : make_int4(0, 0, 0, 0);
    body(p, raw, rr);
  }
  if (oct_partial != nullptr) {
    sm[threadIdx.x * 2] = s1;
    sm[threadIdx.x * 2 + 1] = s2;
    __syncthreads();
    if (threadIdx.x < oct) {
      float a1 = 0.f, a2 = 0.f;
      for (int l = 0; l < lanes; ++l) {
        a1 += sm[(l * oct + threadIdx.x) * 2];
        a2 += sm[(l * oct + threadIdx.x) * 2 + 1];
      }
      float2* dst = reinterpret_cast<float2*>(oct_partial) + ((long)b * nblk + blockIdx.x) * oct + threadIdx.x;
      *dst = make_float2(a1, a2);
    }
  }
}

// standalone octet statistics (fallback when the producer kernel could not fuse them)
__global__ void oct_stats_kernel(const h16* __restrict__ x, long HW, int C, float* __restrict__ partial, int nblk) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  extern __shared__ float sm[];
  const int oct = C >> 3;
  const int lanes = blockDim.x / oct;
  const int o = threadIdx.x % oct;
  const int pl = threadIdx.x / oct;
  const int b = blockIdx.y;
  const long per = (HW + nblk - 1) / nblk;
  const long p0 = (long)blockIdx.x * per;
  const long p1 = p0 + per < HW ? p0 + per : HW;
  const h16* xb = x + (long)b * HW * C + (long)o * 8;
  float s = 0.f, ss = 0.f;
  long p = p0 + pl;
  for (; p + 3L * lanes < p1; p += 4L * lanes) {
    int4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = ld_stream(xb + (p + (long)u * lanes) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v[8];
      h16x8_to_float(*reinterpret_cast<h16x8*>(&raw[u]), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += v[j];
        ss = fmaf(v[j], v[j], ss);
      }
    }
  }
  for (; p < p1; p += lanes) {
    int4 raw = ld_stream(xb + p * C);
    float v[8];
    h16x8_to_float(*reinterpret_cast<h16x8*>(&raw), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s += v[j];
      ss = fmaf(v[j], v[j], ss);
    }
  }
  sm[threadIdx.x * 2] = s;
  sm[threadIdx.x * 2 + 1] = ss;
  __syncthreads();
  if (threadIdx.x < oct) {
    float a1 = 0.f, a2 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      a1 += sm[(l * oct + threadIdx.x) * 2];
      a2 += sm[(l * oct + threadIdx.x) * 2 + 1];
    }
    float2* dst = reinterpret_cast<float2*>(partial) + ((long)b * nblk + blockIdx.x) * oct + threadIdx.x;
    *dst = make_float2(a1, a2);
  }
}

// first-level reduction of partial rows: grid (NS splits, B); each block sums a contiguous range of the logical rows of its
// batch image for all octets (thread = fixed octet, `lanes` threads per octet) -> out[b][split][oct]; fixed order throughout
__global__ void oct_reduce_kernel(const float* __restrict__ partial, int rpt, int tiles, int TB, int n_oct, int NS,
                                  float* __restrict__ out) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  extern __shared__ double sred[];  // [T][2]
  const int b = blockIdx.y, sp = blockIdx.x;
  const int o = threadIdx.x % n_oct, l = threadIdx.x / n_oct;
  const int lanes = blockDim.x / n_oct;
  const int tile_b = b / TB, sub = b % TB, rpb = rpt / TB;
  const long count = (long)tiles * rpb;
  const long per = (count + NS - 1) / NS;
  const long r0 = (long)sp * per, r1 = (r0 + per < count) ? r0 + per : count;
  double s1 = 0.0, s2 = 0.0;
  const float2* pp = reinterpret_cast<const float2*>(partial);
  auto row_of = [&](long i) { return ((long)tile_b * tiles + i / rpb) * rpt + (long)sub * rpb + i % rpb; };
  long i = r0 + l;
  for (; i + 7L * lanes < r1; i += 8L * lanes) {
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(pp + row_of(i + (long)u * lanes) * n_oct + o);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s1 += v[u].x;
      s2 += v[u].y;
    }
  }
  for (; i < r1; i += lanes) {
    const float2 v = __ldg(pp + row_of(i) * n_oct + o);
    s1 += v.x;
    s2 += v.y;
  }
  sred[threadIdx.x * 2] = s1;
  sred[threadIdx.x * 2 + 1] = s2;
  __syncthreads();
  if (threadIdx.x < n_oct) {
    double a1 = 0.0, a2 = 0.0;
    for (int k = 0; k < lanes; ++k) {
      a1 += sred[(k * n_oct + threadIdx.x) * 2];
      a2 += sred[(k * n_oct + threadIdx.x) * 2 + 1];
    }
    reinterpret_cast<float2*>(out)[((long)b * NS + sp) * n_oct + threadIdx.x] = make_float2((float)a1, (float)a2);
  }
}

// per-lane share of the {sum, sumsq} of one GroupNorm group from the split sums of up to two concatenated sources: the
// (octet, split) pairs of the group are dealt to the 32 lanes (independent loads in flight instead of a serial loop over up
// to 64 splits), each lane adds its pairs in index order, the caller finishes with the fixed shuffle tree.  Used by both the
// standalone and the single-launch finalize so that they stay bit-identical.
__device__ __forceinline__ void group_octet_sums(int g, int lane, int group_size, const float* sa, int na, int nsa, float scale_a,
                                                 const float* sb, int nb, int nsb, float scale_b, int b, double& s, double& ss) {
  const int o0 = (g * group_size) >> 3, o1 = ((g + 1) * group_size) >> 3;  // octets of the group in concat order
  const int a0 = o0 < na ? o0 : na, a1 = o1 < na ? o1 : na;                // ... that fall into source a
  const int items_a = (a1 - a0) * nsa;
  for (int i = lane; i < items_a; i += 32) {
    const int o = a0 + i / nsa, k = i - (i / nsa) * nsa;
    const float2 v = __ldcg(reinterpret_cast<const float2*>(sa) + ((long)b * nsa + k) * na + o);
    s += (double)v.x * scale_a;
    ss += (double)v.y * scale_a * scale_a;
  }
  if (nb > 0) {
    const int b0 = (o0 > na ? o0 : na) - na, b1 = (o1 > na ? o1 : na) - na;
    const int items_b = (b1 - b0) * nsb;
    for (int i = lane; i < items_b; i += 32) {
      const int o = b0 + i / nsb, k = i - (i / nsb) * nsb;
      const float2 v = __ldcg(reinterpret_cast<const float2*>(sb) + ((long)b * nsb + k) * nb + o);
      s += (double)v.x * scale_b;
      ss += (double)v.y * scale_b * scale_b;
    }
  }
}

// mean / rstd per (b, group) from reduced octet sums of up to two concatenated sources: one warp per group
__global__ void gn_finalize_oct_kernel(const float* __restrict__ sa, int na, int nsa, float scale_a, const float* __restrict__ sb, int nb,
                                       int nsb, float scale_b, int G, int group_size, double count, float eps,
                                       float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ scale_shift, long ss_stride,
                                       float2* __restrict__ coef) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int b = blockIdx.x;
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (g >= G) return;
  double s = 0.0, ss = 0.0;
  group_octet_sums(g, lane, group_size, sa, na, nsa, scale_a, sb, nb, nsb, scale_b, b, s, ss);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    ss += __shfl_xor_sync(0xffffffffu, ss, off);
  }
  const double mean_d = s / count;
  double var = ss / count - mean_d * mean_d;
  if (var < 0.0) var = 0.0;
  const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var + (double)eps));  // identical in every lane
  if (lane == 0) {
    mean_rstd[((long)b * G + g) * 2] = mean;
    mean_rstd[((long)b * G + g) * 2 + 1] = rstd;
  }
  if (coef != nullptr) {
    // per-channel affine of GroupNorm (+ time scale/shift) on the RAW source tensors: y = A * x + B, same arithmetic as
    // gn_apply_kernel; consumed by the convolution's fused pre-activation (kd_conv_gemm_fused pre_coef)
    const int Ctot = (na + nb) * 8, Ca = na * 8;
    for (int c = g * group_size + lane; c < (g + 1) * group_size; c += 32) {
      const float ga = gamma[c], be = beta[c];
      float a = rstd * ga, cc = be - mean * rstd * ga;
      if (scale_shift != nullptr) {
        const float sc = scale_shift[(long)b * ss_stride + c] + 1.0f;
        const float sh = scale_shift[(long)b * ss_stride + Ctot + c];
        a *= sc;
        cc = cc * sc + sh;
      }
      coef[(long)b * Ctot + c] = make_float2(a * (c < Ca ? scale_a : scale_b), cc);
    }
  }
}

// ---- one launch for "reduce the producer's octet partials + finalize": grid (NS_a + NS_b, B) blocks of 256 threads do the
// first-level sums of the two sources; the last block to finish for an image (self-resetting arrival counter) runs the
// finalize.  Same arithmetic, in the same order, as kd_oct_reduce followed by kd_gn_finalize_oct (bit-identical), one launch
// instead of two or three per GroupNorm (104 GroupNorms per 1024^2 step).
struct OctSrc {
  const float* partial;
  int rpt, tiles, TB, n_oct, NS;
  float* out;  // [B][NS][n_oct][2]
  float scale;
};

__global__ void __launch_bounds__(256) gn_reduce_finalize_kernel(const OctSrc a, const OctSrc b2, int G, int group_size, double count,
                                                                 float eps, float* __restrict__ mean_rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ scale_shift, long ss_stride,
                                                                 float2* __restrict__ coef, unsigned int* __restrict__ counter) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  __shared__ double sred[256 * 2];
  __shared__ bool s_last;
  const int b = blockIdx.y;
  {
    const bool first = (int)blockIdx.x < a.NS;
    const OctSrc& s = first ? a : b2;
    const int sp = first ? (int)blockIdx.x : (int)blockIdx.x - a.NS;
    const int n_oct = s.n_oct, lanes = 256 / n_oct;
    const int o = threadIdx.x % n_oct, l = threadIdx.x / n_oct;
    const int tile_b = b / s.TB, sub = b % s.TB, rpb = s.rpt / s.TB;
    const long cnt = (long)s.tiles * rpb;
    const long per = (cnt + s.NS - 1) / s.NS;
    const long r0 = (long)sp * per, r1 = (r0 + per < cnt) ? r0 + per : cnt;
    double s1 = 0.0, s2 = 0.0;
    if (l < lanes) {
      const float2* pp = reinterpret_cast<const float2*>(s.partial);
      auto row_of = [&](long i) { return ((long)tile_b * s.tiles + i / rpb) * s.rpt + (long)sub * rpb + i % rpb; };
      long i = r0 + l;
      // eight rows in flight per thread (the loop was one exposed L2 / HBM latency per row); added in the same order
      for (; i + 7L * lanes < r1; i += 8L * lanes) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(pp + row_of(i + (long)u * lanes) * n_oct + o);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          s1 += v[u].x;
          s2 += v[u].y;
        }
      }
      for (; i < r1; i += lanes) {
        const float2 v = __ldg(pp + row_of(i) * n_oct + o);
        s1 += v.x;
        s2 += v.y;
      }
    }
    sred[threadIdx.x * 2] = s1;
    sred[threadIdx.x * 2 + 1] = s2;
    __syncthreads();
    if ((int)threadIdx.x < n_oct) {
      double a1 = 0.0, a2 = 0.0;
      for (int k = 0; k < lanes; ++k) {
        a1 += sred[(k * n_oct + threadIdx.x) * 2];
        a2 += sred[(k * n_oct + threadIdx.x) * 2 + 1];
      }
      reinterpret_cast<float2*>(s.out)[((long)b * s.NS + sp) * n_oct + threadIdx.x] = make_float2((float)a1, (float)a2);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = (unsigned int)(a.NS + b2.NS);
    s_last = atomicAdd(&counter[b], 1u) == total - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int lane = threadIdx.x & 31;
  const int na = a.n_oct, nb = b2.NS > 0 ? b2.n_oct : 0;
  for (int g = threadIdx.x >> 5; g < G; g += 8) {
    double s = 0.0, ss = 0.0;
    group_octet_sums(g, lane, group_size, a.out, na, a.NS, a.scale, b2.out, nb, b2.NS, b2.scale, b, s, ss);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    const double mean_d = s / count;
    double var = ss / count - mean_d * mean_d;
    if (var < 0.0) var = 0.0;
    const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (lane == 0) {
      mean_rstd[((long)b * G + g) * 2] = mean;
      mean_rstd[((long)b * G + g) * 2 + 1] = rstd;
    }
    if (coef != nullptr) {
      const int Ctot = (na + nb) * 8, Ca = na * 8;
      for (int c = g * group_size + lane; c < (g + 1) * group_size; c += 32) {
        const float ga = gamma[c], be = beta[c];
        float aa = rstd * ga, cc = be - mean * rstd * ga;
        if (scale_shift != nullptr) {
          const float sc = scale_shift[(long)b * ss_stride + c] + 1.0f;
          const float sh = scale_shift[(long)b * ss_stride + Ctot + c];
          aa *= sc;
          cc = cc * sc + sh;
        }
        coef[(long)b * Ctot + c] = make_float2(aa * (c < Ca ? a.scale : b2.scale), cc);
      }
    }
  }
  if (threadIdx.x == 0) counter[b] = 0;  // ready for the next launch (stream order)
}

// ------------------------------------------------------------------------------------------------ LayerNorm (one warp per token)
template <bool F32>
__global__ void layernorm_kernel(const void* __restrict__ x_, const float* __restrict__ g, const float* __restrict__ bias,
                                 const void* __restrict__ res_, void* __restrict__ y_, long M, int C, float eps) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  float s = 0.f, ss = 0.f;
  if (F32) {
    const float* x = reinterpret_cast<const float*>(x_) + row * C;
    for (int c = lane; c < C; c += 32) {
      const float v = x[c];
      s += v;
      ss += v * v;
    }
  } else {
    const h16* x = reinterpret_cast<const h16*>(x_) + row * C;
    for (int o = lane; o < (C >> 3); o += 32) {
      int4 raw = *reinterpret_cast<const int4*>(x + o * 8);
      float v[8];
      h16x8_to_float(*reinterpret_cast<h16x8*>(&raw), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += v[j];
        ss += v[j] * v[j];
      }
    }
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const float mean = s / C;
  float var = ss / C - mean * mean;
  var = var < 0.f ? 0.f : var;
  const float rstd = rsqrtf(var + eps);
  if (F32) {
    const float* x = reinterpret_cast<const float*>(x_) + row * C;
    float* y = reinterpret_cast<float*>(y_) + row * C;
    for (int c = lane; c < C; c += 32) {
      float v = (x[c] - mean) * rstd * g[c];
      if (bias) v += bias[c];
      y[c] = v;
    }
  } else {
    const h16* x = reinterpret_cast<const h16*>(x_) + row * C;
    const h16* res = reinterpret_cast<const h16*>(res_);
    h16* y = reinterpret_cast<h16*>(y_) + row * C;
    for (int o = lane; o < (C >> 3); o += 32) {
      int4 raw = *reinterpret_cast<const int4*>(x + o * 8);
      float v[8];
      h16x8_to_float(*reinterpret_cast<h16x8*>(&raw), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = (v[j] - mean) * rstd * g[o * 8 + j];
        if (bias) v[j] += bias[o * 8 + j];
      }
      if (res) {
        int4 rr = *reinterpret_cast<const int4*>(res + row * C + o * 8);
        float r[8];
        h16x8_to_float(*reinterpret_cast<h16x8*>(&rr), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += r[j];
      }
      h16x8 o8 = float_to_h16x8(v);
      *reinterpret_cast<int4*>(y + o * 8) = *reinterpret_cast<int4*>(&o8);
    }
  }
}

int pick_nblk(long HW, int lanes, int B) {
  // enough blocks to fill 148 SMs a few times over, but at least 16 pixels per pixel-lane of a block (the per-block
  // reduction epilogue is a fixed cost: 592 blocks of 7 pixels each took 60 us on a 64 x 64 x 1024 tensor)
  long want = (long)kd_num_sms() * 4 / (B > 0 ? B : 1);
  if (want < 1) want = 1;
  long maxb = (HW + 16L * lanes - 1) / (16L * lanes);
  if (want > maxb) want = maxb;
  if (want < 1) want = 1;
  return (int)want;
}

// ---- GlobalContext tail in ONE launch: softmax-pool finalize -> Conv1x1(C -> hid) + SiLU -> Conv1x1(hid -> C) + sigmoid.
// One cluster of GG_CL CTAs per batch image.  Every CTA merges the pooling partials (cheap, redundant); the two matrix-vector
// products are split by output rows over the cluster's CTAs (the weights, up to 2 x 2 MB fp32 at C = 1024, are streamed once per
// image by 8 SMs instead of by one); the hidden vector travels between the CTAs through distributed shared memory.
// Replaces kd_gca_finalize + 2 x kd_linear_small (three dependent ~5 us launches per ResnetBlock; 42 blocks per 1024^2 step).
constexpr int GG_CL = 8;
__device__ __forceinline__ void gg_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(GG_CL, 1, 1) __launch_bounds__(256)
gca_gate_kernel(const float* __restrict__ part, const float* __restrict__ ml, int nblk, int C, int hid, const float* __restrict__ w0,
                const float* __restrict__ b0, const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ gate) {
  extern __shared__ float gsm[];  // pooled[C] | hidden_full[hid] | hidden_local[hp] | f[nblk]
  __shared__ float s_red[8];
  __shared__ float s_M, s_L;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int b = blockIdx.x / GG_CL;
  const int hp = (hid + GG_CL - 1) / GG_CL, cp = (C + GG_CL - 1) / GG_CL;
  float* pooled = gsm;
  float* hidden_full = pooled + C;
  float* hidden_local = hidden_full + hid;
  float* f = hidden_local + hp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  kd_pdl_wait();
  kd_pdl_trigger();
  // ---- phase 1: merge the online-softmax partials of kd_gca_pool (fixed order)
  const float* mlb = ml + (long)b * nblk * 2;
  float m = -INFINITY;
  for (int k = threadIdx.x; k < nblk; k += 256) m = fmaxf(m, mlb[k * 2]);
  m = warp_max(m);
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = -INFINITY;
    for (int i = 0; i < 8; ++i) mm = fmaxf(mm, s_red[i]);
    s_M = mm;
  }
  __syncthreads();
  const float M = s_M;
  float l = 0.f;
  for (int k = threadIdx.x; k < nblk; k += 256) {
    const float mk = mlb[k * 2];
    const float fk = (mk == -INFINITY) ? 0.f : __expf(mk - M);
    f[k] = fk;
    l += fk * mlb[k * 2 + 1];
  }
  l = warp_sum(l);
  __syncthreads();
  if (lane == 0) s_red[warp] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float ll = 0.f;
    for (int i = 0; i < 8; ++i) ll += s_red[i];
    s_L = ll;
  }
  __syncthreads();
  const float inv_L = 1.0f / s_L;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float* pp = part + (long)b * nblk * C + c;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 3 < nblk; k += 4) {
      s0 = fmaf(pp[(long)k * C], f[k], s0);
      s1 = fmaf(pp[(long)(k + 1) * C], f[k + 1], s1);
      s2 = fmaf(pp[(long)(k + 2) * C], f[k + 2], s2);
      s3 = fmaf(pp[(long)(k + 3) * C], f[k + 3], s3);
    }
    for (; k < nblk; ++k) s0 = fmaf(pp[(long)k * C], f[k], s0);
    pooled[c] = ((s0 + s1) + (s2 + s3)) * inv_L;
  }
  __syncthreads();
  // ---- phase 2: this CTA's rows of the hidden layer
  const int j0 = (int)rank * hp, j1 = min(hid, j0 + hp);
  for (int j = j0 + warp; j < j1; j += 8) {
    const float* wr = w0 + (long)j * C;
    float acc = 0.f;
    for (int k = lane * 4; k < C; k += 128) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + k));
      const float4 xv = *reinterpret_cast<const float4*>(pooled + k);
      acc += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) hidden_local[j - j0] = silu_f(acc + b0[j]);
  }
  gg_cluster_sync();
  // ---- phase 3: gather the whole hidden vector through distributed shared memory
  for (int idx = threadIdx.x; idx < hid; idx += 256) {
    const uint32_t r = (uint32_t)(idx / hp);
    uint32_t remote;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                 : "=r"(remote)
                 : "r"((uint32_t)__cvta_generic_to_shared(hidden_local + (idx - (int)r * hp))), "r"(r));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    hidden_full[idx] = v;
  }
  gg_cluster_sync();  // nobody leaves (or reuses smem) while a peer still reads its hidden_local
  // ---- phase 4: this CTA's channels of the gate
  const int c0 = (int)rank * cp, c1 = min(C, c0 + cp);
  for (int c = c0 + warp; c < c1; c += 8) {
    const float* wr = w1 + (long)c * hid;
    float acc = 0.f;
    for (int k = lane * 4; k < hid; k += 128) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + k));
      const float4 xv = *reinterpret_cast<const float4*>(hidden_full + k);
      acc += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) gate[(long)b * C + c] = sigmoid_f(acc + b1[c]);
  }
}

}  // namespace

#define KD_CHECK_OCT(C)                                                                                   \
  KD_REQUIRE((C) > 0 && (C) % 8 == 0 && (C) / 8 <= 256, "channel count %d must be a multiple of 8 and <= 2048", (C))

extern "C" int kd_gn_stats(const void* x, int B, long HW, int C, int c_offset, int group_size, int num_groups, float* partial,
                           int nblk, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && partial && B > 0 && HW > 0 && nblk > 0, "kd_gn_stats: bad argument");
  KD_CHECK_OCT(C);
  KD_REQUIRE(group_size % 8 == 0 && c_offset % 8 == 0 && num_groups <= 32, "kd_gn_stats: group_size/c_offset must be multiples of 8");
  const int T = threads_for_oct(C / 8);
  gn_stats_kernel<<<dim3(nblk, B), T, T * 2 * sizeof(float), stream>>>(reinterpret_cast<const h16*>(x), HW, C, c_offset,
                                                                        group_size, num_groups, partial, nblk);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_gn_finalize(const float* partial_a, int nblk_a, float scale_a, const float* partial_b, int nblk_b, float scale_b,
                              int B, int num_groups, double count, float eps, float* mean_rstd, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(partial_a && mean_rstd && B > 0 && num_groups > 0 && num_groups <= 32 && count > 0, "kd_gn_finalize: bad argument");
  gn_finalize_kernel<<<B, 32 * num_groups, 0, stream>>>(partial_a, nblk_a, scale_a, partial_b, nblk_b, scale_b, num_groups, count, eps, mean_rstd);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_gn_apply(const void* x, void* y, int B, long HW, int C, int c_offset, int group_size, int num_groups,
                           float src_scale, const float* mean_rstd, const float* gamma, const float* beta, const float* scale_shift,
                           long ss_stride, int Ctot, int act, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && y && mean_rstd && gamma && beta && B > 0 && HW > 0, "kd_gn_apply: bad argument");
  KD_CHECK_OCT(C);
  KD_REQUIRE(group_size % 8 == 0 && c_offset % 8 == 0, "kd_gn_apply: group_size/c_offset must be multiples of 8");
  const int T = threads_for_oct(C / 8);
  const int nblk = pick_nblk(HW, T / (C / 8), B);
  KD_CUDA(kd_launch(gn_apply_kernel, dim3(nblk, B), dim3(T), 0, stream, reinterpret_cast<const h16*>(x), reinterpret_cast<h16*>(y), HW, C, c_offset, group_size, num_groups, src_scale, mean_rstd, gamma, beta, scale_shift, ss_stride, Ctot, act, nblk));
  return KD_OK;
}

extern "C" int kd_rowdot(const void* x, const float* w, const float* bias, float* out, int B, long HW, int C, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && w && out && B > 0 && HW > 0, "kd_rowdot: bad argument");
  KD_CHECK_OCT(C);
  const int oct = C / 8;
  const long rows = (long)B * HW;
  const int sub = (oct < 32 && (oct & (oct - 1)) == 0) ? oct : 32;
  long blocks = (rows / (32 / sub) + 7) / 8;  // 8 warps per block
  const long cap = (long)kd_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  KD_CUDA(kd_launch(rowdot_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, reinterpret_cast<const h16*>(x), w, bias, out, rows, C));
  return KD_OK;
}

extern "C" int kd_gca_pool(const void* x, const float* logits, int n_parts, int B, long HW, int C, int nblk, float* part, float* ml,
                           kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && logits && part && ml && B > 0 && HW > 0 && nblk > 0 && n_parts >= 1, "kd_gca_pool: bad argument");
  KD_CHECK_OCT(C);
  const int T = threads_for_oct(C / 8);
  const long per = (HW + nblk - 1) / nblk;
  const int e_cache = per <= 8192 ? (int)per : 0;  // chunk logits cached in shared memory (<= 32 KB)
  const size_t smem = sizeof(float) * ((size_t)T * 8 + e_cache);
  KD_CUDA(kd_launch(gca_pool_kernel, dim3(nblk, B), dim3(T), smem, stream, reinterpret_cast<const h16*>(x), logits, n_parts, (long)B * HW, HW, C, nblk, part, ml, e_cache));
  return KD_OK;
}

extern "C" int kd_gca_gate(const float* part, const float* ml, int B, int nblk, int C, int hid, const float* w0, const float* b0,
                           const float* w1, const float* b1, float* gate, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(part && ml && w0 && b0 && w1 && b1 && gate && B > 0 && nblk > 0 && nblk <= 8192, "kd_gca_gate: bad argument");
  KD_REQUIRE(C > 0 && C % 4 == 0 && hid > 0 && hid % 4 == 0, "kd_gca_gate: C (%d) and hid (%d) must be multiples of 4", C, hid);
  const size_t smem = sizeof(float) * ((size_t)C + hid + (hid + GG_CL - 1) / GG_CL + nblk + 8);
  KD_REQUIRE(smem <= 48 * 1024, "kd_gca_gate: shape too large for shared memory");
  KD_CUDA(kd_launch(gca_gate_kernel, dim3(B * GG_CL), dim3(256), smem, stream, part, ml, nblk, C, hid, w0, b0, w1, b1, gate));
  return KD_OK;
}

extern "C" int kd_gca_finalize(const float* part, const float* ml, int B, int nblk, int C, float* pooled, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(part && ml && pooled && B > 0 && nblk > 0 && C > 0, "kd_gca_finalize: bad argument");
  KD_REQUIRE(nblk <= 8192, "kd_gca_finalize: nblk too large");
  KD_CUDA(kd_launch(gca_finalize_kernel, dim3(kd_ceil_div(C, 64), B), dim3(256), (nblk + 256) * sizeof(float), stream, part, ml, nblk, C, pooled));
  return KD_OK;
}

extern "C" int kd_elementwise_blocks(long HW, int C) {
  if (C <= 0 || C % 8 != 0 || C / 8 > 256 || HW <= 0) return 0;
  const int T = threads_for_oct(C / 8);
  return pick_nblk(HW, T / (C / 8), 1);  // independent of the batch size (batch-invariant reduction order)
}

extern "C" int kd_gate_residual(const void* h, const float* gate, const void* res, void* out, float* oct_partial, int B, long HW, int C,
                                kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(h && out && B > 0 && HW > 0, "kd_gate_residual: bad argument");
  KD_CHECK_OCT(C);
  const int T = threads_for_oct(C / 8);
  const int nblk = kd_elementwise_blocks(HW, C);
  KD_CUDA(kd_launch(gate_residual_kernel, dim3(nblk, B), dim3(T), oct_partial ? T * 2 * sizeof(float) : 0, stream,
                    reinterpret_cast<const h16*>(h), gate, reinterpret_cast<const h16*>(res), reinterpret_cast<h16*>(out), oct_partial, HW, C,
                    nblk));
  return KD_OK;
}

extern "C" int kd_oct_stats(const void* x, int B, long HW, int C, float* partial, int nblk, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && partial && B > 0 && HW > 0 && nblk > 0, "kd_oct_stats: bad argument");
  KD_CHECK_OCT(C);
  const int T = threads_for_oct(C / 8);
  KD_CUDA(kd_launch(oct_stats_kernel, dim3(nblk, B), dim3(T), T * 2 * sizeof(float), stream, reinterpret_cast<const h16*>(x), HW, C, partial, nblk));
  return KD_OK;
}

extern "C" int kd_oct_reduce_splits(int rpt, int tiles, int TB) {
  if (rpt <= 0 || tiles <= 0 || TB <= 0) return 0;
  const long count = (long)tiles * (rpt / TB);
  // >= 64 partial rows per split, at most one split per SM (16 rows per split was tried for the small maps: no measurable change
  // of the in-graph step, the reads hit L2 and the launch overlaps its predecessor)
  long ns = count / 64;
  if (ns < 1) ns = 1;
  if (ns > 144) ns = 144;
  return (int)ns;
}

extern "C" int kd_oct_reduce(const float* partial, int rpt, int tiles, int TB, int B, int n_oct, float* out, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(partial && out && rpt > 0 && tiles > 0 && TB > 0 && rpt % TB == 0 && B > 0 && n_oct > 0 && n_oct <= 256,
             "kd_oct_reduce: bad argument");
  const int NS = kd_oct_reduce_splits(rpt, tiles, TB);
  const int T = threads_for_oct(n_oct);
  KD_CUDA(kd_launch(oct_reduce_kernel, dim3(NS, B), dim3(T), T * 2 * sizeof(double), stream, partial, rpt, tiles, TB, n_oct, NS, out));
  return KD_OK;
}

extern "C" int kd_gn_finalize_oct(const float* sum_a, int n_oct_a, int ns_a, float scale_a, const float* sum_b, int n_oct_b, int ns_b,
                                  float scale_b, int B, int num_groups, int group_size, double count, float eps, float* mean_rstd,
                                  const float* gamma, const float* beta, const float* scale_shift, long ss_stride, float* coef,
                                  kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(sum_a && mean_rstd && B > 0 && num_groups > 0 && num_groups <= 32 && group_size % 8 == 0 && count > 0 && ns_a > 0,
             "kd_gn_finalize_oct: bad argument");
  KD_REQUIRE(coef == nullptr || (gamma && beta), "kd_gn_finalize_oct: coefficients need gamma and beta");
  KD_CUDA(kd_launch(gn_finalize_oct_kernel, dim3(B), dim3(32 * num_groups), 0, stream, sum_a, n_oct_a, ns_a, scale_a, sum_b, sum_b ? n_oct_b : 0, sum_b ? ns_b : 0, scale_b, num_groups, group_size, count, eps, mean_rstd, gamma, beta, scale_shift, ss_stride, reinterpret_cast<float2*>(coef)));
  return KD_OK;
}

extern "C" int kd_gn_reduce_finalize(const float* partial_a, int rpt_a, int tiles_a, int TB_a, int n_oct_a, float scale_a,
                                     const float* partial_b, int rpt_b, int tiles_b, int TB_b, int n_oct_b, float scale_b, int B,
                                     int num_groups, int group_size, double count, float eps, float* scratch, unsigned int* counter,
                                     float* mean_rstd, const float* gamma, const float* beta, const float* scale_shift, long ss_stride,
                                     float* coef, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(partial_a && scratch && counter && mean_rstd && B > 0 && num_groups > 0 && num_groups <= 32 && group_size % 8 == 0 && count > 0,
             "kd_gn_reduce_finalize: bad argument");
  KD_REQUIRE(n_oct_a > 0 && n_oct_a <= 256 && rpt_a > 0 && tiles_a > 0 && TB_a > 0 && rpt_a % TB_a == 0, "kd_gn_reduce_finalize: bad source a");
  KD_REQUIRE(!partial_b || (n_oct_b > 0 && n_oct_b <= 256 && rpt_b > 0 && tiles_b > 0 && TB_b > 0 && rpt_b % TB_b == 0),
             "kd_gn_reduce_finalize: bad source b");
  KD_REQUIRE(coef == nullptr || (gamma && beta), "kd_gn_reduce_finalize: coefficients need gamma and beta");
  OctSrc a, b2;
  a.partial = partial_a; a.rpt = rpt_a; a.tiles = tiles_a; a.TB = TB_a; a.n_oct = n_oct_a; a.scale = scale_a;
  a.NS = kd_oct_reduce_splits(rpt_a, tiles_a, TB_a);
  a.out = scratch;
  b2 = a;
  b2.NS = 0;
  if (partial_b) {
    b2.partial = partial_b; b2.rpt = rpt_b; b2.tiles = tiles_b; b2.TB = TB_b; b2.n_oct = n_oct_b; b2.scale = scale_b;
    b2.NS = kd_oct_reduce_splits(rpt_b, tiles_b, TB_b);
    b2.out = scratch + (size_t)B * a.NS * n_oct_a * 2;
  }
  KD_CUDA(kd_launch(gn_reduce_finalize_kernel, dim3(a.NS + b2.NS, B), dim3(256), 0, stream, a, b2, num_groups, group_size, count, eps, mean_rstd, gamma, beta, scale_shift, ss_stride, reinterpret_cast<float2*>(coef), counter));
  return KD_OK;
}

extern "C" int kd_layernorm_h16(const void* x, const float* g, const float* bias, const void* residual, void* y, long M, int C,
                                 float eps, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && g && y && M > 0 && C > 0 && C % 8 == 0, "kd_layernorm_h16: bad argument (C=%d)", C);
  KD_CUDA(kd_launch(layernorm_kernel<false>, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, stream, x, g, bias, residual, y, M, C, eps));
  return KD_OK;
}

extern "C" int kd_layernorm_f32(const float* x, const float* g, const float* bias, float* y, long M, int C, float eps,
                                kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && g && y && M > 0 && C > 0, "kd_layernorm_f32: bad argument");
  KD_CUDA(kd_launch(layernorm_kernel<true>, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, stream, x, g, bias, nullptr, y, M, C, eps));
  return KD_OK;
}
