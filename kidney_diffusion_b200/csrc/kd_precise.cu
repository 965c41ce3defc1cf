// fp32 ("precise") path of the UNet step: the same operations as the fp16 tensor-core path, on fp32 NHWC activations with
// fp32 weights and fp32 FFMA accumulation on the CUDA cores.  It exists for the 1e-4 parity bar of BASELINE.json's north_star
// ("1e-4 for the fp32 path") and as an on-GPU cross-check of the fast path; it is 20-50x slower and is never chosen implicitly
// (Imagen.set_precision("fp32") / Unet.precision).  Every output element is accumulated by one thread in a fixed order, and
// every reduction is a fixed tree over a chunk count that depends on the per-sample shape only: results do not depend on the
// batch size.
//
//   kd_conv_f32           Conv2d k x k / stride / pad over the (virtual) channel concat of two sources, + bias, activation,
//                         gate * addend, optional pixel-shuffle store        (every Conv2d / Linear over pixels of Unet.forward)
//   kd_gn_stats_f32 / kd_gn_finalize_f32 / kd_gn_apply_f32     GroupNorm (+ time scale / shift) + SiLU   (Block.forward)
//   kd_rowdot_f32, kd_softmax_stats_f32, kd_softmax_pool_f32, kd_pool_finalize_f32, kd_gate_residual_f32   (GlobalContext)
//   kd_attn_f32           softmax attention, head dim 64, explicit key / value strides  (Attention: one shared K/V head;
//                         CrossAttention: per-head K/V)
//   kd_dwconv3x3_f32, kd_linattn_f32   LinearAttention / LinearCrossAttention (depthwise 3x3 of q | k | v; softmax_n(k)^T v context;
//                         softmax_d(q) context)
#include "kd_common.cuh"

namespace {

__device__ __forceinline__ float act_precise(float x, int act) {
  if (act == KD_ACT_SILU) return x / (1.0f + expf(-x));
  if (act == KD_ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  if (act == KD_ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
  return x;
}

struct ConvF32 {
  const float* xa;
  const float* xb;
  const float* w;
  const float* bias;
  const float* addend;
  const float* addend_scale;
  float* out;
  int B, Hin, Win, Ca, Cb, Cout, ks, stride, pad, Ho, Wo, act, out_mode;
};

// 64 output pixels x 64 output channels per block, K walked 16 at a time; thread (ty, tx) owns a 4 x 4 micro-tile.
constexpr int CT = 64, CK = 16;
template <bool VEC>
__global__ void __launch_bounds__(256) conv_f32_kernel(const ConvF32 p) {
  __shared__ float As[CK][CT + 4];
  __shared__ float Bs[CK][CT + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int Ctot = p.Ca + p.Cb;
  const int K = p.ks * p.ks * Ctot;
  const long M = (long)p.B * p.Ho * p.Wo;
  const long m0 = (long)blockIdx.x * CT;
  const int n0 = blockIdx.y * CT;
  // loader role: row r of the tile, k offsets kq .. kq + 3
  const int r = tid >> 2, kq = (tid & 3) * 4;
  const long m = m0 + r;
  const bool m_ok = m < M;
  int b = 0, oy = 0, ox = 0;
  if (m_ok) {
    b = (int)(m / ((long)p.Ho * p.Wo));
    const int rem = (int)(m - (long)b * p.Ho * p.Wo);
    oy = rem / p.Wo;
    ox = rem - oy * p.Wo;
  }
  const int n_row = n0 + r;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += CK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = k0 + kq;
    if (VEC) {  // Ca, Cb multiples of 4: the four k share one tap and one source, and K is a multiple of 4
      if (k < K) {
        const int tap = k / Ctot, c = k - tap * Ctot;
        const int ky = tap / p.ks, kx = tap - ky * p.ks;
        const int iy = oy * p.stride - p.pad + ky, ix = ox * p.stride - p.pad + kx;
        if (m_ok && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) {
          const long pix = ((long)b * p.Hin + iy) * p.Win + ix;
          const float4 v = c < p.Ca ? *reinterpret_cast<const float4*>(p.xa + pix * p.Ca + c)
                                    : *reinterpret_cast<const float4*>(p.xb + pix * p.Cb + (c - p.Ca));
          av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
        }
        if (n_row < p.Cout) {
          const float4 v = *reinterpret_cast<const float4*>(p.w + (long)n_row * K + k);
          bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = k + i;
        if (kk < K) {
          const int tap = kk / Ctot, c = kk - tap * Ctot;
          const int ky = tap / p.ks, kx = tap - ky * p.ks;
          const int iy = oy * p.stride - p.pad + ky, ix = ox * p.stride - p.pad + kx;
          if (m_ok && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) {
            const long pix = ((long)b * p.Hin + iy) * p.Win + ix;
            av[i] = c < p.Ca ? p.xa[pix * p.Ca + c] : p.xb[pix * p.Cb + (c - p.Ca)];
          }
          if (n_row < p.Cout) bv[i] = p.w[(long)n_row * K + kk];
        }
      }
    }
    __syncthreads();  // previous step's reads are done
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[kq + i][r] = av[i];
      Bs[kq + i][r] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
  }

  const int Cq = p.Cout / 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
    const int ob = (int)(mm / ((long)p.Ho * p.Wo));
    const int rem = (int)(mm - (long)ob * p.Ho * p.Wo);
    const int y = rem / p.Wo, x = rem - y * p.Wo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j] + (p.bias ? p.bias[n] : 0.0f);
      v = act_precise(v, p.act);
      long idx;
      if (p.out_mode == 1) {  // PixelShuffle(2) store: conv channel n = (dy * 2 + dx) * Cq + c
        const int q = n / Cq, c = n - q * Cq;
        idx = (((long)ob * 2 * p.Ho + 2 * y + (q >> 1)) * (2 * p.Wo) + 2 * x + (q & 1)) * Cq + c;
      } else {
        idx = mm * p.Cout + n;
      }
      if (p.addend) v = fmaf(p.addend[idx], p.addend_scale ? p.addend_scale[(long)ob * p.Cout + n] : 1.0f, v);
      p.out[idx] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------- GroupNorm
// partial[b][g][chunk] = {sum, sum of squares} (fp64) of group g of sample b over the chunk's pixels; value = x * scale of its source
__global__ void __launch_bounds__(256) gn_stats_f32_kernel(const float* __restrict__ xa, int Ca, const float* __restrict__ xb, int Cb,
                                                           float scale_b, long HW, int G, int gs, int nch, double* __restrict__ partial) {
  const int bg = blockIdx.x, ch = blockIdx.y;
  const int b = bg / G, g = bg - b * G;
  const long per = (HW + nch - 1) / nch;
  const long p0 = (long)ch * per, p1 = p0 + per < HW ? p0 + per : HW;
  double s = 0.0, ss = 0.0;
  const long n = (p1 > p0 ? (p1 - p0) : 0) * gs;
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    const long pix = p0 + i / gs;
    const int c = g * gs + (int)(i % gs);
    const float v = c < Ca ? xa[((long)b * HW + pix) * Ca + c] : xb[((long)b * HW + pix) * Cb + (c - Ca)] * scale_b;
    s += (double)v;
    ss += (double)v * (double)v;
  }
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[((long)bg * nch + ch) * 2] = sh[0][0];
    partial[((long)bg * nch + ch) * 2 + 1] = sh[1][0];
  }
}

__global__ void gn_finalize_f32_kernel(const double* __restrict__ partial, int nch, int BG, double count, float eps, float* __restrict__ mean_rstd) {
  const int bg = blockIdx.x * blockDim.x + threadIdx.x;
  if (bg >= BG) return;
  double s = 0.0, ss = 0.0;
  for (int c = 0; c < nch; ++c) {
    s += partial[((long)bg * nch + c) * 2];
    ss += partial[((long)bg * nch + c) * 2 + 1];
  }
  const double mean = s / count;
  double var = ss / count - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_rstd[bg * 2] = (float)mean;
  mean_rstd[bg * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// y = act(((x * src_scale - mean) * rstd * gamma + beta) * (scale + 1) + shift) for the C channels [c_offset, c_offset + C) of a
// ctot-channel GroupNorm; scale_shift row b = [scale (ctot) | shift (ctot)]
__global__ void __launch_bounds__(256) gn_apply_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long HW, int C, int c_offset,
                                                           int gs, int G, float src_scale, const float* __restrict__ mean_rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ ss, long ss_stride, int ctot, int act, long total) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    const int b = (int)(i / ((long)HW * C));
    const int cg = c_offset + c;
    const int g = cg / gs;
    const float mean = mean_rstd[((long)b * G + g) * 2], rstd = mean_rstd[((long)b * G + g) * 2 + 1];
    float v = (x[i] * src_scale - mean) * rstd * gamma[cg] + beta[cg];
    if (ss) v = v * (ss[(long)b * ss_stride + cg] + 1.0f) + ss[(long)b * ss_stride + ctot + cg];
    y[i] = act_precise(v, act);
  }
}

// ---------------------------------------------------------------------------------------------- GlobalContext
__global__ void __launch_bounds__(256) rowdot_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ out, long M, int C) {
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float s = 0.0f;
  for (int c = lane; c < C; c += 32) s = fmaf(x[row * C + c], w[c], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s + (bias ? bias[0] : 0.0f);
}

// ml[b] = {max, sum exp(l - max)} of the HW logits of sample b
__global__ void __launch_bounds__(256) softmax_stats_f32_kernel(const float* __restrict__ logits, long HW, float* __restrict__ ml) {
  const int b = blockIdx.x;
  const float* l = logits + (long)b * HW;
  __shared__ float sh[256];
  float mx = -INFINITY;
  for (long i = threadIdx.x; i < HW; i += 256) mx = fmaxf(mx, l[i]);
  sh[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] = fmaxf(sh[threadIdx.x], sh[threadIdx.x + o]);
    __syncthreads();
  }
  mx = sh[0];
  __syncthreads();
  float s = 0.0f;
  for (long i = threadIdx.x; i < HW; i += 256) s += expf(l[i] - mx);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ml[b * 2] = mx;
    ml[b * 2 + 1] = sh[0];
  }
}

// part[b][chunk][c] = sum over the chunk's pixels of exp(l - max) * x[pixel][c]; thread = channel, pixels in order
__global__ void __launch_bounds__(128) softmax_pool_f32_kernel(const float* __restrict__ x, const float* __restrict__ logits,
                                                               const float* __restrict__ ml, long HW, int C, int nch, float* __restrict__ part) {
  const int b = blockIdx.z, ch = blockIdx.y;
  const int c = blockIdx.x * 128 + threadIdx.x;
  const long per = (HW + nch - 1) / nch;
  const long p0 = (long)ch * per, p1 = p0 + per < HW ? p0 + per : HW;
  const float mx = ml[b * 2];
  float acc = 0.0f;
  if (c < C)
    for (long pix = p0; pix < p1; ++pix) acc = fmaf(expf(logits[(long)b * HW + pix] - mx), x[((long)b * HW + pix) * C + c], acc);
  if (c < C) part[((long)b * nch + ch) * C + c] = acc;
}

__global__ void pool_finalize_f32_kernel(const float* __restrict__ part, const float* __restrict__ ml, int nch, int C, float* __restrict__ pooled,
                                         int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / C, c = i - b * C;
  float s = 0.0f;
  for (int k = 0; k < nch; ++k) s += part[((long)b * nch + k) * C + c];
  pooled[i] = s / ml[b * 2 + 1];
}

__global__ void __launch_bounds__(256) gate_residual_f32_kernel(const float* __restrict__ h, const float* __restrict__ gate,
                                                                const float* __restrict__ res, float* __restrict__ out, long HW, int C, long total) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    const long b = i / (HW * C);
    out[i] = fmaf(h[i], gate[b * C + c], res[i]);
  }
}

// ---------------------------------------------------------------------------------------------- attention (head dim 64)
// one warp per (sample, head, query); lane owns dims 2 * lane, 2 * lane + 1; keys in order, online softmax in fp32
__global__ void __launch_bounds__(256) attn_f32_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ k, long ldk, long k_batch,
                                                       int k_head, const float* __restrict__ v, long ldv, long v_batch, int v_head,
                                                       float* __restrict__ out, int B, int N, int J, int heads, float scale) {
  const long wid = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= (long)B * N * heads) return;
  const int h = (int)(wid % heads);
  const long bn = wid / heads;
  const int b = (int)(bn / N);
  const float2 qv = *reinterpret_cast<const float2*>(q + bn * ldq + h * 64 + 2 * lane);
  const float q0 = qv.x * scale, q1 = qv.y * scale;
  const float* kp = k + (long)b * k_batch + (long)h * k_head + 2 * lane;
  const float* vp = v + (long)b * v_batch + (long)h * v_head + 2 * lane;
  float m = -INFINITY, l = 0.0f, a0 = 0.0f, a1 = 0.0f;
  for (int j = 0; j < J; ++j) {
    const float2 kv = *reinterpret_cast<const float2*>(kp + (long)j * ldk);
    float s = fmaf(q0, kv.x, q1 * kv.y);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn), pj = expf(s - mn);
    const float2 vv = *reinterpret_cast<const float2*>(vp + (long)j * ldv);
    l = fmaf(l, corr, pj);
    a0 = fmaf(a0, corr, pj * vv.x);
    a1 = fmaf(a1, corr, pj * vv.y);
    m = mn;
  }
  const float inv = 1.0f / l;
  *reinterpret_cast<float2*>(out + (bn * heads + h) * 64 + 2 * lane) = make_float2(a0 * inv, a1 * inv);
}

// ---------------------------------------------------------------------------------------------- linear attention (head dim 64)
__global__ void __launch_bounds__(256) dwconv3x3_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                                            int H, int W, int C, long total) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    long pix = i / C;
    const int xw = (int)(pix % W);
    pix /= W;
    const int yh = (int)(pix % H);
    const long b = pix / H;
    float acc = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = yh + ky - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = xw + kx - 1;
        if (ix < 0 || ix >= W) continue;
        acc = fmaf(x[((b * H + iy) * W + ix) * C + c], w[c * 9 + ky * 3 + kx], acc);
      }
    }
    y[i] = acc;
  }
}

// ctx[b][h][d][e] = sum_n softmax_n(k[., d])[n] * v[n][e]: one block per (sample, head); thread -> (d = tid / 4, 16 values of e);
// positions in order, fp32
__global__ void __launch_bounds__(256) linattn_ctx_f32_kernel(const float* __restrict__ k, const float* __restrict__ v, long ld, long batch_stride,
                                                              int N, int heads, float* __restrict__ ctx) {
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const float* kp = k + (long)b * batch_stride + h * 64;
  const float* vp = v + (long)b * batch_stride + h * 64;
  __shared__ float s_max[4][64];
  {
    const int d = threadIdx.x & 63, part = threadIdx.x >> 6;
    float m = -INFINITY;
    for (int n = part; n < N; n += 4) m = fmaxf(m, kp[(long)n * ld + d]);
    s_max[part][d] = m;
  }
  __syncthreads();
  const int d = threadIdx.x >> 2, e0 = (threadIdx.x & 3) * 16;
  const float m = fmaxf(fmaxf(s_max[0][d], s_max[1][d]), fmaxf(s_max[2][d], s_max[3][d]));
  float acc[16], l = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
  for (int n = 0; n < N; ++n) {
    const float wgt = expf(kp[(long)n * ld + d] - m);
    const float4* v4 = reinterpret_cast<const float4*>(vp + (long)n * ld + e0);
    l += wgt;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = v4[q];
      acc[q * 4 + 0] = fmaf(wgt, t.x, acc[q * 4 + 0]);
      acc[q * 4 + 1] = fmaf(wgt, t.y, acc[q * 4 + 1]);
      acc[q * 4 + 2] = fmaf(wgt, t.z, acc[q * 4 + 2]);
      acc[q * 4 + 3] = fmaf(wgt, t.w, acc[q * 4 + 3]);
    }
  }
  const float inv = 1.0f / l;
  float* o = ctx + (((long)b * heads + h) * 64 + d) * 64 + e0;
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = acc[j] * inv;
}

// out[b][n][h*64 + e] = act(scale * sum_d softmax_d(q[b][n][h*64 + .])[d] * ctx[b][h][d][e]): one warp per (sample, position, head)
__global__ void __launch_bounds__(256) linattn_apply_f32_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ ctx,
                                                                float* __restrict__ out, int B, int N, int heads, float scale, int act) {
  const long wid = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= (long)B * N * heads) return;
  const int h = (int)(wid % heads);
  const long bn = wid / heads;
  const int b = (int)(bn / N);
  const float2 qv = *reinterpret_cast<const float2*>(q + bn * ldq + h * 64 + 2 * lane);
  float m = fmaxf(qv.x, qv.y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float p0 = expf(qv.x - m), p1 = expf(qv.y - m);
  float l = p0 + p1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  const float* cp = ctx + ((long)b * heads + h) * 4096 + 2 * lane;
  float a0 = 0.0f, a1 = 0.0f;
  for (int d = 0; d < 64; ++d) {
    const float pd = __shfl_sync(0xffffffffu, (d & 1) ? p1 : p0, d >> 1);
    const float2 c = *reinterpret_cast<const float2*>(cp + d * 64);
    a0 = fmaf(pd, c.x, a0);
    a1 = fmaf(pd, c.y, a1);
  }
  const float f = scale / l;
  *reinterpret_cast<float2*>(out + (bn * heads + h) * 64 + 2 * lane) = make_float2(act_precise(a0 * f, act), act_precise(a1 * f, act));
}

inline unsigned grid_1d(long total, int block, int cap_per_sm = 16) {
  long blocks = (total + block - 1) / block;
  const long cap = (long)kd_num_sms() * cap_per_sm;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" int kd_conv_f32(const float* xa, int Ca, const float* xb, int Cb, const float* w, const float* bias, const float* addend,
                           const float* addend_scale, float* out, int B, int Hin, int Win, int Cout, int ksize, int stride, int pad,
                           int act, int out_mode, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(xa && w && out && B > 0 && Hin > 0 && Win > 0 && Ca > 0 && Cb >= 0 && Cout > 0 && ksize > 0 && stride > 0 && pad >= 0,
             "kd_conv_f32: bad argument");
  KD_REQUIRE(Cb == 0 || xb, "kd_conv_f32: second source missing");
  KD_REQUIRE(out_mode == 0 || (out_mode == 1 && Cout % 4 == 0), "kd_conv_f32: pixel-shuffle store needs Cout %% 4 == 0");
  ConvF32 p;
  p.xa = xa; p.xb = xb; p.w = w; p.bias = bias; p.addend = addend; p.addend_scale = addend_scale; p.out = out;
  p.B = B; p.Hin = Hin; p.Win = Win; p.Ca = Ca; p.Cb = Cb; p.Cout = Cout; p.ks = ksize; p.stride = stride; p.pad = pad;
  p.Ho = (Hin + 2 * pad - ksize) / stride + 1;
  p.Wo = (Win + 2 * pad - ksize) / stride + 1;
  p.act = act; p.out_mode = out_mode;
  KD_REQUIRE(p.Ho > 0 && p.Wo > 0, "kd_conv_f32: empty output");
  const long M = (long)B * p.Ho * p.Wo;
  const dim3 grid((unsigned)((M + CT - 1) / CT), (unsigned)((Cout + CT - 1) / CT));
  const bool vec = Ca % 4 == 0 && Cb % 4 == 0 && (reinterpret_cast<uintptr_t>(xa) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (!xb || (reinterpret_cast<uintptr_t>(xb) & 15) == 0);
  if (vec)
    conv_f32_kernel<true><<<grid, 256, 0, stream>>>(p);
  else
    conv_f32_kernel<false><<<grid, 256, 0, stream>>>(p);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_gn_chunks_f32(long HW) {
  long n = HW / 1024;
  return (int)(n < 1 ? 1 : (n > 64 ? 64 : n));
}

extern "C" int kd_gn_stats_f32(const float* xa, int Ca, const float* xb, int Cb, float scale_b, int B, long HW, int G, double* partial,
                               kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(xa && partial && B > 0 && HW > 0 && G > 0 && (Ca + Cb) % G == 0 && (Cb == 0 || xb), "kd_gn_stats_f32: bad argument");
  const int nch = kd_gn_chunks_f32(HW);
  gn_stats_f32_kernel<<<dim3((unsigned)(B * G), (unsigned)nch), 256, 0, stream>>>(xa, Ca, xb, Cb, scale_b, HW, G, (Ca + Cb) / G, nch, partial);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_gn_finalize_f32(const double* partial, int B, long HW, int G, int group_size, float eps, float* mean_rstd, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(partial && mean_rstd && B > 0 && G > 0, "kd_gn_finalize_f32: bad argument");
  gn_finalize_f32_kernel<<<(B * G + 127) / 128, 128, 0, stream>>>(partial, kd_gn_chunks_f32(HW), B * G, (double)group_size * (double)HW, eps, mean_rstd);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_gn_apply_f32(const float* x, float* y, int B, long HW, int C, int c_offset, int group_size, int G, float src_scale,
                               const float* mean_rstd, const float* gamma, const float* beta, const float* scale_shift, long ss_stride,
                               int ctot, int act, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && y && mean_rstd && gamma && beta && B > 0 && HW > 0 && C > 0, "kd_gn_apply_f32: bad argument");
  const long total = (long)B * HW * C;
  gn_apply_f32_kernel<<<grid_1d(total, 256), 256, 0, stream>>>(x, y, HW, C, c_offset, group_size, G, src_scale, mean_rstd, gamma, beta, scale_shift,
                                                               ss_stride, ctot, act, total);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_rowdot_f32(const float* x, const float* w, const float* bias, float* out, long M, int C, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && w && out && M > 0 && C > 0, "kd_rowdot_f32: bad argument");
  rowdot_f32_kernel<<<(unsigned)((M + 7) / 8), 256, 0, stream>>>(x, w, bias, out, M, C);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_softmax_pool_f32(const float* x, const float* logits, int B, long HW, int C, float* ml /* [B,2] */,
                                   float* part /* [B, kd_gn_chunks_f32(HW), C] */, float* pooled /* [B,C] */, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && logits && ml && part && pooled && B > 0 && HW > 0 && C > 0, "kd_softmax_pool_f32: bad argument");
  const int nch = kd_gn_chunks_f32(HW);
  softmax_stats_f32_kernel<<<B, 256, 0, stream>>>(logits, HW, ml);
  KD_LAUNCH_CHECK();
  softmax_pool_f32_kernel<<<dim3((unsigned)((C + 127) / 128), (unsigned)nch, (unsigned)B), 128, 0, stream>>>(x, logits, ml, HW, C, nch, part);
  KD_LAUNCH_CHECK();
  pool_finalize_f32_kernel<<<(B * C + 127) / 128, 128, 0, stream>>>(part, ml, nch, C, pooled, B * C);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_gate_residual_f32(const float* h, const float* gate, const float* res, float* out, int B, long HW, int C, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(h && gate && res && out && B > 0 && HW > 0 && C > 0, "kd_gate_residual_f32: bad argument");
  const long total = (long)B * HW * C;
  gate_residual_f32_kernel<<<grid_1d(total, 256), 256, 0, stream>>>(h, gate, res, out, HW, C, total);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_attn_f32(const float* q, long ldq, const float* k, long ldk, long k_batch, int k_head, const float* v, long ldv, long v_batch,
                           int v_head, float* out, int B, int N, int J, int heads, float scale, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(q && k && v && out && B > 0 && N > 0 && J > 0 && heads > 0, "kd_attn_f32: bad argument");
  KD_REQUIRE(ldq % 2 == 0 && ldk % 2 == 0 && ldv % 2 == 0 && k_head % 2 == 0 && v_head % 2 == 0 && k_batch % 2 == 0 && v_batch % 2 == 0 &&
                 (reinterpret_cast<uintptr_t>(q) & 7) == 0 && (reinterpret_cast<uintptr_t>(k) & 7) == 0 && (reinterpret_cast<uintptr_t>(v) & 7) == 0,
             "kd_attn_f32: 8-byte alignment of q / k / v rows required");
  const long warps = (long)B * N * heads;
  attn_f32_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(q, ldq, k, ldk, k_batch, k_head, v, ldv, v_batch, v_head, out, B, N, J, heads, scale);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_dwconv3x3_f32(const float* x, const float* w, float* y, int B, int H, int W, int C, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0 && C > 0, "kd_dwconv3x3_f32: bad argument");
  const long total = (long)B * H * W * C;
  dwconv3x3_f32_kernel<<<grid_1d(total, 256), 256, 0, stream>>>(x, w, y, H, W, C, total);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_linattn_f32(const float* q, long ldq, const float* k, const float* v, long ldkv, long kv_batch, int B, int N, int J, int heads,
                              float scale, int act, float* ctx /* [B, heads, 64, 64] */, float* out /* [B, N, heads * 64] */,
                              kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(q && k && v && ctx && out && B > 0 && N > 0 && J > 0 && heads > 0, "kd_linattn_f32: bad argument");
  KD_REQUIRE(ldq % 2 == 0 && ldkv % 4 == 0 && kv_batch % 4 == 0 && (reinterpret_cast<uintptr_t>(q) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 3) == 0,
             "kd_linattn_f32: alignment of q (8 bytes) / v rows (16 bytes) required");
  linattn_ctx_f32_kernel<<<B * heads, 256, 0, stream>>>(k, v, ldkv, kv_batch, J, heads, ctx);
  KD_LAUNCH_CHECK();
  const long warps = (long)B * N * heads;
  linattn_apply_f32_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(q, ldq, ctx, out, B, N, heads, scale, act);
  KD_LAUNCH_CHECK();
  return KD_OK;
}
