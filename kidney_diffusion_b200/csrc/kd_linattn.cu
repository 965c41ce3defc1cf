// Linear attention of imagen-pytorch's LinearAttention block (Unet(use_linear_attn=...)): NHWC fp16 activations, fp32 statistics.
//
//   q, k, v = depthwise3x3(conv1x1(ChanLayerNorm(x)))            [B, N, heads*64] each (N = H*W pixels); 1x1 conv = kd_conv_gemm
//   k <- softmax over positions n (pixels + context tokens), per (head, d);   q <- softmax over d per (pixel, head), * scale
//   ctx[h][d][e] = sum_n k[n, h, d] * v[n, h, e];   out[n, h, e] = SiLU(sum_d q[n, h, d] * ctx[h][d][e])
//
// The matrix products are 64 x 64 per head -- far below a tensor-core tile -- and the block is bound by streaming q / k / v once
// (HBM), so these are CUDA-core kernels:  kd_dwconv3x3 (q|k|v in one pass), kd_linattn_kmax (column max of k), kd_linattn_ctx
// (exp-weighted k^T v partials per pixel chunk, fixed order), kd_linattn_merge (sum of partials, normalise), kd_linattn_apply.
#include "kd_common.cuh"

namespace {

constexpr int LA_D = 64;  // head dimension (the reference's attn_dim_head)

// y[b,h,w,c] = sum_{ky,kx} x[b,h+ky-1,w+kx-1,c] * wt[c][ky][kx], zero padding; 8 channels per thread
__global__ void dwconv3x3_kernel(const h16* __restrict__ x, const float* __restrict__ wt, h16* __restrict__ y, int H, int W, int C, long total8) {
  const int oct = C >> 3;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
    const int o = (int)(i % oct);
    long p = i / oct;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const long b = p / H;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ww = w + kx - 1;
        if (ww < 0 || ww >= W) continue;
        const int4 raw = *reinterpret_cast<const int4*>(x + (((b * H + hh) * W + ww) * (long)C + o * 8));
        float v[8];
        h16x8_to_float(*reinterpret_cast<const h16x8*>(&raw), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], __ldg(wt + (o * 8 + j) * 9 + ky * 3 + kx), acc[j]);
      }
    }
    *reinterpret_cast<h16x8*>(y + (((b * H + h) * W + w) * (long)C + o * 8)) = float_to_h16x8(acc);
  }
}

// column max of k over a chunk of pixels: part[b][blk][inner]; the context tokens are folded in by the consumers
__global__ void linattn_kmax_kernel(const h16* __restrict__ qkv, long ld, int k_col, int N, int inner, int nblk, float* __restrict__ part) {
  const int b = blockIdx.y, blk = blockIdx.x;
  const long per = ((long)N + nblk - 1) / nblk;
  const long n0 = blk * per, n1 = min((long)N, n0 + per);
  for (int c = threadIdx.x; c < inner; c += blockDim.x) {
    float m = -INFINITY;
    const h16* p = qkv + (long)b * N * ld + k_col + c;
    for (long n = n0; n < n1; ++n) m = fmaxf(m, __half2float(p[n * ld]));
    part[((long)b * nblk + blk) * inner + c] = m;
  }
}

// One block per (pixel chunk, head, image): ctx partial [64][64] = sum_n exp(k[n,d] - M[d]) * v[n,e], l[d] = sum_n exp(k[n,d] - M[d]).
// 256 threads: thread t -> d = t >> 2, e in [16 * (t & 3), +16).  Block 0 of every (head, image) also adds the context tokens.
__global__ void __launch_bounds__(256) linattn_ctx_kernel(const h16* __restrict__ qkv, long ld, int k_col, int v_col, int N, int heads,
                                                          const float* __restrict__ ctx_kv /* [B][J][2*inner] or NULL */, int J, int nblk_max,
                                                          const float* __restrict__ kmax_part, int nblk, float* __restrict__ ctx_part,
                                                          float* __restrict__ l_part) {
  __shared__ float s_k[32][LA_D + 1];
  __shared__ float s_v[32][LA_D];
  __shared__ float s_M[LA_D];
  const int blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int inner = heads * LA_D;
  const int d = threadIdx.x >> 2, eg = (threadIdx.x & 3) * 16;
  if (threadIdx.x < LA_D) {  // global column max over all pixel chunks and the context tokens (same value in every block)
    const int c = h * LA_D + threadIdx.x;
    float m = -INFINITY;
    for (int q = 0; q < nblk_max; ++q) m = fmaxf(m, kmax_part[((long)b * nblk_max + q) * inner + c]);
    for (int j = 0; j < J; ++j) m = fmaxf(m, ctx_kv[((long)b * J + j) * 2 * inner + c]);
    s_M[threadIdx.x] = m;
  }
  __syncthreads();
  const float M = s_M[d];
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  float l = 0.f;
  const long per = ((long)N + nblk - 1) / nblk;
  const long n0 = blk * per, n1 = min((long)N, n0 + per);
  for (long t0 = n0; t0 < n1; t0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 16; i += 256) {  // 32 pixels x (8 k-octets + 8 v-octets)
      const int r = i >> 4, o = i & 15;
      const long n = t0 + r;
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (n < n1) {
        const int col = (o < 8 ? k_col : v_col) + h * LA_D + (o & 7) * 8;
        const int4 raw = *reinterpret_cast<const int4*>(qkv + ((long)b * N + n) * ld + col);
        h16x8_to_float(*reinterpret_cast<const h16x8*>(&raw), f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (o < 8) s_k[r][(o & 7) * 8 + j] = (n < n1) ? f[j] : -INFINITY;
        else s_v[r][(o & 7) * 8 + j] = f[j];
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const float w = __expf(s_k[r][d] - M);  // exp(-inf) = 0 for rows past the chunk
      l += w;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(w, s_v[r][eg + j], acc[j]);
    }
  }
  if (blk == 0) {  // context tokens (fp32 rows: k | v)
    for (int j = 0; j < J; ++j) {
      const float* row = ctx_kv + ((long)b * J + j) * 2 * inner;
      const float w = __expf(row[h * LA_D + d] - M);
      l += w;
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = fmaf(w, row[inner + h * LA_D + eg + e], acc[e]);
    }
  }
  float* out = ctx_part + ((((long)b * heads + h) * nblk + blk) * LA_D + d) * LA_D + eg;
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = acc[j];
  if ((threadIdx.x & 3) == 0) l_part[(((long)b * heads + h) * nblk + blk) * LA_D + d] = l;
}

// ctx[b][h][d][e] = (sum over chunks, fixed order) / (sum of l): the normalised softmax-over-positions context
__global__ void linattn_merge_kernel(const float* __restrict__ ctx_part, const float* __restrict__ l_part, int nblk, float* __restrict__ ctx) {
  const long bh = blockIdx.x;
  for (int i = threadIdx.x; i < LA_D * LA_D; i += blockDim.x) {
    const int d = i / LA_D;
    float s = 0.f, l = 0.f;
    for (int q = 0; q < nblk; ++q) {
      s += ctx_part[((bh * nblk + q) * LA_D * LA_D) + i];
      l += l_part[(bh * nblk + q) * LA_D + d];
    }
    ctx[bh * LA_D * LA_D + i] = s / l;
  }
}

// out[n, h*64 + e] = act(scale * sum_d softmax_d(q[n, h, :])[d] * ctx[h][d][e]); block = 32 pixels x one head
__global__ void __launch_bounds__(256) linattn_apply_kernel(const h16* __restrict__ q, long ld, int q_col, int N, int heads,
                                                            const float* __restrict__ ctx, float scale, int act, h16* __restrict__ out) {
  __shared__ float s_ctx[LA_D][LA_D];
  __shared__ float s_p[32][LA_D + 1];
  const int h = blockIdx.y, b = blockIdx.z;
  const long t0 = (long)blockIdx.x * 32;
  const int inner = heads * LA_D;
  const float* cp = ctx + ((long)b * heads + h) * LA_D * LA_D;
  for (int i = threadIdx.x; i < LA_D * LA_D; i += 256) s_ctx[i / LA_D][i % LA_D] = cp[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < 32; r += 8) {  // softmax over the 64 head channels of one pixel per warp pass
    const long n = t0 + r;
    float a = -INFINITY, c = -INFINITY;
    if (n < N) {
      const h16* qp = q + ((long)b * N + n) * ld + q_col + h * LA_D;
      a = __half2float(qp[lane]);
      c = __half2float(qp[lane + 32]);
    }
    const float m = warp_max(fmaxf(a, c));
    const float ea = (n < N) ? __expf(a - m) : 0.f, ec = (n < N) ? __expf(c - m) : 0.f;
    const float inv = 1.0f / fmaxf(warp_sum(ea + ec), 1e-30f);
    s_p[r][lane] = ea * inv;
    s_p[r][lane + 32] = ec * inv;
  }
  __syncthreads();
  const int r = threadIdx.x >> 3, e0 = (threadIdx.x & 7) * 8;
  const long n = t0 + r;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 8
  for (int d = 0; d < LA_D; ++d) {
    const float p = s_p[r][d];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(p, s_ctx[d][e0 + j], acc[j]);
  }
  if (n < N) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= scale;
    apply_act8(acc, act);
    *reinterpret_cast<h16x8*>(out + ((long)b * N + n) * inner + h * LA_D + e0) = float_to_h16x8(acc);
  }
}

unsigned la_blocks(long n) {
  long b = (n + 255) / 256;
  const long cap = (long)kd_num_sms() * 16;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int kd_dwconv3x3(const void* x, const float* w, void* y, int B, int H, int W, int C, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "kd_dwconv3x3: bad argument");
  const long total8 = (long)B * H * W * (C / 8);
  dwconv3x3_kernel<<<la_blocks(total8), 256, 0, stream>>>(reinterpret_cast<const h16*>(x), w, reinterpret_cast<h16*>(y), H, W, C, total8);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_linattn_blocks(int N) {
  int nblk = (N + 255) / 256;  // >= 256 pixels per chunk; a function of the per-sample size only (batch-invariant order)
  if (nblk > 64) nblk = 64;
  return nblk < 1 ? 1 : nblk;
}

extern "C" int kd_linattn_context(const void* qkv, long ld, int k_col, int v_col, int B, int N, int heads, const float* ctx_kv, int J,
                                  float* workspace, size_t ws_bytes, float* ctx, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(qkv && workspace && ctx && B > 0 && N >= 0 && heads > 0 && J >= 0 && N + J > 0 && (J == 0 || ctx_kv),
             "kd_linattn_context: bad argument");
  KD_REQUIRE(ld % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0, "kd_linattn_context: ld / column offsets must be multiples of 8");
  const int nblk = kd_linattn_blocks(N), inner = heads * LA_D;
  const size_t need = sizeof(float) * ((size_t)B * nblk * inner + (size_t)B * heads * nblk * (LA_D * LA_D + LA_D));
  KD_REQUIRE(ws_bytes >= need, "kd_linattn_context: workspace too small (%zu < %zu)", ws_bytes, need);
  float* kmax = workspace;
  float* ctx_part = kmax + (size_t)B * nblk * inner;
  float* l_part = ctx_part + (size_t)B * heads * nblk * LA_D * LA_D;
  const h16* p = reinterpret_cast<const h16*>(qkv);
  linattn_kmax_kernel<<<dim3(nblk, B), 256, 0, stream>>>(p, ld, k_col, N, inner, nblk, kmax);
  KD_LAUNCH_CHECK();
  linattn_ctx_kernel<<<dim3(nblk, heads, B), 256, 0, stream>>>(p, ld, k_col, v_col, N, heads, ctx_kv, J, nblk, kmax, nblk, ctx_part, l_part);
  KD_LAUNCH_CHECK();
  linattn_merge_kernel<<<B * heads, 256, 0, stream>>>(ctx_part, l_part, nblk, ctx);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" size_t kd_linattn_workspace_bytes(int B, int N, int heads) {
  const int nblk = kd_linattn_blocks(N), inner = heads * LA_D;
  return sizeof(float) * ((size_t)B * nblk * inner + (size_t)B * heads * nblk * (LA_D * LA_D + LA_D));
}

extern "C" int kd_linattn_apply(const void* q, long ld, int q_col, const float* ctx, void* out, int B, int N, int heads, float scale, int act,
                                kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(q && ctx && out && B > 0 && N > 0 && heads > 0 && ld % 8 == 0 && q_col % 8 == 0, "kd_linattn_apply: bad argument");
  linattn_apply_kernel<<<dim3((N + 31) / 32, heads, B), 256, 0, stream>>>(reinterpret_cast<const h16*>(q), ld, q_col, N, heads, ctx, scale, act,
                                                                         reinterpret_cast<h16*>(out));
  KD_LAUNCH_CHECK();
  return KD_OK;
}
