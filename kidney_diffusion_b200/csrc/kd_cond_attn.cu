// Conditioning towers (small-M fp32 linear, sinusoidal embedding) and attention kernels.
#include <mma.h>

#include "kd_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ small-M linear
// One warp per output column n; the weight row is streamed once (coalesced float4) per group of 8 input rows.
template <int ROWS>
__global__ void linear_small_kernel(const float* __restrict__ x, int M, int K, long ldx, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ y, int N, long ldy, int pre_act,
                                    int post_act) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const float* wr = w + (long)n * K;
  const bool vec = (K % 4 == 0) && (ldx % 4 == 0);
  for (int m0 = blockIdx.y * ROWS; m0 < M; m0 += gridDim.y * ROWS) {
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
    if (vec) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 wv = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          if (m0 + r < M) {
            float4 xv = *reinterpret_cast<const float4*>(x + (long)(m0 + r) * ldx + k);
            if (pre_act) {
              xv.x = apply_act(xv.x, pre_act); xv.y = apply_act(xv.y, pre_act);
              xv.z = apply_act(xv.z, pre_act); xv.w = apply_act(xv.w, pre_act);
            }
            acc[r] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
          }
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float wv = wr[k];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
          if (m0 + r < M) acc[r] += apply_act(x[(long)(m0 + r) * ldx + k], pre_act) * wv;
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const float s = warp_sum(acc[r]);
      if (lane == 0 && m0 + r < M) y[(long)(m0 + r) * ldy + n] = apply_act(s + (bias ? bias[n] : 0.f), post_act);
    }
  }
}

__global__ void sinu_emb_kernel(const float* __restrict__ t, const float* __restrict__ w, int B, int half, float* __restrict__ out) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int D = 2 * half + 1;
  if (i >= B * D) return;
  const int b = i / D, j = i % D;
  const float x = t[b];
  float v;
  if (j == 0) {
    v = x;
  } else {
    const int k = (j - 1) % half;
    // reference: freqs = x * w * 2 * pi (left to right, fp32)
    const float f = __fmul_rn(__fmul_rn(__fmul_rn(x, w[k]), 2.0f), 3.14159265358979323846f);
    v = (j - 1) < half ? sinf(f) : cosf(f);
  }
  out[i] = v;
}

// ------------------------------------------------------------------------------------------------ K/V assembly for MQA
// kv_out[b] = [ ctx rows (Jc) | null row | token rows (N) ], each row = 64 k values then 64 v values (h16)
__global__ void kv_assemble_kernel(const h16* __restrict__ qkv, long ld, int kv_col, const float* __restrict__ ctx_kv, int Jc,
                                   const float* __restrict__ null_kv, h16* __restrict__ kv_out, int N) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  const int b = blockIdx.y;
  const int J = Jc + 1 + N;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (row, 8-col group): 16 groups per row
  if (idx >= (long)J * 16) return;
  const int row = (int)(idx >> 4), g = (int)(idx & 15);
  h16* dst = kv_out + ((long)b * J + row) * 128 + g * 8;
  if (row < Jc) {
    const float* src = ctx_kv + ((long)b * Jc + row) * 128 + g * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = src[j];
    *reinterpret_cast<h16x8*>(dst) = float_to_h16x8(v);
  } else if (row == Jc) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = null_kv[g * 8 + j];  // [2,64] row-major = k(64) then v(64)
    *reinterpret_cast<h16x8*>(dst) = float_to_h16x8(v);
  } else {
    const h16* src = qkv + ((long)b * N + (row - Jc - 1)) * ld + kv_col + g * 8;
    *reinterpret_cast<int4*>(dst) = *reinterpret_cast<const int4*>(src);
  }
}

// ------------------------------------------------------------------------------------------------ MQA flash attention
// One CTA = 64 queries of one (b, head); 4 warps x 16 query rows; keys processed in tiles of 64 with online softmax.
// Tensor cores via mma.sync.m16n8k16 (h16 in, fp32 accumulate).  K/V tiles ([64 keys][64] k and v) staged in smem.
__device__ __forceinline__ void mma_f16_16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int AT_BQ = 64, AT_BK = 64, AT_D = 64, AT_PAD = 8;  // smem row = 72 h16 (144 B) to avoid bank conflicts

__global__ void __launch_bounds__(128) attn_mqa_kernel(const h16* __restrict__ q, long ldq, const h16* __restrict__ kv,
                                                       h16* __restrict__ out, int N, int J, int heads, float scale_log2) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  __shared__ __align__(16) h16 sK[AT_BK][AT_D + AT_PAD];
  __shared__ __align__(16) h16 sV[AT_BK][AT_D + AT_PAD];
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * AT_BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;  // mma fragment coordinates
  const int row_a = q0 + warp * 16 + gid;     // this thread's first query row (second is +8)

  // Q fragments (A operand, 16 x 64 per warp): 4 k-steps x 4 regs, loaded straight from global
  uint32_t qa[4][4];
  const h16* qb = q + (long)b * N * ldq + head * AT_D;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int c = ks * 16 + tig * 2;
    const int r0 = row_a, r1 = row_a + 8;
    qa[ks][0] = r0 < N ? *reinterpret_cast<const uint32_t*>(qb + (long)r0 * ldq + c) : 0u;
    qa[ks][1] = r1 < N ? *reinterpret_cast<const uint32_t*>(qb + (long)r1 * ldq + c) : 0u;
    qa[ks][2] = r0 < N ? *reinterpret_cast<const uint32_t*>(qb + (long)r0 * ldq + c + 8) : 0u;
    qa[ks][3] = r1 < N ? *reinterpret_cast<const uint32_t*>(qb + (long)r1 * ldq + c + 8) : 0u;
  }

  float o_acc[8][4];  // 16 x 64 output: 8 n-tiles of 8 columns
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  const h16* kvb = kv + (long)b * J * 128;
  for (int j0 = 0; j0 < J; j0 += AT_BK) {
    __syncthreads();
    // stage K and V tiles: 64 rows x (8 + 8) 16-byte vectors
    for (int i = threadIdx.x; i < AT_BK * 16; i += 128) {
      const int r = i >> 4, g = i & 15;
      int4 v = make_int4(0, 0, 0, 0);
      if (j0 + r < J) v = *reinterpret_cast<const int4*>(kvb + (long)(j0 + r) * 128 + g * 8);
      if (g < 8)
        *reinterpret_cast<int4*>(&sK[r][g * 8]) = v;
      else
        *reinterpret_cast<int4*>(&sV[r][(g - 8) * 8]) = v;
    }
    __syncthreads();

    // S = Q K^T : 16 x 64 per warp = 8 n-tiles
    float s_acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s_acc[nt][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bfrag[2];
        // B (col-major k x n): element (k, n) = K[n][k]; thread holds k = ks*16 + tig*2 (+1), (+8,+9) for n = nt*8 + gid
        bfrag[0] = *reinterpret_cast<const uint32_t*>(&sK[nt * 8 + gid][ks * 16 + tig * 2]);
        bfrag[1] = *reinterpret_cast<const uint32_t*>(&sK[nt * 8 + gid][ks * 16 + tig * 2 + 8]);
        mma_f16_16816(s_acc[nt], qa[ks], bfrag);
      }
    }
    // online softmax (rows gid and gid+8); logits scaled into log2 domain
    float m_new[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = j0 + nt * 8 + tig * 2 + (j & 1);
        float v = s_acc[nt][j] * scale_log2;
        if (col >= J) v = -INFINITY;
        s_acc[nt][j] = v;
        m_new[j >> 1] = fmaxf(m_new[j >> 1], v);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
    }
    float corr[2], l_add[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) corr[r] = exp2f(m_run[r] - m_new[r]);
    uint32_t pa[4][4];  // P as A fragments for P @ V: 16 x 64 keys = 4 k-steps
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float pv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pv[j] = exp2f(s_acc[nt][j] - m_new[j >> 1]);
        l_add[j >> 1] += pv[j];
      }
      // accumulator layout (row gid: cols 2*tig,2*tig+1 ; row gid+8: same) maps onto the A fragment of k-step nt/2
      const int ks = nt >> 1, hi = nt & 1;
      pa[ks][hi * 2 + 0] = pack_h16x2(pv[0], pv[1]);
      pa[ks][hi * 2 + 1] = pack_h16x2(pv[2], pv[3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] = l_run[r] * corr[r] + l_add[r];
      m_run[r] = m_new[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o_acc[nt][0] *= corr[0];
      o_acc[nt][1] *= corr[0];
      o_acc[nt][2] *= corr[1];
      o_acc[nt][3] *= corr[1];
    }
    // O += P V : B operand element (k = key, n = d) = V[key][d]
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int k0 = ks * 16 + tig * 2;
        const int n = nt * 8 + gid;
        uint32_t bfrag[2];
        h162 t0, t1;
        t0.x = sV[k0][n];
        t0.y = sV[k0 + 1][n];
        t1.x = sV[k0 + 8][n];
        t1.y = sV[k0 + 9][n];
        bfrag[0] = *reinterpret_cast<uint32_t*>(&t0);
        bfrag[1] = *reinterpret_cast<uint32_t*>(&t1);
        mma_f16_16816(o_acc[nt], pa[ks], bfrag);
      }
    }
  }
  // finalize: row sums live in 4 lanes (tig), reduce
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  h16* ob = out + (long)b * N * heads * AT_D + head * AT_D;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + tig * 2;
    if (row_a < N) *reinterpret_cast<uint32_t*>(ob + (long)row_a * heads * AT_D + c) = pack_h16x2(o_acc[nt][0] * inv0, o_acc[nt][1] * inv0);
    if (row_a + 8 < N)
      *reinterpret_cast<uint32_t*>(ob + (long)(row_a + 8) * heads * AT_D + c) = pack_h16x2(o_acc[nt][2] * inv1, o_acc[nt][3] * inv1);
  }
}

// ------------------------------------------------------------------------------------------------ cross attention, J <= 64
// warp = one head, lane = one token; K/V of all heads staged in smem as fp32 pairs -> broadcast reads.
__global__ void __launch_bounds__(256) attn_cross_kernel(const h16* __restrict__ q, long ldq, const float* __restrict__ kv,
                                                         const float* __restrict__ null_kv, h16* __restrict__ out, int N, int Jc,
                                                         int heads, float scale) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  extern __shared__ float skv[];  // [J][heads][2][64]
  const int b = blockIdx.y;
  const int J = Jc + 1;
  const int HD = heads * 64;
  for (int i = threadIdx.x; i < J * heads * 128; i += blockDim.x) {
    const int d = i & 63, kvsel = (i >> 6) & 1, h = (i >> 7) % heads, j = i / (heads * 128);
    float v;
    if (j == 0)
      v = null_kv[kvsel * 64 + d];
    else
      v = kv[((long)b * Jc + (j - 1)) * 2 * HD + kvsel * HD + h * 64 + d];
    skv[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tokens_per_block = 32 * (blockDim.x / 32 / heads);
  const int tsub = warp / heads, h = warp % heads;
  const int n = blockIdx.x * tokens_per_block + tsub * 32 + lane;
  if (n >= N) return;
  float qv[64];
  const h16* qp = q + ((long)b * N + n) * ldq + h * 64;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    int4 raw = *reinterpret_cast<const int4*>(qp + g * 8);
    h16x8_to_float(*reinterpret_cast<h16x8*>(&raw), qv + g * 8);
  }
  float m = -INFINITY, l = 0.f;
  float acc[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
  for (int j = 0; j < J; ++j) {
    const float* kp = skv + ((long)(j * heads + h) * 2) * 64;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) s = fmaf(qv[d], kp[d], s);
    s *= scale;
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn), pj = __expf(s - mn);
    l = l * corr + pj;
    const float* vp = kp + 64;
#pragma unroll
    for (int d = 0; d < 64; ++d) acc[d] = fmaf(pj, vp[d], acc[d] * corr);
    m = mn;
  }
  const float inv = 1.f / l;
  h16* op = out + ((long)b * N + n) * HD + h * 64;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[g * 8 + j] * inv;
    *reinterpret_cast<h16x8*>(op + g * 8) = float_to_h16x8(v);
  }
}


// ------------------------------------------------------------------------------------------------ small fp32 attention
// PerceiverResampler attention of the text-conditioning tower (36 latent queries x <= 512 keys, once per sample() call):
// one warp per (b, head, query); lanes stride over the keys with a private online softmax, merged by shuffles.
__global__ void __launch_bounds__(128) attn_small_f32_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                             float* __restrict__ out, int Nq, int J, int heads, float scale) {
  const int lane = threadIdx.x & 31;
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int HD = heads * 64;
  const int b = blockIdx.z;
  const long local = wid;
  if (local >= (long)Nq * heads) return;
  const int h = (int)(local % heads), n = (int)(local / heads);
  const float* qp = q + ((long)b * Nq + n) * HD + h * 64;
  float qv[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) qv[d] = qp[d] * scale;
  float m = -INFINITY, l = 0.f, acc[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
  for (int j = lane; j < J; j += 32) {
    const float* kp = kv + ((long)b * J + j) * 2 * HD + h * 64;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) s = fmaf(qv[d], kp[d], s);
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn), pj = __expf(s - mn);
    l = l * corr + pj;
    const float* vp = kp + HD;
#pragma unroll
    for (int d = 0; d < 64; ++d) acc[d] = fmaf(pj, vp[d], acc[d] * corr);
    m = mn;
  }
  // merge the 32 lanes' partial softmaxes
  const float M = warp_max(m);
  const float f = (m == -INFINITY) ? 0.f : __expf(m - M);
  const float L = warp_sum(l * f);
  float* op = out + ((long)b * Nq + n) * HD + h * 64;
#pragma unroll
  for (int d = 0; d < 64; ++d) {
    const float v = warp_sum(acc[d] * f);
    if (lane == (d & 31)) op[d] = v / L;
  }
}

__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float a, float b, float* __restrict__ out, long n) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = a * x[i] + b * y[i];
}

}  // namespace

extern "C" int kd_linear_small(const float* x, int M, int K, long ldx, const float* w, const float* bias, float* y, int N, long ldy,
                               int pre_act, int post_act, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && w && y && M > 0 && K > 0 && N > 0 && M <= 65536, "kd_linear_small: bad argument (M=%d K=%d N=%d)", M, K, N);
  const int row_groups = kd_ceil_div(M, 8) < 64 ? kd_ceil_div(M, 8) : 64;
  dim3 grid(kd_ceil_div(N, 8), row_groups);
  KD_CUDA(kd_launch(linear_small_kernel<8>, dim3(grid), dim3(256), 0, stream, x, M, K, ldx, w, bias, y, N, ldy, pre_act, post_act));
  return KD_OK;
}

extern "C" int kd_sinu_emb(const float* t, const float* weights, int B, int half, float* out, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(t && weights && out && B > 0 && half > 0, "kd_sinu_emb: bad argument");
  const int total = B * (2 * half + 1);
  KD_CUDA(kd_launch(sinu_emb_kernel, dim3(kd_ceil_div(total, 128)), dim3(128), 0, stream, t, weights, B, half, out));
  return KD_OK;
}

extern "C" int kd_kv_assemble(const void* qkv, long ld, int kv_col, const float* ctx_kv, int Jc, const float* null_kv, void* kv_out,
                              int B, int N, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(qkv && null_kv && kv_out && B > 0 && N > 0 && Jc >= 0 && (Jc == 0 || ctx_kv), "kd_kv_assemble: bad argument");
  KD_REQUIRE(ld % 8 == 0 && kv_col % 8 == 0, "kd_kv_assemble: ld / kv_col must be multiples of 8");
  const long total = (long)(Jc + 1 + N) * 16;
  KD_CUDA(kd_launch(kv_assemble_kernel, dim3((unsigned)((total + 255) / 256), B), dim3(256), 0, stream, reinterpret_cast<const h16*>(qkv), ld, kv_col, ctx_kv, Jc, null_kv, reinterpret_cast<h16*>(kv_out), N));
  return KD_OK;
}

extern "C" int kd_attn_mqa(const void* q, long ldq, const void* kv, void* out, int B, int N, int J, int heads, float scale,
                           kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(q && kv && out && B > 0 && N > 0 && J > 0 && heads > 0, "kd_attn_mqa: bad argument");
  KD_REQUIRE(ldq % 8 == 0, "kd_attn_mqa: ldq must be a multiple of 8");
  dim3 grid(kd_ceil_div(N, AT_BQ), heads, B);
  KD_CUDA(kd_launch(attn_mqa_kernel, dim3(grid), dim3(128), 0, stream, reinterpret_cast<const h16*>(q), ldq, reinterpret_cast<const h16*>(kv), reinterpret_cast<h16*>(out), N, J, heads, scale * 1.4426950408889634f));
  return KD_OK;
}

extern "C" int kd_attn_cross(const void* q, long ldq, const float* kv, const float* null_kv, void* out, int B, int N, int Jc,
                             int heads, float scale, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(q && kv && null_kv && out && B > 0 && N > 0 && Jc > 0, "kd_attn_cross: bad argument");
  KD_REQUIRE(heads == 8 || heads == 4 || heads == 2 || heads == 1, "kd_attn_cross: heads must divide 8 (got %d)", heads);
  KD_REQUIRE(ldq % 8 == 0, "kd_attn_cross: ldq must be a multiple of 8");
  const size_t smem = (size_t)(Jc + 1) * heads * 128 * sizeof(float);
  KD_REQUIRE(smem <= 200 * 1024, "kd_attn_cross: %d context tokens exceed the shared-memory budget", Jc);
  static bool configured = false;
  if (!configured) {
    KD_CUDA(cudaFuncSetAttribute(attn_cross_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  const int tokens_per_block = 32 * (8 / heads);
  dim3 grid(kd_ceil_div(N, tokens_per_block), B);
  KD_CUDA(kd_launch(attn_cross_kernel, dim3(grid), dim3(256), smem, stream, reinterpret_cast<const h16*>(q), ldq, kv, null_kv, reinterpret_cast<h16*>(out), N, Jc, heads, scale));
  return KD_OK;
}

extern "C" int kd_attn_small_f32(const float* q, const float* kv, float* out, int B, int Nq, int J, int heads, float scale,
                                 kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(q && kv && out && B > 0 && Nq > 0 && J > 0 && heads > 0, "kd_attn_small_f32: bad argument");
  const long warps = (long)Nq * heads;
  dim3 grid((unsigned)((warps + 3) / 4), 1, B);
  attn_small_f32_kernel<<<grid, 128, 0, stream>>>(q, kv, out, Nq, J, heads, scale);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_axpby(const float* x, const float* y, float a, float b, float* out, long n, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && y && out && n > 0, "kd_axpby: bad argument");
  long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  axpby_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, y, a, b, out, n);
  KD_LAUNCH_CHECK();
  return KD_OK;
}
