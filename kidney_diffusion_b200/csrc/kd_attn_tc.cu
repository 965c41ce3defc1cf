// Multi-query self attention on 5th-gen tensor cores (imagen-pytorch Attention.forward: one shared 64-d K/V head, null k/v and
// optional context k/v prepended; N <= 4096 tokens at the 64^2 level of the base / SR UNets).
//
//   one CTA = 128 queries of one (image, head);   keys in tiles of 128:
//     S  = Q K^T          tcgen05.mma  M 128, N 128, K 64   -> TMEM columns [0, 128)       Q, K tiles by TMA (K-major, 128B swizzle)
//     P  = exp2((S - m) c)  softmax warps: tcgen05.ld, online max / sum in registers, fp16 P -> shared memory (K-major A operand)
//     PV = P V            tcgen05.mma  M 128, N 64,  K 128  -> TMEM columns [128, 192)     V^T tiles by TMA (kd_kv_transpose_v)
//     O  = O * corr + PV    in registers (one query row per thread): no rescaling of an accumulator in TMEM
//   192 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax (one TMEM lane quarter each); 97 KB of shared
//   memory and 256 TMEM columns per CTA, so two CTAs share an SM and one's MMAs overlap the other's softmax.
// Replaces the mma.sync flash kernel (attn_mqa_kernel in kd_cond_attn.cu) for N >= 256.
#include <cuda.h>
#include <mutex>

#include "kd_common.cuh"
#include "kd_tc.cuh"

namespace {

constexpr int TA_BQ = 128, TA_BK = 128, TA_D = 64;
constexpr int TA_THREADS = 192;
constexpr int TA_TILE_BYTES = 128 * 128;  // 128 rows x 64 h16
constexpr int TA_SMEM = TA_TILE_BYTES /*Q*/ + 2 * TA_TILE_BYTES /*K ring*/ + TA_TILE_BYTES /*V^T: 2 blocks of 64 x 64*/ +
                        2 * TA_TILE_BYTES /*P: 2 blocks of 128 x 64*/ + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// One 128-key tile of one query row: tile max -> running max, P = exp2((S - m) c) as fp16 into the swizzled A-operand tile.
// MASK: only the last tile of a sequence has fewer than 128 valid keys; the full tiles run without per-element predicates.
template <bool MASK>
__device__ __forceinline__ void softmax_tile(uint32_t t_row, uint32_t p_s, int r, int valid, float scale_log2, float& m_run, float& corr,
                                             float& lsum) {
  float tmax = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; c += 2) {  // two 32-column loads in flight per wait
    uint32_t sv[2][32];
    tmem_ld32(t_row + (uint32_t)(c * 32), sv[0]);
    tmem_ld32(t_row + (uint32_t)(c * 32 + 32), sv[1]);
    tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (!MASK || (c + u) * 32 + j < valid) tmax = fmaxf(tmax, __uint_as_float(sv[u][j]));
  }
  const float m_new = fmaxf(m_run, tmax);
  const float mc = m_new * scale_log2;
  corr = ex2_fast((m_run - m_new) * scale_log2);  // m_run = -inf on the first tile: exp2(-inf) = 0
  m_run = m_new;
  lsum = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t sv[32];
    tmem_ld32(t_row + (uint32_t)(c * 32), sv);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float p0 = ex2_fast(fmaf(__uint_as_float(sv[j]), scale_log2, -mc));
      float p1 = ex2_fast(fmaf(__uint_as_float(sv[j + 1]), scale_log2, -mc));
      if (MASK) {
        if (c * 32 + j >= valid) p0 = 0.f;
        if (c * 32 + j + 1 >= valid) p1 = 0.f;
      }
      lsum += p0 + p1;
      pk[j >> 1] = pack_h16x2(p0, p1);
    }
    // P as a K-major, 128B-swizzled A operand: block = c / 2 (64 keys each), 16-byte chunk = (c & 1) * 4 + q
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t chunk = (uint32_t)((c & 1) * 4 + q);
      const uint32_t dst = p_s + (uint32_t)(c >> 1) * TA_TILE_BYTES + (uint32_t)r * 128u + ((chunk ^ ((uint32_t)r & 7u)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]),
                   "r"(pk[4 * q + 3])
                   : "memory");
    }
  }
}

__global__ void __launch_bounds__(TA_THREADS, 2)
attn_mqa_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_vt, h16* __restrict__ out, int N, int J, int heads, float scale_log2) {
  constexpr uint32_t IDESC_S = (1u << 4) | ((uint32_t)(TA_BK >> 3) << 17) | ((uint32_t)(TA_BQ >> 4) << 24);
  constexpr uint32_t IDESC_PV = (1u << 4) | ((uint32_t)(TA_D >> 3) << 17) | ((uint32_t)(TA_BQ >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t q_s = base, k_s = base + TA_TILE_BYTES, vt_s = base + 3 * TA_TILE_BYTES, p_s = base + 4 * TA_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + 6 * TA_TILE_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* k_empty = bars + 3;     // 2
  uint64_t* v_full = bars + 5;      // 1
  uint64_t* v_empty = bars + 6;     // 1
  uint64_t* s_full = bars + 7;      // 1
  uint64_t* p_ready = bars + 8;     // 1 (count 4: one lane per softmax warp)
  uint64_t* pv_full = bars + 9;     // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * TA_BQ;
  const int T = (J + TA_BK - 1) / TA_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_vt);
    mbar_init(smem_u32(q_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1);
      mbar_init(smem_u32(&k_empty[s]), 1);
    }
    mbar_init(smem_u32(v_full), 1);
    mbar_init(smem_u32(v_empty), 1);
    mbar_init(smem_u32(s_full), 1);
    mbar_init(smem_u32(p_ready), 4);
    mbar_init(smem_u32(pv_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  kd_pdl_wait();
  kd_pdl_trigger();

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      mbar_expect_tx(smem_u32(q_full), TA_TILE_BYTES);
      tma_load_3d(q_s, &map_q, smem_u32(q_full), head * TA_D, q0, b);
      for (int t = 0; t < T; ++t) {
        const int s = t & 1;
        mbar_wait_relaxed(smem_u32(&k_empty[s]), (((uint32_t)t >> 1) & 1u) ^ 1u);
        mbar_expect_tx(smem_u32(&k_full[s]), TA_TILE_BYTES);
        tma_load_3d(k_s + s * TA_TILE_BYTES, &map_k, smem_u32(&k_full[s]), 0, t * TA_BK, b);
        mbar_wait_relaxed(smem_u32(v_empty), ((uint32_t)t & 1u) ^ 1u);
        mbar_expect_tx(smem_u32(v_full), TA_TILE_BYTES);
        tma_load_3d(vt_s, &map_vt, smem_u32(v_full), t * TA_BK, 0, b);
        tma_load_3d(vt_s + TA_TILE_BYTES / 2, &map_vt, smem_u32(v_full), t * TA_BK + 64, 0, b);
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (warp-convergent, one elected lane issues)
    mbar_wait(smem_u32(q_full), 0);
    for (int t = 0; t < T; ++t) {
      const int s = t & 1;
      mbar_wait(smem_u32(&k_full[s]), ((uint32_t)t >> 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_desc = make_sw128_desc(q_s), b_desc = make_sw128_desc(k_s + s * TA_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < TA_D / 16; ++k) umma_f16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), IDESC_S, k != 0 ? 1u : 0u);
        umma_commit(smem_u32(&k_empty[s]));
        umma_commit(smem_u32(s_full));
      }
      __syncwarp();
      mbar_wait(smem_u32(p_ready), (uint32_t)t & 1u);
      mbar_wait(smem_u32(v_full), (uint32_t)t & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t a_desc = make_sw128_desc(p_s + kb * TA_TILE_BYTES), b_desc = make_sw128_desc(vt_s + kb * (TA_TILE_BYTES / 2));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + 128, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), IDESC_PV, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(v_empty));
        umma_commit(smem_u32(pv_full));
      }
      __syncwarp();
    }
  } else {
    // ================================================================ softmax + output: one query row per thread
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f;
    float o[TA_D];
#pragma unroll
    for (int j = 0; j < TA_D; ++j) o[j] = 0.f;
    for (int t = 0; t < T; ++t) {
      mbar_wait(smem_u32(s_full), (uint32_t)t & 1u);
      tc_fence_after();
      const int valid = min(TA_BK, J - t * TA_BK);  // keys of this tile that exist (>= 1)
      float corr, lsum;
      if (valid == TA_BK) softmax_tile<false>(t_row, p_s, r, valid, scale_log2, m_run, corr, lsum);   // full tile: no per-element predicates
      else softmax_tile<true>(t_row, p_s, r, valid, scale_log2, m_run, corr, lsum);
      l_run = fmaf(l_run, corr, lsum);
#pragma unroll
      for (int j = 0; j < TA_D; ++j) o[j] *= corr;
      tc_fence_before();
      fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(p_ready));
      mbar_wait(smem_u32(pv_full), (uint32_t)t & 1u);
      tc_fence_after();
      {
        uint32_t pv[2][32];
        tmem_ld32(t_row + 128u, pv[0]);
        tmem_ld32(t_row + 160u, pv[1]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j = 0; j < 32; ++j) o[c * 32 + j] += __uint_as_float(pv[c][j]);
      }
      tc_fence_before();
    }
    if (q0 + r < N) {
      const float inv = 1.0f / l_run;
      h16* dst = out + ((long)b * N + q0 + r) * ((long)heads * TA_D) + head * TA_D;
#pragma unroll
      for (int g = 0; g < TA_D / 8; ++g) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = o[g * 8 + j] * inv;
        *reinterpret_cast<h16x8*>(dst + g * 8) = float_to_h16x8(v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// vT[b][d][j] = kv[b][j][64 + d] for j < J, 0 for J <= j < Jpad: the K-major B operand of the PV product
__global__ void kv_transpose_v_kernel(const h16* __restrict__ kv, h16* __restrict__ vt, int J, int Jpad) {
  kd_pdl_wait();
  kd_pdl_trigger();
  __shared__ h16 tile[64][TA_D + 2];
  const int b = blockIdx.y, j0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * TA_D; i += blockDim.x) {
    const int jj = i / TA_D, d = i % TA_D;
    tile[jj][d] = (j0 + jj < J) ? kv[((long)b * J + j0 + jj) * 128 + 64 + d] : __float2half(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * TA_D; i += blockDim.x) {
    const int d = i / 64, jj = i % 64;
    if (j0 + jj < Jpad) vt[((long)b * TA_D + d) * Jpad + j0 + jj] = tile[jj][d];
  }
}

}  // namespace

extern "C" int kd_attn_vt_elems(int B, int J) { return B * TA_D * (((J + 7) / 8) * 8); }

extern "C" int kd_attn_mqa_tc(const void* q, long ldq, const void* kv, void* vt_scratch, void* out, int B, int N, int J, int heads, float scale,
                              kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(q && kv && vt_scratch && out && B > 0 && N > 0 && J > 0 && heads > 0, "kd_attn_mqa_tc: bad argument");
  KD_REQUIRE(ldq % 8 == 0 && ldq >= (long)heads * TA_D, "kd_attn_mqa_tc: ldq must be a multiple of 8 and >= heads * 64");
  const int Jpad = ((J + 7) / 8) * 8;
  KD_CUDA(kd_launch(kv_transpose_v_kernel, dim3((Jpad + 63) / 64, B), dim3(256), 0, stream, reinterpret_cast<const h16*>(kv),
                    reinterpret_cast<h16*>(vt_scratch), J, Jpad));
  CUtensorMap mq, mk, mv;
  {
    const uint64_t dims[3] = {(uint64_t)ldq, (uint64_t)N, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)N * ldq * 2};
    const uint32_t box[3] = {64u, 128u, 1u};
    int rc = kd_encode_tiled_h16(&mq, q, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {128ull, (uint64_t)J, (uint64_t)B};
    const uint64_t str[2] = {256ull, (uint64_t)J * 256};
    const uint32_t box[3] = {64u, 128u, 1u};
    int rc = kd_encode_tiled_h16(&mk, kv, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Jpad, (uint64_t)TA_D, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)Jpad * 2, (uint64_t)TA_D * Jpad * 2};
    const uint32_t box[3] = {64u, 64u, 1u};
    int rc = kd_encode_tiled_h16(&mv, vt_scratch, 3, dims, str, box);
    if (rc) return rc;
  }
  static bool configured = false;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!configured) {
      KD_CUDA(cudaFuncSetAttribute(attn_mqa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
      configured = true;
    }
  }
  KD_CUDA(kd_launch(attn_mqa_tc_kernel, dim3((N + TA_BQ - 1) / TA_BQ, heads, B), dim3(TA_THREADS), TA_SMEM, stream, mq, mk, mv,
                    reinterpret_cast<h16*>(out), N, J, heads, scale * 1.4426950408889634f));
  return KD_OK;
}
