// K1: implicit-GEMM convolution on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands fed by TMA).
//
//   M = B*H*W output pixels, N = Cout, K = taps * (Ca + Cb); A = NHWC fp16 activations (one or two concatenated sources),
//   B = packed weights [Cout, K] (K-major, K ordered (tap, channel)); D = fp32 accumulator in tensor memory.
//   Both operands land in shared memory through TMA with the 128-byte swizzle, i.e. as canonical K-major SWIZZLE_128B UMMA
//   operands; out-of-bounds TMA coordinates are zero-filled, which implements the convolution's zero padding.
//
// Three kernels share the PTX wrappers of kd_tc.cuh:
//   conv_gemm_kernel       128 x BN tile per CTA, 192 threads, 2 CTAs/SM            (Cout < 128: small models / tests)
//   conv_gemm_pair_kernel  cta_group::2, UMMA M = 256, persistent, tap-loop A loads  (1x1, 2x2-stride-2, tiny images)
//   conv_gemm_halo_kernel  as the pair kernel, but all nine taps of a 3x3 read ONE halo tile per 64-channel chunk, and the
//                          GroupNorm + scale/shift + SiLU of the consuming Block can be applied to that tile in shared
//                          memory (PRE)                                              (every 3x3 on >= 16 x 8 images)
// and one epilogue (pair_epilogue_role): TMEM -> registers -> swizzled staging -> one TMA store per 128 x 64 item, with
// fused bias / activation / gated residual (addend tile TMA-loaded one item ahead) / GroupNorm octet statistics /
// GlobalContext logits.  MMA issue loops are warp-convergent with one elected lane (uniform-register descriptors).
#include <cuda.h>
#include <mutex>

#include "kd_common.cuh"
#include "kd_tc.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 h16 = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 192;

struct ConvParams {
  int mode, B, H, W, Ca, Cb, Cout, ksize, act, out_mode, out_f32, addend_f32;
  int TW, TH, TB;
  int tiles_w, tiles_h, tiles_b, n_tiles;
  int chunks_a, chunks_per_tap, num_kb;
  const float* bias;
  const void* addend;
  const float* addend_scale;
  void* out;
  float* stats;  // optional per-(row group, channel octet) {sum, sumsq} of the stored output (fused GroupNorm statistics)
  const float* logit_w;  // optional GlobalContext to_k weight [Cout]: the epilogue also emits per-pixel partial dot products
  float* logit_parts;    // [Cout / 64][B*H*W] fp32, one partial per 64-column group (summed in fixed order by kd_gca_pool)
  const float2* pre_coef;  // optional [B][Ca+Cb] {A, B}: the A operand becomes SiLU(A * x + B) (fused GroupNorm apply, halo kernel)
  int kb_per_split;        // split-K: k-blocks per split (blockIdx.y selects the split); 0 = the whole K range in one CTA
  float* splitk_ws;        // split-K: fp32 partial tiles [split][tile][128][BN]
};

// ------------------------------------------------------------------------------------------------ kernel
// SPLIT: the CTA accumulates only k-blocks [blockIdx.y * kb_per_split, ...) and stores its raw fp32 accumulator tile to the
// split-K workspace; splitk_finish_kernel adds the splits in fixed order and applies the epilogue.
template <int BN, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 2)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_w, const ConvParams p) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  // instruction descriptor: D = fp32 (bit 4), A = B = h16 (bits 7, 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // control block after the operand ring
  uint8_t* ctrl = smem_gen + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* bias_smem = reinterpret_cast<float*>(ctrl + 128);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int n_tile = blockIdx.x % p.n_tiles;
  int m_tile = blockIdx.x / p.n_tiles;
  const int tile_w = m_tile % p.tiles_w;
  m_tile /= p.tiles_w;
  const int tile_h = m_tile % p.tiles_h;
  const int tile_b = m_tile / p.tiles_h;
  const int w0 = tile_w * p.TW, h0 = tile_h * p.TH, b0 = tile_b * p.TB;
  const int n0 = n_tile * BN;
  const int kb_lo = SPLIT ? (int)blockIdx.y * p.kb_per_split : 0;
  const int kb_hi = SPLIT ? min(p.num_kb, kb_lo + p.kb_per_split) : p.num_kb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    if (p.Cb > 0) tma_prefetch_desc(&map_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
  if (warp >= 2) {
    for (int j = threadIdx.x - 64; j < BN; j += 128) {
      const int n = n0 + j;
      bias_smem[j] = (p.bias != nullptr && n < p.Cout) ? p.bias[n] : 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  kd_pdl_wait();     // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  kd_pdl_trigger();

  if (warp == 0) {
    // ================================================================ TMA producer (one elected lane)
    if (lane == 0) {
      const int pad = (p.mode == 0) ? (p.ksize >> 1) : 0;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        const int s = (kb - kb_lo) % STAGES;
        const uint32_t ph = ((kb - kb_lo) / STAGES) & 1;
        mbar_wait_relaxed(smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, STAGE_BYTES);
        const int tap = kb / p.chunks_per_tap;
        const int ch = kb - tap * p.chunks_per_tap;
        const bool src_b = ch >= p.chunks_a;
        const CUtensorMap* map = src_b ? &map_b : &map_a;
        const int c0 = (src_b ? (ch - p.chunks_a) : ch) * BK;
        const uint32_t a_dst = smem_base + s * STAGE_BYTES;
        if (p.mode == 1) {
          // input viewed as [B, H, 2(dy), W, 2C]: tap = dy*2 + dx selects the dx half of the 2C axis
          const int dy = tap >> 1, dx = tap & 1;
          const int C = src_b ? p.Cb : p.Ca;
          tma_load_5d(a_dst, map, fb, dx * C + c0, w0, dy, h0, b0);
        } else {
          const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
          tma_load_5d(a_dst, map, fb, c0, w0 + kx - pad, h0 + ky - pad, b0, 0);
        }
        tma_load_2d(a_dst + A_STAGE_BYTES, &map_w, fb, kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (warp-convergent loop, one elected lane issues)
    for (int kb = kb_lo; kb < kb_hi; ++kb) {
      const int s = (kb - kb_lo) % STAGES;
      const uint32_t ph = ((kb - kb_lo) / STAGES) & 1;
      mbar_wait(smem_u32(&full_bar[s]), ph);
      tc_fence_after();
      const uint32_t a_addr = smem_base + s * STAGE_BYTES;
      const uint64_t a_desc = make_sw128_desc(a_addr);
      const uint64_t b_desc = make_sw128_desc(a_addr + A_STAGE_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_f16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), IDESC, ((kb - kb_lo) | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[s]));  // frees the smem slot once these MMAs have read it
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(tmem_full_bar));  // accumulator complete
    __syncwarp();
  } else {
    // ================================================================ epilogue: TMEM -> registers -> global
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    const int tw = r % p.TW;
    const int th = (r / p.TW) % p.TH;
    const int tb = r / (p.TW * p.TH);
    const int b = b0 + tb, h = h0 + th, w = w0 + tw;
    const bool row_ok = (b < p.B) && (h < p.H) && (w < p.W);

    mbar_wait(smem_u32(tmem_full_bar), 0);
    tc_fence_after();

    const int Cq = p.Cout >> 2;  // channels per pixel-shuffle quadrant
#pragma unroll 1
    for (int chunk = 0; chunk < BN / 32; ++chunk) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chunk * 32), acc);
      tmem_ld_wait();
      const int nc = n0 + chunk * 32;
      if (SPLIT) {  // raw partial tile, every row (rows outside the image hold exact zeros: their A rows were zero-filled)
        float4* dst = reinterpret_cast<float4*>(p.splitk_ws + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * BM + r) * BN + chunk * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), __uint_as_float(acc[4 * j + 2]),
                               __uint_as_float(acc[4 * j + 3]));
        continue;
      }
      if (!row_ok || nc >= p.Cout) continue;

      long long out_off;   // element offset of column nc for this row
      long long add_off;
      if (p.out_mode == 1) {
        const int q4 = nc / Cq, c = nc - q4 * Cq;
        const int dy = q4 >> 1, dx = q4 & 1;
        out_off = (((long long)b * (2 * p.H) + (2 * h + dy)) * (2 * p.W) + (2 * w + dx)) * Cq + c;
      } else {
        out_off = (((long long)b * p.H + h) * p.W + w) * p.Cout + nc;
      }
      add_off = out_off;
      const float* gate = p.addend_scale ? p.addend_scale + (long long)b * p.Cout + nc : nullptr;

#pragma unroll
      for (int g = 0; g < 4; ++g) {  // 4 groups of 8 columns
        if (nc + g * 8 + 8 > p.Cout) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g * 8 + j]) + bias_smem[chunk * 32 + g * 8 + j];
        apply_act8(v, p.act);
        if (p.addend != nullptr) {
          float a[8];
          if (p.addend_f32) {
            const float4* ap = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.addend) + add_off + g * 8);
            float4 a0 = ap[0], a1 = ap[1];
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
          } else {
            h16x8 raw = *reinterpret_cast<const h16x8*>(reinterpret_cast<const h16*>(p.addend) + add_off + g * 8);
            h16x8_to_float(raw, a);
          }
          if (gate != nullptr) {
            const float4 g0 = *reinterpret_cast<const float4*>(gate + g * 8);
            const float4 g1 = *reinterpret_cast<const float4*>(gate + g * 8 + 4);
            a[0] *= g0.x; a[1] *= g0.y; a[2] *= g0.z; a[3] *= g0.w; a[4] *= g1.x; a[5] *= g1.y; a[6] *= g1.z; a[7] *= g1.w;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += a[j];
        }
        if (p.out_f32) {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off + g * 8);
          op[0] = make_float4(v[0], v[1], v[2], v[3]);
          op[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
          *reinterpret_cast<h16x8*>(reinterpret_cast<h16*>(p.out) + out_off + g * 8) = float_to_h16x8(v);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ================================================================================================ conv v2
// CTA-pair kernel (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN output tile with UMMA M = 256.
// Each CTA loads only its own 128 A rows and its own half (BN/2 rows) of the weight tile, and the pair's tensor cores
// read both halves -- this halves the per-SM operand feed from L2, which is what bounded the single-CTA kernel
// (32 KB per k-block at ~64 B/clk/SM = 512 cycles against 256 cycles of MMA).  CTAs are persistent (static round-robin
// tile scheduler) and the accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps
// the main loop of tile i+1.
//
// Barrier protocol (identical smem layout in both CTAs; "leader" = cluster rank 0):
//   full[s]      leader only, count 2: leader producer's arrive.expect_tx(bytes of BOTH CTAs) + peer producer's remote
//                arrive; all four TMA loads of a stage complete_tx on the leader's barrier (.cta_group::2 loads).
//   empty[s]     both CTAs, count 1: tcgen05.commit multicast from the leader's MMA thread.
//   tmem_full[a] both CTAs, count 1: commit multicast after the last k-block of a tile.
//   tmem_empty[a] leader only, count 16: one elected lane of each of the 8 epilogue warps of both CTAs (remote arrive).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_remote_arrive(uint32_t local_bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32  remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64  _, [remAddr32];\n\t"
      "}" ::"r"(local_bar),
      "r"(cta_rank)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2,
                                                int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(dst),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (count 1) on the barrier at this offset in BOTH CTAs of the pair once all prior MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

// TMA tensor store smem -> global (bulk async group), used by the coalesced epilogue
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void epi_barrier(int set) { asm volatile("bar.sync %0, 128;" ::"r"(set + 1) : "memory"); }  // 4 warps of one set

constexpr int NUM_THREADS2 = 384;  // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-7 and 8-11: two epilogue sets
constexpr int EPI_COLS = 64;                          // columns per staged group = one 128-byte swizzle row of h16
constexpr int EPI_STAGE_BYTES = BM * EPI_COLS * 2;    // 16 KB staging buffer per epilogue set (two sets)

// epilogue of one 32-column chunk of one accumulator row (shared by both kernels' store paths)
__device__ __forceinline__ void epilogue_store_chunk(const ConvParams& p, const uint32_t* acc, int nc, int b, int h, int w) {
  const int Cq = p.Cout >> 2;
  long long out_off;
  if (p.out_mode == 1) {
    const int q4 = nc / Cq, c = nc - q4 * Cq;
    const int dy = q4 >> 1, dx = q4 & 1;
    out_off = (((long long)b * (2 * p.H) + (2 * h + dy)) * (2 * p.W) + (2 * w + dx)) * Cq + c;
  } else {
    out_off = (((long long)b * p.H + h) * p.W + w) * p.Cout + nc;
  }
  const float* gate = p.addend_scale ? p.addend_scale + (long long)b * p.Cout + nc : nullptr;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (nc + g * 8 + 8 > p.Cout) break;
    float v[8];
    if (p.bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + nc + g * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + nc + g * 8 + 4));
      v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += __uint_as_float(acc[g * 8 + j]);
    apply_act8(v, p.act);
    if (p.addend != nullptr) {
      float a[8];
      if (p.addend_f32) {
        const float4* ap = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.addend) + out_off + g * 8);
        float4 a0 = ap[0], a1 = ap[1];
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      } else {
        h16x8 raw = *reinterpret_cast<const h16x8*>(reinterpret_cast<const h16*>(p.addend) + out_off + g * 8);
        h16x8_to_float(raw, a);
      }
      if (gate != nullptr) {
        const float4 g0 = *reinterpret_cast<const float4*>(gate + g * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gate + g * 8 + 4);
        a[0] *= g0.x; a[1] *= g0.y; a[2] *= g0.z; a[3] *= g0.w; a[4] *= g1.x; a[5] *= g1.y; a[6] *= g1.z; a[7] *= g1.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += a[j];
    }
    if (p.out_f32) {
      float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off + g * 8);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      *reinterpret_cast<h16x8*>(reinterpret_cast<h16*>(p.out) + out_off + g * 8) = float_to_h16x8(v);
    }
  }
}

// accumulator stages in TMEM (512 columns): BN = 256 -> 2, BN = 128 -> 4 (the MMA can run up to three tiles ahead of the
// epilogue, which at BN = 128 takes about as long per tile as the MMAs and otherwise alternates with them)
template <int BN>
struct AccStages {
  static constexpr uint32_t N = 512 / BN;
};

// ---- epilogue role of the CTA-pair kernels (warps 4..11): drains the double-buffered TMEM accumulator tile by tile.
// Two sets of 4 warps (one warp of each set per SM sub-partition): set 0 takes the even 64-column groups of a tile, set 1 the
// odd ones; an "item" is one such group: 128 rows x 64 columns, staged in a 16 KB 128B-swizzled smem tile and written with
// one TMA store.
// TADD: the h16 addend tile of every item is fetched by TMA straight into the (double-buffered) staging tile one item ahead
// and updated in place, instead of 8 row-strided LDG.128 per thread (32 distinct 128-byte lines per warp instruction: the
// LSU queue, not HBM, bounded the convolutions with a residual -- profiles/r01c ncu of the 1x1 res conv).
// Bias, gate rows and GlobalContext to_k weights of the whole tile width live in smem and are reloaded only when the
// tile's (n_tile, batch group) changes, which is rare in the persistent tile order.
// aux smem: bias[BN] + aux[2][BN] floats

template <int BN, bool TADD>
__device__ __forceinline__ void pair_epilogue_role(const ConvParams& p, const CUtensorMap* map_out_ptr, const CUtensorMap* map_add_ptr,
                                                   uint64_t* add_bar, const uint32_t tmem_base, uint64_t* tmem_full_bar,
                                                   uint64_t* tmem_empty_bar, const uint32_t epi_smem, float* epi_aux, const int warp,
                                                   const int lane, const uint32_t rank, const int cluster_id, const int num_clusters,
                                                   const int num_pair_tiles) {
  constexpr uint32_t ACC = AccStages<BN>::N;
  const CUtensorMap& map_out = *map_out_ptr;
  const int quarter = warp & 3;
  const int set = (warp - 4) >> 2;
  const bool issuer = (quarter == 0) && (lane == 0);  // this set's TMA thread
  const uint32_t stage0 = epi_smem + set * (TADD ? 2 : 1) * EPI_STAGE_BYTES;
  float* bias_all = epi_aux;        // [BN]
  float* aux_all = epi_aux + BN;    // [2][BN]: gate rows of the <= 2 batch images of a tile, or [0] = to_k weights
  const int r = quarter * 32 + lane;
  const int et = (warp - 4) * 32 + lane;  // 0..255 over both sets
  const int tw = r % p.TW;
  const int th = (r / p.TW) % p.TH;
  const int tb = r / (p.TW * p.TH);
  const bool gate_smem = (p.addend_scale != nullptr) && (p.TB <= 2);
  const int Cq = p.Cout >> 2;

  auto set_has_group = [&](int t2) { return (t2 % p.n_tiles) * BN + set * EPI_COLS < p.Cout; };
  auto load_addend_tile = [&](int t2, int g2, uint32_t buf) {  // issuer thread only
    int m2 = (t2 / p.n_tiles) * 2 + (int)rank;
    const int tile_w2 = m2 % p.tiles_w;
    m2 /= p.tiles_w;
    const int tile_h2 = m2 % p.tiles_h;
    const int tile_b2 = m2 / p.tiles_h;
    const uint32_t bar = smem_u32(&add_bar[set * 2 + buf]);
    mbar_expect_tx(bar, EPI_STAGE_BYTES);
    tma_load_5d(stage0 + buf * EPI_STAGE_BYTES, map_add_ptr, bar, (t2 % p.n_tiles) * BN + g2 * EPI_COLS, tile_w2 * p.TW, tile_h2 * p.TH,
                tile_b2 * p.TB, 0);
  };
  // (tn, gn): the next item of this set in visiting order (TADD: the one whose addend tile is requested ahead)
  int tn = cluster_id, gn = set;
  uint32_t item = 0;
  if (TADD && !p.out_f32) {
    while (tn < num_pair_tiles && !set_has_group(tn)) tn += num_clusters;
    if (issuer && tn < num_pair_tiles) load_addend_tile(tn, gn, 0);
  }

  int aux_key = -1;
  uint32_t tile_iter = 0;
  for (int t = cluster_id; t < num_pair_tiles; t += num_clusters, ++tile_iter) {
    const uint32_t as = tile_iter % ACC, aph = (tile_iter / ACC) & 1;
    const int n_tile = t % p.n_tiles;
    int m_tile = (t / p.n_tiles) * 2 + (int)rank;
    const int tile_w = m_tile % p.tiles_w;
    m_tile /= p.tiles_w;
    const int tile_h = m_tile % p.tiles_h;
    const int tile_b = m_tile / p.tiles_h;
    const int b = tile_b * p.TB + tb, h = tile_h * p.TH + th, w = tile_w * p.TW + tw;
    const bool row_ok = (b < p.B) && (h < p.H) && (w < p.W);
    const int n0 = n_tile * BN;

    if (!p.out_f32) {
      const int key = gate_smem ? tile_b * p.n_tiles + n_tile : n_tile;
      if (key != aux_key) {  // uniform over all 8 epilogue warps (both sets walk the same tiles)
        aux_key = key;
        asm volatile("bar.sync 3, 256;" ::: "memory");  // nobody still reads the old values
        for (int c = et; c < BN; c += 256) {
          const bool in = n0 + c < p.Cout;
          bias_all[c] = (p.bias != nullptr && in) ? __ldg(p.bias + n0 + c) : 0.f;
          if (gate_smem) {
            const int bb = tile_b * p.TB;
            aux_all[c] = (in && bb < p.B) ? __ldg(p.addend_scale + (long long)bb * p.Cout + n0 + c) : 0.f;
            aux_all[BN + c] = (in && bb + 1 < p.B) ? __ldg(p.addend_scale + (long long)(bb + 1) * p.Cout + n0 + c) : 0.f;
          } else if (p.logit_w != nullptr) {
            aux_all[c] = in ? __ldg(p.logit_w + n0 + c) : 0.f;
          }
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
      }
    }

    mbar_wait_relaxed(smem_u32(&tmem_full_bar[as]), aph);
    tc_fence_after();
    if (p.out_f32) {
      // fp32 output (tests / small tensors): direct per-row stores
#pragma unroll 1
      for (int chunk = set; chunk < BN / 32; chunk += 2) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + as * BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chunk * 32), acc);
        tmem_ld_wait();
        const int nc = n0 + chunk * 32;
        if (row_ok && nc < p.Cout) epilogue_store_chunk(p, acc, nc, b, h, w);
      }
    } else {
      const bool tile_in_range = tile_b < p.tiles_b;
#pragma unroll 1
      for (int g = set; g < BN / EPI_COLS; g += 2) {
        const int nc0 = n0 + g * EPI_COLS;
        if (nc0 >= p.Cout) break;  // uniform across the set
        const uint32_t buf = TADD ? (item & 1u) : 0u;
        const uint32_t stage = stage0 + buf * EPI_STAGE_BYTES;
        if (!TADD) {
          // the TMA store that last read this set's staging buffer must have finished reading it
          if (issuer) tma_store_wait_read<0>();
          epi_barrier(set);
        } else {
          // advance (tn, gn) to the item after this one and request its addend tile into the other buffer
          gn += 2;
          if (!(gn < BN / EPI_COLS && n0 + gn * EPI_COLS < p.Cout)) {
            gn = set;
            for (tn += num_clusters; tn < num_pair_tiles && !set_has_group(tn); tn += num_clusters) {
            }
          }
          epi_barrier(set);  // every thread of the set is done with the previous item (its statistics read the other buffer)
          if (issuer && tn < num_pair_tiles) {
            tma_store_wait_read<0>();  // ... and so is the bulk store that drained it
            load_addend_tile(tn, gn, buf ^ 1u);
          }
          mbar_wait(smem_u32(&add_bar[set * 2 + buf]), (item >> 1) & 1u);  // this item's addend tile has landed
        }
        ++item;
        const float* bias_s = bias_all + g * EPI_COLS;
        float lacc = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t acc[32];
          tmem_ld32(tmem_base + as * BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * EPI_COLS + half * 32), acc);
          tmem_ld_wait();
          const int nc = nc0 + half * 32;
          const float* gate = nullptr;
          if (p.addend_scale != nullptr)
            gate = gate_smem ? (aux_all + tb * BN + g * EPI_COLS + half * 32)
                             : (p.addend_scale + (long long)(row_ok ? b : 0) * p.Cout + nc);
          const float* lw = aux_all + g * EPI_COLS + half * 32;
          long long add_off = 0;
          if (!TADD && p.addend != nullptr && row_ok) {
            if (p.out_mode == 1) {
              const int q4 = nc / Cq, c = nc - q4 * Cq;
              add_off = (((long long)b * (2 * p.H) + (2 * h + (q4 >> 1))) * (2 * p.W) + (2 * w + (q4 & 1))) * Cq + c;
            } else {
              add_off = (((long long)b * p.H + h) * p.W + w) * p.Cout + nc;
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // 4 groups of 8 columns -> one 16-byte staging store each
            float v[8];
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + half * 32 + q * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + half * 32 + q * 8 + 4);
            v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __uint_as_float(acc[q * 8 + j]);
            apply_act8(v, p.act);
            const uint32_t chunk16 = (uint32_t)(half * 4 + q);
            const uint32_t dst = stage + (uint32_t)r * 128u + ((chunk16 ^ ((uint32_t)r & 7u)) << 4);
            if (p.addend != nullptr && row_ok && nc + q * 8 + 8 <= p.Cout) {
              float a[8];
              if (TADD) {
                int4 raw;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "r"(dst));
                h16x8_to_float(*reinterpret_cast<const h16x8*>(&raw), a);
              } else if (p.addend_f32) {
                const float4* ap = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.addend) + add_off + q * 8);
                float4 a0 = ap[0], a1 = ap[1];
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
              } else {  // h16 addend of a pixel-shuffle output (not on the UNet's path): plain row loads
                const int4 raw = ld_stream(reinterpret_cast<const h16*>(p.addend) + add_off + q * 8);
                h16x8_to_float(*reinterpret_cast<const h16x8*>(&raw), a);
              }
              if (gate != nullptr) {
                const float4 g0 = *reinterpret_cast<const float4*>(gate + q * 8);
                const float4 g1 = *reinterpret_cast<const float4*>(gate + q * 8 + 4);
                a[0] *= g0.x; a[1] *= g0.y; a[2] *= g0.z; a[3] *= g0.w; a[4] *= g1.x; a[5] *= g1.y; a[6] *= g1.z; a[7] *= g1.w;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += a[j];
            }
            if (!row_ok) {  // rows outside the image are clipped by the TMA store; zeros keep the fused statistics exact
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
            const h16x8 o8 = float_to_h16x8(v);
            const int4 ov = *reinterpret_cast<const int4*>(&o8);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(ov.x), "r"(ov.y), "r"(ov.z), "r"(ov.w) : "memory");
            if (p.logit_w != nullptr) {  // dot product with the ROUNDED values, as a pass over the stored tensor would see them
              float f[8];
              h16x8_to_float(o8, f);
              const float4 l0 = *reinterpret_cast<const float4*>(lw + q * 8);
              const float4 l1 = *reinterpret_cast<const float4*>(lw + q * 8 + 4);
              lacc += f[0] * l0.x + f[1] * l0.y + f[2] * l0.z + f[3] * l0.w + f[4] * l1.x + f[5] * l1.y + f[6] * l1.z + f[7] * l1.w;
            }
          }
        }
        if (p.logit_w != nullptr && row_ok)
          p.logit_parts[(long long)(nc0 / EPI_COLS) * ((long long)p.B * p.H * p.W) + ((long long)b * p.H + h) * p.W + w] = lacc;
        fence_proxy_async_smem();
        epi_barrier(set);
        if (issuer) {
          if (tile_in_range) {
            if (p.out_mode == 1) {
              const int q4 = nc0 / Cq, c0 = nc0 - q4 * Cq;
              tma_store_5d(&map_out, stage, (q4 & 1) * Cq + c0, tile_w * p.TW, q4 >> 1, tile_h * p.TH, tile_b * p.TB);
            } else {
              tma_store_5d(&map_out, stage, nc0, tile_w * p.TW, tile_h * p.TH, tile_b * p.TB, 0);
            }
          }
          tma_store_commit();
        }
        if (p.stats != nullptr && tile_in_range) {
          // fused GroupNorm statistics: per warp (32 rows) and channel octet, {sum, sumsq} of the ROUNDED values just
          // staged (exactly what a separate pass over the stored tensor would see); fixed order -> deterministic.
          // lane -> (octet = lane & 7, 8-row slice = lane >> 3); the 4 slices of a warp are merged by two shuffles.
          const int o = lane & 7, sl = lane >> 3;
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t row = (uint32_t)(quarter * 32 + sl * 8 + i);
            const uint32_t src = stage + row * 128u + (((uint32_t)o ^ (row & 7u)) << 4);
            int4 raw;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "r"(src));
            float f[8];
            h16x8_to_float(*reinterpret_cast<const h16x8*>(&raw), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              s1 += f[j];
              s2 = fmaf(f[j], f[j], s2);
            }
          }
          s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
          s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
          if (sl == 0 && nc0 + o * 8 < p.Cout) {
            const long long m_tile_lin = ((long long)tile_b * p.tiles_h + tile_h) * p.tiles_w + tile_w;
            float2* dstp = reinterpret_cast<float2*>(p.stats) + (m_tile_lin * 4 + quarter) * (p.Cout >> 3) + ((nc0 >> 3) + o);
            *dstp = make_float2(s1, s2);
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_remote_arrive(smem_u32(&tmem_empty_bar[as]), 0);  // accumulator stage drained (leader's barrier)
  }
  if (issuer) tma_store_wait_read<0>();  // staging smem must outlive the last bulk stores
}

template <int BN, int STAGES, bool TADD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
                      const __grid_constant__ CUtensorMap map_add, const ConvParams p, const int num_pair_tiles) {
  constexpr int EPI_BUFS = TADD ? 4 : 2;  // 16 KB staging tiles: one per epilogue set, two with the TMA-fed addend
  constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_HALF_BYTES;
  constexpr uint32_t ACC = AccStages<BN>::N;
  constexpr uint32_t TMEM_COLS = 512;  // ACC accumulator stages
  // instruction descriptor: D fp32, A = B = h16, K-major, N = BN, M = 256 (pair)
  constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_smem = smem_base + STAGES * STAGE_BYTES;  // swizzled output staging tiles (1024-aligned)
  uint8_t* ctrl = smem_gen + STAGES * STAGE_BYTES + EPI_BUFS * EPI_STAGE_BYTES;
  uint64_t* add_bar = reinterpret_cast<uint64_t*>(ctrl + 320);  // [set][buffer]: addend tile landed (TADD)
  float* epi_aux = reinterpret_cast<float*>(ctrl + 384);        // per epilogue set: bias[64] + gate[2][64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  cluster_sync_all();  // both CTAs are resident before the pair-wide TMEM allocation
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_out);
    if (TADD) tma_prefetch_desc(&map_add);
    if (p.Cb > 0) tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 2);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
#pragma unroll
    for (int a = 0; a < (int)ACC; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), 16);  // 8 epilogue warps x 2 CTAs
    }
    for (int a = 0; a < 4; ++a) mbar_init(smem_u32(&add_bar[a]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc_2sm(smem_u32(tmem_ptr_smem), TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  kd_pdl_wait();     // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  kd_pdl_trigger();

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs, one lane each)
    if (lane == 0) {
      const int pad = (p.mode == 0) ? (p.ksize >> 1) : 0;
      uint32_t it = 0;
      for (int t = cluster_id; t < num_pair_tiles; t += num_clusters) {
        const int n_tile = t % p.n_tiles;
        int m_tile = (t / p.n_tiles) * 2 + (int)rank;
        const int tile_w = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int tile_h = m_tile % p.tiles_h;
        const int tile_b = m_tile / p.tiles_h;  // may be == tiles_b for the odd tail: TMA zero-fills, epilogue masks
        const int w0 = tile_w * p.TW, h0 = tile_h * p.TH, b0 = tile_b * p.TB;
        const int n0 = n_tile * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait_relaxed(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb_local = smem_u32(&full_bar[s]);
          const uint32_t fb_leader = fb_local & kPeerBitMask;
          const int tap = kb / p.chunks_per_tap;
          if (rank == 0) mbar_expect_tx(fb_local, 2 * STAGE_BYTES);
          const int ch = kb - tap * p.chunks_per_tap;
          const bool src_b = ch >= p.chunks_a;
          const CUtensorMap* map = src_b ? &map_b : &map_a;
          const int c0 = (src_b ? (ch - p.chunks_a) : ch) * BK;
          const uint32_t a_dst = smem_base + s * STAGE_BYTES;
          if (p.mode == 1) {
            const int dy = tap >> 1, dx = tap & 1;
            const int C = src_b ? p.Cb : p.Ca;
            tma_load_5d_2sm(a_dst, map, fb_leader, dx * C + c0, w0, dy, h0, b0);
          } else {
            const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
            tma_load_5d_2sm(a_dst, map, fb_leader, c0, w0 + kx - pad, h0 + ky - pad, b0, 0);
          }
          tma_load_2d_2sm(a_dst + A_STAGE_BYTES, &map_w, fb_leader, kb * BK, n0);
          if (rank != 0) mbar_remote_arrive(fb_local, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA): warp-convergent loop, one
    // elected lane issues (descriptors stay in uniform registers; a lane-0-only loop cost a R2UR waterfall per MMA)
    if (rank == 0) {
      uint32_t it = 0, tile_iter = 0;
      for (int t = cluster_id; t < num_pair_tiles; t += num_clusters, ++tile_iter) {
        const uint32_t as = tile_iter % ACC, aph = (tile_iter / ACC) & 1;
        mbar_wait(smem_u32(&tmem_empty_bar[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * STAGE_BYTES;
          const uint64_t a_desc = make_sw128_desc(a_addr);
          const uint64_t b_desc = make_sw128_desc(a_addr + A_STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), IDESC, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(smem_u32(&empty_bar[s]));
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(smem_u32(&tmem_full_bar[as]));
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    pair_epilogue_role<BN, TADD>(p, &map_out, &map_add, add_bar, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, epi_aux, warp, lane, rank, cluster_id,
                           num_clusters, num_pair_tiles);
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its partner can still signal its barriers / read its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}


// ================================================================================================ conv v3 (3x3, halo reuse)
// Same CTA-pair / persistent / double-buffered-TMEM structure as conv_gemm_pair_kernel, but the A operand of all nine taps
// of a 64-channel chunk comes from ONE halo tile in shared memory: the CTA's M tile is 16 rows x 8 columns of pixels, the
// TMA box is the (16+2) x (8+2) pixel halo (180 rows of 128 B, out-of-bounds = zero padding), and tap (ky,kx) is the UMMA
// descriptor with start address + (ky*10 + kx) * 128 B and SBO = 10 * 128 B (next 8-pixel group = next image row).  This
// works because the 128-byte swizzle is a function of the absolute shared-memory address for both the TMA write and the
// tcgen05.mma read (verified on B200 by profiles/halo_probe.py: rel-L2 7e-7).  L2 -> SMEM traffic per k-block drops from
// 16 KB (A) + B to 2.5 KB + B, which is what bounded the tap-loop kernel (profiles/README.md).
//
// Two operand rings: A halo tiles (AS stages, one per 64-channel chunk) and weight tiles (BS stages, one per tap).
constexpr int HALO_TW = 8, HALO_TH = 16, HALO_W = HALO_TW + 2, HALO_H = HALO_TH + 2;
constexpr int HALO_BYTES = HALO_H * HALO_W * 128;                 // 23 040 B actually transferred per chunk
constexpr int HALO_STAGE_BYTES = ((HALO_BYTES + 1023) / 1024) * 1024;  // stage stride keeps 1024-byte alignment

// TPS = taps per B-ring stage: with BN = 128 one MMA is only 64 tensor-pipe cycles, and a barrier wait + commit per tap
// (4 MMAs) left the issuing warp, not the pipe, as the limit (ncu: pipe 52 % active, no barrier ever spun); three taps per
// stage amortise that over 12 MMAs.
// PRE: GroupNorm + scale/shift + SiLU of the consuming Block applied to the halo tile in shared memory (4 extra warps
// 12..15 per CTA: a_full -> transform in place -> a_ready -> MMA), so the activated tensor never exists in HBM.
constexpr int NUM_THREADS3 = 512;
template <int BN, int AS, int BS, bool TADD, int TPS, bool PRE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PRE ? NUM_THREADS3 : NUM_THREADS2, 1)
conv_gemm_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
                      const __grid_constant__ CUtensorMap map_add, const ConvParams p, const int num_pair_tiles) {
  constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  constexpr uint32_t ACC = AccStages<BN>::N;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + AS * HALO_STAGE_BYTES;
  constexpr int EPI_BUFS = TADD ? 4 : 2;
  constexpr int B_STAGE = TPS * B_HALF_BYTES;
  const uint32_t epi_smem = b_base + BS * B_STAGE;
  uint8_t* ctrl = smem_gen + AS * HALO_STAGE_BYTES + BS * B_STAGE + EPI_BUFS * EPI_STAGE_BYTES;
  uint64_t* add_bar = reinterpret_cast<uint64_t*>(ctrl + 320);
  uint64_t* a_ready = reinterpret_cast<uint64_t*>(ctrl + 352);  // PRE: leader only, 4 transform warps x 2 CTAs
  float* epi_aux = reinterpret_cast<float*>(ctrl + 384);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* a_empty = a_full + AS;
  uint64_t* b_full = a_empty + AS;
  uint64_t* b_empty = b_full + BS;
  uint64_t* tmem_full_bar = b_empty + BS;
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC);
  static_assert((2 * AS + 2 * BS + 2 * ACC) * 8 + 4 <= 320, "barrier block overflows its 320 bytes");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int Ctot = p.Ca + p.Cb;

  cluster_sync_all();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_out);
    if (TADD) tma_prefetch_desc(&map_add);
    if (p.Cb > 0) tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
#pragma unroll
    for (int s = 0; s < AS; ++s) {
      mbar_init(smem_u32(&a_full[s]), PRE ? 1 : 2);  // PRE: each CTA tracks its own halo tile (its transform warps wait on it)
      mbar_init(smem_u32(&a_empty[s]), 1);
      mbar_init(smem_u32(&a_ready[s]), 8);
    }
#pragma unroll
    for (int s = 0; s < BS; ++s) {
      mbar_init(smem_u32(&b_full[s]), 2);
      mbar_init(smem_u32(&b_empty[s]), 1);
    }
#pragma unroll
    for (int a = 0; a < (int)ACC; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), 16);
    }
    for (int a = 0; a < 4; ++a) mbar_init(smem_u32(&add_bar[a]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc_2sm(smem_u32(tmem_ptr_smem), TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  kd_pdl_wait();     // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  kd_pdl_trigger();

  if (warp == 0) {
    // ================================================================ weight (B) TMA producer (both CTAs, one lane each)
    if (lane == 0) {
      uint32_t itb = 0;
      for (int t = cluster_id; t < num_pair_tiles; t += num_clusters) {
        const int n0 = (t % p.n_tiles) * BN + (int)rank * (BN / 2);
        for (int ch = 0; ch < p.chunks_per_tap; ++ch) {
#pragma unroll 1
          for (int tg = 0; tg < 9 / TPS; ++tg, ++itb) {
            const uint32_t sb = itb % BS;
            const uint32_t phb = (itb / BS) & 1;
            mbar_wait_relaxed(smem_u32(&b_empty[sb]), phb ^ 1u);
            const uint32_t fb_local = smem_u32(&b_full[sb]);
            if (rank == 0) mbar_expect_tx(fb_local, 2 * B_STAGE);
#pragma unroll
            for (int j = 0; j < TPS; ++j)
              tma_load_2d_2sm(b_base + sb * B_STAGE + j * B_HALF_BYTES, &map_w, fb_local & kPeerBitMask,
                              (tg * TPS + j) * Ctot + ch * BK, n0);
            if (rank != 0) mbar_remote_arrive(fb_local, 0);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ================================================================ halo (A) TMA producer: its own warp, so the A ring runs
    // AS chunks ahead of the MMA independently of the weight ring (the fused pre-activation adds a transform stage to the
    // A pipeline, which a load issued only after the previous chunk's nine weight loads could not hide)
    if (lane == 0) {
      uint32_t ita = 0;
      for (int t = cluster_id; t < num_pair_tiles; t += num_clusters) {
        int m_tile = (t / p.n_tiles) * 2 + (int)rank;
        const int tile_w = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int tile_h = m_tile % p.tiles_h;
        const int tile_b = m_tile / p.tiles_h;
        const int w0 = tile_w * HALO_TW, h0 = tile_h * HALO_TH;
        for (int ch = 0; ch < p.chunks_per_tap; ++ch, ++ita) {
          const uint32_t sa = ita % AS;
          const uint32_t pha = (ita / AS) & 1;
          mbar_wait_relaxed(smem_u32(&a_empty[sa]), pha ^ 1u);
          const uint32_t fa_local = smem_u32(&a_full[sa]);
          const bool src_b = ch >= p.chunks_a;
          const CUtensorMap* map = src_b ? &map_b : &map_a;
          const int c0 = (src_b ? (ch - p.chunks_a) : ch) * BK;
          if (PRE) {
            mbar_expect_tx(fa_local, HALO_BYTES);
            tma_load_5d(a_base + sa * HALO_STAGE_BYTES, map, fa_local, c0, w0 - 1, h0 - 1, tile_b, 0);
          } else {
            if (rank == 0) mbar_expect_tx(fa_local, 2 * HALO_BYTES);
            tma_load_5d_2sm(a_base + sa * HALO_STAGE_BYTES, map, fa_local & kPeerBitMask, c0, w0 - 1, h0 - 1, tile_b, 0);
            if (rank != 0) mbar_remote_arrive(fa_local, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA): the whole warp runs the loop
    // convergently (all lanes poll the barriers and compute the same, uniform, descriptors), one elected lane issues
    if (rank == 0) {
      uint32_t ita = 0, itb = 0, tile_iter = 0;
      for (int t = cluster_id; t < num_pair_tiles; t += num_clusters, ++tile_iter) {
        const uint32_t as = tile_iter % ACC, aph = (tile_iter / ACC) & 1;
        mbar_wait(smem_u32(&tmem_empty_bar[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int ch = 0; ch < p.chunks_per_tap; ++ch, ++ita) {
          const uint32_t sa = ita % AS;
          const uint32_t pha = (ita / AS) & 1;
          mbar_wait(smem_u32(PRE ? &a_ready[sa] : &a_full[sa]), pha);
          tc_fence_after();
          const uint32_t a_stage = a_base + sa * HALO_STAGE_BYTES;
#pragma unroll
          for (int tg = 0; tg < 9 / TPS; ++tg) {
            const uint32_t sb = itb % BS;
            const uint32_t phb = (itb / BS) & 1;
            ++itb;
            mbar_wait(smem_u32(&b_full[sb]), phb);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < TPS; ++j) {
                const int tap = tg * TPS + j;
                const int ky = tap / 3, kx = tap - ky * 3;
                const uint64_t a_desc = make_sw128_desc(a_stage + (uint32_t)(ky * HALO_W + kx) * 128u, HALO_W * 128);
                const uint64_t b_desc = make_sw128_desc(b_base + sb * B_STAGE + j * B_HALF_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  umma_f16_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), IDESC, (ch | tap | k) != 0 ? 1u : 0u);
              }
              umma_commit_2sm(smem_u32(&b_empty[sb]));
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit_2sm(smem_u32(&a_empty[sa]));
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(smem_u32(&tmem_full_bar[as]));
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 12) {
    pair_epilogue_role<BN, TADD>(p, &map_out, &map_add, add_bar, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, epi_aux, warp, lane, rank, cluster_id,
                           num_clusters, num_pair_tiles);
  } else if (PRE && warp >= 12) {
    // ================================================================ pre-activation transform (both CTAs, 128 threads)
    // thread -> one physical 16-byte unit column (8 channels) and rows rbase + 16 i of the 180-row halo tile.  The 128B
    // swizzle XORs the unit index with (row & 7), and (rbase + 16 i) & 7 is constant, so every thread's 8 channels -- and its
    // 8 {A, B} coefficient pairs -- are fixed for the whole chunk.
    const int tt = threadIdx.x - 384;
    const int pu = tt & 7, rbase = tt >> 3;
    const int lu = pu ^ (rbase & 7);  // logical channel octet of this thread
    uint32_t ita = 0;
    for (int t = cluster_id; t < num_pair_tiles; t += num_clusters) {
      int m_tile = (t / p.n_tiles) * 2 + (int)rank;
      const int tile_w = m_tile % p.tiles_w;
      m_tile /= p.tiles_w;
      const int tile_h = m_tile % p.tiles_h;
      const int tile_b = m_tile / p.tiles_h;
      const int x0 = tile_w * HALO_TW - 1, y0 = tile_h * HALO_TH - 1;
      const bool live = tile_b < p.B;
      for (int ch = 0; ch < p.chunks_per_tap; ++ch, ++ita) {
        const uint32_t sa = ita % AS;
        const uint32_t pha = (ita / AS) & 1;
        float4 cf[4];
        if (live) {
          const float4* cp = reinterpret_cast<const float4*>(p.pre_coef + (long long)tile_b * Ctot + ch * BK + lu * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            cf[j] = __ldg(cp + j);  // {A0, B0, A1, B1} ...; halved: the SiLU below takes (A x + B) / 2
            cf[j].x *= 0.5f; cf[j].y *= 0.5f; cf[j].z *= 0.5f; cf[j].w *= 0.5f;
          }
        }
        mbar_wait_relaxed(smem_u32(&a_full[sa]), pha);
        if (live) {
          const uint32_t col = a_base + sa * HALO_STAGE_BYTES + (uint32_t)pu * 16u;
          // three units per step, all loads first: 24 independent FMA -> SiLU chains keep the MUFU pipe busy (one unit at a
          // time left the warp latency-bound at ~550 cycles per unit instead of the 128-cycle MUFU issue cost)
          constexpr int UNR = 3, ROWS = HALO_H * HALO_W;
          static_assert(((ROWS + 15) / 16) % UNR == 0, "unit loop must split evenly");
#pragma unroll 1
          for (int i0 = 0; i0 < (ROWS + 15) / 16; i0 += UNR) {
            int4 raw[UNR];
            bool ok[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
              const int r = rbase + 16 * (i0 + u);
              const int hy = r / HALO_W, hx = r - hy * HALO_W;
              const int y = y0 + hy, x = x0 + hx;
              // rows past the tile and zero padding of the ACTIVATED tensor (pixels outside the image) stay untouched
              ok[u] = (r < ROWS) && y >= 0 && y < p.H && x >= 0 && x < p.W;
              raw[u] = make_int4(0, 0, 0, 0);
              if (ok[u]) {
                const uint32_t addr = col + (uint32_t)r * 128u;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(raw[u].x), "=r"(raw[u].y), "=r"(raw[u].z), "=r"(raw[u].w)
                             : "r"(addr));
              }
            }
            float v[UNR][8];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
              h16x8_to_float(*reinterpret_cast<const h16x8*>(&raw[u]), v[u]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[u][2 * j] = fmaf(cf[j].x, v[u][2 * j], cf[j].y);
                v[u][2 * j + 1] = fmaf(cf[j].z, v[u][2 * j + 1], cf[j].w);
              }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[u][j] = silu_from_half_arg(v[u][j]);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
              if (ok[u]) {
                const int r = rbase + 16 * (i0 + u);
                const uint32_t addr = col + (uint32_t)r * 128u;
                const h16x8 o8 = float_to_h16x8(v[u]);
                const int4 ov = *reinterpret_cast<const int4*>(&o8);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(ov.x), "r"(ov.y), "r"(ov.z), "r"(ov.w)
                             : "memory");
              }
            }
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_remote_arrive(smem_u32(&a_ready[sa]), 0);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) KD_FAIL(KD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdims[5], gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) KD_FAIL(KD_ERR_BAD_ARG, "TMA box dim %d = %u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] % 16 != 0) KD_FAIL(KD_ERR_BAD_ARG, "TMA stride %d = %llu not a multiple of 16 bytes", i, (unsigned long long)gstr[i]);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) KD_FAIL(KD_ERR_BAD_ARG, "TMA base pointer not 16-byte aligned");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) KD_FAIL(KD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return KD_OK;
}

int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// activation tensor map for one source with C channels
int make_act_map(CUtensorMap* m, const KdConvDesc* d, const void* x, int C, int TW, int TH, int TB, bool halo = false) {
  if (halo) {  // one box = the (TH+2) x (TW+2) pixel halo of a 16 x 8 tile
    const uint64_t dims[5] = {(uint64_t)C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B, 1ull};
    const uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)d->W * C * 2, (uint64_t)d->H * d->W * C * 2,
                             (uint64_t)d->B * d->H * d->W * C * 2};
    const uint32_t box[5] = {(uint32_t)BK, (uint32_t)HALO_W, (uint32_t)HALO_H, 1u, 1u};
    return encode_map(m, x, 5, dims, str, box);
  }
  if (d->mode == 1) {
    const uint64_t Hi = 2ull * d->H, Wi = 2ull * d->W;
    const uint64_t dims[5] = {2ull * C, (uint64_t)d->W, 2ull, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t str[4] = {2ull * C * 2, Wi * C * 2, 2 * Wi * C * 2, Hi * Wi * C * 2};
    const uint32_t box[5] = {(uint32_t)BK, (uint32_t)TW, 1u, (uint32_t)TH, (uint32_t)TB};
    return encode_map(m, x, 5, dims, str, box);
  }
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B, 1ull};
  const uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)d->W * C * 2, (uint64_t)d->H * d->W * C * 2,
                           (uint64_t)d->B * d->H * d->W * C * 2};
  const uint32_t box[5] = {(uint32_t)BK, (uint32_t)TW, (uint32_t)TH, (uint32_t)TB, 1u};
  return encode_map(m, x, 5, dims, str, box);
}

template <int BN, int STAGES, bool SPLIT = false>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mw, const ConvParams& p, long long grid,
           cudaStream_t stream, int splits = 1) {
  constexpr int SMEM = STAGES * (A_STAGE_BYTES + BN * BK * 2) + 1024 /*align slack*/ + 128 /*barriers*/ + BN * 4 /*bias*/;
  static bool configured = false;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!configured) {
      KD_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      configured = true;
    }
  }
  KD_CUDA(kd_launch(conv_gemm_kernel<BN, STAGES, SPLIT>, dim3((unsigned)grid, (unsigned)splits), dim3(NUM_THREADS), SMEM, stream, ma, mb, mw, p));
  return KD_OK;
}

// ---- split-K second pass: out = epilogue(sum over splits, in split order) for one (m-tile, 64-column group) per CTA.
// warp = 32-row quarter of the tile; lane -> (channel octet o = lane & 7, 8-row slice sl = lane >> 3), the mapping of the fused
// statistics in pair_epilogue_role: the statistics rows written here have the same geometry as the conv epilogue's.
__global__ void __launch_bounds__(128) splitk_finish_kernel(const ConvParams p, const int S, const int BN, const int tiles_total) {
  const int n_g64 = (p.Cout + 63) >> 6;
  const int g64 = blockIdx.x % n_g64;
  const int m_lin = blockIdx.x / n_g64;
  const int n_tile = (g64 * 64) / BN, col0 = (g64 * 64) % BN;
  const int tile = m_lin * p.n_tiles + n_tile;
  int m_tile = m_lin;
  const int tile_w = m_tile % p.tiles_w;
  m_tile /= p.tiles_w;
  const int tile_h = m_tile % p.tiles_h;
  const int tile_b = m_tile / p.tiles_h;
  const int quarter = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = lane & 7, sl = lane >> 3;
  const int nc = g64 * 64 + o * 8;
  const bool col_ok = nc + 8 <= p.Cout;
  const int Cq = p.Cout >> 2;
  kd_pdl_wait();
  kd_pdl_trigger();
  float bias8[8], lw8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bias8[j] = (p.bias != nullptr && col_ok) ? __ldg(p.bias + nc + j) : 0.f;
    lw8[j] = (p.logit_w != nullptr && col_ok) ? __ldg(p.logit_w + nc + j) : 0.f;
  }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 2
  for (int i = 0; i < 8; ++i) {
    const int r = quarter * 32 + sl * 8 + i;
    const int tw = r % p.TW, th = (r / p.TW) % p.TH, tb = r / (p.TW * p.TH);
    const int b = tile_b * p.TB + tb, h = tile_h * p.TH + th, w = tile_w * p.TW + tw;
    const bool row_ok = (b < p.B) && (h < p.H) && (w < p.W);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    const float* src = p.splitk_ws + ((size_t)tile * BM + r) * BN + col0 + o * 8;
    const size_t split_stride = (size_t)tiles_total * BM * BN;
    // fixed order: the result does not depend on how many CTAs ran concurrently.  Eight splits' loads are issued before the
    // first add: with one dependent load per add the kernel sat at S x 8 rows L2 round trips (22 us for a 64-pixel image).
    for (int sp0 = 0; sp0 < S; sp0 += 8) {
      float4 a0[8], a1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (sp0 + u < S) {
          a0[u] = __ldg(reinterpret_cast<const float4*>(src + (sp0 + u) * split_stride));
          a1[u] = __ldg(reinterpret_cast<const float4*>(src + (sp0 + u) * split_stride) + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (sp0 + u < S) {
          v[0] += a0[u].x; v[1] += a0[u].y; v[2] += a0[u].z; v[3] += a0[u].w;
          v[4] += a1[u].x; v[5] += a1[u].y; v[6] += a1[u].z; v[7] += a1[u].w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += bias8[j];
    apply_act8(v, p.act);
    long long out_off = 0;
    if (row_ok && col_ok) {
      if (p.out_mode == 1) {
        const int q4 = nc / Cq, c = nc - q4 * Cq;
        out_off = (((long long)b * (2 * p.H) + (2 * h + (q4 >> 1))) * (2 * p.W) + (2 * w + (q4 & 1))) * Cq + c;
      } else {
        out_off = (((long long)b * p.H + h) * p.W + w) * p.Cout + nc;
      }
      if (p.addend != nullptr) {
        float a[8];
        if (p.addend_f32) {
          const float4* ap = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.addend) + out_off);
          const float4 a0 = ap[0], a1 = ap[1];
          a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        } else {
          const h16x8 raw = *reinterpret_cast<const h16x8*>(reinterpret_cast<const h16*>(p.addend) + out_off);
          h16x8_to_float(raw, a);
        }
        if (p.addend_scale != nullptr) {
          const float* gate = p.addend_scale + (long long)b * p.Cout + nc;
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] *= __ldg(gate + j);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += a[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    float lacc = 0.f;
    if (p.out_f32) {
      if (row_ok && col_ok) {
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off);
        op[0] = make_float4(v[0], v[1], v[2], v[3]);
        op[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    } else {
      const h16x8 o8 = float_to_h16x8(v);
      if (row_ok && col_ok) *reinterpret_cast<h16x8*>(reinterpret_cast<h16*>(p.out) + out_off) = o8;
      float f[8];
      h16x8_to_float(o8, f);  // statistics / logits of the ROUNDED values, as a pass over the stored tensor would see them
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1 += f[j];
        s2 = fmaf(f[j], f[j], s2);
        lacc = fmaf(f[j], lw8[j], lacc);
      }
    }
    if (p.logit_w != nullptr) {  // one partial per (row, 64-column group): merge the 8 octet lanes of the row
      lacc += __shfl_xor_sync(0xffffffffu, lacc, 1);
      lacc += __shfl_xor_sync(0xffffffffu, lacc, 2);
      lacc += __shfl_xor_sync(0xffffffffu, lacc, 4);
      if (o == 0 && row_ok) p.logit_parts[(long long)g64 * ((long long)p.B * p.H * p.W) + ((long long)b * p.H + h) * p.W + w] = lacc;
    }
  }
  if (p.stats != nullptr) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
    if (sl == 0 && col_ok && tile_b < p.tiles_b)
      reinterpret_cast<float2*>(p.stats)[((long long)m_lin * 4 + quarter) * (p.Cout >> 3) + (nc >> 3)] = make_float2(s1, s2);
  }
}

// output tensor map for the coalesced epilogue (h16 NHWC, or its pixel-shuffle view [B, H, 2(dy), W, 2(dx)*Cq])
int make_out_map(CUtensorMap* m, const ConvParams& p, void* out) {
  if (p.out_mode == 1) {
    const uint64_t Cq = (uint64_t)p.Cout / 4, Ho = 2ull * p.H, Wo = 2ull * p.W;
    const uint64_t dims[5] = {2 * Cq, (uint64_t)p.W, 2ull, (uint64_t)p.H, (uint64_t)p.B};
    const uint64_t str[4] = {2 * Cq * 2, Wo * Cq * 2, 2 * Wo * Cq * 2, Ho * Wo * Cq * 2};
    const uint32_t box[5] = {(uint32_t)EPI_COLS, (uint32_t)p.TW, 1u, (uint32_t)p.TH, (uint32_t)p.TB};
    return encode_map(m, out, 5, dims, str, box);
  }
  const uint64_t C = (uint64_t)p.Cout;
  const uint64_t dims[5] = {C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B, 1ull};
  const uint64_t str[4] = {C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2, (uint64_t)p.B * p.H * p.W * C * 2};
  const uint32_t box[5] = {(uint32_t)EPI_COLS, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TB, 1u};
  return encode_map(m, out, 5, dims, str, box);
}

// TADD (h16 addend, plain NHWC output): the addend tile is TMA-fed into double-buffered staging, paid for with ring stages
template <int BN, int STAGES, bool TADD>
int launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mw, const ConvParams& p, cudaStream_t stream) {
  constexpr int SMEM = STAGES * (A_STAGE_BYTES + (BN / 2) * BK * 2) + (TADD ? 4 : 2) * EPI_STAGE_BYTES + 1024 /*align slack*/ +
                       384 /*barriers*/ + 3 * BN * 4 /*bias + gate / to_k staging*/;
  static_assert(SMEM <= 232448, "pair kernel exceeds the 227 KB shared-memory limit");
  CUtensorMap mo = ma, madd = ma;
  if (!p.out_f32) {
    int rc = make_out_map(&mo, p, p.out);
    if (rc) return rc;
    madd = mo;
    if (TADD) {
      rc = make_out_map(&madd, p, const_cast<void*>(p.addend));
      if (rc) return rc;
    }
  }
  static bool configured = false;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!configured) {
      KD_CUDA(cudaFuncSetAttribute(conv_gemm_pair_kernel<BN, STAGES, TADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      configured = true;
    }
  }
  const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_b;
  const long long pair_tiles = ((m_tiles + 1) / 2) * p.n_tiles;
  KD_REQUIRE(pair_tiles < 2147483647LL, "kd_conv_gemm: too many tiles");
  int clusters = kd_num_sms() / 2;
  if (pair_tiles < clusters) clusters = (int)pair_tiles;
  KD_CUDA(kd_launch(conv_gemm_pair_kernel<BN, STAGES, TADD>, dim3(2 * clusters), dim3(NUM_THREADS2), SMEM, stream, ma, mb, mw, mo, madd, p,
                    (int)pair_tiles));
  return KD_OK;
}

template <int BN, int AS, int BS, bool TADD, int TPS, bool PRE>
int launch_halo(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mw, const ConvParams& p, cudaStream_t stream) {
  constexpr int SMEM = AS * HALO_STAGE_BYTES + BS * TPS * (BN / 2) * BK * 2 + (TADD ? 4 : 2) * EPI_STAGE_BYTES + 1024 + 384 +
                       3 * BN * 4;
  static_assert(SMEM <= 232448, "halo kernel exceeds the 227 KB shared-memory limit");
  CUtensorMap mo, madd;
  int rc = make_out_map(&mo, p, p.out);
  if (rc) return rc;
  madd = mo;
  if (TADD) {
    rc = make_out_map(&madd, p, const_cast<void*>(p.addend));
    if (rc) return rc;
  }
  static bool configured = false;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!configured) {
      KD_CUDA(cudaFuncSetAttribute(conv_gemm_halo_kernel<BN, AS, BS, TADD, TPS, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      configured = true;
    }
  }
  const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_b;
  const long long pair_tiles = ((m_tiles + 1) / 2) * p.n_tiles;
  KD_REQUIRE(pair_tiles < 2147483647LL, "kd_conv_gemm: too many tiles");
  int clusters = kd_num_sms() / 2;
  if (pair_tiles < clusters) clusters = (int)pair_tiles;
  KD_CUDA(kd_launch(conv_gemm_halo_kernel<BN, AS, BS, TADD, TPS, PRE>, dim3(2 * clusters), dim3(PRE ? NUM_THREADS3 : NUM_THREADS2), SMEM, stream,
                    ma, mb, mw, mo, madd, p, (int)pair_tiles));
  return KD_OK;
}

int g_conv_impl = 0;  // 0 = auto, 1 = single-CTA kernel, 2 = pair (halo where possible), 4 = tap-loop pair kernel only (no halo)

}  // namespace

int kd_encode_tiled_h16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box) {
  return encode_map(m, base, rank, dims, strides_bytes, box);
}

extern "C" int kd_set_conv_impl(int impl) {
  if (!(impl == 0 || impl == 1 || impl == 2 || impl == 4 || impl == 8)) KD_FAIL(KD_ERR_BAD_ARG, "kd_set_conv_impl: impl must be 0, 1, 2, 4 or 8");
  g_conv_impl = impl;
  return KD_OK;
}

namespace {
// tiling the dispatcher will use for `d` (shared by kd_conv_gemm_stats and kd_conv_stats_layout)
struct Tiling {
  bool use_pair, use_halo;
  int TW, TH, TB, tiles_w, tiles_h, tiles_b;
  int splits, kb_per_split, BN;  // split-K (splits > 1): single-CTA 128 x BN tiles, K range divided into `splits` CTAs
};
// Split-K applies to convolutions on tiny images (<= 8 x 8 output pixels per sample by default: the 8^2 level of the base UNet),
// where M = B * H * W gives a handful of tiles while K = 9 * Cin is 7 000 - 18 000: without it one or two CTA pairs walk the
// whole K range (65 us for 1 GFLOP).  The split count is a function of the PER-SAMPLE shape only -- never of the batch size --
// so that a sample's accumulation order, hence its bits, does not depend on which other patches share its batch.
void choose_split(const KdConvDesc* d, Tiling* t) {
  t->splits = 1;
  t->kb_per_split = 0;
  if (g_conv_impl != 0 || d->mode == 2) return;
  const long hw = (long)d->H * d->W;
  static int tgt_small = -1, tgt_mid = -1;  // k-blocks per split for <= 8x8 / <= 16x16 images (KD_SPLITK_TARGETS="a,b": tuning hook; 0 = off)
  if (tgt_small < 0) {
    // measured on the 64^2 / 256^2 stage UNets (profiles/README.md round 2, item 4): 18 k-blocks per split for <= 16x16 images is the
    // fastest at B = 1-4 (the wavefront's chain-bound batches: 4.97 -> 4.68 ms, 4.11 -> 3.69 ms) and costs 10 % at B = 16
    int a = 18, b = 18;
    if (const char* e = getenv("KD_SPLITK_TARGETS")) sscanf(e, "%d,%d", &a, &b);
    tgt_small = a;
    tgt_mid = b;
  }
  const int target = hw <= 64 ? tgt_small : (hw <= 256 ? tgt_mid : 0);
  if (target == 0) return;
  const int taps = (d->mode == 1) ? 4 : d->ksize * d->ksize;
  const int num_kb = taps * ((d->Ca + d->Cb) / BK);
  int S = num_kb / target;
  if (S < 2) return;
  if (S > 32) S = 32;
  t->kb_per_split = (num_kb + S - 1) / S;
  t->splits = (num_kb + t->kb_per_split - 1) / t->kb_per_split;
}
Tiling choose_tiling(const KdConvDesc* d) {
  Tiling t;
  choose_split(d, &t);
  t.use_pair = t.splits == 1 && ((g_conv_impl == 2) || (g_conv_impl == 4) || ((g_conv_impl == 0 || g_conv_impl == 8) && d->Cout >= 128));
  t.use_halo = t.use_pair && g_conv_impl != 4 && d->mode == 0 && d->ksize == 3 && d->H >= HALO_TH &&
               d->W >= HALO_TW && !d->out_f32 && d->out_mode == 0;
  if (t.use_halo) {
    t.TW = HALO_TW; t.TH = HALO_TH; t.TB = 1;
  } else {
    t.TH = pow2_ceil(d->H) < 8 ? pow2_ceil(d->H) : 8;
    const int wmax = BM / t.TH;
    t.TW = pow2_ceil(d->W) < wmax ? pow2_ceil(d->W) : wmax;
    t.TB = BM / (t.TH * t.TW);
  }
  t.tiles_w = kd_ceil_div(d->W, t.TW);
  t.tiles_h = kd_ceil_div(d->H, t.TH);
  t.tiles_b = kd_ceil_div(d->B, t.TB);
  t.BN = t.use_pair ? (d->Cout >= 256 ? 256 : 128) : (d->Cout >= 128 ? 128 : 64);
  return t;
}
}  // namespace

// layout[0] = total rows of the statistics buffer ([rows][Cout/8][2] fp32), layout[1] = m-tiles per batch group,
// layout[2] = batch images per tile (TB); 0 rows = this shape does not produce fused statistics
extern "C" int kd_conv_stats_layout(const KdConvDesc* d, int* layout) {
  KD_REQUIRE(d && layout, "kd_conv_stats_layout: null argument");
  const Tiling t = choose_tiling(d);
  layout[0] = layout[1] = layout[2] = 0;
  layout[3] = t.use_halo ? 1 : 0;
  if (!(t.use_pair || t.splits > 1) || d->out_f32 || d->out_mode != 0 || t.TB > 2 || d->Cout % 8 != 0) return KD_OK;
  layout[0] = t.tiles_w * t.tiles_h * t.tiles_b * 4;
  layout[1] = t.tiles_w * t.tiles_h;
  layout[2] = t.TB;
  return KD_OK;
}

extern "C" int kd_conv_gemm(const KdConvDesc* d, const void* xa, const void* xb, const void* w, const float* bias,
                            const void* addend, const float* addend_scale, void* out, kd_stream_t stream_) {
  return kd_conv_gemm_fused(d, xa, xb, w, bias, addend, addend_scale, out, nullptr, stream_);
}

extern "C" size_t kd_conv_splitk_workspace_bytes(const KdConvDesc* d) {
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Cout <= 0) return 0;
  const Tiling t = choose_tiling(d);
  if (t.splits <= 1) return 0;
  return (size_t)t.splits * t.tiles_w * t.tiles_h * t.tiles_b * kd_ceil_div(d->Cout, t.BN) * BM * t.BN * sizeof(float);
}

extern "C" int kd_conv_gemm_fused(const KdConvDesc* d, const void* xa, const void* xb, const void* w, const float* bias,
                                  const void* addend, const float* addend_scale, void* out, const KdConvFusion* fusion,
                                  kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  float* stats = fusion ? fusion->stats : nullptr;
  const float* logit_w = fusion ? fusion->logit_w : nullptr;
  float* logit_parts = fusion ? fusion->logit_parts : nullptr;
  const float* pre_coef = fusion ? fusion->pre_coef : nullptr;
  KD_REQUIRE(d && xa && w && out, "kd_conv_gemm: null argument");
  KD_REQUIRE(d->mode >= 0 && d->mode <= 2, "kd_conv_gemm: bad mode %d", d->mode);
  KD_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, "kd_conv_gemm: bad output shape %dx%dx%d", d->B, d->H, d->W);
  KD_REQUIRE(d->Ca > 0 && d->Ca % BK == 0 && d->Cb >= 0 && d->Cb % BK == 0,
             "kd_conv_gemm: channel counts must be multiples of %d (Ca=%d Cb=%d)", BK, d->Ca, d->Cb);
  KD_REQUIRE(d->Cb == 0 || xb != nullptr, "kd_conv_gemm: Cb > 0 but xb is null");
  KD_REQUIRE(d->Cout > 0 && d->Cout % 8 == 0, "kd_conv_gemm: Cout=%d must be a multiple of 8", d->Cout);
  KD_REQUIRE(d->mode != 0 || d->ksize == 1 || d->ksize == 3, "kd_conv_gemm: ksize must be 1 or 3");
  KD_REQUIRE(d->out_mode == 0 || (d->out_mode == 1 && d->Cout % 128 == 0),
             "kd_conv_gemm: pixel-shuffle output needs Cout %% 128 == 0 (got %d)", d->Cout);
  KD_REQUIRE(d->mode != 2 || (d->H == 1 && d->B == 1 && d->Cb == 0), "kd_conv_gemm: mode 2 expects B=H=1, single source");

  const int taps = (d->mode == 1) ? 4 : (d->mode == 0 ? d->ksize * d->ksize : 1);
  ConvParams p;
  p.mode = d->mode == 2 ? 0 : d->mode;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Ca = d->Ca; p.Cb = d->Cb; p.Cout = d->Cout;
  p.ksize = (d->mode == 0) ? d->ksize : 1;
  p.act = d->act; p.out_mode = d->out_mode; p.out_f32 = d->out_f32; p.addend_f32 = d->addend_f32;
  // kernel choice (one place: choose_tiling): CTA-pair tiles (256 x 256 / 256 x 128) whenever the layer is wide enough, else
  // the single-CTA kernel; 3x3 convolutions on images of at least 16 x 8 pixels use the halo-reuse variant
  const Tiling tl = choose_tiling(d);
  const bool use_pair = tl.use_pair, use_halo = tl.use_halo;
  KD_REQUIRE(!((g_conv_impl == 2 || g_conv_impl == 4) && d->Cout < 128), "kd_conv_gemm: the CTA-pair kernel needs Cout >= 128");
  p.TW = tl.TW; p.TH = tl.TH; p.TB = tl.TB;
  p.tiles_w = tl.tiles_w; p.tiles_h = tl.tiles_h; p.tiles_b = tl.tiles_b;
  p.chunks_a = d->Ca / BK;
  p.chunks_per_tap = (d->Ca + d->Cb) / BK;
  p.num_kb = taps * p.chunks_per_tap;
  p.bias = bias; p.addend = addend; p.addend_scale = addend_scale; p.out = out; p.stats = stats; p.logit_w = logit_w; p.logit_parts = logit_parts;
  p.pre_coef = reinterpret_cast<const float2*>(pre_coef);

  const int BN = tl.BN;
  p.n_tiles = kd_ceil_div(d->Cout, BN);
  p.kb_per_split = tl.splits > 1 ? tl.kb_per_split : 0;
  p.splitk_ws = nullptr;
  if (tl.splits > 1) {
    const size_t need = kd_conv_splitk_workspace_bytes(d);
    KD_REQUIRE(fusion && fusion->splitk_ws && fusion->splitk_ws_bytes >= need,
               "kd_conv_gemm: this shape runs split-K and needs a %zu-byte workspace in KdConvFusion (kd_conv_splitk_workspace_bytes)", need);
    p.splitk_ws = reinterpret_cast<float*>(fusion->splitk_ws);
  }
  const long long grid = (long long)p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles;
  KD_REQUIRE(grid > 0 && grid < 2147483647LL, "kd_conv_gemm: grid too large");

  KdConvDesc dd = *d;
  if (d->mode == 2) dd.mode = 0;
  CUtensorMap ma, mb, mw;
  int rc = make_act_map(&ma, &dd, xa, d->Ca, p.TW, p.TH, p.TB, use_halo);
  if (rc) return rc;
  if (d->Cb > 0) {
    rc = make_act_map(&mb, &dd, xb, d->Cb, p.TW, p.TH, p.TB, use_halo);
    if (rc) return rc;
  } else {
    mb = ma;
  }
  {
    const uint64_t Ktot = (uint64_t)taps * (d->Ca + d->Cb);
    const uint64_t dims[2] = {Ktot, (uint64_t)d->Cout};
    const uint64_t str[1] = {Ktot * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)(use_pair ? BN / 2 : BN)};  // pair kernel: each CTA loads half the tile rows
    rc = encode_map(&mw, w, 2, dims, str, box);
    if (rc) return rc;
  }
  KD_REQUIRE(pre_coef == nullptr || use_halo, "kd_conv_gemm_fused: pre_coef needs the 3x3 halo kernel (kd_conv_stats_layout[3])");
  if (stats != nullptr || logit_w != nullptr) {
    int lay[4];
    kd_conv_stats_layout(d, lay);
    KD_REQUIRE(lay[0] > 0, "kd_conv_gemm_fused: this shape / kernel does not produce fused statistics (see kd_conv_stats_layout)");
    KD_REQUIRE(logit_w == nullptr || (logit_parts != nullptr && addend_scale == nullptr && d->Cout % 64 == 0),
               "kd_conv_gemm_fused: fused GlobalContext logits need Cout %% 64 == 0, an output buffer and no gate");
  }
  if (tl.splits > 1) {
    rc = (BN == 128) ? launch<128, 3, true>(ma, mb, mw, p, grid, stream, tl.splits) : launch<64, 4, true>(ma, mb, mw, p, grid, stream, tl.splits);
    if (rc) return rc;
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    KD_CUDA(kd_launch(splitk_finish_kernel, dim3((unsigned)(m_tiles * ((d->Cout + 63) / 64))), dim3(128), 0, stream, p, tl.splits, BN, (int)grid));
    return KD_OK;
  }
  const bool tadd = use_pair && addend != nullptr && !d->addend_f32 && !d->out_f32 && d->out_mode == 0;
  if (use_halo) {
    // <BN, A stages, B stages, TMA addend, taps per B stage, fused pre-activation>; smem: A 23 KB / stage, B 8 (16) KB / tap
    if (p.pre_coef != nullptr) {
      if (tadd) {
        if (BN == 256) return launch_halo<256, 3, 5, true, 1, true>(ma, mb, mw, p, stream);
        return launch_halo<128, 4, 2, true, 3, true>(ma, mb, mw, p, stream);
      }
      if (BN == 256) return launch_halo<256, 3, 7, false, 1, true>(ma, mb, mw, p, stream);
      return launch_halo<128, 4, 4, false, 3, true>(ma, mb, mw, p, stream);
    }
    if (tadd) {
      if (BN == 256) return launch_halo<256, 3, 5, true, 1, false>(ma, mb, mw, p, stream);
      return launch_halo<128, 3, 3, true, 3, false>(ma, mb, mw, p, stream);
    }
    if (BN == 256) return launch_halo<256, 3, 7, false, 1, false>(ma, mb, mw, p, stream);
    return launch_halo<128, 3, 4, false, 3, false>(ma, mb, mw, p, stream);
  }
  if (use_pair) {
    if (tadd) {
      if (BN == 256) return launch_pair<256, 4, true>(ma, mb, mw, p, stream);
      return launch_pair<128, 6, true>(ma, mb, mw, p, stream);
    }
    if (BN == 256) return launch_pair<256, 5, false>(ma, mb, mw, p, stream);
    return launch_pair<128, 8, false>(ma, mb, mw, p, stream);
  }
  if (BN == 128) return launch<128, 3>(ma, mb, mw, p, grid, stream);
  return launch<64, 4>(ma, mb, mw, p, grid, stream);
}
