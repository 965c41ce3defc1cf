// EXPERIMENT (not part of libkidney_b200.so): tcgen05 formulation of Unet.final_conv.  Correct (rel-L2 < 1e-5 vs F.conv2d on five
// shapes, batch-invariant) but 1.8x SLOWER than the shipped mma.sync kernel: 3.0 ms vs 1.67 ms at B = 8, 1024^2.  Every N = 16 MMA
// costs ~140 cycles regardless of how many independent TMEM accumulators are used (1 or 8: same time), i.e. a per-instruction floor
// of the tensor pipe for M = 128 operands, not a dependency chain: 72 MMAs per 128-pixel tile = 10 000 cycles against 1 600 cycles
// of HBM time.  The same floor (~105 cycles per MMA) is what holds the Cout = 128 convolutions at 0.66 of cuBLAS's duty cycle
// (profiles/README.md, round 2, item 9).  Kept for the record; it would need cta_group::2 (M = 256 per instruction) to break even.
// Unet.final_conv (3x3, Cout <= 4 output channels) on cat(x [NHWC fp16, Ca channels], lowres_cond_img [NCHW fp32, Cb <= 4]) as a
// tcgen05 kernel: the op is HBM-bound (one read of the 128-channel activation, 268 MB per 1024^2 image), so the design goal is to
// stream halo tiles at memory speed with as few issue slots as possible -- which the mma.sync version (N = 8 fragments through
// ldmatrix, 65 % of the shared-memory LSU pipe) could not.
//
//   M tile = 16 x 8 output pixels; A = the 18 x 10 halo of a 64-channel chunk brought by ONE TMA box (as in conv_gemm_halo_kernel:
//   the nine taps are UMMA descriptors into it, start + (ky * 10 + kx) * 128 B, SBO = 10 * 128 B);
//   B = the filter as a 16-row operand: rows [0, Cout) hold fp16(w), rows [4, 4 + Cout) hold fp16(w - fp16(w)) -- the hi / lo split
//   that keeps the filter's fp32 precision -- all (chunk, tap) blocks (36 KB at Ca = 128) resident in shared memory for the whole
//   persistent CTA;  D = 16 fp32 columns in TMEM, double-buffered.  The epilogue adds hi + lo columns, the bias and the <= 4 fp32
//   low-res channels (27 FMAs per output from L1-cached global loads) and writes NCHW fp32.
#include <cuda.h>
#include <mutex>

#include "kd_common.cuh"
#include "kd_tc.cuh"

namespace {

constexpr int FT_TW = 8, FT_TH = 16, FT_HW = FT_TW + 2, FT_HH = FT_TH + 2;
constexpr int FT_HALO_BYTES = FT_HH * FT_HW * 128;                         // 23 040 B per 64-channel chunk
constexpr int FT_STAGE_BYTES = ((FT_HALO_BYTES + 1023) / 1024) * 1024;     // 23 552
constexpr int FT_AS = 4;                                                   // halo ring stages
constexpr int FT_WBLK = 16 * 128;                                          // one (chunk, tap) filter block: 16 rows x 64 k
constexpr int FT_THREADS = 192;                                            // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int FT_MAXC = 4;
constexpr int FT_NACC = 8;   // independent accumulators per tile (16 TMEM columns each), used round-robin by the (chunk, tap) MMAs: a chain of
                             // dependent N = 16 MMAs into ONE accumulator is latency-bound (~140 cycles each); the epilogue adds them up
constexpr int FT_TCOLS = 2 * FT_NACC * 16;  // two accumulator stages

struct FinalParams {
  const float* xb;    // [B, Cb, H, W] fp32 or null
  const float* wb;    // [Cout][9][Cb] fp32 (filter taps of the fp32 channels)
  const float* bias;  // [Cout] or null
  float* out;         // [B, Cout, H, W]
  int B, H, W, Ca, Cb, Cout, tiles_w, tiles_h, n_tiles;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(FT_THREADS, 1)
final_conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const FinalParams p) {
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // M 128, N 16, fp16 -> fp32
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int chunks = p.Ca / 64;
  const uint32_t a_s = base, w_s = base + FT_AS * FT_STAGE_BYTES;
  uint8_t* ctrl = gen + FT_AS * FT_STAGE_BYTES + chunks * 9 * FT_WBLK;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);   // FT_AS
  uint64_t* a_empty = a_full + FT_AS;                      // FT_AS
  uint64_t* w_full = a_empty + FT_AS;                      // 1
  uint64_t* t_full = w_full + 1;                           // 2
  uint64_t* t_empty = t_full + 2;                          // 2 (count 4: one lane per epilogue warp)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* s_wb = reinterpret_cast<float*>(ctrl + 128);      // [Cout][9][Cb] + bias[Cout]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < FT_AS; ++s) {
      mbar_init(smem_u32(&a_full[s]), 1);
      mbar_init(smem_u32(&a_empty[s]), 1);
    }
    mbar_init(smem_u32(w_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&t_full[s]), 1);
      mbar_init(smem_u32(&t_empty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), FT_TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  kd_pdl_wait();
  kd_pdl_trigger();
  if (warp >= 2) {  // fp32 side filter + bias (read-only weights: safe to stage after the dependency wait)
    const int nwb = p.Cout * 9 * p.Cb;
    for (int i = threadIdx.x - 64; i < nwb + p.Cout; i += 128) s_wb[i] = i < nwb ? p.wb[i] : (p.bias ? p.bias[i - nwb] : 0.f);
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }

  if (warp == 0) {
    // ================================================================ TMA producer: the whole filter once, then the halo ring
    if (lane == 0) {
      mbar_expect_tx(smem_u32(w_full), (uint32_t)(chunks * 9 * FT_WBLK));
      for (int ch = 0; ch < chunks; ++ch)
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(w_s + (ch * 9 + tap) * FT_WBLK, &map_w, smem_u32(w_full), tap * p.Ca + ch * 64, 0);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, b = t / (p.tiles_w * p.tiles_h);
        for (int ch = 0; ch < chunks; ++ch, ++it) {
          const uint32_t s = it % FT_AS, ph = (it / FT_AS) & 1u;
          mbar_wait_relaxed(smem_u32(&a_empty[s]), ph ^ 1u);
          mbar_expect_tx(smem_u32(&a_full[s]), FT_HALO_BYTES);
          tma_load_5d(a_s + s * FT_STAGE_BYTES, &map_a, smem_u32(&a_full[s]), ch * 64, tw * FT_TW - 1, th * FT_TH - 1, b, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (warp-convergent, one elected lane issues)
    mbar_wait(smem_u32(w_full), 0);
    uint32_t it = 0, tile_iter = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++tile_iter) {
      const uint32_t as = tile_iter & 1u, aph = (tile_iter >> 1) & 1u;
      mbar_wait(smem_u32(&t_empty[as]), aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * (FT_NACC * 16);
      for (int ch = 0; ch < chunks; ++ch, ++it) {
        const uint32_t s = it % FT_AS, ph = (it / FT_AS) & 1u;
        mbar_wait(smem_u32(&a_full[s]), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_stage = a_s + s * FT_STAGE_BYTES;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint64_t a_desc = make_sw128_desc(a_stage + (uint32_t)(ky * FT_HW + kx) * 128u, FT_HW * 128);
            const uint64_t b_desc = make_sw128_desc(w_s + (uint32_t)(ch * 9 + tap) * FT_WBLK);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int seq = (ch * 9 + tap) * 4 + k;  // accumulator seq % FT_NACC; its first use overwrites, later uses accumulate
              umma_f16(d_tmem + (uint32_t)(seq % FT_NACC) * 16u, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), IDESC,
                       seq >= FT_NACC ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&a_empty[s]));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(&t_full[as]));
      __syncwarp();
    }
  } else {
    // ================================================================ epilogue: one output pixel per thread
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int pw = r % FT_TW, ph_ = r / FT_TW;
    const float* s_bias = s_wb + p.Cout * 9 * p.Cb;
    uint32_t tile_iter = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++tile_iter) {
      const uint32_t as = tile_iter & 1u, aph = (tile_iter >> 1) & 1u;
      const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, b = t / (p.tiles_w * p.tiles_h);
      const int h = th * FT_TH + ph_, w = tw * FT_TW + pw;
      const bool ok = h < p.H && w < p.W;
      // the fp32 channels first: independent of the accumulator, overlaps the MMAs of this tile
      float extra[FT_MAXC] = {0.f, 0.f, 0.f, 0.f};
      if (ok && p.Cb > 0) {
        for (int tap = 0; tap < 9; ++tap) {
          const int y = h + tap / 3 - 1, x = w + tap % 3 - 1;
          if (y < 0 || y >= p.H || x < 0 || x >= p.W) continue;
          for (int c = 0; c < p.Cb; ++c) {
            const float a = __ldg(p.xb + (((long)b * p.Cb + c) * p.H + y) * p.W + x);
#pragma unroll
            for (int co = 0; co < FT_MAXC; ++co)
              if (co < p.Cout) extra[co] = fmaf(a, s_wb[(co * 9 + tap) * p.Cb + c], extra[co]);
          }
        }
      }
      mbar_wait_relaxed(smem_u32(&t_full[as]), aph);
      tc_fence_after();
      float sum[2 * FT_MAXC];
#pragma unroll
      for (int j = 0; j < 2 * FT_MAXC; ++j) sum[j] = 0.f;
#pragma unroll
      for (int a = 0; a < FT_NACC; a += 2) {  // fixed order: the result does not depend on timing
        uint32_t acc[2][16];
        tmem_ld16(tmem_base + as * (FT_NACC * 16) + (uint32_t)a * 16u + ((uint32_t)(quarter * 32) << 16), acc[0]);
        tmem_ld16(tmem_base + as * (FT_NACC * 16) + (uint32_t)(a + 1) * 16u + ((uint32_t)(quarter * 32) << 16), acc[1]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int j = 0; j < 2 * FT_MAXC; ++j) sum[j] += __uint_as_float(acc[u][j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&t_empty[as]));
      if (ok) {
#pragma unroll
        for (int co = 0; co < FT_MAXC; ++co)
          if (co < p.Cout)
            p.out[(((long)b * p.Cout + co) * p.H + h) * p.W + w] =
                (sum[co] + sum[FT_MAXC + co]) + extra[co] + s_bias[co];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FT_TCOLS);
  }
}

// filter -> [16][9 * Ca] fp16, K ordered (tap, channel): rows [0, Cout) = hi, rows [4, 4 + Cout) = lo, other rows zero
__global__ void final_conv_tc_pack_kernel(const float* __restrict__ w, int Cout, int Ca, int Ctot, h16* __restrict__ wp, float* __restrict__ wb) {
  const int K = 9 * Ca;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 16 * K; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i % K;
    const int tap = k / Ca, c = k % Ca;
    float v = 0.f;
    const int co = n & 3;
    if (n < 8 && co < Cout) {
      const float wf = w[((long)co * 9 + tap) * Ctot + c];
      const float hi = __half2float(__float2half_rn(wf));
      v = n < 4 ? hi : wf - hi;
    }
    wp[i] = __float2half_rn(v);
  }
  const int Cb = Ctot - Ca;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Cout * 9 * Cb; i += gridDim.x * blockDim.x) {
    const int c = i % Cb, tap = (i / Cb) % 9, co = i / (Cb * 9);
    wb[i] = w[((long)co * 9 + tap) * Ctot + Ca + c];
  }
}

}  // namespace

extern "C" long kd_final_conv_tc_pack_elems(int Ca) { return 16L * 9 * Ca + 2 * FT_MAXC * 9 * FT_MAXC; /* fp16 filter + (as fp16 slots) the fp32 side filter */ }

extern "C" int kd_final_conv_tc_supported(int Ca, int Cb, int Cout, int H, int W) {
  return (Ca > 0 && Ca % 64 == 0 && Ca <= 256 && Cb >= 0 && Cb <= FT_MAXC && Cout > 0 && Cout <= FT_MAXC && H >= FT_TH && W >= FT_TW) ? 1 : 0;
}

extern "C" int kd_final_conv_tc_pack(const float* w, int Cout, int Ca, int Cb, void* w_packed, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(w && w_packed && Cout > 0 && Cout <= FT_MAXC && Ca % 64 == 0 && Cb >= 0 && Cb <= FT_MAXC, "kd_final_conv_tc_pack: bad argument");
  h16* wp = reinterpret_cast<h16*>(w_packed);
  float* wb = reinterpret_cast<float*>(wp + 16L * 9 * Ca);
  final_conv_tc_pack_kernel<<<64, 256, 0, stream>>>(w, Cout, Ca, Ca + Cb, wp, wb);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_final_conv_tc(const void* xa, int Ca, const float* xb, int Cb, const void* w_packed, const float* bias, float* out, int B,
                                int H, int W, int Cout, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(xa && w_packed && out && B > 0, "kd_final_conv_tc: bad argument");
  KD_REQUIRE(kd_final_conv_tc_supported(Ca, Cb, Cout, H, W) && (Cb == 0 || xb), "kd_final_conv_tc: unsupported shape (see kd_final_conv_tc_supported)");
  const h16* wp = reinterpret_cast<const h16*>(w_packed);
  FinalParams p;
  p.xb = xb;
  p.wb = reinterpret_cast<const float*>(wp + 16L * 9 * Ca);
  p.bias = bias;
  p.out = out;
  p.B = B; p.H = H; p.W = W; p.Ca = Ca; p.Cb = Cb; p.Cout = Cout;
  p.tiles_w = kd_ceil_div(W, FT_TW);
  p.tiles_h = kd_ceil_div(H, FT_TH);
  const long long n_tiles = (long long)p.tiles_w * p.tiles_h * B;
  KD_REQUIRE(n_tiles < 2147483647LL, "kd_final_conv_tc: too many tiles");
  p.n_tiles = (int)n_tiles;
  CUtensorMap ma, mw;
  {
    const uint64_t dims[5] = {(uint64_t)Ca, (uint64_t)W, (uint64_t)H, (uint64_t)B, 1ull};
    const uint64_t str[4] = {(uint64_t)Ca * 2, (uint64_t)W * Ca * 2, (uint64_t)H * W * Ca * 2, (uint64_t)B * H * W * Ca * 2};
    const uint32_t box[5] = {64u, (uint32_t)FT_HW, (uint32_t)FT_HH, 1u, 1u};
    int rc = kd_encode_tiled_h16(&ma, xa, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * Ca, 16ull};
    const uint64_t str[1] = {(uint64_t)9 * Ca * 2};
    const uint32_t box[2] = {64u, 16u};
    int rc = kd_encode_tiled_h16(&mw, wp, 2, dims, str, box);
    if (rc) return rc;
  }
  const int smem = FT_AS * FT_STAGE_BYTES + (Ca / 64) * 9 * FT_WBLK + 1024 /*align*/ + 128 /*barriers*/ + (FT_MAXC * 9 * FT_MAXC + FT_MAXC) * 4;
  static int configured_smem = 0;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (configured_smem < smem) {
      KD_CUDA(cudaFuncSetAttribute(final_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured_smem = smem;
    }
  }
  int grid = kd_num_sms();
  if (n_tiles < grid) grid = (int)n_tiles;
  KD_CUDA(kd_launch(final_conv_tc_kernel, dim3(grid), dim3(FT_THREADS), (size_t)smem, stream, ma, mw, p));
  return KD_OK;
}
