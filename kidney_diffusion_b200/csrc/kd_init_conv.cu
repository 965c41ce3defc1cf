// K2: the CrossEmbedLayer input convolution (image channels C <= 3, kernels 3 / 7 / 15 merged into one ks x ks filter)
// as an implicit GEMM on tcgen05 with NO materialised im2col panel.
//
//   out[b, y, x, n] = bias[n] + addend[b, y, x, n] + sum_{ky, c, kx} img[b, c, y + ky - pad, x + kx - pad] * Wp[n, (ky*C + c)*16 + kx]
//
// With only 3 channels the K axis has to run over the filter window, and a window row is a Toeplitz matrix
// A[x][kx] = halo[x + kx]: consecutive pixels read overlapping, element-shifted spans, which no 16-byte-granular UMMA
// descriptor can express on a plain row.  So every halo row is kept in shared memory as 8 element-shifted copies
// interleaved at 16-byte granularity: block j (128 B) = units s = 0..7, unit s = halo[8j + s .. 8j + s + 7].  Block j is
// then exactly the 8-row x 16-byte core matrix of pixels 8j .. 8j+7 at kx = 0..7, and block j + 1 the one at kx = 8..15,
// i.e. a K-major SWIZZLE_NONE operand with LBO = SBO = 128 bytes: one tcgen05.mma (M = 128 pixels of an image row,
// K = 16 = one (ky, c) window row) reads it directly.  A CTA slides down a 128-column strip: a ring of ks + 3 halo rows
// stays resident, every pass computes two output rows (two TMEM accumulators share each TMA-streamed weight chunk), and
// each new output row costs one new halo row of SIMT fill (7 KB) instead of a 180 KB im2col panel row.
//
// Warp roles (448 threads, persistent CTA, one per SM): warp 0 = weight TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2-9 = epilogue (one set of 4 per output row: TMEM -> bias / addend -> fp16 NHWC), warps 10-13 = halo-row fillers
// (NCHW fp32 -> shifted fp16).
#include "kd_tc.cuh"

namespace {

constexpr int IC_M = 128;                       // pixels of one image row per MMA
constexpr int IC_BLOCKS = 17;                   // 128-byte blocks per (halo row, channel): covers x + kx <= 127 + 15
constexpr int IC_ROWC_BYTES = IC_BLOCKS * 128;  // 2176
constexpr int IC_HALO_W = IC_BLOCKS * 8 + 7;    // 143 source elements
constexpr int IC_BSTAGES = 4;
constexpr int IC_THREADS = 448;
constexpr int IC_FILL_THREADS = 128;

struct InitParams {
  const float* x;
  int B, C, H, W, ks, Cout;
  int rows_seg, n_strips, n_segs, n_units;
  int n_pairs, n_chunks, ring;
  const float* bias;
  const h16* addend;
  h16* out;
};

template <int BN>
__global__ void __launch_bounds__(IC_THREADS, 1) init_conv_kernel(const __grid_constant__ CUtensorMap map_w, const InitParams p) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  constexpr int B_STAGE_BYTES = BN * 128;
  constexpr uint32_t TMEM_COLS = 4 * BN;  // 2 buffers x 2 output rows
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(IC_M >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ring_base = smem_base + IC_BSTAGES * B_STAGE_BYTES;
  uint8_t* ring_gen = smem_gen + IC_BSTAGES * B_STAGE_BYTES;
  const int ring_bytes = p.ring * p.C * IC_ROWC_BYTES;
  uint8_t* stage_gen = ring_gen + ring_bytes;  // 8 epilogue warps x 4 KB
  uint8_t* ctrl = stage_gen + 8 * 4096;
  uint64_t* b_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* b_empty = b_full + IC_BSTAGES;
  uint64_t* rows_full = b_empty + IC_BSTAGES;  // [2]
  uint64_t* tmem_full = rows_full + 2;         // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_smem = reinterpret_cast<float*>(ctrl + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < IC_BSTAGES; ++s) {
      mbar_init(smem_u32(&b_full[s]), 1);
      mbar_init(smem_u32(&b_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&rows_full[a]), IC_FILL_THREADS);
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
  if (warp >= 2 && warp < 6) {
    for (int j = threadIdx.x - 64; j < BN; j += 128) bias_smem[j] = (p.bias != nullptr && j < p.Cout) ? p.bias[j] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int pad = p.ks >> 1;

  if (warp == 0) {
    // ================================================================ weight chunks: [BN, 64] boxes, re-streamed from L2 per pass
    if (lane == 0) {
      uint32_t cc = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int seg = (u / p.n_strips) % p.n_segs;
        const int r0 = seg * p.rows_seg;
        const int rows = min(p.rows_seg, p.H - r0);
        const int passes = (rows + 1) >> 1;
        for (int pl = 0; pl < passes; ++pl) {
          for (int j = 0; j < p.n_chunks; ++j, ++cc) {
            const int s = cc % IC_BSTAGES;
            mbar_wait_relaxed(smem_u32(&b_empty[s]), ((cc / IC_BSTAGES) & 1) ^ 1u);
            const uint32_t fb = smem_u32(&b_full[s]);
            mbar_expect_tx(fb, B_STAGE_BYTES);
            tma_load_2d(smem_base + s * B_STAGE_BYTES, &map_w, fb, j * 64, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    // (one thread issues ~90 MMAs per pass: all operand addresses come from wrap-around counters, no div / mod on this path)
    {
      uint32_t cc = 0, q = 0;
      const uint64_t a_desc0 = make_nosw_desc(ring_base, 128, 128);
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int seg = (u / p.n_strips) % p.n_segs;
        const int r0 = seg * p.rows_seg;
        const int rows = min(p.rows_seg, p.H - r0);
        const int passes = (rows + 1) >> 1;
        int s0 = 0;  // ring slot of halo row 2 * pl
        for (int pl = 0; pl < passes; ++pl, ++q) {
          const uint32_t a = q & 1, ph = (q >> 1) & 1;
          mbar_wait(smem_u32(&tmem_empty[a]), ph ^ 1u);
          mbar_wait(smem_u32(&rows_full[a]), ph);
          tc_fence_after();
          const uint32_t d0 = tmem_base + a * (2 * BN);
          int c = 0, slot = s0, pair = 0;
          for (int j = 0; j < p.n_chunks; ++j, ++cc) {
            const int s = cc % IC_BSTAGES;
            mbar_wait(smem_u32(&b_full[s]), (cc / IC_BSTAGES) & 1);
            tc_fence_after();
            const uint64_t b_desc = make_sw128_desc(smem_base + s * B_STAGE_BYTES);
            const bool leader = elect_one();
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (pair < p.n_pairs) {
                const int slot1 = (slot + 1 == p.ring) ? 0 : slot + 1;
                if (leader) {
                  const uint32_t acc = pair != 0 ? 1u : 0u;
                  umma_f16(d0, a_desc0 + (uint64_t)((slot * p.C + c) * (IC_ROWC_BYTES >> 4)), b_desc + (uint64_t)(2 * t), IDESC, acc);
                  umma_f16(d0 + BN, a_desc0 + (uint64_t)((slot1 * p.C + c) * (IC_ROWC_BYTES >> 4)), b_desc + (uint64_t)(2 * t), IDESC,
                           acc);
                }
                ++pair;
                if (++c == p.C) {
                  c = 0;
                  slot = slot1;
                }
              }
            }
            if (leader) umma_commit(smem_u32(&b_empty[s]));
            __syncwarp();
          }
          if (elect_one()) umma_commit(smem_u32(&tmem_full[a]));
          __syncwarp();
          s0 += 2;
          if (s0 >= p.ring) s0 -= p.ring;
        }
      }
    }
  } else if (warp < 10) {
    // ================================================================ epilogue: warp set o (4 warps) owns output row o of a pass
    const int quarter = warp & 3;
    const int o = (warp - 2) >> 2;
    const int col = quarter * 32 + lane;
    uint8_t* stage_w = stage_gen + (warp - 2) * 4096;
    uint32_t q = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int strip = u % p.n_strips;
      const int seg = (u / p.n_strips) % p.n_segs;
      const int b = u / (p.n_strips * p.n_segs);
      const int r0 = seg * p.rows_seg;
      const int rows = min(p.rows_seg, p.H - r0);
      const int passes = (rows + 1) >> 1;
      const int x = strip * IC_M + col;
      for (int pl = 0; pl < passes; ++pl, ++q) {
        const uint32_t a = q & 1, ph = (q >> 1) & 1;
        const int y = r0 + 2 * pl + o;
        const bool ok = (y < r0 + rows) && (x < p.W);
        const long long off = (((long long)b * p.H + y) * p.W + x) * p.Cout;
        // the whole addend row of this pixel is requested before waiting for the accumulator (hides the HBM latency)
        int4 add[BN / 8];
        if (ok && p.addend != nullptr) {
#pragma unroll
          for (int g = 0; g < BN / 8; ++g) add[g] = *reinterpret_cast<const int4*>(p.addend + off + g * 8);  // may alias out
        }
        mbar_wait_relaxed(smem_u32(&tmem_full[a]), ph);
        tc_fence_after();
        // 64-channel halves are transposed through a per-warp 4 KB staging tile (16-byte units XOR-swizzled by pixel) so that
        // every global store instruction writes four whole 128-byte pixel-halves
        const int x_warp = strip * IC_M + quarter * 32;
        const bool row_live = (y < r0 + rows);
        const long long off_warp = (((long long)b * p.H + y) * p.W + x_warp) * p.Cout;
#pragma unroll
        for (int chunk = 0; chunk < BN / 32; ++chunk) {
          uint32_t acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + a * (2 * BN) + o * BN + chunk * 32, acc);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g * 8 + j]) + bias_smem[chunk * 32 + g * 8 + j];
            if (ok && p.addend != nullptr) {
              float av[8];
              h16x8_to_float(*reinterpret_cast<const h16x8*>(&add[chunk * 4 + g]), av);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += av[j];
            }
            h16x8 o8 = float_to_h16x8(v);
            const int unit = (chunk & 1) * 4 + g;
            *reinterpret_cast<h16x8*>(stage_w + lane * 128 + ((unit ^ (lane & 7)) << 4)) = o8;
          }
          if (chunk & 1) {
            __syncwarp();
            if (row_live) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int px = i * 4 + (lane >> 3), unit = lane & 7;
                const int4 val = *reinterpret_cast<const int4*>(stage_w + px * 128 + ((unit ^ (px & 7)) << 4));
                if (x_warp + px < p.W) st_stream(p.out + off_warp + (long long)px * p.Cout + (chunk >> 1) * 64 + unit * 8, val);
              }
            }
            __syncwarp();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[a]));
      }
    }
  } else {
    // ================================================================ halo-row fillers
    const int f = threadIdx.x - 320;
    const int units_per_row = p.C * IC_BLOCKS * 8;  // 16-byte units of one halo row (all channels)
    uint32_t q = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int strip = u % p.n_strips;
      const int seg = (u / p.n_strips) % p.n_segs;
      const int b = u / (p.n_strips * p.n_segs);
      const int r0 = seg * p.rows_seg;
      const int rows = min(p.rows_seg, p.H - r0);
      const int passes = (rows + 1) >> 1;
      const int gx0 = strip * IC_M - pad;
      const float* img = p.x + (long long)b * p.C * p.H * p.W;
      for (int pl = 0; pl < passes; ++pl, ++q) {
        int rr0, nrows;
        if (pl == 0) {
          rr0 = 0;
          nrows = p.ks + 1;
          if (q >= 1) mbar_wait_relaxed(smem_u32(&tmem_full[(q - 1) & 1]), ((q - 1) >> 1) & 1);
        } else {
          rr0 = 2 * pl + p.ks - 1;
          nrows = 2;
          if (q >= 2) mbar_wait_relaxed(smem_u32(&tmem_full[q & 1]), ((q - 2) >> 1) & 1);
        }
        const int total = nrows * units_per_row;
        for (int i = f; i < total; i += IC_FILL_THREADS) {
          const int rr = rr0 + i / units_per_row;
          const int rem = i % units_per_row;
          const int c = rem / (IC_BLOCKS * 8);
          const int un = rem % (IC_BLOCKS * 8);  // block * 8 + shift
          const int e0 = (un >> 3) * 8 + (un & 7);
          const int gy = r0 + rr - pad;
          float v[8];
          if (gy >= 0 && gy < p.H) {
            const float* src = img + ((long long)c * p.H + gy) * p.W;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int gx = gx0 + e0 + j;
              v[j] = (gx >= 0 && gx < p.W) ? __ldg(src + gx) : 0.0f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.0f;
          }
          h16x8 o8 = float_to_h16x8(v);
          const int slot = rr % p.ring;
          *reinterpret_cast<h16x8*>(ring_gen + (slot * p.C + c) * IC_ROWC_BYTES + un * 16) = o8;
        }
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&rows_full[q & 1]));
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

template <int BN>
int launch_init(const CUtensorMap& mw, const InitParams& p, cudaStream_t stream) {
  const size_t smem = 1024 + IC_BSTAGES * BN * 128 + (size_t)p.ring * p.C * IC_ROWC_BYTES + 8 * 4096 + 128 + BN * sizeof(float);
  KD_REQUIRE(smem <= 227 * 1024, "kd_init_conv: shared memory %zu exceeds the SM limit", smem);
  static size_t configured = 0;
  if (smem > configured) {
    KD_CUDA(cudaFuncSetAttribute(init_conv_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = p.n_units < kd_num_sms() ? p.n_units : kd_num_sms();
  KD_CUDA(kd_launch(init_conv_kernel<BN>, dim3(grid), dim3(IC_THREADS), smem, stream, mw, p));
  return KD_OK;
}

}  // namespace

extern "C" int kd_init_conv_kp(int C, int ksize) { return ((ksize * C * 16 + 63) / 64) * 64; }

extern "C" int kd_init_conv(const float* x, int B, int C, int H, int W, int ksize, const void* w_packed, const float* bias,
                            const void* addend, void* out, int Cout, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && w_packed && out && B > 0 && H > 0 && W > 0, "kd_init_conv: bad argument");
  KD_REQUIRE(C >= 1 && C <= 3, "kd_init_conv: C=%d image channels per call must be 1..3", C);
  KD_REQUIRE(ksize % 2 == 1 && ksize >= 1 && ksize <= 15, "kd_init_conv: ksize must be odd and <= 15");
  KD_REQUIRE(Cout == 64 || Cout == 128, "kd_init_conv: Cout=%d must be 64 or 128", Cout);
  InitParams p;
  p.x = x;
  p.B = B; p.C = C; p.H = H; p.W = W; p.ks = ksize; p.Cout = Cout;
  p.n_strips = kd_ceil_div(W, IC_M);
  // rows per work unit: as long as possible (the ks - 1 halo rows are refilled per unit) while still giving every SM work
  int rows_seg = 64;
  while (rows_seg > 16 && (long long)p.n_strips * kd_ceil_div(H, rows_seg) * B < 3LL * kd_num_sms()) rows_seg >>= 1;
  p.rows_seg = rows_seg;
  p.n_segs = kd_ceil_div(H, rows_seg);
  p.n_units = p.n_strips * p.n_segs * B;
  p.n_pairs = ksize * C;
  p.n_chunks = kd_ceil_div(p.n_pairs, 4);
  p.ring = ksize + 3;
  p.bias = bias;
  p.addend = reinterpret_cast<const h16*>(addend);
  p.out = reinterpret_cast<h16*>(out);
  CUtensorMap mw;
  const uint64_t dims[2] = {(uint64_t)p.n_chunks * 64, (uint64_t)Cout};
  const uint64_t str[1] = {(uint64_t)p.n_chunks * 64 * 2};
  const uint32_t box[2] = {64, (uint32_t)Cout};
  int rc = kd_encode_tiled_h16(&mw, w_packed, 2, dims, str, box);
  if (rc != KD_OK) return rc;
  return Cout == 64 ? launch_init<64>(mw, p, stream) : launch_init<128>(mw, p, stream);
}
