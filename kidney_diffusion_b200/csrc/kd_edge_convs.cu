// The two "edge" convolutions of the UNet whose shapes are not tensor-core friendly on their own:
//   * CrossEmbedLayer input convs (Cin = 3..10 image channels, kernels 3/7/15): an im2col panel builder (NCHW fp32 ->
//     [pixels, Kp] h16) feeding kd_conv_gemm mode 2, where the three kernels are merged into one 15x15 weight matrix.
//   * final_conv (3x3, Cout = 3) on cat(x, lowres_cond_img): HBM-bound, shared-memory halo tiles + mma.sync with a split
//     (hi + lo) fp16 filter; also converts NHWC h16 -> NCHW fp32.
#include "kd_common.cuh"

namespace {

constexpr int IC_TW = 64;  // pixels of one image row per block

__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, int C, int H, int W, int ks, h16* __restrict__ out,
                                                     int Kp) {
  extern __shared__ float sm[];
  const int pad = ks >> 1;
  const int halo_w = IC_TW + ks - 1;
  float* tile = sm;                                              // [C][ks][halo_w]
  int* lut = reinterpret_cast<int*>(sm + (size_t)C * ks * halo_w);  // [Kp] -> offset in tile or -1
  const int b = blockIdx.z, h = blockIdx.y, w0 = blockIdx.x * IC_TW;
  const int Kreal = ks * ks * C;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    int v = -1;
    if (k < Kreal) {
      const int tap = k / C, c = k - tap * C;
      const int ky = tap / ks, kx = tap - ky * ks;
      v = (c * ks + ky) * halo_w + kx;
    }
    lut[k] = v;
  }
  const int tile_n = C * ks * halo_w;
  for (int i = threadIdx.x; i < tile_n; i += blockDim.x) {
    const int xx = i % halo_w;
    const int ky = (i / halo_w) % ks;
    const int c = i / (halo_w * ks);
    const int gy = h + ky - pad, gx = w0 + xx - pad;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = x[(((long)b * C + c) * H + gy) * W + gx];
    tile[i] = v;
  }
  __syncthreads();
  const int groups = Kp >> 3;
  const int npx = (W - w0) < IC_TW ? (W - w0) : IC_TW;
  for (int i = threadIdx.x; i < npx * groups; i += blockDim.x) {
    const int px = i / groups, g = i - px * groups;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int off = lut[g * 8 + j];
      v[j] = off >= 0 ? tile[off + px] : 0.f;
    }
    h16x8 o8 = float_to_h16x8(v);
    h16* dst = out + (((long)b * H + h) * W + w0 + px) * Kp + g * 8;
    st_stream(dst, *reinterpret_cast<int4*>(&o8));
  }
}

// ------------------------------------------------------------------------------------------------ final conv
// 3x3, Cout <= 4: HBM-bound (one read of the 128-channel activation).  Per 8 x 32 pixel tile the halo is staged in shared
// memory 32 channels at a time; each warp owns one tile row (two m16 pixel groups) and runs mma.sync m16n8k16 with the
// fp32 filter split into fp16 hi + lo parts (two MMAs), so the filter keeps ~22 significant bits as in the fp32 reference.
// The <= 4 fp32 NCHW extra channels (lowres_cond_img) and the bias are added per pixel in fp32 SIMT, which is also the
// coalesced NCHW writer.
constexpr int FC_TH = 8, FC_TW = 32, FC_CH = 32;           // 256 pixels per block, 32-channel chunks
constexpr int FC_HH = FC_TH + 2, FC_HW = FC_TW + 2;         // halo tile
constexpr int FC_PIX_STRIDE = FC_CH + 8;                    // h16 elements per halo pixel (80 B: conflict-free ldmatrix rows)
constexpr int FC_MAXCO = 4;
constexpr int FC_WK = FC_CH + 8;                            // padded k stride of the staged filter rows

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// filter pre-split for the tensor cores, once per model: [chunk][hi | lo][tap][n = 8][FC_WK] fp16 (rows n >= Cout and the
// k padding are zero), so a block fetches a chunk's filter with 720 16-byte copies instead of 864 scalar loads + splits
constexpr int FC_WCHUNK = 2 * 9 * 8 * FC_WK;  // h16 elements per 32-channel chunk (11 520 B)

__global__ void final_conv_pack_kernel(const float* __restrict__ w, int Cout, int Ca, int Ctot, h16* __restrict__ wp) {
  const int total = (Ca / FC_CH) * FC_WCHUNK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % FC_WK, n = (i / FC_WK) % 8, tap = (i / (FC_WK * 8)) % 9, part = (i / (FC_WK * 72)) % 2, chunk = i / FC_WCHUNK;
    float v = 0.f;
    if (k < FC_CH && n < Cout) {
      const float wf = w[((long)n * 9 + tap) * Ctot + chunk * FC_CH + k];
      const h16 hi = __float2half_rn(wf);
      v = part == 0 ? __half2float(hi) : wf - __half2float(hi);
    }
    wp[i] = __float2half_rn(v);
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled (conv padding)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(256) final_conv_kernel(const h16* __restrict__ xa, int Ca, const float* __restrict__ xb, int Cb,
                                                         const float* __restrict__ w, const h16* __restrict__ wp,
                                                         const float* __restrict__ bias, float* __restrict__ out, int H, int W,
                                                         int Cout) {
  kd_pdl_wait();  // programmatic dependent launch: this grid may have been scheduled while its stream predecessor drains
  kd_pdl_trigger();
  extern __shared__ __align__(16) uint8_t fc_smem[];
  constexpr int ACT_BYTES = FC_HH * FC_HW * FC_PIX_STRIDE * 2;  // 27 200
  constexpr int W_BYTES = FC_WCHUNK * 2;                       // 11 520
  // [2 x activations][2 x filter] double-buffered by cp.async, then the fp32 tails
  h16* s_w_base = reinterpret_cast<h16*>(fc_smem + 2 * ACT_BYTES);
  float* s_xb = reinterpret_cast<float*>(fc_smem + 2 * ACT_BYTES + 2 * W_BYTES);
  float* s_wb = s_xb + 4 * FC_HH * FC_HW;
  float* s_res = s_wb + FC_MAXCO * 9 * 4;  // [FC_MAXCO][FC_TH][FC_TW]
  const uint32_t s_u32 = (uint32_t)__cvta_generic_to_shared(fc_smem);
  const int Ctot = Ca + Cb;
  const int b = blockIdx.z, h0 = blockIdx.y * FC_TH, w0 = blockIdx.x * FC_TW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};

  auto issue_chunk = [&](int chunk, int buf) {
    const uint32_t act_dst = s_u32 + buf * ACT_BYTES;
    for (int i = threadIdx.x; i < FC_HH * FC_HW * 4; i += 256) {
      const int v = i & 3, pix = i >> 2;
      const int py = pix / FC_HW, px = pix % FC_HW;
      const int gy = h0 + py - 1, gx = w0 + px - 1;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const h16* src = xa + (((long)b * H + (ok ? gy : 0)) * W + (ok ? gx : 0)) * Ca + chunk * FC_CH + v * 8;
      cp_async16(act_dst + (pix * FC_PIX_STRIDE + v * 8) * 2, src, ok);
    }
    const uint32_t w_dst = s_u32 + 2 * ACT_BYTES + buf * W_BYTES;
    const h16* wsrc = wp + (long)chunk * FC_WCHUNK;
    for (int i = threadIdx.x; i < W_BYTES / 16; i += 256) cp_async16(w_dst + i * 16, wsrc + i * 8, true);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // ldmatrix lane addressing of an m16 x k16 A fragment: matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15)
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int a_kof = (lane >> 4) * 8;
  const int n_chunks = Ca / FC_CH;
  issue_chunk(0, 0);
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int buf = chunk & 1;
    if (chunk + 1 < n_chunks) {
      issue_chunk(chunk + 1, buf ^ 1);  // the buffer was released by the barrier that ended chunk - 1
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t act_u32 = s_u32 + buf * ACT_BYTES;
    const h16* sw = s_w_base + buf * FC_WCHUNK;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int py = warp + tap / 3, dx = tap % 3;
#pragma unroll
      for (int ks = 0; ks < FC_CH / 16; ++ks) {
        const h16* wh = sw + ((0 * 9 + tap) * 8 + (lane >> 2)) * FC_WK + ks * 16 + (lane & 3) * 2;
        const h16* wl = sw + ((1 * 9 + tap) * 8 + (lane >> 2)) * FC_WK + ks * 16 + (lane & 3) * 2;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(wh), bh1 = *reinterpret_cast<const uint32_t*>(wh + 8);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(wl), bl1 = *reinterpret_cast<const uint32_t*>(wl + 8);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t a0, a1, a2, a3;
          const int px = mt * 16 + a_row + dx;
          ldmatrix_x4(act_u32 + ((py * FC_HW + px) * FC_PIX_STRIDE + ks * 16 + a_kof) * 2, a0, a1, a2, a3);
          mma_16816(acc[mt], a0, a1, a2, a3, bh0, bh1);
          mma_16816(acc[mt], a0, a1, a2, a3, bl0, bl1);
        }
      }
    }
    __syncthreads();
  }
  // accumulator fragment: rows lane/4 and lane/4 + 8 of the m-tile, columns (lane%4)*2 + {0, 1}
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int n0 = (lane & 3) * 2, r = lane >> 2;
    if (n0 < FC_MAXCO) {
      s_res[(n0 * FC_TH + warp) * FC_TW + mt * 16 + r] = acc[mt][0];
      s_res[((n0 + 1) * FC_TH + warp) * FC_TW + mt * 16 + r] = acc[mt][1];
      s_res[(n0 * FC_TH + warp) * FC_TW + mt * 16 + r + 8] = acc[mt][2];
      s_res[((n0 + 1) * FC_TH + warp) * FC_TW + mt * 16 + r + 8] = acc[mt][3];
    }
  }
  const int ty = threadIdx.x / FC_TW, tx = threadIdx.x % FC_TW;
  float extra[FC_MAXCO] = {0.f, 0.f, 0.f, 0.f};
  if (Cb > 0) {
    for (int i = threadIdx.x; i < Cb * FC_HH * FC_HW; i += 256) {
      const int pix = i % (FC_HH * FC_HW), c = i / (FC_HH * FC_HW);
      const int py = pix / FC_HW, px = pix % FC_HW;
      const int gy = h0 + py - 1, gx = w0 + px - 1;
      s_xb[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? xb[(((long)b * Cb + c) * H + gy) * W + gx] : 0.f;
    }
    for (int i = threadIdx.x; i < Cout * 9 * Cb; i += 256) {
      const int c = i % Cb, tap = (i / Cb) % 9, co = i / (Cb * 9);
      s_wb[i] = w[((long)co * 9 + tap) * Ctot + Ca + c];
    }
  }
  __syncthreads();
  if (Cb > 0) {
    for (int tap = 0; tap < 9; ++tap) {
      const int py = ty + tap / 3, px = tx + tap % 3;
      for (int c = 0; c < Cb; ++c) {
        const float a = s_xb[c * FC_HH * FC_HW + py * FC_HW + px];
#pragma unroll
        for (int co = 0; co < FC_MAXCO; ++co)
          if (co < Cout) extra[co] += a * s_wb[(co * 9 + tap) * Cb + c];
      }
    }
  }
  const int gy = h0 + ty, gx = w0 + tx;
  if (gy < H && gx < W) {
#pragma unroll
    for (int co = 0; co < FC_MAXCO; ++co)
      if (co < Cout) out[(((long)b * Cout + co) * H + gy) * W + gx] = s_res[(co * FC_TH + ty) * FC_TW + tx] + extra[co] + (bias ? bias[co] : 0.f);
  }
}

constexpr int FC_SMEM = 2 * (FC_HH * FC_HW * FC_PIX_STRIDE * 2) + 2 * (FC_WCHUNK * 2) + (4 * FC_HH * FC_HW + FC_MAXCO * 9 * 4 + FC_MAXCO * FC_TH * FC_TW) * 4;

}  // namespace

extern "C" int kd_im2col_nchw(const float* x, int B, int C, int H, int W, int ksize, void* out, int Kp, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0, "kd_im2col_nchw: bad argument");
  KD_REQUIRE(ksize % 2 == 1 && ksize <= 15, "kd_im2col_nchw: ksize must be odd and <= 15");
  KD_REQUIRE(Kp % 64 == 0 && Kp >= ksize * ksize * C, "kd_im2col_nchw: Kp=%d must be a multiple of 64 covering %d", Kp, ksize * ksize * C);
  const size_t smem = sizeof(float) * (size_t)C * ksize * (IC_TW + ksize - 1) + sizeof(int) * (size_t)Kp;
  KD_REQUIRE(smem <= 160 * 1024, "kd_im2col_nchw: tile does not fit shared memory (C=%d)", C);
  static bool configured = false;
  if (!configured) {
    KD_CUDA(cudaFuncSetAttribute(im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  KD_REQUIRE(H <= 65535 && B <= 65535, "kd_im2col_nchw: H / B exceed grid limits");
  dim3 grid(kd_ceil_div(W, IC_TW), H, B);
  im2col_kernel<<<grid, 256, smem, stream>>>(x, C, H, W, ksize, reinterpret_cast<h16*>(out), Kp);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" long kd_final_conv_pack_elems(int Ca) { return (long)(Ca / FC_CH) * FC_WCHUNK; }

extern "C" int kd_final_conv_pack(const float* w, int Cout, int Ca, int Cb, void* w_split, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(w && w_split && Cout >= 1 && Cout <= FC_MAXCO && Ca > 0 && Ca % FC_CH == 0 && Cb >= 0, "kd_final_conv_pack: bad argument");
  final_conv_pack_kernel<<<64, 256, 0, stream>>>(w, Cout, Ca, Ca + Cb, reinterpret_cast<h16*>(w_split));
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_final_conv(const void* xa, int Ca, const float* xb, int Cb, const float* w, const void* w_split, const float* bias,
                             float* out, int B, int H, int W, int Cout, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(xa && w && w_split && out && B > 0 && H > 0 && W > 0, "kd_final_conv: bad argument");
  KD_REQUIRE(Ca > 0 && Ca % FC_CH == 0, "kd_final_conv: Ca=%d must be a multiple of %d", Ca, FC_CH);
  KD_REQUIRE(Cb >= 0 && Cb <= 4 && (Cb == 0 || xb), "kd_final_conv: Cb must be <= 4");
  KD_REQUIRE(Cout >= 1 && Cout <= FC_MAXCO, "kd_final_conv: Cout must be <= %d", FC_MAXCO);
  KD_REQUIRE(kd_ceil_div(H, FC_TH) <= 65535 && B <= 65535, "kd_final_conv: grid too large");
  static bool configured = false;
  if (!configured) {
    KD_CUDA(cudaFuncSetAttribute(final_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM));
    configured = true;
  }
  dim3 grid(kd_ceil_div(W, FC_TW), kd_ceil_div(H, FC_TH), B);
  KD_CUDA(kd_launch(final_conv_kernel, dim3(grid), dim3(256), FC_SMEM, stream, reinterpret_cast<const h16*>(xa), Ca, xb, Cb, w, reinterpret_cast<const h16*>(w_split), bias, out, H, W, Cout));
  return KD_OK;
}
