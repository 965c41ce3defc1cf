// The two "edge" convolutions of the UNet whose shapes are not tensor-core friendly on their own:
//   * CrossEmbedLayer input convs (Cin = 3..10 image channels, kernels 3/7/15): an im2col panel builder (NCHW fp32 ->
//     [pixels, Kp] h16) feeding kd_conv_gemm mode 2, where the three kernels are merged into one 15x15 weight matrix.
//   * final_conv (3x3, Cout = 3) on cat(x, lowres_cond_img): HBM-bound, a shared-memory tiled SIMT kernel that also
//     converts NHWC h16 -> NCHW fp32.
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
constexpr int FC_TH = 8, FC_TW = 32, FC_CH = 32;           // 256 pixels per block, 32-channel chunks
constexpr int FC_HH = FC_TH + 2, FC_HW = FC_TW + 2;         // halo tile
constexpr int FC_PIX_STRIDE = FC_CH + 8;                    // h16 elements per halo pixel (80 B: conflict-free 16 B reads)
constexpr int FC_MAXCO = 4;

__global__ void __launch_bounds__(256) final_conv_kernel(const h16* __restrict__ xa, int Ca, const float* __restrict__ xb, int Cb,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ out, int H, int W, int Cout) {
  __shared__ __align__(16) h16 s_act[FC_HH * FC_HW * FC_PIX_STRIDE];
  __shared__ __align__(16) float s_w[FC_MAXCO * 9 * FC_CH];
  __shared__ float s_xb[4 * FC_HH * FC_HW];
  const int Ctot = Ca + Cb;
  const int b = blockIdx.z, h0 = blockIdx.y * FC_TH, w0 = blockIdx.x * FC_TW;
  const int ty = threadIdx.x / FC_TW, tx = threadIdx.x % FC_TW;
  float acc[FC_MAXCO] = {0.f, 0.f, 0.f, 0.f};

  for (int c0 = 0; c0 < Ca; c0 += FC_CH) {
    __syncthreads();
    // halo activations: FC_HH*FC_HW pixels x 4 vectors of 8 channels
    for (int i = threadIdx.x; i < FC_HH * FC_HW * 4; i += 256) {
      const int v = i & 3, pix = i >> 2;
      const int py = pix / FC_HW, px = pix % FC_HW;
      const int gy = h0 + py - 1, gx = w0 + px - 1;
      int4 val = make_int4(0, 0, 0, 0);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) val = ld_stream(xa + (((long)b * H + gy) * W + gx) * Ca + c0 + v * 8);
      *reinterpret_cast<int4*>(&s_act[pix * FC_PIX_STRIDE + v * 8]) = val;
    }
    for (int i = threadIdx.x; i < Cout * 9 * FC_CH; i += 256) {
      const int c = i % FC_CH, tap = (i / FC_CH) % 9, co = i / (FC_CH * 9);
      s_w[i] = w[((long)co * 9 + tap) * Ctot + c0 + c];
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int py = ty + tap / 3, px = tx + tap % 3;
      const h16* ap = &s_act[(py * FC_HW + px) * FC_PIX_STRIDE];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float a[8];
        h16x8_to_float(*reinterpret_cast<const h16x8*>(ap + v * 8), a);
#pragma unroll
        for (int co = 0; co < FC_MAXCO; ++co) {
          if (co < Cout) {
            const float4 wa = *reinterpret_cast<const float4*>(&s_w[(co * 9 + tap) * FC_CH + v * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&s_w[(co * 9 + tap) * FC_CH + v * 8 + 4]);
            acc[co] += a[0] * wa.x + a[1] * wa.y + a[2] * wa.z + a[3] * wa.w + a[4] * wb.x + a[5] * wb.y + a[6] * wb.z + a[7] * wb.w;
          }
        }
      }
    }
  }
  if (Cb > 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < Cb * FC_HH * FC_HW; i += 256) {
      const int pix = i % (FC_HH * FC_HW), c = i / (FC_HH * FC_HW);
      const int py = pix / FC_HW, px = pix % FC_HW;
      const int gy = h0 + py - 1, gx = w0 + px - 1;
      s_xb[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? xb[(((long)b * Cb + c) * H + gy) * W + gx] : 0.f;
    }
    for (int i = threadIdx.x; i < Cout * 9 * Cb; i += 256) {
      const int c = i % Cb, tap = (i / Cb) % 9, co = i / (Cb * 9);
      s_w[i] = w[((long)co * 9 + tap) * Ctot + Ca + c];
    }
    __syncthreads();
    for (int tap = 0; tap < 9; ++tap) {
      const int py = ty + tap / 3, px = tx + tap % 3;
      for (int c = 0; c < Cb; ++c) {
        const float a = s_xb[c * FC_HH * FC_HW + py * FC_HW + px];
#pragma unroll
        for (int co = 0; co < FC_MAXCO; ++co)
          if (co < Cout) acc[co] += a * s_w[(co * 9 + tap) * Cb + c];
      }
    }
  }
  const int gy = h0 + ty, gx = w0 + tx;
  if (gy < H && gx < W) {
#pragma unroll
    for (int co = 0; co < FC_MAXCO; ++co)
      if (co < Cout) out[(((long)b * Cout + co) * H + gy) * W + gx] = acc[co] + (bias ? bias[co] : 0.f);
  }
}

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

extern "C" int kd_final_conv(const void* xa, int Ca, const float* xb, int Cb, const float* w, const float* bias, float* out, int B,
                             int H, int W, int Cout, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(xa && w && out && B > 0 && H > 0 && W > 0, "kd_final_conv: bad argument");
  KD_REQUIRE(Ca > 0 && Ca % FC_CH == 0, "kd_final_conv: Ca=%d must be a multiple of %d", Ca, FC_CH);
  KD_REQUIRE(Cb >= 0 && Cb <= 4 && (Cb == 0 || xb), "kd_final_conv: Cb must be <= 4");
  KD_REQUIRE(Cout >= 1 && Cout <= FC_MAXCO, "kd_final_conv: Cout must be <= %d", FC_MAXCO);
  KD_REQUIRE(kd_ceil_div(H, FC_TH) <= 65535 && B <= 65535, "kd_final_conv: grid too large");
  dim3 grid(kd_ceil_div(W, FC_TW), kd_ceil_div(H, FC_TH), B);
  final_conv_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const h16*>(xa), Ca, xb, Cb, w, bias, out, H, W, Cout);
  KD_LAUNCH_CHECK();
  return KD_OK;
}
