// Patch-grid sampler support (sample_ultra_res.py:213-261, 304-400, 430-446), one process per GPU:
//   * peer mailbox: device memory exported with CUDA IPC and mapped by the other ranks of the node; a producer copies the
//     overlap strip a remote dependent needs straight into the consumer's mailbox over NVLink and then raises a flag word
//     there; the consumer's stream waits on the flag on the device.  No host synchronisation, no collective, no matching
//     send / receive order (replaces the reference's Manager-dict pickling of whole patches through the CPU, :204-205).
//   * get_cond_images as a 1024^2 window gather (roll + fill + centre-crop of :358-395 folded into index arithmetic).
//   * stitch (:440-446) as owner-computes paste: every canvas pixel is written exactly once, by the last patch in list order
//     that covers it (what "later patches overwrite earlier ones" leaves behind), or by the bilinear background where no
//     patch covers it -- race-free for any number of ranks writing into rank 0's canvas.
#include "kd_common.cuh"

namespace {

__global__ void strip_copy_kernel(const float* __restrict__ src, long cs, long rs, int C, int rows, int cols, float* __restrict__ dst) {
  const long n = (long)C * rows * cols;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int x = (int)(i % cols);
    const long t = i / cols;
    const int y = (int)(t % rows), c = (int)(t / rows);
    dst[i] = src[c * cs + (long)y * rs + x];
  }
}

// Runs after the copy kernel in stream order: every strip byte is performed before the flag becomes visible system-wide.
__global__ void peer_signal_kernel(uint32_t* flag, uint32_t value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void flag_wait_kernel(const uint32_t* flag, uint32_t value, unsigned long long timeout_ns, uint32_t* status) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned ns = 64;
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= value) return;
    __nanosleep(ns);
    if (ns < 4096) ns <<= 1;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (timeout_ns && t - t0 > timeout_ns) {  // never hang the GPU: report and let the host raise
      if (status) atomicAdd(status, 1u);
      return;
    }
  }
}

struct AxisMap {
  int off;    // position in the shifted image of output index 0 (CenterCrop offset; negative when the image is padded)
  int shift;  // torch.roll shift along this axis
};

__device__ __forceinline__ bool axis_lookup(int o, const AxisMap& m, int W, int* src, bool* filled) {
  const int p = o + m.off;
  if (p < 0 || p >= W) return false;  // CenterCrop zero padding (W < 1024)
  // :380-388 -- rows [0, shift) when shift > 0, rows [W + shift, W) when shift < 0, and the WHOLE axis when shift == 0
  // (`img[:, 0:, :] = FILL`, a reference quirk kept bit for bit)
  *filled = m.shift > 0 ? (p < m.shift) : (m.shift < 0 ? (p >= W + m.shift) : true);
  int s = (p - m.shift) % W;
  *src = s < 0 ? s + W : s;
  return true;
}

__device__ __forceinline__ float cond_value(const float* __restrict__ z, int W, int c, int y, int x, AxisMap my, AxisMap mx, float fill) {
  int sy, sx;
  bool fy, fx;
  if (!axis_lookup(y, my, W, &sy, &fy) || !axis_lookup(x, mx, W, &sx, &fx)) return 0.f;
  return (fy || fx) ? fill : z[((long)c * W + sy) * W + sx];
}

__global__ void cond_gather_kernel(const float* __restrict__ z, int W, float* __restrict__ out, int P, int extra, AxisMap my, AxisMap mx,
                                   float fill, int center_top, float nearest_scale, int patch_width) {
  const long n = (long)P * P;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int y = (int)(i / P), x = (int)(i % P);
    for (int c = 0; c < 3; ++c) out[(long)c * n + i] = cond_value(z, W, c, y, x, my, mx, fill);
    if (extra) {  // v2 (:392-395): CenterCrop(patch_width) of the cond image, nearest-upsampled to P
      const int ny = min((int)floorf(y * nearest_scale), patch_width - 1) + center_top;
      const int nx = min((int)floorf(x * nearest_scale), patch_width - 1) + center_top;
      for (int c = 0; c < 3; ++c) out[(long)(3 + c) * n + i] = cond_value(z, W, c, ny, nx, my, mx, fill);
    }
  }
}

// list index of the patch that owns canvas pixel (Y, X): the largest list index among the patches covering it, -1 if none
__device__ __forceinline__ int canvas_owner(int Y, int X, const int* __restrict__ cell, int n, int d, int P) {
  int best = -1;
  const int i_hi = min(n - 1, Y / d), j_hi = min(n - 1, X / d);
  for (int i = i_hi; i >= 0 && i * d + P > Y; --i)
    for (int j = j_hi; j >= 0 && j * d + P > X; --j) best = max(best, cell[i * n + j]);
  return best;
}

__global__ void canvas_fill_kernel(const float* __restrict__ z, int W, float* __restrict__ canvas, int Wc, const int* __restrict__ cell, int n,
                                   int d, int P, float scale) {
  const long npx = (long)Wc * Wc;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) {
    const int Y = (int)(i / Wc), X = (int)(i % Wc);
    if (canvas_owner(Y, X, cell, n, d, P) >= 0) continue;
    if (!z) {
      for (int c = 0; c < 3; ++c) canvas[(long)c * npx + i] = 0.f;
      continue;
    }
    // F.interpolate(mode='bilinear', align_corners=False): src = scale * (dst + 0.5) - 0.5 clamped at 0
    const float fy = fmaxf(scale * (Y + 0.5f) - 0.5f, 0.f), fx = fmaxf(scale * (X + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < W - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    for (int c = 0; c < 3; ++c) {
      const float* p = z + (long)c * W * W;
      canvas[(long)c * npx + i] = hy * (hx * p[(long)y0 * W + x0] + lx * p[(long)y0 * W + x1]) + ly * (hx * p[(long)y1 * W + x0] + lx * p[(long)y1 * W + x1]);
    }
  }
}

__global__ void patch_paste_kernel(const float* __restrict__ patch, float* __restrict__ canvas, int Wc, const int* __restrict__ cell, int n, int d,
                                   int P, int k, int pi, int pj) {
  const long n_in = (long)P * P, npx = (long)Wc * Wc;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    const int y = (int)(i / P), x = (int)(i % P);
    const int Y = pi * d + y, X = pj * d + x;
    if (Y >= Wc || X >= Wc || canvas_owner(Y, X, cell, n, d, P) != k) continue;
    for (int c = 0; c < 3; ++c) canvas[(long)c * npx + (long)Y * Wc + X] = patch[(long)c * n_in + i];
  }
}

unsigned grid_blocks(long n) {
  long b = (n + 255) / 256;
  const long cap = (long)kd_num_sms() * 16;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

// ------------------------------------------------------------------------------------------------ peer mailbox
extern "C" int kd_peer_alloc(size_t bytes, void** ptr) {
  KD_REQUIRE(ptr && bytes > 0, "kd_peer_alloc: bad argument");
  KD_CUDA(cudaMalloc(ptr, bytes));
  KD_CUDA(cudaMemset(*ptr, 0, bytes));
  KD_CUDA(cudaDeviceSynchronize());
  return KD_OK;
}

extern "C" int kd_peer_free(void* ptr) {
  if (ptr) KD_CUDA(cudaFree(ptr));
  return KD_OK;
}

extern "C" int kd_peer_export(const void* ptr, uint8_t* handle) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  KD_REQUIRE(ptr && handle, "kd_peer_export: bad argument");
  cudaIpcMemHandle_t h;
  KD_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle, &h, sizeof(h));
  return KD_OK;
}

extern "C" int kd_peer_open(const uint8_t* handle, void** ptr) {
  KD_REQUIRE(ptr && handle, "kd_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  KD_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return KD_OK;
}

extern "C" int kd_peer_close(void* ptr) {
  if (ptr) KD_CUDA(cudaIpcCloseMemHandle(ptr));
  return KD_OK;
}

extern "C" int kd_strip_push(const float* src, long cs, long rs, int C, int rows, int cols, float* dst, uint32_t* flag, uint32_t value,
                             kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(src && dst && flag && C > 0 && rows > 0 && cols > 0 && value > 0, "kd_strip_push: bad argument");
  strip_copy_kernel<<<grid_blocks((long)C * rows * cols), 256, 0, stream>>>(src, cs, rs, C, rows, cols, dst);
  KD_LAUNCH_CHECK();
  peer_signal_kernel<<<1, 1, 0, stream>>>(flag, value);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_flag_wait(const uint32_t* flag, uint32_t value, double timeout_s, uint32_t* status, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(flag && value > 0 && timeout_s >= 0, "kd_flag_wait: bad argument");
  flag_wait_kernel<<<1, 1, 0, stream>>>(flag, value, (unsigned long long)(timeout_s * 1e9), status);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

// ------------------------------------------------------------------------------------------------ get_cond_images / stitch
extern "C" int kd_cond_gather(const float* zoomed, int W, float* out, int channels_out, int P, int off, int shift_y, int shift_x, float fill,
                              int patch_width, int center_top, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(zoomed && out && W > 0 && P > 0 && (channels_out == 3 || channels_out == 6), "kd_cond_gather: bad argument");
  KD_REQUIRE(channels_out == 3 || (patch_width > 0 && center_top >= 0 && center_top + patch_width <= P), "kd_cond_gather: bad v2 crop");
  const AxisMap my{off, shift_y}, mx{off, shift_x};
  const float nearest_scale = channels_out == 6 ? (float)patch_width / (float)P : 0.f;
  cond_gather_kernel<<<grid_blocks((long)P * P), 256, 0, stream>>>(zoomed, W, out, P, channels_out == 6, my, mx, fill, center_top, nearest_scale,
                                                                   patch_width);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_canvas_fill(const float* zoomed, int W, float* canvas, int Wc, const int* cell_index, int n, int patch_dist, int P,
                              kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(canvas && cell_index && Wc > 0 && n > 0 && patch_dist > 0 && P > 0 && (!zoomed || W > 0), "kd_canvas_fill: bad argument");
  const float scale = zoomed ? (float)W / (float)Wc : 0.f;
  canvas_fill_kernel<<<grid_blocks((long)Wc * Wc), 256, 0, stream>>>(zoomed, W, canvas, Wc, cell_index, n, patch_dist, P, scale);
  KD_LAUNCH_CHECK();
  return KD_OK;
}

extern "C" int kd_patch_paste(const float* patch, float* canvas, int Wc, const int* cell_index, int n, int patch_dist, int P, int k, int i,
                              int j, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(patch && canvas && cell_index && Wc > 0 && n > 0 && patch_dist > 0 && P > 0 && k >= 0 && i >= 0 && j >= 0 && i < n && j < n,
             "kd_patch_paste: bad argument");
  patch_paste_kernel<<<grid_blocks((long)P * P), 256, 0, stream>>>(patch, canvas, Wc, cell_index, n, patch_dist, P, k, i, j);
  KD_LAUNCH_CHECK();
  return KD_OK;
}
