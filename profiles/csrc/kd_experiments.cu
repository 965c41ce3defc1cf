// Hardware-semantics probe (test-only entry point, not used by the product path):
// can one 3x3-conv A operand be served from a single halo tile in shared memory, i.e. does tcgen05.mma accept a K-major
// SWIZZLE_128B descriptor whose start address is offset by whole 128-byte rows (tap shift) with SBO = 16 rows?
// Variant 0: base_offset = 0; variant 1: base_offset = (start >> 7) & 7 (PTX matrix-descriptor "base offset").
#include <cuda.h>

#include "kd_common.cuh"

namespace {
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void bar_expect(uint32_t bar, uint32_t b) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

__global__ void __launch_bounds__(128, 1) halo_probe_kernel(const __grid_constant__ CUtensorMap map_x,
                                                            const __grid_constant__ CUtensorMap map_w, float* out, int h0, int w0,
                                                            int variant) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (s_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - s_u32(raw));
  constexpr int A_BYTES = 18 * 16 * 128, B_BYTES = 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + A_BYTES + B_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    bar_init(s_u32(&bars[0]), 1);
    bar_init(s_u32(&bars[1]), 1);
    bar_init(s_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_ptr)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_ptr;
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    bar_expect(s_u32(&bars[0]), (variant == 2) ? 18 * 10 * 128 : A_BYTES);
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(base),
        "l"(&map_x), "r"(s_u32(&bars[0])), "r"(0), "r"(w0 - 1), "r"(h0 - 1), "r"(0), "r"(0)
        : "memory");
    bar_wait(s_u32(&bars[0]), 0);
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      bar_expect(s_u32(&bars[1]), B_BYTES);
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       base + A_BYTES),
                   "l"(&map_w), "r"(s_u32(&bars[1])), "r"(tap * 64), "r"(0)
                   : "memory");
      bar_wait(s_u32(&bars[1]), tap & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int hw = (variant == 2) ? 10 : 16;  // halo tile width in pixels
      const uint32_t a_addr = base + (uint32_t)(ky * hw + kx) * 128u;
      uint64_t a_desc = 0;
      a_desc |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
      a_desc |= (uint64_t)1 << 16;
      a_desc |= (uint64_t)((hw * 128) >> 4) << 32;  // SBO: next 8-pixel group = next image row of the halo tile
      a_desc |= (uint64_t)1 << 46;
      if (variant == 1) a_desc |= (uint64_t)((a_addr >> 7) & 7) << 49;
      a_desc |= (uint64_t)2 << 61;
      uint64_t b_desc = 0;
      b_desc |= (uint64_t)(((base + A_BYTES) & 0x3FFFF) >> 4);
      b_desc |= (uint64_t)1 << 16;
      b_desc |= (uint64_t)(1024 >> 4) << 32;
      b_desc |= (uint64_t)1 << 46;
      b_desc |= (uint64_t)2 << 61;
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = (tap | k) != 0;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(a_desc + (uint64_t)(2 * k)), "l"(b_desc + (uint64_t)(2 * k)), "r"(IDESC), "r"(acc)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(&bars[2])) : "memory");
      bar_wait(s_u32(&bars[2]), tap & 1);
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int chunk = 0; chunk < 4; ++chunk) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(chunk * 32))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + chunk * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
}  // namespace

// x: fp16 NHWC [1,H,W,64]; w: fp16 [128, 9*64]; out: fp32 [128 pixels (16 rows x 8 cols from (h0,w0)), 128]
extern "C" int kd_exp_halo_probe(const void* x, int H, int W, const void* w, float* out, int h0, int w0, int variant,
                                 kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  KD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
  PFN_enc enc = reinterpret_cast<PFN_enc>(ptr);
  CUtensorMap mx, mw;
  cuuint64_t dx[5] = {64, (cuuint64_t)W, (cuuint64_t)H, 1, 1};
  cuuint64_t sx[4] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128, (cuuint64_t)H * W * 128};
  cuuint32_t bx[5] = {64, (cuuint32_t)(variant == 2 ? 10 : 16), 18, 1, 1}, es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(x), dx, sx, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) KD_FAIL(KD_ERR_CUDA, "halo probe: x map encode failed %d", (int)r);
  cuuint64_t dw[2] = {9 * 64, 128};
  cuuint64_t sw[1] = {9 * 64 * 2};
  cuuint32_t bw[2] = {64, 128};
  r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dw, sw, bw, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) KD_FAIL(KD_ERR_CUDA, "halo probe: w map encode failed %d", (int)r);
  const int smem = 18 * 16 * 128 + 128 * 128 + 1024 + 64;
  KD_CUDA(cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  halo_probe_kernel<<<1, 128, smem, stream>>>(mx, mw, out, h0, w0, variant);
  KD_LAUNCH_CHECK();
  return KD_OK;
}
