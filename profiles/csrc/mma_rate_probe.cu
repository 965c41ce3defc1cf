// EXPERIMENT (not part of libkidney_b200.so): issue-rate probe for tcgen05.mma kind::f16 with operands resident in shared memory.
// One CTA per SM issues `iters` x 72 MMAs (the K loop of a 3x3 conv tile over 2 channel chunks) for a given shape and prints the
// cycles per MMA:  variant 0: cta_group::1, M 128 x N 256 (Cout-on-M formulation for Cout = 128: A = filter block, B = 256 pixels of a
// 34 x 10 halo through row-offset descriptors);  variant 1: cta_group::1, M 128 x N 128 (pixels on M, A through halo descriptors);
// variants 2, 3, 4: M 128 x N 16 / 64 / 32 (pixels on M).  Operand contents are irrelevant (uninitialised smem): only the issue / completion rate is measured.
#include <cuda.h>
#include <cstdio>

#include "kd_common.cuh"
#include "kd_tc.cuh"

namespace {
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int variant, int iters, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_ptr), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  const int N = variant == 0 ? 256 : (variant == 1 ? 128 : (variant == 2 ? 16 : (variant == 3 ? 64 : 32)));
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t halo = base;                 // 34 x 10 x 128 B = 43 520 B
  const uint32_t filt = base + 44 * 1024;     // 9 taps x (128 rows x 128 B) = 147 456 B
  if (warp == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap % 3;
            const uint64_t hdesc = make_sw128_desc(halo + (uint32_t)(ky * 10 + kx) * 128u, 10 * 128);
            const uint64_t fdesc = make_sw128_desc(filt + (uint32_t)tap * 16384u);
            const uint64_t a_desc = variant == 0 ? fdesc : hdesc, b_desc = variant == 0 ? hdesc : fdesc;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + (uint32_t)(it & 1) * 256u, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (ch | tap | k) != 0 ? 1u : 0u);
          }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&done_bar));
    __syncwarp();
    mbar_wait(smem_u32(&done_bar), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
}  // namespace

int main() {
  const int smem = 44 * 1024 + 9 * 16384 + 2048;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  for (int variant = 0; variant < 5; ++variant) {
    const int iters = 200;
    mma_rate_kernel<<<148, 128, smem>>>(variant, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const int N = variant == 0 ? 256 : (variant == 1 ? 128 : (variant == 2 ? 16 : (variant == 3 ? 64 : 32)));
    const double cyc = (double)mx / (iters * 72.0);
    printf("variant %d (M 128 x N %d, cta_group::1): %s, %.1f cycles per MMA (math at full rate: %.0f; operand bytes %d -> %.0f B/clk)\n", variant, N,
           cudaGetErrorString(e), cyc, N / 2.0, (128 + N) * 32, (128 + N) * 32 / cyc);
  }
  return 0;
}
