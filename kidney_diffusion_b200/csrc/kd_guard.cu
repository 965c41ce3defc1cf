// Overflow guard of the fp16 activation path.  Every conversion to the 16-bit activation type saturates at +-65504
// (cvt.rn.satfinite, kd_common.cuh) so an outlier can never become inf / NaN -- but silent clipping would still corrupt a
// sample.  kd_count_saturated counts the elements of a stored activation tensor that sit AT the saturation value (or are not
// finite); the host enables it per sample() call (Imagen.check_saturation) and reports the total.
#include "kd_common.cuh"

namespace {
__global__ void count_saturated_kernel(const h16* __restrict__ x, long n8, long n, unsigned long long* __restrict__ counter) {
  unsigned int local = 0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int4 raw = ld_stream(x + i * 8);
    const uint32_t w[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // |v| >= 65504  <=>  (bits & 0x7fff) >= 0x7bff (max finite; 0x7c00.. are inf / NaN)
      local += ((w[k] & 0x7fffu) >= 0x7bffu) + (((w[k] >> 16) & 0x7fffu) >= 0x7bffu);
    }
  }
  if (blockIdx.x == 0)
    for (long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) local += ((__half_as_ushort(x[i]) & 0x7fffu) >= 0x7bffu);
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(counter, (unsigned long long)local);
}
}  // namespace

extern "C" int kd_count_saturated(const void* x, long n, unsigned long long* counter, kd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  KD_REQUIRE(x && counter && n > 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "kd_count_saturated: bad argument");
  const long n8 = n / 8;
  long blocks = (n8 + 255) / 256;
  const long cap = (long)kd_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  count_saturated_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const h16*>(x), n8, n, counter);
  KD_LAUNCH_CHECK();
  return KD_OK;
}
