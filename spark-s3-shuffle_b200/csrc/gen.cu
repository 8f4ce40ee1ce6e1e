// gen.cu — synthetic TeraGen-shaped record generator for bench.py (NOT on the product path).
// Byte-identical to orc_gen_terasort() in oracle/b2s_oracle.c (counter-based splitmix64), so the CPU baseline
// sample and the GPU run see the same data.  One thread per 104-byte record.
#include "kernels.h"

namespace b2s {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void gen_terasort_kernel(uint8_t* __restrict__ dst, uint64_t first, uint64_t n, uint64_t seed) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t g = first + i;
  uint8_t r[104];
  const uint64_t a = mix64(seed ^ mix64(g * 4 + 0)), b = mix64(seed ^ mix64(g * 4 + 1)),
                 c = mix64(seed ^ mix64(g * 4 + 2));
  r[0] = 0x01;
  r[1] = 0x0B;
#pragma unroll
  for (int k = 0; k < 8; k++) r[2 + k] = (uint8_t)(a >> (8 * k));
  r[10] = (uint8_t)b;
  r[11] = (uint8_t)(b >> 8);
  r[12] = 0x01;
  r[13] = 0x5B;
  r[14] = 0x00;
  r[15] = 0x11;
#pragma unroll
  for (int k = 0; k < 32; k++) {
    const uint32_t nib = (uint32_t)(g >> (4 * (31 - (k < 16 ? 16 : k)))) & 15u;
    r[16 + k] = (k < 16) ? (uint8_t)'0' : (uint8_t)(nib < 10 ? '0' + nib : 'A' + nib - 10);
  }
  r[48] = 0x88;
  r[49] = 0x99;
  r[50] = 0xAA;
  r[51] = 0xBB;
#pragma unroll
  for (int k = 0; k < 12; k++) {
    const uint32_t nib = (uint32_t)(c >> (4 * k)) & 15u;
    const uint8_t ch = (uint8_t)(nib < 10 ? '0' + nib : 'A' + nib - 10);
    r[52 + 4 * k] = r[53 + 4 * k] = r[54 + 4 * k] = r[55 + 4 * k] = ch;
  }
  r[100] = 0xCC;
  r[101] = 0xDD;
  r[102] = 0xEE;
  r[103] = 0xFF;
  // 104 = 26 words; record start is 8-byte aligned when dst is
  uint32_t* o = reinterpret_cast<uint32_t*>(dst + i * 104);
#pragma unroll
  for (int k = 0; k < 26; k++)
    o[k] = (uint32_t)r[4 * k] | ((uint32_t)r[4 * k + 1] << 8) | ((uint32_t)r[4 * k + 2] << 16) |
           ((uint32_t)r[4 * k + 3] << 24);
}

void launch_gen_terasort(uint8_t* d_dst, uint64_t first_record, uint64_t n_records, uint64_t seed, cudaStream_t st) {
  if (!n_records) return;
  gen_terasort_kernel<<<(unsigned)((n_records + 255) / 256), 256, 0, st>>>(d_dst, first_record, n_records, seed);
}

}  // namespace b2s
