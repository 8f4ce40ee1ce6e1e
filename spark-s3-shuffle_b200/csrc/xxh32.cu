// xxh32.cu — K2: XXH32 per LZ4Block codec block.
//
// Replaces lz4-java's StreamingXXHash32(seed 0x9747b28c).asChecksum() [U], which LZ4BlockOutputStream stores (masked
// to 28 bits) in every block header and LZ4BlockInputStream re-computes after decoding.
//
// XXH32 is four serial accumulator chains per block (the round rotl(acc + x*P2, 13)*P1 is not associative), so the
// parallelism is across blocks: one thread per block, four independent chains per thread for ILP, 16 bytes per step.
// A warp touches 32 different cache lines per step, each line is then reused for 8 steps out of L1, so DRAM traffic
// stays at 1 byte per input byte.
#include "kernels.h"

namespace b2s {

constexpr uint32_t XP1 = 2654435761u, XP2 = 2246822519u, XP3 = 3266489917u, XP4 = 668265263u, XP5 = 374761393u;

__device__ __forceinline__ uint32_t xxh_round(uint32_t acc, uint32_t x) { return rotl32(acc + x * XP2, 13) * XP1; }

// 4 words at word offset K (0..3) and bit shift sh (0, 8, 16, 24) inside the 8 words lo|hi
template <int K>
__device__ __forceinline__ void pick4(const uint4& lo, const uint4& hi, unsigned sh, uint32_t& a, uint32_t& b,
                                      uint32_t& c, uint32_t& d) {
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  a = __funnelshift_r(w[K], w[K + 1], sh);
  b = __funnelshift_r(w[K + 1], w[K + 2], sh);
  c = __funnelshift_r(w[K + 2], w[K + 3], sh);
  d = __funnelshift_r(w[K + 3], w[K + 4], sh);
}

// the 16-byte stripes of [p, p + 16*stripes) for ANY alignment of p: one aligned 128-bit load per stripe (the previous
// chunk is carried in registers) + funnel shifts.  One thread per block means every load instruction is 32 wavefronts on
// the LSU pipe, so the 4 x 32-bit loads per stripe of the first version were the bottleneck (1.6 TB/s), not HBM.
// Reads only aligned 16-byte chunks that contain bytes of the range.
template <int K>
__device__ __forceinline__ void xxh_stripes(const uint4* __restrict__ q, unsigned sh, uint32_t stripes, bool tail_chunk,
                                            uint32_t& v1, uint32_t& v2, uint32_t& v3, uint32_t& v4) {
  uint4 lo = __ldg(q);
#pragma unroll 4
  for (uint32_t s = 0; s < stripes; s++) {
    // the chunk after the last stripe holds range bytes only when the range is not 16-byte aligned (K or sh != 0)
    const uint4 hi = (s + 1 < stripes || tail_chunk) ? __ldg(q + s + 1) : make_uint4(0, 0, 0, 0);
    uint32_t a, b, c, d;
    pick4<K>(lo, hi, sh, a, b, c, d);
    v1 = xxh_round(v1, a);
    v2 = xxh_round(v2, b);
    v3 = xxh_round(v3, c);
    v4 = xxh_round(v4, d);
    lo = hi;
  }
}

// p may have any alignment; reads only aligned chunks/words that contain bytes of [p, p+n)
__device__ uint32_t xxh32_device(const uint8_t* p, uint32_t n, uint32_t seed) {
  uint32_t h;
  uint32_t i = 0;  // bytes consumed
  if (n >= 16) {
    uint32_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
    const uint32_t stripes = n >> 4;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned sh = (a & 3u) * 8u;
    const int k = (int)((a >> 2) & 3u);
    const uint4* q = reinterpret_cast<const uint4*>(a & ~uintptr_t(15));
    // bytes of the range inside the chunk that follows the last full stripe's chunk?
    const bool tail_chunk = (a & 15u) != 0;
    switch (k) {
      case 0: xxh_stripes<0>(q, sh, stripes, tail_chunk, v1, v2, v3, v4); break;
      case 1: xxh_stripes<1>(q, sh, stripes, tail_chunk, v1, v2, v3, v4); break;
      case 2: xxh_stripes<2>(q, sh, stripes, tail_chunk, v1, v2, v3, v4); break;
      default: xxh_stripes<3>(q, sh, stripes, tail_chunk, v1, v2, v3, v4); break;
    }
    i = stripes << 4;
    h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
  } else {
    h = seed + XP5;
  }
  h += n;
  for (; i + 4 <= n; i += 4) h = rotl32(h + ld32u(p + i) * XP3, 17) * XP4;
  for (; i < n; i++) h = rotl32(h + p[i] * XP5, 11) * XP1;
  h ^= h >> 15;
  h *= XP2;
  h ^= h >> 13;
  h *= XP3;
  h ^= h >> 16;
  return h;
}

// largest i with blk_base[i] <= b  (blk_base has n_streams+1 entries, non-decreasing)
__device__ __forceinline__ uint32_t find_stream(const uint32_t* __restrict__ blk_base, uint32_t n_streams, uint32_t b) {
  uint32_t lo = 0, hi = n_streams;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (blk_base[mid] <= b) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128) xxh32_encode_kernel(const uint8_t* __restrict__ src_base,
                                                           const uint64_t* __restrict__ src_off,
                                                           const uint64_t* __restrict__ src_len,
                                                           const uint32_t* __restrict__ blk_base, uint32_t n_streams,
                                                           uint32_t n_blocks, uint32_t block_size, uint32_t seed,
                                                           uint32_t* __restrict__ hash) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  uint32_t i = find_stream(blk_base, n_streams, b);
  uint64_t o = (uint64_t)(b - blk_base[i]) * block_size;
  uint64_t rem = src_len[i] - o;
  uint32_t n = (uint32_t)(rem < block_size ? rem : block_size);
  hash[b] = xxh32_device(src_base + src_off[i] + o, n, seed);
}

__global__ void __launch_bounds__(128) xxh32_verify_kernel(const BlockDesc* __restrict__ desc, uint32_t n_blocks,
                                                           const uint8_t* __restrict__ dst_base, uint32_t seed,
                                                           uint32_t mask, int32_t* __restrict__ status) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const BlockDesc d = desc[b];
  const uint32_t stream = d.stream & 0x7fffffffu;
  if (status[stream] != 0) return;  // already failed: decoded bytes are not meaningful
  uint32_t h = xxh32_device(dst_base + d.dst, d.olen, seed) & mask;
  if (h != d.check) set_status(status, stream, B2S_E_CORRUPT);
}

void launch_xxh32_encode(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len,
                         const uint32_t* d_blk_base, uint32_t n_streams, uint32_t n_blocks, uint32_t block_size,
                         uint32_t seed, uint32_t* d_hash, cudaStream_t st, uint64_t* launches) {
  if (!n_blocks) return;
  xxh32_encode_kernel<<<(n_blocks + 127) / 128, 128, 0, st>>>(src_base, d_src_off, d_src_len, d_blk_base, n_streams,
                                                             n_blocks, block_size, seed, d_hash);
  *launches += 1;
}

void launch_xxh32_verify(const BlockDesc* d_desc, uint32_t n_blocks, const uint8_t* dst_base, uint32_t seed,
                         uint32_t mask, int32_t* d_status, cudaStream_t st, uint64_t* launches) {
  if (!n_blocks) return;
  xxh32_verify_kernel<<<(n_blocks + 127) / 128, 128, 0, st>>>(d_desc, n_blocks, dst_base, seed, mask, d_status);
  *launches += 1;
}

}  // namespace b2s
