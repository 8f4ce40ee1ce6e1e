// scan.cu — device exclusive prefix sum over uint64 (three small kernels; inputs are per-block/per-stream sizes,
// at most a few million elements, so this is launch-latency bound and never on the roofline).
#include "kernels.h"

namespace b2s {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total) {
  __shared__ uint64_t warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint64_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    uint64_t winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += t;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = winc - w;  // exclusive warp offsets
    if (lane == kScanThreads / 32 - 1) *total = winc;
  }
  __syncthreads();
  uint64_t r = warp_sums[wid] + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint64_t* __restrict__ v, size_t n,
                                                                   uint64_t* __restrict__ partial) {
  __shared__ uint64_t total;
  const size_t base = (size_t)blockIdx.x * kScanTile;
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    size_t i = base + (size_t)threadIdx.x * kScanItems + k;
    if (i < n) s += v[i];
  }
  (void)block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// single CTA: exclusive scan of the partials in place, grand total out
// (base: optional running offset added to every result; grand may alias base => chained chunk scans)
__global__ void __launch_bounds__(kScanThreads) scan_partials_kernel(uint64_t* __restrict__ partial, size_t m,
                                                                     uint64_t* grand, const uint64_t* base) {
  __shared__ uint64_t total;
  uint64_t carry = base ? *base : 0;
  __syncthreads();
  for (size_t base = 0; base < m; base += kScanThreads) {
    size_t i = base + threadIdx.x;
    uint64_t x = i < m ? partial[i] : 0;
    uint64_t e = block_exclusive_scan(x, &total);
    if (i < m) partial[i] = carry + e;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint64_t* __restrict__ v, size_t n,
                                                                  const uint64_t* __restrict__ partial) {
  __shared__ uint64_t total;
  const size_t base = (size_t)blockIdx.x * kScanTile;
  uint64_t x[kScanItems];
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    size_t i = base + (size_t)threadIdx.x * kScanItems + k;
    x[k] = i < n ? v[i] : 0;
    s += x[k];
  }
  uint64_t e = block_exclusive_scan(s, &total) + partial[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    size_t i = base + (size_t)threadIdx.x * kScanItems + k;
    if (i < n) v[i] = e;
    e += x[k];
  }
}

size_t scan_ws_elems(size_t n) { return (n + kScanTile - 1) / kScanTile + 1; }

void launch_exclusive_scan_u64(uint64_t* d_v, size_t n, uint64_t* d_total, uint64_t* d_ws, cudaStream_t st,
                               uint64_t* launches, const uint64_t* d_base) {
  if (n == 0) {
    if (!d_base) cudaMemsetAsync(d_total, 0, sizeof(uint64_t), st);
    else if (d_base != d_total) cudaMemcpyAsync(d_total, d_base, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st);
    return;
  }
  const size_t m = (n + kScanTile - 1) / kScanTile;
  scan_reduce_kernel<<<(unsigned)m, kScanThreads, 0, st>>>(d_v, n, d_ws);
  scan_partials_kernel<<<1, kScanThreads, 0, st>>>(d_ws, m, d_total, d_base);
  scan_apply_kernel<<<(unsigned)m, kScanThreads, 0, st>>>(d_v, n, d_ws);
  if (launches) *launches += 3;
}

}  // namespace b2s
