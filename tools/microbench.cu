// microbench.cu — dev tool (not product): latency/throughput of warp primitives used by the LZ kernels on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_match_lat(uint32_t* out, int iters, uint32_t seed) {
  uint32_t h = (threadIdx.x * 2654435761u + seed) >> 20;
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < iters; i++) { unsigned m = __match_any_sync(0xffffffffu, h); h = (h + m) & 4095; acc += m; }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = (uint32_t)((t1 - t0) / iters); out[1] = acc; }
}
__global__ void k_ballot_lat(uint32_t* out, int iters, uint32_t seed) {
  uint32_t h = threadIdx.x + seed; long long t0 = clock64(); uint32_t acc = 0;
  for (int i = 0; i < iters; i++) { unsigned m = __ballot_sync(0xffffffffu, h & 1); h = h * 3 + m; acc += m; }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[2] = (uint32_t)((t1 - t0) / iters); out[3] = acc; }
}
__global__ void k_shfl_lat(uint32_t* out, int iters, uint32_t seed) {
  uint32_t h = threadIdx.x + seed; long long t0 = clock64();
  for (int i = 0; i < iters; i++) { h = __shfl_sync(0xffffffffu, h, (h + 1) & 31) + 1; }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[4] = (uint32_t)((t1 - t0) / iters); out[5] = h; }
}
// throughput: many warps issuing independent match_any
__global__ void k_match_tput(uint32_t* out, int iters, uint32_t seed) {
  uint32_t h0 = (threadIdx.x * 2654435761u + seed) >> 20, h1 = h0 ^ 77, h2 = h0 ^ 99, h3 = h0 + 5;
  uint32_t acc = 0;
  for (int i = 0; i < iters; i++) {
    acc += __match_any_sync(0xffffffffu, h0); acc += __match_any_sync(0xffffffffu, h1);
    acc += __match_any_sync(0xffffffffu, h2); acc += __match_any_sync(0xffffffffu, h3);
    h0 += 3; h1 += 5; h2 += 7; h3 += 11;
  }
  if (acc == 0x12345) out[8] = acc;
}
__global__ void k_sts_winner(uint32_t* out) {
  __shared__ uint32_t s[4];
  s[threadIdx.x & 3] = 0xffffffffu; __syncwarp();
  s[0] = threadIdx.x; __syncwarp();            // all 32 lanes same address
  if (threadIdx.x == 0) out[6] = s[0];
  __syncwarp();
  if (threadIdx.x >= 5 && threadIdx.x < 21 && (threadIdx.x & 1)) s[1] = threadIdx.x; __syncwarp();
  if (threadIdx.x == 0) out[7] = s[1];
}
int main() {
  uint32_t* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  k_match_lat<<<1, 32>>>(d, 10000, 1); k_ballot_lat<<<1, 32>>>(d, 10000, 1); k_shfl_lat<<<1, 32>>>(d, 10000, 1); k_sts_winner<<<1, 32>>>(d);
  uint32_t h[16]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("match_any latency %u cyc | ballot dep-chain %u cyc | shfl dep-chain %u cyc | STS same-addr winner lane (32 lanes) %u, (odd lanes 5..19) %u\n", h[0], h[2], h[4], h[6], h[7]);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int iters = 20000; k_match_tput<<<148 * 8, 256>>>(d, 10, 3); cudaEventRecord(a); k_match_tput<<<148 * 8, 256>>>(d, iters, 3); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double insts = 148.0 * 8 * 8 * iters * 4; // warp-level match instructions
  printf("match_any throughput: %.2f warp-inst/cycle/SM at 1.965 GHz (%.3f ms)\n", insts / (ms * 1e-3) / 1.965e9 / 148, ms);
  return 0;
}
