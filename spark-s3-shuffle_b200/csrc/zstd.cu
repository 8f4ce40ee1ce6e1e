// zstd.cu — K6: Zstandard frame decoding on the device (decoder core: zstd_core.h, shared with the host unit test).
//
// Replaces com.github.luben.zstd.ZstdInputStreamNoFinalizer [U] (zstd-jni -> libzstd ZSTD_decompressStream) under
// serializerManager.wrapStream at storage/S3ShuffleReader.scala:107-109 for spark.io.compression.codec=zstd.
//
// One WARP per compressed stream (= shuffle block; a stream is one or more frames, and blocks inside a frame depend on
// each other through the window and the repeat-offset / table history, so a frame is a serial unit).  All 32 lanes run
// the decoder core in lock step on the same data — the serial parts (headers, table builds, the FSE sequence decoder)
// are computed redundantly, which costs a warp what it would cost one lane — so that
//   * the FSE / Huffman tables live in SHARED memory (11 KB per warp; the thread-per-stream first version kept them in
//     global memory and spent ~1500 cycles per output byte on table look-ups),
//   * the four Huffman streams of a literals section are decoded by four lanes,
//   * literal and match copies of every sequence are spread over the 32 lanes.
// Parallelism across streams comes from the batch (16,000 shuffle blocks in BASELINE config 2; ~3,000 in flight).
#define B2S_ZSTD_WARP 1
#include "kernels.h"
#include "zstd_core.h"

namespace b2s {

__device__ __forceinline__ int32_t zstd_status(int64_t r) {
  return r == zstd::kErrDstTooSmall ? B2S_E_DST_TOO_SMALL : r == zstd::kErrUnsupported ? B2S_E_UNSUPPORTED : B2S_E_CORRUPT;
}

constexpr int kZWarps = 4;  // warps (= streams in flight) per CTA: 4 x sizeof(Workspace) of dynamic shared memory
constexpr uint32_t kZLitBytes = zstd::kBlockMax + 64;

template <bool SIZE_ONLY>
__global__ void __launch_bounds__(kZWarps * 32) zstd_stream_kernel(const uint8_t* __restrict__ src_base,
                                                                   const uint64_t* __restrict__ src_off,
                                                                   const uint64_t* __restrict__ src_len, uint32_t n,
                                                                   uint8_t* __restrict__ lit_pool, uint32_t n_workers,
                                                                   uint64_t* __restrict__ olen,
                                                                   uint8_t* __restrict__ dst_base,
                                                                   const uint64_t* __restrict__ dst_off,
                                                                   uint64_t dst_cap, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char zsmem[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const uint32_t worker = blockIdx.x * kZWarps + wic;
  if (worker >= n_workers) return;
  zstd::Workspace* w = reinterpret_cast<zstd::Workspace*>(zsmem) + wic;
  w->lit = lit_pool + (size_t)worker * kZLitBytes;
  __syncwarp();
  for (uint32_t i = worker; i < n; i += n_workers) {
    const int32_t st0 = status[i];
    if (st0 != 0) {  // failed checksum verification (or an earlier phase): nothing of this block is decoded
      if (SIZE_ONLY && lane == 0) olen[i] = 0;
      continue;
    }
    const uint8_t* s = src_base + src_off[i];
    if (SIZE_ONLY) {
      const int64_t r = zstd::decode_stream(w, s, src_len[i], nullptr, 0, true);
      if (lane == 0) {
        if (r < 0) {
          status[i] = zstd_status(r);
          olen[i] = 0;
        } else {
          olen[i] = (uint64_t)r;
        }
      }
    } else {
      const uint64_t want = olen[i];
      if (dst_off[i] + want > dst_cap) {
        if (lane == 0) status[i] = B2S_E_DST_TOO_SMALL;
        continue;
      }
      const int64_t r = zstd::decode_stream(w, s, src_len[i], dst_base + dst_off[i], want, false);
      if (lane == 0) {
        if (r < 0) status[i] = zstd_status(r);
        else if ((uint64_t)r != want) status[i] = B2S_E_CORRUPT;
      }
    }
    __syncwarp();
  }
}

uint32_t zstd_workers(uint32_t n) { return n < 4096u ? n : 4096u; }  // warps; ~3000 fit on the machine at once
size_t zstd_ws_bytes(uint32_t n) { return (size_t)zstd_workers(n) * kZLitBytes + 256; }

template <bool SIZE_ONLY>
static void launch_zstd_t(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len, uint32_t n,
                          uint8_t* d_ws, uint64_t* d_olen, uint8_t* dst_base, const uint64_t* d_dst_off,
                          uint64_t dst_cap, int32_t* d_status, cudaStream_t st) {
  const uint32_t workers = zstd_workers(n);
  const size_t smem = (size_t)kZWarps * sizeof(zstd::Workspace);
  cudaFuncSetAttribute(zstd_stream_kernel<SIZE_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device and cheap: set on every launch
  zstd_stream_kernel<SIZE_ONLY><<<(workers + kZWarps - 1) / kZWarps, kZWarps * 32, smem, st>>>(
      src_base, d_src_off, d_src_len, n, d_ws, workers, d_olen, dst_base, d_dst_off, dst_cap, d_status);
}

void launch_zstd_sizes(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len, uint32_t n,
                       uint8_t* d_ws, uint64_t* d_olen, int32_t* d_status, cudaStream_t st, uint64_t* launches) {
  if (!n) return;
  launch_zstd_t<true>(src_base, d_src_off, d_src_len, n, d_ws, d_olen, nullptr, nullptr, 0, d_status, st);
  *launches += 1;
}

void launch_zstd_decode(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len, uint32_t n,
                        uint8_t* d_ws, uint64_t* d_olen, uint8_t* dst_base, const uint64_t* d_dst_off, uint64_t dst_cap,
                        int32_t* d_status, cudaStream_t st, uint64_t* launches) {
  if (!n) return;
  launch_zstd_t<false>(src_base, d_src_off, d_src_len, n, d_ws, d_olen, dst_base, d_dst_off, dst_cap, d_status, st);
  *launches += 1;
}

}  // namespace b2s
