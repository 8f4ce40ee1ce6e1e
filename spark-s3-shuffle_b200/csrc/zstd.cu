// zstd.cu — K6: Zstandard frame decoding on the device (decoder core: zstd_core.h + zstd_par.h, shared with the host
// unit test tests/test_zstd_core.py).
//
// Replaces com.github.luben.zstd.ZstdInputStreamNoFinalizer [U] (zstd-jni -> libzstd ZSTD_decompressStream) under
// serializerManager.wrapStream at storage/S3ShuffleReader.scala:107-109 for spark.io.compression.codec=zstd.
//
// Four kernels (zstd_par.h explains the split):
//   zstd_walk_kernel<FILL>     thread per stream: block headers only -> BlockInfo records with table provenance
//   zstd_entropy_kernel<SIZE>  WARP per Zstandard block, every block of every stream at once: Huffman literals (four
//                              streams on four lanes) and the FSE sequence decoder, tables in SHARED memory (6.6 KB per
//                              warp); all 32 lanes run the serial parts in lock step — redundantly, which costs a warp
//                              what it costs one lane — and store literals / sequence triples for the next kernel
//   zstd_sum_kernel            thread per stream: decoded size = sum of its blocks' sizes (this ends the size pass)
//   zstd_execute_kernel        WARP per stream, blocks in order: repeat offsets, literal and match copies, 32 sequences
//                              at a time (lz_batch.cuh)
// The first version decoded a whole stream in one warp (parallelism = number of shuffle blocks only): 12 GB/s on 4,096
// streams of 64 KiB but 0.13 GB/s on 16 streams of 16 MiB (profiles/r1z_zstd_sweep.json).
#define B2S_ZSTD_WARP 1
#define B2S_ZSTD_UNION_TABLES 1
#include "kernels.h"
#include "lz_batch.cuh"
#include "zstd_par.h"

namespace b2s {

__device__ __forceinline__ int32_t zstd_status(int64_t r) {
  return r == zstd::kErrDstTooSmall ? B2S_E_DST_TOO_SMALL : r == zstd::kErrUnsupported ? B2S_E_UNSUPPORTED : B2S_E_CORRUPT;
}

constexpr int kZWarps = 4;  // warps per CTA: 4 x sizeof(Workspace) (6.6 KB) of dynamic shared memory

// cnt / base: three arrays of n u64 each — blocks, sequences, literal-workspace bytes per stream (counts / exclusive scans)
template <bool FILL>
__global__ void __launch_bounds__(128) zstd_walk_kernel(const uint8_t* __restrict__ src_base,
                                                        const uint64_t* __restrict__ src_off,
                                                        const uint64_t* __restrict__ src_len, uint32_t n,
                                                        uint64_t* __restrict__ cnt, const uint64_t* __restrict__ base,
                                                        zstd::BlockInfo* __restrict__ blocks,
                                                        int32_t* __restrict__ status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  zstd::StreamTotals t{0, 0, 0};
  if (status[i] == 0) {  // else: failed checksum verification — nothing of this stream is decoded
    int rc;
    if (FILL) {
      if (cnt[i] == 0) return;
      rc = zstd::walk_stream(src_base + src_off[i], src_len[i], blocks + base[i], (uint32_t)base[i], i, src_off[i],
                             base[2 * (size_t)n + i], base[(size_t)n + i], &t);
    } else {
      rc = zstd::walk_stream(src_base + src_off[i], src_len[i], nullptr, 0, i, 0, 0, 0, &t);
    }
    if (rc < 0) {
      status[i] = zstd_status(rc);
      t = zstd::StreamTotals{0, 0, 0};
    }
  }
  if (!FILL) {
    cnt[i] = t.nblk;
    cnt[(size_t)n + i] = t.nseq;
    cnt[2 * (size_t)n + i] = t.lit;
  }
}

template <bool SIZE_ONLY>
__global__ void __launch_bounds__(kZWarps * 32) zstd_entropy_kernel(const uint8_t* __restrict__ src_base,
                                                                    zstd::BlockInfo* __restrict__ blocks, uint32_t nb,
                                                                    uint32_t n_workers, uint8_t* __restrict__ lit_ws,
                                                                    uint32_t* __restrict__ sq_ll,
                                                                    uint32_t* __restrict__ sq_ml,
                                                                    uint32_t* __restrict__ sq_ofv,
                                                                    int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char zsmem[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const uint32_t worker = blockIdx.x * kZWarps + wic;
  if (worker >= n_workers) return;
  zstd::Workspace* w = reinterpret_cast<zstd::Workspace*>(zsmem) + wic;
  for (uint32_t b = worker; b < nb; b += n_workers) {
    if (blocks[b].type != 2) continue;  // raw / RLE blocks: size known from the header
    const uint32_t stream = blocks[b].stream;
    if (status[stream] != 0) continue;  // another block of the stream already failed
    const int64_t r = zstd::entropy_block(w, blocks, b, src_base, lit_ws, sq_ll, sq_ml, sq_ofv, SIZE_ONLY);
    if (lane == 0) {
      if (r < 0) {
        status[stream] = zstd_status(r);
        blocks[b].out_size = 0;
      } else {
        blocks[b].out_size = (uint32_t)r;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) zstd_sum_kernel(const zstd::BlockInfo* __restrict__ blocks,
                                                       const uint64_t* __restrict__ cnt,
                                                       const uint64_t* __restrict__ base, uint32_t n,
                                                       uint64_t* __restrict__ olen, int32_t* __restrict__ status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t sum = 0;
  if (status[i] == 0) {
    const int64_t r = zstd::stream_size(blocks + base[i], cnt[i]);
    if (r < 0) status[i] = zstd_status(r);  // a frame's blocks do not add up to its Frame_Content_Size
    else sum = (uint64_t)r;
  }
  olen[i] = sum;
}

__global__ void __launch_bounds__(128) zstd_execute_kernel(const uint8_t* __restrict__ src_base,
                                                           const zstd::BlockInfo* __restrict__ blocks,
                                                           const uint64_t* __restrict__ cnt,
                                                           const uint64_t* __restrict__ base, uint32_t n,
                                                           uint32_t n_workers, const uint8_t* __restrict__ lit_ws,
                                                           const uint32_t* __restrict__ sq_ll,
                                                           const uint32_t* __restrict__ sq_ml,
                                                           const uint32_t* __restrict__ sq_ofv,
                                                           const uint64_t* __restrict__ olen, uint8_t* dst_base,
                                                           const uint64_t* __restrict__ dst_off, uint64_t dst_cap,
                                                           int32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const uint32_t worker = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (worker >= n_workers) return;
  for (uint32_t i = worker; i < n; i += n_workers) {
    if (status[i] != 0) continue;
    const uint64_t want = olen[i];
    if (dst_off[i] + want > dst_cap) {
      if (lane == 0) status[i] = B2S_E_DST_TOO_SMALL;
      continue;
    }
    const int64_t r = zstd::execute_stream(blocks + base[i], (uint32_t)cnt[i], src_base, lit_ws, sq_ll, sq_ml, sq_ofv,
                                           dst_base + dst_off[i], want);
    if (lane == 0) {
      if (r < 0) status[i] = zstd_status(r);
      else if ((uint64_t)r != want) status[i] = B2S_E_CORRUPT;
    }
    __syncwarp();
  }
}

size_t zstd_block_info_bytes() { return sizeof(zstd::BlockInfo); }
// literals workspace + three u32 arrays of nseq entries
size_t zstd_ws_bytes(uint64_t lit_bytes, uint64_t nseq) { return (size_t)lit_bytes + 256 + ((size_t)nseq + 64) * 12; }

void launch_zstd_count(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len, uint32_t n,
                       uint64_t* d_cnt, int32_t* d_status, cudaStream_t st, uint64_t* launches) {
  if (!n) return;
  zstd_walk_kernel<false><<<(n + 127) / 128, 128, 0, st>>>(src_base, d_src_off, d_src_len, n, d_cnt, nullptr, nullptr,
                                                           d_status);
  *launches += 1;
}
void launch_zstd_fill(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len, uint32_t n,
                      uint64_t* d_cnt, const uint64_t* d_base, void* d_blocks, int32_t* d_status, cudaStream_t st,
                      uint64_t* launches) {
  if (!n) return;
  zstd_walk_kernel<true><<<(n + 127) / 128, 128, 0, st>>>(src_base, d_src_off, d_src_len, n, d_cnt, d_base,
                                                          (zstd::BlockInfo*)d_blocks, d_status);
  *launches += 1;
}
void launch_zstd_entropy(bool size_only, const uint8_t* src_base, void* d_blocks, uint64_t nb, uint8_t* d_ws,
                         uint64_t lit_bytes, uint64_t nseq, int32_t* d_status, cudaStream_t st, uint64_t* launches) {
  if (!nb) return;
  const uint32_t workers = (uint32_t)(nb < 8192 ? nb : 8192);  // warps; ~3000 are resident at once
  const size_t smem = (size_t)kZWarps * sizeof(zstd::Workspace);
  uint32_t* sq = size_only ? nullptr : reinterpret_cast<uint32_t*>(d_ws + ((lit_bytes + 255) & ~(uint64_t)255));
  const size_t stride = (size_t)nseq + 64;
  const dim3 grid((workers + kZWarps - 1) / kZWarps);
  if (size_only) {
    cudaFuncSetAttribute(zstd_entropy_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device and cheap: set on every launch
    zstd_entropy_kernel<true><<<grid, kZWarps * 32, smem, st>>>(src_base, (zstd::BlockInfo*)d_blocks, (uint32_t)nb,
                                                                workers, nullptr, nullptr, nullptr, nullptr, d_status);
  } else {
    cudaFuncSetAttribute(zstd_entropy_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    zstd_entropy_kernel<false><<<grid, kZWarps * 32, smem, st>>>(src_base, (zstd::BlockInfo*)d_blocks, (uint32_t)nb,
                                                                 workers, d_ws, sq, sq + stride, sq + 2 * stride,
                                                                 d_status);
  }
  *launches += 1;
}
void launch_zstd_sum(const void* d_blocks, const uint64_t* d_cnt, const uint64_t* d_base, uint32_t n, uint64_t* d_olen,
                     int32_t* d_status, cudaStream_t st, uint64_t* launches) {
  if (!n) return;
  zstd_sum_kernel<<<(n + 127) / 128, 128, 0, st>>>((const zstd::BlockInfo*)d_blocks, d_cnt, d_base, n, d_olen, d_status);
  *launches += 1;
}
void launch_zstd_execute(const uint8_t* src_base, const void* d_blocks, const uint64_t* d_cnt, const uint64_t* d_base,
                         uint32_t n, const uint8_t* d_ws, uint64_t lit_bytes, uint64_t nseq, const uint64_t* d_olen,
                         uint8_t* dst_base, const uint64_t* d_dst_off, uint64_t dst_cap, int32_t* d_status,
                         cudaStream_t st, uint64_t* launches) {
  if (!n) return;
  const uint32_t workers = n < 16384u ? n : 16384u;
  const uint32_t* sq = reinterpret_cast<const uint32_t*>(d_ws + ((lit_bytes + 255) & ~(uint64_t)255));
  const size_t stride = (size_t)nseq + 64;
  zstd_execute_kernel<<<(workers + 3) / 4, 128, 0, st>>>(src_base, (const zstd::BlockInfo*)d_blocks, d_cnt, d_base, n,
                                                         workers, d_ws, sq, sq + stride, sq + 2 * stride, d_olen,
                                                         dst_base, d_dst_off, dst_cap, d_status);
  *launches += 1;
}

}  // namespace b2s
