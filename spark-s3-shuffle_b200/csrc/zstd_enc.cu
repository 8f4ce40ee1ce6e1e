// zstd_enc.cu — K7: Zstandard frame ENCODING on the device (encoder core: zstd_enc_core.h, shared with the host model).
//
// Replaces com.github.luben.zstd.ZstdOutputStreamNoFinalizer [U] (zstd-jni -> libzstd ZSTD_compressStream2) under
// SerializerManager.wrapStream on the streams of shuffle/S3ShuffleMapOutputWriter.scala:140-146 for
// spark.io.compression.codec=zstd.  Staged encoder: valid frames first — the shared LZ match finder and
// greedy parse (lz4_match_kernel, lz4_parse_kernel<2>), Raw_Literals + FSE sequences with the block's own tables where
// they pay (zstd_enc_core.h: histogram -> normalised counts -> table description + coding tables, per thread),
// Raw_Block fallback.  Huffman literals are the follow-up (worth nothing on terasort keys, ~20 % on text).
//   stream = frame header (6 B: magic, descriptor, 128 KiB window) | one block per codec block | empty last Raw_Block
//
//   zstd_seqenc_kernel  THREAD per block: code histograms, the block's own FSE tables (thread-private, 1.6 KB), then the
//                       sequence bitstream — a serial state machine (three interleaved FSE states, written in reverse
//                       sequence order) -> per-block scratch [modes + table descriptions | bitstream], Raw vs Compressed
//   zstd_emit_kernel    warp per block, lane per sequence: block + section headers, gathers the literals, moves the
//                       bitstream to its final packed position
// CPU model with identical output: tests/native/zstd_core_host.cpp::zc_compress_model.
#include "kernels.h"
#include "zstd_enc_core.h"

namespace b2s {

using zstdenc::CTables;
using zstdenc::kSeqHeaderMax;
static_assert(zstdenc::kSeqHeaderMax == 192, "zstd_bits_stride() in kernels.h reserves 192 bytes for the header");

__device__ __forceinline__ uint32_t find_stream_z(const uint32_t* __restrict__ blk_base, uint32_t n_streams, uint32_t b) {
  uint32_t lo = 0, hi = n_streams;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (blk_base[mid] <= b) lo = mid; else hi = mid;
  }
  return lo;
}

// Sequence i of a block as the encoder core wants it, read from the parse records and the off[] table.  The core walks
// the sequences strictly in order (forwards for the histograms, backwards for the bitstream); a record load and the
// dependent off[] gather are two trips to L2/DRAM, and issued on demand they were the whole kernel: ~2,800 cycles per
// sequence, 12.7 ms per 32,768 blocks wherever the tables lived (profiles/r2z_zstd.md).  The reader keeps six records and
// three offsets in flight in registers, i.e. every load is issued three to six sequences before its use.
struct SeqReader {
  const uint2* __restrict__ seq;
  const uint16_t* __restrict__ offp;
  uint32_t omask;
  int n, next, dir;
  uint2 r0, r1, r2, r3, r4, r5;
  uint32_t o0, o1, o2;
  __device__ __forceinline__ uint2 rec(int i) const { return seq[i < 0 ? 0 : (i >= n ? n - 1 : i)]; }
  __device__ __forceinline__ uint32_t off_of(const uint2 r) const { return offp[(r.x & 0xffffu) + (r.x >> 16)] & omask; }
  __device__ __forceinline__ void start(int i, int d) {
    next = i;
    dir = d;
    r0 = rec(i); r1 = rec(i + d); r2 = rec(i + 2 * d); r3 = rec(i + 3 * d); r4 = rec(i + 4 * d); r5 = rec(i + 5 * d);
    o0 = off_of(r0); o1 = off_of(r1); o2 = off_of(r2);
  }
  __device__ __forceinline__ zstdenc::Seq operator()(uint32_t i) {
    if ((int)i != next) start((int)i, i == 0 ? 1 : -1);
    zstdenc::Seq q;
    q.ll = r0.x >> 16;
    q.ml = r0.y & 0xffffu;
    q.off = o0;
    r0 = r1; r1 = r2; r2 = r3; r3 = r4; r4 = r5;
    r5 = rec(next + 6 * dir);
    o0 = o1; o1 = o2;
    o2 = off_of(r2);
    next += dir;
    return q;
  }
};

// records from lz4_parse_kernel<2>: x = literal start | literal count << 16 ; y = match length | literal position << 16
__global__ void __launch_bounds__(64) zstd_seqenc_kernel(
    const uint64_t* __restrict__ src_len, const uint32_t* __restrict__ blk_base, uint32_t n_streams, uint32_t b0,
    uint32_t m, uint32_t block_size, uint32_t stride, uint32_t max_seq, uint32_t bits_stride,
    const uint16_t* __restrict__ offarr, const uint2* __restrict__ seqarr, const uint32_t* __restrict__ nseq,
    const CTables* __restrict__ T, uint8_t* __restrict__ bits, uint32_t* __restrict__ nbits_out,
    uint32_t* __restrict__ csize, uint64_t* __restrict__ sizes) {
  const uint32_t bl = blockIdx.x * blockDim.x + threadIdx.x;
  if (bl >= m) return;
  const uint32_t b = b0 + bl;
  const uint32_t si = find_stream_z(blk_base, n_streams, b);
  const uint64_t rem = src_len[si] - (uint64_t)(b - blk_base[si]) * block_size;
  const uint32_t n = (uint32_t)(rem < block_size ? rem : block_size);
  const uint2* __restrict__ seq = seqarr + (size_t)bl * max_seq;
  const uint16_t* __restrict__ offp = offarr + (size_t)bl * stride;
  const uint32_t ns = nseq[b];  // >= 1: the last record holds the trailing literals
  const uint2 tail = seq[ns - 1];
  const uint32_t total_lit = (tail.y >> 16) + (tail.x >> 16);
  const uint32_t nreal = ns - 1;
  bool raw = nreal == 0;
  uint32_t nb = 0, cs = 0, hb = 0;
  if (!raw) {
    // The block's own coding tables are thread-private (local memory, 1.6 KB).  One BlockTables per thread in shared
    // memory was measured too: 2 CTAs of 64 threads per SM instead of 8, compress step 401 vs 207 ms per 10 GiB — this
    // kernel lives on occupancy (profiles/r2z_zstd.md).
    zstdenc::BlockTables local_tables;
    zstdenc::BlockTables* B = &local_tables;
    uint8_t* scratch = bits + (size_t)bl * bits_stride;
    SeqReader get;
    get.seq = seq;
    get.offp = offp;
    get.omask = stride <= 32768u ? 0x7fffu : 0xffffu;  // bit 15: parse flag (lz4_compress.cu)
    get.n = (int)nreal;
    get.next = -1;
    nb = zstdenc::encode_block_sequences(T, B, nreal, get, scratch, &hb, scratch + kSeqHeaderMax, n);
    cs = zstdenc::raw_literals_header_bytes(total_lit) + total_lit + zstdenc::nseq_header_bytes(nreal) + hb + nb;
    if (nb > n || cs >= n) raw = true;
  }
  nbits_out[b] = nb | (hb << 20);  // nb <= 2^16, hb <= kSeqHeaderMax
  csize[b] = raw ? (n | 0x80000000u) : cs;
  sizes[b] = 3u + (uint64_t)(raw ? n : cs);
}

constexpr int kZEmitThreads = 256;
__global__ void __launch_bounds__(kZEmitThreads) zstd_emit_kernel(
    const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off, const uint64_t* __restrict__ src_len,
    const uint32_t* __restrict__ blk_base, uint32_t n_streams, uint32_t b0, uint32_t m, uint32_t block_size,
    uint32_t max_seq, uint32_t bits_stride, const uint2* __restrict__ seqarr, const uint32_t* __restrict__ nseq,
    const uint8_t* __restrict__ bits, const uint32_t* __restrict__ nbits, const uint32_t* __restrict__ csize,
    const uint64_t* __restrict__ scan, uint8_t* __restrict__ dst_base, uint64_t dst_cap) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const uint32_t bl = blockIdx.x * (kZEmitThreads / 32) + (threadIdx.x >> 5);
  if (bl >= m) return;
  const uint32_t b = b0 + bl;
  const uint32_t si = find_stream_z(blk_base, n_streams, b);
  const uint64_t boff = (uint64_t)(b - blk_base[si]) * block_size;
  const uint64_t rem = src_len[si] - boff;
  const uint32_t n = (uint32_t)(rem < block_size ? rem : block_size);
  const uint8_t* __restrict__ s = src_base + src_off[si] + boff;
  const uint32_t cs = csize[b];
  const bool raw = cs & 0x80000000u;
  const uint32_t payload = cs & 0x7fffffffu;
  // stream i: 6-byte frame header, its blocks, 3-byte end block  =>  9 bytes of overhead per stream before mine
  const uint64_t o0 = scan[b] + 9ull * si + zstdenc::kFrameHeaderBytes;
  if (o0 + 3ull + payload > dst_cap) return;  // the stream-meta kernel reports B2S_E_DST_TOO_SMALL
  uint8_t* __restrict__ o = dst_base + o0;
  if (lane == 0) zstdenc::put_block_header(o, 0, raw ? 0 : 2, payload);
  if (raw) {
    group_copy<32>(o + 3, s, n, lane);
    return;
  }
  const uint32_t ns = nseq[b];
  const uint2* __restrict__ seq = seqarr + (size_t)bl * max_seq;
  const uint2 tail = seq[ns - 1];
  const uint32_t total_lit = (tail.y >> 16) + (tail.x >> 16);
  const uint32_t hl = zstdenc::raw_literals_header_bytes(total_lit);
  if (lane == 0) zstdenc::put_raw_literals_header(o + 3, total_lit);
  uint8_t* __restrict__ lits = o + 3 + hl;
  for (uint32_t i0 = 0; i0 < ns; i0 += 32) {  // literal gather: lane per sequence (the tail record included)
    const uint32_t i = i0 + lane;
    int anchor = 0, lit = 0, lpos = 0;
    if (i < ns) {
      const uint2 r = seq[i];
      anchor = (int)(r.x & 0xffffu);
      lit = (int)(r.x >> 16);
      lpos = (int)(r.y >> 16);
      if (lit <= 16)
        for (int j = 0; j < lit; j++) lits[lpos + j] = __ldg(s + anchor + j);
    }
    unsigned big = __ballot_sync(FULL, lit > 16);
    while (big) {
      const int l = __ffs(big) - 1;
      big &= big - 1;
      const int a_r = __shfl_sync(FULL, anchor, l), n_r = __shfl_sync(FULL, lit, l), p_r = __shfl_sync(FULL, lpos, l);
      if (n_r >= 96) group_copy<32>(lits + p_r, s + a_r, (uint32_t)n_r, lane);
      else
        for (int j = lane; j < n_r; j += 32) lits[p_r + j] = __ldg(s + a_r + j);
    }
  }
  uint8_t* __restrict__ q = lits + total_lit;
  const uint32_t nreal = ns - 1;
  const uint32_t hb = zstdenc::nseq_header_bytes(nreal);
  if (lane == 0) zstdenc::put_nseq(q, nreal);
  // Compression_Modes + the block's own table descriptions, then the bitstream (zstd_seqenc_kernel's scratch layout)
  const uint8_t* __restrict__ scratch = bits + (size_t)bl * bits_stride;
  const uint32_t packed = nbits[b], nbs = packed & 0xfffffu, th = packed >> 20;
  for (uint32_t j = lane; j < th; j += 32) q[hb + j] = scratch[j];
  group_copy<32>(q + hb + th, scratch + kSeqHeaderMax, nbs, lane);
}

// per stream: frame header at the start, empty last Raw_Block at the end, packed offset/length, capacity check
__global__ void zstd_stream_meta_kernel(const uint32_t* __restrict__ blk_base, uint32_t n_streams, uint32_t n_blocks,
                                        const uint64_t* __restrict__ scan, const uint64_t* __restrict__ scan_total,
                                        uint8_t* __restrict__ dst_base, uint64_t dst_cap,
                                        uint64_t* __restrict__ dst_off, uint64_t* __restrict__ dst_len,
                                        int32_t* __restrict__ status) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_streams) return;
  const uint32_t b0 = blk_base[i], b1 = blk_base[i + 1];
  const uint64_t s0 = b0 < n_blocks ? scan[b0] : *scan_total;
  const uint64_t s1 = b1 < n_blocks ? scan[b1] : *scan_total;
  const uint64_t off = s0 + 9ull * i;
  const uint64_t len = (s1 - s0) + 9ull;
  dst_off[i] = off;
  dst_len[i] = len;
  if (off + len > dst_cap) {
    status[i] = B2S_E_DST_TOO_SMALL;
    return;
  }
  zstdenc::put_frame_header(dst_base + off);
  zstdenc::put_block_header(dst_base + off + len - 3, 1, 0, 0);
}

int zstd_ctables_create(void** d_tables) {
  CTables h;
  zstdenc::build_predefined(&h);
  if (cudaMalloc(d_tables, sizeof(CTables)) != cudaSuccess) return -1;
  if (cudaMemcpy(*d_tables, &h, sizeof(CTables), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  return 0;
}
void zstd_ctables_destroy(void* d_tables) {
  if (d_tables) cudaFree(d_tables);
}

static const CTables* g_ctables_for_device[64] = {nullptr};
void zstd_set_ctables(int ordinal, const void* d_tables) {
  if (ordinal >= 0 && ordinal < 64) g_ctables_for_device[ordinal] = reinterpret_cast<const CTables*>(d_tables);
}

void launch_zstd_seqenc(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len,
                        const uint32_t* d_blk_base, uint32_t n_streams, uint32_t b0, uint32_t m, uint32_t block_size,
                        uint32_t stride, uint32_t max_seq, const uint16_t* d_off, const uint2* d_seq,
                        const uint32_t* d_nseq, uint8_t* d_bits, uint32_t* d_nbits, uint32_t* d_csize,
                        uint64_t* d_sizes, cudaStream_t st, uint64_t* launches) {
  (void)src_base;
  (void)d_src_off;
  if (!m) return;
  int dev = 0;
  cudaGetDevice(&dev);
  zstd_seqenc_kernel<<<(m + 63) / 64, 64, 0, st>>>(d_src_len, d_blk_base, n_streams, b0, m, block_size, stride, max_seq,
                                                  (uint32_t)zstd_bits_stride(stride), d_off, d_seq, d_nseq,
                                                  g_ctables_for_device[dev & 63], d_bits, d_nbits, d_csize, d_sizes);
  *launches += 1;
}

void launch_zstd_emit(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len,
                      const uint32_t* d_blk_base, uint32_t n_streams, uint32_t b0, uint32_t m, uint32_t block_size,
                      uint32_t stride, uint32_t max_seq, const uint2* d_seq, const uint32_t* d_nseq,
                      const uint8_t* d_bits, const uint32_t* d_nbits, const uint32_t* d_csize, const uint64_t* d_scan,
                      uint8_t* dst_base, uint64_t dst_cap, cudaStream_t st, uint64_t* launches) {
  if (!m) return;
  zstd_emit_kernel<<<(m + kZEmitThreads / 32 - 1) / (kZEmitThreads / 32), kZEmitThreads, 0, st>>>(
      src_base, d_src_off, d_src_len, d_blk_base, n_streams, b0, m, block_size, max_seq, (uint32_t)zstd_bits_stride(stride),
      d_seq, d_nseq,
      d_bits, d_nbits, d_csize, d_scan, dst_base, dst_cap);
  *launches += 1;
}

void launch_zstd_stream_meta(const uint32_t* d_blk_base, uint32_t n_streams, uint32_t n_blocks, const uint64_t* d_scan,
                             const uint64_t* d_scan_total, uint8_t* dst_base, uint64_t dst_cap, uint64_t* d_dst_off,
                             uint64_t* d_dst_len, int32_t* d_status, cudaStream_t st, uint64_t* launches) {
  if (!n_streams) return;
  zstd_stream_meta_kernel<<<(n_streams + 255) / 256, 256, 0, st>>>(d_blk_base, n_streams, n_blocks, d_scan, d_scan_total,
                                                                   dst_base, dst_cap, d_dst_off, d_dst_len, d_status);
  *launches += 1;
}

}  // namespace b2s
