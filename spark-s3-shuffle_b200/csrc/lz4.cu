// lz4.cu — K3 (LZ4 block compress), K4 (LZ4 block decompress) and the lz4-java LZ4Block stream framing.
//
// Replaces, for spark.io.compression.codec=lz4, what Spark's SerializerManager.wrapStream puts around the streams
// of shuffle/S3ShuffleMapOutputWriter.scala:140-146 (write) and storage/S3ShuffleReader.scala:107-109 (read):
//   net.jpountz.lz4.LZ4BlockOutputStream -> LZ4JNICompressor  -> liblz4 LZ4_compress_default      [U]
//   net.jpountz.lz4.LZ4BlockInputStream  -> LZ4JNIFastDecompressor -> liblz4 LZ4_decompress_fast  [U]
// Wire format (oracle/b2s_oracle.c restates it): per <= blockSize bytes of input a 21-byte header
//   "LZ4Block" | token = method(0x10 RAW, 0x20 LZ4) | level | compressedLen LE32 | originalLen LE32 | XXH32&0x0FFFFFFF LE32
// then the payload; a stream ends with a header whose three ints are zero.
//
// Parallelism: every codec block (<= 64 KiB here) is independent.  A *tile* of TILE lanes (8/16/32) owns one block.
//  compress : the tile scans TILE consecutive positions per step (hash -> shared-memory u16 table -> verify 4 bytes),
//             the lowest matching lane wins (ballot/ffs), the tile extends the match TILE bytes per ballot, emits
//             the sequence cooperatively, then inserts the window's positions.  Deterministic; the CPU model
//             orc_lz4_compress_block_tile() in oracle/ produces identical bytes.
//  decompress: token chain is serial; the tile parses uniformly and copies literals / (overlapping) matches
//             TILE bytes per step.
// This is latency/issue-bound byte-stream work (no tensor cores, HBM far from saturated by one block per tile), so
// the levers are blocks in flight per SM (small per-tile state: 4-8 KiB hash table, no staged copy of the block)
// and instructions per sequence.
#include "kernels.h"

namespace b2s {

constexpr int kLz4Threads = 128;
constexpr int kMinMatch = 4, kMFLimit = 12, kLastLiterals = 5;

__device__ __forceinline__ uint32_t find_stream32(const uint32_t* __restrict__ blk_base, uint32_t n_streams,
                                                  uint32_t b) {
  uint32_t lo = 0, hi = n_streams;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (blk_base[mid] <= b) lo = mid; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------------------
// K3: compress
// ------------------------------------------------------------------------------------------------------------
template <int TILE>
__device__ __forceinline__ void tile_store_len_ext(uint8_t* o, int nb, int r, int lane) {
  // nb bytes: 255 ... 255, r - 255*(nb-1)
  for (int j = lane; j < nb; j += TILE) o[j] = (j == nb - 1) ? (uint8_t)(r - 255 * (nb - 1)) : (uint8_t)255;
}

template <int TILE>
__device__ __forceinline__ void tile_copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int n,
                                                int lane) {
  if (n >= 96) {
    group_copy<TILE>(dst, src, (uint32_t)n, lane);
  } else {
    for (int j = lane; j < n; j += TILE) dst[j] = __ldg(src + j);
  }
}

template <int TILE, int HLOG>
__global__ void __launch_bounds__(kLz4Threads) lz4_compress_kernel(
    const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off, const uint64_t* __restrict__ src_len,
    const uint32_t* __restrict__ blk_base, uint32_t n_streams, uint32_t n_blocks, uint32_t block_size,
    uint8_t* __restrict__ scratch, uint32_t* __restrict__ csize, uint64_t* __restrict__ sizes,
    unsigned int* __restrict__ work_counter) {
  extern __shared__ __align__(16) uint16_t smem_tables[];
  const int lane = threadIdx.x % TILE;
  const int tile_in_cta = threadIdx.x / TILE;
  uint16_t* table = smem_tables + (size_t)tile_in_cta * (1 << HLOG);

  for (;;) {
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1u);
    b = tile_shfl<TILE>(b, 0);
    if (b >= n_blocks) break;

    const uint32_t si = find_stream32(blk_base, n_streams, b);
    const uint64_t boff = (uint64_t)(b - blk_base[si]) * block_size;
    const uint64_t rem = src_len[si] - boff;
    const int n = (int)(rem < block_size ? rem : block_size);
    const uint8_t* __restrict__ s = src_base + src_off[si] + boff;
    uint8_t* __restrict__ out = scratch + (uint64_t)b * block_size;
    const int cap = n - 1;  // LZ4BlockOutputStream stores RAW when compressedLength >= originalLength

    // zero the table (16 bytes per lane per step)
    {
      uint4* t4 = reinterpret_cast<uint4*>(table);
      for (int j = lane; j < (1 << HLOG) / 8; j += TILE) t4[j] = make_uint4(0, 0, 0, 0);
    }
    tile_sync<TILE>();

    int op = 0, anchor = 0, pos = 0;
    bool fail = false;
    if (n >= kMFLimit + 1) {
      const int mflimit = n - kMFLimit;
      const int matchlimit = n - kLastLiterals;
      while (pos <= mflimit) {
        const int p = pos + lane;
        const bool valid = p <= mflimit;
        const uint32_t v = valid ? ld32u_ro(s + p) : 0u;
        const uint32_t h = (v * 2654435761u) >> (32 - HLOG);
        const int cand = table[h];
        const bool ok = valid && cand < p && ld32u_ro(s + cand) == v;
        const unsigned bal = tile_ballot<TILE>(ok);
        int next;
        if (bal) {
          const int first = __ffs(bal) - 1;
          const int m = pos + first;
          const int c = tile_shfl<TILE>(cand, first);
          // forward extension, TILE bytes per ballot
          int mlen;
          {
            const int maxl = matchlimit - m;
            int k = kMinMatch + lane;
            for (;;) {
              const bool ne = (k >= maxl) || (__ldg(s + m + k) != __ldg(s + c + k));
              const unsigned nb = tile_ballot<TILE>(ne);
              if (nb) {
                mlen = k - lane + __ffs(nb) - 1;
                break;
              }
              k += TILE;
            }
          }
          // emit: token | litlen ext | literals | offset | matchlen ext
          const int lit = m - anchor;
          const int ml = mlen - kMinMatch;
          const int nbL = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
          const int nbM = ml >= 15 ? (ml - 15) / 255 + 1 : 0;
          const int need = 1 + nbL + lit + 2 + nbM;
          if (op + need > cap) {
            fail = true;
            break;
          }
          uint8_t* o = out + op;
          if (lane == 0) o[0] = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (ml < 15 ? ml : 15));
          if (nbL) tile_store_len_ext<TILE>(o + 1, nbL, lit - 15, lane);
          tile_copy_bytes<TILE>(o + 1 + nbL, s + anchor, lit, lane);
          uint8_t* q = o + 1 + nbL + lit;
          const int off = m - c;
          if (lane == 0) q[0] = (uint8_t)off;
          if (lane == 1 % TILE) q[1] = (uint8_t)(off >> 8);
          if (nbM) tile_store_len_ext<TILE>(q + 2, nbM, ml - 15, lane);
          op += need;
          next = m + mlen;
          anchor = next;
        } else {
          next = pos + TILE;
        }
        tile_sync<TILE>();  // all lanes have read the table for this window
        // insert the window's positions that precede the next scan position.  Two lanes may hash to one slot and the
        // hardware's same-address store winner is unspecified, so losers with a *later* position re-store until the
        // slot holds the highest position (= what a sequential scan would leave).  Usually one read-back, no retry.
        const bool ins = valid && p < next;
        if (ins) table[h] = (uint16_t)p;
        tile_sync<TILE>();
        for (;;) {
          const bool lost = ins && table[h] < (uint16_t)p;
          if (!tile_ballot<TILE>(lost)) break;
          if (lost) table[h] = (uint16_t)p;
          tile_sync<TILE>();
        }
        pos = next;
      }
    }
    if (!fail) {
      const int lit = n - anchor;
      const int nbL = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
      const int need = 1 + nbL + lit;
      if (op + need > cap) {
        fail = true;
      } else {
        uint8_t* o = out + op;
        if (lane == 0) o[0] = (uint8_t)((lit < 15 ? lit : 15) << 4);
        if (nbL) tile_store_len_ext<TILE>(o + 1, nbL, lit - 15, lane);
        tile_copy_bytes<TILE>(o + 1 + nbL, s + anchor, lit, lane);
        op += need;
      }
    }
    if (lane == 0) {
      csize[b] = fail ? ((uint32_t)n | 0x80000000u) : (uint32_t)op;
      sizes[b] = 21u + (uint64_t)(fail ? n : op);
    }
    tile_sync<TILE>();
  }
}

template <int TILE, int HLOG>
static void launch_lz4_compress_t(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len,
                                  const uint32_t* d_blk_base, uint32_t n_streams, uint32_t n_blocks,
                                  uint32_t block_size, uint8_t* d_scratch, uint32_t* d_csize, uint64_t* d_sizes,
                                  unsigned int* d_counter, cudaStream_t st) {
  constexpr int kTiles = kLz4Threads / TILE;
  const size_t smem = (size_t)kTiles * (2u << HLOG);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(lz4_compress_kernel<TILE, HLOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lz4_compress_kernel<TILE, HLOG>, kLz4Threads, smem);
  if (per_sm < 1) per_sm = 1;
  uint64_t want = ((uint64_t)n_blocks + kTiles - 1) / kTiles;
  uint64_t grid = (uint64_t)kSMs * per_sm;
  if (grid > want) grid = want;
  lz4_compress_kernel<TILE, HLOG><<<(unsigned)grid, kLz4Threads, smem, st>>>(
      src_base, d_src_off, d_src_len, d_blk_base, n_streams, n_blocks, block_size, d_scratch, d_csize, d_sizes,
      d_counter);
}

int g_lz4_tile = 16;   // tuning knobs (api.cu reads B2S_LZ4_TILE / B2S_LZ4_HLOG / B2S_LZ4D_TILE once at init)
int g_lz4_hlog = 12;
int g_lz4d_tile = 16;

void launch_lz4_compress(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len,
                         const uint32_t* d_blk_base, uint32_t n_streams, uint32_t n_blocks, uint32_t block_size,
                         uint8_t* d_scratch, uint32_t* d_csize, uint64_t* d_sizes, unsigned int* d_counter,
                         cudaStream_t st, uint64_t* launches) {
  if (!n_blocks) return;
  cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), st);
#define B2S_LZ4C(T, H)                                                                                              \
  launch_lz4_compress_t<T, H>(src_base, d_src_off, d_src_len, d_blk_base, n_streams, n_blocks, block_size, d_scratch, \
                              d_csize, d_sizes, d_counter, st)
  const int key = g_lz4_tile * 100 + g_lz4_hlog;
  switch (key) {
    case 811: B2S_LZ4C(8, 11); break;
    case 812: B2S_LZ4C(8, 12); break;
    case 1611: B2S_LZ4C(16, 11); break;
    case 1613: B2S_LZ4C(16, 13); break;
    case 3211: B2S_LZ4C(32, 11); break;
    case 3212: B2S_LZ4C(32, 12); break;
    case 3213: B2S_LZ4C(32, 13); break;
    case 1612:
    default: B2S_LZ4C(16, 12); break;
  }
#undef B2S_LZ4C
  *launches += 1;
}

// ------------------------------------------------------------------------------------------------------------
// LZ4Block framing, write side
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lz4b_level(uint32_t block_size) {
  int bits = 32 - __clz(block_size - 1);
  int lvl = bits - 10;
  return lvl < 0 ? 0 : lvl;
}

__device__ __forceinline__ uint8_t lz4b_header_byte(int j, int token, uint32_t clen, uint32_t olen, uint32_t check) {
  // "LZ4Block" = 4C 5A 34 42 6C 6F 63 6B
  const uint64_t magic = 0x6B636F6C42345A4Cull;
  if (j < 8) return (uint8_t)(magic >> (8 * j));
  if (j == 8) return (uint8_t)token;
  if (j < 13) return (uint8_t)(clen >> (8 * (j - 9)));
  if (j < 17) return (uint8_t)(olen >> (8 * (j - 13)));
  return (uint8_t)(check >> (8 * (j - 17)));
}

// per stream: packed offset/length, end mark, capacity check
__global__ void lz4block_stream_meta_kernel(const uint32_t* __restrict__ blk_base, uint32_t n_streams,
                                            uint32_t n_blocks, uint32_t block_size, const uint64_t* __restrict__ scan,
                                            const uint64_t* __restrict__ scan_total, uint8_t* __restrict__ dst_base,
                                            uint64_t dst_cap, uint64_t* __restrict__ dst_off,
                                            uint64_t* __restrict__ dst_len, int32_t* __restrict__ status) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_streams) return;
  const uint32_t b0 = blk_base[i], b1 = blk_base[i + 1];
  const uint64_t s0 = b0 < n_blocks ? scan[b0] : *scan_total;
  const uint64_t s1 = b1 < n_blocks ? scan[b1] : *scan_total;
  const uint64_t off = s0 + 21ull * i;
  const uint64_t len = (s1 - s0) + 21ull;
  dst_off[i] = off;
  dst_len[i] = len;
  if (off + len > dst_cap) {
    status[i] = B2S_E_DST_TOO_SMALL;
    return;
  }
  uint8_t* e = dst_base + off + len - 21;
  const int token = 0x10 | lz4b_level(block_size);
#pragma unroll
  for (int j = 0; j < 21; j++) e[j] = lz4b_header_byte(j, token, 0, 0, 0);
}

// one warp per codec block: header + payload to the packed position
__global__ void __launch_bounds__(256) lz4block_pack_kernel(
    const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off, const uint64_t* __restrict__ src_len,
    const uint32_t* __restrict__ blk_base, uint32_t n_streams, uint32_t n_blocks, uint32_t block_size,
    const uint8_t* __restrict__ scratch, const uint32_t* __restrict__ csize, const uint32_t* __restrict__ hash,
    const uint64_t* __restrict__ scan, uint8_t* __restrict__ dst_base, uint64_t dst_cap,
    const int32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= n_blocks) return;
  const uint32_t si = find_stream32(blk_base, n_streams, b);
  if (status[si] != 0) return;
  const uint64_t boff = (uint64_t)(b - blk_base[si]) * block_size;
  const uint64_t rem = src_len[si] - boff;
  const uint32_t olen = (uint32_t)(rem < block_size ? rem : block_size);
  const uint32_t cs = csize[b];
  const bool raw = cs & 0x80000000u;
  const uint32_t clen = cs & 0x7fffffffu;
  uint8_t* o = dst_base + scan[b] + 21ull * si;
  const int token = (raw ? 0x10 : 0x20) | lz4b_level(block_size);
  if (lane < 21) o[lane] = lz4b_header_byte(lane, token, clen, olen, hash[b] & 0x0FFFFFFFu);
  const uint8_t* payload = raw ? src_base + src_off[si] + boff : scratch + (uint64_t)b * block_size;
  group_copy<32>(o + 21, payload, clen, lane);
}

void launch_lz4block_pack(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len,
                          const uint32_t* d_blk_base, uint32_t n_streams, uint32_t n_blocks, uint32_t block_size,
                          const uint8_t* d_scratch, const uint32_t* d_csize, const uint32_t* d_hash,
                          const uint64_t* d_scan, const uint64_t* d_scan_total, uint8_t* dst_base, uint64_t dst_cap,
                          uint64_t* d_dst_off, uint64_t* d_dst_len, int32_t* d_status, cudaStream_t st,
                          uint64_t* launches) {
  if (!n_streams) return;
  lz4block_stream_meta_kernel<<<(n_streams + 255) / 256, 256, 0, st>>>(d_blk_base, n_streams, n_blocks, block_size,
                                                                       d_scan, d_scan_total, dst_base, dst_cap,
                                                                       d_dst_off, d_dst_len, d_status);
  *launches += 1;
  if (n_blocks) {
    lz4block_pack_kernel<<<(n_blocks + 7) / 8, 256, 0, st>>>(src_base, d_src_off, d_src_len, d_blk_base, n_streams,
                                                            n_blocks, block_size, d_scratch, d_csize, d_hash, d_scan,
                                                            dst_base, dst_cap, d_status);
    *launches += 1;
  }
}

// ------------------------------------------------------------------------------------------------------------
// LZ4Block framing, read side: header walk (LZ4BlockInputStream.refill [U])
// ------------------------------------------------------------------------------------------------------------
struct Lz4bHeader {
  int method;
  int level;
  int32_t clen, olen;
  uint32_t check;
  bool magic_ok;
};
__device__ __forceinline__ Lz4bHeader lz4b_read_header(const uint8_t* p) {
  Lz4bHeader h;
  h.magic_ok = ld32u_ro(p) == 0x42345A4Cu && ld32u_ro(p + 4) == 0x6B636F6Cu;
  const int token = p[8];
  h.method = token & 0xF0;
  h.level = 10 + (token & 0x0F);
  h.clen = (int32_t)ld32u_ro(p + 9);
  h.olen = (int32_t)ld32u_ro(p + 13);
  h.check = ld32u_ro(p + 17);
  return h;
}
__device__ __forceinline__ bool lz4b_header_valid(const Lz4bHeader& h) {
  if (!h.magic_ok) return false;
  if (h.method != 0x10 && h.method != 0x20) return false;
  if (h.olen > (1 << h.level) || h.olen < 0 || h.clen < 0 || (h.olen == 0 && h.clen != 0) ||
      (h.olen != 0 && h.clen == 0) || (h.method == 0x10 && h.olen != h.clen))
    return false;
  return true;
}

// FILL=false: count blocks/bytes per stream.  FILL=true: write descriptors.
template <bool FILL>
__global__ void lz4block_walk_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                                     const uint64_t* __restrict__ src_len, uint32_t n, uint64_t* __restrict__ nblk,
                                     uint64_t* __restrict__ olen_total, const uint64_t* __restrict__ blk_base,
                                     const uint64_t* __restrict__ dst_off, uint64_t dst_cap,
                                     int32_t* __restrict__ status, BlockDesc* __restrict__ desc) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (FILL && status[i] != 0 && status[i] != B2S_E_DST_TOO_SMALL) return;
  if (!FILL && status[i] != 0) {  // failed checksum verification: nothing of this block is decoded
    nblk[i] = 0;
    olen_total[i] = 0;
    return;
  }
  const uint8_t* p = src_base + src_off[i];
  const uint64_t len = src_len[i];
  uint64_t ip = 0, cnt = 0, total = 0;
  bool bad = false;
  bool too_small = false;
  if (FILL) too_small = (dst_off[i] + olen_total[i] > dst_cap);
  while (ip < len) {
    if (len - ip < 21) {
      bad = true;
      break;
    }
    const Lz4bHeader h = lz4b_read_header(p + ip);
    if (!lz4b_header_valid(h)) {
      bad = true;
      break;
    }
    ip += 21;
    if (h.olen == 0 && h.clen == 0) {
      if (h.check != 0) {
        bad = true;
        break;
      }
      continue;  // end mark; stopOnEmptyBlock=false: a concatenated stream may follow
    }
    if ((uint64_t)h.clen > len - ip) {
      bad = true;
      break;
    }
    if (FILL) {
      BlockDesc d;
      d.src = src_off[i] + ip;
      d.dst = dst_off[i] + total;
      d.clen = too_small ? 0u : (uint32_t)h.clen;
      d.olen = too_small ? 0u : (uint32_t)h.olen;
      d.check = h.check;
      d.stream = i | (h.method == 0x10 ? 0x80000000u : 0u);
      desc[blk_base[i] + cnt] = d;
    }
    cnt++;
    total += (uint64_t)h.olen;
    ip += (uint64_t)h.clen;
  }
  if (!FILL) {
    if (bad) {
      status[i] = B2S_E_CORRUPT;
      cnt = 0;
      total = 0;
    }
    nblk[i] = cnt;
    olen_total[i] = total;
  } else if (too_small) {
    status[i] = B2S_E_DST_TOO_SMALL;
  }
}

void launch_lz4block_count(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len, uint32_t n,
                           uint64_t* d_nblk, uint64_t* d_olen, int32_t* d_status, cudaStream_t st,
                           uint64_t* launches) {
  if (!n) return;
  lz4block_walk_kernel<false><<<(n + 127) / 128, 128, 0, st>>>(src_base, d_src_off, d_src_len, n, d_nblk, d_olen,
                                                              nullptr, nullptr, 0, d_status, nullptr);
  *launches += 1;
}

void launch_lz4block_fill(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len, uint32_t n,
                          const uint64_t* d_blk_base, const uint64_t* d_dst_off, uint64_t* d_olen, uint64_t dst_cap,
                          int32_t* d_status, BlockDesc* d_desc, cudaStream_t st, uint64_t* launches) {
  if (!n) return;
  lz4block_walk_kernel<true><<<(n + 127) / 128, 128, 0, st>>>(src_base, d_src_off, d_src_len, n, nullptr, d_olen,
                                                             d_blk_base, d_dst_off, dst_cap, d_status, d_desc);
  *launches += 1;
}

// ------------------------------------------------------------------------------------------------------------
// K4: decompress (LZ4_decompress_fast semantics: driven by originalLen, must consume exactly compressedLen)
// ------------------------------------------------------------------------------------------------------------
template <int TILE>
__global__ void __launch_bounds__(kLz4Threads) lz4_decompress_kernel(const BlockDesc* __restrict__ desc,
                                                                     uint32_t n_blocks,
                                                                     const uint8_t* __restrict__ src_base,
                                                                     uint8_t* __restrict__ dst_base,
                                                                     int32_t* __restrict__ status,
                                                                     unsigned int* __restrict__ work_counter) {
  const int lane = threadIdx.x % TILE;
  for (;;) {
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1u);
    b = tile_shfl<TILE>(b, 0);
    if (b >= n_blocks) break;
    const BlockDesc d = desc[b];
    const int olen = (int)d.olen, clen = (int)d.clen;
    if (olen == 0) continue;
    const uint8_t* __restrict__ in = src_base + d.src;
    uint8_t* out = dst_base + d.dst;
    if (d.stream & 0x80000000u) {  // stored RAW
      group_copy<TILE>(out, in, (uint32_t)olen, lane);
      continue;
    }
    int ip = 0, op = 0;
    bool err = false;
    for (;;) {
      if (ip >= clen) { err = true; break; }
      const int token = __ldg(in + ip++);
      int ll = token >> 4;
      if (ll == 15) {
        int bb;
        do {
          if (ip >= clen) { err = true; break; }
          bb = __ldg(in + ip++);
          ll += bb;
        } while (bb == 255);
        if (err) break;
      }
      if (ll > olen - op || ll > clen - ip) { err = true; break; }
      if (ll >= 96) {
        group_copy<TILE>(out + op, in + ip, (uint32_t)ll, lane);
      } else {
        for (int j = lane; j < ll; j += TILE) out[op + j] = __ldg(in + ip + j);
      }
      op += ll;
      ip += ll;
      if (olen - op < kMFLimit) {
        if (op != olen) err = true;  // last match must start >= 12 bytes before the end of the block
        break;
      }
      if (ip + 2 > clen) { err = true; break; }
      const int off = __ldg(in + ip) | (__ldg(in + ip + 1) << 8);
      ip += 2;
      int ml = token & 15;
      if (ml == 15) {
        int bb;
        do {
          if (ip >= clen) { err = true; break; }
          bb = __ldg(in + ip++);
          ml += bb;
        } while (bb == 255);
        if (err) break;
      }
      ml += kMinMatch;
      if (ml > olen - op || off == 0 || off > op) { err = true; break; }
      tile_sync<TILE>();  // literals of this sequence visible to the whole tile
      {
        const uint8_t* msrc = out + op - off;
        if (off >= ml) {
          for (int j = lane; j < ml; j += TILE) out[op + j] = msrc[j];
        } else {
          // overlapping match: every byte comes from the already complete window [op-off, op)
          for (int j = lane; j < ml; j += TILE) out[op + j] = msrc[j % off];
        }
      }
      op += ml;
      tile_sync<TILE>();
      if (olen - op < kLastLiterals) { err = true; break; }
    }
    if (err || ip != clen) {
      if (lane == 0) set_status(status, d.stream & 0x7fffffffu, B2S_E_CORRUPT);
    }
  }
}

template <int TILE>
static void launch_lz4_decompress_t(const BlockDesc* d_desc, uint32_t n_blocks, const uint8_t* src_base,
                                    uint8_t* dst_base, int32_t* d_status, unsigned int* d_counter, cudaStream_t st) {
  constexpr int kTiles = kLz4Threads / TILE;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lz4_decompress_kernel<TILE>, kLz4Threads, 0);
  if (per_sm < 1) per_sm = 1;
  uint64_t want = ((uint64_t)n_blocks + kTiles - 1) / kTiles;
  uint64_t grid = (uint64_t)kSMs * per_sm;
  if (grid > want) grid = want;
  lz4_decompress_kernel<TILE><<<(unsigned)grid, kLz4Threads, 0, st>>>(d_desc, n_blocks, src_base, dst_base, d_status,
                                                                     d_counter);
}

void launch_lz4_decompress(const BlockDesc* d_desc, uint32_t n_blocks, const uint8_t* src_base, uint8_t* dst_base,
                           int32_t* d_status, unsigned int* d_counter, cudaStream_t st, uint64_t* launches) {
  if (!n_blocks) return;
  cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), st);
  switch (g_lz4d_tile) {
    case 8: launch_lz4_decompress_t<8>(d_desc, n_blocks, src_base, dst_base, d_status, d_counter, st); break;
    case 32: launch_lz4_decompress_t<32>(d_desc, n_blocks, src_base, dst_base, d_status, d_counter, st); break;
    default: launch_lz4_decompress_t<16>(d_desc, n_blocks, src_base, dst_base, d_status, d_counter, st); break;
  }
  *launches += 1;
}

}  // namespace b2s
