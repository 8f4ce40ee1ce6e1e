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
// Parallelism: every codec block (<= 64 KiB here) is independent.
//  compress : one warp per block, window-batched greedy parse (see K3 below).  Deterministic; the CPU model
//             orc_lz4_compress_block_win() in oracle/ produces identical bytes.
//  decompress: a tile of TILE lanes (4/8/16/32) per block; the token chain is serial, the tile parses uniformly and
//             copies literals / (overlapping) matches TILE bytes per step.
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
// K3: compress — one warp per codec block, "window-batched greedy" (specification: orc_lz4_compress_block_win)
//
//   per window of 32 positions:
//     1. every lane hashes its 4 bytes, reads the candidate from the shared-memory table (state before the window;
//        byte runs use an explicit offset-1 candidate), verifies 4 bytes and measures the match locally up to
//        8 bytes                                                                          (all lanes in parallel)
//     2. the warp walks the window greedily with uniform bit operations on the ballot mask: lowest matching
//        position >= cursor, its length by shuffle (cooperative extension only when the local 8 bytes were all
//        equal) — about a dozen instructions per selected sequence
//     3. literal counts, encoded sizes and output offsets of all selected sequences by one warp suffix-sum
//     4. each selected lane emits its own sequence (token, <= 16 literals, offset, <= 2 length bytes); rare longer
//        literal runs / lengths go through a cooperative slow path
//     5. all 32 positions are inserted; a slot hit twice is settled by a read-back so the highest position wins
//   Measured motivation (profiles/): terasort-shaped data has ~4400 sequences per 32 KiB block (7.5 B each), so the
//   per-sequence instruction count is what bounds this kernel, not HBM.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_store_len_ext(uint8_t* o, int nb, int r, int lane) {
  for (int j = lane; j < nb; j += 32) o[j] = (j == nb - 1) ? (uint8_t)(r - 255 * (nb - 1)) : (uint8_t)255;
}
__device__ __forceinline__ void warp_copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int n,
                                                int lane) {
  if (n >= 96) {
    group_copy<32>(dst, src, (uint32_t)n, lane);
  } else {
    for (int j = lane; j < n; j += 32) dst[j] = __ldg(src + j);
  }
}
// cooperative emit of one sequence (mlen == 0: final literal run without a match part)
__device__ __forceinline__ void warp_emit_seq(uint8_t* __restrict__ o, const uint8_t* __restrict__ lit_src, int lit,
                                              int off, int mlen, int lane) {
  const int ml = mlen - kMinMatch;
  const int nbL = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
  if (lane == 0) o[0] = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (mlen ? (ml < 15 ? ml : 15) : 0));
  if (nbL) warp_store_len_ext(o + 1, nbL, lit - 15, lane);
  warp_copy_bytes(o + 1 + nbL, lit_src, lit, lane);
  if (mlen) {
    uint8_t* q = o + 1 + nbL + lit;
    if (lane == 0) q[0] = (uint8_t)off;
    if (lane == 1) q[1] = (uint8_t)(off >> 8);
    if (ml >= 15) warp_store_len_ext(q + 2, (ml - 15) / 255 + 1, ml - 15, lane);
  }
}
template <int HLOG>
__global__ void __launch_bounds__(kLz4Threads, (HLOG >= 13 ? 3 : HLOG == 12 ? 6 : 8)) lz4_compress_kernel(
    const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off, const uint64_t* __restrict__ src_len,
    const uint32_t* __restrict__ blk_base, uint32_t n_streams, uint32_t n_blocks, uint32_t block_size,
    uint8_t* __restrict__ scratch, uint32_t* __restrict__ csize, uint64_t* __restrict__ sizes,
    unsigned int* __restrict__ work_counter) {
  extern __shared__ __align__(16) uint16_t smem_tables[];
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  uint16_t* table = smem_tables + (size_t)(threadIdx.x >> 5) * (1 << HLOG);

  for (;;) {
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1u);
    b = __shfl_sync(FULL, b, 0);
    if (b >= n_blocks) break;

    const uint32_t si = find_stream32(blk_base, n_streams, b);
    const uint64_t boff = (uint64_t)(b - blk_base[si]) * block_size;
    const uint64_t rem = src_len[si] - boff;
    const int n = (int)(rem < block_size ? rem : block_size);
    const uint8_t* __restrict__ s = src_base + src_off[si] + boff;
    uint8_t* __restrict__ out = scratch + (uint64_t)b * block_size;
    const int cap = n - 1;  // LZ4BlockOutputStream stores RAW when compressedLength >= originalLength

    {
      uint4* t4 = reinterpret_cast<uint4*>(table);
      for (int j = lane; j < (1 << HLOG) / 8; j += 32) t4[j] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();

    int op = 0, anchor = 0, pos = 0;
    bool fail = false;
    if (n >= kMFLimit + 1) {
      const int mflimit = n - kMFLimit;
      const int matchlimit = n - kLastLiterals;
      while (pos <= mflimit) {
        // ---- 1. lookup + local match length (<= 8) for all 32 positions.  Lane l owns window position r = 31 - l:
        //         the hardware resolves same-address shared stores in favour of the lowest lane, so with this
        //         mapping a contested hash slot receives the highest position on the first store (step 4).
        const int r_me = 31 - lane;
        const int p = pos + r_me;
        const bool valid = p <= mflimit;
        uint32_t v = 0, v2 = 0, h = 0;
        int cand = 0, ml = 0;
        bool ok = false;
        if (valid) {
          const uintptr_t a = reinterpret_cast<uintptr_t>(s + p);
          const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
          const unsigned sh = (a & 3u) * 8u;
          const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = sh ? __ldg(w + 2) : 0u;
          v = __funnelshift_r(w0, w1, sh);
          v2 = __funnelshift_r(w1, w2, sh);
          h = (v * 2654435761u) >> (32 - HLOG);
          cand = table[h];
          const uint32_t prev = p > 0 ? (uint32_t)__ldg(s + p - 1) : (~v & 0xffu);
          uint32_t x = 1;
          if (v == prev * 0x01010101u) {  // byte run: offset-1 candidate (what a sequential hash table would hold)
            ok = true;
            cand = p - 1;
            x = __funnelshift_r(v, v2, 24) ^ v2;  // bytes p+3..p+6 against p+4..p+7
          } else if (cand < p) {
            const uintptr_t ca = reinterpret_cast<uintptr_t>(s + cand);
            const uint32_t* cw = reinterpret_cast<const uint32_t*>(ca & ~uintptr_t(3));
            const unsigned csh = (ca & 3u) * 8u;
            const uint32_t c0 = __ldg(cw), c1 = __ldg(cw + 1);
            if (__funnelshift_r(c0, c1, csh) == v) {
              ok = true;
              const uint32_t c2 = csh ? __ldg(cw + 2) : 0u;
              x = __funnelshift_r(c1, c2, csh) ^ v2;
            }
          }
          if (ok) {
            ml = x ? 4 + ((__ffs(x) - 1) >> 3) : 8;
            const int lim = matchlimit - p;
            if (ml > lim) ml = lim;
          }
        }
        const unsigned posmask = __brev(__ballot_sync(FULL, ok));  // bit r <-> window position r

        // ---- 2. greedy walk over the window: uniform bit operations, ~a dozen instructions per selected sequence
        int cur = 0;          // window-relative parse cursor
        unsigned selmask = 0; // positions whose match the parse takes
        for (;;) {
          const unsigned t = posmask & (FULL << cur);
          if (!t) break;
          const int r = __ffs(t) - 1;
          int mlr = __shfl_sync(FULL, ml, 31 - r);
          if (mlr == 8) {  // local 8 bytes all equal: extend cooperatively, 32 bytes per ballot
            const int m = pos + r;
            const int maxl = matchlimit - m;
            if (maxl > 8) {
              const int c = __shfl_sync(FULL, cand, 31 - r);
              int k = 8 + lane;
              for (;;) {
                const bool ne = (k >= maxl) || (__ldg(s + m + k) != __ldg(s + c + k));
                const unsigned nb = __ballot_sync(FULL, ne);
                if (nb) {
                  mlr = k - lane + __ffs(nb) - 1;
                  break;
                }
                k += 32;
              }
              if (r == r_me) ml = mlr;
            }
          }
          selmask |= 1u << r;
          cur = r + mlr;
          if (cur >= 32) break;
        }

        // ---- 3. sizes and output offsets for all selected sequences at once (suffix sum over lanes = prefix over positions)
        const bool sel = (selmask >> r_me) & 1u;
        const unsigned below = selmask & ((1u << r_me) - 1u);  // selected positions before mine
        const int rp = 31 - __clz(below);                      // -1 when none
        const int ml_prev = __shfl_sync(FULL, ml, below ? 31 - rp : lane);
        const int lit = p - (below ? pos + rp + ml_prev : anchor);
        const int mlc = ml - kMinMatch;
        int sz = 0;
        if (sel) {
          sz = 3 + lit;
          if (lit >= 15) sz += (lit - 15) / 255 + 1;
          if (mlc >= 15) sz += (mlc - 15) / 255 + 1;
        }
        int suf = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_down_sync(FULL, suf, d);
          if (lane + d < 32) suf += t;
        }
        const int total = __shfl_sync(FULL, suf, 0);
        if (op + total > cap) {
          fail = true;
          break;
        }
        const int my_op = op + suf - sz;  // sequences at earlier positions (higher lanes) come first

        // ---- 4. emit: every selected lane writes its own sequence; only very long literal runs / lengths go cooperative
        bool slow = false;
        if (sel) {
          if (lit <= 16 && mlc < 15 + 510) {
            uint8_t* q = out + my_op;
            *q++ = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (mlc < 15 ? mlc : 15));
            if (lit >= 15) *q++ = (uint8_t)(lit - 15);
            const uint8_t* ls = s + p - lit;
            for (int j = 0; j < lit; j++) q[j] = __ldg(ls + j);
            q += lit;
            const int off = p - cand;
            q[0] = (uint8_t)off;
            q[1] = (uint8_t)(off >> 8);
            if (mlc >= 15) {
              int rem = mlc - 15;
              q += 2;
              if (rem >= 255) {
                *q++ = 255;
                rem -= 255;
              }
              *q = (uint8_t)rem;
            }
          } else {
            slow = true;
          }
        }
        unsigned slowmask = __ballot_sync(FULL, slow);
        while (slowmask) {
          const int l = __ffs(slowmask) - 1;
          slowmask &= slowmask - 1;
          const int lit_r = __shfl_sync(FULL, lit, l);
          const int ml_r = __shfl_sync(FULL, ml, l);
          const int op_r = __shfl_sync(FULL, my_op, l);
          const int cand_r = __shfl_sync(FULL, cand, l);
          const int m = pos + 31 - l;
          warp_emit_seq(out + op_r, s + m - lit_r, lit_r, m - cand_r, ml_r, lane);
        }
        op += total;
        if (selmask) anchor = pos + cur;

        // ---- 5. insert the window; a read-back settles slots hit twice so that the highest position wins
        __syncwarp();
        if (valid) table[h] = (uint16_t)p;
        __syncwarp();
        for (;;) {
          const bool lost = valid && table[h] < (uint16_t)p;
          if (!__ballot_sync(FULL, lost)) break;
          if (lost) table[h] = (uint16_t)p;
          __syncwarp();
        }
        pos += cur > 32 ? cur : 32;
      }
    }
    if (!fail) {
      const int lit = n - anchor;
      const int need = 1 + lit + (lit >= 15 ? (lit - 15) / 255 + 1 : 0);
      if (op + need > cap) {
        fail = true;
      } else {
        warp_emit_seq(out + op, s + anchor, lit, 0, 0, lane);
        op += need;
      }
    }
    if (lane == 0) {
      csize[b] = fail ? ((uint32_t)n | 0x80000000u) : (uint32_t)op;
      sizes[b] = 21u + (uint64_t)(fail ? n : op);
    }
    __syncwarp();
  }
}

template <int HLOG>
static void launch_lz4_compress_t(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len,
                                  const uint32_t* d_blk_base, uint32_t n_streams, uint32_t n_blocks,
                                  uint32_t block_size, uint8_t* d_scratch, uint32_t* d_csize, uint64_t* d_sizes,
                                  unsigned int* d_counter, cudaStream_t st) {
  constexpr int kWarps = kLz4Threads / 32;
  const size_t smem = (size_t)kWarps * (2u << HLOG);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(lz4_compress_kernel<HLOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lz4_compress_kernel<HLOG>, kLz4Threads, smem);
  if (per_sm < 1) per_sm = 1;
  uint64_t want = ((uint64_t)n_blocks + kWarps - 1) / kWarps;
  uint64_t grid = (uint64_t)kSMs * per_sm;
  if (grid > want) grid = want;
  lz4_compress_kernel<HLOG><<<(unsigned)grid, kLz4Threads, smem, st>>>(src_base, d_src_off, d_src_len, d_blk_base,
                                                                      n_streams, n_blocks, block_size, d_scratch,
                                                                      d_csize, d_sizes, d_counter);
}

int g_lz4_hlog = 12;   // tuning knobs (api.cu reads B2S_LZ4_HLOG / B2S_LZ4D_TILE once at init); 12 is the specified default
int g_lz4d_tile = 8;

void launch_lz4_compress(const uint8_t* src_base, const uint64_t* d_src_off, const uint64_t* d_src_len,
                         const uint32_t* d_blk_base, uint32_t n_streams, uint32_t n_blocks, uint32_t block_size,
                         uint8_t* d_scratch, uint32_t* d_csize, uint64_t* d_sizes, unsigned int* d_counter,
                         cudaStream_t st, uint64_t* launches) {
  if (!n_blocks) return;
  cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), st);
#define B2S_LZ4C(H)                                                                                                  \
  launch_lz4_compress_t<H>(src_base, d_src_off, d_src_len, d_blk_base, n_streams, n_blocks, block_size, d_scratch,  \
                           d_csize, d_sizes, d_counter, st)
  switch (g_lz4_hlog) {
    case 10: B2S_LZ4C(10); break;
    case 11: B2S_LZ4C(11); break;
    case 13: B2S_LZ4C(13); break;
    default: B2S_LZ4C(12); break;
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
//
// A warp carries 32/TILE independent blocks ("tiles" of TILE lanes) through ONE warp-uniform instruction stream:
// every iteration decodes one sequence per tile (token, literal run, offset, match), with the copy loops bounded by
// the warp-wide maximum (REDUX.MAX) and predicated per tile.  All synchronisation is full-mask __syncwarp().
// (Per-tile control flow with partial-mask sync looked natural but ptxas guards every non-uniform-mask
// __syncwarp/__ballot with MATCH.ANY — measured at ~1 per 50 cycles per SM on B200, profiles/ — so it is avoided.)
// Sequences in shuffle data are short (terasort: 0.8 literal + 6.7 match bytes), hence small tiles.
// ------------------------------------------------------------------------------------------------------------
template <int TILE>
__global__ void __launch_bounds__(kLz4Threads) lz4_decompress_kernel(const BlockDesc* __restrict__ desc,
                                                                     uint32_t n_blocks,
                                                                     const uint8_t* __restrict__ src_base,
                                                                     uint8_t* __restrict__ dst_base,
                                                                     int32_t* __restrict__ status,
                                                                     unsigned int* __restrict__ work_counter) {
  constexpr unsigned FULL = 0xffffffffu;
  const int tl = threadIdx.x % TILE;  // lane within the tile
  bool active = false, exhausted = false;
  const uint8_t* __restrict__ in = nullptr;
  uint8_t* out = nullptr;
  int ip = 0, op = 0, clen = 0, olen = 0;
  uint32_t stream = 0;

  for (;;) {
    // ---- refill tiles that have no block (rare: once per codec block)
    if (!active && !exhausted) {
      uint32_t b = 0;
      if (tl == 0) b = atomicAdd(work_counter, 1u);
      b = __shfl_sync(tile_mask<TILE>(), b, 0, TILE);
      if (b >= n_blocks) {
        exhausted = true;
      } else {
        const BlockDesc d = desc[b];
        if (d.olen != 0) {
          in = src_base + d.src;
          out = dst_base + d.dst;
          if (d.stream & 0x80000000u) {  // stored RAW
            group_copy<TILE>(out, in, d.olen, tl);
          } else {
            active = true;
            ip = op = 0;
            clen = (int)d.clen;
            olen = (int)d.olen;
            stream = d.stream;
          }
        }
      }
    }
    if (__all_sync(FULL, exhausted && !active)) break;

    // ---- one sequence per active tile, warp-uniform control flow
    bool err = false;
    int token = 0, ll = 0;
    if (active) {
      if (ip >= clen) err = true;
      else {
        token = __ldg(in + ip++);
        ll = token >> 4;
      }
    }
    bool more = active && !err && ll == 15;
    while (__any_sync(FULL, more)) {
      if (more) {
        if (ip >= clen) {
          err = true;
          more = false;
        } else {
          const int bb = __ldg(in + ip++);
          ll += bb;
          more = (bb == 255);
        }
      }
    }
    if (active && !err && (ll > olen - op || ll > clen - ip)) err = true;
    {
      const int cl = (active && !err) ? ll : 0;
      const int maxl = __reduce_max_sync(FULL, cl);
      for (int j = tl; j < maxl; j += TILE)
        if (j < cl) out[op + j] = __ldg(in + ip + j);
    }
    if (active && !err) {
      op += ll;
      ip += ll;
    }
    const bool last = active && !err && (olen - op < kMFLimit);  // last match must start >= 12 bytes before the end
    if (last && (op != olen || ip != clen)) err = true;
    const bool cont = active && !err && !last;
    int off = 0, ml = 0;
    if (cont) {
      if (ip + 2 > clen) err = true;
      else {
        off = __ldg(in + ip) | (__ldg(in + ip + 1) << 8);
        ip += 2;
        ml = token & 15;
      }
    }
    more = cont && !err && ml == 15;
    while (__any_sync(FULL, more)) {
      if (more) {
        if (ip >= clen) {
          err = true;
          more = false;
        } else {
          const int bb = __ldg(in + ip++);
          ml += bb;
          more = (bb == 255);
        }
      }
    }
    if (cont && !err) {
      ml += kMinMatch;
      if (ml > olen - op || off == 0 || off > op) err = true;
    }
    __syncwarp();  // literals of this sequence are visible to the whole warp
    {
      const int cm = (cont && !err) ? ml : 0;
      const int maxm = __reduce_max_sync(FULL, cm);
      const uint8_t* msrc = out + op - off;
      for (int j = tl; j < maxm; j += TILE) {
        if (j < cm) {
          // overlapping match (off < ml): every byte comes from the already complete window [op-off, op)
          const int k = (j >= off) ? (int)((unsigned)j % (unsigned)off) : j;
          out[op + j] = msrc[k];
        }
      }
    }
    if (cont && !err) {
      op += ml;
      if (olen - op < kLastLiterals) err = true;  // the last 5 bytes of a block are literals
    }
    __syncwarp();
    if (active && (err || last)) {
      if (err && tl == 0) set_status(status, stream & 0x7fffffffu, B2S_E_CORRUPT);
      active = false;
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
    case 4: launch_lz4_decompress_t<4>(d_desc, n_blocks, src_base, dst_base, d_status, d_counter, st); break;
    case 16: launch_lz4_decompress_t<16>(d_desc, n_blocks, src_base, dst_base, d_status, d_counter, st); break;
    case 32: launch_lz4_decompress_t<32>(d_desc, n_blocks, src_base, dst_base, d_status, d_counter, st); break;
    default: launch_lz4_decompress_t<8>(d_desc, n_blocks, src_base, dst_base, d_status, d_counter, st); break;
  }
  *launches += 1;
}

}  // namespace b2s
