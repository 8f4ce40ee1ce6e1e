// lz4_parse_core.h — the lane-local part of the sub-chunk parallel parse (B2S_LZ4_PIPE=4), written once for host and
// device.
//
// Replaces the sequence selection of liblz4's LZ4_compress_default / snappy's RawCompress / (with CODEC 2) the match
// selection in front of the Zstandard sequence coder, as driven by the compressed output streams Spark puts on
// shuffle/S3ShuffleMapOutputWriter.scala:140-146.  Specification: orc_lz4_compress_block_win_sub /
// orc_snappy_compress_raw_win_sub in oracle/ (phase B with `sub` > 0).
//
// nvcc compiles seq_size() into lz4_parse4_kernel (lz4_compress.cu, one WARP per codec block, lane per sub-chunk — the
// product when that generation is selected); g++ compiles walk_subchunk() into tests/native/lz4_parse_host.cpp, where
// tests/test_parse_core.py runs 32 lanes + the kernel's stitch against the oracle without a GPU.  The kernel's own loop
// is the same walk restructured for lock-step lanes (one sequence per lane and trip).  Nothing here is a CPU fallback:
// the C ABI only launches the device build.
//
// Input per position: off[p] (u16, 0 = no match; for blocks <= 32 KiB bit 15 = "exactly 4 long", which spares the
// measurement for the commonest sequences) and one 32-bit mask per window (bit r = position r matched), both written by
// lz4_match2_kernel.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2S_PHD __host__ __device__ __forceinline__
#else
#define B2S_PHD inline
struct uint2 {
  uint32_t x, y;
};
struct uint4 {
  uint32_t x, y, z, w;
};
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
#endif

namespace b2s {
namespace lzparse {

constexpr int kMinMatch = 4, kMFLimit = 12, kLastLiterals = 5;

struct Result {
  uint32_t nseq;   // records written
  uint32_t csize;  // LZ4: payload bytes, bit 31 = store RAW; Snappy: raw block bytes; Zstandard: unused
  uint64_t size;   // framed bytes of the block (LZ4Block: 21 + payload; xerial: 4 + payload)
};

B2S_PHD uint32_t funnel_r(uint32_t lo, uint32_t hi, unsigned sh) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
B2S_PHD int first_set(uint32_t x) {  // index of the lowest set bit; x != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// Sub-chunk parallel parse (B2S_LZ4_PIPE=4).  A thread-per-block walk that measures match lengths itself is bound by the
// LATENCY of one thread's chain over 32 KiB (two variants were built and measured, profiles/r2_compress_generations.md:
// the event-driven one executed 40 % fewer instructions than the fixed-trip one and was slower, 9.5 ms per 32,768 blocks
// at 8 % issue utilisation — every iteration ends in dependent loads).  This generation puts the parallelism INSIDE the
// block: the block is cut into 32 sub-chunks of S = stride / 32
// positions (1 KiB for 32 KiB blocks), lane k of a WARP parses sub-chunk k greedily from its first position, and a match
// neither starts in the last three positions of a sub-chunk nor extends past its end.  That is a (slightly) different
// compressor — at most one cut match per KiB, ~0.5 % more output on the terasort shape — so the executable
// specification in oracle/ states the same rule (orc_lz4_compress_block_win, `sub`), and the bytes still have to agree.
// The warp then stitches the 32 record lists together (lz4_parse4_kernel: literal runs that span sub-chunks, output
// offsets by warp scans, in-place compaction of the staged records); 32,768 warps instead of 1,024 per GiB hide the
// latency this walk cannot avoid.
//
// walk_subchunk(): lane-local part, host/device.  Records are staged at `stage[0..]` in the final format except that
// offsets are relative to the sub-chunk's first record and record 0 carries its match position instead of an anchor:
//   record 0 : x = pm0                    y = len0
//   record i : x = anchor | lit << 16     y = len | rel_op << 16    (rel_op = sum of the sizes of records 1..i-1)
struct SubResult {
  int nrec;       // records staged
  int rest;       // sum of the sizes of records 1..nrec-1 (CODEC 2: their literal bytes)
  int pm0, len0, d0;  // first match of the sub-chunk (valid if nrec > 0)
  int last_end;   // end of the last match, -1 if none
};

template <int CODEC>
B2S_PHD int seq_size(int lit, int len, int d) {
  constexpr bool SNAPPY = CODEC == 1, ZSTD = CODEC == 2;
  if (ZSTD) return lit;
  if (SNAPPY) {
    int size = lit ? lit + (lit - 1 < 60 ? 1 : lit - 1 < 256 ? 2 : 3) : 0;
    int l = len;
    if (l >= 68) {
      const int k = (l - 68) / 64 + 1;
      size += 3 * k;
      l -= 64 * k;
    }
    if (l > 64) {
      size += 3;
      l -= 60;
    }
    return size + ((l < 12 && d < 2048) ? 2 : 3);
  }
  const int mlc = len - kMinMatch;
  int size = 3 + lit;
  if (lit >= 15) size += (lit - 15) / 255 + 1;
  if (mlc >= 15) size += (mlc - 15) / 255 + 1;
  return size;
}

template <int CODEC, class Mem>
B2S_PHD SubResult walk_subchunk(const Mem& mem, int n, int sb, uint32_t stride, int lo, int hi, uint2* stage) {
  const bool flag4 = stride <= 32768u;
  const uint32_t omask = flag4 ? 0x7fffu : 0xffffu;
  SubResult R;
  R.nrec = 0;
  R.rest = 0;
  R.pm0 = R.len0 = R.d0 = 0;
  R.last_end = -1;
  const int mflimit = n - kMFLimit, matchlimit = n - kLastLiterals;
  const int plim = mflimit < hi - 4 ? mflimit : hi - 4;       // last position a match may start at
  const int elim = matchlimit < hi ? matchlimit : hi;         // a match ends at or before this position
  if (lo > plim) return R;
  const int last_w = plim >> 5;
  int wi = lo >> 5;
  uint32_t mw = mem.mask(wi);
  int pm = -1, ed = 0, e = 0, anchor = lo, rel = 0;
  bool x4 = false, done = false;
  uint2* sp = stage;
  while (!(done && pm < 0)) {
    if (pm < 0) {
      if (mw == 0u) {
        wi++;
        if (wi > last_w) done = true;
        else mw = mem.mask(wi);
      } else {
        const int q = (wi << 5) + first_set(mw);
        if (q > plim) {
          done = true;
        } else {
          pm = q;
          const uint32_t o16 = mem.off16(q);
          ed = (int)(o16 & omask);
          e = q + 4;
          x4 = (flag4 && (o16 & 0x8000u)) || e >= elim;  // nothing (more) to compare
        }
      }
    }
    if (pm >= 0) {
      int eq = 0;
      bool end = true;
      if (!x4) {
        const int ci = sb + e, cj = ci - ed;
        const unsigned sha = (unsigned)(ci & 3) * 8u, shc = (unsigned)(cj & 3) * 8u;
        const int ka = ci >> 2, kc = cj >> 2;
        const uint32_t a0 = mem.cand_word(ka), a1 = mem.word(ka + 1), a2 = mem.word(ka + 2);
        const uint32_t c0 = mem.cand_word(kc), c1 = mem.cand_word(kc + 1), c2 = mem.cand_word(kc + 2);
        const uint32_t xl = funnel_r(a0, a1, sha) ^ funnel_r(c0, c1, shc);
        const uint32_t xh = funnel_r(a1, a2, sha) ^ funnel_r(c1, c2, shc);
        eq = xl ? (first_set(xl) >> 3) : xh ? 4 + (first_set(xh) >> 3) : 8;
        const int room = elim - e;  // > 0
        eq = eq < room ? eq : room;
        end = eq < 8 || eq == room;
      }
      if (end) {
        const int len = e + eq - pm;
        if (sp == stage) {
          R.pm0 = pm;
          R.len0 = len;
          R.d0 = ed;
          *sp++ = make_uint2((uint32_t)pm, (uint32_t)len);
        } else {
          const int lit = pm - anchor;
          *sp++ = make_uint2((uint32_t)anchor | ((uint32_t)lit << 16), (uint32_t)len | ((uint32_t)rel << 16));
          rel += seq_size<CODEC>(lit, len, ed);
        }
        const int p = pm + len;
        anchor = p;
        pm = -1;
        const int nw = p >> 5;
        if (nw != wi) {
          wi = nw;
          if (wi > last_w) {
            done = true;
            mw = 0u;
          } else {
            mw = mem.mask(wi);
          }
        }
        mw &= ~0u << (p & 31);
      } else {
        e += 8;
      }
    }
  }
  R.nrec = (int)(sp - stage);
  R.rest = rel;
  R.last_end = R.nrec ? anchor : -1;
  return R;
}

}  // namespace lzparse
}  // namespace b2s
