// lz4_parse_core.h — the greedy parse of one codec block (K3-B, second generation), written once for host and device.
//
// Replaces the sequence selection of liblz4's LZ4_compress_default / snappy's RawCompress / (with CODEC 2) the match
// selection in front of the Zstandard sequence coder, as driven by the compressed output streams Spark puts on
// shuffle/S3ShuffleMapOutputWriter.scala:140-146.  Specification: orc_lz4_compress_block_win / orc_snappy_compress_raw_win
// in oracle/ (phase B): lowest p >= cursor with off[p] != 0, extended on the source up to matchlimit.
//
// nvcc compiles parse_block() into lz4_parse2_kernel (lz4_compress.cu, one THREAD per codec block — the product); g++
// compiles the same function into tests/native/lz4_parse_host.cpp, where tests/test_parse_core.py checks it against the
// oracle without a GPU.  Nothing here is a CPU fallback: the C ABI only launches the device build.
//
// Shape of the loop (why it looks the way it does): 32 lanes = 32 different blocks execute this in lock step, so the
// walk is a FIXED-TRIP loop over groups of FOUR positions.  A match is >= 4 long, hence at most one sequence starts
// per group and its extension (bytes p+4 ...) begins in the NEXT group: every group iteration is "finish or continue
// the open match on this group's four source bytes, then maybe take one match", the same straight-line code for all
// lanes.  Input per position: off[p] (u16, 0 = no match; for blocks <= 32 KiB bit 15 = "exactly 4 long", which spares
// the extension and its dependent candidate load for the commonest sequences).  Source words and off[] vectors come
// through a register ring (the loads of trip t+1 are issued at the top of trip t); the candidate words of the next
// group are requested one iteration ahead.
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

// record: x = literal start | literal count << 16 ; y = match length (0 = final literal run) | output offset << 16
// (CODEC 2: the running literal count instead of an output offset).  CODEC: 0 LZ4, 1 Snappy, 2 Zstandard.
// Mem: off8(i) = vector i (8 positions) of the block's off[] row, word(k) = aligned source word k (block byte q is
// stream byte sb + q), cand_word(k) = the same stream for candidate reads (never past the word of the current byte).
template <int CODEC, class Mem>
B2S_PHD Result parse_block(const Mem& mem, int n, int sb, uint32_t stride, uint2* seq) {
  constexpr bool SNAPPY = CODEC == 1, ZSTD = CODEC == 2;
  const bool flag4 = stride <= 32768u;
  const uint32_t omask = flag4 ? 0x7fffu : 0xffffu;
  const int cap = (SNAPPY || ZSTD) ? 0x7fffffff : n - 1;  // LZ4BlockOutputStream stores RAW when compressedLength >= originalLength
  int p = 0, anchor = 0, op = SNAPPY ? (n < 128 ? 1 : n < 16384 ? 2 : 3) : 0;  // Snappy: varint(n) comes first
  uint32_t ns = 0;
  bool fail = false;
  const int mflimit = n - kMFLimit, matchlimit = n - kLastLiterals;

  // one sequence: literals [anchor, pm), match (len, d)
  auto emit = [&](int pm, int len, int d) {
    const int lit = pm - anchor;
    int size;
    if (ZSTD) {
      size = lit;  // "op" counts literal bytes: where this sequence's literals go in the literals section
    } else if (SNAPPY) {
      size = lit ? lit + (lit - 1 < 60 ? 1 : lit - 1 < 256 ? 2 : 3) : 0;
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
      size += (l < 12 && d < 2048) ? 2 : 3;
    } else {
      const int mlc = len - kMinMatch;
      size = 3 + lit;
      if (lit >= 15) size += (lit - 15) / 255 + 1;
      if (mlc >= 15) size += (mlc - 15) / 255 + 1;
    }
    if (op + size > cap) {
      fail = true;
    } else {
      seq[ns++] = make_uint2((uint32_t)anchor | ((uint32_t)lit << 16), (uint32_t)len | ((uint32_t)op << 16));
      op += size;
    }
    p = anchor = pm + len;
  };

  if (n >= kMFLimit + 1) {  // shorter blocks hold no match; the match kernel skipped them (nothing was written to off[])
    const unsigned sh = (unsigned)sb * 8u;
    const int trips = (matchlimit + 15) >> 4;  // groups up to the one holding byte matchlimit - 1

    // open match: starts at pm, offset ed, e = next byte to compare (lies in the current group), x4 = flagged "exactly
    // 4 long" (nothing to compare: it is closed in the group that follows its start); cn0/cn1 = the candidate words of
    // the current group, requested one iteration ago.  Sequences are emitted at ONE place — when an open match is
    // closed — because the 32 lanes of a warp sit in different phases and every divergent path is paid by all of them.
    bool ext = false, x4 = false;
    int pm = 0, ed = 0, e = 0;
    uint32_t cn0 = 0, cn1 = 0;
    unsigned cnsh = 0;
    auto request_cand = [&](int base) {  // words holding bytes base - ed .. base - ed + 3
      const int ci = sb + base - ed;     // >= sb + 1
      cnsh = (unsigned)(ci & 3) * 8u;
      cn0 = mem.cand_word(ci >> 2);
      cn1 = mem.cand_word((ci >> 2) + 1);  // at most the word of byte base + 2
    };

    // one group: o_lo/o_hi = its four off entries, sw = its four source bytes
    auto body = [&](uint32_t o_lo, uint32_t o_hi, uint32_t sw, int base) {
      if (fail) return;
      if (ext) {
        const int k0 = e - base;  // 0..3
        int eq = 0;
        bool end = true;
        if (!x4) {
          const uint32_t cw = funnel_r(cn0, cn1, cnsh);
          const uint32_t x = (sw ^ cw) >> (8 * k0);
          eq = x ? (first_set(x) >> 3) : 4 - k0;  // equal bytes from e on, within this group
          const int room = matchlimit - e;         // > 0
          eq = eq < room ? eq : room;
          end = eq < 4 - k0 || eq == room;
        }
        if (end) {
          ext = false;
          emit(pm, e + eq - pm, ed);
        } else {
          e = base + 4;
          request_cand(base + 4);
        }
      }
      if (!ext && !fail && p < base + 4 && base <= mflimit) {
        const int rel = p > base ? p - base : 0;
        const unsigned nz = ((o_lo & 0xffffu) ? 1u : 0u) | ((o_lo >> 16) ? 2u : 0u) | ((o_hi & 0xffffu) ? 4u : 0u) |
                            ((o_hi >> 16) ? 8u : 0u);
        const unsigned mm = nz & (0xfu << rel);
        if (mm) {
          const int j = first_set(mm);
          const uint32_t w = j < 2 ? o_lo : o_hi;
          const uint32_t o16 = (j & 1) ? w >> 16 : w & 0xffffu;
          ext = true;
          pm = base + j;
          ed = (int)(o16 & omask);
          e = pm + 4;  // lies in the next group
          x4 = flag4 && (o16 & 0x8000u);
          if (!x4) request_cand(base + 4);
        } else {
          p = base + 4;
        }
      }
    };

    uint4 oa = mem.off8(0), ob = mem.off8(1);
    uint32_t wc = mem.word(0);
    uint32_t w1 = mem.word(1), w2 = mem.word(2), w3 = mem.word(3), w4 = mem.word(4);
    for (int t = 0; t < trips && !fail; t++) {
      const uint4 a = oa, bq = ob;
      const uint32_t x0 = wc, x1 = w1, x2 = w2, x3 = w3, x4 = w4;
      oa = mem.off8(2 * t + 2);
      ob = mem.off8(2 * t + 3);
      wc = x4;
      w1 = mem.word(4 * t + 5);
      w2 = mem.word(4 * t + 6);
      w3 = mem.word(4 * t + 7);
      w4 = mem.word(4 * t + 8);
      const int base = 16 * t;
      body(a.x, a.y, funnel_r(x0, x1, sh), base);
      body(a.z, a.w, funnel_r(x1, x2, sh), base + 4);
      body(bq.x, bq.y, funnel_r(x2, x3, sh), base + 8);
      body(bq.z, bq.w, funnel_r(x3, x4, sh), base + 12);
    }
    // (an open match is always closed inside the loop: eq == room in the group holding byte matchlimit - 1)
  }
  Result r;
  if (ZSTD) {  // trailing literals as a final ml == 0 record; sizes are decided by the entropy stage
    seq[ns++] = make_uint2((uint32_t)anchor | ((uint32_t)(n - anchor) << 16), (uint32_t)op << 16);
    r.nseq = ns;
    r.csize = 0;
    r.size = 0;
    return r;
  }
  if (SNAPPY) {
    const int lit = n - anchor;
    if (lit) {
      seq[ns++] = make_uint2((uint32_t)anchor | ((uint32_t)lit << 16), (uint32_t)op << 16);
      op += lit + (lit - 1 < 60 ? 1 : lit - 1 < 256 ? 2 : 3);
    }
    r.nseq = ns;
    r.csize = (uint32_t)op;
    r.size = 4u + (uint64_t)op;  // BE32 chunk length + raw snappy block
    return r;
  }
  if (!fail) {
    const int lit = n - anchor;
    const int need = 1 + lit + (lit >= 15 ? (lit - 15) / 255 + 1 : 0);
    if (op + need > cap) {
      fail = true;
    } else {
      seq[ns++] = make_uint2((uint32_t)anchor | ((uint32_t)lit << 16), (uint32_t)op << 16);
      op += need;
    }
  }
  r.nseq = ns;
  r.csize = fail ? ((uint32_t)n | 0x80000000u) : (uint32_t)op;
  r.size = 21u + (uint64_t)(fail ? n : op);
  return r;
}

// ---------------------------------------------------------------------------------------------------------------
// Event-driven variant (B2S_LZ4_PIPE=3).  The fixed-trip walk above visits every group of four positions (8192 per
// 32 KiB block) and, because the 32 lanes of a warp sit in different phases, pays every path of its body each time
// (ncu: ~130 warp-instructions per group).  This variant iterates per EVENT instead: the match kernel also stores one
// 32-bit word per window with a bit per matched position, so "next match at or after the cursor" is a find-first-set,
// and an iteration is "take the next match (one off[] gather) and/or compare eight more bytes of the open match, emit
// if it closed".  A sequence flagged "exactly 4" costs one iteration, a longer one 1 + (len - 4) / 8; empty windows
// cost one cheap iteration each.  Same sequences, same records as parse_block().
// Mem additionally provides mask(w) (window w's match bits), off16(q) and word(k) without the ring.
template <int CODEC, class Mem>
B2S_PHD Result parse_block_ev(const Mem& mem, int n, int sb, uint32_t stride, uint2* seq) {
  constexpr bool SNAPPY = CODEC == 1, ZSTD = CODEC == 2;
  const bool flag4 = stride <= 32768u;
  const uint32_t omask = flag4 ? 0x7fffu : 0xffffu;
  int anchor = 0, op = SNAPPY ? (n < 128 ? 1 : n < 16384 ? 2 : 3) : 0;
  uint2* sp = seq;  // next record
  const int mflimit = n - kMFLimit, matchlimit = n - kLastLiterals;
  // LZ4: the RAW decision (compressedLength >= originalLength) is taken once at the end — op only grows, and a block
  // that fails it is re-emitted from the source, its records are ignored (op may then exceed the record's 16 bits)

  if (n >= kMFLimit + 1) {
    const int last_w = mflimit >> 5;
    int wi = 0;
    uint32_t mw = mem.mask(0);
    // pm < 0: searching; otherwise the open match starts at pm, offset ed, next byte to compare e (x4: nothing to compare)
    int pm = -1, ed = 0, e = 0;
    bool x4 = false, done = false;
    while (!(done && pm < 0)) {
      if (pm < 0) {
        if (mw == 0u) {
          wi++;
          if (wi > last_w) done = true;
          else mw = mem.mask(wi);
        } else {
          pm = (wi << 5) + first_set(mw);
          const uint32_t o16 = mem.off16(pm);
          ed = (int)(o16 & omask);
          e = pm + 4;
          x4 = flag4 && (o16 & 0x8000u);
        }
      }
      if (pm >= 0) {
        int eq = 0;
        bool end = true;
        if (!x4) {
          // eight bytes at e against the eight at e - ed: three aligned words each; only the source side can run past
          // the block's last word (clamped; the surplus bytes are cut off by room)
          const int ci = sb + e, cj = ci - ed;
          const unsigned sha = (unsigned)(ci & 3) * 8u, shc = (unsigned)(cj & 3) * 8u;
          const int ka = ci >> 2, kc = cj >> 2;
          const uint32_t a0 = mem.cand_word(ka), a1 = mem.word(ka + 1), a2 = mem.word(ka + 2);
          const uint32_t c0 = mem.cand_word(kc), c1 = mem.cand_word(kc + 1), c2 = mem.cand_word(kc + 2);
          const uint32_t xl = funnel_r(a0, a1, sha) ^ funnel_r(c0, c1, shc);
          const uint32_t xh = funnel_r(a1, a2, sha) ^ funnel_r(c1, c2, shc);
          eq = xl ? (first_set(xl) >> 3) : xh ? 4 + (first_set(xh) >> 3) : 8;
          const int room = matchlimit - e;  // > 0
          eq = eq < room ? eq : room;
          end = eq < 8 || eq == room;
        }
        if (end) {
          const int len = e + eq - pm;
          const int lit = pm - anchor;
          int size;
          if (ZSTD) {
            size = lit;
          } else if (SNAPPY) {
            size = lit ? lit + (lit - 1 < 60 ? 1 : lit - 1 < 256 ? 2 : 3) : 0;
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
            size += (l < 12 && ed < 2048) ? 2 : 3;
          } else {
            const int mlc = len - kMinMatch;
            size = 3 + lit;
            if (lit >= 15) size += (lit - 15) / 255 + 1;
            if (mlc >= 15) size += (mlc - 15) / 255 + 1;
          }
          *sp++ = make_uint2((uint32_t)anchor | ((uint32_t)lit << 16), (uint32_t)len | ((uint32_t)op << 16));
          op += size;
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
  }
  Result r;
  if (ZSTD) {
    *sp++ = make_uint2((uint32_t)anchor | ((uint32_t)(n - anchor) << 16), (uint32_t)op << 16);
    r.nseq = (uint32_t)(sp - seq);
    r.csize = 0;
    r.size = 0;
    return r;
  }
  if (SNAPPY) {
    const int lit = n - anchor;
    if (lit) {
      *sp++ = make_uint2((uint32_t)anchor | ((uint32_t)lit << 16), (uint32_t)op << 16);
      op += lit + (lit - 1 < 60 ? 1 : lit - 1 < 256 ? 2 : 3);
    }
    r.nseq = (uint32_t)(sp - seq);
    r.csize = (uint32_t)op;
    r.size = 4u + (uint64_t)op;
    return r;
  }
  {
    const int lit = n - anchor;
    *sp++ = make_uint2((uint32_t)anchor | ((uint32_t)lit << 16), (uint32_t)op << 16);
    op += 1 + lit + (lit >= 15 ? (lit - 15) / 255 + 1 : 0);
  }
  const bool fail = op > n - 1;
  r.nseq = (uint32_t)(sp - seq);
  r.csize = fail ? ((uint32_t)n | 0x80000000u) : (uint32_t)op;
  r.size = 21u + (uint64_t)(fail ? n : op);
  return r;
}

// ---------------------------------------------------------------------------------------------------------------
// Sub-chunk parallel parse (B2S_LZ4_PIPE=4, the default).  Both thread-per-block walks above are bound by the LATENCY
// of one thread's chain over 32 KiB (ncu, profiles/r2c_*: the event-driven walk executes 40 % fewer instructions than
// the fixed-trip one and is slower, 9.5 ms per 32,768 blocks at 8 % issue utilisation — every iteration ends in
// dependent loads).  The fix is parallelism INSIDE the block: the block is cut into 32 sub-chunks of S = stride / 32
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
