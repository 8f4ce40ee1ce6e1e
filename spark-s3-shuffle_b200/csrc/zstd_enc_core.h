// zstd_enc_core.h — Zstandard (RFC 8878) block ENCODING helpers, written once for host and device.
//
// Replaces, for spark.io.compression.codec=zstd on the write side, com.github.luben.zstd.ZstdOutputStreamNoFinalizer [U]
// (zstd-jni -> libzstd ZSTD_compressStream2).  Staged encoder (SURVEY.md §7 "hard parts": valid frames first, ratio
// second): matches come from the shared LZ match finder, a block is
//   Raw_Literals section  +  sequences coded with, per table, either the block's own FSE distribution (histogram ->
//   normalised counts -> table description in the block, FSE_Compressed_Mode) or the PREDEFINED one, whichever the
//   estimate says is smaller,
// or a Raw_Block when that is not smaller.  Any conforming decoder — libzstd / zstd-jni included — reads it.
// (Measured on the terasort shape, tools/zstd_ratio_probe.py: the sequence CODES cost ~12.5 bits per sequence with the
// predefined tables and ~5 with per-block tables, 14 sequences per 104-byte record: ratio 0.43 -> 0.30.  Repeat offsets
// and Huffman literals are worth < 1 % each on this data — the literals are the random key bytes — and are not done.)
// The same functions are compiled into zstd_enc.cu (product) and into the host unit test (tests/native), where the
// frames they produce are decoded by libzstd.so.1.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2S_HD __host__ __device__
#else
#ifndef B2S_HD
#define B2S_HD
#endif
#endif

namespace b2s {
namespace zstdenc {

// ---- FSE compression tables for the three predefined distributions (RFC 8878 3.1.1.3.2.2) ---------------------------
struct SymbolTT {
  int32_t deltaFindState;
  uint32_t deltaNbBits;
};
// Accuracy logs of the per-block tables: the predefined ones use 6 / 5 / 6; the format allows up to 9 / 8 / 9.
// Measured on the terasort shape (tools/zstd_ratio_probe.py): ratio 0.3037 / 0.2996 / 0.2981 / 0.2978 at 6,5,6 / 7,6,7 /
// 8,7,8 / 9,8,9 — 7,6,7 keeps a thread's private tables at 1.6 KB.
#ifndef B2S_ZSTD_LOGS
#define B2S_ZSTD_LOGS 7, 6, 7
#endif
constexpr int kZstdLogs[3] = {B2S_ZSTD_LOGS};  // tools/zstd_ratio_probe.py -DB2S_ZSTD_LOGS=9,8,9 measures other choices
constexpr int kLLLog = kZstdLogs[0], kOFLog = kZstdLogs[1], kMLLog = kZstdLogs[2];
constexpr int kLLSyms = 36, kOFSyms = 32, kMLSyms = 53;  // OF: offsets of a <= 64 KiB block need codes <= 16
struct CTables {  // built once on the host, uploaded to the device: predefined distributions
  uint16_t ll_state[64], of_state[32], ml_state[64];
  SymbolTT ll_tt[36], of_tt[29], ml_tt[53];
  uint16_t ll_cost[36], of_cost[32], ml_cost[53];  // bits per symbol in 1/256 bit under the predefined distributions
};
struct BlockTables {  // one block's own tables (thread-private on the device)
  SymbolTT ll_tt[kLLSyms], of_tt[kOFSyms], ml_tt[kMLSyms];
  uint16_t ll_state[1 << kLLLog], of_state[1 << kOFLog], ml_state[1 << kMLLog];
};

B2S_HD inline int hb32(uint32_t v) {
  int r = 0;
  while (v >>= 1) r++;
  return r;
}

// FSE_buildCTable for one distribution of accuracy `log` <= 9; `sym` is scratch of 1 << log bytes.
// norm[s] == -1 ("less than one") only occurs in the predefined distributions.
B2S_HD inline void build_ctable(const int16_t* norm, int nsym, int log, uint16_t* stateTable, SymbolTT* tt, uint8_t* sym) {
  const int size = 1 << log;
  int cumul[64];
  int high = size - 1;
  cumul[0] = 0;
  for (int u = 1; u <= nsym; u++) {
    if (norm[u - 1] == -1) {
      cumul[u] = cumul[u - 1] + 1;
      sym[high--] = (uint8_t)(u - 1);
    } else {
      cumul[u] = cumul[u - 1] + norm[u - 1];
    }
  }
  const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
  int pos = 0;
  for (int s = 0; s < nsym; s++)
    for (int i = 0; i < norm[s]; i++) {
      sym[pos] = (uint8_t)s;
      do pos = (pos + step) & mask;
      while (pos > high);
    }
  for (int u = 0; u < size; u++) {
    const int s = sym[u];
    stateTable[cumul[s]++] = (uint16_t)(size + u);
  }
  int total = 0;
  for (int s = 0; s < nsym; s++) {
    const int n = norm[s];
    if (n == 0) {
      tt[s].deltaNbBits = (uint32_t)(((log + 1) << 16) - (1 << log));
      tt[s].deltaFindState = 0;
    } else if (n == -1 || n == 1) {
      tt[s].deltaNbBits = (uint32_t)((log << 16) - (1 << log));
      tt[s].deltaFindState = total - 1;
      total++;
    } else {
      const int maxBitsOut = log - hb32((uint32_t)(n - 1));
      const int minStatePlus = n << maxBitsOut;
      tt[s].deltaNbBits = (uint32_t)((maxBitsOut << 16) - minStatePlus);
      tt[s].deltaFindState = total - n;
      total += n;
    }
  }
}

// log2(x) in 1/256 bit for x >= 1 (piecewise linear between powers of two: error < 0.09 bit, used for estimates only)
B2S_HD inline uint32_t log2_q8(uint32_t x) {
  const int h = hb32(x);
  return ((uint32_t)h << 8) + (h >= 8 ? (x >> (h - 8)) - 256u : (x << (8 - h)) - 256u);
}

inline void build_predefined(CTables* t) {
  const int16_t LL[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
  const int16_t OF[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
  const int16_t ML[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                          1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
  uint8_t scratch[64];
  build_ctable(LL, 36, 6, t->ll_state, t->ll_tt, scratch);
  build_ctable(OF, 29, 5, t->of_state, t->of_tt, scratch);
  build_ctable(ML, 53, 6, t->ml_state, t->ml_tt, scratch);
  for (int s = 0; s < 36; s++) t->ll_cost[s] = (uint16_t)((6u << 8) - log2_q8((uint32_t)(LL[s] < 0 ? 1 : LL[s])));
  for (int s = 0; s < 32; s++) t->of_cost[s] = s < 29 ? (uint16_t)((5u << 8) - log2_q8((uint32_t)(OF[s] < 0 ? 1 : OF[s]))) : 0xffff;
  for (int s = 0; s < 53; s++) t->ml_cost[s] = (uint16_t)((6u << 8) - log2_q8((uint32_t)(ML[s] < 0 ? 1 : ML[s])));
}

// ---- symbol codes ---------------------------------------------------------------------------------------------------
B2S_HD inline int highbit(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz(v);
#else
  int r = 0;
  while (v >>= 1) r++;
  return r;
#endif
}
// literal length -> code, extra bits, extra value
B2S_HD inline void ll_encode(uint32_t ll, int* code, int* nbits, uint32_t* extra) {
  int c;
  if (ll < 16) c = (int)ll;
  else if (ll < 64) {
    // 16,16,17,17,18,18,19,19,20,20,20,20,21,21,21,21,22 x8,23 x8,24 x16
    c = ll < 24 ? 16 + (int)((ll - 16) >> 1) : ll < 32 ? 20 + (int)((ll - 24) >> 2) : ll < 48 ? 22 + (int)((ll - 32) >> 3) : 24;
  } else {
    c = highbit(ll) + 19;
  }
  // baseline / extra bits in closed form (RFC 8878 table of literal-length codes): no per-call lookup table
  uint32_t base;
  int nb;
  if (c < 16) { base = (uint32_t)c; nb = 0; }
  else if (c < 20) { base = 16u + 2u * (uint32_t)(c - 16); nb = 1; }
  else if (c < 22) { base = 24u + 4u * (uint32_t)(c - 20); nb = 2; }
  else if (c < 24) { base = 32u + 8u * (uint32_t)(c - 22); nb = 3; }
  else if (c == 24) { base = 48u; nb = 4; }
  else if (c == 25) { base = 64u; nb = 6; }
  else { base = 1u << (c - 19); nb = c - 19; }
  *code = c;
  *nbits = nb;
  *extra = ll - base;
}
// match length (>= 3) -> code, extra bits, extra value
B2S_HD inline void ml_encode(uint32_t ml, int* code, int* nbits, uint32_t* extra) {
  const uint32_t mb = ml - 3;
  int c;
  if (mb < 32) c = (int)mb;
  else if (mb < 128) {
    // 32,32,33,33,34,34,35,35,36 x4,37 x4,38 x8,39 x8,40 x16,41 x16,42 x32
    c = mb < 40 ? 32 + (int)((mb - 32) >> 1) : mb < 48 ? 36 + (int)((mb - 40) >> 2) : mb < 64 ? 38 + (int)((mb - 48) >> 3)
        : mb < 96 ? 40 + (int)((mb - 64) >> 4) : 42;
  } else {
    c = highbit(mb) + 36;
  }
  uint32_t base;
  int nb;
  if (c < 32) { base = (uint32_t)c + 3u; nb = 0; }
  else if (c < 36) { base = 35u + 2u * (uint32_t)(c - 32); nb = 1; }
  else if (c < 38) { base = 43u + 4u * (uint32_t)(c - 36); nb = 2; }
  else if (c < 40) { base = 51u + 8u * (uint32_t)(c - 38); nb = 3; }
  else if (c < 42) { base = 67u + 16u * (uint32_t)(c - 40); nb = 4; }
  else if (c == 42) { base = 99u; nb = 5; }
  else if (c == 43) { base = 131u; nb = 7; }
  else { base = (1u << (c - 36)) + 3u; nb = c - 36; }
  *code = c;
  *nbits = nb;
  *extra = ml - base;
}

// ---- forward bit writer (the sequence bitstream is written forwards, sequences in reverse order) ------------------
struct BitWriter {
  uint8_t* p;       // 4-byte aligned, with at least 8 bytes of slack beyond `cap`
  uint32_t cap, n;  // bytes written so far; overflow => n ends > cap (nothing is stored past cap + 4)
  uint64_t acc;
  int fill;
  B2S_HD void init(uint8_t* dst, uint32_t capacity) {
    p = dst;
    cap = capacity;
    n = 0;
    acc = 0;
    fill = 0;
  }
  // Flushes 32 bits at a time with one aligned word store.  (The first version flushed byte by byte in a loop whose
  // trip count differed between the 32 blocks a warp encodes: ~900 warp-instructions per sequence, profiles/r1z.)
  B2S_HD void add(uint32_t v, int nb) {  // nb <= 24 per call, fill < 32 on entry
    acc |= (uint64_t)(v & ((1u << nb) - 1u)) << fill;
    fill += nb;
    if (fill >= 32) {
      if (n <= cap) {
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint32_t*>(p + n) = (uint32_t)acc;
#else
        for (int k = 0; k < 4; k++) p[n + k] = (uint8_t)(acc >> (8 * k));
#endif
      }
      n += 4;
      acc >>= 32;
      fill -= 32;
    }
  }
  B2S_HD void close() {  // final 1-bit marker, then the last partial word byte by byte
    add(1, 1);
    while (fill > 0) {
      if (n <= cap) p[n] = (uint8_t)acc;
      n++;
      acc >>= 8;
      fill -= 8;
    }
    fill = 0;
  }
};

struct FseState {
  uint32_t v;
};
B2S_HD inline void fse_init(FseState* s, const uint16_t* st, const SymbolTT* tt, int sym) {
  const uint32_t nb = (tt[sym].deltaNbBits + (1u << 15)) >> 16;
  const uint32_t value = (nb << 16) - tt[sym].deltaNbBits;
  s->v = st[(value >> nb) + tt[sym].deltaFindState];
}
B2S_HD inline void fse_encode(BitWriter* bw, FseState* s, const uint16_t* st, const SymbolTT* tt, int sym) {
  const uint32_t nb = (s->v + tt[sym].deltaNbBits) >> 16;
  bw->add(s->v, (int)nb);
  s->v = st[(s->v >> nb) + tt[sym].deltaFindState];
}

// One sequence as the encoder sees it
struct Seq {
  uint32_t ll, ml, off;  // literal length, match length (>= 4 here), real offset (>= 1)
};

// One of the three tables as the bitstream coder sees it: predefined (CTables) or the block's own (BlockTables)
struct EncTab {
  const uint16_t* st;
  const SymbolTT* tt;
  int log;
};

// Encodes `nseq` sequences (get(i) returns sequence i, 0 <= i < nseq, in parse order) into dst; returns the byte
// count, or cap + 1 when it does not fit.
template <typename Get>
B2S_HD inline uint32_t encode_sequences(const EncTab L, const EncTab O, const EncTab M, uint32_t nseq, Get get,
                                        uint8_t* dst, uint32_t cap) {
  BitWriter bw;
  bw.init(dst, cap);
  FseState sl, so, sm;
  int lc, ln, mc, mn;
  uint32_t le, me;
  {
    const Seq q = get(nseq - 1);
    const uint32_t ofv = q.off + 3;
    const int oc = highbit(ofv);
    ll_encode(q.ll, &lc, &ln, &le);
    ml_encode(q.ml, &mc, &mn, &me);
    fse_init(&sm, M.st, M.tt, mc);
    fse_init(&so, O.st, O.tt, oc);
    fse_init(&sl, L.st, L.tt, lc);
    bw.add(le, ln);
    bw.add(me, mn);
    bw.add(ofv - (1u << oc), oc);
  }
  for (uint32_t i = nseq - 1; i-- > 0;) {
    const Seq q = get(i);
    const uint32_t ofv = q.off + 3;
    const int oc = highbit(ofv);
    ll_encode(q.ll, &lc, &ln, &le);
    ml_encode(q.ml, &mc, &mn, &me);
    fse_encode(&bw, &so, O.st, O.tt, oc);
    fse_encode(&bw, &sm, M.st, M.tt, mc);
    fse_encode(&bw, &sl, L.st, L.tt, lc);
    bw.add(le, ln);
    bw.add(me, mn);
    bw.add(ofv - (1u << oc), oc);
    if (bw.n > cap) return cap + 1;
  }
  bw.add(sm.v, M.log);
  bw.add(so.v, O.log);
  bw.add(sl.v, L.log);
  bw.close();
  return bw.n > cap ? cap + 1 : bw.n;
}

// ---- a block's own distributions (FSE_Compressed_Mode, RFC 8878 3.1.1.3.2.1 / 4.1.1) ---------------------------------
// counts -> probabilities that sum to 1 << log, every symbol that occurs >= 1 (no "less than one" entries).  Floor of the
// exact share, at least 1; what is left over goes to the most frequent symbol, what was over-spent by the "at least 1"
// rule is taken back from the largest entries.  Needs at least one symbol with cnt > 0 and at most 1 << log of them.
B2S_HD inline void normalize_counts(const uint16_t* cnt, int nsym, uint32_t total, int log, int16_t* norm) {
  const uint32_t size = 1u << log;
  int32_t sum = 0;
  int largest = 0;
  for (int s = 0; s < nsym; s++) {
    uint32_t v = 0;
    if (cnt[s]) {
      v = (uint32_t)(((uint64_t)cnt[s] << log) / total);
      if (v == 0) v = 1;
      if (cnt[s] > cnt[largest]) largest = s;
    }
    norm[s] = (int16_t)v;
    sum += (int32_t)v;
  }
  int32_t rem = (int32_t)size - sum;
  if (rem >= 0) {
    norm[largest] = (int16_t)(norm[largest] + rem);
    return;
  }
  while (rem < 0) {
    int big = 0;
    for (int s = 1; s < nsym; s++)
      if (norm[s] > norm[big]) big = s;
    norm[big]--;
    rem++;
  }
}

// table description (the inverse of fse_read_header in zstd_core.h); returns its bytes
B2S_HD inline uint32_t write_ncount(uint8_t* dst, const int16_t* norm, int log) {
  uint64_t acc = (uint64_t)(log - 5);
  int fill = 4;
  uint32_t n = 0;
  int remaining = 1 << log, s = 0;
  while (remaining > 0) {  // the probabilities sum to 1 << log: a symbol with norm > 0 always follows
    const int bits = hb32((uint32_t)(remaining + 1)) + 1;
    const uint32_t lower = (1u << (bits - 1)) - 1u, threshold = (1u << bits) - 1u - (uint32_t)(remaining + 1);
    const uint32_t v = (uint32_t)norm[s] + 1u;
    if (v < threshold) {
      acc |= (uint64_t)v << fill;
      fill += bits - 1;
    } else {
      acc |= (uint64_t)(v <= lower ? v : v + threshold) << fill;
      fill += bits;
    }
    remaining -= norm[s];
    const bool zero = norm[s] == 0;
    s++;
    if (zero) {  // how many more zero-probability symbols follow, in 2-bit digits (3 = "3 and continue")
      int z = 0;
      while (norm[s + z] == 0) z++;
      s += z;
      for (; z >= 3; z -= 3) {
        acc |= (uint64_t)3 << fill;
        fill += 2;
        if (fill >= 32) {
          for (int k = 0; k < 4; k++) dst[n++] = (uint8_t)(acc >> (8 * k));
          acc >>= 32;
          fill -= 32;
        }
      }
      acc |= (uint64_t)z << fill;
      fill += 2;
    }
    if (fill >= 32) {
      for (int k = 0; k < 4; k++) dst[n++] = (uint8_t)(acc >> (8 * k));
      acc >>= 32;
      fill -= 32;
    }
  }
  for (; fill > 0; fill -= 8) {
    dst[n++] = (uint8_t)acc;
    acc >>= 8;
  }
  return n;
}

// estimated size of one field's codes in 1/256 bit: under the block's own probabilities / under the predefined ones
B2S_HD inline uint32_t cost_own_q8(const uint16_t* cnt, const int16_t* norm, int nsym, int log) {
  uint32_t c = 0;
  for (int s = 0; s < nsym; s++)
    if (cnt[s]) c += cnt[s] * (((uint32_t)log << 8) - log2_q8((uint32_t)norm[s]));
  return c;
}
B2S_HD inline uint32_t cost_predefined_q8(const uint16_t* cnt, const uint16_t* cost, int nsym) {
  uint32_t c = 0;
  for (int s = 0; s < nsym; s++) c += (uint32_t)cnt[s] * cost[s];
  return c;
}

// Decides one field: own table (description appended to hdr, coding table built into st / tt) or predefined.
// Returns the compression mode (0 predefined, 2 FSE_Compressed).
B2S_HD inline int choose_table(const uint16_t* cnt, int nsym, uint32_t nseq, int log, const uint16_t* predef_cost,
                               uint8_t* hdr, uint32_t* h, uint16_t* st, SymbolTT* tt, uint8_t* scratch) {
  int distinct = 0;
  for (int s = 0; s < nsym; s++) distinct += cnt[s] != 0;
  if (distinct < 2) return 0;  // a single symbol would be RLE_Mode; the predefined table codes it in a few bits
  int16_t norm[kMLSyms + 1];
  normalize_counts(cnt, nsym, nseq, log, norm);
  norm[nsym] = 1;  // sentinel for write_ncount's zero-run scan
  const uint32_t nb = write_ncount(hdr + *h, norm, log);
  if (cost_own_q8(cnt, norm, nsym, log) + ((nb * 8u) << 8) >= cost_predefined_q8(cnt, predef_cost, nsym)) return 0;
  *h += nb;
  build_ctable(norm, nsym, log, st, tt, scratch);
  return 2;
}

constexpr uint32_t kSeqHeaderMax = 192;  // Compression_Modes byte + three table descriptions (< 1 + 46 + 33 + 68 bytes)

// The sequences section of one block after its Number_of_Sequences field: hdr receives the Compression_Modes byte and
// the table descriptions (*hdr_bytes of them, <= kSeqHeaderMax), dst the bitstream (returned byte count, cap + 1 when it
// does not fit).  dst must be 4-byte aligned.  B is scratch for the block's own tables.
template <typename Get>
B2S_HD inline uint32_t encode_block_sequences(const CTables* T, BlockTables* B, uint32_t nseq, Get get, uint8_t* hdr,
                                              uint32_t* hdr_bytes, uint8_t* dst, uint32_t cap) {
  uint16_t cl[kLLSyms], co[kOFSyms], cm[kMLSyms];
  for (int s = 0; s < kLLSyms; s++) cl[s] = 0;
  for (int s = 0; s < kOFSyms; s++) co[s] = 0;
  for (int s = 0; s < kMLSyms; s++) cm[s] = 0;
  for (uint32_t i = 0; i < nseq; i++) {
    const Seq q = get(i);
    int c, nb;
    uint32_t e;
    ll_encode(q.ll, &c, &nb, &e);
    cl[c]++;
    ml_encode(q.ml, &c, &nb, &e);
    cm[c]++;
    co[highbit(q.off + 3)]++;
  }
  uint8_t scratch[1 << (kLLLog > kMLLog ? kLLLog : kMLLog)];
  uint32_t h = 1;
  EncTab L{T->ll_state, T->ll_tt, 6}, O{T->of_state, T->of_tt, 5}, M{T->ml_state, T->ml_tt, 6};
  const int ml_ = choose_table(cl, kLLSyms, nseq, kLLLog, T->ll_cost, hdr, &h, B->ll_state, B->ll_tt, scratch);
  if (ml_) L = EncTab{B->ll_state, B->ll_tt, kLLLog};
  const int mo_ = choose_table(co, kOFSyms, nseq, kOFLog, T->of_cost, hdr, &h, B->of_state, B->of_tt, scratch);
  if (mo_) O = EncTab{B->of_state, B->of_tt, kOFLog};
  const int mm_ = choose_table(cm, kMLSyms, nseq, kMLLog, T->ml_cost, hdr, &h, B->ml_state, B->ml_tt, scratch);
  if (mm_) M = EncTab{B->ml_state, B->ml_tt, kMLLog};
  hdr[0] = (uint8_t)((ml_ << 6) | (mo_ << 4) | (mm_ << 2));
  *hdr_bytes = h;
  return encode_sequences(L, O, M, nseq, get, dst, cap);
}

// ---- headers -----------------------------------------------------------------------------------------------------
constexpr uint32_t kFrameHeaderBytes = 6;  // magic + Frame_Header_Descriptor (no FCS, no dict, no checksum) + window
constexpr uint32_t kEndBlockBytes = 3;     // every stream ends with an empty Raw_Block carrying Last_Block
B2S_HD inline void put_frame_header(uint8_t* p) {
  p[0] = 0x28; p[1] = 0xB5; p[2] = 0x2F; p[3] = 0xFD;
  p[4] = 0x00;  // FCS flag 0, single segment 0, no checksum, no dictionary
  p[5] = 0x38;  // window descriptor: exponent 7, mantissa 0 -> 128 KiB (>= any offset a <= 64 KiB block produces)
}
B2S_HD inline void put_block_header(uint8_t* p, int last, int type, uint32_t size) {
  const uint32_t v = (uint32_t)last | ((uint32_t)type << 1) | (size << 3);
  p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
}
B2S_HD inline uint32_t raw_literals_header_bytes(uint32_t n) { return n < 32 ? 1 : n < 4096 ? 2 : 3; }
B2S_HD inline void put_raw_literals_header(uint8_t* p, uint32_t n) {
  if (n < 32) {
    p[0] = (uint8_t)(n << 3);
  } else if (n < 4096) {
    const uint32_t v = (n << 4) | (1u << 2);
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8);
  } else {
    const uint32_t v = (n << 4) | (3u << 2);
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
  }
}
B2S_HD inline uint32_t nseq_header_bytes(uint32_t nseq) { return nseq < 128 ? 1 : nseq < 0x7F00 ? 2 : 3; }
B2S_HD inline void put_nseq(uint8_t* p, uint32_t nseq) {
  if (nseq < 128) {
    p[0] = (uint8_t)nseq;
  } else if (nseq < 0x7F00) {
    p[0] = (uint8_t)((nseq >> 8) + 128);
    p[1] = (uint8_t)nseq;
  } else {
    p[0] = 255;
    p[1] = (uint8_t)(nseq - 0x7F00);
    p[2] = (uint8_t)((nseq - 0x7F00) >> 8);
  }
}

}  // namespace zstdenc
}  // namespace b2s
