// Host build of spark-s3-shuffle_b200/csrc/lz4_parse_core.h for tests/test_parse_core.py: the lane-local walk of the
// sub-chunk parallel parse (B2S_LZ4_PIPE=4) compiled by g++, run for the 32 lanes of a warp and stitched like
// lz4_parse4_kernel does, so the result can be checked against the oracle (orc_lz4_compress_block_win_sub /
// orc_snappy_compress_raw_win_sub) without a GPU.
// Test infrastructure only — the C ABI never runs this.
//
// The off[] input is produced here by a plain restatement of what lz4_match2_kernel writes (the oracle's
// win_find_offsets, oracle/b2s_oracle.c, plus the "exactly 4" flag in bit 15 for rows of <= 32768 positions), and the
// records are turned into an LZ4 block by a plain restatement of lz4_emit_kernel's byte layout.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../spark-s3-shuffle_b200/csrc/lz4_parse_core.h"

namespace {

uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

void match_model(const uint8_t* s, int n, int hash_log, uint32_t stride, uint16_t* off, uint32_t* mask) {
  memset(off, 0xEE, stride * 2);  // rows are reused between chunks: whatever the match kernel does not write is garbage
  memset(mask, 0xEE, stride / 8);
  if (n < 13) return;
  const int mflimit = n - 12;
  std::vector<uint16_t> table((size_t)1 << hash_log, 0);
  const bool flag4 = stride <= 32768u;
  for (int pos = 0; pos <= mflimit; pos += 32) {
    for (int p = pos; p < pos + 32; p++) {
      uint32_t o = 0;
      if (p <= mflimit) {
        const uint32_t v = rd32(s + p);
        const int c = table[(v * 2654435761u) >> (32 - hash_log)];
        if (p > 0 && rd32(s + p - 1) == v) o = 1;
        else if (c < p && rd32(s + c) == v) o = (uint32_t)(p - c);
        if (flag4 && o && s[p + 4] != s[p + 4 - o]) o |= 0x8000u;
      }
      off[p] = (uint16_t)o;  // the kernel stores all 32 lanes of the window (invalid positions: 0)
      if (p == pos) mask[pos >> 5] = 0;
      if (o) mask[pos >> 5] |= 1u << (p - pos);  // the window's match bits (one word per window)
    }
    for (int p = pos; p < pos + 32 && p <= mflimit; p++) table[(rd32(s + p) * 2654435761u) >> (32 - hash_log)] = (uint16_t)p;
  }
}

struct MemHost {
  const uint32_t* maskrow;
  uint32_t mask(int w) const { return maskrow[w]; }
  uint32_t off16(int q) const { return off[q]; }
  const uint16_t* off;
  const uint8_t* aligned;  // 4-byte aligned start of the word stream
  int ovmax, kmax;
  uint32_t word(int k) const {
    uint32_t r;
    memcpy(&r, aligned + 4 * (size_t)(k < kmax ? k : kmax), 4);
    return r;
  }
  uint32_t cand_word(int k) const {
    if (k > kmax) abort();  // the device build would read past the block's last word
    uint32_t r;
    memcpy(&r, aligned + 4 * (size_t)k, 4);
    return r;
  }
};

}  // namespace

// Generation 4: walk_subchunk() for the 32 lanes of a warp, then the kernel's stitch (lz4_parse4_kernel steps 1-4)
// restated with plain loops.  stride / S exactly as the kernel derives them from the codec block size.
extern "C" int ph_parse4(int codec, const unsigned char* src, int n, int sb, int hash_log, unsigned block_size,
                         unsigned int* nseq, unsigned int* csize, unsigned long long* size, unsigned int* records) {
  using namespace b2s::lzparse;
  const uint32_t stride = (block_size + 31u) & ~31u;
  const int S = (int)(((stride >> 5) + 31u) & ~31u), slot = S / 4 + 1;
  std::vector<uint32_t> store((size_t)n / 4 + 4, 0xA5A5A5A5u);
  uint8_t* base = reinterpret_cast<uint8_t*>(store.data());
  memcpy(base + sb, src, (size_t)n);
  std::vector<uint16_t> off(stride);
  std::vector<uint32_t> mask(stride / 32);
  match_model(base + sb, n, hash_log, stride, off.data(), mask.data());
  MemHost mem{mask.data(), off.data(), base, (int)(stride >> 3) - 1, n > 0 ? (sb + n - 1) >> 2 : 0};
  std::vector<uint2> seq((size_t)32 * slot + 2);
  SubResult R[32];
  for (int k = 0; k < 32; k++) {
    R[k] = SubResult{0, 0, 0, 0, 0, -1};
    const int lo = k * S;
    if (n >= 13 && lo < n) {
      const int hi = lo + S < n ? lo + S : n;
      if (codec == 0) R[k] = walk_subchunk<0>(mem, n, sb, stride, lo, hi, seq.data() + (size_t)k * slot);
      else if (codec == 1) R[k] = walk_subchunk<1>(mem, n, sb, stride, lo, hi, seq.data() + (size_t)k * slot);
      else R[k] = walk_subchunk<2>(mem, n, sb, stride, lo, hi, seq.data() + (size_t)k * slot);
    }
  }
  int anchor_in[32], size0[32], basek[32], nb[32];
  int amax = -1, total = codec == 1 ? (n < 128 ? 1 : n < 16384 ? 2 : 3) : 0, N = 0;
  for (int k = 0; k < 32; k++) {
    anchor_in[k] = amax < 0 ? 0 : amax;
    if (R[k].last_end > amax) amax = R[k].last_end;
    size0[k] = 0;
    if (R[k].nrec) {
      const int lit = R[k].pm0 - anchor_in[k];
      size0[k] = codec == 0 ? seq_size<0>(lit, R[k].len0, R[k].d0) : codec == 1 ? seq_size<1>(lit, R[k].len0, R[k].d0)
                                                                                : seq_size<2>(lit, R[k].len0, R[k].d0);
    }
    basek[k] = total;
    nb[k] = N;
    total += size0[k] + R[k].rest;
    N += R[k].nrec;
  }
  const int anchor_f = amax < 0 ? 0 : amax;
  std::vector<uint2> dense((size_t)N + 1);
  for (int k = 0; k < 32; k++)
    for (int j = 0; j < R[k].nrec; j++) {
      uint2 r = seq[(size_t)k * slot + j];
      if (j == 0) {
        const int pm = (int)r.x, len = (int)r.y;
        r.x = (uint32_t)anchor_in[k] | ((uint32_t)(pm - anchor_in[k]) << 16);
        r.y = (uint32_t)len | ((uint32_t)basek[k] << 16);
      } else {
        r.y += (uint32_t)(basek[k] + size0[k]) << 16;
      }
      dense[(size_t)nb[k] + j] = r;
    }
  uint32_t ns = (uint32_t)N;
  const int lit = n - anchor_f;
  if (codec == 2) {
    dense[ns++] = make_uint2((uint32_t)anchor_f | ((uint32_t)lit << 16), (uint32_t)total << 16);
    *csize = 0;
    *size = 0;
  } else if (codec == 1) {
    if (lit) {
      dense[ns++] = make_uint2((uint32_t)anchor_f | ((uint32_t)lit << 16), (uint32_t)total << 16);
      total += lit + (lit - 1 < 60 ? 1 : lit - 1 < 256 ? 2 : 3);
    }
    *csize = (uint32_t)total;
    *size = 4u + (uint64_t)total;
  } else {
    dense[ns++] = make_uint2((uint32_t)anchor_f | ((uint32_t)lit << 16), (uint32_t)total << 16);
    total += 1 + lit + (lit >= 15 ? (lit - 15) / 255 + 1 : 0);
    const bool fail = total > n - 1;
    *csize = fail ? ((uint32_t)n | 0x80000000u) : (uint32_t)total;
    *size = 21u + (uint64_t)(fail ? n : total);
  }
  *nseq = ns;
  const uint32_t omask = stride <= 32768u ? 0x7fffu : 0xffffu;
  for (uint32_t i = 0; i < ns; i++) {
    records[2 * i] = dense[i].x;
    records[2 * i + 1] = dense[i].y;
    const uint32_t anchor = dense[i].x & 0xffffu, l = dense[i].x >> 16, ml = dense[i].y & 0xffffu;
    records[2 * (size_t)(n / 4 + 34) + i] = ml ? (off[anchor + l] & omask) : 0;
  }
  return 0;
}
