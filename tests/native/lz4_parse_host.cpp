// Host build of spark-s3-shuffle_b200/csrc/lz4_parse_core.h for tests/test_parse_core.py: the very function
// lz4_parse2_kernel runs per thread, compiled by g++ so the greedy parse + in-parse extension can be checked against
// the oracle (orc_lz4_compress_block_win / orc_snappy_compress_raw_win) without a GPU.
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

void match_model(const uint8_t* s, int n, int hash_log, uint32_t stride, uint16_t* off) {
  memset(off, 0xEE, stride * 2);  // rows are reused between chunks: whatever the match kernel does not write is garbage
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
    }
    for (int p = pos; p < pos + 32 && p <= mflimit; p++) table[(rd32(s + p) * 2654435761u) >> (32 - hash_log)] = (uint16_t)p;
  }
}

struct MemHost {
  const uint16_t* off;
  const uint8_t* aligned;  // 4-byte aligned start of the word stream
  int ovmax, kmax;
  uint4 off8(int i) const {
    uint4 r;
    memcpy(&r, off + 8 * (size_t)(i < ovmax ? i : ovmax), 16);
    return r;
  }
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

extern "C" {

// parses `n` bytes placed at byte phase `sb` (0..3) of an aligned buffer; returns the result fields and the records
int ph_parse(int codec, const unsigned char* src, int n, int sb, int hash_log, unsigned int* nseq, unsigned int* csize,
             unsigned long long* size, unsigned int* records /* 2 x (n/4+2) */) {
  const uint32_t stride = ((uint32_t)n + 31u) & ~31u;
  std::vector<uint32_t> store((size_t)n / 4 + 4, 0xA5A5A5A5u);
  uint8_t* base = reinterpret_cast<uint8_t*>(store.data());
  memcpy(base + sb, src, (size_t)n);
  std::vector<uint16_t> off(stride ? stride : 32);
  match_model(base + sb, n, hash_log, stride ? stride : 32, off.data());
  MemHost mem{off.data(), base, (int)((stride ? stride : 32) >> 3) - 1, n > 0 ? (sb + n - 1) >> 2 : 0};
  std::vector<uint2> seq((size_t)stride / 4 + 2);
  b2s::lzparse::Result r;
  if (codec == 0) r = b2s::lzparse::parse_block<0>(mem, n, sb, stride, seq.data());
  else if (codec == 1) r = b2s::lzparse::parse_block<1>(mem, n, sb, stride, seq.data());
  else r = b2s::lzparse::parse_block<2>(mem, n, sb, stride, seq.data());
  *nseq = r.nseq;
  *csize = r.csize;
  *size = r.size;
  for (uint32_t i = 0; i < r.nseq; i++) {
    records[2 * i] = seq[i].x;
    records[2 * i + 1] = seq[i].y;
  }
  // the offsets the emit kernels would look up (bit 15 masked as they do)
  const uint32_t omask = stride <= 32768u ? 0x7fffu : 0xffffu;
  for (uint32_t i = 0; i < r.nseq; i++) {
    const uint32_t anchor = seq[i].x & 0xffffu, lit = seq[i].x >> 16, ml = seq[i].y & 0xffffu;
    records[2 * (size_t)(n / 4 + 2) + i] = ml ? (off[anchor + lit] & omask) : 0;
  }
  return 0;
}
}
