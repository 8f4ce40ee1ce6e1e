// tma_ring.cuh — per-THREAD shared-memory ring over a private, sequentially consumed byte stream, filled by the TMA
// unit with bulk copies (cp.async.bulk.shared::cluster.global, SASS UBLKCP) that complete on per-thread mbarriers.
//
// Why: the thread-per-block kernels of this library (token walks, parses, per-block hashes) stream through one private
// byte range per lane.  With ordinary loads every lane of a warp touches a different cache line, so each load
// instruction is 32 L1 wavefronts and the chain of dependent loads runs at L1/L2 latency.  Here the stream arrives in
// 64-byte pieces ahead of the cursor, asynchronously and without occupying the LSU's global path, and the walk reads
// it from shared memory (one wavefront per distinct bank, ~29 cycles).
//
// Contract: the stream is bytes [0, len) at global address g (any alignment).  The ring holds the 16-byte-aligned
// stream [g & ~15, ...): stream position u = (g & 15) + byte index.  The consumer never looks back: `consume_to(u)`
// declares positions below u dead (their slots are re-armed for the pieces 4 ahead), `ensure(u)` blocks until position u
// has landed.  A piece's tail is rounded up to 16 bytes, i.e. up to 15 bytes past the stream's end are read — inside the
// same 256-byte allocation granule (the library's arenas carry >= 64 bytes of slack; cudaMalloc rounds sizes up).
#pragma once
#include <stdint.h>

namespace b2s {

constexpr int kRingStages = 4, kRingPiece = 64, kRingBytes = kRingStages * kRingPiece;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct TmaRing {
  uint32_t ring, bars;   // shared-window addresses of this thread's ring and of its kRingStages mbarriers
  const uint8_t* g16;    // 16-byte aligned global start of the stream
  int npieces;           // pieces holding bytes of the stream
  int total;             // stream bytes from g16 on, rounded up to 16
  int next_issue, ready; // pieces < next_issue have been requested, pieces < ready have landed (and been waited for)

  __device__ __forceinline__ void issue(int p) {
    const int slot = p & (kRingStages - 1);
    int bytes = total - p * kRingPiece;
    bytes = bytes < kRingPiece ? bytes : kRingPiece;
    const uint32_t bar = bars + 8u * slot, dst = ring + (uint32_t)(slot * kRingPiece);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(g16 + (size_t)p * kRingPiece), "r"(bytes), "r"(bar)
                 : "memory");
  }
  __device__ __forceinline__ void wait(int p) {
    const uint32_t bar = bars + 8u * (p & (kRingStages - 1));
    const uint32_t parity = (uint32_t)(p / kRingStages) & 1u;
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred q;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    }
  }
  // ring_smem: this thread's kRingBytes (16-byte aligned); bar_smem: its kRingStages 8-byte mbarriers
  __device__ __forceinline__ void init(uint8_t* ring_smem, uint64_t* bar_smem, const uint8_t* g, int len) {
    ring = smem_u32(ring_smem);
    bars = smem_u32(bar_smem);
    const uintptr_t a = reinterpret_cast<uintptr_t>(g);
    g16 = reinterpret_cast<const uint8_t*>(a & ~uintptr_t(15));
    total = ((int)(a & 15u) + len + 15) & ~15;
    npieces = (total + kRingPiece - 1) / kRingPiece;
    for (int s = 0; s < kRingStages; s++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8u * s) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the TMA unit sees the initialised barriers
    next_issue = ready = 0;
    for (; next_issue < kRingStages && next_issue < npieces; next_issue++) issue(next_issue);
  }
  // stream position u (and everything below it that is still in the ring) is readable afterwards
  __device__ __forceinline__ void ensure(int u) {
    int p = u / kRingPiece;
    p = p < npieces - 1 ? p : npieces - 1;
    while (ready <= p) {
      wait(ready);
      ready++;
    }
  }
  // positions below u are dead: their pieces' slots are handed to the pieces kRingStages ahead.  A slot is only
  // re-armed after its previous copy has completed (waited for here if the walk skipped over it).
  __device__ __forceinline__ void consume_to(int u) {
    const int keep = u / kRingPiece;  // oldest piece still needed
    while (next_issue - kRingStages < keep && next_issue < npieces) {
      const int oldest = next_issue - kRingStages;
      if (oldest >= ready) {
        wait(oldest);
        ready = oldest + 1;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my reads of the slot precede the unit's writes
      issue(next_issue);
      next_issue++;
    }
  }
  // every copy in flight must have landed before the thread exits (the barriers and the ring die with the CTA)
  __device__ __forceinline__ void drain() {
    for (; ready < next_issue; ready++) wait(ready);
  }
  __device__ __forceinline__ uint32_t word(int u4) const {  // aligned word at stream position u4 (multiple of 4)
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ring + ((uint32_t)u4 & (kRingBytes - 1))));
    return v;
  }
  __device__ __forceinline__ uint32_t byte(int u) const {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(ring + ((uint32_t)u & (kRingBytes - 1))));
    return v;
  }
};

}  // namespace b2s
