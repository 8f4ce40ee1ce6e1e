// Host-only harness for spark-s3-shuffle_b200/host/coalesce.h (tests/test_host_coalesce.py): the group-commit queue
// with the two C-ABI batch calls replaced by stand-ins DEFINED HERE (a byte-wise "codec" that takes a couple of
// milliseconds, like a busy GPU), so that merging, result hand-back, leadership hand-over and error propagation can be
// checked — and run under ThreadSanitizer — without a device.  Test infrastructure only; nothing of this is linked
// into the product libraries.
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200shuffle.h"

static std::atomic<int> g_fail_next{0};
static std::atomic<long> g_calls{0}, g_streams{0};

extern "C" {
const char* b2s_last_error(void) { return "stand-in failure"; }
// "compress": dst = src with every byte + 1, checksum = number of bytes; status -3 when the slot is too small
int b2s_compress_batch(uint32_t, int32_t, uint32_t, uint32_t, uint32_t n, const uint8_t* const* src, const uint64_t* src_len,
                       uint8_t* const* dst, const uint64_t* dst_cap, uint64_t* dst_len, uint64_t* checksum_out,
                       int32_t* status) {
  g_calls++;
  g_streams += n;
  usleep(2000);
  if (g_fail_next.exchange(0)) return B2S_E_CUDA;
  for (uint32_t i = 0; i < n; i++) {
    if (dst_cap[i] < src_len[i]) {
      status[i] = B2S_E_DST_TOO_SMALL;
      dst_len[i] = 0;
      continue;
    }
    for (uint64_t k = 0; k < src_len[i]; k++) dst[i][k] = (uint8_t)(src[i][k] + 1);
    dst_len[i] = src_len[i];
    if (checksum_out) checksum_out[i] = src_len[i];
    status[i] = 0;
  }
  return 0;
}
// "decompress": the inverse; bad_slice = stream length modulo 7 so that per-stream hand-back of that array is visible
int b2s_decompress_batch(uint32_t, uint32_t, uint32_t n, const uint8_t* const* src, const uint64_t* src_len, const uint32_t*,
                         const uint64_t* const*, const uint64_t* const*, uint8_t* const* dst, const uint64_t* dst_cap,
                         uint64_t* dst_len, int32_t* status, int32_t* bad_slice) {
  g_calls++;
  g_streams += n;
  usleep(1000);
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t m = std::min(src_len[i], dst_cap[i]);
    for (uint64_t k = 0; k < m; k++) dst[i][k] = (uint8_t)(src[i][k] - 1);
    dst_len[i] = m;
    status[i] = 0;
    if (bad_slice) bad_slice[i] = (int32_t)(src_len[i] % 7);
  }
  return 0;
}
}

#include "../../spark-s3-shuffle_b200/host/coalesce.h"

using namespace b2s::host;

// every thread: `reps` rounds of compress (5 streams of thread-specific content) + decompress of the result; returns the
// number of mismatches.  stats: queue calls, batches, max merged, streams, stand-in calls.
extern "C" int coalesce_run(int threads, int reps, int fail_one, uint64_t* stats) {
  CoalescingQueue q;
  std::atomic<int> bad{0}, failed{0};
  g_calls = 0;
  g_streams = 0;
  if (fail_one) g_fail_next = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      for (int r = 0; r < reps; r++) {
        const uint32_t n = 5;
        std::vector<std::vector<uint8_t>> in(n), mid(n), out(n);
        std::vector<const uint8_t*> sp(n);
        std::vector<uint8_t*> dp(n);
        std::vector<uint64_t> len(n), cap(n), dlen(n), cks(n);
        std::vector<int32_t> st(n, -99);
        for (uint32_t i = 0; i < n; i++) {
          in[i].resize(1000 + 37 * t + 11 * i + r);
          for (size_t k = 0; k < in[i].size(); k++) in[i][k] = (uint8_t)(k * 7 + t * 31 + i * 3 + r);
          mid[i].resize(in[i].size() + (i == 3 ? 0 : 8));
          sp[i] = in[i].data();
          dp[i] = mid[i].data();
          len[i] = in[i].size();
          cap[i] = i == 4 && r == 1 ? 10 : mid[i].size();  // one slot too small: a per-stream status, not a call failure
        }
        CodecRequest c;
        c.op = 0;
        c.codec = 1;
        c.block_size = 32768;
        c.checksum_alg = 3;
        c.n = n;
        c.src = sp.data();
        c.src_len = len.data();
        c.dst = dp.data();
        c.dst_cap = cap.data();
        c.dst_len = dlen.data();
        c.checksum_out = cks.data();
        c.status = st.data();
        const int rc = q.submit(c);
        if (rc != 0) {
          if (rc == B2S_E_CUDA && c.error == "stand-in failure") failed++;
          else bad++;
          continue;
        }
        for (uint32_t i = 0; i < n; i++) {
          const bool small = i == 4 && r == 1;
          if (small) {
            if (st[i] != B2S_E_DST_TOO_SMALL) bad++;
            continue;
          }
          if (st[i] != 0 || dlen[i] != len[i] || cks[i] != len[i]) bad++;
          for (size_t k = 0; k < in[i].size(); k++)
            if (mid[i][k] != (uint8_t)(in[i][k] + 1)) {
              bad++;
              break;
            }
        }
        // read it back through a decompress request (a different key: merges only with other decompress requests)
        std::vector<const uint8_t*> sp2(n);
        std::vector<uint8_t*> dp2(n);
        std::vector<uint64_t> len2(n), cap2(n), dlen2(n);
        std::vector<int32_t> st2(n, -99), badsl(n, -99);
        for (uint32_t i = 0; i < n; i++) {
          out[i].resize(in[i].size());
          sp2[i] = mid[i].data();
          len2[i] = in[i].size();
          dp2[i] = out[i].data();
          cap2[i] = out[i].size();
        }
        CodecRequest d;
        d.op = 1;
        d.codec = 1;
        d.n = n;
        d.src = sp2.data();
        d.src_len = len2.data();
        d.dst = dp2.data();
        d.dst_cap = cap2.data();
        d.dst_len = dlen2.data();
        d.status = st2.data();
        d.bad_slice = badsl.data();
        if (q.submit(d) != 0) {
          bad++;
          continue;
        }
        for (uint32_t i = 0; i < n; i++) {
          if (i == 4 && r == 1) continue;
          if (st2[i] != 0 || dlen2[i] != len2[i] || badsl[i] != (int32_t)(len2[i] % 7) || out[i] != in[i]) bad++;
        }
      }
    });
  for (auto& x : th) x.join();
  const CoalescingQueue::Statistics s = q.statistics();
  stats[0] = s.calls;
  stats[1] = s.batches;
  stats[2] = s.maxMerged;
  stats[3] = s.streams;
  stats[4] = (uint64_t)g_calls.load();
  stats[5] = (uint64_t)failed.load();
  return bad.load();
}

int main() {
  uint64_t st[6];
  const int bad = coalesce_run(8, 20, 1, st);
  printf("bad=%d calls=%llu batches=%llu maxMerged=%llu streams=%llu backend_calls=%llu failed=%llu\n", bad,
         (unsigned long long)st[0], (unsigned long long)st[1], (unsigned long long)st[2], (unsigned long long)st[3],
         (unsigned long long)st[4], (unsigned long long)st[5]);
  return bad != 0;
}
