// coalesce.h — group commit for the codec calls of concurrent task threads (internal to shuffle_host.cpp).
//
// SURVEY.md §8(b), threading row: an executor runs spark.executor.cores task threads, each with its own writer / reader
// (shuffle/S3ShuffleMapOutputWriter.scala is per map task, storage/S3ShuffleReader.scala per reduce task), and a
// task-sized call is small for a B200: one map task's 200 partitions take 8 ms of which most is the fixed cost of the
// thread-per-block kernels and un-overlapped copies.  The C ABI serialises callers per device, so N task
// threads would pay that floor N times.
//
// CoalescingQueue merges them instead, the way a database group-commits: the first caller becomes the leader and runs
// its request at once (no timer, no added latency when there is one thread); whatever other threads submit while the
// GPU is busy is taken TOGETHER in the leader's next round — one b2s_compress_batch / b2s_decompress_batch over the
// concatenated stream lists — and the results are handed back per request.  Leadership passes to a waiting thread as
// soon as the leader's own request is done.  Requests merge only when codec, level, block size and checksum algorithm
// agree (one executor = one SparkConf, so they do).
#pragma once

namespace b2s {
namespace host {

struct CodecRequest {
  int op = 0;  // 0 = compress, 1 = decompress
  uint32_t codec = 0, block_size = 0, checksum_alg = 0, n = 0;
  int32_t level = 0;
  const uint8_t* const* src = nullptr;
  const uint64_t* src_len = nullptr;
  uint8_t* const* dst = nullptr;
  const uint64_t* dst_cap = nullptr;
  uint64_t* dst_len = nullptr;
  uint64_t* checksum_out = nullptr;             // compress
  const uint32_t* n_slices = nullptr;           // decompress, checksum_alg != 0
  const uint64_t* const* slice_len = nullptr;
  const uint64_t* const* slice_checksum = nullptr;
  int32_t* bad_slice = nullptr;                 // decompress, optional
  int32_t* status = nullptr;
  // result
  int rc = 0;
  std::string error;
  bool done = false;
  bool sameKey(const CodecRequest& o) const {
    return op == o.op && codec == o.codec && block_size == o.block_size && checksum_alg == o.checksum_alg &&
           level == o.level;
  }
};

class CoalescingQueue {
 public:
  struct Statistics {
    uint64_t calls = 0, batches = 0, maxMerged = 0, streams = 0, isolated = 0;  // isolated: merged rounds re-run per request after a failure
  };

  // blocks until the request has been executed (by this thread or by another thread's round); returns its rc
  int submit(CodecRequest& r) {
    std::unique_lock<std::mutex> lk(m_);
    pending_.push_back(&r);
    stats_.calls++;
    stats_.streams += r.n;
    cv_.wait(lk, [&] { return r.done || !leader_; });
    if (r.done) return r.rc;
    leader_ = true;
    while (!r.done) runRound(lk);
    leader_ = false;  // hand over: a waiting thread whose request is still pending takes the next round
    cv_.notify_all();
    return r.rc;
  }
  Statistics statistics() {
    std::unique_lock<std::mutex> lk(m_);
    return stats_;
  }

 private:
  void runRound(std::unique_lock<std::mutex>& lk) {
    // everything pending with the key of the oldest request
    std::vector<CodecRequest*> round;
    const CodecRequest key = *pending_.front();
    for (auto it = pending_.begin(); it != pending_.end();) {
      if ((*it)->sameKey(key)) {
        round.push_back(*it);
        it = pending_.erase(it);
      } else {
        ++it;
      }
    }
    lk.unlock();
    size_t total = 0;
    for (auto* q : round) total += q->n;
    int rc = 0;
    std::string err;
    if (round.size() == 1) {
      rc = call(*round[0], round[0]->n, round[0]->src, round[0]->src_len, round[0]->dst, round[0]->dst_cap,
                round[0]->dst_len, round[0]->checksum_out, round[0]->n_slices, round[0]->slice_len,
                round[0]->slice_checksum, round[0]->status, round[0]->bad_slice);
      if (rc != 0) err = b2s_last_error();
    } else {
      std::vector<const uint8_t*> src(total);
      std::vector<uint64_t> len(total), cap(total), dlen(total), cks(total);
      std::vector<uint8_t*> dst(total);
      std::vector<uint32_t> nsl(total);
      std::vector<const uint64_t*> sl(total), sc(total);
      std::vector<int32_t> status(total), bad(total, -1);
      size_t at = 0;
      const bool slices = key.op == 1 && key.checksum_alg != 0;
      for (auto* q : round) {
        for (uint32_t i = 0; i < q->n; i++, at++) {
          src[at] = q->src[i];
          len[at] = q->src_len[i];
          dst[at] = q->dst[i];
          cap[at] = q->dst_cap[i];
          if (slices) {
            nsl[at] = q->n_slices[i];
            sl[at] = q->slice_len[i];
            sc[at] = q->slice_checksum[i];
          }
        }
      }
      rc = call(key, (uint32_t)total, src.data(), len.data(), dst.data(), cap.data(), dlen.data(), cks.data(),
                slices ? nsl.data() : nullptr, slices ? sl.data() : nullptr, slices ? sc.data() : nullptr, status.data(),
                bad.data());
      if (rc != 0) {
        // A call-level failure of a MERGED batch (one task's truncated block, a header claiming an impossible size,
        // ...) must not fail the tasks it happened to be merged with: the reference surfaces such errors per block,
        // to the task that owns it.  Re-run the round one request at a time; only the offender keeps the error.
        std::vector<int> rcs(round.size(), 0);
        std::vector<std::string> errs(round.size());
        for (size_t k = 0; k < round.size(); k++) {
          CodecRequest* q = round[k];
          rcs[k] = call(*q, q->n, q->src, q->src_len, q->dst, q->dst_cap, q->dst_len, q->checksum_out, q->n_slices,
                        q->slice_len, q->slice_checksum, q->status, q->bad_slice);
          if (rcs[k] != 0) errs[k] = b2s_last_error();
        }
        lk.lock();
        stats_.batches += 1 + round.size();
        stats_.isolated++;
        stats_.maxMerged = std::max<uint64_t>(stats_.maxMerged, round.size());
        for (size_t k = 0; k < round.size(); k++) {
          round[k]->rc = rcs[k];
          round[k]->error = errs[k];
          round[k]->done = true;
        }
        cv_.notify_all();
        return;
      }
      at = 0;
      for (auto* q : round) {
        for (uint32_t i = 0; i < q->n; i++, at++) {
          q->dst_len[i] = dlen[at];
          q->status[i] = status[at];
          if (q->checksum_out) q->checksum_out[i] = cks[at];
          if (q->bad_slice) q->bad_slice[i] = bad[at];
        }
      }
    }
    lk.lock();
    stats_.batches++;
    stats_.maxMerged = std::max<uint64_t>(stats_.maxMerged, round.size());
    for (auto* q : round) {
      q->rc = rc;
      q->error = err;
      q->done = true;
    }
    cv_.notify_all();
  }

  static int call(const CodecRequest& k, uint32_t n, const uint8_t* const* src, const uint64_t* len, uint8_t* const* dst,
                  const uint64_t* cap, uint64_t* dlen, uint64_t* cks, const uint32_t* nsl, const uint64_t* const* sl,
                  const uint64_t* const* sc, int32_t* status, int32_t* bad) {
    if (k.op == 0)
      return b2s_compress_batch(k.codec, k.level, k.block_size, k.checksum_alg, n, src, len, dst, cap, dlen, cks, status);
    return b2s_decompress_batch(k.codec, k.checksum_alg, n, src, len, nsl, sl, sc, dst, cap, dlen, status, bad);
  }

  std::mutex m_;
  std::condition_variable cv_;
  std::deque<CodecRequest*> pending_;
  bool leader_ = false;
  Statistics stats_;
};

}  // namespace host
}  // namespace b2s
