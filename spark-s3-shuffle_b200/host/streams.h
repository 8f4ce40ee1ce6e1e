// streams.h — the stream classes of the reference's write and read path, C++ mirrors (internal to shuffle_host.cpp:
// included after the exception types and BlockId are defined).
//
//   S3MeasureOutputStream         shuffle/S3MeasureOutputStream.scala:8-65
//   S3ShuffleBlockStream          storage/S3ShuffleBlockStream.scala:16-111
//   S3BufferedInputStreamAdaptor  storage/S3BufferedInputStreamAdaptor.scala:7-59
//   S3BufferedPrefetchIterator    storage/S3BufferedPrefetchIterator.scala:16-213 (ThreadPredictor :29-67)
//
// What differs from the reference, on purpose:
//   * nextBatch(): SURVEY.md §8(f)-2 — the consumer drains every block that is complete at that moment in one call,
//     so the reader can hand K blocks to one b2s_decompress_batch instead of pulling them one by one.
//   * a failure inside a prefetch thread is carried to the consumer and rethrown by next()/nextBatch(); in the
//     reference the thread dies and the consumer waits forever.
//   * the iterator can be destroyed early (task cancellation): threads are joined.
#pragma once

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <exception>
#include <functional>
#include <list>
#include <numeric>
#include <thread>

namespace b2s {
namespace host {

static inline int64_t nanoTime() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

// ---- S3MeasureOutputStream -----------------------------------------------------------------------------------
// Wall time spent inside write/flush/close of the wrapped stream plus a byte counter; close() produces the
// reference's log line.  The wrapped stream is the .data object behind a BufferedOutputStream(bufferSize)
// (shuffle/S3ShuffleMapOutputWriter.scala:43-49): a FILE* with a bufferSize stdio buffer.
class S3MeasureOutputStream {
 public:
  S3MeasureOutputStream(const std::string& path, std::string label, size_t bufferSize) : label_(std::move(label)) {
    out_ = fopen(path.c_str(), "wb");
    if (!out_) throw IOException("cannot create " + path);
    if (bufferSize) setvbuf(out_, nullptr, _IOFBF, bufferSize);
  }
  ~S3MeasureOutputStream() {
    if (out_) fclose(out_);
  }
  void write(const uint8_t* b, uint64_t len) {  // :36-42
    checkOpen();
    const int64_t now = nanoTime();
    if (len && fwrite(b, 1, len, out_) != len) throw IOException("short write on " + label_);
    timings_ += nanoTime() - now;
    bytes_ += (int64_t)len;
  }
  void flush() {  // :44-49
    checkOpen();
    const int64_t now = nanoTime();
    fflush(out_);
    timings_ += nanoTime() - now;
  }
  void close() {  // :51-64
    if (!isOpen_) return;
    const int64_t now = nanoTime();
    fflush(out_);
    fclose(out_);
    timings_ += nanoTime() - now;
    out_ = nullptr;
    isOpen_ = false;
    const int64_t t = timings_ / 1000000;
    const double bw = (double)bytes_ / ((double)t / 1000) / (1024 * 1024);
    std::ostringstream o;
    o << "Statistics: Stage 0.0 TID 0 -- Writing " << label_ << " " << bytes_ << " took " << t << " ms (" << bw
      << " MiB/s)";
    statistics_ = o.str();
  }
  int64_t timings() const { return timings_; }
  int64_t bytes() const { return bytes_; }
  const std::string& statistics() const { return statistics_; }

 private:
  void checkOpen() const {
    if (!isOpen_) throw IOException("The stream is already closed!");  // :17-21
  }
  FILE* out_ = nullptr;
  std::string label_, statistics_;
  bool isOpen_ = true;
  int64_t timings_ = 0, bytes_ = 0;
};

// ---- S3ShuffleBlockStream ------------------------------------------------------------------------------------
// The byte range [accumulatedPositions(startReduceId), accumulatedPositions(endReduceId)) of one .data object, read
// with positioned readFully.  The object is opened lazily (:25-34) and closed when the range is exhausted (:83-85).
class S3ShuffleBlockStream {
 public:
  S3ShuffleBlockStream(std::string dataPath, int64_t startPosition, int64_t endPosition)
      : maxBytes(endPosition - startPosition), path_(std::move(dataPath)), startPosition_(startPosition),
        streamClosed_(startPosition == endPosition) {}  // :36-40
  ~S3ShuffleBlockStream() { close(); }
  S3ShuffleBlockStream(const S3ShuffleBlockStream&) = delete;

  const int64_t maxBytes;

  void close() {  // :45-52
    if (streamClosed_) return;
    if (fd_ >= 0) ::close(fd_);
    fd_ = -1;
    streamClosed_ = true;
  }
  // :73-92 — returns the number of bytes read, or -1 at the end of the range.  As in the reference, an I/O error
  // closes the stream and reads as end-of-stream; the truncated block then fails in the codec (corrupt stream) or in
  // the checksum validation, which is where the JVM reader would notice it too.
  int64_t read(uint8_t* b, int64_t len) {
    if (streamClosed_ || numBytes_ >= maxBytes) return -1;
    const int64_t length = std::min(maxBytes - numBytes_, len);
    if (fd_ < 0) {
      fd_ = ::open(path_.c_str(), O_RDONLY);
      if (fd_ < 0) throw IOException("File does not exist: " + path_);  // dispatcher.openBlock (:27-33) rethrows
    }
    int64_t done = 0;
    while (done < length) {  // readFully
      const ssize_t k = ::pread(fd_, b + done, (size_t)(length - done), (off_t)(startPosition_ + numBytes_ + done));
      if (k <= 0) {
        close();
        return -1;
      }
      done += k;
    }
    numBytes_ += length;
    if (numBytes_ >= maxBytes) close();
    return length;
  }
  int64_t skip(int64_t n) {  // :94-102
    if (streamClosed_ || numBytes_ >= maxBytes || n <= 0) return 0;
    const int64_t toSkip = std::min(maxBytes - numBytes_, n);
    numBytes_ += toSkip;
    return toSkip;
  }
  int64_t available() const { return streamClosed_ ? 0 : maxBytes - numBytes_; }  // :104-109

 private:
  std::string path_;
  int64_t startPosition_;
  bool streamClosed_;
  int64_t numBytes_ = 0;
  int fd_ = -1;
};

// ---- S3BufferedInputStreamAdaptor ----------------------------------------------------------------------------
// Owns the block's buffer of bufferSize bytes, filled by ONE read at construction (prefill, :13-21); close() hands
// bufferSize back to the prefetcher's budget (:49-58).  Bytes past bufferSize (a block larger than the task's budget)
// are read through from the block stream, as BufferedInputStream does once its buffer is drained.
class S3BufferedInputStreamAdaptor {
 public:
  S3BufferedInputStreamAdaptor(std::unique_ptr<S3ShuffleBlockStream> in, int64_t bufferSize,
                               std::function<void(int64_t)> onClose)
      : in_(std::move(in)), bufferSize_(bufferSize), onClose_(std::move(onClose)) {
    buf_.reset(new uint8_t[(size_t)std::max<int64_t>(bufferSize, 1)]);
    const int64_t k = in_->read(buf_.get(), bufferSize);  // prefill
    count_ = k > 0 ? k : 0;
  }
  ~S3BufferedInputStreamAdaptor() {
    if (buf_) close();
  }
  int64_t read(uint8_t* b, int64_t len) {  // :34-37
    checkOpen();
    if (pos_ < count_) {
      const int64_t k = std::min(len, count_ - pos_);
      memcpy(b, buf_.get() + pos_, (size_t)k);
      pos_ += k;
      return k;
    }
    return in_->read(b, len);
  }
  int64_t available() const { return (count_ - pos_) + in_->available(); }
  // zero-copy view of the prefilled bytes, for the batch decoder
  const uint8_t* buffered() const { return buf_.get(); }
  int64_t bufferedBytes() const { return count_; }
  int64_t totalBytes() const { return in_->maxBytes; }
  int64_t bufferSize() const { return bufferSize_; }
  bool doubleCloseSeen() const { return doubleClose_; }
  void close() {  // :49-58
    if (!buf_) {
      doubleClose_ = true;  // "Double close detected. Ignoring."
      return;
    }
    in_->close();
    buf_.reset();
    onClose_(bufferSize_);
  }

 private:
  void checkOpen() const {
    if (!buf_) throw IOException("Stream is closed");  // EOFException (:23-27)
  }
  std::unique_ptr<S3ShuffleBlockStream> in_;
  int64_t bufferSize_;
  std::function<void(int64_t)> onClose_;
  std::unique_ptr<uint8_t[]> buf_;
  int64_t count_ = 0, pos_ = 0;
  bool doubleClose_ = false;
};

// ---- S3BufferedPrefetchIterator ------------------------------------------------------------------------------
class S3BufferedPrefetchIterator {
 public:
  struct Source {  // one element of the wrapped iterator: (BlockId, S3ShuffleBlockStream) + the caller's tag
    BlockId id;
    std::unique_ptr<S3ShuffleBlockStream> stream;
    size_t tag;
  };
  struct Fetched {  // (BlockId, InputStream)
    BlockId id;
    std::unique_ptr<S3BufferedInputStreamAdaptor> stream;
    size_t tag;
    std::exception_ptr error;
  };
  struct Statistics {
    int64_t totalRuntime = 0, timeWaiting = 0, timePrefetching = 0, numStreams = 0, bytesRead = 0,
            activeThreads = 0, peakMemoryUsage = 0, threadsStarted = 0, peakThreads = 0;
    std::string line;
  };

  S3BufferedPrefetchIterator(std::deque<Source> iter, int64_t maxBufferSize, int maxConcurrencyTask)
      : iter_(std::move(iter)), maxBufferSize_(maxBufferSize), startTime_(nanoTime()), hasItem_(!iter_.empty()),
        threadPredictor_(maxConcurrencyTask < 1 ? 1 : maxConcurrencyTask) {
    std::unique_lock<std::mutex> lk(mon_);
    configureThreads(-1, lk);  // :102-103 make sure that there's at least a single thread running
  }
  ~S3BufferedPrefetchIterator() {
    {
      std::unique_lock<std::mutex> lk(mon_);
      stop_ = true;
      cv_.notify_all();
    }
    for (auto& t : threads_) t.join();
    completed_.clear();  // closes whatever the consumer never took (needs mon_ unlocked: onCloseStream locks)
  }

  bool hasNext() {  // :176-182
    std::unique_lock<std::mutex> lk(mon_);
    const bool result = hasItem_ || activeTasks_ > 0 || !completed_.empty();
    if (!result) printStatistics();
    return result;
  }
  Fetched next() {  // :184-203
    std::vector<Fetched> v = take(1);
    return std::move(v[0]);
  }
  // §8(f)-2: every block that is complete right now (at least one, at most maxBlocks)
  std::vector<Fetched> nextBatch(size_t maxBlocks) { return take(maxBlocks ? maxBlocks : (size_t)-1); }

  Statistics statistics() {
    std::unique_lock<std::mutex> lk(mon_);
    Statistics s = stats_;
    s.totalRuntime = nanoTime() - startTime_;
    s.timeWaiting = timeWaiting_;
    s.timePrefetching = timePrefetching_;
    s.numStreams = numStreams_;
    s.bytesRead = bytesRead_;
    s.activeThreads = desiredActiveThreads_.load();
    s.peakMemoryUsage = peakMemoryUsage_;
    s.threadsStarted = (int64_t)threads_.size();
    s.peakThreads = peakThreads_;
    return s;
  }

 private:
  // :29-67 — hill climbing on the consumer's wait time: 20 measurements per step; move towards the neighbouring
  // thread count whose recorded total was lower.
  class ThreadPredictor {
   public:
    explicit ThreadPredictor(int maxThreads) : latencies_((size_t)maxThreads + 2, 0), measurementsNS_(20, 0) {
      latencies_[0] = INT64_MAX;
      latencies_[(size_t)maxThreads + 1] = INT64_MAX;
    }
    int addMeasurementAndPredict(int64_t latencyNS) {  // :60-66
      if (latencyNS >= 0) {
        measurementsNS_[(size_t)numMeasurements_ % measurementsNS_.size()] = latencyNS;
        numMeasurements_ += 1;
      }
      return predict();
    }

   private:
    int predict() {  // :38-58
      if (numMeasurements_ < (int)measurementsNS_.size() + currentThreads_) return currentThreads_;
      const int64_t current = std::accumulate(measurementsNS_.begin(), measurementsNS_.end(), (int64_t)0);
      if (current < 500) return currentThreads_;  // less than 25 ns latency for each request
      latencies_[(size_t)currentThreads_] = current;
      const int64_t prevValue = latencies_[(size_t)currentThreads_ - 1];
      const int64_t nextValue = latencies_[(size_t)currentThreads_ + 1];
      numMeasurements_ = 0;
      if (prevValue < current) currentThreads_ -= 1;
      else if (nextValue < current) currentThreads_ += 1;
      return currentThreads_;
    }
    int currentThreads_ = 1;
    std::vector<int64_t> latencies_, measurementsNS_;
    int numMeasurements_ = 0;
  };

  // :78-92 (caller holds mon_)
  void configureThreads(int64_t latency, std::unique_lock<std::mutex>&) {
    if (desiredActiveThreads_.load() != currentActiveThreads_.load()) return;
    const int64_t nThreads = threadPredictor_.addMeasurementAndPredict(latency);
    const int64_t activeThreads = desiredActiveThreads_.exchange(nThreads);
    if (nThreads > activeThreads && !stop_) {
      pendingStarts_++;
      threads_.emplace_back([this, nThreads] { prefetchThread(nThreads); });
    }
  }

  void onCloseStream(int64_t bufferSize) {  // :96-100
    std::unique_lock<std::mutex> lk(mon_);
    memoryUsage_ -= bufferSize;
    cv_.notify_all();
  }

  void prefetchThread(int64_t threadId) {  // :102-160
    {
      std::unique_lock<std::mutex> lk(mon_);
      pendingStarts_--;
      const int64_t now = currentActiveThreads_.fetch_add(1) + 1;
      peakThreads_ = std::max(peakThreads_, now);
    }
    bool have = false;
    Source nextElement;
    while (true) {
      {
        std::unique_lock<std::mutex> lk(mon_);
        if (stop_ || (iter_.empty() && !have)) {
          if (!stop_) hasItem_ = false;
          currentActiveThreads_.fetch_sub(1);
          return;
        }
        if (!have) {
          if (threadId > desiredActiveThreads_.load()) {  // scaled down: the highest ids leave
            currentActiveThreads_.fetch_sub(1);
            return;
          }
          nextElement = std::move(iter_.front());
          iter_.pop_front();
          have = true;
          activeTasks_ += 1;
          hasItem_ = !iter_.empty();
        }
      }
      bool fetchNext = false;
      const int64_t bsize = std::min(maxBufferSize_, nextElement.stream->maxBytes);
      {
        std::unique_lock<std::mutex> lk(mon_);
        if (memoryUsage_ + bsize > maxBufferSize_) {
          if (!stop_) cv_.wait(lk);
        } else {
          fetchNext = true;
          memoryUsage_ += bsize;
          peakMemoryUsage_ = std::max(peakMemoryUsage_, memoryUsage_);
        }
      }
      if (fetchNext) {
        Fetched f;
        f.id = nextElement.id;
        f.tag = nextElement.tag;
        have = false;
        const int64_t now = nanoTime();
        try {
          f.stream.reset(new S3BufferedInputStreamAdaptor(std::move(nextElement.stream), bsize,
                                                          [this](int64_t n) { onCloseStream(n); }));
        } catch (...) {
          f.error = std::current_exception();
        }
        const int64_t dt = nanoTime() - now;
        std::unique_lock<std::mutex> lk(mon_);
        if (f.error) memoryUsage_ -= bsize;
        timePrefetching_ += dt;
        bytesRead_ += bsize;
        completed_.push_front(std::move(f));  // LinkedList.push: LIFO (:146)
        activeTasks_ -= 1;
        cv_.notify_all();
      }
    }
  }

  std::vector<Fetched> take(size_t maxBlocks) {
    std::unique_lock<std::mutex> lk(mon_);
    const int64_t now = nanoTime();
    while (completed_.empty()) {
      if (!(hasItem_ || activeTasks_ > 0)) throw RuntimeException("next on empty iterator");
      cv_.wait(lk);
    }
    const int64_t latency = nanoTime() - now;
    configureThreads(latency, lk);
    timeWaiting_ += latency;
    std::vector<Fetched> out;
    while (!completed_.empty() && out.size() < maxBlocks) {
      out.push_back(std::move(completed_.front()));  // pop (:209)
      completed_.pop_front();
      numStreams_ += 1;
    }
    cv_.notify_all();
    lk.unlock();
    for (auto& f : out)
      if (f.error) std::rethrow_exception(f.error);
    return out;
  }

  void printStatistics() {  // :162-192 (caller holds mon_)
    const int64_t totalRuntime = nanoTime() - startTime_;
    std::ostringstream o;
    if (numStreams_ == 0) {
      o << "Unable to print statistics: / by zero.";
    } else {
      const int64_t tR = totalRuntime / 1000000, wPer = totalRuntime ? 100 * timeWaiting_ / totalRuntime : 0;
      const int64_t tW = timeWaiting_ / 1000000, tP = timePrefetching_ / 1000000, bR = bytesRead_, r = numStreams_;
      const int64_t atP = tP / r, atW = tW / r, bs = bR / r;
      const double bW = (double)bR / ((double)tP / 1000) / (1024 * 1024);
      o << "Statistics: Stage 0.0 TID 0 -- " << bR << " bytes, " << tW << " ms waiting (" << atW << " avg), " << tP
        << " ms prefetching (avg: " << atP << " ms - " << bs << " block size - " << bW << " MiB/s). Total: " << tR
        << " ms - " << wPer << "% waiting. " << desiredActiveThreads_.load() << " active threads.";
    }
    stats_.line = o.str();
  }

  std::mutex mon_;
  std::condition_variable cv_;
  std::deque<Source> iter_;
  const int64_t maxBufferSize_;
  const int64_t startTime_;
  int64_t memoryUsage_ = 0, peakMemoryUsage_ = 0;
  bool hasItem_;
  int64_t timeWaiting_ = 0, timePrefetching_ = 0, numStreams_ = 0, bytesRead_ = 0, activeTasks_ = 0;
  std::list<Fetched> completed_;
  ThreadPredictor threadPredictor_;
  std::atomic<int64_t> currentActiveThreads_{0}, desiredActiveThreads_{0};
  int64_t pendingStarts_ = 0, peakThreads_ = 0;
  bool stop_ = false;
  std::vector<std::thread> threads_;
  Statistics stats_;
};

}  // namespace host
}  // namespace b2s
