// codec_adapter.h — SURVEY.md §8(f)-1: the Spark CompressionCodec seam on the same C ABI (internal to
// shuffle_host.cpp).  Mirrors org.apache.spark.io.CompressionCodec [U]:
//     compressedOutputStream(s: OutputStream): OutputStream      compressedInputStream(s: InputStream): InputStream
// which is where BypassMergeSortShuffleWriter / UnsafeShuffleWriter compress *before* the plugin's writer sees bytes
//  and where the reader wraps every block (storage/S3ShuffleReader.scala:107-109).
//
// A GPU codec cannot work in 32 KiB pulls, so both streams are whole-buffer adapters:
//   output: bytes are collected in a pinned arena; every `flushBytes` (and at close) the collected bytes become ONE
//           complete stream in the codec's JVM wire format, made by one b2s_compress_packed call.  Several such
//           streams back to back are legal input for all three JVM readers — LZ4BlockInputStream is opened by Spark
//           with concatenation enabled, SnappyInputStream accepts a re-occurring header, ZstdInputStream runs
//           continuous — which is exactly what supportsConcatenationOfSerializedStreams (the capability the reference
//           asks for at storage/S3ShuffleReader.scala:57-60) promises.
//   input:  the first read() drains the source, sizes and decodes it with one b2s_decompress_packed call, and serves
//           reads from the decoded arena.
#pragma once

namespace b2s {
namespace host {

using SinkFn = std::function<void(const uint8_t*, uint64_t)>;
using SourceFn = std::function<int64_t(uint8_t*, uint64_t)>;  // bytes read; <= 0 at the end

class B200CompressionCodec {
 public:
  B200CompressionCodec(int codecId, uint32_t blockSize, uint64_t flushBytes)
      : codec_(codecId), blockSize_(blockSize), flushBytes_(flushBytes < blockSize ? blockSize : flushBytes) {  // at least one codec block: 0 (an unparsable size) would spin write() forever
    if (codec_ == B2S_CODEC_NONE) throw UnsupportedOperationException("no compression codec configured");
  }
  int codecId() const { return codec_; }
  uint32_t blockSize() const { return blockSize_; }
  uint64_t flushBytes() const { return flushBytes_; }
  // CompressionCodec.supportsConcatenationOfSerializedStreams [U]: true for lz4, snappy, zstd (and lzf)
  static bool supportsConcatenationOfSerializedStreams(int) { return true; }

 private:
  int codec_;
  uint32_t blockSize_;
  uint64_t flushBytes_;
};

class B200CompressedOutputStream {
 public:
  B200CompressedOutputStream(const B200CompressionCodec& c, SinkFn sink) : c_(c), sink_(std::move(sink)) {}
  void write(const uint8_t* b, uint64_t n) {
    if (closed_) throw IOException("Stream is closed");
    while (n) {
      const uint64_t room = c_.flushBytes() - in_.size();
      const uint64_t k = n < room ? n : room;
      in_.append(b, k);
      b += k;
      n -= k;
      if (in_.size() >= c_.flushBytes()) emit();
    }
  }
  // flush(): a GPU stream has no cheap partial flush; bytes become visible downstream at flushBytes or close, like
  // LZ4BlockOutputStream without syncFlush (Spark passes syncFlush=false [U]).
  void flush() {
    if (closed_) throw IOException("Stream is closed");
  }
  void close() {
    if (closed_) return;
    if (in_.size() || !emitted_) emit();  // an empty stream still has to be a valid, self-terminated stream
    closed_ = true;
  }
  uint64_t bytesIn() const { return bytesIn_; }
  uint64_t bytesOut() const { return bytesOut_; }
  uint32_t streamsEmitted() const { return emitted_; }

 private:
  void emit() {
    int rc = b2s_init(0, 0, 0);
    if (rc != 0) throw CodecException(std::string("b2s_init: ") + b2s_strerror(rc) + ": " + b2s_last_error());
    uint64_t off = 0, len = in_.size(), doff = 0, dlen = 0, total = 0;
    int32_t status = 0;
    const uint64_t bound = b2s_compress_bound((uint32_t)c_.codecId(), c_.blockSize(), len);
    out_.resize(bound);
    if (!in_.size()) in_.reserve(1);
    rc = b2s_compress_packed((uint32_t)c_.codecId(), 0, c_.blockSize(), 0, 1, in_.data(), &off, &len, out_.data(),
                             bound, &doff, &dlen, &total, nullptr, &status);
    if (rc != 0) throw CodecException(std::string("b2s_compress_packed: ") + b2s_strerror(rc) + ": " + b2s_last_error());
    if (status != 0) throw IOException(std::string("compress failed: ") + b2s_strerror(status));
    sink_(out_.data() + doff, dlen);
    bytesIn_ += len;
    bytesOut_ += dlen;
    emitted_++;
    in_.clear();
  }
  const B200CompressionCodec c_;  // by value: a stream may outlive the codec object that made it
  SinkFn sink_;
  PinnedArena in_, out_;
  bool closed_ = false;
  uint32_t emitted_ = 0;
  uint64_t bytesIn_ = 0, bytesOut_ = 0;
};

class B200CompressedInputStream {
 public:
  B200CompressedInputStream(const B200CompressionCodec& c, SourceFn source) : c_(c), source_(std::move(source)) {}
  // InputStream.read(b, off, len): bytes read, or -1 at the end of the stream
  int64_t read(uint8_t* b, uint64_t len) {
    if (closed_) throw IOException("Stream is closed");
    if (!decoded_) decode();
    if (pos_ >= total_) return -1;
    const uint64_t k = std::min<uint64_t>(len, total_ - pos_);
    memcpy(b, out_.data() + pos_, k);
    pos_ += k;
    return (int64_t)k;
  }
  int64_t available() {
    if (!decoded_) decode();
    return (int64_t)(total_ - pos_);
  }
  void close() { closed_ = true; }

 private:
  void decode() {
    decoded_ = true;
    while (true) {  // drain the source
      if (in_.size() == in_.capacity()) in_.reserve(in_.capacity() ? in_.capacity() * 2 : (1u << 20));
      const uint64_t at = in_.size();
      const int64_t k = source_(in_.data() + at, in_.capacity() - at);
      if (k <= 0) break;
      in_.resize(at + (uint64_t)k);
    }
    if (!in_.size()) return;  // an empty source decodes to nothing
    int rc = b2s_init(0, 0, 0);
    if (rc != 0) throw CodecException(std::string("b2s_init: ") + b2s_strerror(rc) + ": " + b2s_last_error());
    const uint8_t* p = in_.data();
    uint64_t len = in_.size(), olen = 0, off = 0, doff = 0, dlen = 0;
    int32_t status = 0, bad = 0;
    rc = b2s_decompressed_size_batch((uint32_t)c_.codecId(), 1, &p, &len, &olen, &status);
    if (rc != 0) throw CodecException(std::string("b2s_decompressed_size_batch: ") + b2s_strerror(rc) + ": " + b2s_last_error());
    if (status != 0) throw IOException("Stream is corrupted");
    out_.resize(olen ? olen : 1);
    rc = b2s_decompress_packed((uint32_t)c_.codecId(), 0, 1, in_.data(), &off, &len, nullptr, nullptr, nullptr,
                               out_.data(), olen ? olen : 1, &doff, &dlen, &total_, &status, &bad);
    if (rc != 0) throw CodecException(std::string("b2s_decompress_packed: ") + b2s_strerror(rc) + ": " + b2s_last_error());
    if (status != 0) throw IOException("Stream is corrupted");
    total_ = dlen;
  }
  const B200CompressionCodec c_;
  SourceFn source_;
  PinnedArena in_, out_;
  bool decoded_ = false, closed_ = false;
  uint64_t total_ = 0, pos_ = 0;
};

}  // namespace host
}  // namespace b2s
