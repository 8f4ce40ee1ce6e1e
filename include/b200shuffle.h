/*
 * b200shuffle.h — flat C ABI of libb200shuffle.so, the B200-native shuffle-block codec path.
 *
 * This is the drop-in boundary for IBM/spark-s3-shuffle's codec hot path.  The reference has no FFI of its own
 * (it is pure Scala and reaches native code only through lz4-java / snappy-java / zstd-jni / java.util.zip); each
 * entry point below names the reference interface (file:line under /root/reference/src/main/scala/org/apache/spark)
 * whose work it replaces.  INTEGRATION.md shows the JNI stub + Scala classes a maintainer adds on the JVM side.
 *
 *   write side  (map task)   : per-partition  compress  + checksum          -> b2s_compress_*
 *       replaces the CompressionCodec.compressedOutputStream + MutableCheckedOutputStream chain that feeds
 *       shuffle/S3ShuffleMapOutputWriter.scala:140-146,168-202 and whose results arrive at :91,113-115
 *   read side   (reduce task): checksum-verify + decompress per block        -> b2s_decompress_*
 *       replaces storage/S3ChecksumValidationStream.scala:54-86 and serializerManager.wrapStream at
 *       storage/S3ShuffleReader.scala:99-110
 *   checksums only                                                           -> b2s_checksum_*
 *       replaces helper/S3ShuffleHelper.scala:94-103 (ADLER32 | CRC32) and adds CRC32C
 *
 * Conventions: plain pointers and sizes, no C++ or torch types.  Every call is synchronous (returns when results
 * are in the caller's buffers) and re-entrant.  Function return: 0 = call completed (inspect per-block status[]),
 * negative B2S_E_* = the call itself failed.  The library never aborts the process and has NO CPU fallback: without
 * a usable CUDA device every compute entry point returns B2S_E_CUDA.
 *
 * "_packed" variants take/produce one contiguous arena plus offsets — the layout of a .data object
 * (concatenated per-partition streams, ascending reduceId) — so one H2D/D2H moves a whole
 * batch and dst_len[] *is* the partitionLengths array handed to helper/S3ShuffleHelper.scala:44-47.
 * "_dev" variants take device pointers for data (descriptor arrays stay on the host) and are what the
 * device-resident roofline measurement in bench.py drives.
 */
#ifndef B200SHUFFLE_H
#define B200SHUFFLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_VERSION 0x000100

/* codecs — Spark's spark.io.compression.codec short names (storage/S3ShuffleReader.scala:57-60 probes the codec) */
#define B2S_CODEC_NONE 0
#define B2S_CODEC_LZ4BLOCK 1      /* "lz4"   : lz4-java LZ4Block stream, XXH32 seed 0x9747b28c & 0x0FFFFFFF */
#define B2S_CODEC_SNAPPY_XERIAL 2 /* "snappy": xerial SnappyOutputStream framing over raw snappy */
#define B2S_CODEC_ZSTD 3          /* "zstd"  : RFC 8878 frames, concatenation allowed */

/* checksum algorithms — spark.shuffle.checksum.algorithm (helper/S3ShuffleHelper.scala:94-103) */
#define B2S_CHECKSUM_NONE 0
#define B2S_CHECKSUM_ADLER32 1
#define B2S_CHECKSUM_CRC32 2
#define B2S_CHECKSUM_CRC32C 3 /* north-star addition; the Scala shim adds the "CRC32C" case */

/* per-block status / call errors.  The Scala shim maps them to the reference's exception types:            */
#define B2S_OK 0
#define B2S_E_CORRUPT (-1)       /* IOException("Stream is corrupted")  (LZ4BlockInputStream et al. [U])   */
#define B2S_E_CHECKSUM (-2)      /* SparkException("Invalid checksum detected for <block>") S3ChecksumValidationStream.scala:72-74 */
#define B2S_E_DST_TOO_SMALL (-3) /* retry with a larger destination                                          */
#define B2S_E_UNSUPPORTED (-4)   /* UnsupportedOperationException (S3ShuffleHelper.scala:100-101)            */
#define B2S_E_ARG (-5)           /* RuntimeException("Precondition: ...")                                    */
#define B2S_E_CUDA (-6)          /* no device / CUDA failure; b2s_last_error() has the text                  */
#define B2S_E_NOT_INIT (-7)
#define B2S_E_NOMEM (-8)

/* ---- lifecycle: call once per executor, beside S3ShuffleDataIO.initializeExecutor (shuffle/S3ShuffleDataIO.scala:30-32) ---- */
/* gpu_mask: bit i selects CUDA device i (0 = all visible devices).  The per-stream-pointer calls (b2s_*_batch) shard
 * their streams round-robin over the selected devices (stream i -> device i mod D), one host thread, pinned staging and
 * streams per device, nothing exchanged between devices.  Packed and device-resident calls run on ONE device: the calling
 * thread's (b2s_set_thread_device, default 0) — an executor pins each task thread to a device, or, as bench.py does,
 * runs one process per GPU.  streams_per_gpu = pipeline slots (stream pair + staging) per lane, 2..8, 0 = default 6;
 * pinned_bytes_per_gpu = the library's own pinned descriptor blocks are allocated up front from this budget (0 = grow on
 * demand; payload staging is the caller's, b2s_host_alloc).  Every device runs two LANES — write-side calls take one,
 * read-side calls the other — so a compress and a decompress call from different threads overlap on the full-duplex
 * link.  Idempotent. */
int b2s_init(uint32_t gpu_mask, uint64_t pinned_bytes_per_gpu, uint32_t streams_per_gpu);
void b2s_shutdown(void);
int b2s_device_count(void); /* devices selected by b2s_init, or B2S_E_NOT_INIT */
int b2s_set_thread_device(uint32_t dev_index); /* device used by this thread's packed / host-pointer calls */
/* NUMA placement (SURVEY.md §8e "NUMA-pin staging buffers to the GPU's socket"; the reference's task threads are the
 * ones of storage/S3BufferedPrefetchIterator.scala:78-92 and the map task).  b2s_bind_thread_to_device = set_thread_device
 * + pin the calling thread to the CPUs of the device's NUMA node + prefer that node for its allocations (best effort;
 * a no-op returning 0 when the topology is unknown or B2S_NUMA=0).  b2s_host_alloc always places its pages next to the
 * calling thread's device.  b2s_device_numa_node: the node, -1000 if unknown. */
int b2s_bind_thread_to_device(uint32_t dev_index);
int b2s_device_numa_node(uint32_t dev_index);
const char* b2s_strerror(int32_t code);
const char* b2s_last_error(void); /* thread-local text of the last B2S_E_CUDA / B2S_E_ARG */
uint32_t b2s_version(void);

/* pinned host memory the JVM side wraps as direct ByteBuffers (replaces the byte[] of
 * storage/S3BufferedInputStreamAdaptor.scala:11 and the BufferedOutputStream of S3ShuffleMapOutputWriter.scala:46-47) */
void* b2s_host_alloc(uint64_t bytes);
void b2s_host_free(void* p);
int b2s_host_register(void* p, uint64_t bytes);
int b2s_host_unregister(void* p);

/* ---- sizing ---- */
/* worst-case compressed size of one stream of src_len bytes (codec_block_size 0 = Spark default 32 KiB) */
uint64_t b2s_compress_bound(uint32_t codec, uint32_t codec_block_size, uint64_t src_len);
/* decompressed size of each compressed stream (LZ4Block: sum of originalLen; Snappy: sum of chunk varints) */
int b2s_decompressed_size_batch(uint32_t codec, uint32_t n, const uint8_t* const* src, const uint64_t* src_len,
                                uint64_t* out_len, int32_t* status);

/* ---- checksums: one value per slice of bytes (low 32 bits significant, as in the .checksum file) ---- */
int b2s_checksum_batch(uint32_t alg, uint32_t n, const uint8_t* const* src, const uint64_t* len, uint64_t* out);
int b2s_checksum_packed(uint32_t alg, uint32_t n, const uint8_t* base, const uint64_t* off, const uint64_t* len,
                        uint64_t* out);

/* ---- write side ---- */
/* n independent streams (one per (map,reduce) partition).  Each output is a complete, self-terminated stream in the
 * codec's JVM wire format; checksum_out[i] (if checksum_alg != 0) is over the *compressed* bytes of stream i. */
int b2s_compress_batch(uint32_t codec, int32_t level, uint32_t codec_block_size, uint32_t checksum_alg, uint32_t n,
                       const uint8_t* const* src, const uint64_t* src_len, uint8_t* const* dst,
                       const uint64_t* dst_cap, uint64_t* dst_len, uint64_t* checksum_out, int32_t* status);
/* packed: stream i is src_base[src_off[i] .. +src_len[i]); outputs are written back to back into dst_base in index
 * order: dst_off[i], dst_len[i].  *dst_total receives the arena bytes used. */
int b2s_compress_packed(uint32_t codec, int32_t level, uint32_t codec_block_size, uint32_t checksum_alg, uint32_t n,
                        const uint8_t* src_base, const uint64_t* src_off, const uint64_t* src_len, uint8_t* dst_base,
                        uint64_t dst_cap, uint64_t* dst_off, uint64_t* dst_len, uint64_t* dst_total,
                        uint64_t* checksum_out, int32_t* status);

/* ---- read side ---- */
/* n compressed blocks as produced by S3BufferedPrefetchIterator.next() (storage/S3BufferedPrefetchIterator.scala:196-212).
 * Block i covers n_slices[i] consecutive reduce partitions (1 for ShuffleBlockId, >1 for ShuffleBlockBatchId,
 * storage/S3ChecksumValidationStream.scala:22-27); slice_len[i][k] / slice_checksum[i][k] are the .index differences and
 * .checksum values of those partitions.  With checksum_alg != 0 every slice is verified over the compressed bytes
 * before decoding; a mismatch yields status[i] = B2S_E_CHECKSUM and bad_slice[i] = k (may be NULL).
 * n_slices / slice_* may be NULL when checksum_alg == 0. */
int b2s_decompress_batch(uint32_t codec, uint32_t checksum_alg, uint32_t n, const uint8_t* const* src,
                         const uint64_t* src_len, const uint32_t* n_slices, const uint64_t* const* slice_len,
                         const uint64_t* const* slice_checksum, uint8_t* const* dst, const uint64_t* dst_cap,
                         uint64_t* dst_len, int32_t* status, int32_t* bad_slice);
/* packed: slices are flattened — block i owns slice_len[slice_base[i] .. slice_base[i+1]) */
int b2s_decompress_packed(uint32_t codec, uint32_t checksum_alg, uint32_t n, const uint8_t* src_base,
                          const uint64_t* src_off, const uint64_t* src_len, const uint32_t* slice_base,
                          const uint64_t* slice_len, const uint64_t* slice_checksum, uint8_t* dst_base,
                          uint64_t dst_cap, uint64_t* dst_off, uint64_t* dst_len, uint64_t* dst_total,
                          int32_t* status, int32_t* bad_slice);

/* ---- device-resident variants (data pointers are device memory on device `dev_index` of the b2s_init selection;
 *      descriptor arrays are host memory).  Synchronous; timings retrievable with b2s_last_timing. ---- */
int b2s_checksum_dev(uint32_t dev_index, uint32_t alg, uint32_t n, const void* d_base, const uint64_t* off,
                     const uint64_t* len, uint64_t* out);
int b2s_compress_dev(uint32_t dev_index, uint32_t codec, int32_t level, uint32_t codec_block_size,
                     uint32_t checksum_alg, uint32_t n, const void* d_src_base, const uint64_t* src_off,
                     const uint64_t* src_len, void* d_dst_base, uint64_t dst_cap, uint64_t* dst_off,
                     uint64_t* dst_len, uint64_t* dst_total, uint64_t* checksum_out, int32_t* status);
int b2s_decompress_dev(uint32_t dev_index, uint32_t codec, uint32_t checksum_alg, uint32_t n, const void* d_src_base,
                       const uint64_t* src_off, const uint64_t* src_len, const uint32_t* slice_base,
                       const uint64_t* slice_len, const uint64_t* slice_checksum, void* d_dst_base, uint64_t dst_cap,
                       uint64_t* dst_off, uint64_t* dst_len, uint64_t* dst_total, int32_t* status,
                       int32_t* bad_slice);
void* b2s_dev_alloc(uint32_t dev_index, uint64_t bytes);
void b2s_dev_free(uint32_t dev_index, void* p);
int b2s_dev_memcpy(uint32_t dev_index, void* dst, const void* src, uint64_t bytes, int kind /*1=H2D 2=D2H 3=D2D*/);

/* ---- observability (the reference logs bytes/ms/MiB/s per block: shuffle/S3MeasureOutputStream.scala:55-63,
 *      storage/S3BufferedPrefetchIterator.scala:155-186) ---- */
typedef struct b2s_timing {
  double total_ms;        /* host wall time of the last call on this thread */
  double h2d_ms, d2h_ms;  /* CUDA-event time of the copies (0 for _dev calls) */
  double kernel_ms;       /* CUDA-event time from first to last kernel of the call */
  double top_kernel_ms;   /* CUDA-event time of the codec step (lz4 match+parse+emit / decompress / checksum) */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t kernel_launches; /* kernels launched by the call */
  uint64_t src_bytes, dst_bytes;
  double dominant_ms;       /* CUDA-event time of the single dominant kernel (lz4_match_kernel / lz4 decode), summed over its launches */
  uint64_t dominant_launches;
} b2s_timing;
int b2s_last_timing(b2s_timing* out);
/* CUDA-event stopwatch on the library's own stream of device dev_index (torch.cuda.Event only sees torch's stream):
 * b2s_mark(dev, 0) ... calls ... b2s_mark(dev, 1); b2s_marks_elapsed_ms(dev, &ms) synchronises on mark 1 */
int b2s_mark(uint32_t dev_index, uint32_t which);
int b2s_marks_elapsed_ms(uint32_t dev_index, double* ms);
uint64_t b2s_total_kernel_launches(void); /* process-wide counter since b2s_init */

/* ---- synthetic workload generator for the benchmark (device-side TeraGen-style 104-byte records; not on the
 *      product path).  Writes n_records*104 bytes at d_dst. ---- */
int b2s_gen_terasort_dev(uint32_t dev_index, void* d_dst, uint64_t first_record, uint64_t n_records, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
