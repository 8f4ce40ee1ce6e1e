/*
 * b200shuffle_jni.c — the thin JNI layer between the Scala side of IBM/spark-s3-shuffle and libb200shuffle.so.
 *
 * One native method per C entry point a JVM needs (include/b200shuffle.h); the Scala object that declares them is
 * org.apache.spark.shuffle.gpu.B200Codec.  Call sites on the reference side:
 *   init / shutdown / bindThreadToDevice   shuffle/S3ShuffleDataIO.scala:30-32 (initializeExecutor), task threads
 *   compressPacked                         shuffle/S3ShuffleMapOutputWriter.scala:91-118 (commitAllPartitions)
 *   decompressPacked / decompressedSize    storage/S3ShuffleReader.scala:98-110 (the codec seam, batched drain of
 *                                          storage/S3BufferedPrefetchIterator.scala:196-212)
 *   checksumPacked                         shuffle/S3SingleSpillShuffleMapOutputWriter.scala:54-63,
 *                                          helper/S3ShuffleHelper.scala:94-103
 * Payload travels in pinned direct ByteBuffers (hostAlloc/wrap): the methods take raw addresses (jlong), so no
 * jbyteArray is ever pinned or copied; the small descriptor arrays (offsets, lengths, checksums, status) are Java
 * long[]/int[] accessed with Get/ReleasePrimitiveArrayCritical — every Get is paired with a Release on every path.
 *
 * Build (reference side):  cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *                              jni/b200shuffle_jni.c -L. -lb200shuffle -o libb200shuffle_jni.so
 * Here (no JDK): tests/test_jni_shim.py compiles this file against jni/stub/jni.h with -Wall -Werror, so signatures,
 * argument order against b200shuffle.h and the Get/Release pairing are at least compiler- and script-checked.
 */
#include <jni.h>
#include <stdint.h>

#include "b200shuffle.h"

#define J(name) Java_org_apache_spark_shuffle_gpu_B200Codec_##name
#define PTR(T, a) ((T*)(uintptr_t)(a))

/* critical-section helper: NULL arrays are allowed where the C ABI allows NULL */
static void* crit_get(JNIEnv* e, jarray a) { return a ? (*e)->GetPrimitiveArrayCritical(e, a, 0) : 0; }
static void crit_put(JNIEnv* e, jarray a, void* p, jint mode) {
  if (a && p) (*e)->ReleasePrimitiveArrayCritical(e, a, p, mode);
}

/* ---- lifecycle ---- */
JNIEXPORT jint JNICALL J(init)(JNIEnv* e, jclass c, jint gpuMask, jlong pinnedBytesPerGpu, jint streamsPerGpu) {
  (void)e; (void)c;
  return b2s_init((uint32_t)gpuMask, (uint64_t)pinnedBytesPerGpu, (uint32_t)streamsPerGpu);
}
JNIEXPORT void JNICALL J(shutdown)(JNIEnv* e, jclass c) {
  (void)e; (void)c;
  b2s_shutdown();
}
JNIEXPORT jint JNICALL J(deviceCount)(JNIEnv* e, jclass c) {
  (void)e; (void)c;
  return b2s_device_count();
}
JNIEXPORT jint JNICALL J(setThreadDevice)(JNIEnv* e, jclass c, jint dev) {
  (void)e; (void)c;
  return b2s_set_thread_device((uint32_t)dev);
}
JNIEXPORT jint JNICALL J(bindThreadToDevice)(JNIEnv* e, jclass c, jint dev) {
  (void)e; (void)c;
  return b2s_bind_thread_to_device((uint32_t)dev);
}
JNIEXPORT jstring JNICALL J(lastError)(JNIEnv* e, jclass c) {
  (void)c;
  return (*e)->NewStringUTF(e, b2s_last_error());
}
JNIEXPORT jstring JNICALL J(strerror)(JNIEnv* e, jclass c, jint code) {
  (void)c;
  return (*e)->NewStringUTF(e, b2s_strerror((int32_t)code));
}

/* ---- pinned host memory, handed to the JVM as direct ByteBuffers ---- */
JNIEXPORT jlong JNICALL J(hostAlloc)(JNIEnv* e, jclass c, jlong bytes) {
  (void)e; (void)c;
  return (jlong)(uintptr_t)b2s_host_alloc((uint64_t)bytes);
}
JNIEXPORT void JNICALL J(hostFree)(JNIEnv* e, jclass c, jlong addr) {
  (void)e; (void)c;
  b2s_host_free(PTR(void, addr));
}
JNIEXPORT jobject JNICALL J(wrap)(JNIEnv* e, jclass c, jlong addr, jlong bytes) {
  (void)c;
  return (*e)->NewDirectByteBuffer(e, PTR(void, addr), bytes);
}
JNIEXPORT jlong JNICALL J(addressOf)(JNIEnv* e, jclass c, jobject directBuffer) {
  (void)c;
  return (jlong)(uintptr_t)(*e)->GetDirectBufferAddress(e, directBuffer);
}
JNIEXPORT jint JNICALL J(hostRegister)(JNIEnv* e, jclass c, jlong addr, jlong bytes) {
  (void)e; (void)c;
  return b2s_host_register(PTR(void, addr), (uint64_t)bytes);
}
JNIEXPORT jint JNICALL J(hostUnregister)(JNIEnv* e, jclass c, jlong addr) {
  (void)e; (void)c;
  return b2s_host_unregister(PTR(void, addr));
}

/* ---- sizing ---- */
JNIEXPORT jlong JNICALL J(compressBound)(JNIEnv* e, jclass c, jint codec, jint blockSize, jlong n) {
  (void)e; (void)c;
  return (jlong)b2s_compress_bound((uint32_t)codec, (uint32_t)blockSize, (uint64_t)n);
}

/* ---- write side: n partition streams of one map task, packed arena in, packed .data arena out.
 *      off/len/dstOff/dstLen/checksums: long[n]; meta: long[1] = {dst_total}; status: int[n] ---- */
JNIEXPORT jint JNICALL J(compressPacked)(JNIEnv* e, jclass c, jint codec, jint level, jint blockSize, jint alg, jint n,
                                         jlong src, jlongArray off, jlongArray len, jlong dst, jlong dstCap,
                                         jlongArray dstOff, jlongArray dstLen, jlongArray meta, jlongArray checksums,
                                         jintArray status) {
  (void)c;
  jlong* o = crit_get(e, off);
  jlong* l = crit_get(e, len);
  jlong* dO = crit_get(e, dstOff);
  jlong* dL = crit_get(e, dstLen);
  jlong* m = crit_get(e, meta);
  jlong* k = crit_get(e, checksums);
  jint* st = crit_get(e, status);
  const int rc = b2s_compress_packed((uint32_t)codec, (int32_t)level, (uint32_t)blockSize, (uint32_t)alg, (uint32_t)n,
                                     PTR(const uint8_t, src), (const uint64_t*)o, (const uint64_t*)l, PTR(uint8_t, dst),
                                     (uint64_t)dstCap, (uint64_t*)dO, (uint64_t*)dL, (uint64_t*)m, (uint64_t*)k,
                                     (int32_t*)st);
  crit_put(e, status, st, 0);
  crit_put(e, checksums, k, 0);
  crit_put(e, meta, m, 0);
  crit_put(e, dstLen, dL, 0);
  crit_put(e, dstOff, dO, 0);
  crit_put(e, len, l, JNI_ABORT); /* inputs: nothing to copy back */
  crit_put(e, off, o, JNI_ABORT);
  return rc;
}

/* ---- read side: n prefetched blocks; block i owns slices [sliceBase[i], sliceBase[i+1]) of sliceLen/sliceChecksum
 *      (.index differences and .checksum values).  meta: long[1] = {dst_total}; status, badSlice: int[n] ---- */
JNIEXPORT jint JNICALL J(decompressPacked)(JNIEnv* e, jclass c, jint codec, jint alg, jint n, jlong src, jlongArray off,
                                           jlongArray len, jintArray sliceBase, jlongArray sliceLen,
                                           jlongArray sliceChecksum, jlong dst, jlong dstCap, jlongArray dstOff,
                                           jlongArray dstLen, jlongArray meta, jintArray status, jintArray badSlice) {
  (void)c;
  jlong* o = crit_get(e, off);
  jlong* l = crit_get(e, len);
  jint* sb = crit_get(e, sliceBase);
  jlong* sl = crit_get(e, sliceLen);
  jlong* sc = crit_get(e, sliceChecksum);
  jlong* dO = crit_get(e, dstOff);
  jlong* dL = crit_get(e, dstLen);
  jlong* m = crit_get(e, meta);
  jint* st = crit_get(e, status);
  jint* bad = crit_get(e, badSlice);
  const int rc = b2s_decompress_packed((uint32_t)codec, (uint32_t)alg, (uint32_t)n, PTR(const uint8_t, src),
                                       (const uint64_t*)o, (const uint64_t*)l, (const uint32_t*)sb, (const uint64_t*)sl,
                                       (const uint64_t*)sc, PTR(uint8_t, dst), (uint64_t)dstCap, (uint64_t*)dO,
                                       (uint64_t*)dL, (uint64_t*)m, (int32_t*)st, (int32_t*)bad);
  crit_put(e, badSlice, bad, 0);
  crit_put(e, status, st, 0);
  crit_put(e, meta, m, 0);
  crit_put(e, dstLen, dL, 0);
  crit_put(e, dstOff, dO, 0);
  crit_put(e, sliceChecksum, sc, JNI_ABORT);
  crit_put(e, sliceLen, sl, JNI_ABORT);
  crit_put(e, sliceBase, sb, JNI_ABORT);
  crit_put(e, len, l, JNI_ABORT);
  crit_put(e, off, o, JNI_ABORT);
  return rc;
}

/* decoded size of each of n compressed streams laid out in one arena (sizes the destination of decompressPacked) */
JNIEXPORT jint JNICALL J(decompressedSizePacked)(JNIEnv* e, jclass c, jint codec, jint n, jlong src, jlongArray off,
                                                 jlongArray len, jlongArray outLen, jintArray status) {
  (void)c;
  enum { kStack = 256 };
  const uint8_t* ptrs_stack[kStack];
  const uint8_t** ptrs = ptrs_stack;
  jlong* o = crit_get(e, off);
  jlong* l = crit_get(e, len);
  jlong* out = crit_get(e, outLen);
  jint* st = crit_get(e, status);
  int rc = 0;
  /* the C entry point takes per-stream pointers; chunks of kStack streams keep the critical section allocation-free */
  for (jint i0 = 0; i0 < n && rc == 0; i0 += kStack) {
    const jint m = n - i0 < kStack ? n - i0 : kStack;
    for (jint i = 0; i < m; i++) ptrs[i] = PTR(const uint8_t, src) + (uint64_t)o[i0 + i];
    rc = b2s_decompressed_size_batch((uint32_t)codec, (uint32_t)m, ptrs, (const uint64_t*)(l + i0),
                                     (uint64_t*)(out + i0), (int32_t*)(st + i0));
  }
  crit_put(e, status, st, 0);
  crit_put(e, outLen, out, 0);
  crit_put(e, len, l, JNI_ABORT);
  crit_put(e, off, o, JNI_ABORT);
  return rc;
}

/* ---- checksums only: one value per slice of a packed arena ---- */
JNIEXPORT jint JNICALL J(checksumPacked)(JNIEnv* e, jclass c, jint alg, jint n, jlong base, jlongArray off, jlongArray len,
                                         jlongArray out) {
  (void)c;
  jlong* o = crit_get(e, off);
  jlong* l = crit_get(e, len);
  jlong* r = crit_get(e, out);
  const int rc = b2s_checksum_packed((uint32_t)alg, (uint32_t)n, PTR(const uint8_t, base), (const uint64_t*)o,
                                     (const uint64_t*)l, (uint64_t*)r);
  crit_put(e, out, r, 0);
  crit_put(e, len, l, JNI_ABORT);
  crit_put(e, off, o, JNI_ABORT);
  return rc;
}

/* ---- observability: {total_ms, h2d_ms, d2h_ms, kernel_ms} x 1000 (microseconds) + byte counters of the calling
 *      thread's last call, for the MiB/s log lines of shuffle/S3MeasureOutputStream.scala:55-63 ---- */
JNIEXPORT jint JNICALL J(lastTiming)(JNIEnv* e, jclass c, jlongArray out8) {
  (void)c;
  b2s_timing t;
  const int rc = b2s_last_timing(&t);
  if ((*e)->GetArrayLength(e, out8) < 8) return B2S_E_ARG;
  jlong* o = crit_get(e, out8);
  if (o) {
    o[0] = (jlong)(t.total_ms * 1000.0);
    o[1] = (jlong)(t.h2d_ms * 1000.0);
    o[2] = (jlong)(t.d2h_ms * 1000.0);
    o[3] = (jlong)(t.kernel_ms * 1000.0);
    o[4] = (jlong)t.h2d_bytes;
    o[5] = (jlong)t.d2h_bytes;
    o[6] = (jlong)t.src_bytes;
    o[7] = (jlong)t.dst_bytes;
  }
  crit_put(e, out8, o, 0);
  return rc;
}
