// Host build of spark-s3-shuffle_b200/csrc/zstd_core.h for tests/test_zstd_core.py: the very functions the CUDA
// kernels of zstd.cu call, compiled by g++ so they can be checked against libzstd.so.1 without a GPU.
// Test infrastructure only — the C ABI never runs this.
#include <stdlib.h>

#include "../../spark-s3-shuffle_b200/csrc/zstd_core.h"

extern "C" {
long long zc_decode(const unsigned char* src, unsigned long long n, unsigned char* dst, unsigned long long cap) {
  b2s::zstd::Workspace* w = (b2s::zstd::Workspace*)malloc(sizeof(b2s::zstd::Workspace));
  long long r = b2s::zstd::decode_stream(w, src, n, dst, cap, false);
  free(w);
  return r;
}
long long zc_size(const unsigned char* src, unsigned long long n) {
  b2s::zstd::Workspace* w = (b2s::zstd::Workspace*)malloc(sizeof(b2s::zstd::Workspace));
  long long r = b2s::zstd::decode_stream(w, src, n, nullptr, 0, true);
  free(w);
  return r;
}
unsigned long long zc_workspace_bytes() { return sizeof(b2s::zstd::Workspace); }
}
