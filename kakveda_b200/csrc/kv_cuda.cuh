// CUDA helpers shared by the kernel translation units.
#pragma once
#include "kv_internal.h"

#include <cuda_runtime.h>

#define KV_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return kv_fail(KV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                     __FILE__, __LINE__);                                                     \
  } while (0)

// Growable device array (amortised doubling) -- the raw CSR of an append-only index.
template <typename T>
struct DevVec {
  T *p = nullptr;
  int64_t n = 0, cap = 0;
  cudaError_t reserve(int64_t want, cudaStream_t s) {
    if (want <= cap) return cudaSuccess;
    int64_t nc = cap ? cap : 1024;
    while (nc < want) nc *= 2;
    T *q = nullptr;
    cudaError_t e = cudaMalloc(&q, (size_t)nc * sizeof(T));
    if (e != cudaSuccess) return e;
    if (n) {
      e = cudaMemcpyAsync(q, p, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) { cudaFree(q); return e; }
      e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { cudaFree(q); return e; }
    }
    cudaFree(p);
    p = q;
    cap = nc;
    return cudaSuccess;
  }
  void release() { cudaFree(p); p = nullptr; n = cap = 0; }
};

// Fixed-size device buffer re-allocated only when it must grow (scratch reused across calls).
template <typename T>
struct DevBuf {
  T *p = nullptr;
  int64_t cap = 0;
  cudaError_t ensure(int64_t want) {
    if (want <= cap) return cudaSuccess;
    cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, (size_t)(want > 0 ? want : 1) * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T>
struct PinnedBuf {
  T *p = nullptr;
  int64_t cap = 0;
  cudaError_t ensure(int64_t want) {
    if (want <= cap) return cudaSuccess;
    cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaHostAlloc(&p, (size_t)(want > 0 ? want : 1) * sizeof(T), cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { cudaFreeHost(p); p = nullptr; cap = 0; }
};
