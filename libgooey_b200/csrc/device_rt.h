// device_rt.h — small CUDA runtime helpers shared by the host-side classes:
// error latch, device buffers, pinned buffers, SoA state upload.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>

namespace gh {

std::string& last_error();
inline void set_error(const std::string& s) { last_error() = s; }

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define GH_CUDA(call)                                                                            \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      char _b[512];                                                                              \
      snprintf(_b, sizeof _b, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); \
      throw gh::CudaError(_b);                                                                   \
    }                                                                                            \
  } while (0)

template <class T> struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  // Growth frees and re-allocates, and cudaFree waits for the whole device: a per-piece table that outgrows its buffer by a few
  // entries in the middle of a render drains the pipeline (measured: ~0.5 s at the bar boundary of an 8192-engine bounce).
  // Small buffers therefore grow with 50 % headroom; large ones (voice / output blocks) are sized exactly.
  void alloc(size_t count) {
    if (count <= n && p) return;
    static const bool trace = getenv("GOOEY_B200_TRACE") != nullptr;
    if (trace && p) fprintf(stderr, "[gooey trace] device buffer regrown: %zu -> %zu bytes (cudaFree drains the device)\n", n * sizeof(T), count * sizeof(T));
    if (p) { cudaFree(p); p = nullptr; }
    if (count * sizeof(T) <= ((size_t)64 << 20)) count += count / 2 + 64;
    n = count;
    if (count) GH_CUDA(cudaMalloc(&p, count * sizeof(T)));
  }
  void upload(const T* src, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) GH_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(cudaStream_t s) { if (p && n) GH_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

// AoS host states -> word-interleaved SoA on the device: dev[w * n_pad + v].
template <class S> void upload_states(DevBuf<uint32_t>& dev, const std::vector<S>& host, int n_pad, cudaStream_t s) {
  static_assert(sizeof(S) % 4 == 0, "state must be a whole number of 32-bit words");
  constexpr size_t W = sizeof(S) / 4;
  std::vector<uint32_t> t(W * (size_t)n_pad, 0u);
  for (size_t v = 0; v < host.size(); v++) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&host[v]);
    for (size_t i = 0; i < W; i++) t[i * n_pad + v] = w[i];
  }
  dev.alloc(t.size());
  GH_CUDA(cudaMemcpyAsync(dev.p, t.data(), t.size() * 4, cudaMemcpyHostToDevice, s));
  GH_CUDA(cudaStreamSynchronize(s));  // t is a temporary
}

inline int pad32(int n) { return (n + 31) & ~31; }

}  // namespace gh
