// pool.cuh — host-side state pools and launch staging.
//
// Pool<S>: every instance of one voice type (or every engine's mix state) on one device, stored
// word-interleaved (SoA) with a fixed row pitch (capacity).  Slots are allocated on the host, their
// constructor-built initial states (voices.cuh init functions run on the host through the bit-exact
// gm:: math) are uploaded lazily in bulk.  LaunchSet<V>: the per-render list of (slot, output slot,
// events) for one type bucket and the launch of voice_kernel<V>.  Type bucketing keeps warps
// type-homogeneous (no cross-instrument divergence).
#pragma once
#include <algorithm>
#include <atomic>
#include "device_rt.h"
#include "kernels.cuh"
#include "../../include/gooey_batch.h"

namespace gh {

extern std::atomic<uint64_t> g_launches;

__global__ void scatter_words_kernel(uint32_t* __restrict__ dst, long long cap, const uint32_t* __restrict__ src, const uint32_t* __restrict__ slots,
                                     int n, int words) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t s = slots[k];
  for (int w = 0; w < words; w++) dst[(long long)w * cap + s] = src[(long long)w * n + k];
}

template <class S> struct Pool {
  static_assert(sizeof(S) % 4 == 0, "state must be whole 32-bit words");
  static constexpr int W = sizeof(S) / 4;
  DevBuf<uint32_t> d;
  int cap = 0, count = 0;
  std::vector<int> free_list;
  std::vector<std::pair<int, S>> pending;

  int alloc(const S& init) {
    int slot;
    if (!free_list.empty()) { slot = free_list.back(); free_list.pop_back(); } else slot = count++;
    pending.emplace_back(slot, init);
    return slot;
  }
  void release(int slot) {
    for (size_t i = 0; i < pending.size(); i++) if (pending[i].first == slot) { pending.erase(pending.begin() + i); break; }
    free_list.push_back(slot);
  }
  // grow (content preserving) and upload every pending initial state
  void flush(cudaStream_t st) {
    if (count > cap) {
      int ncap = std::max(pad32(count), cap * 2);
      DevBuf<uint32_t> nd;
      nd.alloc((size_t)W * ncap);
      GH_CUDA(cudaMemsetAsync(nd.p, 0, (size_t)W * ncap * 4, st));
      if (cap > 0) GH_CUDA(cudaMemcpy2DAsync(nd.p, (size_t)ncap * 4, d.p, (size_t)cap * 4, (size_t)cap * 4, W, cudaMemcpyDeviceToDevice, st));
      GH_CUDA(cudaStreamSynchronize(st));
      std::swap(d.p, nd.p); std::swap(d.n, nd.n);
      cap = ncap;
    }
    if (pending.empty()) return;
    const int n = (int)pending.size();
    std::vector<uint32_t> words((size_t)W * n), slots(n);
    for (int k = 0; k < n; k++) {
      slots[k] = (uint32_t)pending[k].first;
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&pending[k].second);
      for (int i = 0; i < W; i++) words[(size_t)i * n + k] = w[i];
    }
    DevBuf<uint32_t> dw, ds;
    dw.upload(words.data(), words.size(), st);
    ds.upload(slots.data(), slots.size(), st);
    scatter_words_kernel<<<(n + 127) / 128, 128, 0, st>>>(d.p, cap, dw.p, ds.p, n, W);
    g_launches.fetch_add(1);
    GH_CUDA(cudaGetLastError());
    GH_CUDA(cudaStreamSynchronize(st));
    pending.clear();
    pending.shrink_to_fit();
  }
};

inline gd::VoiceEvent make_event(uint32_t frame, uint32_t kind, uint32_t param, float value, uint32_t aux = 0) {
  gd::VoiceEvent e; e.frame = frame; e.kind = (uint16_t)kind; e.param = (uint16_t)param; e.value = value; e.aux = aux;
  return e;
}

// Flattened per-render event lists of n launch items.
struct EventStage {
  std::vector<uint32_t> begin;
  std::vector<gd::VoiceEvent> flat;
  DevBuf<gd::VoiceEvent> d_events;
  DevBuf<uint32_t> d_begin, d_cursor;
  void reset() { begin.clear(); flat.clear(); begin.push_back(0); }
  // append one item's events (already ordered by frame)
  void push_item(const std::vector<gd::VoiceEvent>& ev) {
    flat.insert(flat.end(), ev.begin(), ev.end());
    begin.push_back((uint32_t)flat.size());
  }
  void upload(cudaStream_t st) {
    if (begin.size() <= 1) return;
    if (flat.empty()) flat.push_back(make_event(0xffffffffu, 0xffff, 0, 0.0f));
    d_events.upload(flat.data(), flat.size(), st);
    d_begin.upload(begin.data(), begin.size(), st);
    d_cursor.upload(begin.data(), begin.size() - 1, st);
  }
};

// One type bucket of one render call.
template <class V> struct LaunchSet {
  using State = typename V::State;
  std::vector<uint32_t> slots, out_slots;
  EventStage ev;
  DevBuf<uint32_t> d_slots, d_out_slots;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  int n() const { return (int)slots.size(); }
  void reset() { slots.clear(); out_slots.clear(); ev.reset(); }
  void add(uint32_t slot, uint32_t out_slot, const std::vector<gd::VoiceEvent>& events) {
    slots.push_back(slot); out_slots.push_back(out_slot); ev.push_item(events);
  }
  void upload(cudaStream_t st) {
    if (slots.empty()) return;
    if (!stream) {
      GH_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      GH_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    }
    d_slots.upload(slots.data(), slots.size(), st);
    d_out_slots.upload(out_slots.data(), out_slots.size(), st);
    ev.upload(st);
  }
  ~LaunchSet() { if (done) cudaEventDestroy(done); if (stream) cudaStreamDestroy(stream); }
  // Forks from `parent` (after `start`), launches on this bucket's stream, joins back into `parent`.
  void launch(Pool<State>& pool, cudaStream_t parent, cudaEvent_t start, const gd::RateCtx& rc, uint32_t frame0, int frames, float* out,
              long long stride, int layout) {
    const int cnt = n();
    if (cnt == 0 || frames <= 0) return;
    cudaStream_t st = stream;
    GH_CUDA(cudaStreamWaitEvent(st, start, 0));
    gd::VoiceLaunch L;
    L.state = pool.d.p; L.n = cnt; L.n_pad = pool.cap;
    L.slots = d_slots.p;
    L.out_slots = layout == gd::OUT_TIME_MAJOR ? d_out_slots.p : nullptr;
    L.events = ev.d_events.p; L.ev_begin = ev.d_begin.p; L.ev_cursor = ev.d_cursor.p;
    L.frame0 = frame0; L.frames = frames; L.out = out; L.stride = stride; L.layout = layout; L.slot0 = 0;
    L.rows = layout == gd::OUT_VOICE_MAJOR ? d_out_slots.p : nullptr;
    L.rc = rc;
    // Small buckets: one warp per block so the warps spread over all 148 SMs; large: 128-thread blocks.
    if (cnt <= 148 * 32 * 4) gd::voice_kernel<V, 32><<<(cnt + 31) / 32, 32, 0, st>>>(L);
    else gd::voice_kernel<V, 128><<<(cnt + 127) / 128, 128, 0, st>>>(L);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GH_CUDA(cudaGetLastError());
    GH_CUDA(cudaEventRecord(done, st));
    GH_CUDA(cudaStreamWaitEvent(parent, done, 0));
  }
};

}  // namespace gh
