// voice_batch.cuh — host-side owner of type-bucketed voice groups.
//
// A VoiceGroup<V> is every voice of one instrument type in a batch: host-built
// initial states (constructors restated in voices.cuh, run on the host through
// the bit-exact gm:: math), the device SoA state image, the per-voice event
// lists, and the launch of voice_kernel<V>.  Type bucketing keeps warps
// type-homogeneous (no cross-instrument divergence).
#pragma once
#include <algorithm>
#include <atomic>
#include "device_rt.h"
#include "kernels.cuh"
#include "../../include/gooey_batch.h"

namespace gh {

extern std::atomic<uint64_t> g_launches;

struct EventList {
  // per-voice pending events for the next render, frames relative to its start
  std::vector<std::vector<gd::VoiceEvent>> per_voice;
  void resize(size_t n) { per_voice.resize(n); }
  void add(uint32_t v, uint32_t frame, uint32_t kind, uint32_t param, float value) {
    gd::VoiceEvent e; e.frame = frame; e.kind = (uint16_t)kind; e.param = (uint16_t)param; e.value = value;
    per_voice[v].push_back(e);
  }
};

template <class V> struct VoiceGroup {
  using State = typename V::State;
  std::vector<State> init_states;   // until first upload
  std::vector<int> rows;            // output row (voice-major) or slot (time-major) of each voice
  EventList events;
  DevBuf<uint32_t> d_state;
  DevBuf<gd::VoiceEvent> d_events;
  DevBuf<uint32_t> d_ev_begin, d_ev_cursor, d_rows;
  cudaStream_t stream = nullptr;   // each group renders on its own stream so type buckets overlap
  cudaEvent_t done = nullptr;
  int n = 0, n_pad = 0;
  bool uploaded = false;

  int add(const State& s, int row) {
    init_states.push_back(s);
    rows.push_back(row);
    n = (int)init_states.size();
    events.resize(n);
    return n - 1;
  }
  void ensure_uploaded(cudaStream_t st) {
    if (uploaded || n == 0) return;
    n_pad = pad32(n);
    upload_states(d_state, init_states, n_pad, st);
    std::vector<uint32_t> r(rows.begin(), rows.end());
    d_rows.upload(r.data(), r.size(), st);
    GH_CUDA(cudaStreamSynchronize(st));
    GH_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    GH_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    init_states.clear();
    init_states.shrink_to_fit();
    uploaded = true;
  }
  // flatten + upload this render's events (stable by frame), reset cursors
  void stage_events(cudaStream_t st) {
    if (n == 0) return;
    std::vector<uint32_t> begin(n + 1, 0);
    size_t total = 0;
    for (int v = 0; v < n; v++) { begin[v] = (uint32_t)total; total += events.per_voice[v].size(); }
    begin[n] = (uint32_t)total;
    std::vector<gd::VoiceEvent> flat;
    flat.reserve(total + 1);
    for (int v = 0; v < n; v++) {
      auto& ev = events.per_voice[v];
      std::stable_sort(ev.begin(), ev.end(), [](const gd::VoiceEvent& a, const gd::VoiceEvent& b) { return a.frame < b.frame; });
      flat.insert(flat.end(), ev.begin(), ev.end());
      ev.clear();
    }
    if (flat.empty()) flat.push_back(gd::VoiceEvent{0xffffffffu, 0, 0, 0.0f});
    d_events.upload(flat.data(), flat.size(), st);
    d_ev_begin.upload(begin.data(), begin.size(), st);
    d_ev_cursor.upload(begin.data(), n, st);  // cursor starts at each voice's begin
    GH_CUDA(cudaStreamSynchronize(st));        // host vectors are temporaries
  }
  // rows must be contiguous from rows[0] (voice-major) — groups are laid out that way by their owners
  ~VoiceGroup() { if (done) cudaEventDestroy(done); if (stream) cudaStreamDestroy(stream); }
  // Forks from `parent` (waits on `start`), launches on the group's stream, and makes `parent` wait for completion.
  void launch(cudaStream_t parent, cudaEvent_t start, const gd::RateCtx& rc, uint32_t frame0, int frames, float* out, long long stride, int layout, int slot0, bool use_rows) {
    if (n == 0 || frames <= 0) return;
    cudaStream_t st = stream;
    GH_CUDA(cudaStreamWaitEvent(st, start, 0));
    gd::VoiceLaunch L;
    L.state = d_state.p; L.n = n; L.n_pad = n_pad; L.slots = nullptr; L.out_slots = nullptr;
    L.events = d_events.p; L.ev_begin = d_ev_begin.p; L.ev_cursor = d_ev_cursor.p;
    L.frame0 = frame0; L.frames = frames; L.out = out; L.stride = stride; L.layout = layout; L.slot0 = slot0; L.rows = use_rows ? d_rows.p : nullptr; L.rc = rc;
    // Small batches: one warp per block so the warps spread over all 148 SMs; large: 128-thread blocks.
    if (n <= 148 * 32 * 4) {
      gd::voice_kernel<V, 32><<<(n + 31) / 32, 32, 0, st>>>(L);
    } else {
      gd::voice_kernel<V, 128><<<(n + 127) / 128, 128, 0, st>>>(L);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GH_CUDA(cudaGetLastError());
    GH_CUDA(cudaEventRecord(done, st));
    GH_CUDA(cudaStreamWaitEvent(parent, done, 0));
  }
};

}  // namespace gh
