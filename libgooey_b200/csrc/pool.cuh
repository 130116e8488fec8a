// pool.cuh — host-side state pools, the engine clock table and the per-type launch sequence.
//
// Pool<S>: every instance of one voice type (or every engine's mix state) on one device, stored
// word-interleaved (SoA) with a fixed row pitch (capacity).  Slots are allocated on the host, their
// constructor-built initial states (voices.cuh init functions run on the host through the bit-exact
// gm:: math) are uploaded lazily in bulk.
// TypeRunner<V>: the per-render list of (slot, output row, events) for one type bucket and the
// A -> (B | C pipelined over time chunks) + S launch sequence of kernels.cuh.  Type bucketing keeps
// warps type-homogeneous (no cross-instrument divergence).
#pragma once
#include <algorithm>
#include <atomic>
#include <map>
#include <type_traits>
#include <set>
#include <mutex>
#include "device_rt.h"
#include "wave.cuh"
#include "coop.cuh"
#include "gran_wave.cuh"
#include "bass_wave.cuh"
#ifndef GOOEY_WAVE_CTA_WARPS_DEFAULT
#define GOOEY_WAVE_CTA_WARPS_DEFAULT 1
#endif
#ifndef GOOEY_WAVE_G_KICK
#define GOOEY_WAVE_G_KICK 32
#define GOOEY_WAVE_G_SNARE 32
#define GOOEY_WAVE_G_HAT 32
#define GOOEY_WAVE_G_TOM 32
#endif
#include "../../include/gooey_batch.h"

namespace gh {

extern std::atomic<uint64_t> g_launches;

// Per-kernel device time, measured with CUDA events on the stream the kernel is launched on (bench.py's roofline line).
struct KernelStat { uint64_t launches = 0; double ms = 0.0; double voice_frames = 0.0; };
std::map<std::string, KernelStat>& kernel_stats();
std::mutex& kernel_stats_mutex();

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
      // Row pitch of the word-interleaved state (and of every ring arena that follows the mix pool's capacity).  Kept an ODD
      // multiple of 32 slots: with a power-of-two pitch the words of one slot (pitch * 4 bytes apart) fall into the same L1 / L2
      // sets and evict each other (measured: 1600 granulators ran 1.8x slower in a pool that had grown to 2048 slots).
      int ncap = std::max(pad32(count), cap * 2);
      if (((ncap / 32) & 1) == 0) ncap += 32;
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

// The engine clock: tt[k] = value of a f64 accumulator after k additions of 1/sr, exactly as the reference's
// `current_time += 1.0 / sample_rate` (bounce.rs:48-53, ffi.rs:1379).  The host keeps one checkpoint every CK additions
// (8 bytes per 1.5 s of audio) and rebuilds any window [k0, k0 + n) from the nearest checkpoint by the same repeated
// addition, so memory does not grow with how long an engine has been streaming.
struct ClockTable {
  static constexpr uint64_t CK = 65536;
  double dt = 0.0;
  std::vector<double> ckpt;   // ckpt[i] = value after i * CK additions
  std::mutex mu;
  void fill(uint64_t k0, size_t n, double* out) {
    std::lock_guard<std::mutex> lk(mu);
    if (ckpt.empty()) ckpt.push_back(0.0);
    const uint64_t need = k0 / CK;
    while (ckpt.size() <= need) { double t = ckpt.back(); for (uint64_t i = 0; i < CK; i++) t += dt; ckpt.push_back(t); }
    double t = ckpt[need];
    for (uint64_t k = need * CK; k < k0; k++) t += dt;
    for (size_t i = 0; i < n; i++) { out[i] = t; t += dt; }
  }
};
// A caller-owned device copy of tt[k0 .. k0 + n): view() returns a pointer `p` such that p[k] is valid for every k of the
// requested range (kernels index the clock with absolute k).  Uploads are ordered on the caller's stream; a window that
// already covers the range is reused (a bounce always asks for [0, frames]).
struct ClockWindow {
  DevBuf<double> d;
  std::vector<double> h;
  uint64_t k0 = 0; size_t n = 0;
  const double* view(ClockTable& T, uint64_t kmin, uint64_t kmax, cudaStream_t st) {
    if (kmin > 0) kmin -= 1;                                   // the poly synth's catch-up reads tt[k - 1]
    if (!(n && kmin >= k0 && kmax < k0 + n)) {
      const size_t want = (size_t)(kmax - kmin + 1);
      const size_t cap = std::max<size_t>((want + 8191) & ~(size_t)8191, 1u << 16);
      if (d.n < cap) { GH_CUDA(cudaStreamSynchronize(st)); d.alloc(cap); }
      h.resize(want);
      T.fill(kmin, want, h.data());
      GH_CUDA(cudaMemcpyAsync(d.p, h.data(), want * sizeof(double), cudaMemcpyHostToDevice, st));
      GH_CUDA(cudaStreamSynchronize(st));                      // h is reused by the next call
      k0 = kmin; n = want;
    }
    return d.p - k0;
  }
};
ClockTable& clock_table(float sr);

// Which voice types have a scan back-end (wave.cuh); the others use back_kernel (one voice per lane).  ID indexes the
// per-type group width table below.
template <class V> struct WaveOf { static constexpr bool has = false; static constexpr int ID = -1; static constexpr const char* name = "back_kernel"; };
template <> struct WaveOf<gd::KickV> { static constexpr bool has = true; static constexpr int ID = 0; using w32 = gd::w32::KickW; static constexpr const char* name = "wave_kernel<KickW>";
#ifdef GOOEY_WAVE_ALL_WIDTHS
  using w16 = gd::w16::KickW; using w8 = gd::w8::KickW;
#endif
};
template <> struct WaveOf<gd::SnareV> { static constexpr bool has = true; static constexpr int ID = 1; using w32 = gd::w32::SnareW; static constexpr const char* name = "wave_kernel<SnareW>";
#ifdef GOOEY_WAVE_ALL_WIDTHS
  using w16 = gd::w16::SnareW; using w8 = gd::w8::SnareW;
#endif
};
template <> struct WaveOf<gd::HatV> { static constexpr bool has = true; static constexpr int ID = 2; using w32 = gd::w32::HatW; static constexpr const char* name = "wave_kernel<HatW>";
#ifdef GOOEY_WAVE_ALL_WIDTHS
  using w16 = gd::w16::HatW; using w8 = gd::w8::HatW;
#endif
};
template <> struct WaveOf<gd::TomV> { static constexpr bool has = true; static constexpr int ID = 3; using w32 = gd::w32::TomW; static constexpr const char* name = "wave_kernel<TomW>";
#ifdef GOOEY_WAVE_ALL_WIDTHS
  using w16 = gd::w16::TomW; using w8 = gd::w8::TomW;
#endif
};
// Which voice types have a CTA-cooperative back end (coop.cuh).
template <class V> struct CoopOf { static constexpr bool has = false; };
template <> struct CoopOf<gd::TomV> { static constexpr bool has = true; using C = gd::coop::TomC; static constexpr const char* name = "coop_kernel<TomC>"; };
// GOOEY_B200_BACKEND: "serial" = per-sample-order kernel C everywhere (the A/B reference of the parity tests), "wave" = the
// warp-per-voice scan back end of round 1, anything else / unset = the CTA-cooperative back end where a type has one.
inline bool wave_backend() { const char* e = getenv("GOOEY_B200_BACKEND"); return e && strcmp(e, "wave") == 0; }
// Voices per CTA of the cooperative back end (4, 8 or 16; GOOEY_B200_COOP_NV overrides the default).
inline int coop_nv() { const char* e = getenv("GOOEY_B200_COOP_NV"); if (e) { int v = atoi(e); if (v == 4 || v == 8 || v == 16) return v; } return 16; }
// The cooperative back end issues ~3x fewer instructions per voice-frame but walks a block through barrier-separated phases,
// so a voice's timeline is slower than in the warp-per-voice back end while voices are too few to fill the device with CTAs:
// it takes over above this many voices of one type (measured crossover; GOOEY_B200_COOP_ABOVE overrides).
inline int coop_above() { if (const char* e = getenv("GOOEY_B200_COOP_ABOVE")) { int v = atoi(e); if (v >= 0) return v; } return 4096; }
template <class C, int NV> inline void coop_launch(int cnt, cudaStream_t s, const gd::VoiceLaunch& L) {
  using SM = gd::coop::Smem<C, NV>;
  auto kernel = gd::coop::coop_kernel<C, NV>;
  static std::mutex mu;
  static std::set<int> done;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (done.insert(dev).second) {
      GH_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM)));
      GH_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    }
  }
  kernel<<<(cnt + NV - 1) / NV, NV * 32, sizeof(SM), s>>>(L);
}

// Lanes per voice for kick / snare / hat / tom (see wave.cuh "Group width").  GOOEY_B200_WAVE_G="32,32,8,8" overrides
// (tuning only; every width renders the same audio up to the scans' re-association noise).
inline int wave_group_width(int id) {
  static int table[4] = {0, 0, 0, 0};
  static std::once_flag once;
  std::call_once(once, [] {
    const int def[4] = {GOOEY_WAVE_G_KICK, GOOEY_WAVE_G_SNARE, GOOEY_WAVE_G_HAT, GOOEY_WAVE_G_TOM};
    for (int i = 0; i < 4; i++) table[i] = def[i];
    if (const char* e = getenv("GOOEY_B200_WAVE_G")) {
      int v[4];
      if (sscanf(e, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4)
        for (int i = 0; i < 4; i++) if (v[i] == 8 || v[i] == 16 || v[i] == 32) table[i] = v[i];
    }
#ifndef GOOEY_WAVE_ALL_WIDTHS
    for (int i = 0; i < 4; i++) table[i] = 32;
#endif
  });
  return table[id];
}
// GOOEY_B200_BACKEND=serial forces the per-sample-order back-end (kernel C) everywhere: the A/B switch used by the
// parity tests to compare the two back-ends.
inline bool serial_backend() { const char* e = getenv("GOOEY_B200_BACKEND"); return e && strcmp(e, "serial") == 0; }
// Above this many voices of one type the per-sample-order back end with one voice per THREAD wins: the warp-per-voice
// back end spends 32 lanes on the replayed recurrences of one voice, which pays only while voices are too few to fill the
// device with threads (measured, 32768 voices x 2 s: tom 234 vs 526 ms, hi-hat 117 vs 193 ms; at 16384: 176 vs 263 and
// 104 vs 97 ms).  GOOEY_B200_SERIAL_ABOVE overrides.
// Warps (= voices) per CTA of the scan back end: 8 keeps an SM's warps in step (one CTA barrier per block) so they share
// instruction fetches, 1 lets every voice run free.  Measured on C2: no gain from the lock-step (isolated buckets equal,
// the mix 54.7 vs 50.1 ms), so 1 is the default; build with -DGOOEY_WAVE_LOCKSTEP and set GOOEY_B200_WAVE_WARPS=8 to try it.
inline int wave_cta_warps() {
  const char* e = getenv("GOOEY_B200_WAVE_WARPS");
  if (e && atoi(e) == 1) return 1;
  if (e && atoi(e) == 8) return 8;
  return GOOEY_WAVE_CTA_WARPS_DEFAULT;
}
inline int serial_above() {
  if (const char* e = getenv("GOOEY_B200_SERIAL_ABOVE")) { int v = atoi(e); if (v > 0) return v; }
  return 12288;
}

// Kernels whose shared-memory carve-out differs cannot share an SM: the L1 / shared split is an SM-wide setting, so a
// back end with 19 KB of scan tables per CTA and one with 3.5 KB would be given disjoint SMs instead of interleaving
// their warps.  Every kernel of the render is therefore pinned to the same (maximum shared) carve-out, once per device.
template <class K> inline void same_carveout(K kernel) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  if (done.insert({dev, (const void*)kernel}).second)
    GH_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
}
#define GH_LAUNCH(kernel, grid, block, stream, ...) do { same_carveout(kernel); kernel<<<grid, block, 0, stream>>>(__VA_ARGS__); } while (0)

// Rank of a voice type by the length of its per-chunk dependency chain (tools/type_scaling.py: tom 31 ms, snare 22 ms,
// kick 11 ms, hi-hat 10 ms per 1024 voices x 2 s); used for stream priorities.
template <class V> struct PrioOf { static constexpr int rank = 4; };
template <> struct PrioOf<gd::TomV> { static constexpr int rank = 0; };
template <> struct PrioOf<gd::SnareV> { static constexpr int rank = 1; };
template <> struct PrioOf<gd::KickV> { static constexpr int rank = 2; };
template <> struct PrioOf<gd::HatV> { static constexpr int rank = 3; };

// Frames per front/back pipeline stage.  GOOEY_B200_CHUNK overrides (tuning).
inline int default_chunk_frames() {
  if (const char* e = getenv("GOOEY_B200_CHUNK")) { int v = atoi(e); if (v >= 1024 && v <= (1 << 20)) return (v + 31) & ~31; }
  return 8192;
}
// One type bucket of one render call.
template <class V> struct TypeRunner {
  using State = typename V::State;
  Pool<State> pool;
  std::vector<uint32_t> slots, rows, ev_begin, span_off;
  std::vector<gd::VoiceEvent> ev_flat;
  std::vector<gd::ModRoute> routes_flat; std::vector<uint32_t> route_begin;       // LFO routes per voice of this launch
  DevBuf<gd::ModRoute> d_routes; DevBuf<uint32_t> d_route_begin;
  const float* mod_planes = nullptr; long long mod_pitch = 0; int mod_frame0 = 0;  // set by the engine path before launch()
  DevBuf<uint32_t> d_slots, d_rows, d_ev_begin, d_span_off, d_n_spans, d_span_cursor;
  DevBuf<gd::VoiceEvent> d_events;
  DevBuf<uint8_t> d_mode, d_spans;
  DevBuf<float> d_planes[2];
  cudaStream_t sB = nullptr, sC = nullptr, sS = nullptr;
  cudaEvent_t evA = nullptr, evB[2] = {nullptr, nullptr}, evC[2] = {nullptr, nullptr}, evDoneC = nullptr, evDoneS = nullptr;
  std::vector<cudaEvent_t> evT0, evT1;   // timing brackets of the back-end launch of chunk i (on sC)
  std::vector<double> timed_units;       // voice-frames of that launch
  std::vector<cudaEvent_t> evChunk;      // evChunk[i]: frames of chunk i are final in the output (fast-path voices)
  int chunk_frames = default_chunk_frames();
  int launched_chunks = 0;               // chunks of the most recent launch (0 = one undivided launch)
  const char* backend_name = "back_kernel";   // the back-end kernel of the most recent launch (kernel statistics key)
  static int chunk_of(int frames, int chunk_frames) { return std::min(chunk_frames, (frames + 31) & ~31); }
  // Makes `s` wait until every voice of this bucket has written frames [i * chunk, (i + 1) * chunk) of the last launch.
  void wait_chunk(cudaStream_t s, int i) {
    if (n() == 0) return;
    if (launched_chunks == 0) { GH_CUDA(cudaStreamWaitEvent(s, evDoneC, 0)); return; }
    GH_CUDA(cudaStreamWaitEvent(s, evDoneS, 0));
    GH_CUDA(cudaStreamWaitEvent(s, evChunk[std::min(i, std::abs(launched_chunks) - 1)], 0));
  }

  int n() const { return (int)slots.size(); }
  void reset() { slots.clear(); rows.clear(); ev_flat.clear(); ev_begin.clear(); ev_begin.push_back(0); span_off.clear(); span_off.push_back(0);
                 routes_flat.clear(); route_begin.clear(); route_begin.push_back(0); }
  // events must be ordered by frame and lie inside the call
  void add(uint32_t slot, uint32_t out_row, const std::vector<gd::VoiceEvent>& events, const std::vector<gd::ModRoute>* routes = nullptr) {
    slots.push_back(slot); rows.push_back(out_row);
    if (routes) routes_flat.insert(routes_flat.end(), routes->begin(), routes->end());
    route_begin.push_back((uint32_t)routes_flat.size());
    uint32_t distinct = 0, last = 0xffffffffu;
    for (const auto& e : events) { if (e.frame != last) { distinct++; last = e.frame; } }
    ev_flat.insert(ev_flat.end(), events.begin(), events.end());
    ev_begin.push_back((uint32_t)ev_flat.size());
    span_off.push_back(span_off.back() + distinct + 1);
  }
  void ensure_streams() {
    if (sC) return;
    // Stream priorities.  The buckets overlap on the device and the step ends when the slowest bucket's chunk chain
    // does (tom on C2), so the bucket with the longest chain should win the block scheduler's slots whenever one frees up
    // and the cheaper buckets fill the gaps: PrioOf<V> ranks the types by per-voice cost, 0 = longest chain.
    // Measured on C2: 54.4 ms ranked vs 55.1 ms equal (noise) — the tom chain is slowed by sharing issue slots, not by
    // waiting for slots — so equal priorities stay the default; GOOEY_B200_PRIO=1 enables the ranking (tuning knob).
    int prio_lo = 0, prio_hi = 0;
    GH_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));   // numerically lower = higher priority
    const char* pe = getenv("GOOEY_B200_PRIO");
    int prio = prio_lo;
    if (pe && pe[0] == '1') prio = std::min(prio_lo, prio_hi + PrioOf<V>::rank);
    GH_CUDA(cudaStreamCreateWithPriority(&sB, cudaStreamNonBlocking, prio));
    GH_CUDA(cudaStreamCreateWithPriority(&sC, cudaStreamNonBlocking, prio));
    GH_CUDA(cudaStreamCreateWithFlags(&sS, cudaStreamNonBlocking));
    cudaEvent_t* evs[] = {&evA, &evB[0], &evB[1], &evC[0], &evC[1], &evDoneC, &evDoneS};
    for (auto e : evs) GH_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  ~TypeRunner() {
    cudaEvent_t evs[] = {evA, evB[0], evB[1], evC[0], evC[1], evDoneC, evDoneS};
    for (auto e : evs) if (e) cudaEventDestroy(e);
    for (auto e : evChunk) cudaEventDestroy(e);
    for (auto e : evT0) cudaEventDestroy(e);
    for (auto e : evT1) cudaEventDestroy(e);
    if (sB) cudaStreamDestroy(sB);
    if (sC) cudaStreamDestroy(sC);
    if (sS) cudaStreamDestroy(sS);
  }
  // After the launch's streams have been synchronised: add the back-end launches of the last call to the global table.
  void collect_stats() {
    const char* name = backend_name;
    if (launched_chunks <= 0) return;
    std::lock_guard<std::mutex> lk(kernel_stats_mutex());
    KernelStat& k = kernel_stats()[name];
    for (int i = 0; i < launched_chunks; i++) {
      float ms = 0.0f;
      if (cudaEventElapsedTime(&ms, evT0[i], evT1[i]) != cudaSuccess) { cudaGetLastError(); continue; }
      k.launches++; k.ms += ms; k.voice_frames += timed_units[i];
    }
    launched_chunks = -launched_chunks;   // collected once
  }
  // Forks from `parent` (after `start`), runs the bucket on its own streams, joins back into `parent`.
  void launch(cudaStream_t parent, cudaEvent_t start, const gd::RateCtx& rc, const double* tt, int frames, float* out, long long stride) {
    const int cnt = n();
    launched_chunks = 0;
    if (cnt == 0 || frames <= 0) return;
    ensure_streams();
    GH_CUDA(cudaStreamWaitEvent(sC, start, 0));
    pool.flush(sC);
    d_slots.upload(slots.data(), slots.size(), sC);
    d_rows.upload(rows.data(), rows.size(), sC);
    if (ev_flat.empty()) ev_flat.push_back(make_event(0xffffffffu, 0xffff, 0, 0.0f));
    d_events.upload(ev_flat.data(), ev_flat.size(), sC);
    d_ev_begin.upload(ev_begin.data(), ev_begin.size(), sC);
    gd::VoiceLaunch L;
    memset(&L, 0, sizeof L);
    L.state = pool.d.p; L.n = cnt; L.n_pad = pool.cap;
    L.slots = d_slots.p; L.rows = d_rows.p; L.row0 = 0;
    L.events = d_events.p; L.ev_begin = d_ev_begin.p; L.tt = tt; L.frames = frames; L.out = out; L.stride = stride; L.rc = rc;
    const bool modulated = !routes_flat.empty() && mod_planes;
    if (modulated) {
      d_routes.upload(routes_flat.data(), routes_flat.size(), sC);
      d_route_begin.upload(route_begin.data(), route_begin.size(), sC);
      L.routes = d_routes.p; L.route_begin = d_route_begin.p; L.mod_planes = mod_planes; L.mod_pitch = mod_pitch; L.mod_frame0 = mod_frame0;
    }
    if constexpr (V::FAST) {
      using Span = typename V::Span;
      d_span_off.upload(span_off.data(), span_off.size(), sC);
      d_n_spans.alloc(cnt); d_span_cursor.alloc(cnt); d_mode.alloc(cnt);
      d_spans.alloc((size_t)span_off.back() * sizeof(Span));
      const int chunk = chunk_of(frames, chunk_frames);
      const int rows_pad = pad32(cnt);
      const size_t plane_floats = (size_t)rows_pad * chunk;
      d_planes[0].alloc(plane_floats * V::NPL); d_planes[1].alloc(plane_floats * V::NPL);
      L.mode = d_mode.p; L.spans = d_spans.p; L.span_off = d_span_off.p; L.n_spans = d_n_spans.p; L.span_cursor = d_span_cursor.p;
      L.plane_stride = (long long)plane_floats; L.pitch = chunk;
      GH_LAUNCH((gd::plan_kernel<V>), (cnt + 63) / 64, 64, sC, L);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      GH_CUDA(cudaGetLastError());
      GH_CUDA(cudaEventRecord(evA, sC));
      GH_CUDA(cudaStreamWaitEvent(sB, evA, 0));
      GH_CUDA(cudaStreamWaitEvent(sS, evA, 0));
      // general path for the voices A could not plan (whole call, one launch)
      GH_LAUNCH((gd::slow_kernel<V, 32>), (cnt + 31) / 32, 32, sS, L);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      GH_CUDA(cudaGetLastError());
      GH_CUDA(cudaEventRecord(evDoneS, sS));
      int i = 0;
      for (int c0 = 0; c0 < frames; c0 += chunk, i++) {
        const int b = i & 1;
        L.chunk0 = c0; L.chunk_frames = std::min(chunk, frames - c0);
        L.planes = d_planes[b].p;
        if (i >= 2) GH_CUDA(cudaStreamWaitEvent(sB, evC[b], 0));   // plane buffer b is free once C of chunk i-2 is done
        const int bpv = (L.chunk_frames + 255) / 256;
        GH_LAUNCH((gd::front_kernel<V, 256>), (unsigned)((size_t)cnt * bpv), 256, sB, L, bpv);
        GH_CUDA(cudaGetLastError());
        GH_CUDA(cudaEventRecord(evB[b], sB));
        GH_CUDA(cudaStreamWaitEvent(sC, evB[b], 0));
        if ((int)evT0.size() <= i) { cudaEvent_t e0, e1; GH_CUDA(cudaEventCreate(&e0)); GH_CUDA(cudaEventCreate(&e1)); evT0.push_back(e0); evT1.push_back(e1); timed_units.push_back(0.0); }
        timed_units[i] = (double)cnt * L.chunk_frames;
        GH_CUDA(cudaEventRecord(evT0[i], sC));
        bool launched = false;
        backend_name = WaveOf<V>::has ? WaveOf<V>::name : "back_kernel";
        if constexpr (CoopOf<V>::has) {
          if (!serial_backend() && !wave_backend() && cnt >= coop_above()) {
            const int nv = coop_nv();
            if (nv == 4) coop_launch<typename CoopOf<V>::C, 4>(cnt, sC, L);
            else if (nv == 8) coop_launch<typename CoopOf<V>::C, 8>(cnt, sC, L);
            else coop_launch<typename CoopOf<V>::C, 16>(cnt, sC, L);
            backend_name = CoopOf<V>::name;
            launched = true;
          }
        }
        if (launched) {}
        else if constexpr (WaveOf<V>::has) {
          if (!serial_backend() && cnt < serial_above()) {
            const int g = wave_group_width(WaveOf<V>::ID);          // voices per warp = 32 / g, one warp per CTA
            const int warps = (cnt * g + 31) / 32;
#ifdef GOOEY_WAVE_ALL_WIDTHS
            if (g == 16) GH_LAUNCH((gd::w16::wave_kernel<typename WaveOf<V>::w16, 1>), warps, 32, sC, L);
            else if (g == 8) GH_LAUNCH((gd::w8::wave_kernel<typename WaveOf<V>::w8, 1>), warps, 32, sC, L);
            else
#endif
#ifdef GOOEY_WAVE_LOCKSTEP          // compiled on request only (it doubles the build time of the back ends)
            if (wave_cta_warps() == 8) GH_LAUNCH((gd::w32::wave_kernel<typename WaveOf<V>::w32, 8>), (warps + 7) / 8, 256, sC, L);
            else
#endif
            GH_LAUNCH((gd::w32::wave_kernel<typename WaveOf<V>::w32, 1>), warps, 32, sC, L);
          } else GH_LAUNCH((gd::back_kernel<V, 32>), (cnt + 31) / 32, 32, sC, L);
        } else GH_LAUNCH((gd::back_kernel<V, 32>), (cnt + 31) / 32, 32, sC, L);   // one voice per lane, one warp per CTA
        GH_CUDA(cudaGetLastError());
        GH_CUDA(cudaEventRecord(evT1[i], sC));
        GH_CUDA(cudaEventRecord(evC[b], sC));
        if ((int)evChunk.size() <= i) { cudaEvent_t e; GH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evChunk.push_back(e); }
        GH_CUDA(cudaEventRecord(evChunk[i], sC));
        g_launches.fetch_add(2, std::memory_order_relaxed);
      }
      launched_chunks = i;
      GH_CUDA(cudaEventRecord(evDoneC, sC));
      GH_CUDA(cudaStreamWaitEvent(parent, evDoneC, 0));
      GH_CUDA(cudaStreamWaitEvent(parent, evDoneS, 0));
      if (getenv("GOOEY_B200_TRACE_TYPES")) {      // diagnostic (serialises the buckets): time of the time-parallel chain and of the per-sample kernel, voices on the latter
        cudaEvent_t t0, t1, t2;
        cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventCreate(&t2);
        cudaEventRecord(t1, sC); cudaEventRecord(t2, sS);
        GH_CUDA(cudaStreamSynchronize(sC)); GH_CUDA(cudaStreamSynchronize(sS));
        std::vector<uint8_t> mode(cnt);
        GH_CUDA(cudaMemcpy(mode.data(), d_mode.p, cnt, cudaMemcpyDeviceToHost));
        int slow = 0; for (uint8_t m : mode) slow += m != 0;
        float ms_all = 0.0f;
        cudaEventElapsedTime(&ms_all, evT0[0], t1);
        float ms_slow_end = 0.0f;
        cudaEventElapsedTime(&ms_slow_end, evT0[0], t2);
        fprintf(stderr, "[gooey trace]   %s: %d voices x %d frames, %d on the per-sample path; back ends done %.1f ms, per-sample kernel done %.1f ms after the first back-end launch\n",
                backend_name, cnt, frames, slow, ms_all, ms_slow_end);
        cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(t2);
      }
    } else {
      if constexpr (std::is_same<V, gd::GranV>::value) {
        const char* ge = getenv("GOOEY_B200_GRAN");
        if (!(ge && strcmp(ge, "serial") == 0)) {       // one warp per granulator, grains on lanes (gran_wave.cuh)
          GH_LAUNCH((gd::gran_wave_kernel<4>), (cnt + 3) / 4, 128, sC, L);
          g_launches.fetch_add(1, std::memory_order_relaxed);
          GH_CUDA(cudaGetLastError());
          GH_CUDA(cudaEventRecord(evDoneC, sC));
          GH_CUDA(cudaStreamWaitEvent(parent, evDoneC, 0));
          return;
        }
      }
      if constexpr (std::is_same<V, gd::BassV>::value) {
        const char* be = getenv("GOOEY_B200_BASS");
        if (!(be && strcmp(be, "serial") == 0) && !modulated) {      // LFO-routed basses take the per-sample kernel       // one warp per bass voice, lane = frame (bass_wave.cuh)
          const bool ttrace = getenv("GOOEY_B200_TRACE_TYPES") != nullptr;
          cudaEvent_t t0 = nullptr, t1 = nullptr;
          if (ttrace) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0, sC); }
          GH_LAUNCH((gd::bass_wave_kernel<2>), (cnt + 1) / 2, 64, sC, L);
          if (ttrace) {
            cudaEventRecord(t1, sC); GH_CUDA(cudaStreamSynchronize(sC));
            float ms = 0.0f; cudaEventElapsedTime(&ms, t0, t1);
            fprintf(stderr, "[gooey trace]   bass_wave_kernel: %d voices x %d frames, %.1f ms\n", cnt, frames, ms);
            cudaEventDestroy(t0); cudaEventDestroy(t1);
          }
          g_launches.fetch_add(1, std::memory_order_relaxed);
          GH_CUDA(cudaGetLastError());
          GH_CUDA(cudaEventRecord(evDoneC, sC));
          GH_CUDA(cudaStreamWaitEvent(parent, evDoneC, 0));
          return;
        }
      }
      if (cnt <= 148 * 32 * 4) GH_LAUNCH((gd::slow_kernel<V, 32>), (cnt + 31) / 32, 32, sC, L);
      else GH_LAUNCH((gd::slow_kernel<V, 128>), (cnt + 127) / 128, 128, sC, L);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      GH_CUDA(cudaGetLastError());
      GH_CUDA(cudaEventRecord(evDoneC, sC));
      GH_CUDA(cudaStreamWaitEvent(parent, evDoneC, 0));
    }
  }
};

}  // namespace gh
