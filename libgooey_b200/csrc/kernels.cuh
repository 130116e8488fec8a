// kernels.cuh — the voice-per-thread render kernel and its launch helpers.
//
// Mapping: one voice per thread; the voice's whole DSP
// state is loaded word-interleaved (SoA, coalesced) into registers/local
// memory, ticked sample-serially for the launch's frame range, and stored
// back.  Output leaves in one of two layouts:
//   OUT_VOICE_MAJOR  out[voice * stride + frame]  — final, host-facing.  Each
//                    warp stages a 32 voices x 32 frames tile in shared memory
//                    and writes it as 128-bit stores, 128 B contiguous per voice.
//   OUT_TIME_MAJOR   out[frame * stride + slot]    — intermediate buffers the
//                    engine mix kernel consumes; lanes write adjacent slots, so
//                    the store is coalesced without staging.
#pragma once
#include <cuda_runtime.h>
#include "voices2.cuh"

namespace gd {

enum { OUT_VOICE_MAJOR = 0, OUT_TIME_MAJOR = 1 };

struct VoiceLaunch {
  uint32_t* state;        // [words][n_pad]  (n_pad = pool capacity / row pitch in words)
  int n, n_pad;
  const uint32_t* slots;     // optional launch index -> pool slot (nullptr = identity)
  const uint32_t* out_slots; // optional launch index -> time-major output slot (nullptr = slot0 + index)
  const VoiceEvent* events;
  const uint32_t* ev_begin;  // [n+1] offsets into events
  uint32_t* ev_cursor;       // [n] running cursor (persists across chunked launches of one render)
  uint32_t frame0;           // first frame of this launch (event frames are relative to render start)
  int frames;                // frames in this launch
  float* out;
  long long stride;
  int layout;
  int slot0;                 // first slot (time-major) / first row (voice-major)
  const uint32_t* rows;      // optional voice -> output row map (voice-major); nullptr = slot0 + voice
  RateCtx rc;
};

struct KickV  { using State = KickState;  static __device__ __forceinline__ float tick(State& s, const RateCtx& rc) { return kick_tick(s, rc); }
                static __device__ __forceinline__ void event(State& s, const VoiceEvent& e, const RateCtx&) { kick_event(s, e); } };
struct SnareV { using State = SnareState; static __device__ __forceinline__ float tick(State& s, const RateCtx& rc) { return snare_tick(s, rc); }
                static __device__ __forceinline__ void event(State& s, const VoiceEvent& e, const RateCtx&) { snare_event(s, e); } };
struct HatV   { using State = HatState;   static __device__ __forceinline__ float tick(State& s, const RateCtx& rc) { return hat_tick(s, rc); }
                static __device__ __forceinline__ void event(State& s, const VoiceEvent& e, const RateCtx&) { hat_event(s, e); } };
struct TomV   { using State = TomState;   static __device__ __forceinline__ float tick(State& s, const RateCtx& rc) { return tom_tick(s, rc); }
                static __device__ __forceinline__ void event(State& s, const VoiceEvent& e, const RateCtx& rc) { tom_event(s, e, rc.sr); } };

template <class S> __device__ __forceinline__ void load_state(S& s, const uint32_t* base, int v, int n_pad) {
  constexpr int W = sizeof(S) / 4;
  uint32_t* w = reinterpret_cast<uint32_t*>(&s);
#pragma unroll 8
  for (int i = 0; i < W; i++) w[i] = base[(size_t)i * n_pad + v];
}
template <class S> __device__ __forceinline__ void store_state(const S& s, uint32_t* base, int v, int n_pad) {
  constexpr int W = sizeof(S) / 4;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&s);
#pragma unroll 8
  for (int i = 0; i < W; i++) base[(size_t)i * n_pad + v] = w[i];
}

constexpr int TILE = 32;

// Warp-cooperative store of a 32x32 tile (tile[lane][frame]) to voice-major output.
__device__ __forceinline__ void store_tile_voice_major(const float* tile /*[32][33]*/, float* out, long long stride, int row0,
                                                        const uint32_t* rows, int n_rows, int f0, int nf, int lane) {
  const bool vec_ok = (nf == TILE) && ((stride & 3) == 0) && ((f0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec_ok) {
    const int c = (lane & 7) * 4;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int r = k * 4 + (lane >> 3);
      if (r < n_rows) {
        const float* src = tile + r * 33 + c;
        float4 v = make_float4(src[0], src[1], src[2], src[3]);
        const long long row = rows ? (long long)rows[r] : (long long)(row0 + r);
        *reinterpret_cast<float4*>(out + row * stride + f0 + c) = v;
      }
    }
  } else {
    for (int r = 0; r < n_rows; r++) {
      const long long row = rows ? (long long)rows[r] : (long long)(row0 + r);
      if (lane < nf) out[row * stride + f0 + lane] = tile[r * 33 + lane];
    }
  }
}

template <class V, int BLOCK>
__global__ void __launch_bounds__(BLOCK) voice_kernel(const VoiceLaunch L) {
  __shared__ float tiles[BLOCK / 32][TILE * 33];
  const int v = blockIdx.x * BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool valid = v < L.n;
  const int warp_v0 = v - lane;
  if (warp_v0 >= L.n) return;  // whole warp idle
  typename V::State st;
  uint32_t ev = 0, ev_end = 0;
  int sv = v, ov = L.slot0 + v;
  if (valid) {
    if (L.slots) sv = (int)L.slots[v];
    if (L.out_slots) ov = (int)L.out_slots[v];
    load_state(st, L.state, sv, L.n_pad);
    ev = L.ev_cursor[v];
    ev_end = L.ev_begin[v + 1];
  }
  float* tile = tiles[warp];
  const int n_rows = min(32, L.n - warp_v0);
  for (int f0 = 0; f0 < L.frames; f0 += TILE) {
    const int nf = min(TILE, L.frames - f0);
    if (valid) {
      for (int j = 0; j < nf; j++) {
        const uint32_t frame = L.frame0 + f0 + j;
        while (ev < ev_end && L.events[ev].frame <= frame) {
          const VoiceEvent e = L.events[ev];
          if (e.kind == EV_SET_TIME) st.t = (double)e.value; else V::event(st, e, L.rc);
          ev++;
        }
        const float y = V::tick(st, L.rc);
        if (L.layout == OUT_TIME_MAJOR) L.out[(long long)(f0 + j) * L.stride + ov] = y;
        else tile[lane * 33 + j] = y;
      }
    }
    if (L.layout == OUT_VOICE_MAJOR) {
      __syncwarp();
      store_tile_voice_major(tile, L.out, L.stride, L.slot0 + warp_v0, L.rows ? L.rows + warp_v0 : nullptr, n_rows, L.frame0 + f0, nf, lane);
      __syncwarp();
    }
  }
  if (valid) {
    store_state(st, L.state, sv, L.n_pad);
    L.ev_cursor[v] = ev;
  }
}

}  // namespace gd
