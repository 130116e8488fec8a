// tests/emu/emu.cpp — TEST INFRASTRUCTURE ONLY (never loaded by the product).
// Host execution of the SAME device functions the kernels call (libgooey_b200/csrc/*.cuh compiled by g++), driven by
// plain loops that mirror kernels.cuh: mode 0 = general path (slow_kernel: tick per sample), mode 1 = split path
// (plan_kernel -> front_kernel -> back_kernel).  Lets the CPU test-suite check the kernel logic — span planning,
// envelope latch search, pure front evaluation — against the oracle without a GPU.
#include <algorithm>
#include <cstdio>
#include <vector>
#include "../../libgooey_b200/csrc/patch.h"
#include "../../libgooey_b200/csrc/halfband_design.h"

namespace gd {
float g_hb_host[8];
double g_midi_freq_host[128];
}
using namespace gd;

static std::vector<double> make_clock(float sr, size_t n) {
  std::vector<double> t(n);
  double dt = 1.0 / (double)sr, x = 0.0;
  for (size_t i = 0; i < n; i++) { t[i] = x; x += dt; }
  return t;
}

struct KickE { using State = KickState; using Span = KickSpan; enum { NPL = KICK_PLANES };
  static bool settled(const KickCtl& c) { return params_settled<K_NP>(c.cur, c.tgt); }
  static void event(KickCtl& c, const VoiceEvent& e, const double* tt, uint32_t& r) { kick_event(c, e, tt, r); }
  static void slow_event(State& s, const VoiceEvent& e, const double* tt, const RateCtx&) { uint32_t r = 0; kick_event(s.c, e, tt, r); kick_span_begin(s.a, s.c, r); }
  static float tick(State& s, const double* tt, const RateCtx& rc) { return kick_tick(s, tt, rc); }
  static void plan(KickCtl& c, uint32_t r, const double* tt, int ja, int jb, Span& sp) { kick_plan(c, r, tt, ja, jb, sp); }
  static int front_end(const Span& sp) { return sp.j_act; }
  static void front(const Span& sp, double now, uint32_t, float sr, float* o) { KickFront f = kick_front(sp.c, sp.d, now, sr); o[0] = f.p1; o[1] = f.raw_click; o[2] = f.ne; o[3] = f.amp; }
  static void span_begin(KickAud& a, const Span& sp, float) { kick_span_begin(a, sp.c, sp.resets); }
  static float back(KickAud& a, const Span& sp, int j, const float* p, const RateCtx& rc) {
    if (j >= sp.j_act) return 0.0f;
    KickFront f; f.p1 = p[0]; f.raw_click = p[1]; f.ne = p[2]; f.amp = p[3];
    return kick_back(a, sp.d, f, rc); } };
struct SnareE { using State = SnareState; using Span = SnareSpan; enum { NPL = SNARE_PLANES };
  static bool settled(const SnareCtl& c) { return params_settled<S_NP>(c.cur, c.tgt); }
  static void event(SnareCtl& c, const VoiceEvent& e, const double* tt, uint32_t& r) { snare_event(c, e, tt, r); }
  static void slow_event(State& s, const VoiceEvent& e, const double* tt, const RateCtx&) { uint32_t r = 0; snare_event(s.c, e, tt, r); snare_span_begin(s.a, s.c, r); }
  static float tick(State& s, const double* tt, const RateCtx& rc) { return snare_tick(s, tt, rc); }
  static void plan(SnareCtl& c, uint32_t r, const double* tt, int ja, int jb, Span& sp) { snare_plan(c, r, tt, ja, jb, sp); }
  static int front_end(const Span& sp) { return sp.j_act; }
  static void front(const Span& sp, double now, uint32_t, float sr, float* o) { SnareFront f = snare_front(sp.c, sp.d, now, sr); o[0] = f.tonal_out; o[1] = f.raw_noise; o[2] = f.cne; o[3] = f.crack_out; o[4] = f.amp; }
  static void span_begin(SnareAud& a, const Span& sp, float) { snare_span_begin(a, sp.c, sp.resets); }
  static float back(SnareAud& a, const Span& sp, int j, const float* p, const RateCtx& rc) {
    if (j >= sp.j_act) return 0.0f;
    SnareFront f; f.tonal_out = p[0]; f.raw_noise = p[1]; f.cne = p[2]; f.crack_out = p[3]; f.amp = p[4];
    return snare_back(a, sp.d, f, rc); } };
struct HatE { using State = HatState; using Span = HatSpan; enum { NPL = HAT_PLANES };
  static bool settled(const HatCtl& c) { return params_settled<H_NP>(c.cur, c.tgt); }
  static void event(HatCtl& c, const VoiceEvent& e, const double* tt, uint32_t& r) { hat_event(c, e, tt, r); }
  static void slow_event(State& s, const VoiceEvent& e, const double* tt, const RateCtx&) { uint32_t r = 0; hat_event(s.c, e, tt, r); hat_span_begin(s.a, s.c, r); }
  static float tick(State& s, const double* tt, const RateCtx& rc) { return hat_tick(s, tt, rc); }
  static void plan(HatCtl& c, uint32_t r, const double* tt, int ja, int jb, Span& sp) { hat_plan(c, r, tt, ja, jb, sp); }
  static int front_end(const Span& sp) { return sp.j_env < sp.j1 ? sp.j_env : sp.j1; }
  static void front(const Span& sp, double now, uint32_t, float sr, float* o) { o[0] = hat_front(sp.c, sp.d, now, sr).env; }
  static void span_begin(HatAud& a, const Span& sp, float) { hat_span_begin(a, sp.c, sp.resets); }
  static float back(HatAud& a, const Span& sp, int j, const float* p, const RateCtx& rc) {
    if (!a.active) return 0.0f;
    HatFront f; f.env = j < sp.j_env ? p[0] : sp.env_final;
    return hat_back(a, sp.d, f, j >= sp.j_env, rc); } };
struct TomE { using State = TomState; using Span = TomSpan; enum { NPL = TOM_PLANES };
  static bool settled(const TomCtl&) { return true; }
  static void event(TomCtl& c, const VoiceEvent& e, const double* tt, uint32_t& r) { tom_event(c, e, tt, r); }
  static void slow_event(State& s, const VoiceEvent& e, const double* tt, const RateCtx& rc) {
    uint32_t r = 0; tom_event(s.c, e, tt, r); tom_span_begin(s.a, s.c.mem_q_scale, s.c.mem_gain_scale, s.c.mem_dirty, r, rc.sr); s.c.mem_dirty = 0; }
  static float tick(State& s, const double* tt, const RateCtx& rc) { return tom_tick(s, tt, rc); }
  static void plan(TomCtl& c, uint32_t r, const double* tt, int ja, int jb, Span& sp) { tom_plan(c, r, tt, ja, jb, sp); }
  static int front_end(const Span& sp) { return sp.d.membrane > 0.0f ? sp.j1 : (sp.j_env < sp.j1 ? sp.j_env + 1 : sp.j1); }
  static void front(const Span& sp, double now, uint32_t k, float, float* o) { TomFront f = tom_front(sp.c, sp.d, now, k); o[0] = f.env; o[1] = f.noise; o[2] = f.rnd; }
  static void span_begin(TomAud& a, const Span& sp, float sr) { tom_span_begin(a, sp.mem_q_scale, sp.mem_gain_scale, sp.mem_dirty, sp.resets, sr); }
  static float back(TomAud& a, const Span& sp, int j, const float* p, const RateCtx& rc) {
    if (!a.active) return 0.0f;
    TomFront f; f.env = j < sp.j_env ? p[0] : sp.env_final; f.noise = p[1]; f.rnd = p[2];
    return tom_back(a, sp.d, f, j >= sp.j_env, rc); } };

// returns 1 if the split path was taken, 0 if the voice had to use the general path
template <class V> static int run_voice(typename V::State& st, const std::vector<VoiceEvent>& ev, const double* tt, const RateCtx& rc, int frames, int mode, float* out) {
  using Span = typename V::Span;
  if (mode == 1) {
    // A
    auto c = st.c;
    std::vector<Span> spans;
    size_t e = 0;
    bool fast = true;
    int j = 0;
    while (fast && j < frames) {
      uint32_t resets = 0;
      while (e < ev.size() && ev[e].frame <= (uint32_t)j) { V::event(c, ev[e], tt, resets); e++; }
      if (!V::settled(c)) { fast = false; break; }
      int jn = frames;
      if (e < ev.size() && ev[e].frame < (uint32_t)jn) jn = (int)ev[e].frame;
      Span sp; memset(&sp, 0, sizeof sp);
      V::plan(c, resets, tt, j, jn, sp);
      spans.push_back(sp);
      j = jn;
    }
    if (fast) {
      st.c = c;
      // B: planes; entries that B does not write are poisoned with NaN so that C reading them shows up
      std::vector<float> planes((size_t)V::NPL * frames, NAN);
      for (const Span& sp : spans)
        for (int jj = sp.j0; jj < std::min(V::front_end(sp), sp.j1); jj++) {
          float o[V::NPL];
          uint32_t k = sp.kbase + (uint32_t)jj;
          V::front(sp, tt[k], k, rc.sr, o);
          for (int p = 0; p < V::NPL; p++) planes[(size_t)p * frames + jj] = o[p];
        }
      // C
      for (const Span& sp : spans) {
        V::span_begin(st.a, sp, rc.sr);
        for (int jj = sp.j0; jj < sp.j1; jj++) {
          float p[V::NPL];
          for (int q = 0; q < V::NPL; q++) p[q] = planes[(size_t)q * frames + jj];
          out[jj] = V::back(st.a, sp, jj, p, rc);
        }
      }
      return 1;
    }
  }
  size_t e = 0;
  for (int j = 0; j < frames; j++) {
    while (e < ev.size() && ev[e].frame <= (uint32_t)j) { V::slow_event(st, ev[e], tt, rc); e++; }
    out[j] = V::tick(st, tt, rc);
  }
  return 0;
}

extern "C" {

// Same event vocabulary as orc_render_voices (kind 0 trigger, 1 set_param, 2 set_param + snap).  The render is done in
// `n_calls` consecutive calls of frames/n_calls frames (state carried) to exercise call boundaries.  took_fast[v] is
// the number of calls of voice v that went through the split path.
int emu_render_voices(const GooeyVoicePatch* patches, uint32_t n, float sr, uint32_t frames, uint32_t n_events,
                      const uint32_t* ev_voice, const uint32_t* ev_frame, const uint32_t* ev_kind, const uint32_t* ev_param,
                      const float* ev_value, float* out, int mode, int n_calls, int* took_fast) {
  design_halfband8(g_hb_host);
  const RateCtx rc = make_rate_ctx(sr);
  std::vector<double> clock = make_clock(sr, (size_t)frames + 2);
  const double* tt = clock.data();
  if (n_calls < 1) n_calls = 1;
  for (uint32_t v = 0; v < n; v++) {
    std::vector<VoiceEvent> all;
    for (uint32_t e = 0; e < n_events; e++) {
      if (ev_voice[e] != v) continue;
      auto add = [&](uint32_t kind, uint32_t p, float val) { VoiceEvent x; x.frame = ev_frame[e]; x.kind = (uint16_t)kind; x.param = (uint16_t)p; x.value = val; x.aux = 0; all.push_back(x); };
      if (ev_kind[e] == 0) add(EV_TRIGGER, 0, ev_value[e]);
      else { bool known = gh::ffi_param_to_events(patches[v].instrument, ev_param[e], ev_value[e], add); if (known && ev_kind[e] == 2) add(EV_SNAP, 0, 0.0f); }
    }
    std::stable_sort(all.begin(), all.end(), [](const VoiceEvent& a, const VoiceEvent& b) { return a.frame < b.frame; });
    float* o = out + (size_t)v * frames;
    int fast_calls = 0;
    auto run_calls = [&](auto tag, auto& st) {
      using V = decltype(tag);
      uint32_t f0 = 0;
      for (int c = 0; c < n_calls; c++) {
        uint32_t f1 = c == n_calls - 1 ? frames : (uint32_t)((uint64_t)frames * (c + 1) / n_calls);
        std::vector<VoiceEvent> ev;
        for (auto x : all) if (x.frame >= f0 && x.frame < f1) { x.frame -= f0; ev.push_back(x); }
        fast_calls += run_voice<V>(st, ev, tt, rc, (int)(f1 - f0), mode, o + f0);
        f0 = f1;
      }
    };
    switch (patches[v].instrument) {
      case GOOEY_INSTRUMENT_KICK: { KickState s; gh::init_from_patch(s, patches[v], sr); run_calls(KickE{}, s); } break;
      case GOOEY_INSTRUMENT_SNARE: { SnareState s; gh::init_from_patch(s, patches[v], sr); run_calls(SnareE{}, s); } break;
      case GOOEY_INSTRUMENT_HIHAT: { HatState s; gh::init_from_patch(s, patches[v], sr); run_calls(HatE{}, s); } break;
      case GOOEY_INSTRUMENT_TOM: { TomState s; gh::init_from_patch(s, patches[v], sr); run_calls(TomE{}, s); } break;
      case GOOEY_INSTRUMENT_BASS: {      // no split path: the per-sample tick (what bass_wave_kernel falls back to, block by block)
        BassState s; gh::init_from_patch(s, patches[v], sr);
        size_t e = 0;
        for (uint32_t j = 0; j < frames; j++) {
          while (e < all.size() && all[e].frame <= j) { bass_event(s, all[e], tt); e++; }
          o[j] = bass_tick(s, tt, rc);
        }
      } break;
      default: return -1;
    }
    if (took_fast) took_fast[v] = fast_calls;
  }
  return 0;
}

// Bare poly synth / granulator through the device functions of voices2.cuh (what slow_kernel<PolyV> / the granulator's control lane run),
// same event vocabulary and clock as orc_poly_render / orc_gran_render.
int emu_poly_render(uint32_t preset, float sr, uint32_t n_ev, const uint32_t* ev_frame, const uint32_t* ev_kind, const float* ev_a, const float* ev_b,
                    uint32_t frames, float* out) {
  static const float T[5][14] = {
      {0.0f, 0.2f, 0.6f, 0.15f, 0.3f, 0.55f, 0.7f, 0.7f, 0.8f, 0.5f, 0.65f, 0.4f, 0.75f, 0.7f},
      {0.0f, 0.4f, 0.45f, 0.2f, 0.2f, 0.8f, 0.75f, 0.8f, 0.85f, 0.75f, 0.7f, 0.5f, 0.8f, 0.6f},
      {0.3f, 0.1f, 0.7f, 0.25f, 0.6f, 0.0f, 0.75f, 0.0f, 0.65f, 0.0f, 0.7f, 0.1f, 0.65f, 0.7f},
      {0.5f, 0.15f, 0.55f, 0.1f, 0.4f, 0.35f, 0.7f, 0.5f, 0.75f, 0.3f, 0.65f, 0.3f, 0.7f, 0.7f},
      {0.0f, 0.5f, 0.5f, 0.1f, 0.15f, 0.85f, 0.7f, 0.9f, 0.85f, 0.8f, 0.7f, 0.6f, 0.8f, 0.5f}};   // PolySynthConfig presets (poly_synth.rs:49-142), as engine.cuh passes them
  for (int n = 0; n < 128; n++) g_midi_freq_host[n] = 440.0 * pow(2.0, ((double)n - 69.0) / 12.0);
  const RateCtx rc = make_rate_ctx(sr);
  std::vector<double> clock = make_clock(sr, (size_t)frames + 2);
  const double* tt = clock.data();
  PolyState s; memset(&s, 0, sizeof s);
  poly_init(s, T[preset < 5 ? preset : 0], sr);
  uint32_t e = 0;
  for (uint32_t j = 0; j < frames; j++) {
    while (e < n_ev && ev_frame[e] <= j) {
      VoiceEvent x; memset(&x, 0, sizeof x); x.frame = j;
      if (ev_kind[e] == 0) { x.kind = EV_POLY_NOTE; x.param = (uint16_t)ev_a[e]; x.value = ev_b[e]; }
      else if (ev_kind[e] == 1) x.kind = EV_POLY_RELEASE;
      else { x.kind = EV_SET_TARGET; x.param = (uint16_t)ev_a[e]; x.value = ev_b[e]; }
      poly_event(s, x, tt);
      e++;
    }
    out[j] = poly_tick(s, tt, rc);
  }
  return 0;
}
int emu_gran_render(float sr, const float* buf, uint32_t buf_len, float buf_sr, uint32_t n_ev, const uint32_t* ev_frame, const uint32_t* ev_kind,
                    const float* ev_a, const float* ev_b, uint32_t frames, float* out) {
  design_halfband8(g_hb_host);
  const RateCtx rc = make_rate_ctx(sr);
  std::vector<double> clock = make_clock(sr, (size_t)frames + 2);
  const double* tt = clock.data();
  GranState s; memset(&s, 0, sizeof s);
  gran_init(s, sr);
  auto ev = [&](uint16_t kind, uint16_t param, float value, uint32_t aux) { VoiceEvent x; memset(&x, 0, sizeof x); x.kind = kind; x.param = param; x.value = value; x.aux = aux; gran_event(s, x, tt); };
  const uint64_t addr = (uint64_t)(uintptr_t)buf;                       // set_buffer as the engine encodes it: pointer halves + length / rate
  { float lo; uint32_t lo_bits = (uint32_t)(addr & 0xffffffffu); memcpy(&lo, &lo_bits, 4); ev(EV_GRAN_BUFFER, 0, lo, (uint32_t)(addr >> 32)); }
  ev(EV_SET_AUX, AUX_GRAN_BUFINFO, buf_sr, buf_len);
  uint32_t e = 0;
  for (uint32_t j = 0; j < frames; j++) {
    while (e < n_ev && ev_frame[e] <= j) {
      switch (ev_kind[e]) {
        case 0: ev(EV_TRIGGER, 0, ev_b[e], 0); break;
        case 2: ev(EV_SET_TARGET, (uint16_t)ev_a[e], ev_b[e], 0); break;
        case 3: ev(EV_SNAP, 0, 0.0f, 0); break;
        case 4: ev(EV_GRAN_SEED, 0, 0.0f, (uint32_t)ev_a[e]); break;
      }
      e++;
    }
    out[j] = gran_tick(s, tt, rc);
  }
  return 0;
}

// The front end's arithmetic shortcuts, evaluated by the host build of the very functions the kernels call (gmath.cuh):
// kind 0 = g_div_by(a, b, 1 / b) (must equal the IEEE quotient bit for bit), 1 = g_sinf_fast(a), 2 = the additive triangle with
// FAST = true (osc_triangle<true>(a, b, sr)), 3 = the same with the bit-exact sine.
int emu_math(int kind, const float* a, const float* b, float sr, float* out, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) {
    switch (kind) {
      case 0: { volatile float y = 1.0f / b[i]; out[i] = gm::g_div_by(a[i], b[i], y); } break;
      case 1: out[i] = gm::g_sinf_fast(a[i]); break;
      case 2: out[i] = osc_triangle<true>(a[i], b[i], sr); break;
      case 3: out[i] = osc_triangle<false>(a[i], b[i], sr); break;
      default: return -1;
    }
  }
  return 0;
}

}  // extern "C"

// ---- sample-playback sources (libgooey_b200/csrc/loops.cuh): the loop mixer and the sampler rack ticks on the host ----------
#include "../../libgooey_b200/csrc/loops.cuh"
static LoopChan g_pending[LOOP_CHANNELS];     // queued takes for the next emu_loop_mixer call (has_pending = 0: none)
static uint32_t g_swaps[LOOP_CHANNELS];
extern "C" {
void emu_loop_queue(int k, const float* left, const float* right, uint32_t len, float buf_sr, double warp, double warp_pp, uint32_t divisions) {
  LoopChan& p = g_pending[k];
  p.pend_left = left; p.pend_right = right; p.pend_len = len; p.pend_buf_sr = buf_sr; p.pend_warp = warp; p.pend_warp_pp = warp_pp; p.pend_div = divisions; p.has_pending = left ? 1u : 0u;
}
uint32_t emu_loop_swaps(int k) { return g_swaps[k]; }
// One LoopMixer descriptor driven for `frames` frames, exactly like ext_source_kernel<LoopMixer> drives it.  Per-channel arrays
// of 4; left[k] == nullptr: nothing loaded.  cursor / gain_ct / active_ct ([4][2] = current, target) are read and written back.
int emu_loop_mixer(const float* const* left, const float* const* right, const uint32_t* len, const float* buf_sr, double* cursor, const double* warp,
                   const float* loop_start, const float* loop_end, const float* speed, const uint32_t* playing, float* gain_ct, float* active_ct,
                   float engine_sr, int frames, float* out_l, float* out_r, const uint32_t* preserve, const double* warp_pp) {
  LoopMixer m;
  memset(&m, 0, sizeof m);
  // PreservePitch channels start with no stretcher (built at their first tick); the window table as the host side of the product builds it
  const uint32_t hop = wsola_hop_len(engine_sr);
  std::vector<float> hann(2 * hop);
  for (uint32_t i = 0; i < 2 * hop; i++) hann[i] = wsola_window_coeff(i, 2 * hop);
  std::vector<std::vector<float>> stretch(LOOP_CHANNELS, std::vector<float>(9 * (size_t)hop, 123.0f));   // junk: a fresh stretcher must not read it
  for (int k = 0; k < LOOP_CHANNELS; k++) {
    m.ch[k].preserve = preserve ? preserve[k] : 0u; m.ch[k].warp_pp = warp_pp ? warp_pp[k] : 1.0;
    m.ch[k].hop = hop; m.ch[k].hann = hann.data(); m.ch[k].st_buf = stretch[k].data();
    const LoopChan& q = g_pending[k];
    m.ch[k].pend_left = q.pend_left; m.ch[k].pend_right = q.pend_right; m.ch[k].pend_len = q.pend_len; m.ch[k].pend_buf_sr = q.pend_buf_sr;
    m.ch[k].pend_warp = q.pend_warp; m.ch[k].pend_warp_pp = q.pend_warp_pp; m.ch[k].pend_div = q.pend_div; m.ch[k].has_pending = q.has_pending;
    g_pending[k].has_pending = 0;
  }
  for (int k = 0; k < LOOP_CHANNELS; k++) {
    LoopChan& c = m.ch[k];
    c.left = left[k]; c.right = right[k]; c.len = len[k]; c.buf_sr = buf_sr[k]; c.cursor = cursor[k]; c.warp = warp[k];
    c.loop_start = loop_start[k]; c.loop_end = loop_end[k]; c.speed = speed[k]; c.playing = playing[k];
    c.gain = {gain_ct[2 * k], gain_ct[2 * k + 1]}; c.active = {active_ct[2 * k], active_ct[2 * k + 1]};
  }
  const float coeff15 = smooth_coeff(engine_sr, LOOP_FADER_MS);
  for (int f = 0; f < frames; f++) loop_mixer_tick(m, engine_sr, coeff15, out_l[f], out_r[f]);
  for (int k = 0; k < LOOP_CHANNELS; k++) {
    g_swaps[k] = m.ch[k].swaps;
    cursor[k] = m.ch[k].cursor;
    gain_ct[2 * k] = m.ch[k].gain.c; gain_ct[2 * k + 1] = m.ch[k].gain.t; active_ct[2 * k] = m.ch[k].active.c; active_ct[2 * k + 1] = m.ch[k].active.t;
  }
  return 0;
}
// the WSOLA window table as the product's host side builds it (EngineBank::attach_stretcher); returns hop_len
int emu_wsola_window(float engine_sr, float* out, int cap) {
  const uint32_t hop = wsola_hop_len(engine_sr);
  for (uint32_t i = 0; i < 2 * hop && (int)i < cap; i++) out[i] = wsola_window_coeff(i, 2 * hop);
  return (int)hop;
}
// One SamplerRack descriptor with a pad table and a hit list (what the host resolves a rack pattern into), rendered in `n_pieces` kernel-like
// passes of `piece` frames with the descriptor carried over: the device starts the voices itself at the hit frames.
int emu_sampler_rack_hits(int n_pads, const float* const* pads, const uint32_t* frames, const uint32_t* channels, const double* inc,
                          int n_hits, const uint32_t* hit_frame, const uint32_t* hit_slot, const float* hit_vel, int piece, int n_pieces, float* out_l, float* out_r) {
  SamplerRack r;
  memset(&r, 0, sizeof r);
  for (int k = 0; k < n_pads && k < SAMPLER_SLOTS; k++) r.slots[k] = SamplerSlotRef{pads[k], frames[k], channels[k], inc[k]};
  std::vector<SamplerHit> hits(n_hits);
  for (int k = 0; k < n_hits; k++) hits[k] = SamplerHit{hit_frame[k], hit_slot[k], hit_vel[k], 0u};
  r.hits = hits.data(); r.n_hits = (uint32_t)n_hits;
  for (int p = 0; p < n_pieces; p++) {
    SamplerRack d = r;                                  // the kernel's local copy, written back at the end of the launch
    for (int f = 0; f < piece; f++) sampler_rack_tick(d, out_l[p * piece + f], out_r[p * piece + f]);
    r = d;
  }
  int alive = 0;
  for (int v = 0; v < SAMPLER_VOICES; v++) alive += r.v[v].samples != nullptr;
  return alive;
}
double emu_window_fold(float loop_start, float loop_end, double len, double p) { return window_fold(loop_window(loop_start, loop_end, len), p); }
double emu_window_lo(float loop_start, float loop_end, double len) { return loop_window(loop_start, loop_end, len).lo; }
// One SamplerRack descriptor: n_voices (<= 32) voices started at position 0 (voice v plays pads[v], frames[v] x channels[v], increment inc[v],
// velocity vel[v]); returns the number of voices still sounding after `n` frames.
int emu_sampler_rack(int n_voices, const float* const* pads, const uint32_t* frames, const uint32_t* channels, const double* inc, const float* vel,
                     const double* start_pos, int n, float* out_l, float* out_r) {
  SamplerRack r;
  memset(&r, 0, sizeof r);
  for (int v = 0; v < n_voices && v < SAMPLER_VOICES; v++) {
    SampleVoice& s = r.v[v];
    s.samples = pads[v]; s.frames = frames[v]; s.channels = channels[v]; s.increment = inc[v]; s.velocity = vel[v]; s.position = start_pos[v]; s.age = v + 1;
  }
  for (int f = 0; f < n; f++) sampler_rack_tick(r, out_l[f], out_r[f]);
  int alive = 0;
  for (int v = 0; v < SAMPLER_VOICES; v++) alive += r.v[v].samples != nullptr;
  return alive;
}
}
