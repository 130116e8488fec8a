// voices2.cuh — melodic / sample-based voices: BassSynth, PolySynth ("chord oscillators"), Granulator.
// Same contract as voices.cuh: {State, init, event, tick}, one voice per thread.
// Reference: src/instruments/{bass,poly_synth,granulator}.rs.
#pragma once
#include "voices.cuh"

namespace gd {

#ifdef __CUDA_ARCH__
#define G_LDG(p) __ldg(p)
#else
#define G_LDG(p) (*(p))
#endif

// polyBLEP on f64 phases (gen/polyblep.rs:8-40)
G_HD double poly_blep(double t, double dt) {
  if (t < dt) { t = t / dt; return 2.0 * t - t * t - 1.0; }
  if (t > 1.0 - dt) { t = (t - 1.0) / dt; return t * t + 2.0 * t + 1.0; }
  return 0.0;
}
G_HD float polyblep_saw(double ph, double inc) { return (float)((2.0 * ph - 1.0) - poly_blep(ph, inc)); }
G_HD float polyblep_square(double ph, double inc) {
  double naive = ph < 0.5 ? 1.0 : -1.0;
  double b1 = poly_blep(ph, inc);
  double p2 = fmod(ph + 0.5, 1.0);
  double b2 = poly_blep(p2, inc);
  return (float)(naive + b1 - b2);
}

// =========================================== Bass ===========================================
enum { B_FREQ, B_SUB, B_OSC, B_DETUNE_LEVEL, B_DETUNE_AMT, B_SHAPE, B_CUTOFF, B_RES, B_FENV_AMT, B_FENV_DECAY, B_FENV_CURVE,
       B_AMP_DECAY, B_AMP_CURVE, B_OVERDRIVE, B_VOLUME, B_TUNING, B_NP };
struct BassState {
  float cur[B_NP], tgt[B_NP];
  double sub_phase, osc_phase, detune_phase;
  Tpt filter;
  Env amp_env, flt_env;
  WShaper ws;
  float velocity, trig_freq;
  uint32_t active;
  float saved_freq; uint32_t has_saved;
  uint32_t k, pad_k;
  // pure functions of one smoothed parameter, re-evaluated only when it moves (three powf per tick otherwise)
  float memo_tuning_in, memo_tuning_out, memo_cents_in, memo_cents_out, memo_cutoff_in, memo_cutoff_out;
};
G_HD float exp_denorm(float n, float mn, float mx) { return mn * gm::g_powf(mx / mn, clampf(n, 0.0f, 1.0f)); }
// BassSynth::with_config (bass.rs:613-634); cfg = BassConfig::new order (15)
G_HD void bass_init(BassState& s, const float* cfg, float sr) {
  for (int i = 0; i < 15; i++) { float c = clampf(cfg[i], 0.0f, 1.0f); s.cur[i] = s.tgt[i] = c; }
  s.cur[B_TUNING] = s.tgt[B_TUNING] = 0.5f;
  s.sub_phase = s.osc_phase = s.detune_phase = 0.0;
  tpt_init(s.filter, sr, exp_denorm(s.cur[B_CUTOFF], 20.0f, 18000.0f), denorm(s.cur[B_RES], 0.5f, 15.0f));
  env_init(s.amp_env); env_init(s.flt_env);
  ws_init(s.ws, s.cur[B_OVERDRIVE], 1.0f);
  s.velocity = 1.0f; s.trig_freq = denorm(s.cur[B_FREQ], 30.0f, 200.0f); s.active = 0;
  s.saved_freq = 0.0f; s.has_saved = 0; s.k = 0; s.pad_k = 0;
  s.memo_tuning_in = s.memo_cents_in = s.memo_cutoff_in = -1.0f;   // no valid input is negative
  s.memo_tuning_out = s.memo_cents_out = s.memo_cutoff_out = 0.0f;
}
G_HD void bass_trigger(BassState& s, float velocity, double time) {  // bass.rs:747-791
  s.velocity = clampf(velocity, 0.0f, 1.0f);
  s.active = 1;
  s.sub_phase = s.osc_phase = s.detune_phase = 0.0;
  s.trig_freq = denorm(s.cur[B_FREQ], 30.0f, 200.0f);
  float amp_decay = denorm(s.cur[B_AMP_DECAY], 0.05f, 4.0f);
  float amp_curve = denorm(s.cur[B_AMP_CURVE], 0.1f, 10.0f);
  env_config_raw(s.amp_env, 0.002f, amp_decay, 0.0f, amp_decay * 0.1f, CURVE_LINEAR, amp_curve);
  env_trigger(s.amp_env, time);
  float fd = denorm(s.cur[B_FENV_DECAY], 0.01f, 2.0f);
  float fc = denorm(s.cur[B_FENV_CURVE], 0.1f, 8.0f);
  env_config_raw(s.flt_env, 0.001f, fd, 0.0f, fd * 0.1f, CURVE_LINEAR, fc);
  env_trigger(s.flt_env, time);
  s.filter.ic1 = s.filter.ic2 = 0.0f;
  s.ws.drive = clampf(1.0f + s.cur[B_OVERDRIVE] * 9.0f, 1.0f, 10.0f);
}
G_HD void bass_event(BassState& s, const VoiceEvent& e, const double* tt) {
  switch (e.kind) {
    case EV_TRIGGER: bass_trigger(s, e.value, tt[s.k]); break;
    case EV_SET_TIME: s.k = e.aux; break;
    case EV_SET_TARGET: { const float c = clampf(e.value, 0.0f, 1.0f);
      for (int i = 0; i < B_NP; i++) if ((uint32_t)i == e.param && fabsf(s.tgt[i] - c) > 1e-8f) s.tgt[i] = c; } break;   // select, not tgt[param]: keeps the state in registers
    case EV_SNAP: for (int i = 0; i < B_NP; i++) s.cur[i] = s.tgt[i]; break;
    case EV_NOTE_FREQ: {
      if (!s.has_saved) { s.saved_freq = s.cur[B_FREQ]; s.has_saved = 1; }
      float c = clampf(e.value, 0.0f, 1.0f);
      if (fabsf(s.tgt[B_FREQ] - c) > 1e-8f) s.tgt[B_FREQ] = c;
      for (int i = 0; i < B_NP; i++) s.cur[i] = s.tgt[i];
    } break;
    case EV_RESTORE_FREQ:
      if (s.has_saved) {
        s.has_saved = 0;
        float c = clampf(s.saved_freq, 0.0f, 1.0f);
        if (fabsf(s.tgt[B_FREQ] - c) > 1e-8f) s.tgt[B_FREQ] = c;
        for (int i = 0; i < B_NP; i++) s.cur[i] = s.tgt[i];
      }
      break;
    case EV_SET_AUX: if (e.param == AUX_OVERSAMPLING) { uint32_t m = (uint32_t)e.value; if (s.ws.os.mode != m) { s.ws.os.mode = m; os_reset(s.ws.os); } } break;
    default: break;
  }
}
G_D float bass_tick(BassState& s, const double* tt, const RateCtx& rc) {  // bass.rs:793-877
  const double now = tt[s.k];
  s.k += 1;
#pragma unroll
  for (int i = 0; i < B_NP; i++) smooth_tick(s.cur[i], s.tgt[i], rc.smooth15);
  if (!s.active) return 0.0f;
  const float sr = rc.sr;
  if (s.cur[B_TUNING] != s.memo_tuning_in) { s.memo_tuning_in = s.cur[B_TUNING]; s.memo_tuning_out = tuning_to_multiplier(s.cur[B_TUNING]); }
  float freq = s.trig_freq * s.memo_tuning_out;
  float sub_level = s.cur[B_SUB], osc_level = s.cur[B_OSC], detune_level = s.cur[B_DETUNE_LEVEL];
  float detune_cents = denorm(s.cur[B_DETUNE_AMT], 0.0f, 30.0f);
  float osc_shape = s.cur[B_SHAPE];
  if (detune_cents != s.memo_cents_in) { s.memo_cents_in = detune_cents; s.memo_cents_out = gm::g_powf(2.0f, detune_cents / 1200.0f); }
  float detune_ratio = s.memo_cents_out;
  float detune_freq = freq * detune_ratio;
  double dt = 1.0 / (double)sr;
  double sub_inc = (double)freq * dt, osc_inc = (double)freq * dt, det_inc = (double)detune_freq * dt;
  s.sub_phase += sub_inc; s.sub_phase -= floor(s.sub_phase);
  s.osc_phase += osc_inc; s.osc_phase -= floor(s.osc_phase);
  s.detune_phase += det_inc; s.detune_phase -= floor(s.detune_phase);
  float sub_out = (float)sin(s.sub_phase * 6.283185307179586476925286766559);
  float saw_m = polyblep_saw(s.osc_phase, osc_inc), sq_m = polyblep_square(s.osc_phase, osc_inc);
  float osc_out = saw_m * (1.0f - osc_shape) + sq_m * osc_shape;
  float saw_d = polyblep_saw(s.detune_phase, det_inc), sq_d = polyblep_square(s.detune_phase, det_inc);
  float det_out = saw_d * (1.0f - osc_shape) + sq_d * osc_shape;
  float mix = sub_out * sub_level + osc_out * osc_level + det_out * detune_level;
  float od = s.cur[B_OVERDRIVE];
  s.ws.drive = clampf(1.0f + od * 9.0f, 1.0f, 10.0f);
  float sat = od > 0.001f ? ws_process(s.ws, mix) : mix;
  float fenv = env_amp(s.flt_env, now);
  if (s.cur[B_CUTOFF] != s.memo_cutoff_in) { s.memo_cutoff_in = s.cur[B_CUTOFF]; s.memo_cutoff_out = exp_denorm(s.cur[B_CUTOFF], 20.0f, 18000.0f); }
  float base_cutoff = s.memo_cutoff_out;
  float env_offset = (18000.0f - base_cutoff) * s.cur[B_FENV_AMT] * fenv;
  float cutoff = clampf(base_cutoff + env_offset, 20.0f, 18000.0f);
  tpt_set(s.filter, sr, cutoff, denorm(s.cur[B_RES], 0.5f, 15.0f));
  float lo, bd, hi;
  tpt_process(s.filter, sat, lo, bd, hi);
  float amp_env = env_amp(s.amp_env, now);
  float out = lo * amp_env * sqrtf(s.velocity) * s.cur[B_VOLUME];
  if (!env_active(s.amp_env)) s.active = 0;
  return out;
}

// =========================================== PolySynth ===========================================
enum { P_SHAPE, P_DETUNE, P_CUTOFF, P_RES, P_FENV_AMT, P_AMP_A, P_AMP_D, P_AMP_S, P_AMP_R, P_FLT_A, P_FLT_D, P_FLT_S, P_FLT_R, P_VOLUME, P_NP };
struct PolyVoice {
  double frequency, phase_a, phase_b;
  Env amp_env, flt_env;
  Tpt filter;
  float velocity;
  uint32_t midi_note, active;
  uint32_t order_lo, order_hi;   // trigger_order (u64)
};
struct PolyState {
  float cur[P_NP], tgt[P_NP];
  PolyVoice v[6];
  uint32_t counter_lo, counter_hi;
  double last_tick_time;          // PolySynth.current_time: time of the most recent tick (poly_synth.rs:513)
  uint32_t k, pad_k;
};
#ifdef __CUDACC__
__constant__ double c_midi_freq[128];   // 440 * 2^((n-69)/12) in f64, filled by the host with the platform libm (music/note.rs:81-83)
#define G_MIDI_FREQ(n) c_midi_freq[n]
#else
extern double g_midi_freq_host[128];
#define G_MIDI_FREQ(n) g_midi_freq_host[n]
#endif
G_HD void poly_init(PolyState& s, const float* cfg, float sr) {  // PolySynth::with_config :268-280
  for (int i = 0; i < P_NP; i++) { float c = clampf(cfg[i], 0.0f, 1.0f); s.cur[i] = s.tgt[i] = c; }
  for (int k = 0; k < 6; k++) {
    PolyVoice& v = s.v[k];
    v.frequency = 440.0; v.phase_a = v.phase_b = 0.0;
    env_init(v.amp_env); env_init(v.flt_env);
    tpt_init(v.filter, sr, 1000.0f, 1.0f);
    v.velocity = 1.0f; v.midi_note = 0; v.active = 0; v.order_lo = v.order_hi = 0;
  }
  s.counter_lo = s.counter_hi = 0; s.last_tick_time = 0.0; s.k = 0; s.pad_k = 0;
}
G_HD float poly_env_time(float n) { return 0.001f * gm::g_powf(5000.0f, n); }
G_D void poly_trigger_note(PolyState& s, uint32_t note, float velocity) {  // :309-342
  double time = s.last_tick_time;
  int idx = -1;
  for (int k = 0; k < 6; k++) if (!s.v[k].active) { idx = k; break; }
  if (idx < 0) {
    idx = 0;
    uint64_t best = ((uint64_t)s.v[0].order_hi << 32) | s.v[0].order_lo;
    for (int k = 1; k < 6; k++) { uint64_t o = ((uint64_t)s.v[k].order_hi << 32) | s.v[k].order_lo; if (o < best) { best = o; idx = k; } }
  }
  PolyVoice& v = s.v[idx];
  v.midi_note = note;
  v.frequency = G_MIDI_FREQ(note & 127);
  v.phase_a = v.phase_b = 0.0;
  v.velocity = velocity;
  v.active = 1;
  v.order_lo = s.counter_lo; v.order_hi = s.counter_hi;
  uint64_t c = (((uint64_t)s.counter_hi << 32) | s.counter_lo) + 1;
  s.counter_lo = (uint32_t)c; s.counter_hi = (uint32_t)(c >> 32);
  env_config(v.amp_env, poly_env_time(s.cur[P_AMP_A]), poly_env_time(s.cur[P_AMP_D]), s.cur[P_AMP_S], poly_env_time(s.cur[P_AMP_R]), CURVE_LINEAR, 0.5f);
  env_trigger(v.amp_env, time);
  env_config(v.flt_env, poly_env_time(s.cur[P_FLT_A]), poly_env_time(s.cur[P_FLT_D]), s.cur[P_FLT_S], poly_env_time(s.cur[P_FLT_R]), CURVE_LINEAR, 0.5f);
  env_trigger(v.flt_env, time);
  v.filter.ic1 = v.filter.ic2 = 0.0f;
}
G_D void poly_event(PolyState& s, const VoiceEvent& e, const double* tt) {
  switch (e.kind) {
    // param 1: the voice was not ticked while untouched; catch up `current_time` = time of the engine's previous tick
    case EV_SET_TIME: s.k = e.aux; if (e.param == 1) s.last_tick_time = e.aux ? tt[e.aux - 1] : 0.0; break;
    case EV_POLY_NOTE: poly_trigger_note(s, e.param, e.value); break;
    case EV_POLY_RELEASE: { double t = s.last_tick_time; for (int k = 0; k < 6; k++) if (s.v[k].active) { env_release(s.v[k].amp_env, t); env_release(s.v[k].flt_env, t); } } break;
    case EV_SET_TARGET: { const float c = clampf(e.value, 0.0f, 1.0f);
      for (int i = 0; i < P_NP; i++) if ((uint32_t)i == e.param && fabsf(s.tgt[i] - c) > 1e-8f) s.tgt[i] = c; } break;   // select, not tgt[param]: keeps the state in registers
    case EV_SNAP: for (int i = 0; i < P_NP; i++) s.cur[i] = s.tgt[i]; break;
    default: break;
  }
}
G_D float poly_tick(PolyState& s, const double* tt, const RateCtx& rc) {  // :437-525
  const double now = tt[s.k];
  s.k += 1;
  s.last_tick_time = now;
#pragma unroll
  for (int i = 0; i < P_NP; i++) smooth_tick(s.cur[i], s.tgt[i], rc.smooth15);
  const float sr = rc.sr;
  float out = 0.0f;
  const float osc_shape = s.cur[P_SHAPE], detune = s.cur[P_DETUNE], volume = s.cur[P_VOLUME];
  for (int k = 0; k < 6; k++) {
    PolyVoice& v = s.v[k];
    float y = 0.0f;
    if (v.active) {
      float amp_env = env_amp(v.amp_env, now);
      if (!env_active(v.amp_env)) { v.active = 0; }
      else {
        float flt_env = env_amp(v.flt_env, now);
        double freq = v.frequency;
        double detune_ratio = 1.0 + (double)detune * 0.0175;
        double dt = 1.0 / (double)sr;
        double inc_a = freq * dt, inc_b = freq * detune_ratio * dt;
        float saw_a = polyblep_saw(v.phase_a, inc_a), sq_a = polyblep_square(v.phase_a, inc_a);
        float osc_a = saw_a * (1.0f - osc_shape) + sq_a * osc_shape;
        float saw_b = polyblep_saw(v.phase_b, inc_b), sq_b = polyblep_square(v.phase_b, inc_b);
        float osc_b = saw_b * (1.0f - osc_shape) + sq_b * osc_shape;
        float osc_mix = (osc_a + osc_b) * 0.5f;
        v.phase_a += inc_a; v.phase_a -= floor(v.phase_a);
        v.phase_b += inc_b; v.phase_b -= floor(v.phase_b);
        float base_cutoff = 20.0f * gm::g_powf(18000.0f / 20.0f, s.cur[P_CUTOFF]);
        float mod_cutoff = base_cutoff + s.cur[P_FENV_AMT] * flt_env * (18000.0f - base_cutoff);
        float q = 0.5f + s.cur[P_RES] * 14.5f;
        tpt_set(v.filter, sr, clampf(mod_cutoff, 20.0f, 18000.0f), q);
        float lo, bd, hi;
        tpt_process(v.filter, osc_mix, lo, bd, hi);
        y = lo * amp_env * sqrtf(v.velocity) * volume;
      }
    }
    out += y;
  }
  return out * (1.0f / 4.0f);
}

// =========================================== Granulator ===========================================
enum { G_SCAN, G_LENGTH, G_SPRAY, G_PITCH, G_DENSITY, G_TEXTURE, G_DIRECTION, G_CLOUD, G_VOLUME, G_RAND_TIMING, G_RAND_AMP, G_DRIVE, G_NP };
struct Grain { float source_pos, age, duration, speed, direction, window_shape, velocity, release_samples, release_total; uint32_t active; };
struct GranState {
  float cur[G_NP], tgt[G_NP];
  Grain grains[80];              // 64 main slots + 16 release slots (granulator.rs:13-14)
  float gc_cur, gc_tgt;          // gain_compensation (10 ms smoother)
  uint32_t cloud_active;
  double cloud_end, next_grain;
  float velocity;
  uint32_t rng;
  WShaper drive;
  uint32_t buf_lo, buf_hi, buf_len;   // device pointer to the (shared, read-only) source buffer
  float buf_sr;
  uint32_t k, pad_k;
};
G_HD void gran_init(GranState& s, float sr) {  // Granulator::with_config :332-352 with GranulatorConfig::default :188-205
  const float D[12] = {0.5f, 0.16f, 0.12f, 0.5f, 0.35f, 0.25f, 0.0f, 0.35f, 0.8f, 0.0f, 0.0f, 0.0f};
  for (int i = 0; i < G_NP; i++) s.cur[i] = s.tgt[i] = D[i];
  for (int i = 0; i < 80; i++) { Grain& g = s.grains[i]; g.source_pos = 0; g.age = 0; g.duration = 1; g.speed = 1; g.direction = 1; g.window_shape = 1; g.velocity = 1; g.release_samples = 0; g.release_total = 0; g.active = 0; }
  s.gc_cur = s.gc_tgt = 1.0f;
  s.cloud_active = 0; s.cloud_end = 0.0; s.next_grain = 0.0; s.velocity = 1.0f; s.rng = 0x1234abcdu;
  ws_init(s.drive, 4.0f, 0.0f);
  s.buf_lo = s.buf_hi = 0; s.buf_len = 0; s.buf_sr = 44100.0f; s.k = 0; s.pad_k = 0;
  (void)sr;
}
G_HD float gran_next_f32(GranState& s) { uint32_t x = s.rng; x ^= x << 13; x ^= x >> 17; x ^= x << 5; s.rng = x; return (float)x / 4294967296.0f; }
G_HD int imax(int a, int b) { return a > b ? a : b; }
G_HD int imin(int a, int b) { return a < b ? a : b; }
G_D float gran_sample(const float* buf, uint32_t len, float pos) {  // SampleBuffer::sample_interpolated :161-177
  if (len == 1) return buf[0];
  float last = (float)len - 1.0f;
  pos = clampf(pos, 0.0f, last);
  int idx = (int)floorf(pos);
  float frac = pos - (float)idx;
  int lasti = (int)len - 1;
  float p0 = G_LDG(buf + imax(0, imin(lasti, idx - 1))), p1 = G_LDG(buf + imax(0, imin(lasti, idx)));
  float p2 = G_LDG(buf + imax(0, imin(lasti, idx + 1))), p3 = G_LDG(buf + imax(0, imin(lasti, idx + 2)));
  float a0 = -0.5f * p0 + 1.5f * p1 - 1.5f * p2 + 0.5f * p3;
  float a1 = p0 - 2.5f * p1 + 2.0f * p2 - 0.5f * p3;
  float a2 = -0.5f * p0 + 0.5f * p2;
  return ((a0 * frac + a1) * frac + a2) * frac + p1;
}
G_D bool gran_steal(GranState& s, float sr) {  // :626-659
  int victim = -1; float shortest = INFINITY;
  for (int i = 0; i < 64; i++) { if (!s.grains[i].active) continue; float rem = fmaxf(s.grains[i].duration - s.grains[i].age, 0.0f); if (rem < shortest) { shortest = rem; victim = i; } }
  if (victim < 0) return false;
  int rs = -1;
  for (int i = 64; i < 80; i++) if (!s.grains[i].active) { rs = i; break; }
  if (rs < 0) return false;
  float release = fmaxf(4.0f * 0.001f * sr, 1.0f);
  float remaining = fmaxf(s.grains[victim].duration - s.grains[victim].age, 1.0f);
  release = fminf(release, remaining);
  s.grains[rs] = s.grains[victim];
  s.grains[rs].release_samples = release; s.grains[rs].release_total = release;
  s.grains[victim].active = 0;
  return true;
}
G_D void gran_spawn(GranState& s, float sr) {  // :546-620
  float amp_jitter = gran_next_f32(s);
  int slot = -1;
  for (int i = 0; i < 64; i++) if (!s.grains[i].active) { slot = i; break; }
  if (slot < 0) {
    if (!gran_steal(s, sr)) return;
    for (int i = 0; i < 64; i++) if (!s.grains[i].active) { slot = i; break; }
    if (slot < 0) return;
  }
  float last_sample = (float)(s.buf_len - 1);
  float scan = clampf(s.cur[G_SCAN], 0.0f, 1.0f) * last_sample;
  float sp = clampf(s.cur[G_SPRAY], 0.0f, 1.0f);
  float spray_samples = sp * sp * sp * 10.0f * s.buf_sr;
  float spray_offset = (gran_next_f32(s) * 2.0f - 1.0f) * spray_samples;
  float req = clampf(scan + spray_offset, 0.0f, last_sample);
  float direction = gran_next_f32(s) < s.cur[G_DIRECTION] ? -1.0f : 1.0f;
  float speed = 0.25f * gm::g_powf(4.0f / 0.25f, clampf(s.cur[G_PITCH], 0.0f, 1.0f)) * (s.buf_sr / sr);
  float gl = clampf(s.cur[G_LENGTH], 0.0f, 1.0f);
  float duration = fmaxf((5.0f + gl * gl * (3000.0f - 5.0f)) * 0.001f * sr, 1.0f);
  float wshape = 0.5f + clampf(s.cur[G_TEXTURE], 0.0f, 1.0f) * 3.5f;
  float travel = duration * speed;
  float source_pos;
  if (travel >= last_sample) { duration = fmaxf(last_sample / speed, 1.0f); source_pos = direction < 0.0f ? last_sample : 0.0f; }
  else if (direction < 0.0f) source_pos = clampf(req, travel, last_sample);
  else source_pos = clampf(req, 0.0f, last_sample - travel);
  float random_amp = clampf(s.cur[G_RAND_AMP], 0.0f, 1.0f);
  float amp_factor = 1.0f - random_amp * amp_jitter;
  Grain& g = s.grains[slot];
  g.active = 1; g.source_pos = source_pos; g.age = 0.0f; g.duration = duration; g.speed = speed; g.direction = direction;
  g.window_shape = wshape; g.velocity = s.velocity * amp_factor; g.release_samples = 0.0f; g.release_total = 0.0f;
}
G_D void gran_event(GranState& s, const VoiceEvent& e, const double* tt) {
  switch (e.kind) {
    case EV_SET_TIME: s.k = e.aux; break;
    case EV_TRIGGER: {  // :722-728 (cloud length from the TARGET of cloud_duration)
      s.velocity = clampf(e.value, 0.0f, 1.0f);
      s.cloud_active = 1;
      float c = clampf(s.tgt[G_CLOUD], 0.0f, 1.0f);
      s.cloud_end = tt[s.k] + (double)(50.0f + c * c * (8000.0f - 50.0f)) * 0.001;
      s.next_grain = tt[s.k];
    } break;
    case EV_SET_TARGET: { const float c = clampf(e.value, 0.0f, 1.0f);
      for (int i = 0; i < G_NP; i++) if ((uint32_t)i == e.param && fabsf(s.tgt[i] - c) > 1e-8f) s.tgt[i] = c; } break;   // select, not tgt[param]: keeps the state in registers
    case EV_SNAP: for (int i = 0; i < G_NP; i++) s.cur[i] = s.tgt[i]; s.gc_cur = s.gc_tgt; break;
    case EV_GRAN_SEED: s.rng = e.aux == 0 ? 0x6d2b79f5u : e.aux; break;
    case EV_GRAN_BUFFER:
      s.buf_lo = gm::asuint(e.value); s.buf_hi = e.aux;
      for (int i = 0; i < 80; i++) s.grains[i].active = 0;
      s.cloud_active = 0;
      break;
    case EV_SET_AUX: if (e.param == AUX_GRAN_BUFINFO) { s.buf_sr = e.value; s.buf_len = e.aux; } break;
    default: break;
  }
}
G_D float gran_tick(GranState& s, const double* tt, const RateCtx& rc) {  // :730-742
  const double now = tt[s.k];
  s.k += 1;
#pragma unroll
  for (int i = 0; i < G_NP; i++) smooth_tick(s.cur[i], s.tgt[i], rc.smooth15);
  const float sr = rc.sr;
  const float* buf = reinterpret_cast<const float*>(((uint64_t)s.buf_hi << 32) | s.buf_lo);
  // spawn_due_grains :511-544
  if (s.cloud_active) {
    if (now > s.cloud_end) s.cloud_active = 0;
    else {
      float density = clampf(s.cur[G_DENSITY], 0.0f, 1.0f) * 80.0f;
      if (density > 0.0f && buf != nullptr && s.buf_len > 0) {
        double interval = 1.0 / (double)density;
        double random_timing = (double)clampf(s.cur[G_RAND_TIMING], 0.0f, 1.0f);
        int guard = 0;
        while (s.cloud_active && now + 1e-12 >= s.next_grain && guard < 8) {
          gran_spawn(s, sr);
          s.next_grain += interval;
          if (random_timing > 0.0) {
            double jitter = ((double)gran_next_f32(s) * 2.0 - 1.0) * interval * random_timing;
            s.next_grain = fmax(s.next_grain + jitter, now);
          }
          if (s.next_grain > s.cloud_end) s.cloud_active = 0;
          guard++;
        }
      }
    }
  }
  // tick_grains :661-718
  int active = 0;
  for (int i = 0; i < 80; i++) active += s.grains[i].active;
  float raw = 0.0f;
  if (active == 0) {
    if (fabsf(s.gc_tgt - 1.0f) > 1e-8f) s.gc_tgt = 1.0f;
    smooth_tick(s.gc_cur, s.gc_tgt, rc.smooth10);
  } else {
    float tg = clampf(1.0f / sqrtf((float)active), 0.0f, 1.0f);
    if (fabsf(s.gc_tgt - tg) > 1e-8f) s.gc_tgt = tg;
    smooth_tick(s.gc_cur, s.gc_tgt, rc.smooth10);
    const float gc = s.gc_cur;
    for (int i = 0; i < 80; i++) {
      Grain& g = s.grains[i];
      if (!g.active) continue;
      if (g.age >= g.duration) { g.active = 0; continue; }
      float phase = clampf(g.age / g.duration, 0.0f, 1.0f);
      float window = gm::g_powf(fmaxf(gm::g_sinf(PI_F * clampf(phase, 0.0f, 1.0f)), 0.0f), g.window_shape);
      float rg = g.release_total > 0.0f ? clampf(g.release_samples / g.release_total, 0.0f, 1.0f) : 1.0f;
      float smp = gran_sample(buf, s.buf_len, g.source_pos);
      raw += smp * window * rg * g.velocity * gc;
      g.source_pos += g.speed * g.direction;
      g.age += 1.0f;
      if (g.release_samples > 0.0f) { g.release_samples -= 1.0f; if (g.release_samples <= 0.0f) g.active = 0; }
    }
  }
  s.drive.mix = clampf(s.cur[G_DRIVE], 0.0f, 1.0f);
  float driven = ws_process(s.drive, raw);
  return driven * s.cur[G_VOLUME];
}

}  // namespace gd
