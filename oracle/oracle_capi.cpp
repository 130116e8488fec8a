// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE ONLY.
// C entry points over the CPU restatement so tests/ and bench.py's cpu_baseline /
// --impl reference legs can drive it through ctypes.  Never linked into or called
// by the product library.
#include <thread>
#include <atomic>
#include <memory>
#include "synths.hpp"
#include "../include/gooey_batch.h"

using namespace orc;

static std::unique_ptr<Instrument> make_voice(const GooeyVoicePatch& p, float sr) {
  switch (p.instrument) {
    case GOOEY_INSTRUMENT_KICK: {
      KickConfig c;
      for (int i = 0; i < 18; i++) c.v[i] = clampf(p.params[i], 0.0f, 1.0f);
      auto k = std::make_unique<KickDrum>(sr, c);
      if (p.aux & 0x100) k->p[K_TUNING].set_immediate(p.params[23]);
      return k;
    }
    case GOOEY_INSTRUMENT_SNARE: {
      const float* a = p.params;
      float ft = a[13];
      int t = !(ft == ft) ? 0 : (ft <= 0.0f ? 0 : (ft >= 255.0f ? 255 : (int)ft));
      SnareConfig c = SnareConfig::full(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], (uint8_t)t,
                                        a[14], a[15], a[16], a[17], a[18]);
      auto s = std::make_unique<SnareDrum>(sr, c);
      if (p.aux & 0x100) s->p[S_TUNING].set_immediate(p.params[23]);
      return s;
    }
    case GOOEY_INSTRUMENT_HIHAT: {
      HiHat2Config c = HiHat2Config::make(p.params[0], p.params[1], p.params[2], (p.aux & 1) != 0, (p.aux & 2) == 0, p.params[3]);
      c.volume = clampf(p.params[4], 0.0f, 1.0f);
      auto h = std::make_unique<HiHat2>(sr, c);
      if (p.aux & 0x100) h->p[H_TUNING].set_immediate(p.params[23]);
      return h;
    }
    case GOOEY_INSTRUMENT_TOM: {
      auto t = std::make_unique<Tom2>(sr);
      if (p.aux & 1) t->set_config({p.params[0], p.params[1], p.params[2], p.params[3], p.params[4], p.params[5], p.params[6], p.params[7]});
      if (p.aux & 0x100) t->tuning = clampf(p.params[23], 0.0f, 1.0f);
      return t;
    }
    case GOOEY_INSTRUMENT_BASS: {
      BassConfig c;
      for (int i = 0; i < 15; i++) c.v[i] = clampf(p.params[i], 0.0f, 1.0f);
      auto b = std::make_unique<BassSynth>(sr, c);
      if (p.aux & 0x100) b->p[B_TUNING].set_immediate(p.params[23]);
      return b;
    }
    default: return nullptr;
  }
}

extern "C" {

// Events: kind 0 = trigger(value = velocity), 1 = set_param(param, value), 2 = set_param + snap_params.
// Events of one voice must be sorted by frame.  out[v * frames + i].  Returns 0, or -1 on a bad patch.
int orc_render_voices(const GooeyVoicePatch* patches, uint32_t n, float sr, uint32_t frames, uint32_t n_events,
                      const uint32_t* ev_voice, const uint32_t* ev_frame, const uint32_t* ev_kind, const uint32_t* ev_param,
                      const float* ev_value, float* out, int n_threads) {
  std::vector<std::vector<uint32_t>> per_voice(n);
  for (uint32_t e = 0; e < n_events; e++) if (ev_voice[e] < n) per_voice[ev_voice[e]].push_back(e);
  std::atomic<uint32_t> next{0};
  std::atomic<int> bad{0};
  auto work = [&]() {
    for (;;) {
      uint32_t v = next.fetch_add(1);
      if (v >= n) break;
      auto inst = make_voice(patches[v], sr);
      if (!inst) { bad = 1; continue; }
      double t = 0.0;
      const double dt = 1.0 / (double)sr;
      size_t cur = 0;
      const auto& evs = per_voice[v];
      float* o = out + (size_t)v * frames;
      for (uint32_t i = 0; i < frames; i++) {
        while (cur < evs.size() && ev_frame[evs[cur]] <= i) {
          uint32_t e = evs[cur++];
          if (ev_kind[e] == 0) inst->trigger_with_velocity(t, ev_value[e]);
          else { inst->set_param(ev_param[e], ev_value[e]); if (ev_kind[e] == 2) inst->snap_params(); }
        }
        o[i] = inst->tick(t);
        t += dt;
      }
    }
  };
  if (n_threads <= 1) work();
  else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; i++) th.emplace_back(work);
    for (auto& x : th) x.join();
  }
  return bad ? -1 : 0;
}

// ---- pins used by tests/test_oracle_pins.py ------------------------------------------------------
uint64_t orc_siphash(uint64_t m, uint64_t k0, uint64_t k1, int c, int d) { return siphash_u64(m, k0, k1, c, d); }
float orc_hash_noise(uint64_t idx) { return hash_noise(idx); }
float orc_max_curve(float p, float c) { return max_curve(p, c); }
void orc_halfband_coefs(float* out8) { const float* c = halfband_coefs(); for (int i = 0; i < 8; i++) out8[i] = c[i]; }
float orc_smoother_coeff(float sr, float ms) { return SmoothedParam::calculate_coeff(sr, ms); }
void orc_click_table(float* out64) { for (int i = 0; i < 64; i++) out64[i] = TOM_IMPULSE[i]; }
// Impulse responses of the half-band pair, in the layout rust/examples/dump_golden.rs writes (tests/golden/ref):
// up[2n], up[2n+1] = Upsampler8::process(impulse[n]); down[0..n) = Downsampler8 fed ((1,0),(0,0)...), down[n..2n) a fresh one fed ((0,1),(0,0)...)
void orc_halfband_impulse(float* up, float* down, uint32_t n) {
  Upsampler8 u;
  for (uint32_t i = 0; i < n; i++) u.process(i == 0 ? 1.0f : 0.0f, up[2 * i], up[2 * i + 1]);
  Downsampler8 d0, d1;
  for (uint32_t i = 0; i < n; i++) down[i] = d0.process(i == 0 ? 1.0f : 0.0f, 0.0f);
  for (uint32_t i = 0; i < n; i++) down[n + i] = d1.process(0.0f, i == 0 ? 1.0f : 0.0f);
}
// Oversampler2x then Oversampler4x around tanh(drive x) over the same n samples: out[0..n) and out[n..2n)
void orc_oversampler_tanh(const float* in, float* out, uint32_t n, float drive) {
  Oversampler o2, o4;
  o2.set_mode(OversamplingMode::X2); o4.set_mode(OversamplingMode::X4);
  for (uint32_t i = 0; i < n; i++) out[i] = o2.process(in[i], [&](float x) { return tanhf(x * drive); });
  for (uint32_t i = 0; i < n; i++) out[n + i] = o4.process(in[i], [&](float x) { return tanhf(x * drive); });
}
// Oversampler: mode 0/2/4, f(x) = tanh(x*drive) (drive<=0: identity); processes n samples.
void orc_oversample(int mode, float drive, const float* in, float* out, uint32_t n) {
  Oversampler os;
  os.set_mode(mode == 0 ? OversamplingMode::Off : (mode == 2 ? OversamplingMode::X2 : OversamplingMode::X4));
  for (uint32_t i = 0; i < n; i++) out[i] = os.process(in[i], [&](float x) { return drive > 0.0f ? tanhf(x * drive) : x; });
}
void orc_pink(float sr, float* out, uint32_t n) { PinkNoise p(sr); for (uint32_t i = 0; i < n; i++) out[i] = p.tick(); }

}  // extern "C"

// =====================================================================================================
// Engine-level mirror of the reference C FFI subset (include/gooey.h) with an orc_ prefix, so a test can
// run one call script against both the oracle and the product.
// =====================================================================================================
#include "engine.hpp"

extern "C" {

void* orc_engine_new(float sr) { return new FfiEngine(sr); }
void orc_engine_free(void* e) { delete (FfiEngine*)e; }
#define E ((FfiEngine*)e)
static void set_typed(void* e, uint32_t type, uint32_t param, float v) { if (!e) return; if (VoiceStrip* s = E->by_type(type)) s->inst->set_param(param, v); }
void orc_engine_set_kick_param(void* e, uint32_t p, float v) { set_typed(e, 0, p, v); }
void orc_engine_set_snare_param(void* e, uint32_t p, float v) { set_typed(e, 1, p, v); }
void orc_engine_set_hihat_param(void* e, uint32_t p, float v) { set_typed(e, 2, p, v); }
void orc_engine_set_tom_param(void* e, uint32_t p, float v) { set_typed(e, 3, p, v); }
void orc_engine_set_bass_param(void* e, uint32_t p, float v) { set_typed(e, 4, p, v); }
void orc_engine_set_channel_param(void* e, uint32_t ch, uint32_t p, float v) { if (e && ch < 5) E->voices[ch].inst->set_param(p, v); }
void orc_engine_set_channel_instrument_type(void* e, uint32_t ch, uint32_t type) {
  if (!e || ch >= 5 || type > 4 || E->voices[ch].type == type) return;
  VoiceStrip& v = E->voices[ch];
  v.inst = make_instrument(type, E->sample_rate);
  v.type = type;
  v.blender.default_for_type(type);                                    // ffi.rs:2334-2342
  if (v.blend_enabled) v.blend_and_apply(v.blend_x, v.blend_y);
}
// ---- LFO pool (ffi.rs:4616-4993) ----
void orc_engine_set_lfo_enabled(void* e, uint32_t i, bool on) { if (e && i < 8) E->lfo_enabled[i] = on; }
void orc_engine_set_lfo_timing(void* e, uint32_t i, uint32_t timing) { if (e && i < 8 && timing < 8) { E->lfos[i].synced = true; E->lfos[i].division = timing; } }
void orc_engine_set_lfo_frequency(void* e, uint32_t i, float hz) { if (e && i < 8) { E->lfos[i].synced = false; E->lfos[i].hz = hz; } }
void orc_engine_set_lfo_amount(void* e, uint32_t i, float a) { if (e && i < 8) E->lfos[i].amount = a; }
void orc_engine_set_lfo_offset(void* e, uint32_t i, float o) { if (e && i < 8) E->lfos[i].offset = o; }
uint32_t orc_engine_add_lfo_route(void* e, uint32_t i, uint32_t inst, uint32_t param, float depth) {
  if (!e || i >= 8 || E->lfo_routes[i].size() >= 16) return 0xFFFFFFFFu;
  const uint32_t id = E->lfo_next_route_id[i]++;
  E->lfo_routes[i].push_back({id, inst, param, depth});
  return id;
}
bool orc_engine_remove_lfo_route(void* e, uint32_t i, uint32_t id) {
  if (!e || i >= 8) return false;
  auto& r = E->lfo_routes[i];
  for (size_t k = 0; k < r.size(); k++) if (r[k].id == id) { r.erase(r.begin() + k); return true; }
  return false;
}
void orc_engine_clear_lfo_routes(void* e, uint32_t i) { if (e && i < 8) E->lfo_routes[i].clear(); }
void orc_engine_reset_lfo_phase(void* e, uint32_t i) { if (e && i < 8) E->lfos[i].phase = 0.0f; }
float orc_engine_get_lfo_phase(void* e, uint32_t i) { return (e && i < 8) ? E->lfos[i].phase : -1.0f; }
// ---- preset blend (ffi.rs:5245-5490) and per-step blend (:4009-4075) ----
void orc_engine_blend_enable(void* e, uint32_t i) { if (e && i < 5) E->voices[i].blend_enabled = true; }
void orc_engine_blend_disable(void* e, uint32_t i) { if (e && i < 5) E->voices[i].blend_enabled = false; }
bool orc_engine_blend_is_enabled(void* e, uint32_t i) { return e && i < 5 && E->voices[i].blend_enabled; }
void orc_engine_blend_set_position(void* e, uint32_t i, float x, float y) {
  if (!e || i >= 5) return;
  VoiceStrip& v = E->voices[i];
  if (!v.blend_enabled) return;
  v.blend_x = clampf(x, 0.0f, 1.0f); v.blend_y = clampf(y, 0.0f, 1.0f);
  v.blend_and_apply(v.blend_x, v.blend_y);
}
float orc_engine_blend_get_position_x(void* e, uint32_t i) { return (e && i < 5) ? E->voices[i].blend_x : -1.0f; }
float orc_engine_blend_get_position_y(void* e, uint32_t i) { return (e && i < 5) ? E->voices[i].blend_y : -1.0f; }
void orc_engine_blend_set_corner_preset(void* e, uint32_t i, uint32_t corner, uint32_t id) {
  if (!e || i >= 5 || corner >= 4) return;
  E->voices[i].blender.corner_ids[corner] = id;
  E->voices[i].blender.set_corner_preset(corner, id);
}
uint32_t orc_engine_blend_get_corner_preset(void* e, uint32_t i, uint32_t corner) { return (e && i < 5 && corner < 4) ? E->voices[i].blender.corner_ids[corner] : 0xFFFFFFFFu; }
void orc_engine_blend_reset_corners(void* e, uint32_t i) { if (e && i < 5) E->voices[i].blender.default_for_type(E->voices[i].type); }
void orc_engine_sequencer_set_instrument_step_blend(void* e, uint32_t inst, uint32_t step, float x, float y) {
  if (!e || inst >= 5 || step >= E->voices[inst].seq.pattern.size()) return;
  SeqStep& st = E->voices[inst].seq.pattern[step];
  st.has_blend = true; st.bx = clampf(x, 0.0f, 1.0f); st.by = clampf(y, 0.0f, 1.0f);
}
void orc_engine_sequencer_clear_instrument_step_blend(void* e, uint32_t inst, uint32_t step) {
  if (!e || inst >= 5 || step >= E->voices[inst].seq.pattern.size()) return;
  E->voices[inst].seq.pattern[step].has_blend = false;
}
void orc_engine_load_bass_preset(void* e, uint32_t id) {
  if (!e || id > 3) return;
  VoiceStrip* s = E->by_type(4);
  if (!s) return;
  BassConfig c = id == 0 ? BassConfig::acid() : id == 1 ? BassConfig::sub() : id == 2 ? BassConfig::reese() : BassConfig::stab();
  static_cast<BassSynth*>(s->inst.get())->set_config(c);
}
void orc_engine_set_bpm(void* e, float b) { if (e) E->set_bpm(b); }
void orc_engine_set_swing(void* e, float s) { if (e) E->set_swing(s); }
void orc_engine_set_master_gain(void* e, float g) { if (e && std::isfinite(g)) E->master_gain.set_target(g); }
void orc_engine_sequencer_set_instrument_step_settings(void* e, uint32_t inst, uint32_t step, bool enabled, bool set_vel, float vel, bool set_blend,
                                                       float bx, float by, bool set_note, uint8_t note) {
  if (!e || inst >= 5) return;
  VoiceStrip* s = &E->voices[inst];  // sequencer_for_instrument = channel index (ffi.rs:3767-3770)
  if (step >= s->seq.pattern.size()) return;
  SeqStep& st = s->seq.pattern[step];
  st.enabled = enabled;
  if (set_vel) st.velocity = clampf(vel, 0.0f, 1.0f);
  if (set_blend) { st.has_blend = true; st.bx = clampf(bx, 0, 1); st.by = clampf(by, 0, 1); }
  if (set_note) { if (note == 255) st.has_note = false; else { st.has_note = true; st.note = note; } }
}
void orc_engine_sequencer_set_instrument_step(void* e, uint32_t inst, uint32_t step, bool enabled) {
  if (!e || inst >= 5) return;
  VoiceStrip* s = &E->voices[inst];
  if (step < s->seq.pattern.size()) s->seq.pattern[step].enabled = enabled;
}
void orc_engine_sequencer_set_step(void* e, uint32_t step, bool enabled) { orc_engine_sequencer_set_instrument_step(e, 0, step, enabled); }   // kick only (ffi.rs:3695-3708)
void orc_engine_sequencer_set_instrument_step_with_velocity(void* e, uint32_t inst, uint32_t step, bool enabled, float vel) {  // sequencer.rs:705-712: keeps blend and note
  if (!e || inst >= 5) return;
  VoiceStrip* s = &E->voices[inst];
  if (step < s->seq.pattern.size()) { s->seq.pattern[step].enabled = enabled; s->seq.pattern[step].velocity = clampf(vel, 0.0f, 1.0f); }
}
void orc_engine_sequencer_set_instrument_step_note(void* e, uint32_t inst, uint32_t step, uint8_t note) {  // ffi.rs:3975-3995
  if (!e || inst >= 5) return;
  VoiceStrip* s = &E->voices[inst];
  if (step < s->seq.pattern.size()) { if (note == 255) s->seq.pattern[step].has_note = false; else { s->seq.pattern[step].has_note = true; s->seq.pattern[step].note = note; } }
}
void orc_engine_sequencer_set_instrument_pattern(void* e, uint32_t inst, const bool* pattern) {  // sequencer.rs:822-828
  if (!e || !pattern || inst >= 5) return;
  Sequencer& q = E->voices[inst].seq;
  q.pattern.assign(16, SeqStep());
  for (int i = 0; i < 16; i++) { q.pattern[i].enabled = pattern[i]; q.pattern[i].velocity = 1.0f; }
  if (q.current_step >= q.pattern.size()) q.current_step = 0;
}
void orc_engine_sequencer_start(void* e) { if (e) E->sequencer_start(); }
void orc_engine_sequencer_stop(void* e) { if (e) E->sequencer_stop(); }
void orc_engine_sequencer_reset(void* e) { if (e) E->sequencer_reset(); }
void orc_engine_set_sequencer_triggers_enabled(void* e, bool on) { if (e) E->seq_triggers_enabled = on; }   // ffi.rs:2188-2198
bool orc_engine_get_sequencer_triggers_enabled(void* e) { return e ? E->seq_triggers_enabled : true; }    // ffi.rs:2205-2215
void orc_engine_set_global_effect_param(void* e, uint32_t fx, uint32_t p, float v) {
  if (!e) return;
  switch (fx) { case 1: E->delay.set_param(p, v); break; case 4: E->tilt.set_param(p, v); break; case 6: E->reverb.set_param(p, v); break;
    case 9: E->plate.set_param(p, v); break; case 5: if (p == 0) E->limiter.set_threshold(v); break;
    case 0: E->lowpass.set_param(p, v); break; case 2: E->saturation.set_param(p, v); break; case 3: E->compressor.set_param(p, v); break;
    case 7: if (p == 0) E->waveshaper.set_drive(v); else if (p == 1) E->waveshaper.set_mix(v); break;
    case 8: {
      if (p == 0) E->feedback_waveshaper.set_drive(v); else if (p == 1) E->feedback_waveshaper.set_feedback(v);
      else if (p == 2) E->feedback_waveshaper.set_filter_cutoff(v); else if (p == 3) E->feedback_waveshaper.set_mix(v);
    } break; }
}
void orc_engine_set_global_effect_enabled(void* e, uint32_t fx, bool on) {
  if (!e) return;
  switch (fx) { case 1: E->delay_enabled = on; break; case 4: E->tilt_enabled = on; break; case 6: E->reverb_enabled = on; break;
    case 9: E->plate_enabled = on; break; case 5: E->limiter_enabled = on; break;
    case 0: E->lowpass_enabled = on; break; case 2: E->saturation_enabled = on; break; case 3: E->compressor_enabled = on; break;
    case 7: E->waveshaper_enabled = on; break; case 8: E->feedback_waveshaper_enabled = on; break; }
}
void orc_engine_set_compressor_sidechain(void* e, uint32_t inst) { if (e) E->compressor_sidechain = inst; }
bool orc_engine_set_effect_order(void* e, const uint32_t* ids, uint32_t len) {
  if (!e || !ids || len != 9) return false;
  for (uint32_t i = 0; i < 9; i++) { if (ids[i] > 9 || ids[i] == 5) return false; for (uint32_t j = 0; j < i; j++) if (ids[j] == ids[i]) return false; }
  for (uint32_t i = 0; i < 9; i++) E->effect_order[i] = ids[i];
  E->reset_effect_states();
  return true;
}
void orc_engine_set_instrument_gain(void* e, uint32_t i, float g) { if (e && i < 5) E->voices[i].channel_gain.set_target(clampf(g, 0, 1)); }
void orc_engine_set_instrument_pan(void* e, uint32_t i, float p) { if (e && i < 5) E->voices[i].pan.set_target(clampf(p, 0, 1)); }
void orc_engine_set_instrument_mute(void* e, uint32_t i, bool m) { if (e && i < 5) E->voices[i].muted = m; }
void orc_engine_set_instrument_solo(void* e, uint32_t i, bool s) { if (e && i < 5) E->voices[i].soloed = s; }
void orc_engine_trigger_instrument_with_velocity(void* e, uint32_t i, float v) { if (e && i < 5) { E->voices[i].trigger_velocity = clampf(v, 0, 1); E->voices[i].trigger_pending = true; } }
void orc_engine_trigger_instrument(void* e, uint32_t i) { orc_engine_trigger_instrument_with_velocity(e, i, 1.0f); }
bool orc_engine_move_effect(void* e, uint32_t id, uint32_t new_position) {  // ffi.rs:4544-4581
  if (!e || id > 9 || id == 5 || new_position >= 9) return false;
  int cur = -1;
  for (int i = 0; i < 9; i++) if (E->effect_order[i] == id) cur = i;
  if (cur < 0) return false;
  int np = (int)new_position;
  if (cur == np) return true;
  if (np > cur) for (int i = cur; i < np; i++) E->effect_order[i] = E->effect_order[i + 1];
  else for (int i = cur - 1; i >= np; i--) E->effect_order[i + 1] = E->effect_order[i];
  E->effect_order[np] = id;
  E->reset_effect_states();
  return true;
}
bool orc_engine_track_effect_remove(void* e, uint32_t t, uint32_t slot) {
  if (!e || t >= E->graph.tracks.size() || slot >= E->graph.tracks[t].rack.size()) return false;
  E->graph.tracks[t].rack.erase(E->graph.tracks[t].rack.begin() + slot);
  return true;
}
bool orc_engine_track_effect_move(void* e, uint32_t t, uint32_t slot, uint32_t np) {
  if (!e || t >= E->graph.tracks.size() || slot >= E->graph.tracks[t].rack.size()) return false;
  auto& r = E->graph.tracks[t].rack;
  auto fx = std::move(r[slot]);
  r.erase(r.begin() + slot);
  size_t dest = np < r.size() ? np : r.size();
  r.insert(r.begin() + dest, std::move(fx));
  return true;
}
int32_t orc_engine_mixer_add_track(void* e, const char*) { return e ? (int32_t)E->graph.add_track() : -1; }
bool orc_engine_mixer_route_source(void* e, uint32_t src, uint32_t track) { return e ? E->graph.route(src, track) : false; }
// graph layout and strip getters (ffi.rs:6291-6320, 6427-6455, 6472-6566)
bool orc_engine_mixer_unroute_source(void* e, uint32_t src) { return e ? E->graph.unroute(src) : false; }
int32_t orc_engine_mixer_get_source_route(void* e, uint32_t src) { return e ? (int32_t)E->graph.route_of(src) : -1; }
void orc_engine_mixer_clear_layout(void* e) { if (e) E->graph.reset(); }
void orc_engine_mixer_reset_default_layout(void* e) {
  if (!e) return;
  E->graph = MixerGraph(E->sample_rate, E->bpm);
  E->graph.default_layout();
  for (uint32_t r = 0; r < 4; r++) if (E->samplers[r]) E->graph.register_source(5 + r);
}
float orc_engine_mixer_get_track_gain(void* e, uint32_t t) { return (e && t < E->graph.tracks.size()) ? E->graph.tracks[t].gain.target : 1.0f; }
float orc_engine_mixer_get_track_pan(void* e, uint32_t t) { return (e && t < E->graph.tracks.size()) ? E->graph.tracks[t].pan.target : 0.5f; }
bool orc_engine_mixer_get_track_mute(void* e, uint32_t t) { return e && t < E->graph.tracks.size() && E->graph.tracks[t].muted; }
bool orc_engine_mixer_get_track_solo(void* e, uint32_t t) { return e && t < E->graph.tracks.size() && E->graph.tracks[t].soloed; }
void orc_engine_mixer_set_track_gain(void* e, uint32_t t, float g) { if (e && t < E->graph.tracks.size()) E->graph.tracks[t].gain.set_target(clampf(g, 0.0f, 2.0f)); }
void orc_engine_mixer_set_track_pan(void* e, uint32_t t, float p) { if (e && t < E->graph.tracks.size()) E->graph.tracks[t].pan.set_target(clampf(p, 0.0f, 1.0f)); }
void orc_engine_mixer_set_track_mute(void* e, uint32_t t, bool m) { if (e && t < E->graph.tracks.size()) E->graph.tracks[t].muted = m; }
void orc_engine_mixer_set_track_solo(void* e, uint32_t t, bool s) { if (e && t < E->graph.tracks.size()) E->graph.tracks[t].soloed = s; }
int32_t orc_engine_track_effect_add(void* e, uint32_t t, uint32_t fx) {
  if (!e || t >= E->graph.tracks.size()) return -1;
  auto eff = make_channel_effect(fx, E->sample_rate, E->graph.bpm);
  if (!eff) return -1;
  E->graph.tracks[t].rack.push_back(std::move(eff));
  return (int32_t)E->graph.tracks[t].rack.size() - 1;
}
void orc_engine_track_effect_set_param(void* e, uint32_t t, uint32_t slot, uint32_t p, float v) {
  if (e && t < E->graph.tracks.size() && slot < E->graph.tracks[t].rack.size()) E->graph.tracks[t].rack[slot]->set_param(p, v);
}
bool orc_engine_granulator_set_buffer(void* e, const float* s, uint32_t len, float sr) {
  if (!e || !s || len == 0 || !std::isfinite(sr) || sr <= 0.0f) return false;
  for (uint32_t i = 0; i < len; i++) if (!std::isfinite(s[i])) return false;
  E->granulator.set_buffer(std::make_shared<std::vector<float>>(s, s + len), sr);
  return true;
}
void orc_engine_granulator_trigger(void* e, float v) { if (e) E->granulator.trigger_with_velocity(E->current_time, clampf(v, 0, 1)); }
void orc_engine_granulator_set_param(void* e, uint32_t p, float v) { if (e) E->granulator.set_param(p, v); }
void orc_engine_granulator_set_seed(void* e, uint32_t s) { if (e) E->granulator.set_seed(s); }
void orc_engine_granulator_snap_params(void* e) { if (e) E->granulator.snap_params(); }
void orc_engine_poly_trigger_notes(void* e, const uint8_t* notes, uint32_t n, uint32_t preset, float vel) {
  if (!e) return;
  vel = clampf(vel, 0, 1);
  E->poly.set_config(PolyConfig::preset(preset));
  E->poly.release_all();
  for (uint32_t i = 0; i < n; i++) E->poly.trigger_note(notes[i], vel);
}
void orc_engine_poly_release(void* e) { if (e) E->poly.release_all(); }
void orc_engine_poly_set_preset(void* e, uint32_t p) { if (e) E->poly.set_config(PolyConfig::preset(p)); }
void orc_engine_poly_set_param(void* e, uint32_t p, float v) { if (e) E->poly.set_param(p, v); }
void orc_engine_render(void* e, float* buf, uint32_t frames) { if (e && buf) E->render(buf, frames); }
// ffi.rs:2572-2584, 6573-6580, 2145-2167
void orc_engine_get_channel_peaks(void* e, float* out, uint32_t count) {
  if (!e || !out) return;
  for (uint32_t i = 0; i < count && i < 5; i++) { out[i] = E->voices[i].peak; E->voices[i].peak = 0.0f; }
}
float orc_engine_mixer_get_track_peak(void* e, uint32_t t) {
  if (!e || t >= E->graph.tracks.size()) return 0.0f;
  float p = E->graph.tracks[t].peak; E->graph.tracks[t].peak = 0.0f; return p;
}
uint32_t orc_engine_drain_midi_events(void* e, void* out, uint32_t max_events) {
  if (!e || !out || max_events == 0) return 0;
  auto& q = E->pending_midi_events;
  const size_t n = q.size() < max_events ? q.size() : max_events;
  if (n) { memcpy(out, q.data(), n * sizeof(FfiEngine::MidiEvent)); q.erase(q.begin(), q.begin() + n); }
  return (uint32_t)n;
}
float* orc_engine_bounce_to_buffer(void* e, uint32_t bars, uint32_t* out_len) {
  if (!e || !out_len) return nullptr;
  std::vector<float> v = E->bounce_to_buffer(bars);
  float* p = (float*)malloc(v.size() * sizeof(float) + 4);
  memcpy(p, v.data(), v.size() * sizeof(float));
  *out_len = (uint32_t)v.size();
  return p;
}
void orc_engine_free_buffer(float* p, uint32_t) { free(p); }
// sequencer trigger table of channel `ch` over `frames` samples after reset+start (bit-exact gate)
uint32_t orc_engine_trigger_table(void* e, uint32_t ch, uint32_t frames, uint32_t* out_frames, float* out_vel, uint32_t cap) {
  if (!e || ch >= 5) return 0;
  Sequencer s = E->voices[ch].seq;
  s.reset(); s.start();
  uint32_t n = 0;
  for (uint32_t f = 0; f < frames; f++) { SeqTrigger t; if (s.tick(t)) { if (n < cap) { out_frames[n] = f; out_vel[n] = t.velocity; } n++; } }
  return n;
}
// ---- loop mixer (ffi.rs:7150-7535, 8006-8048) and sampler racks (ffi.rs:6000-6172) ----------------------------------
bool orc_engine_loop_load(void* e, uint32_t ch, const float* samples, uint32_t frames, uint32_t channels, float sr) {
  if (!e || !samples || frames == 0 || channels == 0) return false;
  auto b = StereoSampleBuffer::from_interleaved(samples, (size_t)frames * channels, channels, sr);
  if (!b) return false;
  LoopChannel* c = E->mixer.ch(ch);
  if (!c) return false;
  c->set_buffer(b);
  return true;
}
#define LCH(body) do { if (e) { if (LoopChannel* c = E->mixer.ch(ch)) { body; } } } while (0)
void orc_engine_loop_set_playing(void* e, uint32_t ch, bool v) { LCH(c->playing = v); }
void orc_engine_loop_set_gain(void* e, uint32_t ch, float v) { LCH(c->set_gain(v)); }
void orc_engine_loop_set_mute(void* e, uint32_t ch, bool v) { LCH(c->muted = v); }
void orc_engine_loop_set_solo(void* e, uint32_t ch, bool v) { LCH(c->soloed = v); }
void orc_engine_loop_set_start(void* e, uint32_t ch, float v) { LCH(c->set_loop_start(v)); }
void orc_engine_loop_set_end(void* e, uint32_t ch, float v) { LCH(c->set_loop_end(v)); }
void orc_engine_loop_set_speed(void* e, uint32_t ch, float v) { LCH(c->set_speed(v)); }
void orc_engine_loop_set_source_bpm(void* e, uint32_t ch, float v) { LCH(if (c->buffer) c->buffer->set_source_bpm(v > 0.0f, v)); }
void orc_engine_loop_set_pitch_mode(void* e, uint32_t ch, uint32_t m) { LCH(c->set_pitch_mode(m == 1 ? PITCH_RESAMPLE : (m == 2 ? PITCH_PRESERVE : PITCH_OFF))); }
void orc_engine_loop_restart(void* e, uint32_t ch) { LCH(c->restart()); }
void orc_engine_loop_set_position(void* e, uint32_t ch, float v) { LCH(c->set_position(v)); }
void orc_engine_loop_cancel_queued_swap(void* e, uint32_t ch) { LCH(c->cancel_queued_swap()); }
#undef LCH
bool orc_engine_loop_queue_swap(void* e, uint32_t ch, const float* samples, uint32_t frames, uint32_t channels, float sr, float source_bpm, uint32_t divisions) {  // ffi.rs:7449-7476
  if (!e || !samples || frames == 0 || channels == 0) return false;
  auto b = StereoSampleBuffer::from_interleaved(samples, (size_t)frames * channels, channels, sr);
  if (!b) return false;
  b->set_source_bpm(source_bpm > 0.0f, source_bpm);
  LoopChannel* c = E->mixer.ch(ch);
  if (!c) return false;
  c->queue_swap(b, divisions);
  return true;
}
uint32_t orc_engine_loop_swaps_completed(void* e, uint32_t ch) { if (!e) return 0; LoopChannel* c = E->mixer.ch(ch); return c ? c->swaps_completed : 0u; }
float orc_engine_loop_get_source_bpm(void* e, uint32_t ch) { if (!e) return 0.0f; LoopChannel* c = E->mixer.ch(ch); return (c && c->buffer && c->buffer->has_source_bpm) ? c->buffer->source_bpm : 0.0f; }
uint32_t orc_engine_loop_get_pitch_mode(void* e, uint32_t ch) { if (!e) return 0; LoopChannel* c = E->mixer.ch(ch); return c ? (uint32_t)c->pitch_mode : 0u; }
float orc_engine_loop_get_position(void* e, uint32_t ch) { if (!e) return 0.0f; LoopChannel* c = E->mixer.ch(ch); return c ? c->position_normalized() : 0.0f; }
// render_channel_to_interleaved (mixer/mod.rs:444-476), the audio gooey_engine_loop_render_to_wav writes as 32-bit float stereo
bool orc_engine_loop_render(void* e, uint32_t ch, uint32_t frames, uint32_t preroll, float* out) {
  if (!e || !out || frames == 0) return false;
  std::vector<float> v;
  if (!E->mixer.render_channel_to_interleaved(ch, frames, preroll, v)) return false;
  memcpy(out, v.data(), v.size() * sizeof(float));
  return true;
}
int32_t orc_engine_sampler_register(void* e) {
  if (!e) return -1;
  for (int i = 0; i < 4; i++)
    if (!E->samplers[i]) {   // SamplerRack::new(engine.sample_rate, engine.bpm, ..): its sequencer starts at the engine's CURRENT tempo, swing 0.5
      E->samplers[i] = std::make_unique<SamplerRack>(E->sample_rate);
      E->rack_pat[i] = FfiEngine::RackPattern(E->bpm, E->sample_rate);
      E->graph.register_source(5 + i);
      return i;
    }
  return -1;
}
uint32_t orc_engine_sampler_get_source_id(void* e, uint32_t rack) { return (e && rack < 4 && E->samplers[rack]) ? 5u + rack : 0xFFFFFFFFu; }
bool orc_engine_sampler_set_slot_buffer(void* e, uint32_t rack, uint32_t slot, const float* s, uint32_t frames, uint32_t channels, float sr) {
  if (!e || !s) return false;
  auto b = SamplerBuffer::from_interleaved(s, frames, channels, sr);
  if (!b || rack >= 4 || !E->samplers[rack]) return false;
  return E->samplers[rack]->set_buffer(slot, b);
}
bool orc_engine_sampler_clear_slot(void* e, uint32_t rack, uint32_t slot) { return e && rack < 4 && E->samplers[rack] && E->samplers[rack]->clear_slot(slot); }
bool orc_engine_sampler_slot_is_loaded(void* e, uint32_t rack, uint32_t slot) { return e && rack < 4 && E->samplers[rack] && slot < 16 && E->samplers[rack]->slots[slot]; }
uint32_t orc_engine_sampler_slot_frames(void* e, uint32_t rack, uint32_t slot) { return orc_engine_sampler_slot_is_loaded(e, rack, slot) ? (uint32_t)E->samplers[rack]->slots[slot]->frames : 0u; }
uint32_t orc_engine_sampler_slot_channels(void* e, uint32_t rack, uint32_t slot) { return orc_engine_sampler_slot_is_loaded(e, rack, slot) ? (uint32_t)E->samplers[rack]->slots[slot]->channels : 0u; }
float orc_engine_sampler_slot_sample_rate(void* e, uint32_t rack, uint32_t slot) { return orc_engine_sampler_slot_is_loaded(e, rack, slot) ? E->samplers[rack]->slots[slot]->sample_rate : 0.0f; }
// the rack's step pattern (ffi.rs:6173-6290; sampler.rs:232-310)
bool orc_engine_sampler_set_step(void* e, uint32_t rack, uint32_t step, bool enabled, uint32_t slot, float vel) {
  if (!e || rack >= 4 || !E->samplers[rack] || step >= 16 || slot >= 16) return false;
  SeqStep& s = E->rack_pat[rack].seq.pattern[step];
  s.enabled = enabled; s.velocity = clampf(vel, 0.0f, 1.0f); s.has_note = true; s.note = (uint8_t)slot;
  return true;
}
bool orc_engine_sampler_get_step(void* e, uint32_t rack, uint32_t step, bool* en, uint32_t* slot, float* vel) {
  if (!e || !en || !slot || !vel || rack >= 4 || !E->samplers[rack] || step >= 16) return false;
  const SeqStep& s = E->rack_pat[rack].seq.pattern[step];
  *en = s.enabled; *slot = s.has_note ? s.note : 0u; *vel = s.velocity;
  return true;
}
bool orc_engine_sampler_start_pattern(void* e, uint32_t rack, uint32_t quantization) {
  if (!e || quantization > 2) return false;
  const double interval = quantization == 0 ? 0.25 : (quantization == 1 ? 1.0 : 4.0);
  const double target = E->quantized_target(interval);
  if (rack >= 4 || !E->samplers[rack]) return false;
  auto& p = E->rack_pat[rack];                      // SamplerRack::schedule_start (:254-262)
  if (!std::isfinite(target) || target < 0.0) return false;
  p.pattern_running = false; p.seq.stop(); p.has_pending = true; p.pending_start_beat = target;
  return true;
}
bool orc_engine_sampler_stop_pattern(void* e, uint32_t rack) {
  if (!e || rack >= 4 || !E->samplers[rack]) return false;
  auto& p = E->rack_pat[rack];
  p.has_pending = false; p.pattern_running = false; p.seq.stop(); E->rack_stop_all((int)rack);
  return true;
}
bool orc_engine_sampler_cancel_pattern_start(void* e, uint32_t rack) { if (!e || rack >= 4 || !E->samplers[rack]) return false; E->rack_pat[rack].has_pending = false; return true; }
double orc_engine_sampler_get_pending_start_beat(void* e, uint32_t rack) { return (e && rack < 4 && E->samplers[rack] && E->rack_pat[rack].has_pending) ? E->rack_pat[rack].pending_start_beat : -1.0; }
bool orc_engine_sampler_is_pattern_running(void* e, uint32_t rack) { return e && rack < 4 && E->samplers[rack] && E->rack_pat[rack].pattern_running; }
// test hooks: the pattern hits logged so far (frames counted from the engine's first rendered frame) and the transport beat
uint32_t orc_engine_sampler_hit_log(void* e, uint32_t rack, uint32_t* frames, uint32_t* slots, float* vel, uint32_t cap) {
  if (!e || rack >= 4) return 0;
  uint32_t n = 0;
  for (const auto& h : E->rack_hits[rack]) { if (n < cap) { frames[n] = (uint32_t)h.frame; slots[n] = h.slot; vel[n] = h.velocity; } n++; }
  return n;
}
double orc_engine_transport_beat(void* e) { return e ? E->transport_beat : 0.0; }
double orc_engine_transport_get_beat_position(void* e) { return e ? E->transport_beat : 0.0; }   // ffi.rs:7143-7150
void orc_engine_capture_rack0(void* e, bool on) { if (e) { E->capture_rack0 = on; E->rack0_capture.clear(); } }
uint32_t orc_engine_rack0_capture(void* e, float* out_interleaved, uint32_t cap_frames) {
  if (!e) return 0;
  uint32_t n = (uint32_t)(E->rack0_capture.size() / 2);
  if (out_interleaved) memcpy(out_interleaved, E->rack0_capture.data(), sizeof(float) * 2 * (n < cap_frames ? n : cap_frames));
  return n;
}
bool orc_engine_sampler_trigger(void* e, uint32_t rack, uint32_t slot, float vel) { return e && rack < 4 && E->samplers[rack] && E->samplers[rack]->trigger(slot, vel); }
// test hooks: the loop mixer / one sampler rack ticked on their own (what the product's ext_source_kernel computes per descriptor)
void orc_engine_loop_mixer_tick(void* e, uint32_t frames, float* out_interleaved) {
  if (!e || !out_interleaved) return;
  for (uint32_t f = 0; f < frames; f++) { StereoFrame s = E->mixer.tick(E->sample_rate); out_interleaved[2 * f] = s.l; out_interleaved[2 * f + 1] = s.r; }
}
void orc_engine_sampler_rack_tick(void* e, uint32_t rack, uint32_t frames, float* out_interleaved) {
  if (!e || !out_interleaved || rack >= 4 || !E->samplers[rack]) return;
  for (uint32_t f = 0; f < frames; f++) { StereoFrame s = E->samplers[rack]->tick(); out_interleaved[2 * f] = s.l; out_interleaved[2 * f + 1] = s.r; }
}
uint32_t orc_engine_sampler_active_voices(void* e, uint32_t rack) {
  if (!e || rack >= 4 || !E->samplers[rack]) return 0;
  uint32_t n = 0;
  for (auto& v : E->samplers[rack]->voices) n += v.active();
  return n;
}
double orc_engine_loop_get_cursor(void* e, uint32_t ch) { if (!e) return 0.0; LoopChannel* c = E->mixer.ch(ch); return c ? c->cursor : 0.0; }
#undef E

// ---- Rust-API engine (engine/mod.rs + bounce.rs): instruments by type with optional config, one sequencer each ----
void* orc_rust_engine_new(float sr) { return new RustEngine(sr); }
void orc_rust_engine_free(void* e) { delete (RustEngine*)e; }
#define R ((RustEngine*)e)
int orc_rust_engine_add_instrument(void* e, const char* name, const GooeyVoicePatch* patch) {
  auto inst = make_voice(*patch, R->sample_rate);
  if (!inst) return -1;
  R->instruments.emplace_back(std::string(name), std::move(inst));
  return 0;
}
void orc_rust_engine_add_sequencer(void* e, const char* name, const uint8_t* enabled, const float* velocity, uint32_t steps) {
  Sequencer s(R->bpm, R->sample_rate, steps, false);
  for (uint32_t i = 0; i < steps; i++) { s.pattern[i].enabled = enabled[i] != 0; s.pattern[i].velocity = velocity ? clampf(velocity[i], 0, 1) : 1.0f; }
  R->sequencers.emplace_back(std::move(s), std::string(name));
}
void orc_rust_engine_set_bpm(void* e, float b) { R->bpm = b; }
void orc_rust_engine_set_master_gain(void* e, float g) { R->master_gain.set_target(g); }
void orc_rust_engine_clear_global_effects(void* e) { R->limiter_on = false; }
void orc_rust_engine_bounce_samples(void* e, uint32_t n, float* out) { auto v = R->bounce_samples(n); memcpy(out, v.data(), n * sizeof(float)); }
#undef R

// ---- stand-alone effects (tests/test_reference_units_cpu.py restates the reference's inline unit tests of src/effects/*.rs on these) ----
// kind = FFI effect id (ffi.rs:1548-1575); ctor[] = the effect's constructor arguments after the sample rate, in the reference's order:
//   0 low-pass (cutoff, res) · 1 delay (timing, bpm, feedback, mix, cutoff) · 2 saturation (drive, warmth, mix) ·
//   3 compressor (threshold, ratio, attack, release, mix) · 4 tilt () · 6 spring (decay, mix, damping) · 7 waveshaper (drive, mix) ·
//   8 feedback waveshaper (drive, feedback, cutoff, mix) · 9 plate (decay, mix, damping)
struct FxBox { uint32_t kind; std::unique_ptr<StereoEffect> fx; std::unique_ptr<Waveshaper> ws; std::unique_ptr<FeedbackWaveshaper> fb; };
void* orc_fx_new(uint32_t kind, float sr, const float* c) {
  auto b = std::make_unique<FxBox>();
  b->kind = kind;
  switch (kind) {
    case 0: b->fx.reset(new LowpassFilterEffect(sr, c[0], c[1])); break;
    case 1: b->fx.reset(new DelayEffect(sr, (uint32_t)c[0], c[1], c[2], c[3], c[4])); break;
    case 2: b->fx.reset(new TubeSaturation(sr, c[0], c[1], c[2])); break;
    case 3: b->fx.reset(new TubeCompressor(sr, c[0], c[1], c[2], c[3], c[4])); break;
    case 4: b->fx.reset(new TiltFilterEffect(sr)); break;
    case 6: b->fx.reset(new SpringReverbEffect(sr, c[0], c[1], c[2])); break;
    case 7: b->ws.reset(new Waveshaper(c[0], c[1])); break;
    case 8: b->fb.reset(new FeedbackWaveshaper(sr, c[0], c[1], c[2], c[3])); break;
    case 9: b->fx.reset(new PlateReverbEffect(sr, c[0], c[1], c[2])); break;
    default: return nullptr;
  }
  return b.release();
}
void orc_fx_free(void* h) { delete (FxBox*)h; }
void orc_fx_set_param(void* h, uint32_t p, float v) {
  FxBox* b = (FxBox*)h;
  if (b->fx) b->fx->set_param(p, v);
  else if (b->ws) { if (p == 0) b->ws->set_drive(v); else if (p == 1) b->ws->set_mix(v); }
  else if (b->fb) { if (p == 0) b->fb->set_drive(v); else if (p == 1) b->fb->set_feedback(v); else if (p == 2) b->fb->set_filter_cutoff(v); else if (p == 3) b->fb->set_mix(v); }
}
void orc_fx_set_bpm(void* h, float bpm) { FxBox* b = (FxBox*)h; if (b->fx) b->fx->set_bpm(bpm); }
void orc_fx_reset(void* h) {
  FxBox* b = (FxBox*)h;
  if (b->ws) { b->ws->reset(); return; }
  if (b->fb) { b->fb->reset(); return; }
  switch (b->kind) {
    case 0: static_cast<LowpassFilterEffect*>(b->fx.get())->reset(); break;
    case 1: static_cast<DelayEffect*>(b->fx.get())->reset(); break;
    case 2: static_cast<TubeSaturation*>(b->fx.get())->reset(); break;
    case 3: static_cast<TubeCompressor*>(b->fx.get())->reset(); break;
    case 4: static_cast<TiltFilterEffect*>(b->fx.get())->reset(); break;
    case 6: static_cast<SpringReverbEffect*>(b->fx.get())->reset(); break;
    case 9: static_cast<PlateReverbEffect*>(b->fx.get())->reset(); break;
  }
}
void orc_fx_process(void* h, const float* in, float* out, uint32_t n) {     // the mono `process` of the effect
  FxBox* b = (FxBox*)h;
  for (uint32_t i = 0; i < n; i++) out[i] = b->fx ? b->fx->process(in[i]) : (b->ws ? b->ws->process(in[i]) : b->fb->process(in[i]));
}
void orc_fx_process_stereo(void* h, const float* l, const float* r, float* ol, float* orr, uint32_t n) {
  FxBox* b = (FxBox*)h;
  if (!b->fx) return;
  for (uint32_t i = 0; i < n; i++) { StereoFrame f; f.l = l[i]; f.r = r[i]; f = b->fx->process_stereo(f); ol[i] = f.l; orr[i] = f.r; }
}
// ---- stand-alone filters and generators (the reference's inline unit tests of src/filters/*.rs and src/gen/*.rs) ----
// kind: 0 Chamberlin SVF (mode = type 0 lp / 1 bp / 2 hp), 1 TPT SVF (mode), 2 resonant low-pass, 3 band-pass biquad (a = freq, b = q, c = gain),
// 4 high-pass biquad (a = freq, b = q), 5 membrane resonator (a = gain scale, <= 0: default)
void orc_filter_run(uint32_t kind, float sr, float a, float b, float c, uint32_t mode, const float* in, float* out, uint32_t n) {
  switch (kind) {
    case 0: { StateVariableFilter f(sr, a, b); for (uint32_t i = 0; i < n; i++) out[i] = f.process_mode(in[i], (uint8_t)mode); } break;
    case 1: { StateVariableFilterTpt f(sr, a, b); for (uint32_t i = 0; i < n; i++) out[i] = f.process_mode(in[i], (uint8_t)mode); } break;
    case 2: { ResonantLowpassFilter f(sr, a, b); for (uint32_t i = 0; i < n; i++) out[i] = f.process(in[i]); } break;
    case 3: { BiquadBandpass f(sr); f.set_params(a, b, c); for (uint32_t i = 0; i < n; i++) out[i] = f.process(in[i]); } break;
    case 4: { BiquadHighpass f(sr); f.set_params(a, b); for (uint32_t i = 0; i < n; i++) out[i] = f.process(in[i]); } break;
    case 5: { MembraneResonator f(sr); if (a > 0.0f) f.set_gain_scale(a); for (uint32_t i = 0; i < n; i++) out[i] = f.process(in[i]); } break;
    default: for (uint32_t i = 0; i < n; i++) out[i] = 0.0f;
  }
}
void orc_polyblep(int square, double inc, float* out, uint32_t n) {   // gen/polyblep.rs tests: phase += inc; phase -= floor(phase)
  double phase = 0.0;
  for (uint32_t i = 0; i < n; i++) { out[i] = square ? polyblep_square(phase, inc) : polyblep_saw(phase, inc); phase += inc; phase -= floor(phase); }
}
void orc_morph_osc(float sr, float freq, float morph, float color, float tone, float* out, uint32_t n) {
  MorphOsc o(sr);
  for (uint32_t i = 0; i < n; i++) out[i] = o.tick(freq, morph, color, tone);
}
// ---- SmoothedParam, StereoFrame and MixerGraph micro-operations (inline unit tests of utils/smoother.rs, frame.rs, mixer/graph.rs) ----
// ops: a script of (code, value): 0 set_target, 1 set_immediate, 2 snap, 3 set_normalized, 4 set_bipolar, 5 tick x (int)value.
// out = {current, target, settled}
void orc_smoother_script(float init, float mn, float mx, float sr, float ms, const uint32_t* codes, const float* values, uint32_t n, float* out3) {
  SmoothedParam p(init, mn, mx, sr, ms);
  for (uint32_t i = 0; i < n; i++) {
    switch (codes[i]) {
      case 0: p.set_target(values[i]); break;
      case 1: p.set_immediate(values[i]); break;
      case 2: p.snap(); break;
      case 3: p.set_normalized(values[i]); break;
      case 4: p.set_bipolar(values[i]); break;
      case 5: for (int k = 0; k < (int)values[i]; k++) p.tick(); break;
    }
  }
  out3[0] = p.get(); out3[1] = p.target; out3[2] = p.is_settled() ? 1.0f : 0.0f;
}
void orc_frame_panned(float x, float pan, float* lr) { StereoFrame f = StereoFrame::panned(x, pan); lr[0] = f.l; lr[1] = f.r; }
float orc_frame_downmix(float l, float r) { StereoFrame f; f.l = l; f.r = r; return f.downmix(); }
// One-track graph fed by the drum-kit source: gain / pan / mute / solo of track 0 (and an optional second track that is soloed),
// strips snapped, one frame scattered and mixed down; out = {l, r, peak after, peak after second read}
void orc_graph_one_frame(float gain, float pan, int mute, int second_track_solo, float in_l, float in_r, float* out4) {
  MixerGraph g(44100.0f, 120.0f);
  size_t t = g.add_track();
  g.route(0, t);
  g.tracks[t].gain.set_target(gain);
  g.tracks[t].pan.set_target(pan);
  g.tracks[t].muted = mute != 0;
  if (second_track_solo) { size_t s2 = g.add_track(); g.tracks[s2].soloed = true; }
  g.snap_strip_params();
  g.clear_scratch();
  StereoFrame f; f.l = in_l; f.r = in_r;
  g.scatter(0, f);
  StereoFrame o = g.mix_down();
  out4[0] = o.l; out4[1] = o.r;
  out4[2] = g.tracks[t].peak; g.tracks[t].peak = 0.0f; out4[3] = g.tracks[t].peak;
}
// tests/aliasing.rs render_oversampled_naive_square: a naive square generated INSIDE the oversampler's callback at the oversampled rate
void orc_oversample_square(int mode, double sub_dt, float* out, uint32_t n) {
  Oversampler os;
  os.set_mode(mode == 0 ? OversamplingMode::Off : (mode == 2 ? OversamplingMode::X2 : OversamplingMode::X4));
  double phase = 0.0;
  for (uint32_t i = 0; i < n; i++)
    out[i] = os.process(0.0f, [&](float) { float v = phase < 0.5 ? 1.0f : -1.0f; phase = phase + sub_dt; phase -= floor(phase); return v; });
}
// max_curve.rs test_envelope_basic: two-segment MaxCurveEnvelope sampled at the given times after trigger(0)
void orc_maxcurve_envelope(const float* seg6, const double* times, float* out, uint32_t n) {
  MaxCurveEnvelope env({{seg6[0], seg6[1], seg6[2]}, {seg6[3], seg6[4], seg6[5]}});
  env.trigger(0.0);
  for (uint32_t i = 0; i < n; i++) out[i] = env.get_value(times[i]);
}
// utils/blendable.rs tests: a two-field config blended bilinearly over four corners (ChannelBlender::blend with n = 2, no discrete field)
void orc_blend2(const float* corners8, float x, float y, float* out2) {
  ChannelBlender b;
  b.type = 0; b.n = 2;
  for (int c = 0; c < 4; c++) { b.corner[c][0] = corners8[2 * c]; b.corner[c][1] = corners8[2 * c + 1]; }
  float o[24];
  b.blend(x, y, o);
  out2[0] = o[0]; out2[1] = o[1];
}
// ---- bare PolySynth / Granulator renders (tests/test_emu_cpu.py: the product's poly_tick / gran_tick host builds must match bit for bit) ----
// events sorted by frame; kind 0 = trigger_note(a = midi note, b = velocity), 1 = release_all, 2 = set_param(a = id, b = value).  The clock is the
// engine's: t += 1/sr by repeated addition; an event at frame j is applied before tick j (trigger_note reads current_time = the previous tick's time).
void orc_poly_render(uint32_t preset, float sr, uint32_t n_ev, const uint32_t* ev_frame, const uint32_t* ev_kind, const float* ev_a, const float* ev_b,
                     uint32_t frames, float* out) {
  PolySynth syn(sr, PolyConfig::preset(preset));
  double t = 0.0; const double dt = 1.0 / (double)sr;
  uint32_t e = 0;
  for (uint32_t j = 0; j < frames; j++) {
    while (e < n_ev && ev_frame[e] <= j) {
      if (ev_kind[e] == 0) syn.trigger_note((uint8_t)ev_a[e], ev_b[e]); else if (ev_kind[e] == 1) syn.release_all(); else syn.set_param((uint32_t)ev_a[e], ev_b[e]);
      e++;
    }
    out[j] = syn.tick(t);
    t += dt;
  }
}
// kind 0 = trigger(b = velocity), 2 = set_param(a, b), 3 = snap_params, 4 = set_seed(a)
void orc_gran_render(float sr, const float* buf, uint32_t buf_len, float buf_sr, uint32_t n_ev, const uint32_t* ev_frame, const uint32_t* ev_kind,
                     const float* ev_a, const float* ev_b, uint32_t frames, float* out) {
  Granulator g(sr);
  g.set_buffer(std::make_shared<std::vector<float>>(buf, buf + buf_len), buf_sr);
  double t = 0.0; const double dt = 1.0 / (double)sr;
  uint32_t e = 0;
  for (uint32_t j = 0; j < frames; j++) {
    while (e < n_ev && ev_frame[e] <= j) {
      switch (ev_kind[e]) {
        case 0: g.trigger_with_velocity(t, ev_b[e]); break;
        case 2: g.set_param((uint32_t)ev_a[e], ev_b[e]); break;
        case 3: g.snap_params(); break;
        case 4: g.set_seed((uint32_t)ev_a[e]); break;
      }
      e++;
    }
    out[j] = g.tick(t);
    t += dt;
  }
}
float orc_limiter(float threshold, float x) { SoftLimiter lim(1.0f); lim.set_threshold(threshold); return lim.process(x); }

}  // extern "C"
