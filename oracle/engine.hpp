// oracle/engine.hpp — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the control + mix layer: Sequencer (engine/sequencer.rs), MixerGraph (mixer/graph.rs),
// the Rust-API Engine + bounce (engine/mod.rs, bounce.rs) and the C-FFI GooeyEngine (ffi.rs:570-1541, 7833-7884).
// Out of scope here exactly as in SURVEY.md §2: performance recorder, loop mixer,
// samplers, host-time arm (all default-off / contribute +0.0 on the bounce path).
#pragma once
#include <memory>
#include <string>
#include "synths.hpp"
#include "sources.hpp"
#include "effects.hpp"
#include "loops.hpp"

namespace orc {

// ---- engine/sequencer.rs -----------------------------------------------------------------------------
struct SeqStep { bool enabled = true; float velocity = 1.0f; bool has_blend = false; float bx = 0, by = 0; bool has_note = false; uint8_t note = 0; };
struct SeqTrigger { float velocity; bool has_blend; float bx, by; bool has_note; uint8_t note; };
struct Sequencer {
  float bpm, sample_rate;
  uint64_t sample_count = 0, next_trigger_sample = 0;
  float samples_per_step;
  uint64_t step_start_sample = 0;
  std::vector<SeqStep> pattern;
  size_t current_step = 0, playhead_step = 0;
  bool is_running = false;
  SmoothedParam swing;
  static float calc_sps(float bpm, float sr) { float s16 = (60.0f / bpm) / 4.0f; return s16 * sr; }  // :583-588
  Sequencer(float bpm_, float sr, size_t steps, bool enabled) : bpm(bpm_), sample_rate(sr), samples_per_step(calc_sps(bpm_, sr)), swing(0.5f, 0.0f, 1.0f, sr, 15.0f) {
    pattern.resize(steps);
    for (auto& s : pattern) s.enabled = enabled;
  }
  void start() { is_running = true; next_trigger_sample = sample_count; }
  void stop() { is_running = false; }
  void reset() { sample_count = 0; next_trigger_sample = 0; step_start_sample = 0; current_step = 0; playhead_step = 0; }
  void set_bpm(float b) { bpm = b; samples_per_step = calc_sps(b, sample_rate); }
  void set_swing(float s) { swing.set_target(clampf(s, 0.0f, 1.0f)); }
  void set_beat_position(double beat) {  // :658-682 (a step is a 16th = 1/4 beat)
    size_t n = pattern.size();
    if (n == 0) return;
    double step_f = beat * 4.0, fl = std::floor(step_f);
    size_t idx = (size_t)fl % n;
    double frac = step_f - fl;
    current_step = idx; playhead_step = idx;
    double off = frac * (double)samples_per_step;
    sample_count = off <= 0.0 ? 0 : (uint64_t)off;
    step_start_sample = 0;
    double nt = std::round((double)samples_per_step - frac * (double)samples_per_step);
    next_trigger_sample = nt <= 0.0 ? 0 : (uint64_t)nt;
  }
  bool tick(SeqTrigger& out) {  // tick_with_settings :883-952 (armed start not modelled)
    if (!is_running || pattern.empty()) { sample_count += 1; return false; }
    swing.tick();
    bool fired = false;
    if (sample_count >= next_trigger_sample) {
      step_start_sample = sample_count;
      playhead_step = current_step;
      const SeqStep& st = pattern[current_step];
      if (st.enabled) { out = {st.velocity, st.has_blend, st.bx, st.by, st.has_note, st.note}; fired = true; }
      current_step = (current_step + 1) % pattern.size();
      float swing_offset = (swing.get() - 0.5f) * 2.0f * samples_per_step;
      float signed_off = (current_step % 2 == 1) ? swing_offset : -swing_offset;
      float nx = roundf((float)next_trigger_sample + samples_per_step + signed_off);
      next_trigger_sample = f32_as_u64(nx);
    }
    sample_count += 1;
    return fired;
  }
};

// ---- mixer/graph.rs --------------------------------------------------------------------------------------
static inline std::unique_ptr<StereoEffect> make_channel_effect(uint32_t id, float sr, float bpm) {  // effect_chain.rs:57-109
  switch (id) {
    case 0: return std::make_unique<LowpassFilterEffect>(sr, 20000.0f, 0.0f);
    case 1: return std::make_unique<DelayEffect>(sr, 2, bpm, 0.3f, 0.3f, 8000.0f);
    case 2: return std::make_unique<TubeSaturation>(sr, 0.3f, 0.4f, 0.5f);
    case 3: return std::make_unique<TubeCompressor>(sr, -12.0f, 4.0f, 5.0f, 100.0f, 0.5f);
    case 4: return std::make_unique<TiltFilterEffect>(sr);
    case 6: return std::make_unique<SpringReverbEffect>(sr, 0.5f, 0.3f, 0.5f);
    case 7: return std::make_unique<WaveshaperPair>();
    case 8: return std::make_unique<FeedbackWaveshaperPair>(sr);
    case 9: return std::make_unique<PlateReverbEffect>(sr, 0.5f, 0.3f, 0.5f);
    default: return nullptr;  // the master limiter is not a channel effect
  }
}
struct Track {
  SmoothedParam gain, pan, mute_gain;
  bool muted = false, soloed = false;
  float peak = 0.0f;                                  // post-strip peak, read-and-reset (graph.rs:93-98, 233-237)
  std::vector<std::unique_ptr<StereoEffect>> rack;
  explicit Track(float sr) : gain(1.0f, 0.0f, 2.0f, sr, 10.0f), pan(0.5f, 0.0f, 1.0f, sr, 10.0f), mute_gain(1.0f, 0.0f, 1.0f, sr, 10.0f) {}
};
struct MixerGraph {
  static const int SOURCE_CAPACITY = 9;
  std::vector<Track> tracks;
  int routes[SOURCE_CAPACITY];
  std::vector<StereoFrame> scratch;
  float sample_rate, bpm;
  bool active_sources[SOURCE_CAPACITY];               // loop mixer and below always; sampler racks once registered (graph.rs:120, 269-282)
  MixerGraph(float sr, float b) : sample_rate(sr), bpm(b) { for (int& r : routes) r = -1; for (int i = 0; i < SOURCE_CAPACITY; i++) active_sources[i] = i < 5; }
  bool register_source(uint32_t src) { if (src >= (uint32_t)SOURCE_CAPACITY) return false; active_sources[src] = true; return true; }
  bool source_is_active(uint32_t src) const { return src < (uint32_t)SOURCE_CAPACITY && active_sources[src]; }
  size_t add_track() { tracks.emplace_back(sample_rate); scratch.push_back({}); return tracks.size() - 1; }
  void default_layout() { add_track(); add_track(); add_track(); add_track(); route(0, 0); route(1, 1); route(2, 2); route(3, 3); route(4, 3); }
  bool route(uint32_t src, size_t track) { if (source_is_active(src) && track < tracks.size()) { routes[src] = (int)track; return true; } return false; }
  bool unroute(uint32_t src) { if (!source_is_active(src)) return false; bool had = routes[src] >= 0; routes[src] = -1; return had; }   // :253-259
  int route_of(uint32_t src) const { return source_is_active(src) ? routes[src] : -1; }                                             // :262-266
  void reset() { tracks.clear(); scratch.clear(); for (int& r : routes) r = -1; }                                                     // :145-149
  void set_bpm(float b) { bpm = b; for (auto& t : tracks) for (auto& e : t.rack) e->set_bpm(b); }
  void clear_scratch() { for (auto& s : scratch) s = {}; }
  void scatter(uint32_t src, StereoFrame f) { if (source_is_active(src) && routes[src] >= 0 && (size_t)routes[src] < scratch.size()) scratch[routes[src]] += f; }
  void update_mute_solo_targets() {
    bool any = false;
    for (auto& t : tracks) any |= t.soloed;
    for (auto& t : tracks) t.mute_gain.set_target(t.soloed ? 1.0f : ((any || t.muted) ? 0.0f : 1.0f));
  }
  void snap_strip_params() { update_mute_solo_targets(); for (auto& t : tracks) { t.gain.snap(); t.pan.snap(); t.mute_gain.snap(); } }
  static StereoFrame balanced(StereoFrame f, float pan) {  // :50-58
    pan = clampf(pan, 0.0f, 1.0f);
    float lg = rust_min(2.0f * (1.0f - pan), 1.0f), rg = rust_min(2.0f * pan, 1.0f);
    return {f.l * lg, f.r * rg};
  }
  StereoFrame mix_down() {  // :385-399
    StereoFrame master;
    for (size_t i = 0; i < tracks.size(); i++) {
      Track& t = tracks[i];
      float g = t.gain.tick() * t.mute_gain.tick();
      StereoFrame f = scratch[i].scaled(g);
      f = balanced(f, t.pan.tick());
      for (auto& e : t.rack) f = e->process_stereo(f);
      { float lv = rust_max(fabsf(f.l), fabsf(f.r)); if (lv > t.peak) t.peak = lv; }   // record_peak(f.l.abs().max(f.r.abs())) (:395)
      master += f;
    }
    return master;
  }
};

// ---- utils/blendable.rs PresetBlender + ffi.rs ChannelBlender (:405-560) --------------------------------------
// Presets as flat values in config-field order; ids ffi.rs:1882-1998.  Returns the number of fields (0 = unknown id).
static inline int preset_flat(uint32_t type, uint32_t id, float* v) {
  if (id > 3) return 0;
  switch (type) {
    case 0: { KickConfig c = id == 0 ? KickConfig::tight() : id == 1 ? KickConfig::punch() : id == 2 ? KickConfig::loose() : KickConfig::dirt();
      for (int i = 0; i < 18; i++) { v[i] = c.v[i]; }
      return 18; }
    case 1: { SnareConfig c = id == 0 ? SnareConfig::tight() : id == 1 ? SnareConfig::loose() : id == 2 ? SnareConfig::hiss() : SnareConfig::smack();
      const float a[19] = {c.frequency, c.tonal_amount, c.noise_amount, c.crack_amount, c.decay, c.pitch_drop, c.volume, c.tonal_decay, c.tonal_decay_curve,
                           c.noise_decay, c.noise_tail_decay, c.filter_cutoff, c.filter_resonance, (float)c.filter_type, c.xfade, c.phase_mod_amount,
                           c.overdrive_amount, c.amp_decay, c.amp_decay_curve};
      for (int i = 0; i < 19; i++) { v[i] = a[i]; }
      return 19; }
    case 2: { HiHat2Config c = id == 0 ? HiHat2Config::short_() : id == 1 ? HiHat2Config::loose() : id == 2 ? HiHat2Config::dark() : HiHat2Config::soft();
      const float a[7] = {c.pitch, c.decay, c.attack, c.pink ? 1.0f : 0.0f, c.db24 ? 1.0f : 0.0f, c.tone, c.volume};
      for (int i = 0; i < 7; i++) { v[i] = a[i]; }
      return 7; }
    case 3: {  // Tom2Config::{derp,ring,brush,void_preset} (tom2.rs:119-172)
      static const float T[4][8] = {{60.0f, 70.0f, 50.0f, 0.0f, 20.0f, 0.0f, 50.0f, 100.0f}, {80.0f, 20.0f, 10.0f, 0.0f, 100.0f, 60.0f, 70.0f, 100.0f},
                                    {40.0f, 20.0f, 10.0f, 90.0f, 30.0f, 0.0f, 50.0f, 100.0f}, {60.0f, 30.0f, 100.0f, 50.0f, 90.0f, 40.0f, 80.0f, 100.0f}};
      for (int i = 0; i < 8; i++) { v[i] = T[id][i]; }
      return 8; }
    case 4: { BassConfig c = id == 0 ? BassConfig::acid() : id == 1 ? BassConfig::sub() : id == 2 ? BassConfig::reese() : BassConfig::stab();
      for (int i = 0; i < 15; i++) { v[i] = c.v[i]; }
      return 15; }
    default: return 0;
  }
}
struct ChannelBlender {
  uint32_t type = 0; int n = 0;
  float corner[4][24];                       // bottom_left, bottom_right, top_left, top_right
  uint32_t corner_ids[4] = {0, 1, 2, 3};     // default_corner_preset_ids: presets 0..3 of the type
  void default_for_type(uint32_t t) { type = t; for (uint32_t c = 0; c < 4; c++) { n = preset_flat(t, c, corner[c]); corner_ids[c] = c; } }
  bool discrete(int i) const { return (type == 1 && i == 13) || (type == 2 && (i == 3 || i == 4)); }   // filter_type / noise_color / filter_slope: `if t < 0.5 { self } else { other }`
  void lerp(const float* a, const float* b, float t, float* o) const {
    t = clampf(t, 0.0f, 1.0f);
    const float inv_t = 1.0f - t;
    for (int i = 0; i < n; i++) o[i] = discrete(i) ? (t < 0.5f ? a[i] : b[i]) : a[i] * inv_t + b[i] * t;
  }
  void blend(float x, float y, float* o) const {   // blendable.rs:73-86
    x = clampf(x, 0.0f, 1.0f); y = clampf(y, 0.0f, 1.0f);
    float bottom[24], top[24];
    lerp(corner[0], corner[1], x, bottom);
    lerp(corner[2], corner[3], x, top);
    lerp(bottom, top, y, o);
  }
  void set_corner_preset(uint32_t c, uint32_t id) { float v[24]; if (c < 4 && preset_flat(type, id, v) > 0) for (int i = 0; i < n; i++) corner[c][i] = v[i]; }
};

// ---- engine/lfo.rs Lfo (sine, free-running or tempo-synced) + the FFI's LFO pool (ffi.rs:33-54, 716-719, 1238-1251) ----
struct Lfo {
  bool synced = true; uint32_t division = 4; float hz = 1.0f;   // Lfo::with_sample_rate: BpmSync(Quarter) (lfo.rs:86-97)
  float bpm = 120.0f, phase = 0.0f, sample_rate;
  float amount = 1.0f, offset = 0.0f;
  explicit Lfo(float sr) : sample_rate(sr) {}
  static float beats(uint32_t d) { const float B[8] = {16.0f, 8.0f, 4.0f, 2.0f, 1.0f, 0.5f, 0.25f, 0.125f}; return B[d < 8 ? d : 4]; }   // lfo.rs:14-25
  float frequency() const { if (!synced) return hz; float bps = bpm / 60.0f; return bps / beats(division); }                            // :27-33, 155-160
  float tick() {                                                                                                                          // :170-185
    float value = sinf(phase * 2.0f * 3.14159265358979323846f);
    float inc = frequency() / sample_rate;
    phase += inc;
    if (phase >= 1.0f) phase -= 1.0f;
    return offset + (value * amount);
  }
};
struct LfoRoute { uint32_t id, instrument, param; float depth; };

// ---- ffi.rs GooeyEngine -------------------------------------------------------------------------------------
struct VoiceStrip {
  std::unique_ptr<Instrument> inst;
  uint32_t type;
  Sequencer seq;
  SmoothedParam channel_gain, mute_gain, pan;
  bool muted = false, soloed = false, trigger_pending = false;
  float trigger_velocity = 1.0f;
  bool has_saved = false;
  float saved_freq = 0;
  float peak = 0.0f;                                  // pre-pan mono peak, read-and-reset (ffi.rs:654-659, 2572-2584)
  ChannelBlender blender; bool blend_enabled = false; float blend_x = 0.5f, blend_y = 0.5f;   // ffi.rs:598-600, 635-637
  VoiceStrip(std::unique_ptr<Instrument> i, uint32_t t, float bpm, float sr)
      : inst(std::move(i)), type(t), seq(bpm, sr, 16, false), channel_gain(1.0f, 0, 1, sr, 10.0f), mute_gain(1.0f, 0, 1, sr, 10.0f), pan(0.5f, 0, 1, sr, 10.0f) { blender.default_for_type(t); }
  void blend_and_apply(float x, float y) { float v[24]; blender.blend(x, y, v); inst->set_config_flat(v); }   // ffi.rs:419-428
};
static inline std::unique_ptr<Instrument> make_instrument(uint32_t type, float sr) {
  switch (type) {
    case 0: return std::make_unique<KickDrum>(sr);
    case 1: return std::make_unique<SnareDrum>(sr);
    case 2: return std::make_unique<HiHat2>(sr);
    case 3: return std::make_unique<Tom2>(sr);
    case 4: return std::make_unique<BassSynth>(sr);
    default: return nullptr;
  }
}

struct FfiEngine {
  float sample_rate, bpm = 120.0f, swing = 0.5f;
  double current_time = 0.0;
  std::vector<VoiceStrip> voices;  // 0..3 kit, 4 bass
  DelayEffect delay; bool delay_enabled = false;
  TiltFilterEffect tilt; bool tilt_enabled = false;
  SpringReverbEffect reverb; bool reverb_enabled = false;
  PlateReverbEffect plate; bool plate_enabled = false;
  SoftLimiter limiter; bool limiter_enabled = false;
  LowpassFilterEffect lowpass; bool lowpass_enabled = false;
  TubeSaturation saturation; bool saturation_enabled = false;
  TubeCompressor compressor; bool compressor_enabled = false; uint32_t compressor_sidechain = 0xFFFFFFFFu;
  Waveshaper waveshaper; bool waveshaper_enabled = false;                          // ONE instance: L then R through the same state (ffi.rs:1344-1349)
  FeedbackWaveshaper feedback_waveshaper; bool feedback_waveshaper_enabled = false;
  uint32_t effect_order[9] = {7, 2, 0, 4, 1, 3, 8, 6, 9};  // DEFAULT_EFFECT_ORDER (ffi.rs:1583-1593)
  SmoothedParam master_gain;
  bool seq_triggers_enabled = true;
  PolySynth poly;
  Granulator granulator;
  MixerGraph graph;
  LoopMixer mixer;                                   // ffi.rs:768 (`mixer: Mixer`)
  std::unique_ptr<SamplerRack> samplers[4];          // ffi.rs:770
  // the rack's own 16-step sequencer (step note = pad) and its transport-armed start (sampler.rs:160-175, 232-310)
  struct RackPattern {
    Sequencer seq; bool pattern_running = false, has_pending = false; double pending_start_beat = 0.0;
    RackPattern(float bpm, float sr) : seq(bpm, sr, 16, false) {}
  };
  std::vector<RackPattern> rack_pat;
  struct RackHitLog { uint64_t frame; uint32_t slot; float velocity; };   // test hook: the pattern hits of rack 0..3 and the frames rendered so far
  std::vector<RackHitLog> rack_hits[4]; uint64_t frames_rendered = 0;
  bool capture_rack0 = false; std::vector<float> rack0_capture;          // test hook: rack 0's own stereo frames while the engine renders
  // the mixer's clip-grid transport, as far as the racks need it (clip_grid.rs:144-195, 526-579, 656-660): a monotonic beat clock
  bool transport_running = false; double transport_beat = 0.0; float transport_bpm = 120.0f;
  double beats_per_sample() const { return (double)rust_max(transport_bpm, 0.0f) / (60.0 * (double)rust_max(sample_rate, 1.0f)); }
  double quantized_target(double interval) const {  // clip_grid.rs:174-191
    if (!transport_running) return 0.0;
    double scaled = transport_beat / interval, nearest = std::round(scaled);
    double base = std::fabs(scaled - nearest) <= 1.0e-9 ? nearest : std::floor(scaled);
    return (base + 1.0) * interval;
  }
  void rack_stop_all(int r) { if (samplers[r]) for (auto& v : samplers[r]->voices) v.buffer.reset(); }
  void sequencer_start() { for (auto& v : voices) v.seq.start(); for (int r = 0; r < 4; r++) if (samplers[r]) rack_pat[r].seq.start(); transport_running = true; }   // ffi.rs:3501-3512
  void sequencer_stop() {   // :3515-3530
    for (auto& v : voices) v.seq.stop();
    for (int r = 0; r < 4; r++) if (samplers[r]) { auto& p = rack_pat[r]; p.seq.stop(); p.has_pending = false; p.pattern_running = false; p.seq.stop(); rack_stop_all(r); }
    transport_running = false;
  }
  void sequencer_reset() {  // :3547-3561
    for (auto& v : voices) v.seq.reset();
    for (int r = 0; r < 4; r++) if (samplers[r]) { auto& p = rack_pat[r]; p.seq.reset(); p.has_pending = false; p.pattern_running = false; p.seq.reset(); rack_stop_all(r); }
    transport_beat = 0.0;
  }
  std::vector<Lfo> lfos; bool lfo_enabled[8] = {false}; std::vector<LfoRoute> lfo_routes[8]; uint32_t lfo_next_route_id[8] = {0};
  struct MidiEvent { uint32_t instrument_index; float velocity; uint32_t sample_offset; };   // GooeyMidiEvent (ffi.rs:78-83)
  std::vector<MidiEvent> pending_midi_events;                                                // capacity 64, cleared by every render (:71, :1045)
  void push_midi_event(uint32_t ch, float vel, uint32_t off) { if (pending_midi_events.size() < 64) pending_midi_events.push_back({ch, vel, off}); }
  explicit FfiEngine(float sr)
      : sample_rate(sr), delay(sr, 2, 120.0f, 0.0f, 0.0f, 20000.0f), tilt(sr), reverb(sr, 0.5f, 0.0f, 0.5f), plate(sr, 0.5f, 0.0f, 0.5f),
        limiter(1.0f), lowpass(sr, 20000.0f, 0.0f), saturation(sr, 0.3f, 0.4f, 0.5f), compressor(sr, -12.0f, 4.0f, 5.0f, 100.0f, 0.5f),
        waveshaper(1.0f, 0.0f), feedback_waveshaper(sr, 1.0f, 0.0f, 2000.0f, 0.0f),
        master_gain(0.25f, 0.0f, 2.0f, sr, 30.0f), poly(sr), granulator(sr), graph(sr, 120.0f), mixer(sr) {
    for (uint32_t t = 0; t < 5; t++) voices.emplace_back(make_instrument(t, sr), t, bpm, sr);
    for (int i = 0; i < 8; i++) lfos.emplace_back(sr);
    for (int r = 0; r < 4; r++) rack_pat.emplace_back(bpm, sr);
    graph.default_layout();
  }
  VoiceStrip* by_type(uint32_t t) { for (auto& v : voices) if (v.type == t) return &v; return nullptr; }
  void set_bpm(float b) { bpm = b; for (auto& v : voices) v.seq.set_bpm(b); for (auto& l : lfos) l.bpm = b; delay.set_bpm(b); graph.set_bpm(b); mixer.set_bpm(b);
    for (int r = 0; r < 4; r++) if (samplers[r]) rack_pat[r].seq.set_bpm(b);
    if (std::isfinite(b) && b > 0.0f) transport_bpm = b; }   // :3337-3364, clip_grid.rs:520-524
  void set_swing(float s) { swing = clampf(s, 0.0f, 1.0f); for (auto& v : voices) v.seq.set_swing(swing); for (int r = 0; r < 4; r++) if (samplers[r]) rack_pat[r].seq.set_swing(swing); }
  void reset_effect_states() { saturation.reset(); lowpass.reset(); tilt.reset(); delay.reset(); compressor.reset(); reverb.reset(); plate.reset(); }  // ffi.rs:1417-1425
  static bool freq_range(uint32_t type, float& mn, float& mx) {  // :1511-1518
    if (type == 4) { mn = 30.0f; mx = 200.0f; return true; }
    if (type == 0) { mn = 30.0f; mx = 120.0f; return true; }
    if (type == 3) { mn = 40.0f; mx = 600.0f; return true; }
    return false;
  }
  static float midi_to_norm(uint8_t note, float mn, float mx) { float hz = 440.0f * powf(2.0f, ((float)note - 69.0f) / 12.0f); return clampf((hz - mn) / (mx - mn), 0.0f, 1.0f); }

  void render(float* buffer, size_t frames) {  // ffi.rs:1043-1382
    pending_midi_events.clear();
    for (size_t ch = 0; ch < voices.size(); ch++) {
      VoiceStrip& v = voices[ch];
      if (v.trigger_pending) { v.trigger_pending = false; push_midi_event((uint32_t)ch, v.trigger_velocity, 0); v.inst->trigger_with_velocity(current_time, v.trigger_velocity); }
    }
    const double period = 1.0 / (double)sample_rate;
    bool any_solo = false;
    for (auto& v : voices) any_solo |= v.soloed;
    for (auto& v : voices) v.mute_gain.set_target(v.soloed ? 1.0f : (any_solo ? 0.0f : (v.muted ? 0.0f : 1.0f)));
    graph.update_mute_solo_targets();
    for (size_t f = 0; f < frames; f++) {
      if (transport_running)   // a rack start is owned by the render clock (:1139-1147; SamplerRack::activate_start_if_due)
        for (int r = 0; r < 4; r++) {
          if (!samplers[r]) continue;
          RackPattern& p = rack_pat[r];
          if (!p.has_pending || transport_beat + 1.0e-8 < p.pending_start_beat) continue;
          p.has_pending = false;
          p.seq.set_beat_position(p.pending_start_beat);
          p.seq.start();
          p.pattern_running = true;
        }
      SeqTrigger trig[5];
      bool fired[5];
      for (int ch = 0; ch < 5; ch++) fired[ch] = voices[ch].seq.tick(trig[ch]);
      for (int r = 0; r < 4; r++) {   // SamplerRack::tick_sequencer (:1199-1210): ticks only while its pattern runs; hits count only with the triggers enabled
        if (!samplers[r] || !rack_pat[r].pattern_running) continue;
        SeqTrigger t{};
        if (rack_pat[r].seq.tick(t) && seq_triggers_enabled) {
          samplers[r]->trigger(t.has_note ? (size_t)t.note : 0, t.velocity);
          rack_hits[r].push_back({frames_rendered, t.has_note ? (uint32_t)t.note : 0u, t.velocity});
        }
      }
      if (seq_triggers_enabled) {
        double time = current_time;
        for (int ch = 0; ch < 5; ch++) {
          if (!fired[ch]) continue;
          VoiceStrip& v = voices[ch];
          // apply_sequencer_blend_setting (ffi.rs:1384-1402) and the snap that follows an applied blend (:1168-1171)
          if (trig[ch].has_blend) v.blend_and_apply(clampf(trig[ch].bx, 0.0f, 1.0f), clampf(trig[ch].by, 0.0f, 1.0f));
          else if (v.blend_enabled) v.blend_and_apply(v.blend_x, v.blend_y);
          if (trig[ch].has_blend || v.blend_enabled) v.inst->snap_params();
          if (trig[ch].has_note) {
            float mn, mx;
            if (freq_range(v.type, mn, mx)) {
              if (!v.has_saved) { float fq; if (v.inst->get_freq_param(fq)) { v.saved_freq = fq; v.has_saved = true; } }
              v.inst->set_param(0, midi_to_norm(trig[ch].note, mn, mx));
              v.inst->snap_params();
            }
          } else if (v.has_saved) {
            v.has_saved = false;
            v.inst->set_param(0, v.saved_freq);
            v.inst->snap_params();
          }
          v.inst->trigger_with_velocity(time, trig[ch].velocity);
          push_midi_event((uint32_t)ch, trig[ch].velocity, (uint32_t)f);
        }
      }
      for (int li = 0; li < 8; li++) {   // LFO pool (:1238-1251): after the triggers, before the voices tick
        if (!lfo_enabled[li]) continue;
        const float lv = lfos[li].tick();
        for (const LfoRoute& r : lfo_routes[li]) if (r.instrument < voices.size()) voices[r.instrument].inst->apply_modulation(r.param, lv * r.depth);
      }
      StereoFrame kit, bassf;
      double time = current_time;
      float channel_outs[5];
      for (int ch = 0; ch < 5; ch++) {
        VoiceStrip& v = voices[ch];
        float out = v.inst->tick(time) * v.channel_gain.tick() * v.mute_gain.tick();
        channel_outs[ch] = out;
        if (fabsf(out) > v.peak) v.peak = fabsf(out);      // record_peak(ch_out.abs()) (:1281-1282)
        StereoFrame p = StereoFrame::panned(out, v.pan.tick());
        if (ch < 4) kit += p; else bassf += p;
      }
      StereoFrame polyf = StereoFrame::panned(poly.tick(time), 0.5f);
      StereoFrame granf = StereoFrame::panned(granulator.tick(time), 0.5f);
      graph.clear_scratch();
      graph.scatter(0, kit);
      graph.scatter(1, bassf);
      graph.scatter(2, polyf);
      graph.scatter(3, granf);
      StereoFrame sampler_frames[4];                 // :1289-1294
      for (int r = 0; r < 4; r++) if (samplers[r]) sampler_frames[r] = samplers[r]->tick();
      if (capture_rack0) { rack0_capture.push_back(sampler_frames[0].l); rack0_capture.push_back(sampler_frames[0].r); }
      StereoFrame loop_frame = mixer.tick(sample_rate);   // :1296
      if (transport_running) transport_beat += beats_per_sample();   // ClipGrid::after_tick, inside Mixer::tick (mod.rs:73)
      frames_rendered += 1;
      graph.scatter(4, loop_frame);
      for (int r = 0; r < 4; r++) graph.scatter(5 + r, sampler_frames[r]);
      StereoFrame st = graph.mix_down();
      st = st.scaled(master_gain.tick());
      for (uint32_t id : effect_order) {
        if (id == 2 && saturation_enabled) st = saturation.process_stereo(st);
        else if (id == 0 && lowpass_enabled) st = lowpass.process_stereo(st);
        else if (id == 3 && compressor_enabled) {
          if (compressor_sidechain < 5) { StereoFrame sc; sc.l = sc.r = channel_outs[compressor_sidechain]; st = compressor.process_stereo_with_sidechain(st, sc); }
          else st = compressor.process_stereo(st);
        }
        else if (id == 7 && waveshaper_enabled) { float l = waveshaper.process(st.l); float r = waveshaper.process(st.r); st.l = l; st.r = r; }
        else if (id == 8 && feedback_waveshaper_enabled) { float l = feedback_waveshaper.process(st.l); float r = feedback_waveshaper.process(st.r); st.l = l; st.r = r; }
        else if (id == 4 && tilt_enabled) st = tilt.process_stereo(st);
        else if (id == 1 && delay_enabled) st = delay.process_stereo(st);
        else if (id == 6 && reverb_enabled) st = reverb.process_stereo(st);
        else if (id == 9 && plate_enabled) st = plate.process_stereo(st);
      }
      if (limiter_enabled) { st.l = limiter.process(st.l); st.r = limiter.process(st.r); }
      buffer[2 * f] = st.l;
      buffer[2 * f + 1] = st.r;
      current_time += period;
    }
  }

  std::vector<float> bounce_to_buffer(uint32_t bars) {  // ffi.rs:7835-7884
    double spb = 4.0 * (60.0 / (double)bpm) * (double)sample_rate;
    double tot = round((double)bars * spb);
    size_t total = tot <= 0 ? 0 : (size_t)tot;
    current_time = 0.0;
    for (auto& v : voices) { v.seq.reset(); v.seq.start(); }
    for (int r = 0; r < 4; r++) if (samplers[r]) { rack_pat[r].seq.reset(); rack_pat[r].seq.start(); }   // sequencers_iter_mut covers the racks (:3777-3788)
    for (auto& v : voices) { v.mute_gain.snap(); v.channel_gain.snap(); v.pan.snap(); }
    graph.snap_strip_params();
    master_gain.snap();
    std::vector<float> out;
    out.reserve(total);
    std::vector<float> chunk(512 * 2);
    size_t remaining = total;
    while (remaining > 0) {
      size_t n = remaining < 512 ? remaining : 512;
      std::fill(chunk.begin(), chunk.begin() + n * 2, 0.0f);
      render(chunk.data(), n);
      for (size_t i = 0; i < n; i++) out.push_back(0.5f * (chunk[2 * i] + chunk[2 * i + 1]));
      remaining -= n;
    }
    for (auto& v : voices) v.seq.stop();
    for (int r = 0; r < 4; r++) if (samplers[r]) rack_pat[r].seq.stop();
    return out;
  }
};

// ---- engine/mod.rs Engine + bounce.rs (Rust API) ------------------------------------------------------------------
struct RustEngine {
  float sample_rate, bpm = 120.0f;
  std::vector<std::pair<std::string, std::unique_ptr<Instrument>>> instruments;  // insertion order (HashMap order is unspecified)
  std::vector<std::pair<Sequencer, std::string>> sequencers;
  bool limiter_on = true;  // Engine::new pushes SoftLimiter(1.0) (engine/mod.rs:109-112)
  SoftLimiter limiter;
  SmoothedParam master_gain;
  explicit RustEngine(float sr) : sample_rate(sr), limiter(1.0f), master_gain(0.25f, 0.0f, 2.0f, sr, 30.0f) {}
  Instrument* find(const std::string& n) { for (auto& p : instruments) if (p.first == n) return p.second.get(); return nullptr; }
  float tick(double t) {  // engine/mod.rs:343-415
    for (auto& sp : sequencers) {
      SeqTrigger tr;
      if (sp.first.tick(tr)) { if (Instrument* i = find(sp.second)) i->trigger_with_velocity(t, tr.velocity); }
    }
    float out = 0.0f;
    for (auto& p : instruments) out += p.second->tick(t);
    out += 0.0f;  // loop mixer downmix: nothing loaded
    out *= master_gain.tick();
    if (limiter_on) out = limiter.process(out);
    return out;
  }
  std::vector<float> bounce_samples(size_t total) {  // bounce.rs:41-59 with BounceLength::Samples
    for (auto& sp : sequencers) { sp.first.reset(); sp.first.start(); }
    master_gain.snap();
    std::vector<float> buf;
    buf.reserve(total);
    double t = 0.0, step = 1.0 / (double)sample_rate;
    for (size_t i = 0; i < total; i++) { buf.push_back(tick(t)); t += step; }
    for (auto& sp : sequencers) sp.first.stop();
    return buf;
  }
};

}  // namespace orc
