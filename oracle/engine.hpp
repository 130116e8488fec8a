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
  MixerGraph(float sr, float b) : sample_rate(sr), bpm(b) { for (int& r : routes) r = -1; }
  size_t add_track() { tracks.emplace_back(sample_rate); scratch.push_back({}); return tracks.size() - 1; }
  void default_layout() { add_track(); add_track(); add_track(); add_track(); route(0, 0); route(1, 1); route(2, 2); route(3, 3); route(4, 3); }
  bool route(uint32_t src, size_t track) { if (src < 5 && track < tracks.size()) { routes[src] = (int)track; return true; } return false; }
  void set_bpm(float b) { bpm = b; for (auto& t : tracks) for (auto& e : t.rack) e->set_bpm(b); }
  void clear_scratch() { for (auto& s : scratch) s = {}; }
  void scatter(uint32_t src, StereoFrame f) { if (src < 5 && routes[src] >= 0 && (size_t)routes[src] < scratch.size()) scratch[routes[src]] += f; }
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
  VoiceStrip(std::unique_ptr<Instrument> i, uint32_t t, float bpm, float sr)
      : inst(std::move(i)), type(t), seq(bpm, sr, 16, false), channel_gain(1.0f, 0, 1, sr, 10.0f), mute_gain(1.0f, 0, 1, sr, 10.0f), pan(0.5f, 0, 1, sr, 10.0f) {}
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
  struct MidiEvent { uint32_t instrument_index; float velocity; uint32_t sample_offset; };   // GooeyMidiEvent (ffi.rs:78-83)
  std::vector<MidiEvent> pending_midi_events;                                                // capacity 64, cleared by every render (:71, :1045)
  void push_midi_event(uint32_t ch, float vel, uint32_t off) { if (pending_midi_events.size() < 64) pending_midi_events.push_back({ch, vel, off}); }
  explicit FfiEngine(float sr)
      : sample_rate(sr), delay(sr, 2, 120.0f, 0.0f, 0.0f, 20000.0f), tilt(sr), reverb(sr, 0.5f, 0.0f, 0.5f), plate(sr, 0.5f, 0.0f, 0.5f),
        limiter(1.0f), lowpass(sr, 20000.0f, 0.0f), saturation(sr, 0.3f, 0.4f, 0.5f), compressor(sr, -12.0f, 4.0f, 5.0f, 100.0f, 0.5f),
        waveshaper(1.0f, 0.0f), feedback_waveshaper(sr, 1.0f, 0.0f, 2000.0f, 0.0f),
        master_gain(0.25f, 0.0f, 2.0f, sr, 30.0f), poly(sr), granulator(sr), graph(sr, 120.0f) {
    for (uint32_t t = 0; t < 5; t++) voices.emplace_back(make_instrument(t, sr), t, bpm, sr);
    graph.default_layout();
  }
  VoiceStrip* by_type(uint32_t t) { for (auto& v : voices) if (v.type == t) return &v; return nullptr; }
  void set_bpm(float b) { bpm = b; for (auto& v : voices) v.seq.set_bpm(b); delay.set_bpm(b); graph.set_bpm(b); }
  void set_swing(float s) { swing = clampf(s, 0.0f, 1.0f); for (auto& v : voices) v.seq.set_swing(swing); }
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
      SeqTrigger trig[5];
      bool fired[5];
      for (int ch = 0; ch < 5; ch++) fired[ch] = voices[ch].seq.tick(trig[ch]);
      if (seq_triggers_enabled) {
        double time = current_time;
        for (int ch = 0; ch < 5; ch++) {
          if (!fired[ch]) continue;
          VoiceStrip& v = voices[ch];
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
      graph.scatter(4, StereoFrame{});
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
