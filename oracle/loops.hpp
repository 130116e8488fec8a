// oracle/loops.hpp — TEST INFRASTRUCTURE ONLY (tests/, smoke() and bench.py's cpu_baseline leg; never the product).
// CPU restatement of the reference's sample-playback sources:
//   mixer/stereo_buffer.rs (StereoSampleBuffer), mixer/loop_channel.rs (LoopWindow, LoopChannel), mixer/mod.rs (Mixer),
//   instruments/sampler.rs (SamplerBuffer, SampleVoice, SamplerRack).
// Restated: everything a bounce touches with PitchMode Off / Resample and empty per-channel effect chains.  Not restated
// (the product refuses the same requests): PitchMode::PreservePitch (mixer/wsola.rs), queued swaps, per-channel effect
// chains, the clip grid, transport-armed sampler patterns.  Parity unpinned like the rest of the oracle (no reference
// build here); the reference's own unit tests of these files are restated in tests/test_samples_cpu.py.
#pragma once
#include <memory>
#include <vector>
#include "prims.hpp"

namespace orc {

static inline double rem_euclid(double a, double b) { double r = std::fmod(a, b); return r < 0.0 ? r + std::fabs(b) : r; }
static inline double clampd(double x, double lo, double hi) { if (x < lo) return lo; if (x > hi) return hi; return x; }

// ---- mixer/stereo_buffer.rs -------------------------------------------------------------------------------------------
struct StereoSampleBuffer {
  std::vector<float> left, right;
  float sample_rate = 44100.0f;
  bool has_source_bpm = false; float source_bpm = 0.0f;
  // from_interleaved (:57-87) + from_channels (:22-55); nullptr on the reference's Err paths
  static std::shared_ptr<StereoSampleBuffer> from_interleaved(const float* samples, size_t total, size_t channels, float sr) {
    if (channels == 0 || total == 0) return nullptr;
    size_t frames = total / channels;
    if (frames == 0) return nullptr;
    auto b = std::make_shared<StereoSampleBuffer>();
    for (size_t f = 0; f < frames; f++) {
      const float* fr = samples + f * channels;
      b->left.push_back(fr[0]);
      b->right.push_back(channels == 1 ? fr[0] : fr[1]);
    }
    if (!std::isfinite(sr) || sr <= 0.0f) return nullptr;
    for (float s : b->left) if (!std::isfinite(s)) return nullptr;
    for (float s : b->right) if (!std::isfinite(s)) return nullptr;
    b->sample_rate = sr;
    return b;
  }
  size_t len() const { return left.size(); }
  void set_source_bpm(bool some, float bpm) { has_source_bpm = some && std::isfinite(bpm) && bpm > 0.0f; source_bpm = has_source_bpm ? bpm : 0.0f; }   // :178-180
  static float clamped(const std::vector<float>& c, long long i) { long long last = (long long)c.size() - 1; return c[(size_t)(i < 0 ? 0 : (i > last ? last : i))]; }
  static float wrapped(const std::vector<float>& c, long long i) { long long n = (long long)c.size(); long long r = i % n; if (r < 0) r += n; return c[(size_t)r]; }
  StereoFrame read_interpolated(double position) const {  // :198-226
    if (left.size() == 1) return {left[0], right[0]};
    double last = (double)(left.size() - 1);
    position = clampd(position, 0.0, last);
    long long index = (long long)std::floor(position);
    float frac = (float)(position - (double)index);
    auto rd = [&](const std::vector<float>& c) { return cubic_interpolate(clamped(c, index - 1), clamped(c, index), clamped(c, index + 1), clamped(c, index + 2), frac); };
    return {rd(left), rd(right)};
  }
  StereoFrame read_wrapped(double position) const {  // :228-262
    if (left.size() == 1) return {left[0], right[0]};
    double len = (double)left.size();
    position = rem_euclid(position, len);
    long long index = (long long)std::floor(position);
    float frac = (float)(position - (double)index);
    auto rd = [&](const std::vector<float>& c) { return cubic_interpolate(wrapped(c, index - 1), wrapped(c, index), wrapped(c, index + 1), wrapped(c, index + 2), frac); };
    return {rd(left), rd(right)};
  }
};

// ---- mixer/loop_channel.rs ---------------------------------------------------------------------------------------------
struct LoopWindow {  // :59-116
  double lo, hi, span; bool wraps; double len;
  double to_virtual(double p) const { return rem_euclid(p - lo, len); }
  double to_physical(double v) const { return rem_euclid(lo + v, len); }
  bool contains(double p) const { return wraps ? (p >= lo || p < hi) : (p >= lo && p < hi); }
  double fold(double p) const {
    if (contains(p)) return p;
    if (wraps) return (p - hi) <= (lo - p) ? hi : lo;
    return clampd(p, lo, hi);
  }
};
enum PitchMode { PITCH_OFF = 0, PITCH_RESAMPLE = 1, PITCH_PRESERVE = 2 };
struct LoopChannel {
  std::shared_ptr<StereoSampleBuffer> buffer;
  double cursor = 0.0;
  float loop_start = 0.0f, loop_end = 1.0f;
  bool playing = false;
  float speed = 1.0f;
  SmoothedParam gain, active_gain;
  bool muted = false, soloed = false;
  int pitch_mode = PITCH_OFF;
  float engine_bpm = 120.0f;
  explicit LoopChannel(float sr) : gain(1.0f, 0.0f, 2.0f, sr, 15.0f), active_gain(1.0f, 0.0f, 1.0f, sr, 15.0f) {}   // :157-178
  bool has_buffer() const { return buffer && buffer->len() > 0; }
  LoopWindow window(double len) const {  // :293-307
    double lo = clampd((double)loop_start * len, 0.0, len);
    double hi = clampd((double)loop_end * len, 0.0, len);
    bool wraps = hi < lo;
    double span = wraps ? len - lo + hi : hi - lo;
    return {lo, hi, span, wraps, len};
  }
  double warp_ratio() const {  // :282-291
    if (pitch_mode == PITCH_OFF) return 1.0;
    if (buffer && buffer->has_source_bpm && buffer->source_bpm > 0.0f && engine_bpm > 0.0f) return (double)engine_bpm / (double)buffer->source_bpm;
    return 1.0;
  }
  void advance(float engine_sr) {  // :233-279 (no queued swap)
    if (!buffer) return;
    double len = (double)buffer->len(), source_sr = (double)buffer->sample_rate;
    LoopWindow w = window(len);
    double span = w.span > 1.0 ? w.span : 1.0;
    double ratio = source_sr / (double)rust_max(engine_sr, 1.0f);
    double warp = pitch_mode == PITCH_RESAMPLE ? warp_ratio() : 1.0;
    double delta = (double)speed * ratio * warp;
    if (w.wraps) {
      double prev_v = w.to_virtual(cursor);
      double raw = prev_v + delta;
      double cur_v = rem_euclid(raw, span);
      cursor = w.to_physical(cur_v);
    } else {
      cursor += delta;
      if (cursor >= w.hi) cursor = w.lo + rem_euclid(cursor - w.lo, span);
      else if (cursor < w.lo) cursor = w.hi - rem_euclid(w.lo - cursor, span);
    }
  }
  StereoFrame tick(float engine_sr) {  // :181-208
    StereoFrame dry;
    if (playing && has_buffer()) {
      LoopWindow w = window((double)buffer->len());
      dry = w.wraps ? buffer->read_wrapped(cursor) : buffer->read_interpolated(cursor);
      advance(engine_sr);
    }
    StereoFrame gained = dry.scaled(gain.tick());
    return gained.scaled(active_gain.tick());          // empty EffectChain in between
  }
  void set_buffer(std::shared_ptr<StereoSampleBuffer> b) { double len = (double)b->len(); buffer = std::move(b); cursor = window(len).lo; }   // :311-316
  void set_gain(float g) { gain.set_target(clampf(g, 0.0f, 2.0f)); }
  void set_loop_start(float n) { loop_start = clampf(n, 0.0f, 1.0f); }
  void set_loop_end(float n) { loop_end = clampf(n, 0.0f, 1.0f); }
  void set_speed(float s) { speed = clampf(s, -4.0f, 4.0f); }
  void restart() { if (!buffer) return; cursor = window((double)buffer->len()).lo; }
  void set_position(float n) {  // :388-397
    if (!buffer) return;
    double len = (double)buffer->len();
    LoopWindow w = window(len);
    cursor = w.fold((double)clampf(n, 0.0f, 1.0f) * len);
  }
  float position_normalized() const { return (buffer && buffer->len() > 1) ? (float)(cursor / (double)buffer->len()) : 0.0f; }   // :497-502
  void prepare_offline_render() { playing = true; gain.snap(); active_gain.set_target(1.0f); active_gain.snap(); restart(); }      // :444-451
};

// ---- mixer/mod.rs -----------------------------------------------------------------------------------------------------
struct LoopMixer {
  std::vector<LoopChannel> channels;
  float sample_rate, bpm = 120.0f;
  explicit LoopMixer(float sr) : sample_rate(sr) { for (int i = 0; i < 4; i++) channels.emplace_back(sr); }
  StereoFrame tick(float engine_sr) {  // :60-75
    bool any_solo = false;
    for (auto& c : channels) any_solo |= c.soloed;
    StereoFrame out;
    for (auto& c : channels) {
      bool audible = any_solo ? c.soloed : !c.muted;
      c.active_gain.set_target(audible ? 1.0f : 0.0f);
      out += c.tick(engine_sr);
    }
    return out;
  }
  void set_bpm(float b) { bpm = b; for (auto& c : channels) c.engine_bpm = b; }
  LoopChannel* ch(size_t i) { return i < channels.size() ? &channels[i] : nullptr; }
  bool render_channel_to_interleaved(size_t channel, size_t frames, size_t preroll, std::vector<float>& out) {  // :444-476
    LoopChannel* c = ch(channel);
    if (!c || !c->has_buffer()) return false;
    c->prepare_offline_render();
    for (size_t i = 0; i < preroll; i++) c->tick(sample_rate);
    c->restart();
    out.clear();
    for (size_t i = 0; i < frames; i++) { StereoFrame f = c->tick(sample_rate); out.push_back(f.l); out.push_back(f.r); }
    return true;
  }
};

// ---- instruments/sampler.rs ------------------------------------------------------------------------------------------
struct SamplerBuffer {
  std::vector<float> samples; size_t frames = 0, channels = 0; float sample_rate = 0.0f;
  static std::shared_ptr<SamplerBuffer> from_interleaved(const float* s, size_t frames, size_t channels, float sr) {  // :25-50
    if (!(channels == 1 || channels == 2) || frames == 0 || !std::isfinite(sr) || sr <= 0.0f) return nullptr;
    for (size_t i = 0; i < frames * channels; i++) if (!std::isfinite(s[i])) return nullptr;
    auto b = std::make_shared<SamplerBuffer>();
    b->samples.assign(s, s + frames * channels); b->frames = frames; b->channels = channels; b->sample_rate = sr;
    return b;
  }
  StereoFrame frame(double position) const {  // :64-81
    position = clampd(position, 0.0, (double)(frames - 1));
    size_t i0 = (size_t)std::floor(position);
    size_t i1 = i0 + 1 < frames - 1 ? i0 + 1 : frames - 1;
    float frac = (float)(position - (double)i0);
    auto at = [&](size_t f, size_t c) { return samples[f * channels + c]; };
    auto lerp = [&](float a, float b) { return a + (b - a) * frac; };
    if (channels == 1) return StereoFrame::mono(lerp(at(i0, 0), at(i1, 0)));
    return {lerp(at(i0, 0), at(i1, 0)), lerp(at(i0, 1), at(i1, 1))};
  }
};
struct SampleVoice {  // :84-150
  std::shared_ptr<SamplerBuffer> buffer;
  size_t slot = 0; double position = 0.0, increment = 1.0; float velocity = 0.0f; uint64_t age = 0;
  bool active() const { return (bool)buffer; }
  void start(size_t s, std::shared_ptr<SamplerBuffer> b, float engine_rate, float vel, uint64_t a) {
    slot = s; position = 0.0; increment = (double)b->sample_rate / (double)engine_rate; velocity = clampf(vel, 0.0f, 1.0f); age = a; buffer = std::move(b);
  }
  StereoFrame tick() {
    if (!buffer) return {};
    StereoFrame f = buffer->frame(position);
    double fade = 32.0, end = (double)buffer->frames;
    double tail = (end - position) / fade;
    if (!(tail > 0.0)) tail = 0.0;                                   // .max(0.0)
    double g = position / fade;
    if (tail < g) g = tail;                                          // .min(...)
    if (1.0 < g) g = 1.0;                                            // .min(1.0)
    float gain = (float)g * velocity;
    position += increment;
    if (position >= end) buffer.reset();
    return f.scaled(gain);
  }
};
struct SamplerRack {
  float sample_rate;
  std::shared_ptr<SamplerBuffer> slots[16];
  SampleVoice voices[32];
  uint64_t next_age = 0;
  explicit SamplerRack(float sr) : sample_rate(sr) {}
  void stop_slot(size_t slot) { for (auto& v : voices) if (v.active() && v.slot == slot) v.buffer.reset(); }
  bool set_buffer(size_t slot, std::shared_ptr<SamplerBuffer> b) { if (slot >= 16) return false; slots[slot] = std::move(b); stop_slot(slot); return true; }   // :178-185
  bool clear_slot(size_t slot) { if (slot >= 16) return false; slots[slot].reset(); stop_slot(slot); return true; }
  bool trigger(size_t slot, float velocity) {  // :200-223
    if (slot >= 16 || !slots[slot]) return false;
    int vi = -1;
    for (int i = 0; i < 32; i++) if (!voices[i].active()) { vi = i; break; }
    if (vi < 0) { vi = 0; for (int i = 1; i < 32; i++) if (voices[i].age < voices[vi].age) vi = i; }   // min_by_key: first minimum
    next_age += 1;
    voices[vi].start(slot, slots[slot], sample_rate, velocity, next_age);
    return true;
  }
  StereoFrame tick() { StereoFrame out; for (auto& v : voices) out += v.tick(); return out; }   // :225-229
};

}  // namespace orc
