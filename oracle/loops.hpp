// oracle/loops.hpp — TEST INFRASTRUCTURE ONLY (tests/, smoke() and bench.py's cpu_baseline leg; never the product).
// CPU restatement of the reference's sample-playback sources:
//   mixer/stereo_buffer.rs (StereoSampleBuffer), mixer/loop_channel.rs (LoopWindow, LoopChannel), mixer/mod.rs (Mixer),
//   instruments/sampler.rs (SamplerBuffer, SampleVoice, SamplerRack).
//   mixer/wsola.rs (WsolaStretcher, PitchMode::PreservePitch).
// Restated: everything a bounce touches with empty per-channel effect chains.  Not restated (the product refuses the same
// requests): per-channel effect chains, the clip grid, transport-armed sampler patterns.  Parity unpinned like the rest of the oracle (no reference
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

// ---- mixer/wsola.rs ---------------------------------------------------------------------------------------------------
struct WsolaStretcher {
  size_t hop_len, window_len;
  std::vector<float> window;
  std::vector<StereoFrame> out_scratch, grain_scratch, prev_tail;
  std::vector<float> prev_tail_mono;
  bool have_prev_tail = false;
  size_t drain_idx;
  double analysis_cursor;
  WsolaStretcher(float engine_sr, double initial_cursor) {  // :70-98
    double sr = (double)rust_max(engine_sr, 1.0f);
    double h = std::round(((double)20.0f / 1000.0) * sr);
    hop_len = (size_t)(h > 1.0 ? h : 1.0);
    window_len = hop_len * 2;
    window.resize(window_len);
    // raised_sine_window(i / window_len, 2.0): `.powf(2.0)` with the literal exponent of the inlined call is lowered to x * x by
    // LLVM (pow(x, 2.0) -> x * x needs no fast-math), as GCC does for the same C expression; written out so no optimiser decides it.
    for (size_t i = 0; i < window_len; i++) {
      float sn = rust_max(sinf(3.14159265358979323846f * clampf((float)i / (float)window_len, 0.0f, 1.0f)), 0.0f);
      window[i] = sn * sn;
    }
    out_scratch.assign(hop_len, {}); grain_scratch.assign(window_len, {}); prev_tail.assign(hop_len, {}); prev_tail_mono.assign(hop_len, 0.0f);
    drain_idx = hop_len;
    analysis_cursor = initial_cursor;
  }
  bool needs_refill() const { return drain_idx >= hop_len; }
  StereoFrame drain() { StereoFrame f; if (drain_idx < out_scratch.size()) f = out_scratch[drain_idx]; drain_idx += 1; return f; }
  static double maxd(double a, double b) { return a > b ? a : b; }
  static double mind(double a, double b) { return a < b ? a : b; }
  // score_at of both searches (:330-347, :398-415): `pos_of(i)` yields the stereo read of tap i
  template <class Read> float score(Read read_at) const {
    float num = 0.0f, ref_energy = 0.0f, cand_energy = 0.0f;
    for (size_t i = 0; i < prev_tail_mono.size(); i++) {
      float reference = prev_tail_mono[i];
      StereoFrame raw = read_at(i);
      float cand = raw.l + raw.r;
      num += cand * reference;
      ref_energy += reference * reference;
      cand_energy += cand * cand;
    }
    const float EPS = 1.1920929e-7f;
    if (ref_energy <= EPS || cand_energy <= EPS) return 0.0f;
    return num / (sqrtf(ref_energy) * sqrtf(cand_energy));
  }
  template <class Score> static double coarse_to_fine(double lo_bound, double hi_bound, Score score_at) {  // :349-378, :417-446
    double span = hi_bound - lo_bound;
    double coarse_stride = maxd(span / 64.0, 1.0);
    double best = lo_bound; float best_score = -3.40282347e+38f;
    for (double c = lo_bound; c <= hi_bound; c += coarse_stride) { float sc = score_at(c); if (sc > best_score) { best_score = sc; best = c; } }
    double refine_lo = maxd(best - coarse_stride, lo_bound), refine_hi = mind(best + coarse_stride, hi_bound);
    for (double c = refine_lo; c <= refine_hi; c += 1.0) { float sc = score_at(c); if (sc > best_score) { best_score = sc; best = c; } }
    return best;
  }
  double search_best_start(const StereoSampleBuffer& b, double center, double step, double loop_lo, double max_start) const {  // :386-448
    double radius = maxd(std::round(((double)10.0f / 1000.0) * (double)b.sample_rate), 1.0);
    double lo_bound = maxd(center - radius, loop_lo), hi_bound = mind(center + radius, max_start);
    if (hi_bound <= lo_bound) return clampd(center, loop_lo, max_start);
    return coarse_to_fine(lo_bound, hi_bound, [&](double start) {
      return score([&](size_t i) { return b.read_interpolated(clampd(start + (double)i * step, loop_lo, max_start + step)); }); });
  }
  double search_best_start_wrapped(const StereoSampleBuffer& b, const LoopWindow& w, double center, double step, double max_start) const {  // :314-380
    double radius = maxd(std::round(((double)10.0f / 1000.0) * (double)b.sample_rate), 1.0);
    double lo_bound = maxd(center - radius, 0.0), hi_bound = mind(center + radius, max_start);
    if (hi_bound <= lo_bound) return clampd(center, 0.0, max_start);
    return coarse_to_fine(lo_bound, hi_bound, [&](double start) {
      return score([&](size_t i) { return b.read_wrapped(w.to_physical(clampd(start + (double)i * step, 0.0, max_start + step))); }); });
  }
  void overlap_add_and_carry() {  // :206-231, :284-307
    for (size_t i = 0; i < hop_len; i++) {
      StereoFrame prev = have_prev_tail ? prev_tail[i] : StereoFrame{};
      out_scratch[i] = {prev.l + grain_scratch[i].l, prev.r + grain_scratch[i].r};
    }
    for (size_t i = 0; i < hop_len; i++) { prev_tail[i] = grain_scratch[hop_len + i]; prev_tail_mono[i] = prev_tail[i].l + prev_tail[i].r; }
    have_prev_tail = true;
    drain_idx = 0;
  }
  double synthesize_next_hop(const StereoSampleBuffer& b, const LoopWindow& w, double sr_ratio, double speed, double warp) {  // :121-135
    double step = maxd(sr_ratio * maxd(speed, 0.0), 1e-6);
    double hop_source_span = (double)hop_len * step;
    double grain_source_span = ((double)window_len - 1.0) * step + 1.0;
    if (!w.wraps) {  // :140-236
      double loop_lo = w.lo, loop_hi = w.hi;
      double max_start = maxd(loop_hi - grain_source_span, loop_lo);
      double raw_target = analysis_cursor + hop_source_span * maxd(warp, 0.0);
      double search_center; bool wrapped;
      if (raw_target > max_start || max_start <= loop_lo) { search_center = loop_lo; wrapped = true; } else { search_center = maxd(raw_target, loop_lo); wrapped = false; }
      if (wrapped) have_prev_tail = false;
      double best_start = have_prev_tail ? search_best_start(b, search_center, step, loop_lo, max_start) : search_center;
      for (size_t i = 0; i < window_len; i++) {
        StereoFrame raw = b.read_interpolated(clampd(best_start + (double)i * step, loop_lo, loop_hi));
        grain_scratch[i] = {raw.l * window[i], raw.r * window[i]};
      }
      overlap_add_and_carry();
      analysis_cursor = best_start;
      return best_start;
    }
    // :246-311
    double span = w.span;
    double max_start = maxd(span - grain_source_span, 0.0);
    double cursor_v = w.to_virtual(analysis_cursor);
    double raw_target = cursor_v + hop_source_span * maxd(warp, 0.0);
    double search_center; bool wrapped;
    if (raw_target > max_start || max_start <= 0.0) { search_center = 0.0; wrapped = true; } else { search_center = maxd(raw_target, 0.0); wrapped = false; }
    if (wrapped) have_prev_tail = false;
    double best_start = have_prev_tail ? search_best_start_wrapped(b, w, search_center, step, max_start) : search_center;
    for (size_t i = 0; i < window_len; i++) {
      StereoFrame raw = b.read_wrapped(w.to_physical(clampd(best_start + (double)i * step, 0.0, span)));
      grain_scratch[i] = {raw.l * window[i], raw.r * window[i]};
    }
    overlap_add_and_carry();
    double phys = w.to_physical(best_start);
    analysis_cursor = phys;
    return phys;
  }
};
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
  std::shared_ptr<WsolaStretcher> stretcher;            // lazily built in PreservePitch mode, dropped whenever the cursor is moved from outside
  std::shared_ptr<StereoSampleBuffer> pending;          // queue_swap (:413-423): replaces `buffer` at the next grid boundary
  uint32_t pending_divisions = 1; bool has_pending = false; uint32_t swaps_completed = 0;
  void queue_swap(std::shared_ptr<StereoSampleBuffer> b, uint32_t divisions) { pending_divisions = divisions > 1 ? divisions : 1; pending = std::move(b); has_pending = true; }
  void cancel_queued_swap() { has_pending = false; pending.reset(); }
  void maybe_swap_pending(double prev_v, double cur_v, double span, bool wrapped) {  // :249-276
    if (!has_pending) return;
    double grid = (double)(pending_divisions > 1 ? pending_divisions : 1);
    double prev_idx = std::floor((prev_v / span) * grid), new_idx = std::floor((cur_v / span) * grid);
    if (!(wrapped || new_idx != prev_idx)) return;
    if (pending) {
      double new_lo = window((double)pending->len()).lo;
      buffer = std::move(pending); pending.reset();
      cursor = new_lo;
      stretcher.reset();
      swaps_completed += 1;
      has_pending = false;
    }
  }
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
  void advance(float engine_sr) {  // :233-279
    if (!buffer) return;
    double len = (double)buffer->len(), source_sr = (double)buffer->sample_rate;
    LoopWindow w = window(len);
    double span = w.span > 1.0 ? w.span : 1.0;
    double ratio = source_sr / (double)rust_max(engine_sr, 1.0f);
    double warp = pitch_mode == PITCH_RESAMPLE ? warp_ratio() : 1.0;
    double delta = (double)speed * ratio * warp;
    double prev = cursor, prev_v, cur_v; bool wrapped;
    if (w.wraps) {
      prev_v = w.to_virtual(prev);
      double raw = prev_v + delta;
      wrapped = !(0.0 <= raw && raw < span);
      cur_v = rem_euclid(raw, span);
      cursor = w.to_physical(cur_v);
    } else {
      cursor += delta;
      wrapped = false;
      if (cursor >= w.hi) { cursor = w.lo + rem_euclid(cursor - w.lo, span); wrapped = true; }
      else if (cursor < w.lo) { cursor = w.hi - rem_euclid(w.lo - cursor, span); wrapped = true; }
      prev_v = prev - w.lo; cur_v = cursor - w.lo;
    }
    maybe_swap_pending(prev_v, cur_v, span, wrapped);
  }
  StereoFrame tick(float engine_sr) {  // :181-208
    StereoFrame dry;
    if (playing && has_buffer()) {
      if (pitch_mode == PITCH_PRESERVE && speed >= 0.0f) dry = tick_preserve_pitch(engine_sr);
      else {
        LoopWindow w = window((double)buffer->len());
        dry = w.wraps ? buffer->read_wrapped(cursor) : buffer->read_interpolated(cursor);
        advance(engine_sr);
      }
    }
    StereoFrame gained = dry.scaled(gain.tick());
    return gained.scaled(active_gain.tick());          // empty EffectChain in between
  }
  StereoFrame tick_preserve_pitch(float engine_sr) {  // :215-262
    double len = (double)buffer->len();
    LoopWindow w = window(len);
    double sr_ratio = (double)buffer->sample_rate / (double)rust_max(engine_sr, 1.0f);
    double warp = warp_ratio(), sp = (double)speed;
    if (!stretcher) stretcher = std::make_shared<WsolaStretcher>(engine_sr, cursor);
    double prev = cursor; bool wrapped = false;
    if (stretcher->needs_refill()) {
      cursor = stretcher->synthesize_next_hop(*buffer, w, sr_ratio, sp, warp);
      wrapped = w.wraps ? w.to_virtual(cursor) < w.to_virtual(prev) : cursor < prev;
    }
    StereoFrame out = stretcher->drain();
    double span = w.span > 1.0 ? w.span : 1.0;
    double prev_v = w.wraps ? w.to_virtual(prev) : prev - w.lo, cur_v = w.wraps ? w.to_virtual(cursor) : cursor - w.lo;
    maybe_swap_pending(prev_v, cur_v, span, wrapped);        // hop granularity in this mode (:236-261)
    return out;
  }
  void set_pitch_mode(int m) { if (pitch_mode == PITCH_PRESERVE && m != PITCH_PRESERVE) stretcher.reset(); pitch_mode = m; }   // :341-346
  void set_buffer(std::shared_ptr<StereoSampleBuffer> b) { double len = (double)b->len(); buffer = std::move(b); cursor = window(len).lo; stretcher.reset(); }   // :311-316
  void set_gain(float g) { gain.set_target(clampf(g, 0.0f, 2.0f)); }
  void set_loop_start(float n) { loop_start = clampf(n, 0.0f, 1.0f); }
  void set_loop_end(float n) { loop_end = clampf(n, 0.0f, 1.0f); }
  void set_speed(float s) { speed = clampf(s, -4.0f, 4.0f); }
  void restart() { if (!buffer) return; cursor = window((double)buffer->len()).lo; stretcher.reset(); }
  void set_position(float n) {  // :388-397
    if (!buffer) return;
    double len = (double)buffer->len();
    LoopWindow w = window(len);
    cursor = w.fold((double)clampf(n, 0.0f, 1.0f) * len);
    stretcher.reset();
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
