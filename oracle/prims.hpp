// oracle/prims.hpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// CPU restatement of libgooey's DSP primitives, one class per reference
// struct, same evaluation order, plain scalar f32/f64, glibc libm (the libm
// rustc links on linux-gnu).  Compile with -O2 -ffp-contract=off.
// Every class cites the reference file:line it follows (paths relative to
// /root/reference/src).  Parity status: the reference cannot be built here
// (no Rust toolchain) and stores no golden audio, so the restatement is pinned
// against the reference's known-answer constants and metamorphic properties
// (tests/test_oracle_*.py); the half-band filters (third-party crate
// `halfband 0.2.0`, source absent) are a reconstruction — "parity unpinned".
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <array>
#include <algorithm>

namespace orc {

static inline float clampf(float x, float lo, float hi) {
  // Rust f32::clamp: NaN passes through.
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}
static inline float rust_max(float a, float b) { return fmaxf(a, b); }
static inline float rust_min(float a, float b) { return fminf(a, b); }
static inline float fract(float x) { return x - truncf(x); }
static inline uint64_t f32_as_u64(float x) {  // Rust `as u64`: truncate, saturate, NaN->0
  if (!(x == x)) return 0;
  if (x <= 0.0f) return 0;
  if (x >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)x;
}

// ---- utils/smoother.rs:13-137 ------------------------------------------------
struct SmoothedParam {
  float current = 0, target = 0, coeff = 1;
  bool settled = true;
  float min = 0, max = 1;
  static float calculate_coeff(float sr, float ms) {  // :69-77
    if (ms <= 0.0f) return 1.0f;
    float n = (ms / 1000.0f) * sr;
    return 1.0f - expf(-1.0f / n);
  }
  SmoothedParam() : SmoothedParam(0.0f, 0.0f, 1.0f, 44100.0f, 15.0f) {}
  SmoothedParam(float init, float mn, float mx, float sr, float ms) {  // :38-55
    coeff = calculate_coeff(sr, ms);
    float c = clampf(init, mn, mx);
    current = target = c;
    settled = true;
    min = mn;
    max = mx;
  }
  void set_target(float t) {  // :80-86
    float c = clampf(t, min, max);
    if (fabsf(target - c) > 1e-8f) { target = c; settled = false; }
  }
  void set_immediate(float v) { float c = clampf(v, min, max); current = target = c; settled = true; }
  void snap() { current = target; settled = true; }
  void set_normalized(float n) { set_target(min + clampf(n, 0.0f, 1.0f) * (max - min)); }
  void set_bipolar(float b) { set_normalized((clampf(b, -1.0f, 1.0f) + 1.0f) * 0.5f); }
  float tick() {  // :120-137
    if (settled) return current;
    current += coeff * (target - current);
    if (fabsf(current - target) < 1e-4f) { current = target; settled = true; }
    return current;
  }
  float get() const { return current; }
  bool is_settled() const { return settled; }
};

// ---- utils/mod.rs:14-44 ------------------------------------------------------
static inline float tuning_to_multiplier(float n) {
  float semis = (clampf(n, 0.0f, 1.0f) - 0.5f) * 24.0f;
  return powf(2.0f, semis / 12.0f);
}
static inline float cubic_interpolate(float p0, float p1, float p2, float p3, float t) {
  float a0 = -0.5f * p0 + 1.5f * p1 - 1.5f * p2 + 0.5f * p3;
  float a1 = p0 - 2.5f * p1 + 2.0f * p2 - 0.5f * p3;
  float a2 = -0.5f * p0 + 0.5f * p2;
  float a3 = p1;
  return ((a0 * t + a1) * t + a2) * t + a3;
}
static inline float raised_sine_window(float phase, float shape) {
  return powf(rust_max(sinf(3.14159265358979323846f * clampf(phase, 0.0f, 1.0f)), 0.0f), shape);
}

// ---- envelope.rs -------------------------------------------------------------
struct EnvelopeCurve {  // :6-26
  bool exponential = false;
  float c = 1.0f;
  static EnvelopeCurve Linear() { return {}; }
  static EnvelopeCurve Exponential(float c) { EnvelopeCurve e; e.exponential = true; e.c = c; return e; }
  float apply(float p) const { return exponential ? powf(p, clampf(c, 0.1f, 10.0f)) : p; }
};
struct ADSRConfig {  // :29-68
  float attack_time, decay_time, sustain_level, release_time;
  EnvelopeCurve attack_curve, decay_curve;
  ADSRConfig(float a, float d, float s, float r)
      : attack_time(rust_max(a, 0.001f)), decay_time(rust_max(d, 0.001f)),
        sustain_level(clampf(s, 0.0f, 1.0f)), release_time(rust_max(r, 0.001f)) {}
  ADSRConfig with_attack_curve(EnvelopeCurve c) const { ADSRConfig x = *this; x.attack_curve = c; return x; }
  ADSRConfig with_decay_curve(EnvelopeCurve c) const { ADSRConfig x = *this; x.decay_curve = c; return x; }
  // struct-literal construction (no floors), used by bass.rs
  static ADSRConfig raw(float a, float d, float s, float r, EnvelopeCurve ac, EnvelopeCurve dc) {
    ADSRConfig x(0, 0, 0, 0);
    x.attack_time = a; x.decay_time = d; x.sustain_level = s; x.release_time = r;
    x.attack_curve = ac; x.decay_curve = dc;
    return x;
  }
};
struct Envelope {  // :70-211
  float attack_time, decay_time, sustain_level, release_time;
  EnvelopeCurve attack_curve, decay_curve;
  float current_time = 0;
  bool is_active = false;
  double trigger_time = 0;
  bool has_release = false;
  double release_time_start = 0;
  Envelope() { set_config(ADSRConfig(0.01f, 0.3f, 0.7f, 0.5f)); }
  explicit Envelope(const ADSRConfig& c) { set_config(c); }
  void set_config(const ADSRConfig& c) {
    attack_time = c.attack_time; decay_time = c.decay_time; sustain_level = c.sustain_level;
    release_time = c.release_time; attack_curve = c.attack_curve; decay_curve = c.decay_curve;
  }
  void set_decay_time(float d) { decay_time = d; }
  void set_release_time(float r) { release_time = r; }
  void trigger(double t) { is_active = true; trigger_time = t; current_time = 0; has_release = false; }
  void release(double t) { if (is_active && !has_release) { has_release = true; release_time_start = t; } }
  float get_amplitude(double now) {  // :154-211
    if (!is_active) return 0.0f;
    float elapsed = (float)(now - trigger_time);
    current_time = elapsed;
    if (has_release) {
      float rel_el = (float)(now - release_time_start);
      if (rel_el < release_time) {
        float ra;
        if (elapsed < attack_time) {
          ra = attack_curve.apply(elapsed / attack_time);
        } else if (elapsed < attack_time + decay_time) {
          float de = elapsed - attack_time;
          float dp = de / decay_time;
          float cp = decay_curve.apply(dp);
          ra = 1.0f - (1.0f - sustain_level) * cp;
        } else {
          ra = sustain_level;
        }
        float rp = rel_el / release_time;
        return ra * (1.0f - rp);
      } else {
        is_active = false;
        return 0.0f;
      }
    } else {
      if (elapsed < attack_time) {
        return attack_curve.apply(elapsed / attack_time);
      } else if (elapsed < attack_time + decay_time) {
        float de = elapsed - attack_time;
        float dp = de / decay_time;
        float cp = decay_curve.apply(dp);
        return 1.0f - (1.0f - sustain_level) * cp;
      } else {
        if (sustain_level == 0.0f && !has_release) { has_release = true; release_time_start = now; }
        return sustain_level;
      }
    }
  }
};

// ---- max_curve.rs ------------------------------------------------------------
static inline float max_curve(float progress, float curve) {  // :21-48
  progress = clampf(progress, 0.0f, 1.0f);
  if (fabsf(curve) < 1e-6f) return progress;
  float hp = powf((fabsf(curve) + 1e-20f) * 1.2f, 0.41f) * 0.91f;
  float fp = hp / (1.0f - hp);
  if (fabsf(fp) < 1e-6f) return progress;
  float gp = expm1f(fp * progress) / expm1f(fp);
  if (curve < 0.0f) return 1.0f - max_curve(1.0f - progress, -curve);
  return gp;
}
struct EnvelopeSegment { float target_value, duration_secs, curve; };
struct MaxCurveEnvelope {  // :64-179
  std::vector<EnvelopeSegment> segments;
  size_t current_segment = 0;
  double segment_start_time = 0;
  float segment_start_value = 0, current_value = 0;
  bool is_active = false;
  double trigger_time = 0;
  float initial_value = 0;
  MaxCurveEnvelope() {}
  explicit MaxCurveEnvelope(std::initializer_list<std::array<float, 3>> segs) {
    for (auto& s : segs) segments.push_back({s[0], s[1] / 1000.0f, s[2]});
  }
  void set_initial_value(float v) { initial_value = v; }
  void set_segment_duration_ms(size_t i, float ms) {
    if (i < segments.size()) segments[i].duration_secs = rust_max(ms / 1000.0f, 0.0f);
  }
  void trigger(double t) {
    is_active = true; trigger_time = t; current_segment = 0; segment_start_time = t;
    segment_start_value = initial_value; current_value = initial_value;
  }
  float get_value(double now) {  // :133-174
    if (!is_active) return current_value;
    for (;;) {
      if (current_segment >= segments.size()) { is_active = false; return current_value; }
      const EnvelopeSegment& seg = segments[current_segment];
      float el = (float)(now - segment_start_time);
      if (el >= seg.duration_secs) {
        segment_start_value = seg.target_value;
        current_value = seg.target_value;
        segment_start_time += (double)seg.duration_secs;
        current_segment += 1;
        continue;
      }
      float progress = seg.duration_secs > 0.0f ? el / seg.duration_secs : 1.0f;
      float cp = max_curve(progress, seg.curve);
      float range = seg.target_value - segment_start_value;
      current_value = segment_start_value + range * cp;
      return current_value;
    }
  }
  bool is_complete() const { return !is_active && current_segment >= segments.size(); }
};

// ---- Rust std DefaultHasher = SipHash-1-3, k0=k1=0, of one u64 ----------------
// gen/oscillator.rs:187-196, gen/morph_osc.rs:42-47.
static inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
static inline uint64_t siphash_u64(uint64_t m, uint64_t k0 = 0, uint64_t k1 = 0, int c_rounds = 1, int d_rounds = 3) {
  uint64_t v0 = k0 ^ 0x736f6d6570736575ull, v1 = k1 ^ 0x646f72616e646f6dull;
  uint64_t v2 = k0 ^ 0x6c7967656e657261ull, v3 = k1 ^ 0x7465646279746573ull;
  auto round = [&]() {
    v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32);
    v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;
    v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;
    v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32);
  };
  v3 ^= m; for (int i = 0; i < c_rounds; i++) round(); v0 ^= m;
  uint64_t b = 8ull << 56;
  v3 ^= b; for (int i = 0; i < c_rounds; i++) round(); v0 ^= b;
  v2 ^= 0xff; for (int i = 0; i < d_rounds; i++) round();
  return v0 ^ v1 ^ v2 ^ v3;
}
static inline float hash_noise(uint64_t idx) {
  uint64_t h = siphash_u64(idx);
  float normalized = (float)h / 18446744073709551616.0f;  // (u64::MAX as f32) == 2^64
  return normalized * 2.0f - 1.0f;
}

// ---- gen/polyblep.rs:8-40 ------------------------------------------------------
static inline double poly_blep(double t, double dt) {
  if (t < dt) { t = t / dt; return 2.0 * t - t * t - 1.0; }
  else if (t > 1.0 - dt) { t = (t - 1.0) / dt; return t * t + 2.0 * t + 1.0; }
  return 0.0;
}
static inline float polyblep_saw(double phase, double inc) {
  double naive = 2.0 * phase - 1.0;
  return (float)(naive - poly_blep(phase, inc));
}
static inline float polyblep_square(double phase, double inc) {
  double naive = phase < 0.5 ? 1.0 : -1.0;
  double b1 = poly_blep(phase, inc);
  double p2 = fmod(phase + 0.5, 1.0);
  double b2 = poly_blep(p2, inc);
  return (float)(naive + b1 - b2);
}

// ---- gen/oscillator.rs -----------------------------------------------------------
enum class Waveform { Sine, Square, Saw, Triangle, RingMod, Noise };
struct Oscillator {
  float sample_rate;
  Waveform waveform = Waveform::Square;
  float current_sample_index = 0;
  float frequency_hz;
  Envelope envelope;
  float volume = 1.0f;
  float modulator_frequency_hz;
  bool enabled = true, antialias = true;
  Oscillator(float sr, float f) : sample_rate(sr), frequency_hz(f), modulator_frequency_hz(f * 0.5f) {}
  float sine_from_freq(float freq) const {  // :42-46
    float two_pi = 2.0f * 3.14159265358979323846f;
    return sinf(current_sample_index * freq * two_pi / sample_rate);
  }
  bool above_nyquist(float mult) const { return frequency_hz * mult > sample_rate / 2.0f; }
  float generative(int inc, float gain_exp) const {  // :106-131
    float output = 0.0f;
    int i = 1;
    float nyquist = sample_rate / 2.0f;
    float q = nyquist / frequency_hz;
    int max_h;  // Rust `as i32`: saturating, NaN -> 0
    if (!(q == q)) max_h = 0; else if (q >= 2147483648.0f) max_h = INT32_MAX; else if (q <= -2147483648.0f) max_h = INT32_MIN; else max_h = (int)q;
    while (i <= max_h && !above_nyquist((float)i)) {
      float gain = 1.0f / powf((float)i, gain_exp);
      float hf = frequency_hz * (float)i;
      float ratio = hf / nyquist;
      float taper;
      if (ratio > 0.75f) { float t = (ratio - 0.75f) / 0.25f; taper = 1.0f - t * t; } else taper = 1.0f;
      output += gain * taper * sine_from_freq(hf);
      i += inc;
    }
    return output;
  }
  void polyblep_phase(double& phase, double& inc) const {
    inc = (double)frequency_hz / (double)sample_rate;
    phase = fmod((double)current_sample_index * inc, 1.0);
  }
  float noise() const { return hash_noise(f32_as_u64(current_sample_index)); }
  void trigger(double t) { envelope.trigger(t); current_sample_index = 0.0f; }
  void release(double t) { envelope.release(t); }
  void set_volume(float v) { volume = clampf(v, 0.0f, 1.0f); }
  void set_adsr(const ADSRConfig& c) { envelope.set_config(c); }
  float tick(double now) {  // :242-286
    if (!enabled) return 0.0f;
    float el = envelope.is_active ? (float)(now - envelope.trigger_time) : 0.0f;
    current_sample_index = el * sample_rate;
    float raw = 0;
    double ph, inc;
    switch (waveform) {
      case Waveform::Sine: raw = sine_from_freq(frequency_hz); break;
      case Waveform::Square:
        polyblep_phase(ph, inc);
        raw = antialias ? polyblep_square(ph, inc) : (ph < 0.5 ? 1.0f : -1.0f);
        break;
      case Waveform::Saw:
        polyblep_phase(ph, inc);
        raw = antialias ? polyblep_saw(ph, inc) : (float)(2.0 * ph - 1.0);
        break;
      case Waveform::Triangle:
        if (antialias) raw = generative(2, 2.0f);
        else { polyblep_phase(ph, inc); float p = (float)ph; raw = p < 0.5f ? 4.0f * p - 1.0f : 3.0f - 4.0f * p; }
        break;
      case Waveform::RingMod: raw = sine_from_freq(frequency_hz) * sine_from_freq(modulator_frequency_hz); break;
      case Waveform::Noise: raw = noise(); break;
    }
    float amp = envelope.get_amplitude(now);
    return raw * amp * volume;
  }
};

// ---- gen/pink_noise.rs -------------------------------------------------------------
struct PinkNoise {
  uint64_t rng_state = 0x123456789abcdef0ull;
  float filter_state[3] = {0, 0, 0};
  float poles[3], gains[3];
  explicit PinkNoise(float sr) {  // :24-46
    const float RP[3] = {0.99765f, 0.96300f, 0.57000f};
    const float RG[3] = {0.0990460f, 0.2965164f, 1.0526913f};
    sr = rust_max(sr, 1.0f);
    float ratio = 44100.0f / sr;
    for (int i = 0; i < 3; i++) {
      poles[i] = powf(RP[i], ratio);
      gains[i] = RG[i] * sqrtf((1.0f - poles[i] * poles[i]) / (1.0f - RP[i] * RP[i]));
    }
  }
  void reset() { rng_state = 0x123456789abcdef0ull; filter_state[0] = filter_state[1] = filter_state[2] = 0; }
  float next_white() {  // :66-79
    uint64_t x = rng_state;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    rng_state = x;
    uint64_t h = x * 0x2545f4914f6cdd1dull;
    float n = (float)(h >> 40) / (float)((1u << 24) - 1);
    return n * 2.0f - 1.0f;
  }
  float tick() {  // :56-64
    float w = next_white();
    for (int i = 0; i < 3; i++) filter_state[i] = poles[i] * filter_state[i] + gains[i] * w;
    float s = 0.0f;  // iter().sum::<f32>() folds from 0.0
    for (int i = 0; i < 3; i++) s += filter_state[i];
    return (s + w * 0.1848f) * 0.11f;
  }
};

// ---- gen/click_osc.rs ----------------------------------------------------------------
static const float TOM_IMPULSE[64] = {
    0.884058f, 0.942029f, 0.913043f, 0.869565f, 0.833333f, 0.797101f, 0.772947f, 0.748792f, 0.724638f,
    0.695652f, 0.666667f, 0.637681f, 0.619565f, 0.601449f, 0.583333f, 0.565217f, 0.536232f, 0.507246f,
    0.478261f, 0.449275f, 0.42029f,  0.391304f, 0.371981f, 0.352657f, 0.333333f, 0.304348f, 0.275362f,
    0.23913f,  0.202899f, 0.181159f, 0.15942f,  0.137681f, 0.115942f, 0.101449f, 0.086957f, 0.072464f,
    0.057971f, 0.043478f, 0.028986f, 0.014493f, 0.009662f, 0.004831f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f,
    0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.014493f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
struct ClickOsc {
  size_t position = 0;
  bool is_playing = false;
  void trigger() { position = 0; is_playing = true; }
  float tick() {  // :58-78
    if (!is_playing) return 0.0f;
    if (position >= 64) { is_playing = false; return 0.0f; }
    float s = TOM_IMPULSE[position];
    position += 1;
    if (position >= 64) is_playing = false;
    return s;
  }
};

// ---- gen/morph_osc.rs --------------------------------------------------------------------
struct MorphOsc {
  float sample_rate;
  float main_sine_phase = 0, tri_phase = 0, fixed_sine_phase = 0;
  uint64_t noise_counter = 0;
  float rand_phase = 0, rand_current = 0, rand_target = 0, gated_sine_phase = 0;
  explicit MorphOsc(float sr) : sample_rate(sr) {}
  void reset() { main_sine_phase = tri_phase = fixed_sine_phase = 0; noise_counter = 0; rand_phase = rand_current = rand_target = gated_sine_phase = 0; }
  static float sine(float p) { return sinf(p * 2.0f * 3.14159265358979323846f); }
  static float triangle(float p) { float t = fract(p); return t < 0.5f ? 4.0f * t - 1.0f : 3.0f - 4.0f * t; }
  static float mtof(float m) { return 440.0f * powf(2.0f, (m - 69.0f) / 12.0f); }
  static void advance(float& ph, float f, float sr) { ph += f / sr; if (ph >= 1.0f) ph -= 1.0f; }
  static float mix3(float c, float a, float b, float d) {
    float w1 = clampf(-c, 0.0f, 1.0f), w2 = clampf(1.0f - fabsf(c), 0.0f, 1.0f), w3 = clampf(c, 0.0f, 1.0f);
    return a * w1 + b * w2 + d * w3;
  }
  float tick(float frequency, float mix_control, float color_freq, float tone) {  // :137-202
    float main_sine = sine(main_sine_phase) * 0.5f;
    advance(main_sine_phase, frequency, sample_rate);
    float tri = triangle(tri_phase) * 0.5f;
    advance(tri_phase, frequency, sample_rate);
    float fixed_sine = sine(fixed_sine_phase) * 0.5f;
    advance(fixed_sine_phase, 190.0f, sample_rate);
    noise_counter += 1;
    float noise = hash_noise(noise_counter) * 0.2f;
    float rand_freq = mtof(color_freq);
    float prev = rand_phase;
    advance(rand_phase, rand_freq, sample_rate);
    if (rand_phase < prev) { rand_current = rand_target; rand_target = hash_noise(noise_counter + 0x12345678ull); }
    float rand_value = rand_current + (rand_target - rand_current) * rand_phase;
    float noise_combined = (noise + rand_value) * 0.4f;
    float gated = tone < 99.0f ? sine(gated_sine_phase) * 0.2f : 0.0f;
    advance(gated_sine_phase, frequency, sample_rate);
    float ch1 = main_sine * fixed_sine;
    float ch2 = tri + noise_combined;
    float ch3 = noise_combined + gated;
    return mix3(mix_control, ch1, ch2, ch3);
  }
};

// ---- filters -----------------------------------------------------------------------------
static const float PI_F = 3.14159265358979323846f;

struct StateVariableFilter {  // filters/state_variable.rs (Chamberlin, 2x)
  float sample_rate, cutoff_freq, resonance, low = 0, band = 0, f = 0, q = 0;
  StateVariableFilter(float sr, float c, float r) : sample_rate(sr), cutoff_freq(clampf(c, 20.0f, 20000.0f)), resonance(rust_max(r, 0.5f)) { update(); }
  void reset() { low = band = 0; }
  void update() {  // :53-60
    float nf = rust_min(cutoff_freq / sample_rate, 0.45f);
    f = 2.0f * sinf(PI_F * nf);
    q = 1.0f / resonance;
  }
  void process_all(float in, float& lo, float& bd, float& hi) {  // :78-90
    float high = 0.0f;
    for (int i = 0; i < 2; i++) {
      low = low + f * band;
      high = in - low - q * band;
      band = f * high + band;
    }
    lo = low; bd = band; hi = high;
  }
  float process_mode(float in, uint8_t type) {
    float lo, bd, hi;
    process_all(in, lo, bd, hi);
    switch (type) { case 0: return lo; case 1: return bd; case 2: return hi; case 3: return lo + hi; default: return bd; }
  }
  void set_params(float c, float r) { cutoff_freq = clampf(c, 20.0f, 20000.0f); resonance = rust_max(r, 0.5f); update(); }  // :131-135
};

struct StateVariableFilterTpt {  // filters/state_variable_tpt.rs
  float sample_rate, cutoff_freq, resonance, g = 0, r = 0, h = 0, ic1eq = 0, ic2eq = 0;
  StateVariableFilterTpt(float sr, float c, float res) : sample_rate(sr), cutoff_freq(clampf(c, 20.0f, 20000.0f)), resonance(rust_max(res, 0.5f)) { update(); }
  void reset() { ic1eq = ic2eq = 0; }
  void update() {  // :42-53
    float cutoff = clampf(cutoff_freq, 20.0f, sample_rate * 0.45f);
    float q = rust_max(resonance, 0.5f);
    float gg = tanf(PI_F * cutoff / sample_rate);
    float rr = 1.0f / q;
    float hh = 1.0f / (1.0f + rr * gg + gg * gg);
    g = gg; r = rr; h = hh;
  }
  void process_all(float in, float& lo, float& bd, float& hi) {  // :56-69
    float v1 = (g * (in - ic2eq) + ic1eq) * h;
    float v2 = ic2eq + g * v1;
    ic1eq = 2.0f * v1 - ic1eq;
    ic2eq = 2.0f * v2 - ic2eq;
    lo = v2; bd = v1; hi = in - (r * v1 + v2);
  }
  float process_mode(float in, uint8_t type) {
    float lo, bd, hi;
    process_all(in, lo, bd, hi);
    switch (type) { case 0: return lo; case 1: return bd; case 2: return hi; case 3: return lo + hi; default: return bd; }
  }
  void set_params(float c, float res) {  // :83-92
    float nc = clampf(c, 20.0f, sample_rate * 0.45f);
    float nr = rust_max(res, 0.5f);
    if (fabsf(nc - cutoff_freq) > 0.001f || fabsf(nr - resonance) > 0.001f) { cutoff_freq = nc; resonance = nr; update(); }
  }
};

struct ResonantLowpassFilter {  // filters/resonant_lowpass.rs
  float sample_rate, cutoff_freq, resonance, g = 0, r = 0, h = 0, ic1eq = 0, ic2eq = 0;
  ResonantLowpassFilter(float sr, float c, float res) : sample_rate(sr), cutoff_freq(clampf(c, 20.0f, 20000.0f)), resonance(clampf(res, 0.5f, 10.0f)) { update(); }
  void reset() { ic1eq = ic2eq = 0; }
  float process(float in) {  // :49-61
    float v1 = (g * (in - ic2eq) + ic1eq) * h;
    float v2 = ic2eq + g * v1;
    ic1eq = 2.0f * v1 - ic1eq;
    ic2eq = 2.0f * v2 - ic2eq;
    return fabsf(v2) < 1e-15f ? 0.0f : v2;
  }
  void set_params(float c, float res) {  // :79-91
    c = clampf(c, 20.0f, 20000.0f);
    res = clampf(res, 0.5f, 10.0f);
    if (fabsf(c - cutoff_freq) > 0.001f || fabsf(res - resonance) > 0.001f) { cutoff_freq = c; resonance = res; update(); }
  }
  void update() {  // :93-101
    float sr = rust_max(sample_rate, 1.0f);
    float cutoff = clampf(cutoff_freq, 20.0f, sr * 0.45f);
    float q = clampf(resonance, 0.5f, 10.0f);
    g = tanf(PI_F * cutoff / sr);
    r = 1.0f / q;
    h = 1.0f / (1.0f + r * g + g * g);
  }
};

struct ResonantHighpassFilter {  // filters/resonant_highpass.rs:22-54
  float sample_rate, cutoff_freq, resonance, filter_state = 0;
  ResonantHighpassFilter(float sr, float c, float r) : sample_rate(sr), cutoff_freq(c), resonance(r) {}
  void reset() { filter_state = 0; }
  float process(float in) {
    float alpha = 1.0f - expf(-2.0f * PI_F * cutoff_freq / sample_rate);
    float hp = in - filter_state;
    filter_state += alpha * hp;
    return hp * (1.0f + resonance * 0.1f);
  }
};

struct BiquadBandpass {  // filters/biquad_bandpass.rs
  float sample_rate, b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0, x1 = 0, x2 = 0, y1 = 0, y2 = 0;
  float last_freq = -1, last_q = -1, last_gain = -1;
  explicit BiquadBandpass(float sr) : sample_rate(sr) { set_params(1000.0f, 1.0f, 1.0f); }
  void reset() { x1 = x2 = y1 = y2 = 0; }
  void set_params(float freq, float q, float gain) {  // :73-87
    if (fabsf(freq - last_freq) < 0.01f && fabsf(q - last_q) < 0.001f && fabsf(gain - last_gain) < 0.001f) return;
    last_freq = freq; last_q = q; last_gain = gain;
    calc(freq, q, gain);
  }
  void calc(float freq, float q, float gain) {  // :89-119
    float nyq = sample_rate * 0.5f;
    freq = clampf(freq, 20.0f, nyq * 0.95f);
    q = clampf(q, 0.1f, 100.0f);
    float w0 = 2.0f * PI_F * freq / sample_rate;
    float sn = sinf(w0), cs = cosf(w0);
    float alpha = sn / (2.0f * q);
    float B0 = q * alpha * gain, B1 = 0.0f, B2 = -q * alpha * gain;
    float A0 = 1.0f + alpha, A1 = -2.0f * cs, A2 = 1.0f - alpha;
    b0 = B0 / A0; b1 = B1 / A0; b2 = B2 / A0; a1 = A1 / A0; a2 = A2 / A0;
  }
  float process(float in) {  // :122-145
    float out = b0 * in + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
    x2 = x1; x1 = in; y2 = y1; y1 = out;
    if (fabsf(out) < 1e-15f) return 0.0f;
    return out;
  }
};

struct BiquadHighpass {  // filters/biquad_highpass.rs
  float sample_rate, b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0, x1 = 0, x2 = 0, y1 = 0, y2 = 0;
  float last_freq = -1, last_q = -1;
  explicit BiquadHighpass(float sr) : sample_rate(sr) { set_params(1000.0f, 1.0f); }
  void reset() { x1 = x2 = y1 = y2 = 0; }
  void set_params(float freq, float q) {  // :68-77
    if (fabsf(freq - last_freq) < 0.01f && fabsf(q - last_q) < 0.001f) return;
    last_freq = freq; last_q = q;
    float nyq = sample_rate * 0.5f;
    freq = clampf(freq, 20.0f, nyq * 0.95f);
    q = clampf(q, 0.1f, 100.0f);
    float w0 = 2.0f * PI_F * freq / sample_rate;
    float sn = sinf(w0), cs = cosf(w0);
    float alpha = sn / (2.0f * q);
    float B0 = (1.0f + cs) / 2.0f, B1 = -(1.0f + cs), B2 = (1.0f + cs) / 2.0f;
    float A0 = 1.0f + alpha, A1 = -2.0f * cs, A2 = 1.0f - alpha;
    b0 = B0 / A0; b1 = B1 / A0; b2 = B2 / A0; a1 = A1 / A0; a2 = A2 / A0;
  }
  float process(float in) {  // :98-110
    float out = b0 * in + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
    x2 = x1; x1 = in; y2 = y1; y1 = out;
    if (fabsf(out) < 1e-15f) return 0.0f;
    return out;
  }
};

struct MembraneResonator {  // filters/membrane_resonator.rs
  std::vector<BiquadBandpass> filters;
  float params[5][3] = {{275.0f, 165.0f, 376.0f}, {220.0f, 228.0f, 205.0f}, {79.0f, 294.0f, 143.0f}, {65.0f, 320.0f, 129.0f}, {57.0f, 326.0f, 141.0f}};
  float q_scale = 0.01f, gain_scale = 0.0031f, ring_level = 0;
  explicit MembraneResonator(float sr) { for (int i = 0; i < 5; i++) filters.emplace_back(sr); update_filters(); }
  void reset() { for (auto& f : filters) f.reset(); ring_level = 0; }
  void update_filters() {  // :87-93
    for (int i = 0; i < 5; i++) {
      float sq = clampf(params[i][2] * q_scale, 0.1f, 100.0f);
      float sg = params[i][0] * gain_scale;
      filters[i].set_params(params[i][1], sq, sg);
    }
  }
  void set_q_scale(float s) { q_scale = clampf(s, 0.001f, 1.0f); update_filters(); }
  void set_gain_scale(float s) { gain_scale = clampf(s, 0.0001f, 0.1f); update_filters(); }
  bool is_ringing() const { return ring_level > 0.0001f; }
  float fade_multiplier() const {  // :171-185
    if (ring_level >= 0.005f) return 1.0f;
    if (ring_level <= 0.0001f) return 0.0f;
    return (ring_level - 0.0001f) / (0.005f - 0.0001f);
  }
  float process(float in) {  // :189-200
    float out = 0.0f;
    for (auto& f : filters) out += f.process(in);
    float clipped = tanhf(out);
    ring_level = ring_level * 0.999f + fabsf(clipped) * 0.001f;
    return clipped;
  }
};

// ---- halfband 0.2.0 (third-party, source absent — RECONSTRUCTION, parity unpinned) ---------
// utils/oversampler.rs:1 uses halfband::iir::{Upsampler8, Downsampler8}.  Restated as
// the HIIR polyphase-allpass pair the crate is understood to port: 8 coefficients
// designed by de Soras' elliptic method for a transition bandwidth giving the
// "94 dB" the reference quotes (oversampler.rs:37); even coefficients on path 0,
// odd on path 1; each section y[n] = c*(x[n]-y[n-1]) + x[n-1].
struct HalfbandDesign {
  float c[8];
  static double ipowp(double x, long n) { double r = 1.0; while (n > 0) { if (n & 1) r *= x; x *= x; n >>= 1; } return r; }
  HalfbandDesign() {
    const double transition = 0.0343747;
    const int n = 8, order = n * 2 + 1;
    double k = tan((1.0 - transition * 2.0) * M_PI / 4.0); k *= k;
    double kk = pow(1.0 - k * k, 0.25);
    double e = 0.5 * (1.0 - kk) / (1.0 + kk), e2 = e * e, e4 = e2 * e2;
    double q = e * (1.0 + e4 * (2.0 + e4 * (15.0 + 150.0 * e4)));
    for (int idx = 0; idx < n; idx++) {
      int cc = idx + 1;
      double num = 0, den = 0, v;
      int i = 0, j = 1;
      do { v = ipowp(q, (long)i * (i + 1)) * sin((i * 2 + 1) * cc * M_PI / order) * j; num += v; j = -j; i++; } while (fabs(v) > 1e-100);
      num *= pow(q, 0.25);
      i = 1; j = -1;
      do { v = ipowp(q, (long)i * i) * cos(i * 2 * cc * M_PI / order) * j; den += v; j = -j; i++; } while (fabs(v) > 1e-100);
      den += 0.5;
      double ww = num / den; ww *= ww;
      double x = sqrt((1.0 - ww * k) * (1.0 - ww / k)) / (1.0 + ww);
      c[idx] = (float)((1.0 - x) / (1.0 + x));
    }
  }
};
static inline const float* halfband_coefs() { static HalfbandDesign d; return d.c; }

struct Halfband8 {  // shared section bank: x[i] = previous input of section i, y[i] = previous output
  float x[8] = {0}, y[8] = {0};
  void clear() { for (int i = 0; i < 8; i++) x[i] = y[i] = 0; }
  inline void run(float& p0, float& p1) {
    const float* c = halfband_coefs();
    for (int i = 0; i < 8; i += 2) {
      float t0 = (p0 - y[i]) * c[i] + x[i];
      float t1 = (p1 - y[i + 1]) * c[i + 1] + x[i + 1];
      x[i] = p0; x[i + 1] = p1; y[i] = t0; y[i + 1] = t1;
      p0 = t0; p1 = t1;
    }
  }
};
struct Upsampler8 { Halfband8 s; void clear() { s.clear(); } void process(float in, float& o0, float& o1) { float a = in, b = in; s.run(a, b); o0 = a; o1 = b; } };
struct Downsampler8 { Halfband8 s; void clear() { s.clear(); } float process(float i0, float i1) { float a = i1, b = i0; s.run(a, b); return 0.5f * (a + b); } };

enum class OversamplingMode : uint8_t { Off = 0, X2 = 2, X4 = 4 };
struct Oversampler {  // utils/oversampler.rs:38-175
  OversamplingMode mode = OversamplingMode::X4;
  Upsampler8 x2_up; Downsampler8 x2_down;
  Upsampler8 outer_up, inner_up; Downsampler8 inner_down, outer_down;
  template <class F> float process(float in, F f) {
    if (mode == OversamplingMode::Off) return f(in);
    if (mode == OversamplingMode::X2) { float s0, s1; x2_up.process(in, s0, s1); float a = f(s0), b = f(s1); return x2_down.process(a, b); }
    float o0, o1, i0, i1, i2, i3;
    outer_up.process(in, o0, o1);
    inner_up.process(o0, i0, i1);
    float fa = f(i0), fb = f(i1);
    float d0 = inner_down.process(fa, fb);
    inner_up.process(o1, i2, i3);
    float fc = f(i2), fd = f(i3);
    float d1 = inner_down.process(fc, fd);
    return outer_down.process(d0, d1);
  }
  void reset() { x2_up.clear(); x2_down.clear(); outer_up.clear(); inner_up.clear(); inner_down.clear(); outer_down.clear(); }
  void set_mode(OversamplingMode m) { if (mode != m) { mode = m; reset(); } }
};

// ---- effects/waveshaper.rs:48-72 ---------------------------------------------------------------
struct Waveshaper {
  float drive, mix;
  Oversampler oversampler;
  Waveshaper(float d, float m) : drive(clampf(d, 1.0f, 10.0f)), mix(clampf(m, 0.0f, 1.0f)) {}
  void set_drive(float d) { drive = clampf(d, 1.0f, 10.0f); }
  void set_mix(float m) { mix = clampf(m, 0.0f, 1.0f); }
  void reset() { oversampler.reset(); }
  float process(float in) {
    if (!std::isfinite(in)) { reset(); return 0.0f; }
    if (mix <= 0.0001f || drive <= 1.0f) return in;
    float d = drive;
    float reference = 0.5f;
    float comp = tanhf(reference) / tanhf(reference * d);
    float sat = oversampler.process(in, [&](float x) { return tanhf(x * d) * comp; });
    return in * (1.0f - mix) + sat * mix;
  }
};

// ---- effects/feedback_waveshaper.rs ---------------------------------------------------------------
struct FeedbackWaveshaper {
  float drive, mix, feedback, sample_rate, filter_cutoff, filter_coeff, env_att_coeff, env_rel_coeff;
  float last_out = 0, filter_state = 0, dc_x1 = 0, dc_y1 = 0, env = 0;
  Oversampler oversampler;
  static float compute_filter_coeff(float c, float sr) { float g = 1.0f - expf(-2.0f * PI_F * c / sr); return clampf(g, 0.0f, 0.9f); }
  static float compute_env_coeff(float ms, float sr) { return expf(-1.0f / (ms / 1000.0f * sr)); }
  FeedbackWaveshaper(float sr, float d, float fb, float cutoff, float m)
      : drive(clampf(d, 1.0f, 100.0f)), mix(clampf(m, 0.0f, 1.0f)), feedback(clampf(fb, 0.0f, 0.98f)), sample_rate(sr),
        filter_cutoff(clampf(cutoff, 200.0f, 20000.0f)) {
    filter_coeff = compute_filter_coeff(filter_cutoff, sr);
    env_att_coeff = compute_env_coeff(1.0f, sr);
    env_rel_coeff = compute_env_coeff(120.0f, sr);
  }
  void reset() { last_out = filter_state = dc_x1 = dc_y1 = env = 0; oversampler.reset(); }
  void set_drive(float d) { drive = clampf(d, 1.0f, 100.0f); }
  void set_feedback(float f) { feedback = clampf(f, 0.0f, 0.98f); }
  void set_filter_cutoff(float c) { filter_cutoff = clampf(c, 200.0f, 20000.0f); filter_coeff = compute_filter_coeff(filter_cutoff, sample_rate); }
  void set_mix(float m) { mix = clampf(m, 0.0f, 1.0f); }
  static float gain_compensation(float env, float drive, float feedback) {  // :247-259
    float reference = rust_max(env, 0.05f);
    float driven = rust_max(fabsf(tanhf(reference * drive)), 1e-6f);
    float comp_no_fb = tanhf(reference) / driven;
    float drive_norm = clampf((drive - 1.0f) / 99.0f, 0.0f, 1.0f);
    float fb_norm = clampf(feedback / 0.98f, 0.0f, 1.0f);
    float high_end = powf(drive_norm, 1.35f) * powf(fb_norm, 2.0f);
    float makeup = powf(10.0f, 5.1f * high_end / 20.0f);
    float taming = 1.0f / (1.0f + comp_no_fb * feedback * 0.25f);
    return rust_min(comp_no_fb * taming * makeup, 3.0f);
  }
  float process(float in) {  // :109-169
    if (!std::isfinite(in)) { reset(); return 0.0f; }
    if (mix <= 0.0001f || drive <= 1.0f) return in;
    float fb_in = drive * in + feedback * last_out;
    float shaped = oversampler.process(fb_in, [](float x) { return tanhf(x); });
    float rect = fabsf(in);
    float coeff = rect > env ? env_att_coeff : env_rel_coeff;
    env += (1.0f - coeff) * (rect - env);
    if (fabsf(env) < 1e-15f) env = 0.0f;
    float comp = gain_compensation(env, drive, feedback);
    float compensated = shaped * comp;
    float out = compensated - dc_x1 + 0.995f * dc_y1;  // dc_block :262-271
    dc_x1 = compensated;
    dc_y1 = fabsf(out) < 1e-15f ? 0.0f : out;
    float dc_blocked = out;
    filter_state += filter_coeff * (dc_blocked - filter_state);
    if (fabsf(filter_state) < 1e-15f) filter_state = 0.0f;
    last_out = filter_state;
    if (!std::isfinite(last_out) || fabsf(last_out) > 50.0f) { reset(); return in; }
    return in * (1.0f - mix) + dc_blocked * mix;
  }
};

// ---- instruments/fm_snap.rs:102-169 -----------------------------------------------------------------
struct PhaseModulator {
  float attack_time = 0.001f, decay_time = 0.005f, attack_curve = 0.3f, decay_curve = 0.4f;
  double trigger_time = 0;
  bool is_active = false;
  void trigger(double t) { trigger_time = t; is_active = true; }
  float tick(double now) {
    if (!is_active) return 0.0f;
    float el = (float)(now - trigger_time);
    float total = attack_time + decay_time;
    if (el > total) { is_active = false; return 0.0f; }
    if (el < attack_time) return powf(el / attack_time, attack_curve);
    float de = el - attack_time;
    return 1.0f - powf(de / decay_time, decay_curve);
  }
};

// ---- effects/limiter.rs:44-78 ---------------------------------------------------------------------------
struct SoftLimiter {
  float threshold, inv_threshold;
  explicit SoftLimiter(float t) { threshold = rust_max(t, 0.001f); inv_threshold = 1.0f / threshold; }
  void set_threshold(float t) { if (!std::isfinite(t)) return; threshold = clampf(t, 0.001f, 1.0f); inv_threshold = 1.0f / threshold; }
  float process(float x) const { return tanhf(x * inv_threshold) * threshold; }
};

// ---- frame.rs:13-53 --------------------------------------------------------------------------------------
struct StereoFrame {
  float l = 0, r = 0;
  static StereoFrame mono(float x) { return {x, x}; }
  static StereoFrame panned(float x, float pan) {
    float angle = clampf(pan, 0.0f, 1.0f) * 1.57079632679489661923f;
    return {x * cosf(angle), x * sinf(angle)};
  }
  float downmix() const { return 0.5f * (l + r); }
  StereoFrame scaled(float g) const { return {l * g, r * g}; }
  StereoFrame& operator+=(const StereoFrame& o) { l += o.l; r += o.r; return *this; }
};

}  // namespace orc
