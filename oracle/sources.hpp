// oracle/sources.hpp — TEST INFRASTRUCTURE ONLY.
// CPU restatement of PolySynth (instruments/poly_synth.rs) and Granulator (instruments/granulator.rs).
#pragma once
#include <memory>
#include "synths.hpp"

namespace orc {

// ============================ PolySynth ============================
enum PolyP { P_SHAPE, P_DETUNE, P_CUTOFF, P_RES, P_FENV_AMT, P_AMP_A, P_AMP_D, P_AMP_S, P_AMP_R, P_FLT_A, P_FLT_D, P_FLT_S, P_FLT_R, P_VOLUME, P_NPARAMS };
struct PolyConfig {
  float v[14];
  static PolyConfig preset(uint32_t id) {  // poly_synth.rs:49-142 ; ffi preset ids 0 default,1 pad,2 pluck,3 keys,4 strings
    static const float T[5][14] = {
        {0.0f, 0.2f, 0.6f, 0.15f, 0.3f, 0.55f, 0.7f, 0.7f, 0.8f, 0.5f, 0.65f, 0.4f, 0.75f, 0.7f},
        {0.0f, 0.4f, 0.45f, 0.2f, 0.2f, 0.8f, 0.75f, 0.8f, 0.85f, 0.75f, 0.7f, 0.5f, 0.8f, 0.6f},
        {0.3f, 0.1f, 0.7f, 0.25f, 0.6f, 0.0f, 0.75f, 0.0f, 0.65f, 0.0f, 0.7f, 0.1f, 0.65f, 0.7f},
        {0.5f, 0.15f, 0.55f, 0.1f, 0.4f, 0.35f, 0.7f, 0.5f, 0.75f, 0.3f, 0.65f, 0.3f, 0.7f, 0.7f},
        {0.0f, 0.5f, 0.5f, 0.1f, 0.15f, 0.85f, 0.7f, 0.9f, 0.85f, 0.8f, 0.7f, 0.6f, 0.8f, 0.5f}};
    PolyConfig c;
    const float* t = T[id < 5 ? id : 0];
    for (int i = 0; i < 14; i++) c.v[i] = t[i];
    return c;
  }
};
struct PolySynth {
  struct Voice {
    uint8_t midi_note = 0; double frequency = 440.0, phase_a = 0, phase_b = 0;
    Envelope amp_env, flt_env; StateVariableFilterTpt filter; float velocity = 1.0f; bool active = false; uint64_t trigger_order = 0;
    explicit Voice(float sr) : filter(sr, 1000.0f, 1.0f) {}
  };
  float sample_rate;
  SmoothedParam p[P_NPARAMS];
  std::vector<Voice> voices;
  uint64_t trigger_counter = 0;
  double current_time = 0.0;
  explicit PolySynth(float sr, const PolyConfig& c = PolyConfig::preset(0)) : sample_rate(sr) {
    for (int i = 0; i < 14; i++) p[i] = SmoothedParam(clampf(c.v[i], 0, 1), 0.0f, 1.0f, sr, 15.0f);
    for (int i = 0; i < 6; i++) voices.emplace_back(sr);
  }
  static float env_time(float n) { return 0.001f * powf(5000.0f, n); }
  void set_config(const PolyConfig& c) { for (int i = 0; i < 14; i++) p[i].set_target(c.v[i]); }
  void set_param(uint32_t id, float v) { if (id < 14) p[id].set_target(clampf(v, 0.0f, 1.0f)); }
  size_t allocate_voice() const {  // :422-435
    for (size_t i = 0; i < voices.size(); i++) if (!voices[i].active) return i;
    size_t best = 0;
    for (size_t i = 1; i < voices.size(); i++) if (voices[i].trigger_order < voices[best].trigger_order) best = i;
    return best;
  }
  void trigger_note(uint8_t note, float velocity) {  // :309-342
    double time = current_time;
    Voice& v = voices[allocate_voice()];
    v.midi_note = note;
    v.frequency = 440.0 * pow(2.0, ((double)note - 69.0) / 12.0);
    v.phase_a = v.phase_b = 0.0;
    v.velocity = velocity;
    v.active = true;
    v.trigger_order = trigger_counter++;
    v.amp_env.set_config(ADSRConfig(env_time(p[P_AMP_A].get()), env_time(p[P_AMP_D].get()), p[P_AMP_S].get(), env_time(p[P_AMP_R].get())).with_decay_curve(EnvelopeCurve::Exponential(0.5f)));
    v.amp_env.trigger(time);
    v.flt_env.set_config(ADSRConfig(env_time(p[P_FLT_A].get()), env_time(p[P_FLT_D].get()), p[P_FLT_S].get(), env_time(p[P_FLT_R].get())).with_decay_curve(EnvelopeCurve::Exponential(0.5f)));
    v.flt_env.trigger(time);
    v.filter.reset();
  }
  void release_all() { double t = current_time; for (auto& v : voices) if (v.active) { v.amp_env.release(t); v.flt_env.release(t); } }
  float generate_voice(Voice& v, double now) {  // :437-502
    float osc_shape = p[P_SHAPE].get(), detune = p[P_DETUNE].get(), cutoff_norm = p[P_CUTOFF].get(), res_norm = p[P_RES].get();
    float fenv_amt = p[P_FENV_AMT].get(), volume = p[P_VOLUME].get();
    if (!v.active) return 0.0f;
    float amp_env = v.amp_env.get_amplitude(now);
    if (!v.amp_env.is_active) { v.active = false; return 0.0f; }
    float flt_env = v.flt_env.get_amplitude(now);
    double freq = v.frequency;
    double detune_ratio = 1.0 + (double)detune * 0.0175;
    double dt = 1.0 / (double)sample_rate;
    double inc_a = freq * dt, inc_b = freq * detune_ratio * dt;
    float saw_a = polyblep_saw(v.phase_a, inc_a), sq_a = polyblep_square(v.phase_a, inc_a);
    float osc_a = saw_a * (1.0f - osc_shape) + sq_a * osc_shape;
    float saw_b = polyblep_saw(v.phase_b, inc_b), sq_b = polyblep_square(v.phase_b, inc_b);
    float osc_b = saw_b * (1.0f - osc_shape) + sq_b * osc_shape;
    float osc_mix = (osc_a + osc_b) * 0.5f;
    v.phase_a += inc_a; v.phase_a -= floor(v.phase_a);
    v.phase_b += inc_b; v.phase_b -= floor(v.phase_b);
    float base_cutoff = 20.0f * powf(18000.0f / 20.0f, cutoff_norm);
    float max_cutoff = 18000.0f;
    float mod_cutoff = base_cutoff + fenv_amt * flt_env * (max_cutoff - base_cutoff);
    float q = 0.5f + res_norm * 14.5f;
    v.filter.set_params(clampf(mod_cutoff, 20.0f, 18000.0f), q);
    float lo, bd, hi;
    v.filter.process_all(osc_mix, lo, bd, hi);
    return lo * amp_env * sqrtf(v.velocity) * volume;
  }
  float tick(double now) {  // :512-525
    current_time = now;
    for (auto& s : p) s.tick();
    float out = 0.0f;
    for (auto& v : voices) out += generate_voice(v, now);
    return out * (1.0f / 4.0f);
  }
};

// ============================ Granulator ============================
enum GranP { G_SCAN, G_LENGTH, G_SPRAY, G_PITCH, G_DENSITY, G_TEXTURE, G_DIRECTION, G_CLOUD, G_VOLUME, G_RAND_TIMING, G_RAND_AMP, G_DRIVE, G_NPARAMS };
struct Grain { bool active = false; float source_pos = 0, age = 0, duration = 1, speed = 1, direction = 1, window_shape = 1, velocity = 1, release_samples = 0, release_total = 0; };
struct Granulator {
  float sample_rate;
  std::shared_ptr<std::vector<float>> buffer;
  float buffer_sr;
  SmoothedParam p[G_NPARAMS];
  Grain grains[64], release_grains[16];
  SmoothedParam gain_comp;
  bool cloud_active = false;
  double cloud_end_time = 0, next_grain_time = 0;
  float current_velocity = 1.0f;
  uint32_t rng = 0x1234abcdu;
  Waveshaper drive_shaper;
  explicit Granulator(float sr) : sample_rate(sr), buffer(std::make_shared<std::vector<float>>(1, 0.0f)), buffer_sr(44100.0f),
                                  gain_comp(1.0f, 0.0f, 1.0f, sr, 10.0f), drive_shaper(4.0f, 0.0f) {
    const float D[12] = {0.5f, 0.16f, 0.12f, 0.5f, 0.35f, 0.25f, 0.0f, 0.35f, 0.8f, 0.0f, 0.0f, 0.0f};  // GranulatorConfig::default :188-205
    for (int i = 0; i < 12; i++) p[i] = SmoothedParam(D[i], 0.0f, 1.0f, sr, 15.0f);
  }
  void set_buffer(std::shared_ptr<std::vector<float>> b, float sr) { buffer = b; buffer_sr = sr; for (auto& g : grains) g.active = false; for (auto& g : release_grains) g.active = false; cloud_active = false; }
  void set_seed(uint32_t s) { rng = s == 0 ? 0x6d2b79f5u : s; }
  void snap_params() { for (auto& s : p) s.snap(); gain_comp.snap(); }
  void set_param(uint32_t id, float v) { if (id < 12) p[id].set_target(clampf(v, 0.0f, 1.0f)); }
  float next_f32() { uint32_t x = rng; x ^= x << 13; x ^= x >> 17; x ^= x << 5; rng = x; return (float)x / 4294967296.0f; }  // u32::MAX as f32 == 2^32
  static float grain_length_ms(float v) { v = clampf(v, 0, 1); return 5.0f + v * v * (3000.0f - 5.0f); }
  static float spray_seconds(float v) { v = clampf(v, 0, 1); return v * v * v * 10.0f; }
  static float pitch_ratio(float v) { v = clampf(v, 0, 1); return 0.25f * powf(4.0f / 0.25f, v); }
  static float cloud_ms(float v) { v = clampf(v, 0, 1); return 50.0f + v * v * (8000.0f - 50.0f); }
  float sample_clamped(long i) const { long last = (long)buffer->size() - 1; if (i < 0) i = 0; if (i > last) i = last; return (*buffer)[i]; }
  float sample_interpolated(float pos) const {
    if (buffer->size() == 1) return (*buffer)[0];
    float last = (float)buffer->size() - 1.0f;
    pos = clampf(pos, 0.0f, last);
    long idx = (long)floorf(pos);
    float frac = pos - (float)idx;
    return cubic_interpolate(sample_clamped(idx - 1), sample_clamped(idx), sample_clamped(idx + 1), sample_clamped(idx + 2), frac);
  }
  void trigger_with_velocity(double t, float vel) {  // :722-728
    current_velocity = clampf(vel, 0.0f, 1.0f);
    cloud_active = true;
    cloud_end_time = t + (double)cloud_ms(p[G_CLOUD].target) * 0.001;
    next_grain_time = t;
  }
  int free_slot() const { for (int i = 0; i < 64; i++) if (!grains[i].active) return i; return -1; }
  bool steal_grain() {  // :626-659
    int victim = -1; float shortest = INFINITY;
    for (int i = 0; i < 64; i++) { if (!grains[i].active) continue; float rem = rust_max(grains[i].duration - grains[i].age, 0.0f); if (rem < shortest) { shortest = rem; victim = i; } }
    if (victim < 0) return false;
    int rs = -1;
    for (int i = 0; i < 16; i++) if (!release_grains[i].active) { rs = i; break; }
    if (rs < 0) return false;
    float release = rust_max(4.0f * 0.001f * sample_rate, 1.0f);
    float remaining = rust_max(grains[victim].duration - grains[victim].age, 1.0f);
    release = rust_min(release, remaining);
    Grain moved = grains[victim];
    moved.release_samples = release; moved.release_total = release;
    release_grains[rs] = moved;
    grains[victim].active = false;
    return true;
  }
  void spawn_grain() {  // :546-620
    float amp_jitter = next_f32();
    int slot = free_slot();
    if (slot < 0) { if (!steal_grain()) return; slot = free_slot(); if (slot < 0) return; }
    float last_sample = (float)(buffer->size() - 1);
    float scan = clampf(p[G_SCAN].get(), 0, 1) * last_sample;
    float spray_samples = spray_seconds(p[G_SPRAY].get()) * buffer_sr;
    float spray_offset = (next_f32() * 2.0f - 1.0f) * spray_samples;
    float req = clampf(scan + spray_offset, 0.0f, last_sample);
    float direction = next_f32() < p[G_DIRECTION].get() ? -1.0f : 1.0f;
    float speed = pitch_ratio(p[G_PITCH].get()) * (buffer_sr / sample_rate);
    float duration = rust_max(grain_length_ms(p[G_LENGTH].get()) * 0.001f * sample_rate, 1.0f);
    float wshape = 0.5f + clampf(p[G_TEXTURE].get(), 0, 1) * 3.5f;
    float travel = duration * speed;
    float source_pos;
    if (travel >= last_sample) { duration = rust_max(last_sample / speed, 1.0f); source_pos = direction < 0.0f ? last_sample : 0.0f; }
    else if (direction < 0.0f) source_pos = clampf(req, travel, last_sample);
    else source_pos = clampf(req, 0.0f, last_sample - travel);
    float random_amp = clampf(p[G_RAND_AMP].get(), 0, 1);
    float amp_factor = 1.0f - random_amp * amp_jitter;
    Grain g; g.active = true; g.source_pos = source_pos; g.age = 0; g.duration = duration; g.speed = speed; g.direction = direction;
    g.window_shape = wshape; g.velocity = current_velocity * amp_factor; g.release_samples = 0; g.release_total = 0;
    grains[slot] = g;
  }
  void spawn_due_grains(double now) {  // :511-544
    if (!cloud_active) return;
    if (now > cloud_end_time) { cloud_active = false; return; }
    float density = clampf(p[G_DENSITY].get(), 0, 1) * 80.0f;
    if (density <= 0.0f) return;
    double interval = 1.0 / (double)density;
    double random_timing = (double)clampf(p[G_RAND_TIMING].get(), 0, 1);
    int guard = 0;
    while (cloud_active && now + 1e-12 >= next_grain_time && guard < 8) {
      spawn_grain();
      next_grain_time += interval;
      if (random_timing > 0.0) {
        double jitter = ((double)next_f32() * 2.0 - 1.0) * interval * random_timing;
        next_grain_time = fmax(next_grain_time + jitter, now);
      }
      if (next_grain_time > cloud_end_time) cloud_active = false;
      guard++;
    }
  }
  void tick_slice(Grain* g, int n, float gc, float& out) {  // :683-718
    for (int i = 0; i < n; i++) {
      Grain& gr = g[i];
      if (!gr.active) continue;
      if (gr.age >= gr.duration) { gr.active = false; continue; }
      float phase = clampf(gr.age / gr.duration, 0.0f, 1.0f);
      float window = raised_sine_window(phase, gr.window_shape);
      float rg = gr.release_total > 0.0f ? clampf(gr.release_samples / gr.release_total, 0.0f, 1.0f) : 1.0f;
      float s = sample_interpolated(gr.source_pos);
      out += s * window * rg * gr.velocity * gc;
      gr.source_pos += gr.speed * gr.direction;
      gr.age += 1.0f;
      if (gr.release_samples > 0.0f) { gr.release_samples -= 1.0f; if (gr.release_samples <= 0.0f) gr.active = false; }
    }
  }
  float tick_grains() {  // :661-681
    int active = 0;
    for (auto& g : grains) active += g.active;
    for (auto& g : release_grains) active += g.active;
    if (active == 0) { gain_comp.set_target(1.0f); gain_comp.tick(); return 0.0f; }
    gain_comp.set_target(1.0f / sqrtf((float)active));
    float gc = gain_comp.tick();
    float out = 0.0f;
    tick_slice(grains, 64, gc, out);
    tick_slice(release_grains, 16, gc, out);
    return out;
  }
  float tick(double now) {  // :730-742
    for (auto& s : p) s.tick();
    spawn_due_grains(now);
    float raw = tick_grains();
    drive_shaper.set_mix(p[G_DRIVE].get());
    float driven = drive_shaper.process(raw);
    return driven * p[G_VOLUME].get();
  }
};

}  // namespace orc
