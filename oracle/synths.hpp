// oracle/synths.hpp — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the melodic voices: BassSynth (instruments/bass.rs).
#pragma once
#include "drums.hpp"

namespace orc {

enum BassP { B_FREQ, B_SUB, B_OSC, B_DETUNE_LEVEL, B_DETUNE_AMT, B_SHAPE, B_CUTOFF, B_RES, B_FENV_AMT, B_FENV_DECAY, B_FENV_CURVE,
             B_AMP_DECAY, B_AMP_CURVE, B_OVERDRIVE, B_VOLUME, B_TUNING, B_NPARAMS };

struct BassConfig {  // bass.rs:51-269, BassConfig::new order
  float v[15];
  static BassConfig make(std::initializer_list<float> a) { BassConfig c; int i = 0; for (float x : a) c.v[i++] = clampf(x, 0.0f, 1.0f); return c; }
  static BassConfig acid() { return make({0.24f, 0.40f, 0.80f, 0.00f, 0.00f, 0.10f, 0.15f, 0.70f, 0.85f, 0.15f, 0.08f, 0.35f, 0.10f, 0.30f, 0.80f}); }
  static BassConfig sub() { return make({0.18f, 1.00f, 0.15f, 0.00f, 0.00f, 0.00f, 0.70f, 0.05f, 0.10f, 0.30f, 0.20f, 0.60f, 0.15f, 0.00f, 0.85f}); }
  static BassConfig reese() { return make({0.18f, 0.30f, 0.80f, 0.80f, 0.50f, 0.05f, 0.35f, 0.30f, 0.50f, 0.40f, 0.15f, 0.55f, 0.12f, 0.60f, 0.80f}); }
  static BassConfig stab() { return make({0.30f, 0.20f, 0.90f, 0.00f, 0.00f, 0.90f, 0.20f, 0.40f, 0.90f, 0.08f, 0.05f, 0.20f, 0.08f, 0.20f, 0.80f}); }
};
static inline float exp_denorm(float n, float mn, float mx) { return mn * powf(mx / mn, clampf(n, 0.0f, 1.0f)); }

struct BassSynth : Instrument {
  float sample_rate;
  SmoothedParam p[B_NPARAMS];
  double sub_phase = 0, osc_phase = 0, detune_phase = 0;
  StateVariableFilterTpt filter;
  Envelope amp_envelope, filter_envelope;
  Waveshaper waveshaper;
  bool active = false;
  float current_velocity = 1.0f, triggered_frequency;
  BassSynth(float sr, const BassConfig& c = BassConfig::acid())
      : sample_rate(sr), filter(sr, exp_denorm(c.v[6], 20.0f, 18000.0f), denorm(c.v[7], 0.5f, 15.0f)), waveshaper(c.v[13], 1.0f),
        triggered_frequency(denorm(c.v[0], 30.0f, 200.0f)) {
    for (int i = 0; i < 15; i++) p[i] = SmoothedParam(c.v[i], 0.0f, 1.0f, sr, 15.0f);
    p[B_TUNING] = SmoothedParam(0.5f, 0.0f, 1.0f, sr, 15.0f);
  }
  void set_config(const BassConfig& c) { for (int i = 0; i < 15; i++) p[i].set_target(c.v[i]); }
  void set_config_flat(const float* v) override { BassConfig c; for (int i = 0; i < 15; i++) c.v[i] = v[i]; set_config(c); }
  void snap_params() override { for (auto& s : p) s.snap(); }
  void set_param(uint32_t id, float v) override { if (id < 16) p[id].set_target(clampf(v, 0.0f, 1.0f)); }  // ffi.rs:232-249
  void apply_modulation(uint32_t id, float v) override { if (id < 16) p[id].set_bipolar(v); }  // ffi.rs:386-403
  bool get_freq_param(float& f) const override { f = p[B_FREQ].get(); return true; }
  bool is_active() const override { return active; }
  void trigger_with_velocity(double time, float velocity) override {  // bass.rs:747-791
    current_velocity = clampf(velocity, 0.0f, 1.0f);
    active = true;
    sub_phase = osc_phase = detune_phase = 0.0;
    triggered_frequency = denorm(p[B_FREQ].get(), 30.0f, 200.0f);
    float amp_decay = denorm(p[B_AMP_DECAY].get(), 0.05f, 4.0f);
    float amp_curve = denorm(p[B_AMP_CURVE].get(), 0.1f, 10.0f);
    amp_envelope.set_config(ADSRConfig::raw(0.002f, amp_decay, 0.0f, amp_decay * 0.1f, EnvelopeCurve::Linear(), EnvelopeCurve::Exponential(amp_curve)));
    amp_envelope.trigger(time);
    float fd = denorm(p[B_FENV_DECAY].get(), 0.01f, 2.0f);
    float fc = denorm(p[B_FENV_CURVE].get(), 0.1f, 8.0f);
    filter_envelope.set_config(ADSRConfig::raw(0.001f, fd, 0.0f, fd * 0.1f, EnvelopeCurve::Linear(), EnvelopeCurve::Exponential(fc)));
    filter_envelope.trigger(time);
    filter.reset();
    waveshaper.set_drive(1.0f + p[B_OVERDRIVE].get() * 9.0f);
  }
  float tick(double now) override {  // bass.rs:793-877
    for (auto& s : p) s.tick();
    if (!active) return 0.0f;
    float freq = triggered_frequency * tuning_to_multiplier(p[B_TUNING].get());
    float sub_level = p[B_SUB].get(), osc_level = p[B_OSC].get(), detune_level = p[B_DETUNE_LEVEL].get();
    float detune_cents = denorm(p[B_DETUNE_AMT].get(), 0.0f, 30.0f);
    float osc_shape = p[B_SHAPE].get();
    float detune_ratio = powf(2.0f, detune_cents / 1200.0f);
    float detune_freq = freq * detune_ratio;
    double dt = 1.0 / (double)sample_rate;
    double sub_inc = (double)freq * dt, osc_inc = (double)freq * dt, det_inc = (double)detune_freq * dt;
    sub_phase += sub_inc; sub_phase -= floor(sub_phase);
    osc_phase += osc_inc; osc_phase -= floor(osc_phase);
    detune_phase += det_inc; detune_phase -= floor(detune_phase);
    float sub_out = (float)sin(sub_phase * 6.283185307179586476925286766559);
    float saw_m = polyblep_saw(osc_phase, osc_inc), sq_m = polyblep_square(osc_phase, osc_inc);
    float osc_out = saw_m * (1.0f - osc_shape) + sq_m * osc_shape;
    float saw_d = polyblep_saw(detune_phase, det_inc), sq_d = polyblep_square(detune_phase, det_inc);
    float det_out = saw_d * (1.0f - osc_shape) + sq_d * osc_shape;
    float mix = sub_out * sub_level + osc_out * osc_level + det_out * detune_level;
    float od = p[B_OVERDRIVE].get();
    waveshaper.set_drive(1.0f + od * 9.0f);
    float sat = od > 0.001f ? waveshaper.process(mix) : mix;
    float fenv = filter_envelope.get_amplitude(now);
    float base_cutoff = exp_denorm(p[B_CUTOFF].get(), 20.0f, 18000.0f);
    float env_amount = p[B_FENV_AMT].get();
    float env_offset = (18000.0f - base_cutoff) * env_amount * fenv;
    float cutoff = clampf(base_cutoff + env_offset, 20.0f, 18000.0f);
    float resonance = denorm(p[B_RES].get(), 0.5f, 15.0f);
    filter.set_params(cutoff, resonance);
    float lo, bd, hi;
    filter.process_all(sat, lo, bd, hi);
    float amp_env = amp_envelope.get_amplitude(now);
    float va = sqrtf(current_velocity);
    float out = lo * amp_env * va * p[B_VOLUME].get();
    if (!amp_envelope.is_active) active = false;
    return out;
  }
};

}  // namespace orc
