// oracle/drums.hpp — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the four drum voices (instruments/{kick,snare,hihat2,tom2}.rs).
#pragma once
#include "prims.hpp"

namespace orc {

static inline float denorm(float n, float mn, float mx) { return mn + clampf(n, 0.0f, 1.0f) * (mx - mn); }

struct Instrument {
  virtual ~Instrument() {}
  virtual void trigger_with_velocity(double t, float vel) = 0;
  virtual float tick(double t) = 0;
  virtual bool is_active() const = 0;
  virtual void snap_params() {}
  virtual void set_param(uint32_t, float) {}       // FFI param ids (ffi.rs:168-250)
  virtual bool get_freq_param(float&) const { return false; }
  // `set_config(cfg)` with the config given as flat values in the order of its fields (kick.rs:905-941, snare.rs:818-855,
  // hihat2.rs:382-390, tom2.rs:400-411, bass.rs:636-662); enum / integer fields are passed as floats.  Used by the preset blender.
  virtual void set_config_flat(const float*) {}
  // ChannelInstrument::apply_modulation (ffi.rs:322-405): bipolar -1..1 onto the FFI parameter's range (LFO routes)
  virtual void apply_modulation(uint32_t, float) {}
};

// ============================ KickDrum (instruments/kick.rs) ============================
enum KickP { K_FREQ, K_PUNCH, K_SUB, K_CLICK, K_OSC_DECAY, K_PITCH_ENV_AMT, K_PITCH_ENV_CURVE, K_VOLUME,
             K_PITCH_START_RATIO, K_PHASE_MOD, K_NOISE_AMT, K_NOISE_CUTOFF, K_NOISE_RES, K_OVERDRIVE,
             K_FEEDBACK, K_FB_CUTOFF, K_AMP_DECAY, K_AMP_DECAY_CURVE, K_TUNING, K_NPARAMS };

struct KickConfig {  // kick.rs:66-350 (normalized 0-1), order = new_full args
  float v[18];
  static KickConfig full(std::initializer_list<float> a) { KickConfig c; int i = 0; for (float x : a) c.v[i++] = clampf(x, 0.0f, 1.0f); return c; }
  static KickConfig tight() { return full({0.22f, 0.00f, 1.00f, 0.00f, 0.12f, 0.70f, 0.01f, 0.85f, 0.64f, 1.00f, 0.07f, 0.01f, 0.02f, 0.20f, 0.00f, 0.47f, 0.12f, 0.02f}); }
  static KickConfig punch() { return full({0.50f, 0.20f, 1.00f, 0.20f, 0.12f, 0.60f, 0.10f, 0.85f, 0.24f, 1.00f, 0.07f, 0.11f, 0.42f, 0.20f, 0.00f, 0.47f, 0.12f, 0.02f}); }
  static KickConfig loose() { return full({0.32f, 0.40f, 1.00f, 0.00f, 0.62f, 0.20f, 0.12f, 0.85f, 0.84f, 1.00f, 0.07f, 0.01f, 0.02f, 0.25f, 0.00f, 0.47f, 0.12f, 0.12f}); }
  static KickConfig dirt() { return full({0.62f, 0.10f, 1.00f, 0.10f, 0.10f, 0.60f, 0.10f, 0.85f, 0.44f, 1.00f, 0.20f, 0.10f, 0.82f, 0.20f, 0.00f, 0.47f, 0.10f, 0.10f}); }
};
static inline float overdrive_to_drive(float a) { return 1.0f + a * a * a * 40.0f; }

struct KickDrum : Instrument {
  float sample_rate;
  SmoothedParam p[K_NPARAMS];
  Oscillator sub_osc, punch_osc, click_osc;
  Envelope pitch_envelope;
  float triggered_pitch_multiplier;
  ResonantHighpassFilter click_filter;
  PhaseModulator phase_modulator;
  PinkNoise pink_noise;
  ResonantLowpassFilter noise_filter;
  Envelope noise_envelope;
  FeedbackWaveshaper waveshaper;
  Envelope amplitude_envelope;
  bool active = false;
  float current_velocity = 1.0f, velocity_to_decay = 0.5f, velocity_to_pitch = 0.7f;

  float freq_hz() const { return denorm(p[K_FREQ].get(), 30.0f, 120.0f); }
  float osc_decay_secs() const { return denorm(p[K_OSC_DECAY].get(), 0.01f, 4.0f); }

  KickDrum(float sr, const KickConfig& c = KickConfig::tight())
      : sample_rate(sr),
        sub_osc(sr, denorm(c.v[0], 30.0f, 120.0f)), punch_osc(sr, denorm(c.v[0], 30.0f, 120.0f) * 2.5f),
        click_osc(sr, denorm(c.v[0], 30.0f, 120.0f) * 40.0f),
        click_filter(sr, 8000.0f, 4.0f), pink_noise(sr),
        noise_filter(sr, denorm(c.v[11], 20.0f, 10000.0f), denorm(c.v[12], 0.0f, 5.0f)),
        waveshaper(sr, overdrive_to_drive(c.v[13]), c.v[14] * 0.98f, 200.0f + c.v[15] * 3800.0f, 1.0f) {
    for (int i = 0; i < 18; i++) p[i] = SmoothedParam(c.v[i], 0.0f, 1.0f, sr, 15.0f);
    p[K_TUNING] = SmoothedParam(0.5f, 0.0f, 1.0f, sr, 15.0f);
    float ratio = denorm(c.v[8], 1.0f, 10.0f);
    triggered_pitch_multiplier = 1.0f + (ratio - 1.0f) * c.v[5];
    configure_oscillators();
  }
  void configure_oscillators() {  // kick.rs:777-817
    float decay = osc_decay_secs();
    sub_osc.waveform = Waveform::Sine;
    sub_osc.set_adsr(ADSRConfig(0.001f, decay, 0.0f, decay * 0.2f));
    punch_osc.waveform = Waveform::Triangle;
    punch_osc.set_adsr(ADSRConfig(0.001f, decay, 0.0f, decay * 0.2f));
    click_osc.waveform = Waveform::Noise;
    click_osc.set_adsr(ADSRConfig(0.001f, decay * 0.2f, 0.0f, decay * 0.02f));
    float pd = decay * 0.6f;
    pitch_envelope.set_config(ADSRConfig(0.001f, pd, 0.0f, pd * 0.1f));
    noise_envelope.set_config(ADSRConfig(0.001f, decay, 0.0f, decay * 0.2f));
  }
  void apply_params() {  // :820-835
    float punch = p[K_PUNCH].get(), sub = p[K_SUB].get(), click = p[K_CLICK].get();
    float cvs = 0.6f + 0.4f * current_velocity;
    sub_osc.set_volume(sub);
    punch_osc.set_volume(punch * 0.7f);
    click_osc.set_volume(click * 0.15f * cvs);
  }
  void set_config(const KickConfig& c) { for (int i = 0; i < 18; i++) p[i].set_target(c.v[i]); }  // :837-879
  void set_config_flat(const float* v) override { KickConfig c; for (int i = 0; i < 18; i++) c.v[i] = v[i]; set_config(c); }
  void snap_params() override { for (auto& s : p) s.snap(); }
  void set_param(uint32_t id, float v) override {  // ffi.rs:170-180, ids ffi.rs:1737-1751
    static const int map[8] = {K_FREQ, K_PUNCH, K_SUB, K_CLICK, K_OSC_DECAY, K_PITCH_ENV_AMT, K_VOLUME, K_TUNING};
    if (id < 8) p[map[id]].set_target(clampf(v, 0.0f, 1.0f));
  }
  bool get_freq_param(float& f) const override { f = p[K_FREQ].get(); return true; }
  bool is_active() const override { return active; }
  void apply_modulation(uint32_t id, float v) override {  // ffi.rs:324-336: every FFI id except the pitch envelope (5)
    static const int map[8] = {K_FREQ, K_PUNCH, K_SUB, K_CLICK, K_OSC_DECAY, -1, K_VOLUME, K_TUNING};
    if (id < 8 && map[id] >= 0) p[map[id]].set_bipolar(v);
  }

  void trigger_with_velocity(double time, float velocity) override {  // :971-1086
    current_velocity = clampf(velocity, 0.0f, 1.0f);
    active = true;
    float vel = current_velocity;
    float vel2 = vel * vel;
    float decay_scale = 1.0f - (velocity_to_decay * vel2);
    float base_decay = osc_decay_secs() * decay_scale;
    float base_freq = freq_hz();
    float pea = p[K_PITCH_ENV_AMT].get();
    float psr = denorm(p[K_PITCH_START_RATIO].get(), 1.0f, 10.0f);
    triggered_pitch_multiplier = 1.0f + (psr - 1.0f) * pea;
    float pcv = denorm(p[K_PITCH_ENV_CURVE].get(), 0.1f, 4.0f);
    EnvelopeCurve dc = fabsf(pcv - 1.0f) < 0.01f ? EnvelopeCurve::Linear() : EnvelopeCurve::Exponential(pcv);
    pitch_envelope.set_config(ADSRConfig(0.001f, base_decay, 0.0f, base_decay * 0.2f).with_decay_curve(dc));
    sub_osc.set_adsr(ADSRConfig(0.001f, base_decay, 0.0f, base_decay * 0.2f));
    punch_osc.set_adsr(ADSRConfig(0.001f, base_decay, 0.0f, base_decay * 0.2f));
    click_osc.set_adsr(ADSRConfig(0.001f, base_decay * 0.2f, 0.0f, base_decay * 0.02f));
    sub_osc.frequency_hz = base_freq;
    punch_osc.frequency_hz = base_freq * 2.5f;
    click_osc.frequency_hz = base_freq * 40.0f;
    sub_osc.trigger(time);
    punch_osc.trigger(time);
    click_osc.trigger(time);
    pitch_envelope.trigger(time);
    if (p[K_PHASE_MOD].get() > 0.001f) phase_modulator.trigger(time);
    noise_envelope.set_config(ADSRConfig(0.001f, base_decay, 0.0f, base_decay * 0.2f));
    noise_envelope.trigger(time);
    float amp_decay = denorm(p[K_AMP_DECAY].get(), 0.0f, 4.0f) * decay_scale;
    float adc = denorm(p[K_AMP_DECAY_CURVE].get(), 0.1f, 10.0f);
    EnvelopeCurve aac = EnvelopeCurve::Exponential(0.5f);
    EnvelopeCurve adcv = fabsf(adc - 1.0f) < 0.01f ? EnvelopeCurve::Linear() : EnvelopeCurve::Exponential(adc);
    amplitude_envelope.set_config(ADSRConfig(0.001f, amp_decay, 0.0f, amp_decay * 0.2f).with_attack_curve(aac).with_decay_curve(adcv));
    amplitude_envelope.trigger(time);
    click_filter.reset();
    noise_filter.reset();
    pink_noise.reset();
  }

  float tick(double now) override {  // :1097-1232
    for (auto& s : p) s.tick();
    if (!active) return 0.0f;
    apply_params();
    float vel2 = current_velocity * current_velocity;
    float decay_scale = 1.0f - (velocity_to_decay * vel2);
    float base_decay = osc_decay_secs() * decay_scale;
    sub_osc.envelope.set_decay_time(base_decay);
    sub_osc.envelope.set_release_time(base_decay * 0.2f);
    punch_osc.envelope.set_decay_time(base_decay);
    punch_osc.envelope.set_release_time(base_decay * 0.2f);
    click_osc.envelope.set_decay_time(base_decay * 0.2f);
    click_osc.envelope.set_release_time(base_decay * 0.02f);
    noise_envelope.set_decay_time(base_decay);
    noise_envelope.set_release_time(base_decay * 0.2f);
    pitch_envelope.set_decay_time(base_decay);
    pitch_envelope.set_release_time(base_decay * 0.2f);
    float base_frequency = freq_hz() * tuning_to_multiplier(p[K_TUNING].get());
    float pev = pitch_envelope.get_amplitude(now);
    float fm = 1.0f + (triggered_pitch_multiplier - 1.0f) * pev;
    float pma = p[K_PHASE_MOD].get();
    if (pma > 0.001f) {
      float pm = phase_modulator.tick(now);
      fm *= 1.0f + (pm * pma * 2.0f);
    }
    sub_osc.frequency_hz = base_frequency * fm;
    punch_osc.frequency_hz = base_frequency * 2.5f * fm;
    float cpm = 1.0f + (fm - 1.0f) * 0.3f;
    click_osc.frequency_hz = base_frequency * 40.0f * cpm;
    float sub_out = sub_osc.tick(now);
    float punch_out = punch_osc.tick(now);
    float raw_click = click_osc.tick(now);
    float filt_click = click_filter.process(raw_click);
    float noise_amount = p[K_NOISE_AMT].get();
    float noise_out;
    if (noise_amount > 0.001f) {
      float pn = pink_noise.tick();
      noise_filter.set_params(denorm(p[K_NOISE_CUTOFF].get(), 20.0f, 10000.0f), denorm(p[K_NOISE_RES].get(), 0.0f, 5.0f));
      float fn = noise_filter.process(pn);
      float ne = noise_envelope.get_amplitude(now);
      noise_out = fn * ne * noise_amount * 0.5f;
    } else noise_out = 0.0f;
    float total = sub_out + punch_out + filt_click + noise_out;
    waveshaper.set_drive(overdrive_to_drive(p[K_OVERDRIVE].get()));
    waveshaper.set_feedback(p[K_FEEDBACK].get() * 0.98f);
    waveshaper.set_filter_cutoff(200.0f + p[K_FB_CUTOFF].get() * 3800.0f);
    float od = waveshaper.process(total);
    float amp_env = amplitude_envelope.get_amplitude(now);
    float va = sqrtf(current_velocity);
    float volume = p[K_VOLUME].get();
    float out = od * amp_env * va * volume;
    if (!amplitude_envelope.is_active) active = false;
    return out;
  }
};

// ============================ SnareDrum (instruments/snare.rs) ============================
enum SnareP { S_FREQ, S_DECAY, S_BRIGHTNESS, S_VOLUME, S_TONAL, S_NOISE, S_PITCH_DROP, S_TONAL_DECAY, S_TONAL_DECAY_CURVE,
              S_NOISE_DECAY, S_NOISE_TAIL_DECAY, S_FILTER_CUTOFF, S_FILTER_RES, S_XFADE, S_PHASE_MOD, S_OVERDRIVE,
              S_AMP_DECAY, S_AMP_DECAY_CURVE, S_TUNING, S_NPARAMS };

struct SnareConfig {  // snare.rs:68-351
  float frequency, tonal_amount, noise_amount, crack_amount, decay, pitch_drop, volume;
  float tonal_decay, tonal_decay_curve, noise_decay, noise_tail_decay, filter_cutoff, filter_resonance;
  uint8_t filter_type;
  float xfade, phase_mod_amount, overdrive_amount, amp_decay, amp_decay_curve;
  static SnareConfig basic(float f, float t, float n, float c, float d, float pd, float v) {  // ::new :99-132
    SnareConfig s;
    s.frequency = clampf(f, 0, 1); s.tonal_amount = clampf(t, 0, 1); s.noise_amount = clampf(n, 0, 1);
    s.crack_amount = clampf(c, 0, 1); s.decay = clampf(d, 0, 1); s.pitch_drop = clampf(pd, 0, 1); s.volume = clampf(v, 0, 1);
    s.tonal_decay = d * 0.8f; s.tonal_decay_curve = 0.091f; s.noise_decay = d * 0.6f; s.noise_tail_decay = d;
    s.filter_cutoff = 0.495f; s.filter_resonance = 0.053f; s.filter_type = 1; s.xfade = 0.5f; s.phase_mod_amount = 0.0f;
    s.overdrive_amount = 0.0f; s.amp_decay = 0.125f; s.amp_decay_curve = 0.02f;
    return s;
  }
  static SnareConfig full(float f, float t, float n, float c, float d, float pd, float v, float td, float tdc, float nd, float ntd,
                          float fc, float fr, uint8_t ft, float xf, float pm, float od, float ad, float adc) {  // new_full :135-180
    SnareConfig s;
    s.frequency = clampf(f, 0, 1); s.tonal_amount = clampf(t, 0, 1); s.noise_amount = clampf(n, 0, 1);
    s.crack_amount = clampf(c, 0, 1); s.decay = clampf(d, 0, 1); s.pitch_drop = clampf(pd, 0, 1); s.volume = clampf(v, 0, 1);
    s.tonal_decay = clampf(td, 0, 1); s.tonal_decay_curve = clampf(tdc, 0, 1); s.noise_decay = clampf(nd, 0, 1);
    s.noise_tail_decay = clampf(ntd, 0, 1); s.filter_cutoff = clampf(fc, 0, 1); s.filter_resonance = clampf(fr, 0, 1);
    s.filter_type = ft > 3 ? 3 : ft; s.xfade = clampf(xf, 0, 1); s.phase_mod_amount = clampf(pm, 0, 1);
    s.overdrive_amount = clampf(od, 0, 1); s.amp_decay = clampf(ad, 0, 1); s.amp_decay_curve = clampf(adc, 0, 1);
    return s;
  }
  static SnareConfig tight() { return basic(0.2f, 0.4f, 0.7f, 0.5f, 0.029f, 0.3f, 0.8f); }
  static SnareConfig loose() { return full(0.16f, 0.80f, 0.60f, 0.30f, 0.79f, 0.10f, 0.90f, 0.33f, 0.20f, 0.23f, 0.34f, 0.55f, 0.05f, 1, 0.50f, 0.00f, 0.10f, 0.12f, 0.02f); }
  static SnareConfig hiss() { return full(0.16f, 0.00f, 0.60f, 0.30f, 0.04f, 0.40f, 0.90f, 0.53f, 0.09f, 0.38f, 0.29f, 0.29f, 0.45f, 1, 0.50f, 1.00f, 0.20f, 0.18f, 0.02f); }
  static SnareConfig smack() { return full(0.2f, 0.3f, 0.8f, 0.0f, 0.029f, 0.3f, 0.85f, 0.014f, 0.091f, 0.034f, 0.086f, 0.293f, 0.158f, 1, 0.4f, 0.5f, 0.0f, 0.125f, 0.02f); }
  float get(int i) const {
    const float a[18] = {frequency, decay, crack_amount, volume, tonal_amount, noise_amount, pitch_drop, tonal_decay, tonal_decay_curve,
                         noise_decay, noise_tail_decay, filter_cutoff, filter_resonance, xfade, phase_mod_amount, overdrive_amount,
                         amp_decay, amp_decay_curve};
    return a[i];
  }
};

struct SnareDrum : Instrument {
  float sample_rate;
  SmoothedParam p[S_NPARAMS];
  uint8_t filter_type;
  Oscillator tonal_osc, noise_osc, crack_osc;
  Envelope pitch_envelope;
  float pitch_start_multiplier;
  bool active = false;
  float current_velocity = 0.5f, velocity_to_decay = 0.45f, velocity_to_pitch = 0.5f;
  StateVariableFilter noise_filter;
  PhaseModulator phase_modulator;
  Envelope noise_tail_envelope, tonal_envelope, main_noise_envelope;
  Waveshaper waveshaper;
  Envelope amplitude_envelope;

  SnareDrum(float sr, const SnareConfig& c = SnareConfig::tight())
      : sample_rate(sr), filter_type(c.filter_type),
        tonal_osc(sr, denorm(c.frequency, 100.0f, 600.0f)), noise_osc(sr, denorm(c.frequency, 100.0f, 600.0f) * 8.0f),
        crack_osc(sr, denorm(c.frequency, 100.0f, 600.0f) * 25.0f),
        pitch_start_multiplier(1.0f + c.pitch_drop * 1.5f),
        noise_filter(sr, denorm(c.filter_cutoff, 100.0f, 10000.0f), denorm(c.filter_resonance, 0.5f, 10.0f)),
        waveshaper(1.0f, 1.0f) {
    for (int i = 0; i < 18; i++) p[i] = SmoothedParam(c.get(i), 0.0f, 1.0f, sr, 15.0f);
    p[S_TUNING] = SmoothedParam(0.5f, 0.0f, 1.0f, sr, 15.0f);
    tonal_osc.waveform = Waveform::Triangle;
    noise_osc.waveform = Waveform::Noise;
    crack_osc.waveform = Waveform::Noise;
  }
  void set_config_flat(const float* v) override {
    SnareConfig c;
    c.frequency = v[0]; c.tonal_amount = v[1]; c.noise_amount = v[2]; c.crack_amount = v[3]; c.decay = v[4]; c.pitch_drop = v[5]; c.volume = v[6];
    c.tonal_decay = v[7]; c.tonal_decay_curve = v[8]; c.noise_decay = v[9]; c.noise_tail_decay = v[10]; c.filter_cutoff = v[11]; c.filter_resonance = v[12];
    c.filter_type = (uint8_t)v[13]; c.xfade = v[14]; c.phase_mod_amount = v[15]; c.overdrive_amount = v[16]; c.amp_decay = v[17]; c.amp_decay_curve = v[18];
    set_config(c);
  }
  void set_config(const SnareConfig& c) {  // :818-855
    pitch_start_multiplier = 1.0f + c.pitch_drop * 1.5f;
    for (int i = 0; i < 18; i++) p[i].set_target(c.get(i));
    filter_type = c.filter_type;
  }
  void snap_params() override { for (auto& s : p) s.snap(); }
  void set_param(uint32_t id, float v) override {  // ffi.rs:181-203; ids ffi.rs:1775-1813
    // FFI id order: freq, decay, brightness, volume, tonal, noise, pitch_drop, tonal_decay, noise_decay, noise_tail_decay,
    // filter_cutoff, filter_resonance, filter_type, xfade, phase_mod, overdrive, amp_decay, amp_decay_curve, tonal_decay_curve, tuning
    static const int map[20] = {S_FREQ, S_DECAY, S_BRIGHTNESS, S_VOLUME, S_TONAL, S_NOISE, S_PITCH_DROP, S_TONAL_DECAY, S_NOISE_DECAY,
                                S_NOISE_TAIL_DECAY, S_FILTER_CUTOFF, S_FILTER_RES, -1, S_XFADE, S_PHASE_MOD, S_OVERDRIVE, S_AMP_DECAY,
                                S_AMP_DECAY_CURVE, S_TONAL_DECAY_CURVE, S_TUNING};
    if (id >= 20) return;
    if (id == 12) {  // value as u8 (saturating), then .min(3)
      float f = v;
      int t = !(f == f) ? 0 : (f <= 0.0f ? 0 : (f >= 255.0f ? 255 : (int)f));
      filter_type = (uint8_t)(t > 3 ? 3 : t);
      return;
    }
    p[map[id]].set_target(clampf(v, 0.0f, 1.0f));
  }
  void apply_modulation(uint32_t id, float v) override {  // ffi.rs:337-358: every FFI id except the filter type (12)
    static const int map[20] = {S_FREQ, S_DECAY, S_BRIGHTNESS, S_VOLUME, S_TONAL, S_NOISE, S_PITCH_DROP, S_TONAL_DECAY, S_NOISE_DECAY,
                                S_NOISE_TAIL_DECAY, S_FILTER_CUTOFF, S_FILTER_RES, -1, S_XFADE, S_PHASE_MOD, S_OVERDRIVE, S_AMP_DECAY,
                                S_AMP_DECAY_CURVE, S_TONAL_DECAY_CURVE, S_TUNING};
    if (id < 20 && map[id] >= 0) p[map[id]].set_bipolar(v);
  }
  bool is_active() const override { return active; }
  float freq_hz() const { return denorm(p[S_FREQ].get(), 100.0f, 600.0f); }
  float decay_secs() const { return denorm(p[S_DECAY].get(), 0.05f, 3.5f); }

  void trigger_with_velocity(double time, float velocity) override {  // :873-1027
    current_velocity = clampf(velocity, 0.0f, 1.0f);
    active = true;
    float vel = current_velocity, vel2 = vel * vel;
    float decay_scale = 1.0f - (velocity_to_decay * vel2);
    float pitch_decay_scale = 1.0f - (velocity_to_pitch * vel2);
    float base_freq = freq_hz();
    float base_decay = decay_secs();
    float brightness = p[S_BRIGHTNESS].get(), tonal_amount = p[S_TONAL].get(), noise_amount = p[S_NOISE].get();
    float pitch_drop = p[S_PITCH_DROP].get();
    float tonal_decay = denorm(p[S_TONAL_DECAY].get(), 0.0f, 3.5f);
    float tonal_decay_curve = denorm(p[S_TONAL_DECAY_CURVE].get(), 0.1f, 10.0f);
    float noise_decay = denorm(p[S_NOISE_DECAY].get(), 0.0f, 3.5f);
    float noise_tail_decay = denorm(p[S_NOISE_TAIL_DECAY].get(), 0.0f, 3.5f);
    float amp_decay = denorm(p[S_AMP_DECAY].get(), 0.0f, 4.0f);
    float amp_decay_curve = denorm(p[S_AMP_DECAY_CURVE].get(), 0.1f, 10.0f);
    float scaled_decay = base_decay * decay_scale;
    pitch_start_multiplier = 1.0f + pitch_drop * 1.5f;
    float pdt = rust_min(scaled_decay * 0.3f * pitch_decay_scale, scaled_decay * 0.25f);
    pitch_envelope.set_config(ADSRConfig(0.001f, pdt, 0.0f, pdt * 0.1f));
    tonal_osc.frequency_hz = base_freq;
    tonal_osc.set_volume(tonal_amount);
    tonal_osc.set_adsr(ADSRConfig(0.001f, 0.001f, 1.0f, scaled_decay * 0.4f));
    noise_osc.frequency_hz = base_freq * 8.0f;
    noise_osc.set_volume(noise_amount * 0.8f);
    noise_osc.set_adsr(ADSRConfig(0.001f, 0.001f, 1.0f, scaled_decay * 0.3f));
    float cvs = 0.7f + 0.3f * vel;
    crack_osc.frequency_hz = base_freq * 25.0f;
    crack_osc.set_volume(brightness * 0.4f * cvs);
    crack_osc.set_adsr(ADSRConfig(0.001f, scaled_decay * 0.2f, 0.0f, scaled_decay * 0.1f));
    float std_ = tonal_decay * decay_scale;
    ADSRConfig tc(0.001f, std_, 0.0f, std_ * 0.2f);
    tc.decay_curve = EnvelopeCurve::Exponential(tonal_decay_curve);
    tonal_envelope.set_config(tc);
    float snd = noise_decay * decay_scale;
    main_noise_envelope.set_config(ADSRConfig(0.001f, snd, 0.0f, snd * 0.2f));
    float stl = noise_tail_decay * decay_scale;
    noise_tail_envelope.set_config(ADSRConfig(0.001f, stl, 0.0f, stl * 0.3f));
    float sad = amp_decay * decay_scale;
    ADSRConfig ac(0.001f, sad, 0.0f, sad * 0.2f);
    ac.decay_curve = EnvelopeCurve::Exponential(amp_decay_curve);
    amplitude_envelope.set_config(ac);
    tonal_osc.trigger(time);
    noise_osc.trigger(time);
    crack_osc.trigger(time);
    pitch_envelope.trigger(time);
    tonal_envelope.trigger(time);
    main_noise_envelope.trigger(time);
    noise_tail_envelope.trigger(time);
    amplitude_envelope.trigger(time);
    if (p[S_PHASE_MOD].get() > 0.001f) phase_modulator.trigger(time);
    noise_filter.reset();
  }
  void apply_params() {  // :1206-1220
    float cvs = 0.7f + 0.3f * current_velocity;
    tonal_osc.set_volume(p[S_TONAL].get());
    noise_osc.set_volume(p[S_NOISE].get() * 0.8f);
    crack_osc.set_volume(p[S_BRIGHTNESS].get() * 0.4f * cvs);
  }
  float tick(double now) override {  // :1044-1198
    for (auto& s : p) s.tick();
    bool changing = false;
    for (auto& s : p) if (!s.is_settled()) changing = true;
    if (!active) return 0.0f;
    if (changing) apply_params();
    float vel2 = current_velocity * current_velocity;
    float decay_scale = 1.0f - (velocity_to_decay * vel2);
    float pitch_decay_scale = 1.0f - (velocity_to_pitch * vel2);
    float scaled_decay = decay_secs() * decay_scale;
    float pdt = rust_min(scaled_decay * 0.3f * pitch_decay_scale, scaled_decay * 0.25f);
    pitch_envelope.set_decay_time(pdt);
    pitch_envelope.set_release_time(pdt * 0.1f);
    tonal_osc.envelope.set_release_time(scaled_decay * 0.4f);
    noise_osc.envelope.set_release_time(scaled_decay * 0.3f);
    crack_osc.envelope.set_decay_time(scaled_decay * 0.2f);
    crack_osc.envelope.set_release_time(scaled_decay * 0.1f);
    float std_ = denorm(p[S_TONAL_DECAY].get(), 0.0f, 3.5f) * decay_scale;
    tonal_envelope.set_decay_time(std_);
    tonal_envelope.set_release_time(std_ * 0.2f);
    float snd = denorm(p[S_NOISE_DECAY].get(), 0.0f, 3.5f) * decay_scale;
    main_noise_envelope.set_decay_time(snd);
    main_noise_envelope.set_release_time(snd * 0.2f);
    float stl = denorm(p[S_NOISE_TAIL_DECAY].get(), 0.0f, 3.5f) * decay_scale;
    noise_tail_envelope.set_decay_time(stl);
    noise_tail_envelope.set_release_time(stl * 0.3f);
    float sad = denorm(p[S_AMP_DECAY].get(), 0.0f, 4.0f) * decay_scale;
    amplitude_envelope.set_decay_time(sad);
    amplitude_envelope.set_release_time(sad * 0.2f);
    float base_frequency = freq_hz() * tuning_to_multiplier(p[S_TUNING].get());
    float pev = pitch_envelope.get_amplitude(now);
    float fm = 1.0f + (pitch_start_multiplier - 1.0f) * pev;
    float pma = p[S_PHASE_MOD].get();
    if (pma > 0.001f) {
      float pm = phase_modulator.tick(now);
      fm *= 1.0f + (pm * pma * 1.0f);
    }
    tonal_osc.frequency_hz = base_frequency * fm;
    noise_filter.set_params(denorm(p[S_FILTER_CUTOFF].get(), 100.0f, 10000.0f), denorm(p[S_FILTER_RES].get(), 0.5f, 10.0f));
    float xfade = p[S_XFADE].get();
    float tonal_mix = 1.0f - xfade, noise_mix = xfade;
    float raw_tonal = tonal_osc.tick(now);
    float tonal_env = tonal_envelope.get_amplitude(now);
    float tonal_out = raw_tonal * tonal_env * tonal_mix;
    float raw_noise = noise_osc.tick(now);
    float filtered = noise_filter.process_mode(raw_noise, filter_type);
    float ne = main_noise_envelope.get_amplitude(now);
    float te = noise_tail_envelope.get_amplitude(now);
    float cne = (ne * 0.7f) + (te * 0.3f);
    float noise_out = filtered * cne * noise_mix;
    float crack_out = crack_osc.tick(now);
    float total = tonal_out + noise_out + crack_out;
    float drive = 1.0f + (p[S_OVERDRIVE].get() * 9.0f);
    waveshaper.set_drive(drive);
    float od = waveshaper.process(total);
    float amp_env = amplitude_envelope.get_amplitude(now);
    float va = sqrtf(current_velocity);
    float volume = p[S_VOLUME].get();
    float out = od * amp_env * va * volume;
    bool classic = tonal_osc.envelope.is_active || noise_osc.envelope.is_active || crack_osc.envelope.is_active;
    bool ds = tonal_envelope.is_active || main_noise_envelope.is_active || noise_tail_envelope.is_active ||
              amplitude_envelope.is_active || phase_modulator.is_active;
    if (!classic && !ds) active = false;
    return out;
  }
};

// ============================ HiHat2 (instruments/hihat2.rs) ============================
enum HatP { H_PITCH, H_DECAY, H_ATTACK, H_TONE, H_VOLUME, H_TUNING, H_NPARAMS };
struct HiHat2Config {
  float pitch, decay, attack; bool pink; bool db24; float tone, volume;
  static HiHat2Config make(float p, float d, float a, bool pink, bool db24, float t) {
    return {clampf(p, 0, 1), clampf(d, 0, 1), clampf(a, 0, 1), pink, db24, clampf(t, 0, 1), 1.0f};
  }
  static HiHat2Config short_() { return make(0.76f, 0.05f, 0.00f, false, true, 1.00f); }
  static HiHat2Config loose() { return make(0.76f, 0.30f, 0.00f, false, true, 1.00f); }
  static HiHat2Config dark() { return make(0.41f, 0.05f, 0.00f, false, true, 0.15f); }
  static HiHat2Config soft() { return make(0.41f, 0.05f, 0.15f, false, true, 0.60f); }
};
struct PhaseModOsc {  // hihat2.rs:258-287
  float sample_rate, frequency_hz, phase_cycle = 0;
  PhaseModOsc(float sr, float f) : sample_rate(sr), frequency_hz(f) {}
  void set_frequency(float f) { frequency_hz = rust_max(f, 0.0f); }
  float tick(float pm) {
    float inc = frequency_hz / sample_rate;
    phase_cycle = fmodf(phase_cycle + inc, 1.0f);
    float ph = phase_cycle + pm;
    ph -= floorf(ph);
    return sinf(2.0f * PI_F * ph);
  }
};
struct AsymmetricSmoother {  // :289-322
  float current = 0, down_coeff;
  explicit AsymmetricSmoother(float n) { down_coeff = n <= 0.0f ? 1.0f : 1.0f - expf(-1.0f / n); }
  float process(float t) { if (t >= current) current = t; else current += down_coeff * (t - current); return current; }
};
struct HiHat2 : Instrument {
  float sample_rate;
  SmoothedParam p[H_NPARAMS];
  bool pink, db24;
  PhaseModOsc mod_osc, main_osc;
  MaxCurveEnvelope envelope;
  AsymmetricSmoother env_smoother;
  BiquadHighpass hpf1, hpf2;
  StateVariableFilterTpt svf;
  uint64_t white_state = 0x123456789abcdef0ull;
  PinkNoise pink_noise;
  bool active = false;
  float current_velocity = 1.0f;
  static float pitch_hz_of(float pitch) { return denorm(pitch * pitch, 3500.0f, 10000.0f); }
  HiHat2(float sr, const HiHat2Config& c = HiHat2Config::short_())
      : sample_rate(sr), pink(c.pink), db24(c.db24), mod_osc(sr, pitch_hz_of(c.pitch) * 0.1f), main_osc(sr, pitch_hz_of(c.pitch)),
        env_smoother(100.0f), hpf1(sr), hpf2(sr), svf(sr, denorm(c.tone, 500.0f, 10000.0f), 0.5f), pink_noise(sr) {
    p[H_PITCH] = SmoothedParam(c.pitch, 0, 1, sr, 15.0f); p[H_DECAY] = SmoothedParam(c.decay, 0, 1, sr, 15.0f);
    p[H_ATTACK] = SmoothedParam(c.attack, 0, 1, sr, 15.0f); p[H_TONE] = SmoothedParam(c.tone, 0, 1, sr, 15.0f);
    p[H_VOLUME] = SmoothedParam(c.volume, 0, 1, sr, 15.0f); p[H_TUNING] = SmoothedParam(0.5f, 0, 1, sr, 15.0f);
  }
  void set_config_flat(const float* v) override { HiHat2Config c{v[0], v[1], v[2], v[3] != 0.0f, v[4] != 0.0f, v[5], v[6]}; set_config(c); }
  void set_config(const HiHat2Config& c) {  // :390-398 (tuning untouched)
    p[H_PITCH].set_target(c.pitch); p[H_DECAY].set_target(c.decay); p[H_ATTACK].set_target(c.attack);
    p[H_TONE].set_target(c.tone); p[H_VOLUME].set_target(c.volume); pink = c.pink; db24 = c.db24;
  }
  void snap_params() override { for (auto& s : p) s.snap(); }
  void set_param(uint32_t id, float v) override {  // ffi.rs:204-212; ids :1758-1768 pitch,decay,attack,volume,tone,tuning
    switch (id) {
      case 0: p[H_PITCH].set_target(v); break;
      case 1: p[H_DECAY].set_target(v); break;
      case 2: p[H_ATTACK].set_target(v); break;
      case 3: p[H_TONE].set_target(v); break;
      case 4: p[H_VOLUME].set_target(clampf(v, 0.0f, 1.0f)); break;
      case 5: p[H_TUNING].set_target(clampf(v, 0.0f, 1.0f)); break;
    }
  }
  void apply_modulation(uint32_t id, float v) override {  // ffi.rs:359-367
    static const int map[6] = {H_PITCH, H_DECAY, H_ATTACK, H_TONE, H_VOLUME, H_TUNING};
    if (id < 6) p[map[id]].set_bipolar(v);
  }
  bool is_active() const override { return active; }
  float attack_ms() const { return denorm(p[H_ATTACK].get(), 0.5f, 200.0f); }
  float decay_ms() const { return denorm(p[H_DECAY].get(), 0.5f, 4000.0f); }
  void trigger_with_velocity(double time, float velocity) override {  // :434-451
    active = true;
    current_velocity = clampf(velocity, 0.0f, 1.0f);
    envelope = MaxCurveEnvelope({{1.0f, attack_ms(), -0.3f}, {0.0f, decay_ms(), -0.8f}});
    envelope.set_initial_value(0.0f);
    envelope.trigger(time);
    env_smoother.current = 0.0f;
    mod_osc.phase_cycle = 0; main_osc.phase_cycle = 0;
    hpf1.reset(); hpf2.reset(); svf.reset();
  }
  float white_tick() {  // :514-525
    uint64_t x = white_state;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    white_state = x;
    uint64_t h = x * 0x2545F4914F6CDD1Dull;
    float n = (float)h / 18446744073709551616.0f;
    return n * 2.0f - 1.0f;
  }
  float tick(double now) override {  // :453-508
    for (auto& s : p) s.tick();
    if (!active) return 0.0f;
    envelope.set_segment_duration_ms(0, attack_ms());
    envelope.set_segment_duration_ms(1, decay_ms());
    float pitch_hz = pitch_hz_of(p[H_PITCH].get()) * tuning_to_multiplier(p[H_TUNING].get());
    mod_osc.set_frequency(pitch_hz * 0.1f);
    main_osc.set_frequency(pitch_hz);
    float noise = pink ? pink_noise.tick() : white_tick();
    float mod_out = mod_osc.tick(noise * 0.25f);
    float main_out = main_osc.tick(mod_out * 0.75f);
    hpf1.set_params(pitch_hz, 1.0f);
    float filtered = hpf1.process(main_out);
    if (db24) { hpf2.set_params(pitch_hz, 1.0f); filtered = hpf2.process(filtered) * 0.8f; }
    float env = envelope.get_value(now);
    env = env_smoother.process(env);
    float out = filtered * env * current_velocity * 0.35f;
    svf.set_params(denorm(p[H_TONE].get(), 500.0f, 10000.0f), 0.5f);
    float lo, bd, hi;
    svf.process_all(out, lo, bd, hi);
    float o = hi * p[H_VOLUME].get();
    if (envelope.is_complete() && env_smoother.current < 1e-4f) active = false;
    return o;
  }
};

// ============================ Tom2 (instruments/tom2.rs) ============================
struct Tom2Config { float tune, bend, tone, color, decay, membrane, membrane_q, volume; };
struct Tom2 : Instrument {
  float sample_rate;
  MorphOsc morph;
  ClickOsc click;
  BiquadBandpass bp;
  MaxCurveEnvelope envelope;
  bool active = false;
  float tri_phase = 0;
  bool past_attack = false;
  float tune = 50, bend = 30, tone = 50, color = 50, decay = 50;
  bool triangle_enabled = true;
  MembraneResonator membrane_res;
  float membrane = 0, membrane_q = 50, tuning = 0.5f, volume = 100;
  bool main_sound_done = false;
  explicit Tom2(float sr) : sample_rate(sr), morph(sr), bp(sr), envelope({{1.0f, 1.0f, 0.8f}, {0.0f, 2000.0f, -0.83f}}), membrane_res(sr) { update_membrane(); }
  void update_membrane() {  // :405-411
    float qs = 0.005f + (membrane_q / 100.0f) * 0.015f;
    membrane_res.set_q_scale(qs);
    membrane_res.set_gain_scale(0.003f);
  }
  void set_config_flat(const float* v) override { set_config(Tom2Config{v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]}); }
  void set_config(const Tom2Config& c) { tune = c.tune; bend = c.bend; tone = c.tone; color = c.color; decay = c.decay; membrane = c.membrane; membrane_q = c.membrane_q; volume = c.volume; update_membrane(); }
  void set_param(uint32_t id, float v) override {  // ffi.rs:213-231; ids :1820-1836
    float s = clampf(v, 0.0f, 1.0f) * 100.0f;
    switch (id) {
      case 0: tune = clampf(s, 0, 100); break;
      case 1: bend = clampf(s, 0, 100); break;
      case 2: tone = clampf(s, 0, 100); break;
      case 3: color = clampf(s, 0, 100); break;
      case 4: decay = clampf(s, 0, 100); break;
      case 5: membrane = clampf(s, 0, 100); break;
      case 6: membrane_q = clampf(s, 0, 100); update_membrane(); break;
      case 7: volume = clampf(s, 0, 100); break;
      case 8: tuning = clampf(clampf(v, 0.0f, 1.0f), 0.0f, 1.0f); break;
    }
  }
  void apply_modulation(uint32_t id, float v) override {  // ffi.rs:368-385: set_*(value * 100), tuning = clamp01(value)
    const float s = v * 100.0f;
    switch (id) {
      case 0: tune = clampf(s, 0, 100); break;
      case 1: bend = clampf(s, 0, 100); break;
      case 2: tone = clampf(s, 0, 100); break;
      case 3: color = clampf(s, 0, 100); break;
      case 4: decay = clampf(s, 0, 100); break;
      case 5: membrane = clampf(s, 0, 100); break;
      case 6: membrane_q = clampf(s, 0, 100); update_membrane(); break;
      case 7: volume = clampf(s, 0, 100); break;
      case 8: tuning = clampf(clampf(v, 0.0f, 1.0f), 0.0f, 1.0f); break;
    }
  }
  bool get_freq_param(float& f) const override { f = tune; return true; }
  bool is_active() const override { return active; }
  static float tri(float ph) { float t = fract(ph); return t < 0.5f ? 4.0f * t - 1.0f : 3.0f - 4.0f * t; }
  void trigger_with_velocity(double time, float) override {  // :428-448
    active = true; past_attack = false;
    morph.reset(); click.trigger(); tri_phase = 0; bp.reset();
    membrane_res.reset(); main_sound_done = false;
    float decay_ms = 0.5f + (decay / 100.0f) * (4000.0f - 0.5f);
    envelope = MaxCurveEnvelope({{1.0f, 1.0f, 0.8f}, {0.0f, decay_ms, -0.83f}});
    envelope.trigger(time);
  }
  float tick(double now) override {  // :450-585
    if (!active) return 0.0f;
    float env = envelope.get_value(now);
    if (env > 0.9f) past_attack = true;
    float tn = tune / 100.0f;
    float base_frequency = (40.0f + (tn * tn) * (600.0f - 40.0f)) * tuning_to_multiplier(tuning);
    float bend_scaled = (bend / 100.0f) * 2.0f;
    float eb = env * bend_scaled;
    float pitch_mod = eb * eb;
    float raw_freq = base_frequency * (1.0f + pitch_mod);
    bool stop = envelope.is_complete() || (past_attack && raw_freq < 20.0f);
    if (stop) main_sound_done = true;
    if (main_sound_done && !membrane_res.is_ringing()) { active = false; return 0.0f; }
    float fade = (past_attack && raw_freq < 40.0f) ? (raw_freq - 20.0f) / (40.0f - 20.0f) : 1.0f;
    float mf = rust_max(raw_freq, 40.0f);
    float click_out = click.tick() * 1.1f;
    float tri_out = triangle_enabled ? tri(tri_phase) * 0.5f : 0.0f;
    MorphOsc::advance(tri_phase, mf, sample_rate);
    float mix_control = (tone / 100.0f) * 2.0f - 1.0f;
    float color_midi = 30.0f + (color / 100.0f) * 20.0f;
    float cf1 = 440.0f * powf(2.0f, (color_midi - 69.0f) / 12.0f);
    float morph_out = morph.tick(mf, mix_control, cf1, tone);
    float mixed = click_out + tri_out + morph_out;
    float ff = rust_max(mf, 20.0f);
    float cn = color / 100.0f;
    float fq = 1.0f + cn * cn;
    bp.set_params(ff, fq, 1.1f);
    float filtered = bp.process(mixed);
    float mem_out;
    if (membrane > 0.0f) {
      float mi = main_sound_done ? 0.0f : filtered * env;
      mem_out = membrane_res.process(mi);
    } else mem_out = 0.0f;
    if (main_sound_done) {
      float mm = membrane / 100.0f;
      float fd = membrane_res.fade_multiplier();
      return mem_out * mm * fd * 0.7f * (volume / 100.0f);
    }
    float mm = membrane / 100.0f;
    float dry_gain = 1.0f - mm, wet_gain = mm;
    float dry = filtered * env;
    float fs = dry * dry_gain + mem_out * wet_gain;
    return fs * fade * 0.7f * (volume / 100.0f);
  }
};

}  // namespace orc
