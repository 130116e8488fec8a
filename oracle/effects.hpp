// oracle/effects.hpp — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the effects on the hot path: TiltFilterEffect, LowpassFilterEffect, TubeSaturation, TubeCompressor,
// DelayEffect, SpringReverbEffect, PlateReverbEffect (effects/{tilt_filter,lowpass_filter,saturation,compressor,delay,
// reverb,plate_reverb}.rs) and the per-channel waveshaper pairs of the track racks (mixer/effect_chain.rs).
#pragma once
#include "prims.hpp"

namespace orc {

struct StereoEffect {
  virtual ~StereoEffect() {}
  virtual StereoFrame process_stereo(StereoFrame in) = 0;
  virtual float process(float in) = 0;
  virtual void set_param(uint32_t, float) {}
  virtual void set_bpm(float) {}
};

// ---- effects/tilt_filter.rs:87-139 ---------------------------------------------------------------
struct TiltFilterEffect : StereoEffect {
  struct St { SmoothedParam cutoff, res; StateVariableFilterTpt svf; St(float sr) : cutoff(0.5f, 0, 1, sr, 30.0f), res(0.0f, 0, 1, sr, 30.0f), svf(sr, 1000.0f, 0.5f) {} };
  St st[2];
  float cutoff_target = 0.5f, res_target = 0.0f;
  explicit TiltFilterEffect(float sr) : st{St(sr), St(sr)} {}
  void set_param(uint32_t p, float v) override { if (p == 0) cutoff_target = clampf(v, 0, 1); else if (p == 1) res_target = clampf(v, 0, 1); }
  float one(St& s, float in) {
    s.cutoff.set_target(cutoff_target);
    s.res.set_target(res_target);
    float knob = s.cutoff.tick();
    float resonance = s.res.tick();
    float mix, freq;
    bool lp;
    if (knob < 0.5f) { mix = 1.0f - (knob * 2.0f); float t = knob * 2.0f; freq = 80.0f * powf(20000.0f / 80.0f, t); lp = true; }
    else { mix = (knob - 0.5f) * 2.0f; float t = (knob - 0.5f) * 2.0f; freq = 20.0f * powf(8000.0f / 20.0f, t); lp = false; }
    if (mix < 0.001f) return in;
    float q = 0.5f + resonance * 8.0f;
    s.svf.set_params(freq, q);
    float lo, bd, hi;
    s.svf.process_all(in, lo, bd, hi);
    float wet = lp ? lo : hi;
    float out = in * (1.0f - mix) + wet * mix;
    if (!std::isfinite(out)) { s.svf.reset(); return 0.0f; }
    if (fabsf(out) < 1e-15f) return 0.0f;
    return out;
  }
  void reset() { for (auto& s : st) s.svf.reset(); }   // tilt_filter.rs:79-84
  float process(float in) override { return one(st[0], in); }
  StereoFrame process_stereo(StereoFrame in) override { StereoFrame o; o.l = one(st[0], in.l); o.r = one(st[1], in.r); return o; }
};

// ---- effects/lowpass_filter.rs:129-194 (two-pole "Moog-style" low-pass with tanh feedback) -----------------------------
struct LowpassFilterEffect : StereoEffect {
  struct St { SmoothedParam cutoff, res; float stage1 = 0, stage2 = 0; };
  float sample_rate;
  St st[2];
  float cutoff_target, res_target;
  LowpassFilterEffect(float sr, float cutoff, float res) : sample_rate(sr) {
    cutoff = clampf(cutoff, 20.0f, 20000.0f); res = clampf(res, 0.0f, 0.95f);
    for (auto& s : st) { s.cutoff = SmoothedParam(cutoff, 20.0f, 20000.0f, sr, 30.0f); s.res = SmoothedParam(res, 0.0f, 0.95f, sr, 30.0f); }
    cutoff_target = cutoff; res_target = res;
  }
  void set_param(uint32_t p, float v) override { if (p == 0) cutoff_target = clampf(v, 20.0f, 20000.0f); else if (p == 1) res_target = clampf(v, 0.0f, 0.95f); }
  void reset() { for (auto& s : st) s.stage1 = s.stage2 = 0.0f; }
  float one(St& s, float in) {
    s.cutoff.set_target(cutoff_target); s.res.set_target(res_target);
    float cutoff = s.cutoff.tick(), resonance = s.res.tick();
    float max_cutoff = sample_rate * 0.40f;
    float safe_cutoff = rust_min(cutoff, max_cutoff);
    float nf = safe_cutoff / sample_rate;
    float g = 1.0f - expf(-2.0f * PI_F * nf);
    g = clampf(g, 0.0f, 0.90f);
    float freq_ratio = rust_min(safe_cutoff / 5000.0f, 1.0f);
    float resonance_scale = 1.0f - (freq_ratio * freq_ratio * 0.7f);
    float effective = resonance * resonance_scale;
    float feedback = effective * 3.5f;
    float fbs = s.stage2 * feedback;
    float iwf = in - tanhf(fbs) * rust_min(feedback, 1.0f);
    s.stage1 += g * (iwf - s.stage1);
    s.stage2 += g * (s.stage1 - s.stage2);
    float out = tanhf(s.stage2);
    if (fabsf(s.stage1) < 1e-15f) s.stage1 = 0.0f;
    if (fabsf(s.stage2) < 1e-15f) s.stage2 = 0.0f;
    if (!std::isfinite(out)) { s.stage1 = s.stage2 = 0.0f; return 0.0f; }
    return out;
  }
  float process(float in) override { return one(st[0], in); }
  StereoFrame process_stereo(StereoFrame in) override { StereoFrame o; o.l = one(st[0], in.l); o.r = one(st[1], in.r); return o; }
};

static const float FRAC_2_PI_F = 0.63661977236758134308f;
static inline float dc_block(float in, float& x1, float& y1) {  // saturation.rs:135-146, compressor.rs:120-130
  float out = in - x1 + 0.995f * y1;
  x1 = in;
  y1 = fabsf(out) < 1e-15f ? 0.0f : out;
  return out;
}

// ---- effects/saturation.rs:202-255 ------------------------------------------------------------------------------------
struct TubeSaturation : StereoEffect {
  struct St { SmoothedParam drive, warmth, mix; float dc_x1 = 0, dc_y1 = 0; Oversampler os; };
  St st[2];
  float drive_target, warmth_target, mix_target;
  TubeSaturation(float sr, float drive, float warmth, float mix) {
    drive = clampf(drive, 0, 1); warmth = clampf(warmth, 0, 1); mix = clampf(mix, 0, 1);
    for (auto& s : st) { s.drive = SmoothedParam(drive, 0, 1, sr, 30.0f); s.warmth = SmoothedParam(warmth, 0, 1, sr, 30.0f); s.mix = SmoothedParam(mix, 0, 1, sr, 30.0f); }
    drive_target = drive; warmth_target = warmth; mix_target = mix;
  }
  void set_param(uint32_t p, float v) override { v = clampf(v, 0, 1); if (p == 0) drive_target = v; else if (p == 1) warmth_target = v; else if (p == 2) mix_target = v; }
  void reset() { for (auto& s : st) { s.dc_x1 = s.dc_y1 = 0.0f; s.os.reset(); } }
  static float saturate(float in, float drive, float bias) {  // :104-123
    float driven = in * drive;
    float biased = driven + bias * fabsf(driven);
    float soft = atanf(biased) * FRAC_2_PI_F;
    float second = (soft * soft) * copysignf(1.0f, soft) * 0.15f;   // powi(2) * signum()
    return soft + second * bias;
  }
  float one(St& s, float in) {
    if (!std::isfinite(in)) { s.dc_x1 = s.dc_y1 = 0.0f; s.os.reset(); return 0.0f; }
    s.os.set_mode(OversamplingMode::X4);
    s.drive.set_target(drive_target); s.warmth.set_target(warmth_target); s.mix.set_target(mix_target);
    float drive = 1.0f + s.drive.tick() * 7.0f;
    float warmth = s.warmth.tick() * 0.4f;
    float mix = s.mix.tick();
    if (mix < 0.0001f) return in;
    float sat = s.os.process(in, [&](float x) { return saturate(x, drive, warmth); });
    float dcb = dc_block(sat, s.dc_x1, s.dc_y1);
    float out = in * (1.0f - mix) + dcb * mix;
    if (!std::isfinite(out)) { s.dc_x1 = s.dc_y1 = 0.0f; s.os.reset(); return 0.0f; }
    return out;
  }
  float process(float in) override { return one(st[0], in); }
  StereoFrame process_stereo(StereoFrame in) override { StereoFrame o; o.l = one(st[0], in.l); o.r = one(st[1], in.r); return o; }
};

// ---- effects/compressor.rs:133-250 ------------------------------------------------------------------------------------
struct TubeCompressor : StereoEffect {
  struct St { SmoothedParam threshold, ratio, attack, release, mix; float envelope = 0, gain = 1.0f, dc_x1 = 0, dc_y1 = 0; Oversampler os; };
  float sample_rate;
  St st[2];
  float threshold_target, ratio_target, attack_target, release_target, mix_target;
  TubeCompressor(float sr, float th, float ratio, float att, float rel, float mix) : sample_rate(sr) {
    th = clampf(th, -60.0f, 0.0f); ratio = clampf(ratio, 1.0f, 20.0f); att = clampf(att, 0.1f, 100.0f); rel = clampf(rel, 5.0f, 1000.0f); mix = clampf(mix, 0, 1);
    for (auto& s : st) {
      s.threshold = SmoothedParam(th, -60.0f, 0.0f, sr, 30.0f); s.ratio = SmoothedParam(ratio, 1.0f, 20.0f, sr, 30.0f);
      s.attack = SmoothedParam(att, 0.1f, 100.0f, sr, 30.0f); s.release = SmoothedParam(rel, 5.0f, 1000.0f, sr, 30.0f); s.mix = SmoothedParam(mix, 0, 1, sr, 30.0f);
    }
    threshold_target = th; ratio_target = ratio; attack_target = att; release_target = rel; mix_target = mix;
  }
  void set_param(uint32_t p, float v) override {
    switch (p) { case 0: threshold_target = clampf(v, -60.0f, 0.0f); break; case 1: ratio_target = clampf(v, 1.0f, 20.0f); break;
      case 2: attack_target = clampf(v, 0.1f, 100.0f); break; case 3: release_target = clampf(v, 5.0f, 1000.0f); break; case 4: mix_target = clampf(v, 0, 1); break; }
  }
  void reset() { for (auto& s : st) { s.envelope = 0.0f; s.gain = 1.0f; s.dc_x1 = s.dc_y1 = 0.0f; s.os.reset(); } }
  static float time_to_coeff(float ms, float sr) { return expf(-1.0f / (ms * 0.001f * sr)); }
  static float gain_reduction_db(float over_db, float ratio) {  // :103-117
    float slope = 1.0f - 1.0f / ratio;
    if (over_db <= -3.0f) return 0.0f;
    if (over_db >= 3.0f) return over_db * slope;
    float x = over_db + 3.0f;
    return x * x / (2.0f * 6.0f) * slope;
  }
  float inner(St& s, float in, float sidechain) {
    if (!std::isfinite(in) || !std::isfinite(sidechain)) return 0.0f;
    s.threshold.set_target(threshold_target); s.ratio.set_target(ratio_target); s.attack.set_target(attack_target);
    s.release.set_target(release_target); s.mix.set_target(mix_target);
    float threshold_db = s.threshold.tick(), ratio = s.ratio.tick(), attack_ms = s.attack.tick(), release_ms = s.release.tick(), mix = s.mix.tick();
    if (mix < 0.0001f) return in;
    float sc = fabsf(sidechain);
    float coeff = sc > s.envelope ? time_to_coeff(attack_ms, sample_rate) : time_to_coeff(release_ms, sample_rate);
    s.envelope = coeff * s.envelope + (1.0f - coeff) * sc;
    if (s.envelope < 1e-15f) s.envelope = 0.0f;
    float env_db = 20.0f * log10f(s.envelope + 1e-20f);
    float over_db = env_db - threshold_db;
    float gr = gain_reduction_db(over_db, ratio);
    float gain_linear = powf(10.0f, -gr * 0.05f);
    s.gain += 0.05f * (gain_linear - s.gain);
    float compressed = in * s.gain;
    s.os.set_mode(OversamplingMode::X4);
    float colored_os = s.os.process(compressed, [](float x) { return atanf(x) * FRAC_2_PI_F * 1.1f; });
    float colored = s.gain < 0.99f ? colored_os : compressed;
    float dcb = dc_block(colored, s.dc_x1, s.dc_y1);
    float out = in * (1.0f - mix) + dcb * mix;
    if (!std::isfinite(out)) { s.dc_x1 = s.dc_y1 = 0.0f; s.envelope = 0.0f; s.gain = 1.0f; return 0.0f; }
    return out;
  }
  float process(float in) override { return inner(st[0], in, in); }
  StereoFrame process_stereo(StereoFrame in) override { StereoFrame o; o.l = inner(st[0], in.l, in.l); o.r = inner(st[1], in.r, in.r); return o; }
  StereoFrame process_stereo_with_sidechain(StereoFrame in, StereoFrame sc) { StereoFrame o; o.l = inner(st[0], in.l, sc.l); o.r = inner(st[1], in.r, sc.r); return o; }
};

// Rack variants of the two waveshapers: one instance per channel (effect_chain.rs:49-50, 141-156)
struct WaveshaperPair : StereoEffect {
  Waveshaper ws[2];
  WaveshaperPair() : ws{Waveshaper(1.0f, 0.0f), Waveshaper(1.0f, 0.0f)} {}
  void set_param(uint32_t p, float v) override { for (auto& w : ws) { if (p == 0) w.set_drive(v); else if (p == 1) w.set_mix(v); } }
  float process(float in) override { return ws[0].process(in); }
  StereoFrame process_stereo(StereoFrame in) override { StereoFrame o; o.l = ws[0].process(in.l); o.r = ws[1].process(in.r); return o; }
};
struct FeedbackWaveshaperPair : StereoEffect {
  FeedbackWaveshaper fb[2];
  explicit FeedbackWaveshaperPair(float sr) : fb{FeedbackWaveshaper(sr, 1.0f, 0.0f, 2000.0f, 0.0f), FeedbackWaveshaper(sr, 1.0f, 0.0f, 2000.0f, 0.0f)} {}
  void set_param(uint32_t p, float v) override {
    for (auto& w : fb) { if (p == 0) w.set_drive(v); else if (p == 1) w.set_feedback(v); else if (p == 2) w.set_filter_cutoff(v); else if (p == 3) w.set_mix(v); }
  }
  float process(float in) override { return fb[0].process(in); }
  StereoFrame process_stereo(StereoFrame in) override { StereoFrame o; o.l = fb[0].process(in.l); o.r = fb[1].process(in.r); return o; }
};

// ---- effects/delay.rs ---------------------------------------------------------------------------------
static inline float delay_beats(uint32_t t) {
  switch (t) { case 0: return 4.0f; case 1: return 2.0f; case 2: return 1.0f; case 3: return 0.5f; case 4: return 0.25f;
    case 5: return 4.0f / 3.0f; case 6: return 2.0f / 3.0f; case 7: return 1.0f / 3.0f; case 8: return 1.0f / 6.0f; default: return 1.0f; }
}
static inline float delay_seconds(uint32_t timing, float bpm) { float spb = 60.0f / bpm; return rust_min(spb * delay_beats(timing), 5.0f); }

struct DelayEffect : StereoEffect {
  struct St {
    std::vector<float> buffer; size_t write_index = 0; float z1 = 0, z2 = 0; uint32_t previous_timing;
    SmoothedParam time, feedback, mix, cutoff;
  };
  float sample_rate;
  St st[2];
  uint32_t timing_target; float bpm_target, feedback_target, mix_target, cutoff_target; bool pingpong = false;
  DelayEffect(float sr, uint32_t timing, float bpm, float fb, float mix, float cutoff) : sample_rate(sr) {
    float time = delay_seconds(timing, bpm);
    float fbc = clampf(fb, 0.0f, 0.95f), mc = clampf(mix, 0.0f, 1.0f), cc = clampf(cutoff, 20.0f, 20000.0f);
    size_t n = (size_t)(sr * 5.0f) + 1;
    for (auto& s : st) {
      s.buffer.assign(n, 0.0f); s.previous_timing = timing;
      s.time = SmoothedParam(time, 0.0f, 5.0f, sr, 50.0f); s.feedback = SmoothedParam(fbc, 0.0f, 0.95f, sr, 30.0f);
      s.mix = SmoothedParam(mc, 0.0f, 1.0f, sr, 30.0f); s.cutoff = SmoothedParam(cc, 20.0f, 20000.0f, sr, 30.0f);
    }
    timing_target = timing; bpm_target = bpm; feedback_target = fbc; mix_target = mc; cutoff_target = cc;
  }
  void set_bpm(float b) override { bpm_target = b; }
  void reset() { for (auto& s : st) { std::fill(s.buffer.begin(), s.buffer.end(), 0.0f); s.write_index = 0; s.z1 = s.z2 = 0.0f; } }   // delay.rs:229-238
  void set_param(uint32_t p, float v) override {  // ffi.rs:3006-3017
    switch (p) {
      case 0: { uint32_t t = (uint32_t)f32_as_u64(v) ; if (f32_as_u64(v) > 0xffffffffull) t = 0xffffffffu; if (t <= 8) timing_target = t; } break;
      case 1: feedback_target = clampf(v, 0.0f, 0.95f); break;
      case 2: mix_target = clampf(v, 0.0f, 1.0f); break;
      case 3: cutoff_target = clampf(v, 20.0f, 20000.0f); break;
      case 4: pingpong = v >= 0.5f; break;
    }
  }
  struct Step { float filtered, feedback, mix; };
  Step step_read(St& s) {  // :321-399
    uint32_t tc = timing_target;
    float time_target = delay_seconds(tc <= 8 ? tc : 2, bpm_target);
    if (tc != s.previous_timing) {
      s.previous_timing = tc;
      std::fill(s.buffer.begin(), s.buffer.end(), 0.0f);
      s.z1 = s.z2 = 0.0f;
      s.time.set_immediate(time_target);
    }
    s.time.set_target(time_target); s.feedback.set_target(feedback_target); s.mix.set_target(mix_target); s.cutoff.set_target(cutoff_target);
    float time = s.time.tick(), feedback = s.feedback.tick(), mix = s.mix.tick(), cutoff = s.cutoff.tick();
    float ds = time * sample_rate;
    size_t di = (size_t)f32_as_u64(ds);
    float df = ds - (float)di;
    size_t len = s.buffer.size();
    size_t r1 = (s.write_index + len - di) % len;
    size_t r2 = (s.write_index + len - di - 1) % len;
    float s1 = s.buffer[r1], s2 = s.buffer[r2];
    float delayed = s1 * (1.0f - df) + s2 * df;
    float g = 1.0f - expf(-2.0f * PI_F * cutoff / sample_rate);
    float resonance = 0.3f;
    float rfb = resonance * (s.z1 - s.z2);
    s.z1 = s.z1 + g * (delayed + rfb - s.z1);
    s.z2 = s.z2 + g * (s.z1 - s.z2);
    float filtered = s.z2;
    if (fabsf(s.z1) < 1e-15f) s.z1 = 0.0f;
    if (fabsf(s.z2) < 1e-15f) s.z2 = 0.0f;
    return {filtered, feedback, mix};
  }
  float step_write(St& s, float dry, float inject, const Step& st_, float tap) {  // :407-439
    size_t len = s.buffer.size();
    float w = inject + tap * st_.feedback;
    w = (std::isfinite(w) && fabsf(w) > 1e-15f) ? w : 0.0f;
    s.buffer[s.write_index] = w;
    s.write_index = (s.write_index + 1) % len;
    float out = dry * (1.0f - st_.mix) + st_.filtered * st_.mix;
    if (!std::isfinite(out)) return dry;
    return out;
  }
  float one(St& s, float in) { in = std::isfinite(in) ? in : 0.0f; Step t = step_read(s); return step_write(s, in, in, t, t.filtered); }
  float process(float in) override { return one(st[0], in); }
  StereoFrame process_stereo(StereoFrame in) override {  // :460-491
    StereoFrame o;
    if (!pingpong) { o.l = one(st[0], in.l); o.r = one(st[1], in.r); return o; }
    float l = std::isfinite(in.l) ? in.l : 0.0f, r = std::isfinite(in.r) ? in.r : 0.0f;
    Step sl = step_read(st[0]), sr_ = step_read(st[1]);
    o.l = step_write(st[0], l, l, sl, sr_.filtered);
    o.r = step_write(st[1], r, 0.0f, sr_, sl.filtered);
    return o;
  }
};

// ---- effects/reverb.rs (spring) ------------------------------------------------------------------------------
struct SpringReverbEffect : StereoEffect {
  struct AP { std::vector<float> buf; size_t idx = 0; float process(float in, float g) { float d = buf[idx]; float v = in - g * d; float o = g * v + d; buf[idx] = v; idx = (idx + 1) % buf.size(); return o; } };
  struct St { AP ap[6]; float fb = 0, damp = 0; SmoothedParam decay, mix, damping; };
  St st[2];
  float decay_target, mix_target, damping_target;
  SpringReverbEffect(float sr, float decay, float mix, float damping) {
    decay = clampf(decay, 0, 1); mix = clampf(mix, 0, 1); damping = clampf(damping, 0, 1);
    const size_t DL[6] = {131, 251, 389, 521, 617, 787}, DR[6] = {127, 263, 397, 541, 631, 797};
    float scale = sr / 44100.0f;
    for (int c = 0; c < 2; c++) {
      for (int i = 0; i < 6; i++) { size_t len = (size_t)f32_as_u64(rust_max((float)(c == 0 ? DL[i] : DR[i]) * scale, 1.0f)); st[c].ap[i].buf.assign(len, 0.0f); }
      st[c].decay = SmoothedParam(clampf(decay, 0, 1), 0, 1, sr, 15.0f); st[c].mix = SmoothedParam(clampf(mix, 0, 1), 0, 1, sr, 15.0f);
      st[c].damping = SmoothedParam(clampf(damping, 0, 1), 0, 1, sr, 15.0f);
    }
    decay_target = decay; mix_target = mix; damping_target = damping;
  }
  void set_param(uint32_t p, float v) override { v = clampf(v, 0, 1); if (p == 0) decay_target = v; else if (p == 1) mix_target = v; else if (p == 2) damping_target = v; }
  void reset() { for (auto& s : st) { for (auto& a : s.ap) { std::fill(a.buf.begin(), a.buf.end(), 0.0f); a.idx = 0; } s.fb = s.damp = 0.0f; } }   // reverb.rs:148-159
  float one(St& s, float in) {  // :162-217
    const float G[6] = {0.70f, 0.68f, 0.65f, 0.62f, 0.60f, 0.58f};
    in = std::isfinite(in) ? in : 0.0f;
    s.decay.set_target(decay_target); s.mix.set_target(mix_target); s.damping.set_target(damping_target);
    float decay = s.decay.tick(), mix = s.mix.tick(), damping = s.damping.tick();
    float feedback = powf(decay, 0.4f) * 0.95f;
    float d1 = damping, d2 = 1.0f - damping;
    float sig = in + s.fb;
    for (int i = 0; i < 6; i++) sig = s.ap[i].process(sig, G[i]);
    s.damp = sig * d2 + s.damp * d1;
    if (fabsf(s.damp) < 1e-15f) s.damp = 0.0f;
    s.fb = s.damp * feedback;
    if (fabsf(s.fb) < 1e-15f) s.fb = 0.0f;
    float res = in * (1.0f - mix) + sig * mix;
    return std::isfinite(res) ? res : in;
  }
  float process(float in) override { return one(st[0], in); }
  StereoFrame process_stereo(StereoFrame in) override { StereoFrame o; o.l = one(st[0], in.l); o.r = one(st[1], in.r); return o; }
};

// ---- effects/plate_reverb.rs (Dattorro) ------------------------------------------------------------------------
struct PlateReverbEffect : StereoEffect {
  struct DL {
    std::vector<float> buf; size_t idx = 0;
    void init(size_t cap) { buf.assign(cap < 4 ? 4 : cap, 0.0f); idx = 0; }
    void write(float x) { buf[idx] = x; idx = (idx + 1) % buf.size(); }
    float read_frac(float off) const {
      size_t len = buf.size();
      off = clampf(off, 1.0f, (float)(len - 2));
      size_t w = (size_t)f32_as_u64(off);
      float fr = off - (float)w;
      float a = buf[(idx + len - w) % len], b = buf[(idx + len - w - 1) % len];
      return a + fr * (b - a);
    }
    float tap_frac(float off) const {
      size_t len = buf.size();
      off = clampf(off, 0.0f, (float)(len - 2));
      size_t w = (size_t)f32_as_u64(off);
      float fr = off - (float)w;
      float a = buf[(idx + len - 1 - w) % len], b = buf[(idx + len - 2 - w) % len];
      return a + fr * (b - a);
    }
    float allpass(float in, float g, float d) { float dl = read_frac(d); float v = in - g * dl; write(v); return g * v + dl; }
  };
  DL predelay, in_ap[4], mod_ap_a, delay1_a, ap2_a, delay2_a, mod_ap_b, delay1_b, ap2_b, delay2_b;
  float in_ap_delay[4];
  float bandwidth_state = 0, damp_a = 0, damp_b = 0, fb_a = 0, fb_b = 0, lfo_pa = 0, lfo_pb = 0, lfo_ia, lfo_ib;
  float len_ap1_a, len_d1_a, len_ap2_a, len_d2_a, len_ap1_b, len_d1_b, len_ap2_b, len_d2_b, excursion, sr_scale, sample_rate;
  SmoothedParam decay_s, mix_s, damping_s, predelay_s, width_s, size_s;
  float decay_t, mix_t, damping_t, predelay_t = 0.0f, width_t = 1.0f, size_t_ = 0.5f;
  static float size_to_scale(float s) { return s <= 0.5f ? powf(4.0f, 2.0f * s - 1.0f) : powf(2.0f, 2.0f * s - 1.0f); }
  PlateReverbEffect(float sr, float decay, float mix, float damping) : sample_rate(sr) {
    decay = clampf(decay, 0, 1); mix = clampf(mix, 0, 1); damping = clampf(damping, 0, 1);
    sr_scale = sr / 29761.0f;
    excursion = 16.0f * sr_scale;
    auto fixed = [&](float base) { return (size_t)f32_as_u64(ceilf(base * sr_scale)) + 4; };
    auto sized = [&](float base, float head) { return (size_t)f32_as_u64(ceilf(base * 2.0f * sr_scale + head)) + 4; };
    const float IAD[4] = {142.0f, 107.0f, 379.0f, 277.0f};
    predelay.init((size_t)f32_as_u64(ceilf(200.0f * 0.001f * sr)) + 8);
    for (int i = 0; i < 4; i++) { in_ap[i].init(fixed(IAD[i])); in_ap_delay[i] = rust_max(IAD[i] * sr_scale, 1.0f); }
    mod_ap_a.init(sized(672.0f, excursion)); delay1_a.init(sized(4453.0f, 0)); ap2_a.init(sized(1800.0f, 0)); delay2_a.init(sized(3720.0f, 0));
    mod_ap_b.init(sized(908.0f, excursion)); delay1_b.init(sized(4217.0f, 0)); ap2_b.init(sized(2656.0f, 0)); delay2_b.init(sized(3163.0f, 0));
    lfo_ia = 0.50f / sr; lfo_ib = 0.71f / sr;
    len_ap1_a = 672.0f * sr_scale; len_d1_a = 4453.0f * sr_scale; len_ap2_a = 1800.0f * sr_scale; len_d2_a = 3720.0f * sr_scale;
    len_ap1_b = 908.0f * sr_scale; len_d1_b = 4217.0f * sr_scale; len_ap2_b = 2656.0f * sr_scale; len_d2_b = 3163.0f * sr_scale;
    decay_s = SmoothedParam(decay, 0, 1, sr, 15.0f); mix_s = SmoothedParam(mix, 0, 1, sr, 15.0f); damping_s = SmoothedParam(damping, 0, 1, sr, 15.0f);
    predelay_s = SmoothedParam(0.0f, 0, 1, sr, 15.0f); width_s = SmoothedParam(1.0f, 0, 1, sr, 15.0f); size_s = SmoothedParam(0.5f, 0, 1, sr, 15.0f);
    decay_t = decay; mix_t = mix; damping_t = damping;
  }
  void set_param(uint32_t p, float v) override {
    v = clampf(v, 0, 1);
    switch (p) { case 0: decay_t = v; break; case 1: mix_t = v; break; case 2: damping_t = v; break; case 3: predelay_t = v; break; case 4: width_t = v; break; case 5: size_t_ = v; break; }
  }
  void reset() {   // plate_reverb.rs:380-402
    DL* all[13] = {&predelay, &in_ap[0], &in_ap[1], &in_ap[2], &in_ap[3], &mod_ap_a, &delay1_a, &ap2_a, &delay2_a, &mod_ap_b, &delay1_b, &ap2_b, &delay2_b};
    for (DL* d : all) { std::fill(d->buf.begin(), d->buf.end(), 0.0f); d->idx = 0; }
    bandwidth_state = damp_a = damp_b = fb_a = fb_b = lfo_pa = lfo_pb = 0.0f;
  }
  static void flush(float& x) { if (fabsf(x) < 1e-15f) x = 0.0f; }
  void tick_tank(float in, float& wl, float& wr, float& mix) {  // :406-534
    in = std::isfinite(in) ? in : 0.0f;
    decay_s.set_target(decay_t); mix_s.set_target(mix_t); damping_s.set_target(damping_t);
    predelay_s.set_target(predelay_t); width_s.set_target(width_t); size_s.set_target(size_t_);
    float decay_knob = decay_s.tick(); mix = mix_s.tick(); float damping = damping_s.tick();
    float predelay_knob = predelay_s.tick(); float width = width_s.tick(); float size = size_to_scale(size_s.tick());
    float decay_gain = decay_knob * 0.95f;
    float dd2 = clampf(decay_gain + 0.15f, 0.25f, 0.50f);
    float damp = damping * 0.95f;
    predelay.write(in);
    float pds = predelay_knob * 200.0f * 0.001f * sample_rate;
    float delayed = predelay.tap_frac(pds);
    bandwidth_state += 0.9995f * (delayed - bandwidth_state);
    flush(bandwidth_state);
    float sig = bandwidth_state;
    const float IAG[4] = {0.750f, 0.750f, 0.625f, 0.625f};
    for (int i = 0; i < 4; i++) sig = in_ap[i].allpass(sig, IAG[i], in_ap_delay[i]);
    lfo_pa = fract(lfo_pa + lfo_ia); lfo_pb = fract(lfo_pb + lfo_ib);
    const float TAU_F = 6.28318530717958647692f;
    float lfo_a = sinf(TAU_F * lfo_pa), lfo_b = sinf(TAU_F * lfo_pb);
    float in_a = sig + fb_b, in_b = sig + fb_a;
    float a1 = mod_ap_a.allpass(in_a, 0.70f, len_ap1_a * size + lfo_a * excursion);
    float d1a = delay1_a.read_frac(len_d1_a * size);
    delay1_a.write(a1);
    damp_a = d1a * (1.0f - damp) + damp_a * damp; flush(damp_a);
    float a2 = ap2_a.allpass(damp_a * decay_gain, dd2, len_ap2_a * size);
    float d2a = delay2_a.read_frac(len_d2_a * size);
    delay2_a.write(a2);
    float b1 = mod_ap_b.allpass(in_b, 0.70f, len_ap1_b * size + lfo_b * excursion);
    float d1b = delay1_b.read_frac(len_d1_b * size);
    delay1_b.write(b1);
    damp_b = d1b * (1.0f - damp) + damp_b * damp; flush(damp_b);
    float b2 = ap2_b.allpass(damp_b * decay_gain, dd2, len_ap2_b * size);
    float d2b = delay2_b.read_frac(len_d2_b * size);
    delay2_b.write(b2);
    fb_a = d2a * decay_gain; flush(fb_a);
    fb_b = d2b * decay_gain; flush(fb_b);
    float ts = sr_scale * size;
    float yl = 0.6f * (delay1_b.tap_frac(266.0f * ts) + delay1_b.tap_frac(2974.0f * ts) - ap2_b.tap_frac(1913.0f * ts) + delay2_b.tap_frac(1996.0f * ts)
                       - delay1_a.tap_frac(1990.0f * ts) - ap2_a.tap_frac(187.0f * ts) - delay2_a.tap_frac(1066.0f * ts));
    float yr = 0.6f * (delay1_a.tap_frac(353.0f * ts) + delay1_a.tap_frac(3627.0f * ts) - ap2_a.tap_frac(1228.0f * ts) + delay2_a.tap_frac(2673.0f * ts)
                       - delay1_b.tap_frac(2111.0f * ts) - ap2_b.tap_frac(335.0f * ts) - delay2_b.tap_frac(121.0f * ts));
    float mid = 0.5f * (yl + yr);
    float side = 0.5f * (yl - yr) * width;
    wl = mid + side; wr = mid - side;
  }
  float process(float in) override {
    in = std::isfinite(in) ? in : 0.0f;
    float wl, wr, mix;
    tick_tank(in, wl, wr, mix);
    float r = in * (1.0f - mix) + 0.5f * (wl + wr) * mix;
    return std::isfinite(r) ? r : in;
  }
  StereoFrame process_stereo(StereoFrame in) override {
    float l = std::isfinite(in.l) ? in.l : 0.0f, r = std::isfinite(in.r) ? in.r : 0.0f;
    float wl, wr, mix;
    tick_tank(0.5f * (l + r), wl, wr, mix);
    float ol = l * (1.0f - mix) + wl * mix, orr = r * (1.0f - mix) + wr * mix;
    StereoFrame o; o.l = std::isfinite(ol) ? ol : l; o.r = std::isfinite(orr) ? orr : r;
    return o;
  }
};

}  // namespace orc
