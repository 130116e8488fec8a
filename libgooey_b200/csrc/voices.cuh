// voices.cuh — per-voice state, trigger and per-sample tick of the drum voices
// (kick, snare, hi-hat, tom), one voice per thread.
//
// Each voice is {State (persistent across launches), init(), on_event(), tick()}.
// Parameters are smoothed on the device (cur/tgt pairs); the host only ships
// set-target / snap / trigger events (events.h).  Reference: src/instruments/
// {kick,snare,hihat2,tom2}.rs; citations inline.
#pragma once
#include "dsp.cuh"
#include "events.h"

namespace gd {

// Per-launch constants derived from the sample rate (uniform across a batch).
struct RateCtx {
  float sr;
  double dt;          // 1.0 / (sr as f64)  (bounce.rs:46, ffi.rs:1096)
  float smooth15;     // SmoothedParam coeff, 15 ms (DEFAULT_SMOOTH_TIME_MS)
  float smooth10, smooth30, smooth50;
  float click_alpha;  // kick click HP: 1 - exp(-2pi*8000/sr) (resonant_highpass.rs:44-45)
  float asym_down;    // hihat AsymmetricSmoother(100 samples) (hihat2.rs:295-305)
  PinkCoef pink;
};
G_HD RateCtx make_rate_ctx(float sr) {
  RateCtx c;
  c.sr = sr;
  c.dt = 1.0 / (double)sr;
  c.smooth15 = smooth_coeff(sr, 15.0f);
  c.smooth10 = smooth_coeff(sr, 10.0f);
  c.smooth30 = smooth_coeff(sr, 30.0f);
  c.smooth50 = smooth_coeff(sr, 50.0f);
  c.click_alpha = 1.0f - gm::g_expf(-2.0f * PI_F * 8000.0f / sr);
  c.asym_down = 1.0f - gm::g_expf(-1.0f / 100.0f);
  c.pink = pink_coefs(sr);
  return c;
}

// =========================================== Kick ===========================================
enum { K_FREQ, K_PUNCH, K_SUB, K_CLICK, K_OSC_DECAY, K_PITCH_ENV_AMT, K_PITCH_ENV_CURVE, K_VOLUME,
       K_PITCH_START_RATIO, K_PHASE_MOD, K_NOISE_AMT, K_NOISE_CUTOFF, K_NOISE_RES, K_OVERDRIVE,
       K_FEEDBACK, K_FB_CUTOFF, K_AMP_DECAY, K_AMP_DECAY_CURVE, K_TUNING, K_NP };

struct KickState {
  float cur[K_NP], tgt[K_NP];
  Env sub_env, punch_env, click_env, pitch_env, noise_env, amp_env;
  float tpm;          // triggered_pitch_multiplier
  float click_hp;     // ResonantHighpassFilter.filter_state
  PhaseMod pm;
  Pink pink;
  Tpt noise_lp;
  FbShaper ws;
  float velocity;
  uint32_t active;
  float saved_freq; uint32_t has_saved;   // VoiceStrip.saved_global_freq (ffi.rs:596, 1176-1194)
  double t;
};

G_HD float overdrive_to_drive(float a) { return 1.0f + a * a * a * 40.0f; }

// KickDrum::with_config (kick.rs:712-775) + configure_oscillators (:777-817); cfg = 18 normalized values
G_HD void kick_init(KickState& s, const float* cfg, float sr) {
  for (int i = 0; i < 18; i++) { float c = clampf(cfg[i], 0.0f, 1.0f); s.cur[i] = s.tgt[i] = c; }
  s.cur[K_TUNING] = s.tgt[K_TUNING] = 0.5f;
  float ratio = denorm(s.cur[K_PITCH_START_RATIO], 1.0f, 10.0f);
  s.tpm = 1.0f + (ratio - 1.0f) * s.cur[K_PITCH_ENV_AMT];
  float decay = denorm(s.cur[K_OSC_DECAY], 0.01f, 4.0f);
  env_init(s.sub_env); env_init(s.punch_env); env_init(s.click_env); env_init(s.pitch_env); env_init(s.noise_env); env_init(s.amp_env);
  env_config(s.sub_env, 0.001f, decay, 0.0f, decay * 0.2f);
  env_config(s.punch_env, 0.001f, decay, 0.0f, decay * 0.2f);
  env_config(s.click_env, 0.001f, decay * 0.2f, 0.0f, decay * 0.02f);
  float pd = decay * 0.6f;
  env_config(s.pitch_env, 0.001f, pd, 0.0f, pd * 0.1f);
  env_config(s.noise_env, 0.001f, decay, 0.0f, decay * 0.2f);
  s.click_hp = 0.0f;
  s.pm.trig = 0.0; s.pm.active = 0;
  pink_reset(s.pink);
  rlp_init(s.noise_lp, sr, denorm(s.cur[K_NOISE_CUTOFF], 20.0f, 10000.0f), denorm(s.cur[K_NOISE_RES], 0.0f, 5.0f));
  fbws_init(s.ws, sr, overdrive_to_drive(s.cur[K_OVERDRIVE]), s.cur[K_FEEDBACK] * 0.98f, 200.0f + s.cur[K_FB_CUTOFF] * 3800.0f, 1.0f);
  s.velocity = 1.0f; s.active = 0; s.saved_freq = 0.0f; s.has_saved = 0; s.t = 0.0;
}

// KickDrum::trigger_with_velocity (kick.rs:971-1086)
G_HD void kick_trigger(KickState& s, float velocity) {
  double time = s.t;
  s.velocity = clampf(velocity, 0.0f, 1.0f);
  s.active = 1;
  float vel = s.velocity, vel2 = vel * vel;
  float decay_scale = 1.0f - (0.5f * vel2);
  float base_decay = denorm(s.cur[K_OSC_DECAY], 0.01f, 4.0f) * decay_scale;
  float psr = denorm(s.cur[K_PITCH_START_RATIO], 1.0f, 10.0f);
  s.tpm = 1.0f + (psr - 1.0f) * s.cur[K_PITCH_ENV_AMT];
  float pcv = denorm(s.cur[K_PITCH_ENV_CURVE], 0.1f, 4.0f);
  float dc = fabsf(pcv - 1.0f) < 0.01f ? CURVE_LINEAR : pcv;
  env_config(s.pitch_env, 0.001f, base_decay, 0.0f, base_decay * 0.2f, CURVE_LINEAR, dc);
  env_config(s.sub_env, 0.001f, base_decay, 0.0f, base_decay * 0.2f);
  env_config(s.punch_env, 0.001f, base_decay, 0.0f, base_decay * 0.2f);
  env_config(s.click_env, 0.001f, base_decay * 0.2f, 0.0f, base_decay * 0.02f);
  env_trigger(s.sub_env, time); env_trigger(s.punch_env, time); env_trigger(s.click_env, time);
  env_trigger(s.pitch_env, time);
  if (s.cur[K_PHASE_MOD] > 0.001f) { s.pm.trig = time; s.pm.active = 1; }
  env_config(s.noise_env, 0.001f, base_decay, 0.0f, base_decay * 0.2f);
  env_trigger(s.noise_env, time);
  float amp_decay = denorm(s.cur[K_AMP_DECAY], 0.0f, 4.0f) * decay_scale;
  float adc = denorm(s.cur[K_AMP_DECAY_CURVE], 0.1f, 10.0f);
  float adcv = fabsf(adc - 1.0f) < 0.01f ? CURVE_LINEAR : adc;
  env_config(s.amp_env, 0.001f, amp_decay, 0.0f, amp_decay * 0.2f, 0.5f, adcv);
  env_trigger(s.amp_env, time);
  s.click_hp = 0.0f;
  s.noise_lp.ic1 = s.noise_lp.ic2 = 0.0f;
  pink_reset(s.pink);
}

G_HD void kick_event(KickState& s, const VoiceEvent& e) {
  switch (e.kind) {
    case EV_TRIGGER: kick_trigger(s, e.value); break;
    case EV_SET_TARGET: if (e.param < K_NP) { float c = clampf(e.value, 0.0f, 1.0f); if (fabsf(s.tgt[e.param] - c) > 1e-8f) s.tgt[e.param] = c; } break;
    case EV_SNAP: for (int i = 0; i < K_NP; i++) s.cur[i] = s.tgt[i]; break;
    case EV_SET_AUX: if (e.param == AUX_OVERSAMPLING) { uint32_t m = (uint32_t)e.value; if (s.ws.os.mode != m) { s.ws.os.mode = m; os_reset(s.ws.os); } } break;
    case EV_NOTE_FREQ: {
      if (!s.has_saved) { s.saved_freq = s.cur[K_FREQ]; s.has_saved = 1; }
      float c = clampf(e.value, 0.0f, 1.0f);
      if (fabsf(s.tgt[K_FREQ] - c) > 1e-8f) s.tgt[K_FREQ] = c;
      for (int i = 0; i < K_NP; i++) s.cur[i] = s.tgt[i];
    } break;
    case EV_RESTORE_FREQ:
      if (s.has_saved) {
        s.has_saved = 0;
        float c = clampf(s.saved_freq, 0.0f, 1.0f);
        if (fabsf(s.tgt[K_FREQ] - c) > 1e-8f) s.tgt[K_FREQ] = c;
        for (int i = 0; i < K_NP; i++) s.cur[i] = s.tgt[i];
      }
      break;
    default: break;
  }
}

// KickDrum::tick (kick.rs:1097-1232)
G_D float kick_tick(KickState& s, const RateCtx& rc) {
  const double now = s.t;
  s.t = now + rc.dt;
#pragma unroll
  for (int i = 0; i < K_NP; i++) smooth_tick(s.cur[i], s.tgt[i], rc.smooth15);
  if (!s.active) return 0.0f;
  const float sr = rc.sr;
  // apply_params (:820-835)
  float cvs = 0.6f + 0.4f * s.velocity;
  float sub_vol = clampf(s.cur[K_SUB], 0.0f, 1.0f);
  float punch_vol = clampf(s.cur[K_PUNCH] * 0.7f, 0.0f, 1.0f);
  float click_vol = clampf(s.cur[K_CLICK] * 0.15f * cvs, 0.0f, 1.0f);
  // live decay re-application (:1111-1136), no floors
  float vel2 = s.velocity * s.velocity;
  float decay_scale = 1.0f - (0.5f * vel2);
  float base_decay = denorm(s.cur[K_OSC_DECAY], 0.01f, 4.0f) * decay_scale;
  s.sub_env.decay = base_decay; s.sub_env.release = base_decay * 0.2f;
  s.punch_env.decay = base_decay; s.punch_env.release = base_decay * 0.2f;
  s.click_env.decay = base_decay * 0.2f; s.click_env.release = base_decay * 0.02f;
  s.noise_env.decay = base_decay; s.noise_env.release = base_decay * 0.2f;
  s.pitch_env.decay = base_decay; s.pitch_env.release = base_decay * 0.2f;
  float base_frequency = denorm(s.cur[K_FREQ], 30.0f, 120.0f) * tuning_to_multiplier(s.cur[K_TUNING]);
  float pev = env_amp(s.pitch_env, now);
  float fm = 1.0f + (s.tpm - 1.0f) * pev;
  float pma = s.cur[K_PHASE_MOD];
  if (pma > 0.001f) {
    float pm = phasemod_tick(s.pm, now);
    fm *= 1.0f + (pm * pma * 2.0f);
  }
  float sub_f = base_frequency * fm;
  float punch_f = base_frequency * 2.5f * fm;
  // Oscillator::tick x3 (oscillator.rs:242-286).  An oscillator whose envelope amplitude or volume is exactly 0
  // contributes raw*0 = +-0, so its waveform is skipped (raw is always finite).
  float sub_out = 0.0f, punch_out = 0.0f, raw_click = 0.0f;
  {
    float idx = env_active(s.sub_env) ? (float)(now - s.sub_env.trig) * sr : 0.0f;
    float amp = env_amp(s.sub_env, now);
    if (amp != 0.0f && sub_vol != 0.0f) sub_out = osc_sine(idx, sub_f, sr) * amp * sub_vol;
  }
  {
    float idx = env_active(s.punch_env) ? (float)(now - s.punch_env.trig) * sr : 0.0f;
    float amp = env_amp(s.punch_env, now);
    if (amp != 0.0f && punch_vol != 0.0f) punch_out = osc_triangle(idx, punch_f, sr) * amp * punch_vol;
  }
  {
    float idx = env_active(s.click_env) ? (float)(now - s.click_env.trig) * sr : 0.0f;
    float amp = env_amp(s.click_env, now);
    if (amp != 0.0f && click_vol != 0.0f) raw_click = hash_noise(f32_to_u64_sat(idx)) * amp * click_vol;
  }
  // click HP (resonant_highpass.rs:22-54), resonance 4.0
  float hp = raw_click - s.click_hp;
  s.click_hp += rc.click_alpha * hp;
  float filt_click = hp * (1.0f + 4.0f * 0.1f);
  float noise_amount = s.cur[K_NOISE_AMT];
  float noise_out = 0.0f;
  if (noise_amount > 0.001f) {
    float pn = pink_tick(s.pink, rc.pink);
    rlp_set(s.noise_lp, sr, denorm(s.cur[K_NOISE_CUTOFF], 20.0f, 10000.0f), denorm(s.cur[K_NOISE_RES], 0.0f, 5.0f));
    float fn = rlp_process(s.noise_lp, pn);
    float ne = env_amp(s.noise_env, now);
    noise_out = fn * ne * noise_amount * 0.5f;
  }
  float total = sub_out + punch_out + filt_click + noise_out;
  s.ws.drive = clampf(overdrive_to_drive(s.cur[K_OVERDRIVE]), 1.0f, 100.0f);
  s.ws.feedback = clampf(s.cur[K_FEEDBACK] * 0.98f, 0.0f, 0.98f);
  fbws_set_cutoff(s.ws, sr, 200.0f + s.cur[K_FB_CUTOFF] * 3800.0f);
  float od = fbws_process(s.ws, total);
  float amp_env = env_amp(s.amp_env, now);
  float va = sqrtf(s.velocity);
  float out = od * amp_env * va * s.cur[K_VOLUME];
  if (!env_active(s.amp_env)) s.active = 0;
  return out;
}

// =========================================== Snare ===========================================
enum { S_FREQ, S_DECAY, S_BRIGHTNESS, S_VOLUME, S_TONAL, S_NOISE, S_PITCH_DROP, S_TONAL_DECAY, S_TONAL_DECAY_CURVE,
       S_NOISE_DECAY, S_NOISE_TAIL_DECAY, S_FILTER_CUTOFF, S_FILTER_RES, S_XFADE, S_PHASE_MOD, S_OVERDRIVE,
       S_AMP_DECAY, S_AMP_DECAY_CURVE, S_TUNING, S_NP };

struct SnareState {
  float cur[S_NP], tgt[S_NP];
  uint32_t filter_type;
  Env tonal_osc_env, noise_osc_env, crack_env, pitch_env, tail_env, tonal_env, main_noise_env, amp_env;
  float tonal_vol, noise_vol, crack_vol;   // Oscillator.volume (only refreshed while params move, snare.rs:1052-1055)
  float psm;                               // pitch_start_multiplier
  Chamb filt;
  PhaseMod pm;
  WShaper ws;
  float velocity;
  uint32_t active;
  double t;
};

// SnareDrum::with_config (snare.rs:769-809); cfg = 18 normalized values in S_* order, aux = filter_type
G_HD void snare_init(SnareState& s, const float* cfg, uint32_t filter_type, float sr) {
  for (int i = 0; i < 18; i++) { float c = clampf(cfg[i], 0.0f, 1.0f); s.cur[i] = s.tgt[i] = c; }
  s.cur[S_TUNING] = s.tgt[S_TUNING] = 0.5f;
  s.filter_type = filter_type > 3 ? 3 : filter_type;
  env_init(s.tonal_osc_env); env_init(s.noise_osc_env); env_init(s.crack_env); env_init(s.pitch_env);
  env_init(s.tail_env); env_init(s.tonal_env); env_init(s.main_noise_env); env_init(s.amp_env);
  s.tonal_vol = s.noise_vol = s.crack_vol = 1.0f;  // Oscillator::new volume
  s.psm = 1.0f + s.cur[S_PITCH_DROP] * 1.5f;
  chamb_init(s.filt, sr, denorm(s.cur[S_FILTER_CUTOFF], 100.0f, 10000.0f), denorm(s.cur[S_FILTER_RES], 0.5f, 10.0f));
  s.pm.trig = 0.0; s.pm.active = 0;
  ws_init(s.ws, 1.0f, 1.0f);
  s.velocity = 0.5f; s.active = 0; s.t = 0.0;
}

// SnareDrum::trigger_with_velocity (snare.rs:873-1027)
G_HD void snare_trigger(SnareState& s, float velocity) {
  double time = s.t;
  s.velocity = clampf(velocity, 0.0f, 1.0f);
  s.active = 1;
  float vel = s.velocity, vel2 = vel * vel;
  float decay_scale = 1.0f - (0.45f * vel2);
  float pitch_decay_scale = 1.0f - (0.5f * vel2);
  float base_decay = denorm(s.cur[S_DECAY], 0.05f, 3.5f);
  float tonal_decay = denorm(s.cur[S_TONAL_DECAY], 0.0f, 3.5f);
  float tonal_decay_curve = denorm(s.cur[S_TONAL_DECAY_CURVE], 0.1f, 10.0f);
  float noise_decay = denorm(s.cur[S_NOISE_DECAY], 0.0f, 3.5f);
  float noise_tail_decay = denorm(s.cur[S_NOISE_TAIL_DECAY], 0.0f, 3.5f);
  float amp_decay = denorm(s.cur[S_AMP_DECAY], 0.0f, 4.0f);
  float amp_decay_curve = denorm(s.cur[S_AMP_DECAY_CURVE], 0.1f, 10.0f);
  float scaled_decay = base_decay * decay_scale;
  s.psm = 1.0f + s.cur[S_PITCH_DROP] * 1.5f;
  float pdt = fminf(scaled_decay * 0.3f * pitch_decay_scale, scaled_decay * 0.25f);
  env_config(s.pitch_env, 0.001f, pdt, 0.0f, pdt * 0.1f);
  s.tonal_vol = clampf(s.cur[S_TONAL], 0.0f, 1.0f);
  env_config(s.tonal_osc_env, 0.001f, 0.001f, 1.0f, scaled_decay * 0.4f);
  s.noise_vol = clampf(s.cur[S_NOISE] * 0.8f, 0.0f, 1.0f);
  env_config(s.noise_osc_env, 0.001f, 0.001f, 1.0f, scaled_decay * 0.3f);
  float cvs = 0.7f + 0.3f * vel;
  s.crack_vol = clampf(s.cur[S_BRIGHTNESS] * 0.4f * cvs, 0.0f, 1.0f);
  env_config(s.crack_env, 0.001f, scaled_decay * 0.2f, 0.0f, scaled_decay * 0.1f);
  float std_ = tonal_decay * decay_scale;
  env_config(s.tonal_env, 0.001f, std_, 0.0f, std_ * 0.2f, CURVE_LINEAR, tonal_decay_curve);
  float snd = noise_decay * decay_scale;
  env_config(s.main_noise_env, 0.001f, snd, 0.0f, snd * 0.2f);
  float stl = noise_tail_decay * decay_scale;
  env_config(s.tail_env, 0.001f, stl, 0.0f, stl * 0.3f);
  float sad = amp_decay * decay_scale;
  env_config(s.amp_env, 0.001f, sad, 0.0f, sad * 0.2f, CURVE_LINEAR, amp_decay_curve);
  env_trigger(s.tonal_osc_env, time); env_trigger(s.noise_osc_env, time); env_trigger(s.crack_env, time);
  env_trigger(s.pitch_env, time); env_trigger(s.tonal_env, time); env_trigger(s.main_noise_env, time);
  env_trigger(s.tail_env, time); env_trigger(s.amp_env, time);
  if (s.cur[S_PHASE_MOD] > 0.001f) { s.pm.trig = time; s.pm.active = 1; }
  s.filt.low = s.filt.band = 0.0f;
}

G_HD void snare_event(SnareState& s, const VoiceEvent& e) {
  switch (e.kind) {
    case EV_TRIGGER: snare_trigger(s, e.value); break;
    case EV_SET_TARGET: if (e.param < S_NP) { float c = clampf(e.value, 0.0f, 1.0f); if (fabsf(s.tgt[e.param] - c) > 1e-8f) s.tgt[e.param] = c; } break;
    case EV_SNAP: for (int i = 0; i < S_NP; i++) s.cur[i] = s.tgt[i]; break;
    case EV_SET_AUX:
      if (e.param == AUX_SNARE_FILTER_TYPE) { uint32_t t = (uint32_t)e.value; s.filter_type = t > 3 ? 3 : t; }
      else if (e.param == AUX_OVERSAMPLING) { uint32_t m = (uint32_t)e.value; if (s.ws.os.mode != m) { s.ws.os.mode = m; os_reset(s.ws.os); } }
      else if (e.param == AUX_SNARE_PITCH_START) s.psm = e.value;  // SnareDrum::set_config (:820)
      break;
    default: break;
  }
}

// SnareDrum::tick (snare.rs:1044-1198)
G_D float snare_tick(SnareState& s, const RateCtx& rc) {
  const double now = s.t;
  s.t = now + rc.dt;
  bool changing = false;
#pragma unroll
  for (int i = 0; i < S_NP; i++) { smooth_tick(s.cur[i], s.tgt[i], rc.smooth15); changing |= (s.cur[i] != s.tgt[i]); }
  if (!s.active) return 0.0f;
  const float sr = rc.sr;
  if (changing) {  // apply_params (:1206-1220)
    float cvs = 0.7f + 0.3f * s.velocity;
    s.tonal_vol = clampf(s.cur[S_TONAL], 0.0f, 1.0f);
    s.noise_vol = clampf(s.cur[S_NOISE] * 0.8f, 0.0f, 1.0f);
    s.crack_vol = clampf(s.cur[S_BRIGHTNESS] * 0.4f * cvs, 0.0f, 1.0f);
  }
  float vel2 = s.velocity * s.velocity;
  float decay_scale = 1.0f - (0.45f * vel2);
  float pitch_decay_scale = 1.0f - (0.5f * vel2);
  float scaled_decay = denorm(s.cur[S_DECAY], 0.05f, 3.5f) * decay_scale;
  float pdt = fminf(scaled_decay * 0.3f * pitch_decay_scale, scaled_decay * 0.25f);
  s.pitch_env.decay = pdt; s.pitch_env.release = pdt * 0.1f;
  s.tonal_osc_env.release = scaled_decay * 0.4f;
  s.noise_osc_env.release = scaled_decay * 0.3f;
  s.crack_env.decay = scaled_decay * 0.2f; s.crack_env.release = scaled_decay * 0.1f;
  float std_ = denorm(s.cur[S_TONAL_DECAY], 0.0f, 3.5f) * decay_scale;
  s.tonal_env.decay = std_; s.tonal_env.release = std_ * 0.2f;
  float snd = denorm(s.cur[S_NOISE_DECAY], 0.0f, 3.5f) * decay_scale;
  s.main_noise_env.decay = snd; s.main_noise_env.release = snd * 0.2f;
  float stl = denorm(s.cur[S_NOISE_TAIL_DECAY], 0.0f, 3.5f) * decay_scale;
  s.tail_env.decay = stl; s.tail_env.release = stl * 0.3f;
  float sad = denorm(s.cur[S_AMP_DECAY], 0.0f, 4.0f) * decay_scale;
  s.amp_env.decay = sad; s.amp_env.release = sad * 0.2f;
  float base_frequency = denorm(s.cur[S_FREQ], 100.0f, 600.0f) * tuning_to_multiplier(s.cur[S_TUNING]);
  float pev = env_amp(s.pitch_env, now);
  float fm = 1.0f + (s.psm - 1.0f) * pev;
  float pma = s.cur[S_PHASE_MOD];
  if (pma > 0.001f) {
    float pm = phasemod_tick(s.pm, now);
    fm *= 1.0f + (pm * pma * 1.0f);
  }
  float tonal_f = base_frequency * fm;
  chamb_set(s.filt, sr, denorm(s.cur[S_FILTER_CUTOFF], 100.0f, 10000.0f), denorm(s.cur[S_FILTER_RES], 0.5f, 10.0f));
  float xfade = s.cur[S_XFADE];
  float tonal_mix = 1.0f - xfade, noise_mix = xfade;
  // tonal: osc (sustain-1 envelope) * tonal_env * mix.  The heavy additive sum is skipped whenever any
  // factor of the product is exactly 0 (result would be +-0).
  float osc_amp, tonal_env;
  float idx_t = env_active(s.tonal_osc_env) ? (float)(now - s.tonal_osc_env.trig) * sr : 0.0f;
  osc_amp = env_amp(s.tonal_osc_env, now);
  tonal_env = env_amp(s.tonal_env, now);
  float tonal_out = 0.0f;
  if (osc_amp != 0.0f && s.tonal_vol != 0.0f && tonal_env != 0.0f && tonal_mix != 0.0f)
    tonal_out = osc_triangle(idx_t, tonal_f, sr) * osc_amp * s.tonal_vol * tonal_env * tonal_mix;
  // noise
  float idx_n = env_active(s.noise_osc_env) ? (float)(now - s.noise_osc_env.trig) * sr : 0.0f;
  float namp = env_amp(s.noise_osc_env, now);
  float raw_noise = 0.0f;
  if (namp != 0.0f && s.noise_vol != 0.0f) raw_noise = hash_noise(f32_to_u64_sat(idx_n)) * namp * s.noise_vol;
  float filtered = chamb_process(s.filt, raw_noise, s.filter_type);
  float ne = env_amp(s.main_noise_env, now);
  float te = env_amp(s.tail_env, now);
  float cne = (ne * 0.7f) + (te * 0.3f);
  float noise_out = filtered * cne * noise_mix;
  // crack
  float idx_c = env_active(s.crack_env) ? (float)(now - s.crack_env.trig) * sr : 0.0f;
  float camp = env_amp(s.crack_env, now);
  float crack_out = 0.0f;
  if (camp != 0.0f && s.crack_vol != 0.0f) crack_out = hash_noise(f32_to_u64_sat(idx_c)) * camp * s.crack_vol;
  float total = tonal_out + noise_out + crack_out;
  s.ws.drive = clampf(1.0f + (s.cur[S_OVERDRIVE] * 9.0f), 1.0f, 10.0f);
  float od = ws_process(s.ws, total);
  float amp_env = env_amp(s.amp_env, now);
  float va = sqrtf(s.velocity);
  float out = od * amp_env * va * s.cur[S_VOLUME];
  bool classic = env_active(s.tonal_osc_env) || env_active(s.noise_osc_env) || env_active(s.crack_env);
  bool ds = env_active(s.tonal_env) || env_active(s.main_noise_env) || env_active(s.tail_env) || env_active(s.amp_env) || s.pm.active;
  if (!classic && !ds) s.active = 0;
  return out;
}

// =========================================== HiHat2 ===========================================
enum { H_PITCH, H_DECAY, H_ATTACK, H_TONE, H_VOLUME, H_TUNING, H_NP };
struct HatState {
  float cur[H_NP], tgt[H_NP];
  uint32_t pink_on, db24;
  float mod_phase, main_phase;
  MaxEnv2 env;
  float env_smooth;
  Biquad hp1, hp2;
  Tpt svf;
  uint64_t white;
  Pink pink;
  float velocity;
  uint32_t active;
  double t;
};
G_HD float hat_pitch_hz(float p) { return denorm(p * p, 3500.0f, 10000.0f); }

// HiHat2::with_config (hihat2.rs:354-376); cfg = pitch, decay, attack, tone, volume
G_HD void hat_init(HatState& s, const float* cfg, uint32_t pink_on, uint32_t db24, float sr) {
  for (int i = 0; i < 5; i++) { float c = clampf(cfg[i], 0.0f, 1.0f); s.cur[i] = s.tgt[i] = c; }
  s.cur[H_TUNING] = s.tgt[H_TUNING] = 0.5f;
  s.pink_on = pink_on; s.db24 = db24;
  s.mod_phase = s.main_phase = 0.0f;
  s.env.target[0] = s.env.target[1] = 0; s.env.dur[0] = s.env.dur[1] = 0; s.env.curve[0] = s.env.curve[1] = 0;
  s.env.seg_start = 0.0; s.env.seg_start_val = 0; s.env.cur_val = 0; s.env.seg = 2; s.env.active = 0;  // MaxCurveEnvelope::new(vec![])
  s.env_smooth = 0.0f;
  hp_init(s.hp1, sr); hp_init(s.hp2, sr);
  tpt_init(s.svf, sr, denorm(s.cur[H_TONE], 500.0f, 10000.0f), 0.5f);
  s.white = XS_SEED;
  pink_reset(s.pink);
  s.velocity = 1.0f; s.active = 0; s.t = 0.0;
}
G_HD void hat_trigger(HatState& s, float velocity) {  // :434-451
  s.active = 1;
  s.velocity = clampf(velocity, 0.0f, 1.0f);
  float attack_ms = denorm(s.cur[H_ATTACK], 0.5f, 200.0f);
  float decay_ms = denorm(s.cur[H_DECAY], 0.5f, 4000.0f);
  maxenv_init(s.env, 1.0f, attack_ms, -0.3f, 0.0f, decay_ms, -0.8f);
  maxenv_trigger(s.env, s.t);
  s.env_smooth = 0.0f;
  s.mod_phase = s.main_phase = 0.0f;
  biquad_reset(s.hp1); biquad_reset(s.hp2);
  s.svf.ic1 = s.svf.ic2 = 0.0f;
}
G_HD void hat_event(HatState& s, const VoiceEvent& e) {
  switch (e.kind) {
    case EV_TRIGGER: hat_trigger(s, e.value); break;
    case EV_SET_TARGET: if (e.param < H_NP) { float c = clampf(e.value, 0.0f, 1.0f); if (fabsf(s.tgt[e.param] - c) > 1e-8f) s.tgt[e.param] = c; } break;
    case EV_SNAP: for (int i = 0; i < H_NP; i++) s.cur[i] = s.tgt[i]; break;
    case EV_SET_AUX:
      if (e.param == AUX_HAT_PINK) s.pink_on = e.value != 0.0f;
      else if (e.param == AUX_HAT_DB24) s.db24 = e.value != 0.0f;
      break;
    default: break;
  }
}
G_D float hat_osc(float& phase_cycle, float freq, float sr, float pm) {  // PhaseModOsc::tick (:277-286); set_frequency floors at 0
  float inc = fmaxf(freq, 0.0f) / sr;
  phase_cycle = fmodf(phase_cycle + inc, 1.0f);
  float ph = phase_cycle + pm;
  ph -= floorf(ph);
  return gm::g_sinf(2.0f * PI_F * ph);
}
// HiHat2::tick (hihat2.rs:453-508)
G_D float hat_tick(HatState& s, const RateCtx& rc) {
  const double now = s.t;
  s.t = now + rc.dt;
#pragma unroll
  for (int i = 0; i < H_NP; i++) smooth_tick(s.cur[i], s.tgt[i], rc.smooth15);
  if (!s.active) return 0.0f;
  const float sr = rc.sr;
  maxenv_set_dur_ms(s.env, 0, denorm(s.cur[H_ATTACK], 0.5f, 200.0f));
  maxenv_set_dur_ms(s.env, 1, denorm(s.cur[H_DECAY], 0.5f, 4000.0f));
  float pitch_hz = hat_pitch_hz(s.cur[H_PITCH]) * tuning_to_multiplier(s.cur[H_TUNING]);
  float noise;
  if (s.pink_on) noise = pink_tick(s.pink, rc.pink);
  else { float n = u64_to_f32(xorshift64s_next(s.white)) / 18446744073709551616.0f; noise = (n * 2.0f) - 1.0f; }
  float mod_out = hat_osc(s.mod_phase, pitch_hz * 0.1f, sr, noise * 0.25f);
  float main_out = hat_osc(s.main_phase, pitch_hz, sr, mod_out * 0.75f);
  hp_set(s.hp1, sr, pitch_hz, 1.0f);
  float filtered = biquad_process(s.hp1, main_out);
  if (s.db24) { hp_set(s.hp2, sr, pitch_hz, 1.0f); filtered = biquad_process(s.hp2, filtered) * 0.8f; }
  float env = maxenv_value(s.env, now);
  if (env >= s.env_smooth) s.env_smooth = env; else s.env_smooth += rc.asym_down * (env - s.env_smooth);
  env = s.env_smooth;
  float out = filtered * env * s.velocity * 0.35f;
  tpt_set(s.svf, sr, denorm(s.cur[H_TONE], 500.0f, 10000.0f), 0.5f);
  float lo, bd, hi;
  tpt_process(s.svf, out, lo, bd, hi);
  float o = hi * s.cur[H_VOLUME];
  if (maxenv_complete(s.env) && s.env_smooth < 1e-4f) s.active = 0;
  return o;
}

// =========================================== Tom2 ===========================================
enum { T_TUNE, T_BEND, T_TONE, T_COLOR, T_DECAY, T_MEMBRANE, T_MEMBRANE_Q, T_VOLUME, T_TUNING, T_NP };
struct TomState {
  float p[T_NP];      // plain f32 parameters, 0-100 (tuning 0-1), applied instantly (tom2.rs:66-79)
  // MorphOsc (gen/morph_osc.rs)
  float main_sine_phase, mtri_phase, fixed_sine_phase, rand_phase, rand_current, rand_target, gated_sine_phase;
  uint64_t noise_counter;
  uint32_t click_pos, click_playing;
  Biquad bp;
  MaxEnv2 env;
  float tri_phase;
  uint32_t past_attack, main_done, active, tri_enabled;
  Biquad mem[5];
  float mem_q_scale, mem_gain_scale, ring_level;
  float saved_freq; uint32_t has_saved;
  double t;
};
#ifdef __CUDACC__
__constant__ float c_tom_impulse[64] = {
#else
static const float c_tom_impulse[64] = {
#endif
    0.884058f, 0.942029f, 0.913043f, 0.869565f, 0.833333f, 0.797101f, 0.772947f, 0.748792f, 0.724638f,
    0.695652f, 0.666667f, 0.637681f, 0.619565f, 0.601449f, 0.583333f, 0.565217f, 0.536232f, 0.507246f,
    0.478261f, 0.449275f, 0.42029f,  0.391304f, 0.371981f, 0.352657f, 0.333333f, 0.304348f, 0.275362f,
    0.23913f,  0.202899f, 0.181159f, 0.15942f,  0.137681f, 0.115942f, 0.101449f, 0.086957f, 0.072464f,
    0.057971f, 0.043478f, 0.028986f, 0.014493f, 0.009662f, 0.004831f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f,
    0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.014493f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};

// MembraneResonator::update_filters (membrane_resonator.rs:87-93)
G_HD void tom_membrane_update(TomState& s, float sr) {
  const float P[5][3] = {{275.0f, 165.0f, 376.0f}, {220.0f, 228.0f, 205.0f}, {79.0f, 294.0f, 143.0f}, {65.0f, 320.0f, 129.0f}, {57.0f, 326.0f, 141.0f}};
  for (int i = 0; i < 5; i++) {
    float sq = clampf(P[i][2] * s.mem_q_scale, 0.1f, 100.0f);
    float sg = P[i][0] * s.mem_gain_scale;
    bp_set(s.mem[i], sr, P[i][1], sq, sg);
  }
}
G_HD void tom_update_membrane_params(TomState& s, float sr) {  // tom2.rs:405-411
  float qs = 0.005f + (s.p[T_MEMBRANE_Q] / 100.0f) * 0.015f;
  s.mem_q_scale = clampf(qs, 0.001f, 1.0f);
  tom_membrane_update(s, sr);
  s.mem_gain_scale = clampf(0.003f, 0.0001f, 0.1f);
  tom_membrane_update(s, sr);
}
// Tom2::new (tom2.rs:199-235), then optional set_config (:413-423) with cfg (0-100 units) if cfg != nullptr
G_HD void tom_init(TomState& s, const float* cfg, float sr) {
  s.p[T_TUNE] = 50.0f; s.p[T_BEND] = 30.0f; s.p[T_TONE] = 50.0f; s.p[T_COLOR] = 50.0f; s.p[T_DECAY] = 50.0f;
  s.p[T_MEMBRANE] = 0.0f; s.p[T_MEMBRANE_Q] = 50.0f; s.p[T_VOLUME] = 100.0f; s.p[T_TUNING] = 0.5f;
  s.main_sine_phase = s.mtri_phase = s.fixed_sine_phase = s.rand_phase = s.rand_current = s.rand_target = s.gated_sine_phase = 0.0f;
  s.noise_counter = 0; s.click_pos = 0; s.click_playing = 0;
  bp_init(s.bp, sr);
  maxenv_init(s.env, 1.0f, 1.0f, 0.8f, 0.0f, 2000.0f, -0.83f);
  s.tri_phase = 0.0f; s.past_attack = 0; s.main_done = 0; s.active = 0; s.tri_enabled = 1;
  for (int i = 0; i < 5; i++) bp_init(s.mem[i], sr);
  s.mem_q_scale = 0.01f; s.mem_gain_scale = 0.0031f; s.ring_level = 0.0f;
  tom_membrane_update(s, sr);           // MembraneResonator::with_params
  tom_update_membrane_params(s, sr);    // Tom2::new tail
  s.saved_freq = 0.0f; s.has_saved = 0; s.t = 0.0;
  if (cfg) {
    for (int i = 0; i < 8; i++) s.p[i] = cfg[i];
    tom_update_membrane_params(s, sr);
  }
}
G_HD void tom_trigger(TomState& s, float sr) {  // :428-448 (velocity ignored)
  s.active = 1; s.past_attack = 0;
  s.main_sine_phase = s.mtri_phase = s.fixed_sine_phase = s.rand_phase = s.rand_current = s.rand_target = s.gated_sine_phase = 0.0f;
  s.noise_counter = 0;
  s.click_pos = 0; s.click_playing = 1;
  s.tri_phase = 0.0f;
  biquad_reset(s.bp);
  for (int i = 0; i < 5; i++) biquad_reset(s.mem[i]);
  s.ring_level = 0.0f;
  s.main_done = 0;
  float decay_ms = 0.5f + (s.p[T_DECAY] / 100.0f) * (4000.0f - 0.5f);
  maxenv_init(s.env, 1.0f, 1.0f, 0.8f, 0.0f, decay_ms, -0.83f);
  maxenv_trigger(s.env, s.t);
}
G_HD void tom_event(TomState& s, const VoiceEvent& e, float sr) {
  switch (e.kind) {
    case EV_TRIGGER: tom_trigger(s, sr); break;
    case EV_SET_TARGET:  // value already in the voice's internal units (0-100; tuning 0-1)
      if (e.param < T_NP) {
        s.p[e.param] = e.param == T_TUNING ? clampf(e.value, 0.0f, 1.0f) : clampf(e.value, 0.0f, 100.0f);
        if (e.param == T_MEMBRANE_Q) tom_update_membrane_params(s, sr);
      }
      break;
    case EV_NOTE_FREQ:   // get_freq_param() = tune (0-100); set_param(0, v) = clamp01(v)*100 (ffi.rs:131-137, 213-217)
      if (!s.has_saved) { s.saved_freq = s.p[T_TUNE]; s.has_saved = 1; }
      s.p[T_TUNE] = clampf(clampf(e.value, 0.0f, 1.0f) * 100.0f, 0.0f, 100.0f);
      break;
    case EV_RESTORE_FREQ:
      if (s.has_saved) { s.has_saved = 0; s.p[T_TUNE] = clampf(clampf(s.saved_freq, 0.0f, 1.0f) * 100.0f, 0.0f, 100.0f); }
      break;
    case EV_SET_AUX:
      if (e.param >= AUX_TOM_RAW_PARAM0 && e.param < AUX_TOM_RAW_PARAM0 + 8) s.p[e.param - AUX_TOM_RAW_PARAM0] = e.value;  // set_config: unclamped
      else if (e.param == AUX_TOM_CONFIG_DONE) tom_update_membrane_params(s, sr);
      break;
    default: break;
  }
}
G_D void phase_advance(float& ph, float f, float sr) { ph += f / sr; if (ph >= 1.0f) ph -= 1.0f; }
G_D float tri_wave(float ph) { float t = fract(ph); return t < 0.5f ? 4.0f * t - 1.0f : 3.0f - 4.0f * t; }
G_D float unit_sine(float ph) { return gm::g_sinf(ph * 2.0f * PI_F); }

// Tom2::tick (tom2.rs:450-585) with MorphOsc::tick (morph_osc.rs:137-202) inlined
G_D float tom_tick(TomState& s, const RateCtx& rc) {
  const double now = s.t;
  s.t = now + rc.dt;
  if (!s.active) return 0.0f;
  const float sr = rc.sr;
  float env = maxenv_value(s.env, now);
  if (env > 0.9f) s.past_attack = 1;
  float tn = s.p[T_TUNE] / 100.0f;
  float base_frequency = (40.0f + (tn * tn) * (600.0f - 40.0f)) * tuning_to_multiplier(s.p[T_TUNING]);
  float bend_scaled = (s.p[T_BEND] / 100.0f) * 2.0f;
  float eb = env * bend_scaled;
  float raw_freq = base_frequency * (1.0f + eb * eb);
  if (maxenv_complete(s.env) || (s.past_attack && raw_freq < 20.0f)) s.main_done = 1;
  if (s.main_done && !(s.ring_level > 0.0001f)) { s.active = 0; return 0.0f; }
  float fade = (s.past_attack && raw_freq < 40.0f) ? (raw_freq - 20.0f) / (40.0f - 20.0f) : 1.0f;
  float mf = fmaxf(raw_freq, 40.0f);
  // click
  float click = 0.0f;
  if (s.click_playing) {
    if (s.click_pos >= 64) s.click_playing = 0;
    else { click = c_tom_impulse[s.click_pos]; s.click_pos += 1; if (s.click_pos >= 64) s.click_playing = 0; }
  }
  float click_out = click * 1.1f;
  float tri_out = s.tri_enabled ? tri_wave(s.tri_phase) * 0.5f : 0.0f;
  phase_advance(s.tri_phase, mf, sr);
  float tone = s.p[T_TONE], color = s.p[T_COLOR];
  float mix_control = (tone / 100.0f) * 2.0f - 1.0f;
  float color_midi = 30.0f + (color / 100.0f) * 20.0f;
  float cf1 = 440.0f * gm::g_powf(2.0f, (color_midi - 69.0f) / 12.0f);
  // MorphOsc::tick
  float main_sine = unit_sine(s.main_sine_phase) * 0.5f;
  phase_advance(s.main_sine_phase, mf, sr);
  float mtri = tri_wave(s.mtri_phase) * 0.5f;
  phase_advance(s.mtri_phase, mf, sr);
  float fixed_sine = unit_sine(s.fixed_sine_phase) * 0.5f;
  phase_advance(s.fixed_sine_phase, 190.0f, sr);
  s.noise_counter += 1;
  float noise = hash_noise(s.noise_counter) * 0.2f;
  float rand_freq = 440.0f * gm::g_powf(2.0f, (cf1 - 69.0f) / 12.0f);
  float prev = s.rand_phase;
  phase_advance(s.rand_phase, rand_freq, sr);
  if (s.rand_phase < prev) { s.rand_current = s.rand_target; s.rand_target = hash_noise(s.noise_counter + 0x12345678ull); }
  float rand_value = s.rand_current + (s.rand_target - s.rand_current) * s.rand_phase;
  float noise_combined = (noise + rand_value) * 0.4f;
  float gated = tone < 99.0f ? unit_sine(s.gated_sine_phase) * 0.2f : 0.0f;
  phase_advance(s.gated_sine_phase, mf, sr);
  float ch1 = main_sine * fixed_sine, ch2 = mtri + noise_combined, ch3 = noise_combined + gated;
  float w1 = clampf(-mix_control, 0.0f, 1.0f), w2 = clampf(1.0f - fabsf(mix_control), 0.0f, 1.0f), w3 = clampf(mix_control, 0.0f, 1.0f);
  float morph_out = ch1 * w1 + ch2 * w2 + ch3 * w3;
  float mixed = click_out + tri_out + morph_out;
  float ff = fmaxf(mf, 20.0f);
  float cn = color / 100.0f;
  float fq = 1.0f + cn * cn;
  bp_set(s.bp, sr, ff, fq, 1.1f);
  float filtered = biquad_process(s.bp, mixed);
  float membrane = s.p[T_MEMBRANE];
  float mem_out = 0.0f;
  if (membrane > 0.0f) {  // MembraneResonator::process (membrane_resonator.rs:189-200)
    float mi = s.main_done ? 0.0f : filtered * env;
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 5; i++) acc += biquad_process(s.mem[i], mi);
    float clipped = gm::g_tanhf(acc);
    s.ring_level = s.ring_level * 0.999f + fabsf(clipped) * 0.001f;
    mem_out = clipped;
  }
  float vol = s.p[T_VOLUME] / 100.0f;
  float mm = membrane / 100.0f;
  if (s.main_done) {
    float fd = s.ring_level >= 0.005f ? 1.0f : (s.ring_level <= 0.0001f ? 0.0f : (s.ring_level - 0.0001f) / (0.005f - 0.0001f));
    return mem_out * mm * fd * 0.7f * vol;
  }
  float dry_gain = 1.0f - mm;
  float dry = filtered * env;
  float fs = dry * dry_gain + mem_out * mm;
  return fs * fade * 0.7f * vol;
}

}  // namespace gd
