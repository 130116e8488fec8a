// bass_wave.cuh — the bass voice, one warp per voice, lane = frame (reference: src/instruments/bass.rs:793-877).
//
// A bass tick is three f64 phase accumulators -> f64 sine + two polyBLEP saw/square pairs -> optional oversampled
// waveshaper -> TPT low-pass whose cutoff follows the filter envelope (coefficients recomputed, through a change
// threshold, while the envelope moves) -> amp envelope.  Per 32-frame block the warp
//   * replays the phase accumulators and the state-variable filter in the reference's order (cheap recurrences),
//   * evaluates everything else one frame per lane: sines, polyBLEPs, both envelopes (from lane-local copies whose
//     latches are advanced analytically, dsp.cuh env_advance), tan() of the filter coefficients, the waveshaper through
//     the half-band scans of wave.cuh.
// Blocks are cut at event frames and at the frame where the amp envelope ends.  Whenever a parameter is still gliding
// (the smoothers then change the tick's inputs every sample) or the voice is idle, lane 0 simply runs the per-sample
// path (voices2.cuh bass_tick) for the block, so the two paths share one state and alternate freely.
#pragma once
#include "wave.cuh"

namespace gd {

constexpr int BASS_WORDS = sizeof(BassState) / 4;

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) bass_wave_kernel(const VoiceLaunch L) {
  __shared__ w32::GeoTables T;
  __shared__ BassState states[WARPS];
  __shared__ float stashes[WARPS][6][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  w32::geo_tables_init(T, L.rc, threadIdx.x, WARPS * 32);
  __syncthreads();
  const int v = blockIdx.x * WARPS + warp;
  if (v >= L.n) return;
  const int sv = L.slots ? (int)L.slots[v] : v;
  BassState& s = states[warp];
  float (*st)[32] = stashes[warp];
  {
    uint32_t* w = reinterpret_cast<uint32_t*>(&s);
    for (int i = lane; i < BASS_WORDS; i += 32) w[i] = L.state[(size_t)i * L.n_pad + sv];
  }
  __syncwarp();
  uint32_t ev = L.ev_begin[v];
  const uint32_t ev_end = L.ev_begin[v + 1];
  const long long row = L.rows ? (long long)L.rows[v] : (long long)(L.row0 + v);
  float* out = L.out + row * L.stride;
  const RateCtx& rc = L.rc;
  const float sr = rc.sr;
  int j = 0;
  while (j < L.frames) {
    // events due at frame j, then the block runs up to the next event
    if (ev < ev_end && L.events[ev].frame <= (uint32_t)j) {
      if (lane == 0) while (ev < ev_end && L.events[ev].frame <= (uint32_t)j) { bass_event(s, L.events[ev], L.tt); ev++; }
      ev = __shfl_sync(0xffffffffu, ev, 0);
      __syncwarp();
    }
    int end = min(j + 32, L.frames);
    if (ev < ev_end) end = min(end, (int)L.events[ev].frame);
    int nl = end - j;
    bool ok = lane < B_NP ? s.cur[lane] == s.tgt[lane] : true;
    const bool settled = __all_sync(0xffffffffu, ok);
    bool parallel = settled && s.active != 0;
    int j_off = J_NONE;
    const uint32_t k0 = s.k;
    if (parallel) {                       // cut the block at the frame whose tick ends the amp envelope
      Env ae = s.amp_env;
      j_off = env_advance(ae, L.tt, k0, 0, nl);
      if (j_off != J_NONE) nl = j_off + 1;
    }
    const int last = nl - 1;
    float y = 0.0f;
    // ---- parallel block ----
    if (parallel) {
      const float freq = s.trig_freq * tuning_to_multiplier(s.cur[B_TUNING]);
      const float sub_level = s.cur[B_SUB], osc_level = s.cur[B_OSC], detune_level = s.cur[B_DETUNE_LEVEL];
      const float detune_cents = denorm(s.cur[B_DETUNE_AMT], 0.0f, 30.0f);
      const float osc_shape = s.cur[B_SHAPE];
      const float detune_ratio = gm::g_powf(2.0f, detune_cents / 1200.0f);
      const float detune_freq = freq * detune_ratio;
      const double dt = 1.0 / (double)sr;
      const double sub_inc = (double)freq * dt, osc_inc = (double)freq * dt, det_inc = (double)detune_freq * dt;
      double sp = s.sub_phase, op = s.osc_phase, dp = s.detune_phase, my_sp = 0.0, my_op = 0.0, my_dp = 0.0;
#pragma unroll 4
      for (int n = 0; n < nl; n++) {      // bass.rs:820-828: advance, then read
        sp += sub_inc; sp -= floor(sp);
        op += osc_inc; op -= floor(op);
        dp += det_inc; dp -= floor(dp);
        if (n == lane) { my_sp = sp; my_op = op; my_dp = dp; }
      }
      const float sub_out = (float)sin(my_sp * 6.283185307179586476925286766559);
      const float saw_m = polyblep_saw(my_op, osc_inc), sq_m = polyblep_square(my_op, osc_inc);
      const float osc_out = saw_m * (1.0f - osc_shape) + sq_m * osc_shape;
      const float saw_d = polyblep_saw(my_dp, det_inc), sq_d = polyblep_square(my_dp, det_inc);
      const float det_out = saw_d * (1.0f - osc_shape) + sq_d * osc_shape;
      const float mix = sub_out * sub_level + osc_out * osc_level + det_out * detune_level;
      const float od = s.cur[B_OVERDRIVE];
      const float drive = clampf(1.0f + od * 9.0f, 1.0f, 10.0f);
      const bool shaping = od > 0.001f && !(s.ws.mix <= 0.0001f || drive <= 1.0f);
      const bool finite = __all_sync(0xffffffffu, lane >= nl || isfinite(mix));
      if (od > 0.001f && !finite) parallel = false;     // Waveshaper resets on a non-finite input: per-sample order (nothing was written yet)
      else {
        float sat = mix;
        Oversamp os = s.ws.os;
        if (shaping) {
          const float comp = gm::g_tanhf(0.5f) / gm::g_tanhf(0.5f * drive), wmix = s.ws.mix;
          const float shaped = w32::os_scan(os, mix, [drive, comp](float x) { return w32::w_tanh(x * drive) * comp; }, T, lane, last);
          sat = mix * (1.0f - wmix) + shaped * wmix;
        }
        // envelopes: lane-local copies advanced through the frames before mine
        Env fe = s.flt_env, ae = s.amp_env;
        env_advance(fe, L.tt, k0, 0, lane);
        env_advance(ae, L.tt, k0, 0, lane);
        const double now = L.tt[k0 + (uint32_t)min(lane, last)];
        const float fenv = env_value(fe, now);
        const float amp_env = env_value(ae, now);
        const float base_cutoff = exp_denorm(s.cur[B_CUTOFF], 20.0f, 18000.0f);
        const float env_offset = (18000.0f - base_cutoff) * s.cur[B_FENV_AMT] * fenv;
        const float cutoff = clampf(base_cutoff + env_offset, 20.0f, 18000.0f);
        // filter coefficients this frame would get if the change threshold lets them through (state_variable_tpt.rs:83-92)
        Tpt spec;
        spec.cutoff = clampf(cutoff, 20.0f, sr * 0.45f);
        spec.res = fmaxf(denorm(s.cur[B_RES], 0.5f, 15.0f), 0.5f);
        spec.ic1 = spec.ic2 = 0.0f;
        tpt_update(spec, sr);
        __syncwarp();
        st[0][lane] = spec.cutoff; st[1][lane] = spec.g; st[2][lane] = spec.r; st[3][lane] = spec.h; st[4][lane] = sat;
        __syncwarp();
        Tpt f = s.filter;
        const float nr = spec.res;
        float lo_l = 0.0f;
#pragma unroll 4
        for (int n = 0; n < nl; n++) {
          const float nc = st[0][n];
          const bool upd = fabsf(nc - f.cutoff) > 0.001f || fabsf(nr - f.res) > 0.001f;
          f.cutoff = upd ? nc : f.cutoff; f.res = upd ? nr : f.res;
          f.g = upd ? st[1][n] : f.g; f.r = upd ? st[2][n] : f.r; f.h = upd ? st[3][n] : f.h;
          float lo, bd, hi;
          tpt_process(f, st[4][n], lo, bd, hi);
          if (n == lane) lo_l = lo;
        }
        y = lo_l * amp_env * sqrtf(s.velocity) * s.cur[B_VOLUME];
        __syncwarp();
        if (lane == 0) {                  // commit the block
          s.sub_phase = sp; s.osc_phase = op; s.detune_phase = dp;
          s.filter = f;
          s.ws.drive = drive;
          if (shaping) s.ws.os = os;
          Env fe2 = s.flt_env, ae2 = s.amp_env;
          env_advance(fe2, L.tt, k0, 0, nl);
          env_advance(ae2, L.tt, k0, 0, nl);
          s.flt_env = fe2; s.amp_env = ae2;
          if (!env_active(ae2)) s.active = 0;
          s.k = k0 + (uint32_t)nl;
        }
        __syncwarp();
      }
    }
    // ---- per-sample path for this block (gliding parameters, idle voice, non-finite waveshaper input) ----
    if (!parallel) {
      if (lane == 0) for (int n = 0; n < nl; n++) st[5][n] = bass_tick(s, L.tt, rc);
      __syncwarp();
      y = st[5][min(lane, last)];
      __syncwarp();
    }
    if (lane < nl) out[j + lane] = y;
    j += nl;
  }
  __syncwarp();
  {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&s);
    for (int i = lane; i < BASS_WORDS; i += 32) L.state[(size_t)i * L.n_pad + sv] = w[i];
  }
}

}  // namespace gd
