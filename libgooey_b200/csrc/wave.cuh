// wave.cuh — kernel W: the warp-per-voice back-end.  One warp owns one voice and walks its timeline in blocks of
// (up to) 32 frames, lane = frame.  Within a block
//   * every linear time-invariant recurrence (one-poles, all-pass sections of the half-band oversampler, TPT / Chamberlin
//     state-variable filters, RBJ biquads, DC blocker) is evaluated as a Kogge-Stone prefix scan over the lanes:
//     y[n] = a y[n-1] + b[n] for scalars, s[n] = M s[n-1] + f[n] over 2x2 state matrices for second-order sections;
//   * every memoryless stage between them (tanh waveshaping at 4x rate, gain compensation, sine lookups of phase-
//     modulated oscillators, mixes) is plain SIMT work, one frame per lane;
//   * the few genuinely nonlinear / integer recurrences (xorshift RNGs, f32 phase accumulators, the envelope follower's
//     data-dependent coefficient, the hi-hat's asymmetric smoother) are replayed in the reference's exact operation order
//     over the block (a handful of instructions per frame; every lane runs the same loop and keeps its own frame).
// This turns the serial tail that bounds kernel C (1024 voices = 32 warps on 148 SMs, IPC 0.19) into 1024 warps of mostly
// data-parallel work.  Scans re-associate the f32 sums, so W is not bit-identical to the per-sample order (C / S are);
// the difference is the rounding noise of the recurrence itself (measured <= 3e-6 through the 4x oversampler at drive 40),
// inside the 1e-5 parity bar.  State is the same Aud struct the serial paths use, so calls can alternate between them.
#pragma once
#include "kernels.cuh"

namespace gd {

constexpr unsigned FULLMASK = 0xffffffffu;

// ---- first-order scans -----------------------------------------------------------------------------------
// Per-lane constants of the recurrence y[n] = a y[n-1] + b[n] over 32 lanes.
struct Geo1 { float m[5]; float ap; };            // m[i] = lane >= 2^i ? a^(2^i) : 0 ;  ap = a^(lane+1)
// Two frames per lane (64-long sequence, lane holds elements 2l and 2l+1): lane ratio a^2.
struct Geo2 { float a; float m[5]; float ap; };   // m[i] = lane >= 2^i ? (a^2)^(2^i) : 0 ;  ap = (a^2)^lane

__device__ __forceinline__ Geo1 make_geo1(float a, int lane) {
  Geo1 g; float p = a, ap = 1.0f; const int e = lane + 1;
#pragma unroll
  for (int i = 0; i < 5; i++) { g.m[i] = lane >= (1 << i) ? p : 0.0f; if (e & (1 << i)) ap *= p; p *= p; }
  if (e & 32) ap *= p;
  g.ap = ap;
  return g;
}
__device__ __forceinline__ Geo2 make_geo2(float a, int lane) {
  Geo2 g; g.a = a; float p = a * a, ap = 1.0f;
#pragma unroll
  for (int i = 0; i < 5; i++) { g.m[i] = lane >= (1 << i) ? p : 0.0f; if (lane & (1 << i)) ap *= p; p *= p; }
  g.ap = ap;
  return g;
}
__device__ __forceinline__ float scan1(float b, const Geo1& g, float y_prev) {
#pragma unroll
  for (int i = 0; i < 5; i++) { float t = __shfl_up_sync(FULLMASK, b, 1 << i); b = __fmaf_rn(g.m[i], t, b); }
  return __fmaf_rn(g.ap, y_prev, b);
}
__device__ __forceinline__ void scan2(float& b0, float& b1, const Geo2& g, float y_prev, int lane) {
  float B = __fmaf_rn(g.a, b0, b1);
#pragma unroll
  for (int i = 0; i < 5; i++) { float t = __shfl_up_sync(FULLMASK, B, 1 << i); B = __fmaf_rn(g.m[i], t, B); }
  float yp = __shfl_up_sync(FULLMASK, B, 1);
  if (lane == 0) yp = 0.0f;
  yp = __fmaf_rn(g.ap, y_prev, yp);
  b0 = __fmaf_rn(g.a, yp, b0);
  b1 = __fmaf_rn(g.a, b0, b1);
}
__device__ __forceinline__ float prev_lane(float v, float carry, int lane) { float t = __shfl_up_sync(FULLMASK, v, 1); return lane == 0 ? carry : t; }

// Shared-memory table of the per-lane constants that do not depend on the voice.
enum { GT_HB = 0, GT_CLICK = 8, GT_DC = 9, GT_PINK = 10, GT_RING = 13, GT_N1 = 14 };
struct GeoTables {
  float g1[GT_N1][6][32];   // half-band sections at the section's own rate (ratio -c_i), then the fixed one-poles
  float g2[8][7][32];       // the half-band sections over two frames per lane
};
__device__ __forceinline__ void geo_tables_init(GeoTables& T, const RateCtx& rc, int tid, int nthreads) {
  for (int idx = tid; idx < GT_N1 * 32; idx += nthreads) {
    const int i = idx >> 5, lane = idx & 31;
    float ratio;
    if (i < 8) ratio = -c_hb[i];
    else if (i == GT_CLICK) ratio = 1.0f - rc.click_alpha;
    else if (i == GT_DC) ratio = 0.995f;
    else if (i == GT_PINK) ratio = rc.pink.p0;
    else if (i == GT_PINK + 1) ratio = rc.pink.p1;
    else if (i == GT_PINK + 2) ratio = rc.pink.p2;
    else ratio = 0.999f;
    const Geo1 a = make_geo1(ratio, lane);
#pragma unroll
    for (int k = 0; k < 5; k++) T.g1[i][k][lane] = a.m[k];
    T.g1[i][5][lane] = a.ap;
    if (i < 8) {
      const Geo2 b = make_geo2(ratio, lane);
#pragma unroll
      for (int k = 0; k < 5; k++) T.g2[i][1 + k][lane] = b.m[k];
      T.g2[i][0][lane] = b.a; T.g2[i][6][lane] = b.ap;
    }
  }
}
__device__ __forceinline__ Geo1 load_geo1(const GeoTables& T, int i, int lane) {
  Geo1 g;
#pragma unroll
  for (int k = 0; k < 5; k++) g.m[k] = T.g1[i][k][lane];
  g.ap = T.g1[i][5][lane];
  return g;
}
__device__ __forceinline__ Geo2 load_geo2(const GeoTables& T, int i, int lane) {
  Geo2 g; g.a = T.g2[i][0][lane];
#pragma unroll
  for (int k = 0; k < 5; k++) g.m[k] = T.g2[i][1 + k][lane];
  g.ap = T.g2[i][6][lane];
  return g;
}

// One all-pass section y[n] = c (x[n] - y[n-1]) + x[n-1] (dsp.cuh hb_run) over a block; (xs, ys) = section state.
__device__ __forceinline__ float ap_sec1(float x, float c, const Geo1& g, float& xs, float& ys, int lane, int last) {
  const float xp = prev_lane(x, xs, lane);
  const float y = scan1(__fmaf_rn(c, x, xp), g, ys);
  xs = __shfl_sync(FULLMASK, x, last); ys = __shfl_sync(FULLMASK, y, last);
  return y;
}
// two frames per lane; `last` = last valid lane, both of its frames valid
__device__ __forceinline__ void ap_sec2(float& x0, float& x1, float c, const Geo2& g, float& xs, float& ys, int lane, int last) {
  const float xp = prev_lane(x1, xs, lane);
  float b0 = __fmaf_rn(c, x0, xp), b1 = __fmaf_rn(c, x1, x0);
  const float nx = __shfl_sync(FULLMASK, x1, last);
  scan2(b0, b1, g, ys, lane);
  xs = nx; ys = __shfl_sync(FULLMASK, b1, last);
  x0 = b0; x1 = b1;
}
// hb_run over a block at the stage's own rate: path 0 through coefficients 0,2,4,6, path 1 through 1,3,5,7
__device__ __forceinline__ void hb_run1(Hb8& s, float& p0, float& p1, const GeoTables& T, int lane, int last) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    p0 = ap_sec1(p0, c_hb[i], load_geo1(T, i, lane), s.x[i], s.y[i], lane, last);
    p1 = ap_sec1(p1, c_hb[i + 1], load_geo1(T, i + 1, lane), s.x[i + 1], s.y[i + 1], lane, last);
  }
}
__device__ __forceinline__ void hb_run2(Hb8& s, float& a0, float& a1, float& b0, float& b1, const GeoTables& T, int lane, int last) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    ap_sec2(a0, a1, c_hb[i], load_geo2(T, i, lane), s.x[i], s.y[i], lane, last);
    ap_sec2(b0, b1, c_hb[i + 1], load_geo2(T, i + 1, lane), s.x[i + 1], s.y[i + 1], lane, last);
  }
}
// Oversampler::process over a block (utils/oversampler.rs:38-113; dsp.cuh os_process); f is the memoryless nonlinearity
template <class F> __device__ __forceinline__ float os_scan(Oversamp& o, float in, F f, const GeoTables& T, int lane, int last) {
  if (o.mode == 0) return f(in);
  float s0 = in, s1 = in;
  hb_run1(o.outer_up, s0, s1, T, lane, last);          // 2x stream: s0, s1 per frame
  if (o.mode == 2) {
    float a = f(s1), b = f(s0);
    hb_run1(o.outer_down, a, b, T, lane, last);
    return 0.5f * (a + b);
  }
  float i00 = s0, i01 = s1, i10 = s0, i11 = s1;        // inner_up path 0 / path 1 over the 2x stream
  hb_run2(o.inner_up, i00, i01, i10, i11, T, lane, last);
  // 4x stream per frame: i00, i10, i01, i11.  inner_down takes (second, first) of each pair
  float a0 = f(i10), a1 = f(i11), b0 = f(i00), b1 = f(i01);
  hb_run2(o.inner_down, a0, a1, b0, b1, T, lane, last);
  const float d0 = 0.5f * (a0 + b0), d1 = 0.5f * (a1 + b1);
  float a = d1, b = d0;
  hb_run1(o.outer_down, a, b, T, lane, last);
  return 0.5f * (a + b);
}

// ---- second-order LTI scans: s[n] = M s[n-1] + f[n],  M = [m0 m1; m2 m3] ------------------------------------
struct Lin2 {            // per-warp, in shared memory
  float P[5][4];         // M^(2^i)
  float L[4][32];        // M^(lane+1)
};
__device__ __forceinline__ void mat_mul(const float* a, const float* b, float* c) {
  c[0] = a[0] * b[0] + a[1] * b[2]; c[1] = a[0] * b[1] + a[1] * b[3];
  c[2] = a[2] * b[0] + a[3] * b[2]; c[3] = a[2] * b[1] + a[3] * b[3];
}
__device__ __forceinline__ void lin2_init(Lin2& S, float m0, float m1, float m2, float m3, int lane) {
  float p[4] = {m0, m1, m2, m3}, acc[4] = {1.0f, 0.0f, 0.0f, 1.0f}, t[4];
  const int e = lane + 1;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    if (lane == 0) { S.P[i][0] = p[0]; S.P[i][1] = p[1]; S.P[i][2] = p[2]; S.P[i][3] = p[3]; }
    if (e & (1 << i)) { mat_mul(p, acc, t); acc[0] = t[0]; acc[1] = t[1]; acc[2] = t[2]; acc[3] = t[3]; }
    mat_mul(p, p, t); p[0] = t[0]; p[1] = t[1]; p[2] = t[2]; p[3] = t[3];
  }
  if (e & 32) { mat_mul(p, acc, t); acc[0] = t[0]; acc[1] = t[1]; acc[2] = t[2]; acc[3] = t[3]; }
  S.L[0][lane] = acc[0]; S.L[1][lane] = acc[1]; S.L[2][lane] = acc[2]; S.L[3][lane] = acc[3];
  __syncwarp();
}
// in: forcing (f0, f1) per lane; out: state after each frame.  (s0p, s1p) = state before the block.
__device__ __forceinline__ void lin2_scan(const Lin2& S, float& f0, float& f1, float s0p, float s1p, int lane) {
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const float t0 = __shfl_up_sync(FULLMASK, f0, 1 << i), t1 = __shfl_up_sync(FULLMASK, f1, 1 << i);
    if (lane >= (1 << i)) {
      f0 = __fmaf_rn(S.P[i][0], t0, __fmaf_rn(S.P[i][1], t1, f0));
      f1 = __fmaf_rn(S.P[i][2], t0, __fmaf_rn(S.P[i][3], t1, f1));
    }
  }
  f0 = __fmaf_rn(S.L[0][lane], s0p, __fmaf_rn(S.L[1][lane], s1p, f0));
  f1 = __fmaf_rn(S.L[2][lane], s0p, __fmaf_rn(S.L[3][lane], s1p, f1));
}

// Per-warp read-only tables in shared memory (written by span_setup, then only read).
struct WaveScratch { Lin2 lin[3]; float stash[6][32]; };

// The Aud state is held in REGISTERS, replicated in every lane (all lanes compute the same state updates from
// shuffled values), so there are no shared-memory hazards; the exact-order serial sections below are executed by all
// lanes redundantly, which costs the same issue slots as running them on one lane.

__device__ __forceinline__ float wrap01(float x) { return x < 1.0f ? x : (x < 2.0f ? x - 1.0f : fmodf(x, 1.0f)); }   // fmodf(x, 1) for x >= 0, exact

// TPT SVF as a linear map of (ic1, ic2) (state_variable_tpt.rs:56-69, resonant_lowpass.rs:49-103):
//   v1 = (g (in - ic2) + ic1) h ; v2 = ic2 + g v1 ; ic1' = 2 v1 - ic1 ; ic2' = 2 v2 - ic2
struct TptLin { float in0, in1; };
__device__ __forceinline__ TptLin tpt_lin_init(Lin2& S, const Tpt& f, int lane) {
  const float g = f.g, h = f.h, gh = g * h, g2h = g * gh;
  lin2_init(S, 2.0f * h - 1.0f, -2.0f * gh, 2.0f * gh, 1.0f - 2.0f * g2h, lane);
  TptLin t; t.in0 = 2.0f * gh; t.in1 = 2.0f * g2h;
  return t;
}
// runs the filter over the block; returns this lane's PRE-state so the caller can form the outputs exactly as the
// serial code does from it
__device__ __forceinline__ void tpt_scan(const Lin2& S, const TptLin& t, Tpt& f, float in, float& ic1_pre, float& ic2_pre, int lane, int last) {
  float s0 = t.in0 * in, s1 = t.in1 * in;
  lin2_scan(S, s0, s1, f.ic1, f.ic2, lane);
  ic1_pre = prev_lane(s0, f.ic1, lane); ic2_pre = prev_lane(s1, f.ic2, lane);
  f.ic1 = __shfl_sync(FULLMASK, s0, last); f.ic2 = __shfl_sync(FULLMASK, s1, last);
}
// RBJ biquad, direct form I (dsp.cuh biquad_process): y[n] = b0 x[n] + b1 x[n-1] + b2 x[n-2] - a1 y[n-1] - a2 y[n-2]
__device__ __forceinline__ void biquad_lin_init(Lin2& S, const Biquad& b, int lane) { lin2_init(S, -b.a1, -b.a2, 1.0f, 0.0f, lane); }
__device__ __forceinline__ float biquad_scan(const Lin2& S, Biquad& b, float x, int lane, int last) {
  const float xm1 = prev_lane(x, b.x1, lane);
  float xm2 = __shfl_up_sync(FULLMASK, x, 2);
  xm2 = lane == 0 ? b.x2 : (lane == 1 ? b.x1 : xm2);
  float f0 = b.b0 * x + b.b1 * xm1 + b.b2 * xm2, f1 = 0.0f;
  lin2_scan(S, f0, f1, b.y1, b.y2, lane);
  const float nx1 = __shfl_sync(FULLMASK, x, last), nx2 = __shfl_sync(FULLMASK, xm1, last);
  const float ny1 = __shfl_sync(FULLMASK, f0, last), ny2 = __shfl_sync(FULLMASK, f1, last);
  b.x1 = nx1; b.x2 = nx2; b.y1 = ny1; b.y2 = ny2;
  return fabsf(f0) < 1e-15f ? 0.0f : f0;
}

// Replay helper: value of `v` held by lane n, for a loop index n that is uniform across the warp.
#define LANE_VAL(v, n) __shfl_sync(FULLMASK, (v), (n))

// ---- kick ------------------------------------------------------------------------------------------------
struct KickW {
  using V = KickV;
  struct Span2 { bool noise, shaper, serial; };
  static __device__ __forceinline__ int active_end(const KickV::Run& r) { return r.j_act; }
  static __device__ __forceinline__ void span_setup(KickAud& a, const KickV::Run& r, Span2& w, WaveScratch&, const RateCtx& rc, int) {
    w.noise = r.d.noise_amount > 0.001f;
    a.ws.drive = r.d.drive; a.ws.feedback = r.d.feedback;
    fbws_set_cutoff(a.ws, rc.sr, r.d.fb_cutoff);
    w.shaper = !(a.ws.mix <= 0.0001f || a.ws.drive <= 1.0f);
    w.serial = w.shaper && a.ws.feedback != 0.0f;     // tanh inside the feedback loop: no scan (feedback_waveshaper.rs:122-159)
    if (w.noise) rlp_set(a.noise_lp, rc.sr, r.d.noise_cut, r.d.noise_res);
  }
  // one block: lanes [0, nl) hold frames j .. j+nl-1.  p[] = this lane's front planes.  Returns the output sample.
  static __device__ __forceinline__ float block(KickAud& a, const KickV::Run& r, const Span2& w, WaveScratch&, const GeoTables& T,
                                                const float* p, int, int& nl, const RateCtx& rc, int lane) {
    const int last = nl - 1;
    const float p1 = p[0], raw_click = p[1], ne = p[2], amp = p[3];
    // click high-pass and the pink-noise layer: cheap recurrences, replayed in the reference's order (kick.rs:1171-1193)
    float hp_l = 0.0f, fn_l = 0.0f;
#pragma unroll 4
    for (int n = 0; n < nl; n++) {
      const float raw = LANE_VAL(raw_click, n);
      const float hp = raw - a.click_hp;
      a.click_hp += rc.click_alpha * hp;
      float fn = 0.0f;
      if (w.noise) fn = rlp_process(a.noise_lp, pink_tick(a.pink, rc.pink));
      if (n == lane) { hp_l = hp; fn_l = fn; }
    }
    const float filt_click = hp_l * (1.0f + 4.0f * 0.1f);
    const float noise_out = w.noise ? fn_l * ne * r.d.noise_amount * 0.5f : 0.0f;
    const float total = p1 + filt_click + noise_out;
    float od = total;
    const bool finite = __all_sync(FULLMASK, isfinite(total));
    if (w.serial || (w.shaper && !finite)) {       // feedback > 0 (or a non-finite input): the reference's per-sample order
#pragma unroll 4
      for (int n = 0; n < nl; n++) {
        const float y = fbws_process(a.ws, LANE_VAL(total, n));
        if (n == lane) od = y;
      }
    } else if (w.shaper) {
      FbShaper& ws = a.ws;
      const float fb_in = ws.drive * total + ws.feedback * ws.last_out;   // feedback == 0
      const float shaped = os_scan(ws.os, fb_in, [](float x) { return gm::g_tanhf(x); }, T, lane, last);
      // envelope follower: the coefficient depends on a comparison with the running state -> replay
      const float rect_l = fabsf(total);
      float env_l = 0.0f;
#pragma unroll 4
      for (int n = 0; n < nl; n++) {
        const float rect = LANE_VAL(rect_l, n);
        const float coeff = rect > ws.env ? ws.env_att : ws.env_rel;
        ws.env += (1.0f - coeff) * (rect - ws.env);
        if (fabsf(ws.env) < 1e-15f) ws.env = 0.0f;
        if (n == lane) env_l = ws.env;
      }
      const float comp = fbws_gain_comp(ws, env_l);
      const float compensated = shaped * comp;
      // DC blocker and the one-pole that feeds last_out (feedback_waveshaper.rs:262-271, 151-159): replay
      float out_l = 0.0f;
#pragma unroll 4
      for (int n = 0; n < nl; n++) {
        const float c = LANE_VAL(compensated, n);
        const float out = c - ws.dc_x1 + 0.995f * ws.dc_y1;
        ws.dc_x1 = c;
        ws.dc_y1 = fabsf(out) < 1e-15f ? 0.0f : out;
        ws.filter_state += ws.filter_coeff * (out - ws.filter_state);
        if (fabsf(ws.filter_state) < 1e-15f) ws.filter_state = 0.0f;
        if (n == lane) out_l = out;
      }
      ws.last_out = ws.filter_state;
      od = total * (1.0f - ws.mix) + out_l * ws.mix;
    }
    return od * amp * r.d.va * r.d.volume;
  }
};

// ---- snare -----------------------------------------------------------------------------------------------
struct SnareW {
  using V = SnareV;
  struct Span2 { float comp; bool shaper; };
  static __device__ __forceinline__ int active_end(const SnareV::Run& r) { return r.j_act; }
  static __device__ __forceinline__ void span_setup(SnareAud& a, const SnareV::Run& r, Span2& w, WaveScratch&, const RateCtx& rc, int) {
    chamb_set(a.filt, rc.sr, r.d.cutoff, r.d.res);
    a.ws.drive = r.d.drive;
    w.shaper = !(a.ws.mix <= 0.0001f || a.ws.drive <= 1.0f);
    w.comp = w.shaper ? gm::g_tanhf(0.5f) / gm::g_tanhf(0.5f * a.ws.drive) : 1.0f;
  }
  static __device__ __forceinline__ float block(SnareAud& a, const SnareV::Run& r, const Span2& w, WaveScratch&, const GeoTables& T,
                                                const float* p, int, int& nl, const RateCtx&, int lane) {
    const int last = nl - 1;
    const float tonal_out = p[0], raw_noise = p[1], cne = p[2], crack_out = p[3], amp = p[4];
    // Chamberlin SVF: replayed exactly (it is unstable for high cutoff x low resonance, and the reference's
    // blow-up -> NaN -> Waveshaper guard -> 0 sequence has to be reproduced sample for sample)
    float low_l = 0.0f, band_l = 0.0f, high_l = 0.0f;
    {
      const float f = a.filt.f, q = a.filt.q;
      float low = a.filt.low, band = a.filt.band;
#pragma unroll 4
      for (int n = 0; n < nl; n++) {
        const float in = LANE_VAL(raw_noise, n);
        float high = 0.0f;
#pragma unroll
        for (int i = 0; i < 2; i++) { low = low + f * band; high = in - low - q * band; band = f * high + band; }
        if (n == lane) { low_l = low; band_l = band; high_l = high; }
      }
      a.filt.low = low; a.filt.band = band;
    }
    const float filtered = svf_pick(low_l, band_l, high_l, r.d.filter_type);
    const float noise_out = filtered * cne * r.d.noise_mix;
    const float total = tonal_out + noise_out + crack_out;
    float od = total;
    const bool finite = __all_sync(FULLMASK, isfinite(total));
    if (!finite) {          // waveshaper.rs:49-53 resets its oversampler on a non-finite input: per-sample order
#pragma unroll 4
      for (int n = 0; n < nl; n++) {
        const float y = ws_process(a.ws, LANE_VAL(total, n));
        if (n == lane) od = y;
      }
    } else if (w.shaper) {
      const float d = a.ws.drive, comp = w.comp;
      const float sat = os_scan(a.ws.os, total, [d, comp](float x) { return gm::g_tanhf(x * d) * comp; }, T, lane, last);
      od = total * (1.0f - a.ws.mix) + sat * a.ws.mix;
    }
    return od * amp * r.d.va * r.d.volume;
  }
};

// ---- hi-hat ----------------------------------------------------------------------------------------------
struct HatW {
  using V = HatV;
  struct Span2 { TptLin svf; };
  static __device__ __forceinline__ int active_end(const HatV::Run&) { return 0x7fffffff; }
  static __device__ __forceinline__ void span_setup(HatAud& a, const HatV::Run& r, Span2& w, WaveScratch& sc, const RateCtx& rc, int lane) {
    hp_set(a.hp1, rc.sr, r.d.pitch_hz, 1.0f);
    biquad_lin_init(sc.lin[0], a.hp1, lane);
    if (r.d.db24) { hp_set(a.hp2, rc.sr, r.d.pitch_hz, 1.0f); biquad_lin_init(sc.lin[1], a.hp2, lane); }
    tpt_set(a.svf, rc.sr, r.d.tone_hz, 0.5f);
    w.svf = tpt_lin_init(sc.lin[2], a.svf, lane);
  }
  // j = frame of lane 0.  May shorten nl when the voice deactivates inside the block (frames after it are silent).
  static __device__ __forceinline__ float block(HatAud& a, const HatV::Run& r, const Span2& w, WaveScratch& sc, const GeoTables&,
                                                const float* p, int j, int& nl, const RateCtx& rc, int lane) {
    const float sr = rc.sr;
    // envelope through the asymmetric smoother and the deactivation test (hihat2.rs:489-506): replay
    const float env_l = (j + lane) < r.j_env ? p[0] : r.env_final;
    float es_l = 0.0f;
    int nv = nl;
#pragma unroll 4
    for (int n = 0; n < nl; n++) {
      const float e = LANE_VAL(env_l, n);
      if (e >= a.env_smooth) a.env_smooth = e; else a.env_smooth += rc.asym_down * (e - a.env_smooth);
      if (n == lane) es_l = a.env_smooth;
      if ((j + n) >= r.j_env && a.env_smooth < 1e-4f) { nv = n + 1; a.active = 0; break; }
    }
    nl = nv;
    // noise and the two phase accumulators: replay
    float noise = 0.0f, mph = 0.0f, nph = 0.0f;
    const float inc_mod = fmaxf(r.d.pitch_hz * 0.1f, 0.0f) / sr, inc_main = fmaxf(r.d.pitch_hz, 0.0f) / sr;
#pragma unroll 4
    for (int n = 0; n < nv; n++) {
      float x;
      if (r.d.pink_on) x = pink_tick(a.pink, rc.pink);
      else { float u = u64_to_f32(xorshift64s_next(a.white)) / 18446744073709551616.0f; x = (u * 2.0f) - 1.0f; }
      a.mod_phase = wrap01(a.mod_phase + inc_mod);
      a.main_phase = wrap01(a.main_phase + inc_main);
      if (n == lane) { noise = x; mph = a.mod_phase; nph = a.main_phase; }
    }
    // PhaseModOsc outputs are memoryless given the phases (hihat2.rs:277-286): one frame per lane
    float ph = mph + noise * 0.25f; ph -= floorf(ph);
    const float mod_out = gm::g_sinf(2.0f * PI_F * ph);
    ph = nph + mod_out * 0.75f; ph -= floorf(ph);
    const float main_out = gm::g_sinf(2.0f * PI_F * ph);
    // two RBJ high-passes (Q = 1), envelope, TPT high-pass (Q = 0.5): well damped LTI sections -> 2x2 scans
    const int last = nv - 1;
    float filtered = biquad_scan(sc.lin[0], a.hp1, main_out, lane, last);
    if (r.d.db24) filtered = biquad_scan(sc.lin[1], a.hp2, filtered, lane, last) * 0.8f;
    const float out = filtered * es_l * r.d.vel035 * 0.35f;
    float ic1p, ic2p;
    tpt_scan(sc.lin[2], w.svf, a.svf, out, ic1p, ic2p, lane, last);
    const float v1 = (a.svf.g * (out - ic2p) + ic1p) * a.svf.h;
    const float v2 = ic2p + a.svf.g * v1;
    const float hi = out - (a.svf.r * v1 + v2);
    return hi * r.d.volume;
  }
};

// ---- tom -------------------------------------------------------------------------------------------------
struct TomW {
  using V = TomV;
  struct Span2 { int dummy; };
  static __device__ __forceinline__ int active_end(const TomV::Run&) { return 0x7fffffff; }
  static __device__ __forceinline__ void span_setup(TomAud&, const TomV::Run&, Span2&, WaveScratch&, const RateCtx&, int) {}
  static __device__ __forceinline__ float block(TomAud& a, const TomV::Run& r, const Span2&, WaveScratch& sc, const GeoTables&,
                                                const float* p, int j, int& nl, const RateCtx& rc, int lane) {
    const TomDer& d = r.d;
    const float sr = rc.sr;
    const float env_l = (j + lane) < r.j_env ? p[0] : r.env_final;
    const float noise_l = p[1], rnd_l = p[2];
    const bool complete_l = (j + lane) >= r.j_env;
    const unsigned valid = nl >= 32 ? FULLMASK : ((1u << nl) - 1u);
    const unsigned le = (2u << lane) - 1u;     // lanes 0..lane
    // pure per-frame quantities (tom2.rs:455-489)
    const float eb = env_l * d.bend_scaled;
    const float raw_freq = d.base_frequency * (1.0f + eb * eb);
    const unsigned m_att = __ballot_sync(FULLMASK, env_l > 0.9f) & valid;
    const bool past_l = a.past_attack || (m_att & le) != 0u;
    const unsigned m_done = __ballot_sync(FULLMASK, complete_l || (past_l && raw_freq < 20.0f)) & valid;
    const int first_done = a.main_done ? 0 : (m_done ? __ffs(m_done) - 1 : 32);
    if (a.main_done || (first_done < nl && (d.membrane > 0.0f || a.ring_level > 0.0001f))) {
      // the ringing tail (membrane) is gated by its own level: the reference's per-sample order, replayed
      float y_l = 0.0f;
      int nv = nl;
#pragma unroll 4
      for (int n = 0; n < nl; n++) {
        TomFront f; f.env = LANE_VAL(env_l, n); f.noise = LANE_VAL(noise_l, n); f.rnd = LANE_VAL(rnd_l, n);
        const float y = tom_back(a, d, f, (j + n) >= r.j_env, rc);
        if (n == lane) y_l = y;
        if (!a.active) { nv = n + 1; break; }
      }
      nl = nv;
      return y_l;
    }
    // no main_done inside [0, nv): the membrane (if any) is driven, the voice cannot deactivate before frame nv
    const int nv = first_done < nl ? first_done : nl;
    if (first_done < nl) { a.main_done = 1; a.active = 0; a.past_attack = (m_att & ((2u << first_done) - 1u)) != 0u || a.past_attack; }
    else a.past_attack = a.past_attack || m_att != 0u;
    nl = nv;
    if (nv == 0) return 0.0f;
    const float fade = (past_l && raw_freq < 40.0f) ? (raw_freq - 20.0f) / (40.0f - 20.0f) : 1.0f;
    const float mf_l = fmaxf(raw_freq, 40.0f);
    float click = 0.0f;
    if (a.click_playing) {
      const uint32_t idx = a.click_pos + (uint32_t)lane;
      if (idx < 64u) click = c_tom_impulse[idx];
      a.click_pos = min(64u, a.click_pos + (uint32_t)nv);
      if (a.click_pos >= 64u) a.click_playing = 0;
    }
    // phase accumulators and the rand~ sample-and-hold: replay (morph_osc.rs:137-202).  The per-frame increments f / sr
    // are formed once per lane (same f32 division as phase_advance) instead of six times per replayed frame.
    float tp = 0.0f, msp = 0.0f, mtp = 0.0f, fsp = 0.0f, gsp = 0.0f, rp = 0.0f, rcur = 0.0f, rtgt = 0.0f;
    const float inc_l = mf_l / sr, inc_fixed = 190.0f / sr, inc_rand = d.rand_freq / sr;
#pragma unroll 4
    for (int n = 0; n < nv; n++) {
      const float inc = LANE_VAL(inc_l, n), rnd = LANE_VAL(rnd_l, n);
      if (n == lane) { tp = a.tri_phase; msp = a.main_sine_phase; mtp = a.mtri_phase; fsp = a.fixed_sine_phase; gsp = a.gated_sine_phase; }
      a.tri_phase += inc; if (a.tri_phase >= 1.0f) a.tri_phase -= 1.0f;
      a.main_sine_phase += inc; if (a.main_sine_phase >= 1.0f) a.main_sine_phase -= 1.0f;
      a.mtri_phase += inc; if (a.mtri_phase >= 1.0f) a.mtri_phase -= 1.0f;
      a.fixed_sine_phase += inc_fixed; if (a.fixed_sine_phase >= 1.0f) a.fixed_sine_phase -= 1.0f;
      const float prev = a.rand_phase;
      a.rand_phase += inc_rand; if (a.rand_phase >= 1.0f) a.rand_phase -= 1.0f;
      if (a.rand_phase < prev) { a.rand_current = a.rand_target; a.rand_target = rnd; }
      a.gated_sine_phase += inc; if (a.gated_sine_phase >= 1.0f) a.gated_sine_phase -= 1.0f;
      if (n == lane) { rp = a.rand_phase; rcur = a.rand_current; rtgt = a.rand_target; }
    }
    const float click_out = click * 1.1f;
    const float tri_out = a.tri_enabled ? tri_wave(tp) * 0.5f : 0.0f;
    const float main_sine = unit_sine(msp) * 0.5f;
    const float mtri = tri_wave(mtp) * 0.5f;
    const float fixed_sine = unit_sine(fsp) * 0.5f;
    const float noise = noise_l * 0.2f;
    const float rand_value = rcur + (rtgt - rcur) * rp;
    const float noise_combined = (noise + rand_value) * 0.4f;
    const float gated = d.tone < 99.0f ? unit_sine(gsp) * 0.2f : 0.0f;
    const float ch1 = main_sine * fixed_sine, ch2 = mtri + noise_combined, ch3 = noise_combined + gated;
    const float morph_out = ch1 * d.w1 + ch2 * d.w2 + ch3 * d.w3;
    const float mixed = click_out + tri_out + morph_out;
    const float ff_l = fmaxf(mf_l, 20.0f);
    // pitch-tracking band-pass: every lane prepares the coefficients its frame would get; the replay applies the
    // reference's change thresholds (biquad_bandpass.rs:73-87) and picks them up when an update is due
    Biquad spec;
    bp_compute(spec, sr, ff_l, d.fq, 1.1f);
    float filtered = 0.0f;
#pragma unroll 4
    for (int n = 0; n < nv; n++) {
      const float ffn = LANE_VAL(ff_l, n);
      if (!bp_unchanged(a.bp, ffn, d.fq, 1.1f)) {
        a.bp.last_freq = ffn; a.bp.last_q = d.fq; a.bp.last_gain = 1.1f;
        a.bp.b0 = LANE_VAL(spec.b0, n); a.bp.b1 = LANE_VAL(spec.b1, n); a.bp.b2 = LANE_VAL(spec.b2, n);
        a.bp.a1 = LANE_VAL(spec.a1, n); a.bp.a2 = LANE_VAL(spec.a2, n);
      }
      const float y = biquad_process(a.bp, LANE_VAL(mixed, n));
      if (n == lane) filtered = y;
    }
    float mem_out = 0.0f;
    if (d.membrane > 0.0f) {  // MembraneResonator::process (membrane_resonator.rs:189-200); main_done is false here
      // Five fixed band-passes on the same input.  At 165-326 Hz a direct-form-I section amplifies its own f32
      // rounding noise ~40x, so any re-association (scan) drifts 1e-4 from the reference: they are replayed in the
      // reference's operation order — but the five independent filters run on five LANES at once (lane i owns
      // filter i), and frame n then sums the five outputs in the reference's order.
      const float mi_l = filtered * env_l;
      Biquad mine = a.mem[4];
      if (lane == 0) mine = a.mem[0]; else if (lane == 1) mine = a.mem[1]; else if (lane == 2) mine = a.mem[2]; else if (lane == 3) mine = a.mem[3];
#pragma unroll 4
      for (int n = 0; n < nv; n++) {
        const float y = biquad_process(mine, LANE_VAL(mi_l, n));
        if (lane < 5) sc.stash[lane][n] = y;
      }
      __syncwarp();
      float acc_l = 0.0f;
#pragma unroll
      for (int i = 0; i < 5; i++) acc_l += sc.stash[i][lane];
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 5; i++) {
        a.mem[i].x1 = LANE_VAL(mine.x1, i); a.mem[i].x2 = LANE_VAL(mine.x2, i);
        a.mem[i].y1 = LANE_VAL(mine.y1, i); a.mem[i].y2 = LANE_VAL(mine.y2, i);
      }
      const float clipped = gm::g_tanhf(acc_l);
      const float ac = fabsf(clipped);
#pragma unroll 4
      for (int n = 0; n < nv; n++) a.ring_level = a.ring_level * 0.999f + LANE_VAL(ac, n) * 0.001f;
      mem_out = clipped;
    }
    const float dry_gain = 1.0f - d.mm;
    const float dry = filtered * env_l;
    const float fs = dry * dry_gain + mem_out * d.mm;
    return fs * fade * 0.7f * d.vol;
  }
};

// ---- the kernel ----------------------------------------------------------------------------------------------
// One warp per voice; WARPS voices per CTA.
#ifndef GOOEY_WAVE_MIN_CTAS
#define GOOEY_WAVE_MIN_CTAS 1
#endif
#ifndef GOOEY_WAVE_WARPS
#define GOOEY_WAVE_WARPS 1
#endif
template <class W, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, GOOEY_WAVE_MIN_CTAS) wave_kernel(const VoiceLaunch L) {
  using V = typename W::V; using Span = typename V::Span; using Aud = typename V::Aud;
  constexpr int WC = sizeof(typename V::Ctl) / 4;
  __shared__ GeoTables T;
  __shared__ WaveScratch scr[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  geo_tables_init(T, L.rc, threadIdx.x, WARPS * 32);
  __syncthreads();
  const int v = blockIdx.x * WARPS + warp;
  if (v >= L.n || L.mode[v] != 0) return;
  const int sv = L.slots ? (int)L.slots[v] : v;
  WaveScratch& sc = scr[warp];
  Aud a;
  load_words(a, L.state, sv, L.n_pad, WC);
  const Span* spans = reinterpret_cast<const Span*>(L.spans) + L.span_off[v];
  const uint32_t ns = L.n_spans[v];
  uint32_t cur = L.span_cursor[v];
  typename V::Run run;
  typename W::Span2 w2;
  bool ready = false;
  if (cur != 0xffffffffu) V::span_resume(spans[cur], run);
  int next_j0 = cur + 1u < ns ? spans[cur + 1u].j0 : 0x7fffffff;
  const long long row = L.rows ? (long long)L.rows[v] : (long long)(L.row0 + v);
  float* out = L.out + row * L.stride;
  const float* planes = L.planes + (size_t)v * L.pitch;
  const int c1 = L.chunk0 + min(L.chunk_frames, L.frames - L.chunk0);
  int j = L.chunk0;
  while (j < c1) {
    while (j == next_j0) {
      cur += 1u;
      V::span_begin(a, spans[cur], run, L.rc.sr);
      ready = false;
      next_j0 = cur + 1u < ns ? spans[cur + 1u].j0 : 0x7fffffff;
    }
    const int span_end = min(next_j0, c1);
    const int act_end = min(span_end, W::active_end(run));
    if (cur == 0xffffffffu || !V::is_active(a, run, j) || j >= act_end) {      // silent until the next span: zero-fill
      for (int k = j + lane; k < span_end; k += 32) out[k] = 0.0f;
      j = span_end;
      continue;
    }
    if (!ready) { __syncwarp(); W::span_setup(a, run, w2, sc, L.rc, lane); ready = true; }
    int nl = min(32, act_end - j);
    const int nl0 = nl;
    float p[V::NPL];
    const int jj = j + min(lane, nl - 1);       // lanes past the block re-read its last frame: finite, never stored
#pragma unroll
    for (int q = 0; q < V::NPL; q++) p[q] = planes[(size_t)q * L.plane_stride + (jj - L.chunk0)];
    const float y = W::block(a, run, w2, sc, T, p, j, nl, L.rc, lane);
    if (lane < nl0) out[j + lane] = lane < nl ? y : 0.0f;
    j += nl0;
  }
  if (lane == 0) {
    store_words(a, L.state, sv, L.n_pad, WC);
    L.span_cursor[v] = cur;
  }
}

}  // namespace gd
