// loops.cuh — sample-playback sources of the engine path: the 4-channel stereo loop mixer
// (mixer/mod.rs:60-75, mixer/loop_channel.rs:181-208 and :262-300, mixer/stereo_buffer.rs:198-262) and the sampler racks
// (instruments/sampler.rs:64-225).  Both feed the MixerGraph as stereo sources (ffi.rs:1289-1308): source 4 = loop mixer,
// sources 5-8 = sampler racks.
//
// Shape of the work: every source is a gather from a host-supplied PCM buffer at a cursor that advances by a constant f64
// increment per output frame — `cursor += delta` accumulated frame by frame in the reference, with a modulo at the loop
// seam — so the cursor recurrence is replayed exactly as written (one thread per engine walks its own cursors; the adds
// are one f64 add per frame and channel) and the reads are 4-tap cubic (loops) or 2-tap linear (samplers) gathers from
// buffers that stay L2-resident (a 10 s stereo loop is 3.5 MB).  A warp's 32 engines stage 32 frames of their stereo rows
// in shared memory and store them transposed, so the planes the mixer reads are written with full 128-byte lines.
//
// The playback state (cursors, fader and gate smoothers, sampler voices) is host-authoritative between render calls: the
// FFI setters act on it directly, exactly like the reference's setters act on the LoopChannel / SamplerRack structs;
// engines_render uploads one descriptor per engine with a loaded source, the kernel advances it piece by piece, and the
// state is read back when the call ends.
//
// PitchMode::PreservePitch (mixer/wsola.rs) runs inside the same tick: every hop (20 ms) the channel's thread searches the best
// aligned grain start (<= ~130 candidates x hop taps of normalised cross-correlation), windows a 2-hop grain and overlap-adds it;
// the stretcher's buffers (9 hops of floats per channel) live in a device buffer owned by the channel.
//
// Not built (a request latches the engine's sticky error instead of rendering different audio): per-channel effect chains and
// the clip grid.
#pragma once
#include "dsp.cuh"

namespace gd {

constexpr int LOOP_CHANNELS = 4;                      // mixer/mod.rs:32
constexpr int SAMPLER_RACKS = 4;                      // ffi.rs:585
constexpr int SAMPLER_SLOTS = 16, SAMPLER_VOICES = 32;   // sampler.rs:14-15
constexpr int EXT_SOURCES = 1 + SAMPLER_RACKS;        // graph sources 4 .. 8 (graph.rs:35-42)
constexpr float LOOP_MAX_SPEED = 4.0f, LOOP_MAX_GAIN = 2.0f, LOOP_FADER_MS = 15.0f;   // loop_channel.rs:20-24

struct LSm { float c, t; };                           // SmoothedParam: current, target (settled <=> c == t)
G_HD void lsm_set(LSm& s, float v, float lo, float hi) { float c = clampf(v, lo, hi); if (fabsf(s.t - c) > 1e-8f) s.t = c; }   // smoother.rs:84-95
G_HD float lsm_tick(LSm& s, float coeff) { smooth_tick(s.c, s.t, coeff); return s.c; }

G_HD double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
G_HD double rem_euclid_d(double a, double b) { double r = fmod(a, b); return r < 0.0 ? r + fabs(b) : r; }   // f64::rem_euclid
G_HD long long rem_euclid_i(long long a, long long b) { long long r = a % b; return r < 0 ? r + b : r; }    // isize::rem_euclid (b > 0)

// utils/mod.rs:26-32
G_HD float cubic_interp(float p0, float p1, float p2, float p3, float t) {
  float a0 = -0.5f * p0 + 1.5f * p1 - 1.5f * p2 + 0.5f * p3;
  float a1 = p0 - 2.5f * p1 + 2.0f * p2 - 0.5f * p3;
  float a2 = -0.5f * p0 + 0.5f * p2;
  float a3 = p1;
  return ((a0 * t + a1) * t + a2) * t + a3;
}

// ---- loop_channel.rs LoopWindow (:59-116) and LoopChannel::window (:293-307) -------------------------------------------
struct LoopWindow { double lo, hi, span, len; bool wraps; };
G_HD LoopWindow loop_window(float loop_start, float loop_end, double len) {
  LoopWindow w;
  w.lo = clampd((double)loop_start * len, 0.0, len);
  w.hi = clampd((double)loop_end * len, 0.0, len);
  w.wraps = w.hi < w.lo;
  w.span = w.wraps ? len - w.lo + w.hi : w.hi - w.lo;
  w.len = len;
  return w;
}
G_HD double window_to_virtual(const LoopWindow& w, double p) { return rem_euclid_d(p - w.lo, w.len); }
G_HD double window_to_physical(const LoopWindow& w, double v) { return rem_euclid_d(w.lo + v, w.len); }
G_HD bool window_contains(const LoopWindow& w, double p) { return w.wraps ? (p >= w.lo || p < w.hi) : (p >= w.lo && p < w.hi); }
G_HD double window_fold(const LoopWindow& w, double p) {   // :99-115
  if (window_contains(w, p)) return p;
  if (w.wraps) return (p - w.hi) <= (w.lo - p) ? w.hi : w.lo;
  return clampd(p, w.lo, w.hi);
}

// One loop channel of one engine.  `left` / `right` are device planes of `len` frames (nullptr: nothing loaded).
struct LoopChan {
  const float* left; const float* right;
  double cursor;              // playback position in source frames
  double warp;                // warp_ratio() when the pitch mode is Resample, else 1.0 (:282-291, :233-237)
  uint32_t len; float buf_sr;
  float loop_start, loop_end, speed;
  uint32_t playing;
  LSm gain, active;           // user fader; mute / solo gate written by Mixer::tick (mod.rs:63-72)
  // PitchMode::PreservePitch (mixer/wsola.rs): the channel's WsolaStretcher.  st_buf holds its buffers as planes of `hop` floats:
  // out l, out r | grain l (2 hop), grain r (2 hop) | tail l, tail r | tail mono = 9 * hop floats; hann = the window table [2 * hop]
  float* st_buf; const float* hann;
  double st_cursor;           // analysis_cursor
  double warp_pp;             // warp_ratio() of this mode (tempo enters through the hop-to-hop jump only)
  uint32_t hop;               // hop_len = round(20 ms * engine rate)
  uint32_t preserve;          // the pitch mode is PreservePitch (the stretcher plays while speed >= 0, loop_channel.rs:184)
  uint32_t st_valid;          // a stretcher exists (0: built from the cursor at the next tick, :222-224)
  uint32_t st_have_prev, st_drain, pad;
  // a buffer queued to replace this one at the next grid boundary of the loop (queue_swap, loop_channel.rs:413-423): the take's
  // planes, length, rate and the warp ratios its tempo tag gives; `swaps` counts the swaps that landed during this call
  const float* pend_left; const float* pend_right;
  double pend_warp, pend_warp_pp;
  uint32_t pend_len; float pend_buf_sr;
  uint32_t pend_div, has_pending, swaps, pad2;
};
struct LoopMixer { LoopChan ch[LOOP_CHANNELS]; uint32_t row; uint32_t pad; };   // row = index of the stereo row pair this mixer writes

// stereo_buffer.rs:198-262
G_HD void loop_read_interpolated(const float* L, const float* R, uint32_t len, double position, float& ol, float& orr) {
  if (len == 1) { ol = L[0]; orr = R[0]; return; }
  const double last = (double)(len - 1);
  position = clampd(position, 0.0, last);
  const long long index = (long long)floor(position);
  const float frac = (float)(position - (double)index);
  const long long lastI = (long long)len - 1;
  const long long i0 = index - 1 < 0 ? 0 : (index - 1 > lastI ? lastI : index - 1);
  const long long i1 = index < 0 ? 0 : (index > lastI ? lastI : index);
  const long long i2 = index + 1 > lastI ? lastI : index + 1;
  const long long i3 = index + 2 > lastI ? lastI : index + 2;
  ol = cubic_interp(L[i0], L[i1], L[i2], L[i3], frac);
  orr = cubic_interp(R[i0], R[i1], R[i2], R[i3], frac);
}
G_HD void loop_read_wrapped(const float* L, const float* R, uint32_t len, double position, float& ol, float& orr) {
  if (len == 1) { ol = L[0]; orr = R[0]; return; }
  const double flen = (double)len;
  position = rem_euclid_d(position, flen);
  const long long index = (long long)floor(position);
  const float frac = (float)(position - (double)index);
  const long long n = (long long)len;
  const long long i0 = rem_euclid_i(index - 1, n), i1 = rem_euclid_i(index, n), i2 = rem_euclid_i(index + 1, n), i3 = rem_euclid_i(index + 2, n);
  ol = cubic_interp(L[i0], L[i1], L[i2], L[i3], frac);
  orr = cubic_interp(R[i0], R[i1], R[i2], R[i3], frac);
}

// maybe_swap_pending (:249-276): the queued take lands when this sample crossed a boundary of the loop split into `divisions`
// equal parts (or wrapped); the phrase restarts from the new buffer's loop start and any stretcher is dropped
G_HD void loop_maybe_swap(LoopChan& c, double prev_v, double cur_v, double span, bool wrapped) {
  if (!c.has_pending) return;
  const double grid = (double)(c.pend_div > 1u ? c.pend_div : 1u);
  const double prev_idx = floor((prev_v / span) * grid), new_idx = floor((cur_v / span) * grid);
  if (!(wrapped || new_idx != prev_idx)) return;
  c.left = c.pend_left; c.right = c.pend_right; c.len = c.pend_len; c.buf_sr = c.pend_buf_sr; c.warp = c.pend_warp; c.warp_pp = c.pend_warp_pp;
  c.cursor = loop_window(c.loop_start, c.loop_end, (double)c.len).lo;
  c.st_valid = 0; c.swaps += 1; c.has_pending = 0;
}
// LoopChannel::advance (:233-279)
G_HD void loop_advance(LoopChan& c, float engine_sr) {
  const double len = (double)c.len, source_sr = (double)c.buf_sr;
  const LoopWindow w = loop_window(c.loop_start, c.loop_end, len);
  const double span = w.span > 1.0 ? w.span : 1.0;                 // window.span.max(1.0)
  const double ratio = source_sr / (double)fmaxf(engine_sr, 1.0f);
  const double delta = (double)c.speed * ratio * c.warp;
  const double prev = c.cursor;
  double prev_v, cur_v; bool wrapped;
  if (w.wraps) {
    prev_v = window_to_virtual(w, prev);
    const double raw = prev_v + delta;
    wrapped = !(raw >= 0.0 && raw < span);
    cur_v = rem_euclid_d(raw, span);
    c.cursor = window_to_physical(w, cur_v);
  } else {
    c.cursor += delta;
    wrapped = false;
    if (c.cursor >= w.hi) { c.cursor = w.lo + rem_euclid_d(c.cursor - w.lo, span); wrapped = true; }
    else if (c.cursor < w.lo) { c.cursor = w.hi - rem_euclid_d(w.lo - c.cursor, span); wrapped = true; }
    prev_v = prev - w.lo; cur_v = c.cursor - w.lo;
  }
  loop_maybe_swap(c, prev_v, cur_v, span, wrapped);
}
// ---- mixer/wsola.rs: WSOLA time-stretch (tempo follows engine_bpm / source_bpm, pitch does not) ----------------------------
G_HD uint32_t wsola_hop_len(float engine_sr) {   // :73-74
  const double sr = (double)fmaxf(engine_sr, 1.0f);
  const double h = round(((double)20.0f / 1000.0) * sr);
  return (uint32_t)(h > 1.0 ? h : 1.0);
}
G_HD float raised_sine_window(float phase, float shape) {   // utils/mod.rs:39-44
  return gm::g_powf(fmaxf(gm::g_sinf(PI_F * clampf(phase, 0.0f, 1.0f)), 0.0f), shape);
}
// Periodic Hann (:80-82): raised_sine_window(i / window_len, 2.0).  The exponent is a literal at the (inlined) call site, and both
// LLVM and GCC lower pow(x, 2.0) to x * x whatever the math flags, so the table is sin^2 by one multiplication (correctly rounded;
// glibc's powf(x, 2) differs from it by 1 ulp on about one entry in a thousand).  A debug build of the reference would call powf.
G_HD float wsola_window_coeff(uint32_t i, uint32_t window_len) {
  const float s = fmaxf(gm::g_sinf(PI_F * clampf((float)i / (float)window_len, 0.0f, 1.0f)), 0.0f);
  return s * s;
}
struct WsolaView { float *out_l, *out_r, *grain_l, *grain_r, *tail_l, *tail_r, *mono; };
G_HD WsolaView wsola_view(float* b, uint32_t hop) {
  WsolaView v;
  v.out_l = b; v.out_r = b + hop; v.grain_l = b + 2 * hop; v.grain_r = b + 4 * hop; v.tail_l = b + 6 * hop; v.tail_r = b + 7 * hop; v.mono = b + 8 * hop;
  return v;
}
G_HD double maxd(double a, double b) { return a > b ? a : b; }
G_HD double mind(double a, double b) { return a < b ? a : b; }
// A source read of either synthesis path: physical frames (linear window) or the window's virtual coordinates (wrap-around window)
G_HD void wsola_read(const LoopChan& c, const LoopWindow& w, double pos, float& l, float& r) {
  if (w.wraps) loop_read_wrapped(c.left, c.right, c.len, window_to_physical(w, pos), l, r);
  else loop_read_interpolated(c.left, c.right, c.len, pos, l, r);
}
// score_at (:330-347, :398-415): normalised cross-correlation of the candidate's next hop with the previous grain's tail
G_HD float wsola_score(const LoopChan& c, const LoopWindow& w, const float* mono, double start, double step, double lo_floor, double hi_clamp) {
  float num = 0.0f, ref_energy = 0.0f, cand_energy = 0.0f;
  for (uint32_t i = 0; i < c.hop; i++) {
    const float reference = mono[i];
    float l, r;
    wsola_read(c, w, clampd(start + (double)i * step, lo_floor, hi_clamp), l, r);
    const float cand = l + r;
    num += cand * reference;
    ref_energy += reference * reference;
    cand_energy += cand * cand;
  }
  if (ref_energy <= 1.1920929e-7f || cand_energy <= 1.1920929e-7f) return 0.0f;
  return num / (sqrtf(ref_energy) * sqrtf(cand_energy));
}
// search_best_start / search_best_start_wrapped (:314-380, :386-448): coarse pass of <= 65 candidates, then a +-stride refinement
G_HD double wsola_search(const LoopChan& c, const LoopWindow& w, const float* mono, double center, double step, double lo_floor, double max_start) {
  const double radius = maxd(round(((double)10.0f / 1000.0) * (double)c.buf_sr), 1.0);
  const double lo_bound = maxd(center - radius, lo_floor), hi_bound = mind(center + radius, max_start);
  if (hi_bound <= lo_bound) return clampd(center, lo_floor, max_start);
  const double span = hi_bound - lo_bound;
  const double coarse_stride = maxd(span / 64.0, 1.0);
  double best = lo_bound;
  float best_score = -3.40282347e+38f;
  for (double k = lo_bound; k <= hi_bound; k += coarse_stride) {
    const float sc = wsola_score(c, w, mono, k, step, lo_floor, max_start + step);
    if (sc > best_score) { best_score = sc; best = k; }
  }
  const double refine_lo = maxd(best - coarse_stride, lo_bound), refine_hi = mind(best + coarse_stride, hi_bound);
  for (double k = refine_lo; k <= refine_hi; k += 1.0) {
    const float sc = wsola_score(c, w, mono, k, step, lo_floor, max_start + step);
    if (sc > best_score) { best_score = sc; best = k; }
  }
  return best;
}
// synthesize_next_hop (:121-311): both paths in one body — the linear window works in physical frames [lo, hi], the wrap-around
// window in virtual frames [0, span] read back through to_physical; returns the new (physical) analysis cursor
G_HD double wsola_synthesize(LoopChan& c, const LoopWindow& w, double sr_ratio, double speed, double warp) {
  const WsolaView v = wsola_view(c.st_buf, c.hop);
  const uint32_t hop = c.hop, window_len = 2 * c.hop;
  const double step = maxd(sr_ratio * maxd(speed, 0.0), 1e-6);
  const double hop_source_span = (double)hop * step;
  const double grain_source_span = ((double)window_len - 1.0) * step + 1.0;
  const double lo_floor = w.wraps ? 0.0 : w.lo, hi_ceil = w.wraps ? w.span : w.hi;
  const double max_start = maxd(hi_ceil - grain_source_span, lo_floor);
  const double from = w.wraps ? window_to_virtual(w, c.st_cursor) : c.st_cursor;
  const double raw_target = from + hop_source_span * maxd(warp, 0.0);
  double search_center;
  if (raw_target > max_start || max_start <= lo_floor) { search_center = lo_floor; c.st_have_prev = 0; }   // restart at the loop start with a fresh grain
  else search_center = maxd(raw_target, lo_floor);
  const double best_start = c.st_have_prev ? wsola_search(c, w, v.mono, search_center, step, lo_floor, max_start) : search_center;
  for (uint32_t i = 0; i < window_len; i++) {
    float l, r;
    wsola_read(c, w, clampd(best_start + (double)i * step, lo_floor, hi_ceil), l, r);
    const float wi = c.hann[i];
    v.grain_l[i] = l * wi; v.grain_r[i] = r * wi;
  }
  for (uint32_t i = 0; i < hop; i++) {
    const float pl = c.st_have_prev ? v.tail_l[i] : 0.0f, pr = c.st_have_prev ? v.tail_r[i] : 0.0f;
    v.out_l[i] = pl + v.grain_l[i]; v.out_r[i] = pr + v.grain_r[i];
  }
  for (uint32_t i = 0; i < hop; i++) {
    v.tail_l[i] = v.grain_l[hop + i]; v.tail_r[i] = v.grain_r[hop + i];
    v.mono[i] = v.tail_l[i] + v.tail_r[i];
  }
  c.st_have_prev = 1; c.st_drain = 0;
  const double phys = w.wraps ? window_to_physical(w, best_start) : best_start;
  c.st_cursor = phys;
  return phys;
}
// LoopChannel::tick_preserve_pitch (:215-262, no queued swap): one frame out of the stretcher, a new hop first if it ran dry
G_HD void wsola_tick(LoopChan& c, float engine_sr, float& ol, float& orr) {
  const LoopWindow w = loop_window(c.loop_start, c.loop_end, (double)c.len);
  const double sr_ratio = (double)c.buf_sr / (double)fmaxf(engine_sr, 1.0f);
  if (!c.st_valid) { c.st_valid = 1; c.st_have_prev = 0; c.st_drain = c.hop; c.st_cursor = c.cursor; }   // WsolaStretcher::new(engine rate, cursor)
  const double prev = c.cursor;
  bool wrapped = false;
  if (c.st_drain >= c.hop) {
    c.cursor = wsola_synthesize(c, w, sr_ratio, (double)c.speed, c.warp_pp);
    wrapped = w.wraps ? window_to_virtual(w, c.cursor) < window_to_virtual(w, prev) : c.cursor < prev;     // forward only: moving back = the hop wrapped the loop
  }
  if (c.st_drain < c.hop) { ol = c.st_buf[c.st_drain]; orr = c.st_buf[c.hop + c.st_drain]; }
  else { ol = 0.0f; orr = 0.0f; }
  c.st_drain += 1;
  if (c.has_pending) {   // a queued take lands at hop granularity in this mode (:236-261)
    const double span = w.span > 1.0 ? w.span : 1.0;
    const double prev_v = w.wraps ? window_to_virtual(w, prev) : prev - w.lo, cur_v = w.wraps ? window_to_virtual(w, c.cursor) : c.cursor - w.lo;
    loop_maybe_swap(c, prev_v, cur_v, span, wrapped);
  }
}

// LoopChannel::tick (:181-208) with an empty effect chain (EffectChain::process of no effects is the identity, effect_chain.rs:294-299)
G_HD void loop_chan_tick(LoopChan& c, float engine_sr, float coeff15, float& ol, float& orr) {
  float dl = 0.0f, dr = 0.0f;
  if (c.playing && c.left) {
    if (c.preserve && c.speed >= 0.0f) wsola_tick(c, engine_sr, dl, dr);
    else {
      const LoopWindow w = loop_window(c.loop_start, c.loop_end, (double)c.len);
      if (w.wraps) loop_read_wrapped(c.left, c.right, c.len, c.cursor, dl, dr);
      else loop_read_interpolated(c.left, c.right, c.len, c.cursor, dl, dr);
      loop_advance(c, engine_sr);
    }
  }
  const float g = lsm_tick(c.gain, coeff15);
  const float gl = dl * g, gr = dr * g;
  const float a = lsm_tick(c.active, coeff15);
  ol = gl * a; orr = gr * a;
}
// Mixer::tick (mod.rs:60-75) after the gate targets were written (they only change between render calls); the clip grid holds nothing
G_HD void loop_mixer_tick(LoopMixer& m, float engine_sr, float coeff15, float& ol, float& orr) {
  float l = 0.0f, r = 0.0f;
#pragma unroll
  for (int k = 0; k < LOOP_CHANNELS; k++) {
    float cl, cr;
    loop_chan_tick(m.ch[k], engine_sr, coeff15, cl, cr);
    l += cl; r += cr;
  }
  ol = l; orr = r;
}

// ---- instruments/sampler.rs ----------------------------------------------------------------------------------------
// SamplerBuffer::frame (:64-81): linear read of an interleaved buffer, position clamped to the last frame
struct SampleVoice {          // :84-150
  const float* samples;       // device copy of the slot's interleaved PCM the voice was started on (nullptr: inactive)
  double position, increment;
  unsigned long long age;
  uint32_t frames, channels, slot;
  float velocity;
};
// The rack's step pattern is resolved on the host (its sequencer is the engine's 16th-note sequencer, engine.cuh HostSeq) into hits:
// (frame of the call, pad, velocity); the device starts the voices itself at those frames — it is the side that knows which voices
// still sound (SamplerRack::trigger, :200-223).
struct SamplerHit { uint32_t frame, slot; float velocity; uint32_t pad; };
struct SamplerSlotRef { const float* samples; uint32_t frames, channels; double increment; };   // a loaded pad: increment = pad rate / engine rate
struct SamplerRack {
  SampleVoice v[SAMPLER_VOICES];
  uint32_t row; uint32_t rack;
  SamplerSlotRef slots[SAMPLER_SLOTS];
  const SamplerHit* hits; uint32_t n_hits, next_hit;
  unsigned long long next_age;
  uint32_t cur_frame, pad;      // frames of this call already rendered (the kernel runs once per piece)
};

G_HD void sample_voice_tick(SampleVoice& v, float& ol, float& orr) {
  if (!v.samples) { ol = 0.0f; orr = 0.0f; return; }
  const double pos = clampd(v.position, 0.0, (double)(v.frames - 1));
  const uint32_t i0 = (uint32_t)floor(pos);
  const uint32_t i1 = i0 + 1 < v.frames - 1 ? i0 + 1 : v.frames - 1;
  const float frac = (float)(pos - (double)i0);
  float fl, fr;
  if (v.channels == 1) {
    const float a = v.samples[i0], b = v.samples[i1];
    fl = fr = a + (b - a) * frac;
  } else {
    const float a = v.samples[2 * i0], b = v.samples[2 * i1];
    fl = a + (b - a) * frac;
    const float c = v.samples[2 * i0 + 1], d = v.samples[2 * i1 + 1];
    fr = c + (d - c) * frac;
  }
  const double fade = 32.0, end = (double)v.frames;               // fixed click guard (:134-141)
  const double tail = (end - v.position) / fade;
  const double g = fmin(fmin(v.position / fade, tail > 0.0 ? tail : 0.0), 1.0);
  const float gain = (float)g * v.velocity;
  v.position += v.increment;
  if (v.position >= end) v.samples = nullptr;
  ol = fl * gain; orr = fr * gain;
}
// SamplerRack::trigger (:200-223): first free voice, else the oldest (first of equals); false when the pad is empty
G_HD bool sampler_rack_trigger(SamplerRack& r, uint32_t slot, float velocity) {
  if (slot >= (uint32_t)SAMPLER_SLOTS || !r.slots[slot].samples) return false;
  int vi = -1;
  for (int k = 0; k < SAMPLER_VOICES; k++) if (!r.v[k].samples) { vi = k; break; }
  if (vi < 0) { vi = 0; for (int k = 1; k < SAMPLER_VOICES; k++) if (r.v[k].age < r.v[vi].age) vi = k; }
  r.next_age += 1;
  SampleVoice& v = r.v[vi];
  const SamplerSlotRef& s = r.slots[slot];
  v.samples = s.samples; v.frames = s.frames; v.channels = s.channels; v.slot = slot;
  v.position = 0.0; v.increment = s.increment; v.velocity = clampf(velocity, 0.0f, 1.0f); v.age = r.next_age;
  return true;
}
// SamplerRack::tick (:225-229): fold over the 32 voices in index order — after the pattern hits due at this frame (ffi.rs:1199-1204)
G_HD void sampler_rack_tick(SamplerRack& r, float& ol, float& orr) {
  while (r.next_hit < r.n_hits && r.hits[r.next_hit].frame <= r.cur_frame) { sampler_rack_trigger(r, r.hits[r.next_hit].slot, r.hits[r.next_hit].velocity); r.next_hit++; }
  r.cur_frame++;
  float l = 0.0f, rr = 0.0f;
  for (int k = 0; k < SAMPLER_VOICES; k++) {
    float vl, vr;
    sample_voice_tick(r.v[k], vl, vr);
    l += vl; rr += vr;
  }
  ol = l; orr = rr;
}

#ifdef __CUDACC__
// Stereo rows of the sample-playback sources of one piece: descriptor with row r writes ext[2 * r][frame] (left) and
// ext[2 * r + 1][frame] (right); MixCfg::ext_row tells the mixer which pair belongs to which (engine, graph source).
// One thread per descriptor; a warp stages 32 frames of its 32 descriptors' stereo output in shared memory (pitch 33) and
// writes each row's 32 frames with one coalesced store.
struct ExtTickCtx { float engine_sr, coeff15; };
__device__ __forceinline__ void ext_tick(LoopMixer& m, const ExtTickCtx& c, float& l, float& r) { loop_mixer_tick(m, c.engine_sr, c.coeff15, l, r); }
__device__ __forceinline__ void ext_tick(SamplerRack& m, const ExtTickCtx&, float& l, float& r) { sampler_rack_tick(m, l, r); }

template <class Desc>
__global__ void __launch_bounds__(64) ext_source_kernel(Desc* descs, int n, float* ext, long long stride, int frames, ExtTickCtx ctx) {
  __shared__ float tile_l[2][32][33], tile_r[2][32][33];
  float (*tl)[33] = tile_l[threadIdx.x >> 5];
  float (*tr)[33] = tile_r[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  Desc d;
  unsigned row_mine = 0u;
  if (valid) { d = descs[i]; row_mine = d.row; }
  for (int t0 = 0; t0 < frames; t0 += 32) {
    const int nf = min(32, frames - t0);
    if (valid)
      for (int j = 0; j < nf; j++) { float l, r; ext_tick(d, ctx, l, r); tl[lane][j] = l; tr[lane][j] = r; }
    __syncwarp();
    for (int q = 0; q < 32; q++) {
      const unsigned row = __shfl_sync(0xffffffffu, row_mine, q);
      const int ok = __shfl_sync(0xffffffffu, (int)valid, q);
      if (!ok || lane >= nf) continue;
      float* pl = ext + (long long)(2u * row) * stride + t0 + lane;
      pl[0] = tl[q][lane]; pl[stride] = tr[q][lane];
    }
    __syncwarp();
  }
  if (valid) descs[i] = d;
}
#endif

}  // namespace gd
