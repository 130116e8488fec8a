// chain.cuh — the global effect chain of SETTLED engines: tilt filter, delay and spring reverb (any order, any subset) + soft limiter.
//
// mix_kernel (mix.cuh) is the general per-engine mixer: events, racks, every effect kind, gliding smoothers, state in shared
// memory behind run-time slot indices — ~1000+ warp-instructions per frame of its 32 engines on a single warp per CTA (profiles/r2_mix_*).  Once
// the parameter smoothers of an engine's chain have settled (SmoothedParam::tick returns the same value every sample,
// smoother.rs:120-137) the chain of ffi.rs:1317-1364 is the same few recurrences every frame with constant coefficients.  This
// kernel runs exactly those, in the reference's operation order (effects/tilt_filter.rs:87-139, delay.rs:321-491,
// reverb.rs:162-217, limiter.rs), with
//   * one engine per thread, both channels interleaved in registers (two independent dependency chains per thread),
//   * every delay-line read of the NEXT 16-frame tile already in flight (cp.async into shared memory, double buffered): ring
//     reads are addressed by state that advances one word per frame, never by audio, and every line is longer than two tiles,
//     so the words a tile reads were written before the previous tile started;
//   * ring words laid out [word][engine] (mix.cuh), so the 32 engines of a warp read and write one 128-byte row per access;
//   * the pre-chain stereo mix (strips -> graph -> master, mix_fast_kernel) read through the same pipeline and the output
//     written through a shared-memory tile, 64-byte row segments.
// (Tried and dropped: requesting a line's cell of frame n + 4 straight into the register frame n has just consumed, no shared-memory
// staging — half the instructions, but slower, 2.0 vs 1.25 us per frame alone: a warp has six scoreboards, so a wait on the oldest
// of 56 outstanding loads is a wait on the newest one sharing its scoreboard and the kernel runs at DRAM latency.  cp.async groups
// do not have that limit.)
// A warp takes its 32 engines only if ALL of them qualify with the same chain (same effects in the same order); it then marks
// them done (fast[i] = 1) and mix_kernel, launched next on the same stream, skips them.  Everything else — first pieces of a
// bounce while FFI edits glide, racks, the other effect kinds, ping-pong, mid-piece events — stays with mix_kernel.
#pragma once
#include <cuda_pipeline.h>
#include "mix.cuh"

namespace gd {
#ifdef __CUDACC__

constexpr int CF_T = 16;     // frames per tile
constexpr int CF_NL = 14;    // ring words read per engine-frame: delay L, R; spring L0..5, R0..5
struct ChainSmem {
  float ring[2][CF_T][CF_NL][32];      // [stage][frame][line][engine]
  float pre[2][2][32][CF_T + 1];       // [stage][plane][engine][frame]
  float out[2][32][CF_T + 1];          // [channel][engine][frame]
};
constexpr size_t CHAIN_SMEM = sizeof(ChainSmem);

#define MS_W(member) ((int)(offsetof(MixState, member) / 4))
// sm_set would not move the target and sm_tick is the identity: either cur == tgt, or cur is a FIXED POINT of the tick one
// rounding step short of the target (a cutoff of 8 kHz has an ulp of 5e-4: `cur + coeff * (tgt - cur)` rounds back to cur and
// the 1e-4 snap of smoother.rs:131 never fires — the reference keeps returning cur, and so does this).
__device__ __forceinline__ bool sm_settled(const Sm& s, float target, float lo, float hi, float coeff) {
  if (fabsf(s.t - clampf(target, lo, hi)) > 1e-8f) return false;
  float c = s.c;
  smooth_tick(c, s.t, coeff);
  return __float_as_uint(c) == __float_as_uint(s.c);
}

__global__ void __launch_bounds__(32) chain_fast_kernel(const MixLaunch L) {
  extern __shared__ __align__(16) unsigned char chain_smem_raw[];
  ChainSmem& S = *reinterpret_cast<ChainSmem*>(chain_smem_raw);
  const int lane = threadIdx.x;
  const int warp_i0 = blockIdx.x * 32;
  const int i = warp_i0 + lane;
  if (warp_i0 >= L.n || !L.premix) return;
  const bool valid = i < L.n;
  const int n_rows = min(32, L.n - warp_i0);
  const RateCtx& rc = L.rc;
  const FxGeom& geo = L.geo;
  // ---- eligibility ----
  bool ok = true;
  uint32_t sig = 0;
  int es = 0;
  uint32_t limiter_on = 0; float lim_th = 0.0f, lim_inv = 0.0f;
  if (valid) {
    ok = L.fast[i] == 2;
    es = (int)L.slots[i];
    const MixCfg& cfg = L.cfg[es];
    limiter_on = cfg.limiter_on; lim_th = cfg.lim_th; lim_inv = cfg.lim_inv;
    for (int o = 0; o < 9 && ok; o++) {
      const uint32_t id = cfg.order[o];
      const uint32_t slot = id < 12u ? cfg.gslot[id] : 0xffu;
      if (slot >= (uint32_t)MAX_FX || !cfg.fx_enabled[slot]) continue;
      const uint32_t want = id == FXK_TILT ? FXS_TILT : id == FXK_DELAY ? FXS_DELAY : id == FXK_SPRING ? FXS_SPRING : 0xffu;
      ok = ok && want != 0xffu && slot == want && cfg.fx_kind[slot] == id && (id == FXK_TILT || L.ring[slot] != nullptr);   // the tilt filter has no delay line
      sig = (sig << 4) | (id + 1u);
    }
  }
  const uint32_t sig0 = __shfl_sync(0xffffffffu, sig, 0);
  ok = ok && (!valid || sig == sig0);
  if (!__all_sync(0xffffffffu, ok) || sig0 == 0u) return;
  bool has_tilt = false, has_delay = false, has_spring = false;
  int n_fx = 0;
  for (uint32_t s = sig0; s; s >>= 4, n_fx++) { const uint32_t id = (s & 15u) - 1u; has_tilt |= id == FXK_TILT; has_delay |= id == FXK_DELAY; has_spring |= id == FXK_SPRING; }

  // ---- constants and carried state (what tilt_one / delay_prep / spring_one compute per frame from settled smoothers) ----
  auto stw = [&](int w, uint32_t v) { L.state[(size_t)w * L.state_cap + es] = v; };
  auto stf = [&](int w, float v) { stw(w, __float_as_uint(v)); };
  // tilt
  bool t_on[2] = {false, false}, t_lp[2] = {false, false};
  float t_keep[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};   // (cutoff, res, g, r, h) of the two sections after tpt_set: written back if the warp proceeds
  float t_mix[2] = {0, 0}, t_g[2] = {0, 0}, t_h[2] = {0, 0}, t_r[2] = {0, 0}, t_ic1[2] = {0, 0}, t_ic2[2] = {0, 0};
  if (has_tilt && valid) {
    TiltDyn d;
    load_words(d, L.state, es, L.state_cap, MS_W(fx[FXS_TILT]));
#pragma unroll
    for (int c = 0; c < 2; c++) {
      ok = ok && sm_settled(d.cutoff[c], d.cutoff_target, 0.0f, 1.0f, rc.smooth30) && sm_settled(d.res[c], d.res_target, 0.0f, 1.0f, rc.smooth30);
      const float knob = d.cutoff[c].c, resonance = d.res[c].c;
      const float pw = knob < 0.5f ? gm::g_powf(20000.0f / 80.0f, knob * 2.0f) : gm::g_powf(8000.0f / 20.0f, (knob - 0.5f) * 2.0f);
      float mix, freq;
      if (knob < 0.5f) { mix = 1.0f - (knob * 2.0f); freq = 80.0f * pw; t_lp[c] = true; } else { mix = (knob - 0.5f) * 2.0f; freq = 20.0f * pw; t_lp[c] = false; }
      t_mix[c] = mix;
      t_on[c] = !(mix < 0.001f);
      if (t_on[c]) {
        const float q = 0.5f + resonance * 8.0f;
        tpt_set(d.svf[c], rc.sr, freq, q);       // hysteresis on the stored (cutoff, res): applied once, then idempotent
        t_g[c] = d.svf[c].g; t_h[c] = d.svf[c].h; t_r[c] = d.svf[c].r; t_ic1[c] = d.svf[c].ic1; t_ic2[c] = d.svf[c].ic2;
      }
    }
#pragma unroll
    for (int c = 0; c < 2; c++) { t_keep[c][0] = d.svf[c].cutoff; t_keep[c][1] = d.svf[c].res; t_keep[c][2] = d.svf[c].g; t_keep[c][3] = d.svf[c].r; t_keep[c][4] = d.svf[c].h; }
  }
  // delay
  uint32_t d_wi[2] = {0, 0}, d_di[2] = {0, 0};
  float d_df[2] = {0, 0}, d_fb[2] = {0, 0}, d_mix[2] = {0, 0}, d_g[2] = {0, 0}, d_z1[2] = {0, 0}, d_z2[2] = {0, 0}, d_s2[2] = {0, 0};
  const uint32_t dlen = geo.delay_len;
  if (has_delay && valid) {
    DelayDyn d;
    load_words(d, L.state, es, L.state_cap, MS_W(fx[FXS_DELAY]));
    const uint32_t tc = d.timing_target;
    const float time_target = delay_seconds(tc <= 8 ? tc : 2, d.bpm_target);
    ok = ok && !d.pingpong;
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const DelayCh& s = d.ch[c];
      ok = ok && tc == s.prev_timing && sm_settled(s.time, time_target, 0.0f, 5.0f, rc.smooth50) && sm_settled(s.fb, d.fb_target, 0.0f, 0.95f, rc.smooth30)
              && sm_settled(s.mix, d.mix_target, 0.0f, 1.0f, rc.smooth30) && sm_settled(s.cutoff, d.cutoff_target, 20.0f, 20000.0f, rc.smooth30);
      const float ds = s.time.c * rc.sr;
      const uint32_t di = (uint32_t)f32_to_u64_sat(ds);
      d_di[c] = di; d_df[c] = ds - (float)di;
      ok = ok && di >= 2u * CF_T && di < dlen && s.write_index < dlen;
      d_fb[c] = s.fb.c; d_mix[c] = s.mix.c;
      d_g[c] = 1.0f - gm::g_expf(-2.0f * PI_F * s.cutoff.c / rc.sr);
      d_wi[c] = s.write_index; d_z1[c] = s.z1; d_z2[c] = s.z2;
    }
  }
  // spring
  uint32_t s_idx[2][6];
  float s_fbk[2] = {0, 0}, s_mix[2] = {0, 0}, s_d1[2] = {0, 0}, s_d2[2] = {0, 0}, s_fb[2] = {0, 0}, s_damp[2] = {0, 0};
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int k = 0; k < 6; k++) s_idx[c][k] = 0;
  if (has_spring) {
    for (int k = 0; k < 12; k++) ok = ok && geo.spring_len[k] >= 2u * CF_T;
    if (valid) {
      SpringDyn d;
      load_words(d, L.state, es, L.state_cap, MS_W(fx[FXS_SPRING]));
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const SpringCh& s = d.ch[c];
        ok = ok && sm_settled(s.decay, d.decay_target, 0.0f, 1.0f, rc.smooth15) && sm_settled(s.mix, d.mix_target, 0.0f, 1.0f, rc.smooth15) && sm_settled(s.damping, d.damping_target, 0.0f, 1.0f, rc.smooth15);
        s_fbk[c] = gm::g_powf(s.decay.c, 0.4f) * 0.95f;
        s_mix[c] = s.mix.c; s_d1[c] = s.damping.c; s_d2[c] = 1.0f - s.damping.c;
        s_fb[c] = s.fb; s_damp[c] = s.damp;
#pragma unroll
        for (int k = 0; k < 6; k++) { s_idx[c][k] = s.idx[k]; ok = ok && s.idx[k] < geo.spring_len[c * 6 + k]; }
      }
    }
  }
  if (!__all_sync(0xffffffffu, ok)) return;
  __syncwarp();
  // the warp takes its engines
  if (lane == 0 && L.chain_units) atomicAdd(L.chain_units, (unsigned long long)n_rows * (unsigned long long)L.frames);
  if (valid) {
    L.fast[i] = 1;
    if (has_tilt) {
#pragma unroll
      for (int c = 0; c < 2; c++)
        if (t_on[c]) {
          const int w = c ? MS_W(fx[FXS_TILT].tilt.svf[1]) : MS_W(fx[FXS_TILT].tilt.svf[0]);
          stf(w + 0, t_keep[c][0]); stf(w + 1, t_keep[c][1]); stf(w + 2, t_keep[c][2]); stf(w + 3, t_keep[c][3]); stf(w + 4, t_keep[c][4]);
        }
    }
  }
  __syncwarp();

  // ---- ring cursors, as words of the slot's arena (line offset included); the bases carry the engine column ----
  const long long cap = L.ring_cap;
  float* const dring = L.ring[FXS_DELAY] ? L.ring[FXS_DELAY] + es : nullptr;
  float* const sring = L.ring[FXS_SPRING] ? L.ring[FXS_SPRING] + es : nullptr;
  uint32_t d_pf[2] = {0, 0};           // prefetch cursor of the delay read
  if (has_delay && valid) {
#pragma unroll
    for (int c = 0; c < 2; c++) {
      d_pf[c] = c * dlen + wrap2(d_wi[c] + dlen - d_di[c], dlen);
      d_s2[c] = dring[(long long)(c * dlen + wrap2(d_wi[c] + dlen - d_di[c] - 1u, dlen)) * cap];   // the frame-0 second tap; later frames carry it
      d_wi[c] += c * dlen;
    }
  }
  uint32_t s_pf[2][6], s_w[2][6];      // prefetch / write cursors of the spring lines
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int k = 0; k < 6; k++) s_pf[c][k] = s_w[c][k] = geo.spring_off[c * 6 + k] + s_idx[c][k];
#define CF_NEXT(w, lo, len) ((w) + 1u == (lo) + (len) ? (lo) : (w) + 1u)

  const int frames = L.frames;
  const int ntiles = (frames + CF_T - 1) / CF_T;
  const float* const pre_l = L.premix;
  const float* const pre_r = L.premix + (long long)L.n_lpad * L.premix_stride;
  auto issue = [&](int tile) {
    const int st = tile & 1;
    const int f0 = tile * CF_T;
    const int nf = min(CF_T, frames - f0);
    if (valid) {
      for (int j = 0; j < nf; j++) {
        if (has_delay) {
#pragma unroll
          for (int c = 0; c < 2; c++) {
            __pipeline_memcpy_async(&S.ring[st][j][c][lane], dring + (long long)d_pf[c] * cap, 4);
            d_pf[c] = CF_NEXT(d_pf[c], c * dlen, dlen);
          }
        }
        if (has_spring) {
#pragma unroll
          for (int c = 0; c < 2; c++)
#pragma unroll
            for (int k = 0; k < 6; k++) {
              __pipeline_memcpy_async(&S.ring[st][j][2 + c * 6 + k][lane], sring + (long long)s_pf[c][k] * cap, 4);
              s_pf[c][k] = CF_NEXT(s_pf[c][k], geo.spring_off[c * 6 + k], geo.spring_len[c * 6 + k]);
            }
        }
      }
    }
    // pre-chain mix: half a warp per row, 16 consecutive frames
    const int jj = lane & (CF_T - 1);
    if (jj < nf) {
#pragma unroll 4
      for (int q = 0; q < 16; q++) {
        const int r = (lane >> 4) + 2 * q;
        if (r < n_rows) {
          const long long off = (long long)(warp_i0 + r) * L.premix_stride + f0 + jj;
          __pipeline_memcpy_async(&S.pre[st][0][r][jj], pre_l + off, 4);
          __pipeline_memcpy_async(&S.pre[st][1][r][jj], pre_r + off, 4);
        }
      }
    }
    __pipeline_commit();
  };

  issue(0);
  for (int tile = 0; tile < ntiles; tile++) {
    if (tile + 1 < ntiles) issue(tile + 1); else __pipeline_commit();
    __pipeline_wait_prior(1);
    __syncwarp();
    const int st = tile & 1;
    const int f0 = tile * CF_T;
    const int nf = min(CF_T, frames - f0);
    if (valid) {
      for (int j = 0; j < nf; j++) {
        float x[2] = {S.pre[st][0][lane][j], S.pre[st][1][lane][j]};
#pragma unroll 1
        for (int e = n_fx - 1; e >= 0; e--) {      // the signature holds the chain most-significant nibble first
          {
            const uint32_t id = ((sig0 >> (4 * e)) & 15u) - 1u;
            if (id == FXK_TILT) {
#pragma unroll
              for (int c = 0; c < 2; c++) {
                if (!t_on[c]) continue;
                const float in = x[c];
                const float v1 = (t_g[c] * (in - t_ic2[c]) + t_ic1[c]) * t_h[c];
                const float v2 = t_ic2[c] + t_g[c] * v1;
                t_ic1[c] = 2.0f * v1 - t_ic1[c];
                t_ic2[c] = 2.0f * v2 - t_ic2[c];
                const float hi = in - (t_r[c] * v1 + v2);
                const float wet = t_lp[c] ? v2 : hi;
                float out = in * (1.0f - t_mix[c]) + wet * t_mix[c];
                if (!isfinite(out)) { t_ic1[c] = t_ic2[c] = 0.0f; out = 0.0f; }
                else if (fabsf(out) < 1e-15f) out = 0.0f;
                x[c] = out;
              }
            } else if (id == FXK_DELAY) {
#pragma unroll
              for (int c = 0; c < 2; c++) {
                const float dry = isfinite(x[c]) ? x[c] : 0.0f;
                const float s1 = S.ring[st][j][c][lane];
                const float delayed = s1 * (1.0f - d_df[c]) + d_s2[c] * d_df[c];
                d_s2[c] = s1;
                const float rfb = 0.3f * (d_z1[c] - d_z2[c]);
                d_z1[c] = d_z1[c] + d_g[c] * (delayed + rfb - d_z1[c]);
                d_z2[c] = d_z2[c] + d_g[c] * (d_z1[c] - d_z2[c]);
                const float filtered = d_z2[c];
                if (fabsf(d_z1[c]) < 1e-15f) d_z1[c] = 0.0f;
                if (fabsf(d_z2[c]) < 1e-15f) d_z2[c] = 0.0f;
                float w = dry + filtered * d_fb[c];
                w = (isfinite(w) && fabsf(w) > 1e-15f) ? w : 0.0f;
                dring[(long long)d_wi[c] * cap] = w;
                d_wi[c] = CF_NEXT(d_wi[c], c * dlen, dlen);
                const float out = dry * (1.0f - d_mix[c]) + filtered * d_mix[c];
                x[c] = isfinite(out) ? out : dry;
              }
            } else {   // FXK_SPRING
              const float G[6] = {0.70f, 0.68f, 0.65f, 0.62f, 0.60f, 0.58f};
#pragma unroll
              for (int c = 0; c < 2; c++) {
                const float in = isfinite(x[c]) ? x[c] : 0.0f;
                float sgl = in + s_fb[c];
#pragma unroll
                for (int k = 0; k < 6; k++) {
                  const float delayed = S.ring[st][j][2 + c * 6 + k][lane];
                  const float v = sgl - G[k] * delayed;
                  sgl = G[k] * v + delayed;
                  sring[(long long)s_w[c][k] * cap] = v;
                  s_w[c][k] = CF_NEXT(s_w[c][k], geo.spring_off[c * 6 + k], geo.spring_len[c * 6 + k]);
                }
                s_damp[c] = sgl * s_d2[c] + s_damp[c] * s_d1[c];
                if (fabsf(s_damp[c]) < 1e-15f) s_damp[c] = 0.0f;
                s_fb[c] = s_damp[c] * s_fbk[c];
                if (fabsf(s_fb[c]) < 1e-15f) s_fb[c] = 0.0f;
                const float res = in * (1.0f - s_mix[c]) + sgl * s_mix[c];
                x[c] = isfinite(res) ? res : in;
              }
            }
          }
        }
        if (limiter_on) { x[0] = gm::g_tanhf(x[0] * lim_inv) * lim_th; x[1] = gm::g_tanhf(x[1] * lim_inv) * lim_th; }
        if (L.out_mode == 0) S.out[0][lane][j] = 0.5f * (x[0] + x[1]);
        else { S.out[0][lane][j] = x[0]; S.out[1][lane][j] = x[1]; }
      }
    }
    __syncwarp();
    if (L.out_mode == 0) {
      const int jj = lane & (CF_T - 1);
      if (jj < nf) {
#pragma unroll 4
        for (int q = 0; q < 16; q++) {
          const int r = (lane >> 4) + 2 * q;
          if (r < n_rows) {
            const long long row = L.out_rows ? (long long)L.out_rows[warp_i0 + r] : (long long)(warp_i0 + r);
            L.out[row * L.out_stride + f0 + jj] = S.out[0][r][jj];
          }
        }
      }
    } else {
      for (int r = 0; r < n_rows; r++) {
        const long long row = L.out_rows ? (long long)L.out_rows[warp_i0 + r] : (long long)(warp_i0 + r);
        float* dst = L.out + row * L.out_stride + 2LL * f0;
        if (lane < 2 * nf) dst[lane] = S.out[lane & 1][r][lane >> 1];
      }
    }
    __syncwarp();
  }
  __pipeline_wait_prior(0);
  // cursors back to the state's own units
#pragma unroll
  for (int c = 0; c < 2; c++) {
    d_wi[c] -= c * dlen;
#pragma unroll
    for (int k = 0; k < 6; k++) s_idx[c][k] = s_w[c][k] - geo.spring_off[c * 6 + k];
  }
#undef CF_NEXT

  // ---- carried state back to the pool ----
  if (valid) {
    if (has_tilt) {
#pragma unroll
      for (int c = 0; c < 2; c++)
        if (t_on[c]) {
          const int w = c ? MS_W(fx[FXS_TILT].tilt.svf[1]) : MS_W(fx[FXS_TILT].tilt.svf[0]);
          stf(w + 5, t_ic1[c]); stf(w + 6, t_ic2[c]);
        }
    }
    if (has_delay) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int w = c ? MS_W(fx[FXS_DELAY].delay.ch[1]) : MS_W(fx[FXS_DELAY].delay.ch[0]);
        stw(w + 0, d_wi[c]); stf(w + 1, d_z1[c]); stf(w + 2, d_z2[c]);
      }
    }
    if (has_spring) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int w = c ? MS_W(fx[FXS_SPRING].spring.ch[1]) : MS_W(fx[FXS_SPRING].spring.ch[0]);
#pragma unroll
        for (int k = 0; k < 6; k++) stw(w + k, s_idx[c][k]);
        stf(w + 6, s_fb[c]); stf(w + 7, s_damp[c]);
      }
    }
  }
}
#undef MS_W
#endif  // __CUDACC__
}  // namespace gd
