// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE ONLY.
// C entry points over the CPU restatement so tests/ and bench.py's cpu_baseline /
// --impl reference legs can drive it through ctypes.  Never linked into or called
// by the product library.
#include <thread>
#include <atomic>
#include <memory>
#include "drums.hpp"
#include "../include/gooey_batch.h"

using namespace orc;

static std::unique_ptr<Instrument> make_voice(const GooeyVoicePatch& p, float sr) {
  switch (p.instrument) {
    case GOOEY_INSTRUMENT_KICK: {
      KickConfig c;
      for (int i = 0; i < 18; i++) c.v[i] = clampf(p.params[i], 0.0f, 1.0f);
      auto k = std::make_unique<KickDrum>(sr, c);
      if (p.aux & 0x100) k->p[K_TUNING].set_immediate(p.params[23]);
      return k;
    }
    case GOOEY_INSTRUMENT_SNARE: {
      const float* a = p.params;
      float ft = a[13];
      int t = !(ft == ft) ? 0 : (ft <= 0.0f ? 0 : (ft >= 255.0f ? 255 : (int)ft));
      SnareConfig c = SnareConfig::full(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], (uint8_t)t,
                                        a[14], a[15], a[16], a[17], a[18]);
      auto s = std::make_unique<SnareDrum>(sr, c);
      if (p.aux & 0x100) s->p[S_TUNING].set_immediate(p.params[23]);
      return s;
    }
    case GOOEY_INSTRUMENT_HIHAT: {
      HiHat2Config c = HiHat2Config::make(p.params[0], p.params[1], p.params[2], (p.aux & 1) != 0, (p.aux & 2) == 0, p.params[3]);
      c.volume = clampf(p.params[4], 0.0f, 1.0f);
      auto h = std::make_unique<HiHat2>(sr, c);
      if (p.aux & 0x100) h->p[H_TUNING].set_immediate(p.params[23]);
      return h;
    }
    case GOOEY_INSTRUMENT_TOM: {
      auto t = std::make_unique<Tom2>(sr);
      if (p.aux & 1) t->set_config({p.params[0], p.params[1], p.params[2], p.params[3], p.params[4], p.params[5], p.params[6], p.params[7]});
      if (p.aux & 0x100) t->tuning = clampf(p.params[23], 0.0f, 1.0f);
      return t;
    }
    default: return nullptr;
  }
}

extern "C" {

// Events: kind 0 = trigger(value = velocity), 1 = set_param(param, value), 2 = set_param + snap_params.
// Events of one voice must be sorted by frame.  out[v * frames + i].  Returns 0, or -1 on a bad patch.
int orc_render_voices(const GooeyVoicePatch* patches, uint32_t n, float sr, uint32_t frames, uint32_t n_events,
                      const uint32_t* ev_voice, const uint32_t* ev_frame, const uint32_t* ev_kind, const uint32_t* ev_param,
                      const float* ev_value, float* out, int n_threads) {
  std::vector<std::vector<uint32_t>> per_voice(n);
  for (uint32_t e = 0; e < n_events; e++) if (ev_voice[e] < n) per_voice[ev_voice[e]].push_back(e);
  std::atomic<uint32_t> next{0};
  std::atomic<int> bad{0};
  auto work = [&]() {
    for (;;) {
      uint32_t v = next.fetch_add(1);
      if (v >= n) break;
      auto inst = make_voice(patches[v], sr);
      if (!inst) { bad = 1; continue; }
      double t = 0.0;
      const double dt = 1.0 / (double)sr;
      size_t cur = 0;
      const auto& evs = per_voice[v];
      float* o = out + (size_t)v * frames;
      for (uint32_t i = 0; i < frames; i++) {
        while (cur < evs.size() && ev_frame[evs[cur]] <= i) {
          uint32_t e = evs[cur++];
          if (ev_kind[e] == 0) inst->trigger_with_velocity(t, ev_value[e]);
          else { inst->set_param(ev_param[e], ev_value[e]); if (ev_kind[e] == 2) inst->snap_params(); }
        }
        o[i] = inst->tick(t);
        t += dt;
      }
    }
  };
  if (n_threads <= 1) work();
  else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; i++) th.emplace_back(work);
    for (auto& x : th) x.join();
  }
  return bad ? -1 : 0;
}

// ---- pins used by tests/test_oracle_pins.py ------------------------------------------------------
uint64_t orc_siphash(uint64_t m, uint64_t k0, uint64_t k1, int c, int d) { return siphash_u64(m, k0, k1, c, d); }
float orc_hash_noise(uint64_t idx) { return hash_noise(idx); }
float orc_max_curve(float p, float c) { return max_curve(p, c); }
void orc_halfband_coefs(float* out8) { const float* c = halfband_coefs(); for (int i = 0; i < 8; i++) out8[i] = c[i]; }
float orc_smoother_coeff(float sr, float ms) { return SmoothedParam::calculate_coeff(sr, ms); }
void orc_click_table(float* out64) { for (int i = 0; i < 64; i++) out64[i] = TOM_IMPULSE[i]; }
// Oversampler: mode 0/2/4, f(x) = tanh(x*drive) (drive<=0: identity); processes n samples.
void orc_oversample(int mode, float drive, const float* in, float* out, uint32_t n) {
  Oversampler os;
  os.set_mode(mode == 0 ? OversamplingMode::Off : (mode == 2 ? OversamplingMode::X2 : OversamplingMode::X4));
  for (uint32_t i = 0; i < n; i++) out[i] = os.process(in[i], [&](float x) { return drive > 0.0f ? tanhf(x * drive) : x; });
}
void orc_pink(float sr, float* out, uint32_t n) { PinkNoise p(sr); for (uint32_t i = 0; i < n; i++) out[i] = p.tick(); }

}  // extern "C"
