// gooey_b200.cu — libgooey_b200.so: the C ABI (include/gooey_batch.h, include/gooey.h)
// over the sm_100a kernels.  Single translation unit: the kernels are header
// templates (kernels.cuh), this file instantiates them and owns the host side.
#include <mutex>
#include <memory>
#include "voice_batch.cuh"
#include "halfband_design.h"

namespace gh {

std::string& last_error() { static thread_local std::string e; return e; }
std::atomic<uint64_t> g_launches{0};
static float g_last_kernel_ms = 0.0f;

// ---- device bring-up --------------------------------------------------------------------------
static std::mutex g_dev_mutex;
static std::vector<char> g_dev_ready;

static int device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// Select `device` and make sure its constant memory holds the half-band coefficients.
static void use_device(int device) {
  int n = device_count();
  if (n <= 0) throw CudaError("no CUDA device visible: libgooey_b200 has no CPU fallback");
  if (device < 0 || device >= n) throw CudaError("device index out of range");
  GH_CUDA(cudaSetDevice(device));
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  if ((int)g_dev_ready.size() < n) g_dev_ready.resize(n, 0);
  if (!g_dev_ready[device]) {
    float hb[8];
    gd::design_halfband8(hb);
    GH_CUDA(cudaMemcpyToSymbol(gd::c_hb, hb, sizeof hb));
    g_dev_ready[device] = 1;
  }
}

// FFI parameter id -> smoother index (ffi.rs:168-250 with ids ffi.rs:1737-1836)
static const int kKickFfi[8] = {gd::K_FREQ, gd::K_PUNCH, gd::K_SUB, gd::K_CLICK, gd::K_OSC_DECAY, gd::K_PITCH_ENV_AMT, gd::K_VOLUME, gd::K_TUNING};
static const int kSnareFfi[20] = {gd::S_FREQ, gd::S_DECAY, gd::S_BRIGHTNESS, gd::S_VOLUME, gd::S_TONAL, gd::S_NOISE, gd::S_PITCH_DROP,
                                  gd::S_TONAL_DECAY, gd::S_NOISE_DECAY, gd::S_NOISE_TAIL_DECAY, gd::S_FILTER_CUTOFF, gd::S_FILTER_RES, -1,
                                  gd::S_XFADE, gd::S_PHASE_MOD, gd::S_OVERDRIVE, gd::S_AMP_DECAY, gd::S_AMP_DECAY_CURVE,
                                  gd::S_TONAL_DECAY_CURVE, gd::S_TUNING};
static const int kHatFfi[6] = {gd::H_PITCH, gd::H_DECAY, gd::H_ATTACK, gd::H_TONE, gd::H_VOLUME, gd::H_TUNING};

static inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

// Translate `ChannelInstrument::set_param(param, value)` into voice events.  Returns false for unknown ids
// (the reference ignores them silently).
template <class AddFn> static bool ffi_param_to_events(uint32_t instrument, uint32_t param, float value, AddFn add) {
  switch (instrument) {
    case GOOEY_INSTRUMENT_KICK:
      if (param >= 8) return false;
      add(gd::EV_SET_TARGET, kKickFfi[param], clamp01(value));
      return true;
    case GOOEY_INSTRUMENT_SNARE:
      if (param >= 20) return false;
      if (param == 12) {  // `value as u8` (saturating) then .min(3)
        int t = !(value == value) ? 0 : (value <= 0.0f ? 0 : (value >= 255.0f ? 255 : (int)value));
        add(gd::EV_SET_AUX, gd::AUX_SNARE_FILTER_TYPE, (float)(t > 3 ? 3 : t));
      } else add(gd::EV_SET_TARGET, kSnareFfi[param], clamp01(value));
      return true;
    case GOOEY_INSTRUMENT_HIHAT:
      if (param >= 6) return false;
      add(gd::EV_SET_TARGET, kHatFfi[param], clamp01(value));
      return true;
    case GOOEY_INSTRUMENT_TOM:
      if (param >= 9) return false;
      add(gd::EV_SET_TARGET, param, param == 8 ? clamp01(value) : clamp01(value) * 100.0f);
      return true;
    default: return false;
  }
}

}  // namespace gh

using namespace gh;

// =================================================================================================
// Voice batch
// =================================================================================================
struct GooeyVoiceBatch {
  int device = 0;
  float sr = 44100.0f;
  gd::RateCtx rc;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  uint32_t n = 0;
  // voice v -> (type, index in its group); rows are regrouped so each group is contiguous: kick rows first, ...
  std::vector<uint8_t> vtype;
  std::vector<uint32_t> vindex;
  std::vector<uint32_t> row_of_voice;   // position in the type-sorted row space
  VoiceGroup<gd::KickV> kicks;
  VoiceGroup<gd::SnareV> snares;
  VoiceGroup<gd::HatV> hats;
  VoiceGroup<gd::TomV> toms;
  DevBuf<float> d_sorted;               // [n][stride] type-sorted rows when the caller's order is mixed
  DevBuf<uint32_t> d_row_map;
  bool identity_rows = true;
  float* pinned = nullptr;
  size_t pinned_floats = 0;
  ~GooeyVoiceBatch() {
    if (pinned) cudaFreeHost(pinned);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace gh {

// out[v][f] = sorted[row_of_voice[v]][f]  — only needed when patches of different types interleave.
__global__ void unsort_rows_kernel(const float* __restrict__ sorted, float* __restrict__ out, const uint32_t* __restrict__ row_of_voice,
                                   uint32_t n, uint32_t frames, size_t sstride, size_t ostride) {
  const uint32_t v = blockIdx.y;
  const size_t src = (size_t)row_of_voice[v] * sstride;
  const size_t dst = (size_t)v * ostride;
  for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < frames; f += gridDim.x * blockDim.x) out[dst + f] = sorted[src + f];
}

static void snare_cfg_from_patch(const float* p, float* cfg, uint32_t& filter_type) {
  // new_full order -> S_* order (snare.rs:135-180 / SnareParams::from_config :420-545)
  // p: 0 freq,1 tonal,2 noise,3 crack,4 decay,5 pitch_drop,6 volume,7 tonal_decay,8 tonal_decay_curve,9 noise_decay,
  //    10 noise_tail_decay,11 filter_cutoff,12 filter_res,13 filter_type,14 xfade,15 phase_mod,16 overdrive,17 amp_decay,18 amp_decay_curve
  cfg[gd::S_FREQ] = p[0]; cfg[gd::S_TONAL] = p[1]; cfg[gd::S_NOISE] = p[2]; cfg[gd::S_BRIGHTNESS] = p[3]; cfg[gd::S_DECAY] = p[4];
  cfg[gd::S_PITCH_DROP] = p[5]; cfg[gd::S_VOLUME] = p[6]; cfg[gd::S_TONAL_DECAY] = p[7]; cfg[gd::S_TONAL_DECAY_CURVE] = p[8];
  cfg[gd::S_NOISE_DECAY] = p[9]; cfg[gd::S_NOISE_TAIL_DECAY] = p[10]; cfg[gd::S_FILTER_CUTOFF] = p[11]; cfg[gd::S_FILTER_RES] = p[12];
  float ft = p[13];
  int t = !(ft == ft) ? 0 : (ft <= 0.0f ? 0 : (ft >= 255.0f ? 255 : (int)ft));
  filter_type = (uint32_t)(t > 3 ? 3 : t);
  cfg[gd::S_XFADE] = p[14]; cfg[gd::S_PHASE_MOD] = p[15]; cfg[gd::S_OVERDRIVE] = p[16]; cfg[gd::S_AMP_DECAY] = p[17]; cfg[gd::S_AMP_DECAY_CURVE] = p[18];
}

static void voice_batch_render_impl(GooeyVoiceBatch* b, uint32_t frames, float* out_dev, size_t stride) {
  use_device(b->device);
  cudaStream_t st = b->stream;
  b->kicks.ensure_uploaded(st); b->snares.ensure_uploaded(st); b->hats.ensure_uploaded(st); b->toms.ensure_uploaded(st);
  b->kicks.stage_events(st); b->snares.stage_events(st); b->hats.stage_events(st); b->toms.stage_events(st);
  GH_CUDA(cudaEventRecord(b->ev0, st));
  const bool rows = !b->identity_rows;
  int row = 0;
  b->kicks.launch(st, b->ev0, b->rc, 0, (int)frames, out_dev, (long long)stride, gd::OUT_VOICE_MAJOR, row, rows); row += b->kicks.n;
  b->snares.launch(st, b->ev0, b->rc, 0, (int)frames, out_dev, (long long)stride, gd::OUT_VOICE_MAJOR, row, rows); row += b->snares.n;
  b->hats.launch(st, b->ev0, b->rc, 0, (int)frames, out_dev, (long long)stride, gd::OUT_VOICE_MAJOR, row, rows); row += b->hats.n;
  b->toms.launch(st, b->ev0, b->rc, 0, (int)frames, out_dev, (long long)stride, gd::OUT_VOICE_MAJOR, row, rows); row += b->toms.n;
  GH_CUDA(cudaEventRecord(b->ev1, st));
}

}  // namespace gh

#define GOOEY_TRY try {
#define GOOEY_CATCH                                                              \
  } catch (const gh::CudaError& e) { gh::set_error(e.what());                    \
    return std::string(e.what()).find("no CUDA device") != std::string::npos ? GOOEY_E_NO_DEVICE : GOOEY_E_CUDA; \
  } catch (const std::exception& e) { gh::set_error(e.what()); return GOOEY_E_INVALID; }

extern "C" {

const char* gooey_b200_last_error(void) { return gh::last_error().c_str(); }
int gooey_b200_device_count(void) { return gh::device_count(); }
uint64_t gooey_b200_launch_count(void) { return gh::g_launches.load(); }
float gooey_b200_last_kernel_ms(void) { return gh::g_last_kernel_ms; }

int gooey_voice_batch_new(float sample_rate, uint32_t n_voices, const GooeyVoicePatch* patches, int device, GooeyVoiceBatch** out_batch) {
  GOOEY_TRY
  if (!out_batch || (!patches && n_voices)) { set_error("null argument"); return GOOEY_E_INVALID; }
  *out_batch = nullptr;
  if (!(sample_rate > 0.0f)) { set_error("sample_rate must be > 0"); return GOOEY_E_INVALID; }
  use_device(device);
  std::unique_ptr<GooeyVoiceBatch> b(new GooeyVoiceBatch);
  b->device = device; b->sr = sample_rate; b->n = n_voices;
  b->rc = gd::make_rate_ctx(sample_rate);
  GH_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  GH_CUDA(cudaEventCreate(&b->ev0));
  GH_CUDA(cudaEventCreate(&b->ev1));
  b->vtype.resize(n_voices); b->vindex.resize(n_voices); b->row_of_voice.resize(n_voices);
  for (uint32_t v = 0; v < n_voices; v++) {
    const GooeyVoicePatch& p = patches[v];
    switch (p.instrument) {
      case GOOEY_INSTRUMENT_KICK: {
        gd::KickState s; memset(&s, 0, sizeof s);
        gd::kick_init(s, p.params, sample_rate);
        if (p.aux & 0x100) s.cur[gd::K_TUNING] = s.tgt[gd::K_TUNING] = clamp01(p.params[23]);
        b->vindex[v] = b->kicks.add(s, (int)v);
      } break;
      case GOOEY_INSTRUMENT_SNARE: {
        gd::SnareState s; memset(&s, 0, sizeof s);
        float cfg[18]; uint32_t ft;
        snare_cfg_from_patch(p.params, cfg, ft);
        gd::snare_init(s, cfg, ft, sample_rate);
        if (p.aux & 0x100) s.cur[gd::S_TUNING] = s.tgt[gd::S_TUNING] = clamp01(p.params[23]);
        b->vindex[v] = b->snares.add(s, (int)v);
      } break;
      case GOOEY_INSTRUMENT_HIHAT: {
        gd::HatState s; memset(&s, 0, sizeof s);
        gd::hat_init(s, p.params, p.aux & 1, (p.aux & 2) ? 0 : 1, sample_rate);
        if (p.aux & 0x100) s.cur[gd::H_TUNING] = s.tgt[gd::H_TUNING] = clamp01(p.params[23]);
        b->vindex[v] = b->hats.add(s, (int)v);
      } break;
      case GOOEY_INSTRUMENT_TOM: {
        gd::TomState s; memset(&s, 0, sizeof s);
        gd::tom_init(s, (p.aux & 1) ? p.params : nullptr, sample_rate);
        if (p.aux & 0x100) s.p[gd::T_TUNING] = clamp01(p.params[23]);
        b->vindex[v] = b->toms.add(s, (int)v);
      } break;
      default: set_error("unsupported instrument id in voice patch"); return GOOEY_E_INVALID;
    }
    b->vtype[v] = (uint8_t)p.instrument;
  }
  // type-sorted row space: kicks, snares, hats, toms
  uint32_t base[4] = {0, (uint32_t)b->kicks.n, (uint32_t)(b->kicks.n + b->snares.n), (uint32_t)(b->kicks.n + b->snares.n + b->hats.n)};
  b->identity_rows = true;
  for (uint32_t v = 0; v < n_voices; v++) {
    b->row_of_voice[v] = base[b->vtype[v]] + b->vindex[v];
    if (b->row_of_voice[v] != v) b->identity_rows = false;
  }
  if (!b->identity_rows) {
    b->d_row_map.upload(b->row_of_voice.data(), n_voices, b->stream);
    GH_CUDA(cudaStreamSynchronize(b->stream));
  }
  *out_batch = b.release();
  return GOOEY_E_OK;
  GOOEY_CATCH
}

void gooey_voice_batch_free(GooeyVoiceBatch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  delete b;
}

static void vb_add_event(GooeyVoiceBatch* b, uint32_t v, uint32_t frame, uint32_t kind, uint32_t param, float value) {
  const uint32_t i = b->vindex[v];
  switch (b->vtype[v]) {
    case GOOEY_INSTRUMENT_KICK: b->kicks.events.add(i, frame, kind, param, value); break;
    case GOOEY_INSTRUMENT_SNARE: b->snares.events.add(i, frame, kind, param, value); break;
    case GOOEY_INSTRUMENT_HIHAT: b->hats.events.add(i, frame, kind, param, value); break;
    case GOOEY_INSTRUMENT_TOM: b->toms.events.add(i, frame, kind, param, value); break;
  }
}

int gooey_voice_batch_trigger(GooeyVoiceBatch* b, uint32_t voice, uint32_t frame, float velocity) {
  GOOEY_TRY
  if (!b || voice >= b->n) { set_error("bad batch/voice"); return GOOEY_E_INVALID; }
  vb_add_event(b, voice, frame, gd::EV_TRIGGER, 0, velocity);
  return GOOEY_E_OK;
  GOOEY_CATCH
}

int gooey_voice_batch_trigger_all(GooeyVoiceBatch* b, uint32_t frame, const float* velocities) {
  GOOEY_TRY
  if (!b) { set_error("null batch"); return GOOEY_E_INVALID; }
  for (uint32_t v = 0; v < b->n; v++) vb_add_event(b, v, frame, gd::EV_TRIGGER, 0, velocities ? velocities[v] : 1.0f);
  return GOOEY_E_OK;
  GOOEY_CATCH
}

int gooey_voice_batch_set_param(GooeyVoiceBatch* b, uint32_t voice, uint32_t frame, uint32_t param, float value, int snap) {
  GOOEY_TRY
  if (!b || voice >= b->n) { set_error("bad batch/voice"); return GOOEY_E_INVALID; }
  bool known = ffi_param_to_events(b->vtype[voice], param, value,
                                   [&](uint32_t kind, uint32_t p, float v) { vb_add_event(b, voice, frame, kind, p, v); });
  if (known && snap) vb_add_event(b, voice, frame, gd::EV_SNAP, 0, 0.0f);
  return GOOEY_E_OK;
  GOOEY_CATCH
}

int gooey_voice_batch_render_device(GooeyVoiceBatch* b, uint32_t frames, float* out_dev, size_t stride) {
  GOOEY_TRY
  if (!b || !out_dev || stride < frames) { set_error("bad arguments"); return GOOEY_E_INVALID; }
  voice_batch_render_impl(b, frames, out_dev, stride);
  GH_CUDA(cudaStreamSynchronize(b->stream));
  GH_CUDA(cudaEventElapsedTime(&gh::g_last_kernel_ms, b->ev0, b->ev1));
  return GOOEY_E_OK;
  GOOEY_CATCH
}

int gooey_voice_batch_render(GooeyVoiceBatch* b, uint32_t frames, float* out_host) {
  GOOEY_TRY
  if (!b || !out_host) { set_error("bad arguments"); return GOOEY_E_INVALID; }
  use_device(b->device);
  const size_t stride = (frames + 3) & ~(size_t)3;
  const size_t total = (size_t)b->n * stride;
  static thread_local DevBuf<float>* scratch = nullptr;
  DevBuf<float> local;
  (void)scratch;
  local.alloc(total);
  voice_batch_render_impl(b, frames, local.p, stride);
  GH_CUDA(cudaMemcpy2DAsync(out_host, (size_t)frames * 4, local.p, stride * 4, (size_t)frames * 4, b->n, cudaMemcpyDeviceToHost, b->stream));
  GH_CUDA(cudaStreamSynchronize(b->stream));
  GH_CUDA(cudaEventElapsedTime(&gh::g_last_kernel_ms, b->ev0, b->ev1));
  return GOOEY_E_OK;
  GOOEY_CATCH
}

}  // extern "C"

// =================================================================================================
// Self-test hooks (tests/test_gmath_gpu.py): evaluate the gm:: routines on the device so the test can
// compare them bit-for-bit with the host libm the reference links.
// =================================================================================================
namespace gh {
__global__ void math_selftest_kernel(int kind, const float* x, const float* y, float* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r;
  switch (kind) {
    case 0: r = gm::g_powf(x[i], y[i]); break;
    case 1: r = gm::g_sinf(x[i]); break;
    case 2: r = gm::g_cosf(x[i]); break;
    case 3: r = gm::g_expf(x[i]); break;
    case 4: r = gd::hash_noise((uint64_t)x[i]); break;
    case 5: r = gd::max_curve(x[i], y[i]); break;
    case 6: r = gm::g_tanhf(x[i]); break;
    case 7: r = gm::g_tanf(x[i]); break;
    case 8: r = gm::g_expm1f(x[i]); break;
    default: r = 0.0f;
  }
  out[i] = r;
}
}  // namespace gh

extern "C" int gooey_b200_selftest_math(int kind, const float* x_host, const float* y_host, float* out_host, uint32_t n, int device) {
  GOOEY_TRY
  use_device(device);
  DevBuf<float> dx, dy, dout;
  dx.upload(x_host, n, 0);
  dy.upload(y_host ? y_host : x_host, n, 0);
  dout.alloc(n);
  gh::math_selftest_kernel<<<(n + 255) / 256, 256>>>(kind, dx.p, dy.p, dout.p, n);
  g_launches.fetch_add(1);
  GH_CUDA(cudaGetLastError());
  GH_CUDA(cudaMemcpy(out_host, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return GOOEY_E_OK;
  GOOEY_CATCH
}
