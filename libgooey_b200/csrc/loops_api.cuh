// loops_api.cuh — extern "C" surface of the sample-playback sources (include/gooey.h, "loop mixer" and "sampler racks"
// sections; reference: src/ffi.rs:6000-6172 and :7150-7535, :8006-8048).  The setters edit the host-side channel / rack
// structs exactly like the reference's setters edit LoopChannel / SamplerRack; the audio work is loops.cuh, driven by
// engines_render (engine.cuh).  Requests for the parts that are not built latch the engine's sticky error.
#pragma once
#include "engine_api.cuh"

namespace gh {
static GooeyEngine::LoopHost* loop_ch(GooeyEngine* e, uint32_t ch) { return (e && ch < (uint32_t)gd::LOOP_CHANNELS) ? &e->loops[ch] : nullptr; }
static const GooeyEngine::LoopHost* loop_ch(const GooeyEngine* e, uint32_t ch) { return (e && ch < (uint32_t)gd::LOOP_CHANNELS) ? &e->loops[ch] : nullptr; }
static GooeyEngine::SamplerHost* rack_of(GooeyEngine* e, uint32_t rack) { return (e && rack < (uint32_t)gd::SAMPLER_RACKS && e->samplers[rack].registered) ? &e->samplers[rack] : nullptr; }
static const GooeyEngine::SamplerHost* rack_of(const GooeyEngine* e, uint32_t rack) { return (e && rack < (uint32_t)gd::SAMPLER_RACKS && e->samplers[rack].registered) ? &e->samplers[rack] : nullptr; }
// host PCM -> a device buffer of the engine's bank; the bank's stream is drained so no in-flight render reads a buffer that is replaced
static std::shared_ptr<DevBuf<float>> upload_pcm(GooeyEngine* e, const std::vector<float>& host) {
  use_device(e->bank->device);
  auto buf = std::make_shared<DevBuf<float>>();
  buf->alloc(host.size());
  GH_CUDA(cudaMemcpy(buf->p, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
  GH_CUDA(cudaStreamSynchronize(e->bank->stream));
  return buf;
}
static void sampler_stop_slot(GooeyEngine::SamplerHost& R, uint32_t slot) {   // sampler.rs:321-327
  for (int v = 0; v < gd::SAMPLER_VOICES; v++) if (R.voices[v].samples && R.voices[v].slot == slot) R.voices_release(v);
}
}  // namespace gh

extern "C" {

// ---- loop mixer --------------------------------------------------------------------------------------------------------
bool gooey_engine_loop_load(GooeyEngine* e, uint32_t channel, const float* samples, uint32_t frames, uint32_t channels, float sample_rate) {   /* :7184-7202 */
  if (!e || !samples || frames == 0 || channels == 0) return false;
  // StereoSampleBuffer::from_interleaved + from_channels (stereo_buffer.rs:22-87): channels 0 / 1 become left / right (mono is
  // duplicated), the rate and every kept sample must be finite
  if (!std::isfinite(sample_rate) || !(sample_rate > 0.0f)) return false;
  std::vector<float> planes((size_t)2 * frames);
  for (uint32_t f = 0; f < frames; f++) {
    const float* fr = samples + (size_t)f * channels;
    const float l = fr[0], r = channels == 1 ? fr[0] : fr[1];
    if (!std::isfinite(l) || !std::isfinite(r)) return false;
    planes[f] = l; planes[(size_t)frames + f] = r;
  }
  GooeyEngine::LoopHost* c = gh::loop_ch(e, channel);
  if (!c) return false;                                                           // Mixer::load: bad channel index
  try {
    auto buf = gh::upload_pcm(e, planes);
    c->buf = buf; c->len = frames; c->buf_sr = sample_rate;
    c->has_source_bpm = false; c->source_bpm = 0.0f;                              // a new buffer carries no tempo tag
    c->cursor = c->window().lo; c->st_valid = false;                              // set_buffer (loop_channel.rs:311-316)
  } catch (const std::exception& ex) { gh::set_error(ex.what()); gh::engine_fail(e, ex.what()); return false; }
  return true;
}
// libgooey_b200 addition: channel `dst_channel` of `dst` plays the buffer already loaded into `src` (same device) without another
// copy — a batch of engines over one loop keeps one resident buffer instead of one each (as gooey_b200_granulator_share_buffer does
// for the granulator source).  Otherwise identical to gooey_engine_loop_load: the playhead goes to the loop start, no tempo tag.
bool gooey_b200_loop_share_buffer(GooeyEngine* dst, uint32_t dst_channel, const GooeyEngine* src, uint32_t src_channel) {
  GooeyEngine::LoopHost* d = gh::loop_ch(dst, dst_channel);
  const GooeyEngine::LoopHost* s = gh::loop_ch(src, src_channel);
  if (!d || !s || !s->buf || s->len == 0 || dst->bank->device != src->bank->device) return false;
  try { gh::use_device(dst->bank->device); GH_CUDA(cudaStreamSynchronize(dst->bank->stream)); } catch (const std::exception& ex) { gh::set_error(ex.what()); return false; }
  d->buf = s->buf; d->len = s->len; d->buf_sr = s->buf_sr;
  d->has_source_bpm = false; d->source_bpm = 0.0f;
  d->cursor = d->window().lo; d->st_valid = false;
  return true;
}
void gooey_engine_loop_set_playing(GooeyEngine* e, uint32_t ch, bool playing) { if (auto* c = gh::loop_ch(e, ch)) c->playing = playing; }
void gooey_engine_loop_set_gain(GooeyEngine* e, uint32_t ch, float gain) { if (auto* c = gh::loop_ch(e, ch)) gd::lsm_set(c->gain, gd::clampf(gain, 0.0f, gd::LOOP_MAX_GAIN), 0.0f, gd::LOOP_MAX_GAIN); }
void gooey_engine_loop_set_mute(GooeyEngine* e, uint32_t ch, bool muted) { if (auto* c = gh::loop_ch(e, ch)) c->muted = muted; }
void gooey_engine_loop_set_solo(GooeyEngine* e, uint32_t ch, bool soloed) { if (auto* c = gh::loop_ch(e, ch)) c->soloed = soloed; }
void gooey_engine_loop_set_start(GooeyEngine* e, uint32_t ch, float n) { if (auto* c = gh::loop_ch(e, ch)) c->loop_start = gd::clampf(n, 0.0f, 1.0f); }
void gooey_engine_loop_set_end(GooeyEngine* e, uint32_t ch, float n) { if (auto* c = gh::loop_ch(e, ch)) c->loop_end = gd::clampf(n, 0.0f, 1.0f); }
void gooey_engine_loop_set_speed(GooeyEngine* e, uint32_t ch, float s) { if (auto* c = gh::loop_ch(e, ch)) c->speed = gd::clampf(s, -gd::LOOP_MAX_SPEED, gd::LOOP_MAX_SPEED); }
void gooey_engine_loop_set_source_bpm(GooeyEngine* e, uint32_t ch, float bpm) {   /* :7331-7344; StereoSampleBuffer::set_source_bpm :178-180 */
  auto* c = gh::loop_ch(e, ch);
  if (!c || !c->buf) return;
  c->has_source_bpm = bpm > 0.0f && std::isfinite(bpm);
  c->source_bpm = c->has_source_bpm ? bpm : 0.0f;
}
float gooey_engine_loop_get_source_bpm(const GooeyEngine* e, uint32_t ch) { const auto* c = gh::loop_ch(e, ch); return (c && c->buf && c->has_source_bpm) ? c->source_bpm : 0.0f; }
void gooey_engine_loop_set_pitch_mode(GooeyEngine* e, uint32_t ch, uint32_t mode) {   /* :7368-7381 */
  auto* c = gh::loop_ch(e, ch);
  if (!c) return;
  const uint32_t m = mode == GOOEY_PITCH_MODE_RESAMPLE ? 1u : (mode == GOOEY_PITCH_MODE_PRESERVE_PITCH ? 2u : 0u);
  if (c->pitch_mode == 2u && m != 2u) c->st_valid = false;                        // leaving PreservePitch drops the stretcher (loop_channel.rs:341-346)
  c->pitch_mode = m;
}
uint32_t gooey_engine_loop_get_pitch_mode(const GooeyEngine* e, uint32_t ch) { const auto* c = gh::loop_ch(e, ch); return c ? c->pitch_mode : 0u; }
void gooey_engine_loop_restart(GooeyEngine* e, uint32_t ch) { auto* c = gh::loop_ch(e, ch); if (c && c->buf) { c->cursor = c->window().lo; c->st_valid = false; } }
void gooey_engine_loop_set_position(GooeyEngine* e, uint32_t ch, float n) {   /* loop_channel.rs:388-397 */
  auto* c = gh::loop_ch(e, ch);
  if (!c || !c->buf) return;
  const double len = (double)c->len;
  c->cursor = gd::window_fold(c->window(), (double)gd::clampf(n, 0.0f, 1.0f) * len);
  c->st_valid = false;
}
float gooey_engine_loop_get_position(const GooeyEngine* e, uint32_t ch) {     /* position_normalized :497-502 */
  const auto* c = gh::loop_ch(e, ch);
  return (c && c->buf && c->len > 1) ? (float)(c->cursor / (double)c->len) : 0.0f;
}
bool gooey_engine_loop_queue_swap(GooeyEngine* e, uint32_t channel, const float* samples, uint32_t frames, uint32_t channels, float sample_rate,
                                  float source_bpm, uint32_t divisions) {   /* :7449-7476; LoopChannel::queue_swap loop_channel.rs:413-418 */
  if (!e || !samples || frames == 0 || channels == 0) return false;
  if (!std::isfinite(sample_rate) || !(sample_rate > 0.0f)) return false;
  std::vector<float> planes((size_t)2 * frames);
  for (uint32_t f = 0; f < frames; f++) {
    const float* fr = samples + (size_t)f * channels;
    const float l = fr[0], r = channels == 1 ? fr[0] : fr[1];
    if (!std::isfinite(l) || !std::isfinite(r)) return false;
    planes[f] = l; planes[(size_t)frames + f] = r;
  }
  GooeyEngine::LoopHost* c = gh::loop_ch(e, channel);
  if (!c) return false;
  try {
    auto buf = gh::upload_pcm(e, planes);
    c->pend_buf = buf; c->pend_len = frames; c->pend_sr = sample_rate;
    c->pend_has_bpm = source_bpm > 0.0f && std::isfinite(source_bpm); c->pend_bpm = c->pend_has_bpm ? source_bpm : 0.0f;   // tagged before it lands
    c->pend_div = divisions > 1u ? divisions : 1u;
    c->has_pending = true;
  } catch (const std::exception& ex) { gh::set_error(ex.what()); gh::engine_fail(e, ex.what()); return false; }
  return true;
}
void gooey_engine_loop_cancel_queued_swap(GooeyEngine* e, uint32_t ch) {      /* :7483-7490 */
  auto* c = gh::loop_ch(e, ch);
  if (!c) return;
  try { gh::use_device(e->bank->device); GH_CUDA(cudaStreamSynchronize(e->bank->stream)); } catch (const std::exception& ex) { gh::set_error(ex.what()); }
  c->has_pending = false; c->pend_buf.reset(); c->pend_len = 0;
}
uint32_t gooey_engine_loop_swaps_completed(const GooeyEngine* e, uint32_t ch) { const auto* c = gh::loop_ch(e, ch); return c ? c->swaps_completed : 0u; }   /* :7500-7508 */
/* not built: it latches the sticky error and reports failure */
int32_t gooey_engine_loop_effect_add(GooeyEngine* e, uint32_t, uint32_t) {
  if (e) gh::engine_fail(e, "libgooey_b200: per-loop-channel effect chains (ffi.rs:7536-7655) are not built; put the effect on the track the loop mixer is routed to");
  return -1;
}

// Mixer::render_channel_to_interleaved (mixer/mod.rs:444-476): `frames` stereo frames of one loop channel from its loop start,
// ignoring mute / solo, after prepare_offline_render (playing, fader snapped, gate open) — the audio of
// gooey_engine_loop_render_to_wav.  With no per-channel effects the preroll only moves the cursor and the stretcher, and both
// are reset by the restart that follows it — unless a queued take lands during it, so then it is rendered (and discarded) too.
bool gooey_engine_loop_render(GooeyEngine* e, uint32_t channel, uint32_t frames, uint32_t preroll, float* out_interleaved) {
  auto* c = gh::loop_ch(e, channel);
  if (!c || !out_interleaved || frames == 0 || !c->buf || c->len == 0) return false;
  gh::EngineBank& B = *e->bank;
  try {
    std::lock_guard<std::recursive_mutex> lk(B.mu);
    gh::use_device(B.device);
    c->playing = true; c->gain.c = c->gain.t; c->active = {1.0f, 1.0f};
    c->cursor = c->window().lo; c->st_valid = false;                              // prepare_offline_render (loop_channel.rs:444-451)
    const uint32_t longest = std::max(frames, c->has_pending ? preroll : 0u);
    const size_t stride = ((size_t)longest + 31) & ~(size_t)31;
    B.d_ext[0].alloc(2 * stride);
    gd::LoopMixer m;
    auto pass = [&](uint32_t nf) {                                                 // nf ticks of the channel on its own into the row pair 0
      memset(&m, 0, sizeof m);
      c->describe(m.ch[0], e->loop_engine_bpm);
      B.attach_stretcher(*c, m.ch[0]);
      m.row = 0;
      B.d_loop_descs.upload(&m, 1, B.stream);
      gd::ext_source_kernel<gd::LoopMixer><<<1, 64, 0, B.stream>>>(B.d_loop_descs.p, 1, B.d_ext[0].p, (long long)stride, (int)nf, gd::ExtTickCtx{e->sr, B.rc.smooth15});
      gh::g_launches.fetch_add(1, std::memory_order_relaxed);
      GH_CUDA(cudaGetLastError());
      GH_CUDA(cudaMemcpyAsync(&m, B.d_loop_descs.p, sizeof m, cudaMemcpyDeviceToHost, B.stream));
      GH_CUDA(cudaStreamSynchronize(B.stream));
      c->absorb(m.ch[0]);
    };
    // the discarded preroll can only matter through a queued take that lands during it (there are no per-channel effects to warm)
    if (c->has_pending && preroll) pass(preroll);
    c->cursor = c->window().lo; c->st_valid = false;                              // restart() after the preroll (mod.rs:465-466)
    pass(frames);
    std::vector<float> planes(2 * stride);
    GH_CUDA(cudaMemcpy(planes.data(), B.d_ext[0].p, planes.size() * 4, cudaMemcpyDeviceToHost));
    for (uint32_t f = 0; f < frames; f++) { out_interleaved[2 * (size_t)f] = planes[f]; out_interleaved[2 * (size_t)f + 1] = planes[stride + f]; }
  } catch (const std::exception& ex) { gh::set_error(ex.what()); gh::engine_fail(e, ex.what()); return false; }
  return true;
}
bool gooey_engine_loop_render_to_wav(GooeyEngine* e, uint32_t channel, uint32_t frame_count, uint32_t preroll_frame_count, const char* utf8_path) {   /* :8006-8048 */
  if (!e || !utf8_path || !utf8_path[0] || frame_count == 0) return false;
  std::vector<float> il((size_t)2 * frame_count);
  if (!gooey_engine_loop_render(e, channel, frame_count, preroll_frame_count, il.data())) return false;
  return gooey_b200_write_wav_f32(utf8_path, il.data(), frame_count, 2, (uint32_t)gd::f32_to_u64_sat(e->sr)) == GOOEY_E_OK;
}

// ---- sampler racks -----------------------------------------------------------------------------------------------------
int32_t gooey_engine_sampler_register(GooeyEngine* e) {   /* :6007-6027 */
  if (!e) return -1;
  for (int r = 0; r < gd::SAMPLER_RACKS; r++)
    if (!e->samplers[r].registered) {               // SamplerRack::new(engine.sample_rate, engine.bpm): a stopped 16-step sequencer, every step off
      e->samplers[r].registered = true;
      e->samplers[r].pat.seq.init(e->bpm, e->sr);
      return r;
    }
  return -1;
}
uint32_t gooey_engine_sampler_get_source_id(const GooeyEngine* e, uint32_t rack) { return gh::rack_of(e, rack) ? GOOEY_SOURCE_SAMPLER_BASE + rack : 0xFFFFFFFFu; }
bool gooey_engine_sampler_set_slot_buffer(GooeyEngine* e, uint32_t rack, uint32_t slot, const float* samples, uint32_t frames, uint32_t channels, float sample_rate) {   /* :6044-6072 */
  if (!e || !samples) return false;
  // SamplerBuffer::from_interleaved (sampler.rs:25-50)
  if (!(channels == 1 || channels == 2) || frames == 0 || !std::isfinite(sample_rate) || !(sample_rate > 0.0f)) return false;
  const size_t count = (size_t)frames * channels;
  for (size_t i = 0; i < count; i++) if (!std::isfinite(samples[i])) return false;
  auto* R = gh::rack_of(e, rack);
  if (!R || slot >= (uint32_t)gd::SAMPLER_SLOTS) return false;
  try {
    auto buf = gh::upload_pcm(e, std::vector<float>(samples, samples + count));
    auto& S = R->slots[slot];
    S.buf = buf; S.frames = frames; S.channels = channels; S.sr = sample_rate;
  } catch (const std::exception& ex) { gh::set_error(ex.what()); gh::engine_fail(e, ex.what()); return false; }
  gh::sampler_stop_slot(*R, slot);
  return true;
}
bool gooey_engine_sampler_clear_slot(GooeyEngine* e, uint32_t rack, uint32_t slot) {
  auto* R = gh::rack_of(e, rack);
  if (!R || slot >= (uint32_t)gd::SAMPLER_SLOTS) return false;
  try { gh::use_device(e->bank->device); GH_CUDA(cudaStreamSynchronize(e->bank->stream)); } catch (const std::exception& ex) { gh::set_error(ex.what()); return false; }
  R->slots[slot] = GooeyEngine::SamplerHost::Slot();
  gh::sampler_stop_slot(*R, slot);
  return true;
}
bool gooey_engine_sampler_slot_is_loaded(const GooeyEngine* e, uint32_t rack, uint32_t slot) { const auto* R = gh::rack_of(e, rack); return R && slot < (uint32_t)gd::SAMPLER_SLOTS && R->slots[slot].buf; }
uint32_t gooey_engine_sampler_slot_frames(const GooeyEngine* e, uint32_t rack, uint32_t slot) { return gooey_engine_sampler_slot_is_loaded(e, rack, slot) ? e->samplers[rack].slots[slot].frames : 0u; }
uint32_t gooey_engine_sampler_slot_channels(const GooeyEngine* e, uint32_t rack, uint32_t slot) { return gooey_engine_sampler_slot_is_loaded(e, rack, slot) ? e->samplers[rack].slots[slot].channels : 0u; }
float gooey_engine_sampler_slot_sample_rate(const GooeyEngine* e, uint32_t rack, uint32_t slot) { return gooey_engine_sampler_slot_is_loaded(e, rack, slot) ? e->samplers[rack].slots[slot].sr : 0.0f; }
bool gooey_engine_sampler_trigger(GooeyEngine* e, uint32_t rack, uint32_t slot, float velocity) {   /* :6150-6168; SamplerRack::trigger sampler.rs:200-223 */
  auto* R = gh::rack_of(e, rack);
  if (!R || slot >= (uint32_t)gd::SAMPLER_SLOTS || !R->slots[slot].buf) return false;
  int vi = -1;
  for (int v = 0; v < gd::SAMPLER_VOICES; v++) if (!R->voices[v].samples) { vi = v; break; }
  if (vi < 0) { vi = 0; for (int v = 1; v < gd::SAMPLER_VOICES; v++) if (R->voices[v].age < R->voices[vi].age) vi = v; }   // oldest voice, first on ties
  R->next_age += 1;
  const auto& S = R->slots[slot];
  gd::SampleVoice& V = R->voices[vi];
  V.samples = S.buf->p; V.frames = S.frames; V.channels = S.channels; V.slot = slot;
  V.position = 0.0; V.increment = (double)S.sr / (double)e->sr;
  V.velocity = gd::clampf(velocity, 0.0f, 1.0f); V.age = R->next_age;
  R->voice_buf[vi] = S.buf;
  return true;
}
// ---- the rack's step pattern (:6173-6290; sampler.rs:232-310): 16 steps of (enabled, pad, velocity) on the engine's 16th-note grid; a start
// is armed on the transport beat (quantised to the next sixteenth / quarter / bar while the transport runs, beat 0 while it is stopped)
// and fires inside the render at the first frame whose beat has reached it ----
bool gooey_engine_sampler_set_step(GooeyEngine* e, uint32_t rack, uint32_t step, bool enabled, uint32_t slot, float velocity) {
  auto* R = gh::rack_of(e, rack);
  if (!R || step >= (uint32_t)gd::SAMPLER_SLOTS || slot >= (uint32_t)gd::SAMPLER_SLOTS) return false;
  gh::SeqStep& s = R->pat.seq.pattern[step];
  s.enabled = enabled; s.velocity = gd::clampf(velocity, 0.0f, 1.0f); s.has_note = true; s.note = (uint8_t)slot;
  return true;
}
bool gooey_engine_sampler_get_step(const GooeyEngine* e, uint32_t rack, uint32_t step, bool* out_enabled, uint32_t* out_slot, float* out_velocity) {
  if (!out_enabled || !out_slot || !out_velocity) return false;
  const auto* R = gh::rack_of(e, rack);
  if (!R || step >= (uint32_t)gd::SAMPLER_SLOTS) return false;
  const gh::SeqStep& s = R->pat.seq.pattern[step];
  *out_enabled = s.enabled; *out_slot = s.has_note ? s.note : 0u; *out_velocity = s.velocity;
  return true;
}
bool gooey_engine_sampler_start_pattern(GooeyEngine* e, uint32_t rack, uint32_t quantization) {   /* :6192-6207 */
  if (!e || quantization > GOOEY_CLIP_QUANTIZE_BAR) return false;
  const double interval = quantization == GOOEY_CLIP_QUANTIZE_SIXTEENTH ? 0.25 : (quantization == GOOEY_CLIP_QUANTIZE_QUARTER ? 1.0 : 4.0);
  const double target = e->transport.quantized_target(interval);
  auto* R = gh::rack_of(e, rack);
  if (!R || !std::isfinite(target) || target < 0.0) return false;
  R->pat.pattern_running = false; R->pat.seq.stop(); R->pat.has_pending = true; R->pat.pending_beat = target;   // SamplerRack::schedule_start (sampler.rs:254-262)
  return true;
}
bool gooey_engine_sampler_stop_pattern(GooeyEngine* e, uint32_t rack) {      /* :6211-6221; stops the rack's voices too */
  auto* R = gh::rack_of(e, rack);
  if (!R) return false;
  R->pat.has_pending = false; R->pat.pattern_running = false; R->pat.seq.stop(); R->stop_all();
  return true;
}
bool gooey_engine_sampler_cancel_pattern_start(GooeyEngine* e, uint32_t rack) { auto* R = gh::rack_of(e, rack); if (!R) return false; R->pat.has_pending = false; return true; }
double gooey_engine_sampler_get_pending_start_beat(const GooeyEngine* e, uint32_t rack) { const auto* R = gh::rack_of(e, rack); return (R && R->pat.has_pending) ? R->pat.pending_beat : -1.0; }
double gooey_engine_transport_get_beat_position(const GooeyEngine* e) { return e ? const_cast<GooeyEngine*>(e)->transport.now() : 0.0; }   /* :7143-7150 */
bool gooey_engine_sampler_is_pattern_running(const GooeyEngine* e, uint32_t rack) { const auto* R = gh::rack_of(e, rack); return R && R->pat.pattern_running; }

}  // extern "C"
