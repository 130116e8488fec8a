/* gooey.h — the subset of libgooey's C FFI that drives the offline bounce path, as exported by libgooey_b200.so.
 *
 * The reference generates its header with cbindgen at build time (build.rs:1-28) and does not check it in; this file
 * restates, by hand, the entry points of src/ffi.rs that the bounce path needs, with the same
 * names, argument meaning and error behaviour: setters ignore a null engine / bad index / unknown id, getters return
 * sentinels, nothing unwinds.  Every function cites the reference definition it replaces (src/ffi.rs:LINE).
 * Rendering happens on an NVIDIA B200; there is no CPU fallback: without a CUDA device gooey_engine_new returns NULL
 * and gooey_b200_last_error() says why.
 *
 * Out of scope in this build: UI getters, samplers, loop mixer, clip grid, MIDI / Link.  A call that asks for
 * something this build cannot render (more effect instances than it has slots) latches the sticky error and fires the error
 * callback instead of rendering different audio silently.
 */
#ifndef GOOEY_H
#define GOOEY_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include "gooey_batch.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct GooeyEngine GooeyEngine;

/* ---- ABI constants (src/ffi.rs:1547-2007) ---- */
#define GOOEY_OUTPUT_CHANNELS 2u              /* :2047 interleaved [L, R] */
#define GOOEY_INSTRUMENT_COUNT 5u             /* :1853 */
#define GOOEY_EFFECT_LOWPASS_FILTER 0u        /* :1548-1593 */
#define GOOEY_EFFECT_DELAY 1u
#define GOOEY_EFFECT_SATURATION 2u
#define GOOEY_EFFECT_COMPRESSOR 3u
#define GOOEY_EFFECT_TILT_FILTER 4u
#define GOOEY_EFFECT_LIMITER 5u
#define GOOEY_EFFECT_REVERB 6u
#define GOOEY_EFFECT_WAVESHAPER 7u
#define GOOEY_EFFECT_FEEDBACK_WAVESHAPER 8u
#define GOOEY_EFFECT_PLATE_REVERB 9u
#define GOOEY_DELAY_PARAM_TIMING 0u           /* :1600-1730 */
#define GOOEY_DELAY_PARAM_FEEDBACK 1u
#define GOOEY_DELAY_PARAM_MIX 2u
#define GOOEY_DELAY_PARAM_FILTER_CUTOFF 3u
#define GOOEY_DELAY_PARAM_PINGPONG 4u
#define GOOEY_TILT_PARAM_CUTOFF 0u
#define GOOEY_TILT_PARAM_RESONANCE 1u
#define GOOEY_REVERB_PARAM_DECAY 0u
#define GOOEY_REVERB_PARAM_MIX 1u
#define GOOEY_REVERB_PARAM_DAMPING 2u
#define GOOEY_PLATE_PARAM_DECAY 0u
#define GOOEY_PLATE_PARAM_MIX 1u
#define GOOEY_PLATE_PARAM_DAMPING 2u
#define GOOEY_PLATE_PARAM_PREDELAY 3u
#define GOOEY_PLATE_PARAM_WIDTH 4u
#define GOOEY_PLATE_PARAM_SIZE 5u
#define GOOEY_LIMITER_PARAM_THRESHOLD 0u
#define GOOEY_FILTER_PARAM_CUTOFF 0u          /* low-pass: Hz 20-20000, resonance 0-0.95 */
#define GOOEY_FILTER_PARAM_RESONANCE 1u
#define GOOEY_SATURATION_PARAM_DRIVE 0u
#define GOOEY_SATURATION_PARAM_WARMTH 1u
#define GOOEY_SATURATION_PARAM_MIX 2u
#define GOOEY_COMPRESSOR_PARAM_THRESHOLD 0u   /* dB -60..0, ratio 1..20, attack ms 0.1..100, release ms 5..1000, mix 0..1 */
#define GOOEY_COMPRESSOR_PARAM_RATIO 1u
#define GOOEY_COMPRESSOR_PARAM_ATTACK 2u
#define GOOEY_COMPRESSOR_PARAM_RELEASE 3u
#define GOOEY_COMPRESSOR_PARAM_MIX 4u
#define GOOEY_COMPRESSOR_SIDECHAIN_NONE 0xFFFFFFFFu
#define GOOEY_WAVESHAPER_PARAM_DRIVE 0u       /* 1-10 */
#define GOOEY_WAVESHAPER_PARAM_MIX 1u
#define GOOEY_FEEDBACK_WAVESHAPER_PARAM_DRIVE 0u          /* 1-100 */
#define GOOEY_FEEDBACK_WAVESHAPER_PARAM_FEEDBACK 1u       /* 0-0.98 */
#define GOOEY_FEEDBACK_WAVESHAPER_PARAM_FILTER_CUTOFF 2u  /* Hz 200-20000 */
#define GOOEY_FEEDBACK_WAVESHAPER_PARAM_MIX 3u
#define GOOEY_SCALE_MAJOR 0u                  /* :5502-5515 */
#define GOOEY_SCALE_MINOR 1u
#define GOOEY_VOICING_ROOT_POSITION 0u
#define GOOEY_VOICING_FIRST_INVERSION 1u
#define GOOEY_VOICING_SECOND_INVERSION 2u
#define GOOEY_VOICING_THIRD_INVERSION 3u
#define GOOEY_VOICING_OPEN 4u
#define GOOEY_VOICING_DROP2 5u
#define GOOEY_VOICING_DROP3 6u
#define GOOEY_VOICING_SPREAD 7u
#define GOOEY_VOICING_SHELL 8u
#define GOOEY_VOICING_ROOTLESS 9u
#define GOOEY_SOURCE_DRUMKIT 0u               /* src/mixer/graph.rs:27-42 */
#define GOOEY_SOURCE_BASS 1u
#define GOOEY_SOURCE_POLYSYNTH 2u
#define GOOEY_SOURCE_GRANULATOR 3u
#define GOOEY_SOURCE_LOOPMIXER 4u
#define GOOEY_SOURCE_SAMPLER_BASE 5u         /* + rack id; routable once the rack is registered (ffi.rs:1872-1873) */
#define GOOEY_STEP_NOTE_NONE 255u             /* :1980 */
#define GOOEY_BASS_PRESET_ACID 0u             /* :1882-1998 */
#define GOOEY_BASS_PRESET_SUB 1u
#define GOOEY_BASS_PRESET_REESE 2u
#define GOOEY_BASS_PRESET_STAB 3u

/* ---- lifetime (:2024-2039) ---- */
GooeyEngine* gooey_engine_new(float sample_rate);
void gooey_engine_free(GooeyEngine* engine);
/* libgooey_b200 addition: CUDA device used by engines created afterwards on this thread (default 0). */
int gooey_b200_set_device(int device);

/* ---- render / bounce (:2067-2122, :7897-7930) ---- */
void gooey_engine_render(GooeyEngine* engine, float* buffer, uint32_t frames);
float* gooey_engine_bounce_to_buffer(GooeyEngine* engine, uint32_t bars, uint32_t* out_length);
void gooey_engine_free_buffer(float* buffer, uint32_t length);
/* :7942-7980 — mono 16-bit PCM WAV of the bounce; false on null arguments or I/O failure. */
bool gooey_engine_bounce_to_wav(GooeyEngine* engine, uint32_t bars, const char* utf8_path);
/* Batch of the two calls above over n engines (libgooey_b200 addition; SURVEY.md section 8b "what calls it"):
 * every engine is bounced `bars` bars in ONE device pass.  out_buffers[i] receives a callee-allocated mono buffer
 * (free with gooey_engine_free_buffer), out_lengths[i] its length.  Returns 0 or a GOOEY_E_* code. */
int gooey_batch_bounce(GooeyEngine* const* engines, uint32_t n, uint32_t bars, float** out_buffers, uint32_t* out_lengths);
/* Same, result left in device memory: out_dev[i * stride + frame]; all engines must share bpm (equal length).
 * *out_frames receives the length. */
int gooey_batch_bounce_device(GooeyEngine* const* engines, uint32_t n, uint32_t bars, float* out_dev, size_t stride, uint32_t* out_frames);
/* Mono bounce of n engines of equal length into ONE pitched host block: out_host[i * pitch + frame] (f32) or the 16-bit PCM
 * of gooey_engine_bounce_to_wav (`(s * 32767).round() as i16`, ffi.rs:7968-7972, quantised on the device).  Finished pieces
 * are drained while later ones render; engines may live on several devices (gooey_b200_set_device before gooey_engine_new):
 * each device's share is rendered and drained by its own host thread, no collective. */
int gooey_batch_bounce_host(GooeyEngine* const* engines, uint32_t n, uint32_t bars, float* out_host, size_t pitch, uint32_t* out_frames);
int gooey_batch_bounce_pcm16(GooeyEngine* const* engines, uint32_t n, uint32_t bars, int16_t* out_host, size_t pitch, uint32_t* out_frames);
/* gooey_engine_bounce_to_wav for n engines in one device pass (paths[i] = UTF-8 path of engine i's mono 16-bit WAV). */
int gooey_batch_bounce_to_wav(GooeyEngine* const* engines, uint32_t n, uint32_t bars, const char* const* utf8_paths);
/* Batch of gooey_engine_render: interleaved stereo, out_host[i * 2 * frames + 2 * f + ch]. */
int gooey_batch_render(GooeyEngine* const* engines, uint32_t n, uint32_t frames, float* out_host);

/* ---- LFO pool (:4616-4993; engine/lfo.rs:170-185): eight sine LFOs synced to the tempo (timing 0..7 = 4 bars, 2 bars, 1 bar, 1/2,
 * 1/4, 1/8, 1/16, 1/32), value = offset + sin(2 pi phase) * amount.  Every frame, after the sequencer's triggers and before the
 * voices tick, each enabled LFO writes `value * depth` (bipolar, -1..1 onto the parameter's range) to its routed channel parameters
 * (:1238-1251, :322-405).  Routed voices are rendered on the per-sample path. */
uint32_t gooey_engine_lfo_count(void);
uint32_t gooey_engine_lfo_timing_count(void);
void gooey_engine_set_lfo_enabled(GooeyEngine* engine, uint32_t lfo_index, bool enabled);
bool gooey_engine_get_lfo_enabled(const GooeyEngine* engine, uint32_t lfo_index);
void gooey_engine_set_lfo_timing(GooeyEngine* engine, uint32_t lfo_index, uint32_t timing);
uint32_t gooey_engine_get_lfo_timing(const GooeyEngine* engine, uint32_t lfo_index);     /* 0xFFFFFFFF for a bad argument */
void gooey_engine_set_lfo_amount(GooeyEngine* engine, uint32_t lfo_index, float amount);
float gooey_engine_get_lfo_amount(const GooeyEngine* engine, uint32_t lfo_index);
void gooey_engine_set_lfo_offset(GooeyEngine* engine, uint32_t lfo_index, float offset);
float gooey_engine_get_lfo_offset(const GooeyEngine* engine, uint32_t lfo_index);
uint32_t gooey_engine_add_lfo_route(GooeyEngine* engine, uint32_t lfo_index, uint32_t instrument, uint32_t param, float depth);   /* route id, 0xFFFFFFFF when full (16) */
bool gooey_engine_remove_lfo_route(GooeyEngine* engine, uint32_t lfo_index, uint32_t route_id);
void gooey_engine_clear_lfo_routes(GooeyEngine* engine, uint32_t lfo_index);
uint32_t gooey_engine_get_lfo_route_count(const GooeyEngine* engine, uint32_t lfo_index);
void gooey_engine_reset_lfo_phase(GooeyEngine* engine, uint32_t lfo_index);
float gooey_engine_get_lfo_phase(const GooeyEngine* engine, uint32_t lfo_index);          /* -1 for a bad argument */

/* ---- preset blend: X/Y pad over four corner presets (:5245-5490; utils/blendable.rs:73-86) and per-step blends (:4009-4075) ----
 * A blend is `<Voice>::set_config(bilinear(corners, x, y))`.  set_position applies it at once (only while enabled); at every
 * sequencer trigger the step's own blend — or, when blending is enabled, the pad position — is applied and followed by
 * snap_params (:1162-1171, :1384-1402).  Corner presets: GOOEY_*_PRESET ids 0..3 of the channel's instrument type. */
void gooey_engine_blend_enable(GooeyEngine* engine, uint32_t instrument);
void gooey_engine_blend_disable(GooeyEngine* engine, uint32_t instrument);
bool gooey_engine_blend_is_enabled(const GooeyEngine* engine, uint32_t instrument);
void gooey_engine_blend_set_position(GooeyEngine* engine, uint32_t instrument, float x, float y);
float gooey_engine_blend_get_position_x(const GooeyEngine* engine, uint32_t instrument);   /* -1 for a bad argument */
float gooey_engine_blend_get_position_y(const GooeyEngine* engine, uint32_t instrument);
void gooey_engine_blend_set_corner_preset(GooeyEngine* engine, uint32_t instrument, uint32_t corner, uint32_t preset_id);
uint32_t gooey_engine_blend_get_corner_preset(const GooeyEngine* engine, uint32_t instrument, uint32_t corner);   /* 0xFFFFFFFF for a bad argument */
void gooey_engine_blend_reset_corners(GooeyEngine* engine, uint32_t instrument);
void gooey_engine_sequencer_set_instrument_step_blend(GooeyEngine* engine, uint32_t instrument, uint32_t step, float x, float y);
void gooey_engine_sequencer_set_instrument_step_blend_override(GooeyEngine* engine, uint32_t instrument, uint32_t step, float x, float y);   /* legacy alias */
void gooey_engine_sequencer_clear_instrument_step_blend(GooeyEngine* engine, uint32_t instrument, uint32_t step);
void gooey_engine_sequencer_clear_instrument_step_blend_override(GooeyEngine* engine, uint32_t instrument, uint32_t step);
float gooey_engine_sequencer_get_instrument_step_blend_x(const GooeyEngine* engine, uint32_t instrument, uint32_t step);   /* -1: no blend on the step */
float gooey_engine_sequencer_get_instrument_step_blend_y(const GooeyEngine* engine, uint32_t instrument, uint32_t step);

/* ---- meters and MIDI export ---- */
/* :2572-2584 — per-channel peak |x| (pre-pan, post gain x mute) since the last call, reset to 0 by the read; count <= 5. */
void gooey_engine_get_channel_peaks(GooeyEngine* engine, float* out_peaks, uint32_t count);
/* :6573-6580 — a mixer track's post-strip peak max(|l|, |r|) since the last call (read and reset); 0 for a bad track. */
float gooey_engine_mixer_get_track_peak(GooeyEngine* engine, uint32_t track);
/* :78-83, :2145-2167 — note-on events of the most recent render call (manual triggers at offset 0, sequencer triggers at their
 * frame), at most 64; drained events are removed.  After a bounce: the events of its last 512-frame chunk (:7855-7870). */
typedef struct GooeyMidiEvent { uint32_t instrument_index; float velocity; uint32_t sample_offset; } GooeyMidiEvent;
uint32_t gooey_engine_drain_midi_events(GooeyEngine* engine, GooeyMidiEvent* out_events, uint32_t max_events);

/* ---- errors (:2236-2284) ---- */
bool gooey_engine_has_error(const GooeyEngine* engine);
const char* gooey_engine_get_error_message(const GooeyEngine* engine);
void gooey_engine_set_error_callback(GooeyEngine* engine, void* context, void (*callback)(void*, const char*));

/* ---- instrument parameters, normalized 0-1 (:2623, :2767, :2689, :2835, :2907, channel :166-250) ---- */
void gooey_engine_set_kick_param(GooeyEngine* engine, uint32_t param, float value);
void gooey_engine_set_snare_param(GooeyEngine* engine, uint32_t param, float value);
void gooey_engine_set_hihat_param(GooeyEngine* engine, uint32_t param, float value);
void gooey_engine_set_tom_param(GooeyEngine* engine, uint32_t param, float value);
void gooey_engine_set_bass_param(GooeyEngine* engine, uint32_t param, float value);
void gooey_engine_set_channel_param(GooeyEngine* engine, uint32_t channel, uint32_t param, float value);
void gooey_engine_load_bass_preset(GooeyEngine* engine, uint32_t preset);                        /* :2933 */
/* :2304-2343 — another synthesizer type on a channel (0-3 kit, 4 bass); the new instrument is `<Voice>::new(sample_rate)`. */
void gooey_engine_set_channel_instrument_type(GooeyEngine* engine, uint32_t channel, uint32_t instrument_type);
uint32_t gooey_engine_get_channel_instrument_type(const GooeyEngine* engine, uint32_t channel);  /* :2360-2372 */

/* ---- transport (:3337-3364, :3469, :3300) ---- */
void gooey_engine_set_bpm(GooeyEngine* engine, float bpm);
float gooey_engine_get_bpm(const GooeyEngine* engine);
void gooey_engine_set_swing(GooeyEngine* engine, float swing);
void gooey_engine_set_master_gain(GooeyEngine* engine, float gain);

/* ---- sequencer (:3695-3708, :3864-3995, :4089, :4191, start/stop/reset) ---- */
void gooey_engine_sequencer_set_step(GooeyEngine* engine, uint32_t step, bool enabled);        /* kick only */
void gooey_engine_sequencer_set_instrument_step(GooeyEngine* engine, uint32_t instrument, uint32_t step, bool enabled);
void gooey_engine_sequencer_set_instrument_step_with_velocity(GooeyEngine* engine, uint32_t instrument, uint32_t step, bool enabled, float velocity);
void gooey_engine_sequencer_set_instrument_step_settings(GooeyEngine* engine, uint32_t instrument, uint32_t step, bool enabled,
                                                         bool set_velocity, float velocity, bool set_blend, float blend_x, float blend_y,
                                                         bool set_note, uint8_t midi_note);
void gooey_engine_sequencer_set_instrument_step_note(GooeyEngine* engine, uint32_t instrument, uint32_t step, uint8_t midi_note);
void gooey_engine_sequencer_set_instrument_pattern(GooeyEngine* engine, uint32_t instrument, const bool* pattern16);
void gooey_engine_sequencer_start(GooeyEngine* engine);
void gooey_engine_sequencer_stop(GooeyEngine* engine);
void gooey_engine_sequencer_reset(GooeyEngine* engine);
/* src/ffi.rs:2188-2215: sequencers keep their clock but fire nothing (and export no MIDI event) while disabled; default true */
void gooey_engine_set_sequencer_triggers_enabled(GooeyEngine* engine, bool enabled);
bool gooey_engine_get_sequencer_triggers_enabled(const GooeyEngine* engine);

/* ---- voice strips (:5036-5210) and manual triggers (:2518-2552; latched, fire at frame 0 of the next render) ---- */
void gooey_engine_set_instrument_gain(GooeyEngine* engine, uint32_t instrument, float gain);
void gooey_engine_set_instrument_pan(GooeyEngine* engine, uint32_t instrument, float pan);
void gooey_engine_set_instrument_mute(GooeyEngine* engine, uint32_t instrument, bool muted);
void gooey_engine_set_instrument_solo(GooeyEngine* engine, uint32_t instrument, bool soloed);
void gooey_engine_trigger_instrument(GooeyEngine* engine, uint32_t instrument);
void gooey_engine_trigger_instrument_with_velocity(GooeyEngine* engine, uint32_t instrument, float velocity);

/* ---- global effect chain (:2988-3072, :3176-3237, :3252-3281, :4498-4620): all ten effects; the nine before the limiter
 * can be reordered, which resets their state (:1417-1425) ---- */
void gooey_engine_set_global_effect_param(GooeyEngine* engine, uint32_t effect, uint32_t param, float value);
void gooey_engine_set_global_effect_enabled(GooeyEngine* engine, uint32_t effect, bool enabled);
bool gooey_engine_get_global_effect_enabled(const GooeyEngine* engine, uint32_t effect);
void gooey_engine_set_compressor_sidechain(GooeyEngine* engine, uint32_t instrument);
uint32_t gooey_engine_get_compressor_sidechain(const GooeyEngine* engine);
bool gooey_engine_set_effect_order(GooeyEngine* engine, const uint32_t* ids, uint32_t len);
bool gooey_engine_move_effect(GooeyEngine* engine, uint32_t effect_id, uint32_t new_position);
uint32_t gooey_engine_get_effect_order(const GooeyEngine* engine, uint32_t* out_ids, uint32_t max_len);

/* ---- mixer graph (:6324-6674) ---- */
int32_t gooey_engine_mixer_add_track(GooeyEngine* engine, const char* name);
uint32_t gooey_engine_mixer_get_track_count(const GooeyEngine* engine);
bool gooey_engine_mixer_route_source(GooeyEngine* engine, uint32_t source, uint32_t track);
bool gooey_engine_mixer_unroute_source(GooeyEngine* engine, uint32_t source);                        /* :6427 */
int32_t gooey_engine_mixer_get_source_route(const GooeyEngine* engine, uint32_t source);             /* :6442, -1 = unrouted / inactive source */
void gooey_engine_mixer_reset_default_layout(GooeyEngine* engine);                                   /* :6295: Drums / Bass / Synth / Loops, fresh strips and routes */
void gooey_engine_mixer_clear_layout(GooeyEngine* engine);                                           /* :6313: no tracks, no routes */
float gooey_engine_mixer_get_track_gain(const GooeyEngine* engine, uint32_t track);                  /* :6472 */
float gooey_engine_mixer_get_track_pan(const GooeyEngine* engine, uint32_t track);                   /* :6501 */
bool gooey_engine_mixer_get_track_mute(const GooeyEngine* engine, uint32_t track);                   /* :6530 */
bool gooey_engine_mixer_get_track_solo(const GooeyEngine* engine, uint32_t track);                   /* :6559 */
void gooey_engine_mixer_set_track_gain(GooeyEngine* engine, uint32_t track, float gain);
void gooey_engine_mixer_set_track_pan(GooeyEngine* engine, uint32_t track, float pan);
void gooey_engine_mixer_set_track_mute(GooeyEngine* engine, uint32_t track, bool muted);
void gooey_engine_mixer_set_track_solo(GooeyEngine* engine, uint32_t track, bool soloed);
int32_t gooey_engine_track_effect_add(GooeyEngine* engine, uint32_t track, uint32_t effect_id);
void gooey_engine_track_effect_set_param(GooeyEngine* engine, uint32_t track, uint32_t slot, uint32_t param, float value);
bool gooey_engine_track_effect_remove(GooeyEngine* engine, uint32_t track, uint32_t slot);
bool gooey_engine_track_effect_move(GooeyEngine* engine, uint32_t track, uint32_t slot, uint32_t new_position);
uint32_t gooey_engine_track_effect_count(const GooeyEngine* engine, uint32_t track);
/* This build: at most 4 effects per track rack and 8 rack + {low-pass, saturation, compressor, waveshaper, feedback waveshaper}
 * instances per engine; one more latches the sticky error. */

/* ---- poly synth ("chord oscillators"; :5571-5648, :5899-5935).  Calls act at the engine's current time, like the reference.
 * gooey_engine_poly_trigger_chord = diatonic seventh chord of (root, scale) at `degree` (src/music/key.rs:55-84), voiced
 * (src/music/voicing.rs:76-172), then gooey_engine_poly_trigger_notes (libgooey_b200 addition: the same with explicit notes). ---- */
#define GOOEY_POLY_PRESET_DEFAULT 0u
#define GOOEY_POLY_PRESET_PAD 1u
#define GOOEY_POLY_PRESET_PLUCK 2u
#define GOOEY_POLY_PRESET_KEYS 3u
#define GOOEY_POLY_PRESET_STRINGS 4u
void gooey_engine_poly_trigger_chord(GooeyEngine* engine, uint32_t root, uint32_t scale_type, uint32_t degree, uint32_t voicing, uint32_t preset,
                                     int32_t octave, float velocity);
void gooey_engine_poly_trigger_notes(GooeyEngine* engine, const uint8_t* midi_notes, uint32_t n, uint32_t preset, float velocity);
/* host-only: the notes the call above would trigger; returns their count */
uint32_t gooey_b200_chord_notes(uint32_t root, uint32_t scale_type, uint32_t degree, uint32_t voicing, int32_t octave, uint8_t* out_notes, uint32_t capacity);
void gooey_engine_poly_release(GooeyEngine* engine);
void gooey_engine_poly_set_preset(GooeyEngine* engine, uint32_t preset);
void gooey_engine_poly_set_param(GooeyEngine* engine, uint32_t param, float value);

/* ---- granulator (:5969-5990 set_buffer copies and validates; :7656-7827).  Parameter ids (:1944-1968): 0 scan position,
 * 1 grain length, 2 spray, 3 pitch, 4 density, 5 texture, 6 direction, 7 cloud duration, 8 volume, 9 random timing,
 * 10 random amp, 11 drive. ---- */
#define GOOEY_GRANULATOR_PARAM_COUNT 12u
bool gooey_engine_granulator_set_buffer(GooeyEngine* engine, const float* samples, uint32_t len, float sample_rate);
uint32_t gooey_engine_granulator_buffer_len(const GooeyEngine* engine);
float gooey_engine_granulator_buffer_sample_rate(const GooeyEngine* engine);
void gooey_engine_granulator_trigger(GooeyEngine* engine, float velocity);
void gooey_engine_granulator_set_param(GooeyEngine* engine, uint32_t param, float value);
void gooey_engine_granulator_set_seed(GooeyEngine* engine, uint32_t seed);
void gooey_engine_granulator_snap_params(GooeyEngine* engine);
/* libgooey_b200 addition: dst plays the buffer already loaded into src, without another device copy. */
bool gooey_b200_granulator_share_buffer(GooeyEngine* dst, const GooeyEngine* src);

/* ---- loop mixer: 4 stereo loop channels summed into graph source GOOEY_SOURCE_LOOPMIXER (ffi.rs:7150-7535; src/mixer/mod.rs,
 * loop_channel.rs, stereo_buffer.rs).  The host passes decoded PCM (interleaved f32; 1 channel is duplicated, 2+ use channels 0 / 1);
 * the buffer is copied to the device.  Calls act immediately, like the reference's.  Playback state is not touched by a bounce
 * (ffi.rs:7840-7854 resets sequencers and snaps strips only).
 * Pitch modes: Off, Resample (tempo warp shifts pitch) and PreservePitch (WSOLA time-stretch, src/mixer/wsola.rs; reverse speeds
 * fall back to the direct read, loop_channel.rs:184).  This build: gooey_engine_loop_effect_add latches the sticky error (per-channel
 * effect chains and the clip grid are not built). ---- */
#define GOOEY_LOOP_CHANNEL_COUNT 4u            /* src/mixer/mod.rs:32 */
#define GOOEY_PITCH_MODE_OFF 0u                /* :7163-7168 */
#define GOOEY_PITCH_MODE_RESAMPLE 1u
#define GOOEY_PITCH_MODE_PRESERVE_PITCH 2u
bool gooey_engine_loop_load(GooeyEngine* engine, uint32_t channel, const float* samples, uint32_t frames, uint32_t channels, float sample_rate);   /* :7184 */
void gooey_engine_loop_set_playing(GooeyEngine* engine, uint32_t channel, bool playing);            /* :7209 */
void gooey_engine_loop_set_gain(GooeyEngine* engine, uint32_t channel, float gain);                 /* :7224, 0 .. 2, 15 ms fader */
void gooey_engine_loop_set_mute(GooeyEngine* engine, uint32_t channel, bool muted);                 /* :7239 */
void gooey_engine_loop_set_solo(GooeyEngine* engine, uint32_t channel, bool soloed);                /* :7255 */
void gooey_engine_loop_set_start(GooeyEngine* engine, uint32_t channel, float normalized);          /* :7275; end < start plays the wrap-around region */
void gooey_engine_loop_set_end(GooeyEngine* engine, uint32_t channel, float normalized);            /* :7295 */
void gooey_engine_loop_set_speed(GooeyEngine* engine, uint32_t channel, float speed);               /* :7311, -4 .. 4, negative = reverse */
void gooey_engine_loop_set_source_bpm(GooeyEngine* engine, uint32_t channel, float source_bpm);     /* :7331, <= 0 clears the tag */
float gooey_engine_loop_get_source_bpm(const GooeyEngine* engine, uint32_t channel);                /* :7352 */
void gooey_engine_loop_set_pitch_mode(GooeyEngine* engine, uint32_t channel, uint32_t mode);        /* :7368 */
uint32_t gooey_engine_loop_get_pitch_mode(const GooeyEngine* engine, uint32_t channel);             /* :7389 */
void gooey_engine_loop_restart(GooeyEngine* engine, uint32_t channel);                              /* :7408 */
void gooey_engine_loop_set_position(GooeyEngine* engine, uint32_t channel, float normalized);       /* :7422 */
float gooey_engine_loop_get_position(const GooeyEngine* engine, uint32_t channel);                  /* :7516 */
/* a take staged to replace the channel's buffer at the next boundary of the loop split into `divisions` equal parts (1 = at the wrap) */
bool gooey_engine_loop_queue_swap(GooeyEngine* engine, uint32_t channel, const float* samples, uint32_t frames, uint32_t channels, float sample_rate,
                                  float source_bpm, uint32_t divisions);                            /* :7449 */
void gooey_engine_loop_cancel_queued_swap(GooeyEngine* engine, uint32_t channel);                   /* :7483 */
uint32_t gooey_engine_loop_swaps_completed(const GooeyEngine* engine, uint32_t channel);            /* :7500 */
int32_t gooey_engine_loop_effect_add(GooeyEngine* engine, uint32_t channel, uint32_t effect_id);    /* :7536, not built: -1 + sticky error */
/* Offline render of one loop channel, ignoring mute / solo, from its loop start (mixer/mod.rs:444-476): a stereo 32-bit float WAV
 * (:8006-8048), or (libgooey_b200 addition) the same frames interleaved into out_interleaved[2 * frames]. */
bool gooey_engine_loop_render_to_wav(GooeyEngine* engine, uint32_t channel, uint32_t frame_count, uint32_t preroll_frame_count, const char* utf8_path);
bool gooey_engine_loop_render(GooeyEngine* engine, uint32_t channel, uint32_t frames, uint32_t preroll, float* out_interleaved);
/* libgooey_b200 addition: dst's channel plays the buffer already loaded into src's channel (same device), without another device copy. */
bool gooey_b200_loop_share_buffer(GooeyEngine* dst, uint32_t dst_channel, const GooeyEngine* src, uint32_t src_channel);

/* ---- sampler racks: up to 4 racks of 16 PCM pads and 32 voices, graph sources GOOEY_SOURCE_SAMPLER_BASE + rack, unrouted until
 * gooey_engine_mixer_route_source (ffi.rs:6000-6172; src/instruments/sampler.rs).  Pads are interleaved f32, 1 or 2 channels.
 * Pads are fired by gooey_engine_sampler_trigger (at once) or by the rack's step pattern below. ---- */
#define GOOEY_SAMPLER_RACK_MAX 4u              /* :585 */
#define GOOEY_SAMPLER_SLOT_COUNT 16u           /* sampler.rs:14 */
int32_t gooey_engine_sampler_register(GooeyEngine* engine);                                          /* :6007, rack id or -1 */
uint32_t gooey_engine_sampler_get_source_id(const GooeyEngine* engine, uint32_t rack);               /* :6031, UINT32_MAX if not registered */
bool gooey_engine_sampler_set_slot_buffer(GooeyEngine* engine, uint32_t rack, uint32_t slot, const float* samples, uint32_t frames, uint32_t channels,
                                          float sample_rate);                                        /* :6044 */
bool gooey_engine_sampler_clear_slot(GooeyEngine* engine, uint32_t rack, uint32_t slot);             /* :6076 */
bool gooey_engine_sampler_slot_is_loaded(const GooeyEngine* engine, uint32_t rack, uint32_t slot);   /* :6090 */
uint32_t gooey_engine_sampler_slot_frames(const GooeyEngine* engine, uint32_t rack, uint32_t slot);  /* :6104 */
uint32_t gooey_engine_sampler_slot_channels(const GooeyEngine* engine, uint32_t rack, uint32_t slot);   /* :6119 */
float gooey_engine_sampler_slot_sample_rate(const GooeyEngine* engine, uint32_t rack, uint32_t slot);   /* :6134 */
bool gooey_engine_sampler_trigger(GooeyEngine* engine, uint32_t rack, uint32_t slot, float velocity);   /* :6150 */
/* the rack's 16-step pattern: a start is armed on the transport beat (gooey_engine_sequencer_start runs the transport) and quantised to
 * the next sixteenth / quarter / bar; with the transport stopped it is armed at beat 0 and fires when the transport starts */
#define GOOEY_CLIP_QUANTIZE_SIXTEENTH 0u       /* src/mixer/clip_grid.rs:8-10 */
#define GOOEY_CLIP_QUANTIZE_QUARTER 1u
#define GOOEY_CLIP_QUANTIZE_BAR 2u
bool gooey_engine_sampler_set_step(GooeyEngine* engine, uint32_t rack, uint32_t step, bool enabled, uint32_t slot, float velocity);   /* :6173 */
bool gooey_engine_sampler_get_step(const GooeyEngine* engine, uint32_t rack, uint32_t step, bool* out_enabled, uint32_t* out_slot, float* out_velocity);   /* :6265 */
bool gooey_engine_sampler_start_pattern(GooeyEngine* engine, uint32_t rack, uint32_t quantization);  /* :6192 */
bool gooey_engine_sampler_stop_pattern(GooeyEngine* engine, uint32_t rack);                          /* :6211 */
bool gooey_engine_sampler_cancel_pattern_start(GooeyEngine* engine, uint32_t rack);                  /* :6225 */
double gooey_engine_sampler_get_pending_start_beat(const GooeyEngine* engine, uint32_t rack);        /* :6239, -1 when none */
bool gooey_engine_sampler_is_pattern_running(const GooeyEngine* engine, uint32_t rack);              /* :6253 */
double gooey_engine_transport_get_beat_position(const GooeyEngine* engine);                          /* :7143: the mixer transport's beat clock */

#ifdef __cplusplus
}
#endif
#endif /* GOOEY_H */
