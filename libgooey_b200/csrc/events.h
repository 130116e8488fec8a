// events.h — the host→device event vocabulary.  The host resolves sequencer
// steps, parameter edits and per-step overrides into per-voice event lists
// (sorted by frame); the kernels apply every event whose frame equals the
// current frame before ticking that frame (ffi.rs:1162-1198 ordering).
#pragma once
#include <stdint.h>

namespace gd {

enum : uint32_t {
  EV_TRIGGER = 0,     // value = velocity                       (Instrument::trigger_with_velocity)
  EV_SET_TARGET = 1,  // param = smoother index, value = target (SmoothedParam::set_target)
  EV_SNAP = 2,        // snap every smoother of the voice       (snap_params)
  EV_SET_AUX = 3,     // non-smoothed fields, param = AUX_*
  EV_SET_TIME = 4,    // value = new engine time (bounce resets current_time to 0; voice state persists)
  EV_NOTE_FREQ = 5,   // per-step note override: save freq param once, set_param(0, value), snap (ffi.rs:1176-1190)
  EV_RESTORE_FREQ = 6,// first later step without a note: set_param(0, saved), snap (ffi.rs:1191-1194)
  EV_POLY_NOTE = 7,   // param = midi note, value = velocity (PolySynth::trigger_note)
  EV_POLY_RELEASE = 8,// PolySynth::release_all
  EV_GRAN_SEED = 9,   // aux = seed (Granulator::set_seed)
  EV_GRAN_BUFFER = 10,// buffer swapped: kill all grains (Granulator::set_buffer)
  // engine-mix events (mix.cuh)
  MX_SET = 32,        // param = MixParam index, value = target (SmoothedParam::set_target on a strip/track/master)
  MX_SNAP = 33,       // param = 0 voice strips, 1 graph strips, 2 master gain
  MX_FX_SET = 34,     // param = (fx slot << 8) | effect param id, value = raw FFI value
  MX_FX_INIT = 35,    // param = fx slot, aux = kind | rack << 8, value = bpm : construct the effect's dynamic state
  MX_TRACK_INIT = 36, // param = track index
  MX_FX_BPM = 37,     // param = fx slot, value = bpm
  MX_FX_RESET = 38,   // reset_effect_states (ffi.rs:1417-1425): every global reorderable effect except the waveshapers
};
enum : uint32_t {
  AUX_OVERSAMPLING = 1,        // 0 / 2 / 4
  AUX_SNARE_FILTER_TYPE = 2,
  AUX_SNARE_PITCH_START = 3,
  AUX_HAT_PINK = 4,
  AUX_HAT_DB24 = 5,
  AUX_TOM_CONFIG_DONE = 6,
  AUX_TOM_RAW_PARAM0 = 16,     // .. +7 : Tom2::set_config fields (unclamped)
  AUX_GRAN_BUFINFO = 32,       // value = buffer sample rate, aux = length
};

struct VoiceEvent {
  uint32_t frame;   // frame index relative to the start of the render call
  uint16_t kind;
  uint16_t param;
  float value;
  uint32_t aux;
};

}  // namespace gd
