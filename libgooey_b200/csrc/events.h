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
  EV_RELEASE = 4,     // note-off (poly / bass release)
  EV_NOTE = 5,        // poly: param = midi note, value = velocity
};
enum : uint32_t {
  AUX_OVERSAMPLING = 1,        // 0 / 2 / 4
  AUX_SNARE_FILTER_TYPE = 2,
  AUX_SNARE_PITCH_START = 3,
  AUX_HAT_PINK = 4,
  AUX_HAT_DB24 = 5,
  AUX_TOM_CONFIG_DONE = 6,
  AUX_TOM_RAW_PARAM0 = 16,     // .. +7 : Tom2::set_config fields (unclamped)
  AUX_BASS_RAW = 32,
};

struct VoiceEvent {
  uint32_t frame;   // frame index relative to the start of the launch's frame range origin
  uint16_t kind;
  uint16_t param;
  float value;
};

}  // namespace gd
