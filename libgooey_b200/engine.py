"""Engine-level API: the reference's C FFI (``gooey_engine_*``, include/gooey.h) bound with ctypes, plus the batch
bounce that renders thousands of engines in one device pass.

``Engine`` mirrors the FFI one method per function (``e.set_kick_param(p, v)`` = ``gooey_engine_set_kick_param``).
The same class drives the CPU oracle in the tests (``prefix="orc_engine_"``), so one call script runs on both sides.
"""
import ctypes

import numpy as np

from ._lib import GooeyError, check, lib

c = ctypes
_SIGS = {
    "set_kick_param": [c.c_uint32, c.c_float], "set_snare_param": [c.c_uint32, c.c_float], "set_hihat_param": [c.c_uint32, c.c_float],
    "set_tom_param": [c.c_uint32, c.c_float], "set_bass_param": [c.c_uint32, c.c_float],
    "set_channel_param": [c.c_uint32, c.c_uint32, c.c_float], "load_bass_preset": [c.c_uint32],
    "set_bpm": [c.c_float], "set_swing": [c.c_float], "set_master_gain": [c.c_float],
    "sequencer_set_instrument_step": [c.c_uint32, c.c_uint32, c.c_bool],
    "sequencer_set_instrument_step_settings": [c.c_uint32, c.c_uint32, c.c_bool, c.c_bool, c.c_float, c.c_bool, c.c_float, c.c_float, c.c_bool, c.c_uint8],
    "sequencer_set_step": [c.c_uint32, c.c_bool],
    "sequencer_set_instrument_step_with_velocity": [c.c_uint32, c.c_uint32, c.c_bool, c.c_float],
    "sequencer_set_instrument_step_note": [c.c_uint32, c.c_uint32, c.c_uint8],
    "sequencer_start": [], "sequencer_stop": [], "sequencer_reset": [],
    "set_sequencer_triggers_enabled": [c.c_bool],
    "set_channel_instrument_type": [c.c_uint32, c.c_uint32], "set_compressor_sidechain": [c.c_uint32],
    "trigger_instrument": [c.c_uint32],
    "set_instrument_gain": [c.c_uint32, c.c_float], "set_instrument_pan": [c.c_uint32, c.c_float],
    "set_instrument_mute": [c.c_uint32, c.c_bool], "set_instrument_solo": [c.c_uint32, c.c_bool],
    "trigger_instrument_with_velocity": [c.c_uint32, c.c_float],
    "set_global_effect_param": [c.c_uint32, c.c_uint32, c.c_float], "set_global_effect_enabled": [c.c_uint32, c.c_bool],
    "mixer_set_track_gain": [c.c_uint32, c.c_float], "mixer_set_track_pan": [c.c_uint32, c.c_float],
    "mixer_set_track_mute": [c.c_uint32, c.c_bool], "mixer_set_track_solo": [c.c_uint32, c.c_bool],
    "track_effect_set_param": [c.c_uint32, c.c_uint32, c.c_uint32, c.c_float],
    "poly_release": [], "poly_set_preset": [c.c_uint32], "poly_set_param": [c.c_uint32, c.c_float],
    "granulator_trigger": [c.c_float], "granulator_set_param": [c.c_uint32, c.c_float], "granulator_set_seed": [c.c_uint32],
    "granulator_snap_params": [],
    "set_lfo_enabled": [c.c_uint32, c.c_bool], "set_lfo_timing": [c.c_uint32, c.c_uint32], "set_lfo_amount": [c.c_uint32, c.c_float],
    "set_lfo_offset": [c.c_uint32, c.c_float], "clear_lfo_routes": [c.c_uint32], "reset_lfo_phase": [c.c_uint32],
    "blend_enable": [c.c_uint32], "blend_disable": [c.c_uint32], "blend_set_position": [c.c_uint32, c.c_float, c.c_float],
    "blend_set_corner_preset": [c.c_uint32, c.c_uint32, c.c_uint32], "blend_reset_corners": [c.c_uint32],
    "sequencer_set_instrument_step_blend": [c.c_uint32, c.c_uint32, c.c_float, c.c_float],
    "sequencer_clear_instrument_step_blend": [c.c_uint32, c.c_uint32],
    # loop mixer (ffi.rs:7150-7535)
    "loop_set_playing": [c.c_uint32, c.c_bool], "loop_set_gain": [c.c_uint32, c.c_float], "loop_set_mute": [c.c_uint32, c.c_bool],
    "loop_set_solo": [c.c_uint32, c.c_bool], "loop_set_start": [c.c_uint32, c.c_float], "loop_set_end": [c.c_uint32, c.c_float],
    "loop_set_speed": [c.c_uint32, c.c_float], "loop_set_source_bpm": [c.c_uint32, c.c_float], "loop_set_pitch_mode": [c.c_uint32, c.c_uint32],
    "loop_restart": [c.c_uint32], "loop_set_position": [c.c_uint32, c.c_float], "loop_cancel_queued_swap": [c.c_uint32],
    "mixer_reset_default_layout": [], "mixer_clear_layout": [],
}
# functions with a return value: name -> (argument types after the handle, result type)
_RSIGS = {
    "loop_swaps_completed": ([c.c_uint32], c.c_uint32),
    "mixer_unroute_source": ([c.c_uint32], c.c_bool), "mixer_get_source_route": ([c.c_uint32], c.c_int32),
    "mixer_get_track_gain": ([c.c_uint32], c.c_float), "mixer_get_track_pan": ([c.c_uint32], c.c_float),
    "mixer_get_track_mute": ([c.c_uint32], c.c_bool), "mixer_get_track_solo": ([c.c_uint32], c.c_bool),
    "loop_get_source_bpm": ([c.c_uint32], c.c_float), "loop_get_pitch_mode": ([c.c_uint32], c.c_uint32), "loop_get_position": ([c.c_uint32], c.c_float),
    "sampler_register": ([], c.c_int32), "sampler_get_source_id": ([c.c_uint32], c.c_uint32),
    "sampler_clear_slot": ([c.c_uint32, c.c_uint32], c.c_bool), "sampler_slot_is_loaded": ([c.c_uint32, c.c_uint32], c.c_bool),
    "sampler_slot_frames": ([c.c_uint32, c.c_uint32], c.c_uint32), "sampler_slot_channels": ([c.c_uint32, c.c_uint32], c.c_uint32),
    "sampler_slot_sample_rate": ([c.c_uint32, c.c_uint32], c.c_float), "sampler_trigger": ([c.c_uint32, c.c_uint32, c.c_float], c.c_bool),
    "sampler_set_step": ([c.c_uint32, c.c_uint32, c.c_bool, c.c_uint32, c.c_float], c.c_bool),
    "sampler_start_pattern": ([c.c_uint32, c.c_uint32], c.c_bool), "sampler_stop_pattern": ([c.c_uint32], c.c_bool),
    "sampler_cancel_pattern_start": ([c.c_uint32], c.c_bool), "sampler_get_pending_start_beat": ([c.c_uint32], c.c_double),
    "sampler_is_pattern_running": ([c.c_uint32], c.c_bool),
    "transport_get_beat_position": ([], c.c_double),
}
_bound = set()


def _bind(L, prefix):
    key = (id(L), prefix)
    if key in _bound:
        return
    for name, args in _SIGS.items():
        f = getattr(L, prefix + name)
        f.argtypes = [c.c_void_p] + args
        f.restype = None
    for name, (args, res) in _RSIGS.items():
        f = getattr(L, prefix + name)
        f.argtypes = [c.c_void_p] + args
        f.restype = res
    f = getattr(L, prefix + "loop_load"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint32, c.c_uint32, c.c_float]; f.restype = c.c_bool
    f = getattr(L, prefix + "loop_queue_swap"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint32, c.c_uint32, c.c_float, c.c_float, c.c_uint32]; f.restype = c.c_bool
    f = getattr(L, prefix + "loop_render"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, c.c_uint32, c.c_void_p]; f.restype = c.c_bool
    f = getattr(L, prefix + "sampler_set_slot_buffer"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, c.c_void_p, c.c_uint32, c.c_uint32, c.c_float]; f.restype = c.c_bool
    getattr(L, prefix + "new").restype = c.c_void_p
    getattr(L, prefix + "new").argtypes = [c.c_float]
    getattr(L, prefix + "free").argtypes = [c.c_void_p]
    getattr(L, prefix + "free").restype = None
    f = getattr(L, prefix + "set_effect_order"); f.argtypes = [c.c_void_p, c.POINTER(c.c_uint32), c.c_uint32]; f.restype = c.c_bool
    f = getattr(L, prefix + "mixer_add_track"); f.argtypes = [c.c_void_p, c.c_char_p]; f.restype = c.c_int32
    f = getattr(L, prefix + "mixer_route_source"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32]; f.restype = c.c_bool
    f = getattr(L, prefix + "track_effect_add"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32]; f.restype = c.c_int32
    f = getattr(L, prefix + "track_effect_remove"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32]; f.restype = c.c_bool
    f = getattr(L, prefix + "track_effect_move"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, c.c_uint32]; f.restype = c.c_bool
    f = getattr(L, prefix + "move_effect"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32]; f.restype = c.c_bool
    f = getattr(L, prefix + "sequencer_set_instrument_pattern"); f.argtypes = [c.c_void_p, c.c_uint32, c.POINTER(c.c_bool)]; f.restype = None
    f = getattr(L, prefix + "poly_trigger_notes"); f.argtypes = [c.c_void_p, c.c_void_p, c.c_uint32, c.c_uint32, c.c_float]; f.restype = None
    f = getattr(L, prefix + "granulator_set_buffer"); f.argtypes = [c.c_void_p, c.c_void_p, c.c_uint32, c.c_float]; f.restype = c.c_bool
    f = getattr(L, prefix + "add_lfo_route"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, c.c_uint32, c.c_float]; f.restype = c.c_uint32
    f = getattr(L, prefix + "remove_lfo_route"); f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32]; f.restype = c.c_bool
    f = getattr(L, prefix + "get_lfo_phase"); f.argtypes = [c.c_void_p, c.c_uint32]; f.restype = c.c_float
    f = getattr(L, prefix + "render"); f.argtypes = [c.c_void_p, c.c_void_p, c.c_uint32]; f.restype = None
    f = getattr(L, prefix + "bounce_to_buffer"); f.argtypes = [c.c_void_p, c.c_uint32, c.POINTER(c.c_uint32)]; f.restype = c.POINTER(c.c_float)
    f = getattr(L, prefix + "free_buffer"); f.argtypes = [c.POINTER(c.c_float), c.c_uint32]; f.restype = None
    _bound.add(key)


class Engine:
    """One GooeyEngine handle.  Methods are the FFI function names without the ``gooey_engine_`` prefix."""

    def __init__(self, sample_rate=44100.0, library=None, prefix="gooey_engine_"):
        self._L = library if library is not None else lib()
        self._prefix = prefix
        _bind(self._L, prefix)
        self._h = getattr(self._L, prefix + "new")(c.c_float(sample_rate))
        if not self._h:
            raise GooeyError("gooey_engine_new failed: " + (lib().gooey_b200_last_error().decode() if library is None else "oracle"))
        self.sample_rate = sample_rate

    def close(self):
        if getattr(self, "_h", None):
            getattr(self._L, self._prefix + "free")(self._h)
            self._h = None

    __del__ = close

    def __getattr__(self, name):
        if name in _SIGS or name in _RSIGS:
            f = getattr(self._L, self._prefix + name)
            return lambda *a: f(self._h, *a)
        raise AttributeError(name)

    # ---- sample-playback sources: the host's decoded PCM as float32 arrays [frames] (mono) or [frames, channels] ----
    @staticmethod
    def _pcm(samples):
        a = np.ascontiguousarray(samples, np.float32)
        if a.ndim == 1:
            a = a[:, None]
        return a, a.shape[0], a.shape[1]

    def loop_load(self, channel, samples, sample_rate):
        a, frames, channels = self._pcm(samples)
        return bool(getattr(self._L, self._prefix + "loop_load")(self._h, channel, a.ctypes.data, frames, channels, c.c_float(sample_rate)))

    def loop_queue_swap(self, channel, samples, sample_rate, source_bpm=0.0, divisions=1):
        a, frames, channels = self._pcm(samples)
        return bool(getattr(self._L, self._prefix + "loop_queue_swap")(self._h, channel, a.ctypes.data, frames, channels, c.c_float(sample_rate),
                                                                        c.c_float(source_bpm), divisions))

    # Synthetic PCM by formula, so that a call script (tests/golden/ref/scripts/*.calls) can load buffers without carrying them:
    # rust/examples/dump_golden.rs evaluates the same integer formula inside libgooey.
    @staticmethod
    def synth_pcm(frames, channels, seed):
        k = np.arange(frames, dtype=np.uint64)[:, None]
        ch = np.arange(channels, dtype=np.uint64)[None, :]
        v = ((k * np.uint64(37) + ch * np.uint64(101) + np.uint64(seed) * np.uint64(977)) * (k % np.uint64(89) + np.uint64(3))) % np.uint64(2001)
        return ((v.astype(np.int64) - 1000).astype(np.float32) / np.float32(1024.0)).astype(np.float32)

    def loop_load_synth(self, channel, frames, channels, sample_rate, seed):
        return self.loop_load(channel, self.synth_pcm(frames, channels, seed), sample_rate)

    def loop_queue_swap_synth(self, channel, frames, channels, sample_rate, seed, source_bpm, divisions):
        return self.loop_queue_swap(channel, self.synth_pcm(frames, channels, seed), sample_rate, source_bpm, divisions)

    def sampler_set_slot_synth(self, rack, slot, frames, channels, sample_rate, seed):
        return self.sampler_set_slot_buffer(rack, slot, self.synth_pcm(frames, channels, seed), sample_rate)

    def loop_share_buffer(self, channel, src, src_channel):
        """gooey_b200_loop_share_buffer (product only): play the buffer `src` already holds on the device."""
        f = self._L.gooey_b200_loop_share_buffer
        f.argtypes = [c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint32]; f.restype = c.c_bool
        return bool(f(self._h, channel, src._h, src_channel))

    def loop_render(self, channel, frames, preroll=0):
        """Mixer::render_channel_to_interleaved: [frames, 2] of one loop channel, offline, from its loop start; None on failure."""
        out = np.zeros((frames, 2), np.float32)
        ok = getattr(self._L, self._prefix + "loop_render")(self._h, channel, frames, preroll, out.ctypes.data)
        return out if ok else None

    def sampler_get_step(self, rack, step):
        """(enabled, pad, velocity) of one step of the rack's pattern, or None."""
        en, slot, vel = c.c_bool(False), c.c_uint32(0), c.c_float(0.0)
        f = getattr(self._L, self._prefix + "sampler_get_step")
        f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, c.POINTER(c.c_bool), c.POINTER(c.c_uint32), c.POINTER(c.c_float)]; f.restype = c.c_bool
        return (en.value, slot.value, vel.value) if f(self._h, rack, step, c.byref(en), c.byref(slot), c.byref(vel)) else None

    def sampler_set_slot_buffer(self, rack, slot, samples, sample_rate):
        a, frames, channels = self._pcm(samples)
        return bool(getattr(self._L, self._prefix + "sampler_set_slot_buffer")(self._h, rack, slot, a.ctypes.data, frames, channels, c.c_float(sample_rate)))

    def set_effect_order(self, ids):
        arr = (c.c_uint32 * len(ids))(*ids)
        return bool(getattr(self._L, self._prefix + "set_effect_order")(self._h, arr, len(ids)))

    def mixer_add_track(self, name="track"):
        return int(getattr(self._L, self._prefix + "mixer_add_track")(self._h, name.encode()))

    def mixer_route_source(self, source, track):
        return bool(getattr(self._L, self._prefix + "mixer_route_source")(self._h, source, track))

    def track_effect_add(self, track, effect_id):
        return int(getattr(self._L, self._prefix + "track_effect_add")(self._h, track, effect_id))

    def track_effect_remove(self, track, slot):
        return bool(getattr(self._L, self._prefix + "track_effect_remove")(self._h, track, slot))

    def track_effect_move(self, track, slot, new_position):
        return bool(getattr(self._L, self._prefix + "track_effect_move")(self._h, track, slot, new_position))

    def move_effect(self, effect_id, new_position):
        return bool(getattr(self._L, self._prefix + "move_effect")(self._h, effect_id, new_position))

    def sequencer_set_instrument_pattern(self, instrument, pattern16):
        arr = (c.c_bool * 16)(*[bool(x) for x in pattern16])
        getattr(self._L, self._prefix + "sequencer_set_instrument_pattern")(self._h, instrument, arr)

    def poly_trigger_chord(self, root, scale, degree, voicing, preset=0, octave=4, velocity=1.0):
        """gooey_engine_poly_trigger_chord (product only; the oracle takes the notes: poly_trigger_notes)."""
        f = self._L.gooey_engine_poly_trigger_chord
        f.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, c.c_uint32, c.c_uint32, c.c_uint32, c.c_int32, c.c_float]; f.restype = None
        f(self._h, root, scale, degree, voicing, preset, octave, c.c_float(velocity))

    def poly_trigger_notes(self, notes, preset=0, velocity=1.0):
        """The tail of gooey_engine_poly_trigger_chord once the voicing has produced MIDI notes (ffi.rs:5594-5611)."""
        arr = np.asarray(notes, np.uint8)
        getattr(self._L, self._prefix + "poly_trigger_notes")(self._h, arr.ctypes.data, len(arr), preset, c.c_float(velocity))

    def granulator_set_buffer(self, samples, sample_rate):
        """gooey_engine_granulator_set_buffer (copies; False on empty / non-finite input)."""
        arr = np.ascontiguousarray(samples, np.float32)
        return bool(getattr(self._L, self._prefix + "granulator_set_buffer")(self._h, arr.ctypes.data, len(arr), c.c_float(sample_rate)))

    def granulator_share_buffer(self, other):
        """libgooey_b200 addition: play the buffer already resident for `other` (no second device copy)."""
        f = self._L.gooey_b200_granulator_share_buffer
        f.argtypes = [c.c_void_p, c.c_void_p]; f.restype = c.c_bool
        return bool(f(self._h, other._h))

    def add_lfo_route(self, lfo, instrument, param, depth):
        return int(getattr(self._L, self._prefix + "add_lfo_route")(self._h, lfo, instrument, param, c.c_float(depth)))

    def remove_lfo_route(self, lfo, route_id):
        return bool(getattr(self._L, self._prefix + "remove_lfo_route")(self._h, lfo, route_id))

    def get_lfo_phase(self, lfo):
        return float(getattr(self._L, self._prefix + "get_lfo_phase")(self._h, lfo))

    def get_channel_peaks(self, count=5):
        """gooey_engine_get_channel_peaks: read-and-reset pre-pan peaks of the five voice strips."""
        out = (c.c_float * count)()
        f = getattr(self._L, self._prefix + "get_channel_peaks")
        f.argtypes = [c.c_void_p, c.POINTER(c.c_float), c.c_uint32]; f.restype = None
        f(self._h, out, count)
        return np.array(out[:], np.float32)

    def mixer_get_track_peak(self, track):
        f = getattr(self._L, self._prefix + "mixer_get_track_peak")
        f.argtypes = [c.c_void_p, c.c_uint32]; f.restype = c.c_float
        return float(f(self._h, track))

    def get_sequencer_triggers_enabled(self):
        f = getattr(self._L, self._prefix + "get_sequencer_triggers_enabled")
        f.argtypes = [c.c_void_p]; f.restype = c.c_bool
        return bool(f(self._h))

    def drain_midi_events(self, max_events=64):
        """gooey_engine_drain_midi_events: [(instrument_index, velocity, sample_offset)] of the last render call."""
        class Ev(c.Structure):
            _fields_ = [("instrument_index", c.c_uint32), ("velocity", c.c_float), ("sample_offset", c.c_uint32)]
        buf = (Ev * max_events)()
        f = getattr(self._L, self._prefix + "drain_midi_events")
        f.argtypes = [c.c_void_p, c.c_void_p, c.c_uint32]; f.restype = c.c_uint32
        n = f(self._h, buf, max_events)
        return [(int(buf[i].instrument_index), float(buf[i].velocity), int(buf[i].sample_offset)) for i in range(n)]

    # ---- product-only entry points (no oracle counterpart) ----
    def bounce_to_wav(self, bars, path):
        """gooey_engine_bounce_to_wav: mono 16-bit PCM of the bounce."""
        f = self._L.gooey_engine_bounce_to_wav
        f.argtypes = [c.c_void_p, c.c_uint32, c.c_char_p]; f.restype = c.c_bool
        return bool(f(self._h, bars, str(path).encode()))

    def has_error(self):
        f = self._L.gooey_engine_has_error
        f.argtypes = [c.c_void_p]; f.restype = c.c_bool
        return bool(f(self._h))

    def get_error_message(self):
        f = self._L.gooey_engine_get_error_message
        f.argtypes = [c.c_void_p]; f.restype = c.c_char_p
        m = f(self._h)
        return m.decode() if m else None

    def set_error_callback(self, fn):
        """gooey_engine_set_error_callback(engine, context, callback); `fn(message: str)`."""
        CB = c.CFUNCTYPE(None, c.c_void_p, c.c_char_p)
        self._cb = CB(lambda ctx, msg: fn(msg.decode() if msg else ""))      # keep the thunk alive
        f = self._L.gooey_engine_set_error_callback
        f.argtypes = [c.c_void_p, c.c_void_p, CB]; f.restype = None
        f(self._h, None, self._cb)

    def get_effect_order(self):
        out = (c.c_uint32 * 9)()
        f = self._L.gooey_engine_get_effect_order
        f.argtypes = [c.c_void_p, c.POINTER(c.c_uint32), c.c_uint32]; f.restype = c.c_uint32
        n = f(self._h, out, 9)
        return list(out[:n])

    def render(self, frames):
        """gooey_engine_render: (frames, 2) interleaved stereo."""
        out = np.zeros((frames, 2), np.float32)
        getattr(self._L, self._prefix + "render")(self._h, out.ctypes.data, frames)
        return out

    def bounce_to_buffer(self, bars):
        n = c.c_uint32(0)
        p = getattr(self._L, self._prefix + "bounce_to_buffer")(self._h, bars, c.byref(n))
        if not p:
            raise GooeyError("bounce_to_buffer returned NULL")
        out = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
        getattr(self._L, self._prefix + "free_buffer")(p, n)
        return out


def batch_bounce(engines, bars):
    """gooey_batch_bounce: every engine bounced `bars` bars in one device pass; returns a list of mono arrays."""
    L = lib()
    n = len(engines)
    hs = (c.c_void_p * n)(*[e._h for e in engines])
    bufs = (c.POINTER(c.c_float) * n)()
    lens = (c.c_uint32 * n)()
    L.gooey_batch_bounce.argtypes = [c.POINTER(c.c_void_p), c.c_uint32, c.c_uint32, c.POINTER(c.POINTER(c.c_float)), c.POINTER(c.c_uint32)]
    check(L.gooey_batch_bounce(hs, n, bars, bufs, lens))
    _bind(L, "gooey_engine_")
    out = []
    for i in range(n):
        out.append(np.ctypeslib.as_array(bufs[i], shape=(lens[i],)).copy())
        L.gooey_engine_free_buffer(bufs[i], lens[i])
    return out


def _block_bounce(name, ctype, dtype, engines, bars, out):
    L = lib()
    n = len(engines)
    hs = (c.c_void_p * n)(*[e._h for e in engines])
    f = getattr(L, name)
    f.argtypes = [c.POINTER(c.c_void_p), c.c_uint32, c.c_uint32, c.c_void_p, c.c_size_t, c.POINTER(c.c_uint32)]
    frames = c.c_uint32(0)
    if out is None:
        # length first (host arithmetic of bounce.rs:20-32 done by the library): probe with the reference formula
        L.gooey_engine_get_bpm.restype = c.c_float
        L.gooey_engine_get_bpm.argtypes = [c.c_void_p]
        sr, bpm = float(engines[0].sample_rate), float(L.gooey_engine_get_bpm(engines[0]._h))
        guess = int(round(bars * 4.0 * (60.0 / bpm) * sr))
        out = np.zeros((n, guess), dtype)
    assert out.dtype == dtype and out.flags.c_contiguous and out.shape[0] == n
    check(f(hs, n, bars, out.ctypes.data, out.shape[1], c.byref(frames)))
    return out[:, :frames.value]


def batch_bounce_host(engines, bars, out=None):
    """gooey_batch_bounce_host: mono bounce of every engine into one (n, pitch) float32 block (drain overlapped with the render;
    engines may live on several devices).  Pass `out` (e.g. a HostBuffer view) to choose the memory; returns out[:, :frames]."""
    return _block_bounce("gooey_batch_bounce_host", c.c_float, np.float32, engines, bars, out)


def batch_bounce_pcm16(engines, bars, out=None):
    """gooey_batch_bounce_pcm16: the same as 16-bit PCM quantised on the device (what bounce_to_wav stores)."""
    return _block_bounce("gooey_batch_bounce_pcm16", c.c_int16, np.int16, engines, bars, out)


def batch_bounce_to_wav(engines, bars, paths):
    """gooey_batch_bounce_to_wav: one mono 16-bit WAV per engine, all bounced in one device pass."""
    L = lib()
    n = len(engines)
    hs = (c.c_void_p * n)(*[e._h for e in engines])
    ps = (c.c_char_p * n)(*[str(p).encode() for p in paths])
    L.gooey_batch_bounce_to_wav.argtypes = [c.POINTER(c.c_void_p), c.c_uint32, c.c_uint32, c.POINTER(c.c_char_p)]
    check(L.gooey_batch_bounce_to_wav(hs, n, bars, ps))


def batch_render(engines, frames, out=None):
    """gooey_batch_render: every engine rendered `frames` frames in one device pass; (n, frames, 2) interleaved stereo."""
    L = lib()
    n = len(engines)
    hs = (c.c_void_p * n)(*[e._h for e in engines])
    if out is None:
        out = np.zeros((n, frames, 2), np.float32)
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (n, frames, 2)
    L.gooey_batch_render.argtypes = [c.POINTER(c.c_void_p), c.c_uint32, c.c_uint32, c.c_void_p]
    check(L.gooey_batch_render(hs, n, frames, out.ctypes.data))
    return out


def batch_bounce_device(engines, bars, dev_ptr, stride):
    """gooey_batch_bounce_device: result stays in device memory (rows of `stride` floats); returns frames."""
    L = lib()
    n = len(engines)
    hs = (c.c_void_p * n)(*[e._h for e in engines])
    frames = c.c_uint32(0)
    L.gooey_batch_bounce_device.argtypes = [c.POINTER(c.c_void_p), c.c_uint32, c.c_uint32, c.c_void_p, c.c_size_t, c.POINTER(c.c_uint32)]
    check(L.gooey_batch_bounce_device(hs, n, bars, c.c_void_p(dev_ptr), stride, c.byref(frames)))
    return frames.value


def set_device(device):
    check(lib().gooey_b200_set_device(int(device)))
