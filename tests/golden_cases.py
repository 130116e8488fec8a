"""The inputs of the committed golden fixtures (tests/golden/*.npz), shared by the generator and the tests."""
import numpy as np

from libgooey_b200 import voices as V
import engine_scripts as S

KIT_FRAMES = 4096
ENGINE_KEEP = 16384


def kit_patches():
    """Every preset constructor of the reference's drum voices (kick.rs:257-350, snare.rs:270-351, hihat2.rs:79-96,
    tom2.rs:119-172) plus Tom2::new, one trigger at frame 0."""
    patches, names = [], []
    for name, p in V.KICK_PRESETS.items():
        patches.append(V.patch(V.KICK, p)); names.append("kick." + name)
    for name, p in V.SNARE_PRESETS.items():
        patches.append(V.patch(V.SNARE, p)); names.append("snare." + name)
    for name, p in V.HIHAT_PRESETS.items():
        patches.append(V.patch(V.HIHAT, p)); names.append("hihat." + name)
    for name, p in V.TOM_PRESETS.items():
        patches.append(V.patch(V.TOM, p, aux=1)); names.append("tom." + name)
    patches.append(V.patch(V.TOM)); names.append("tom.new")
    vel = np.linspace(0.4, 1.0, len(patches)).astype(np.float32)
    return patches, vel, names


def _default_pattern(e):
    for s in (0, 4, 8, 12):
        e.sequencer_set_instrument_step(S.KICK, s, True)
    for s in (2, 6, 10, 14):
        e.sequencer_set_instrument_step(S.HIHAT, s, True)
    e.sequencer_set_instrument_step(S.SNARE, 4, True)
    e.sequencer_set_instrument_step(S.BASS, 0, True)


def _fx_pattern(e):
    S.pattern_engine(e, 3, swing=0.58)
    S.fx_chain(e, 4, plate=True, limiter=True)


def _sample_playback(e):
    """Loop mixer (direct read with a queued take, WSOLA time-stretch, reverse Resample warp on a wrap-around window) and a sampler rack
    (a manual hit + its step pattern on the transport) next to the default pattern; every buffer is synth_pcm (libgooey_b200/engine.py)."""
    _default_pattern(e)
    e.set_bpm(128.0)
    e.loop_load_synth(0, 9000, 2, 48000.0, 1)
    e.loop_set_start(0, 0.1); e.loop_set_end(0, 0.85); e.loop_set_speed(0, 0.9); e.loop_set_gain(0, 0.7); e.loop_restart(0); e.loop_set_playing(0, True)
    e.loop_queue_swap_synth(0, 6000, 2, 44100.0, 4, 120.0, 4)
    e.loop_load_synth(1, 30000, 2, 44100.0, 2)
    e.loop_set_source_bpm(1, 100.0); e.loop_set_pitch_mode(1, 2); e.loop_set_gain(1, 0.5); e.loop_set_playing(1, True)
    e.loop_load_synth(2, 5000, 1, 44100.0, 3)
    e.loop_set_start(2, 0.7); e.loop_set_end(2, 0.3); e.loop_restart(2)
    e.loop_set_source_bpm(2, 140.0); e.loop_set_pitch_mode(2, 1); e.loop_set_speed(2, -1.0); e.loop_set_gain(2, 0.4); e.loop_set_playing(2, True)
    e.sampler_register()
    e.mixer_route_source(5, 0)
    e.sampler_set_slot_synth(0, 0, 8000, 1, 44100.0, 5)
    e.sampler_set_slot_synth(0, 1, 3000, 2, 22050.0, 6)
    e.sampler_set_slot_synth(0, 2, 1200, 2, 48000.0, 7)
    for step in range(16):
        e.sampler_set_step(0, step, step % 3 != 1, step % 3, 0.3 + 0.04 * step)
    e.sampler_trigger(0, 0, 0.8)
    e.sampler_start_pattern(0, 2)
    e.sequencer_start()


ENGINE_CASES = {"default_pattern": _default_pattern, "graph_swing_fx": _fx_pattern}
# cases that exist only as reference-vector scripts (tests/golden/ref/scripts): no oracle-made fixture is committed for them
REF_CASES = {"sample_playback": _sample_playback}
