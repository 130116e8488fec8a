"""GPU parity for every exported engine entry point that the round-1 suite did not reach: track effect racks of every kind,
effect reordering, delay ping-pong, instrument solo, track mute / solo, bass presets, per-channel parameters and
instrument swap, every sequencer setter, the batch stereo render, bounce_to_wav, the sticky error + callback, and the
saturation / compressor (+ side-chain) / low-pass / waveshaper / feedback-waveshaper effect slots.
One FFI-named call script drives the CPU oracle and the product; tolerance 1e-5 of full scale (BASELINE.json)."""
import ctypes
import os
import struct

import numpy as np
import pytest

from libgooey_b200 import engine as G
import oracle_lib as O
import engine_scripts as S

pytestmark = pytest.mark.gpu
TOL = 1e-5
FX_LOWPASS, FX_DELAY, FX_SAT, FX_COMP, FX_TILT, FX_LIMITER, FX_REVERB, FX_WS, FX_FBWS, FX_PLATE = range(10)


def busy_pattern(e, snare_overdrive=0.0):
    """Every voice plays, nothing blows up: default (preset) voices, a dense pattern, pans apart so that L != R."""
    for s in (0, 3, 6, 8, 11, 14):
        e.sequencer_set_instrument_step(S.KICK, s, True)
    for s in (4, 12):
        e.sequencer_set_instrument_step(S.SNARE, s, True)
    for s in range(0, 16, 2):
        e.sequencer_set_instrument_step_with_velocity(S.HIHAT, s, True, 0.5 + 0.03 * s)
    for s in (2, 10):
        e.sequencer_set_instrument_step(S.TOM, s, True)
    for s in (0, 7, 9):
        e.sequencer_set_instrument_step(S.BASS, s, True)
    e.set_instrument_pan(S.KICK, 0.35)
    e.set_instrument_pan(S.HIHAT, 0.8)
    e.set_instrument_pan(S.BASS, 0.6)
    e.set_snare_param(15, snare_overdrive)
    e.set_master_gain(0.6)


def both(script, bars=1, stereo_frames=None):
    o = O.oracle_engine()
    g = G.Engine()
    script(o)
    script(g)
    if stereo_frames:
        for e in (o, g):
            e.sequencer_reset(); e.sequencer_start()
        want, got = o.render(stereo_frames), g.render(stereo_frames)
    else:
        want, got = o.bounce_to_buffer(bars), g.bounce_to_buffer(bars)
    assert not g.has_error(), g.get_error_message()
    o.close(); g.close()
    assert got.shape == want.shape and np.isfinite(want).all()
    assert np.abs(want).max() > 0.01
    return got, want


RACK_PARAMS = {
    FX_LOWPASS: [(0, 1800.0), (1, 0.6)], FX_DELAY: [(0, 3.0), (1, 0.5), (2, 0.4), (3, 6000.0)], FX_SAT: [(0, 0.7), (1, 0.6), (2, 0.8)],
    FX_COMP: [(0, -30.0), (1, 8.0), (2, 2.0), (3, 60.0), (4, 0.9)], FX_TILT: [(0, 0.25), (1, 0.4)], FX_REVERB: [(0, 0.6), (1, 0.5), (2, 0.3)],
    FX_WS: [(0, 5.0), (1, 0.8)], FX_FBWS: [(0, 6.0), (1, 0.4), (2, 3000.0), (3, 0.7)], FX_PLATE: [(0, 0.6), (1, 0.5), (2, 0.4), (5, 0.35)],
}


@pytest.mark.parametrize("kind", sorted(RACK_PARAMS))
def test_track_rack_of_every_effect_kind(kind):
    def script(e):
        busy_pattern(e)
        assert e.track_effect_add(0, kind) == 0                 # drum-kit track
        for p, v in RACK_PARAMS[kind]:
            e.track_effect_set_param(0, 0, p, v)
    got, want = both(script, stereo_frames=44100)
    err = np.abs(got - want).max()
    print(f"rack kind {kind}: err {err:.3e} peak {np.abs(want).max():.3f}")
    assert err <= TOL


def test_rack_chain_add_move_remove_and_limiter_is_not_a_rack_effect():
    def script(e):
        busy_pattern(e)
        assert e.track_effect_add(1, FX_LIMITER) == -1
        assert e.track_effect_add(1, FX_SAT) == 0 and e.track_effect_add(1, FX_DELAY) == 1 and e.track_effect_add(1, FX_TILT) == 2
        e.track_effect_set_param(1, 1, 2, 0.5)
        e.track_effect_set_param(1, 2, 0, 0.8)
        assert e.track_effect_move(1, 0, 5)                     # saturation to the end (clamped)
        assert e.track_effect_remove(1, 0)                      # drops the delay
        assert not e.track_effect_remove(1, 7)
        e.track_effect_set_param(1, 1, 0, 0.9)                  # now the saturation
        assert e.track_effect_add(1, FX_REVERB) == 2
    got, want = both(script, bars=1)
    assert np.abs(got - want).max() <= TOL


@pytest.mark.parametrize("order", [[9, 6, 3, 1, 4, 0, 2, 8, 7], [1, 6, 4, 9, 0, 2, 3, 7, 8]])
def test_permuted_effect_order_with_every_global_effect_enabled(order):
    def script(e):
        busy_pattern(e)
        S.fx_chain(e, 21, plate=True)
        for fx, params in RACK_PARAMS.items():
            if fx in (FX_LOWPASS, FX_SAT, FX_COMP, FX_WS, FX_FBWS):
                for p, v in params:
                    e.set_global_effect_param(fx, p, v)
                e.set_global_effect_enabled(fx, True)
        assert not e.set_effect_order([5] + order[1:])           # limiter is not reorderable
        assert not e.set_effect_order(order[:8])
        assert e.set_effect_order(order)
        assert e.move_effect(order[0], 3)
    got, want = both(script, stereo_frames=30000)
    err = np.abs(got - want).max()
    print(f"effect order {order}: err {err:.3e}")
    assert err <= TOL


def test_reorder_after_audio_resets_effect_states():
    def run(e):
        busy_pattern(e)
        S.fx_chain(e, 4, plate=True)
        e.set_global_effect_param(FX_SAT, 2, 0.7); e.set_global_effect_enabled(FX_SAT, True)
        e.sequencer_start()
        a = e.render(12000)
        assert e.set_effect_order([9, 6, 3, 1, 4, 0, 2, 8, 7])   # reset_effect_states: delay lines / tanks cleared mid-stream
        b = e.render(12000)
        return np.concatenate([a, b])
    o = O.oracle_engine(); g = G.Engine()
    want, got = run(o), run(g)
    o.close(); g.close()
    assert np.abs(got - want).max() <= TOL


def test_delay_pingpong_stereo():
    def script(e):
        busy_pattern(e)
        for p, v in [(0, 4.0), (1, 0.6), (2, 0.5), (3, 5000.0), (4, 1.0)]:
            e.set_global_effect_param(FX_DELAY, p, v)
        e.set_global_effect_enabled(FX_DELAY, True)
    got, want = both(script, stereo_frames=44100)
    assert np.abs(want[:, 0] - want[:, 1]).max() > 1e-3          # the echoes really alternate sides
    assert np.abs(got - want).max() <= TOL


def test_instrument_solo_and_track_mute_solo():
    def script(e):
        busy_pattern(e)
        e.set_instrument_solo(S.SNARE, True)
        e.set_instrument_solo(S.BASS, True)
        e.set_instrument_mute(S.BASS, True)                      # solo wins over mute
        t = e.mixer_add_track("aux")
        e.mixer_route_source(1, t)                               # bass -> aux
        e.mixer_set_track_mute(0, True)
        e.mixer_set_track_solo(t, True)
    o = O.oracle_engine(); g = G.Engine()
    script(o); script(g)
    for e in (o, g):
        e.sequencer_start()
    a = [o.render(6000)]; b = [g.render(6000)]
    for e in (o, g):
        e.mixer_set_track_solo(4, False)                         # targets change between renders: mute gains glide
        e.set_instrument_solo(S.SNARE, False)
    a.append(o.render(9000)); b.append(g.render(9000))
    for e in (o, g):
        e.mixer_set_track_mute(0, False)
        e.set_instrument_solo(S.BASS, False)                     # the bass is now plainly muted
    a.append(o.render(9000)); b.append(g.render(9000))
    o.close(); g.close()
    want, got = np.concatenate(a), np.concatenate(b)
    assert np.abs(want).max() > 0.01
    assert np.abs(got - want).max() <= TOL


@pytest.mark.parametrize("preset", [0, 1, 2, 3])
def test_load_bass_preset(preset):
    def script(e):
        for s in (0, 4, 6, 10):
            e.sequencer_set_instrument_step_with_velocity(S.BASS, s, True, 0.9)
        e.sequencer_set_instrument_step_note(S.BASS, 4, 43)
        e.load_bass_preset(preset)
        e.load_bass_preset(9)                                    # unknown id: ignored
    got, want = both(script, bars=1)
    err = np.abs(got - want).max()
    print(f"bass preset {preset}: err {err:.3e} peak {np.abs(want).max():.3f}")
    assert err <= TOL


def test_set_channel_param_and_instrument_type_swap():
    def script(e):
        busy_pattern(e)
        e.set_channel_param(0, 1, 0.9)                           # kick punch through the channel interface
        e.set_channel_param(2, 1, 0.4)                           # hi-hat decay
        e.set_channel_param(9, 0, 0.5)                           # bad channel: ignored
        e.set_channel_instrument_type(3, S.SNARE)                # the tom channel becomes a second snare (SnareDrum::new)
        e.set_channel_param(3, 0, 0.8)                           # ... and takes snare parameter ids
        e.set_channel_instrument_type(1, S.TOM)
        e.set_channel_instrument_type(1, S.TOM)                  # no-op
        e.set_channel_instrument_type(0, 7)                      # unknown type: ignored
    got, want = both(script, bars=1)
    assert np.abs(got - want).max() <= TOL


def test_instrument_type_swap_between_renders():
    def run(e):
        busy_pattern(e)
        e.sequencer_start()
        a = e.render(20000)
        e.set_channel_instrument_type(2, S.BASS)                 # a fresh bass joins at the engine's current time
        e.set_channel_param(2, 0, 0.7)
        b = e.render(30000)
        return np.concatenate([a, b])
    o = O.oracle_engine(); g = G.Engine()
    want, got = run(o), run(g)
    o.close(); g.close()
    assert np.abs(got - want).max() <= TOL


def test_every_sequencer_setter():
    def script(e):
        e.sequencer_set_step(0, True); e.sequencer_set_step(8, True); e.sequencer_set_step(99, True)     # kick only; bad step ignored
        e.sequencer_set_instrument_pattern(S.HIHAT, [i % 3 == 0 for i in range(16)])
        e.sequencer_set_instrument_step_with_velocity(S.SNARE, 4, True, 0.7)
        e.sequencer_set_instrument_step_with_velocity(S.SNARE, 12, True, 1.7)                           # clamped to 1
        e.sequencer_set_instrument_step(S.TOM, 6, True)
        e.sequencer_set_instrument_step_note(S.TOM, 6, 50)
        e.sequencer_set_instrument_step(S.TOM, 14, True)                                                 # no note: restores the tune
        e.sequencer_set_instrument_step_settings(S.BASS, 2, True, True, 0.8, False, 0.0, 0.0, True, 40)
        e.sequencer_set_instrument_step_settings(S.BASS, 10, True, False, 0.0, False, 0.0, 0.0, True, 255)
        e.sequencer_set_instrument_step_note(S.BASS, 2, 255)                                             # cleared again
        e.sequencer_set_instrument_step_note(S.KICK, 8, 38)
    got, want = both(script, bars=2)
    assert np.abs(got - want).max() <= TOL


def test_manual_trigger_default_velocity_and_stop_reset():
    def run(e):
        busy_pattern(e)
        e.trigger_instrument(S.TOM)
        e.sequencer_start()
        a = e.render(15000)
        e.sequencer_stop()
        e.trigger_instrument(S.KICK)
        b = e.render(8000)
        e.sequencer_reset(); e.sequencer_start()
        c = e.render(15000)
        return np.concatenate([a, b, c])
    o = O.oracle_engine(); g = G.Engine()
    want, got = run(o), run(g)
    o.close(); g.close()
    assert np.abs(want).max() > 0.01
    assert np.abs(got - want).max() <= TOL


def test_batch_render_stereo_matches_oracle_renders():
    n, frames = 12, 20000

    def script(e, i):
        busy_pattern(e, snare_overdrive=0.1 * (i % 4))
        e.set_swing(0.5 + 0.02 * i)
        if i % 3 == 0:
            S.fx_chain(e, 40 + i, plate=(i % 2 == 0))
        e.sequencer_start()
    engines = [G.Engine() for _ in range(n)]
    for i, e in enumerate(engines):
        script(e, i)
    got = G.batch_render(engines, frames)
    got2 = G.batch_render(engines, 5000)                          # state and clock carry over
    for e in engines:
        e.close()
    for i in range(n):
        o = O.oracle_engine(); script(o, i)
        want = o.render(frames); want2 = o.render(5000)
        o.close()
        assert np.abs(got[i] - want).max() <= TOL, i
        assert np.abs(got2[i] - want2).max() <= TOL, i


def test_bounce_to_wav_bytes(tmp_path):
    def script(e):
        busy_pattern(e, snare_overdrive=0.3)
    o = O.oracle_engine(); g = G.Engine()
    script(o); script(g)
    want = o.bounce_to_buffer(1)
    path = tmp_path / "bounce.wav"
    assert g.bounce_to_wav(1, path)
    assert not g.bounce_to_wav(1, tmp_path / "no_such_dir" / "x.wav")
    o.close(); g.close()
    raw = path.read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
    fmt = struct.unpack("<HHIIHH", raw[20:36])
    assert fmt == (1, 1, 44100, 88200, 2, 16)                    # PCM, mono, 44.1 kHz, 16 bit (ffi.rs:7957-7964)
    data = np.frombuffer(raw[44:], dtype="<i2")
    assert len(data) == len(want) == 88200 and struct.unpack("<I", raw[40:44])[0] == 2 * 88200
    q = np.clip(np.sign(want) * np.floor(np.abs(want * np.float32(32767.0)) + 0.5), -32768, 32767).astype(np.int16)   # (s * 32767).round() as i16
    assert np.abs(data.astype(np.int32) - q.astype(np.int32)).max() <= 1       # 1e-5 of full scale is a third of one 16-bit step
    assert (data != q).mean() < 0.01


def test_sticky_error_and_callback_when_an_engine_runs_out_of_effect_slots():
    g = G.Engine()
    seen = []
    g.set_error_callback(seen.append)
    assert not g.has_error() and g.get_error_message() is None
    for k in range(4):
        assert g.track_effect_add(0, FX_TILT) == k
    assert g.track_effect_add(0, FX_TILT) == -1                  # a rack holds four here: refused loudly, not silently
    assert g.has_error() and "rack" in g.get_error_message()
    assert len(seen) == 1 and seen[0] == g.get_error_message()
    g.track_effect_add(0, FX_TILT)
    assert len(seen) == 1                                        # fired once (ffi.rs:2236-2284)
    out = g.render(256)
    assert not out.any()                                         # an engine in the error state renders silence
    g.close()


@pytest.mark.parametrize("fx", [FX_LOWPASS, FX_SAT, FX_COMP, FX_WS, FX_FBWS])
def test_each_new_global_effect_alone(fx):
    def script(e):
        busy_pattern(e, snare_overdrive=0.2)
        for p, v in RACK_PARAMS[fx]:
            e.set_global_effect_param(fx, p, v)
        e.set_global_effect_param(fx, 17, 0.5)                   # unknown parameter id: ignored
        e.set_global_effect_enabled(fx, True)
    got, want = both(script, stereo_frames=44100)
    err = np.abs(got - want).max()
    print(f"global effect {fx}: err {err:.3e} peak {np.abs(want).max():.3f}")
    assert err <= TOL


def test_global_effects_glide_from_their_defaults_when_enabled_late():
    def run(e):
        busy_pattern(e)
        e.sequencer_start()
        a = e.render(9000)
        e.set_global_effect_param(FX_SAT, 0, 0.9)
        e.set_global_effect_param(FX_LOWPASS, 0, 900.0)
        e.set_global_effect_param(FX_LOWPASS, 1, 0.8)
        e.set_global_effect_enabled(FX_SAT, True)
        e.set_global_effect_enabled(FX_LOWPASS, True)
        b = e.render(9000)
        e.set_global_effect_enabled(FX_SAT, False)               # state is kept while disabled
        c = e.render(4000)
        e.set_global_effect_enabled(FX_SAT, True)
        d = e.render(9000)
        return np.concatenate([a, b, c, d])
    o = O.oracle_engine(); g = G.Engine()
    want, got = run(o), run(g)
    o.close(); g.close()
    assert np.abs(got - want).max() <= TOL


@pytest.mark.parametrize("sidechain", [0xFFFFFFFF, S.KICK, S.BASS])
def test_bounce_example_chain_saturation_delay_compressor_limiter(sidechain):
    """The chain of examples/bounce.rs:104-123 on the FFI engine: TubeSaturation(0.3, 0.4, 0.5) -> DelayEffect(Eighth, 0.35,
    0.2, 8 kHz) -> TubeCompressor(-12 dB, 4:1, 10 ms, 100 ms, 0.6) -> SoftLimiter(0.95), with and without a side-chain."""
    def script(e):
        busy_pattern(e, snare_overdrive=0.15)
        e.set_bpm(128.0)
        for p, v in [(0, 0.3), (1, 0.4), (2, 0.5)]:
            e.set_global_effect_param(FX_SAT, p, v)
        for p, v in [(0, 3.0), (1, 0.35), (2, 0.2), (3, 8000.0)]:
            e.set_global_effect_param(FX_DELAY, p, v)
        for p, v in [(0, -12.0), (1, 4.0), (2, 10.0), (3, 100.0), (4, 0.6)]:
            e.set_global_effect_param(FX_COMP, p, v)
        e.set_compressor_sidechain(sidechain)
        e.set_global_effect_param(FX_LIMITER, 0, 0.95)
        for fx in (FX_SAT, FX_DELAY, FX_COMP, FX_LIMITER):
            e.set_global_effect_enabled(fx, True)
    got, want = both(script, bars=2)
    err = np.abs(got - want).max()
    print(f"bounce.rs chain, sidechain {sidechain:#x}: err {err:.3e} peak {np.abs(want).max():.3f}")
    assert err <= TOL


def test_poly_trigger_chord_equals_the_voiced_notes():
    def script_notes(e, notes):
        e.poly_trigger_notes(notes, 2, 0.8)
    # C major, degree ii (Dm7), first inversion, octave 4: D4 F4 A4 C5 -> F4 A4 C5 D5 (music/voicing.rs:96-101 on key.rs:55-84)
    o = O.oracle_engine(); g = G.Engine()
    script_notes(o, [65, 69, 72, 74])
    g.poly_trigger_chord(0, 0, 1, 1, preset=2, octave=4, velocity=0.8)
    want, got = o.render(30000), g.render(30000)
    o.close(); g.close()
    assert np.abs(want).max() > 0.01
    assert np.abs(got - want).max() <= TOL


def test_peak_meters_and_midi_export_match_the_oracle():
    """gooey_engine_get_channel_peaks / mixer_get_track_peak (read-and-reset maxima) and gooey_engine_drain_midi_events
    (ffi.rs:2572-2584, 6573-6580, 2145-2167) over several render calls, with manual triggers, an effect rack and mute."""
    def script(e):
        busy_pattern(e)
        e.set_swing(0.57)
        e.set_instrument_gain(S.SNARE, 0.6)
        e.mixer_set_track_gain(0, 1.3); e.mixer_set_track_pan(1, 0.3)
        e.track_effect_add(0, 2)                                   # saturation on the drum track
        e.sequencer_start()

    def run(e):
        script(e)
        log = []
        e.trigger_instrument_with_velocity(S.TOM, 0.7)
        e.render(9000)
        log.append((e.drain_midi_events(), e.get_channel_peaks(), [e.mixer_get_track_peak(t) for t in range(5)]))
        e.set_instrument_mute(S.HIHAT, True)
        e.trigger_instrument(S.KICK)
        e.render(12000)
        log.append((e.drain_midi_events(3), e.drain_midi_events(), e.get_channel_peaks(3), [e.mixer_get_track_peak(t) for t in range(4)]))
        e.render(700)                                              # peaks not read in between: maxima carry over
        e.render(300)
        log.append((e.drain_midi_events(), e.get_channel_peaks(), [e.mixer_get_track_peak(t) for t in range(4)]))
        e.bounce_to_buffer(1)                                      # a bounce is a run of 512-frame renders: the last chunk's events remain
        log.append((e.drain_midi_events(), e.get_channel_peaks(), [e.mixer_get_track_peak(t) for t in range(4)]))
        return log
    o = O.oracle_engine(); g = G.Engine()
    want, got = run(o), run(g)
    o.close(); g.close()
    assert len(want[0][0]) > 3 and want[0][0][0] == (S.TOM, pytest.approx(0.7), 0)

    def same(a, b):
        if len(a) == 0 or len(b) == 0:
            assert len(a) == len(b), (a, b)
        elif isinstance(a, (list, tuple)) and isinstance(a[0], tuple):               # midi events: exact
            assert [(x[0], x[2]) for x in a] == [(x[0], x[2]) for x in b]
            assert np.array_equal(np.array([x[1] for x in a], np.float32), np.array([x[1] for x in b], np.float32))
        else:
            a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
            assert a.shape == b.shape and np.abs(a - b).max() <= TOL, (a, b)
    for w, g_ in zip(want, got):
        assert len(w) == len(g_)
        for a, b in zip(w, g_):
            same(a, b)
    assert max(want[0][1]) > 0.01 and max(want[0][2]) > 0.01


def test_float_wav_writer(tmp_path):
    from libgooey_b200 import bounce as B
    import struct
    x = np.linspace(-1, 1, 2000, dtype=np.float32).reshape(1000, 2)
    B.write_wav_f32(tmp_path / "f.wav", x, 44100)
    raw = (tmp_path / "f.wav").read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and struct.unpack("<HHIIHH", raw[20:36]) == (3, 2, 44100, 44100 * 8, 8, 32)
    assert np.array_equal(np.frombuffer(raw[44:], np.float32).reshape(1000, 2), x)


def test_preset_blend_pad_and_per_step_blends():
    """gooey_engine_blend_* (ffi.rs:5245-5490) and per-step blends (:4009-4075): bilinear blend of four corner presets applied by
    set_position and at every sequencer trigger (followed by snap_params), with corner edits, a cleared step blend, a reset and an
    instrument swap while blending is on."""
    def script(e):
        busy_pattern(e, snare_overdrive=0.2)
        for inst, (x, y) in zip(range(5), [(0.2, 0.8), (0.7, 0.4), (0.5, 0.5), (0.9, 0.1), (0.35, 0.65)]):
            e.blend_enable(inst)
            e.blend_set_position(inst, x, y)
        e.blend_set_corner_preset(S.KICK, 0, 3)            # bottom-left = dirt
        e.blend_set_corner_preset(S.BASS, 3, 1)            # top-right = sub
        e.blend_set_position(S.KICK, 0.25, 0.75)
        e.blend_disable(S.TOM)
        e.blend_set_position(S.TOM, 0.0, 0.0)              # ignored while disabled
        e.sequencer_set_instrument_step_blend(S.TOM, 4, 0.1, 0.9)      # steps with their own blend fire even when the pad is off
        e.sequencer_set_instrument_step_blend(S.KICK, 8, 1.0, 0.0)
        e.sequencer_set_instrument_step_blend(S.SNARE, 4, 0.49, 0.51)  # the filter type switches at t = 0.5
        e.sequencer_set_instrument_step_blend(S.HIHAT, 2, 0.6, 0.6)
        e.sequencer_clear_instrument_step_blend(S.HIHAT, 2)
        e.sequencer_set_instrument_step_settings(S.BASS, 0, True, True, 0.9, True, 0.8, 0.2, False, 0)
    o = O.oracle_engine(); g = G.Engine()
    script(o); script(g)
    w1, g1 = o.bounce_to_buffer(1), g.bounce_to_buffer(1)
    for e in (o, g):
        e.blend_reset_corners(S.KICK)
        e.set_channel_instrument_type(3, S.KICK)           # the tom channel becomes a second kick: default corners
        e.blend_enable(3); e.blend_set_position(3, 0.6, 0.3)
        e.sequencer_set_instrument_step(3, 6, True)
    w2, g2 = o.bounce_to_buffer(1), g.bounce_to_buffer(1)
    o.close(); g.close()
    assert np.abs(w1).max() > 0.01 and np.abs(w2).max() > 0.01
    assert np.abs(g1 - w1).max() <= TOL
    assert np.abs(g2 - w2).max() <= TOL


def test_blend_getters_and_sentinels():
    c = ctypes
    g = G.Engine()
    L = g._L
    L.gooey_engine_blend_is_enabled.restype = c.c_bool; L.gooey_engine_blend_is_enabled.argtypes = [c.c_void_p, c.c_uint32]
    for name in ("gooey_engine_blend_get_position_x", "gooey_engine_blend_get_position_y"):
        getattr(L, name).restype = c.c_float; getattr(L, name).argtypes = [c.c_void_p, c.c_uint32]
    L.gooey_engine_blend_get_corner_preset.restype = c.c_uint32; L.gooey_engine_blend_get_corner_preset.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32]
    L.gooey_engine_sequencer_get_instrument_step_blend_x.restype = c.c_float
    L.gooey_engine_sequencer_get_instrument_step_blend_x.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32]
    assert not L.gooey_engine_blend_is_enabled(g._h, 0)
    assert L.gooey_engine_blend_get_position_x(g._h, 0) == 0.5 and L.gooey_engine_blend_get_position_x(g._h, 9) == -1.0
    g.blend_set_position(0, 0.9, 0.9)                       # ignored: blending is off
    assert L.gooey_engine_blend_get_position_y(g._h, 0) == 0.5
    g.blend_enable(0); g.blend_set_position(0, 2.0, -1.0)   # clamped
    assert L.gooey_engine_blend_is_enabled(g._h, 0)
    assert L.gooey_engine_blend_get_position_x(g._h, 0) == 1.0 and L.gooey_engine_blend_get_position_y(g._h, 0) == 0.0
    assert [L.gooey_engine_blend_get_corner_preset(g._h, 1, k) for k in range(4)] == [0, 1, 2, 3]
    g.blend_set_corner_preset(1, 2, 0)
    assert L.gooey_engine_blend_get_corner_preset(g._h, 1, 2) == 0 and L.gooey_engine_blend_get_corner_preset(g._h, 1, 4) == 0xFFFFFFFF
    assert L.gooey_engine_sequencer_get_instrument_step_blend_x(g._h, 0, 3) == -1.0
    g.sequencer_set_instrument_step_blend(0, 3, 0.25, 0.5)
    assert L.gooey_engine_sequencer_get_instrument_step_blend_x(g._h, 0, 3) == 0.25
    g.close()


def test_lfo_pool_routes_match_the_oracle():
    """gooey_engine_*lfo* (ffi.rs:4616-4993): tempo-synced sine LFOs writing bipolar modulation to routed channel parameters every
    frame, after the triggers and before the voices tick (:1238-1251, :322-405); unrouted LFOs keep running; removed / invalid routes."""
    def script(e):
        busy_pattern(e)
        e.set_bpm(126.0)
        e.set_lfo_enabled(0, True); e.set_lfo_timing(0, 4)
        e.add_lfo_route(0, S.KICK, 0, 0.5)                   # kick frequency
        r = e.add_lfo_route(0, S.SNARE, 10, 0.8)             # snare filter cutoff
        e.add_lfo_route(0, S.SNARE, 3, 0.3)                  # snare volume ...
        e.set_lfo_enabled(1, True); e.set_lfo_timing(1, 6); e.set_lfo_amount(1, 0.7); e.set_lfo_offset(1, 0.2)
        e.add_lfo_route(1, S.HIHAT, 3, 1.0)                  # hat tone
        e.add_lfo_route(1, S.TOM, 0, 0.9)                    # tom tune (0-100 scale)
        e.add_lfo_route(1, S.TOM, 8, 0.5)                    # tom tuning (0-1, not bipolar-mapped)
        e.add_lfo_route(1, S.BASS, 6, 0.6)                   # bass filter cutoff
        e.add_lfo_route(1, S.KICK, 5, 1.0)                   # kick pitch envelope: not modulatable, ignored
        e.add_lfo_route(1, 7, 0, 1.0)                        # no such channel, ignored
        e.set_lfo_enabled(2, True); e.set_lfo_timing(2, 7)   # enabled, unrouted: the phase still runs
        e.set_lfo_timing(3, 2); e.add_lfo_route(3, S.KICK, 6, 1.0)   # routed but disabled
        return r
    o = O.oracle_engine(); g = G.Engine()
    ro, rg = script(o), script(g)
    assert ro == rg == 1
    w1, g1 = o.bounce_to_buffer(1), g.bounce_to_buffer(1)
    assert [g.get_lfo_phase(i) for i in range(4)] == [o.get_lfo_phase(i) for i in range(4)]
    assert o.get_lfo_phase(2) > 0.0 and o.get_lfo_phase(3) == 0.0
    for e in (o, g):
        assert e.remove_lfo_route(0, 1) and not e.remove_lfo_route(0, 1)
        e.clear_lfo_routes(1)
        e.reset_lfo_phase(0)
        e.sequencer_start()
    w2, g2 = o.render(9000), g.render(9000)
    assert [g.get_lfo_phase(i) for i in range(3)] == [o.get_lfo_phase(i) for i in range(3)]
    o.close(); g.close()
    assert np.abs(w1).max() > 0.01 and np.abs(w2).max() > 0.001
    assert np.abs(g1 - w1).max() <= TOL
    assert np.abs(g2 - w2).max() <= TOL


def test_lfo_modulated_and_plain_engines_share_a_batch():
    def script(e, i):
        busy_pattern(e, snare_overdrive=0.1 * (i % 3))
        if i % 2 == 1:
            e.set_lfo_enabled(i % 8, True); e.set_lfo_timing(i % 8, 3 + i % 4)
            e.add_lfo_route(i % 8, i % 5, 0, 0.4 + 0.05 * i)
    n = 7
    es = [G.Engine() for _ in range(n)]
    for i, e in enumerate(es):
        script(e, i)
    got = G.batch_bounce(es, 1)
    [e.close() for e in es]
    wants = O.bounce_many(script, range(n), 1)
    for i in range(n):
        assert np.abs(got[i] - wants[i]).max() <= TOL, i
