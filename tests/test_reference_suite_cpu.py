"""The reference's own integration tests for the bounce / render path, restated against the oracle (oracle/*.hpp behind the
FFI-named Python Engine): every test below cites the Rust test it restates (tests/*.rs of the reference) and asserts what that
test asserts.  The oracle is the checker of every GPU parity test, so the reference's behavioural contract holding for the oracle
is part of what pins it; the reference itself cannot be built here (no Rust toolchain).

Left out: tests that need UI-only getters (`get_instrument_mute`, `get_track_gain`, track names ...), the loop mixer / clip grid /
sampler rack (not built) and the DSL / performance recorder."""
import numpy as np
import pytest

import oracle_lib as O

KICK, SNARE, HIHAT, TOM, BASS = range(5)
FX_LOWPASS, FX_DELAY, FX_SATURATION, FX_COMPRESSOR, FX_TILT, FX_LIMITER, FX_REVERB, FX_WAVESHAPER, FX_FBWS, FX_PLATE = range(10)
KICK_PARAM_VOLUME, SNARE_PARAM_VOLUME, HIHAT_PARAM_VOLUME, TOM_PARAM_VOLUME = 6, 3, 4, 7


@pytest.fixture
def engine():
    made = []

    def make(sr=44100.0):
        e = O.oracle_engine(sr)
        made.append(e)
        return e
    yield make
    for e in made:
        e.close()


def audible(buf, th=0.001):
    return bool((np.abs(buf) > th).any())


def render_n(e, n, frames=1024):
    buf = None
    for _ in range(n):
        buf = e.render(frames)
    return buf


# ------------------------------------------------------------------------------------------------ tests/mute_solo.rs
def test_mute_silences_instrument(engine):                       # mute_solo.rs:24-48
    e = engine()
    e.trigger_instrument(KICK)
    before = e.render(1024)
    e.set_instrument_mute(KICK, True)
    e.trigger_instrument(KICK)
    after = render_n(e, 10)
    assert audible(before) and not audible(after)


def test_solo_isolates_instrument(engine):                       # mute_solo.rs:50-72
    e = engine()
    e.set_instrument_solo(KICK, True)
    e.trigger_instrument(SNARE)
    assert not audible(render_n(e, 10))


def test_solo_allows_soloed_instrument(engine):                  # mute_solo.rs:74-90
    e = engine()
    e.set_instrument_solo(KICK, True)
    e.trigger_instrument(KICK)
    assert audible(e.render(1024))


def test_solo_overrides_mute(engine):                            # mute_solo.rs:92-110
    e = engine()
    e.set_instrument_mute(KICK, True)
    e.set_instrument_solo(KICK, True)
    e.trigger_instrument(KICK)
    assert audible(e.render(1024))


def test_multiple_solos(engine):                                 # mute_solo.rs:112-140
    e = engine()
    e.set_instrument_solo(KICK, True)
    e.set_instrument_solo(SNARE, True)
    e.trigger_instrument(KICK)
    e.trigger_instrument(SNARE)
    assert audible(e.render(1024))


def test_unmute_restores_audio(engine):                          # mute_solo.rs:142-160
    e = engine()
    e.set_instrument_mute(KICK, True)
    e.set_instrument_mute(KICK, False)
    e.trigger_instrument(KICK)
    assert audible(e.render(1024))


def test_unsolo_restores_other_instruments(engine):              # mute_solo.rs:162-183
    e = engine()
    e.set_instrument_solo(KICK, True)
    e.set_instrument_solo(KICK, False)
    e.trigger_instrument(SNARE)
    assert audible(e.render(1024))


def test_invalid_instrument_ids_are_ignored(engine):             # mute_solo.rs:185-199, instrument_gain.rs:150-158
    e = engine()
    e.set_instrument_mute(99, True)
    e.set_instrument_solo(99, True)
    e.set_instrument_gain(99, 0.5)
    e.trigger_instrument(KICK)
    assert audible(e.render(1024))


# ------------------------------------------------------------------------------------------------ tests/instrument_gain.rs
def test_gain_zero_silences_instrument(engine):                  # instrument_gain.rs:28-52
    e = engine()
    e.trigger_instrument(KICK)
    before = e.render(1024)
    e.set_instrument_gain(KICK, 0.0)
    e.trigger_instrument(KICK)
    after = render_n(e, 10)
    assert audible(before) and not audible(after)


def test_gain_reduces_level(engine):                             # instrument_gain.rs:54-86
    full = engine()
    full.trigger_instrument(KICK)
    peak_full = float(np.abs(full.render(1024)).max())
    half = engine()
    half.set_instrument_gain(KICK, 0.5)
    render_n(half, 10)                                           # let the gain smoother settle
    half.trigger_instrument(KICK)
    peak_half = float(np.abs(half.render(1024)).max())
    assert peak_full > 0.001 and peak_half > 0.001 and peak_half < peak_full


def test_gain_with_mute(engine):                                 # instrument_gain.rs:110-128
    e = engine()
    e.set_instrument_gain(KICK, 1.0)
    e.set_instrument_mute(KICK, True)
    e.trigger_instrument(KICK)
    assert not audible(render_n(e, 10))


# ------------------------------------------------------------------------------------------------ tests/volume_zero_mute.rs
VOLUME = [(KICK, "set_kick_param", KICK_PARAM_VOLUME), (SNARE, "set_snare_param", SNARE_PARAM_VOLUME),
          (HIHAT, "set_hihat_param", HIHAT_PARAM_VOLUME), (TOM, "set_tom_param", TOM_PARAM_VOLUME)]


@pytest.mark.parametrize("inst,setter,param", VOLUME)
def test_volume_zero_silences(engine, inst, setter, param):      # volume_zero_mute.rs:3-35, 84-107
    e = engine()
    e.trigger_instrument(inst)
    assert audible(e.render(1024))
    getattr(e, setter)(param, 0.0)
    render_n(e, 10)
    e.trigger_instrument(inst)
    for _ in range(5):
        assert float(np.abs(e.render(1024)).max()) < 1e-6


@pytest.mark.parametrize("inst,setter,param", VOLUME)
def test_volume_zero_mid_playback(engine, inst, setter, param):  # volume_zero_mute.rs:37-64, 108-131
    e = engine()
    e.trigger_instrument(inst)
    assert audible(e.render(1024))
    getattr(e, setter)(param, 0.0)
    render_n(e, 10)
    assert float(np.abs(e.render(1024)).max()) < 1e-6


# ------------------------------------------------------------------------------------------------ tests/panning.rs (through the FFI strip)
def kick_energy(engine, pan):                                    # panning.rs:6-30: settle 8192 frames, trigger, 4096 frames of energy
    e = engine()
    e.set_instrument_pan(KICK, pan)
    e.render(8192)
    e.trigger_instrument(KICK)
    buf = e.render(4096).astype(np.float64)
    return float((buf[:, 0] ** 2).sum()), float((buf[:, 1] ** 2).sum())


def test_hard_left_pan_favors_the_left_channel(engine):          # panning.rs:52-60
    left, right = kick_energy(engine, 0.0)
    assert left > 0.0 and right < left * 1e-6


def test_hard_right_pan_favors_the_right_channel(engine):        # panning.rs:62-70
    left, right = kick_energy(engine, 1.0)
    assert right > 0.0 and left < right * 1e-6


def test_center_pan_is_balanced(engine):                         # panning.rs:72-80
    left, right = kick_energy(engine, 0.5)
    assert left > 0.0 and right > 0.0 and abs(left - right) < left * 1e-6


# ------------------------------------------------------------------------------------------------ tests/effect_order.rs
def test_reordering_changes_audio_output(engine):                # effect_order.rs:245-318
    def render_with_order(order):
        e = engine()
        e.set_global_effect_enabled(FX_SATURATION, True)
        e.set_global_effect_enabled(FX_DELAY, True)
        e.set_global_effect_enabled(FX_LIMITER, False)
        e.set_global_effect_param(FX_SATURATION, 0, 1.0)
        e.set_global_effect_param(FX_SATURATION, 2, 1.0)
        e.set_global_effect_param(FX_DELAY, 1, 0.7)
        e.set_global_effect_param(FX_DELAY, 2, 0.8)
        assert e.set_effect_order(order)
        e.trigger_instrument(KICK)
        return e.render(16384)
    a = render_with_order([FX_WAVESHAPER, FX_SATURATION, FX_DELAY, FX_LOWPASS, FX_TILT, FX_COMPRESSOR, FX_FBWS, FX_REVERB, FX_PLATE])
    b = render_with_order([FX_WAVESHAPER, FX_DELAY, FX_SATURATION, FX_LOWPASS, FX_TILT, FX_COMPRESSOR, FX_FBWS, FX_REVERB, FX_PLATE])
    assert float(np.abs(a - b).max()) > 1e-4


def test_order_rejects_limiter_duplicates_wrong_length_and_unknown_ids(engine):   # effect_order.rs:136-241
    e = engine()
    default = [FX_WAVESHAPER, FX_SATURATION, FX_LOWPASS, FX_TILT, FX_DELAY, FX_COMPRESSOR, FX_FBWS, FX_REVERB, FX_PLATE]
    assert e.set_effect_order(default)
    assert not e.set_effect_order([FX_LIMITER] + default[1:])            # the limiter is pinned to the end of the chain
    assert not e.set_effect_order([FX_DELAY, FX_DELAY] + default[2:])   # duplicates
    assert not e.set_effect_order(default[:-1])                          # wrong length
    assert not e.set_effect_order([42] + default[1:])                    # unknown id
    assert not e.move_effect(FX_LIMITER, 0)
    assert e.move_effect(FX_PLATE, 0)


# ------------------------------------------------------------------------------------------------ tests/mixer_graph.rs
def test_track_peaks_report_and_reset_after_render(engine):      # mixer_graph.rs:380-392
    e = engine()
    e.trigger_instrument(KICK)
    assert float(np.abs(e.render(2048)).max()) > 1e-4
    assert e.mixer_get_track_peak(0) > 1e-4
    assert e.mixer_get_track_peak(0) == 0.0                      # read-and-reset
    assert e.mixer_get_track_peak(99) == 0.0


@pytest.mark.parametrize("edit", ["gain", "mute", "solo_other"])
def test_offline_bounce_snaps_recent_mixer_strip_changes(engine, edit):   # mixer_graph.rs:394-421
    e = engine()
    e.sequencer_set_step(0, True)
    if edit == "gain":
        e.mixer_set_track_gain(0, 0.0)
    elif edit == "mute":
        e.mixer_set_track_mute(0, True)
    else:
        e.mixer_set_track_solo(1, True)
    out = e.bounce_to_buffer(1)
    assert len(out) == 88200
    assert float(np.abs(out).max()) < 1e-6                       # silent from sample zero: the bounce snaps the strip smoothers


def test_track_effect_changes_only_audio_routed_to_that_track(engine):    # mixer_graph.rs:340-378, with the bass as the bright source
    def tail_peak(e):
        e.trigger_instrument(BASS)
        buf = e.render(8192)
        return float(np.abs(buf[4096:]).max())
    e = engine()
    for p, v in [(6, 1.0), (8, 0.0), (11, 1.0), (2, 1.0)]:      # bass: filter wide open, long decay, oscillator up: a bright sustained source
        e.set_bass_param(p, v)
    render_n(e, 10)
    dry = tail_peak(e)
    assert dry > 1e-3
    assert e.track_effect_add(1, FX_LOWPASS) == 0                # the bass feeds track 1 by default (SOURCE_BASS)
    e.track_effect_set_param(1, 0, 0, 60.0)
    tail_peak(e)
    filtered = tail_peak(e)
    assert filtered < dry * 0.6
    t = e.mixer_add_track("Dry Bass")
    assert e.mixer_route_source(1, t)
    rerouted = tail_peak(e)
    assert rerouted > filtered * 1.5                             # the re-routed source bypasses the old track's rack
    assert not e.mixer_route_source(5, 0) and not e.mixer_route_source(0, 99)    # mixer_graph.rs:214-215: SOURCE_COUNT / a missing track


# ------------------------------------------------------------------------------------------------ tests/channel_instrument_swap.rs
def test_swap_produces_audio(engine):                            # channel_instrument_swap.rs:66-77
    e = engine()
    e.set_channel_instrument_type(0, TOM)
    e.trigger_instrument(0)
    assert audible(e.render(1024))


def test_set_channel_param_after_swap(engine):                   # channel_instrument_swap.rs:125-139
    e = engine()
    e.set_channel_instrument_type(0, TOM)
    e.set_channel_param(0, 0, 0.8)
    e.trigger_instrument(0)
    assert audible(e.render(1024))


def test_channel_state_preserved_on_swap(engine):                # channel_instrument_swap.rs:79-90: the strip survives the swap
    e = engine()
    e.set_instrument_mute(0, True)
    e.set_channel_instrument_type(0, HIHAT)
    e.trigger_instrument(0)
    assert not audible(render_n(e, 10))


def test_duplicate_instrument_types(engine):                     # channel_instrument_swap.rs:157-176
    e = engine()
    e.set_channel_instrument_type(2, KICK)
    e.trigger_instrument(0)
    e.trigger_instrument(2)
    both = e.render(1024)
    one = engine()
    one.trigger_instrument(0)
    assert audible(both) and float(np.abs(both).max()) > float(np.abs(one.render(1024)).max())


def test_swap_invalid_arguments_are_ignored(engine):             # channel_instrument_swap.rs:178-201
    e = engine()
    e.set_channel_instrument_type(99, KICK)
    e.set_channel_instrument_type(0, 99)
    e.trigger_instrument(99)
    e.trigger_instrument_with_velocity(99, 1.0)
    e.set_channel_param(99, 0, 0.5)
    e.trigger_instrument(0)
    ref = engine()
    ref.trigger_instrument(0)
    assert np.array_equal(e.render(2048), ref.render(2048))      # channel 0 is still the default kick


# ------------------------------------------------------------------------------------------------ tests/sequencer_triggers_enabled.rs
def test_disabling_triggers_mutes_sequencer_but_keeps_clock_and_host_input(engine):   # sequencer_triggers_enabled.rs:44-143
    sr = 48000
    e = engine(float(sr))
    assert e.get_sequencer_triggers_enabled()                    # default: enabled
    e.sequencer_set_step(0, True)
    e.sequencer_start()
    buf = e.render(sr)
    assert any(ch == 0 for ch, _, _ in e.drain_midi_events())
    assert float(np.abs(buf).max()) > 0.01
    e.set_sequencer_triggers_enabled(False)
    assert not e.get_sequencer_triggers_enabled()
    e.render(256)
    e.drain_midi_events()
    e.render(sr)                                                 # the sounding kick rings out
    assert e.drain_midi_events() == []
    muted = e.render(sr)
    assert e.drain_midi_events() == []
    assert float(np.abs(muted).max()) < 0.001
    e.trigger_instrument_with_velocity(0, 1.0)                   # host triggers still sound
    assert float(np.abs(e.render(sr // 4)).max()) > 0.01
    e.drain_midi_events()
    e.set_sequencer_triggers_enabled(True)
    resumed = e.render(sr)
    assert any(ch == 0 for ch, _, _ in e.drain_midi_events())    # the clock kept running: the pattern comes back
    assert float(np.abs(resumed).max()) > 0.01


# ------------------------------------------------------------------------------------------------ tests/stereo_effects.rs
@pytest.mark.parametrize("fx,params", [(FX_LOWPASS, [(0, 2000.0), (1, 0.5)]), (FX_DELAY, [(1, 0.6), (2, 0.7), (4, 0.0)]),
                                       (FX_SATURATION, [(0, 0.8), (2, 1.0)]), (FX_COMPRESSOR, [(0, -20.0), (1, 8.0), (4, 1.0)]),
                                       (FX_TILT, [(0, 0.2), (1, 0.5)])])
def test_each_effect_keeps_left_equal_right_for_mono_input(engine, fx, params):      # stereo_effects.rs:36-66
    e = engine()
    e.set_global_effect_enabled(fx, True)
    for p, v in params:
        e.set_global_effect_param(fx, p, v)
    e.trigger_instrument(KICK)
    buf = e.render(4096)
    assert float(np.abs(buf).max()) > 0.001
    assert np.array_equal(buf[:, 0], buf[:, 1])


def test_ping_pong_delay_makes_left_and_right_diverge(engine):   # stereo_effects.rs:68-98
    e = engine()
    e.set_global_effect_enabled(FX_DELAY, True)
    for p, v in [(0, 4.0), (1, 0.85), (2, 1.0), (4, 1.0)]:
        e.set_global_effect_param(FX_DELAY, p, v)
    e.trigger_instrument(KICK)
    buf = e.render(32768)
    assert np.isfinite(buf).all() and float(np.abs(buf).max()) < 100.0
    assert float(np.abs(buf[:, 0] - buf[:, 1]).max()) > 1e-4


@pytest.mark.parametrize("fx", [FX_REVERB, FX_PLATE])
def test_reverbs_decorrelate_left_and_right(engine, fx):         # stereo_effects.rs:100-168
    e = engine()
    e.set_global_effect_enabled(fx, True)
    e.set_global_effect_param(fx, 0, 0.7)
    e.set_global_effect_param(fx, 1, 0.8)
    e.trigger_instrument(KICK)
    buf = e.render(32768)
    assert float(np.abs(buf).max()) > 0.001
    assert np.isfinite(buf).all() and float(np.abs(buf).max()) < 100.0
    assert float(np.abs(buf[:, 0] - buf[:, 1]).max()) > 1e-4


def test_ping_pong_off_keeps_delay_dual_mono(engine):            # stereo_effects.rs:170-197
    e = engine()
    e.set_global_effect_enabled(FX_DELAY, True)
    for p, v in [(1, 0.85), (2, 1.0), (4, 0.0)]:
        e.set_global_effect_param(FX_DELAY, p, v)
    e.trigger_instrument(KICK)
    buf = e.render(8192)
    assert np.array_equal(buf[:, 0], buf[:, 1])


# ------------------------------------------------------------------------------------------------ tests/lfo_modulation.rs (audible part, through the FFI LFO pool)
def lfo_render(engine, inst, param, offset):
    """lfo_modulation.rs:149-178: a constant modulation value of -1 / +1 applied every sample — here LFO 0 with amount 0 and that offset,
    routed with depth 1 (ffi.rs:1238-1251) — 2048 frames of settling, a trigger, 4096 frames."""
    e = engine()
    e.set_lfo_amount(0, 0.0)
    e.set_lfo_offset(0, offset)
    e.set_lfo_enabled(0, True)
    assert e.add_lfo_route(0, inst, param, 1.0) != 0xFFFFFFFF
    e.render(2048)
    e.trigger_instrument(inst)
    return e.render(4096)[:, 0]


@pytest.mark.parametrize("inst,param,what", [(KICK, 0, "kick frequency"), (KICK, 4, "kick oscillator decay"), (SNARE, 1, "snare decay"),
                                             (HIHAT, 2, "hi-hat attack"), (HIHAT, 1, "hi-hat decay")])
def test_lfo_on_a_voice_parameter_changes_the_output(engine, inst, param, what):     # lfo_modulation.rs:186-262
    low, high = lfo_render(engine, inst, param, -1.0), lfo_render(engine, inst, param, 1.0)
    assert float(np.abs(low - high).mean()) > 1e-3, what


def test_lfo_route_limits(engine):                               # ffi.rs:4868-4877: eight LFOs, sixteen routes each
    e = engine()
    assert e.add_lfo_route(8, KICK, 0, 1.0) == 0xFFFFFFFF
    ids = [e.add_lfo_route(0, KICK, 0, 0.1) for _ in range(17)]
    assert ids[:16] == list(range(16)) and ids[16] == 0xFFFFFFFF
