"""Call scripts shared by the oracle and the product: each takes an Engine-like object (libgooey_b200.engine.Engine
bound either to libgooey_b200.so or to the oracle) and configures it through FFI-named methods only."""
import numpy as np

KICK, SNARE, HIHAT, TOM, BASS = range(5)
FX_DELAY, FX_TILT, FX_LIMITER, FX_REVERB, FX_PLATE = 1, 4, 5, 6, 9


def pattern_engine(e, seed, swing=None, notes=True, graph=True):
    """C3-style engine: 5 x 16-step Bernoulli(0.35) patterns with velocities, 10 % note overrides on
    kick/tom/bass, a 5th mixer track fed by the bass, random track gains / pans."""
    rng = np.random.default_rng(seed)
    e.set_bpm(120.0)
    if swing is not None:
        e.set_swing(float(swing))
    for inst in range(5):
        for step in range(16):
            on = bool(rng.random() < 0.35)
            vel = float(rng.uniform(0.3, 1.0))
            note = int(rng.integers(36, 61)) if (notes and inst in (KICK, TOM, BASS) and rng.random() < 0.10) else 255
            e.sequencer_set_instrument_step_settings(inst, step, on, True, vel, False, 0.0, 0.0, True, note)
    if graph:
        t = e.mixer_add_track("bass2")
        e.mixer_route_source(1, t)
        for tr in range(5):
            e.mixer_set_track_gain(tr, float(rng.uniform(0.5, 1.5)))
            e.mixer_set_track_pan(tr, float(rng.uniform(0.2, 0.8)))


def random_voice_params(e, seed):
    rng = np.random.default_rng(seed)
    for p in range(7):
        e.set_kick_param(p, float(rng.random()))
    for p in [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 16, 17, 18]:
        e.set_snare_param(p, float(rng.random()))
    e.set_snare_param(12, float(rng.integers(0, 4)))
    e.set_snare_param(15, 0.0)                     # overdrive
    for p in range(5):
        e.set_hihat_param(p, float(rng.random()))
    for p in range(8):
        e.set_tom_param(p, float(rng.random()))
    for p in range(15):
        if p != 13:
            e.set_bass_param(p, float(rng.random()))
    e.set_bass_param(13, 0.0)                      # overdrive


def fx_chain(e, seed, plate=False, spring=True, delay=True, tilt=True, limiter=False):
    """C5-style global chain."""
    rng = np.random.default_rng(seed)
    if tilt:
        e.set_global_effect_param(FX_TILT, 0, float(rng.uniform(0.2, 0.8)))
        e.set_global_effect_param(FX_TILT, 1, float(rng.uniform(0.0, 0.5)))
        e.set_global_effect_enabled(FX_TILT, True)
    if delay:
        e.set_global_effect_param(FX_DELAY, 0, float(rng.integers(2, 5)))
        e.set_global_effect_param(FX_DELAY, 1, float(rng.uniform(0.2, 0.6)))
        e.set_global_effect_param(FX_DELAY, 2, float(rng.uniform(0.1, 0.4)))
        e.set_global_effect_param(FX_DELAY, 3, float(rng.uniform(2000.0, 12000.0)))
        e.set_global_effect_enabled(FX_DELAY, True)
    if spring:
        e.set_global_effect_param(FX_REVERB, 0, float(rng.uniform(0.3, 0.7)))
        e.set_global_effect_param(FX_REVERB, 1, 0.25)
        e.set_global_effect_param(FX_REVERB, 2, 0.5)
        e.set_global_effect_enabled(FX_REVERB, True)
    if plate:
        e.set_global_effect_param(FX_PLATE, 0, float(rng.uniform(0.3, 0.7)))
        e.set_global_effect_param(FX_PLATE, 1, 0.25)
        e.set_global_effect_param(FX_PLATE, 2, 0.5)
        e.set_global_effect_enabled(FX_PLATE, True)
    if limiter:
        e.set_global_effect_param(FX_LIMITER, 0, 0.8)
        e.set_global_effect_enabled(FX_LIMITER, True)
