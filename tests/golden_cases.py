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


ENGINE_CASES = {"default_pattern": _default_pattern, "graph_swing_fx": _fx_pattern}
