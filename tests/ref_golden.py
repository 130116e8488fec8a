"""Access to the reference vectors of tests/golden/ref (written by rust/examples/dump_golden.rs inside a libgooey
checkout).  Absent files mean the oracle is unpinned for that case: the tests skip with that reason."""
import os

import numpy as np
import pytest

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref")
SCRIPTS = os.path.join(REF, "scripts")
UNPINNED = "parity unpinned: no reference vector {} (see tests/golden/ref/README.md)"


def load(name, dtype=np.float32):
    p = os.path.join(REF, name)
    if not os.path.exists(p):
        pytest.skip(UNPINNED.format(name))
    return np.fromfile(p, dtype=np.dtype(dtype).newbyteorder("<"))


def replay(engine, name):
    """Runs tests/golden/ref/scripts/<name>.calls on an Engine-like object; returns the bars to bounce."""
    bars = 1
    with open(os.path.join(SCRIPTS, name + ".calls")) as f:
        for line in f:
            w = line.split()
            if not w or w[0].startswith("#"):
                continue
            if w[0] == "bounce":
                bars = int(w[1]); continue
            if w[0] == "mixer_add_track":
                engine.mixer_add_track(w[1]); continue
            args = []
            for a in w[1:]:
                args.append(float(a) if ("." in a or "e" in a or "inf" in a or "nan" in a) else int(a))
            getattr(engine, w[0])(*_typed(w[0], args))
    return bars


_BOOL_ARGS = {"sequencer_set_instrument_step": (2,), "sequencer_set_instrument_step_settings": (2, 3, 5, 8), "set_global_effect_enabled": (1,),
              "loop_set_playing": (1,), "loop_set_mute": (1,), "loop_set_solo": (1,), "sampler_set_step": (2,)}


def _typed(name, args):
    for i in _BOOL_ARGS.get(name, ()):
        args[i] = bool(args[i])
    return args


def sweep_voices():
    """(raw patches, velocities) of scripts/sweep64.voices."""
    patches, vel = [], []
    with open(os.path.join(SCRIPTS, "sweep64.voices")) as f:
        for line in f:
            w = line.split()
            if not w or w[0].startswith("#"):
                continue
            patches.append((int(w[0]), int(w[1]), [float(x) for x in w[3:]]))
            vel.append(float(w[2]))
    return patches, np.array(vel, np.float32)
