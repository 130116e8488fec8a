#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/, the C++ restatement of the reference).

THESE ARE NOT REFERENCE OUTPUTS.  The Rust reference cannot be built in this image (no cargo) and stores no audio
vectors of its own; the fixtures pin the restatement (so that an accidental change of the oracle is caught on CPU)
and give the GPU tests a target that does not depend on the oracle library being loadable.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
import engine_scripts as S  # noqa: E402
from golden_cases import kit_patches, KIT_FRAMES, ENGINE_CASES, ENGINE_KEEP  # noqa: E402


def main():
    patches, vel, names = kit_patches()
    audio = O.render_voices(patches, KIT_FRAMES, triggers=[(i, 0, float(vel[i])) for i in range(len(patches))])
    np.savez_compressed(os.path.join(HERE, "preset_kit.npz"), audio=audio, velocity=vel, names=np.array(names),
                        note=np.array("oracle render (C++ restatement), not reference output"))
    for name, script in ENGINE_CASES.items():
        o = O.oracle_engine()
        script(o)
        buf = o.bounce_to_buffer(1)
        o.close()
        np.savez_compressed(os.path.join(HERE, f"engine_{name}.npz"), audio=buf[:ENGINE_KEEP], length=np.array(len(buf)),
                            note=np.array("oracle render (C++ restatement), not reference output"))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
