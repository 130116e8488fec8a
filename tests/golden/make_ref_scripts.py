#!/usr/bin/env python
"""Writes the call scripts that rust/examples/dump_golden.rs replays inside a libgooey checkout
(tests/golden/ref/scripts/*.calls: one FFI call per line; sweep64.voices: one voice per line).

The scripts are the SAME Python functions the parity tests run against the oracle and the product
(tests/golden_cases.py, tests/engine_scripts.py, tests/workloads.py), recorded through a proxy, so the reference
vectors dump_golden produces are comparable sample for sample.  Run from the repo root:

    python tests/golden/make_ref_scripts.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

OUT = os.path.join(HERE, "ref", "scripts")


class Recorder:
    """Engine-like object that records FFI-named calls instead of making them."""

    def __init__(self):
        self.lines = []
        self._tracks = 4

    def mixer_add_track(self, name="track"):
        self.lines.append(f"mixer_add_track {name}")
        self._tracks += 1
        return self._tracks - 1

    def sampler_register(self):
        self.lines.append("sampler_register")
        return 0

    def mixer_route_source(self, source, track):
        self.lines.append(f"mixer_route_source {int(source)} {int(track)}")
        return True

    def __getattr__(self, name):
        def call(*args):
            out = [name]
            for a in args:
                if isinstance(a, bool):
                    out.append("1" if a else "0")
                elif isinstance(a, int):
                    out.append(str(a))
                else:
                    out.append(repr(float(a)))
            self.lines.append(" ".join(out))
        return call


def main():
    import golden_cases as GC
    import engine_scripts as S
    from workloads import drum_sweep_raw
    os.makedirs(OUT, exist_ok=True)
    cases = dict(GC.ENGINE_CASES)
    cases.update(GC.REF_CASES)
    cases["c3_style_7"] = lambda e: (S.random_voice_params(e, 1007), S.pattern_engine(e, 2007, swing=0.61))
    for name, script in cases.items():
        r = Recorder()
        script(r)
        with open(os.path.join(OUT, f"{name}.calls"), "w") as f:
            f.write(f"# {name}: replayed by rust/examples/dump_golden.rs on gooey_engine_new(44100), then bounced\nbounce 1\n")
            f.write("\n".join(r.lines) + "\n")
    patches, vel, _ = drum_sweep_raw(64, seed=0x600E7)
    with open(os.path.join(OUT, "sweep64.voices"), "w") as f:
        f.write("# instrument aux velocity p0..p23 (tests/workloads.py drum_sweep_raw(64, seed=0x600E7)); 8192 frames each, trigger at 0\n")
        for (inst, aux, params), v in zip(patches, vel):
            p = list(params) + [0.0] * (24 - len(params))
            f.write(" ".join([str(inst), str(aux), repr(float(v))] + [repr(float(x)) for x in p]) + "\n")
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
