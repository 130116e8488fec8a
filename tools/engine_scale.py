#!/usr/bin/env python
"""Engine-level configurations of BASELINE.json at scale (C3: 1024 engines x 8 bars with 16-step patterns + mixer graph;
C5-style: + global delay / spring / plate / tilt chain), with spot-check parity against the oracle.

    python tools/engine_scale.py --engines 1024 --bars 8 [--fx] [--check 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engines", type=int, default=1024)
    ap.add_argument("--bars", type=int, default=8)
    ap.add_argument("--fx", action="store_true")
    ap.add_argument("--plate", action="store_true")
    ap.add_argument("--check", type=int, default=3)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    from libgooey_b200 import engine as G, lib
    import engine_scripts as S
    import oracle_lib as O

    def script(e, i):
        S.random_voice_params(e, 1000 + i)
        S.pattern_engine(e, 2000 + i, swing=None if i % 2 == 0 else 0.4 + 0.3 * ((i * 37) % 100) / 100.0)
        if args.fx:
            S.fx_chain(e, 3000 + i, plate=args.plate)

    t0 = time.perf_counter()
    engines = [G.Engine() for _ in range(args.engines)]
    for i, e in enumerate(engines):
        script(e, i)
    t_setup = time.perf_counter() - t0
    frames = int(round(args.bars * 4 * 0.5 * 44100.0))
    stride = (frames + 3) & ~3
    out = torch.empty((args.engines, stride), dtype=torch.float32, device="cuda:0")
    times = []
    for r in range(args.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got_frames = G.batch_bounce_device(engines, args.bars, out.data_ptr(), stride)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        assert got_frames == frames
    host = out[:, :frames].cpu().numpy()
    res = {"engines": args.engines, "bars": args.bars, "frames": frames, "fx": args.fx, "plate": args.plate, "setup_s": round(t_setup, 2),
           "bounce_s": [round(t, 3) for t in times], "kernel_ms_last": lib().gooey_b200_last_kernel_ms(),
           "engine_samples_per_s": args.engines * frames / min(times), "voice_samples_per_s": 5 * args.engines * frames / min(times),
           "peak": float(np.nanmax(np.abs(np.where(np.isfinite(host), host, 0.0)))), "non_finite_engines": int((~np.isfinite(host)).any(axis=1).sum())}
    bad = np.nonzero((~np.isfinite(host)).any(axis=1))[0]
    res["non_finite_first"] = [(int(i), int(np.argmax(~np.isfinite(host[i])))) for i in bad[:12]]
    # parity: engines bounced TWICE on the GPU (reps) -> compare the last bounce with an oracle engine bounced reps times
    errs = []
    for i in np.linspace(0, args.engines - 1, args.check).astype(int):
        o = O.oracle_engine()
        script(o, int(i))
        for r in range(args.reps):
            want = o.bounce_to_buffer(args.bars)
        o.close()
        # the reference itself overflows for a few random snare patches (Chamberlin SVF, high cutoff x low resonance):
        # non-finite frames must coincide, finite ones are compared relative to max(1, |want|)
        fin = np.isfinite(want)
        assert np.array_equal(fin, np.isfinite(host[i])), "non-finite frames differ from the oracle"
        errs.append(float((np.abs(host[i][fin] - want[fin]) / np.maximum(1.0, np.abs(want[fin]))).max()))
    res["max_abs_err_vs_oracle"] = errs
    print(json.dumps(res))
    for e in engines:
        e.close()


if __name__ == "__main__":
    main()
