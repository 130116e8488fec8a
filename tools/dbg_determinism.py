"""Run-to-run determinism of a C5-style batch: two independent constructions of the same engines must bounce to identical bits,
device-resident and through the pitched host block (catches races between pieces, streams and the drain)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from libgooey_b200 import engine as G, HostBuffer
import engine_scripts as S
import torch

def script(e, i):
    S.random_voice_params(e, 1000 + i)
    S.pattern_engine(e, 2000 + i, swing=None if i % 2 == 0 else 0.4 + 0.3 * ((i * 37) % 100) / 100.0)
    S.fx_chain(e, 3000 + i, plate=False)

n, bars, frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 2, 176400

def run(host):
    engines = [G.Engine() for _ in range(n)]
    for i, e in enumerate(engines): script(e, i)
    outs = []
    for rep in range(2):
        if host:
            hb = HostBuffer(n * frames * 4)
            o = hb.array((n, frames), np.float32)
            G.batch_bounce_host(engines, bars, out=o)
            outs.append(o.copy()); del o; hb.close()
        else:
            out = torch.empty((n, frames), dtype=torch.float32, device="cuda:0")
            G.batch_bounce_device(engines, bars, out.data_ptr(), frames)
            outs.append(out.cpu().numpy()); del out
    for e in engines: e.close()
    return outs

a = run(False); b = run(False); c = run(True)
for rep in range(2):
    for name, x, y in [("dev vs dev", a[rep], b[rep]), ("dev vs host", a[rep], c[rep])]:
        same = (x.view(np.uint32) == y.view(np.uint32))
        bad = np.nonzero(~same.all(axis=1))[0]
        print(f"bounce {rep + 1} {name}: {len(bad)} engines differ", bad[:10], (np.argmax(~same[bad[0]]) if len(bad) else ""))
