import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from libgooey_b200 import engine as G
import oracle_lib as O, engine_scripts as S
import torch

def script(e, i):
    S.random_voice_params(e, 1000 + i)
    S.pattern_engine(e, 2000 + i, swing=None if i % 2 == 0 else 0.4 + 0.3 * ((i * 37) % 100) / 100.0)
    S.fx_chain(e, 3000 + i, plate=False)
n, bars = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 2
engines = [G.Engine() for _ in range(n)]
for i, e in enumerate(engines): script(e, i)
frames = 176400
out = torch.empty((n, frames), dtype=torch.float32, device="cuda:0")
G.batch_bounce_device(engines, bars, out.data_ptr(), frames)
first = out.cpu().numpy()
second = G.batch_bounce(engines, bars)
for i in range(min(n, 4)):
    o = O.oracle_engine(); script(o, i); w1 = o.bounce_to_buffer(bars); w2 = o.bounce_to_buffer(bars); o.close()
    d1 = np.abs(first[i] - w1); d2 = np.abs(second[i] - w2)
    print(i, "first", d1.max(), int(d1.argmax()), "second", d2.max(), int(d2.argmax()), "peak", np.abs(w1).max(), np.abs(w2).max(), "finite", np.isfinite(w1).all(), np.isfinite(w2).all())
