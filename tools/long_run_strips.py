"""development aid: long-run stability of k row strips exchanging once per call (emulated on one device)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from simplehydrology_b200 import strips

ms, k, calls, every = (int(v) for v in sys.argv[1:5])
bs = [strips.GpuStrip(ms, r, k, 0) for r in range(k)]
for b in bs:
    b.W.synth_terrain(1)
S = strips.LocalStripSet(bs)
t0 = time.time()
for c in range(calls):
    S.erode_cycle(512, 1)
    if (c + 1) % every == 0:
        hs = []
        for b in bs:
            xlo, _ = b.W.stored_rows()
            hs.append(b.W.download_height_q()[b.row0 - xlo:b.row1 - xlo, :, 0])
        h = np.concatenate(hs).astype(np.float64) / 2 ** 26
        print(f"call {c+1}: h[{h.min():.3f},{h.max():.3f}] mean {h.mean():.5f} in flight {S.in_flight()} ({time.time()-t0:.0f}s)", flush=True)
