"""development aid: does the lock-step schedule stay stable over many calls?  Prints the height range of a world every
`every` erode(512) calls (a runaway shows as heights leaving [0, 1])."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import simplehydrology_b200 as shx

ms, calls, every = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tseed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
with shx.World(mapsize=ms) as W:
    W.synth_terrain(tseed)
    t0 = time.time()
    for c in range(calls):
        try:
            st = W.erode(512, 1)
        except shx.ShxError as e:
            print("call", c + 1, "error:", e)
            break
        if (c + 1) % every == 0:
            h = W.download_height_q()[..., 0].astype(np.float64) / 2 ** 26
            m = W.view_maps_download()
            print(f"call {c+1}: h[{h.min():.3f},{h.max():.3f}] mean {h.mean():.5f} max discharge alpha {m[:,0].max():.3f} steps/drop {st.steps/max(st.spawned,1):.0f} ({time.time()-t0:.0f}s)", flush=True)
