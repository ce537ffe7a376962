"""TEST INFRASTRUCTURE: one reference run in its own process (the reference's world is process-global).
usage: python ref_job.py MAPSIZE NCALLS CHECKPOINTS(comma) SHUFFLE(0/1) SEED OUT_PREFIX [TERRAIN.npy]
Marches the hash-spawned drops of erode(512) calls (the positions the CUDA path uses: orc_ls_spawn keyed (seed, call))
through the reference's OWN Drop::descend / World::cascade (oracle/_ref), optionally in another order per call, and
saves height / discharge (tiled pool order) at the checkpoints plus the per-call step counts."""
import ctypes as C
import sys

import numpy as np

import orc

ms, ncalls, cps, shuffle, seed, out = int(sys.argv[1]), int(sys.argv[2]), [int(v) for v in sys.argv[3].split(",")], int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
p = orc.default_params(ms)
R = orc.Ref(ms)
if len(sys.argv) > 7:
    h = np.load(sys.argv[7])
else:
    h = orc.init_terrain(ms, 1)
orc.lib().orc_fill_tiled_from_planar(C.byref(p), np.ascontiguousarray(h, np.float32).ctypes.data, R.cells.ctypes.data)
del h
steps, spawned = [], []
for c in range(ncalls):
    xy = np.zeros((ms * ms * 512, 2), np.float32)
    orc.lib().orc_ls_spawn(C.byref(p), seed, c, 512, xy.ctypes.data)
    if shuffle:
        xy = xy[np.random.default_rng(1000 + c).permutation(len(xy))]
    st = R.erode_spawnlist(xy)
    steps.append(st["steps"])
    spawned.append(st["spawned"])
    if c + 1 in cps:
        np.save(f"{out}_h{c + 1}.npy", R.cells["height"].copy())
        np.save(f"{out}_d{c + 1}.npy", R.cells["discharge"].copy())
np.save(f"{out}_steps.npy", np.array([steps, spawned], np.int64))
print("done")
