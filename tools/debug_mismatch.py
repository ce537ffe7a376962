"""Development aid: locate the first drop whose final record differs between the CUDA batched path
and the lock-step oracle on a small crowded world."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
import simplehydrology_b200 as shx  # noqa: E402
from test_gpu_batched import small_world  # noqa: E402


def run(tilesize, mapsize, n, seed, **kw):
    p, cells = small_world(tilesize, mapsize, seed=tilesize + mapsize)
    size = tilesize * mapsize
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-1.5, size + 1.0, size=(n, 2)).astype(np.float32)
    W = shx.World(params=shx.Params.from_buffer_copy(bytes(p)), max_drops=4096, **kw)
    W.upload(cells)
    ls = orc.Ls(p)
    ls.upload(cells)
    # cycle 0 through the full call (EMA included), then the drops of cycle 1 record by record
    st0 = W.erode_spawnlist(xy)
    so0 = ls.erode_spawnlist(xy)
    h0, h1, f, t = W.download_raw()
    print("   cycle 0: steps", st0.steps, so0.steps, "planes", np.array_equal(h0, ls.height_q(0)), "fields",
          np.array_equal(f.view(np.uint32), ls.field().view(np.uint32)), "max discharge", float(f[..., 0].max()),
          "max |mom|", float(np.abs(f[..., 1:3]).max()))
    if not np.array_equal(f.view(np.uint32), ls.field().view(np.uint32)):
        bad = np.argwhere(f.view(np.uint32) != ls.field().view(np.uint32))
        print("   field mismatches", len(bad), bad[:4], f[tuple(bad[0][:2])], ls.field()[tuple(bad[0][:2])], t[tuple(bad[0][:2])])
    xy = rng.uniform(-1.5, size + 1.0, size=(n, 2)).astype(np.float32)
    drops, _ = ls.make_drops(xy)
    mine = drops.copy().view(shx.DROP_DTYPE)
    st = W.run_drops(mine)
    so, _ = ls.run_drops(drops)
    same = mine.view(np.uint8).reshape(n, 32) == drops.view(np.uint8).reshape(n, 32)
    bad = np.nonzero(~same.all(axis=1))[0]
    h0, h1, f, t = W.download_raw()
    print(f"ts {tilesize} ms {mapsize} n {n} {kw}: steps gpu {st.steps} orc {so.steps} phases {st.phases}/{so.phases} "
          f"bad drops {len(bad)} planes equal {np.array_equal(h0, ls.height_q(0))} {np.array_equal(h1, ls.height_q(1))} "
          f"tracks {np.array_equal(t[..., :3], ls.track_q()[..., :3])}", flush=True)
    for i in bad[:5]:
        print("   drop", i, "gpu", mine[i], "orc", drops[i])
    W.close()


if __name__ == "__main__":
    run(32, 4, 2000, 2000)
    run(32, 4, 1000, 2000)
    run(32, 4, 300, 2000)
    run(128, 1, 2000, 2000)
