"""GPU, batched mode (the product path): bit-exact against the CPU lock-step oracle (orc_ls_*), which
applies the reference's per-step arithmetic under the same schedule.  Integer state (Q5.26 heights,
Q13.18 tracks), fp32 fields and every counter must match exactly, for any launch shape."""
import numpy as np
import pytest

import orc
import simplehydrology_b200 as shx

pytestmark = pytest.mark.gpu

STAT_KEYS = ["spawned", "rejected", "steps", "term_age", "term_vol", "term_oob", "cascade_transfers", "phases",
             "fx_eroded", "fx_deposited", "fx_sed_oob_lost", "fx_sed_deposited", "fx_sed_inflation"]


def assert_state_equal(W, ls, what=""):
    h0, h1, f, t = W.download_raw()
    assert np.array_equal(h0, ls.height_q(0)), f"{what}: height plane 0"
    assert np.array_equal(h1, ls.height_q(1)), f"{what}: height plane 1"
    assert np.array_equal(f.view(np.uint32), ls.field().view(np.uint32)), f"{what}: fields"
    assert np.array_equal(t[..., :3], ls.track_q()[..., :3]), f"{what}: tracks"


def assert_stats_equal(st, so):
    for k in STAT_KEYS:
        assert getattr(st, k) == getattr(so, k), k


def make_world(cells, params, **kw):
    kw.setdefault("max_drops", 4096)
    W = shx.World(params=shx.Params.from_buffer_copy(bytes(params)), **kw)
    W.upload(cells)
    ls = orc.Ls(params)
    ls.upload(cells)
    return W, ls


def small_world(tilesize, mapsize, seed):
    """a reduced geometry (tiles of `tilesize`) with random rolling terrain, a few lakes below 0.1"""
    p = orc.default_params(mapsize)
    p.tilesize = tilesize
    size = tilesize * mapsize
    h = orc.synth_terrain(512, seed)[:size, :size].copy()
    h = (h - h.min()) / (h.max() - h.min())
    return p, orc.planar_to_tiled(p, h)


def test_upload_quantises_like_the_oracle(init_cells):
    W, ls = make_world(init_cells, orc.default_params(1))
    assert_state_equal(W, ls, "after upload")
    back = W.download()
    want = ls.download()
    assert np.array_equal(back.view(np.uint8), want.view(np.uint8))
    assert np.abs(back["height"] - init_cells["height"]).max() <= 2.0 ** -27
    W.close()


def test_single_drop_trace_and_deposits_bit_exact(init_cells):
    W, ls = make_world(init_cells, orc.default_params(1))
    for (x, y) in [(256.0, 256.0), (0.0, 0.0), (511.0, 3.5), (100.25, 300.75)]:
        got = W.trace_drop(x, y)
        drops, _ = ls.make_drops(np.array([[x, y]], np.float32))
        drops["flags"] = orc.DROP_ALIVE  # trace_drop does not apply the spawn rejection
        _, want = ls.run_drops(drops, trace_cap=1024)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (x, y)
        assert_state_equal(W, ls, f"drop {(x, y)}")
    W.close()


@pytest.mark.parametrize("kw", [dict(), dict(block_threads=64), dict(block_threads=128), dict(block_threads=512),
                                dict(block_threads=1024), dict(block_threads=256, variant=1), dict(block_threads=32, grid_blocks=7),
                                dict(variant=2, coop=1), dict(variant=3, coop=1), dict(block_threads=1024, coop=1),
                                dict(block_threads=96, variant=2, coop=1, grid_blocks=3),
                                dict(variant=5), dict(variant=5, block_threads=64), dict(variant=5, block_threads=1024),
                                dict(variant=5, block_threads=256, grid_blocks=9)])
def test_erode_cycles_bit_exact_for_every_launch_shape(golden, init_cells, kw):
    """kw: default = the library's own choice; the others force CTA size, register budget, the
    warp-cooperative gather, the eight-lanes-per-drop kernel (variant 5) and, with grid_blocks, sub-batches
    (7 x 32 = 224 / 3 x 96 = 288 / 9 x 32 = 288 drops)."""
    W, ls = make_world(init_cells, orc.default_params(1), **kw)
    cap = kw["grid_blocks"] * kw["block_threads"] // (8 if kw.get("variant") == 5 else 1) if kw.get("grid_blocks") else None
    for c, xy in enumerate(golden["spawn_lists"][:3]):
        st = W.erode_spawnlist(xy)
        if cap is None:
            so = ls.erode_spawnlist(xy)
        else:  # the library marches sub-batches of `cap` drops one after another, in list order
            so = orc.Stats()
            ls.reset_tracks()
            drops, so = ls.make_drops(xy)
            for s in range(0, len(drops), cap):
                chunk = drops[s:s + cap].copy()
                st_c, _ = ls.run_drops(chunk)
                for k in STAT_KEYS[2:]:
                    setattr(so, k, getattr(so, k) + getattr(st_c, k))
            ls.ema(reset=True)
        assert_stats_equal(st, so)
        assert_state_equal(W, ls, f"cycle {c} {kw}")
    W.close()


def test_hash_spawned_erode_matches_oracle(init_cells):
    W, ls = make_world(init_cells, orc.default_params(1))
    assert np.array_equal(W.spawn(512, 1234, 0), ls.spawn(1234, 0, 512))
    for epoch in range(3):  # the library's call counter is the epoch
        st = W.erode(512, seed=77)
        so = ls.erode(512, 77, epoch)
        assert_stats_equal(st, so)
    assert_state_equal(W, ls, "erode x3")
    assert st.launches == 3  # spawn, descend, ema (+reset fused)
    W.close()


@pytest.mark.parametrize("cap", [0, 300, 1536])
def test_large_calls_run_as_batches(init_cells, cap):
    """erode(cycles) with more cycles than shx_config.max_cycles_per_launch (0 = 512): consecutive lock-step
    batches between one reset and one EMA, batch k holding drops [k*cap, (k+1)*cap) of every node"""
    W, ls = make_world(init_cells, orc.default_params(1), max_cycles_per_launch=cap)
    ls.w.contents.max_cycles_per_launch = cap
    for epoch in range(2):
        st = W.erode(1536, seed=5)
        so = ls.erode(1536, 5, epoch)
        assert_stats_equal(st, so)
    assert_state_equal(W, ls, f"batched call, cap {cap}")
    batches = -(-1536 // (cap or 512))
    assert st.launches == 2 * batches + 1  # (spawn + descend) per batch, one fused EMA/reset
    W.close()


@pytest.mark.parametrize("tilesize,mapsize,n", [(64, 1, 40), (64, 3, 700), (32, 4, 2000), (128, 2, 1500)])
def test_reduced_geometries_and_tiling(tilesize, mapsize, n):
    p, cells = small_world(tilesize, mapsize, seed=tilesize + mapsize)
    W, ls = make_world(cells, p)
    size = tilesize * mapsize
    rng = np.random.default_rng(n)
    for c in range(2):
        xy = rng.uniform(-1.5, size + 1.0, size=(n, 2)).astype(np.float32)  # some spawns outside the map
        st = W.erode_spawnlist(xy)
        so = ls.erode_spawnlist(xy)
        assert_stats_equal(st, so)
        assert_state_equal(W, ls, f"{tilesize}x{mapsize} cycle {c}")
    assert np.array_equal(W.spawn(9, 5, 2), ls.spawn(5, 2, 9))
    back, want = W.download(), ls.download()
    assert np.array_equal(back.view(np.uint8), want.view(np.uint8))  # tiled AoS conversion on the way out
    W.close()


def test_edge_cases(init_cells):
    p = orc.default_params(1)
    W, ls = make_world(init_cells, p)
    # empty call: only EMA + track reset
    st = W.erode_spawnlist(np.zeros((0, 2), np.float32))
    so = ls.erode_spawnlist(np.zeros((0, 2), np.float32))
    assert st.steps == 0 and st.phases == 0
    assert_state_equal(W, ls, "empty")
    # corners, borders, cells below the 0.1 spawn threshold, out-of-map spawns, duplicates
    h = orc.tiled_to_planar(p, init_cells)
    low = np.argwhere(h < 0.1)[:8].astype(np.float32)
    xy = np.concatenate([[[0, 0], [0, 511], [511, 0], [511, 511], [0.99, 510.99], [255, 255], [255, 255], [255, 255]],
                         low, [[-1, 5], [5, -1], [512, 5], [5, 512], [1e6, 3], [-0.5, -0.5]]]).astype(np.float32)
    st = W.erode_spawnlist(xy)
    so = ls.erode_spawnlist(xy)
    assert_stats_equal(st, so)
    assert st.rejected >= 8 + 5
    assert_state_equal(W, ls, "edges")
    # exactly 1024 and 1025 drops: last single-CTA size and first multi-CTA size
    rng = np.random.default_rng(3)
    for n in (1024, 1025):
        xy = rng.integers(0, 512, size=(n, 2)).astype(np.float32)
        assert_stats_equal(W.erode_spawnlist(xy), ls.erode_spawnlist(xy))
        assert_state_equal(W, ls, f"n={n}")
    W.close()


def test_run_drops_roundtrip_and_final_states(init_cells):
    """explicit drop records in, final records out (the hand-off format of the strip exchange)"""
    W, ls = make_world(init_cells, orc.default_params(1))
    rng = np.random.default_rng(8)
    xy = rng.uniform(0, 511, size=(300, 2)).astype(np.float32)
    drops, _ = ls.make_drops(xy)
    mine = drops.copy().view(shx.DROP_DTYPE)
    st = W.run_drops(mine)
    so, _ = ls.run_drops(drops)
    assert np.array_equal(mine.view(np.uint8), drops.view(np.uint8))
    assert st.steps == so.steps
    assert not np.any(mine["flags"] & shx.DROP_ALIVE)
    assert_state_equal(W, ls, "run_drops")
    # tracks are still there (no EMA ran): one more EMA consumes them identically
    W.ema()
    ls.ema(reset=False)
    assert_state_equal(W, ls, "ema")
    W.close()


def test_run_to_run_determinism(init_cells):
    rng = np.random.default_rng(12)
    xys = [rng.integers(0, 512, size=(2048, 2)).astype(np.float32) for _ in range(3)]
    out = []
    for rep in range(2):
        with shx.World(mapsize=1, max_drops=4096) as W:
            W.upload(init_cells)
            for xy in xys:
                W.erode_spawnlist(xy)
            out.append(W.download_raw())
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_drop_order_does_not_matter(init_cells):
    rng = np.random.default_rng(13)
    xy = rng.integers(0, 512, size=(3000, 2)).astype(np.float32)
    out = []
    for perm in (np.arange(3000), rng.permutation(3000)):
        with shx.World(mapsize=1, max_drops=4096) as W:
            W.upload(init_cells)
            W.erode_spawnlist(xy[perm])
            out.append(W.download_raw())
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_synth_terrain_matches_oracle():
    for ms, seed in ((1, 1), (2, 9)):
        with shx.World(mapsize=ms) as W:
            W.synth_terrain(seed)
            h0, h1, f, t = W.download_raw()
        want = orc.synth_terrain(512 * ms, seed)
        L = orc.lib()
        q = np.rint(want.astype(np.float64) * 2 ** 26).astype(np.int32)  # == lrintf(h * 2^26): exact scaling
        assert np.array_equal(h0, q) and np.array_equal(h1, q)
        assert not f.any() and not t.any()
        assert L.orc_ls_quantize_height(0.5) == 2 ** 25
