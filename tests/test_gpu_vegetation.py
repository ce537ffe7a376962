"""GPU, N3: Vegetation::grow on the device (shx_veg_*, csrc/shx_veg_kernels.cuh; reference vegetation.h:122-188).

  * bit for bit against the CPU restatement of the same schedule (orc_veg_grow), coupled with the erosion: the plant
    list in order, the integer root counts, the fp32 rootdensity the erosion reads, the per-frame statistics;
  * capacity handling, uploads, the tree model matrices (SimpleHydrology.cpp:329-335);
  * the coupled frame loop entirely on the device (erode + grow, nothing crosses PCIe but the statistics) against
    the reference's own World::erode + Vegetation::grow over 300 frames: statistical bounds stated below."""
import numpy as np
import pytest

import orc
import simplehydrology_b200 as shx

pytestmark = pytest.mark.gpu


def pair(n=128, seed=3, max_plants=8192):
    p = orc.default_params(1)
    p.tilesize = n
    h = (0.3 + 0.2 * orc.synth_terrain(n, seed)).astype(np.float32)
    cells = orc.planar_to_tiled(p, h)
    W = shx.World(params=shx.Params.from_buffer_copy(bytes(p)))
    W.upload(cells)
    ls = orc.Ls(p)
    ls.upload(cells)
    W.veg_create(max_plants)
    ls.veg_create(max_plants)
    return W, ls


def same_state(W, ls):
    h0, h1, f, t = W.download_raw()
    assert np.array_equal(W.veg_plants().view(np.uint32), ls.veg_plants().view(np.uint32))       # list, in order
    assert np.array_equal(t[..., 3], ls.track_q()[..., 3])                                       # root counts (fifths)
    assert np.array_equal(f.view(np.uint32), ls.field().view(np.uint32))                         # incl. rootdensity
    assert np.array_equal(h0, ls.height_q(0))


def test_grow_is_bit_identical_to_the_oracle_coupled_with_erosion():
    W, ls = pair()
    with W:
        # a standing population first (the reference needs ~100 frames to get going): 400 plants on a grid
        g = np.arange(4, 124, 6)
        seedlings = np.array([[x, y, 0.1 * ((x + y) % 7)] for x in g for y in g] +
                             [[0, 0, 0.3], [127, 127, 0.2], [0, 64, 0.1], [64, 127, 0.5], [127, 0, 0.0]], np.float32)  # corners and
        # edges: Plant::root skips the cells that do not exist (getCell == NULL, vegetation.h:91-120)
        W.veg_upload(seedlings, stamp_roots=True)
        ls.veg_upload(seedlings, stamp_roots=True)
        same_state(W, ls)
        born = died = 0
        for f in range(60):
            st, so = W.erode(128, seed=2), ls.erode(128, 2, f)
            assert (st.steps, st.fx_eroded) == (so.steps, so.fx_eroded)
            vg, vo = W.veg_grow(9, f), ls.veg_grow(9, f)
            assert (vg.plants, vg.born, vg.died, vg.refused) == (vo.plants, vo.born, vo.died, vo.refused), f
            born += vg.born
            died += vg.died
        assert born > 200 and died > 20
        same_state(W, ls)
        # the roots the erosion reads are exactly the stamps of the living plants
        cnt = np.zeros((128, 128), np.int64)
        for x, y, _ in W.veg_plants():
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    if 0 <= x + dx < 128 and 0 <= y + dy < 128:
                        cnt[int(x + dx), int(y + dy)] += {0: 5, 1: 3, 2: 2}[abs(dx) + abs(dy)]
        _, _, f, t = W.download_raw()
        assert np.array_equal(t[..., 3], cnt) and np.array_equal(f[..., 3], cnt.astype(np.float32) / np.float32(5))


def test_from_bare_ground_and_capacity():
    W, ls = pair(max_plants=300)
    with W:
        refused = 0
        for f in range(500):
            try:
                vg = W.veg_grow(7, f)
                r = 0
            except shx.ShxError as e:
                assert e.code == -5  # SHX_ERR_CAPACITY: the frame was done, surplus children refused
                r = 1
            vo = ls.veg_grow(7, f)
            refused += vo.refused
            assert (r == 1) == (vo.refused > 0)
            assert W.veg_count() == ls.nplants <= 300
        assert refused > 0
        same_state(W, ls)


def test_tree_models_and_host_roundtrip():
    W, ls = pair()
    with W:
        for f in range(300):
            W.veg_grow(5, f)
        pl = W.veg_plants()
        assert len(pl) > 50
        m = W.veg_tree_models()
        h = W.download()["height"].reshape(128, 128)  # one tile: pool order == map order
        want = np.zeros((len(pl), 16), np.float32)
        s = pl[:, 2]
        want[:, 0] = want[:, 5] = want[:, 10] = s
        want[:, 12] = pl[:, 0]
        want[:, 13] = s + np.float32(80.0) * h[pl[:, 0].astype(int), pl[:, 1].astype(int)]
        want[:, 14] = pl[:, 1]
        want[:, 15] = 1.0
        assert np.array_equal(m.view(np.uint32), want.view(np.uint32))
        # host round trip: a list uploaded without stamping leaves the map alone and continues identically
        _, _, f0, t0 = W.download_raw()
        W.veg_upload(pl, stamp_roots=False)
        _, _, f1, t1 = W.download_raw()
        assert np.array_equal(t0, t1) and np.array_equal(f0.view(np.uint32), f1.view(np.uint32))
        assert np.array_equal(W.veg_plants(), pl)
    with pytest.raises(shx.ShxError):
        with shx.World(mapsize=1) as W2:
            W2.veg_grow(1, 0)  # no plant store


REF_SCRIPT = r"""
import numpy as np, orc
R = orc.Ref(1, seed=%d)
for f in range(%d):
    R.L.ref_frame(512)
n = R.L.ref_plant_count()
pl = np.zeros((n, 3), np.float32)
R.L.ref_plants.argtypes = [__import__("ctypes").c_void_p]
R.L.ref_plants(pl.ctypes.data)
np.savez("%s", cells=R.cells, plants=pl)
"""


@pytest.mark.skipif(not orc.have_ref(1), reason="oracle/_ref not built (no reference tree at build time)")
def test_device_frame_loop_tracks_the_reference(tmp_path):
    """BASELINE configs[4] with everything on the device: shx_erode + shx_veg_grow per frame, 300 frames of the
    reference's default world, against World::erode + Vegetation::grow of the reference itself"""
    frames, seed = 300, 1
    ref_npz = tmp_path / "ref.npz"
    orc.run_ref_script(REF_SCRIPT % (seed, frames, ref_npz), timeout=900)
    ref = np.load(ref_npz)
    rc, rp = ref["cells"], ref["plants"]
    with shx.World(mapsize=1) as W:
        W.init_terrain(seed)
        W.veg_create(1 << 16)
        for f in range(frames):
            W.erode_async(512, seed)
            W.veg_grow(seed, f)
        cells, plants = W.download(), W.veg_plants()
    h0 = orc.init_terrain(1, seed).ravel().astype(np.float64)
    h, hr = cells["height"].astype(np.float64), rc["height"].astype(np.float64)
    rmse = float(np.sqrt(np.mean((h - hr) ** 2)))
    corr_dh = float(np.corrcoef(h - h0, hr - h0)[0, 1])
    corr_dis = float(np.corrcoef(cells["discharge"], rc["discharge"])[0, 1])
    total_dis = float(cells["discharge"].sum(dtype=np.float64) / rc["discharge"].sum(dtype=np.float64))
    root, root_r = float(cells["rootdensity"].sum(dtype=np.float64)), float(rc["rootdensity"].sum(dtype=np.float64))
    print(f"plants {len(plants)} vs reference {len(rp)}; rootdensity sum {root:.1f} vs {root_r:.1f}; max {cells['rootdensity'].max():.2f} vs "
          f"{rc['rootdensity'].max():.2f}; RMSE(height) {rmse:.5f}; corr(dh) {corr_dh:.3f}; corr(discharge) {corr_dis:.3f}; "
          f"total discharge ratio {total_dis:.3f}")
    # the same bounds as the host-vegetation coupling (tests/test_gpu_coupled.py): two runs of the reference with
    # different rand() streams give plants 4001 vs 3719, RMSE 0.0156, corr(dh) 0.908, corr(discharge) 0.264
    assert 0.75 < len(plants) / len(rp) < 1.33
    assert 0.75 < root / root_r < 1.33
    assert 1.5 < float(cells["rootdensity"].max()) < 4.5
    assert rmse < 0.02 and corr_dh > 0.85 and corr_dis > 0.15 and 0.9 < total_dis < 1.1
