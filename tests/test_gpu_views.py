"""GPU: the per-frame consumers of the eroded map run on the device:
shx_vertex_fill = quad::updatenode (cellpool.h:286-305), shx_view_maps = the dischargeMap /
momentumMap builders (SimpleHydrology.cpp:341-354).  Checked against the committed reference
vectors (tests/golden) and the oracle restatement; float equality (the sign of a zero is not pinned)."""
import numpy as np
import pytest
import torch

import orc
import simplehydrology_b200 as shx

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_vertex_fill_matches_the_reference_vectors(golden, init_cells):
    """sequential mode keeps fp32 heights: the Vertex records equal the reference's own (golden) bit for bit"""
    with shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL) as W:
        W.upload(init_cells)
        v = W.vertex_download() + np.float32(0.0)
        idx = golden["vertex_sample_idx"]
        assert np.array_equal(bits(v[idx]), bits(golden["vertex_fresh_sample"]))
        assert np.array_equal(bits(v), bits(orc.vertex_fill(orc.default_params(1), init_cells) + np.float32(0.0)))
        W.erode_spawnlist(golden["spawn_lists"][0])
        v = W.vertex_download() + np.float32(0.0)
        assert np.array_equal(bits(v[idx]), bits(golden["vertex_eroded_sample"]))


@pytest.mark.parametrize("mapsize", [1, 4])
def test_vertex_fill_and_view_maps_after_batched_erosion(mapsize):
    """batched mode: the kernels read the Q5.26 heights; the oracle gets the downloaded cells (same values)"""
    p = shx.default_params(mapsize)
    with shx.World(params=p) as W:
        W.synth_terrain(3)
        for _ in range(3):
            W.erode(256, 7)
        cells = W.download()
        v = W.vertex_download()
        m = W.view_maps_download()
    po = orc.default_params(mapsize)
    assert np.array_equal(v, orc.vertex_fill(po, cells))
    want = orc.view_maps(po, cells, erf_poly=1)
    assert np.array_equal(m, want)
    # against libm's erf (what the reference calls) the kernels' polynomial stays within 2 ulp of 1
    ref = orc.view_maps(po, cells, erf_poly=0)
    assert np.abs(m - ref).max() <= 2.4e-7
    assert m[:, 0].max() > 0.5  # rivers formed: discharge alpha is not trivially zero
    n = v[:, 3:6]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-6)


def test_vertex_fill_into_a_caller_owned_device_buffer_and_strips():
    """the interop path: the caller hands a device pointer (a mapped VBO in the reference's renderer);
    strips fill their own nodes and together give the whole pool"""
    ms = 2
    with shx.World(mapsize=ms) as W:
        W.synth_terrain(5)
        whole = W.vertex_download()
        buf = torch.zeros(W.owned_cells() * 12, dtype=torch.float32, device="cuda")
        W.set_stream(torch.cuda.current_stream().cuda_stream)
        W.vertex_fill(buf.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(buf.cpu().numpy().reshape(-1, 12), whole)
    parts = []
    for r in range(2):
        with shx.World(mapsize=ms, row0=r * 512, row1=(r + 1) * 512) as S:
            S.synth_terrain(5)
            parts.append(S.vertex_download())
    assert np.array_equal(np.concatenate(parts), whole)


def test_gather_cells_returns_records_and_map_normals(golden, init_cells):
    """shx_gather_cells: the sparse read-back for Vegetation::grow-style host code"""
    cells_xy = [(0, 0), (0, 5), (511, 511), (511, 0), (3, 511), (200, 200), (17, 340), (0, 511)]  # tests/golden NORMAL_CELLS
    with shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL) as W:
        W.upload(init_cells)
        rec, nrm = W.gather_cells(cells_xy)
        assert np.array_equal(bits(nrm + np.float32(0.0)), bits(golden["normals"] + np.float32(0.0)))  # World::map.normal of the reference
        idx = orc.tiled_index_map(orc.default_params(1))
        want = init_cells[[idx[x, y] for x, y in cells_xy]]
        assert np.array_equal(rec.view(np.uint8), want.view(np.uint8))
        # outside the map: map.get() == NULL in the reference -> zeros here
        rec, nrm = W.gather_cells([(-1, 3), (512, 0), (5, 512), (5, -1)])
        assert not rec.view(np.uint8).any() and not nrm.any()
        assert W.gather_cells(np.zeros((0, 2), np.int32), normals=False).size == 0
    # batched mode on a tiled world after erosion: every field equals the full download; normals cross tile borders
    ms = 2
    p = orc.default_params(ms)
    with shx.World(mapsize=ms) as W:
        W.synth_terrain(4)
        W.erode(256, 3)
        full = W.download()
        rng = np.random.default_rng(2)
        xy = np.concatenate([rng.integers(0, 1024, size=(500, 2)), [[511, 300], [512, 300], [300, 511], [300, 512], [0, 0], [1023, 1023]]])
        rec, nrm = W.gather_cells(xy)
    idx = orc.tiled_index_map(p)
    assert np.array_equal(rec.view(np.uint8), full[idx[xy[:, 0], xy[:, 1]]].view(np.uint8))
    S = orc.Seq(full.copy(), params=p)
    want = np.stack([S.normal(int(x), int(y)) for x, y in xy])
    assert np.array_equal(nrm + np.float32(0.0), want + np.float32(0.0))


@pytest.mark.parametrize("mapsize,seed,tilesize", [(1, 1, 512), (2, 20007, 512), (3, 5, 64)])
def test_device_terrain_init_is_the_references(mapsize, seed, tilesize, golden):
    """shx_init_terrain = World::map.init (cellpool.h:349-409) on the device: the Q5.26 heights are the quantised
    heights of the restated (and reference-pinned) init, every other field zero; seed 1 is ./hydrology 1 itself"""
    p = shx.default_params(mapsize)
    p.tilesize = tilesize
    with shx.World(params=p) as W:
        W.init_terrain(seed)
        h0, h1, f, t = W.download_raw()
        cells = W.download()
    want = orc.init_terrain(mapsize, seed, tilesize)
    q = np.rint(want.astype(np.float64) * 2 ** 26).astype(np.int32)
    assert np.array_equal(h0, q) and np.array_equal(h1, q)
    assert not f.any() and not t.any()
    if (mapsize, seed, tilesize) == (1, 1, 512):
        ref = golden["init_height"].reshape(512, 512)
        assert np.abs(cells["height"].reshape(512, 512) - ref).max() <= 2.0 ** -27
        assert int((cells["height"] < 0.1).sum()) == 805  # SURVEY.md 8c probe figure


def test_device_terrain_init_in_sequential_mode_is_bit_exact(golden):
    with shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL) as W:
        W.init_terrain(1)
        cells = W.download()
    assert np.array_equal(cells["height"].view(np.uint32), golden["init_height"].view(np.uint32))


def test_view_textures_are_the_packed_view_maps():
    """shx_view_textures: the dischargeMap / momentumMap RGBA8 textures (SimpleHydrology.cpp:341-354).  Every channel is
    (unsigned char)(255*c) of the pinned float values of shx_view_maps; water colour as model.h:22."""
    with shx.World(mapsize=2) as W:
        W.synth_terrain(4)
        for _ in range(3):
            W.erode(512, 3)
        maps = W.view_maps_download()
        dis, mom = W.view_textures_download()
        dis2, _ = W.view_textures_download(water_rgb=[0.25, 0.5, 1.0])
    q = lambda c: (np.float32(255.0) * c.astype(np.float32)).astype(np.int32).astype(np.uint8)
    water = np.array([92, 133, 142], np.float32) / np.float32(255.0)
    assert np.array_equal(dis[:, 3], q(maps[:, 0])) and (dis[:, :3] == q(water)).all()
    assert np.array_equal(mom[:, 0], q(maps[:, 1])) and np.array_equal(mom[:, 1], q(maps[:, 2]))
    assert (mom[:, 2] == 127).all() and (mom[:, 3] == 255).all()
    assert (dis2[:, :3] == np.array([63, 127, 255], np.uint8)).all() and np.array_equal(dis2[:, 3], dis[:, 3])
    assert dis[:, 3].max() > 100  # rivers did form
