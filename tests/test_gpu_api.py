"""GPU: behaviour of the C-ABI boundary itself (transfers, masks, rootdensity pushes, error codes,
parameters, streams) and of the C++ host adaptor that stands in for World::erode."""
import os
import subprocess

import numpy as np
import pytest

import orc
import simplehydrology_b200 as shx
from simplehydrology_b200 import build as B

pytestmark = pytest.mark.gpu


def random_cells(init_cells, seed=0):
    rng = np.random.default_rng(seed)
    cells = init_cells.copy()
    cells["discharge"] = rng.uniform(0, 5, cells.size).astype(np.float32)
    cells["momentumx"] = rng.normal(size=cells.size).astype(np.float32)
    cells["momentumy"] = rng.normal(size=cells.size).astype(np.float32)
    cells["rootdensity"] = rng.uniform(0, 1.5, cells.size).astype(np.float32)
    return cells


def test_download_masks_only_touch_selected_fields(init_cells):
    cells = random_cells(init_cells)
    with shx.World(mapsize=1) as W:
        W.upload(cells)
        full = W.download()
        for mask, fields in ((shx.F_HEIGHT, ["height"]), (shx.F_DISCHARGE | shx.F_MOMENTUM, ["discharge", "momentumx", "momentumy"]),
                             (shx.F_HEIGHT | shx.F_DISCHARGE | shx.F_MOMENTUM, ["height", "discharge", "momentumx", "momentumy"]),
                             (shx.F_ROOTDENSITY, ["rootdensity"]), (shx.F_TRACKS, ["discharge_track", "momentumx_track", "momentumy_track"])):
            out = np.full(cells.size, -7.0, np.float32).repeat(8).view(shx.CELL_DTYPE).copy()
            W.download(out=out, mask=mask)
            for f in shx.CELL_DTYPE.names:
                if f in fields:
                    assert np.array_equal(out[f], full[f]), (mask, f)
                else:
                    assert np.all(out[f] == -7.0), (mask, f)
    assert np.array_equal(full["discharge"], cells["discharge"]) and np.array_equal(full["rootdensity"], cells["rootdensity"])
    assert np.abs(full["height"] - cells["height"]).max() <= 2.0 ** -27


def test_rootdensity_pushes_follow_plant_root(init_cells):
    """Plant::root (vegetation.h:87-118): += f*{1.0, 0.6, 0.4} on the 3x3, cells outside the map skipped,
    several stamps on one cell add up in list order"""
    p = orc.default_params(1)
    T = orc.tiled_index_map(p)
    with shx.World(mapsize=1) as W:
        W.upload(init_cells)
        xy, dl = [], []
        host = np.zeros((512, 512), np.float32)
        for (px, py, f) in [(10, 10, 1.0), (10, 11, 1.0), (0, 0, 1.0), (511, 511, 1.0), (10, 10, -1.0), (300, 7, 1.0)]:
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    w = np.float32(f) * np.float32(1.0 if (dx, dy) == (0, 0) else (0.6 if 0 in (dx, dy) else 0.4))
                    xy.append((px + dx, py + dy))
                    dl.append(w)
                    if 0 <= px + dx < 512 and 0 <= py + dy < 512:
                        host[px + dx, py + dy] = np.float32(host[px + dx, py + dy] + w)
        W.add_rootdensity(np.array(xy, np.int32), np.array(dl, np.float32))
        got = W.download(mask=shx.F_ROOTDENSITY)["rootdensity"][T.ravel()].reshape(512, 512)
        assert np.array_equal(got, host)
        W.set_rootdensity(np.array([[5, 5], [6, 6], [-1, 3]], np.int32), np.array([2.5, 0.25, 9.0], np.float32))
        got = W.download(mask=shx.F_ROOTDENSITY)["rootdensity"][T.ravel()].reshape(512, 512)
        assert got[5, 5] == 2.5 and got[6, 6] == 0.25
        # rootdensity >= 1 switches deposition off (water.h:86-87 clamp): a rooted world erodes differently
        W.set_rootdensity(np.argwhere(np.ones((512, 512), bool)).astype(np.int32), np.full(512 * 512, 2.0, np.float32))
        before = W.download_height_q()[..., 0].copy()
        st = W.erode(64, seed=1)
        assert st.fx_eroded == 0 and st.fx_deposited == 0  # effD == 0 everywhere: only the cascade moves mass
        after = W.download_height_q()[..., 0]
        assert after.astype(np.int64).sum() == before.astype(np.int64).sum()


def test_error_codes(init_cells):
    with shx.World(mapsize=1, max_drops=100) as W:
        with pytest.raises(shx.ShxError) as e:
            W.upload(init_cells[:1000].copy())
        assert e.value.code == -1
        bad = init_cells.copy()
        bad["height"][123] = 40.0
        with pytest.raises(shx.ShxError) as e:
            W.upload(bad)
        assert e.value.code == -3  # SHX_ERR_RANGE
        W.upload(init_cells)
        with pytest.raises(shx.ShxError) as e:
            W.erode_spawnlist(np.zeros((101, 2), np.float32))
        assert e.value.code == -5  # SHX_ERR_CAPACITY
        with pytest.raises(shx.ShxError) as e:
            W.erode(512)
        assert e.value.code == -5
        assert W.erode(100).spawned + W.read_stats().rejected >= 0
    with pytest.raises(shx.ShxError) as e:
        shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL, row0=0, row1=256)
    assert e.value.code == -4  # SHX_ERR_MODE


def test_track_range_overflow_is_an_error_not_a_wrap():
    """more than ~4096 visits of one cell between a reset and the EMA cannot be represented in Q13.18.  Turn-taking
    admits one drop per cell and phase, so one launch tops out near 400 (sum of 0.999^k over a life): the tracks of
    several launches have to pile up on the cell (a call with many batches; here explicit run_drops calls)."""
    p = orc.default_params(1)
    h = np.full((512, 512), 0.5, np.float32)
    cells = orc.planar_to_tiled(p, h)  # perfectly flat: drops do not move (speed stays 0), every step revisits the cell
    ls = orc.Ls(p)
    ls.upload(cells)
    one, _ = ls.make_drops(np.full((1, 2), 100.0, np.float32))
    with shx.World(mapsize=1, max_drops=4096) as W:
        W.upload(cells)
        for _ in range(8):  # 8 x ~394 = 3150: still fine
            W.run_drops(one.copy().view(shx.DROP_DTYPE))
        W.ema()
        W.read_stats()
        W.reset_tracks()
        for _ in range(14):  # ~5500 > 4096
            W.run_drops(one.copy().view(shx.DROP_DTYPE))
        W.ema()
        with pytest.raises(shx.ShxError) as e:
            W.read_stats()
        assert e.value.code == -3


def test_parameters_are_live(init_cells):
    with shx.World(mapsize=1) as W:
        W.upload(init_cells)
        p = shx.default_params(1)
        p.maxAge = 20.0
        W.set_params(p)
        st = W.erode(512, seed=3)
        # 22 phases (ages 0..21, water.h:74) plus at most free_waits (8) more for drops that queued for a shared cell
        assert 22 <= st.phases <= 22 + 8 and st.steps <= 22 * 512
        p.maxAge = 500.0
        p.evapRate = 0.05  # volume < minVol after ~90 steps: water.h:79-82 becomes the terminator
        W.set_params(p)
        st = W.erode(512, seed=3)
        assert st.term_vol > 0 and st.term_age == 0 and st.phases < 120


def test_runs_on_a_caller_stream(init_cells):
    import torch
    s = torch.cuda.Stream()
    with shx.World(mapsize=1) as W:
        W.upload(init_cells)
        W.set_stream(s.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record()
            W.erode_async(512, seed=1)
            e1.record()
        s.synchronize()
        assert e0.elapsed_time(e1) > 0.05  # the kernels really ran on that stream
        assert W.read_stats().steps > 100000


def test_cpp_host_adaptor_equals_the_python_path(tmp_path, init_cells):
    """simplehydrology_b200/host/shx_world.hpp (Bridge::erode<Drop,World>) is the World::erode drop-in;
    its example driver must produce exactly what the same call sequence gives through ctypes."""
    B.build()
    exe = B.build_host_example()
    hfile = tmp_path / "heights.f32"
    init_cells["height"].astype(np.float32).tofile(hfile)
    out = tmp_path / "cells.bin"
    frames, cycles = 3, 512
    r = subprocess.run([exe, str(hfile), "1", str(frames), str(cycles), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(out, dtype=shx.CELL_DTYPE)
    # same sequence in python: stamp rootdensity in the host pool, push the changed cells, erode, download
    p = orc.default_params(1)
    T = orc.tiled_index_map(p)
    host = init_cells.copy()
    with shx.World(mapsize=1) as W:
        W.upload(host)
        for fr in range(frames):
            px, py = 50 + (37 * fr) % 412, 60 + (91 * fr) % 412
            xy, val = [], []
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    i = T[px + dx, py + dy]
                    host["rootdensity"][i] = np.float32(host["rootdensity"][i] + np.float32(1.0 if (dx, dy) == (0, 0) else (0.6 if 0 in (dx, dy) else 0.4)))
                    xy.append((px + dx, py + dy))
                    val.append(host["rootdensity"][i])
            W.set_rootdensity(np.array(xy, np.int32), np.array(val, np.float32))
            W.erode(cycles, seed=1)
            W.download(out=host, mask=shx.F_ALL)
    assert np.array_equal(got.view(np.uint8), host.view(np.uint8))
    assert "total steps" in r.stdout


@pytest.mark.parametrize("mapsize,nthreads", [(1, 1), (2, 3), (4, 0)])
def test_download_compact_equals_the_first_half_of_the_records(mapsize, nthreads):
    """shx_download_compact: {height, discharge, momentumx, momentumy} of every cell, bit for bit what shx_download
    brings, scattered into the first 16 bytes of the pool's records by host threads; tracks and rootdensity untouched"""
    with shx.World(mapsize=mapsize) as W:
        W.init_terrain(3)
        for _ in range(2):
            W.erode(256, seed=5)
        full = W.download()
        pool = np.zeros(full.size, shx.CELL_DTYPE)
        for f in ("discharge_track", "momentumx_track", "momentumy_track", "rootdensity"):
            pool[f] = 7.5  # must survive
        pool["height"] = -1.0
        W.download_compact(pool, nthreads)
    for f in ("height", "discharge", "momentumx", "momentumy"):
        assert np.array_equal(pool[f].view(np.uint32), full[f].view(np.uint32)), f
    for f in ("discharge_track", "momentumx_track", "momentumy_track", "rootdensity"):
        assert np.all(pool[f] == 7.5), f
    assert float(np.abs(full["discharge"]).max()) > 0.0


def test_download_compact_of_a_strip_fills_only_its_own_tiles():
    with shx.World(mapsize=4, row0=512, row1=1536, halo=2) as W:
        W.synth_terrain(2)
        W.erode(64, seed=1)
        full = np.zeros(W.ncells, shx.CELL_DTYPE)
        W.download(out=full)
        pool = np.zeros(W.ncells, shx.CELL_DTYPE)
        pool["height"] = -1.0
        W.download_compact(pool, 4)
    tile = 512 * 512
    own = slice(4 * tile, 12 * tile)  # tile rows 1 and 2 of four
    for f in ("height", "discharge", "momentumx", "momentumy"):
        assert np.array_equal(pool[f][own].view(np.uint32), full[f][own].view(np.uint32)), f
    assert np.all(pool["height"][:4 * tile] == -1.0) and np.all(pool["height"][12 * tile:] == -1.0)
