"""The CPU restatement (oracle/shx_oracle.c, orc_seq_*) against golden vectors produced by the
reference's own headers (tests/golden/make_golden.py).  Everything is bit-exact: the oracle is a
statement-for-statement restatement of water.h:58-156, world.h:54-168, cellpool.h:181-204."""
import hashlib

import numpy as np

import orc


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).digest(), np.uint8)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_cell_and_drop_layouts():
    # quad::cell is 8 x f32 = 32 B (cellpool.h:207-220); Drop is 28 B + status word
    assert orc.CELL_DTYPE.itemsize == 32
    assert orc.DROP_DTYPE.itemsize == 32
    p = orc.default_params(4)
    assert (p.maxAge, p.evapRate, p.depositionRate, p.minVol) == (500.0, np.float32(0.001), np.float32(0.1), np.float32(0.01))
    assert (p.entrainment, p.gravity, p.momentumTransfer) == (10.0, 1.0, 1.0)
    assert (p.lrate, p.maxdiff, p.settling) == (np.float32(0.1), np.float32(0.01), np.float32(0.8))
    assert (p.mapscale, p.tilesize, p.mapsize, p.lodsize) == (80, 512, 4, 1)


def test_tiled_index_is_node_major_then_x_major():
    p = orc.default_params(4)
    L = orc.lib()
    import ctypes as C
    # cellpool.h:327-336 (node ind = i*mapsize+j owns a contiguous tile), math.h:11-14 (x*res.y + y)
    assert L.orc_tiled_index(C.byref(p), 0, 0) == 0
    assert L.orc_tiled_index(C.byref(p), 0, 1) == 1
    assert L.orc_tiled_index(C.byref(p), 1, 0) == 512
    assert L.orc_tiled_index(C.byref(p), 0, 512) == 512 * 512
    assert L.orc_tiled_index(C.byref(p), 512, 0) == 4 * 512 * 512
    assert L.orc_tiled_index(C.byref(p), 513, 514) == 5 * 512 * 512 + 512 + 2
    T = orc.tiled_index_map(p)
    assert T[513, 514] == 5 * 512 * 512 + 512 + 2 and np.unique(T).size == T.size


def test_initial_world_statistics(golden):
    h = golden["init_height"]
    # SURVEY.md 8c probe figures for ./hydrology 1
    assert h.min() == 0.0 and h.max() == 1.0
    assert abs(float(h.mean(dtype=np.float64)) - 0.494445) < 1e-5
    assert int((h < 0.1).sum()) == 805


def test_normals_match_reference(golden, init_cells):
    S = orc.Seq(init_cells.copy())
    cells = [(0, 0), (0, 5), (511, 511), (511, 0), (3, 511), (200, 200), (17, 340), (0, 511)]
    got = np.stack([S.normal(x, y) for x, y in cells])
    assert np.array_equal(bits(got), bits(golden["normals"]))


def test_single_drop_traces_bit_exact(golden, init_cells):
    cells = init_cells.copy()
    S = orc.Seq(cells)
    for (x, y), n, want in zip(golden["trace_starts"], golden["trace_len"], golden["traces"]):
        got = S.trace_drop(float(x), float(y))
        assert len(got) == n
        assert np.array_equal(bits(got), bits(want[:n]))
    # the drop from (256,256) is the SURVEY.md 8c probe: 502 calls, final state
    t = golden["traces"][0][golden["trace_len"][0] - 1]
    assert golden["trace_len"][0] == 502
    np.testing.assert_allclose(t[1:3], [259.8272, 207.7774], atol=2e-4)
    np.testing.assert_allclose(t[5:7], [0.60577, 0.0051339], rtol=1e-4)
    # height deltas and track deposits left behind by the seven drops
    assert np.array_equal(sha(cells), golden["after_traces_sha"])
    idx = golden["after_traces_idx"]
    assert np.array_equal(bits(cells["height"][idx]), bits(golden["after_traces_height"]))
    tidx = golden["after_traces_track_idx"]
    got = np.stack([cells[f][tidx] for f in ("discharge_track", "momentumx_track", "momentumy_track")], 1)
    assert np.array_equal(bits(got), bits(golden["after_traces_tracks"]))


def test_erode_cycles_bit_exact(golden, init_cells):
    cells = init_cells.copy()
    S = orc.Seq(cells)  # libm erf, as the reference
    for c, xy in enumerate(golden["spawn_lists"]):
        st = S.erode_spawnlist(xy)
        assert [st.spawned, st.rejected, st.steps] == list(golden["cycles_stats"][c])
        assert np.array_equal(sha(cells), golden["cycles_sha"][c]), f"cycle {c}"
    got = cells[golden["cycles_sample_idx"]].view(np.float32).reshape(-1, 8)
    assert np.array_equal(bits(got), bits(golden["cycles_sample_cells"]))


def test_stock_world_erode_is_the_spawnlist_loop(golden, init_cells):
    """World::erode(512) after srand(1) == the explicit-spawn-list loop fed glibc's rand() sequence
    (y consumes the first rand(): argument evaluation order of world.h:69 under g++ 13)."""
    cells = init_cells.copy()
    S = orc.Seq(cells)
    xy = golden["stock_spawns"].reshape(3, 512, 2)
    for f in range(3):
        S.erode_spawnlist(xy[f])
        assert np.array_equal(sha(cells), golden["stock_erode_sha"][f]), f"frame {f}"


def test_cascade_known_answers(golden):
    p = orc.default_params(1)
    idx = orc.tiled_index_map(p)
    for case, want in zip(golden["cascade_cases"], golden["cascade_after"]):
        x0, y0, cx, cy = int(case[0]), int(case[1]), float(case[2]), float(case[3])
        cells = np.zeros(512 * 512, orc.CELL_DTYPE)
        cells["height"][:] = 0.5
        patch = case[4:].reshape(9, 9)
        for i in range(9):
            for j in range(9):
                cells["height"][idx[x0 + i, y0 + j]] = patch[i, j]
        orc.Seq(cells).cascade(cx, cy)
        after = np.array([[cells["height"][idx[x0 + i, y0 + j]] for j in range(9)] for i in range(9)], np.float32)
        assert np.array_equal(bits(after.ravel()), bits(want))


def test_cascade_conserves_and_only_touches_3x3(init_cells):
    cells = init_cells.copy()
    S = orc.Seq(cells)
    before = cells["height"].copy()
    n = S.cascade(300.5, 200.5)
    d = cells["height"].astype(np.float64) - before
    changed = np.nonzero(d)[0]
    p = orc.default_params(1)
    T = orc.tiled_index_map(p)
    allowed = {int(T[300 + i, 200 + j]) for i in (-1, 0, 1) for j in (-1, 0, 1)}
    assert set(changed.tolist()) <= allowed
    assert n == 0 or abs(d.sum()) < 1e-7  # +-transfer pairs (world.h:157-164)


def test_edge_positions(init_cells):
    cells = init_cells.copy()
    S = orc.Seq(cells)
    # pos = -0.4 truncates to cell 0 and is in bounds; -1.0 is out
    d = np.zeros(1, orc.DROP_DTYPE)
    d["px"], d["py"], d["volume"], d["flags"] = -0.4, -0.4, 1.0, orc.DROP_ALIVE
    S.descend(d)
    assert d["flags"][0] != orc.DROP_DONE_NULL
    d = np.zeros(1, orc.DROP_DTYPE)
    d["px"], d["py"], d["volume"], d["flags"] = -1.0, 5.0, 1.0, orc.DROP_ALIVE
    assert S.descend(d) == 0 and d["flags"][0] == orc.DROP_DONE_NULL
    d = np.zeros(1, orc.DROP_DTYPE)
    d["px"], d["py"], d["volume"], d["flags"] = 512.0, 5.0, 1.0, orc.DROP_ALIVE
    assert S.descend(d) == 0 and d["flags"][0] == orc.DROP_DONE_NULL
    # empty spawn list: only reset + EMA happen
    cells2 = init_cells.copy()
    cells2["discharge_track"] = 3.0
    cells2["discharge"] = 1.0
    st = orc.Seq(cells2).erode_spawnlist(np.zeros((0, 2), np.float32))
    assert st.steps == 0 and np.all(cells2["discharge_track"] == 0)
    assert np.allclose(cells2["discharge"], 0.9)


def test_erf_restatement_accuracy():
    from scipy import special
    xs = np.concatenate([np.linspace(0, 4.5, 20001), -np.linspace(0, 4.5, 2001), [0.0, 1e-30, 10.0, 0.875, 4.0]]).astype(np.float32)
    L = orc.lib()
    got = np.array([L.orc_erff_poly(float(x)) for x in xs], np.float64)
    want = special.erf(xs.astype(np.float64))
    ulp = np.spacing(np.maximum(np.abs(want), 1e-30).astype(np.float32)).astype(np.float64)
    assert np.max(np.abs(got - want) / ulp) < 1.6  # tools/fit_erf.py: 1.47 ulp
    assert L.orc_erff_poly(0.0) == 0.0 and L.orc_erff_libm(0.0) == 0.0


def test_vertex_fill_matches_reference_updatenode(golden, init_cells):
    """quad::updatenode (cellpool.h:286-305): 48-byte Vertex records of the fresh map and after one erode cycle"""
    cells = init_cells.copy()
    p = orc.default_params(1)
    S = orc.Seq(cells)
    for tag in ("fresh", "eroded"):
        v = orc.vertex_fill(p, cells) + np.float32(0.0)
        assert np.array_equal(bits(v[golden["vertex_sample_idx"]]), bits(golden[f"vertex_{tag}_sample"])), tag
        assert np.array_equal(sha(v), golden[f"vertex_{tag}_sha"]), tag
        S.erode_spawnlist(golden["spawn_lists"][0])


def test_terrain_init_restatement_reproduces_the_reference_world(golden):
    """orc_init_terrain (map::init restated, cellpool.h:349-409 + the vendored FastNoiseLite.h) against the heights
    the compiled reference produced for ./hydrology 1: bit-identical, incl. the survey's probe figures"""
    got = orc.init_terrain(1, 1)
    want = golden["init_height"].reshape(512, 512)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert int((got < 0.1).sum()) == 805 and abs(float(got.mean()) - 0.494445) < 1e-6 and got.min() == 0.0 and got.max() == 1.0
