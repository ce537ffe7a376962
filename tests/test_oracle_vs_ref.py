"""Live pinning of the CPU restatement against the reference's own headers compiled headless
(oracle/_ref, built from /root/reference by oracle/Makefile).  Skipped where oracle/_ref is absent;
tests/test_oracle_golden.py then carries the same checks through committed fixtures."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.skipif(not orc.have_ref(1), reason="oracle/_ref not built (no reference tree)")

SCRIPT = r"""
import sys, numpy as np, orc
R = orc.Ref(1, seed=1)
cells = R.cells.copy()
S = orc.Seq(cells)
rng = np.random.default_rng(%d)
ok = True
for (x, y) in rng.uniform(0, 511.9, size=(12, 2)):
    a = R.trace_drop(float(x), float(y)); b = S.trace_drop(float(x), float(y))
    ok &= a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
for c in range(3):
    xy = rng.uniform(-2, 514, size=(300, 2)).astype(np.float32)   # includes out-of-map spawns
    sr = R.erode_spawnlist(xy); so = S.erode_spawnlist(xy)
    ok &= (sr["spawned"], sr["rejected"], sr["steps"]) == (so.spawned, so.rejected, so.steps)
    ok &= np.array_equal(R.cells.view(np.uint8), cells.view(np.uint8))
for (x, y) in rng.integers(0, 512, size=(64, 2)):
    ok &= np.array_equal(R.normal(x, y).view(np.uint32), S.normal(x, y).view(np.uint32))
for (x, y) in rng.uniform(0, 511.9, size=(64, 2)):
    R.cascade(float(x), float(y)); S.cascade(float(x), float(y))
ok &= np.array_equal(R.cells.view(np.uint8), cells.view(np.uint8))
print("OK" if ok else "MISMATCH")
"""


@pytest.mark.parametrize("seed", [3, 4])
def test_restatement_is_bit_identical_to_reference(seed):
    assert orc.run_ref_script(SCRIPT % seed).strip().endswith("OK")


def test_params_and_sizes_of_the_compiled_reference():
    out = orc.run_ref_script(
        "import orc, numpy as np\nR = orc.Ref(1)\nprint(R.L.ref_cell_bytes(), R.L.ref_drop_bytes(), R.size, list(R.params()))")
    toks = out.strip().split(" ", 3)
    assert toks[0] == "32" and toks[1] == "28" and toks[2] == "512"  # sizeof(cell) 32, sizeof(Drop) 28
    p = orc.default_params(1)
    want = [p.maxAge, p.minVol, p.evapRate, p.depositionRate, p.entrainment, p.gravity, p.momentumTransfer, p.lrate,
            p.maxdiff, p.settling]
    got = eval(toks[3], {"np": np})
    assert [np.float32(a) for a in got] == [np.float32(b) for b in want]


@pytest.mark.skipif(not orc.have_ref(4), reason="mapsize-4 reference not built")
def test_mapsize4_tiling_matches_reference():
    """2048^2: the tiled pool layout (cellpool.h:327-336,421-426) and a drop crossing tile borders"""
    out = orc.run_ref_script(r"""
import numpy as np, orc
R = orc.Ref(4)
p = orc.default_params(4)
h = orc.synth_terrain(2048, 5)
R.cells[:] = orc.planar_to_tiled(p, h)
cells = R.cells.copy()
S = orc.Seq(cells, p)
ok = True
for (x, y) in [(510.5, 510.5), (1023.2, 1500.7), (2047.0, 0.0), (5.5, 2046.5), (1024.0, 1024.0)]:
    a = R.trace_drop(x, y); b = S.trace_drop(x, y)
    ok &= a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
ok &= np.array_equal(R.cells.view(np.uint8), cells.view(np.uint8))
print("OK" if ok else "MISMATCH")
""")
    assert out.strip().endswith("OK")


VERTEX_SCRIPT = r"""
import numpy as np, orc
R = orc.Ref(%d, seed=1)
rng = np.random.default_rng(5)
ok = True
for rounds in range(2):
    a = R.vertices(); b = orc.vertex_fill(orc.default_params(%d), R.cells)
    ok &= a.shape == b.shape and np.array_equal(a, b)      # float equality: only the sign of a zero may differ
    ok &= float(np.abs(a[:, 3:6]).max()) <= 1.0 and a[:, 4].min() > 0.0
    R.erode_spawnlist(rng.uniform(0, R.size - 1, size=(400, 2)).astype(np.float32))   # vertices of an eroded map too
print("OK" if ok else "MISMATCH")
"""


@pytest.mark.parametrize("mapsize", [1, 4])
def test_vertex_fill_restatement_matches_updatenode(mapsize):
    """quad::updatenode (cellpool.h:286-305) through the reference's own Vertexpool::fill vs orc_vertex_fill"""
    if not orc.have_ref(mapsize):
        pytest.skip("oracle/_ref for this map size not built")
    assert orc.run_ref_script(VERTEX_SCRIPT % (mapsize, mapsize)).strip().endswith("OK")


INIT_SCRIPT = r"""
import numpy as np, orc
R = orc.Ref(%d, seed=%d)
p = orc.default_params(%d)
want = orc.tiled_to_planar(p, R.cells)
got = orc.init_terrain(%d, %d)
print("OK" if np.array_equal(want.view(np.uint32), got.view(np.uint32)) else "MISMATCH", float(np.abs(want - got).max()))
"""


@pytest.mark.parametrize("mapsize,seed", [(1, 7), (1, 12345), (4, 1)])
def test_terrain_init_restatement_matches_map_init(mapsize, seed):
    """map::init (cellpool.h:349-409, FastNoiseLite OpenSimplex2 fBm) through the compiled reference vs
    orc_init_terrain: bit-identical heights, also for a seed above 10000 (SEED % 10000) and a tiled 2048^2 map"""
    if not orc.have_ref(mapsize):
        pytest.skip("oracle/_ref for this map size not built")
    out = orc.run_ref_script(INIT_SCRIPT % (mapsize, seed, mapsize, mapsize, seed))
    assert out.strip().startswith("OK"), out


VIEW_SCRIPT = r"""
import ctypes as C, numpy as np, orc
R = orc.Ref(1, seed=1)
rng = np.random.default_rng(3)
for c in range(6):
    R.erode_spawnlist(rng.integers(0, 512, size=(512, 2)).astype(np.float32))
R.L.ref_discharge.restype = C.c_float
p = orc.default_params(1)
maps = orc.view_maps(p, R.cells, erf_poly=0).reshape(512, 512, 4)
ok = True
worst = 0.0
for (x, y) in rng.integers(0, 512, size=(4000, 2)):
    ok &= np.float32(R.L.ref_discharge(int(x), int(y))) == maps[x, y, 0]
maps_p = orc.view_maps(p, R.cells, erf_poly=1).reshape(512, 512, 4)
worst = float(np.abs(maps_p - maps).max())
print("OK" if ok and maps[..., 0].max() > 0.5 else "MISMATCH", worst)
"""


def test_view_maps_restatement_matches_map_discharge():
    """orc_view_maps' discharge channel against the reference's own World::map.discharge (cellpool.h:242-244,439-443)
    on an eroded world: bit-identical with libm's erf; the kernels' polynomial erf differs by at most 2e-7"""
    out = orc.run_ref_script(VIEW_SCRIPT).strip().split()
    assert out[0] == "OK" and float(out[1]) < 2e-7, out
