"""GPU, sequential mode (SHX_MODE_SEQUENTIAL): one GPU thread marches drops one after another in
fp32.  North-star check (1): a single drop must reproduce the reference's trajectory, height deltas
and track deposits.  On a fresh world erf only ever sees 0, so the comparison with the reference's
golden vectors is BIT-EXACT; once discharge is non-zero the kernels' own erf (<= 1.5 ulp from the
exact function, the reference's libm erf is <= 1 ulp) takes over and the bit-exact partner is the
restatement run with that same erf."""
import hashlib

import numpy as np
import pytest

import orc
import simplehydrology_b200 as shx

pytestmark = pytest.mark.gpu


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).digest(), np.uint8)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_upload_download_roundtrip_is_exact(init_cells):
    rng = np.random.default_rng(0)
    cells = init_cells.copy()
    for f in ("discharge", "momentumx", "momentumy", "discharge_track", "momentumx_track", "momentumy_track", "rootdensity"):
        cells[f] = rng.normal(size=cells.size).astype(np.float32)
    with shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL) as W:
        W.upload(cells)
        back = W.download()
    assert np.array_equal(back.view(np.uint8), cells.view(np.uint8))


def test_single_drop_trajectories_bit_exact_vs_reference(golden, init_cells):
    with shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL) as W:
        W.upload(init_cells)
        for (x, y), n, want in zip(golden["trace_starts"], golden["trace_len"], golden["traces"]):
            got = W.trace_drop(float(x), float(y))
            assert len(got) == n
            assert np.array_equal(bits(got), bits(want[:n])), (x, y)
        cells = W.download()
    # height deltas and discharge / momentum track deposits of the seven drops
    assert np.array_equal(sha(cells), golden["after_traces_sha"])
    idx = golden["after_traces_idx"]
    assert np.array_equal(bits(cells["height"][idx]), bits(golden["after_traces_height"]))


def test_first_erode_cycle_bit_exact_vs_reference(golden, init_cells):
    with shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL, keep_tracks=1) as W:
        W.upload(init_cells)
        st = W.erode_spawnlist(golden["spawn_lists"][0])
        cells = W.download()
    assert [st.spawned, st.rejected, st.steps] == list(golden["cycles_stats"][0])
    assert np.array_equal(sha(cells), golden["cycles_sha"][0])


def test_later_cycles_bit_exact_vs_restatement_and_close_to_reference(golden, init_cells):
    ref = init_cells.copy()
    S_ref = orc.Seq(ref)                      # libm erf == the reference (golden-pinned)
    own = init_cells.copy()
    S_own = orc.Seq(own, erf_poly=True)       # the kernels' erf
    with shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL, keep_tracks=1) as W:
        W.upload(init_cells)
        for c, xy in enumerate(golden["spawn_lists"][:3]):
            S_ref.erode_spawnlist(xy)
            so = S_own.erode_spawnlist(xy)
            st = W.erode_spawnlist(xy)
            cells = W.download()
            assert (st.spawned, st.rejected, st.steps, st.cascade_transfers) == (so.spawned, so.rejected, so.steps, so.cascade_transfers)
            assert np.array_equal(cells.view(np.uint8), own.view(np.uint8)), f"cycle {c}"
            assert (st.fx_sed_inflation, st.fx_sed_deposited, st.fx_sed_oob_lost) == \
                   (so.fx_sed_inflation, so.fx_sed_deposited, so.fx_sed_oob_lost)
            # against the reference itself: identical while erf(0)=0, then the 1-ulp erf difference is
            # amplified by the chaotic dynamics -- bounded like two reference builds
            rmse = float(np.sqrt(np.mean((cells["height"].astype(np.float64) - ref["height"]) ** 2)))
            assert rmse == 0.0 if c == 0 else rmse < 4e-3


def test_trace_of_a_drop_outside_the_map_is_empty(init_cells):
    with shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL) as W:
        W.upload(init_cells)
        assert len(W.trace_drop(-5.0, 10.0)) == 0  # water.h:62-68
        assert len(W.trace_drop(10.0, 512.0)) == 0
        assert len(W.trace_drop(-0.5, -0.5)) >= 1   # truncates to cell (0,0): in bounds


def test_spawn_rejection_sees_what_earlier_drops_of_the_call_left():
    """world.h:71-74 tests `height < 0.1` when each drop is CREATED, i.e. after the earlier drops of the same call
    have run (round-1 advisor finding).  A cell just above 0.1 on a slope: the first drop spawned there erodes it
    below 0.1, the second spawn on the same cell must be rejected."""
    p = orc.default_params(1)
    p.tilesize = 64
    y = np.arange(64, dtype=np.float32)
    h = np.tile(np.float32(0.1003) + np.float32(0.004) * (np.float32(32.0) - y), (64, 1)).astype(np.float32)  # falls along +y
    cells = orc.planar_to_tiled(p, h)
    xy = np.array([[32.0, 32.0], [32.0, 32.0], [10.0, 20.0]], np.float32)
    want = cells.copy()
    so = orc.Seq(want, p, erf_poly=True).erode_spawnlist(xy)
    assert (so.spawned, so.rejected) == (2, 1)  # the restatement follows the reference: second spawn on (32, 32) rejected
    assert want["height"][orc.lib().orc_tiled_index(p, 32, 32)] < np.float32(0.1) <= cells["height"][orc.lib().orc_tiled_index(p, 32, 32)]
    with shx.World(params=shx.Params.from_buffer_copy(bytes(p)), mode=shx.MODE_SEQUENTIAL) as W:
        W.upload(cells)
        st = W.erode_spawnlist(xy)
        got = W.download()
    assert (st.spawned, st.rejected, st.steps) == (so.spawned, so.rejected, so.steps)
    for f in ("height", "discharge", "momentumx", "momentumy"):
        assert np.array_equal(bits(got[f]), bits(want[f])), f
