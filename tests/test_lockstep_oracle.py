"""Properties of the lock-step (batched, fixed-point) restatement orc_ls_* -- the bit-exact checker
of the CUDA path.  It applies the same per-step arithmetic as the sequential oracle under the
batched schedule, so: one drop alone behaves like the reference up to fixed-point rounding, the
result is independent of drop order, and the integer mass ledger closes exactly."""
import numpy as np

import orc

H_LSB = 2.0 ** -26


def test_single_drop_follows_the_reference_until_rounding_separates_them(init_cells):
    p = orc.default_params(1)
    for (x, y) in [(256.0, 256.0), (100.0, 300.0), (37.75, 129.5)]:
        S = orc.Seq(init_cells.copy(), erf_poly=True)
        ls = orc.Ls(p)
        ls.upload(init_cells)
        a = S.trace_drop(x, y)
        drops, _ = ls.make_drops(np.array([[x, y]], np.float32))
        st, b = ls.run_drops(drops, trace_cap=1024)
        n = 64  # first 64 steps: heights differ by <= a few 2^-26, positions by <= 1e-3 cell
        assert len(a) >= n and len(b) >= n
        assert np.abs(a[:n, 1:3] - b[:n, 1:3]).max() < 1e-3
        assert np.abs(a[:n, 5] - b[:n, 5]).max() == 0.0  # volume: identical arithmetic
        assert np.abs(a[:n, 6] - b[:n, 6]).max() < 1e-6
        assert np.array_equal(ls.height_q(0), ls.height_q(1))  # planes agree outside a run


def test_height_quantisation_roundtrip():
    L = orc.lib()
    for h in [0.0, 1.0, 0.5, 0.1, 0.3333333, -0.25, 30.99]:
        q = L.orc_ls_quantize_height(h)
        assert abs(q * H_LSB - np.float32(h)) <= H_LSB / 2 + 1e-12


def test_order_independence_and_determinism(init_cells):
    p = orc.default_params(1)
    rng = np.random.default_rng(21)
    xy = rng.integers(0, 512, size=(300, 2)).astype(np.float32)
    results = []
    for perm in (np.arange(300), rng.permutation(300), rng.permutation(300)):
        ls = orc.Ls(p)
        ls.upload(init_cells)
        st = ls.erode_spawnlist(xy[perm])
        results.append((ls.height_q(0).copy(), ls.field().copy(), st.as_dict()))
    for h, f, st in results[1:]:
        assert np.array_equal(h, results[0][0])
        assert np.array_equal(f.view(np.uint32), results[0][1].view(np.uint32))
        assert st == results[0][2]


def test_integer_mass_ledger_closes_exactly(init_cells):
    p = orc.default_params(1)
    ls = orc.Ls(p)
    ls.upload(init_cells)
    rng = np.random.default_rng(5)
    for _ in range(3):
        before = ls.height_q(0).astype(np.int64).sum()
        st = ls.erode_spawnlist(rng.integers(0, 512, size=(400, 2)).astype(np.float32))
        after = ls.height_q(0).astype(np.int64).sum()
        assert after - before == st.fx_deposited - st.fx_eroded
        assert st.spawned == st.term_age + st.term_vol + st.term_oob
        # sediment budget of the drops (water.h:131,135,139-142): eroded + inflation = deposited + lost,
        # up to one rounding (2^-27) per step of the fp32 sediment against its Q5.26 image
        lhs = st.fx_eroded * H_LSB + st.fx_sed_inflation * 2.0 ** -32
        rhs = (st.fx_sed_deposited + st.fx_sed_oob_lost) * 2.0 ** -32
        assert abs(lhs - rhs) <= st.steps * H_LSB


def test_spawn_is_per_node_and_in_tile():
    p = orc.default_params(4)
    ls = orc.Ls(p)
    xy = ls.spawn(99, 0, 64)
    assert xy.shape == (16 * 64, 2)
    node = np.repeat(np.arange(16), 64)
    assert np.array_equal((xy[:, 0] // 512).astype(int), node // 4)  # cellpool.h:327-333 node origin
    assert np.array_equal((xy[:, 1] // 512).astype(int), node % 4)
    assert np.array_equal(xy, np.floor(xy))
    again = ls.spawn(99, 0, 64)
    other = ls.spawn(99, 1, 64)
    assert np.array_equal(xy, again) and not np.array_equal(xy, other)


def test_rejection_and_out_of_map_spawns(init_cells):
    p = orc.default_params(1)
    ls = orc.Ls(p)
    ls.upload(init_cells)
    h = orc.tiled_to_planar(p, init_cells)
    low = np.argwhere(h < 0.1)[:5].astype(np.float32)
    xy = np.concatenate([low, [[-3.0, 10.0], [600.0, 2.0], [255.0, 255.0]]]).astype(np.float32)
    drops, st = ls.make_drops(xy)
    assert st.rejected == 7 and st.spawned == 1  # world.h:71-72: height() of a missing cell is 0 < 0.1
    assert list(drops["flags"][:7]) == [orc.DROP_REJECTED] * 7 and drops["flags"][7] == orc.DROP_ALIVE


def test_track_overflow_is_reported():
    p = orc.default_params(1)
    ls = orc.Ls(p)
    cells = np.zeros(512 * 512, orc.CELL_DTYPE)
    cells["height"] = 0.5
    cells["discharge_track"][7] = 3000.0
    ls.upload(cells)
    assert ls.ema(reset=False) == 0
    cells["discharge_track"][7] = 5000.0  # beyond 4096: outside the guaranteed Q13.18 range
    ls.upload(cells)
    assert ls.ema(reset=False) == 1


def test_synth_terrain_is_normalised_and_seeded():
    a = orc.synth_terrain(256, 1)
    b = orc.synth_terrain(256, 1)
    c = orc.synth_terrain(256, 2)
    assert a.min() == 0.0 and a.max() == 1.0 and np.array_equal(a, b) and not np.array_equal(a, c)
    assert 0.3 < a.mean() < 0.7


def test_drops_on_one_cell_take_turns(init_cells):
    """Turn-taking: of the drops standing on one cell only one steps per phase; the others wait.
    The first free_waits (8) waits of a drop's life are free, every later one is a step of its life not taken.  k
    drops spawned on the same cell therefore leave it one after another, the call ends after at most
    maxAge + 2 + free_waits phases, and a drop that is alone is not affected at all."""
    p = orc.default_params(1)
    k = 5
    xy = np.tile(np.array([[256.25, 256.5]], np.float32), (k, 1))
    xy += np.linspace(0.0, 0.4, k, dtype=np.float32)[:, None]  # same cell, different positions (different keys)
    ls = orc.Ls(p)
    ls.upload(init_cells)
    drops, _ = ls.make_drops(xy)
    st, _ = ls.run_drops(drops)
    lone = orc.Ls(p)
    lone.upload(init_cells)
    d1, _ = lone.make_drops(xy[:1])
    s1, _ = lone.run_drops(d1)
    assert s1.phases == 502 and 502 < st.phases <= 502 + 8  # free waits prolong the call by at most free_waits
    assert (k - 1) * 100 < st.steps <= k * s1.steps        # every drop went its way
    # with no free waits a phase spent waiting costs a step, and the call never needs more than maxAge + 2 phases
    paid = orc.Ls(p)
    paid.upload(init_cells)
    paid.w.contents.free_waits = 0
    d0, _ = paid.make_drops(xy)
    s0, _ = paid.run_drops(d0)
    assert s0.phases == 502 and s0.steps <= k * s1.steps - (k - 1)
    assert st.term_age + st.term_vol + st.term_oob == k
    # without turn-taking all k step together in phase 0: no step is lost, but the first cell takes a k-fold hit
    free = orc.Ls(p)
    free.upload(init_cells)
    free.w.contents.exclusive_cells = 0
    d2, _ = free.make_drops(xy)
    s2, _ = free.run_drops(d2)
    assert s2.steps >= st.steps and s2.steps > s0.steps
    c0 = (256, 256)
    h0 = init_cells["height"][orc.tiled_index_map(p)[c0]]
    assert abs(free.height_q(0)[c0] * H_LSB - h0) > abs(ls.height_q(0)[c0] * H_LSB - h0) * 0.99  # no smaller hit


def test_twins_take_turns(init_cells):
    """Drops created on bit-identical positions have identical state and identical claim keys: without a tie-break
    they would step together for life, each applying the full erosion.  The k-th copy starts with
    waited = k, so the copies leave the cell one after another; the result does not depend on the list order."""
    p = orc.default_params(1)
    xy = np.array([[256.0, 256.0]] * 3 + [[100.0, 100.0], [300.0, 200.0], [100.0, 100.0]], np.float32)
    ls = orc.Ls(p)
    ls.upload(init_cells)
    drops, _ = ls.make_drops(xy)
    waited = (drops["flags"] >> 16) & 7
    assert sorted(waited[:3]) == [0, 1, 2] and sorted(waited[[3, 5]]) == [0, 1] and waited[4] == 0
    st, tr = ls.run_drops(drops.copy())
    final = drops  # run a permuted list: same maps
    ls2 = orc.Ls(p)
    ls2.upload(init_cells)
    d2, _ = ls2.make_drops(xy[::-1].copy())
    st2, _ = ls2.run_drops(d2)
    assert np.array_equal(ls.height_q(0), ls2.height_q(0)) and st.as_dict() == st2.as_dict()
    # the copies did not march as one: their tracks differ from k times a lone drop's
    lone = orc.Ls(p)
    lone.upload(init_cells)
    d1, _ = lone.make_drops(xy[:1])
    lone.run_drops(d1)
    assert not np.array_equal(ls.track_q()[200:320, 200:320, 0], 3 * lone.track_q()[200:320, 200:320, 0])


def test_turn_taking_keeps_order_independence(init_cells):
    p = orc.default_params(1)
    rng = np.random.default_rng(8)
    xy = np.repeat(rng.integers(100, 400, size=(40, 2)).astype(np.float32), 6, axis=0)  # six drops per cell
    xy += rng.uniform(0.0, 0.9, size=xy.shape).astype(np.float32)
    out = []
    for perm in (np.arange(len(xy)), rng.permutation(len(xy))):
        ls = orc.Ls(p)
        ls.upload(init_cells)
        st = ls.erode_spawnlist(xy[perm])
        out.append((ls.height_q(0).copy(), st.as_dict()))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


def test_crowd_damping_only_acts_next_to_a_higher_key(init_cells):
    """Two drops on NEIGHBOURING cells both step (no waiting), and the one with the lower key moves half as much in
    that phase; far apart, both behave like a lone drop."""
    p = orc.default_params(1)

    def first_step_delta(xy):
        ls = orc.Ls(p)
        ls.upload(init_cells)
        before = ls.height_q(0).copy()
        drops, _ = ls.make_drops(np.asarray(xy, np.float32))
        drops["age"] = 500  # one real step each: with age 501 > maxAge the next call ends the drop (water.h:74)
        st, _ = ls.run_drops(drops)
        return ls.height_q(0).astype(np.int64) - before, st

    lone_a, _ = first_step_delta([[200.5, 200.5]])
    lone_b, _ = first_step_delta([[200.5, 201.5]])
    far, st_far = first_step_delta([[200.5, 200.5], [300.5, 300.5]])
    assert st_far.steps == 4                           # one descend call each, plus the terminating one
    assert np.array_equal(far[190:210, 190:210], lone_a[190:210, 190:210])  # a distant drop changes nothing here
    near, st_near = first_step_delta([[200.5, 200.5], [200.5, 201.5]])
    assert st_near.steps == st_far.steps               # neighbours do not wait for each other
    both = lone_a + lone_b
    assert not np.array_equal(near, both)              # ... but one of them was damped
    assert np.abs(near).sum() < np.abs(both).sum()


def test_schedule_is_as_close_to_the_reference_as_a_reorder(init_cells):
    """North-star check 3 on the CPU: after 10 erode(512) cycles on the reference's default world the lock-step
    schedule (turn-taking, crowd damping, fixed point) differs from the sequential reference semantics by about as
    much as those differ from themselves with the drops of each cycle processed in another order."""
    def metrics(a, b):
        da = a["height"].astype(np.float64) - init_cells["height"]
        db = b["height"].astype(np.float64) - init_cells["height"]
        return (float(np.sqrt(np.mean((da - db) ** 2))), float(np.corrcoef(da, db)[0, 1]),
                float(np.corrcoef(a["discharge"], b["discharge"])[0, 1]))

    rng = np.random.default_rng(7)
    spawns = [rng.integers(0, 512, size=(512, 2)).astype(np.float32) for _ in range(10)]
    ref, shuf = init_cells.copy(), init_cells.copy()
    S, S2 = orc.Seq(ref), orc.Seq(shuf)
    ls = orc.Ls(orc.default_params(1))
    ls.upload(init_cells)
    for c, xy in enumerate(spawns):
        S.erode_spawnlist(xy)
        S2.erode_spawnlist(xy[np.random.default_rng(1000 + c).permutation(512)])
        ls.erode_spawnlist(xy)
    got = ls.download()
    rmse_b, corr_b, cdis_b = metrics(ref, shuf)
    rmse_l, corr_l, cdis_l = metrics(ref, got)
    assert rmse_l <= 1.25 * rmse_b and corr_l >= corr_b - 0.03 and cdis_l >= cdis_b - 0.05
    total = got["discharge"].sum(dtype=np.float64) / ref["discharge"].sum(dtype=np.float64)
    assert abs(total - 1.0) < 0.02  # waiting costs the drops next to no steps at the reference's own density
