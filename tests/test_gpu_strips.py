"""GPU: row strips (the multi-GPU decomposition) emulated as k logical strips on ONE device through
the same C-ABI strip calls bench.py --gpus N uses.  A 2048^2 world (BASELINE configs[2] size) is cut
into 2 and 4 strips of whole tile rows."""
import numpy as np
import pytest
import torch

import simplehydrology_b200 as shx
from simplehydrology_b200 import strips

pytestmark = pytest.mark.gpu

MS, CYCLES, NCYC, SEED = 4, 128, 3, 9


def run_strips(k):
    bs = [strips.GpuStrip(MS, r, k, 0) for r in range(k)]
    for b in bs:
        b.W.synth_terrain(2)
    before = [b.W.download_height_q() for b in bs]
    S = strips.LocalStripSet(bs)
    tot = dict(spawned=0, done=0, ledger=0, steps=0, mig=0)
    rounds = []
    for _ in range(NCYC):
        S.erode(CYCLES, SEED)
        rounds.append(S.rounds)
        for b in bs:
            st = b.W.read_stats()
            tot["spawned"] += st.spawned
            tot["done"] += st.term_age + st.term_vol + st.term_oob
            tot["ledger"] += st.fx_deposited - st.fx_eroded
            tot["steps"] += st.steps
            tot["mig"] += st.migrated_lo + st.migrated_hi
    hs, recs, dsum = [], [], 0
    for b, h0 in zip(bs, before):
        xlo, _ = b.W.stored_rows()
        a, c = b.row0 - xlo, b.row1 - xlo
        hq = b.W.download_height_q()
        assert np.array_equal(hq[..., 0], hq[..., 1])
        dsum += int(hq[a:c, :, 0].astype(np.int64).sum() - h0[a:c, :, 0].astype(np.int64).sum())
        hs.append(hq[a:c, :, 0])
        _, _, f, _ = b.W.download_raw()
        recs.append(f[a:c])
    # halos equal the owners' rows
    for i in range(k - 1):
        lo_b, hi_b = bs[i], bs[i + 1]
        h_lo, h_hi = lo_b.W.download_height_q()[..., 0], hi_b.W.download_height_q()[..., 0]
        x_lo, _ = lo_b.W.stored_rows()
        x_hi, _ = hi_b.W.stored_rows()
        edge = lo_b.row1
        assert np.array_equal(h_lo[edge - x_lo:edge - x_lo + 2], h_hi[edge - x_hi:edge - x_hi + 2])      # lo's hi-halo == hi's first rows
        assert np.array_equal(h_hi[edge - 2 - x_hi:edge - x_hi], h_lo[edge - 2 - x_lo:edge - x_lo])      # hi's lo-halo == lo's last rows
    for b in bs:
        b.W.close()
    return np.concatenate(hs), np.concatenate(recs), tot, dsum, rounds


@pytest.mark.parametrize("k", [2, 4])
def test_strips_conserve_mass_and_account_for_every_drop(k):
    h, f, tot, dsum, rounds = run_strips(k)
    assert dsum == tot["ledger"]               # exact integer ledger over the union of the strips
    assert tot["spawned"] == tot["done"]       # every drop finished somewhere after its hand-offs
    assert tot["mig"] > 0                      # drops did cross strip borders
    assert max(rounds) <= 256
    h2, f2, tot2, dsum2, rounds2 = run_strips(k)
    assert np.array_equal(h, h2) and np.array_equal(f.view(np.uint32), f2.view(np.uint32)) and tot == tot2  # deterministic
    # statistically the single-domain world
    with shx.World(mapsize=MS) as W:
        W.synth_terrain(2)
        h0 = W.download_height_q()[..., 0].astype(np.int64)
        for _ in range(NCYC):
            W.erode(CYCLES, SEED)
        h1 = W.download_height_q()[..., 0].astype(np.int64)
        _, _, f1, _ = W.download_raw()
    d1 = (h1 - h0).astype(np.float64).ravel()
    dk = (h.astype(np.int64) - h0).astype(np.float64).ravel()
    corr = np.corrcoef(d1, dk)[0, 1]
    cdis = np.corrcoef(f1[..., 0].ravel(), f[..., 0].ravel())[0, 1]
    print(f"k={k}: corr(dh) {corr:.4f} corr(discharge) {cdis:.4f} rounds {rounds} migrated {tot['mig']}")
    assert corr > 0.9 and cdis > 0.9


def test_one_strip_is_the_plain_call():
    b = strips.GpuStrip(MS, 0, 1, 0)
    b.W.synth_terrain(2)
    S = strips.LocalStripSet([b])
    with shx.World(mapsize=MS) as W:
        W.synth_terrain(2)
        for _ in range(2):
            S.erode(CYCLES, SEED)
            W.erode(CYCLES, SEED)
        a, c = b.W.download_raw(), W.download_raw()
    for x, y in zip(a, c):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    b.W.close()


def run_cycle_strips(k):
    """one exchange per call (StripExchange.erode_cycle, what bench.py --gpus N runs); carried drops flushed at the end"""
    bs = [strips.GpuStrip(MS, r, k, 0) for r in range(k)]
    for b in bs:
        b.W.synth_terrain(2)
    before = [b.W.download_height_q() for b in bs]
    S = strips.LocalStripSet(bs)
    tot = dict(spawned=0, done=0, ledger=0, steps=0, mig=0)
    carried = []

    def account():
        for b in bs:
            st = b.W.read_stats()
            tot["spawned"] += st.spawned
            tot["done"] += st.term_age + st.term_vol + st.term_oob
            tot["ledger"] += st.fx_deposited - st.fx_eroded
            tot["steps"] += st.steps
            tot["mig"] += st.migrated_lo + st.migrated_hi

    for _ in range(NCYC):
        S.erode_cycle(CYCLES, SEED)
        carried.append(S.in_flight())
        account()
    # the maps a user sees after NCYC calls (the flush calls below run extra EMAs)
    hs, fs = [], []
    for b in bs:
        xlo, _ = b.W.stored_rows()
        a, c = b.row0 - xlo, b.row1 - xlo
        hs.append(b.W.download_height_q()[a:c, :, 0])
        _, _, f, _ = b.W.download_raw()
        fs.append(f[a:c])
    flushes = 0
    while S.in_flight():
        S.erode_cycle(0, SEED)
        account()
        flushes += 1
        assert flushes < 64
    dsum = 0
    for b, h0 in zip(bs, before):
        xlo, _ = b.W.stored_rows()
        a, c = b.row0 - xlo, b.row1 - xlo
        hq = b.W.download_height_q()
        assert np.array_equal(hq[..., 0], hq[..., 1])
        dsum += int(hq[a:c, :, 0].astype(np.int64).sum() - h0[a:c, :, 0].astype(np.int64).sum())
    for b in bs:
        b.W.close()
    return np.concatenate(hs), np.concatenate(fs), tot, dsum, carried, flushes


@pytest.mark.parametrize("k", [2, 4])
def test_cycle_exchange_conserves_mass_and_stays_close_to_one_domain(k):
    h, f, tot, dsum, carried, flushes = run_cycle_strips(k)
    assert dsum == tot["ledger"]               # exact integer ledger over the union of the strips
    assert tot["spawned"] == tot["done"]       # every drop finished somewhere
    assert tot["mig"] > 0 and carried[0] > 0
    assert max(carried) < 0.5 * MS * MS * CYCLES   # the carried population stays a fraction of a call's batch
    h2, f2, tot2, dsum2, carried2, _ = run_cycle_strips(k)
    assert np.array_equal(h, h2) and np.array_equal(f.view(np.uint32), f2.view(np.uint32)) and tot == tot2  # deterministic
    with shx.World(mapsize=MS) as W:
        W.synth_terrain(2)
        h0 = W.download_height_q()[..., 0].astype(np.int64)
        for _ in range(NCYC):
            W.erode(CYCLES, SEED)
        h1 = W.download_height_q()[..., 0].astype(np.int64)
        _, _, f1, _ = W.download_raw()
    d1 = (h1 - h0).astype(np.float64).ravel()
    dk = (h.astype(np.int64) - h0).astype(np.float64).ravel()
    corr = np.corrcoef(d1, dk)[0, 1]
    cdis = np.corrcoef(f1[..., 0].ravel(), f[..., 0].ravel())[0, 1]
    print(f"cycle k={k}: corr(dh) {corr:.4f} corr(discharge) {cdis:.4f} carried {carried} flushes {flushes} migrated {tot['mig']}")
    # Stated bound: 0.85.  After only NCYC calls from a fresh world up to a third of a call's drops are
    # still waiting at a border (512-row strips on a 2048^2 world; measured 0.96 / 0.88 for k = 2 / 4,
    # the exchange-round protocol above gives 0.96 / 0.91)
    assert corr > 0.85 and cdis > 0.85
