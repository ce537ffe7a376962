"""GPU: shx_multi, the single-host-thread multi-GPU driver of the C ABI (one strip context per device, one
exchange per World::erode call by peer stores into the neighbours' inboxes), against the Python orchestration of
the same protocol (strips.LocalStripSet.erode_cycle, itself checked against the single domain and the reference in
tests/test_gpu_strips.py / test_gpu_reference_parity.py).  Both are deterministic for a given number of strips, so
they must agree bit for bit.  With one visible GPU the strips are logical strips on that GPU; with more, strip i
lives on GPU i and the messages cross NVLink."""
import numpy as np
import pytest
import torch

import simplehydrology_b200 as shx
from simplehydrology_b200 import strips

pytestmark = pytest.mark.gpu

MS, CYCLES, NCALLS, SEED = 4, 256, 4, 11


def python_reference(k):
    bs = [strips.GpuStrip(MS, r, k, 0) for r in range(k)]
    for b in bs:
        b.W.init_terrain(3)
    S = strips.LocalStripSet(bs)
    stats = []
    for _ in range(NCALLS):
        S.erode_cycle(CYCLES, SEED)
        one = [b.W.read_stats() for b in bs]
        stats.append((sum(s.steps for s in one), sum(s.spawned for s in one), sum(s.migrated_lo + s.migrated_hi for s in one),
                      sum(s.fx_deposited - s.fx_eroded for s in one)))
    inflight = S.in_flight()
    pool = np.zeros(bs[0].size ** 2, shx.CELL_DTYPE)
    for b in bs:
        b.W.download(out=pool)
    for b in bs:
        b.W.close()
    return pool, stats, inflight


def through_shx_multi(k, devices):
    with shx.MultiWorld(mapsize=MS, ngpu=k, devices=devices) as M:
        M.init_terrain(3)
        stats = []
        for _ in range(NCALLS):
            st = M.erode(CYCLES, SEED)
            stats.append((st.steps, st.spawned, st.migrated_lo + st.migrated_hi, st.fx_deposited - st.fx_eroded))
        inflight = M.in_flight()
        pool = M.download()
    return pool, stats, inflight


@pytest.mark.parametrize("k", [1, 2, 4])
def test_single_thread_driver_equals_the_python_protocol(k):
    ndev = torch.cuda.device_count()
    devices = [i % ndev for i in range(k)]  # real peers where the box has them, logical strips otherwise
    got, st_got, fl_got = through_shx_multi(k, devices)
    if k == 1:
        with shx.World(mapsize=MS) as W:
            W.init_terrain(3)
            st_want = []
            for _ in range(NCALLS):
                st = W.erode(CYCLES, SEED)
                st_want.append((st.steps, st.spawned, 0, st.fx_deposited - st.fx_eroded))
            want, fl_want = W.download(), 0
    else:
        want, st_want, fl_want = python_reference(k)
    assert st_got == st_want
    assert fl_got == fl_want
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    if k > 1:
        assert sum(s[2] for s in st_got) > 0  # drops did cross strip borders


def test_ledger_closes_over_the_union_of_the_strips():
    """sum(height after) - sum(height before) == fx_deposited - fx_eroded over all strips, exactly, once no drop is
    in flight (a drop that crossed a border is finished by the neighbour in the next call)"""
    ndev = torch.cuda.device_count()
    k = 4
    with shx.MultiWorld(mapsize=MS, ngpu=k, devices=[i % ndev for i in range(k)]) as M:
        M.synth_terrain(2)
        before = sum(int(s.download_height_q()[s.cfg.row0 - s.stored_rows()[0]:s.cfg.row1 - s.stored_rows()[0], :, 0].astype(np.int64).sum())
                     for s in M.strips)
        ledger = spawned = done = 0
        for call in range(600):  # a drop that zig-zags along a border crosses once per call: draining takes tens of calls
            st = M.erode(CYCLES if call < 3 else 0, SEED)  # then drain what is still crossing borders
            ledger += st.fx_deposited - st.fx_eroded
            spawned += st.spawned
            done += st.term_age + st.term_vol + st.term_oob
            if call >= 3 and M.in_flight() == 0:
                break
        assert M.in_flight() == 0 and spawned == done
        after = sum(int(s.download_height_q()[s.cfg.row0 - s.stored_rows()[0]:s.cfg.row1 - s.stored_rows()[0], :, 0].astype(np.int64).sum())
                    for s in M.strips)
    assert after - before == ledger
