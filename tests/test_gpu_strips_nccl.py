"""GPU, >= 2 devices: the N > 1 path of bench.py on real ranks.  One process per GPU runs
strips.StripExchange.erode_cycle over NCCL send/recv (tests/strip_worker.py); the protocol is deterministic for a
given number of strips, so the ranks' rows must equal, bit for bit, what the same protocol gives with the k strips
held by one process on one device (strips.LocalStripSet -- itself checked against the single domain, the integer
ledger and the reference in tests/test_gpu_strips.py / test_gpu_reference_parity.py), and what the single-host-thread
C driver shx_multi gives (tests/test_gpu_multi.py compares that one with LocalStripSet too)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from simplehydrology_b200 import strips

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
MS, CYCLES, NCYC, SEED, TSEED = 4, 256, 4, 11, 3


def one_process(k):
    bs = [strips.GpuStrip(MS, r, k, 0) for r in range(k)]
    for b in bs:
        b.W.init_terrain(TSEED)
    S = strips.LocalStripSet(bs)
    per_call = []
    for _ in range(NCYC):
        S.erode_cycle(CYCLES, SEED)
        one = [b.W.read_stats() for b in bs]
        per_call.append([sum(s.steps for s in one), sum(s.spawned for s in one), sum(s.migrated_lo + s.migrated_hi for s in one),
                         sum(s.fx_deposited - s.fx_eroded for s in one), sum(s.term_age + s.term_vol + s.term_oob for s in one)])
    hq, field, track = [], [], []
    for b in bs:
        xlo, _ = b.W.stored_rows()
        a, c = b.row0 - xlo, b.row1 - xlo
        hq.append(b.W.download_height_q()[a:c])
        _, _, f, t = b.W.download_raw()
        field.append(f[a:c])
        track.append(t[a:c])
    inflight = S.in_flight()
    for b in bs:
        b.W.close()
    return np.concatenate(hq), np.concatenate(field), np.concatenate(track), np.array(per_call, dtype=np.int64), inflight


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_ranks_equal_the_one_process_protocol(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = ["timeout", "300", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), os.path.join(HERE, "strip_worker.py"),
           str(tmp_path), str(MS), str(CYCLES), str(NCYC), str(SEED), str(TSEED)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    parts = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    assert [int(p["row0"]) for p in parts] == [k * 512 * MS // world for k in range(world)]
    hq = np.concatenate([p["hq"] for p in parts])
    field = np.concatenate([p["field"] for p in parts])
    track = np.concatenate([p["track"] for p in parts])
    per_call = sum(p["per_call"] for p in parts)
    hq1, field1, track1, per_call1, inflight1 = one_process(world)
    assert np.array_equal(per_call, per_call1)
    assert int(parts[0]["in_flight"]) == inflight1  # all-reduced over the ranks
    assert per_call[:, 2].sum() > 0                 # drops did cross strip borders
    assert np.array_equal(hq, hq1)
    assert np.array_equal(field.view(np.uint32), field1.view(np.uint32))
    assert np.array_equal(track, track1)
    # the union ledger: every height unit that appeared or vanished is accounted for once no drop is in flight is
    # checked in tests/test_gpu_multi.py; here: both planes agree after every call (catch-up complete)
    assert np.array_equal(hq[..., 0], hq[..., 1])
