"""GPU, >= 2 devices: peer mode (one world spread over the GPUs of a box, strips mapped into each
other over NVLink, one cross-GPU barrier per phase) must give the SAME BITS as the one-GPU run."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import simplehydrology_b200 as shx

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CYCLES, NCYC, SEED, TSEED = 256, 3, 5, 2


def single(ms):
    tot = {}
    with shx.World(mapsize=ms) as W:
        W.synth_terrain(TSEED)
        for _ in range(NCYC):
            W.erode(CYCLES, SEED)
            for k, v in W.read_stats().as_dict().items():
                tot[k] = max(tot.get(k, 0), v) if k == "phases" else tot.get(k, 0) + v
        hq = W.download_height_q()
        _, _, field, track = W.download_raw()
    return hq, field, track, tot


@pytest.mark.parametrize("ms,world", [(2, 2), (4, 2), (4, 4), (8, 8)])
def test_peer_mode_is_bit_identical_to_one_gpu(ms, world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + ms * 10 + world
    cmd = ["timeout", "240", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "peer_worker.py"),
           str(tmp_path), str(ms), str(CYCLES), str(NCYC), str(SEED), str(TSEED)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    parts = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    hq = np.concatenate([p["hq"] for p in parts])
    field = np.concatenate([p["field"] for p in parts])
    track = np.concatenate([p["track"] for p in parts])
    hq1, field1, track1, tot1 = single(ms)
    assert np.array_equal(hq[..., 0], hq[..., 1])
    assert np.array_equal(hq, hq1)
    assert np.array_equal(field.view(np.uint32), field1.view(np.uint32))
    assert np.array_equal(track, track1)
    names = [str(n) for n in parts[0]["names"]]
    tot = dict(zip(names, parts[0]["stats"].tolist()))
    for k in ("spawned", "steps", "term_age", "term_vol", "term_oob", "cascade_transfers", "fx_eroded", "fx_deposited", "phases"):
        assert tot[k] == tot1[k], k
    assert tot["migrated_lo"] == 0 and tot["migrated_hi"] == 0  # drops never change rank in peer mode
