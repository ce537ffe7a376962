"""GPU: parity AT the benchmarked configurations (BASELINE configs[2] and [3]: 2048^2 and 8192^2, erode(512), the
reference's own terrain, the launch shape the library picks by itself).

 * bit for bit against the CPU lock-step oracle (orc_ls_*: the reference's per-step arithmetic under the same
   schedule) -- height planes, fields, tracks and every counter, two consecutive calls;
 * check (3) against the reference's OWN loop (oracle/_ref: Drop::descend / World::cascade compiled from its headers)
   at the full density of SimpleHydrology.cpp:319 (512 drops per node and call): 40 calls at 2048^2 and 2 calls at
   8192^2, bounded by the reorder baseline (the reference against itself with each call's drops in another order);
 * the bias of the schedule (queueing costs steps): steps per drop and total discharge within 2 % of the reference's.

The reference runs in child processes (its world is process-global) started when this module is first used, so they
overlap the oracle's own 8192^2 run."""
import os
import subprocess
import sys

import numpy as np
import pytest

import orc
import simplehydrology_b200 as shx

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 31
STAT_KEYS = ["spawned", "rejected", "steps", "term_age", "term_vol", "term_oob", "cascade_transfers", "phases",
             "fx_eroded", "fx_deposited", "fx_sed_oob_lost", "fx_sed_deposited", "fx_sed_inflation"]
have_refs = orc.have_ref(4) and orc.have_ref(16)


@pytest.fixture(scope="module")
def terrain16(tmp_path_factory):
    """map::init at 8192^2 (the oracle's restatement, OpenMP), shared by the oracle run and the reference children"""
    path = tmp_path_factory.mktemp("terrain") / "init16.npy"
    h = orc.init_terrain(16, 1)
    np.save(path, h)
    return str(path), h


@pytest.fixture(scope="module")
def ref_jobs(tmp_path_factory, terrain16):
    if not have_refs:
        return None
    d = tmp_path_factory.mktemp("refjobs")
    env = dict(os.environ, PYTHONPATH=HERE + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="2")
    jobs = {}
    for name, args in (("m4_ref", ["4", "40", "20,40", "0"]), ("m4_shuf", ["4", "40", "20,40", "1"]),
                       ("m16_ref", ["16", "2", "2", "0"]), ("m16_shuf", ["16", "2", "2", "1"])):
        out = str(d / name)
        cmd = [sys.executable, os.path.join(HERE, "ref_job.py")] + args + [str(SEED), out] + ([terrain16[0]] if name.startswith("m16") else [])
        jobs[name] = (subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True), out)
    return jobs


def wait_job(jobs, name):
    pr, out = jobs[name]
    so, se = pr.communicate(timeout=1500)
    assert pr.returncode == 0, se
    return out


@pytest.mark.parametrize("mapsize", [4, 8, 16])
def test_default_launch_is_bit_exact_at_benchmark_size(mapsize, terrain16, ref_jobs):
    p = orc.default_params(mapsize)
    h = terrain16[1] if mapsize == 16 else orc.init_terrain(mapsize, 1)
    ls = orc.Ls(p)
    ls.upload(orc.planar_to_tiled(p, h))
    with shx.World(mapsize=mapsize) as W:
        W.init_terrain(1)
        for epoch in range(2):
            st = W.erode(512, seed=SEED)
            so = ls.erode(512, SEED, epoch)
            for k in STAT_KEYS:
                assert getattr(st, k) == getattr(so, k), (k, epoch)
        grid, block, lanes = W.launch_info()
        h0, h1, f, t = W.download_raw()
    # the shapes bench.py times, all one thread per drop: 8192^2 -> 293 CTAs of 448; 4096^2 (the batch of one strip of
    # a four-GPU run) -> 256 CTAs of 128; 2048^2 -> 128 CTAs of 64.  (Eight lanes per drop serve batches of up to 6 144
    # drops: tests/test_gpu_batched.py and the 512^2 tests.)
    assert (grid, block, lanes) == {16: (293, 448, 1), 8: (256, 128, 1), 4: (128, 64, 1)}[mapsize]
    assert np.array_equal(h0, ls.height_q(0)) and np.array_equal(h1, ls.height_q(1))
    assert np.array_equal(f.view(np.uint32), ls.field().view(np.uint32))
    assert np.array_equal(t[..., :3], ls.track_q()[..., :3])
    assert st.phases <= 502 + 8 and st.steps > 400 * st.spawned


def _metrics(ref_h, ref_d, got_h, got_d, init_h):
    da, db = ref_h.astype(np.float64) - init_h, got_h.astype(np.float64) - init_h
    return (float(np.sqrt(np.mean((da - db) ** 2))), float(np.corrcoef(da, db)[0, 1]), float(np.corrcoef(ref_d, got_d)[0, 1]),
            float(got_d.sum(dtype=np.float64) / ref_d.sum(dtype=np.float64)))


@pytest.mark.skipif(not have_refs, reason="oracle/_ref not built (no reference tree at build time)")
def test_check3_full_density_2048_over_40_calls(ref_jobs):
    """the river-forming regime: rivers exist after ~20 calls, queues and crowd damping act from then on"""
    p = orc.default_params(4)
    init_h = orc.planar_to_tiled(p, orc.init_terrain(4, 1))["height"].astype(np.float64)
    got, steps, spawned = {}, 0, 0
    with shx.World(mapsize=4) as W:
        W.init_terrain(1)
        for c in range(40):
            st = W.erode(512, seed=SEED)
            steps += st.steps
            spawned += st.spawned
            if c + 1 in (20, 40):
                cells = W.download()
                got[c + 1] = (cells["height"].copy(), cells["discharge"].copy())
        assert W.launch_info() == (128, 64, 1)
    ref, shuf = wait_job(ref_jobs, "m4_ref"), wait_job(ref_jobs, "m4_shuf")
    for cp in (20, 40):
        rh, rd = np.load(f"{ref}_h{cp}.npy"), np.load(f"{ref}_d{cp}.npy")
        sh, sd = np.load(f"{shuf}_h{cp}.npy"), np.load(f"{shuf}_d{cp}.npy")
        rmse_b, corr_b, cdis_b, tot_b = _metrics(rh, rd, sh, sd, init_h)
        rmse_g, corr_g, cdis_g, tot_g = _metrics(rh, rd, got[cp][0], got[cp][1], init_h)
        print(f"2048^2 after {cp} calls: reorder baseline rmse {rmse_b:.6f} corr {corr_b:.4f} cdis {cdis_b:.4f} total {tot_b:.4f} | "
              f"gpu rmse {rmse_g:.6f} corr {corr_g:.4f} cdis {cdis_g:.4f} total {tot_g:.4f}")
        assert rmse_g <= 1.15 * rmse_b and corr_g >= corr_b - 0.03 and cdis_g >= cdis_b - 0.05
        assert abs(tot_g - 1.0) < 0.02  # total discharge within 2 % of the reference's
    rs = np.load(f"{ref}_steps.npy")
    ref_spd, gpu_spd = rs[0].sum() / rs[1].sum(), steps / spawned
    print(f"steps per drop over 40 calls: reference {ref_spd:.1f}, gpu {gpu_spd:.1f}")
    assert abs(gpu_spd / ref_spd - 1.0) < 0.02  # the price of waiting, bounded (free_waits = 8)


@pytest.mark.skipif(not have_refs, reason="oracle/_ref not built (no reference tree at build time)")
def test_check3_at_8192_two_calls(ref_jobs, terrain16):
    p = orc.default_params(16)
    init_h = orc.planar_to_tiled(p, terrain16[1])["height"].astype(np.float64)
    with shx.World(mapsize=16) as W:
        W.init_terrain(1)
        steps = spawned = 0
        for c in range(2):
            st = W.erode(512, seed=SEED)
            steps += st.steps
            spawned += st.spawned
        assert W.launch_info() == (293, 448, 1)
        cells = W.download()
    ref, shuf = wait_job(ref_jobs, "m16_ref"), wait_job(ref_jobs, "m16_shuf")
    rh, rd = np.load(f"{ref}_h2.npy"), np.load(f"{ref}_d2.npy")
    sh, sd = np.load(f"{shuf}_h2.npy"), np.load(f"{shuf}_d2.npy")
    rmse_b, corr_b, cdis_b, tot_b = _metrics(rh, rd, sh, sd, init_h)
    rmse_g, corr_g, cdis_g, tot_g = _metrics(rh, rd, cells["height"], cells["discharge"], init_h)
    rs = np.load(f"{ref}_steps.npy")
    print(f"8192^2 after 2 calls: reorder baseline rmse {rmse_b:.6f} corr {corr_b:.4f} cdis {cdis_b:.4f} | gpu rmse {rmse_g:.6f} corr {corr_g:.4f} "
          f"cdis {cdis_g:.4f} total {tot_g:.4f}; steps/drop reference {rs[0].sum() / rs[1].sum():.1f} gpu {steps / spawned:.1f}")
    assert rmse_g <= 1.15 * rmse_b and corr_g >= corr_b - 0.03 and cdis_g >= cdis_b - 0.05
    assert abs(tot_g - 1.0) < 0.02 and abs((steps / spawned) / (rs[0].sum() / rs[1].sum()) - 1.0) < 0.02
