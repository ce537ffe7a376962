"""Ad-hoc first-light check on a GPU box (development aid, not a test): parity of both modes
against the oracles plus rough timings.  Run: gpurun -- python tools/gpu_first_check.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
import simplehydrology_b200 as shx  # noqa: E402


def section(t):
    print(f"\n=== {t}", flush=True)


def main():
    R = orc.Ref(1, seed=1)
    init = R.cells.copy()
    p = orc.default_params(1)

    section("sequential mode vs compiled reference (fresh map: erf sees 0 -> bit exact expected)")
    W = shx.World(mapsize=1, mode=shx.MODE_SEQUENTIAL)
    W.upload(init)
    back = W.download()
    print("upload/download roundtrip exact:", np.array_equal(back.view(np.uint8), init.view(np.uint8)))
    for (x, y) in [(256.0, 256.0), (100.0, 300.0), (400.5, 3.25), (0.0, 0.0), (511.0, 511.0)]:
        a = R.trace_drop(x, y)
        b = W.trace_drop(x, y)
        same = a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
        print((x, y), len(a), len(b), "bit-exact" if same else "DIFF",
              "" if same else np.abs(a[:min(len(a), len(b))] - b[:min(len(a), len(b))]).max())
    g = W.download()
    print("cells after traces exact:", np.array_equal(g.view(np.uint8), R.cells.view(np.uint8)),
          np.abs(g["height"] - R.cells["height"]).max())
    rng = np.random.default_rng(7)
    for c in range(3):
        xy = rng.integers(0, 512, size=(512, 2)).astype(np.float32)
        sr = R.erode_spawnlist(xy)
        t = time.time()
        st = W.erode_spawnlist(xy)
        dt = time.time() - t
        g = W.download()
        print(c, "ref steps", sr["steps"], "gpu steps", st.steps, "exact:", np.array_equal(g.view(np.uint8), R.cells.view(np.uint8)),
              "max|dh|", np.abs(g["height"] - R.cells["height"]).max(), "max|ddis|", np.abs(g["discharge"] - R.cells["discharge"]).max(),
              f"{dt*1e3:.1f} ms")
    W.close()

    section("batched mode vs lock-step oracle (bit exact expected)")
    W = shx.World(mapsize=1, mode=shx.MODE_BATCHED)
    W.upload(init)
    ls = orc.Ls(p)
    ls.upload(init)
    h0, h1, f, t = W.download_raw()
    print("upload planes exact:", np.array_equal(h0, ls.height_q(0)), np.array_equal(h1, ls.height_q(1)))
    a = W.trace_drop(256.0, 256.0)
    drops, _ = ls.make_drops(np.array([[256.0, 256.0]], np.float32))
    _, b = ls.run_drops(drops, trace_cap=1024)
    print("single drop trace", len(a), len(b), np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    h0, h1, f, t = W.download_raw()
    print("planes after trace:", np.array_equal(h0, ls.height_q(0)), np.array_equal(h1, ls.height_q(1)),
          "tracks:", np.array_equal(t, ls.track_q()))
    rng = np.random.default_rng(7)
    for c in range(4):
        xy = rng.integers(0, 512, size=(512, 2)).astype(np.float32)
        so = ls.erode_spawnlist(xy)
        tt = time.time()
        st = W.erode_spawnlist(xy)
        dt = time.time() - tt
        h0, h1, f, t = W.download_raw()
        keys = ["spawned", "rejected", "steps", "term_age", "term_vol", "term_oob", "cascade_transfers", "phases",
                "fx_eroded", "fx_deposited", "fx_sed_oob_lost", "fx_sed_deposited", "fx_sed_inflation"]
        stat_ok = all(getattr(st, k) == getattr(so, k) for k in keys)
        print(c, "planes:", np.array_equal(h0, ls.height_q(0)), np.array_equal(h1, ls.height_q(1)),
              "field:", np.array_equal(f.view(np.uint32), ls.field().view(np.uint32)), "tracks:", np.array_equal(t, ls.track_q()),
              "stats:", stat_ok, f"{dt*1e3:.2f} ms", "launches", st.launches)
        if not stat_ok:
            print("   gpu", {k: getattr(st, k) for k in keys})
            print("   orc", {k: getattr(so, k) for k in keys})
    # erode() with the hash spawn
    xy_gpu = W.spawn(512, 1234, 0)
    xy_orc = ls.spawn(1234, 0, 512)
    print("spawn positions equal:", np.array_equal(xy_gpu, xy_orc))
    W.close()

    section("timings (device-side loop, async erode)")
    import torch
    for ms in (1, 4, 16):
        W = shx.World(mapsize=ms, mode=shx.MODE_BATCHED)
        W.set_stream(torch.cuda.current_stream().cuda_stream)
        W.synth_terrain(1)
        for _ in range(3):
            W.erode_async(512, 1)
        W.sync()
        n = 10 if ms < 16 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 0
        e0.record()
        for i in range(n):
            W.erode_async(512, 1)
        e1.record()
        torch.cuda.synchronize()
        ms_cycle = e0.elapsed_time(e1) / n
        st = W.erode(512, 1)
        print(f"mapsize {ms}: {ms_cycle:.3f} ms/cycle, last cycle steps {st.steps} phases {st.phases} "
              f"-> {st.steps / (ms_cycle * 1e-3) / 1e9:.3f} G steps/s, {st.steps * 88 / (ms_cycle * 1e-3) / 1e9:.1f} GB/s algorithmic")
        W.close()


if __name__ == "__main__":
    main()
