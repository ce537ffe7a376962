"""Generate tests/golden/reference_seed1.npz from the reference's own code (oracle/_ref).

Run in the authoring container, where /root/reference exists and oracle/_ref has been built
(`make -C oracle`).  The reference ships no golden vectors of its own; these are
outputs of its unmodified headers (world.h, water.h, cellpool.h) compiled headless, and they pin
the CPU restatement in oracle/shx_oracle.c on machines without the reference tree.

Every block runs in a fresh interpreter because the reference's world is process-global and
map::init is only faithful once per process.
"""
import ctypes
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))

TRACE_STARTS = [(256.0, 256.0), (100.0, 300.0), (400.5, 3.25), (0.0, 0.0), (511.0, 511.0), (10.0, 500.0), (37.75, 129.5)]
NORMAL_CELLS = [(0, 0), (0, 5), (511, 511), (511, 0), (3, 511), (200, 200), (17, 340), (0, 511)]
SPAWN_SEED = 7
N_CYCLES = 4


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


def spawn_lists():
    rng = np.random.default_rng(SPAWN_SEED)
    return [rng.integers(0, 512, size=(512, 2)).astype(np.float32) for _ in range(N_CYCLES)]


def libc_rand_spawns(seed, n):
    """world.h:69 with g++ 13: ivec2(rand()%512, rand()%512) evaluates the SECOND argument first"""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(seed)
    xy = np.zeros((n, 2), np.float32)
    for i in range(n):
        y = libc.rand() % 512
        x = libc.rand() % 512
        xy[i] = (x, y)
    return xy


def part_world():
    import orc
    R = orc.Ref(1, seed=1)
    out = {"init_height": R.cells["height"].copy()}
    out["normals"] = np.stack([R.normal(x, y) for x, y in NORMAL_CELLS])
    traces, lens = [], []
    for x, y in TRACE_STARTS:
        t = R.trace_drop(x, y)
        lens.append(len(t))
        traces.append(np.pad(t, ((0, 512 - len(t)), (0, 0))))
    out["trace_starts"] = np.array(TRACE_STARTS, np.float32)
    out["trace_len"] = np.array(lens, np.int32)
    out["traces"] = np.stack(traces)
    changed = np.nonzero(R.cells["height"] != out["init_height"])[0]
    out["after_traces_idx"] = changed.astype(np.int32)
    out["after_traces_height"] = R.cells["height"][changed].copy()
    tracked = np.nonzero(R.cells["discharge_track"] != 0)[0]
    out["after_traces_track_idx"] = tracked.astype(np.int32)
    out["after_traces_tracks"] = np.stack([R.cells[f][tracked] for f in ("discharge_track", "momentumx_track", "momentumy_track")], 1)
    out["after_traces_sha"] = np.frombuffer(bytes.fromhex(sha(R.cells)), np.uint8)
    return out


def part_cycles():
    import orc
    R = orc.Ref(1, seed=1)
    shas, steps = [], []
    for xy in spawn_lists():
        st = R.erode_spawnlist(xy)
        shas.append(np.frombuffer(bytes.fromhex(sha(R.cells)), np.uint8))
        steps.append([st["spawned"], st["rejected"], st["steps"]])
    sample = np.arange(0, 512 * 512, 997)
    return {"cycles_sha": np.stack(shas), "cycles_stats": np.array(steps, np.int64),
            "cycles_sample_idx": sample, "cycles_sample_cells": R.cells[sample].copy().view(np.float32).reshape(-1, 8)}


def part_stock_erode():
    """the stock call World::erode(512) with srand(1), 3 frames (SimpleHydrology.cpp:27-30,319)"""
    import orc
    R = orc.Ref(1, seed=1)  # ref_init does srand(seed) like main()
    shas = []
    for _ in range(3):
        R.erode(512)
        shas.append(np.frombuffer(bytes.fromhex(sha(R.cells)), np.uint8))
    return {"stock_erode_sha": np.stack(shas)}


def part_cascade_kat():
    """World::cascade on hand-made 9x9 patches written into an otherwise flat blank world"""
    import orc
    R = orc.Ref(1)  # blank
    rng = np.random.default_rng(11)
    p = orc.default_params(1)
    idx = orc.tiled_index_map(p)
    cases, outs = [], []
    for c in range(24):
        R.cells["height"][:] = 0.5
        base = 0.05 if c % 4 == 0 else 0.5  # some patches straddle the 0.1 threshold (world.h:144)
        patch = (base + rng.normal(0, 0.03 if c % 3 else 0.2, size=(9, 9))).astype(np.float32)
        if c % 5 == 0:
            patch[4, 4] = patch[3, 4]  # exact tie with a neighbour
        corner = c % 6 == 5
        x0, y0 = (0, 0) if corner else (100, 200)
        for i in range(9):
            for j in range(9):
                R.cells["height"][idx[x0 + i, y0 + j]] = patch[i, j]
        cx, cy = (0.5, 0.25) if corner else (x0 + 4.5, y0 + 4.75)
        R.cascade(cx, cy)
        after = np.array([[R.cells["height"][idx[x0 + i, y0 + j]] for j in range(9)] for i in range(9)], np.float32)
        cases.append(np.concatenate([[x0, y0, cx, cy], patch.ravel()]).astype(np.float32))
        outs.append(after.ravel())
    return {"cascade_cases": np.stack(cases), "cascade_after": np.stack(outs)}


def part_vertices():
    """quad::updatenode (cellpool.h:286-305) through the reference's Vertexpool::fill: fresh map and after one cycle"""
    import orc
    R = orc.Ref(1, seed=1)
    rng = np.random.default_rng(77)
    sample = np.concatenate([[0, 1, 511, 512, 513, 512 * 511, 512 * 512 - 1, 512 * 256 + 255],
                             rng.integers(0, 512 * 512, 248)]).astype(np.int64)
    out = {"vertex_sample_idx": sample}
    for tag in ("fresh", "eroded"):
        v = R.vertices() + np.float32(0.0)  # -0 -> +0: only the sign of a zero is not pinned
        out[f"vertex_{tag}_sha"] = np.frombuffer(bytes.fromhex(sha(v)), np.uint8)
        out[f"vertex_{tag}_sample"] = v[sample].copy()
        R.erode_spawnlist(spawn_lists()[0])
    return out


PARTS = {"vertices": part_vertices, "world": part_world, "cycles": part_cycles, "stock": part_stock_erode, "cascade": part_cascade_kat}

if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--part":
        res = PARTS[sys.argv[2]]()
        np.savez(os.path.join(HERE, f"_part_{sys.argv[2]}.npz"), **res)
        sys.exit(0)
    merged = {}
    for name in PARTS:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--part", name], check=True, cwd=ROOT)
        f = os.path.join(HERE, f"_part_{name}.npz")
        with np.load(f) as z:
            merged.update({k: z[k] for k in z.files})
        os.remove(f)
    merged["stock_spawns"] = libc_rand_spawns(1, 3 * 512)
    merged["spawn_lists"] = np.stack(spawn_lists())
    np.savez_compressed(os.path.join(HERE, "reference_seed1.npz"), **merged)
    meta = {k: [list(v.shape), str(v.dtype)] for k, v in merged.items()}
    with open(os.path.join(HERE, "reference_seed1.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (reference headers, g++ -O2 -ffp-contract=off)",
                   "arrays": meta}, fh, indent=1)
    print("wrote reference_seed1.npz", os.path.getsize(os.path.join(HERE, "reference_seed1.npz")), "bytes")
