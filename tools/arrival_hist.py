"""Development aid: distribution of per-warp arrival times at the barrier of one phase (phase 100)
of the descend kernel -- shows whether a phase is dominated by a few stragglers or a broad convoy.
usage: python tools/arrival_hist.py MAPSIZE block:variant:grid"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplehydrology_b200 as shx  # noqa: E402
from simplehydrology_b200 import build as B  # noqa: E402

lib = os.path.join(ROOT, "gpurun_out", "libshx_timing.so")
os.makedirs(os.path.dirname(lib), exist_ok=True)
subprocess.run([B.nvcc()] + B.NVCC_FLAGS + ["-DSHX_PHASE_TIMING", "-o", lib] + B.SOURCES, check=True)
shx.LIB_PATH = lib
L = shx.lib()
ms = int(sys.argv[1])
for cfg in sys.argv[2:]:
    b, v, g, co = (int(x) for x in (cfg.split(":") + ["0"])[:4])
    W = shx.World(mapsize=ms, block_threads=b, variant=v, grid_blocks=g, coop=co)
    W.synth_terrain(1)
    for _ in range(5):
        W.erode(512, 1)
    ns = np.zeros(8192, np.uint64)
    sm = np.zeros(8192, np.uint32)
    L.shx_debug_arrivals(ns.ctypes.data, sm.ctypes.data)
    st = np.zeros(8192, np.uint64)
    L.shx_debug_starts(st.ctypes.data)
    nw = min(8192, (ms * ms * 512 + 31) // 32)
    t0 = float(st[:nw].min())
    s0 = st[:nw].astype(np.float64) - t0
    t = ns[:nw].astype(np.float64) - t0
    pct = [0, 10, 25, 50, 75, 90, 99, 100]
    print(f"mapsize {ms} cfg {cfg}: warps {nw}  (us after the first warp left the previous barrier; p0 p10 p25 p50 p75 p90 p99 p100)")
    print("   phase start :", " ".join(f"{x/1e3:6.1f}" for x in np.percentile(s0, pct)))
    print("   arrival     :", " ".join(f"{x/1e3:6.1f}" for x in np.percentile(t, pct)))
    print("   own duration:", " ".join(f"{x/1e3:6.1f}" for x in np.percentile(t - s0, pct)))
    # per SM: when does its last warp arrive
    s = sm[:nw]
    last = np.array([t[s == i].max() for i in np.unique(s)])
    first = np.array([t[s == i].min() for i in np.unique(s)])
    cnt = np.array([(s == i).sum() for i in np.unique(s)])
    print(f"   per-SM last arrival (us): min {last.min()/1e3:.1f} median {np.median(last)/1e3:.1f} max {last.max()/1e3:.1f};"
          f" first arrival median {np.median(first)/1e3:.1f}; warps per SM {cnt.min()}..{cnt.max()} (only the first 8192 warps are recorded)")
    W.close()
