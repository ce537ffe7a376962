"""Development aid: per-section cycle counts of one phase (thread 0 of block 0), using a build of the
library with -DSHX_PHASE_TIMING.  usage: python tools/phase_timing.py MAPSIZE [block:variant:grid ...]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplehydrology_b200 as shx  # noqa: E402
from simplehydrology_b200 import build as B  # noqa: E402

lib = os.path.join(ROOT, "gpurun_out", "libshx_timing.so")
os.makedirs(os.path.dirname(lib), exist_ok=True)
subprocess.run([B.nvcc()] + B.NVCC_FLAGS + ["-DSHX_PHASE_TIMING", "-o", lib] + B.SOURCES, check=True)
shx.LIB_PATH = lib
L = shx.lib()
names = ["catchup+loads+sts", "cascade", "math(+h2 load)", "atomics+events", "syncthreads", "grid barrier"]
ms = int(sys.argv[1])
for cfg in sys.argv[2:] or ["0:0:0:0"]:
    b, v, g, exp = (int(x) for x in (cfg.split(":") + ["0"])[:4])
    L.shx_debug_set_exp(exp)
    W = shx.World(mapsize=ms, block_threads=b, variant=v, grid_blocks=g)
    W.synth_terrain(1)
    for _ in range(4):
        W.erode(512, 1)
    out = (C.c_ulonglong * 8)()
    L.shx_debug_phase_timing(out, 1)
    import time
    W.sync()
    t0 = time.time()
    for _ in range(4):
        W.erode(512, 1)
    dt = (time.time() - t0) / 4
    L.shx_debug_phase_timing(out, 1)
    n = max(out[7], 1)
    print(f"mapsize {ms} cfg {cfg}: phases sampled {n}; {dt*1e3:.2f} ms/cycle (host timed)")
    for i in range(6):
        print(f"   {names[i]:16s} {out[i]/n:9.0f} cycles")
    print(f"   {'total':16s} {sum(out[:6])/n:9.0f} cycles")
    W.close()
