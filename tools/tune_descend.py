"""Development aid: time erode(512) for several descend-kernel launch shapes.
usage: python tools/tune_descend.py MAPSIZE [block:variant:grid[:coop[:cycles]] ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplehydrology_b200 as shx  # noqa: E402

if os.environ.get("SHX_LIB"):  # tools/variants.py: a build with extra -D switches
    shx.LIB_PATH = os.environ["SHX_LIB"]


def run(ms, block, variant, grid, coop=0, cycles=512, warm=4, n=6):
    W = shx.World(mapsize=ms, block_threads=block, variant=variant, grid_blocks=grid, coop=coop)
    W.set_stream(torch.cuda.current_stream().cuda_stream)
    W.synth_terrain(1)
    for _ in range(warm):
        W.erode_async(cycles, 1)
    W.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        W.erode_async(cycles, 1)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n
    st = W.erode(cycles, 1)
    W.close()
    print(f"mapsize {ms} block {block} variant {variant} grid {grid} coop {coop} cycles {cycles}: {t:.3f} ms/cycle  {t*1e3/max(st.phases,1):.2f} us/phase  "
          f"{st.steps/(t*1e-3)/1e9:.3f} Gsteps/s  launches {st.launches}", flush=True)


if __name__ == "__main__":
    ms = int(sys.argv[1])
    cfgs = sys.argv[2:] or ["256:0:0"]
    for c in cfgs:
        f = [int(x) for x in c.split(":")]
        f += [0, 0, 0, 0, 512][len(f):]  # block:variant:grid:coop:cycles
        run(ms, f[0], f[1], f[2], f[3], cycles=f[4])
