"""development aid: do k logical strips of one world, each lock-stepping on its own stream, overlap on ONE GPU?
(cooperative kernels of different streams co-resident: while one strip waits at its phase barrier the other computes)"""
import sys, time, torch
sys.path.insert(0, ".")
import simplehydrology_b200 as shx
from simplehydrology_b200 import strips

MS, CYC = 16, 512
import itertools
for k, kw in itertools.product((1, 2, 4, 8), ({}, dict(block_threads=448, variant=3, coop=1), dict(block_threads=128, variant=2, coop=1))):
    if k == 1 and kw: continue
    streams = [torch.cuda.Stream() for _ in range(k)]
    bs = []
    for r in range(k):
        with torch.cuda.stream(streams[r]):
            bs.append(strips.GpuStrip(MS, r, k, 0, **kw))
            bs[-1].W.synth_terrain(1)
    torch.cuda.synchronize()
    S = strips.LocalStripSet(bs)
    def cycle():
        for b in bs:
            b.begin(CYC, 1, None)
            b.end()
    for _ in range(3):
        cycle()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        cycle()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    steps = sum(b.W.read_stats().steps for b in bs)
    print(f"k={k} {kw}: {dt*1e3:.3f} ms per cycle (no exchange), {steps/dt/1e9:.2f} G particle-steps/s", flush=True)
    for b in bs:
        b.W.close()
