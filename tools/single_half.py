"""development aid: one-GPU cycle time of the 8192^2 world at N-th of the drops (what one rank of an N-GPU run marches)"""
import sys, torch
sys.path.insert(0, ".")
import simplehydrology_b200 as shx
W = shx.World(mapsize=16)
W.synth_terrain(1)
for cyc in (512, 256, 128, 64):
    for _ in range(3):
        W.erode(cyc, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    W.set_stream(torch.cuda.current_stream().cuda_stream)
    e0.record()
    for _ in range(5):
        W.erode_async(cyc, 1)
    e1.record(); torch.cuda.synchronize()
    print("drops", 256 * cyc, "ms/cycle", e0.elapsed_time(e1) / 5)
