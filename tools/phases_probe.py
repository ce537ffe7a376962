"""development aid: phases, steps and time per erode(512) call on the bench world"""
import sys, time, torch
sys.path.insert(0, ".")
import simplehydrology_b200 as shx
ms = int(sys.argv[1]) if len(sys.argv) > 1 else 16
with shx.World(mapsize=ms) as W:
    W.synth_terrain(1)
    for c in range(int(sys.argv[2]) if len(sys.argv) > 2 else 10):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = W.erode(512, 1)
        dt = time.perf_counter() - t0
        print(f"call {c}: {dt*1e3:.2f} ms phases {st.phases} steps/drop {st.steps/max(st.spawned,1):.1f} term_age {st.term_age} us/phase {dt*1e6/max(st.phases,1):.2f}", flush=True)
