"""One rank of a row-strip run over NCCL (launched by tests/test_gpu_strips_nccl.py through torch.distributed.run):
what `bench.py --gpus N` runs per step -- strips.StripExchange.erode_cycle on a strips.GpuStrip -- for a few calls,
then stores the rows this rank owns."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from simplehydrology_b200 import strips  # noqa: E402


def main():
    out, ms, cycles, ncyc, seed, tseed = sys.argv[1], *(int(v) for v in sys.argv[2:7])
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b = strips.GpuStrip(ms, rank, world, local)
    ex = strips.StripExchange(b, rank, world)
    b.W.init_terrain(tseed)
    dist.barrier()
    torch.cuda.synchronize()
    per_call = []
    for _ in range(ncyc):
        ex.erode_cycle(cycles, seed)
        st = b.W.read_stats()
        per_call.append([st.steps, st.spawned, st.migrated_lo + st.migrated_hi, st.fx_deposited - st.fx_eroded,
                         st.term_age + st.term_vol + st.term_oob])
    xlo, _ = b.W.stored_rows()
    a, c = b.row0 - xlo, b.row1 - xlo
    hq = b.W.download_height_q()[a:c]
    _, _, field, track = b.W.download_raw()
    np.savez(os.path.join(out, f"rank{rank}.npz"), hq=hq, field=field[a:c], track=track[a:c], row0=b.row0, row1=b.row1,
             per_call=np.array(per_call, dtype=np.int64), in_flight=np.int64(ex.in_flight()))
    b.W.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
