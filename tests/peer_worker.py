"""One rank of a peer-mode run (launched by tests/test_gpu_peer.py through torch.distributed.run):
erodes its strip of a synthetic world together with the other ranks and stores what it holds."""
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
    P = strips.PeerWorld(ms, rank, world, local)
    P.W.synth_terrain(tseed)
    dist.barrier()
    tot = None
    for _ in range(ncyc):
        P.erode(cycles, seed)
        st = strips.all_reduce_stats(P.W.read_stats(), torch.device("cuda", local))
        tot = st if tot is None else {k: (max(tot[k], v) if k == "phases" else tot[k] + v) for k, v in st.items()}
    hq = P.W.download_height_q()
    _, _, field, track = P.W.download_raw()
    np.savez(os.path.join(out, f"rank{rank}.npz"), hq=hq, field=field, track=track, row0=P.row0, row1=P.row1,
             stats=np.array([tot[k] for k in sorted(tot)], dtype=np.int64), names=np.array(sorted(tot)))
    P.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
