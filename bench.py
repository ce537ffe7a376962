#!/usr/bin/env python
"""Benchmark of the erosion hot path (World::erode -> Drop::descend -> World::cascade).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU loop (oracle/_ref)

One "step" = one erosion cycle = one World::erode(512)-equivalent (reset + 512 drops per 512^2 node +
EMA, what one reference frame does: SimpleHydrology.cpp:319) on the 8192x8192 configuration of
BASELINE.json (configs[3], the one quoted at 1/2/4/8 GPUs; it fits one GPU), synthetic seeded
terrain.  N > 1 partitions the same map into row strips, one process per GPU (strong scaling).
Metric: particle steps/s (one particle step = one Drop::descend call); cycles/s rides along.

Rank 0 prints ONE JSON line.  `value` is timed with the world resident in HBM; `e2e` goes through
the C-ABI calls the host adaptor makes, with host buffers: per step a rootdensity push from pinned
host memory, shx_erode, and the download of the cell records into the host cell pool.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "particle_steps_per_s"
UNIT = "particle-steps/s"
BYTES_PER_STEP = 88       # 18 words gathered + 4 words scattered per Drop::descend
BYTES_PER_CELL_EMA = 48   # reset + EMA, per cell per cycle
MAPSIZE = 16              # 16 x 16 tiles of 512^2 = 8192^2
CYCLES = 512              # drops per node per erode call (SimpleHydrology.cpp:319: quad::tilesize)
SEED = 1


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    @staticmethod
    def _num(s):
        try:
            return float(s)
        except ValueError:
            return None

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(v for v in (self._num(r[1]) for r in self.rows) if v is not None)
        reasons = set()
        for r in self.rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self._num(self.rows[0][2]),
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------ reference arm

def reference_world(mapsize, seed):
    """the reference's own World (oracle/_ref, mapsize variant) holding the synthetic terrain"""
    import ctypes as C
    import orc
    if not orc.have_ref(mapsize):
        return None
    R = orc.Ref(mapsize)  # blank world: node table as cellpool.h:327-336, heights filled below
    p = orc.default_params(mapsize)
    h = orc.synth_terrain(512 * mapsize, seed)
    orc.lib().orc_fill_tiled_from_planar(C.byref(p), h.ctypes.data, R.cells.ctypes.data)
    return R


def time_reference(R, mapsize, drops_per_step, steps, warmup, seed=SEED):
    """World::erode's loop with explicit spawns (reset, spawn + `while(drop.descend())` per drop, EMA)
    over the reference's own Drop::descend / World::cascade; (particle steps, seconds) of `steps` samples"""
    import numpy as np
    rng = np.random.default_rng(seed)
    size = 512 * mapsize
    total_steps, total_s = 0, 0.0
    for i in range(warmup + steps):
        xy = rng.integers(0, size, size=(drops_per_step, 2)).astype(np.float32)
        t0 = time.perf_counter()
        st = R.erode_spawnlist(xy)
        dt = time.perf_counter() - t0
        if i >= warmup:
            total_steps += st["steps"]
            total_s += dt
    return total_steps, total_s


def replica_worker(seconds, seed):
    """one independent reference world (2048^2, its own seed) eroding for ~`seconds`; prints its particle steps and time"""
    import numpy as np
    R = reference_world(4, seed)
    if R is None:
        print("0 1.0")
        return 0
    rng = np.random.default_rng(seed)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        steps += R.erode_spawnlist(rng.integers(0, 2048, size=(2048, 2)).astype(np.float32))["steps"]
    print(steps, time.perf_counter() - t0)
    return 0


def reference_replicas(ncores, seconds=6.0):
    """The generous many-core figure: the reference loop is sequential and non-reentrant, so more
    cores can only run more worlds -- N independent 2048^2 replicas, one process per core, aggregate steps/s."""
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--replica", str(seconds), "--replica-seed", str(100 + i)],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for i in range(ncores)]
    total = 0.0
    for pr in procs:
        out, _ = pr.communicate(timeout=seconds * 10 + 120)
        try:
            st, dt = out.split()[-2:]
            total += float(st) / float(dt)
        except Exception:
            pass
    return total


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    R = reference_world(MAPSIZE, SEED)
    if R is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libshx_ref_m16.so missing (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    drops = 4096  # bounded sample of one cycle's 131072 drops: ~2 s of CPU per step incl. the full reset+EMA passes
    warm = min(args.warmup, 1)
    nsteps, secs = time_reference(R, MAPSIZE, drops, args.steps, warm)
    value = nsteps / secs
    sample = (f"{drops} of the cycle's {MAPSIZE * MAPSIZE * CYCLES} drops per step on the same 8192^2 synthetic world, "
              "incl. the reset and EMA passes over all cells")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "8192x8192 world (mapsize 16), erode(512)-equivalent cycle, reference CPU loop", "map": "8192x8192",
                       "drops_per_step": drops, "threads": 1,
                       "note": "the reference loop is sequential and non-reentrant (world.h:111 static scratch, global rand): 1 thread is all it can use"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": os.cpu_count()}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ CUDA arm

def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import simplehydrology_b200 as shx
    from simplehydrology_b200 import strips

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    peer = world > 1 and args.multi == "peer"
    if peer:
        # one world over all GPUs: peer-mapped strips, the kernel itself crosses NVLink
        strip = strips.PeerWorld(MAPSIZE, rank, world, local)
        strip.W.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        ex = None
    else:
        strip = strips.GpuStrip(MAPSIZE, rank, world, local)
        ex = strips.StripExchange(strip, rank, world)
    W = strip.W
    W.synth_terrain(SEED)
    names = [n for n, _ in shx.Stats._fields_]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()  # every strip holds its terrain before anybody's kernel reads across a border

    def one_step():
        """one erosion cycle on the device-resident world; returns this rank's counters"""
        if world == 1 or peer:
            W.erode_async(CYCLES, SEED)
        else:
            (ex.erode_cycle if args.multi == "cycle" else ex.erode)(CYCLES, SEED)
        return W.read_stats()  # one stream sync + 128-byte read per cycle

    for _ in range(args.warmup):
        one_step()
    barrier()

    # ---- timed region: K steps, device events on the launching stream, max over ranks
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    W.timing_enable(True)
    W.timing_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = dict.fromkeys(names, 0)
    rounds = 0
    barrier()
    e0.record()
    for _ in range(args.steps):
        st = one_step()
        for n in names:
            acc[n] += int(getattr(st, n))
        rounds += ex.rounds if ex is not None and world > 1 else 0
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    tm = W.timing_read()
    W.timing_enable(False)
    clocks = sampler.summary() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([acc[n] for n in names], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    # per-rank view of the same K steps: strips are static, drops concentrate in valleys (report imbalance)
    per_rank = torch.tensor([ms, tm.descend_ms, float(acc["steps"])], dtype=torch.float64, device=dev)
    ranks = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(ranks, per_rank)
    else:
        ranks = [per_rank]
    ms = float(t.item())
    total = dict(zip(names, [int(v) for v in tot.tolist()]))
    psteps = total["steps"]
    value = psteps / (ms * 1e-3)
    cells = (512 * MAPSIZE) ** 2
    hbm, peak_src = peaks()

    line = None
    if rank == 0:
        descend_ms = tm.descend_ms / max(args.steps, 1)  # rank 0's descend launches per cycle
        rank_steps = acc["steps"] / args.steps
        achieved = rank_steps * BYTES_PER_STEP / (descend_ms * 1e-3) / 1e9 if descend_ms > 0 else 0.0
        ema_ms = tm.ema_ms / max(args.steps, 1)
        ema_gbs = (cells / world) * BYTES_PER_CELL_EMA / (ema_ms * 1e-3) / 1e9 if ema_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "descend_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "8192x8192 world (mapsize 16, BASELINE configs[3]); one step = one erode(512) cycle: 131072 drops, "
                                       "lock-step batched descend + cascade, EMA",
                           "map": "8192x8192", "drops_per_cycle": MAPSIZE * MAPSIZE * CYCLES,
                           "parallelism": {"peer": f"row strips x{world}, peer-mapped over NVLink, one cross-GPU barrier per phase (bit-identical to 1 GPU)",
                                            "cycle": f"row strips x{world}, halo rows and border-crossing drops exchanged once per cycle (NCCL send/recv)",
                                            "rounds": f"row strips x{world}, exchange rounds until no drop is in flight"}[args.multi] if world > 1 else "single GPU",
                           "l2": "inputs larger than L2 (2.7 GB of map state vs 126 MB): no flush needed", "seed": SEED},
                "cycles_per_s": args.steps / (ms * 1e-3),
                "mean_steps_per_drop": psteps / max(total["spawned"], 1),
                "cascade_transfers_per_step": total["cascade_transfers"] / max(psteps, 1),
                "gpu_launches": total["launches"],
                "roofline": {"bound": "hbm", "kernel": "descend_lockstep_kernel", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                             "frac": achieved / hbm, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_particle_step": BYTES_PER_STEP, "kernel_ms_per_cycle": descend_ms,
                             "share_of_step": descend_ms / (ms / args.steps),
                             "ema_kernel": {"achieved": ema_gbs, "frac": ema_gbs / hbm, "ms_per_launch": ema_ms,
                                            "algorithmic_bytes_per_cell": BYTES_PER_CELL_EMA}},
                "clocks": clocks}
        if world > 1:
            if not peer:
                line["exchange_rounds_per_cycle"] = rounds / max(args.steps, 1)
            k = max(args.steps, 1)
            line["per_rank"] = {"ms_per_step": [float(r[0]) / k for r in ranks], "kernel_ms_per_cycle": [float(r[1]) / k for r in ranks],
                                "particle_steps_per_cycle": [float(r[2]) / k for r in ranks]}

    # ---- the per-frame consumers that follow erode in the reference's loop, as device
    # kernels on the resident world: vertex fill (updatenode) and the discharge / momentum maps
    own = (strip.row1 - strip.row0) * 512 * MAPSIZE
    views = {}
    for name, words, nbytes in (("vertex_fill", 12, 52), ("view_maps", 4, 36)):
        buf = torch.empty(own * words, dtype=torch.float32, device=dev)
        fn = (lambda: W.vertex_fill(buf.data_ptr())) if name == "vertex_fill" else (lambda: W.view_maps(buf.data_ptr()))
        for _ in range(3):
            fn()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for _ in range(5):
            fn()
        v1.record()
        torch.cuda.synchronize()
        vms = v0.elapsed_time(v1) / 5
        views[name] = {"ms_per_launch": vms, "achieved": own * nbytes / (vms * 1e-3) / 1e9, "frac": own * nbytes / (vms * 1e-3) / 1e9 / peaks()[0],
                       "algorithmic_bytes_per_cell": nbytes}
        del buf
    if rank == 0:
        line["roofline"]["view_kernels"] = views

    # ---- e2e: the calls the host adaptor makes, with host buffers; with strips every rank downloads
    # its own rows into its own (whole-map sized, as the reference's) pool
    # pinned host memory (cudaHostAlloc through torch; registering a pageable numpy buffer can fail silently under a
    # low RLIMIT_MEMLOCK, which turns the 2-GiB download into a pageable copy at a quarter of the speed)
    pool_t = torch.empty(cells * shx.CELL_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    pool_t.zero_()
    pool = pool_t.numpy().view(shx.CELL_DTYPE)
    nroot = 4096
    rng = np.random.default_rng(5)
    rxy_t = torch.empty((nroot, 2), dtype=torch.int32, pin_memory=True)  # this step's inputs, pinned as well
    rval_t = torch.zeros(nroot, dtype=torch.float32, pin_memory=True)
    rxy, rval = rxy_t.numpy(), rval_t.numpy()
    rxy[:] = np.stack([rng.integers(strip.row0, strip.row1, nroot), rng.integers(0, 512 * MAPSIZE, nroot)], 1)
    mask = shx.F_ALL if args.e2e_mask == "all" else (shx.F_HEIGHT | shx.F_DISCHARGE | shx.F_MOMENTUM)
    rec_bytes = 32 if args.e2e_mask == "all" else 16
    own_cells = (strip.row1 - strip.row0) * 512 * MAPSIZE

    def e2e_step():
        W.set_rootdensity(rxy, rval)     # host -> device: this step's inputs (Plant::root edits)
        n = one_step().steps
        W.download(out=pool, mask=mask)  # device -> host: the records the renderer / vegetation read
        return n

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    mine = 0
    for _ in range(e2e_steps):
        mine += e2e_step()
    barrier()
    dt = time.perf_counter() - t0
    tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
    tsum = torch.tensor([float(mine)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum)
    del pool, pool_t

    # ---- the same loop when the per-frame consumers run on the device: no pool download; the vertex records go
    # into a device buffer (the renderer's VBO through interop) and the host reads back only the cells its
    # vegetation pass looks at (shx_gather_cells).  Reported beside e2e, not instead of it.
    vbuf = torch.empty(own * 12, dtype=torch.float32, device=dev)
    nq = 8192
    qxy = np.stack([rng.integers(strip.row0 + 2, strip.row1 - 2, nq), rng.integers(0, 512 * MAPSIZE, nq)], 1).astype(np.int32)

    def interactive_step():
        W.set_rootdensity(rxy, rval)
        n = one_step().steps
        W.vertex_fill(vbuf.data_ptr())
        W.gather_cells(qxy)  # blocks: cells + normals on the host
        return n

    interactive_step()
    barrier()
    t0_i = time.perf_counter()
    mine_i = 0
    for _ in range(e2e_steps):
        mine_i += interactive_step()
    barrier()
    dt_i = time.perf_counter() - t0_i
    tmax_i = torch.tensor([dt_i], dtype=torch.float64, device=dev)
    tsum_i = torch.tensor([float(mine_i)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax_i, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum_i)
    del vbuf
    if rank == 0:
        line["e2e_device_consumers"] = {
            "value": float(tsum_i.item()) / float(tmax_i.item()), "unit": UNIT, "ms_per_step": 1e3 * float(tmax_i.item()) / e2e_steps,
            "h2d_bytes_per_step": int(rxy.nbytes + rval.nbytes + qxy.nbytes), "d2h_bytes_per_step": int(nq * (32 + 12)),
            "api": "shx_set_rootdensity + shx_erode + shx_vertex_fill (device buffer) + shx_gather_cells (8192 cells and normals to the host)"}
    if rank == 0:
        line["e2e"] = {"value": float(tsum.item()) / float(tmax.item()), "unit": UNIT,
                       "h2d_bytes_per_step": int(rxy.nbytes + rval.nbytes), "d2h_bytes_per_step": int(own_cells * rec_bytes),
                       "ms_per_step": 1e3 * float(tmax.item()) / e2e_steps, "steps": e2e_steps,
                       "api": "shx_set_rootdensity + shx_erode + shx_download (the calls of simplehydrology_b200/host/shx_world.hpp)",
                       "download": "full 32-byte records" if args.e2e_mask == "all" else "height+discharge+momentum (16 of 32 bytes per cell)"}

    # ---- CPU baseline: the reference's own loop on the host cores, bounded sample (rank 0, N == 1)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            W.close()
            R = reference_world(MAPSIZE, SEED)
            if R is not None:
                drops = 8192
                nst, secs = time_reference(R, MAPSIZE, drops, 3, 0)
                line["cpu_baseline"] = {
                    "value": nst / secs, "unit": UNIT, "cores": 1, "kind": "reference", "host_cores": os.cpu_count(),
                    "sample": f"3 x {drops} drops (of the cycle's {MAPSIZE * MAPSIZE * CYCLES}) on the same 8192^2 world through the reference's own "
                              f"Drop::descend / World::cascade (oracle/_ref), incl. reset + EMA passes; {secs:.1f} s of CPU"}
                del R
                ncores = os.cpu_count() or 1
                agg = reference_replicas(ncores)
                line["cpu_baseline"]["replicas"] = {
                    "value": agg, "unit": UNIT, "cores": ncores,
                    "note": "the reference loop cannot use more than one core for one world; this is N independent 2048^2 worlds, "
                            "one process per core, ~6 s each: the generous many-core figure, not the same job"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
        except Exception as e:  # the baseline must never take the GPU numbers down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {e}"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:  # last thing on stdout: NCCL may print its version banner whenever it first builds a communicator
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)  # 40 cycles ~ 0.5 s timed: several clock samples
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--e2e-mask", default="all", choices=["all", "hdm"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replica", type=float, default=0.0, help=argparse.SUPPRESS)
    ap.add_argument("--replica-seed", type=int, default=100, help=argparse.SUPPRESS)
    ap.add_argument("--multi", default="cycle", choices=["cycle", "peer", "rounds"],
                    help="N > 1: strips exchanging once per cycle (default), peer-mapped lock step, or exchange rounds")
    args = ap.parse_args()
    if args.replica > 0:
        return replica_worker(args.replica, args.replica_seed)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
