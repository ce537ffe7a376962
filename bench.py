#!/usr/bin/env python
"""Benchmark of the erosion hot path (World::erode -> Drop::descend -> World::cascade).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU loop (oracle/_ref)

One "step" = one erosion cycle = one World::erode(512) call (reset + 512 drops per 512^2 node + EMA, what one
reference frame does: SimpleHydrology.cpp:319) on the 8192x8192 configuration of BASELINE.json (configs[3], the one
quoted at 1/2/4/8 GPUs; it fits one GPU), on the reference's OWN terrain: World::map.init with SEED 1
(cellpool.h:349-409), generated on the device by shx_init_terrain (bit-identical to the host init).  N > 1 partitions
the same map into row strips, one process per GPU (strong scaling).  Metric: particle steps/s (one particle step = one
Drop::descend call); cycles/s rides along.  Both arms run the same job: the reference arm marches the whole
131 072-drop cycle through the reference's own Drop::descend / World::cascade on one host thread (all it can use).

Rank 0 prints ONE JSON line.  `value` is timed with the world resident in HBM; `e2e` is the C++ host adaptor's
frame (simplehydrology_b200/host/bench_bridge.cpp: shx::Bridge::erode on a HOST cell pool: sparse rootdensity push,
erode, download of the 32-byte records), with a per-call breakdown.  `configs` adds the reference's own map sizes
(512^2, 2048^2: latency regime, set against a measured L2 bandwidth), `mature_world` the same 8192^2 cycle after 200
calls, `roofline.traffic` the DRAM bytes of one descend launch measured by an ncu child process of this run.
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


# ------------------------------------------------------------------------------ the job both arms run

def workload_config():
    """the workload, identically worded in both arms' JSON lines"""
    return {"workload": "8192x8192 world (mapsize 16, BASELINE configs[3]) on the reference's own terrain (World::map.init, SEED 1); "
                        "one step = one World::erode(512) call: reset, 131072 drops (512 per 512^2 node) marched to the end, EMA",
            "map": "8192x8192", "terrain": "World::map.init(SEED=1), cellpool.h:349-409", "drops_per_cycle": MAPSIZE * MAPSIZE * CYCLES,
            "cycles_per_step": 1, "seed": SEED,
            "l2": "inputs larger than L2 (3.2 GB of map state vs 126 MB): no flush needed"}


def node_major_spawns(rng, mapsize, cycles):
    """world.h:64-69: for every node, `cycles` positions node.pos + (rand % 512, rand % 512), node-major"""
    import numpy as np
    nodes = np.arange(mapsize * mapsize)
    ox = np.repeat((nodes // mapsize) * 512, cycles)
    oy = np.repeat((nodes % mapsize) * 512, cycles)
    xy = rng.integers(0, 512, size=(mapsize * mapsize * cycles, 2))
    xy[:, 0] += ox
    xy[:, 1] += oy
    return xy.astype(np.float32)


# ------------------------------------------------------------------------------ reference arm

def reference_world(mapsize, seed, heights=None):
    """the reference's own World (oracle/_ref, mapsize variant) holding the map::init terrain.  The terrain comes
    from `heights` (tiled pool order) or from the oracle's restatement of map::init (OpenMP over the host cores;
    the reference's own single-threaded init takes minutes at 8192^2) -- bit-identical either way."""
    import ctypes as C
    import orc
    if not orc.have_ref(mapsize):
        return None
    R = orc.Ref(mapsize)  # blank world: node table as cellpool.h:327-336, heights filled below
    if heights is not None:
        R.cells["height"] = heights
    else:
        p = orc.default_params(mapsize)
        h = orc.init_terrain(mapsize, seed)
        orc.lib().orc_fill_tiled_from_planar(C.byref(p), h.ctypes.data, R.cells.ctypes.data)
    return R


def time_reference(R, mapsize, cycles, steps, warmup, seed=SEED):
    """World::erode's loop with explicit spawns (reset, spawn + `while(drop.descend())` per drop, EMA) over the
    reference's own Drop::descend / World::cascade; (particle steps, seconds) of `steps` whole calls"""
    import numpy as np
    rng = np.random.default_rng(seed)
    total_steps, total_s = 0, 0.0
    for i in range(warmup + steps):
        xy = node_major_spawns(rng, mapsize, cycles)
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
    import orc
    R = reference_world(4, seed, heights=orc.planar_to_tiled(orc.default_params(4), orc.synth_terrain(2048, seed))["height"])
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
    """the same job on the reference's own CPU implementation: K whole erode(512) calls at 8192^2 after W warm-up
    calls, one host thread (the loop is sequential and non-reentrant: world.h:111 static scratch, global rand())"""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    R = reference_world(MAPSIZE, SEED)
    if R is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libshx_ref_m16.so missing (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    nsteps, secs = time_reference(R, MAPSIZE, CYCLES, args.steps, args.warmup)
    value = nsteps / secs
    sample = (f"{args.steps} whole erode(512) calls ({MAPSIZE * MAPSIZE * CYCLES} drops each) after {args.warmup} warm-up calls, through the "
              "reference's own Drop::descend / World::cascade (oracle/_ref, its headers compiled unmodified), incl. the reset and EMA passes")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(),
            "parallelism": "1 host thread: the reference loop is sequential and non-reentrant (world.h:111 static scratch, global rand())",
            "cycles_per_s": args.steps / secs,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": os.cpu_count()}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ CUDA arm

def timed_calls(W, cycles, calls, warm, seed=SEED):
    """`calls` erode(cycles) calls on a resident world between CUDA events; (ms per call, summed stats of the timed calls)"""
    import torch
    for _ in range(warm):
        W.erode_async(cycles, seed)
    W.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = {}
    e0.record()
    for _ in range(calls):
        W.erode_async(cycles, seed)
        st = W.read_stats()
        for n, _t in st._fields_:
            acc[n] = acc.get(n, 0) + int(getattr(st, n))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / calls, acc


def small_config(shx, torch, mapsize, cycles, calls, warm, l2_gbs):
    """the reference's own map sizes (BASELINE configs[1], [2]): latency regime, map state mostly L2-resident"""
    with shx.World(mapsize=mapsize) as W:
        W.set_stream(torch.cuda.current_stream().cuda_stream)
        W.init_terrain(SEED)
        ms, acc = timed_calls(W, cycles, calls, warm)
    steps = acc["steps"] / calls
    achieved = steps * BYTES_PER_STEP / (ms * 1e-3) / 1e9
    side = 512 * mapsize
    return {"map": f"{side}x{side}", "call": f"erode({cycles})", "drops_per_call": mapsize * mapsize * cycles, "calls_timed": calls,
            "ms_per_call": ms, "calls_per_s": 1e3 / ms, "particle_steps_per_s": steps / (ms * 1e-3),
            "phases_per_call": acc["phases"] / calls, "us_per_phase": 1e3 * ms / max(acc["phases"] / calls, 1),
            "mean_steps_per_drop": acc["steps"] / max(acc["spawned"], 1),
            "map_state_mb": side * side * 48 / 1e6,
            "roofline": {"bound": "l2", "achieved": achieved, "peak": l2_gbs, "unit": "GB/s", "frac": achieved / l2_gbs if l2_gbs else None,
                         "note": "algorithmic 88 B per particle step against the measured L2 read bandwidth (l2_peak below); the regime is "
                                 "bound by the per-phase latency chain, not by bandwidth"}}


def measure_traffic():
    """DRAM bytes of ONE descend launch of this workload, measured now by an ncu child process (never a timed number)"""
    ncu = next((c for c in ("/usr/local/cuda/bin/ncu", "ncu") if os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)), "ncu")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", "regex:descend_lockstep", "-s", "2", "-c", "1", "--csv", sys.executable, os.path.abspath(__file__), "--traffic-probe"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240).stdout
        import csv
        rows = [r for r in csv.reader(out.splitlines()) if len(r) > 10]
        hdr = rows[0]
        vals = {}
        for r in rows[1:]:
            d = dict(zip(hdr, r))
            v = float(d["Metric Value"].replace(",", ""))
            unit = d["Metric Unit"].lower()
            scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1.0)
            vals[d["Metric Name"]] = v * scale
        return vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], "measured by this run: ncu child process, one descend launch of the same workload"
    except Exception as e:  # no ncu on the box, or it failed: say so rather than quote an old number
        return None, f"unavailable ({type(e).__name__}: {e})"


def traffic_probe():
    import torch
    import simplehydrology_b200 as shx
    with shx.World(mapsize=MAPSIZE) as W:
        W.init_terrain(SEED)
        for _ in range(4):
            W.erode(CYCLES, SEED)
    return 0


def run_bridge(ngpu, frames=5, warmup=3):
    """the C++ host adaptor's own frame loop on a host pool (simplehydrology_b200/host/bench_bridge.cpp)"""
    exe = os.path.join(ROOT, "simplehydrology_b200", "host", "bench_bridge")
    if not os.path.exists(exe):
        from simplehydrology_b200 import build as B
        B.build()
        B.build_host_example()
    r = subprocess.run([exe, str(MAPSIZE), str(frames), str(warmup), str(ngpu), str(SEED)], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        raise RuntimeError(f"bench_bridge failed: {r.stderr.strip()[-400:]}")
    return json.loads(r.stdout.strip().splitlines()[-1])


def run_cuda(args):
    import datetime
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
    W.init_terrain(SEED)  # the reference's own terrain, every strip its rows (+ halo)
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
    launch_shape = W.launch_info()
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
        cfg = workload_config()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg,
                "parallelism": {"peer": f"row strips x{world}, peer-mapped over NVLink, one cross-GPU barrier per phase (bit-identical to 1 GPU)",
                                "cycle": f"row strips x{world}, one process per GPU, halo rows and border-crossing drops exchanged once per cycle (NCCL send/recv)",
                                "rounds": f"row strips x{world}, exchange rounds until no drop is in flight"}[args.multi] if world > 1 else "single GPU",
                "cycles_per_s": args.steps / (ms * 1e-3),
                "mean_steps_per_drop": psteps / max(total["spawned"], 1),
                "phases_per_cycle": acc["phases"] / max(args.steps, 1),
                "cascade_transfers_per_step": total["cascade_transfers"] / max(psteps, 1),
                "gpu_launches": total["launches"],
                "roofline": {"bound": "hbm", "kernel": {1: "descend_lockstep_kernel (one thread per drop)",
                                                        8: "descend_group_kernel (eight lanes per drop)"}.get(launch_shape[2], "?"),
                             "launch": {"ctas": launch_shape[0], "threads_per_cta": launch_shape[1], "lanes_per_drop": launch_shape[2]},
                             "achieved": achieved, "peak": hbm, "unit": "GB/s",
                             "frac": achieved / hbm, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_particle_step": BYTES_PER_STEP, "kernel_ms_per_cycle": descend_ms,
                             "us_per_phase": 1e3 * descend_ms / max(acc["phases"] / max(args.steps, 1), 1),
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
    for name, words, nbytes in (("vertex_fill", 12, 52), ("view_maps", 4, 36), ("view_textures", 2, 20)):
        buf = torch.empty(own * words, dtype=torch.float32, device=dev)
        fn = {"vertex_fill": lambda: W.vertex_fill(buf.data_ptr()), "view_maps": lambda: W.view_maps(buf.data_ptr()),
              "view_textures": lambda: W.view_textures(buf.data_ptr(), buf.data_ptr() + 4 * own)}[name]
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

    # ---- mature world: the same cycle after 200 more calls (rivers formed: queues, damping and cascades fire more)
    if world == 1:
        for _ in range(200):
            W.erode_async(CYCLES, SEED)
        W.sync()
        W.timing_enable(True)
        W.timing_read()
        m_ms, m_acc = timed_calls(W, CYCLES, 10, 0)
        m_tm = W.timing_read()
        W.timing_enable(False)
        m_steps = m_acc["steps"] / 10
        line["mature_world"] = {"after_calls": args.warmup + args.steps + 200, "ms_per_step": m_ms, "value": m_steps / (m_ms * 1e-3), "unit": UNIT,
                                "mean_steps_per_drop": m_acc["steps"] / max(m_acc["spawned"], 1), "phases_per_cycle": m_acc["phases"] / 10,
                                "cascade_transfers_per_step": m_acc["cascade_transfers"] / max(m_acc["steps"], 1),
                                "roofline_frac": m_steps * BYTES_PER_STEP / (m_tm.descend_ms / 10 * 1e-3) / 1e9 / hbm}

    # ---- the same loop when the per-frame consumers run on the device: no pool download; the vertex records go
    # into a device buffer (the renderer's VBO through interop) and the host reads back only the cells its
    # vegetation pass looks at (shx_gather_cells).  Reported beside e2e, not instead of it.
    if world == 1:
        rng = np.random.default_rng(5)
        nroot = 4096
        rxy = np.stack([rng.integers(2, 512 * MAPSIZE - 2, nroot), rng.integers(2, 512 * MAPSIZE - 2, nroot)], 1).astype(np.int32)
        rval = np.zeros(nroot, np.float32)
        vbuf = torch.empty(own * 12, dtype=torch.float32, device=dev)
        nq = 8192
        qxy = np.stack([rng.integers(2, 512 * MAPSIZE - 2, nq), rng.integers(0, 512 * MAPSIZE, nq)], 1).astype(np.int32)

        def interactive_step():
            W.set_rootdensity(rxy, rval)
            n = one_step().steps
            W.vertex_fill(vbuf.data_ptr())
            W.gather_cells(qxy)  # blocks: cells + normals on the host
            return n

        interactive_step()
        t0_i = time.perf_counter()
        mine_i = sum(interactive_step() for _ in range(5))
        dt_i = time.perf_counter() - t0_i
        del vbuf
        line["e2e_device_consumers"] = {
            "value": mine_i / dt_i, "unit": UNIT, "ms_per_step": 1e3 * dt_i / 5,
            "h2d_bytes_per_step": int(rxy.nbytes + rval.nbytes + qxy.nbytes), "d2h_bytes_per_step": int(nq * (32 + 12)),
            "api": "shx_set_rootdensity + shx_erode + shx_vertex_fill (device buffer) + shx_gather_cells (8192 cells and normals to the host)"}

    # ---- the reference's own map sizes and the L2 figure they are set against (N == 1)
    if world == 1:
        l2_gbs = W.measure_read_bandwidth(32 << 20, 64)
        hbm_probe = W.measure_read_bandwidth(4 << 30, 1)
        line["l2_peak"] = {"value": l2_gbs, "unit": "GB/s", "how": "64 passes of 16-byte .cg loads over a 32-MiB buffer (L2-resident), best of 5 launches",
                           "hbm_same_probe": hbm_probe}
    W.close()
    if world == 1:
        line["configs"] = {
            "default_512_erode512": small_config(shx, torch, 1, 512, 20, 3, l2_gbs),      # BASELINE configs[1], the reference's frame
            # a call with more cycles runs as consecutive batches of 512 drops per node between one reset and one EMA
            # (SURVEY.md 8d asks for erode(65536): that many visits of one river cell overflow the Q13.18 tracks, which
            # the library reports as SHX_ERR_RANGE instead of wrapping; 4096 is the same regime inside the range)
            "default_512_erode4096": small_config(shx, torch, 1, 4096, 3, 1, l2_gbs),
            "2048_erode512": small_config(shx, torch, 4, 512, 20, 3, l2_gbs)}             # BASELINE configs[2]

    # ---- weak-scaling point (N == 4 only, informational): a 16384^2 world, i.e. every GPU holds as many cells and
    # marches as many drops per call (131 072) as the single GPU does at 8192^2 -- what the strip exchange itself costs,
    # next to the strong-scaling headline whose limit is the latency floor of the lock step
    if world == 4 and not peer:
        big = strips.GpuStrip(2 * MAPSIZE, rank, world, local)
        bex = strips.StripExchange(big, rank, world)
        big.W.init_terrain(SEED)
        barrier()
        for _ in range(3):
            bex.erode_cycle(CYCLES, SEED)
            big.W.read_stats()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wsteps = 0
        barrier()
        w0.record()
        for _ in range(10):
            bex.erode_cycle(CYCLES, SEED)
            wsteps += big.W.read_stats().steps
        w1.record()
        barrier()
        wt = torch.tensor([w0.elapsed_time(w1)], dtype=torch.float64, device=dev)
        ws = torch.tensor([wsteps], dtype=torch.int64, device=dev)
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ws)
        big.W.close()
        if rank == 0:
            line["weak_scaling_point"] = {"map": "16384x16384", "n_gpus": 4, "drops_per_gpu_per_call": 2 * MAPSIZE * 2 * MAPSIZE * CYCLES // 4,
                                          "ms_per_step": float(wt.item()) / 10, "value": float(ws.item()) / (float(wt.item()) * 1e-3),
                                          "unit": UNIT, "note": "per-GPU cells and spawned drops as at N = 1 on 8192^2; compare value / 4 with the N = 1 line. A strip's batch is its 131 072 spawned drops "
                                                  "PLUS the drops its neighbours handed over, and a launch holds at most 131 072: every call pays a second, latency-bound launch"}

    # ---- BASELINE configs[4]: erosion coupled with the vegetation's rootdensity feedback and the vertex-pool update,
    # the whole frame on the device (SimpleHydrology.cpp:319-335: erode, Vegetation::grow, updatenode, tree models)
    if world == 1:
        with shx.World(mapsize=1) as Wc:
            Wc.set_stream(torch.cuda.current_stream().cuda_stream)
            Wc.init_terrain(SEED)
            Wc.veg_create(1 << 16)
            vb = torch.empty(512 * 512 * 12, dtype=torch.float32, device=dev)
            tm_buf = torch.empty((1 << 16) * 16, dtype=torch.float32, device=dev)

            def frame(f):
                Wc.erode_async(CYCLES, SEED)
                Wc.veg_grow(SEED, f)  # reads back 20 bytes of counters (the host needs the plant count)
                Wc.vertex_fill(vb.data_ptr())
                Wc.veg_tree_models(tm_buf.data_ptr())

            for f in range(250):
                frame(f)
            torch.cuda.synchronize()
            t0_c = time.perf_counter()
            for f in range(250, 300):
                frame(f)
            torch.cuda.synchronize()
            dt_c = (time.perf_counter() - t0_c) / 50
            line["configs"]["coupled_512"] = {
                "map": "512x512", "frame": "shx_erode(512) + shx_veg_grow + shx_vertex_fill + shx_veg_tree_models, all on the device",
                "frames": 300, "timed": "frames 250-299 (host wall clock, one 20-byte read-back per frame)", "ms_per_frame": 1e3 * dt_c,
                "frames_per_s": 1.0 / dt_c, "plants": Wc.veg_count(),
                "note": "the reference's own frame (World::erode + Vegetation::grow + updatenode on one core) takes ~85 ms at this size "
                        "(tests/test_gpu_coupled.py runs the 300-frame comparison: 4 001 plants there)"}
            del vb, tm_buf

    # ---- e2e: the C++ host adaptor's own frame on a HOST pool (rank 0 drives all N GPUs from one thread through
    # shx_multi; the other ranks wait on the rendezvous store, not in a GPU kernel, so their devices are free)
    if world > 1:
        store = dist.distributed_c10d._get_default_store()
        if rank != 0:
            store.wait(["shx_e2e_done"], datetime.timedelta(seconds=900))
    if rank == 0:
        try:
            b = run_bridge(world)
            line["e2e"] = {"value": b["value"], "unit": UNIT, "h2d_bytes_per_step": b["h2d_bytes_per_step"], "d2h_bytes_per_step": b["d2h_bytes_per_step"],
                           "ms_per_step": b["ms_per_step"], "steps": b["steps"], "breakdown_ms_per_step": b["breakdown_ms_per_step"],
                           "pool_register_s": b["pool_register_s"], "rootdensity_cells_per_step": b["rootdensity_cells_per_step"], "api": b["api"],
                           "download": b.get("download"), "download_probe_ms": b.get("download_probe_ms"),
                           "bound": "2 GiB of 32-byte records cross PCIe per step; erode and download are sequential because the host reads the step's "
                                    "own result.  (A 1-GiB compact stream scattered by 16 host threads, shx_download_compact, measured 52.3 ms against "
                                    "51.7: the host scatters no faster than PCIe saves.)"}
        except Exception as e:
            line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(e)}
        if world > 1:
            store.set("shx_e2e_done", "1")

    # ---- roofline.traffic: DRAM bytes of one descend launch, measured by an ncu child of this run (N == 1)
    if rank == 0 and world == 1 and not args.no_traffic:
        line["roofline"]["traffic"], line["roofline"]["traffic_source"] = measure_traffic()

    # ---- CPU baseline: the reference's own loop on the host cores, one whole cycle (rank 0, N == 1)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            with shx.World(mapsize=MAPSIZE, mode=shx.MODE_SEQUENTIAL) as W0:  # fp32 heights: the reference's init to the bit
                W0.init_terrain(SEED)
                heights = W0.download(mask=shx.F_HEIGHT)["height"].copy()
            R = reference_world(MAPSIZE, SEED, heights=heights)
            del heights
            if R is not None:
                nst, secs = time_reference(R, MAPSIZE, CYCLES, 1, 0)
                line["cpu_baseline"] = {
                    "value": nst / secs, "unit": UNIT, "cores": 1, "kind": "reference", "host_cores": os.cpu_count(),
                    "sample": f"1 whole erode(512) call ({MAPSIZE * MAPSIZE * CYCLES} drops) on the same 8192^2 world and terrain through the reference's own "
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child process that measures roofline.traffic")
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--replica", type=float, default=0.0, help=argparse.SUPPRESS)
    ap.add_argument("--replica-seed", type=int, default=100, help=argparse.SUPPRESS)
    ap.add_argument("--multi", default="cycle", choices=["cycle", "peer", "rounds"],
                    help="N > 1: strips exchanging once per cycle (default), peer-mapped lock step, or exchange rounds")
    args = ap.parse_args()
    if args.replica > 0:
        return replica_worker(args.replica, args.replica_seed)
    if args.traffic_probe:
        return traffic_probe()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
