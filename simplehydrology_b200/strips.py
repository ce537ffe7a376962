"""Row-strip decomposition of one world over the GPUs of a box: one process per GPU, one strip of
rows (constant x, whole tiles) per rank, neighbour-only exchange through torch.distributed.

Three protocols, all deterministic for a given number of strips:

  StripExchange.erode_cycle   ONE exchange per World::erode call (the default of bench.py --gpus N):
      every rank resets its tracks, spawns the drops of ITS nodes, appends the drops its neighbours
      handed over at the end of the previous call (they sleep until the phase equal to their age) and
      marches them in lock step until they finish or leave the strip; runs the EMA over its own
      rows; then one message per neighbour carries the migrant records, the integer height deltas
      accumulated in the halo rows and the current edge rows.  No collective, no host sync.
  StripExchange.erode         exchange ROUNDS within the call until no drop is in flight anywhere
      (halo deltas -> owner, edge rows -> halos, migrants -> neighbour, one 1-int all-reduce).
  PeerWorld                   no exchange by the host at all: the strips are mapped into each other
      (CUDA IPC) and the descend kernel crosses NVLink itself; bit-identical to one GPU.

The first two are statistically, not bitwise, the single-domain result (a crossing drop pauses,
halos are stale for a round / a call); tests/test_gpu_strips.py and tests/test_gpu_reference_parity.py
state the bounds.  The exchange logic is backend-agnostic: `GpuStrip` drives libshx on a CUDA
device; tests drive the same StripExchange with a CPU stand-in over gloo (tests/test_strips_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as dist


class StripExchange:
    """the per-call protocol; `backend` provides the strip-local operations on torch tensors"""

    def __init__(self, backend, rank, world):
        self.b, self.rank, self.world = backend, rank, world
        self.lo = rank - 1 if rank > 0 else None
        self.hi = rank + 1 if rank < world - 1 else None
        self.rounds = 0
        self.inbox = None  # erode_cycle: the neighbours' messages received at the end of the previous call
        self._counts = None

    def _swap(self, to_lo, to_hi, like_lo, like_hi):
        """send to_lo/to_hi to the neighbours, receive same-shaped tensors from them"""
        from_lo = torch.empty_like(like_lo) if self.lo is not None else None
        from_hi = torch.empty_like(like_hi) if self.hi is not None else None
        ops = []
        if self.lo is not None:
            ops += [dist.P2POp(dist.isend, to_lo, self.lo), dist.P2POp(dist.irecv, from_lo, self.lo)]
        if self.hi is not None:
            ops += [dist.P2POp(dist.isend, to_hi, self.hi), dist.P2POp(dist.irecv, from_hi, self.hi)]
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        return from_lo, from_hi

    def exchange_heights(self):
        d_lo, d_hi = self.b.pack_halo_delta()
        f_lo, f_hi = self._swap(d_lo, d_hi, d_lo, d_hi)  # lower neighbour's hi-halo covers my first rows
        self.b.apply_halo_delta(f_lo, f_hi)
        b_lo, b_hi = self.b.pack_boundary()
        f_lo, f_hi = self._swap(b_lo, b_hi, b_lo, b_hi)
        self.b.set_halo(f_lo, f_hi)

    def exchange_drops(self):
        """returns the number of drops this rank received"""
        out_lo, out_hi = self.b.pack_migrants()  # [n, 8] int32 records
        counts = torch.tensor([out_lo.shape[0], out_hi.shape[0]], dtype=torch.int64, device=out_lo.device)
        z = torch.zeros(2, dtype=torch.int64, device=out_lo.device)
        c_lo, c_hi = self._swap(counts[:1], counts[1:], z[:1], z[:1])
        n_lo = int(c_lo.item()) if c_lo is not None else 0  # what the lower neighbour sends up to me
        n_hi = int(c_hi.item()) if c_hi is not None else 0
        in_lo = out_lo.new_empty((n_lo, 8))
        in_hi = out_lo.new_empty((n_hi, 8))
        ops = []
        if self.lo is not None:
            if out_lo.shape[0]:
                ops.append(dist.P2POp(dist.isend, out_lo, self.lo))
            if n_lo:
                ops.append(dist.P2POp(dist.irecv, in_lo, self.lo))
        if self.hi is not None:
            if out_hi.shape[0]:
                ops.append(dist.P2POp(dist.isend, out_hi, self.hi))
            if n_hi:
                ops.append(dist.P2POp(dist.irecv, in_hi, self.hi))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        received = torch.cat([in_lo, in_hi]) if (n_lo + n_hi) else in_lo
        return received

    def erode(self, cycles, seed=0, max_rounds=1024):
        self.b.begin(cycles, seed)
        self.rounds = 0
        while True:
            self.exchange_heights()
            received = self.exchange_drops()
            flag = torch.tensor([received.shape[0]], dtype=torch.int64, device=received.device)
            if self.world > 1:
                dist.all_reduce(flag)
            self.rounds += 1
            if int(flag.item()) == 0:
                break
            if self.rounds >= max_rounds:
                raise RuntimeError("drops still crossing strip borders after max_rounds exchange rounds")
            self.b.run_drops(received)
        self.b.end()

    # -- one exchange per call ("exchanged each cycle"): strips run a whole call independently
    def erode_cycle(self, cycles, seed=0):
        """World::erode with ONE exchange: every strip marches its own drops (plus the ones handed
        over at the end of the previous call) to the end, then halo deltas / boundary rows / migrants
        are exchanged once.  A drop that crosses a strip border pauses until the next call, where it
        sleeps until the phase equal to its age (so it meets other drops at the same ages as in an
        unsplit run instead of bunching up at phase 0).  No collective, no host-side loop: the strips
        only meet their neighbours once per call.  erode_cycle(0) marches what is waiting without
        spawning (end of a run: repeat until in_flight() == 0)."""
        self.b.begin(cycles, seed, self._take_carry())
        self.b.end()  # EMA of the owned rows: nothing below touches the records
        out_lo, out_hi = self.b.pack_message()
        self.inbox = self._swap(out_lo, out_hi, out_lo, out_hi)  # ONE message per neighbour: migrants + halo deltas + edge rows
        self.b.apply_message(*self.inbox)
        self._counts = _MessageCounts(self.inbox)  # read by the next call (or in_flight), not now
        self.rounds = 1

    def _take_carry(self):
        """the drop records of the messages received at the end of the previous call, [n, 8] int32"""
        if self.inbox is None:
            return None
        return _records(self.inbox, self._counts.get(), self.b.cap)

    def in_flight(self):
        """drops waiting for the next call, summed over ranks"""
        n = sum(self._counts.get()) if self.inbox is not None else 0
        if self.world > 1:
            ref = next((m for m in (self.inbox or ()) if m is not None), None)
            t = torch.tensor([n], dtype=torch.int64, device=ref.device if ref is not None else None)
            dist.all_reduce(t)
            n = int(t.item())
        return n


MSG_HEADER = 8  # int32 words before the drop records of a strip message (include/shx.h)


class _MessageCounts:
    """the record counts in the headers of the two received messages, fetched without stalling the
    stream: an asynchronous copy into pinned memory now, waited for when somebody needs the numbers
    (by then the caller has long synchronised for its stats)"""

    def __init__(self, inbox):
        self.n = None
        heads = [m[:1] for m in inbox if m is not None]
        self.present = [m is not None for m in inbox]
        self.event = None
        if not heads:
            self.n = (0, 0)
            return
        dev = torch.cat(heads)
        if dev.is_cuda:
            self.host = torch.empty(len(heads), dtype=torch.int32, pin_memory=True)
            self.host.copy_(dev, non_blocking=True)
            self.event = torch.cuda.Event()
            self.event.record()
        else:
            self.host = dev.clone()

    def get(self):
        if self.n is None:
            if self.event is not None:
                self.event.synchronize()
            vals = iter(self.host.tolist())
            self.n = tuple(int(next(vals)) if p else 0 for p in self.present)
        return self.n


def _records(inbox, counts, cap):
    parts = []
    for m, n in zip(inbox, counts):
        if n > cap:
            raise RuntimeError(f"a neighbour handed over {n} drops, more than the message capacity {cap}")
        if m is not None and n:
            parts.append(m[MSG_HEADER:MSG_HEADER + 8 * n].view(-1, 8))
    if not parts:
        return None
    return parts[0] if len(parts) == 1 else torch.cat(parts)


class LocalStripSet:
    """k strips held by ONE process (e.g. k logical strips on one GPU): the same protocol as
    StripExchange with the neighbour transport replaced by handing tensors over directly.  Used to
    validate the strip kernels against the single-domain run without a multi-GPU box."""

    def __init__(self, backends):
        self.b = list(backends)
        self.rounds = 0
        self.inbox = None  # erode_cycle: per strip, the (lo, hi) messages received at the end of the previous call

    def _exchange_heights(self):
        k = len(self.b)
        deltas = [b.pack_halo_delta() for b in self.b]
        deltas = [(lo.clone(), hi.clone()) for lo, hi in deltas]
        for i, b in enumerate(self.b):
            b.apply_halo_delta(deltas[i - 1][1] if i > 0 else None, deltas[i + 1][0] if i < k - 1 else None)
        bnds = [b.pack_boundary() for b in self.b]
        bnds = [(lo.clone(), hi.clone()) for lo, hi in bnds]
        for i, b in enumerate(self.b):
            b.set_halo(bnds[i - 1][1] if i > 0 else None, bnds[i + 1][0] if i < k - 1 else None)

    def erode(self, cycles, seed=0, max_rounds=1024):
        k = len(self.b)
        for b in self.b:
            b.begin(cycles, seed)
        self.rounds = 0
        while True:
            self._exchange_heights()
            outs = [b.pack_migrants() for b in self.b]
            outs = [(lo.clone(), hi.clone()) for lo, hi in outs]
            inbox = []
            for i in range(k):
                parts = []
                if i > 0 and outs[i - 1][1].shape[0]:
                    parts.append(outs[i - 1][1])
                if i < k - 1 and outs[i + 1][0].shape[0]:
                    parts.append(outs[i + 1][0])
                inbox.append(torch.cat(parts) if parts else outs[i][0][:0])
            self.rounds += 1
            if sum(x.shape[0] for x in inbox) == 0:
                break
            if self.rounds >= max_rounds:
                raise RuntimeError("drops still crossing strip borders after max_rounds exchange rounds")
            for b, rec in zip(self.b, inbox):
                b.run_drops(rec)
        for b in self.b:
            b.end()

    def erode_cycle(self, cycles, seed=0):
        """StripExchange.erode_cycle for k strips in one process"""
        k = len(self.b)
        if self.inbox is None:
            self.inbox = [None] * k
        for b, box in zip(self.b, self.inbox):
            carried = None
            if box is not None:
                counts = tuple(int(m[0]) if m is not None else 0 for m in box)
                carried = _records(box, counts, b.cap)
            b.begin(cycles, seed, carried)
            b.end()
        outs = [b.pack_message() for b in self.b]
        outs = [(lo.clone() if lo is not None else None, hi.clone() if hi is not None else None) for lo, hi in outs]
        self.inbox = [(outs[i - 1][1] if i > 0 else None, outs[i + 1][0] if i < k - 1 else None) for i in range(k)]
        for b, box in zip(self.b, self.inbox):
            b.apply_message(*box)
        self.rounds = 1

    def in_flight(self):
        if self.inbox is None:
            return 0
        return sum(int(m[0]) for box in self.inbox if box is not None for m in box if m is not None)


class GpuStrip:
    """backend over libshx: one strip context on one CUDA device, buffers are torch CUDA tensors"""

    def __init__(self, mapsize, rank, world, device, halo=2, params=None, drops_per_node=512, **world_kw):
        import simplehydrology_b200 as shx
        self.shx = shx
        p = params if params is not None else shx.default_params(mapsize)
        size = p.mapsize * p.tilesize
        tiles = p.mapsize
        if tiles % world:
            raise ValueError("the number of tile rows must be divisible by the number of strips")
        rows = (tiles // world) * p.tilesize
        self.row0, self.row1 = rank * rows, (rank + 1) * rows
        self.size, self.halo = size, halo
        self.dev = torch.device("cuda", device)
        whole = world == 1
        nodes = (tiles // world) * tiles
        self.cap = max(4096, 4 * nodes * drops_per_node)  # spawned + carried-over drops (narrow strips carry more than they spawn)
        self.W = shx.World(params=p, device=device, row0=0 if whole else self.row0, row1=0 if whole else self.row1, halo=halo,
                           max_drops=self.cap, **world_kw)
        self.has_lo, self.has_hi = rank > 0, rank < world - 1
        n = halo * size
        mk = lambda: torch.zeros(n, dtype=torch.int32, device=self.dev)
        self.d_lo, self.d_hi, self.b_lo, self.b_hi = mk(), mk(), mk(), mk()
        self.out_lo = torch.zeros((self.cap, 8), dtype=torch.int32, device=self.dev)
        self.out_hi = torch.zeros((self.cap, 8), dtype=torch.int32, device=self.dev)
        self.msg = None
        self.W.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)

    @staticmethod
    def _p(t):
        return t.data_ptr() if t is not None else None

    def begin(self, cycles, seed, carried=None):
        if carried is not None and carried.shape[0]:
            carried = carried.contiguous()
            self.W.strip_erode_begin(cycles, seed, carried.data_ptr(), carried.shape[0])
        else:
            self.W.strip_erode_begin(cycles, seed)

    def end(self):
        self.W.strip_erode_end()

    def pack_halo_delta(self):
        self.W.strip_pack_halo_delta(self._p(self.d_lo) if self.has_lo else None, self._p(self.d_hi) if self.has_hi else None)
        return self.d_lo, self.d_hi

    def apply_halo_delta(self, from_lo, from_hi):
        self.W.strip_apply_halo_delta(self._p(from_lo), self._p(from_hi))

    def pack_boundary(self):
        self.W.strip_pack_boundary(self._p(self.b_lo) if self.has_lo else None, self._p(self.b_hi) if self.has_hi else None)
        return self.b_lo, self.b_hi

    def set_halo(self, lo, hi):
        self.W.strip_set_halo(self._p(lo), self._p(hi))

    def pack_message(self):
        if self.msg is None:
            words = self.W.strip_message_words(self.cap)
            self.msg = [torch.zeros(words, dtype=torch.int32, device=self.dev) if has else None for has in (self.has_lo, self.has_hi)]
        self.W.strip_pack_message(self._p(self.msg[0]), self._p(self.msg[1]), self.cap)
        return self.msg[0], self.msg[1]

    def apply_message(self, from_lo, from_hi):
        self.W.strip_apply_message(self._p(from_lo), self._p(from_hi), self.cap)

    def pack_migrants(self):
        n_lo, n_hi = self.W.strip_pack_migrants(self.out_lo.data_ptr(), self.out_hi.data_ptr(), self.cap)
        return self.out_lo[:n_lo], self.out_hi[:n_hi]

    def run_drops(self, records):
        records = records.contiguous()
        self.W.strip_run_device_drops(records.data_ptr(), records.shape[0], want_stats=False)

    def stats(self):
        return self.W.read_stats()


class PeerWorld:
    """One world spread over the GPUs of a box, one process per GPU (`shx_config.peer_world`).

    Unlike the exchange-round strips above, nothing is exchanged by the host: every rank maps the
    other ranks' strips through CUDA IPC once, and the descend kernel itself reads and adds into the
    strip a drop happens to be over (NVLink peer loads / system-scope REDs) and synchronises all
    GPUs at every phase.  The schedule is the single-GPU lock step, so the result is bit-identical
    to a one-GPU run of the same world.  torch.distributed is only used to pass the IPC handles
    around at start-up."""

    def __init__(self, mapsize, rank, world, device, params=None, max_drops=0):
        import simplehydrology_b200 as shx
        p = params if params is not None else shx.default_params(mapsize)
        size = p.mapsize * p.tilesize
        rows = size // world
        self.rank, self.world, self.size = rank, world, size
        self.row0, self.row1 = rank * rows, (rank + 1) * rows
        nodes = (rows // p.tilesize) * p.mapsize
        self.W = shx.World(params=p, device=device, peer_rank=rank, peer_world=world,
                           max_drops=max_drops or max(4096, nodes * 1024))
        blob = self.W.peer_export()
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        self.W.peer_attach(blobs)
        dist.barrier()

    def erode_async(self, cycles, seed=0):
        self.W.erode_async(cycles, seed)

    def erode(self, cycles, seed=0):
        return self.W.erode(cycles, seed)

    def close(self):
        if dist.is_initialized():
            dist.barrier()  # nobody unmaps while a peer may still be inside a kernel
        self.W.close()


def all_reduce_stats(st, device):
    """sum the counters of a Stats struct over ranks (phases: max)"""
    names = [n for n, _ in type(st)._fields_]
    t = torch.tensor([int(getattr(st, n)) for n in names], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        phases = t[names.index("phases")].clone()
        dist.all_reduce(t)
        dist.all_reduce(phases, op=dist.ReduceOp.MAX)
        t[names.index("phases")] = phases
    return dict(zip(names, [int(v) for v in t.tolist()]))
