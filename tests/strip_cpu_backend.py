"""TEST INFRASTRUCTURE: a CPU stand-in for simplehydrology_b200.strips.GpuStrip, built on the
lock-step oracle, so that the strip exchange protocol (StripExchange) can be exercised with
world_size > 1 over gloo on a machine without GPUs.  It implements the same strip-local operations
as the C ABI's shx_strip_* calls, on numpy arrays."""
import numpy as np
import torch

import orc


class CpuStrip:
    def __init__(self, params, cells, rank, world, halo=2):
        self.p = params
        self.ls = orc.Ls(params)
        self.ls.upload(cells)
        self.size = self.ls.size
        rows = self.size // world
        self.row0, self.row1 = rank * rows, (rank + 1) * rows
        self.ls.w.contents.row0, self.ls.w.contents.row1 = self.row0, self.row1
        self.halo = halo
        self.cap = 2048  # drop records per message
        self.has_lo, self.has_hi = rank > 0, rank < world - 1
        self.epoch = 0
        self.stats = orc.Stats()
        self.last = np.zeros(0, orc.DROP_DTYPE)
        self._refresh_ref()

    # rows of the four bands
    def _halo_rows(self, side):
        return slice(self.row0 - self.halo, self.row0) if side == 0 else slice(self.row1, self.row1 + self.halo)

    def _edge_rows(self, side):
        return slice(self.row0, self.row0 + self.halo) if side == 0 else slice(self.row1 - self.halo, self.row1)

    def _refresh_ref(self):
        h = self.ls.height_q(0)
        self.ref = [h[self._halo_rows(0)].copy() if self.has_lo else None, h[self._halo_rows(1)].copy() if self.has_hi else None]

    def _accumulate(self, st):
        for n, _ in orc.Stats._fields_:
            setattr(self.stats, n, getattr(self.stats, n) + getattr(st, n))

    def begin(self, cycles, seed, carried=None):
        self.stats = orc.Stats()
        self.ls.reset_tracks()
        xy = self.ls.spawn(seed, self.epoch, cycles)
        self.epoch += 1
        mine = xy[(xy[:, 0] >= self.row0) & (xy[:, 0] < self.row1)]
        drops, st = self.ls.make_drops(mine)
        self._accumulate(st)
        if carried is not None and carried.shape[0]:  # shx_strip_erode_begin_with: appended to the spawned batch
            drops = np.concatenate([drops, carried.numpy().reshape(-1, 8).copy().view(orc.DROP_DTYPE).reshape(-1)])
            self.ls.w.contents.align_age = 1  # carried drops wait for the phase equal to their age
        st, _ = self.ls.run_drops(drops)
        self.ls.w.contents.align_age = 0
        self._accumulate(st)
        self.last = drops

    def end(self):
        self.ls.ema(reset=True)

    def pack_halo_delta(self):
        h = self.ls.height_q(0)
        out = []
        for side, has in ((0, self.has_lo), (1, self.has_hi)):
            d = (h[self._halo_rows(side)] - self.ref[side]) if has else np.zeros((self.halo, self.size), np.int32)
            out.append(torch.from_numpy(np.ascontiguousarray(d, np.int32).ravel()))
        return out[0], out[1]

    def apply_halo_delta(self, from_lo, from_hi):
        for side, t in ((0, from_lo), (1, from_hi)):
            if t is None:
                continue
            d = t.numpy().reshape(self.halo, self.size)
            for plane in (0, 1):
                self.ls.height_q(plane)[self._edge_rows(side)] += d

    def pack_boundary(self):
        h = self.ls.height_q(0)
        return (torch.from_numpy(np.ascontiguousarray(h[self._edge_rows(0)]).ravel().copy()),
                torch.from_numpy(np.ascontiguousarray(h[self._edge_rows(1)]).ravel().copy()))

    def set_halo(self, lo, hi):
        for side, t in ((0, lo), (1, hi)):
            if t is None:
                continue
            v = t.numpy().reshape(self.halo, self.size)
            for plane in (0, 1):
                self.ls.height_q(plane)[self._halo_rows(side)] = v
            self.ref[side] = v.copy()

    def pack_migrants(self):
        out = []
        for flag in (orc.DROP_MIGRATE_LO, orc.DROP_MIGRATE_HI):
            sel = self.last[(self.last["flags"] & flag) != 0].copy()
            sel["flags"] = (sel["flags"] & ~(orc.DROP_MIGRATE_LO | orc.DROP_MIGRATE_HI)) | orc.DROP_ALIVE
            out.append(torch.from_numpy(sel.view(np.int32).reshape(-1, 8).copy()))
        return out[0], out[1]

    # -- one message per neighbour and call (shx_strip_pack_message / shx_strip_apply_message)
    def pack_message(self):
        h = self.ls.height_q(0)
        mig = self.pack_migrants()
        out = []
        for side, has in ((0, self.has_lo), (1, self.has_hi)):
            if not has:
                out.append(None)
                continue
            n = mig[side].shape[0]
            msg = np.zeros(8 + 8 * self.cap + 2 * self.halo * self.size, np.int32)
            msg[0] = n
            msg[8:8 + 8 * min(n, self.cap)] = mig[side].numpy().ravel()[:8 * self.cap]
            rows = 8 + 8 * self.cap
            band = self.halo * self.size
            msg[rows:rows + band] = (h[self._halo_rows(side)] - self.ref[side]).ravel()
            msg[rows + band:rows + 2 * band] = h[self._edge_rows(side)].ravel()
            out.append(torch.from_numpy(msg))
        return out[0], out[1]

    def apply_message(self, from_lo, from_hi):
        rows = 8 + 8 * self.cap
        band = self.halo * self.size
        for side, t in ((0, from_lo), (1, from_hi)):
            if t is None:
                continue
            m = t.numpy()
            delta = m[rows:rows + band].reshape(self.halo, self.size)
            edge = m[rows + band:rows + 2 * band].reshape(self.halo, self.size)
            mine = self.ls.height_q(0)[self._halo_rows(side)] - self.ref[side]
            v = edge + mine
            for plane in (0, 1):
                self.ls.height_q(plane)[self._halo_rows(side)] = v
                self.ls.height_q(plane)[self._edge_rows(side)] += delta
            self.ref[side] = v.copy()

    def run_drops(self, records):
        drops = records.numpy().reshape(-1, 8).copy().view(orc.DROP_DTYPE).reshape(-1)
        st, _ = self.ls.run_drops(drops)
        self._accumulate(st)
        self.last = drops

    def owned_height_sum(self):
        return int(self.ls.height_q(0)[self.row0:self.row1].astype(np.int64).sum())
