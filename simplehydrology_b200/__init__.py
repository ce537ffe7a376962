"""simplehydrology_b200 -- B200-native erosion hot path for SimpleHydrology worlds.

Python is only the test / benchmark harness here: the product is the C-ABI CUDA library
(include/shx.h -> simplehydrology_b200/libshx.so) and the C++ host adaptor
(simplehydrology_b200/host/shx_world.hpp).  This module binds the C ABI with ctypes and mirrors
the reference's names (World.erode(cycles), Drop / World parameter names).  It never falls
back to a CPU path: a missing library or GPU raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libshx.so")

# == quad::cell (reference source/cellpool.h:207-220)
CELL_DTYPE = np.dtype(
    [(n, np.float32) for n in ("height", "discharge", "momentumx", "momentumy",
                               "discharge_track", "momentumx_track", "momentumy_track", "rootdensity")])
# == struct Drop (reference source/water.h:12-39) + status word
DROP_DTYPE = np.dtype([("px", np.float32), ("py", np.float32), ("sx", np.float32), ("sy", np.float32),
                       ("volume", np.float32), ("sediment", np.float32), ("age", np.int32), ("flags", np.int32)])

MODE_BATCHED, MODE_SEQUENTIAL = 0, 1
F_HEIGHT, F_DISCHARGE, F_MOMENTUM, F_TRACKS, F_ROOTDENSITY, F_ALL = 1, 2, 4, 8, 16, 31
DROP_ALIVE, DROP_CASCADE, DROP_DONE_AGE, DROP_DONE_VOL, DROP_DONE_OOB = 1, 2, 4, 8, 16
DROP_REJECTED, DROP_DONE_NULL, DROP_MIGRATE_LO, DROP_MIGRATE_HI = 32, 64, 128, 256
HEIGHT_FRAC_BITS, TRACK_FRAC_BITS, LEDGER_FRAC_BITS = 26, 18, 32


class ShxError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"shx error {code}: {text}")
        self.code = code


class Params(C.Structure):
    """Drop:: statics (water.h:43-50), World:: statics (world.h:42-44), geometry (cellpool.h:165-179)"""
    _fields_ = [(n, C.c_float) for n in ("maxAge", "minVol", "evapRate", "depositionRate", "entrainment", "gravity",
                                         "momentumTransfer", "lrate", "maxdiff", "settling")] + \
               [(n, C.c_int) for n in ("mapscale", "tilesize", "mapsize", "lodsize")]


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("mode", C.c_int), ("row0", C.c_int), ("row1", C.c_int), ("halo", C.c_int),
                ("max_drops", C.c_size_t), ("block_threads", C.c_int), ("grid_blocks", C.c_int), ("variant", C.c_int),
                ("keep_tracks", C.c_int), ("coop", C.c_int), ("peer_rank", C.c_int), ("peer_world", C.c_int),
                ("max_cycles_per_launch", C.c_int), ("free_waits", C.c_int), ("no_l2_window", C.c_int)]


class PeerHandles(C.Structure):
    _fields_ = [("hq", C.c_ubyte * 64), ("rec", C.c_ubyte * 64), ("inbox", C.c_ubyte * 64),
                ("off_hq", C.c_uint64), ("off_rec", C.c_uint64), ("off_inbox", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("spawned", "rejected", "steps", "term_age", "term_vol", "term_oob",
                                          "cascade_transfers", "phases")] + \
               [(n, C.c_int64) for n in ("fx_eroded", "fx_deposited", "fx_sed_oob_lost", "fx_sed_deposited",
                                         "fx_sed_inflation")] + \
               [(n, C.c_uint64) for n in ("migrated_lo", "migrated_hi", "launches")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class PlantParams(C.Structure):
    """Plant:: statics, vegetation.h:40-44"""
    _fields_ = [(n, C.c_float) for n in ("maxSize", "growRate", "maxSteep", "maxDischarge", "maxTreeHeight")]


class VegStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("plants", "born", "died", "refused")]


class Timing(C.Structure):
    _fields_ = [("spawn_ms", C.c_double), ("descend_ms", C.c_double), ("ema_ms", C.c_double),
                ("descend_launches", C.c_uint64), ("pack_ms", C.c_double), ("d2h_ms", C.c_double), ("push_ms", C.c_double),
                ("scatter_ms", C.c_double)]


_lib = None


def lib():
    """load libshx.so; raises if it has not been built (no fallback of any kind)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ShxError(-2, f"{LIB_PATH} is missing: run `python -m simplehydrology_b200.build` "
                           "(the CUDA library is the only implementation)")
    L = C.CDLL(LIB_PATH)
    vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
    L.shx_last_error.restype = C.c_char_p
    L.shx_default_params.argtypes = [C.POINTER(Params), C.c_int]
    L.shx_default_params.restype = None
    L.shx_default_config.argtypes = [C.POINTER(Config)]
    L.shx_default_config.restype = None
    L.shx_create.argtypes = [C.POINTER(vp), C.POINTER(Params), C.POINTER(Config)]
    L.shx_destroy.argtypes = [vp]
    L.shx_destroy.restype = None
    L.shx_set_params.argtypes = [vp, C.POINTER(Params)]
    L.shx_get_params.argtypes = [vp, C.POINTER(Params)]
    L.shx_set_stream.argtypes = [vp, vp]
    L.shx_sync.argtypes = [vp]
    L.shx_host_register.argtypes = [vp, sz]
    L.shx_host_unregister.argtypes = [vp]
    L.shx_upload.argtypes = [vp, vp, sz]
    L.shx_download.argtypes = [vp, vp, sz, C.c_uint]
    L.shx_download_async.argtypes = [vp, vp, sz, C.c_uint]
    L.shx_download_compact.argtypes = [vp, vp, sz, C.c_int]
    L.shx_erode.argtypes = [vp, C.c_int, u64, C.POINTER(Stats)]
    L.shx_erode_async.argtypes = [vp, C.c_int, u64]
    L.shx_read_stats.argtypes = [vp, C.POINTER(Stats)]
    L.shx_launch_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.shx_timing_enable.argtypes = [vp, C.c_int]
    L.shx_timing_read.argtypes = [vp, C.POINTER(Timing)]
    L.shx_erode_spawnlist.argtypes = [vp, vp, sz, C.POINTER(Stats)]
    L.shx_trace_drop.argtypes = [vp, C.c_float, C.c_float, vp, C.c_int, C.POINTER(C.c_int)]
    L.shx_reset_tracks.argtypes = [vp]
    L.shx_ema.argtypes = [vp]
    L.shx_spawn.argtypes = [vp, C.c_int, u64, u64, vp, C.POINTER(sz)]
    L.shx_run_drops.argtypes = [vp, vp, sz, C.POINTER(Stats)]
    L.shx_add_rootdensity.argtypes = [vp, vp, vp, sz]
    L.shx_set_rootdensity.argtypes = [vp, vp, vp, sz]
    L.shx_multi_last_error.restype = C.c_char_p
    L.shx_multi_create.argtypes = [C.POINTER(vp), C.POINTER(Params), C.c_int, vp, C.POINTER(Config)]
    L.shx_multi_destroy.argtypes = [vp]
    L.shx_multi_destroy.restype = None
    L.shx_multi_strips.argtypes = [vp]
    L.shx_multi_strip.argtypes = [vp, C.c_int]
    L.shx_multi_strip.restype = vp
    L.shx_multi_upload.argtypes = [vp, vp, sz]
    L.shx_multi_download.argtypes = [vp, vp, sz, C.c_uint]
    L.shx_multi_init_terrain.argtypes = [vp, C.c_int]
    L.shx_multi_synth_terrain.argtypes = [vp, C.c_uint32]
    L.shx_multi_set_params.argtypes = [vp, C.POINTER(Params)]
    L.shx_multi_set_rootdensity.argtypes = [vp, vp, vp, sz]
    L.shx_multi_erode.argtypes = [vp, C.c_int, u64, C.POINTER(Stats)]
    L.shx_multi_erode_async.argtypes = [vp, C.c_int, u64]
    L.shx_multi_read_stats.argtypes = [vp, C.POINTER(Stats)]
    L.shx_multi_in_flight.argtypes = [vp, C.POINTER(sz)]
    L.shx_multi_sync.argtypes = [vp]
    L.shx_synth_terrain.argtypes = [vp, C.c_uint32]
    L.shx_init_terrain.argtypes = [vp, C.c_int]
    L.shx_view_textures.argtypes = [vp, vp, vp, vp]
    L.shx_view_textures_download.argtypes = [vp, vp, vp, vp, sz]
    L.shx_measure_read_bandwidth.argtypes = [vp, sz, C.c_int, C.POINTER(C.c_double)]
    L.shx_download_raw.argtypes = [vp, vp, vp]
    L.shx_stored_rows.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.shx_strip_pack_halo_delta.argtypes = [vp, vp, vp]
    L.shx_strip_apply_halo_delta.argtypes = [vp, vp, vp]
    L.shx_strip_pack_boundary.argtypes = [vp, vp, vp]
    L.shx_strip_set_halo.argtypes = [vp, vp, vp]
    L.shx_strip_pack_migrants.argtypes = [vp, vp, vp, sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.shx_strip_run_device_drops.argtypes = [vp, vp, sz, C.POINTER(Stats)]
    L.shx_strip_erode_begin.argtypes = [vp, C.c_int, u64]
    L.shx_vertex_fill.argtypes = [vp, vp]
    L.shx_vertex_download.argtypes = [vp, vp, sz]
    L.shx_view_maps.argtypes = [vp, vp]
    L.shx_view_maps_download.argtypes = [vp, vp, sz]
    L.shx_gather_cells.argtypes = [vp, vp, sz, vp, vp]
    L.shx_default_plant_params.argtypes = [C.POINTER(PlantParams)]
    L.shx_default_plant_params.restype = None
    L.shx_veg_create.argtypes = [vp, sz, C.POINTER(PlantParams)]
    L.shx_veg_set_params.argtypes = [vp, C.POINTER(PlantParams)]
    L.shx_veg_grow.argtypes = [vp, u64, u64, C.POINTER(VegStats)]
    L.shx_veg_count.argtypes = [vp, C.POINTER(sz)]
    L.shx_veg_download.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.shx_veg_upload.argtypes = [vp, vp, sz, C.c_int]
    L.shx_veg_tree_models.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.shx_veg_tree_models_download.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.shx_strip_message_words.argtypes = [vp, sz]
    L.shx_strip_message_words.restype = sz
    L.shx_strip_pack_message.argtypes = [vp, vp, vp, sz]
    L.shx_strip_apply_message.argtypes = [vp, vp, vp, sz]
    L.shx_strip_erode_begin_with.argtypes = [vp, C.c_int, u64, vp, sz]
    L.shx_strip_erode_end.argtypes = [vp]
    L.shx_peer_export.argtypes = [vp, C.POINTER(PeerHandles)]
    L.shx_peer_attach.argtypes = [vp, C.POINTER(PeerHandles)]
    _lib = L
    return L


def default_params(mapsize=1):
    p = Params()
    lib().shx_default_params(C.byref(p), mapsize)
    return p


def _ptr(a):
    return a.ctypes.data if a is not None else None


class World:
    """Device-resident world; `erode(cycles)` is the reference's World::erode (world.h:54-88)."""

    def __init__(self, params=None, mapsize=1, mode=MODE_BATCHED, device=0, row0=0, row1=0, halo=2, max_drops=0,
                 block_threads=0, grid_blocks=0, variant=0, keep_tracks=0, coop=0, peer_rank=0, peer_world=0,
                 max_cycles_per_launch=0, free_waits=None, no_l2_window=0):
        self.L = lib()
        self.params = params if params is not None else default_params(mapsize)
        cfg = Config()
        self.L.shx_default_config(C.byref(cfg))
        cfg.device, cfg.mode, cfg.row0, cfg.row1, cfg.halo = device, mode, row0, row1, halo
        cfg.max_drops, cfg.block_threads, cfg.grid_blocks, cfg.variant = max_drops, block_threads, grid_blocks, variant
        cfg.keep_tracks, cfg.coop = keep_tracks, coop
        cfg.peer_rank, cfg.peer_world = peer_rank, peer_world
        cfg.max_cycles_per_launch = max_cycles_per_launch
        if free_waits is not None:
            cfg.free_waits = free_waits
        cfg.no_l2_window = no_l2_window
        self.cfg = cfg
        self.size = self.params.mapsize * self.params.tilesize
        self.ncells = self.size * self.size
        h = C.c_void_p()
        self._h = None
        self._check(self.L.shx_create(C.byref(h), C.byref(self.params), C.byref(cfg)))
        self._h = h

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise ShxError(rc, self.L.shx_last_error().decode())

    def close(self):
        if self._h:
            self.L.shx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_handle):
        self._check(self.L.shx_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def sync(self):
        self._check(self.L.shx_sync(self._h))

    def set_params(self, params):
        self._check(self.L.shx_set_params(self._h, C.byref(params)))
        self.params = params

    def stored_rows(self):
        a, b = C.c_int(), C.c_int()
        self._check(self.L.shx_stored_rows(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- map transfer (tiled AoS pool, the reference's host layout)
    def upload(self, cells):
        assert cells.dtype == CELL_DTYPE and cells.flags.c_contiguous
        self._check(self.L.shx_upload(self._h, cells.ctypes.data, cells.size))

    def download(self, out=None, mask=F_ALL, asynchronous=False):
        if out is None:
            out = np.zeros(self.ncells, CELL_DTYPE)
        fn = self.L.shx_download_async if asynchronous else self.L.shx_download
        self._check(fn(self._h, out.ctypes.data, out.size, mask))
        return out

    def download_compact(self, out, nthreads=0):
        """height / discharge / momentum of the owned cells into the first 16 bytes of out's records (the rest untouched)"""
        self._check(self.L.shx_download_compact(self._h, out.ctypes.data, out.size, nthreads))
        return out

    def download_raw(self):
        """(plane0, plane1, field[...,4] f32, track[...,4] i32) over the stored rows -- the device's own
        fixed-point state, for bit-exact comparison with the lock-step oracle"""
        _, nrows = self.stored_rows()
        shape = (nrows, self.size)
        hq = np.zeros(shape + (2,), np.int32)
        rec = np.zeros(shape + (8,), np.int32)
        self._check(self.L.shx_download_raw(self._h, hq.ctypes.data, rec.ctypes.data))
        field = np.ascontiguousarray(rec[..., :4]).view(np.float32)
        track = np.ascontiguousarray(rec[..., 4:])
        return np.ascontiguousarray(hq[..., 0]), np.ascontiguousarray(hq[..., 1]), field, track

    def download_height_q(self):
        """the interleaved Q5.26 height planes only: int32 array [rows, size, 2]"""
        _, nrows = self.stored_rows()
        hq = np.zeros((nrows, self.size, 2), np.int32)
        self._check(self.L.shx_download_raw(self._h, hq.ctypes.data, None))
        return hq

    def synth_terrain(self, seed):
        self._check(self.L.shx_synth_terrain(self._h, seed))

    def launch_info(self):
        """(CTAs, threads per CTA, lanes per drop) of the last descend launch"""
        g, b, l = C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.shx_launch_info(self._h, C.byref(g), C.byref(b), C.byref(l)))
        return g.value, b.value, l.value

    def measure_read_bandwidth(self, nbytes, passes):
        g = C.c_double()
        self._check(self.L.shx_measure_read_bandwidth(self._h, nbytes, passes, C.byref(g)))
        return g.value

    def init_terrain(self, seed):
        """World::map.init(..., SEED) (cellpool.h:349-409) on the device"""
        self._check(self.L.shx_init_terrain(self._h, seed))

    # -- the hot path
    def erode(self, cycles, seed=0):
        st = Stats()
        self._check(self.L.shx_erode(self._h, cycles, seed, C.byref(st)))
        return st

    def erode_async(self, cycles, seed=0):
        self._check(self.L.shx_erode_async(self._h, cycles, seed))

    def read_stats(self):
        st = Stats()
        self._check(self.L.shx_read_stats(self._h, C.byref(st)))
        return st

    def timing_enable(self, on=True):
        self._check(self.L.shx_timing_enable(self._h, int(on)))

    def timing_read(self):
        t = Timing()
        self._check(self.L.shx_timing_read(self._h, C.byref(t)))
        return t

    def erode_spawnlist(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        st = Stats()
        self._check(self.L.shx_erode_spawnlist(self._h, xy.ctypes.data, xy.size // 2, C.byref(st)))
        return st

    def trace_drop(self, x, y, max_steps=1024):
        tr = np.zeros((max_steps, 7), np.float32)
        n = C.c_int(0)
        self._check(self.L.shx_trace_drop(self._h, x, y, tr.ctypes.data, max_steps, C.byref(n)))
        return tr[:n.value].copy()

    def reset_tracks(self):
        self._check(self.L.shx_reset_tracks(self._h))

    def ema(self):
        self._check(self.L.shx_ema(self._h))

    def spawn(self, cycles, seed, epoch):
        n = C.c_size_t(0)
        cap = self.params.mapsize ** 2 * cycles
        xy = np.zeros((max(cap, 1), 2), np.float32)
        self._check(self.L.shx_spawn(self._h, cycles, seed, epoch, xy.ctypes.data, C.byref(n)))
        return xy[:n.value].copy()

    def run_drops(self, drops):
        assert drops.dtype == DROP_DTYPE and drops.flags.c_contiguous
        st = Stats()
        self._check(self.L.shx_run_drops(self._h, drops.ctypes.data, drops.size, C.byref(st)))
        return st

    def add_rootdensity(self, xy, delta):
        xy = np.ascontiguousarray(xy, np.int32)
        delta = np.ascontiguousarray(delta, np.float32)
        self._check(self.L.shx_add_rootdensity(self._h, xy.ctypes.data, delta.ctypes.data, delta.size))

    def set_rootdensity(self, xy, value):
        xy = np.ascontiguousarray(xy, np.int32)
        value = np.ascontiguousarray(value, np.float32)
        self._check(self.L.shx_set_rootdensity(self._h, xy.ctypes.data, value.ctypes.data, value.size))

    # -- strip exchange (device pointers as ints)
    def strip_pack_halo_delta(self, lo, hi):
        self._check(self.L.shx_strip_pack_halo_delta(self._h, lo, hi))

    def strip_apply_halo_delta(self, from_lo, from_hi):
        self._check(self.L.shx_strip_apply_halo_delta(self._h, from_lo, from_hi))

    def strip_pack_boundary(self, lo, hi):
        self._check(self.L.shx_strip_pack_boundary(self._h, lo, hi))

    def strip_set_halo(self, lo, hi):
        self._check(self.L.shx_strip_set_halo(self._h, lo, hi))

    def strip_pack_migrants(self, lo, hi, cap):
        a, b = C.c_int(), C.c_int()
        self._check(self.L.shx_strip_pack_migrants(self._h, lo, hi, cap, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- peer mode: one world over the GPUs of a box
    # -- per-frame views (cellpool.h:286-305, SimpleHydrology.cpp:341-354)
    def owned_cells(self):
        r0, r1 = (self.cfg.row0, self.cfg.row1) if self.cfg.row1 else (0, self.size)
        if self.cfg.peer_world > 1:
            rows = self.size // self.cfg.peer_world
            r0, r1 = self.cfg.peer_rank * rows, (self.cfg.peer_rank + 1) * rows
        return (r1 - r0) * self.size

    def vertex_fill(self, dev_ptr):
        self._check(self.L.shx_vertex_fill(self._h, dev_ptr))

    def vertex_download(self):
        n = self.owned_cells()
        out = np.zeros((n, 12), np.float32)
        self._check(self.L.shx_vertex_download(self._h, out.ctypes.data, n))
        return out

    def view_maps(self, dev_ptr):
        self._check(self.L.shx_view_maps(self._h, dev_ptr))

    def view_maps_download(self):
        n = self.owned_cells()
        out = np.zeros((n, 4), np.float32)
        self._check(self.L.shx_view_maps_download(self._h, out.ctypes.data, n))
        return out

    def view_textures(self, dev_discharge, dev_momentum, water_rgb=None):
        w = np.ascontiguousarray(water_rgb, np.float32).ctypes.data if water_rgb is not None else None
        self._check(self.L.shx_view_textures(self._h, w, dev_discharge, dev_momentum))

    def view_textures_download(self, water_rgb=None):
        n = self.owned_cells()
        a, b = np.zeros((n, 4), np.uint8), np.zeros((n, 4), np.uint8)
        w = np.ascontiguousarray(water_rgb, np.float32) if water_rgb is not None else None
        self._check(self.L.shx_view_textures_download(self._h, w.ctypes.data if w is not None else None, a.ctypes.data, b.ctypes.data, n))
        return a, b

    def gather_cells(self, xy, normals=True):
        """records (CELL_DTYPE) and World::map.normal of the cells xy[n, 2] (int32)"""
        xy = np.ascontiguousarray(xy, np.int32)
        n = xy.shape[0]
        out = np.zeros(n, CELL_DTYPE)
        nrm = np.zeros((n, 3), np.float32) if normals else None
        self._check(self.L.shx_gather_cells(self._h, xy.ctypes.data, n, out.ctypes.data, nrm.ctypes.data if normals else None))
        return (out, nrm) if normals else out

    # ---- N3: Vegetation::grow on the device (vegetation.h:122-188)
    def veg_create(self, max_plants=0, plant_params=None):
        self._check(self.L.shx_veg_create(self._h, max_plants, C.byref(plant_params) if plant_params is not None else None))

    def veg_grow(self, seed, frame):
        st = VegStats()
        self._check(self.L.shx_veg_grow(self._h, seed, frame, C.byref(st)))
        return st

    def veg_count(self):
        n = C.c_size_t()
        self._check(self.L.shx_veg_count(self._h, C.byref(n)))
        return int(n.value)

    def veg_plants(self):
        """the plant list as float32 [n, 3] = {pos.x, pos.y, size} (Vegetation::plants)"""
        n = self.veg_count()
        out = np.zeros((n, 3), np.float32)
        got = C.c_size_t()
        self._check(self.L.shx_veg_download(self._h, out.ctypes.data, n, C.byref(got)))
        return out

    def veg_upload(self, plants, stamp_roots=True):
        plants = np.ascontiguousarray(plants, np.float32).reshape(-1, 3)
        self._check(self.L.shx_veg_upload(self._h, plants.ctypes.data, plants.shape[0], 1 if stamp_roots else 0))

    def veg_tree_models(self, dev_ptr=None):
        """glm::mat4 per plant (SimpleHydrology.cpp:329-335); into device memory, or returned as float32 [n, 16]"""
        n = self.veg_count()
        got = C.c_size_t()
        if dev_ptr is not None:
            self._check(self.L.shx_veg_tree_models(self._h, dev_ptr, n, C.byref(got)))
            return n
        out = np.zeros((n, 16), np.float32)
        self._check(self.L.shx_veg_tree_models_download(self._h, out.ctypes.data, n, C.byref(got)))
        return out

    def strip_message_words(self, cap):
        return int(self.L.shx_strip_message_words(self._h, cap))

    def strip_pack_message(self, lo, hi, cap):
        self._check(self.L.shx_strip_pack_message(self._h, lo, hi, cap))

    def strip_apply_message(self, from_lo, from_hi, cap):
        self._check(self.L.shx_strip_apply_message(self._h, from_lo, from_hi, cap))

    def peer_export(self):
        h = PeerHandles()
        self._check(self.L.shx_peer_export(self._h, C.byref(h)))
        return bytes(h)

    def peer_attach(self, handle_blobs):
        """handle_blobs: the peer_export() bytes of every rank, in rank order"""
        arr = (PeerHandles * len(handle_blobs))()
        for i, b in enumerate(handle_blobs):
            C.memmove(C.byref(arr[i]), b, C.sizeof(PeerHandles))
        self._check(self.L.shx_peer_attach(self._h, arr))

    def strip_erode_begin(self, cycles, seed=0, carried_ptr=None, n_carried=0):
        self._check(self.L.shx_strip_erode_begin_with(self._h, cycles, seed, carried_ptr, n_carried))

    def strip_erode_end(self):
        self._check(self.L.shx_strip_erode_end(self._h))

    def strip_run_device_drops(self, dev_ptr, n, want_stats=True):
        st = Stats()
        self._check(self.L.shx_strip_run_device_drops(self._h, dev_ptr, n, C.byref(st) if want_stats else None))
        return st


class _StripView(World):
    """a strip context owned by a MultiWorld: the World methods over a borrowed handle"""

    def __init__(self, L, handle, params, cfg):
        self.L, self._h, self.params, self.cfg = L, handle, params, cfg
        self.size = params.mapsize * params.tilesize
        self.ncells = self.size * self.size

    def close(self):
        self._h = None  # owned by the MultiWorld


class MultiWorld:
    """One world over `ngpu` row strips driven by ONE host thread through shx_multi (include/shx.h): the strips
    exchange once per erode call by peer stores into the neighbours' inboxes.  `devices` may repeat an ordinal
    (k logical strips on one GPU)."""

    def __init__(self, params=None, mapsize=1, ngpu=1, devices=None, **cfg_kw):
        self.L = lib()
        self.params = params if params is not None else default_params(mapsize)
        self.size = self.params.mapsize * self.params.tilesize
        self.ncells = self.size * self.size
        cfg = Config()
        self.L.shx_default_config(C.byref(cfg))
        for k, v in cfg_kw.items():
            setattr(cfg, k, v)
        dev = (C.c_int * ngpu)(*devices) if devices is not None else None
        h = C.c_void_p()
        self._h = None
        self._check(self.L.shx_multi_create(C.byref(h), C.byref(self.params), ngpu, dev, C.byref(cfg)))
        self._h = h
        self.ngpu = ngpu
        rows = (self.params.mapsize // ngpu) * self.params.tilesize
        self.strips = []
        for i in range(ngpu):
            c = Config.from_buffer_copy(bytes(cfg))
            if ngpu > 1:
                c.row0, c.row1 = i * rows, (i + 1) * rows
            self.strips.append(_StripView(self.L, C.c_void_p(self.L.shx_multi_strip(self._h, i)), self.params, c))

    def _check(self, rc):
        if rc != 0:
            raise ShxError(rc, self.L.shx_multi_last_error().decode())

    def close(self):
        if self._h:
            self.L.shx_multi_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, cells):
        assert cells.dtype == CELL_DTYPE and cells.size == self.ncells and cells.flags.c_contiguous
        self._check(self.L.shx_multi_upload(self._h, cells.ctypes.data, cells.size))

    def download(self, out=None, mask=F_ALL):
        if out is None:
            out = np.zeros(self.ncells, CELL_DTYPE)
        self._check(self.L.shx_multi_download(self._h, out.ctypes.data, out.size, mask))
        return out

    def launch_info(self):
        """(CTAs, threads per CTA, lanes per drop) of the last descend launch"""
        g, b, l = C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.shx_launch_info(self._h, C.byref(g), C.byref(b), C.byref(l)))
        return g.value, b.value, l.value

    def measure_read_bandwidth(self, nbytes, passes):
        g = C.c_double()
        self._check(self.L.shx_measure_read_bandwidth(self._h, nbytes, passes, C.byref(g)))
        return g.value

    def init_terrain(self, seed):
        self._check(self.L.shx_multi_init_terrain(self._h, seed))

    def synth_terrain(self, seed):
        self._check(self.L.shx_multi_synth_terrain(self._h, seed))

    def set_rootdensity(self, xy, value):
        xy = np.ascontiguousarray(xy, np.int32)
        value = np.ascontiguousarray(value, np.float32)
        self._check(self.L.shx_multi_set_rootdensity(self._h, xy.ctypes.data, value.ctypes.data, value.size))

    def erode(self, cycles, seed=0):
        st = Stats()
        self._check(self.L.shx_multi_erode(self._h, cycles, seed, C.byref(st)))
        return st

    def erode_async(self, cycles, seed=0):
        self._check(self.L.shx_multi_erode_async(self._h, cycles, seed))

    def read_stats(self):
        st = Stats()
        self._check(self.L.shx_multi_read_stats(self._h, C.byref(st)))
        return st

    def in_flight(self):
        n = C.c_size_t()
        self._check(self.L.shx_multi_in_flight(self._h, C.byref(n)))
        return n.value

    def sync(self):
        self._check(self.L.shx_multi_sync(self._h))
