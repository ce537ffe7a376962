"""ctypes bindings for the CPU oracle (oracle/libshx_oracle.so) and, when it has been
built, the reference's own headers compiled headless (oracle/_ref/libshx_ref_m*.so).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg only.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

CELL_DTYPE = np.dtype(
    [(n, np.float32) for n in ("height", "discharge", "momentumx", "momentumy",
                               "discharge_track", "momentumx_track", "momentumy_track", "rootdensity")])
DROP_DTYPE = np.dtype([("px", np.float32), ("py", np.float32), ("sx", np.float32), ("sy", np.float32),
                       ("volume", np.float32), ("sediment", np.float32), ("age", np.int32), ("flags", np.int32)])

DROP_ALIVE, DROP_CASCADE, DROP_DONE_AGE, DROP_DONE_VOL, DROP_DONE_OOB = 1, 2, 4, 8, 16
DROP_REJECTED, DROP_DONE_NULL, DROP_MIGRATE_LO, DROP_MIGRATE_HI = 32, 64, 128, 256


class Params(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("maxAge", "minVol", "evapRate", "depositionRate", "entrainment", "gravity",
                                         "momentumTransfer", "lrate", "maxdiff", "settling")] + \
               [(n, C.c_int) for n in ("mapscale", "tilesize", "mapsize", "lodsize")]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("spawned", "rejected", "steps", "term_age", "term_vol", "term_oob",
                                          "cascade_transfers", "phases")] + \
               [("fx_eroded", C.c_int64), ("fx_deposited", C.c_int64),
                ("fx_sed_oob_lost", C.c_int64), ("fx_sed_deposited", C.c_int64), ("fx_sed_inflation", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SeqWorld(C.Structure):
    _fields_ = [("p", Params), ("cells", C.c_void_p), ("erf_poly", C.c_int)]


class LsWorld(C.Structure):
    _fields_ = [("p", Params), ("size", C.c_int), ("h", C.POINTER(C.c_int32) * 2), ("field", C.POINTER(C.c_float)),
                ("track", C.POINTER(C.c_int32)), ("row0", C.c_int), ("row1", C.c_int), ("align_age", C.c_int), ("max_cycles_per_launch", C.c_int), ("exclusive_cells", C.c_int), ("cur_damp", C.c_float), ("free_waits", C.c_int), ("recip_evap", C.c_int), ("steps_per_phase", C.c_int)]


class PlantParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("maxSize", "growRate", "maxSteep", "maxDischarge", "maxTreeHeight")]


class VegStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("plants", "born", "died", "refused")]


PLANT_DTYPE = np.dtype([("x", np.int32), ("y", np.int32), ("size", np.float32)])


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True, stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "libshx_oracle.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in ("shx_oracle.c", "shx_oracle.h")]
        if not os.path.exists(path):
            build_oracle()
        elif any(os.path.getmtime(f) > os.path.getmtime(path) for f in srcs):
            try:
                build_oracle()  # make only rebuilds what is out of date
            except Exception:   # e.g. a snapshot that did not keep time stamps on a box without a compiler
                pass
        L = C.CDLL(path)
        L.orc_tiled_index.restype = C.c_size_t
        L.orc_tiled_index.argtypes = [C.POINTER(Params), C.c_int, C.c_int]
        L.orc_erff_libm.restype = C.c_float
        L.orc_erff_libm.argtypes = [C.c_float]
        L.orc_erff_poly.restype = C.c_float
        L.orc_erff_poly.argtypes = [C.c_float]
        L.orc_seq_height.restype = C.c_float
        L.orc_seq_cascade.restype = C.c_uint32
        L.orc_seq_cascade.argtypes = [C.POINTER(SeqWorld), C.c_float, C.c_float]
        L.orc_seq_trace_drop.argtypes = [C.POINTER(SeqWorld), C.c_float, C.c_float, C.c_void_p, C.c_int]
        L.orc_seq_erode_spawnlist.argtypes = [C.POINTER(SeqWorld), C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                              C.POINTER(Stats)]
        L.orc_seq_descend.argtypes = [C.POINTER(SeqWorld), C.c_void_p, C.POINTER(Stats)]
        L.orc_ls_create.restype = C.POINTER(LsWorld)
        L.orc_ls_create.argtypes = [C.POINTER(Params)]
        L.orc_ls_destroy.argtypes = [C.POINTER(LsWorld)]
        L.orc_ls_upload.argtypes = [C.POINTER(LsWorld), C.c_void_p]
        L.orc_ls_download.argtypes = [C.POINTER(LsWorld), C.c_void_p]
        L.orc_ls_quantize_height.restype = C.c_int32
        L.orc_ls_quantize_height.argtypes = [C.c_float]
        L.orc_ls_spawn.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.orc_ls_make_drops.argtypes = [C.POINTER(LsWorld), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(Stats)]
        L.orc_ls_run.argtypes = [C.POINTER(LsWorld), C.c_void_p, C.c_size_t, C.POINTER(Stats), C.c_void_p, C.c_int,
                                 C.POINTER(C.c_int)]
        L.orc_ls_ema.argtypes = [C.POINTER(LsWorld), C.c_int]
        L.orc_ls_reset_tracks.argtypes = [C.POINTER(LsWorld)]
        L.orc_ls_erode.argtypes = [C.POINTER(LsWorld), C.c_int, C.c_uint64, C.c_uint64, C.POINTER(Stats)]
        L.orc_ls_erode_spawnlist.argtypes = [C.POINTER(LsWorld), C.c_void_p, C.c_size_t, C.POINTER(Stats)]
        L.orc_synth_terrain.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.orc_default_plant_params.argtypes = [C.POINTER(PlantParams)]
        L.orc_veg_sync_counts.argtypes = [C.POINTER(LsWorld)]
        L.orc_veg_stamp_list.argtypes = [C.POINTER(LsWorld), C.c_void_p, C.c_size_t]
        L.orc_veg_grow.restype = C.c_size_t
        L.orc_veg_grow.argtypes = [C.POINTER(LsWorld), C.POINTER(PlantParams), C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t,
                                   C.c_size_t, C.POINTER(VegStats)]
        L.orc_fill_tiled_from_planar.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p]
        L.orc_vertex_fill.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p]
        L.orc_view_maps.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def default_params(mapsize=1):
    p = Params()
    lib().orc_default_params(C.byref(p), mapsize)
    return p


def tiled_index_map(p):
    """index array T with T[x, y] = position of world cell (x, y) in the tiled AoS pool"""
    ts, ms = p.tilesize, p.mapsize
    size = ts * ms
    x = np.arange(size)[:, None]
    y = np.arange(size)[None, :]
    return ((x // ts) * ms + (y // ts)) * (ts * ts) + (x % ts) * ts + (y % ts)


def planar_to_tiled(p, planar_height):
    """heights[x, y] -> zeroed tiled AoS cell buffer with that height"""
    size = p.tilesize * p.mapsize
    cells = np.zeros(size * size, CELL_DTYPE)
    cells["height"][tiled_index_map(p).ravel()] = np.asarray(planar_height, np.float32).ravel()
    return cells


def tiled_to_planar(p, cells, field="height"):
    size = p.tilesize * p.mapsize
    return cells[field][tiled_index_map(p).ravel()].reshape(size, size)


def vertex_fill(p, cells):
    """quad::updatenode restated (oracle): [ncells, 12] float32 in pool order"""
    out = np.zeros((cells.size, 12), np.float32)
    lib().orc_vertex_fill(C.byref(p), cells.ctypes.data, out.ctypes.data)
    return out


def view_maps(p, cells, erf_poly=1):
    """dischargeMap / momentumMap values + height: [size*size, 4] float32 in map order"""
    out = np.zeros((cells.size, 4), np.float32)
    lib().orc_view_maps(C.byref(p), cells.ctypes.data, int(erf_poly), out.ctypes.data)
    return out


def synth_terrain(size, seed):
    h = np.empty((size, size), np.float32)
    lib().orc_synth_terrain(h.ctypes.data, size, seed)
    return h


def init_terrain(mapsize, seed, tilesize=512):
    """map::init (cellpool.h:349-409) restated: planar heights of the reference's own terrain"""
    size = mapsize * tilesize
    h = np.empty((size, size), np.float32)
    L = lib()
    L.orc_init_terrain.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.orc_init_terrain.restype = None
    L.orc_init_terrain(h.ctypes.data, mapsize, tilesize, seed)
    return h


class Seq:
    """sequential fp32 oracle over a tiled AoS cell buffer (numpy structured array)"""

    def __init__(self, cells, params=None, erf_poly=False):
        self.p = params or default_params(1)
        self.cells = cells
        assert cells.dtype == CELL_DTYPE and cells.flags.c_contiguous
        assert cells.size == (self.p.mapsize * self.p.tilesize) ** 2
        self.w = SeqWorld(self.p, cells.ctypes.data, int(erf_poly))

    def normal(self, x, y):
        out = np.zeros(3, np.float32)
        lib().orc_seq_normal(C.byref(self.w), int(x), int(y), out.ctypes.data)
        return out

    def cascade(self, px, py):
        return lib().orc_seq_cascade(C.byref(self.w), px, py)

    def trace_drop(self, x, y, max_calls=1024):
        tr = np.zeros((max_calls, 7), np.float32)
        n = lib().orc_seq_trace_drop(C.byref(self.w), x, y, tr.ctypes.data, max_calls)
        return tr[:n].copy()

    def descend(self, drop):
        """drop: 1-element DROP_DTYPE array, updated in place; returns alive"""
        return lib().orc_seq_descend(C.byref(self.w), drop.ctypes.data, None)

    def erode_spawnlist(self, xy, reset=True, ema=True):
        xy = np.ascontiguousarray(xy, np.float32)
        st = Stats()
        lib().orc_seq_erode_spawnlist(C.byref(self.w), xy.ctypes.data, xy.size // 2, int(reset), int(ema), C.byref(st))
        return st


class Ls:
    """lock-step fixed-point oracle (the bit-exact checker for the batched CUDA path)"""

    def __init__(self, params=None):
        self.p = params or default_params(1)
        self.w = lib().orc_ls_create(C.byref(self.p))
        self.size = self.p.mapsize * self.p.tilesize

    def close(self):
        if self.w:
            lib().orc_ls_destroy(self.w)
            self.w = None

    def __del__(self):
        self.close()

    def upload(self, cells):
        assert cells.dtype == CELL_DTYPE and cells.size == self.size ** 2
        lib().orc_ls_upload(self.w, cells.ctypes.data)

    def download(self):
        cells = np.zeros(self.size ** 2, CELL_DTYPE)
        lib().orc_ls_download(self.w, cells.ctypes.data)
        return cells

    def height_q(self, plane=0):
        return np.ctypeslib.as_array(self.w.contents.h[plane], shape=(self.size, self.size))

    def field(self):
        return np.ctypeslib.as_array(self.w.contents.field, shape=(self.size, self.size, 4))

    def track_q(self):
        return np.ctypeslib.as_array(self.w.contents.track, shape=(self.size, self.size, 4))

    def spawn(self, seed, epoch, cycles):
        xy = np.zeros((self.p.mapsize ** 2 * cycles, 2), np.float32)
        lib().orc_ls_spawn(C.byref(self.p), seed, epoch, cycles, xy.ctypes.data)
        return xy

    def erode(self, cycles, seed, epoch):
        st = Stats()
        lib().orc_ls_erode(self.w, cycles, seed, epoch, C.byref(st))
        return st

    def erode_spawnlist(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        st = Stats()
        lib().orc_ls_erode_spawnlist(self.w, xy.ctypes.data, xy.size // 2, C.byref(st))
        return st

    # ---- Vegetation::grow under the device path's schedule (orc_veg_grow)
    def veg_create(self, max_plants, plant_params=None):
        self.plant_params = plant_params or PlantParams()
        if plant_params is None:
            lib().orc_default_plant_params(C.byref(self.plant_params))
        self.plants = np.zeros(max_plants + 1, PLANT_DTYPE)
        self.nplants = 0
        lib().orc_veg_sync_counts(self.w)

    def veg_grow(self, seed, frame):
        st = VegStats()
        self.nplants = int(lib().orc_veg_grow(self.w, C.byref(self.plant_params), seed, frame, self.plants.ctypes.data, self.nplants,
                                              self.plants.size - 1, C.byref(st)))
        return st

    def veg_upload(self, plants, stamp_roots=True):
        plants = np.asarray(plants, np.float32).reshape(-1, 3)
        n = plants.shape[0]
        self.plants["x"][:n] = plants[:, 0].astype(np.int32)
        self.plants["y"][:n] = plants[:, 1].astype(np.int32)
        self.plants["size"][:n] = plants[:, 2]
        self.nplants = n
        if stamp_roots:
            lib().orc_veg_stamp_list(self.w, self.plants.ctypes.data, n)

    def veg_plants(self):
        pl = self.plants[:self.nplants]
        return np.stack([pl["x"].astype(np.float32), pl["y"].astype(np.float32), pl["size"]], 1)

    def run_drops(self, drops, trace_cap=0):
        """march explicit drop records (DROP_DTYPE) to completion, no EMA; returns (stats, trace of drop 0)"""
        st = Stats()
        tr = np.zeros((max(trace_cap, 1), 7), np.float32)
        tn = C.c_int(0)
        lib().orc_ls_run(self.w, drops.ctypes.data, drops.size, C.byref(st),
                         tr.ctypes.data if trace_cap else None, trace_cap, C.byref(tn))
        return st, tr[:tn.value].copy()

    def make_drops(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        drops = np.zeros(xy.size // 2, DROP_DTYPE)
        st = Stats()
        lib().orc_ls_make_drops(self.w, xy.ctypes.data, drops.size, drops.ctypes.data, C.byref(st))
        return drops, st

    def ema(self, reset=False):
        return lib().orc_ls_ema(self.w, int(reset))

    def reset_tracks(self):
        lib().orc_ls_reset_tracks(self.w)


# ----------------------------------------------------------------- the compiled reference

def ref_path(mapsize):
    return os.path.join(ORACLE_DIR, "_ref", f"libshx_ref_m{mapsize}.so")


def have_ref(mapsize=1):
    return os.path.exists(ref_path(mapsize))


class Ref:
    """The reference's own World/Drop code (oracle/_ref).  World state is process-global
    and map.init is only faithful once per process, so use one instance per process per
    map size (tests that need a fresh world run in a subprocess, see run_ref_script)."""

    def __init__(self, mapsize=1, seed=None):
        L = C.CDLL(ref_path(mapsize))
        L.ref_ncells.restype = C.c_size_t
        L.ref_cells.restype = C.c_void_p
        L.ref_trace_drop.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_int]
        L.ref_erode_spawnlist.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.ref_cascade.argtypes = [C.c_float, C.c_float]
        L.ref_normal.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.ref_descend_once.argtypes = [C.c_void_p]
        L.ref_time_erode.restype = C.c_double
        L.ref_time_erode.argtypes = [C.c_int, C.c_int]
        L.ref_height.restype = C.c_float
        L.ref_update_vertices.argtypes = [C.c_void_p]
        self.L = L
        rc = L.ref_init(seed) if seed is not None else L.ref_init_blank()
        if rc != 0:
            raise RuntimeError("reference world already initialised in this process")
        n = L.ref_ncells()
        assert L.ref_cell_bytes() == CELL_DTYPE.itemsize
        buf = (C.c_char * (n * CELL_DTYPE.itemsize)).from_address(L.ref_cells())
        self.cells = np.frombuffer(buf, dtype=CELL_DTYPE)
        self.mapsize = mapsize
        self.size = L.ref_size()

    def params(self):
        out = np.zeros(10, np.float32)
        self.L.ref_get_params(out.ctypes.data)
        return out

    def set_params(self, arr):
        arr = np.ascontiguousarray(arr, np.float32)
        self.L.ref_set_params(arr.ctypes.data)

    def trace_drop(self, x, y, max_calls=1024):
        tr = np.zeros((max_calls, 7), np.float32)
        n = self.L.ref_trace_drop(x, y, tr.ctypes.data, max_calls)
        return tr[:n].copy()

    def erode_spawnlist(self, xy, reset=True, ema=True):
        xy = np.ascontiguousarray(xy, np.float32)
        st = np.zeros(3, np.uint64)
        self.L.ref_erode_spawnlist(xy.ctypes.data, xy.size // 2, int(reset), int(ema), st.ctypes.data)
        return {"spawned": int(st[0]), "rejected": int(st[1]), "steps": int(st[2])}

    def erode(self, cycles):
        self.L.ref_erode(cycles)

    def time_erode(self, cycles, reps):
        return self.L.ref_time_erode(cycles, reps)

    def normal(self, x, y):
        out = np.zeros(3, np.float32)
        self.L.ref_normal(int(x), int(y), out.ctypes.data)
        return out

    def cascade(self, px, py):
        self.L.ref_cascade(px, py)

    def descend_once(self, state7):
        s = np.ascontiguousarray(state7, np.float32).copy()
        alive = self.L.ref_descend_once(s.ctypes.data)
        return alive, s

    def vertices(self):
        out = np.zeros((self.cells.size, 12), np.float32)
        self.L.ref_update_vertices(out.ctypes.data)
        return out


def run_ref_script(code, timeout=600):
    """run a python snippet in a fresh interpreter (fresh reference world); returns stdout"""
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    if r.returncode != 0:
        raise RuntimeError(f"reference subprocess failed:\n{r.stderr}")
    return r.stdout
