/* shx -- B200-native erosion hot path for SimpleHydrology worlds.  C ABI.
 *
 * The reference has no plugin/FFI layer; its seam is the static C++ call
 * World::erode(int cycles) (reference source/world.h:33,54-88; called from
 * SimpleHydrology.cpp:319) operating on the global cell pool
 * (SimpleHydrology.cpp:11,36; source/cellpool.h:102-145,207-220) and the static
 * parameter sets (source/water.h:43-50, source/world.h:42-44).  This header is
 * what a binding for that seam needs: plain pointers and sizes, no CUDA or torch
 * types.  INTEGRATION.md shows the adaptor a maintainer would drop into
 * SimpleHydrology.cpp (simplehydrology_b200/host/shx_world.hpp).
 *
 * All calls are blocking unless named *_async, one context per map (or per row
 * strip of a map), not thread-safe -- the same contract as the reference, which is
 * single-threaded and non-reentrant (world.h:111 static scratch, global rand()).
 * There is no CPU fallback: every entry point fails with SHX_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef SHX_H
#define SHX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHX_VERSION 100

/* error codes (0 = ok).  The reference reports no errors at all (OOB lookups
 * return NULL / 0.0f / false: cellpool.h:90-94,421-443, water.h:62-68); those
 * in-simulation cases keep the reference's silent behaviour, the codes below are
 * for misuse of the boundary itself. */
enum {
  SHX_OK = 0,
  SHX_ERR_ARG = -1,       /* null pointer, bad size, unsupported parameter (lodsize != 1) */
  SHX_ERR_CUDA = -2,      /* CUDA runtime failure; shx_last_error() has the text */
  SHX_ERR_RANGE = -3,     /* a height does not fit the Q5.26 fixed point (|h| >= 32) */
  SHX_ERR_MODE = -4,      /* call not available in this context's mode */
  SHX_ERR_CAPACITY = -5,  /* more drops than the context was sized for */
  SHX_ERR_NOMEM = -6,
  SHX_ERR_PEER = -7       /* peer mode: not attached, or a peer GPU did not reach a phase barrier in time */
};

/* == quad::cell, cellpool.h:207-220: 8 x f32 = 32 B in this order.  Host buffers are
 * the reference's pool layout: node (x/tilesize)*mapsize + (y/tilesize) owns a
 * contiguous tilesize^2 slice (cellpool.h:327-336), x-major inside the tile
 * (include/math.h:11-14). */
typedef struct {
  float height, discharge, momentumx, momentumy;
  float discharge_track, momentumx_track, momentumy_track;
  float rootdensity;
} shx_cell;

/* Drop:: statics (water.h:43-50), World:: statics (world.h:42-44) and the geometry
 * constants that are compile-time in the reference (cellpool.h:165-179). */
typedef struct {
  float maxAge, minVol, evapRate, depositionRate, entrainment, gravity, momentumTransfer;
  float lrate, maxdiff, settling;
  int mapscale, tilesize, mapsize, lodsize;
} shx_params;

/* struct Drop, water.h:12-39 ({age,pos,speed,volume,sediment} = 28 B) plus a status
 * word; also the record handed between row strips. */
typedef struct {
  float px, py, sx, sy, volume, sediment;
  int age;
  int flags; /* SHX_DROP_* */
} shx_drop;

enum {
  SHX_DROP_ALIVE = 1,
  SHX_DROP_CASCADE = 2,      /* World::cascade(pos) of the last step still owed (water.h:151) */
  SHX_DROP_DONE_AGE = 4,     /* water.h:74-77 */
  SHX_DROP_DONE_VOL = 8,     /* water.h:79-82 */
  SHX_DROP_DONE_OOB = 16,    /* water.h:139-142 */
  SHX_DROP_REJECTED = 32,    /* world.h:71-72 */
  SHX_DROP_DONE_NULL = 64,   /* water.h:62-68 */
  SHX_DROP_MIGRATE_LO = 128, /* left this strip towards smaller x */
  SHX_DROP_MIGRATE_HI = 256,
  SHX_DROP_WAITED_SHIFT = 16, /* bits 16-18: phases the drop has been waiting for its cell (batched mode, saturating at 7);
                                 the k-th of several drops created on the same position starts with min(k, 7) */
  SHX_DROP_FREEW_SHIFT = 19,  /* bits 19-22: waits of this drop that did not cost it a step (<= shx_config.free_waits) */
  SHX_DROP_CHECK_SPAWN = 1 << 24 /* sequential mode: apply the height < 0.1 rejection when the drop starts (world.h:71-72) */
};

/* Counters of one call.  fx_* are exact integers: heights in Q5.26 (2^-26 units),
 * sediment sums in Q31.32.  Ledger identity (tests/test_gpu_reference_parity.py::test_check2_mass_ledger):
 *   sum(height_after) - sum(height_before) == fx_deposited - fx_eroded   (exactly, in Q5.26) */
typedef struct {
  uint64_t spawned, rejected, steps, term_age, term_vol, term_oob, cascade_transfers, phases;
  int64_t fx_eroded, fx_deposited;
  int64_t fx_sed_oob_lost, fx_sed_deposited, fx_sed_inflation;
  uint64_t migrated_lo, migrated_hi;
  uint64_t launches; /* kernels launched by this call */
} shx_stats;

#define SHX_HEIGHT_FRAC_BITS 26 /* heights: Q5.26 in an int32, |h| < 31 */
#define SHX_TRACK_FRAC_BITS 18  /* discharge/momentum tracks: Q13.18 in an int32; a call that pushes a
                                   discharge track past 4096 (about 4096 drop visits of ONE cell; the
                                   reference's default world peaks near 60) fails with SHX_ERR_RANGE
                                   instead of wrapping (detection is exact below 16384) */
#define SHX_LEDGER_FRAC_BITS 32 /* fx_sed_* sums: Q31.32 in an int64 */

enum {
  SHX_MODE_BATCHED = 0,   /* all drops of a call advance in lock step; integer atomics; deterministic */
  SHX_MODE_SEQUENTIAL = 1 /* one drop after another in fp32, operation for operation the reference
                             (parity anchor; one GPU thread, slow by construction) */
};

typedef struct {
  int device;          /* CUDA ordinal */
  int mode;            /* SHX_MODE_* */
  int row0, row1;      /* owned global rows [row0,row1) (x range); 0,0 = whole map */
  int halo;            /* halo rows kept on each side of a strip (>= 2); ignored for the whole map */
  size_t max_drops;    /* capacity of the drop buffers; 0 = maparea*1024 */
  /* launch-shape overrides of the descend kernel (tuning / tests); all 0 = the library chooses by drop count.
   * Results do not depend on block_threads / variant / coop: every shape is bit-identical. */
  int block_threads;   /* CTA size */
  int grid_blocks;     /* cap on the grid: a drop list longer than grid_blocks CTAs' worth marches as several launches
                          in list order, each seeing the heights the earlier ones left -- a DIFFERENT schedule, hence
                          different results (tests use it to exercise the split).  Without it a list is split every
                          131 072 drops on every device; a device that cannot keep that many co-resident fails with
                          SHX_ERR_CAPACITY rather than splitting elsewhere */
  int variant;         /* one thread per drop, register budget: 0 = 64 (1024 threads/SM), 1 = 128 (512/SM), 2 = 72 (7x128/SM),
                          3 = 72 (2x448/SM); 5 = eight lanes per drop (what the library picks for batches of up to
                          6 144 drops) in CTAs of block_threads */
  int keep_tracks;     /* 0: erode's EMA pass also zeroes the *_track accumulators (the reset the reference
                          does at the START of the next call, world.h:56-61, hoisted into the same pass);
                          1: leave them readable after erode, at the cost of one more pass per call */
  int coop;            /* 1 = warp-cooperative 3x3 gather (three lines per drop instead of nine lane accesses) */
  /* peer mode: ONE world spread over peer_world GPUs of a box, one process (context) per GPU.  This
   * context owns rows [rank*size/world, (rank+1)*size/world) -- row0/row1/halo are ignored -- and,
   * after shx_peer_attach, reads and adds into the other ranks' strips over NVLink from inside the
   * descend kernel; the per-phase barrier spans all GPUs.  Same schedule as one GPU => same result.
   * Every rank must make the same sequence of erode / run calls.  size/world must be a power of two
   * and a multiple of tilesize. */
  int peer_rank, peer_world;
  /* batched mode: at most this many drops per node march together; a call with more cycles runs as
   * consecutive batches (reset once before, EMA once after).  All drops of a batch read the same
   * frozen heights each step and take turns on shared cells (a waiting phase costs the drop a step
   * of its life), so a batch should stay sparse: 512 per 512^2 node per batch is what the
   * reference's frame loop issues (SimpleHydrology.cpp:319); in much denser batches the drops
   * spend their lives queueing.  0 = 512. */
  int max_cycles_per_launch;
  /* batched mode: the first free_waits (0..15; shx_default_config sets 8) phases a drop spends waiting for its cell do
   * not cost it a step of its life; every later wait does (the reference's drops never wait: without the budget the
   * queues in river cells cost the drops ~3 % of their steps, with 8 free waits ~1 %).  A launch needs at most
   * maxAge + 2 + free_waits phases. */
  int free_waits;
  /* 1: do not put the height / claim words under a persisting-L2 access-policy window.  By default a context whose
   * words fit the device's persisting L2 carve-out (<= 2048^2 on B200) launches its descend kernels with such a
   * window, so the streaming passes between calls cannot evict what every phase gathers. */
  int no_l2_window;
} shx_config;

/* CUDA IPC handles of one rank's strip (opaque bytes; gather them from all ranks with the caller's
 * own transport, e.g. torch.distributed.all_gather_object, and hand the array to shx_peer_attach) */
typedef struct {
  unsigned char hq[64], rec[64], inbox[64]; /* cudaIpcMemHandle_t of the allocations holding the buffers */
  uint64_t off_hq, off_rec, off_inbox;      /* byte offsets of the buffers inside those allocations */
} shx_peer_handles;

typedef struct shx_ctx shx_ctx;

/* field masks for shx_download */
enum {
  SHX_F_HEIGHT = 1, SHX_F_DISCHARGE = 2, SHX_F_MOMENTUM = 4, SHX_F_TRACKS = 8, SHX_F_ROOTDENSITY = 16,
  SHX_F_ALL = 31
};

int shx_version(void);
const char* shx_last_error(void);
void shx_default_params(shx_params* p, int mapsize);
void shx_default_config(shx_config* c);

/* lifetime -- replaces cellpool.reserve()/the World statics' initialisation
 * (SimpleHydrology.cpp:36, world.h:38-44) */
int shx_create(shx_ctx** out, const shx_params* p, const shx_config* cfg /* may be NULL */);
void shx_destroy(shx_ctx* c);
int shx_set_params(shx_ctx* c, const shx_params* p); /* geometry must not change */
int shx_get_params(const shx_ctx* c, shx_params* p);
int shx_set_stream(shx_ctx* c, void* cuda_stream);    /* cudaStream_t; default: the legacy default stream */
int shx_sync(shx_ctx* c);

/* optional: page-lock the caller's pool so up/downloads are true async DMA */
int shx_host_register(void* ptr, size_t bytes);
int shx_host_unregister(void* ptr);

/* map transfer.  `pool` is the whole tiled AoS pool (mapsize^2*tilesize^2 cells) even for
 * a strip context, which picks its rows (+halo).  The caller keeps ownership. */
int shx_upload(shx_ctx* c, const shx_cell* pool, size_t ncells);
int shx_download(shx_ctx* c, shx_cell* pool, size_t ncells, unsigned field_mask);
int shx_download_async(shx_ctx* c, shx_cell* pool, size_t ncells, unsigned field_mask);
/* {height, discharge, momentumx, momentumy} only -- the first 16 bytes of every owned record, i.e. what updatenode,
 * the texture builders and vegetation.h read after World::erode -- in half the PCIe bytes: a dense 16-byte stream is
 * copied into pinned chunks and scattered into the pool by `nthreads` host threads (0 = all) as it lands.  The
 * *_track words and rootdensity of the pool are left untouched.  Blocks.  Faster or slower than shx_download
 * depending on the host's memory bandwidth (every record costs the host a read-for-ownership of its cache line):
 * on this project's B200 boxes (16 host cores) 52.3 ms against 51.7 per 8192^2 frame, i.e. no gain, so shx::Bridge
 * keeps whole records by default and offers this as DownloadMode kCompact / kAuto. */
int shx_download_compact(shx_ctx* c, shx_cell* pool, size_t ncells, int nthreads);

/* == World::erode(cycles), world.h:54-88: reset tracks, `cycles` drops per node, EMA.
 * rand() is replaced by a counter-based hash keyed (seed, call counter, node, i). */
int shx_erode(shx_ctx* c, int cycles, uint64_t seed, shx_stats* out /* may be NULL */);
int shx_erode_async(shx_ctx* c, int cycles, uint64_t seed); /* no host sync; stats via shx_read_stats */
int shx_read_stats(shx_ctx* c, shx_stats* out);             /* syncs; stats of the last *_async call */
/* optional per-kernel device timing of shx_erode / shx_erode_async (CUDA events on the context's
 * stream): sums since the last read, in milliseconds.  shx_timing_read synchronises and resets. */
typedef struct {
  double spawn_ms, descend_ms, ema_ms;
  uint64_t descend_launches; /* timed descend spans (one per erode call) */
  double pack_ms;  /* shx_download: device layout -> tiled AoS staging tiles (sum over tiles) */
  double d2h_ms;   /* shx_download: the device-to-host copies on the copy stream (sum over tiles; overlaps pack_ms) */
  double push_ms;  /* shx_add/set_rootdensity: host-to-device copy + kernel */
  double scatter_ms; /* shx_download_compact: HOST time from the first chunk's arrival to the last record scattered */
} shx_timing;
/* shape of the last descend launch of this context: CTAs, threads per CTA, lanes per drop (1 = one thread per drop,
 * the dense / spread shapes; 8 = eight lanes per drop).  Results never depend on it; tests assert which kernel ran. */
int shx_launch_info(const shx_ctx* c, int* grid, int* block, int* lanes_per_drop);
int shx_timing_enable(shx_ctx* c, int on);
int shx_timing_read(shx_ctx* c, shx_timing* out);
/* same with explicit spawn points (x,y pairs, world coordinates) -- parity mode */
int shx_erode_spawnlist(shx_ctx* c, const float* xy, size_t ndrops, shx_stats* out);
/* one drop, state after every Drop::descend call: 7 floats {age,pos.x,pos.y,speed.x,speed.y,volume,sediment} */
int shx_trace_drop(shx_ctx* c, float x, float y, float* trace7, int max_steps, int* nsteps);

/* the pieces of erode, for tests and for the strip orchestration */
int shx_reset_tracks(shx_ctx* c);                                           /* world.h:56-61 */
int shx_ema(shx_ctx* c);                                                    /* world.h:81-86 */
int shx_spawn(shx_ctx* c, int cycles, uint64_t seed, uint64_t epoch, float* xy_out /* host, may be NULL */, size_t* n);
int shx_run_drops(shx_ctx* c, shx_drop* drops /* host, in/out */, size_t n, shx_stats* out);

/* sparse push of Plant::root stamps (vegetation.h:87-118): rootdensity[x,y] += delta */
int shx_add_rootdensity(shx_ctx* c, const int* xy, const float* delta, size_t n);
/* same, absolute values (rootdensity[x,y] = value); cells must be distinct.  This is what the host adaptor
 * uses after the unchanged Vegetation::grow() has edited the host pool. */
int shx_set_rootdensity(shx_ctx* c, const int* xy, const float* value, size_t n);

/* == World::map.init(vertexpool, cellpool, SEED), cellpool.h:349-409 (SimpleHydrology.cpp:38): the reference's own
 * terrain -- eight layers of FastNoiseLite 3-D OpenSimplex2 fBm, min/max normalised -- generated on the device, bit
 * for bit the heights the reference computes on the host; all other fields zeroed (the reference leaves them
 * uninitialised, cellpool.h:122). */
int shx_init_terrain(shx_ctx* c, int seed);
/* a cheaper device-side seeded synthetic terrain (value-noise fBm normalised to [0,1]); other fields zeroed */
int shx_synth_terrain(shx_ctx* c, uint32_t seed);

/* measurement aid (bench.py): read bandwidth of `passes` streaming passes over a scratch buffer of `bytes` (16-byte
 * .cg loads, best of 5 launches).  A buffer that fits L2 (e.g. 32 MiB) gives the L2 figure the small-map roofline is
 * set against; one far beyond L2 reproduces the HBM figure of MEASURED_PEAKS.json. */
int shx_measure_read_bandwidth(shx_ctx* c, size_t bytes, int passes, double* gbs);

/* raw device state over the stored rows, cell index (x-xlo)*size+y (tests: bit-exact comparison
 * with the lock-step oracle).  Either pointer may be NULL.
 *   hq2:   2 int32 per cell, the two Q5.26 height planes interleaved
 *   rec32: 32 bytes per cell {f32 discharge, momentumx, momentumy, rootdensity,
 *                             i32 Q13.18 discharge_track, momentumx_track, momentumy_track, pad} */
int shx_download_raw(shx_ctx* c, int32_t* hq2, void* rec32);
int shx_stored_rows(const shx_ctx* c, int* xlo, int* nrows);

/* ---- per-frame consumers of the eroded map, run on the device instead of the host
 * shx_vertex_fill replaces the `updatenode(vertexpool, node)` loop (SimpleHydrology.cpp:322-324,
 * source/cellpool.h:286-305): 12 floats per owned cell {position, normal, tangent, bitangent}
 * (the 48-byte Vertex of source/vertexpool.h:6-26) in pool order, written to DEVICE memory --
 * e.g. the vertex pool's VBO mapped with cudaGraphicsResourceGetMappedPointer -- so the interactive
 * path needs no height download.  shx_vertex_download is the same into a host buffer.
 * shx_view_maps replaces the dischargeMap / momentumMap builders (SimpleHydrology.cpp:341-354):
 * 4 floats per owned cell in map order (x*size + y):
 *   {erf(0.4*discharge), 0.5*(1+erf(momentumx)), 0.5*(1+erf(momentumy)), height}. */
int shx_vertex_fill(shx_ctx* c, float* dev_out);
int shx_vertex_download(shx_ctx* c, float* host_out, size_t ncells);
int shx_view_maps(shx_ctx* c, float* dev_out);
int shx_view_maps_download(shx_ctx* c, float* host_out, size_t ncells);
/* the same two maps as the RGBA8 textures the reference uploads (SimpleHydrology.cpp:341-354): dischargeMap =
 * vec4(waterColor, erf(0.4*discharge)), momentumMap = vec4(0.5*(1+erf(mx)), 0.5*(1+erf(my)), 0.5, 1), 4 bytes per
 * owned cell each in map order, channels (unsigned char)(255*c).  water_rgb: 3 floats, NULL = the reference's
 * default (92,133,142)/255 (model.h:22).  The byte packing restates TinyEngine 1.7's image::make, which is not
 * vendored by the reference: unpinned.  Device pointers (e.g. CUDA-GL mapped textures / PBOs). */
int shx_view_textures(shx_ctx* c, const float* water_rgb, uint8_t* dev_discharge_rgba, uint8_t* dev_momentum_rgba);
int shx_view_textures_download(shx_ctx* c, const float* water_rgb, uint8_t* host_discharge_rgba, uint8_t* host_momentum_rgba, size_t ncells);
/* Sparse read-back for host code that looks at a few cells per frame (Vegetation::grow: discharge /
 * height / normal / rootdensity at plant positions, vegetation.h:67-85,160-180) instead of the whole
 * pool: the records of the n queried cells {x, y} and, if normals3 != NULL, World::map.normal there
 * (cellpool.h:181-204 with the map-level oob of :413-419).  Host buffers; blocks.  A query outside
 * the map (map.get() == NULL in the reference) or outside a strip's stored rows returns zeros. */
int shx_gather_cells(shx_ctx* c, const int* xy, size_t n, shx_cell* out, float* normals3);

/* ---- N3: Vegetation::grow() on the device (reference source/vegetation.h:122-188).
 * The plant list and the rootdensity stamps live on the GPU, so a coupled frame (World::erode, Vegetation::grow,
 * updatenode, tree placement: SimpleHydrology.cpp:319-335) needs no pool download and no rootdensity push.  The
 * predicates and arithmetic are the reference's (Plant::grow :67-69, ::die :71-78, ::spawn :80-89, ::root :91-120);
 * the schedule is the parallel one described in simplehydrology_b200/csrc/shx_veg_kernels.cuh: every plant decides
 * from the maps as they are when the call starts, rand() is a counter hash keyed (seed, frame, cell), rootdensity is
 * an exact count of fifths (1.0 / 0.6 / 0.4 = 5 / 3 / 2) whose fp32 value is count/5.  Deterministic and run-to-run
 * identical; against the reference's sequential walk it is statistical (tests/test_gpu_vegetation.py).
 * Whole-map contexts only (not strips). */
typedef struct {
  float maxSize, growRate, maxSteep, maxDischarge, maxTreeHeight; /* Plant:: statics, vegetation.h:40-44 */
} shx_plant_params;
typedef struct {
  uint64_t plants;   /* in the list after the call */
  uint64_t born;     /* seeded + children of this call (root(+1) stamped) */
  uint64_t died;     /* removed by this call (root(-1) stamped) */
  uint64_t refused;  /* children that did not fit max_plants (not created, not stamped) */
} shx_veg_stats;
void shx_default_plant_params(shx_plant_params* pp);
/* allocate the plant store (max_plants == 0: one plant per four cells); pp == NULL: the reference's defaults */
int shx_veg_create(shx_ctx* c, size_t max_plants, const shx_plant_params* pp);
int shx_veg_set_params(shx_ctx* c, const shx_plant_params* pp);
/* == Vegetation::grow(), once; `frame` numbers the calls (it keys the hash together with `seed`).  Returns
 * SHX_ERR_CAPACITY (after doing the frame with the surplus children refused) if the list is full. */
int shx_veg_grow(shx_ctx* c, uint64_t seed, uint64_t frame, shx_veg_stats* out);
int shx_veg_count(shx_ctx* c, size_t* n);
/* the list as {pos.x, pos.y, size} per plant (Vegetation::plants), host buffer of 3*cap floats */
int shx_veg_download(shx_ctx* c, float* xys3, size_t cap, size_t* n);
/* replace the list by the host's plants; stamp_roots != 0 also applies their root(+1) stamps (a list whose roots
 * are already in the uploaded pool's rootdensity passes 0) */
int shx_veg_upload(shx_ctx* c, const float* xys3, size_t n, int stamp_roots);
/* the tree particle system's model matrices (SimpleHydrology.cpp:329-335): translate(pos.x, size +
 * mapscale*height(pos), pos.y) * scale(size), 16 floats per plant, column-major (glm::mat4), into DEVICE memory
 * (e.g. the instance buffer mapped through CUDA-GL interop) / into a host buffer */
int shx_veg_tree_models(shx_ctx* c, float* dev_out16, size_t cap, size_t* n);
int shx_veg_tree_models_download(shx_ctx* c, float* host_out16, size_t cap, size_t* n);

/* ---- row-strip exchange (multi-GPU): buffers are DEVICE pointers owned by the caller
 * (e.g. torch tensors); the transport between ranks is the caller's (NCCL send/recv or P2P). */
/* height deltas this strip accumulated in its halo rows since the last refresh: `halo`*size int32 per side */
int shx_strip_pack_halo_delta(shx_ctx* c, int32_t* dev_lo, int32_t* dev_hi);
/* add a neighbour's halo deltas onto the owned boundary rows (lo: rows row0.., hi: rows ..row1) */
int shx_strip_apply_halo_delta(shx_ctx* c, const int32_t* dev_from_lo, const int32_t* dev_from_hi);
/* current owned boundary rows, to refresh the neighbour's halo copy: `halo`*size int32 per side */
int shx_strip_pack_boundary(shx_ctx* c, int32_t* dev_lo, int32_t* dev_hi);
int shx_strip_set_halo(shx_ctx* c, const int32_t* dev_lo, const int32_t* dev_hi);
/* drops that left the strip during the last run: compacted into dev_lo/dev_hi (capacity cap each);
 * counts are written to the two host ints (syncs) */
int shx_strip_pack_migrants(shx_ctx* c, shx_drop* dev_lo, shx_drop* dev_hi, size_t cap, int* n_lo, int* n_hi);
/* Strips that exchange ONCE per call: one message per neighbour, int32 words
 *   [0] number of drop records  [1..7] unused  [8, 8+8*cap) drop records (shx_drop)
 *   then halo*size halo deltas, then halo*size current edge rows.
 * pack writes both outgoing messages (device buffers of shx_strip_message_words() words; a count
 * above cap means records were lost: the receiver must treat it as SHX_ERR_CAPACITY); apply takes
 * the neighbours' messages: their deltas onto the owned edge rows, and the halo copy becomes what
 * the neighbour holds after taking this strip's deltas.  Neither call synchronises.  The drop
 * records of the received messages go to the next shx_strip_erode_begin_with. */
size_t shx_strip_message_words(const shx_ctx* c, size_t cap);
int shx_strip_pack_message(shx_ctx* c, int32_t* dev_lo, int32_t* dev_hi, size_t cap);
int shx_strip_apply_message(shx_ctx* c, const int32_t* dev_from_lo, const int32_t* dev_from_hi, size_t cap);
/* continue drops received from neighbours (device buffer) until they finish or leave again */
int shx_strip_run_device_drops(shx_ctx* c, const shx_drop* dev_drops, size_t n, shx_stats* out);
/* one strip's share of World::erode split around the exchange rounds:
 *   begin = reset tracks + spawn this strip's nodes + march them (world.h:56-76), no EMA
 *   end   = EMA of the owned rows (world.h:81-86)
 * stats accumulate from begin to end (shx_read_stats after end). */
int shx_strip_erode_begin(shx_ctx* c, int cycles, uint64_t seed);
/* the same with `n_carried` drops (device buffer) appended to the spawned batch: the drops the
 * neighbours handed over at the end of the previous call, when strips exchange once per call */
int shx_strip_erode_begin_with(shx_ctx* c, int cycles, uint64_t seed, const shx_drop* dev_carried, size_t n_carried);
int shx_strip_erode_end(shx_ctx* c);

/* ---- multi-GPU from ONE host thread: what the reference's single frame loop (SimpleHydrology.cpp:314-324) can call.
 * shx_multi owns one strip context per device (row strips of whole tile rows: mapsize must be divisible by ngpu).
 * shx_multi_erode == World::erode(cycles) on all strips at once; the strips meet ONCE per call: each strip's pack
 * kernels store its message (border-crossing drops, the height deltas it put into its halo rows, its edge rows)
 * straight into the neighbour's inbox over NVLink (peer access; cudaMemcpyPeerAsync where a pair of devices has
 * none), CUDA events order pack -> apply across devices, no collective and no host wait inside a call.  A drop that
 * leaves its strip pauses until the next call.  Deterministic for a given ngpu; statistically (not bitwise) the
 * single-GPU result (tests/test_gpu_strips.py, tests/test_gpu_multi.py).  `devices` == NULL: ordinals 0..ngpu-1
 * (wrapping around the visible devices, so k logical strips can share one GPU); `base` may be NULL.
 * cycles must not exceed max_cycles_per_launch (512) when ngpu > 1.  ngpu == 1 is a plain whole-map context. */
typedef struct shx_multi shx_multi;
const char* shx_multi_last_error(void);
int shx_multi_create(shx_multi** out, const shx_params* p, int ngpu, const int* devices, const shx_config* base);
void shx_multi_destroy(shx_multi* m);
int shx_multi_strips(const shx_multi* m);
shx_ctx* shx_multi_strip(shx_multi* m, int i); /* the i-th strip's context (views, gathers, timing ...) */
int shx_multi_upload(shx_multi* m, const shx_cell* pool, size_t ncells);
int shx_multi_download(shx_multi* m, shx_cell* pool, size_t ncells, unsigned field_mask); /* all devices copy concurrently */
int shx_multi_init_terrain(shx_multi* m, int seed);
int shx_multi_synth_terrain(shx_multi* m, uint32_t seed);
int shx_multi_set_params(shx_multi* m, const shx_params* p);
int shx_multi_set_rootdensity(shx_multi* m, const int* xy, const float* value, size_t n);
int shx_multi_erode(shx_multi* m, int cycles, uint64_t seed, shx_stats* out /* summed over strips; may be NULL */);
int shx_multi_erode_async(shx_multi* m, int cycles, uint64_t seed);
int shx_multi_read_stats(shx_multi* m, shx_stats* out);
int shx_multi_in_flight(shx_multi* m, size_t* ndrops); /* drops waiting for the next call */
int shx_multi_sync(shx_multi* m);

/* ---- peer mode (shx_config.peer_world > 1): export this rank's strip, map everybody's */
int shx_peer_export(shx_ctx* c, shx_peer_handles* out);
int shx_peer_attach(shx_ctx* c, const shx_peer_handles* all_ranks /* peer_world entries, index = rank */);

#ifdef __cplusplus
}
#endif
#endif /* SHX_H */
