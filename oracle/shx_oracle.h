/* TEST INFRASTRUCTURE -- CPU oracle for the erosion hot path.
 *
 * Plain-C restatement of the reference's algorithm for the path
 *   World::erode -> Drop::descend -> World::cascade over the cellpool map
 * (reference: source/world.h:54-168, source/water.h:58-156,
 * source/cellpool.h:181-204,207-220,413-447, source/include/math.h:11-14).
 *
 * Nothing under oracle/ is part of the shipped library.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load it, and only as the checker.
 *
 * Pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference's OWN headers compiled headless
 * (oracle/_ref, built by oracle/Makefile from /root/reference where it lies):
 * tests/test_oracle_vs_ref.py requires bit-identical cell buffers and drop traces,
 * and tests/golden/ holds fixtures generated from oracle/_ref by
 * tests/golden/make_golden.py for machines without the reference tree.
 *
 * Two families of functions:
 *   orc_seq_*  sequential fp32 semantics, operation for operation what the
 *              reference executes (drops strictly one after another).
 *   orc_ls_*   the SAME per-step arithmetic (same cited lines) under the batched
 *              "lock-step" schedule the CUDA path uses: all drops advance one
 *              step per phase, heights are Q5.26 integers, every scatter is an
 *              integer add, so the result is independent of drop order.  This is
 *              the bit-exact checker for the batched kernels.
 */
#ifndef SHX_ORACLE_H
#define SHX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == quad::cell, cellpool.h:207-220 (8 x f32 = 32 B, this order) */
typedef struct {
  float height, discharge, momentumx, momentumy;
  float discharge_track, momentumx_track, momentumy_track;
  float rootdensity;
} orc_cell;

/* Drop:: statics water.h:43-50, World:: statics world.h:42-44, geometry cellpool.h:165-179 */
typedef struct {
  float maxAge, minVol, evapRate, depositionRate, entrainment, gravity, momentumTransfer;
  float lrate, maxdiff, settling;
  int mapscale, tilesize, mapsize, lodsize;
} orc_params;

/* struct Drop, water.h:12-39 (+ bookkeeping the batched schedule needs) */
typedef struct {
  float px, py, sx, sy, volume, sediment;
  int age;
  int flags; /* ORC_DROP_* */
} orc_drop;

enum {
  ORC_DROP_ALIVE = 1,        /* still marching */
  ORC_DROP_CASCADE = 2,      /* World::cascade(pos) of the previous step still owed (lock-step only) */
  ORC_DROP_DONE_AGE = 4,     /* water.h:74-77 */
  ORC_DROP_DONE_VOL = 8,     /* water.h:79-82 */
  ORC_DROP_DONE_OOB = 16,    /* water.h:139-142 */
  ORC_DROP_REJECTED = 32,    /* world.h:71-72 */
  ORC_DROP_DONE_NULL = 64,   /* water.h:62-68 (spawned outside the map) */
  ORC_DROP_MIGRATE_LO = 128, /* left the strip towards smaller x (multi-GPU hand-off) */
  ORC_DROP_MIGRATE_HI = 256,
  ORC_DROP_WAITED_SHIFT = 16, /* bits 16-18: phases the drop has waited for its cell (lock-step exclusion), saturating at 7 */
  ORC_DROP_FREEW_SHIFT = 19   /* bits 19-22: waits of this drop that did not cost it a step (at most free_waits <= 15) */
};

typedef struct {
  uint64_t spawned, rejected, steps, term_age, term_vol, term_oob, cascade_transfers, phases;
  int64_t fx_eroded;    /* sum of q(effD*cdiff)  (terrain -> drops), height fixed point */
  int64_t fx_deposited; /* sum of q(sediment) put back at age/volume termination   */
  /* Q31.32 sums of fp32 sediment amounts (each term rounded to 2^-32 once, then exact) */
  int64_t fx_sed_oob_lost;  /* sediment carried out of the map (water.h:139-142)   */
  int64_t fx_sed_deposited; /* sediment handed back at termination (water.h:74-82) */
  int64_t fx_sed_inflation; /* growth of carried sediment by water.h:135           */
} orc_stats;

#define ORC_HEIGHT_FRAC_BITS 26
#define ORC_TRACK_FRAC_BITS 18
#define ORC_LEDGER_FRAC_BITS 32

void orc_default_params(orc_params* p, int mapsize);

/* index of world cell (x,y) inside the tiled AoS pool: node (x/ts)*mapsize+(y/ts)
 * owns a contiguous tilearea slice (cellpool.h:327-336), x-major inside (math.h:11-14) */
size_t orc_tiled_index(const orc_params* p, int x, int y);

float orc_erff_libm(float x);  /* what the reference calls (cellpool.h:243) */
float orc_erff_poly(float x);  /* restatement of the kernels' own erf (see tools/fit_erf.py) */

/* ---- sequential fp32 semantics on the tiled AoS pool (the reference's own layout) */
typedef struct {
  orc_params p;
  orc_cell* cells; /* tiled AoS, caller-owned, p.mapsize^2 * p.tilesize^2 records */
  int erf_poly;    /* 0: libm erff (reference-exact)  1: orc_erff_poly (kernel-exact) */
} orc_seq_world;

int orc_seq_oob(const orc_seq_world* w, int x, int y);                 /* cellpool.h:413-419 */
float orc_seq_height(const orc_seq_world* w, int x, int y);            /* cellpool.h:433-437 */
void orc_seq_normal(const orc_seq_world* w, int x, int y, float* n3);  /* cellpool.h:181-204 */
uint32_t orc_seq_cascade(orc_seq_world* w, float px, float py);        /* world.h:90-168; returns #transfers */
int orc_seq_descend(orc_seq_world* w, orc_drop* d, orc_stats* st);     /* water.h:58-156; 1 = still alive */
void orc_seq_reset_tracks(orc_seq_world* w);                           /* world.h:56-61 */
void orc_seq_ema(orc_seq_world* w);                                    /* world.h:81-86 */
/* world.h:54-88 with the rand() spawn replaced by an explicit list */
void orc_seq_erode_spawnlist(orc_seq_world* w, const float* xy, size_t n, int do_reset, int do_ema, orc_stats* st);
/* 7 floats per descend call: age,px,py,sx,sy,volume,sediment */
int orc_seq_trace_drop(orc_seq_world* w, float x, float y, float* trace, int max_calls);

/* ---- lock-step fixed-point semantics on planar row-major planes */
typedef struct {
  int32_t discharge, momentumx, momentumy, pad;
} orc_track;

typedef struct {
  orc_params p;
  int size;         /* cells per side = mapsize*tilesize */
  int32_t* h[2];    /* height planes, Q5.26, index x*size+y; equal outside a batch */
  float* field;     /* 4 floats per cell: discharge, momentumx, momentumy, rootdensity */
  orc_track* track; /* Q13.18 accumulators */
  int row0, row1;   /* rows [row0,row1) are owned (strip); 0,size for the whole map */
  int align_age;    /* != 0: a drop of age a sleeps until phase a (drops carried over from the previous call) */
  int max_cycles_per_launch; /* orc_ls_erode: drops per node that march together (0 = 512), shx_config's field */
  int exclusive_cells; /* 1: of the drops that stand on the same cell in a phase only the holder of the highest
                          claim key steps, the others wait for the next phase; 2: a drop also waits while a
                          higher key stands on one of the eight cells around it; 3 (default): such a drop steps,
                          with its sediment exchange halved per such cell; 0: no turn-taking */
  float cur_damp;      /* internal: factor on the sediment exchange of the step being made */
  int free_waits;      /* the first free_waits (<= 15, default 8 = shx_config.free_waits) waits of a drop's life do not
                          advance its age; later ones cost a step each, which bounds the phases of a call */
  int recip_evap;      /* 1 (default): water.h:135 as a multiplication by the double 1/(1-evapRate), like the batched
                          kernels; 0: the reference's division */
  int steps_per_phase; /* S >= 1 steps between two global meetings; within a phase a drop reads the frozen
                          plane plus its OWN earlier deltas of the phase (0 is read as 1) */
} orc_ls_world;

orc_ls_world* orc_ls_create(const orc_params* p);
void orc_ls_destroy(orc_ls_world* w);
void orc_ls_upload(orc_ls_world* w, const orc_cell* tiled);  /* tiled AoS -> planes (quantises height) */
void orc_ls_download(const orc_ls_world* w, orc_cell* tiled); /* planes -> tiled AoS (tracks as float) */
int32_t orc_ls_quantize_height(float h);
/* spawn positions of one erode(cycles) call: node-major, `cycles` per node, counter-based hash */
void orc_ls_spawn(const orc_params* p, uint64_t seed, uint64_t epoch, int cycles, float* xy);
/* turn spawn positions into drop records (world.h:71-74: rejection by height < 0.1) */
void orc_ls_make_drops(const orc_ls_world* w, const float* xy, size_t n, orc_drop* drops, orc_stats* st);
/* march all drops to completion, one step per phase */
void orc_ls_run(orc_ls_world* w, orc_drop* drops, size_t n, orc_stats* st, float* trace0, int trace_cap, int* trace_n);
void orc_ls_reset_tracks(orc_ls_world* w); /* world.h:56-61 */
int orc_ls_ema(orc_ls_world* w, int reset); /* world.h:81-86 (+ :56-61 when reset); 1 = track overflow */
/* == erode(cycles): reset tracks, spawn, run, EMA */
void orc_ls_erode(orc_ls_world* w, int cycles, uint64_t seed, uint64_t epoch, orc_stats* st);
void orc_ls_erode_spawnlist(orc_ls_world* w, const float* xy, size_t n, orc_stats* st);

/* ---- Vegetation::grow (vegetation.h:122-188) under the device path's parallel schedule (see shx_veg_kernels.cuh):
 * bit-exact checker of shx_veg_grow.  The rootdensity count (fifths) lives in orc_track.pad. */
typedef struct { float maxSize, growRate, maxSteep, maxDischarge, maxTreeHeight; } orc_plant_params; /* vegetation.h:40-44 */
typedef struct { int x, y; float size; } orc_plant;
typedef struct { uint64_t plants, born, died, refused; } orc_veg_stats;
void orc_default_plant_params(orc_plant_params* pp);
void orc_veg_sync_counts(orc_ls_world* w);
void orc_veg_stamp_list(orc_ls_world* w, const orc_plant* plants, size_t n); /* shx_veg_upload(stamp_roots = 1) */
/* one frame; `plants` (capacity cap) is updated in place, returns the new count */
size_t orc_veg_grow(orc_ls_world* w, const orc_plant_params* pp, uint64_t seed, uint64_t frame, orc_plant* plants, size_t n,
                    size_t cap, orc_veg_stats* st);

/* ---- synthetic seeded terrain (value-noise fBm, normalised to [0,1]); planar x*size+y */
void orc_synth_terrain(float* height, int size, uint32_t seed);
/* map::init (cellpool.h:349-409): the reference's own terrain, 8 layers of 3-octave OpenSimplex2 fBm (vendored
 * FastNoiseLite.h) normalised to [0,1]; planar x*size+y.  orc_init_raw_height is one cell before normalisation. */
float orc_init_raw_height(int x, int y, int tilesize, int seed);
void orc_init_terrain(float* height, int mapsize, int tilesize, int seed);
/* planar -> tiled AoS heights (other fields zero) */
/* quad::updatenode over every node (cellpool.h:286-305): 12 floats per cell {position, normal, tangent,
 * bitangent} in pool order, from the tiled AoS cells; node-local height()/normal() as in the reference */
void orc_vertex_fill(const orc_params* p, const orc_cell* tiled, float* out);
/* dischargeMap / momentumMap values (SimpleHydrology.cpp:341-354) + height, 4 floats per cell in map
 * order (x*size + y); erf_poly as in orc_seq_world */
void orc_view_maps(const orc_params* p, const orc_cell* tiled, int erf_poly, float* out);
void orc_fill_tiled_from_planar(const orc_params* p, const float* planar, orc_cell* tiled);

#ifdef __cplusplus
}
#endif
#endif
