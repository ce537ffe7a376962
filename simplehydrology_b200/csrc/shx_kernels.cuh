// shx CUDA kernels (sm_100a).  Layout in HBM (per context, rows [xlo, xlo+nrows) of the map):
//   h[0], h[1]  int32 Q5.26 height planes, index (x-xlo)*size + y   (x-major like the reference,
//               math.h:11-14, but one global plane instead of 512^2 tiles).  Two planes so a
//               phase can read one while every drop adds into the other (see descend_lockstep).
//   field       float4 {discharge, momentumx, momentumy, rootdensity}  -- read-only inside erode
//   track       32-byte record {int64 discharge, momentumx, momentumy, pad}, Q31.32 accumulators,
//               one sector per cell so the three REDs of a step hit one L2 sector
// Sequential mode reuses h[0] as an fp32 plane and `track` as float4.
#pragma once
#include <cooperative_groups.h>

#include "shx_step.cuh"

namespace shx {

struct __align__(32) Track {
  long long discharge, momentumx, momentumy, pad;
};

struct MapView {
  int32_t* h[2];
  float4* field;
  Track* track;
  int size;        // cells per side of the whole map
  int xlo, nrows;  // stored rows
  int row0, row1;  // owned rows
};

enum StatIndex {
  ST_SPAWNED, ST_REJECTED, ST_STEPS, ST_TERM_AGE, ST_TERM_VOL, ST_TERM_OOB, ST_TRANSFERS, ST_PHASES,
  ST_FX_ERODED, ST_FX_DEPOSITED, ST_FX_SED_OOB, ST_FX_SED_DEPOSITED, ST_FX_SED_INFLATION,
  ST_MIGRATED_LO, ST_MIGRATED_HI, ST_LAUNCHES, ST_COUNT
};
static_assert(sizeof(shx_stats) == ST_COUNT * 8, "shx_stats layout");

struct GridBar {
  unsigned count;
  unsigned active[4];
  unsigned max_steps;  // longest drop of this launch == number of phases that had a live drop
  unsigned pad[2];
};

struct DescendArgs {
  MapView m;
  StepParams P;
  shx_drop* drops;
  unsigned ndrops;
  GridBar* bar;
  unsigned long long* stats;
  float* trace;  // 7 floats per phase of drop 0, or null
  int trace_cap;
  int* trace_n;
};

__device__ __forceinline__ void stat_add(unsigned long long* stats, int i, unsigned long long v) {
  atomicAdd(stats + i, v);
}

// ---------------------------------------------------------------------------------------------
// Grid-wide barrier that also sums a per-CTA count (drops still active).  One arrival per CTA on
// a monotonically increasing counter in L2; `active` is a 4-slot ring so that the sum of phase p
// can be read after the barrier while phase p+1 is already accumulating.  The kernel is launched
// cooperatively (all CTAs co-resident).  Returns the grid-wide sum.
__device__ __forceinline__ unsigned grid_barrier_sum(GridBar* bar, unsigned phase, unsigned block_sum, unsigned* s_total) {
  // caller has just executed a __syncthreads-class barrier (block_sum comes from __syncthreads_count),
  // so every RED of this CTA for this phase has been issued before thread 0 fences.
  if (gridDim.x == 1) return block_sum;
  if (threadIdx.x == 0) {
    const unsigned slot = phase & 3u;
    if (block_sum) atomicAdd(&bar->active[slot], block_sum);
    __threadfence();
    atomicAdd(&bar->count, 1u);
    const unsigned target = (phase + 1u) * gridDim.x;
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&bar->count) : "memory");
    } while (seen < target);
    unsigned total;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(total) : "l"(&bar->active[slot]) : "memory");
    if (blockIdx.x == 0) bar->active[(phase + 2u) & 3u] = 0u;
    *s_total = total;
  }
  __syncthreads();
  return *s_total;
}

// ---------------------------------------------------------------------------------------------
// K3: batched lock-step descend.  One thread per drop, state in registers.  Phase p:
//   * reads heights only from plane p&1 (never written during the phase),
//   * adds this phase's integer height deltas to plane (p+1)&1, together with the deltas of
//     phase p-1 ("catch-up": that plane was the read plane of phase p-1 and has not seen them),
//   * adds volume / momentum to the Q31.32 track records (write-only inside erode),
//   * one grid barrier.
// After the barrier plane (p+1)&1 holds exactly "heights after phase p", so every read is
// independent of thread timing and every write is an integer add: results do not depend on the
// order in which drops are scheduled and are run-to-run identical.
template <int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) descend_lockstep_kernel(const __grid_constant__ DescendArgs a) {
  extern __shared__ int32_t s_pend[];  // [9][blockDim.x] deltas of the previous phase
  __shared__ unsigned s_total;
  const int tid = threadIdx.x, nt = blockDim.x;
  const unsigned gid = blockIdx.x * nt + tid;
  const int size = a.m.size;

  DropRegs d;
  d.px = d.py = d.sx = d.sy = d.vol = d.sed = 0.0f;
  d.age = 0;
  d.flags = 0;
  if (gid < a.ndrops) {
    const float4 lo = reinterpret_cast<const float4*>(a.drops + gid)[0];
    const float4 hi = reinterpret_cast<const float4*>(a.drops + gid)[1];
    d.px = lo.x; d.py = lo.y; d.sx = lo.z; d.sy = lo.w;
    d.vol = hi.x; d.sed = hi.y; d.age = __float_as_int(hi.z); d.flags = __float_as_int(hi.w);
  }
  bool alive = (d.flags & SHX_DROP_ALIVE) != 0;
  bool pend_any = false;
  long long pidx = 0;
  StepAcc acc = {0u, 0u, 0ll, 0ll};
  int tn = 0;

  for (unsigned phase = 0;; ++phase) {
    const int32_t* R = a.m.h[phase & 1u];
    int32_t* W = a.m.h[(phase + 1u) & 1u];

    if (pend_any) {  // catch-up of the previous phase's deltas into the plane read back then
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const int32_t v = s_pend[k * nt + tid];
        if (v) atomicAdd(W + pidx + (long long)(k / 3 - 1) * size + (k % 3 - 1), v);
      }
      pend_any = false;
    }

    if (alive) {
      const int ix = (int)d.px, iy = (int)d.py;  // water.h:60, truncation
      const long long cidx = (long long)(ix - a.m.xlo) * size + iy;
      unsigned inb = 0;
      int32_t B0[9], B[9];
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const int x = ix + k / 3 - 1, y = iy + k % 3 - 1;
        const bool in = x >= 0 && y >= 0 && x < size && y < size;  // cellpool.h:413-419
        inb |= in ? (1u << k) : 0u;
        B0[k] = in ? __ldcg(R + cidx + (long long)(k / 3 - 1) * size + (k % 3 - 1)) : 0;
        B[k] = B0[k];
      }
      const float4 fld = __ldg(a.m.field + cidx);
      StepOut out;
      phase_step<HeightQ>(B, inb, d, fld, a.P, size, ix, iy,
                          [&](int nx, int ny) { return __ldcg(R + (long long)(nx - a.m.xlo) * size + ny); }, acc, out);
      if (out.deposit) {
        Track* t = a.m.track + cidx;
        atomicAdd(reinterpret_cast<unsigned long long*>(&t->discharge), (unsigned long long)t_quantize(out.t_d));
        atomicAdd(reinterpret_cast<unsigned long long*>(&t->momentumx), (unsigned long long)t_quantize(out.t_mx));
        atomicAdd(reinterpret_cast<unsigned long long*>(&t->momentumy), (unsigned long long)t_quantize(out.t_my));
      }
      if (!(d.flags & SHX_DROP_ALIVE)) {  // terminated in this phase (rare: once per drop)
        alive = false;
        atomicMax(&a.bar->max_steps, acc.steps);
        if (d.flags & SHX_DROP_DONE_OOB) {
          stat_add(a.stats, ST_TERM_OOB, 1ull);
          stat_add(a.stats, ST_FX_SED_OOB, (unsigned long long)t_quantize(d.sed));
        } else {
          stat_add(a.stats, (d.flags & SHX_DROP_DONE_AGE) ? ST_TERM_AGE : ST_TERM_VOL, 1ull);
          stat_add(a.stats, ST_FX_DEPOSITED, (unsigned long long)out.fx_event);
          stat_add(a.stats, ST_FX_SED_DEPOSITED, (unsigned long long)t_quantize(d.sed));
        }
      } else {
        const int nix = (int)d.px;
        if (nix < a.m.row0) {  // left the strip: hand over to the neighbour (cascade still owed)
          d.flags = (d.flags & ~SHX_DROP_ALIVE) | SHX_DROP_MIGRATE_LO;
          alive = false;
          atomicMax(&a.bar->max_steps, acc.steps);
          stat_add(a.stats, ST_MIGRATED_LO, 1ull);
        } else if (nix >= a.m.row1) {
          d.flags = (d.flags & ~SHX_DROP_ALIVE) | SHX_DROP_MIGRATE_HI;
          alive = false;
          atomicMax(&a.bar->max_steps, acc.steps);
          stat_add(a.stats, ST_MIGRATED_HI, 1ull);
        }
      }
      bool any = false;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const int32_t dv = B[k] - B0[k];
        s_pend[k * nt + tid] = dv;
        if (dv) {
          atomicAdd(W + cidx + (long long)(k / 3 - 1) * size + (k % 3 - 1), dv);
          any = true;
        }
      }
      pend_any = any;
      pidx = cidx;
      if (a.trace != nullptr && gid == 0 && tn < a.trace_cap) {
        float* t = a.trace + 7 * (size_t)tn++;
        t[0] = (float)d.age; t[1] = d.px; t[2] = d.py; t[3] = d.sx; t[4] = d.sy; t[5] = d.vol; t[6] = d.sed;
      }
    }

    const unsigned block_sum = (unsigned)__syncthreads_count(alive || pend_any);
    const unsigned total = grid_barrier_sum(a.bar, phase, block_sum, &s_total);
    if (total == 0u) {  // every termination's atomicMax happened before the barrier just passed
      if (gid == 0) stat_add(a.stats, ST_PHASES, (unsigned long long)__ldcg(&a.bar->max_steps));
      break;
    }
  }

  if (gid < a.ndrops) {
    float4 lo, hi;
    lo.x = d.px; lo.y = d.py; lo.z = d.sx; lo.w = d.sy;
    hi.x = d.vol; hi.y = d.sed; hi.z = __int_as_float(d.age); hi.w = __int_as_float(d.flags);
    reinterpret_cast<float4*>(a.drops + gid)[0] = lo;
    reinterpret_cast<float4*>(a.drops + gid)[1] = hi;
  }
  if (a.trace_n != nullptr && gid == 0) *a.trace_n = tn;

  // per-step counters: warp reduce, one atomic per warp
  unsigned steps = acc.steps, transfers = acc.transfers;
  long long er = acc.fx_eroded, inf = acc.fx_sed_inflation;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    steps += __shfl_xor_sync(0xffffffffu, steps, o);
    transfers += __shfl_xor_sync(0xffffffffu, transfers, o);
    er += __shfl_xor_sync(0xffffffffu, er, o);
    inf += __shfl_xor_sync(0xffffffffu, inf, o);
  }
  if ((tid & 31) == 0 && steps) {
    stat_add(a.stats, ST_STEPS, steps);
    stat_add(a.stats, ST_TRANSFERS, transfers);
    stat_add(a.stats, ST_FX_ERODED, (unsigned long long)er);
    stat_add(a.stats, ST_FX_SED_INFLATION, (unsigned long long)inf);
  }
}

// ---------------------------------------------------------------------------------------------
// Sequential mode (parity anchor): one thread marches the drops one after another in fp32,
// operation for operation what World::erode's inner loop does (world.h:74-76).  <<<1,1>>>.
struct SequentialArgs {
  float* h;       // fp32 height plane
  float4* field;
  float4* trackf; // {discharge_track, momentumx_track, momentumy_track, -}
  int size;
  StepParams P;
  shx_drop* drops;
  unsigned ndrops;
  unsigned long long* stats;
  float* trace;
  int trace_cap;
  int* trace_n;
};

__global__ void descend_sequential_kernel(const SequentialArgs a) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int size = a.size;
  StepAcc acc = {0u, 0u, 0ll, 0ll};
  int tn = 0;
  for (unsigned i = 0; i < a.ndrops; i++) {
    const shx_drop r = a.drops[i];
    DropRegs d = {r.px, r.py, r.sx, r.sy, r.volume, r.sediment, r.age, r.flags};
    while (d.flags & SHX_DROP_ALIVE) {
      const int ix = (int)d.px, iy = (int)d.py;
      const long long cidx = (long long)ix * size + iy;
      unsigned inb = 0;
      float B[9];
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const int x = ix + k / 3 - 1, y = iy + k % 3 - 1;
        const bool in = x >= 0 && y >= 0 && x < size && y < size;
        inb |= in ? (1u << k) : 0u;
        B[k] = in ? a.h[cidx + (long long)(k / 3 - 1) * size + (k % 3 - 1)] : 0.0f;
      }
      const float4 fld = a.field[cidx];
      StepOut out;
      phase_step<HeightF>(B, inb, d, fld, a.P, size, ix, iy,
                          [&](int nx, int ny) { return a.h[(long long)nx * size + ny]; }, acc, out);
#pragma unroll
      for (int k = 0; k < 9; k++)
        if (inb & (1u << k)) a.h[cidx + (long long)(k / 3 - 1) * size + (k % 3 - 1)] = B[k];
      if (out.deposit) {  // water.h:115-117
        float4 t = a.trackf[2 * cidx];  // 32-byte records, fp32 tracks in the first half
        t.x += out.t_d; t.y += out.t_mx; t.z += out.t_my;
        a.trackf[2 * cidx] = t;
      }
      if (!(d.flags & SHX_DROP_ALIVE)) {
        if (d.flags & SHX_DROP_DONE_OOB) {
          a.stats[ST_TERM_OOB] += 1ull;
          a.stats[ST_FX_SED_OOB] += (unsigned long long)t_quantize(d.sed);
        } else {
          a.stats[(d.flags & SHX_DROP_DONE_AGE) ? ST_TERM_AGE : ST_TERM_VOL] += 1ull;
          a.stats[ST_FX_SED_DEPOSITED] += (unsigned long long)t_quantize(d.sed);
        }
      }
      if (a.trace != nullptr && i == 0 && tn < a.trace_cap) {
        float* t = a.trace + 7 * (size_t)tn++;
        t[0] = (float)d.age; t[1] = d.px; t[2] = d.py; t[3] = d.sx; t[4] = d.sy; t[5] = d.vol; t[6] = d.sed;
      }
    }
    shx_drop w = {d.px, d.py, d.sx, d.sy, d.vol, d.sed, d.age, d.flags};
    a.drops[i] = w;
  }
  if (a.trace_n != nullptr) *a.trace_n = tn;
  a.stats[ST_STEPS] += acc.steps;
  a.stats[ST_TRANSFERS] += acc.transfers;
  a.stats[ST_FX_SED_INFLATION] += (unsigned long long)acc.fx_sed_inflation;
}

// ---------------------------------------------------------------------------------------------
// K2: spawn.  world.h:64-74 with rand() replaced by a counter-based hash keyed
// (seed, epoch, node, i): node-major, `cycles` drops per node, reject where height < 0.1.
// node0/nnodes select the nodes of this strip (all of them for a whole map).
struct SpawnArgs {
  const int32_t* hq;  // batched: Q5.26 plane;  sequential: fp32 plane (hf)
  const float* hf;
  int size, xlo, tilesize, mapsize;
  unsigned node0, nnodes;
  int cycles;
  uint64_t key;
  shx_drop* drops;
  float* xy;  // optional copy of the positions
  unsigned long long* stats;
};

__device__ __forceinline__ shx_drop make_drop(float x, float y, const int32_t* hq, const float* hf, int size, int xlo,
                                              unsigned long long* stats, int row0, int row1) {
  shx_drop d = {x, y, 0.0f, 0.0f, 1.0f, 0.0f, 0, SHX_DROP_ALIVE};  // water.h:14-23
  const int ix = (int)x, iy = (int)y;
  const bool oob = !(x > -1.0f) || !(y > -1.0f) || ix >= size || iy >= size;
  if (!oob && (ix < row0 || ix >= row1)) {  // not this strip's drop
    d.flags = 0;
    return d;
  }
  float h = 0.0f;  // map.height() of a missing cell (cellpool.h:433-437)
  if (!oob) {
    const long long idx = (long long)(ix - xlo) * size + iy;
    h = hq ? h_to_float(hq[idx]) : hf[idx];
  }
  if ((double)h < 0.1) {  // world.h:71-72
    d.flags = SHX_DROP_REJECTED;
    atomicAdd(stats + ST_REJECTED, 1ull);
  } else {
    atomicAdd(stats + ST_SPAWNED, 1ull);
  }
  return d;
}

__global__ void spawn_kernel(const SpawnArgs a, int row0, int row1) {
  const unsigned n = a.nnodes * (unsigned)a.cycles;
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const unsigned node = a.node0 + k / (unsigned)a.cycles, i = k % (unsigned)a.cycles;
    const uint64_t r = mix64(a.key + (((uint64_t)node << 32) | (uint64_t)i));
    const int nx = (int)(node / (unsigned)a.mapsize) * a.tilesize, ny = (int)(node % (unsigned)a.mapsize) * a.tilesize;
    const float x = (float)(nx + (int)((uint32_t)r % (uint32_t)a.tilesize));
    const float y = (float)(ny + (int)((uint32_t)(r >> 32) % (uint32_t)a.tilesize));
    if (a.xy) { a.xy[2 * k] = x; a.xy[2 * k + 1] = y; }
    a.drops[k] = make_drop(x, y, a.hq, a.hf, a.size, a.xlo, a.stats, row0, row1);
  }
}

__global__ void make_drops_kernel(const float* xy, unsigned n, const int32_t* hq, const float* hf, int size, int xlo,
                                  int row0, int row1, shx_drop* drops, unsigned long long* stats) {
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    drops[k] = make_drop(xy[2 * k], xy[2 * k + 1], hq, hf, size, xlo, stats, row0, row1);
}

// ---------------------------------------------------------------------------------------------
// K4: EMA of the discharge / momentum maps (world.h:81-86), streaming over the owned rows.
// Per cell: read 32 B track record + 16 B field, write 16 B field.
__global__ void ema_kernel(float4* __restrict__ field, const Track* __restrict__ track, size_t n, float lrate) {
  const float keep = 1.0f - lrate;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const longlong4 t = *reinterpret_cast<const longlong4*>(track + i);
    float4 f = field[i];
    f.x = keep * f.x + lrate * t_to_float(t.x);
    f.y = keep * f.y + lrate * t_to_float(t.y);
    f.z = keep * f.z + lrate * t_to_float(t.z);
    field[i] = f;
  }
}

__global__ void ema_sequential_kernel(float4* __restrict__ field, const float4* __restrict__ trackf, size_t n, float lrate) {
  const float keep = 1.0f - lrate;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 t = trackf[2 * i];
    float4 f = field[i];
    f.x = keep * f.x + lrate * t.x;
    f.y = keep * f.y + lrate * t.y;
    f.z = keep * f.z + lrate * t.z;
    field[i] = f;
  }
}

// ---------------------------------------------------------------------------------------------
// Boundary conversion: one 512^2 tile of the host's tiled AoS pool (32 B quad::cell records,
// x-major inside the tile) <-> the planar device layout.  Thread per cell; both sides coalesced
// (consecutive threads = consecutive y).
struct TileArgs {
  MapView m;
  int sequential;
  int tilesize, tx0, ty0;  // tile origin in world cells
  int* error_flag;
};

__global__ void unpack_tile_kernel(const TileArgs a, const shx_cell* __restrict__ aos) {
  const int ts = a.tilesize;
  const int n = ts * ts;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    const int x = a.tx0 + c / ts, y = a.ty0 + c % ts;
    if (x < a.m.xlo || x >= a.m.xlo + a.m.nrows) continue;
    const float4 lo = reinterpret_cast<const float4*>(aos + c)[0];  // height discharge momentumx momentumy
    const float4 hi = reinterpret_cast<const float4*>(aos + c)[1];  // tracks x3, rootdensity
    const size_t i = (size_t)(x - a.m.xlo) * a.m.size + y;
    a.m.field[i] = make_float4(lo.y, lo.z, lo.w, hi.w);
    if (a.sequential) {
      reinterpret_cast<float*>(a.m.h[0])[i] = lo.x;
      reinterpret_cast<float4*>(a.m.track)[2 * i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    } else {
      if (!(fabsf(lo.x) < 31.0f)) *a.error_flag = 1;
      const int32_t q = h_quantize(lo.x);
      a.m.h[0][i] = q;
      a.m.h[1][i] = q;
      longlong4 t;
      t.x = t_quantize(hi.x); t.y = t_quantize(hi.y); t.z = t_quantize(hi.z); t.w = 0;
      *reinterpret_cast<longlong4*>(a.m.track + i) = t;
    }
  }
}

__global__ void pack_tile_kernel(const TileArgs a, shx_cell* __restrict__ aos) {
  const int ts = a.tilesize;
  const int n = ts * ts;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    const int x = a.tx0 + c / ts, y = a.ty0 + c % ts;
    if (x < a.m.row0 || x >= a.m.row1) continue;
    const size_t i = (size_t)(x - a.m.xlo) * a.m.size + y;
    const float4 f = a.m.field[i];
    float4 lo, hi;
    if (a.sequential) {
      const float4 t = reinterpret_cast<const float4*>(a.m.track)[2 * i];
      lo = make_float4(reinterpret_cast<const float*>(a.m.h[0])[i], f.x, f.y, f.z);
      hi = make_float4(t.x, t.y, t.z, f.w);
    } else {
      const longlong4 t = *reinterpret_cast<const longlong4*>(a.m.track + i);
      lo = make_float4(h_to_float(a.m.h[0][i]), f.x, f.y, f.z);
      hi = make_float4(t_to_float(t.x), t_to_float(t.y), t_to_float(t.z), f.w);
    }
    reinterpret_cast<float4*>(aos + c)[0] = lo;
    reinterpret_cast<float4*>(aos + c)[1] = hi;
  }
}

// ---------------------------------------------------------------------------------------------
// Plant::root stamps (vegetation.h:87-118).  Applied by ONE thread in list order so that several
// stamps on one cell add up in the same fp32 order as the host's sequential `+=`.
__global__ void add_rootdensity_kernel(const MapView m, const int* xy, const float* delta, size_t n) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (size_t i = 0; i < n; i++) {
    const int x = xy[2 * i], y = xy[2 * i + 1];
    if (x < m.xlo || x >= m.xlo + m.nrows || y < 0 || y >= m.size) continue;  // getCell() == NULL -> skipped
    float* w = &m.field[(size_t)(x - m.xlo) * m.size + y].w;
    *w = *w + delta[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Synthetic seeded terrain: hash-lattice value noise, 8 octaves (wavelength 256..2 cells,
// amplitude 0.6^o -- the reference's layer weights, cellpool.h:361-376), then the reference's
// min/max normalisation (cellpool.h:382-408).  Same arithmetic as oracle orc_synth_terrain.
__device__ __forceinline__ uint32_t hash2(uint32_t x, uint32_t y, uint32_t s) {
  uint32_t h = x * 0x9E3779B1u ^ y * 0x85EBCA77u ^ s * 0xC2B2AE3Du;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
__device__ __forceinline__ float lattice(uint32_t x, uint32_t y, uint32_t s) {
  return (float)(hash2(x, y, s) >> 8) * (1.0f / 8388608.0f) - 1.0f;
}
__device__ __forceinline__ float synth_raw(int x, int y, uint32_t seed) {
  float sum = 0.0f, amp = 0.6f;
  int cell = 256;
#pragma unroll 1
  for (int o = 0; o < 8; o++) {
    const int gx = x / cell, gy = y / cell;
    const float fx = (float)(x % cell) / (float)cell, fy = (float)(y % cell) / (float)cell;
    const float ux = fx * fx * (3.0f - 2.0f * fx), uy = fy * fy * (3.0f - 2.0f * fy);
    const uint32_t s = seed * 8u + (uint32_t)o;
    const float v00 = lattice((uint32_t)gx, (uint32_t)gy, s), v01 = lattice((uint32_t)gx, (uint32_t)gy + 1u, s);
    const float v10 = lattice((uint32_t)gx + 1u, (uint32_t)gy, s), v11 = lattice((uint32_t)gx + 1u, (uint32_t)gy + 1u, s);
    const float p = v00 + (v01 - v00) * uy, q = v10 + (v11 - v10) * uy;
    sum = sum + amp * (p + (q - p) * ux);
    amp = amp * 0.6f;
    cell >>= 1;
  }
  return sum;
}
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// pass 1: global min/max over the WHOLE map (every strip computes the same pair)
__global__ void synth_minmax_kernel(int size, uint32_t seed, unsigned* mnmx) {
  unsigned mn = 0xffffffffu, mx = 0u;
  const size_t n = (size_t)size * size;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned o = f2ord(synth_raw((int)(i / size), (int)(i % size), seed));
    mn = min(mn, o);
    mx = max(mx, o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mnmx, mn);
    atomicMax(mnmx + 1, mx);
  }
}

// pass 2: normalise and store (heights both planes; all other fields zero)
__global__ void synth_fill_kernel(const MapView m, int sequential, uint32_t seed, const unsigned* mnmx) {
  const float mn = ord2f(mnmx[0]), mx = ord2f(mnmx[1]);
  const float range = mx - mn;
  const size_t n = (size_t)m.nrows * m.size;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = m.xlo + (int)(i / m.size), y = (int)(i % m.size);
    const float h = (synth_raw(x, y, seed) - mn) / range;
    m.field[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    longlong4 z;
    z.x = z.y = z.z = z.w = 0;
    *reinterpret_cast<longlong4*>(m.track + i) = z;
    if (sequential) {
      reinterpret_cast<float*>(m.h[0])[i] = h;
    } else {
      const int32_t q = h_quantize(h);
      m.h[0][i] = q;
      m.h[1][i] = q;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row-strip exchange helpers (multi-GPU).  A strip keeps `halo` rows of its neighbours' heights on
// each side.  Cascade transfers of drops on the strip's boundary rows land in those halo rows;
// `halo_ref` remembers what the halo held at the last refresh, so (current - ref) is exactly the
// integer amount this strip owes the owner.
__global__ void strip_halo_delta_kernel(const int32_t* cur, const int32_t* ref, int32_t* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = cur[i] - ref[i];
}
__global__ void strip_add_rows_kernel(int32_t* h0, int32_t* h1, const int32_t* delta, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int32_t v = delta[i];
    if (v) { h0[i] += v; h1[i] += v; }
  }
}
__global__ void strip_copy_rows_kernel(int32_t* dst0, int32_t* dst1, int32_t* ref, const int32_t* src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int32_t v = src[i];
    dst0[i] = v;
    if (dst1) dst1[i] = v;
    if (ref) ref[i] = v;
  }
}
// compact the drops that left the strip into two outboxes (order is irrelevant to the result:
// every scatter downstream is an integer add)
__global__ void strip_pack_migrants_kernel(const shx_drop* drops, unsigned n, shx_drop* lo, shx_drop* hi, unsigned cap,
                                           unsigned* counts) {
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    shx_drop d = drops[k];
    if (d.flags & (SHX_DROP_MIGRATE_LO | SHX_DROP_MIGRATE_HI)) {
      const bool tolo = (d.flags & SHX_DROP_MIGRATE_LO) != 0;
      const unsigned slot = atomicAdd(counts + (tolo ? 0 : 1), 1u);
      d.flags = (d.flags & ~(SHX_DROP_MIGRATE_LO | SHX_DROP_MIGRATE_HI)) | SHX_DROP_ALIVE;
      if (slot < cap) (tolo ? lo : hi)[slot] = d;
    }
  }
}

}  // namespace shx
