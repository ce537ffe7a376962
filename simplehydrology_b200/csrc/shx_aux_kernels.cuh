// shx CUDA kernels, part 2 (included by shx_kernels.cuh): spawn, EMA/reset, boundary conversion of the
// tiled AoS pool, rootdensity pushes, synthetic terrain, row-strip exchange helpers.
#pragma once

namespace shx {

// ---------------------------------------------------------------------------------------------
// K2: spawn.  world.h:64-74 with rand() replaced by a counter-based hash keyed
// (seed, epoch, node, i): node-major, `cycles` drops per node, reject where height < 0.1.
// node0/nnodes select the nodes of this strip (all of them for a whole map).
struct SpawnArgs {
  MapView m;
  int sequential;
  int tilesize, mapsize;
  unsigned node0, nnodes;
  int cycles;      // drops per node in this batch
  int i0;          // index of the batch's first drop within the node's drops of the call
  uint64_t key;
  shx_drop* drops;
  float* xy;  // optional copy of the positions
  unsigned long long* stats;
};

// twin_rank: how many earlier drops of the batch were created on the bit-identical position.  Such copies have
// identical state, hence identical claim keys, and would step together for their whole lives, each applying the full
// erosion to the same cells; the k-th copy therefore starts with `waited` = min(k, 7) and the copies take their first
// turns one after another (oracle: orc_ls_make_drops).
__device__ __forceinline__ shx_drop make_drop(float x, float y, const MapView& m, int sequential, unsigned long long* stats,
                                              unsigned twin_rank) {
  shx_drop d = {x, y, 0.0f, 0.0f, 1.0f, 0.0f, 0, SHX_DROP_ALIVE};  // water.h:14-23
  const int ix = (int)x, iy = (int)y;
  const bool oob = !(x > -1.0f) || !(y > -1.0f) || ix >= m.size || iy >= m.size;
  if (!oob && (ix < m.row0 || ix >= m.row1)) {  // not this strip's drop
    d.flags = 0;
    return d;
  }
  if (sequential) {
    // world.h:71-72 tests the height when the drop is created, i.e. after the earlier drops of the call have run:
    // descend_sequential_kernel applies the rejection (and counts it) right before the drop's first step
    d.flags |= SHX_DROP_CHECK_SPAWN;
    return d;
  }
  float h = 0.0f;  // map.height() of a missing cell (cellpool.h:433-437)
  if (!oob) h = h_to_float(m.hq[(ix - m.xlo) * m.size + iy].x);
  if (!above_tenth(h)) {  // world.h:71-72  (double)h < 0.1
    d.flags = SHX_DROP_REJECTED;
    atomicAdd(stats + ST_REJECTED, 1ull);
  } else {
    atomicAdd(stats + ST_SPAWNED, 1ull);
    d.flags |= (int)(twin_rank < 7u ? twin_rank : 7u) << kWaitedShift;
  }
  return d;
}

__global__ void spawn_kernel(const SpawnArgs a) {
  const unsigned n = a.nnodes * (unsigned)a.cycles;
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const unsigned node = a.node0 + k / (unsigned)a.cycles, i = (unsigned)a.i0 + k % (unsigned)a.cycles;
    const uint64_t base = a.key + ((uint64_t)node << 32);
    const uint64_t r = mix64(base + (uint64_t)i);
    const uint32_t lx = (uint32_t)r % (uint32_t)a.tilesize, ly = (uint32_t)(r >> 32) % (uint32_t)a.tilesize;
    unsigned twins = 0;  // earlier drops of this node's batch on the same cell (the node index decides the tile)
    if (!a.sequential)
      for (unsigned j = (unsigned)a.i0; j < i; j++) {
        const uint64_t q = mix64(base + (uint64_t)j);
        twins += ((uint32_t)q % (uint32_t)a.tilesize == lx && (uint32_t)(q >> 32) % (uint32_t)a.tilesize == ly) ? 1u : 0u;
      }
    const int nx = (int)(node / (unsigned)a.mapsize) * a.tilesize, ny = (int)(node % (unsigned)a.mapsize) * a.tilesize;
    const float x = (float)(nx + (int)lx);
    const float y = (float)(ny + (int)ly);
    if (a.xy) { a.xy[2 * k] = x; a.xy[2 * k + 1] = y; }
    a.drops[k] = make_drop(x, y, a.m, a.sequential, a.stats, twins);
  }
}

// explicit spawn list: the twin rank of entry k is the number of earlier entries with the same bits (tiles of the
// list staged through shared memory; O(n^2 / 2) compares, ~1 ms for the 131 072 drops of an 8192^2 cycle)
__global__ void make_drops_kernel(const float* xy, unsigned n, const MapView m, int sequential, shx_drop* drops,
                                  unsigned long long* stats) {
  __shared__ uint2 s_xy[256];
  const unsigned k0 = blockIdx.x * blockDim.x;
  for (unsigned base = k0; base < n; base += gridDim.x * blockDim.x) {
    const unsigned k = base + threadIdx.x;
    uint2 me = make_uint2(0u, 0u);
    if (k < n) me = make_uint2(__float_as_uint(xy[2 * k]), __float_as_uint(xy[2 * k + 1]));
    unsigned twins = 0;
    if (!sequential) {
      const unsigned last = min(n, base + blockDim.x);  // entries before the last thread of this block
      for (unsigned t0 = 0; t0 < last; t0 += 256u) {
        __syncthreads();
        const unsigned j = t0 + threadIdx.x;
        if (threadIdx.x < 256u && j < n) s_xy[threadIdx.x] = make_uint2(__float_as_uint(xy[2 * j]), __float_as_uint(xy[2 * j + 1]));
        __syncthreads();
        const unsigned lim = k < n ? min(256u, k > t0 ? k - t0 : 0u) : 0u;
        for (unsigned t = 0; t < lim; t++) twins += (s_xy[t].x == me.x && s_xy[t].y == me.y) ? 1u : 0u;
      }
    }
    if (k < n) drops[k] = make_drop(xy[2 * k], xy[2 * k + 1], m, sequential, stats, twins);
  }
}

// ---------------------------------------------------------------------------------------------
// K4 (+K1): EMA of the discharge / momentum maps (world.h:81-86) fused with the track reset
// (world.h:56-61, hoisted from the start of the next call).  Streams the owned rows: 32 B read and
// 32 B written per cell.  flags[0] is raised if a discharge accumulator left the Q13.18 range.
__global__ void ema_kernel(CellRec* __restrict__ rec, size_t n, float lrate, int sequential, int reset, int* flags) {
  const float keep = 1.0f - lrate;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 f = reinterpret_cast<const float4*>(rec + i)[0];
    int4 t = reinterpret_cast<const int4*>(rec + i)[1];
    float td, tx, ty;
    if (sequential) {
      td = __int_as_float(t.x); tx = __int_as_float(t.y); ty = __int_as_float(t.z);
    } else {
      if (t.x < 0 || t.x > kTrackLimit) *flags = 1;  // |momentum| <= sqrt(2)*discharge: checking one is enough
      td = t_to_float(t.x); tx = t_to_float(t.y); ty = t_to_float(t.z);
    }
    f.x = keep * f.x + lrate * td;
    f.y = keep * f.y + lrate * tx;
    f.z = keep * f.z + lrate * ty;
    reinterpret_cast<float4*>(rec + i)[0] = f;
    if (reset) reinterpret_cast<int4*>(rec + i)[1] = make_int4(0, 0, 0, t.w);  // .w: the root count of shx_veg_kernels.cuh
  }
}

__global__ void reset_tracks_kernel(CellRec* __restrict__ rec, size_t n) {  // world.h:56-61
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int4* t = reinterpret_cast<int4*>(rec + i) + 1;
    *t = make_int4(0, 0, 0, t->w);
  }
}

// ---------------------------------------------------------------------------------------------
// Boundary conversion: one 512^2 tile of the host's tiled AoS pool (32 B quad::cell records,
// x-major inside the tile) <-> the device layout.  Thread per cell; both sides coalesced
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
    reinterpret_cast<float4*>(a.m.rec + i)[0] = make_float4(lo.y, lo.z, lo.w, hi.w);
    if (a.sequential) {
      a.m.hq[i] = make_int4(__float_as_int(lo.x), 0, 0, 0);
      reinterpret_cast<float4*>(a.m.rec + i)[1] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    } else {
      if (!(fabsf(lo.x) < 31.0f) || !(fabsf(hi.x) < 4096.0f)) *a.error_flag = 1;
      const int32_t q = h_quantize(lo.x);
      a.m.hq[i] = make_int4(q, 0, q, 0);
      reinterpret_cast<int4*>(a.m.rec + i)[1] = make_int4(t_quantize(hi.x), t_quantize(hi.y), t_quantize(hi.z), 0);
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
    const float4 f = reinterpret_cast<const float4*>(a.m.rec + i)[0];
    const int4 t = reinterpret_cast<const int4*>(a.m.rec + i)[1];
    const int2 hv = *reinterpret_cast<const int2*>(a.m.hq + i);
    float4 lo, hi;
    if (a.sequential) {
      lo = make_float4(__int_as_float(hv.x), f.x, f.y, f.z);
      hi = make_float4(__int_as_float(t.x), __int_as_float(t.y), __int_as_float(t.z), f.w);
    } else {
      lo = make_float4(h_to_float(hv.x), f.x, f.y, f.z);
      hi = make_float4(t_to_float(t.x), t_to_float(t.y), t_to_float(t.z), f.w);
    }
    reinterpret_cast<float4*>(aos + c)[0] = lo;
    reinterpret_cast<float4*>(aos + c)[1] = hi;
  }
}

// The fields host code reads after World::erode -- {height, discharge, momentumx, momentumy}, the first 16 bytes of a
// quad::cell -- of every owned cell, 16 bytes per cell in POOL order (node by node, x-major inside a node): the dense
// stream of shx_download_compact.  first_node: pool index of the first owned node; thread per cell, coalesced stores.
__global__ void pack_fields16_kernel(const MapView m, int sequential, int tilesize, int mapsize, size_t first_cell, size_t ncells,
                                     float4* __restrict__ out) {
  const size_t tile = (size_t)tilesize * tilesize;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = first_cell + i;  // pool index
    const size_t node = p / tile, in = p % tile;
    const int x = (int)(node / (size_t)mapsize) * tilesize + (int)(in / (size_t)tilesize);
    const int y = (int)(node % (size_t)mapsize) * tilesize + (int)(in % (size_t)tilesize);
    const size_t c = (size_t)(x - m.xlo) * m.size + y;
    const float4 f = __ldg(reinterpret_cast<const float4*>(m.rec + c));
    const int hv = __ldg(reinterpret_cast<const int*>(m.hq + c));
    out[i] = make_float4(sequential ? __int_as_float(hv) : h_to_float(hv), f.x, f.y, f.z);
  }
}

// ---------------------------------------------------------------------------------------------
// Plant::root stamps (vegetation.h:87-118).  Applied by ONE thread in list order so that several
// stamps on one cell add up in the same fp32 order as the host's sequential `+=`.
__global__ void set_rootdensity_kernel(const MapView m, const int* xy, const float* value, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = xy[2 * i], y = xy[2 * i + 1];
    if (x < m.xlo || x >= m.xlo + m.nrows || y < 0 || y >= m.size) continue;
    m.rec[(size_t)(x - m.xlo) * m.size + y].rootdensity = value[i];
  }
}

__global__ void add_rootdensity_kernel(const MapView m, const int* xy, const float* delta, size_t n) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (size_t i = 0; i < n; i++) {
    const int x = xy[2 * i], y = xy[2 * i + 1];
    if (x < m.xlo || x >= m.xlo + m.nrows || y < 0 || y >= m.size) continue;  // getCell() == NULL -> skipped
    float* w = &m.rec[(size_t)(x - m.xlo) * m.size + y].rootdensity;
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

// pass 1: global min/max over the WHOLE map (every strip computes the same pair)
__global__ void synth_minmax_kernel(int size, uint32_t seed, unsigned* mnmx) {
  unsigned mn = 0xffffffffu, mx = 0u;
  const size_t n = (size_t)size * size;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned o = f2ord(synth_raw((int)(i / size), (int)(i % size), seed) + 0.0f);
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
    reinterpret_cast<int4*>(m.rec + i)[0] = make_int4(0, 0, 0, 0);
    reinterpret_cast<int4*>(m.rec + i)[1] = make_int4(0, 0, 0, 0);
    if (sequential) {
      m.hq[i] = make_int4(__float_as_int(h), 0, 0, 0);
    } else {
      const int32_t q = h_quantize(h);
      m.hq[i] = make_int4(q, 0, q, 0);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// N4: the reference's own terrain, map::init (cellpool.h:349-409): eight layers (frequency 1..128, weight
// 0.6^(o+1)) of FastNoiseLite's 3-D OpenSimplex2 fBm (3 octaves, lacunarity 2, gain 0.5, bounding 1/1.75, default
// rotation; vendored source/include/FastNoiseLite.h:322-338,687-716,866-886,1054-1148,541-552,487-504), z =
// SEED % 10000, then the min/max normalisation with both extremes starting at 0.  fp32, operation for operation
// (the library is built without FMA contraction), so the heights are the reference's to the bit
// (oracle: orc_init_terrain, pinned against the compiled reference's own init).
__device__ __forceinline__ float fnl_grad3(int g, int c) {  // component c of Gradients3D[g]: the 12 cube-edge vectors
  const int b = g < 60 ? g % 12 : ((0x3918 >> (4 * (g - 60))) & 15);  // last four entries: 8, 1, 9, 3
  const int grp = b >> 2;
  const float s1 = (b & 1) ? -1.0f : 1.0f, s2 = (b & 2) ? -1.0f : 1.0f;
  if (grp == 0) return c == 0 ? 0.0f : (c == 1 ? s1 : s2);
  if (grp == 1) return c == 1 ? 0.0f : (c == 0 ? s1 : s2);
  return c == 2 ? 0.0f : (c == 0 ? s1 : s2);
}
__device__ __forceinline__ float fnl_grad_coord(int seed, int xp, int yp, int zp, float xd, float yd, float zd) {
  int hash = (int)((unsigned)(seed ^ xp ^ yp ^ zp) * 0x27d4eb2du);  // FastNoiseLite.h:496-504
  hash ^= hash >> 15;                                               // :544-545
  const int g = (hash & (63 << 2)) >> 2;
  return xd * fnl_grad3(g, 0) + yd * fnl_grad3(g, 1) + zd * fnl_grad3(g, 2);
}
__device__ __forceinline__ int fnl_round(float f) { return f >= 0.0f ? (int)(f + 0.5f) : (int)(f - 0.5f); }  // :449-450

__device__ float fnl_open_simplex2_3d(int seed, float x, float y, float z) {  // :1054-1148
  constexpr unsigned PX = 501125321u, PY = 1136930381u, PZ = 1720413743u;
  int i = fnl_round(x), j = fnl_round(y), k = fnl_round(z);
  float x0 = x - (float)i, y0 = y - (float)j, z0 = z - (float)k;
  int xs = (int)(-1.0f - x0) | 1, ys = (int)(-1.0f - y0) | 1, zs = (int)(-1.0f - z0) | 1;
  float ax0 = (float)xs * -x0, ay0 = (float)ys * -y0, az0 = (float)zs * -z0;
  i = (int)((unsigned)i * PX); j = (int)((unsigned)j * PY); k = (int)((unsigned)k * PZ);
  float value = 0.0f;
  float a = (0.6f - x0 * x0) - (y0 * y0 + z0 * z0);
#pragma unroll 1
  for (int l = 0;; l++) {
    if (a > 0.0f) value += (a * a) * (a * a) * fnl_grad_coord(seed, i, j, k, x0, y0, z0);
    float b = a + 1.0f;
    int i1 = i, j1 = j, k1 = k;
    float x1 = x0, y1 = y0, z1 = z0;
    if (ax0 >= ay0 && ax0 >= az0) { x1 += (float)xs; b -= (float)(xs * 2) * x1; i1 = (int)((unsigned)i1 - (unsigned)xs * PX); }
    else if (ay0 > ax0 && ay0 >= az0) { y1 += (float)ys; b -= (float)(ys * 2) * y1; j1 = (int)((unsigned)j1 - (unsigned)ys * PY); }
    else { z1 += (float)zs; b -= (float)(zs * 2) * z1; k1 = (int)((unsigned)k1 - (unsigned)zs * PZ); }
    if (b > 0.0f) value += (b * b) * (b * b) * fnl_grad_coord(seed, i1, j1, k1, x1, y1, z1);
    if (l == 1) break;
    ax0 = 0.5f - ax0; ay0 = 0.5f - ay0; az0 = 0.5f - az0;
    x0 = (float)xs * ax0; y0 = (float)ys * ay0; z0 = (float)zs * az0;
    a += (0.75f - ax0) - (ay0 + az0);
    i = (int)((unsigned)i + ((unsigned)(xs >> 1) & PX)); j = (int)((unsigned)j + ((unsigned)(ys >> 1) & PY));
    k = (int)((unsigned)k + ((unsigned)(zs >> 1) & PZ));
    xs = -xs; ys = -ys; zs = -zs;
    seed = ~seed;
  }
  return value * 32.69428253173828125f;
}

__device__ float terrain_raw(int x, int y, int tilesize, int seed) {  // cellpool.h:354-380 for one cell
  const float px = (float)x / (float)tilesize, py = (float)y / (float)tilesize;  // :370
  const float pz = (float)(seed % 10000);
  float h = 0.0f, frequency = 1.0f, scale = 0.6f;
#pragma unroll 1
  for (int o = 0; o < 8; o++) {
    float fx = px * frequency, fy = py * frequency, fz = pz * frequency;  // FastNoiseLite.h:689-691
    const float r = (fx + fy + fz) * (float)(2.0 / 3.0);                  // :708-715
    fx = r - fx; fy = r - fy; fz = r - fz;
    int s = 1337;             // the constructor's seed (:114)
    float sum = 0.0f, amp = 1.0f / 1.75f;  // :129
#pragma unroll 1
    for (int i = 0; i < 3; i++) {  // :872-882 (weighted strength 0: the Lerp factor is exactly 1)
      const float noise = fnl_open_simplex2_3d(s++, fx, fy, fz);
      sum += noise * amp;
      amp *= 1.0f + 0.0f * ((noise + 1.0f) * 0.5f - 1.0f);
      fx *= 2.0f; fy *= 2.0f; fz *= 2.0f;
      amp *= 0.5f;
    }
    h += scale * sum;                      // cellpool.h:371
    frequency *= 2.0f;                     // :375
    scale = (float)((double)scale * 0.6);  // :376
  }
  return h;
}

// pass 1: raw heights of the stored rows (parked in hq[].x as float bits) and min/max over the WHOLE map (every
// strip computes the same pair; rows of other strips are evaluated for the extremes only)
__global__ void terrain_raw_kernel(const MapView m, int tilesize, int seed, unsigned* mnmx) {
  unsigned mn = 0xffffffffu, mx = 0u;
  const size_t n = (size_t)m.size * m.size;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i / m.size), y = (int)(i % m.size);
    const float v = terrain_raw(x, y, tilesize, seed);
    if (x >= m.xlo && x < m.xlo + m.nrows) m.hq[(size_t)(x - m.xlo) * m.size + y].x = __float_as_int(v);
    const unsigned o = f2ord(v + 0.0f);
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

// pass 2: (h - min) / (max - min), cellpool.h:408; all other fields zero
__global__ void terrain_fill_kernel(const MapView m, int sequential, const unsigned* mnmx) {
  const float mn = ord2f(mnmx[0]), mx = ord2f(mnmx[1]);
  const float range = mx - mn;
  const size_t n = (size_t)m.nrows * m.size;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float h = (__int_as_float(m.hq[i].x) - mn) / range;
    reinterpret_cast<int4*>(m.rec + i)[0] = make_int4(0, 0, 0, 0);
    reinterpret_cast<int4*>(m.rec + i)[1] = make_int4(0, 0, 0, 0);
    if (sequential) {
      m.hq[i] = make_int4(__float_as_int(h), 0, 0, 0);
    } else {
      const int32_t q = h_quantize(h);
      m.hq[i] = make_int4(q, 0, q, 0);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row-strip exchange helpers (multi-GPU).  A strip keeps `halo` rows of its neighbours' heights on
// each side.  Cascade transfers of drops on the strip's boundary rows land in those halo rows;
// `halo_ref` remembers what the halo held at the last refresh, so (current - ref) is exactly the
// integer amount this strip owes the owner.  Outside a run both planes are equal: plane 0 is used.
__global__ void strip_halo_delta_kernel(const int4* cur, const int32_t* ref, int32_t* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = cur[i].x - ref[i];
}
__global__ void strip_add_rows_kernel(int4* h, const int32_t* delta, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int32_t v = delta[i];
    if (v) {
      int4 c = h[i];  // {h0, claim0, h1, claim1}: outside a launch the claim words carry nothing that matters
      c.x += v; c.z += v;
      h[i] = c;
    }
  }
}
__global__ void strip_get_rows_kernel(const int4* h, int32_t* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = h[i].x;
}
__global__ void strip_set_rows_kernel(int4* h, int32_t* ref, const int32_t* src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int32_t v = src[i];
    int4 c = h[i];
    c.x = v; c.z = v;
    h[i] = c;
    ref[i] = v;
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


// One message per neighbour and call (strips that exchange once per call), int32 words:
//   [0] number of drop records   [1..7] unused   [8, 8 + 8*cap) drop records
//   then rows*size halo deltas (what this strip moved into its copy of the neighbour's edge rows)
//   then rows*size edge rows (what this strip's own edge rows hold now)
constexpr int kMsgHeader = 8;
__global__ void strip_msg_rows_kernel(const int4* halo, const int32_t* ref, const int4* edge, int32_t* out_delta, int32_t* out_edge,
                                      size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    out_delta[i] = halo[i].x - ref[i];
    out_edge[i] = edge[i].x;
  }
}
__global__ void strip_msg_migrants_kernel(const shx_drop* drops, unsigned n, int32_t* msg_lo, int32_t* msg_hi, unsigned cap) {
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    shx_drop d = drops[k];
    if (d.flags & (SHX_DROP_MIGRATE_LO | SHX_DROP_MIGRATE_HI)) {
      int32_t* msg = (d.flags & SHX_DROP_MIGRATE_LO) ? msg_lo : msg_hi;
      const unsigned slot = (unsigned)atomicAdd(msg, 1);  // a count above cap tells the receiver that records were dropped
      d.flags = (d.flags & ~(SHX_DROP_MIGRATE_LO | SHX_DROP_MIGRATE_HI)) | SHX_DROP_ALIVE;
      if (slot < cap) reinterpret_cast<shx_drop*>(msg + kMsgHeader)[slot] = d;
    }
  }
}
// The owner's edge rows take the neighbour's deltas; this strip's copy of the neighbour's edge rows
// becomes what the neighbour holds once IT has taken this strip's deltas: its rows as sent + ours.
__global__ void strip_msg_apply_kernel(int4* halo, int32_t* ref, int4* edge, const int32_t* in_delta, const int32_t* in_edge, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int4 hc = halo[i];
    const int32_t v = in_edge[i] + (hc.x - ref[i]);
    hc.x = v; hc.z = v;
    halo[i] = hc;
    ref[i] = v;
    const int32_t dv = in_delta[i];
    if (dv) {
      int4 c = edge[i];
      c.x += dv; c.z += dv;
      edge[i] = c;
    }
  }
}

// claim words of every stored cell back to zero (when the launch epoch of the claim keys wraps, every 15 launches).
// The same pass watches the fixed-point range: adds into the Q5.26 planes wrap silently, so a world that is running
// away numerically (|h| beyond 30 of the representable +-32) raises the height flag before it can wrap.
__global__ void clear_claims_kernel(int4* hq, size_t n, int* height_flag) {
  constexpr int kLimit = 30 << kHeightFracBits;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    int4 c = hq[i];
    if (c.x > kLimit || c.x < -kLimit) *height_flag = 1;
    c.y = 0; c.w = 0;
    hq[i] = c;
  }
}

// the two height words of every stored cell, densely (shx_download_raw)
__global__ void copy_heights_kernel(const int4* hq, int2* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int4 c = hq[i];
    out[i] = make_int2(c.x, c.z);
  }
}


// ---------------------------------------------------------------------------------------------
// Bandwidth probe (shx_measure_read_bandwidth): every thread streams 16-byte .cg loads over a buffer `passes`
// times.  With a buffer that fits L2 this measures the L2 read bandwidth the latency-regime configurations are
// set against (MEASURED_PEAKS.json only lists HBM); with a buffer far beyond L2 it reproduces the HBM figure.
__global__ void read_bandwidth_kernel(const int4* __restrict__ buf, size_t n, int passes, int* sink) {
  int acc = 0;
  for (int p = 0; p < passes; p++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      const int4 v = __ldcg(buf + i);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x7fffffff) *sink = acc;  // never true for a zeroed buffer: keeps the loads alive
}

}  // namespace shx
