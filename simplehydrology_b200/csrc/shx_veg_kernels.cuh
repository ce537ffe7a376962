// N3: Vegetation::grow() on the device -- reference source/vegetation.h:122-188 with Plant::grow
// (:67-69), Plant::die (:71-78), Plant::spawn (:80-89) and Plant::root (:91-120).  Included by shx_view_kernels.cuh.
//
// The reference walks its plant vector ONCE per frame, strictly in order, drawing from the global rand() stream, and
// every root() stamp is visible to the plants after it.  The device form keeps the predicates and the arithmetic and
// changes the schedule, like the erosion path does:
//   * every plant of the frame decides from the maps as they were at the START of the frame (heights, discharge and
//     rootdensity are frozen: the stamps of this frame's deaths and births become visible together, afterwards);
//   * rand() is replaced by a counter hash keyed (seed, frame, cell, size bits): a plant's draws do not depend on its
//     place in the list, so the result is independent of the order in which plants are processed;
//   * the frame's one random seeding attempt (vegetation.h:126-137) is appended after the survivors instead of being
//     walked in the same frame;
//   * rootdensity is kept as an INTEGER number of fifths per cell (the stamp weights 1.0 / 0.6 / 0.4 are 5 / 3 / 2
//     fifths; the count lives in the spare word of the cell record) and the fp32 value the erosion reads is
//     count / 5, correctly rounded: stamps are integer atomics, so the map is run-to-run identical and carries none
//     of the +/- rounding drift the reference's repeated fp32 adds leave behind.
// The plant list keeps a deterministic order: survivors in their old order, then the seeded plant, then the children
// in the order of their parents.  The tests hold a bit-exact CPU restatement of this schedule (tests/test_gpu_vegetation.py
// compares the two) and set the population against the reference's own Vegetation::grow (tests/test_oracle_vegetation.py).
#pragma once

namespace shx {

struct PlantParams {  // Plant:: statics, vegetation.h:40-44
  float maxSize, growRate, maxSteep, maxDischarge, maxTreeHeight;
};

struct VegArgs {
  ViewArgs v;
  PlantParams pp;
  uint64_t key;          // mix64 of (seed, frame)
  const int2* pos;       // plants of the frame
  const float* size;
  unsigned n;
  int2* pos_out;         // compacted list of the next frame
  float* size_out;
  unsigned cap;          // capacity of the lists
  unsigned* flags;       // per decision slot (n plants + the seeding attempt): bit 0 survives, bit 1 child, bit 2 died, bit 3 child refused
  int2* child;           // per decision slot: where the child goes
  float* grown;          // per plant: size after Plant::grow
  uint2* block_counts;   // per block of decide: {survivors, children}; after the scan: exclusive prefix sums
  unsigned* totals;      // [0] survivors, [1] children kept, [2] deaths, [3] children refused (capacity), [4] new count
};

constexpr int kVegBlock = 256;

// World::map.normal(p).y (cellpool.h:181-204 with the map-level oob of :413-419), same operations as gather_cells_kernel
__device__ __forceinline__ float veg_normal_y(const ViewArgs& a, int x, int y) {
  const int size = a.m.size;
  const float hc = view_height(a, x, y);
  const bool xm = x > 0, xp = x < size - 1, ym = y > 0, yp = y < size - 1;
  const float hxp = xp ? view_height(a, x + 1, y) : 0.0f, hxm = xm ? view_height(a, x - 1, y) : 0.0f;
  const float hyp = yp ? view_height(a, x, y + 1) : 0.0f, hym = ym ? view_height(a, x, y - 1) : 0.0f;
  const float Bp = a.mapscale * (hxp - hc), Bm = a.mapscale * (hxm - hc);
  const float Ap = a.mapscale * (hyp - hc), Am = a.mapscale * (hym - hc);
  float nx = 0.0f, ny = 0.0f, nz = 0.0f;
  if (xp && yp) { nx += -Bp; ny += 1.0f; nz += -Ap; }
  if (xm && ym) { nx += Bm; ny += 1.0f; nz += Am; }
  if (xp && ym) { nx += -Bp; ny += 1.0f; nz += Am; }
  if (xm && yp) { nx += Bm; ny += 1.0f; nz += -Ap; }
  const float l2 = nx * nx + ny * ny + nz * nz;
  if (l2 > 0.0f) ny *= 1.0f / sqrtf(l2);
  return ny;
}

__device__ __forceinline__ float veg_discharge(const ViewArgs& a, int x, int y) {  // map.discharge(p), cellpool.h:242-244
  return shx_erff(0.4f * __ldg(&a.m.rec[(size_t)(x - a.m.xlo) * a.m.size + y].discharge));
}

// pass 1: one thread per decision slot; slot n is the frame's random seeding attempt
__global__ void __launch_bounds__(kVegBlock) veg_decide_kernel(const VegArgs a) {
  __shared__ unsigned s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const unsigned i = blockIdx.x * kVegBlock + threadIdx.x;
  const int size = a.v.m.size;
  unsigned fl = 0u;
  int2 ch = make_int2(0, 0);
  if (i < a.n) {
    const int2 p = a.pos[i];
    const float s0 = a.size[i];
    a.grown[i] = s0 + a.pp.growRate * (a.pp.maxSize - s0);  // Plant::grow, vegetation.h:67-69
    const uint64_t r = mix64(a.key + (((uint64_t)(uint32_t)p.x << 32) | (uint32_t)p.y) + (uint64_t)__float_as_uint(s0) * 0x9E3779B97F4A7C15ull);
    // Plant::die, vegetation.h:71-78
    const bool die = veg_discharge(a.v, p.x, p.y) >= a.pp.maxDischarge || view_height(a.v, p.x, p.y) >= a.pp.maxTreeHeight ||
                     (uint32_t)r % 1000u == 0u;
    if (die) {
      fl = 4u;
    } else {
      fl = 1u;
      if ((uint32_t)(r >> 32) % 20u == 0u) {  // vegetation.h:157-183
        const uint64_t q = mix64(r);
        const int nx = p.x + (int)((uint32_t)q % 9u) - 4, ny = p.y + (int)((uint32_t)(q >> 32) % 9u) - 4;
        if (nx >= 0 && ny >= 0 && nx < size && ny < size && veg_discharge(a.v, nx, ny) < a.pp.maxDischarge) {
          const uint32_t r5 = (uint32_t)mix64(q) % 1000u;
          const float root = __ldg(&a.v.m.rec[(size_t)(nx - a.v.m.xlo) * size + ny].rootdensity);
          // (float)(rand()%1000)/1000.0 <= rootdensity : a double comparison (vegetation.h:172)
          if (!((double)(float)r5 / 1000.0 <= (double)root) && veg_normal_y(a.v, nx, ny) > a.pp.maxSteep) {
            fl |= 2u;
            ch = make_int2(nx, ny);
          }
        }
      }
    }
  } else if (i == a.n) {  // vegetation.h:126-137: one attempt anywhere on the map, Plant::spawn (:80-89)
    const uint64_t r = mix64(a.key ^ 0x5EED5EED5EED5EEDull);
    const int x = (int)((uint32_t)r % (uint32_t)size), y = (int)((uint32_t)(r >> 32) % (uint32_t)size);
    if (veg_discharge(a.v, x, y) < a.pp.maxDischarge && !(veg_normal_y(a.v, x, y) < a.pp.maxSteep) &&
        view_height(a.v, x, y) < a.pp.maxTreeHeight) {
      fl = 2u;
      ch = make_int2(x, y);
    }
  }
  if (i <= a.n) {
    a.flags[i] = fl;
    a.child[i] = ch;
    if (fl & 1u) atomicAdd(&s_cnt[0], 1u);
    if (fl & 2u) atomicAdd(&s_cnt[1], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) a.block_counts[blockIdx.x] = make_uint2(s_cnt[0], s_cnt[1]);
}

// pass 2: exclusive scan of the per-block counts (one block; the list of a 8192^2 world has a few thousand blocks)
__global__ void __launch_bounds__(1024) veg_scan_kernel(const VegArgs a, unsigned nblocks) {
  __shared__ unsigned s_a[1024], s_b[1024];
  __shared__ unsigned s_base[2];
  if (threadIdx.x == 0) s_base[0] = s_base[1] = 0u;
  __syncthreads();
  for (unsigned b0 = 0; b0 < nblocks; b0 += 1024u) {
    const unsigned b = b0 + threadIdx.x;
    const uint2 c = b < nblocks ? a.block_counts[b] : make_uint2(0u, 0u);
    s_a[threadIdx.x] = c.x;
    s_b[threadIdx.x] = c.y;
    __syncthreads();
    for (unsigned o = 1; o < 1024u; o <<= 1) {  // Hillis-Steele inclusive scan
      const unsigned va = threadIdx.x >= o ? s_a[threadIdx.x - o] : 0u, vb = threadIdx.x >= o ? s_b[threadIdx.x - o] : 0u;
      __syncthreads();
      s_a[threadIdx.x] += va;
      s_b[threadIdx.x] += vb;
      __syncthreads();
    }
    if (b < nblocks) a.block_counts[b] = make_uint2(s_base[0] + s_a[threadIdx.x] - c.x, s_base[1] + s_b[threadIdx.x] - c.y);
    __syncthreads();
    if (threadIdx.x == 1023u) {
      s_base[0] += s_a[1023];
      s_base[1] += s_b[1023];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const unsigned surv = s_base[0], kids = s_base[1];
    const unsigned room = a.cap > surv ? a.cap - surv : 0u;
    a.totals[0] = surv;
    a.totals[1] = kids < room ? kids : room;
    a.totals[2] = a.n - surv;
    a.totals[3] = kids - a.totals[1];
    a.totals[4] = surv + a.totals[1];
  }
}

// Plant::root(f), vegetation.h:91-120, in fifths: +-5 at the cell, +-3 along the axes, +-2 on the diagonals; cells
// outside the map do not exist (getCell == NULL)
__device__ __forceinline__ void veg_stamp(const ViewArgs& a, int x, int y, int sign) {
  const int size = a.m.size;
#pragma unroll
  for (int dx = -1; dx <= 1; dx++)
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
      const int cx = x + dx, cy = y + dy;
      if (cx < 0 || cy < 0 || cx >= size || cy >= size) continue;
      const int w = (dx == 0 && dy == 0) ? 5 : ((dx == 0 || dy == 0) ? 3 : 2);
      atomicAdd(&a.m.rec[(size_t)(cx - a.m.xlo) * size + cy].pad, sign * w);
    }
}

// pass 3: compaction in list order (block-level scan of the flags on top of the scanned block counts) and the stamps
__global__ void __launch_bounds__(kVegBlock) veg_apply_kernel(const VegArgs a) {
  __shared__ unsigned s_a[kVegBlock], s_b[kVegBlock];
  const unsigned i = blockIdx.x * kVegBlock + threadIdx.x;
  const unsigned fl = i <= a.n ? a.flags[i] : 0u;
  s_a[threadIdx.x] = fl & 1u;
  s_b[threadIdx.x] = (fl >> 1) & 1u;
  __syncthreads();
  for (unsigned o = 1; o < kVegBlock; o <<= 1) {
    const unsigned va = threadIdx.x >= o ? s_a[threadIdx.x - o] : 0u, vb = threadIdx.x >= o ? s_b[threadIdx.x - o] : 0u;
    __syncthreads();
    s_a[threadIdx.x] += va;
    s_b[threadIdx.x] += vb;
    __syncthreads();
  }
  const uint2 base = a.block_counts[blockIdx.x];
  const unsigned surv = a.totals[0], kept = a.totals[1];
  if (fl & 1u) {
    const unsigned at = base.x + s_a[threadIdx.x] - 1u;
    a.pos_out[at] = a.pos[i];
    a.size_out[at] = a.grown[i];
  }
  if (fl & 4u) {
    const int2 p = a.pos[i];
    veg_stamp(a.v, p.x, p.y, -1);  // vegetation.h:151: root(-1.0) before the erase
  }
  if (fl & 2u) {
    // the seeding attempt (slot n) comes first among the newcomers (it is appended before the walk, :131-134)
    const bool seeded = i == a.n;
    const unsigned seed_first = (a.flags[a.n] >> 1) & 1u;
    const unsigned rank = seeded ? 0u : seed_first + base.y + s_b[threadIdx.x] - 1u;
    if (rank < kept) {
      const int2 c = a.child[i];
      a.pos_out[surv + rank] = c;
      a.size_out[surv + rank] = 0.0f;  // vegetation.h:16
      veg_stamp(a.v, c.x, c.y, +1);    // :133, :181
    } else {
      a.flags[i] = fl | 8u;  // refused (list full): nothing was stamped, nothing to refresh
    }
  }
}

// shx_veg_upload(stamp_roots): root(+1) of every listed plant (flags bit 1, position in child[])
__global__ void __launch_bounds__(kVegBlock) veg_stamp_list_kernel(const VegArgs a) {
  const unsigned i = blockIdx.x * kVegBlock + threadIdx.x;
  if (i <= a.n && (a.flags[i] & 2u)) veg_stamp(a.v, a.child[i].x, a.child[i].y, +1);
}

// pass 4: the fp32 rootdensity of every cell a stamp of this frame touched = count / 5 (correctly rounded)
__global__ void __launch_bounds__(kVegBlock) veg_refresh_kernel(const VegArgs a) {
  const unsigned i = blockIdx.x * kVegBlock + threadIdx.x;
  if (i > a.n) return;
  const unsigned fl = a.flags[i];
  const int size = a.v.m.size;
  for (int which = 0; which < 2; which++) {
    if (!(fl & (which ? 2u : 4u)) || (which && (fl & 8u))) continue;
    const int2 p = which ? a.child[i] : a.pos[i];
    for (int dx = -1; dx <= 1; dx++)
      for (int dy = -1; dy <= 1; dy++) {
        const int cx = p.x + dx, cy = p.y + dy;
        if (cx < 0 || cy < 0 || cx >= size || cy >= size) continue;
        CellRec* r = a.v.m.rec + (size_t)(cx - a.v.m.xlo) * size + cy;
        r->rootdensity = (float)r->pad / 5.0f;  // several stamps of one cell write the same value
      }
  }
}

// every cell: count = round(5 * rootdensity) (after an upload or a host push of fp32 values)
__global__ void veg_count_from_density_kernel(CellRec* __restrict__ rec, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    rec[i].pad = __float2int_rn(rec[i].rootdensity * 5.0f);
}

// the tree particle system's model matrices (SimpleHydrology.cpp:329-335): translate(pos.x, size + mapscale*height,
// pos.y) * scale(size), 16 floats per plant, column-major like glm::mat4
__global__ void veg_models_kernel(const ViewArgs v, const int2* __restrict__ pos, const float* __restrict__ size, unsigned n,
                                  float4* __restrict__ out) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int2 p = pos[i];
    const float s = size[i];
    const float h = view_height(v, p.x, p.y);
    out[4 * (size_t)i + 0] = make_float4(s, 0.0f, 0.0f, 0.0f);
    out[4 * (size_t)i + 1] = make_float4(0.0f, s, 0.0f, 0.0f);
    out[4 * (size_t)i + 2] = make_float4(0.0f, 0.0f, s, 0.0f);
    out[4 * (size_t)i + 3] = make_float4((float)p.x, s + v.mapscale * h, (float)p.y, 1.0f);
  }
}

// {x, y, size} per plant for the host (Vegetation::plants: pos, size)
__global__ void veg_export_kernel(const int2* __restrict__ pos, const float* __restrict__ size, unsigned n, float* __restrict__ out) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    out[3 * (size_t)i] = (float)pos[i].x;
    out[3 * (size_t)i + 1] = (float)pos[i].y;
    out[3 * (size_t)i + 2] = size[i];
  }
}

}  // namespace shx
