// Per-frame consumers of the eroded map that the reference runs on the host right after
// World::erode (SimpleHydrology.cpp:322-324, 341-354), as streaming kernels over the device map:
//   vertex_fill_kernel   quad::updatenode  (source/cellpool.h:286-305)  -> 48-byte Vertex records
//   view_maps_kernel     dischargeMap / momentumMap builders (SimpleHydrology.cpp:341-354)
// Both are HBM-bound (one pass over the heights / fields, one coalesced write of the result).
#pragma once
#include "shx_kernels.cuh"

namespace shx {

struct ViewArgs {
  MapView m;
  int tilesize, mapsize;  // cells per tile side, tiles per map side
  int sequential;         // heights are fp32 bit patterns instead of Q5.26
  float mapscale;
};

__device__ __forceinline__ float view_height(const ViewArgs& a, int x, int y) {
  const int v = __ldg(reinterpret_cast<const int*>(a.m.hq + (size_t)(x - a.m.xlo) * a.m.size + y));  // plane 0
  return a.sequential ? __int_as_float(v) : h_to_float(v);
}

// quad::updatenode for every node of the owned rows.  out: 12 floats per cell {position, normal,
// tangent, bitangent} (vertexpool.h:6-26) in pool order: node (x/ts)*mapsize + (y/ts), then
// x*ts + y inside the node -- the order of the reference's vertex pool sections.  Everything is
// NODE-local like the reference: node::height() of a cell outside the node is 0 and node::normal()
// only uses the planes whose corner lies inside the node (cellpool.h:227-250).
// One thread per cell; the block's records go through shared memory so that the 48-byte records
// leave as full 16-byte-per-lane coalesced stores.
__global__ void __launch_bounds__(256) vertex_fill_kernel(const ViewArgs a, float* __restrict__ out, const size_t ncells,
                                                          const size_t first_cell) {
  __shared__ __align__(16) float s_v[256 * 12];
  const int ts = a.tilesize, ta = ts * ts;
  for (size_t base = (size_t)blockIdx.x * 256; base < ncells; base += (size_t)gridDim.x * 256) {
    const size_t i = base + threadIdx.x;
    if (i < ncells) {
      const size_t g = first_cell + i;  // index in the pool of the whole map
      const int node = (int)(g / ta), r = (int)(g % ta);
      const int lx = r / ts, ly = r % ts;
      const int x = (node / a.mapsize) * ts + lx, y = (node % a.mapsize) * ts + ly;
      const bool xm = lx > 0, xp = lx < ts - 1, ym = ly > 0, yp = ly < ts - 1;
      const float hc = view_height(a, x, y);
      const float hxp = xp ? view_height(a, x + 1, y) : 0.0f;
      const float hxm = xm ? view_height(a, x - 1, y) : 0.0f;
      const float hyp = yp ? view_height(a, x, y + 1) : 0.0f;
      const float hym = ym ? view_height(a, x, y - 1) : 0.0f;
      // cellpool.h:181-204, the four cross products written out (see shx_step.cuh move_math)
      const float Bp = a.mapscale * (hxp - hc), Bm = a.mapscale * (hxm - hc);
      const float Ap = a.mapscale * (hyp - hc), Am = a.mapscale * (hym - hc);
      float nx = 0.0f, ny = 0.0f, nz = 0.0f;
      if (xp && yp) { nx += -Bp; ny += 1.0f; nz += -Ap; }
      if (xm && ym) { nx += Bm; ny += 1.0f; nz += Am; }
      if (xp && ym) { nx += -Bp; ny += 1.0f; nz += Am; }
      if (xm && yp) { nx += Bm; ny += 1.0f; nz += -Ap; }
      const float l2 = nx * nx + ny * ny + nz * nz;
      if (l2 > 0.0f) {
        const float inv = 1.0f / sqrtf(l2);
        nx *= inv; ny *= inv; nz *= inv;
      }
      const float py = a.mapscale * hc;  // cellpool.h:294-296: P, T, B; T - P and B - P component-wise
      // three 16-byte shared stores per thread at a 48-byte stride: conflict-free per quarter-warp
      float4* v = reinterpret_cast<float4*>(s_v) + threadIdx.x * 3;
      v[0] = make_float4((float)x, py, (float)y, nx);
      v[1] = make_float4(ny, nz, (float)(x + 1) - (float)x, a.mapscale * hxp - py);
      v[2] = make_float4(0.0f, 0.0f, a.mapscale * hyp - py, (float)(y + 1) - (float)y);
    }
    __syncthreads();
    const size_t nblk = (ncells - base < 256 ? ncells - base : 256) * 3;  // float4s of this block
    float4* dst = reinterpret_cast<float4*>(out + base * 12);
    for (size_t k = threadIdx.x; k < nblk; k += 256) dst[k] = reinterpret_cast<const float4*>(s_v)[k];
    __syncthreads();
  }
}

// dischargeMap alpha = erf(0.4f*discharge) (cellpool.h:242-244 via :439) and momentumMap
// {0.5*(1+erf(mx)), 0.5*(1+erf(my))} (SimpleHydrology.cpp:349-353), per cell of the owned rows in
// map order (x*size + y).  out: float4 {discharge alpha, momentum r, momentum g, height}.
__global__ void view_maps_kernel(const ViewArgs a, float4* __restrict__ out, const size_t ncells) {
  const size_t off = (size_t)(a.m.row0 - a.m.xlo) * a.m.size;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += (size_t)gridDim.x * blockDim.x) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(a.m.rec + off + i));  // discharge momentumx momentumy rootdensity
    const int hv = __ldg(reinterpret_cast<const int*>(a.m.hq + off + i));
    float4 o;
    o.x = shx_erff(0.4f * f.x);
    o.y = 0.5f * (1.0f + shx_erff(f.y));
    o.z = 0.5f * (1.0f + shx_erff(f.z));
    o.w = a.sequential ? __int_as_float(hv) : h_to_float(hv);
    out[i] = o;
  }
}

// The two RGBA8 textures the reference uploads every frame (SimpleHydrology.cpp:341-354):
//   dischargeMap  vec4(waterColor, erf(0.4*discharge))                       (cellpool.h:242-244 via :439)
//   momentumMap   vec4(0.5*(1+erf(mx)), 0.5*(1+erf(my)), 0.5, 1.0)
// one texel per cell in map order (x*size + y), each channel (unsigned char)(255*c) -- the byte packing restates
// TinyEngine 1.7's image::make, which is NOT part of /root/reference (parity of the packing is unpinned; the
// float values are pinned through shx_view_maps).  16 of the record's 32 bytes read, 8 bytes written per cell.
__device__ __forceinline__ unsigned pack_rgba8(float r, float g, float b, float a) {
  auto q = [](float c) { return (unsigned)(unsigned char)(int)(255.0f * c); };  // C++ float -> unsigned char via truncation
  return q(r) | (q(g) << 8) | (q(b) << 16) | (q(a) << 24);
}
__global__ void view_textures_kernel(const ViewArgs a, const float3 water, unsigned* __restrict__ discharge_rgba,
                                     unsigned* __restrict__ momentum_rgba, const size_t ncells) {
  const size_t off = (size_t)(a.m.row0 - a.m.xlo) * a.m.size;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += (size_t)gridDim.x * blockDim.x) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(a.m.rec + off + i));  // discharge momentumx momentumy rootdensity
    discharge_rgba[i] = pack_rgba8(water.x, water.y, water.z, shx_erff(0.4f * f.x));
    momentum_rgba[i] = pack_rgba8(0.5f * (1.0f + shx_erff(f.y)), 0.5f * (1.0f + shx_erff(f.z)), 0.5f, 1.0f);
  }
}

// Sparse read-back for host code that only looks at a few cells per frame (Vegetation::grow reads
// discharge / height / normal / rootdensity at plant positions, vegetation.h:67-85,160-180): the
// 32-byte records of the queried cells and, optionally, World::map.normal there (cellpool.h:181-204
// with the MAP-level oob of cellpool.h:413-419, i.e. across tile borders, unlike updatenode).
// A query outside the map (map.get() == NULL) or outside this strip's stored rows returns zeros.
__global__ void gather_cells_kernel(const ViewArgs a, const int* __restrict__ xy, const size_t n, shx_cell* __restrict__ out,
                                    float* __restrict__ normals) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = xy[2 * i], y = xy[2 * i + 1];
    const int size = a.m.size;
    float4 lo = make_float4(0.0f, 0.0f, 0.0f, 0.0f), hi = lo;
    float nx = 0.0f, ny = 0.0f, nz = 0.0f;
    // the normal needs rows x-1 .. x+1: all stored, or outside the map
    const bool have = x >= 0 && x < size && y >= 0 && y < size && x >= a.m.xlo && x < a.m.xlo + a.m.nrows &&
                      (x - 1 >= a.m.xlo || x == 0) && (x + 1 < a.m.xlo + a.m.nrows || x == size - 1);
    if (have) {
      const size_t c = (size_t)(x - a.m.xlo) * size + y;
      const float4 f = __ldg(reinterpret_cast<const float4*>(a.m.rec + c));
      const int4 t = __ldg(reinterpret_cast<const int4*>(a.m.rec + c) + 1);
      const float hc = view_height(a, x, y);
      if (a.sequential) {
        lo = make_float4(hc, f.x, f.y, f.z);
        hi = make_float4(__int_as_float(t.x), __int_as_float(t.y), __int_as_float(t.z), f.w);
      } else {
        lo = make_float4(hc, f.x, f.y, f.z);
        hi = make_float4(t_to_float(t.x), t_to_float(t.y), t_to_float(t.z), f.w);
      }
      const bool xm = x > 0, xp = x < size - 1, ym = y > 0, yp = y < size - 1;
      const float hxp = xp ? view_height(a, x + 1, y) : 0.0f, hxm = xm ? view_height(a, x - 1, y) : 0.0f;
      const float hyp = yp ? view_height(a, x, y + 1) : 0.0f, hym = ym ? view_height(a, x, y - 1) : 0.0f;
      const float Bp = a.mapscale * (hxp - hc), Bm = a.mapscale * (hxm - hc);
      const float Ap = a.mapscale * (hyp - hc), Am = a.mapscale * (hym - hc);
      if (xp && yp) { nx += -Bp; ny += 1.0f; nz += -Ap; }
      if (xm && ym) { nx += Bm; ny += 1.0f; nz += Am; }
      if (xp && ym) { nx += -Bp; ny += 1.0f; nz += Am; }
      if (xm && yp) { nx += Bm; ny += 1.0f; nz += -Ap; }
      const float l2 = nx * nx + ny * ny + nz * nz;
      if (l2 > 0.0f) {
        const float inv = 1.0f / sqrtf(l2);
        nx *= inv; ny *= inv; nz *= inv;
      }
    }
    reinterpret_cast<float4*>(out + i)[0] = lo;
    reinterpret_cast<float4*>(out + i)[1] = hi;
    if (normals) { normals[3 * i] = nx; normals[3 * i + 1] = ny; normals[3 * i + 2] = nz; }
  }
}

}  // namespace shx

#include "shx_veg_kernels.cuh"
