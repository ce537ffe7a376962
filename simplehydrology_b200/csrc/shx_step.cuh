// Per-step arithmetic of the erosion path, shared by both device modes.
//
// Reference semantics (file:line into the reference tree):
//   Drop::descend    source/water.h:58-156
//   World::cascade   source/world.h:90-168
//   quad::_normal    source/cellpool.h:181-204
//   map::height/oob  source/cellpool.h:413-437, node::discharge :242-244
//
// One Drop::descend call is split where it touches memory:
//   move_math      water.h:70-117   normal -> termination checks -> forces -> fixed-length move -> track amounts
//   exchange_math  water.h:120-136  sediment exchange with the OLD cell against the NEW cell's height, evaporation
// followed by water.h:139-154 (out-of-bounds stop, cascade at the new cell, age++).
//
// Both modes run them back to back like the reference.  (Deferring exchange_math of step k to phase
// k+1 -- where the new cell is the centre of the block loaded anyway -- would remove the one dependent
// gather of a step; it was tried and rejected: the erosion of a cell then becomes visible to the
// other drops one phase later, and dense flows (thousands of drops per tile) go unstable.)
#pragma once
#include "../../include/shx.h"
#include "shx_math.cuh"

namespace shx {

struct StepParams {
  float maxAge, minVol, evapRate, depositionRate, entrainment, gravity, momentumTransfer;
  float maxdiff, settling, lod, mapscale, lrate;
  float lim_axis, lim_diag;  // world.h:145: d*maxdiff*lodsize for d = 1 and d = sqrt(2)
  double keep;               // 1.0 - (double)evapRate, water.h:135-136
  double inv_keep;           // 1.0 / keep: the batched kernels multiply where the reference divides
};

struct DropRegs {
  float px, py, sx, sy, vol, sed;
  int age, flags;
};

// (double)h > 0.1  <=>  h >= 0.1f : 0.1f is the smallest float above the double 0.1
// (world.h:71,144 compare a float height against a double literal)
__device__ __forceinline__ bool above_tenth(float h) { return h >= 0.1f; }

struct MoveResult {
  bool moved;             // false: the call terminated by age / volume (water.h:74-82), dheight = +sediment
  bool oob;               // new position outside the map (water.h:121,139)
  float dheight;          // termination only: fp32 amount to ADD to the centre cell
  float t_d, t_mx, t_my;  // track deposits at the old cell (water.h:115-117)
  float effD;             // depositionRate*(1-rootdensity), clamped (water.h:86-87)
};

// water.h:70-117 given the five heights of the normal stencil (after the owed cascade).
// inb: bit k of the 3x3 block (k = (dx+1)*3 + (dy+1)) set if that cell exists.
__device__ __forceinline__ MoveResult move_math(const float hc, const float hxm, const float hxp, const float hym,
                                                const float hyp, const unsigned inb, DropRegs& d, const float4 fld,
                                                const StepParams& P, const int size) {
  MoveResult out;
  out.moved = false;
  out.oob = false;
  out.dheight = 0.0f;
  out.t_d = out.t_mx = out.t_my = 0.0f;
  out.effD = 0.0f;

  // cellpool.h:181-204.  height() of a missing cell is 0 (cellpool.h:433-437; the caller passes 0).
  // Each plane's cross product (cellpool.h:188,191,195,198) written out is (-+80*dh_x, 1, -+80*dh_y);
  // the products with the literal zeros of the generic formula only decide the sign of a zero.
  const float Bp = P.mapscale * (hxp - hc), Bm = P.mapscale * (hxm - hc);
  const float Ap = P.mapscale * (hyp - hc), Am = P.mapscale * (hym - hc);
  float nx = 0.0f, ny = 0.0f, nz = 0.0f;
  if (inb & (1u << 8)) { nx += -Bp; ny += 1.0f; nz += -Ap; }
  if (inb & (1u << 0)) { nx += Bm; ny += 1.0f; nz += Am; }
  if (inb & (1u << 6)) { nx += -Bp; ny += 1.0f; nz += Am; }
  if (inb & (1u << 2)) { nx += Bm; ny += 1.0f; nz += -Ap; }
  {
    const float l2 = nx * nx + ny * ny + nz * nz;
    if (l2 > 0.0f) {  // length(n) > 0  <=>  dot(n,n) > 0;  glm normalize = v * (1/sqrt(dot(v,v)))
      const float inv = 1.0f / sqrtf(l2);
      nx *= inv; ny *= inv; nz *= inv;
    }
  }

  if ((float)d.age > P.maxAge || d.vol < P.minVol) {  // water.h:74-82
    out.dheight = d.sed;
    d.flags = ((float)d.age > P.maxAge) ? SHX_DROP_DONE_AGE : SHX_DROP_DONE_VOL;
    return out;
  }

  float effD = P.depositionRate * (1.0f - fld.w);  // water.h:86-87
  if (effD < 0.0f) effD = 0.0f;
  out.effD = effD;
  {
    const float g = P.lod * P.gravity;  // water.h:95
    d.sx += (g * nx) / d.vol;
    d.sy += (g * nz) / d.vol;
  }
  const float fx = fld.y, fy = fld.z;
  const float f2 = fx * fx + fy * fy, s2 = d.sx * d.sx + d.sy * d.sy;
  if (f2 > 0.0f && s2 > 0.0f) {  // water.h:97-99
    const float fi = 1.0f / sqrtf(f2);
    const float si = 1.0f / sqrtf(s2);
    const float dp = (fx * fi) * (d.sx * si) + (fy * fi) * (d.sy * si);
    const float k = P.lod * P.momentumTransfer * dp / (d.vol + fld.x);
    d.sx += k * fx;
    d.sy += k * fy;
  }
  {
    const float l2 = d.sx * d.sx + d.sy * d.sy;
    if (l2 > 0.0f) {  // water.h:108-109
      const float si = 1.0f / sqrtf(l2);
      const float m = P.lod * sqrtf(2.0f);
      d.sx = m * (d.sx * si);
      d.sy = m * (d.sy * si);
    }
  }
  d.px += d.sx;  // water.h:111
  d.py += d.sy;

  out.moved = true;  // water.h:115-117: old cell, new speed
  out.t_d = d.vol;
  out.t_mx = d.vol * d.sx;
  out.t_my = d.vol * d.sy;

  // truncation as ivec2(vec2); !(x > -1) also catches NaN, which the reference's cvttss2si maps to
  // INT_MIN (out of bounds)
  const int nix = (int)d.px, niy = (int)d.py;
  out.oob = !(d.px > -1.0f) || !(d.py > -1.0f) || nix >= size || niy >= size;
  return out;
}

// water.h:127-136: hc = old cell's height, h2 = new cell's height (or hc - 0.002 when out of bounds,
// water.h:121-122, done by the caller), cap = 1 + entrainment*erf(0.4*discharge) of the old cell
// (water.h:127, cellpool.h:242-244).  Returns the fp32 amount to ADD to the old cell (-effD*cdiff).
// kRecip: water.h:135 as a multiplication by the double 1/(1-evapRate) (batched mode; the fp64 division is a
// ~50-instruction routine on the critical chain of every step; oracle: orc_ls_world.recip_evap).  The sequential
// mode keeps the reference's division.
template <bool kRecip = false>
__device__ __forceinline__ float exchange_math(const float hc, const float h2, const float cap, const float effD, DropRegs& d,
                                               const StepParams& P, float& carried) {
  float c_eq = cap * (hc - h2);  // water.h:127-128
  if (c_eq < 0.0f) c_eq = 0.0f;
  const float cdiff = c_eq - d.sed;
  const float e = effD * cdiff;
  d.sed += e;  // water.h:131
  carried = d.sed;
  d.sed = kRecip ? (float)((double)d.sed * P.inv_keep) : (float)((double)d.sed / P.keep);  // water.h:135
  d.vol = (float)((double)d.vol * P.keep);  // water.h:136
  return -e;                                // water.h:132
}

__device__ __forceinline__ float oob_h2(const float hc) { return (float)((double)hc - 0.002); }  // water.h:121-122

// World::cascade on a private fp32 3x3 block (sequential mode).  Block cell k = (dx+1)*3+(dy+1).
__device__ __forceinline__ unsigned cascade_block_f32(float (&B)[9], const unsigned inb, const StepParams& P) {
  constexpr int nk[8] = {0, 1, 2, 3, 5, 6, 7, 8};  // world.h:94-103 neighbour order
  float h[8], lim[8];
  bool in[8];
  const float hc0 = B[4];
  bool any = false;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    in[j] = (inb >> nk[j]) & 1u;
    h[j] = B[nk[j]];
    const bool diag = (nk[j] == 0 || nk[j] == 2 || nk[j] == 6 || nk[j] == 8);
    lim[j] = above_tenth(h[j]) ? (diag ? P.lim_diag : P.lim_axis) : 0.0f;  // world.h:143-148
    const float diff = hc0 - h[j];
    any |= in[j] && diff != 0.0f && (fabsf(diff) - lim[j]) > 0.0f;
  }
  if (!any) return 0u;  // the centre only changes through a transfer, so nothing can fire
  // world.h:129-131: ascending by height; libstdc++'s sort of <= 16 elements is an insertion sort,
  // i.e. stable: ties keep collection order.  rank = position in that order.
  int rank[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += (in[i] && (h[i] < h[j] || (h[i] == h[j] && i < j))) ? 1 : 0;
    rank[j] = in[j] ? r : 8;
  }
  unsigned transfers = 0;
#pragma unroll 1
  for (int r = 0; r < 8; r++) {
    float hn = 0.0f, ln = 0.0f;
    int sel = -1;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (rank[j] == r) { hn = h[j]; ln = lim[j]; sel = j; }
    if (sel < 0) break;
    const float diff = B[4] - hn;  // world.h:138: centre re-read, neighbour snapshot
    if (diff == 0.0f) continue;
    const float excess = fabsf(diff) - ln;
    if (excess <= 0.0f) continue;
    const float t = P.settling * excess / 2.0f;  // world.h:154
    const bool down = diff > 0.0f;               // world.h:157-164
    B[4] = down ? B[4] - t : B[4] + t;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j == sel) B[nk[j]] = down ? B[nk[j]] + t : B[nk[j]] - t;
    transfers++;
  }
  return transfers;
}

}  // namespace shx
