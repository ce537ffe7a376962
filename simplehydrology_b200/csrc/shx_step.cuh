// One "phase" of one drop: the World::cascade owed by the previous step followed by one
// Drop::descend, both on a private 3x3 block of heights around the drop's cell.
//
// Reference semantics (file:line into the reference tree):
//   Drop::descend    source/water.h:58-156
//   World::cascade   source/world.h:90-168
//   quad::_normal    source/cellpool.h:181-204
//   map::height/oob  source/cellpool.h:413-437, node::discharge :242-244
// In the reference, descend ends with cascade(pos) at the NEW position and the next descend
// starts by reading the normal at that same position, with nothing in between.  Fusing
// "cascade, then the next descend" into one phase therefore reads one 3x3 block per step
// instead of two, and all height changes of a phase (cascade transfers, erosion/deposition at
// the centre, the final sediment drop) land inside that block.
//
// The same template serves both modes:
//   HeightQ  : heights are Q5.26 integers, every change is quantised once and later added
//              atomically (batched lock-step mode; bit-comparable with oracle orc_ls_*)
//   HeightF  : heights are fp32 and changes are plain fp32 adds in the reference's order
//              (sequential mode; bit-comparable with the reference on maps where erf sees 0)
#pragma once
#include "../../include/shx.h"
#include "shx_math.cuh"

namespace shx {

struct StepParams {
  float maxAge, minVol, evapRate, depositionRate, entrainment, gravity, momentumTransfer;
  float maxdiff, settling, lod, mapscale, lrate;
};

struct HeightQ {
  typedef int32_t H;
  static __device__ __forceinline__ float f(H v) { return h_to_float(v); }
  static __device__ __forceinline__ H q(float x) { return h_quantize(x); }
  static __device__ __forceinline__ long long ledger(H v) { return (long long)v; }
};
struct HeightF {
  typedef float H;
  static __device__ __forceinline__ float f(H v) { return v; }
  static __device__ __forceinline__ H q(float x) { return x; }
  static __device__ __forceinline__ long long ledger(H) { return 0; }
};

struct DropRegs {
  float px, py, sx, sy, vol, sed;
  int age, flags;
};

// per-thread accumulators for the per-step counters; rare events go straight to global atomics
struct StepAcc {
  unsigned steps, transfers;
  long long fx_eroded, fx_sed_inflation;
};

// block cell k = (dx+1)*3 + (dy+1); bit k of `inb` says the cell exists (cellpool.h:413-419)
template <class T>
__device__ __forceinline__ unsigned cascade_block(typename T::H (&B)[9], const unsigned inb, const StepParams& P) {
  // world.h:94-103 neighbour order in block indices, and |offset| (world.h:123)
  constexpr int nk[8] = {0, 1, 2, 3, 5, 6, 7, 8};
  const float sq2 = sqrtf(2.0f);
  float h[8], lim[8];
  bool in[8];
  const float hc0 = T::f(B[4]);
  bool any = false;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    in[j] = (inb >> nk[j]) & 1u;
    h[j] = T::f(B[nk[j]]);
    const float d = (nk[j] == 0 || nk[j] == 2 || nk[j] == 6 || nk[j] == 8) ? sq2 : 1.0f;
    // world.h:143-148: slope allowance only where the neighbour is above 0.1 (double compare)
    lim[j] = ((double)h[j] > 0.1) ? d * P.maxdiff * P.lod : 0.0f;
    const float diff = hc0 - h[j];
    any |= in[j] && diff != 0.0f && (fabsf(diff) - lim[j]) > 0.0f;
  }
  // If nothing exceeds its allowance against the untouched centre, no transfer fires at all
  // (the centre only changes through a transfer), so the sort can be skipped.
  if (!any) return 0u;

  // world.h:129-131: ascending by height; libstdc++'s sort of <= 16 elements is an insertion
  // sort, i.e. stable, so ties keep collection order.  rank = position in that order.
  int rank[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += (in[i] && (h[i] < h[j] || (h[i] == h[j] && i < j))) ? 1 : 0;
    rank[j] = in[j] ? r : 8;
  }
  unsigned transfers = 0;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    float hn = 0.0f, ln = 0.0f;
    int sel = -1;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (rank[j] == r) { hn = h[j]; ln = lim[j]; sel = j; }
    if (sel < 0) continue;
    const float diff = T::f(B[4]) - hn;  // world.h:138: centre re-read, neighbour snapshot
    if (diff == 0.0f) continue;
    const float excess = fabsf(diff) - ln;
    if (excess <= 0.0f) continue;
    const typename T::H t = T::q(P.settling * excess / 2.0f);  // world.h:154
    const bool down = diff > 0.0f;                              // world.h:157-164
    B[4] = down ? B[4] - t : B[4] + t;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j == sel) B[nk[j]] = down ? B[nk[j]] + t : B[nk[j]] - t;
    transfers++;
  }
  return transfers;
}

struct StepOut {
  bool deposit;         // tracks to add at the (old) cell
  float t_d, t_mx, t_my;
  long long fx_event;   // Q5.26 amount deposited at termination (ledger), else 0
};

// load_h(nx, ny): height of an in-bounds cell outside the block (h2 two cells away)
template <class T, class LoadH>
__device__ __forceinline__ void phase_step(typename T::H (&B)[9], const unsigned inb, DropRegs& d, const float4 fld,
                                           const StepParams& P, const int size, const int ix, const int iy,
                                           LoadH&& load_h, StepAcc& acc, StepOut& out) {
  out.deposit = false;
  out.fx_event = 0;
  acc.steps++;
  if (d.flags & SHX_DROP_CASCADE) {  // water.h:151 of the previous call
    acc.transfers += cascade_block<T>(B, inb, P);
    d.flags &= ~SHX_DROP_CASCADE;
  }

  // cellpool.h:181-204.  height() of a missing cell is 0 (cellpool.h:433-437).  Each plane's
  // cross product (cellpool.h:188,191,195,198) written out is (-+80*dh_x, 1, -+80*dh_y); the
  // products with the literal zeros of the generic formula only decide the sign of a zero.
  const float hc = T::f(B[4]);
  const float hxm = (inb & (1u << 1)) ? T::f(B[1]) : 0.0f, hxp = (inb & (1u << 7)) ? T::f(B[7]) : 0.0f;
  const float hym = (inb & (1u << 3)) ? T::f(B[3]) : 0.0f, hyp = (inb & (1u << 5)) ? T::f(B[5]) : 0.0f;
  const float Bp = P.mapscale * (hxp - hc), Bm = P.mapscale * (hxm - hc);
  const float Ap = P.mapscale * (hyp - hc), Am = P.mapscale * (hym - hc);
  float nx = 0.0f, ny = 0.0f, nz = 0.0f;
  if (inb & (1u << 8)) { nx += -Bp; ny += 1.0f; nz += -Ap; }
  if (inb & (1u << 0)) { nx += Bm; ny += 1.0f; nz += Am; }
  if (inb & (1u << 6)) { nx += -Bp; ny += 1.0f; nz += Am; }
  if (inb & (1u << 2)) { nx += Bm; ny += 1.0f; nz += -Ap; }
  {
    const float l2 = nx * nx + ny * ny + nz * nz;
    if (sqrtf(l2) > 0.0f) {  // glm normalize = v * (1/sqrt(dot(v,v)))
      const float inv = 1.0f / sqrtf(l2);
      nx *= inv; ny *= inv; nz *= inv;
    }
  }

  if ((float)d.age > P.maxAge || d.vol < P.minVol) {  // water.h:74-82
    const typename T::H q = T::q(d.sed);
    B[4] += q;
    out.fx_event = T::ledger(q);
    d.flags = ((float)d.age > P.maxAge) ? SHX_DROP_DONE_AGE : SHX_DROP_DONE_VOL;
    return;
  }

  float effD = P.depositionRate * (1.0f - fld.w);  // water.h:86-87
  if (effD < 0.0f) effD = 0.0f;
  {
    const float g = P.lod * P.gravity;  // water.h:95
    d.sx += (g * nx) / d.vol;
    d.sy += (g * nz) / d.vol;
  }
  const float fx = fld.y, fy = fld.z;
  if (sqrtf(fx * fx + fy * fy) > 0.0f && sqrtf(d.sx * d.sx + d.sy * d.sy) > 0.0f) {  // water.h:97-99
    const float fi = 1.0f / sqrtf(fx * fx + fy * fy);
    const float si = 1.0f / sqrtf(d.sx * d.sx + d.sy * d.sy);
    const float dp = (fx * fi) * (d.sx * si) + (fy * fi) * (d.sy * si);
    const float k = P.lod * P.momentumTransfer * dp / (d.vol + fld.x);
    d.sx += k * fx;
    d.sy += k * fy;
  }
  if (sqrtf(d.sx * d.sx + d.sy * d.sy) > 0.0f) {  // water.h:108-109
    const float si = 1.0f / sqrtf(d.sx * d.sx + d.sy * d.sy);
    const float m = P.lod * sqrtf(2.0f);
    d.sx = m * (d.sx * si);
    d.sy = m * (d.sy * si);
  }
  d.px += d.sx;  // water.h:111
  d.py += d.sy;

  out.deposit = true;  // water.h:115-117: old cell, new speed
  out.t_d = d.vol;
  out.t_mx = d.vol * d.sx;
  out.t_my = d.vol * d.sy;

  const int nix = (int)d.px, niy = (int)d.py;  // truncation, as ivec2(vec2)
  // !(x >= 0) also catches NaN, which the reference's cvttss2si maps to INT_MIN (out of bounds)
  const bool oob = !(d.px > -1.0f) || !(d.py > -1.0f) || nix >= size || niy >= size;
  float h2;
  if (oob) {
    h2 = (float)((double)hc - 0.002);  // water.h:121-122
  } else {
    const int ddx = nix - ix, ddy = niy - iy;
    if (ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1) {
      const int kk = (ddx + 1) * 3 + (ddy + 1);
      typename T::H v = B[0];
#pragma unroll
      for (int k = 1; k < 9; k++) v = (k == kk) ? B[k] : v;
      h2 = T::f(v);
    } else {
      h2 = T::f(load_h(nix, niy));  // water.h:124
    }
  }
  float c_eq = (1.0f + P.entrainment * shx_erff(0.4f * fld.x)) * (hc - h2);  // water.h:127-128
  if (c_eq < 0.0f) c_eq = 0.0f;
  const float cdiff = c_eq - d.sed;
  const float e = effD * cdiff;
  d.sed += e;  // water.h:131
  {
    const typename T::H q = T::q(e);  // water.h:132
    B[4] -= q;
    acc.fx_eroded += T::ledger(q);
  }
  const float carried = d.sed;
  d.sed = (float)((double)d.sed / (1.0 - (double)P.evapRate));  // water.h:135
  d.vol = (float)((double)d.vol * (1.0 - (double)P.evapRate));  // water.h:136
  acc.fx_sed_inflation += t_quantize_d((double)d.sed - (double)carried);
  if (oob) {  // water.h:139-142
    d.vol = 0.0f;
    d.flags = SHX_DROP_DONE_OOB;
    return;
  }
  d.age++;                       // water.h:153
  d.flags |= SHX_DROP_CASCADE;   // water.h:151, executed at the start of the next phase
}

}  // namespace shx
