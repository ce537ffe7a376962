// shx CUDA kernels (sm_100a).  Layout in HBM (per context, rows [xlo, xlo+nrows) of the map),
// index i = (x - xlo)*size + y (x-major like the reference, math.h:11-14, but one global plane
// instead of 512^2 tiles):
//   hq   int2 per cell: the two Q5.26 height planes interleaved {plane0, plane1}.  A phase reads
//        plane p&1 and adds into plane (p+1)&1 (see descend_lockstep); interleaving puts both in
//        the same 32-byte sector, so the adds hit sectors the phase has just read.
//   rec  32-byte record per cell = one L2 sector:
//        {discharge, momentumx, momentumy, rootdensity}  fp32, read-only inside erode
//        {track_d, track_mx, track_my, pad}              int32 Q13.18 accumulators (RED targets)
// Sequential mode stores fp32 heights in hq[].x and fp32 tracks in the record's second half.
#pragma once
#include "shx_step.cuh"

namespace shx {

struct __align__(32) CellRec {
  float discharge, momentumx, momentumy, rootdensity;
  int32_t track_d, track_mx, track_my, pad;
};

struct MapView {
  int2* hq;
  CellRec* rec;
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
  // The caller has just passed a __syncthreads-class barrier (block_sum comes from
  // __syncthreads_count), so every RED of this CTA for this phase was issued before thread 0's fence.
  if (gridDim.x == 1) return block_sum;
  if (threadIdx.x == 0) {
    const unsigned slot = phase & 3u;
    if (block_sum) atomicAdd(&bar->active[slot], block_sum);
    __threadfence();
    atomicAdd(&bar->count, 1u);
    const unsigned target = (phase + 1u) * gridDim.x;
    unsigned seen;
    do {  // relaxed polling (an acquire load per iteration would invalidate L1 every time)
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&bar->count) : "memory");
    } while (seen < target);
    __threadfence();
    unsigned total;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(total) : "l"(&bar->active[slot]) : "memory");
    if (blockIdx.x == 0) bar->active[(phase + 2u) & 3u] = 0u;
    *s_total = total;
  }
  __syncthreads();
  return *s_total;
}

// order-preserving float <-> uint32 maps (for non-NaN inputs; -0 must be canonicalised by the caller)
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned b = __float_as_uint(f);
  return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float(u ^ (((u >> 31) - 1u) | 0x80000000u));
}

#ifdef SHX_PHASE_TIMING
// development aid: cycle stamps of thread 0 / block 0 accumulated per phase section
__device__ unsigned long long g_phase_timing[8];
__device__ unsigned g_exp;  // experiment switches: 1 no track atomics, 2 no height atomics, 4 no cascade
#define SHX_EXP(bit) (g_exp & (bit))
#define SHX_T(i) do { if (gid == 0) tstamp[i] = clock64(); } while (0)
#else
#define SHX_T(i) do { } while (0)
#define SHX_EXP(bit) 0
#endif

#define SHX_CE(a, b)                                 \
  do {                                               \
    const unsigned long long lo__ = min(a, b);       \
    const unsigned long long hi__ = max(a, b);       \
    a = lo__;                                        \
    b = hi__;                                        \
  } while (0)

// ---------------------------------------------------------------------------------------------
// K3: batched lock-step descend.  One thread per drop, drop state in registers, the 3x3 block of
// the current phase in shared memory.  Phase p:
//   * reads heights only from plane p&1 (never written during the phase),
//   * adds this phase's integer height deltas to plane (p+1)&1, together with the deltas of
//     phase p-1 ("catch-up": that plane was the read plane of phase p-1 and has not seen them),
//   * adds volume / momentum to the int32 track accumulators (write-only inside erode),
//   * one grid barrier.
// After the barrier plane (p+1)&1 holds exactly "heights after phase p", so every read is
// independent of thread timing and every write is an integer add: results do not depend on the
// order in which drops are scheduled and are run-to-run identical.
//
// Shared memory: s_B[9][nt] block heights, s_D[2][8][nt] neighbour deltas of this / the previous phase.
template <int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) descend_lockstep_kernel(const __grid_constant__ DescendArgs a) {
  extern __shared__ int32_t s_mem[];
  __shared__ unsigned s_total;
  const int tid = threadIdx.x, nt = blockDim.x;
  int32_t* s_B = s_mem + tid;                // s_B[k*nt]
  int32_t* s_D = s_mem + 9 * nt + tid;       // s_D[(buf*8 + j)*nt]
  const unsigned gid = blockIdx.x * nt + tid;
  const int size = a.m.size;
  int* const H = reinterpret_cast<int*>(a.m.hq);

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
  int dC_prev = 0;         // centre delta of the previous phase
  unsigned dmask_prev = 0; // neighbours that received a delta in the previous phase
  int pidx = 0;            // its centre cell
  unsigned steps = 0, transfers = 0;
  long long fx_eroded = 0, fx_inflation = 0;
  int tn = 0;

#ifdef SHX_PHASE_TIMING
  long long tstamp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  for (unsigned phase = 0;; ++phase) {
    const int rpar = (int)(phase & 1u), wpar = rpar ^ 1;
    const int cur = rpar * 8, prev = wpar * 8;
    SHX_T(0);

    // Issue this phase's gathers first (they only touch the read plane), then the catch-up adds of
    // the previous phase (write plane): the loads do not queue behind the atomics' round trip.
    const int ix = (int)d.px, iy = (int)d.py;  // water.h:60, truncation
    const int cidx = (ix - a.m.xlo) * size + iy;
    // cellpool.h:413-419 for the 9 cells of the block
    const unsigned xm = ix > 0, xp = ix < size - 1, ym = iy > 0, yp = iy < size - 1;
    const unsigned inb = (xm & ym) | (xm << 1) | ((xm & yp) << 2) | (ym << 3) | (1u << 4) | (yp << 5) |
                         ((xp & ym) << 6) | (xp << 7) | ((xp & yp) << 8);
    int v[9];
    float4 fld = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (alive) {
      const int* c = H + 2 * cidx + rpar;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const int off = (k / 3 - 1) * size + (k % 3 - 1);
        v[k] = ((inb >> k) & 1u) ? __ldcg(c + 2 * off) : 0;
      }
      fld = __ldg(reinterpret_cast<const float4*>(a.m.rec + cidx));
    }

    if (dC_prev | (int)dmask_prev) {  // catch-up of the previous phase's deltas
      if (dC_prev && !SHX_EXP(2)) atomicAdd(H + 2 * pidx + wpar, dC_prev);
      unsigned m = dmask_prev;
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1u;
        const int k = j + (j >> 2);
        const int off = ((k * 11) >> 5) * size - size + (k - 3 * ((k * 11) >> 5)) - 1;
        if (!SHX_EXP(2)) atomicAdd(H + 2 * (pidx + off) + wpar, s_D[(prev + j) * nt]);
      }
      dC_prev = 0;
      dmask_prev = 0;
    }

    if (alive) {
#pragma unroll
      for (int k = 0; k < 9; k++) s_B[k * nt] = v[k];
      int Bc = v[4];
      unsigned dmask = 0;
      steps++;
      SHX_T(1);

      if ((d.flags & SHX_DROP_CASCADE) && !SHX_EXP(4)) {  // World::cascade of the previous call, world.h:90-168
        d.flags &= ~SHX_DROP_CASCADE;
        const float hc0 = h_to_float(Bc);
        unsigned long long key[8];
        bool any = false;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int k = j + (j >> 2);  // world.h:94-103 neighbour order -> block index
          const bool in = (inb >> k) & 1u;
          const float h = h_to_float(v[k]);
          const bool diag = (0xA5u >> j) & 1u;
          const float lim = above_tenth(h) ? (diag ? a.P.lim_diag : a.P.lim_axis) : 0.0f;  // world.h:143-148
          const float diff = hc0 - h;
          any |= in && diff != 0.0f && (fabsf(diff) - lim) > 0.0f;
          // sort key: (height, collection index); missing cells sort last
          key[j] = in ? (((unsigned long long)f2ord(h) << 32) | (unsigned)j) : (0xFFFFFFFF00000008ull | (unsigned)j);
        }
        // The centre only changes through a transfer: if nothing exceeds its allowance against the
        // untouched centre, nothing fires at all.
        if (any) {
          // world.h:129-131 ascending by height; libstdc++ sorts <= 16 elements by insertion, i.e.
          // stably -- the collection index in the low word reproduces that order.  19-comparator network.
          SHX_CE(key[0], key[1]); SHX_CE(key[2], key[3]); SHX_CE(key[4], key[5]); SHX_CE(key[6], key[7]);
          SHX_CE(key[0], key[2]); SHX_CE(key[1], key[3]); SHX_CE(key[4], key[6]); SHX_CE(key[5], key[7]);
          SHX_CE(key[1], key[2]); SHX_CE(key[5], key[6]); SHX_CE(key[0], key[4]); SHX_CE(key[3], key[7]);
          SHX_CE(key[1], key[5]); SHX_CE(key[2], key[6]);
          SHX_CE(key[1], key[4]); SHX_CE(key[3], key[6]);
          SHX_CE(key[2], key[4]); SHX_CE(key[3], key[5]);
          SHX_CE(key[3], key[4]);
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const unsigned j = (unsigned)key[r];
            if (j < 8u) {
              const float hn = ord2f((unsigned)(key[r] >> 32));
              const float diff = h_to_float(Bc) - hn;  // world.h:138: centre re-read, neighbour snapshot
              const bool diag = (0xA5u >> j) & 1u;
              const float lim = above_tenth(hn) ? (diag ? a.P.lim_diag : a.P.lim_axis) : 0.0f;
              const float excess = fabsf(diff) - lim;
              if (diff != 0.0f && excess > 0.0f) {
                const int t = h_quantize(a.P.settling * excess / 2.0f);  // world.h:154
                const int s = diff > 0.0f ? t : -t;                       // world.h:157-164
                Bc -= s;
                const int k = (int)(j + (j >> 2));
                const int kx = (k * 11) >> 5;
                const int off = kx * size - size + (k - 3 * kx) - 1;
                s_B[k * nt] += s;
                s_D[(cur + (int)j) * nt] = s;
                dmask |= 1u << j;
                if (!SHX_EXP(2)) atomicAdd(H + 2 * (cidx + off) + wpar, s);
                transfers++;
              }
            }
          }
          s_B[4 * nt] = Bc;
        }
      }

      SHX_T(2);
      const float hc = h_to_float(Bc);
      const float hxm = xm ? h_to_float(s_B[1 * nt]) : 0.0f, hxp = xp ? h_to_float(s_B[7 * nt]) : 0.0f;
      const float hym = ym ? h_to_float(s_B[3 * nt]) : 0.0f, hyp = yp ? h_to_float(s_B[5 * nt]) : 0.0f;
      const float sed_before = d.sed;
      const StepResult res = descend_math(hc, hxm, hxp, hym, hyp, inb, d, fld, a.P, size, [&](int nix, int niy) {
        const int ddx = nix - ix, ddy = niy - iy;
        int hv;
        if (ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1) hv = s_B[((ddx + 1) * 3 + (ddy + 1)) * nt];
        else hv = __ldcg(H + 2 * ((nix - a.m.xlo) * size + niy) + rpar);
        return h_to_float(hv);
      });
      SHX_T(3);
      const int q = h_quantize(res.dheight);
      const int dC = (Bc - v[4]) + q;
      if (res.moved) {
        int* t = &a.m.rec[cidx].track_d;
        if (!SHX_EXP(1)) {
          atomicAdd(t, t_quantize(res.t_d));
          atomicAdd(t + 1, t_quantize(res.t_mx));
          atomicAdd(t + 2, t_quantize(res.t_my));
        }
        fx_eroded -= (long long)q;
        // growth of the carried sediment by water.h:135 (sediment after the exchange = before + e)
        fx_inflation += l_quantize(d.sed) - l_quantize(sed_before - res.dheight);
      }
      if (!(d.flags & SHX_DROP_ALIVE)) {  // terminated in this phase (once per drop)
        alive = false;
        atomicMax(&a.bar->max_steps, steps);
        if (d.flags & SHX_DROP_DONE_OOB) {
          stat_add(a.stats, ST_TERM_OOB, 1ull);
          stat_add(a.stats, ST_FX_SED_OOB, (unsigned long long)l_quantize(d.sed));
        } else {
          stat_add(a.stats, (d.flags & SHX_DROP_DONE_AGE) ? ST_TERM_AGE : ST_TERM_VOL, 1ull);
          stat_add(a.stats, ST_FX_DEPOSITED, (unsigned long long)(long long)q);
          stat_add(a.stats, ST_FX_SED_DEPOSITED, (unsigned long long)l_quantize(d.sed));
        }
      } else {
        const int nix = (int)d.px;
        if (nix < a.m.row0 || nix >= a.m.row1) {  // left the strip: hand over (cascade still owed)
          const bool tolo = nix < a.m.row0;
          d.flags = (d.flags & ~SHX_DROP_ALIVE) | (tolo ? SHX_DROP_MIGRATE_LO : SHX_DROP_MIGRATE_HI);
          alive = false;
          atomicMax(&a.bar->max_steps, steps);
          stat_add(a.stats, tolo ? ST_MIGRATED_LO : ST_MIGRATED_HI, 1ull);
        }
      }
      if (dC && !SHX_EXP(2)) atomicAdd(H + 2 * cidx + wpar, dC);
      dC_prev = dC;
      dmask_prev = dmask;
      pidx = cidx;
      if (a.trace != nullptr && gid == 0 && tn < a.trace_cap) {
        float* t = a.trace + 7 * (size_t)tn++;
        t[0] = (float)d.age; t[1] = d.px; t[2] = d.py; t[3] = d.sx; t[4] = d.sy; t[5] = d.vol; t[6] = d.sed;
      }
    }

    SHX_T(4);
    const unsigned block_sum = (unsigned)__syncthreads_count(alive || (dC_prev | (int)dmask_prev));
    SHX_T(5);
    const unsigned total = grid_barrier_sum(a.bar, phase, block_sum, &s_total);
    SHX_T(6);
#ifdef SHX_PHASE_TIMING
    if (gid == 0 && alive) {
      for (int i = 0; i < 6; i++) g_phase_timing[i] += (unsigned long long)(tstamp[i + 1] - tstamp[i]);
      g_phase_timing[7] += 1ull;
    }
#endif
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
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    steps += __shfl_xor_sync(0xffffffffu, steps, o);
    transfers += __shfl_xor_sync(0xffffffffu, transfers, o);
    fx_eroded += __shfl_xor_sync(0xffffffffu, fx_eroded, o);
    fx_inflation += __shfl_xor_sync(0xffffffffu, fx_inflation, o);
  }
  if ((tid & 31) == 0 && steps) {
    stat_add(a.stats, ST_STEPS, steps);
    stat_add(a.stats, ST_TRANSFERS, transfers);
    stat_add(a.stats, ST_FX_ERODED, (unsigned long long)fx_eroded);
    stat_add(a.stats, ST_FX_SED_INFLATION, (unsigned long long)fx_inflation);
  }
}

// ---------------------------------------------------------------------------------------------
// Sequential mode (parity anchor): one thread marches the drops one after another in fp32,
// operation for operation what World::erode's inner loop does (world.h:74-76).  <<<1,1>>>.
struct SequentialArgs {
  MapView m;  // hq[].x reinterpreted as fp32 height, record tracks as fp32
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
  const int size = a.m.size;
  float* const H = reinterpret_cast<float*>(a.m.hq);
  unsigned steps = 0, transfers = 0;
  long long fx_inflation = 0;
  int tn = 0;
  for (unsigned i = 0; i < a.ndrops; i++) {
    const shx_drop r = a.drops[i];
    DropRegs d = {r.px, r.py, r.sx, r.sy, r.volume, r.sediment, r.age, r.flags};
    while (d.flags & SHX_DROP_ALIVE) {
      const int ix = (int)d.px, iy = (int)d.py;
      const int cidx = ix * size + iy;
      unsigned inb = 0;
      float B[9];
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const int x = ix + k / 3 - 1, y = iy + k % 3 - 1;
        const bool in = x >= 0 && y >= 0 && x < size && y < size;
        inb |= in ? (1u << k) : 0u;
        B[k] = in ? H[2 * (cidx + (k / 3 - 1) * size + (k % 3 - 1))] : 0.0f;
      }
      CellRec* rec = a.m.rec + cidx;
      const float4 fld = *reinterpret_cast<const float4*>(rec);
      steps++;
      if (d.flags & SHX_DROP_CASCADE) {
        transfers += cascade_block_f32(B, inb, a.P);
        d.flags &= ~SHX_DROP_CASCADE;
      }
      const float sed_before = d.sed;
      const StepResult res = descend_math(B[4], (inb & 2u) ? B[1] : 0.0f, (inb & 128u) ? B[7] : 0.0f,
                                          (inb & 8u) ? B[3] : 0.0f, (inb & 32u) ? B[5] : 0.0f, inb, d, fld, a.P, size,
                                          [&](int nix, int niy) {
                                            const int ddx = nix - ix, ddy = niy - iy;
                                            if (ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1) {
                                              float hv = B[0];
#pragma unroll
                                              for (int k = 1; k < 9; k++) hv = (k == (ddx + 1) * 3 + (ddy + 1)) ? B[k] : hv;
                                              return hv;
                                            }
                                            return H[2 * (nix * size + niy)];
                                          });
      B[4] = B[4] + res.dheight;  // water.h:75,80 (+= sediment) / :132 (-= effD*cdiff)
#pragma unroll
      for (int k = 0; k < 9; k++)
        if (inb & (1u << k)) H[2 * (cidx + (k / 3 - 1) * size + (k % 3 - 1))] = B[k];
      if (res.moved) {  // water.h:115-117
        float* t = reinterpret_cast<float*>(&rec->track_d);
        t[0] += res.t_d; t[1] += res.t_mx; t[2] += res.t_my;
        fx_inflation += l_quantize(d.sed) - l_quantize(sed_before - res.dheight);
      }
      if (!(d.flags & SHX_DROP_ALIVE)) {
        if (d.flags & SHX_DROP_DONE_OOB) {
          a.stats[ST_TERM_OOB] += 1ull;
          a.stats[ST_FX_SED_OOB] += (unsigned long long)l_quantize(d.sed);
        } else {
          a.stats[(d.flags & SHX_DROP_DONE_AGE) ? ST_TERM_AGE : ST_TERM_VOL] += 1ull;
          a.stats[ST_FX_SED_DEPOSITED] += (unsigned long long)l_quantize(d.sed);
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
  a.stats[ST_STEPS] += steps;
  a.stats[ST_TRANSFERS] += transfers;
  a.stats[ST_FX_SED_INFLATION] += (unsigned long long)fx_inflation;
}

// ---------------------------------------------------------------------------------------------
// K2: spawn.  world.h:64-74 with rand() replaced by a counter-based hash keyed
// (seed, epoch, node, i): node-major, `cycles` drops per node, reject where height < 0.1.
// node0/nnodes select the nodes of this strip (all of them for a whole map).
struct SpawnArgs {
  MapView m;
  int sequential;
  int tilesize, mapsize;
  unsigned node0, nnodes;
  int cycles;
  uint64_t key;
  shx_drop* drops;
  float* xy;  // optional copy of the positions
  unsigned long long* stats;
};

__device__ __forceinline__ shx_drop make_drop(float x, float y, const MapView& m, int sequential, unsigned long long* stats) {
  shx_drop d = {x, y, 0.0f, 0.0f, 1.0f, 0.0f, 0, SHX_DROP_ALIVE};  // water.h:14-23
  const int ix = (int)x, iy = (int)y;
  const bool oob = !(x > -1.0f) || !(y > -1.0f) || ix >= m.size || iy >= m.size;
  if (!oob && (ix < m.row0 || ix >= m.row1)) {  // not this strip's drop
    d.flags = 0;
    return d;
  }
  float h = 0.0f;  // map.height() of a missing cell (cellpool.h:433-437)
  if (!oob) {
    const int2 hv = m.hq[(ix - m.xlo) * m.size + iy];
    h = sequential ? __int_as_float(hv.x) : h_to_float(hv.x);
  }
  if (!above_tenth(h)) {  // world.h:71-72  (double)h < 0.1
    d.flags = SHX_DROP_REJECTED;
    atomicAdd(stats + ST_REJECTED, 1ull);
  } else {
    atomicAdd(stats + ST_SPAWNED, 1ull);
  }
  return d;
}

__global__ void spawn_kernel(const SpawnArgs a) {
  const unsigned n = a.nnodes * (unsigned)a.cycles;
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const unsigned node = a.node0 + k / (unsigned)a.cycles, i = k % (unsigned)a.cycles;
    const uint64_t r = mix64(a.key + (((uint64_t)node << 32) | (uint64_t)i));
    const int nx = (int)(node / (unsigned)a.mapsize) * a.tilesize, ny = (int)(node % (unsigned)a.mapsize) * a.tilesize;
    const float x = (float)(nx + (int)((uint32_t)r % (uint32_t)a.tilesize));
    const float y = (float)(ny + (int)((uint32_t)(r >> 32) % (uint32_t)a.tilesize));
    if (a.xy) { a.xy[2 * k] = x; a.xy[2 * k + 1] = y; }
    a.drops[k] = make_drop(x, y, a.m, a.sequential, a.stats);
  }
}

__global__ void make_drops_kernel(const float* xy, unsigned n, const MapView m, int sequential, shx_drop* drops,
                                  unsigned long long* stats) {
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    drops[k] = make_drop(xy[2 * k], xy[2 * k + 1], m, sequential, stats);
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
    if (reset) reinterpret_cast<int4*>(rec + i)[1] = make_int4(0, 0, 0, 0);
  }
}

__global__ void reset_tracks_kernel(CellRec* __restrict__ rec, size_t n) {  // world.h:56-61
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    reinterpret_cast<int4*>(rec + i)[1] = make_int4(0, 0, 0, 0);
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
      a.m.hq[i] = make_int2(__float_as_int(lo.x), 0);
      reinterpret_cast<float4*>(a.m.rec + i)[1] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    } else {
      if (!(fabsf(lo.x) < 31.0f) || !(fabsf(hi.x) < 4096.0f)) *a.error_flag = 1;
      const int32_t q = h_quantize(lo.x);
      a.m.hq[i] = make_int2(q, q);
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
    const int2 hv = a.m.hq[i];
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
      m.hq[i] = make_int2(__float_as_int(h), 0);
    } else {
      const int32_t q = h_quantize(h);
      m.hq[i] = make_int2(q, q);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row-strip exchange helpers (multi-GPU).  A strip keeps `halo` rows of its neighbours' heights on
// each side.  Cascade transfers of drops on the strip's boundary rows land in those halo rows;
// `halo_ref` remembers what the halo held at the last refresh, so (current - ref) is exactly the
// integer amount this strip owes the owner.  Outside a run both planes are equal: plane 0 is used.
__global__ void strip_halo_delta_kernel(const int2* cur, const int32_t* ref, int32_t* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = cur[i].x - ref[i];
}
__global__ void strip_add_rows_kernel(int2* h, const int32_t* delta, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int32_t v = delta[i];
    if (v) {
      int2 c = h[i];
      c.x += v; c.y += v;
      h[i] = c;
    }
  }
}
__global__ void strip_get_rows_kernel(const int2* h, int32_t* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = h[i].x;
}
__global__ void strip_set_rows_kernel(int2* h, int32_t* ref, const int32_t* src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int32_t v = src[i];
    h[i] = make_int2(v, v);
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

}  // namespace shx
