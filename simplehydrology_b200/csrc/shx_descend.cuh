// K3 for small batches (a few thousand drops: the reference's default call): descend_group_kernel, EIGHT LANES PER DROP
// (included by shx_kernels.cuh, which holds the shared pieces: the map view, DescendArgs, the grid barrier, claim
// keys, and the one-thread-per-drop kernel descend_lockstep_kernel used for dense batches and peer mode).
//
// Why: a lock-step phase is a chain of dependent work that every warp walks between two barriers; with one thread
// per drop that chain is ~1 500 warp-instructions (profiles/r1d_ncu_summary.txt) however few drops there are, which
// is what bounds the reference's own sizes (512 drops per call at 512^2, 8 192 at 2048^2) and the strips of a
// multi-GPU run at 6-9 us per phase.  Here lane j of a group of eight owns neighbour j of the drop's 3x3 block
// (world.h:94-103 order): the gather is one load per lane with no shuffles, the cascade's excess test and ranking
// run across the lanes, only its Gauss-Seidel chain stays serial (eight shuffle-broadcast steps), every lane adds
// its own neighbour's delta, and the scalar part of Drop::descend (move_math / exchange_math) is computed
// redundantly by the eight lanes.  Same schedule, same arithmetic, bit-identical results (tests/test_gpu_batched.py).
#pragma once

namespace shx {

// bit j (neighbour order of world.h:94-103: j -> block index k = j + (j >> 2)) set if that cell exists, from the edge
// flags {x > 0, x < size-1, y > 0, y < size-1}
__device__ __forceinline__ unsigned valid8_from_edges(unsigned ef) {
  return ((ef & 1u) ? 0xFFu : 0xF8u) & ((ef & 2u) ? 0xFFu : 0x1Fu) & ((ef & 4u) ? 0xFFu : 0xD6u) & ((ef & 8u) ? 0xFFu : 0x6Bu);
}
// the nine bits move_math takes (bit k of the 3x3 block)
__device__ __forceinline__ unsigned inb9_from_valid8(unsigned v8) { return (v8 & 0xFu) | 0x10u | ((v8 & 0xF0u) << 1); }

__device__ __forceinline__ int neighbour_offset(int j, int size) {  // cell index offset of neighbour j
  const int k = j + (j >> 2), kx = (k * 11) >> 5;
  return (kx - 1) * size + (k - 3 * kx - 1);
}

// Dynamic shared memory per lane: one int2 {height, neighbour index} of the group's cascade order and one int of the
// heights handed back in natural order.
constexpr int kGroupSmemWords = 3;

template <int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) descend_group_kernel(const __grid_constant__ DescendArgs a) {
  extern __shared__ int32_t s_mem[];
  __shared__ unsigned s_total;
  __shared__ unsigned s_hi[2];  // grid barrier bookkeeping, touched by thread 0 only
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  if (tid == 0) s_hi[0] = s_hi[1] = 0u;
  const int j = lane & 7, gl = lane & ~7;  // neighbour index of this lane, first lane of its group
  const unsigned gmask = 0xFFu << gl;
  int2* const s_sorted = reinterpret_cast<int2*>(s_mem) + (tid & ~7);  // [8] per group
  int* const s_back = s_mem + 2 * nt + (tid & ~7);                      // [8] per group
  const int size = a.m.size, xlo = a.m.xlo;
  int* const H = reinterpret_cast<int*>(a.m.hq);
  CellRec* const REC = a.m.rec;
  const int my_off = neighbour_offset(j, size);
  const unsigned my_need = [&] {  // which sides of the map this lane's neighbour needs (edge flags, see below)
    const int k = j + (j >> 2), dx = k / 3 - 1, dy = k % 3 - 1;
    return (dx < 0 ? 1u : 0u) | (dx > 0 ? 2u : 0u) | (dy < 0 ? 4u : 0u) | (dy > 0 ? 8u : 0u);
  }();
  const bool my_diag = (0xA5u >> j) & 1u;
  const unsigned gdrop = blockIdx.x * (unsigned)(nt >> 3) + (unsigned)(tid >> 3);
  const bool leader = j == 0;

  DropRegs d;  // replicated in the eight lanes of the group
  d.px = d.py = d.sx = d.sy = d.vol = d.sed = 0.0f;
  d.age = 0;
  d.flags = 0;
  if (gdrop < a.ndrops) {
    const float4 lo = reinterpret_cast<const float4*>(a.drops + gdrop)[0];
    const float4 hi = reinterpret_cast<const float4*>(a.drops + gdrop)[1];
    d.px = lo.x; d.py = lo.y; d.sx = lo.z; d.sy = lo.w;
    d.vol = hi.x; d.sed = hi.y; d.age = __float_as_int(hi.z); d.flags = __float_as_int(hi.w);
  }
  bool alive = (d.flags & SHX_DROP_ALIVE) != 0;
  bool asleep = a.align_age != 0u && alive && d.age > 0;  // carried over from the previous call: sleeps until phase == age
  alive = alive && !asleep;
  unsigned mykey = 0u;     // the key this drop claimed its cell with for the coming phase
  int pc = 0, dC_prev = 0;  // centre cell and centre delta of the previous phase (owed to the other plane)
  int pend_cell = 0, pend_val = 0;  // this lane's cascade transfer of the previous phase (likewise)
  unsigned steps = 0, transfers = 0;
  long long fx_eroded = 0, fx_inflation = 0;
  int tn = 0;

  auto cell_of = [&](float px, float py) { return ((int)px - xlo) * size + (int)py; };
  auto trace_row = [&]() {
    if (a.trace != nullptr && gdrop == 0 && leader && tn < a.trace_cap) {
      float* t = a.trace + 7 * (size_t)tn++;
      t[0] = (float)d.age; t[1] = d.px; t[2] = d.py; t[3] = d.sx; t[4] = d.sy; t[5] = d.vol; t[6] = d.sed;
    }
  };
  auto claim = [&](int cell, int word, unsigned phase_tag) {  // claim `cell` for the phase with this tag
    mykey = claim_key(a.claim_epoch, phase_tag, d);
    if (leader) red_claim(reinterpret_cast<unsigned*>(H + 4 * (size_t)cell + word + 1), mykey);
  };

  if (alive) claim(cell_of(d.px, d.py), 0, 1u);  // every drop that is awake in phase 0 claims its cell (tag 1, parity 0)
  {
    const unsigned block_sum = (unsigned)__syncthreads_count(alive || asleep);
    grid_barrier_sum<true>(a.bar, 0u, block_sum, &s_total, s_hi[0], s_hi[1]);
  }

  for (unsigned phase = 0;; ++phase) {
    const int rw = 2 * (int)(phase & 1u), ww = 2 - rw;  // word of the read / write plane inside a cell {h0, claim0, h1, claim1}
    if (asleep && (unsigned)d.age <= phase) {
      asleep = false;
      alive = true;
    }
    const int ix = (int)d.px, iy = (int)d.py;  // water.h:60, truncation
    const int cidx = (ix - xlo) * size + iy;
    // cellpool.h:413-419 for the block: which of the four sides exist
    const unsigned ef = (ix > 0 ? 1u : 0u) | (ix < size - 1 ? 2u : 0u) | (iy > 0 ? 4u : 0u) | (iy < size - 1 ? 8u : 0u);
    const bool valid = (ef & my_need) == my_need;  // this lane's neighbour cell exists

    // gather: one {height, claim} pair per lane, the centre pair and the cell record once per group
    int2 nb = make_int2(0, 0), cc = make_int2(0, 0);
    float4 fld = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (alive) {
      if (valid) nb = __ldcg(reinterpret_cast<const int2*>(H + 4 * (size_t)(cidx + my_off) + rw));
      cc = __ldcg(reinterpret_cast<const int2*>(H + 4 * (size_t)cidx + rw));
      fld = __ldg(reinterpret_cast<const float4*>(REC + cidx));
    }
    // "catch-up": the previous phase's deltas for the plane that was being read then
    if (pend_val) {
      red_height(H + 4 * (size_t)pend_cell + ww, pend_val);
      pend_val = 0;
    }
    if (dC_prev) {
      if (leader) red_height(H + 4 * (size_t)pc + ww, dC_prev);
      dC_prev = 0;
    }
    const int hC = cc.x;
    // how many of the eight cells around hold a higher key this phase
    const unsigned crowd = __ballot_sync(0xffffffffu, (unsigned)nb.y > mykey);
    const int crowded = __popc((crowd >> gl) & 0xFFu);
    // whose turn is it on this cell?  (claimed during the previous phase, complete since its barrier)
    const bool turn = alive && (unsigned)cc.y == mykey;

    if (alive && !turn) {  // a drop with a higher key has the cell: wait
      if (wait_one_phase(d, a)) {  // expired in the queue: the sediment stays here (water.h:74-77)
        const int q = h_quantize(d.sed);
        if (leader) {
          if (q) red_height(H + 4 * (size_t)cidx + ww, q);
          atomicMax(&a.bar->max_steps, phase + 1u);
          stat_add(a.stats, ST_TERM_AGE, 1ull);
          stat_add(a.stats, ST_FX_DEPOSITED, (unsigned long long)(long long)q);
          stat_add(a.stats, ST_FX_SED_DEPOSITED, (unsigned long long)l_quantize(d.sed));
        }
        dC_prev = q;  // the other plane gets it in the next phase, like any other delta
        pc = cidx;
        alive = false;
        d.flags = SHX_DROP_DONE_AGE;
      } else {
        claim(cidx, ww, phase + 2u);
      }
    }
    if (asleep && (unsigned)d.age == phase + 1u) claim(cell_of(d.px, d.py), ww, phase + 2u);  // wakes up in the next phase

    if (turn) {
      // 1, or 2^-n next to n cells that hold a higher key: what this drop moves (cascade transfers and the
      // sediment exchange) is scaled down, so that the changes of neighbouring cells in one phase do not add up
      const float damp = __int_as_float((127 - crowded) << 23);
      steps++;
      d.flags &= ~(7 << kWaitedShift);
      int hN = nb.x;  // this lane's neighbour (0 if the cell does not exist)
      int Bc = hC;

      if (d.flags & SHX_DROP_CASCADE) {  // World::cascade of the previous call, world.h:90-168
        d.flags &= ~SHX_DROP_CASCADE;
        const float h = h_to_float(hN);
        const float lim = above_tenth(h) ? (my_diag ? a.P.lim_diag : a.P.lim_axis) : 0.0f;  // world.h:143-148
        // The centre only changes through a transfer: if no neighbour exceeds its allowance against the untouched
        // centre, nothing fires at all (excess > 0 implies diff != 0 because lim >= 0).
        const bool fire = valid && (fabsf(h_to_float(Bc) - h) - lim) > 0.0f;
        if (__ballot_sync(gmask, fire)) {
          // world.h:129-131 ascending by height; libstdc++ sorts <= 16 elements by insertion, i.e. stably: i comes
          // before j iff h_i < h_j, or h_i == h_j and i < j.  Missing cells sort last.
          const float hh = valid ? h : __int_as_float(0x7f800000);
          int rank = 0;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float hi = __shfl_sync(gmask, hh, gl + i);
            rank += (hi < hh || (hi == hh && i < j)) ? 1 : 0;
          }
          s_sorted[rank] = make_int2(hN, valid ? j : 8);
          __syncwarp(gmask);
          const int2 e = s_sorted[j];  // lane r now holds the neighbour of rank r
          const int jn = e.y;
          const float hn = h_to_float(e.x);
          const float lim_n = above_tenth(hn) ? (((0xA5u >> jn) & 1u) ? a.P.lim_diag : a.P.lim_axis) : 0.0f;
          int my_s = 0;
          bool my_fired = false;  // a transfer that rounds to zero height units still counts as one (world.h:154)
#pragma unroll
          for (int r = 0; r < 8; r++) {  // the Gauss-Seidel chain: lane r moves, everybody follows the centre
            const float diff = h_to_float(Bc) - hn;  // world.h:138: centre re-read, neighbour snapshot
            const float excess = fabsf(diff) - lim_n;
            const bool fires = jn < 8 && diff != 0.0f && excess > 0.0f;
            int s = 0;
            if (fires) {
              const int t = h_quantize((a.P.settling * damp) * excess / 2.0f);  // world.h:154
              s = diff > 0.0f ? t : -t;                                         // world.h:157-164
            }
            if (j == r) {
              my_s = s;
              my_fired = fires;
            }
            Bc -= __shfl_sync(gmask, s, gl + r);
          }
          if (jn < 8) s_back[jn] = e.x + my_s;
          if (my_s) {
            pend_cell = cidx + neighbour_offset(jn, size);
            pend_val = my_s;
            red_height(H + 4 * (size_t)pend_cell + ww, my_s);
          }
          transfers += (unsigned)__popc(__ballot_sync(gmask, my_fired));
          __syncwarp(gmask);
          hN = valid ? s_back[j] : 0;
        }
      }

      const float hc = h_to_float(Bc);
      const int q_xm = __shfl_sync(gmask, hN, gl + 1), q_xp = __shfl_sync(gmask, hN, gl + 6);
      const int q_ym = __shfl_sync(gmask, hN, gl + 3), q_yp = __shfl_sync(gmask, hN, gl + 4);
      const float hxm = (ef & 1u) ? h_to_float(q_xm) : 0.0f, hxp = (ef & 2u) ? h_to_float(q_xp) : 0.0f;
      const float hym = (ef & 4u) ? h_to_float(q_ym) : 0.0f, hyp = (ef & 8u) ? h_to_float(q_yp) : 0.0f;
      const MoveResult mv = move_math(hc, hxm, hxp, hym, hyp, inb9_from_valid8(valid8_from_edges(ef)), d, fld, a.P, size);
      int dC = Bc - hC;
      if (!mv.moved) {  // water.h:74-82: aged out / dried up, the sediment stays here
        const int q = h_quantize(mv.dheight);
        dC += q;
        alive = false;
        if (leader) {
          atomicMax(&a.bar->max_steps, phase + 1u);
          stat_add(a.stats, (d.flags & SHX_DROP_DONE_AGE) ? ST_TERM_AGE : ST_TERM_VOL, 1ull);
          stat_add(a.stats, ST_FX_DEPOSITED, (unsigned long long)(long long)q);
          stat_add(a.stats, ST_FX_SED_DEPOSITED, (unsigned long long)l_quantize(d.sed));
        }
        trace_row();
      } else {
        // water.h:124: the new cell (nearest, truncated): inside the block it is one of the lanes' heights,
        // otherwise one dependent load
        const int nix = (int)d.px, niy = (int)d.py;
        const int ncidx = (nix - xlo) * size + niy;
        int hv = 0;
        if (!mv.oob) {
          const int ddx = nix - ix, ddy = niy - iy;
          if (ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1) {
            const int k = (ddx + 1) * 3 + (ddy + 1);
            hv = __shfl_sync(gmask, hN, gl + (k > 4 ? k - 1 : k & 7));
            if (k == 4) hv = Bc;  // a drop without speed stays where it is
          } else {
            hv = __ldcg(H + 4 * (size_t)ncidx + rw);
          }
        }
        const float cap = 1.0f + a.P.entrainment * shx_erff(0.4f * fld.x);  // water.h:127, cellpool.h:242-244
        if (j < 3) {  // water.h:115-117: lanes 0..2 add the three track amounts
          const float tv = j == 0 ? mv.t_d : (j == 1 ? mv.t_mx : mv.t_my);
          red_track32(&REC[cidx].track_d + j, t_quantize(tv));
        }
        const float h2 = mv.oob ? oob_h2(hc) : h_to_float(hv);  // water.h:121-124
        float carried;
        // The exchange is halved for every cell around that holds a higher key: neighbouring cells that change
        // in the same phase form an explicit scheme whose factors (up to 1.1 per cell) must not add up.
        const float dh = exchange_math<true>(hc, h2, cap, mv.effD * damp, d, a.P, carried);  // water.h:127-136
        const int q = h_quantize(dh);
        dC += q;
        fx_eroded -= (long long)q;
        fx_inflation += l_quantize(d.sed) - l_quantize(carried);
        if (mv.oob) {  // water.h:139-142
          alive = false;
          if (leader) {
            atomicMax(&a.bar->max_steps, phase + 1u);
            stat_add(a.stats, ST_TERM_OOB, 1ull);
            stat_add(a.stats, ST_FX_SED_OOB, (unsigned long long)l_quantize(d.sed));
          }
          d.vol = 0.0f;
          d.flags = SHX_DROP_DONE_OOB;
        } else {
          d.age++;                      // water.h:153
          d.flags |= SHX_DROP_CASCADE;  // water.h:151, executed at the start of the next phase
          if (nix >= a.m.row0 && nix < a.m.row1) {
            claim(ncidx, ww, phase + 2u);
          } else {  // left the strip: hand over (cascade still owed)
            const bool tolo = nix < a.m.row0;
            d.flags = (d.flags & ~SHX_DROP_ALIVE) | (tolo ? SHX_DROP_MIGRATE_LO : SHX_DROP_MIGRATE_HI);
            alive = false;
            if (leader) {
              atomicMax(&a.bar->max_steps, phase + 1u);
              stat_add(a.stats, tolo ? ST_MIGRATED_LO : ST_MIGRATED_HI, 1ull);
            }
          }
        }
        trace_row();
      }
      if (j == 3 && dC) red_height(H + 4 * (size_t)cidx + ww, dC);
      dC_prev = dC;
      pc = cidx;
    }

    const unsigned block_sum = (unsigned)__syncthreads_count(alive || asleep || (dC_prev | pend_val));
    if (grid_barrier_sum<true>(a.bar, phase + 1u, block_sum, &s_total, s_hi[0], s_hi[1]) == 0u) break;
    if (phase + 3u >= kMaxPhases) {  // the claim tag would wrap: give up (the host reports SHX_ERR_RANGE)
      if (blockIdx.x == 0 && tid == 0) atomicOr(a.abort_flag, 2);
      break;
    }
  }
  // every termination's atomicMax happened before the last barrier
  if (blockIdx.x == 0 && tid == 0) stat_add(a.stats, ST_PHASES, (unsigned long long)__ldcg(&a.bar->max_steps));

  if (gdrop < a.ndrops && leader) {
    float4 lo, hi;
    lo.x = d.px; lo.y = d.py; lo.z = d.sx; lo.w = d.sy;
    hi.x = d.vol; hi.y = d.sed; hi.z = __int_as_float(d.age); hi.w = __int_as_float(d.flags);
    reinterpret_cast<float4*>(a.drops + gdrop)[0] = lo;
    reinterpret_cast<float4*>(a.drops + gdrop)[1] = hi;
  }
  if (a.trace_n != nullptr && gdrop == 0 && leader) *a.trace_n = tn;

  // per-step counters (kept by every lane, counted once per drop): warp reduce, one atomic per warp
  if (!leader) {
    steps = 0;
    transfers = 0;
    fx_eroded = 0;
    fx_inflation = 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    steps += __shfl_xor_sync(0xffffffffu, steps, o);
    transfers += __shfl_xor_sync(0xffffffffu, transfers, o);
    fx_eroded += __shfl_xor_sync(0xffffffffu, fx_eroded, o);
    fx_inflation += __shfl_xor_sync(0xffffffffu, fx_inflation, o);
  }
  if (lane == 0 && steps) {
    stat_add(a.stats, ST_STEPS, steps);
    stat_add(a.stats, ST_TRANSFERS, transfers);
    stat_add(a.stats, ST_FX_ERODED, (unsigned long long)fx_eroded);
    stat_add(a.stats, ST_FX_SED_INFLATION, (unsigned long long)fx_inflation);
  }
}

}  // namespace shx
