// shx CUDA kernels (sm_100a).  Layout in HBM (per context, rows [xlo, xlo+nrows) of the map),
// index i = (x - xlo)*size + y (x-major like the reference, math.h:11-14, but one global plane
// instead of 512^2 tiles):
//   hq   int4 per cell: the two Q5.26 height planes and the two claim words {plane0, claim0, plane1, claim1}.  A phase reads
//        plane p&1 and adds into plane (p+1)&1 (see descend_lockstep); interleaving puts all four in
//        the same 32-byte sector, so the adds and claims hit sectors the phase has just read.
//   rec  32-byte record per cell = one L2 sector:
//        {discharge, momentumx, momentumy, rootdensity}  fp32, read-only inside erode
//        {track_d, track_mx, track_my, pad}              int32 Q13.18 accumulators (RED targets)
// Sequential mode stores fp32 heights in hq[].x and fp32 tracks in the record's second half.
#pragma once
#include <type_traits>
#include "shx_step.cuh"

namespace shx {

struct __align__(32) CellRec {
  float discharge, momentumx, momentumy, rootdensity;
  int32_t track_d, track_mx, track_my, pad;
};

struct MapView {
  int4* hq;  // {height plane 0, claim word of even phases, height plane 1, claim word of odd phases}
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
  // one 64-bit word per phase parity: low half counts CTA arrivals, high half sums "drops still
  // active"; both monotonically increasing over the launch (zeroed by the host before it)
  unsigned long long word[2];
  unsigned max_steps;  // longest drop of this launch == number of phases that had a live drop
  unsigned abort;      // 1: peer mode, a peer did not show up in time; 2: more than kMaxPhases phases
  // peer mode: {phase tag | box-wide active sum << 32}, stored by the publishing CTA once every GPU has arrived
  unsigned long long release[2];
};

// Peer mode (one world spread over the GPUs of a box, one process per GPU): every rank maps the
// strips of all ranks (CUDA IPC over NVLink).  A drop stays with the rank that spawned it and
// reads / adds into whichever strip it is over; the per-phase barrier spans all GPUs, so the
// schedule -- and therefore the result -- is exactly the single-GPU one.
constexpr int kMaxPeers = 8;
struct PeerView {
  int4* hq[kMaxPeers];                  // strip r holds rows [r*rows, (r+1)*rows)
  CellRec* rec[kMaxPeers];
  unsigned long long* inbox[kMaxPeers]; // rank r's inbox: [parity][sender] words {phase+1 | sum << 32}
  int shift, mask;                      // rows per strip = 1 << shift
  int nranks, rank;
  unsigned tag_base;                    // (launch sequence number) << 20, identical on every rank
};

struct DescendArgs {
  MapView m;
  PeerView pv;
  StepParams P;
  shx_drop* drops;
  unsigned ndrops;
  unsigned align_age;  // != 0: drops with age > 0 wait for the phase equal to their age
  unsigned claim_epoch;  // 1..15, top bits of every claim key of this launch (stale keys of earlier launches lose)
  unsigned free_waits;   // waits per drop that do not cost it a step (shx_config.free_waits, <= 15)
  int* abort_flag;       // context flag word, OR-ed (never overwritten): 1 = a peer timed out, 2 = phase limit
  GridBar* bar;
  unsigned long long* stats;
  float* trace;  // 7 floats per Drop::descend call of drop 0, or null
  int trace_cap;
  int* trace_n;
};

__device__ __forceinline__ void stat_add(unsigned long long* stats, int i, unsigned long long v) {
  atomicAdd(stats + i, v);
}

#ifdef SHX_PHASE_TIMING
// development aid (tools/phase_timing.py, tools/arrival_hist.py): cycle stamps of thread 0 / block 0
// per phase section, per-warp barrier arrival times of one phase, experiment switches
__device__ unsigned long long g_phase_timing[8];
__device__ unsigned long long g_arrival[8192];
__device__ unsigned long long g_start[8192];
__device__ unsigned g_arrival_sm[8192];
__device__ unsigned g_exp;  // 1 no track atomics, 2 no height atomics, 4 no cascade, 8 poll back-off
#define SHX_EXP(bit) (g_exp & (bit))
#define SHX_T(i) do { if (gid == 0) tstamp[i] = clock64(); } while (0)
#else
#define SHX_T(i) do { } while (0)
#define SHX_EXP(bit) 0
#endif

// ---------------------------------------------------------------------------------------------
// Grid-wide barrier that also sums a per-CTA count (drops still active).  One 64-bit RED per CTA:
// arrival count in the low half, the count to be summed in the high half, on the word of this
// phase's parity; thread 0 polls that word, so the arrival count and the sum come from ONE load.
// Phase p+2 reuses the word of phase p, which is safe because nobody can arrive at barrier p+2
// before everybody has left barrier p.  prev_hi[] remembers the high half after the word's previous
// use (identical in every CTA).  The kernel is launched cooperatively (all CTAs co-resident).
// kAcquireLoad: the acquire side as one ld.acquire of the word after the relaxed polling instead of a fence (the same
// acquire pattern in the PTX memory model; measured -4 % per phase for the eight-lane kernel at 512 drops, nothing
// for the dense shapes, which keep the fence).
template <bool kAcquireLoad = false>
__device__ __forceinline__ unsigned grid_barrier_sum(GridBar* bar, unsigned phase, unsigned block_sum, unsigned* s_total,
                                                     unsigned& prev_hi0, unsigned& prev_hi1) {
  // The caller has just passed a __syncthreads-class barrier (block_sum comes from
  // __syncthreads_count), so every RED of this CTA for this phase was issued before thread 0's fence.
  if (gridDim.x == 1) return block_sum;
  if (threadIdx.x == 0) {
    unsigned long long* w = &bar->word[phase & 1u];
    asm volatile("fence.acq_rel.gpu;" ::: "memory");  // release: this CTA's REDs before the arrival
    atomicAdd(w, ((unsigned long long)block_sum << 32) | 1ull);
    const unsigned target = (phase / 2u + 1u) * gridDim.x;
    unsigned long long seen;
    do {  // relaxed polling (an acquire load per iteration would invalidate L1 every time)
      if (SHX_EXP(8)) __nanosleep(200);
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(w) : "memory");
    } while ((int)((unsigned)seen - target) < 0);
    // acquire: the other CTAs' REDs before our reads
    if (kAcquireLoad) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(w) : "memory");
    else asm volatile("fence.acq_rel.gpu;" ::: "memory");
    const unsigned hi = (unsigned)(seen >> 32);
    *s_total = hi - ((phase & 1u) ? prev_hi1 : prev_hi0);
    if (phase & 1u) prev_hi1 = hi;
    else prev_hi0 = hi;
  }
  __syncthreads();
  return *s_total;
}

// The same barrier across the GPUs of the box (peer mode).  Local CTAs arrive on the local word as
// above; block 0 waits for them, publishes {phase tag | local sum << 32} into every rank's inbox over
// NVLink, waits until all ranks' words in its own inbox carry this phase's tag, and releases the
// local CTAs with the box-wide sum through the release word.  Scopes: a CTA that sent REDs into a
// peer's strip during the phase arrives behind a system-scope fence (the others behind a
// device-scope one); block 0 fences at system scope before the inbox stores and after the inbox
// poll; the released CTAs acquire at device scope (causality order is transitive).  System-scope
// fences are expensive (measured: +5 ms per 8192^2 cycle when every CTA issues one per phase), which
// is why only block 0 and the CTAs with off-device REDs pay them.  A peer that does not show up
// within ~2 s aborts the launch (sum 0) instead of hanging the GPU.
__device__ __forceinline__ unsigned peer_barrier_sum(GridBar* bar, const PeerView& pv, unsigned phase, unsigned block_sum,
                                                     unsigned* s_total, unsigned* s_remote, unsigned& prev_hi0,
                                                     unsigned& prev_hi1) {
  if (threadIdx.x == 0) {
    const unsigned par = phase & 1u;
    unsigned long long* w = &bar->word[par];
    // release: system scope only if this CTA sent REDs over NVLink during the phase (the caller's
    // __syncthreads_count ordered every thread's s_remote store before this read)
    if (*s_remote) {
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      *s_remote = 0u;
    } else {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    atomicAdd(w, ((unsigned long long)block_sum << 32) | 1ull);
    const unsigned tag = pv.tag_base + phase + 1u;  // unique per launch and phase: inbox words are never reset
    const long long t0 = clock64();
    bool dead = false;
    if (blockIdx.x == 0) {
      const unsigned target = (phase / 2u + 1u) * gridDim.x;
      unsigned long long seen;
      do {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(w) : "memory");
      } while ((int)((unsigned)seen - target) < 0);
      const unsigned hi = (unsigned)(seen >> 32);
      const unsigned local = hi - (par ? prev_hi1 : prev_hi0);
      if (par) prev_hi1 = hi;
      else prev_hi0 = hi;
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      const unsigned long long msg = ((unsigned long long)local << 32) | tag;
      for (int r = 0; r < pv.nranks; r++) {
        unsigned long long* slot = pv.inbox[r] + par * kMaxPeers + pv.rank;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(msg) : "memory");
      }
      // poll all ranks' words of this GPU's inbox row together: one L2 round trip per sweep
      const unsigned long long* row = pv.inbox[pv.rank] + par * kMaxPeers;
      unsigned total;
      bool all;
      do {
        unsigned long long got[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; r++) {
          got[r] = tag;
          if (r < pv.nranks) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got[r]) : "l"(row + r) : "memory");
        }
        all = true;
        total = 0;
#pragma unroll
        for (int r = 0; r < kMaxPeers; r++) {
          all = all && (unsigned)got[r] == tag;
          total += r < pv.nranks ? (unsigned)(got[r] >> 32) : 0u;
        }
        if (!all && clock64() - t0 > 4000000000ll) dead = true;
      } while (!all && !dead);
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      if (dead) { total = 0; bar->abort = 1u; }  // the caller ORs it into the context's flag word
      const unsigned long long rel = ((unsigned long long)total << 32) | tag;
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&bar->release[par]), "l"(rel) : "memory");
      *s_total = total;
    } else {
      unsigned long long got;
      do {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(got) : "l"(&bar->release[par]) : "memory");
        if ((unsigned)got != tag && clock64() - t0 > 8000000000ll) { dead = true; got = tag; }
      } while ((unsigned)got != tag);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      *s_total = dead ? 0u : (unsigned)(got >> 32);
    }
  }
  __syncthreads();
  return *s_total;
}

// Start-of-call handshake of peer mode, stream-ordered between this rank's track reset / spawn and
// its descend launch: no rank may add into a strip whose owner has not finished resetting it.
// Uses the third inbox row; <<<1, 1>>>.
__global__ void peer_handshake_kernel(const PeerView pv, unsigned tag, unsigned* abort_flag) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  for (int r = 0; r < pv.nranks; r++) {
    unsigned long long* slot = pv.inbox[r] + 2 * kMaxPeers + pv.rank;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot), "l"((unsigned long long)tag) : "memory");
  }
  const long long t0 = clock64();
  for (int r = 0; r < pv.nranks; r++) {
    const unsigned long long* slot = pv.inbox[pv.rank] + 2 * kMaxPeers + r;
    unsigned long long got;
    do {
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(slot) : "memory");
      if ((unsigned)got != tag && clock64() - t0 > 20000000000ll) {  // ~10 s: ranks may reach the call at different times
        atomicOr(abort_flag, 1u);
        return;
      }
    } while ((unsigned)got != tag);
  }
  asm volatile("fence.acq_rel.sys;" ::: "memory");
}

// order-preserving float <-> uint32 maps (for non-NaN inputs; -0 must be canonicalised by the caller)
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned b = __float_as_uint(f);
  return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float(u ^ (((u >> 31) - 1u) | 0x80000000u));
}


// Turn-taking and crowd damping.  A phase reads frozen heights, so two drops standing on one cell would
// both erode it by the full amount, and drops on neighbouring cells (a train along a river) change
// coupled cells at once -- an explicit scheme whose per-cell factors add up.  Left alone this runs
// away (a 2048^2 map within 25 calls, 8192^2 with same-cell turn-taking only after ~100).  Two rules:
//  * of the drops on one cell only the holder of the highest key steps in a phase; the others wait,
//    and a phase spent waiting is a step of the drop's life not taken (bounded phases per launch);
//  * what a stepping drop moves (cascade transfers, sediment exchange) is halved for every one of the
//    eight cells around it that holds a higher key in this phase.
// Keys are claimed one phase ahead with atomicMax on the cell's claim word of the NEXT phase's parity.
// The claim words sit next to the height words of the cell (one 16-byte int4 {h0, claim0, h1, claim1}):
// the word a drop claims belongs to a sector its gather has just pulled into L2, the word it checks
// arrives with its centre height, and the neighbours' claims ride on the cooperative gather's 8-byte
// loads.  (In a plane of their own the claims cost +3.3 ms per 8192^2 cycle: every claim was a DRAM
// read-modify-write.)
//   key = {launch epoch : 4 | phase tag : 12 | phases waited so far, saturating : 3 | state hash : 13}
// i.e. longest waiting first, then an order-independent pseudo-random choice; equal keys all step.
// The epoch makes the keys of earlier launches lose; the host clears the words when it wraps.
constexpr int kWaitedShift = 16;       // SHX_DROP_WAITED_SHIFT
constexpr unsigned kMaxPhases = 4000;  // the tag has 12 bits
__device__ __forceinline__ unsigned claim_key(unsigned epoch, unsigned tag, const DropRegs& d) {
  unsigned h = __float_as_uint(d.px) * 0x9E3779B1u ^ __float_as_uint(d.py) * 0x85EBCA77u ^ (unsigned)d.age * 0xC2B2AE3Du;
  h ^= h >> 15;
  const unsigned waited = ((unsigned)d.flags >> kWaitedShift) & 7u;
  return (epoch << 28) | (tag << 16) | (waited << 13) | (h & 0x1FFFu);
}


constexpr int kFreeWaitShift = 19;  // SHX_DROP_FREEW_SHIFT

// A waiting phase (turn-taking): the wait counter of the key grows; the first free_waits waits of a drop's life are
// free, every later one costs a step of its life (so a launch is bounded by maxAge + 2 + free_waits phases).
// Returns true if the drop expired in the queue (water.h:74-77).
__device__ __forceinline__ bool wait_one_phase(DropRegs& d, const DescendArgs& a) {
  const unsigned waited = ((unsigned)d.flags >> kWaitedShift) & 7u;
  d.flags = (d.flags & ~(7 << kWaitedShift)) | (int)((waited < 7u ? waited + 1u : 7u) << kWaitedShift);
  const unsigned fw = ((unsigned)d.flags >> kFreeWaitShift) & 15u;
  if (fw < a.free_waits) {
    d.flags = (d.flags & ~(15 << kFreeWaitShift)) | (int)((fw + 1u) << kFreeWaitShift);
    return false;
  }
  d.age++;
  return (float)d.age > a.P.maxAge;
}

// L2 eviction-priority hints on the REDs (round 2; measured at 8192^2: 11.4 -> 8.9 ms per cycle).  A drop's next 3x3
// block overlaps its current one and every height add is repeated on the other plane one phase later (catch-up), so
// the sector a height RED or a claim lands in is wanted again within a phase or two: those carry evict_last.  A cell
// record is read and added to in ONE phase and not again until another drop passes: the track REDs carry evict_first.
// Without the hints ~38 MB stream through L2 per phase and the sectors are gone when they are needed again.  The same
// hints on the LOADS do not help (height gathers evict_last: +1 %, record loads evict_first: +6 %); profiles/
// r2_l2_hints.txt holds the sweep.
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void red_height(int* p, int v) {
  asm volatile("red.global.add.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(l2_policy_keep()) : "memory");
}
__device__ __forceinline__ void red_claim(unsigned* p, unsigned v) {
  asm volatile("red.global.max.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(l2_policy_keep()) : "memory");
}
__device__ __forceinline__ void red_track32(int* p, int v) {
  asm volatile("red.global.add.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(l2_policy_stream()) : "memory");
}
__device__ __forceinline__ void red_track64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.global.add.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(l2_policy_stream()) : "memory");
}

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
// Work of one phase for one drop (see shx_step.cuh for the split of Drop::descend):
//   [World::cascade owed by the previous call]   move_math   track REDs   exchange_math
//
// kCoop selects how the 3x3 block is gathered.  With one load per lane and cell, every lane of a
// load instruction touches its own 128-byte line and the SM's L1TEX pipe replays the instruction
// once per line (~2 cycles each): at 28 resident warps per SM that replay time, not DRAM or L2
// bandwidth, bounds the phase.  The cooperative gather lets eight lanes fetch the eight neighbour
// cells of ONE drop (four drops per instruction) while the owner fetches its centre cell -- heights
// and claim words -- with one 16-byte load: three lines per drop instead of nine lane-wavefronts;
// the values travel through shared memory to the lane that owns the drop.
//
// Shared memory per thread: s_B[9] block heights (thread-major, stride 9: conflict-free both for
// the owner's sequential reads and for the cooperative stores), s_D[2][8] neighbour deltas of this
// / the previous phase, s_S[8] (height, index) pairs in cascade order.
template <int kMaxThreads, int kMinBlocks, bool kCoop, bool kPeer = false>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) descend_lockstep_kernel(const __grid_constant__ DescendArgs a) {
  static_assert(kCoop || !kPeer, "peer mode uses the cooperative gather");
  extern __shared__ int32_t s_mem[];
  __shared__ unsigned s_total;
  __shared__ unsigned s_remote;  // peer mode: this CTA added into another GPU's strip during the phase
  const int tid = threadIdx.x, nt = blockDim.x;
  if (kPeer) {
    if (tid == 0) s_remote = 0u;
    __syncthreads();
  }
  int32_t* s_B = s_mem + tid * 9;                               // s_B[k]
  int32_t* s_D = s_mem + 9 * nt + tid;                          // s_D[(buf*8 + j)*nt]
  uint2* s_S = reinterpret_cast<uint2*>(s_mem + 25 * nt) + tid; // s_S[r*nt]
  // cooperative gather: lane -> (which of 4 drops of a group, which of its 8 neighbour cells); the
  // centre cell comes with the owner's own 16-byte load of {heights, claim words}
  const int lane = tid & 31;
  const int co_sub = lane >> 3, co_k = (lane & 7) + ((lane & 7) >> 2);  // block index 0..8 without 4
  const int co_dx = co_k / 3 - 1, co_dy = co_k % 3 - 1;
  const int co_off = co_dx * a.m.size + co_dy;
  int32_t* const s_Bw = s_mem + (tid - lane) * 9;               // first drop of this warp
  const unsigned gid = blockIdx.x * nt + tid;
  const int size = a.m.size;
  int* const H = reinterpret_cast<int*>(a.m.hq);
  // cell (x, y) -> its pair of height words / its record.  Peer mode: the strip x >> shift owns it.
  auto h_at = [&](int x, int y) -> int* {
    if (kPeer) return reinterpret_cast<int*>(a.pv.hq[x >> a.pv.shift]) + 4 * ((x & a.pv.mask) * size + y);
    return H + 4 * ((x - a.m.xlo) * size + y);
  };
  auto rec_at = [&](int x, int y) -> CellRec* {
    if (kPeer) return a.pv.rec[x >> a.pv.shift] + ((x & a.pv.mask) * size + y);
    return a.m.rec + ((x - a.m.xlo) * size + y);
  };
  auto claim_at = [&](int x, int y) -> unsigned* { return reinterpret_cast<unsigned*>(h_at(x, y)) + 1; };
  auto claim_max = [&](unsigned* p, unsigned key, int x) {
    if (kPeer && (x >> a.pv.shift) != a.pv.rank) { atomicMax_system(p, key); s_remote = 1u; }
    else red_claim(p, key);
  };
  // Integer adds.  Peer mode: every add into a strip is performed by the L2 of the GPU that holds it;
  // the owner uses plain device-scope REDs, the others system-scope ones over NVLink, and a CTA
  // that sent anything off-device raises s_remote so that its arrival fence is system scope.
  auto add32 = [&](int* p, int v, int x) {  // heights
    if (kPeer && (x >> a.pv.shift) != a.pv.rank) { atomicAdd_system(p, v); s_remote = 1u; }
    else red_height(p, v);
  };
  auto track32 = [&](int* p, int v, int x) {
    if (kPeer && (x >> a.pv.shift) != a.pv.rank) { atomicAdd_system(p, v); s_remote = 1u; }
    else red_track32(p, v);
  };
  auto add64 = [&](unsigned long long* p, unsigned long long v, int x) {  // tracks
    if (kPeer && (x >> a.pv.shift) != a.pv.rank) { atomicAdd_system(p, v); s_remote = 1u; }
    else red_track64(p, v);
  };

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
  // align_age: a drop carried over from the previous call sleeps until the phase that equals its age,
  // i.e. it meets the drops of this call at the same age at which it met those of its own call
  bool asleep = a.align_age != 0u && alive && d.age > 0;
  alive = alive && !asleep;
  // deltas of the previous phase, still owed to the other plane
  int dC_prev = 0;          // centre
  unsigned dmask_prev = 0;  // neighbours (values in s_D)
  int pix = 0, piy = 0;     // its centre cell
  unsigned bar_hi0 = 0, bar_hi1 = 0;  // grid barrier bookkeeping (thread 0)
  unsigned steps = 0, transfers = 0;
  long long fx_eroded = 0, fx_inflation = 0;
  int tn = 0;

#ifdef SHX_PHASE_TIMING
  long long tstamp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
#define SHX_TRACE_ROW()                                                                                   \
  do {                                                                                                    \
    if (a.trace != nullptr && gid == 0 && tn < a.trace_cap) {                                             \
      float* t__ = a.trace + 7 * (size_t)tn++;                                                            \
      t__[0] = (float)d.age; t__[1] = d.px; t__[2] = d.py; t__[3] = d.sx; t__[4] = d.sy; t__[5] = d.vol; \
      t__[6] = d.sed;                                                                                     \
    }                                                                                                     \
  } while (0)

  // every drop that is awake in phase 0 claims its cell (tag 1, parity 0); barrier number 0
  if (alive) claim_max(claim_at((int)d.px, (int)d.py), claim_key(a.claim_epoch, 1u, d), (int)d.px);
  {
    const unsigned block_sum = (unsigned)__syncthreads_count(alive || asleep);
    if (kPeer) peer_barrier_sum(a.bar, a.pv, 0u, block_sum, &s_total, &s_remote, bar_hi0, bar_hi1);
    else grid_barrier_sum(a.bar, 0u, block_sum, &s_total, bar_hi0, bar_hi1);
  }

  for (unsigned phase = 0;; ++phase) {
    const int rpar = (int)(phase & 1u), wpar = rpar ^ 1;
    const int rw = 2 * rpar, ww = 2 * wpar;  // word of the read / write plane inside a cell {h0, claim0, h1, claim1}
    const int cur = rpar * 8, prev = wpar * 8;
    if (asleep && (unsigned)d.age <= phase) {
      asleep = false;
      alive = true;
    }
    SHX_T(0);
#ifdef SHX_PHASE_TIMING
    if (phase == 100u && (tid & 31) == 0 && (gid >> 5) < 8192u) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      g_start[gid >> 5] = ns;
    }
#endif

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
    unsigned held = 0u;
    int crowded = 0;  // how many of the eight cells around hold a higher key this phase
    if (kCoop) {
      // 8 groups of 4 drops; lane (sub, k) loads neighbour cell k of drop 4g+sub of this warp -- its height
      // and its claim word of this phase, one 8-byte load -- the height into that drop's s_B row; a
      // claim above the drop's own key blocks the drop (one ballot per group tells the owners)
      const unsigned meta = alive ? (inb | 0x200u) : 0u;
      const unsigned mykey = claim_key(a.claim_epoch, phase + 1u, d);
      if (__any_sync(0xffffffffu, alive)) {
        int2 got[8];
#pragma unroll
        for (int g = 0; g < 8; g++) {
          const int src = g * 4 + co_sub;
          const unsigned m = __shfl_sync(0xffffffffu, meta, src);
          const bool ok = ((m >> co_k) & 1u) && (m & 0x200u);
          got[g] = make_int2(0, 0);
          if (kPeer) {
            const int sx = __shfl_sync(0xffffffffu, ix, src), sy = __shfl_sync(0xffffffffu, iy, src);
            if (ok) got[g] = __ldcg(reinterpret_cast<const int2*>(h_at(sx + co_dx, sy + co_dy) + rw));  // coordinates are only valid when ok
          } else {
            const int c = __shfl_sync(0xffffffffu, cidx, src);
            if (ok) got[g] = __ldcg(reinterpret_cast<const int2*>(H + 4 * (c + co_off) + rw));
          }
        }
        if (alive) {
          const int4 cc = __ldcg(reinterpret_cast<const int4*>(h_at(ix, iy)));  // centre heights + claim words
          s_B[4] = rpar ? cc.z : cc.x;
          held = (unsigned)(rpar ? cc.w : cc.y);
          fld = __ldg(reinterpret_cast<const float4*>(rec_at(ix, iy)));
        }
#pragma unroll
        for (int g = 0; g < 8; g++) {
          const int src = g * 4 + co_sub;
          s_Bw[src * 9 + co_k] = got[g].x;
          const unsigned key_src = __shfl_sync(0xffffffffu, mykey, src);
          const unsigned bal = __ballot_sync(0xffffffffu, (unsigned)got[g].y > key_src);
          if ((lane >> 2) == g) crowded = __popc((bal >> (8 * (lane & 3))) & 0xFFu);
        }
      }
    } else if (alive) {
      const int* c = H + 4 * cidx + rw;
      const unsigned mykey = claim_key(a.claim_epoch, phase + 1u, d);
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const int off = (k / 3 - 1) * size + (k % 3 - 1);
        v[k] = ((inb >> k) & 1u) ? __ldcg(c + 4 * off) : 0;
        if (k != 4 && ((inb >> k) & 1u)) crowded += (unsigned)__ldcg(c + 4 * off + 1) > mykey ? 1 : 0;
      }
      fld = __ldg(reinterpret_cast<const float4*>(a.m.rec + cidx));
      held = __ldcg(claim_at(ix, iy) + rw);
    }

    // whose turn is it on this cell?  (claimed during the previous phase, complete since its barrier)
    const bool turn = alive && held == claim_key(a.claim_epoch, phase + 1u, d);

    if (dC_prev | (int)dmask_prev) {  // catch-up of the previous phase's deltas
      if (dC_prev && !SHX_EXP(2)) add32(h_at(pix, piy) + ww, dC_prev, pix);
      unsigned m = dmask_prev;
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1u;
        const int k = j + (j >> 2);
        const int kx = (k * 11) >> 5;
        if (!SHX_EXP(2)) add32(h_at(pix + kx - 1, piy + (k - 3 * kx) - 1) + ww, s_D[(prev + j) * nt], pix + kx - 1);
      }
      dC_prev = 0;
      dmask_prev = 0;
    }

    if (kCoop) __syncwarp();  // the block rows written by the other lanes of the warp are complete

    if (alive && !turn) {  // a drop with a higher key has the cell: wait (the first free_waits waits are free, later ones cost a step)
      if (wait_one_phase(d, a)) {  // expired in the queue: the sediment stays where the drop stands (water.h:74-77)
        const int q = h_quantize(d.sed);
        if (q && !SHX_EXP(2)) add32(h_at(ix, iy) + ww, q, ix);
        dC_prev = q;  // the other plane gets it in the next phase, like any other delta
        pix = ix;
        piy = iy;
        alive = false;
        atomicMax(&a.bar->max_steps, phase + 1u);
        stat_add(a.stats, ST_TERM_AGE, 1ull);
        stat_add(a.stats, ST_FX_DEPOSITED, (unsigned long long)(long long)q);
        stat_add(a.stats, ST_FX_SED_DEPOSITED, (unsigned long long)l_quantize(d.sed));
        d.flags = SHX_DROP_DONE_AGE;
      } else {
        claim_max(claim_at(ix, iy) + ww, claim_key(a.claim_epoch, phase + 2u, d), ix);
      }
    }
    if (asleep && (unsigned)d.age == phase + 1u)  // wakes up in the next phase
      claim_max(claim_at((int)d.px, (int)d.py) + ww, claim_key(a.claim_epoch, phase + 2u, d), (int)d.px);

    if (turn) {
      // 1, or 2^-n next to n cells that hold a higher key: what this drop moves (cascade transfers and the
      // sediment exchange) is scaled down, so that the changes of neighbouring cells in one phase do not add up
      const float damp = __int_as_float((127 - crowded) << 23);
      steps++;
      d.flags &= ~(7 << kWaitedShift);
#pragma unroll
      for (int k = 0; k < 9; k++) {
        if (kCoop) v[k] = s_B[k];
        else s_B[k] = v[k];
      }
      int Bc = v[4];
      unsigned dmask = 0;
      SHX_T(1);

      if ((d.flags & SHX_DROP_CASCADE) && !SHX_EXP(4)) {  // World::cascade of the previous call, world.h:90-168
        d.flags &= ~SHX_DROP_CASCADE;
        const float hc0 = h_to_float(Bc);
        float hh[8];
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
          hh[j] = in ? h : __int_as_float(0x7f800000);  // missing cells sort last
        }
        // The centre only changes through a transfer: if nothing exceeds its allowance against the
        // untouched centre, nothing fires at all.
        if (any) {
          // world.h:129-131 ascending by height; libstdc++ sorts <= 16 elements by insertion, i.e.
          // stably: for i < j, i comes first iff h_i <= h_j.  rank = position in that order.
          int rank[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = i + 1; j < 8; j++) {
              const bool le = hh[i] <= hh[j];
              rank[j] += le ? 1 : 0;
              rank[i] += le ? 0 : 1;
            }
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const unsigned k = (unsigned)(j + (j >> 2));
            s_S[rank[j] * nt] = make_uint2(__float_as_uint(hh[j]), ((inb >> k) & 1u) ? (unsigned)j : 8u);
          }
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const uint2 e = s_S[r * nt];
            const unsigned j = e.y;
            if (j < 8u) {
              const float hn = __uint_as_float(e.x);
              const float diff = h_to_float(Bc) - hn;  // world.h:138: centre re-read, neighbour snapshot
              const bool diag = (0xA5u >> j) & 1u;
              const float lim = above_tenth(hn) ? (diag ? a.P.lim_diag : a.P.lim_axis) : 0.0f;
              const float excess = fabsf(diff) - lim;
              if (diff != 0.0f && excess > 0.0f) {
                const int t = h_quantize((a.P.settling * damp) * excess / 2.0f);  // world.h:154
                const int s = diff > 0.0f ? t : -t;                       // world.h:157-164
                Bc -= s;
                const int k = (int)(j + (j >> 2));
                const int kx = (k * 11) >> 5;
                s_B[k] += s;
                s_D[(cur + (int)j) * nt] = s;
                dmask |= 1u << j;
                if (!SHX_EXP(2)) add32(h_at(ix + kx - 1, iy + (k - 3 * kx) - 1) + ww, s, ix + kx - 1);
                transfers++;
              }
            }
          }
          s_B[4] = Bc;
        }
      }
      SHX_T(2);

      const float hc = h_to_float(Bc);
      const float hxm = xm ? h_to_float(s_B[1]) : 0.0f, hxp = xp ? h_to_float(s_B[7]) : 0.0f;
      const float hym = ym ? h_to_float(s_B[3]) : 0.0f, hyp = yp ? h_to_float(s_B[5]) : 0.0f;
      const MoveResult mv = move_math(hc, hxm, hxp, hym, hyp, inb, d, fld, a.P, size);
      int dC = Bc - v[4];
      SHX_T(3);
      if (!mv.moved) {  // water.h:74-82: aged out / dried up, the sediment stays here
        const int q = h_quantize(mv.dheight);
        dC += q;
        alive = false;
        atomicMax(&a.bar->max_steps, phase + 1u);
        stat_add(a.stats, (d.flags & SHX_DROP_DONE_AGE) ? ST_TERM_AGE : ST_TERM_VOL, 1ull);
        stat_add(a.stats, ST_FX_DEPOSITED, (unsigned long long)(long long)q);
        stat_add(a.stats, ST_FX_SED_DEPOSITED, (unsigned long long)l_quantize(d.sed));
        SHX_TRACE_ROW();
      } else {
        // water.h:124: the new cell (nearest, truncated) may lie outside the block: issue that one
        // dependent gather first and do the erf and the track adds while it is in flight
        const int nix = (int)d.px, niy = (int)d.py;
        int hv = 0;
        if (!mv.oob) {
          const int ddx = nix - ix, ddy = niy - iy;
          if (ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1) hv = s_B[(ddx + 1) * 3 + (ddy + 1)];
          else hv = __ldcg(h_at(nix, niy) + rw);
        }
        const float cap = 1.0f + a.P.entrainment * shx_erff(0.4f * fld.x);  // water.h:127, cellpool.h:242-244
        if (!SHX_EXP(1)) {  // water.h:115-117.  discharge (>= 0, low word) and momentum-x (high word)
          // go out as ONE 64-bit add: the low word cannot carry while the Q13.18 range check holds
          CellRec* rec = rec_at(ix, iy);
          const unsigned long long packed = ((unsigned long long)(unsigned)t_quantize(mv.t_mx) << 32) |
                                            (unsigned long long)(unsigned)t_quantize(mv.t_d);
          add64(reinterpret_cast<unsigned long long*>(&rec->track_d), packed, ix);
          track32(&rec->track_my, t_quantize(mv.t_my), ix);
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
          atomicMax(&a.bar->max_steps, phase + 1u);
          stat_add(a.stats, ST_TERM_OOB, 1ull);
          stat_add(a.stats, ST_FX_SED_OOB, (unsigned long long)l_quantize(d.sed));
          d.vol = 0.0f;
          d.flags = SHX_DROP_DONE_OOB;
        } else {
          d.age++;                      // water.h:153
          d.flags |= SHX_DROP_CASCADE;  // water.h:151, executed at the start of the next phase
          if (kPeer || (nix >= a.m.row0 && nix < a.m.row1)) claim_max(claim_at(nix, niy) + ww, claim_key(a.claim_epoch, phase + 2u, d), nix);
          if (!kPeer && (nix < a.m.row0 || nix >= a.m.row1)) {  // left the strip: hand over (cascade still owed)
            const bool tolo = nix < a.m.row0;
            d.flags = (d.flags & ~SHX_DROP_ALIVE) | (tolo ? SHX_DROP_MIGRATE_LO : SHX_DROP_MIGRATE_HI);
            alive = false;
            atomicMax(&a.bar->max_steps, phase + 1u);
            stat_add(a.stats, tolo ? ST_MIGRATED_LO : ST_MIGRATED_HI, 1ull);
          }
        }
        SHX_TRACE_ROW();
      }
      if (dC && !SHX_EXP(2)) add32(h_at(ix, iy) + ww, dC, ix);
      dC_prev = dC;
      dmask_prev = dmask;
      pix = ix;
      piy = iy;
    }

    SHX_T(4);
#ifdef SHX_PHASE_TIMING
    if (phase == 100u && (tid & 31) == 0) {  // per-warp arrival times of one phase (ns, global timer)
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const unsigned w = gid >> 5;
      if (w < 8192u) { g_arrival[w] = ns; g_arrival_sm[w] = smid; }
    }
#endif
    const unsigned block_sum = (unsigned)__syncthreads_count(alive || asleep || (dC_prev | (int)dmask_prev));
    SHX_T(5);
    const unsigned total = kPeer ? peer_barrier_sum(a.bar, a.pv, phase + 1u, block_sum, &s_total, &s_remote, bar_hi0, bar_hi1)
                                 : grid_barrier_sum(a.bar, phase + 1u, block_sum, &s_total, bar_hi0, bar_hi1);
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
    if (phase + 2u >= kMaxPhases) {  // the claim tag would wrap: give up (the host reports SHX_ERR_RANGE)
      if (gid == 0) atomicOr(a.abort_flag, 2);
      break;
    }
  }
#undef SHX_TRACE_ROW

  if (gid < a.ndrops) {
    float4 lo, hi;
    lo.x = d.px; lo.y = d.py; lo.z = d.sx; lo.w = d.sy;
    hi.x = d.vol; hi.y = d.sed; hi.z = __int_as_float(d.age); hi.w = __int_as_float(d.flags);
    reinterpret_cast<float4*>(a.drops + gid)[0] = lo;
    reinterpret_cast<float4*>(a.drops + gid)[1] = hi;
  }
  if (a.trace_n != nullptr && gid == 0) *a.trace_n = tn;
  if (kPeer && gid == 0 && a.bar->abort) atomicOr(a.abort_flag, (int)a.bar->abort);

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
    if (d.flags & SHX_DROP_CHECK_SPAWN) {  // world.h:71-74: the rejection sees what the earlier drops of the call left
      d.flags &= ~SHX_DROP_CHECK_SPAWN;
      const int sx = (int)d.px, sy = (int)d.py;
      const bool oob = !(d.px > -1.0f) || !(d.py > -1.0f) || sx >= size || sy >= size;
      const float h = oob ? 0.0f : H[4 * (sx * size + sy)];  // map.height() of a missing cell is 0 (cellpool.h:433-437)
      if (!above_tenth(h)) {
        d.flags = SHX_DROP_REJECTED;
        a.stats[ST_REJECTED] += 1ull;
      } else {
        a.stats[ST_SPAWNED] += 1ull;
      }
    }
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
        B[k] = in ? H[4 * (cidx + (k / 3 - 1) * size + (k % 3 - 1))] : 0.0f;
      }
      CellRec* rec = a.m.rec + cidx;
      const float4 fld = *reinterpret_cast<const float4*>(rec);
      steps++;
      if (d.flags & SHX_DROP_CASCADE) {  // water.h:151 of the previous call
        transfers += cascade_block_f32(B, inb, a.P);
        d.flags &= ~SHX_DROP_CASCADE;
      }
      const MoveResult mv = move_math(B[4], (inb & 2u) ? B[1] : 0.0f, (inb & 128u) ? B[7] : 0.0f, (inb & 8u) ? B[3] : 0.0f,
                                      (inb & 32u) ? B[5] : 0.0f, inb, d, fld, a.P, size);
      if (!mv.moved) {
        B[4] = B[4] + mv.dheight;  // water.h:75,80
      } else {
        float* t = reinterpret_cast<float*>(&rec->track_d);  // water.h:115-117
        t[0] += mv.t_d; t[1] += mv.t_mx; t[2] += mv.t_my;
        float h2;
        if (mv.oob) {
          h2 = oob_h2(B[4]);  // water.h:121-122
        } else {
          const int nix = (int)d.px, niy = (int)d.py;  // water.h:124
          const int ddx = nix - ix, ddy = niy - iy;
          if (ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1) {
            h2 = B[0];
#pragma unroll
            for (int k = 1; k < 9; k++) h2 = (k == (ddx + 1) * 3 + (ddy + 1)) ? B[k] : h2;
          } else {
            h2 = H[4 * (nix * size + niy)];
          }
        }
        float carried;
        const float cap = 1.0f + a.P.entrainment * shx_erff(0.4f * fld.x);  // water.h:127, cellpool.h:242-244
        B[4] = B[4] + exchange_math(B[4], h2, cap, mv.effD, d, a.P, carried);  // water.h:127-136
        fx_inflation += l_quantize(d.sed) - l_quantize(carried);
        if (mv.oob) {  // water.h:139-142
          d.vol = 0.0f;
          d.flags = SHX_DROP_DONE_OOB;
        } else {
          d.age++;                      // water.h:153
          d.flags |= SHX_DROP_CASCADE;  // water.h:151
        }
      }
#pragma unroll
      for (int k = 0; k < 9; k++)
        if (inb & (1u << k)) H[4 * (cidx + (k / 3 - 1) * size + (k % 3 - 1))] = B[k];
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

}  // namespace shx

#include "shx_descend.cuh"
#include "shx_aux_kernels.cuh"
