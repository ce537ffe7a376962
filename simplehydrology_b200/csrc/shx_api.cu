// shx C ABI (include/shx.h): context, transfers and kernel launches.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -prec-div=true -prec-sqrt=true
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "shx_kernels.cuh"
#include "shx_view_kernels.cuh"

struct LaunchShape {
  const void* kernel = nullptr;
  int block = 0;
  int cap_blocks = 0;  // co-resident CTAs (cooperative launch limit)
  int lanes = 1;       // lanes per drop: 8 = descend_group_kernel, 1 = one thread per drop (descend_lockstep_kernel)
  size_t smem(int b) const;
  size_t drops_per_block(int b) const { return (size_t)b / (size_t)lanes; }
};

constexpr size_t kDropsPerLaunch = 131072;  // drops that march together in one launch (device-independent split)
constexpr size_t kShapeLimit[3] = {6144, 24576, 49152};  // largest batch of launch shapes [0], [1], [2] (see shx_create)

struct TimingSpan {
  int kind;  // 0 spawn, 1 descend, 2 ema, 3 pack (download), 4 device-to-host copy, 5 rootdensity push
  cudaEvent_t e0, e1;
};

using namespace shx;

// one thread per drop: s_B[9] + s_D[2][8] + s_S[8] x 2 words per thread; eight lanes per drop: 3 words per lane
size_t LaunchShape::smem(int b) const { return (size_t)(lanes > 1 ? kGroupSmemWords : 41) * sizeof(int32_t) * b; }

static thread_local std::string g_err;

static int fail(int code, const char* what) {
  g_err = what ? what : "";
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      char b__[512];                                                                          \
      snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return fail(SHX_ERR_CUDA, b__);                                                         \
    }                                                                                         \
  } while (0)

struct shx_ctx {
  shx_params p;
  shx_config cfg;
  int size = 0;
  MapView m{};
  size_t stored_cells = 0, owned_cells = 0;
  float* d_view = nullptr;  // staging for shx_vertex_download / shx_view_maps_download (allocated on first use)
  size_t view_bytes = 0;
  int halo_lo = 0, halo_hi = 0;  // halo rows actually present on each side
  shx_drop* d_drops = nullptr;
  size_t max_drops = 0;
  float* d_xy = nullptr;
  GridBar* d_bar = nullptr;
  unsigned long long* d_stats = nullptr;
  unsigned long long* h_stats = nullptr;  // pinned
  int* d_flags = nullptr;                 // [0] range error, [1] trace_n
  int* h_flags = nullptr;                 // pinned
  float* d_trace = nullptr;
  int trace_cap = 1024;
  unsigned* d_u32 = nullptr;  // [0..1] min/max, [2..3] migrant counts
  int32_t* d_halo_ref[2] = {nullptr, nullptr};
  shx_cell* d_stage = nullptr;
  shx_cell* d_stage2 = nullptr;      // second tile staging buffer: the pack of tile t+1 overlaps the DMA of tile t
  cudaStream_t copy_stream = nullptr;  // device-to-host copies of shx_download
  cudaEvent_t ev_packed[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;  // legacy default stream unless set
  int sm_count = 0;
  uint64_t epoch = 0;
  uint64_t launches = 0;
  size_t last_n = 0;        // drops of the last run still sitting in d_drops
  bool tracks_clean = true; // all track accumulators are zero (world.h:56-61 already satisfied)
  LaunchShape shape[4];     // multi-CTA launch shapes: [0] eight lanes per drop, [1] CTAs of 64, [2] CTAs of 128, [3] dense (2 x 448 per SM)
  bool forced_shape = false;
  // peer mode
  bool peer = false, peer_attached = false;
  PeerView pv{};
  unsigned long long* d_inbox = nullptr;
  unsigned peer_seq = 0;
  unsigned claim_epoch = 0;  // launch counter mod 15 (+1), see next_claim_epoch
  void* peer_opened[3 * kMaxPeers] = {};
  bool timing = false;
  std::vector<TimingSpan> spans;
  bool strip_open = false;  // between shx_strip_erode_begin and _end
  // sparse pushes (Plant::root edits): persistent device list + pinned bounce buffer, so that a push is two async
  // copies and a launch (round 1 allocated and freed with the stream-ordered allocator and synchronised per call:
  // with the pool's default release threshold that was a real cudaMalloc/cudaFree pair every frame)
  char* d_push = nullptr;
  char* h_push = nullptr;  // pinned
  size_t push_cap = 0;
  cudaEvent_t push_done = nullptr;  // the bounce buffer may be overwritten once this has fired
  int last_grid = 0, last_block = 0, last_lanes = 0;  // shape of the last descend launch (shx_launch_info)
  cudaAccessPolicyWindow l2_window{};  // persisting-L2 window over the height / claim words (maps that fit)
  bool l2_window_on = false;
  // N3: vegetation on the device (shx_veg_*): two plant lists (compaction goes from one to the other), per-slot scratch
  bool veg = false;
  PlantParams plant{};
  size_t veg_cap = 0, veg_n = 0;
  int veg_cur = 0;
  int2* d_plant_pos[2] = {nullptr, nullptr};
  float* d_plant_size[2] = {nullptr, nullptr};
  unsigned* d_veg_flags = nullptr;
  int2* d_veg_child = nullptr;
  float* d_veg_grown = nullptr;
  uint2* d_veg_blocks = nullptr;
  unsigned* d_veg_totals = nullptr;
  unsigned* h_veg_totals = nullptr;  // pinned
  bool root_counts_valid = false;    // the integer root counts mirror the fp32 rootdensity of every cell
  // shx_download_compact: 16 bytes per owned cell on the device, a ring of pinned chunks on the host
  float4* d_compact = nullptr;
  char* h_ring = nullptr;  // pinned, kRingSlots chunks
  cudaEvent_t ev_ring[8] = {};
  double scatter_ms = 0.0;  // host time spent in the scatter since the last shx_timing_read
};

static StepParams step_params(const shx_params& p) {
  StepParams s;
  s.maxAge = p.maxAge; s.minVol = p.minVol; s.evapRate = p.evapRate; s.depositionRate = p.depositionRate;
  s.entrainment = p.entrainment; s.gravity = p.gravity; s.momentumTransfer = p.momentumTransfer;
  s.maxdiff = p.maxdiff; s.settling = p.settling; s.lod = (float)p.lodsize; s.mapscale = (float)p.mapscale;
  s.lrate = p.lrate;
  // world.h:123,145: length(vec2(nn)) * maxdiff * lodsize, fp32 left to right
  s.lim_axis = 1.0f * p.maxdiff * (float)p.lodsize;
  s.lim_diag = sqrtf(2.0f) * p.maxdiff * (float)p.lodsize;
  s.keep = 1.0 - (double)p.evapRate;  // water.h:135-136
  s.inv_keep = 1.0 / s.keep;
  return s;
}

static bool sequential(const shx_ctx* c) { return c->cfg.mode == SHX_MODE_SEQUENTIAL; }

static int grid_for(const shx_ctx* c, size_t n, int block = 256) {
  const size_t want = (n + block - 1) / block;
  const size_t cap = (size_t)c->sm_count * 8;
  return (int)std::max<size_t>(1, std::min(want, cap));
}


// Instantiations {max CTA threads, min CTAs/SM}.  KERNEL_SMALL: the eight-lane kernel as ONE CTA (the barrier is a
// plain __syncthreads): batches of up to 16 drops, and the 1024-thread shape the tests force.  Multi-CTA: the first four cap registers at
// 64 (1024 threads per SM); variant 1 allows 128 registers (512 threads per SM); variant 2 is
// 7 CTAs of 128 threads per SM = 896 threads at 72 registers, just enough for the 886 drops per SM
// of an 8192^2 cycle.
// The last template argument selects the warp-cooperative block gather (see shx_kernels.cuh).
#define KERNEL_SMALL descend_group_kernel<1024, 1>
#define PICK(T, B) (coop ? (const void*)descend_lockstep_kernel<T, B, true> : (const void*)descend_lockstep_kernel<T, B, false>)
static const void* big_kernel(int block, int variant, bool coop) {
  if (variant == 2) return PICK(128, 7);
  if (variant == 3) return PICK(448, 2);  // same 896 threads/SM, 296 CTAs
  if (variant == 1) return block <= 256 ? PICK(256, 2) : PICK(512, 1);
  if (block <= 128) return PICK(128, 8);
  if (block <= 256) return PICK(256, 4);
  if (block <= 512) return PICK(512, 2);
  return PICK(1024, 1);
}
#undef PICK

extern "C" {

int shx_version(void) { return SHX_VERSION; }
const char* shx_last_error(void) { return g_err.c_str(); }

void shx_default_params(shx_params* p, int mapsize) {
  if (!p) return;
  p->evapRate = 0.001f;        // water.h:43
  p->depositionRate = 0.1f;    // water.h:44
  p->minVol = 0.01f;           // water.h:45
  p->maxAge = 500.0f;          // water.h:46
  p->entrainment = 10.0f;      // water.h:48
  p->gravity = 1.0f;           // water.h:49
  p->momentumTransfer = 1.0f;  // water.h:50
  p->lrate = 0.1f;             // world.h:42
  p->maxdiff = 0.01f;          // world.h:43
  p->settling = 0.8f;          // world.h:44
  p->mapscale = 80;            // cellpool.h:165
  p->tilesize = 512;           // cellpool.h:167
  p->mapsize = mapsize;        // cellpool.h:171
  p->lodsize = 1;              // cellpool.h:178
}

void shx_default_config(shx_config* c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->mode = SHX_MODE_BATCHED;
  c->halo = 2;
  c->free_waits = 8;
}

void shx_destroy(shx_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaFree(c->m.hq); cudaFree(c->m.rec);
  cudaFree(c->d_drops); cudaFree(c->d_xy); cudaFree(c->d_bar); cudaFree(c->d_stats); cudaFree(c->d_flags);
  cudaFree(c->d_trace); cudaFree(c->d_u32); cudaFree(c->d_halo_ref[0]); cudaFree(c->d_halo_ref[1]);
  cudaFree(c->d_stage);
  cudaFree(c->d_stage2);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (int i = 0; i < 2; i++) {
    if (c->ev_packed[i]) cudaEventDestroy(c->ev_packed[i]);
    if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
  }
  for (void* p : c->peer_opened)
    if (p) cudaIpcCloseMemHandle(p);
  cudaFree(c->d_inbox);
  cudaFree(c->d_view);
  cudaFree(c->d_push);
  if (c->h_push) cudaFreeHost(c->h_push);
  if (c->push_done) cudaEventDestroy(c->push_done);
  if (c->h_stats) cudaFreeHost(c->h_stats);
  if (c->h_flags) cudaFreeHost(c->h_flags);
  for (int i = 0; i < 2; i++) { cudaFree(c->d_plant_pos[i]); cudaFree(c->d_plant_size[i]); }
  cudaFree(c->d_compact);
  if (c->h_ring) cudaFreeHost(c->h_ring);
  for (cudaEvent_t e : c->ev_ring)
    if (e) cudaEventDestroy(e);
  cudaFree(c->d_veg_flags); cudaFree(c->d_veg_child); cudaFree(c->d_veg_grown); cudaFree(c->d_veg_blocks); cudaFree(c->d_veg_totals);
  if (c->h_veg_totals) cudaFreeHost(c->h_veg_totals);
  delete c;
}

int shx_create(shx_ctx** out, const shx_params* p, const shx_config* cfg_in) {
  if (!out || !p) return fail(SHX_ERR_ARG, "shx_create: null argument");
  *out = nullptr;
  if (p->lodsize != 1) return fail(SHX_ERR_ARG, "only lodsize == 1 is supported (cellpool.h:178)");
  if (p->mapsize < 1 || p->tilesize < 4 || (long long)p->mapsize * p->tilesize > 16384)
    return fail(SHX_ERR_ARG, "bad geometry (side must be <= 16384 cells: cell indices travel in 28 bits)");
  shx_config cfg;
  if (cfg_in) cfg = *cfg_in; else shx_default_config(&cfg);
  const int size = p->mapsize * p->tilesize;
  const bool peer = cfg.peer_world > 1;
  if (peer) {  // one strip of a world shared over NVLink: no halo, rows fixed by the rank
    if (cfg.peer_world > kMaxPeers || cfg.peer_rank < 0 || cfg.peer_rank >= cfg.peer_world) return fail(SHX_ERR_ARG, "bad peer rank/world");
    const int rows = size / cfg.peer_world;
    if (rows * cfg.peer_world != size || (rows & (rows - 1)) || rows % p->tilesize)
      return fail(SHX_ERR_ARG, "peer mode needs size/world to be a power of two and a multiple of tilesize");
    if (cfg.mode == SHX_MODE_SEQUENTIAL) return fail(SHX_ERR_MODE, "peer mode is batched only");
    cfg.row0 = cfg.peer_rank * rows;
    cfg.row1 = cfg.row0 + rows;
    cfg.halo = 0;
  }
  if (cfg.row0 == 0 && cfg.row1 == 0) cfg.row1 = size;
  if (cfg.row0 < 0 || cfg.row1 > size || cfg.row0 >= cfg.row1) return fail(SHX_ERR_ARG, "bad strip rows");
  if (cfg.free_waits < 0 || cfg.free_waits > 15) return fail(SHX_ERR_ARG, "free_waits must be 0..15");
  const bool whole = cfg.row0 == 0 && cfg.row1 == size;
  if (!whole && !peer && cfg.halo < 2) return fail(SHX_ERR_ARG, "strip halo must be >= 2 rows");
  if (!whole && cfg.mode == SHX_MODE_SEQUENTIAL) return fail(SHX_ERR_MODE, "sequential mode is whole-map only");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SHX_ERR_CUDA, "no CUDA device: shx has no CPU fallback");
  if (cfg.device < 0 || cfg.device >= ndev) return fail(SHX_ERR_ARG, "bad device ordinal");
  CU(cudaSetDevice(cfg.device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg.device));
  if (prop.major < 10) return fail(SHX_ERR_CUDA, "shx kernels are built for sm_100a only");

  shx_ctx* c = new (std::nothrow) shx_ctx();
  if (!c) return fail(SHX_ERR_NOMEM, "host allocation failed");
  c->p = *p;
  c->cfg = cfg;
  c->size = size;
  c->sm_count = prop.multiProcessorCount;
  c->halo_lo = whole ? 0 : std::min(cfg.halo, cfg.row0);
  c->halo_hi = whole ? 0 : std::min(cfg.halo, size - cfg.row1);
  c->m.size = size;
  c->m.xlo = cfg.row0 - c->halo_lo;
  c->m.nrows = (cfg.row1 + c->halo_hi) - c->m.xlo;
  c->m.row0 = cfg.row0;
  c->m.row1 = cfg.row1;
  c->stored_cells = (size_t)c->m.nrows * size;
  c->owned_cells = (size_t)(cfg.row1 - cfg.row0) * size;
  c->max_drops = cfg.max_drops ? cfg.max_drops : (size_t)p->mapsize * p->mapsize * 1024;
  const size_t tile_cells = (size_t)p->tilesize * p->tilesize;

#define ALLOC(ptr, bytes)                                                  \
  do {                                                                     \
    if (cudaMalloc((void**)&(ptr), (bytes)) != cudaSuccess) {              \
      shx_destroy(c);                                                      \
      return fail(SHX_ERR_NOMEM, "cudaMalloc failed for " #ptr);           \
    }                                                                      \
  } while (0)
  ALLOC(c->m.hq, c->stored_cells * sizeof(int4));
  ALLOC(c->m.rec, c->stored_cells * sizeof(CellRec));
  ALLOC(c->d_drops, c->max_drops * sizeof(shx_drop));
  ALLOC(c->d_xy, c->max_drops * 2 * sizeof(float));
  ALLOC(c->d_bar, sizeof(GridBar));
  ALLOC(c->d_stats, ST_COUNT * 8);
  ALLOC(c->d_flags, 8 * sizeof(int));
  ALLOC(c->d_trace, (size_t)c->trace_cap * 7 * sizeof(float));
  ALLOC(c->d_u32, 4 * sizeof(unsigned));
  ALLOC(c->d_stage, tile_cells * sizeof(shx_cell));
  if (!whole) {
    ALLOC(c->d_halo_ref[0], (size_t)std::max(1, c->halo_lo) * size * sizeof(int32_t));
    ALLOC(c->d_halo_ref[1], (size_t)std::max(1, c->halo_hi) * size * sizeof(int32_t));
  }
#undef ALLOC
  if (cudaMallocHost((void**)&c->h_stats, ST_COUNT * 8) != cudaSuccess ||
      cudaMallocHost((void**)&c->h_flags, 8 * sizeof(int)) != cudaSuccess) {
    shx_destroy(c);
    return fail(SHX_ERR_NOMEM, "cudaMallocHost failed");
  }
  cudaMemset(c->m.hq, 0, c->stored_cells * sizeof(int4));
  cudaMemset(c->m.rec, 0, c->stored_cells * sizeof(CellRec));
  cudaMemset(c->d_stats, 0, ST_COUNT * 8);
  cudaMemset(c->d_flags, 0, 8 * sizeof(int));
  if (!cfg.no_l2_window && prop.persistingL2CacheMaxSize > 0 &&
      c->stored_cells * sizeof(int4) <= (size_t)prop.persistingL2CacheMaxSize &&
      c->stored_cells * sizeof(int4) <= (size_t)prop.accessPolicyMaxWindowSize) {
    // device-wide carve-out (shared by all contexts of the process on this device)
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize) == cudaSuccess) {
      c->l2_window.base_ptr = c->m.hq;
      c->l2_window.num_bytes = c->stored_cells * sizeof(int4);
      c->l2_window.hitRatio = 1.0f;
      c->l2_window.hitProp = cudaAccessPropertyPersisting;
      c->l2_window.missProp = cudaAccessPropertyStreaming;
      c->l2_window_on = true;
    } else {
      cudaGetLastError();
    }
  }
  c->peer = peer;
  if (peer) {
    if (cudaMalloc((void**)&c->d_inbox, 3 * kMaxPeers * sizeof(unsigned long long)) != cudaSuccess) {
      shx_destroy(c);
      return fail(SHX_ERR_NOMEM, "cudaMalloc failed for the peer inbox");
    }
    cudaMemset(c->d_inbox, 0, 3 * kMaxPeers * sizeof(unsigned long long));
    const int rows = cfg.row1 - cfg.row0;
    int shift = 0;
    while ((1 << shift) < rows) shift++;
    c->pv.shift = shift;
    c->pv.mask = rows - 1;
    c->pv.nranks = cfg.peer_world;
    c->pv.rank = cfg.peer_rank;
    c->pv.hq[cfg.peer_rank] = c->m.hq;
    c->pv.rec[cfg.peer_rank] = c->m.rec;
    c->pv.inbox[cfg.peer_rank] = c->d_inbox;
  }

  // Launch shapes of the multi-CTA descend kernel, both with the cooperative gather.  [0] "spread":
  // CTAs of 64 so that a few thousand drops still cover all SMs (latency-bound regime); [1] "dense":
  // 2 x 448 threads per SM at 72 registers (throughput regime, >= ~256 drops per SM).  An explicit
  // block_threads / variant / coop / grid_blocks in the config forces one shape for both.
  const bool forced = cfg.block_threads > 0 || cfg.variant > 0 || cfg.coop > 0 || cfg.grid_blocks > 0;
  if (forced && peer) {
    shx_destroy(c);
    return fail(SHX_ERR_ARG, "peer mode does not take launch-shape overrides");
  }
  // Launch shapes, chosen by batch size (kShapeLimit; measured with the L2 hints on the REDs, profiles/r2_l2_hints.txt):
  // [0] eight lanes per drop in CTAs of 128 (16 drops): up to 6 144 drops, the reference's default call; one thread
  // per drop beyond, all with the cooperative gather at 72 registers: [1] CTAs of 64 (up to 24 576 drops: a 2048^2
  // world, the strips of an eight-GPU run), [2] CTAs of 128 (up to 49 152: a 4096^2 world, the strips of a four-GPU
  // run), [3] "dense", 2 x 448 threads per SM (the 8192^2 cycle).  An explicit block_threads / variant / coop /
  // grid_blocks in the config forces one shape for all (variant 5 = eight lanes per drop).
  for (int i = 0; i < 4; i++) {
    LaunchShape& ls = c->shape[i];
    if (forced && cfg.variant == 5) {
      ls.lanes = 8;
      ls.block = cfg.block_threads > 0 ? std::min(1024, (cfg.block_threads + 31) / 32 * 32) : 256;
      ls.kernel = ls.block <= 256 ? (const void*)descend_group_kernel<256, 4> : (const void*)descend_group_kernel<1024, 1>;
    } else if (forced) {  // one thread per drop in one of its shapes
      ls.block = cfg.block_threads > 0 ? std::min(1024, (cfg.block_threads + 31) / 32 * 32) : 256;
      if (cfg.variant == 1) ls.block = std::min(ls.block, 512);
      if (cfg.variant == 2) ls.block = std::min(ls.block, 128);
      if (cfg.variant == 3) ls.block = 448;
      ls.kernel = big_kernel(ls.block, cfg.variant, cfg.coop == 1);
    } else if (i == 0 && !peer) {
      ls.lanes = 8;
      ls.block = 128;  // 16 drops per CTA: 512 drops cover 32 SMs (CTAs of 256: 4.50 us per phase, of 64-128: 4.38)
      ls.kernel = (const void*)descend_group_kernel<256, 4>;
    } else if (i <= 2) {
      ls.block = i == 2 ? 128 : 64;
      ls.kernel = peer ? (const void*)descend_lockstep_kernel<128, 7, true, true> : big_kernel(ls.block, 2, true);
    } else {
      ls.block = 448;
      ls.kernel = peer ? (const void*)descend_lockstep_kernel<448, 2, true, true> : big_kernel(448, 3, true);
    }
    int nb = 0;
    if (cudaFuncSetAttribute(ls.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ls.smem(ls.block)) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls.kernel, ls.block, ls.smem(ls.block)) != cudaSuccess || nb < 1) {
      shx_destroy(c);
      return fail(SHX_ERR_CUDA, "descend kernel does not fit on this device");
    }
    ls.cap_blocks = nb * c->sm_count;
  }
  c->forced_shape = forced;
  if (cudaFuncSetAttribute(KERNEL_SMALL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kGroupSmemWords * 4 * 1024)) != cudaSuccess) {
    shx_destroy(c);
    return fail(SHX_ERR_CUDA, "descend kernel does not fit on this device");
  }
  if (cudaDeviceSynchronize() != cudaSuccess) {
    shx_destroy(c);
    return fail(SHX_ERR_CUDA, "context initialisation failed");
  }
  *out = c;
  return SHX_OK;
}

int shx_set_params(shx_ctx* c, const shx_params* p) {
  if (!c || !p) return fail(SHX_ERR_ARG, "null argument");
  if (p->mapsize != c->p.mapsize || p->tilesize != c->p.tilesize || p->lodsize != 1)
    return fail(SHX_ERR_ARG, "geometry cannot change after shx_create");
  c->p = *p;
  return SHX_OK;
}

int shx_get_params(const shx_ctx* c, shx_params* p) {
  if (!c || !p) return fail(SHX_ERR_ARG, "null argument");
  *p = c->p;
  return SHX_OK;
}

int shx_set_stream(shx_ctx* c, void* s) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  c->stream = (cudaStream_t)s;
  return SHX_OK;
}

int shx_sync(shx_ctx* c) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

int shx_host_register(void* ptr, size_t bytes) {
  if (!ptr) return fail(SHX_ERR_ARG, "null pointer");
  CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return SHX_OK;
}
int shx_host_unregister(void* ptr) {
  if (!ptr) return fail(SHX_ERR_ARG, "null pointer");
  CU(cudaHostUnregister(ptr));
  return SHX_OK;
}

int shx_stored_rows(const shx_ctx* c, int* xlo, int* nrows) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  if (xlo) *xlo = c->m.xlo;
  if (nrows) *nrows = c->m.nrows;
  return SHX_OK;
}

// ------------------------------------------------------------------------------- transfers

// ---- optional per-kernel timing with CUDA events on the context's stream (bench.py's roofline)
static int span_begin(shx_ctx* c, int kind, cudaStream_t on = nullptr, bool use_on = false) {
  if (!c->timing) return SHX_OK;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  c->spans.push_back({kind, e0, e1});
  CU(cudaEventRecord(e0, use_on ? on : c->stream));
  return SHX_OK;
}
static int span_end(shx_ctx* c, cudaStream_t on = nullptr, bool use_on = false) {
  if (!c->timing) return SHX_OK;
  CU(cudaEventRecord(c->spans.back().e1, use_on ? on : c->stream));
  return SHX_OK;
}


static int refresh_halo_ref(shx_ctx* c) {
  for (int side = 0; side < 2; side++) {
    const int rows = side == 0 ? c->halo_lo : c->halo_hi;
    if (!rows || !c->d_halo_ref[side]) continue;
    const size_t n = (size_t)rows * c->size;
    const size_t off = side == 0 ? 0 : (size_t)(c->m.row1 - c->m.xlo) * c->size;
    strip_get_rows_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(c->m.hq + off, c->d_halo_ref[side], n);
    c->launches++;
  }
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_upload(shx_ctx* c, const shx_cell* pool, size_t ncells) {
  if (!c || !pool) return fail(SHX_ERR_ARG, "null argument");
  if (ncells != (size_t)c->size * c->size) return fail(SHX_ERR_ARG, "pool size does not match the geometry");
  CU(cudaSetDevice(c->cfg.device));
  const int ts = c->p.tilesize, ms = c->p.mapsize;
  const size_t tile_cells = (size_t)ts * ts;
  CU(cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->stream));
  for (int ti = 0; ti < ms; ti++) {
    if ((ti + 1) * ts <= c->m.xlo || ti * ts >= c->m.xlo + c->m.nrows) continue;
    for (int tj = 0; tj < ms; tj++) {
      const size_t node = (size_t)ti * ms + tj;  // cellpool.h:330
      CU(cudaMemcpyAsync(c->d_stage, pool + node * tile_cells, tile_cells * sizeof(shx_cell), cudaMemcpyHostToDevice, c->stream));
      TileArgs a{c->m, sequential(c) ? 1 : 0, ts, ti * ts, tj * ts, c->d_flags};
      unpack_tile_kernel<<<grid_for(c, tile_cells), 256, 0, c->stream>>>(a, c->d_stage);
      c->launches++;
    }
  }
  CU(cudaGetLastError());
  c->tracks_clean = false;  // the host's track values were taken over as they are
  c->root_counts_valid = false;
  int rc = refresh_halo_ref(c);
  if (rc) return rc;
  CU(cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (c->h_flags[0])
    return fail(SHX_ERR_RANGE, "a height is outside (-31, 31) or a track outside (-1024, 1024): does not fit the fixed point");
  return SHX_OK;
}

int shx_download_async(shx_ctx* c, shx_cell* pool, size_t ncells, unsigned mask) {
  if (!c || !pool) return fail(SHX_ERR_ARG, "null argument");
  if (ncells != (size_t)c->size * c->size) return fail(SHX_ERR_ARG, "pool size does not match the geometry");
  mask &= SHX_F_ALL;
  if (!mask) return SHX_OK;
  CU(cudaSetDevice(c->cfg.device));
  const int ts = c->p.tilesize, ms = c->p.mapsize;
  const size_t tile_cells = (size_t)ts * ts;
  // byte runs of the 32-byte record selected by the mask
  const bool want[8] = {(mask & SHX_F_HEIGHT) != 0, (mask & SHX_F_DISCHARGE) != 0, (mask & SHX_F_MOMENTUM) != 0,
                        (mask & SHX_F_MOMENTUM) != 0, (mask & SHX_F_TRACKS) != 0, (mask & SHX_F_TRACKS) != 0,
                        (mask & SHX_F_TRACKS) != 0, (mask & SHX_F_ROOTDENSITY) != 0};
  // Two staging tiles and a copy stream: the pack kernel of tile t+1 runs while tile t crosses PCIe.
  if (!c->copy_stream) {
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CU(cudaEventCreateWithFlags(&c->ev_packed[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
    }
    if (cudaMalloc((void**)&c->d_stage2, tile_cells * sizeof(shx_cell)) != cudaSuccess) {
      cudaGetLastError();
      return fail(SHX_ERR_NOMEM, "cudaMalloc failed for the second staging tile");
    }
  }
  int turn = 0;
  bool used[2] = {false, false};
  for (int ti = 0; ti < ms; ti++) {
    const int lx0 = std::max(c->m.row0 - ti * ts, 0), lx1 = std::min(c->m.row1 - ti * ts, ts);
    if (lx0 >= lx1) continue;
    for (int tj = 0; tj < ms; tj++) {
      const size_t node = (size_t)ti * ms + tj;
      const int b = turn & 1;
      turn++;
      shx_cell* stage = b ? c->d_stage2 : c->d_stage;
      if (used[b]) CU(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));  // the copy that read this buffer is done
      TileArgs a{c->m, sequential(c) ? 1 : 0, ts, ti * ts, tj * ts, c->d_flags};
      { const int rc_s = span_begin(c, 3); if (rc_s) return rc_s; }
      pack_tile_kernel<<<grid_for(c, tile_cells), 256, 0, c->stream>>>(a, stage);
      { const int rc_s = span_end(c); if (rc_s) return rc_s; }
      c->launches++;
      CU(cudaEventRecord(c->ev_packed[b], c->stream));
      CU(cudaStreamWaitEvent(c->copy_stream, c->ev_packed[b], 0));
      const size_t first = (size_t)lx0 * ts, count = (size_t)(lx1 - lx0) * ts;
      char* dst = reinterpret_cast<char*>(pool + node * tile_cells + first);
      const char* src = reinterpret_cast<const char*>(stage + first);
      { const int rc_s = span_begin(c, 4, c->copy_stream, true); if (rc_s) return rc_s; }
      if (mask == SHX_F_ALL) {
        CU(cudaMemcpyAsync(dst, src, count * sizeof(shx_cell), cudaMemcpyDeviceToHost, c->copy_stream));
      } else {
        for (int f = 0; f < 8;) {
          if (!want[f]) { f++; continue; }
          int g = f;
          while (g < 8 && want[g]) g++;
          CU(cudaMemcpy2DAsync(dst + 4 * f, sizeof(shx_cell), src + 4 * f, sizeof(shx_cell), (size_t)4 * (g - f), count,
                               cudaMemcpyDeviceToHost, c->copy_stream));
          f = g;
        }
      }
      { const int rc_s = span_end(c, c->copy_stream, true); if (rc_s) return rc_s; }
      CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
      used[b] = true;
    }
  }
  // stream order for the caller: whatever follows on the context's stream (and shx_sync) sees the copies done
  for (int b = 0; b < 2; b++)
    if (used[b]) CU(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
  CU(cudaGetLastError());
  return SHX_OK;
}

// Download of what host code reads after World::erode, in half the PCIe bytes: {height, discharge, momentumx,
// momentumy} are the FIRST 16 bytes of a quad::cell (cellpool.h:207-220).  One kernel packs them densely (16 bytes per
// owned cell, pool order), the copy engine streams the dense buffer chunk by chunk into a ring of pinned host chunks,
// and `nthreads` host threads scatter every chunk into the caller's 32-byte records as it lands (the strided 16-of-32
// byte DMA straight into the pool measured 2x SLOWER than whole records).  The *_track words (scratch of the erode
// call, zeroed first thing by the next one, world.h:56-61) and rootdensity (host-owned) are left as they are.
// Whether this beats the whole-record download depends on the host's memory bandwidth (every record costs a
// read-for-ownership of its cache line, and that traffic slows the DMA down): measured on a 16-core B200 host, 8192^2:
// 52.3 ms per frame against 51.7 with whole records -- no gain there, hence opt-in (shx::Bridge::set_download_mode).
int shx_download_compact(shx_ctx* c, shx_cell* pool, size_t ncells, int nthreads) {
  if (!c || !pool) return fail(SHX_ERR_ARG, "null argument");
  if (ncells != (size_t)c->size * c->size) return fail(SHX_ERR_ARG, "pool size does not match the geometry");
  CU(cudaSetDevice(c->cfg.device));
  const int ts = c->p.tilesize, ms = c->p.mapsize;
  if (c->m.row0 % ts || c->m.row1 % ts) return fail(SHX_ERR_ARG, "compact download needs strips of whole tile rows");
  constexpr int kRingSlots = 8;
  const size_t chunk_cells = (size_t)1 << 18;  // 4 MiB of 16-byte records per chunk
  const size_t first_cell = (size_t)(c->m.row0 / ts) * ms * ts * ts;  // owned nodes are contiguous in the pool
  const size_t n = c->owned_cells;
  if (!c->d_compact) {
    if (cudaMalloc((void**)&c->d_compact, n * sizeof(float4)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_ring, kRingSlots * chunk_cells * sizeof(float4)) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(c->d_compact);
      c->d_compact = nullptr;
      return fail(SHX_ERR_NOMEM, "allocation of the compact download buffers failed");
    }
    for (int i = 0; i < kRingSlots; i++) CU(cudaEventCreateWithFlags(&c->ev_ring[i], cudaEventDisableTiming));
  }
  if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  { const int rc_s = span_begin(c, 3); if (rc_s) return rc_s; }
  pack_fields16_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(c->m, sequential(c) ? 1 : 0, ts, ms, first_cell, n, c->d_compact);
  { const int rc_s = span_end(c); if (rc_s) return rc_s; }
  c->launches++;
  CU(cudaGetLastError());
  cudaEvent_t packed;
  CU(cudaEventCreateWithFlags(&packed, cudaEventDisableTiming));
  CU(cudaEventRecord(packed, c->stream));
  CU(cudaStreamWaitEvent(c->copy_stream, packed, 0));
  cudaEventDestroy(packed);  // released once the wait has consumed it

  const size_t nchunks = (n + chunk_cells - 1) / chunk_cells;
  const int T = std::max(1, std::min(nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency(), 64));
  auto enqueue = [&](size_t k) -> int {
    const size_t cells = std::min(chunk_cells, n - k * chunk_cells);
    const int slot = (int)(k % kRingSlots);
    { const int rc_s = span_begin(c, 4, c->copy_stream, true); if (rc_s) return rc_s; }
    CU(cudaMemcpyAsync(c->h_ring + (size_t)slot * chunk_cells * sizeof(float4), c->d_compact + k * chunk_cells, cells * sizeof(float4),
                       cudaMemcpyDeviceToHost, c->copy_stream));
    { const int rc_s = span_end(c, c->copy_stream, true); if (rc_s) return rc_s; }
    CU(cudaEventRecord(c->ev_ring[slot], c->copy_stream));
    return SHX_OK;
  };
  for (size_t k = 0; k < std::min<size_t>(nchunks, kRingSlots); k++) {
    const int rc = enqueue(k);
    if (rc) return rc;
  }
  // worker t scatters the cells [t, t + T, ...) x 4096-cell runs of every chunk; chunk k may be read once `ready` > k
  std::atomic<long> ready{0};
  std::vector<std::atomic<long>> progress(T);
  for (auto& p : progress) p.store(0);
  shx_cell* const dst0 = pool + first_cell;
  const char* const ring = c->h_ring;
  auto scatter_share = [&](int t, size_t k) {
    const size_t cells = std::min(chunk_cells, n - k * chunk_cells);
    const float4* src = reinterpret_cast<const float4*>(ring + (k % kRingSlots) * chunk_cells * sizeof(float4));
    shx_cell* dst = dst0 + k * chunk_cells;
    constexpr size_t kRun = 4096;
    for (size_t r0 = (size_t)t * kRun; r0 < cells; r0 += (size_t)T * kRun) {
      const size_t r1 = std::min(cells, r0 + kRun);
      for (size_t i = r0; i < r1; i++) memcpy(dst + i, src + i, sizeof(float4));  // first half of the 32-byte record
    }
  };
  auto worker = [&](int t) {
    for (size_t k = 0; k < nchunks; k++) {
      while (ready.load(std::memory_order_acquire) <= (long)k) std::this_thread::yield();
      scatter_share(t, k);
      progress[t].store((long)k + 1, std::memory_order_release);
    }
  };
  std::vector<std::thread> pool_threads;
  for (int t = 1; t < T; t++) pool_threads.emplace_back(worker, t);
  int rc = SHX_OK;
  const auto t_host0 = std::chrono::steady_clock::now();
  for (size_t k = 0; k < nchunks; k++) {
    if (cudaEventSynchronize(c->ev_ring[k % kRingSlots]) != cudaSuccess) rc = fail(SHX_ERR_CUDA, "device-to-host copy failed");
    ready.store((long)k + 1, std::memory_order_release);  // (on failure too: the workers must run to the end)
    scatter_share(0, k);
    progress[0].store((long)k + 1, std::memory_order_release);
    if (k + kRingSlots < nchunks && rc == SHX_OK) {  // the slot of chunk k is free once every worker has left it
      for (int t = 1; t < T; t++)
        while (progress[t].load(std::memory_order_acquire) <= (long)k) std::this_thread::yield();
      rc = enqueue(k + kRingSlots);
    }
  }
  if (rc != SHX_OK) ready.store((long)nchunks + 1, std::memory_order_release);
  for (auto& th : pool_threads) th.join();
  c->scatter_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
  return rc;
}

int shx_download(shx_ctx* c, shx_cell* pool, size_t ncells, unsigned mask) {
  int rc = shx_download_async(c, pool, ncells, mask);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

static int view_staging(shx_ctx* c, size_t bytes);

int shx_download_raw(shx_ctx* c, int32_t* hq2, void* rec32) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaStreamSynchronize(c->stream));
  if (hq2) {  // the two height words of every stored cell (the claim words stay behind)
    int rc = view_staging(c, c->stored_cells * sizeof(int2));
    if (rc) return rc;
    const int grid = (int)std::min<size_t>((c->stored_cells + 255) / 256, (size_t)c->sm_count * 16);
    copy_heights_kernel<<<grid, 256, 0, c->stream>>>(c->m.hq, reinterpret_cast<int2*>(c->d_view), c->stored_cells);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hq2, c->d_view, c->stored_cells * sizeof(int2), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  if (rec32) CU(cudaMemcpy(rec32, c->m.rec, c->stored_cells * sizeof(CellRec), cudaMemcpyDeviceToHost));
  return SHX_OK;
}

// ------------------------------------------------------------------------------- per-frame views

static ViewArgs view_args(const shx_ctx* c) {
  ViewArgs a;
  a.m = c->m;
  a.tilesize = c->p.tilesize;
  a.mapsize = c->p.mapsize;
  a.sequential = sequential(c) ? 1 : 0;
  a.mapscale = c->p.mapscale;
  return a;
}

static int view_staging(shx_ctx* c, size_t bytes) {
  if (c->view_bytes >= bytes) return SHX_OK;
  cudaFree(c->d_view);
  c->d_view = nullptr;
  c->view_bytes = 0;
  if (cudaMalloc((void**)&c->d_view, bytes) != cudaSuccess) {
    cudaGetLastError();
    return fail(SHX_ERR_NOMEM, "cudaMalloc failed for the view staging buffer");
  }
  c->view_bytes = bytes;
  return SHX_OK;
}

int shx_vertex_fill(shx_ctx* c, float* dev_out) {  // cellpool.h:286-305 over every owned node
  if (!c || !dev_out) return fail(SHX_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->cfg.device));
  const int ts = c->p.tilesize;
  if (c->m.row0 % ts || c->m.row1 % ts) return fail(SHX_ERR_ARG, "vertex fill needs tile-aligned strips");
  const size_t first = (size_t)(c->m.row0 / ts) * c->p.mapsize * ts * ts;
  const int grid = (int)std::min<size_t>((c->owned_cells + 255) / 256, (size_t)c->sm_count * 16);
  vertex_fill_kernel<<<grid, 256, 0, c->stream>>>(view_args(c), dev_out, c->owned_cells, first);
  c->launches++;
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_vertex_download(shx_ctx* c, float* host_out, size_t ncells) {
  if (!c || !host_out) return fail(SHX_ERR_ARG, "null argument");
  if (ncells != c->owned_cells) return fail(SHX_ERR_ARG, "vertex buffer must hold the owned cells");
  int rc = view_staging(c, c->owned_cells * 12 * sizeof(float));
  if (rc) return rc;
  if ((rc = shx_vertex_fill(c, c->d_view))) return rc;
  CU(cudaMemcpyAsync(host_out, c->d_view, c->owned_cells * 12 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

int shx_view_maps(shx_ctx* c, float* dev_out) {  // SimpleHydrology.cpp:341-354
  if (!c || !dev_out) return fail(SHX_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->cfg.device));
  const int grid = (int)std::min<size_t>((c->owned_cells + 255) / 256, (size_t)c->sm_count * 16);
  view_maps_kernel<<<grid, 256, 0, c->stream>>>(view_args(c), reinterpret_cast<float4*>(dev_out), c->owned_cells);
  c->launches++;
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_view_maps_download(shx_ctx* c, float* host_out, size_t ncells) {
  if (!c || !host_out) return fail(SHX_ERR_ARG, "null argument");
  if (ncells != c->owned_cells) return fail(SHX_ERR_ARG, "map buffer must hold the owned cells");
  int rc = view_staging(c, c->owned_cells * 4 * sizeof(float));
  if (rc) return rc;
  if ((rc = shx_view_maps(c, c->d_view))) return rc;
  CU(cudaMemcpyAsync(host_out, c->d_view, c->owned_cells * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

int shx_view_textures(shx_ctx* c, const float* water_rgb, uint8_t* dev_discharge_rgba, uint8_t* dev_momentum_rgba) {
  if (!c || !dev_discharge_rgba || !dev_momentum_rgba) return fail(SHX_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->cfg.device));
  // model.h:22: waterColor = vec3(92, 133, 142) / 255.0f
  const float3 water = water_rgb ? make_float3(water_rgb[0], water_rgb[1], water_rgb[2])
                                 : make_float3(92.0f / 255.0f, 133.0f / 255.0f, 142.0f / 255.0f);
  const int grid = (int)std::min<size_t>((c->owned_cells + 255) / 256, (size_t)c->sm_count * 16);
  view_textures_kernel<<<grid, 256, 0, c->stream>>>(view_args(c), water, reinterpret_cast<unsigned*>(dev_discharge_rgba),
                                                     reinterpret_cast<unsigned*>(dev_momentum_rgba), c->owned_cells);
  c->launches++;
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_view_textures_download(shx_ctx* c, const float* water_rgb, uint8_t* host_discharge_rgba, uint8_t* host_momentum_rgba, size_t ncells) {
  if (!c || !host_discharge_rgba || !host_momentum_rgba) return fail(SHX_ERR_ARG, "null argument");
  if (ncells != c->owned_cells) return fail(SHX_ERR_ARG, "texture buffers must hold the owned cells");
  int rc = view_staging(c, c->owned_cells * 8);
  if (rc) return rc;
  uint8_t* d = reinterpret_cast<uint8_t*>(c->d_view);
  if ((rc = shx_view_textures(c, water_rgb, d, d + 4 * c->owned_cells))) return rc;
  CU(cudaMemcpyAsync(host_discharge_rgba, d, 4 * c->owned_cells, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(host_momentum_rgba, d + 4 * c->owned_cells, 4 * c->owned_cells, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

int shx_gather_cells(shx_ctx* c, const int* xy, size_t n, shx_cell* out, float* normals3) {
  if (!c || (n && (!xy || !out))) return fail(SHX_ERR_ARG, "null argument");
  if (!n) return SHX_OK;
  CU(cudaSetDevice(c->cfg.device));
  // staging: [xy | cells | normals]
  const size_t b_xy = n * 2 * sizeof(int), b_cells = n * sizeof(shx_cell), b_n = n * 3 * sizeof(float);
  int rc = view_staging(c, b_xy + b_cells + b_n);
  if (rc) return rc;
  char* base = reinterpret_cast<char*>(c->d_view);
  int* d_xy = reinterpret_cast<int*>(base + b_cells);  // cells first: 32-byte aligned
  float* d_n = reinterpret_cast<float*>(base + b_cells + b_xy);
  CU(cudaMemcpyAsync(d_xy, xy, b_xy, cudaMemcpyHostToDevice, c->stream));
  gather_cells_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(view_args(c), d_xy, n, reinterpret_cast<shx_cell*>(base),
                                                            normals3 ? d_n : nullptr);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, base, b_cells, cudaMemcpyDeviceToHost, c->stream));
  if (normals3) CU(cudaMemcpyAsync(normals3, d_n, b_n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

// ------------------------------------------------------------------------------- erode pieces

static int fetch_stats(shx_ctx* c, shx_stats* out) {
  CU(cudaMemcpyAsync(c->h_stats, c->d_stats, ST_COUNT * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->h_flags + 2, c->d_flags + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->h_flags + 4, c->d_flags + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (out) {
    memcpy(out, c->h_stats, sizeof(shx_stats));
    out->launches = c->launches;
  }
  if (c->h_flags[2]) {
    const int why = c->h_flags[2];
    CU(cudaMemsetAsync(c->d_flags + 2, 0, sizeof(int), c->stream));
    if (why == 2) return fail(SHX_ERR_RANGE, "a launch needed more than 4000 phases (the batched mode supports maxAge up to ~3990); the call's result is invalid");
    return fail(SHX_ERR_PEER, "a peer GPU did not reach a phase barrier within the time-out; the call's result is invalid");
  }
  if (c->h_flags[0]) {
    CU(cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->stream));
    return fail(SHX_ERR_RANGE, "a discharge track left the Q13.18 range (more than ~4096 drop visits of one cell in one call)");
  }
  if (c->h_flags[4]) {
    CU(cudaMemsetAsync(c->d_flags + 4, 0, sizeof(int), c->stream));
    return fail(SHX_ERR_RANGE, "a height left (-30, 30): the world is diverging and would wrap the Q5.26 fixed point");
  }
  return SHX_OK;
}

static int begin_call(shx_ctx* c) {
  CU(cudaSetDevice(c->cfg.device));
  c->launches = 0;
  CU(cudaMemsetAsync(c->d_stats, 0, ST_COUNT * 8, c->stream));
  return SHX_OK;
}

int shx_reset_tracks(shx_ctx* c) {  // world.h:56-61
  if (!c) return fail(SHX_ERR_ARG, "null context");
  CU(cudaSetDevice(c->cfg.device));
  const int grid = (int)std::min<size_t>((c->stored_cells + 255) / 256, (size_t)c->sm_count * 16);
  reset_tracks_kernel<<<grid, 256, 0, c->stream>>>(c->m.rec, c->stored_cells);
  c->launches++;
  c->tracks_clean = true;
  CU(cudaGetLastError());
  return SHX_OK;
}

static int ema_launch(shx_ctx* c, bool reset) {  // world.h:81-86 (+ :56-61 for the next call)
  const size_t off = (size_t)(c->m.row0 - c->m.xlo) * c->size;
  const int grid = (int)std::min<size_t>((c->owned_cells + 255) / 256, (size_t)c->sm_count * 16);
  ema_kernel<<<grid, 256, 0, c->stream>>>(c->m.rec + off, c->owned_cells, c->p.lrate, sequential(c) ? 1 : 0, reset ? 1 : 0,
                                           c->d_flags);
  c->launches++;
  c->tracks_clean = reset;
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_ema(shx_ctx* c) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  CU(cudaSetDevice(c->cfg.device));
  return ema_launch(c, false);
}

// Claim keys carry a 4-bit launch epoch so that the keys earlier launches left in the cells lose
// against this launch's; when it wraps, the claim words are cleared (one pass over the height plane
// every 15 launches).  Peer mode: every rank counts the same launches, so the epochs agree.
static int next_claim_epoch(shx_ctx* c) {
  if (++c->claim_epoch > 15u) {
    const int grid = (int)std::min<size_t>((c->stored_cells + 255) / 256, (size_t)c->sm_count * 16);
    clear_claims_kernel<<<grid, 256, 0, c->stream>>>(c->m.hq, c->stored_cells, c->d_flags + 4);
    c->launches++;
    CU(cudaGetLastError());
    c->claim_epoch = 1u;
  }
  return SHX_OK;
}


// Cooperative launch of a descend kernel.  Where the height / claim words of the context fit the persisting part of
// L2 (2048^2: 64 MiB of 16-byte cell words), the launch carries an access-policy window over them: the per-call
// streaming passes (EMA over 32-byte records, vertex fill, view maps) then cannot evict the words every phase of
// every drop gathers, and the phases of a latency-bound call hit L2 instead of HBM.
static cudaError_t launch_descend(shx_ctx* c, const void* kernel, int grid, int block, void** args, size_t smem, int lanes_per_drop) {
  c->last_grid = grid;
  c->last_block = block;
  c->last_lanes = lanes_per_drop;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.numAttrs = 1;
  if (c->l2_window_on) {
    at[1].id = cudaLaunchAttributeAccessPolicyWindow;
    at[1].val.accessPolicyWindow = c->l2_window;
    cfg.numAttrs = 2;
  }
  cfg.attrs = at;
  return cudaLaunchKernelExC(&cfg, kernel, args);
}

// march n drops already in c->d_drops
static int run_device_drops(shx_ctx* c, size_t n, bool trace, bool align_age = false) {
  c->last_n = n;
  if (n > c->max_drops) return fail(SHX_ERR_CAPACITY, "more drops than max_drops");
  if (c->peer) {
    // exactly ONE launch per call on every rank (even with no drops of its own): the kernels of
    // all ranks meet at every phase barrier
    if (!c->peer_attached) return fail(SHX_ERR_PEER, "shx_peer_attach has not been called");
    const LaunchShape& ls = c->shape[n <= kShapeLimit[1] ? 1 : (n <= kShapeLimit[2] ? 2 : 3)];  // peer mode: one thread per drop
    if (n > (size_t)ls.cap_blocks * ls.block) return fail(SHX_ERR_CAPACITY, "peer mode runs a call's drops in one launch");
    c->tracks_clean = false;
    DescendArgs a;
    a.m = c->m;
    a.pv = c->pv;
    a.pv.tag_base = ((++c->peer_seq) & 0xFFFu) << 20;
    a.P = step_params(c->p);
    a.drops = c->d_drops;
    a.ndrops = (unsigned)n;
    a.align_age = 0u;
    a.free_waits = (unsigned)c->cfg.free_waits;
    a.abort_flag = c->d_flags + 2;  // OR-ed by the kernels, cleared when the stats are read
    a.bar = c->d_bar;
    a.stats = c->d_stats;
    a.trace = trace ? c->d_trace : nullptr;
    a.trace_cap = c->trace_cap;
    a.trace_n = trace ? c->d_flags + 1 : nullptr;
    CU(cudaMemsetAsync(c->d_bar, 0, sizeof(GridBar), c->stream));
    { const int rc_epoch = next_claim_epoch(c); if (rc_epoch) return rc_epoch; }
    a.claim_epoch = c->claim_epoch;
    // every strip's reset / spawn is complete before anybody's first phase touches it
    peer_handshake_kernel<<<1, 1, 0, c->stream>>>(a.pv, a.pv.tag_base, reinterpret_cast<unsigned*>(c->d_flags + 2));
    c->launches++;
    void* args[] = {&a};
    const int grid = (int)std::max<size_t>(1, (n + ls.block - 1) / ls.block);
    CU(launch_descend(c, ls.kernel, grid, ls.block, args, ls.smem(ls.block), 1));
    c->launches++;
    return SHX_OK;
  }
  if (n == 0) return SHX_OK;
  c->tracks_clean = false;
  if (sequential(c)) {
    SequentialArgs a;
    a.m = c->m;
    a.P = step_params(c->p);
    a.drops = c->d_drops;
    a.ndrops = (unsigned)n;
    a.stats = c->d_stats;
    a.trace = trace ? c->d_trace : nullptr;
    a.trace_cap = c->trace_cap;
    a.trace_n = trace ? c->d_flags + 1 : nullptr;
    descend_sequential_kernel<<<1, 1, 0, c->stream>>>(a);
    c->launches++;
    CU(cudaGetLastError());
    return SHX_OK;
  }
  // batched: sub-batches of at most `capacity` co-resident threads, in list order
  size_t done = 0;
  while (done < n) {
    const size_t left = n - done;
    DescendArgs a;
    a.m = c->m;
    a.P = step_params(c->p);
    a.drops = c->d_drops + done;
    a.align_age = align_age ? 1u : 0u;
    a.free_waits = (unsigned)c->cfg.free_waits;
    a.abort_flag = c->d_flags + 2;  // OR-ed by the kernels (a later clean launch cannot clear it), reset by fetch_stats
    a.bar = c->d_bar;
    a.stats = c->d_stats;
    a.trace = (trace && done == 0) ? c->d_trace : nullptr;
    a.trace_cap = c->trace_cap;
    a.trace_n = (trace && done == 0) ? c->d_flags + 1 : nullptr;
    CU(cudaMemsetAsync(c->d_bar, 0, sizeof(GridBar), c->stream));
    { const int rc_epoch = next_claim_epoch(c); if (rc_epoch) return rc_epoch; }
    a.claim_epoch = c->claim_epoch;
    void* args[] = {&a};
    size_t take;
    if (left <= 16 && !c->forced_shape) {
      // a handful of drops: one CTA of eight lanes per drop, the per-phase barrier is a plain __syncthreads.  (Only
      // for a CTA of up to 128 threads: 128 drops in ONE CTA of 1024 threads measured 4.73 us per phase, the same
      // drops in 16 CTAs on 16 SMs behind the grid barrier 4.20.)
      take = left;
      a.ndrops = (unsigned)take;
      const int block = (int)((take * 8 + 31) / 32 * 32);
      CU(launch_descend(c, (const void*)KERNEL_SMALL, 1, block, args, (size_t)kGroupSmemWords * 4 * block, 8));
    } else {
      // by batch size (a constant table, not a property of the device: every shape gives the same bits anyway)
      int pick = left <= kShapeLimit[0] ? 0 : (left <= kShapeLimit[1] ? 1 : (left <= kShapeLimit[2] ? 2 : 3));
      if (pick == 0 && c->shape[0].lanes == 1) pick = 1;  // peer mode has no eight-lane form
      while (pick < 3 && left > (size_t)c->shape[pick].cap_blocks * c->shape[pick].drops_per_block(c->shape[pick].block)) pick++;
      const LaunchShape& ls = c->shape[c->forced_shape ? 0 : pick];
      const int block = ls.block;
      const size_t per_block = ls.drops_per_block(block);
      // A list longer than one launch marches as consecutive launches (each sees the heights the earlier ones left).
      // The split must not depend on the device: kDropsPerLaunch is a constant of the library (one 8192^2 cycle), and
      // a device that cannot keep that many drops co-resident reports SHX_ERR_CAPACITY instead of splitting
      // elsewhere.  shx_config.grid_blocks (tests) forces a smaller split on purpose.
      size_t chunk = kDropsPerLaunch;
      if (c->cfg.grid_blocks > 0) chunk = std::min(chunk, (size_t)c->cfg.grid_blocks * per_block);
      take = std::min(left, chunk);
      if (take > (size_t)ls.cap_blocks * per_block)
        return fail(SHX_ERR_CAPACITY, "this device cannot keep a launch's drops co-resident (131072 per launch)");
      a.ndrops = (unsigned)take;
      const int grid = (int)((take + per_block - 1) / per_block);
      CU(launch_descend(c, ls.kernel, grid, block, args, ls.smem(block), ls.lanes));
    }
    c->launches++;
    done += take;
  }
  return SHX_OK;
}

static int batch_cycles(const shx_ctx* c) { return c->cfg.max_cycles_per_launch > 0 ? c->cfg.max_cycles_per_launch : 512; }

static int spawn_device(shx_ctx* c, int cycles, uint64_t seed, uint64_t epoch, size_t* n_out, int i0 = 0) {
  const int ts = c->p.tilesize, ms = c->p.mapsize;
  if (cycles < 0) return fail(SHX_ERR_ARG, "negative cycles");
  if (c->m.row0 % ts || c->m.row1 % ts) return fail(SHX_ERR_ARG, "spawning needs tile-aligned strips");
  const unsigned node0 = (unsigned)(c->m.row0 / ts) * ms, nnodes = (unsigned)((c->m.row1 - c->m.row0) / ts) * ms;
  const size_t n = (size_t)nnodes * cycles;
  if (n > c->max_drops) return fail(SHX_ERR_CAPACITY, "cycles*nodes exceeds max_drops");
  *n_out = n;
  if (!n) return SHX_OK;
  SpawnArgs a;
  a.m = c->m;
  a.sequential = sequential(c) ? 1 : 0;
  a.tilesize = ts; a.mapsize = ms;
  a.node0 = node0; a.nnodes = nnodes; a.cycles = cycles; a.i0 = i0;
  a.key = mix64(mix64(seed) + epoch);
  a.drops = c->d_drops;
  a.xy = c->d_xy;
  a.stats = c->d_stats;
  spawn_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(a);
  c->launches++;
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_launch_info(const shx_ctx* c, int* grid, int* block, int* lanes_per_drop) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  if (grid) *grid = c->last_grid;
  if (block) *block = c->last_block;
  if (lanes_per_drop) *lanes_per_drop = c->last_lanes;
  return SHX_OK;
}

int shx_timing_enable(shx_ctx* c, int on) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  c->timing = on != 0;
  return SHX_OK;
}

int shx_timing_read(shx_ctx* c, shx_timing* out) {
  if (!c || !out) return fail(SHX_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaStreamSynchronize(c->stream));
  memset(out, 0, sizeof(*out));
  for (auto& s : c->spans) {
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, s.e0, s.e1));
    if (s.kind == 0) out->spawn_ms += ms;
    else if (s.kind == 1) { out->descend_ms += ms; out->descend_launches++; }
    else if (s.kind == 2) out->ema_ms += ms;
    else if (s.kind == 3) out->pack_ms += ms;
    else if (s.kind == 4) out->d2h_ms += ms;
    else out->push_ms += ms;
    cudaEventDestroy(s.e0);
    cudaEventDestroy(s.e1);
  }
  c->spans.clear();
  out->scatter_ms = c->scatter_ms;
  c->scatter_ms = 0.0;
  return SHX_OK;
}

int shx_erode_async(shx_ctx* c, int cycles, uint64_t seed) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  int rc = begin_call(c);
  if (rc) return rc;
  if (!c->tracks_clean && (rc = shx_reset_tracks(c))) return rc;  // world.h:56-61
  // world.h:64-76 in batches of at most max_cycles_per_launch drops per node (sequential mode marches
  // the drops one by one anyway).  A peer-mode rank always launches once, even with nothing to spawn.
  const int cap = sequential(c) ? std::max(cycles, 1) : batch_cycles(c);
  for (int i0 = 0; i0 < cycles || i0 == 0; i0 += cap) {
    const int sub = std::max(0, std::min(cap, cycles - i0));
    size_t n = 0;
    if ((rc = span_begin(c, 0))) return rc;
    if ((rc = spawn_device(c, sub, seed, c->epoch, &n, i0))) return rc;
    if ((rc = span_end(c))) return rc;
    if ((rc = span_begin(c, 1))) return rc;
    if ((rc = run_device_drops(c, n, false))) return rc;
    if ((rc = span_end(c))) return rc;
  }
  c->epoch++;
  if ((rc = span_begin(c, 2))) return rc;
  if ((rc = ema_launch(c, !c->cfg.keep_tracks))) return rc;  // world.h:81-86
  return span_end(c);
}

int shx_read_stats(shx_ctx* c, shx_stats* out) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  CU(cudaSetDevice(c->cfg.device));
  return fetch_stats(c, out);
}

int shx_erode(shx_ctx* c, int cycles, uint64_t seed, shx_stats* out) {
  int rc = shx_erode_async(c, cycles, seed);
  if (rc) return rc;
  return fetch_stats(c, out);
}

static int make_drops_from_xy(shx_ctx* c, const float* xy_host, size_t n) {
  if (n > c->max_drops) return fail(SHX_ERR_CAPACITY, "more drops than max_drops");
  if (!n) return SHX_OK;
  CU(cudaMemcpyAsync(c->d_xy, xy_host, n * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  make_drops_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(c->d_xy, (unsigned)n, c->m, sequential(c) ? 1 : 0, c->d_drops, c->d_stats);
  c->launches++;
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_erode_spawnlist(shx_ctx* c, const float* xy, size_t n, shx_stats* out) {
  if (!c || (!xy && n)) return fail(SHX_ERR_ARG, "null argument");
  int rc = begin_call(c);
  if (rc) return rc;
  if (!c->tracks_clean && (rc = shx_reset_tracks(c))) return rc;
  if ((rc = make_drops_from_xy(c, xy, n))) return rc;
  if ((rc = run_device_drops(c, n, false))) return rc;
  if ((rc = ema_launch(c, !c->cfg.keep_tracks))) return rc;
  return fetch_stats(c, out);
}

int shx_trace_drop(shx_ctx* c, float x, float y, float* trace7, int max_steps, int* nsteps) {
  if (!c || !trace7 || !nsteps || max_steps < 1) return fail(SHX_ERR_ARG, "bad argument");
  int rc = begin_call(c);
  if (rc) return rc;
  CU(cudaMemsetAsync(c->d_flags + 1, 0, sizeof(int), c->stream));
  // no rejection here: Drop(pos) followed by while(descend()) exactly as a caller of the reference would
  const shx_drop d = {x, y, 0.0f, 0.0f, 1.0f, 0.0f, 0, SHX_DROP_ALIVE};
  const int ix = (int)x, iy = (int)y;
  if (!(x > -1.0f) || !(y > -1.0f) || ix >= c->size || iy >= c->size || ix < c->m.row0 || ix >= c->m.row1) {
    *nsteps = 0;  // water.h:62-68: NULL node -> descend returns false immediately
    return SHX_OK;
  }
  CU(cudaMemcpyAsync(c->d_drops, &d, sizeof d, cudaMemcpyHostToDevice, c->stream));
  if ((rc = run_device_drops(c, 1, true))) return rc;
  CU(cudaMemcpyAsync(c->h_flags + 1, c->d_flags + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const int n = std::min(c->h_flags[1], max_steps);
  CU(cudaMemcpy(trace7, c->d_trace, (size_t)n * 7 * sizeof(float), cudaMemcpyDeviceToHost));
  *nsteps = n;
  return SHX_OK;
}

int shx_spawn(shx_ctx* c, int cycles, uint64_t seed, uint64_t epoch, float* xy_out, size_t* n_out) {
  if (!c || !n_out) return fail(SHX_ERR_ARG, "null argument");
  int rc = begin_call(c);
  if (rc) return rc;
  size_t n = 0;
  if ((rc = spawn_device(c, cycles, seed, epoch, &n))) return rc;
  *n_out = n;
  if (xy_out && n) CU(cudaMemcpyAsync(xy_out, c->d_xy, n * 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

int shx_run_drops(shx_ctx* c, shx_drop* drops, size_t n, shx_stats* out) {
  if (!c || (!drops && n)) return fail(SHX_ERR_ARG, "null argument");
  if (n > c->max_drops) return fail(SHX_ERR_CAPACITY, "more drops than max_drops");
  int rc = begin_call(c);
  if (rc) return rc;
  if (n) CU(cudaMemcpyAsync(c->d_drops, drops, n * sizeof(shx_drop), cudaMemcpyHostToDevice, c->stream));
  if ((rc = run_device_drops(c, n, false))) return rc;
  if (n) CU(cudaMemcpyAsync(drops, c->d_drops, n * sizeof(shx_drop), cudaMemcpyDeviceToHost, c->stream));
  return fetch_stats(c, out);
}

static int push_rootdensity(shx_ctx* c, const int* xy, const float* delta, size_t n, bool absolute);

int shx_add_rootdensity(shx_ctx* c, const int* xy, const float* delta, size_t n) { return push_rootdensity(c, xy, delta, n, false); }
int shx_set_rootdensity(shx_ctx* c, const int* xy, const float* value, size_t n) { return push_rootdensity(c, xy, value, n, true); }

static int push_rootdensity(shx_ctx* c, const int* xy, const float* delta, size_t n, bool absolute) {
  if (!c || ((!xy || !delta) && n)) return fail(SHX_ERR_ARG, "null argument");
  if (!n) return SHX_OK;
  CU(cudaSetDevice(c->cfg.device));
  const size_t b_xy = n * 2 * sizeof(int), b_val = n * sizeof(float);
  if (c->push_cap < b_xy + b_val) {  // grow-only buffers
    if (c->push_done) CU(cudaEventSynchronize(c->push_done));
    cudaFree(c->d_push);
    if (c->h_push) cudaFreeHost(c->h_push);
    c->d_push = c->h_push = nullptr;
    c->push_cap = 0;
    const size_t cap = std::max<size_t>(2 * (b_xy + b_val), 1 << 16);
    if (cudaMalloc((void**)&c->d_push, cap) != cudaSuccess || cudaMallocHost((void**)&c->h_push, cap) != cudaSuccess) {
      cudaGetLastError();
      return fail(SHX_ERR_NOMEM, "allocation of the rootdensity push buffers failed");
    }
    c->push_cap = cap;
    if (!c->push_done) CU(cudaEventCreateWithFlags(&c->push_done, cudaEventDisableTiming));
  } else {
    CU(cudaEventSynchronize(c->push_done));  // the previous push has left the bounce buffer
  }
  memcpy(c->h_push, xy, b_xy);  // the caller's arrays are free again when this call returns
  memcpy(c->h_push + b_xy, delta, b_val);
  { const int rc_s = span_begin(c, 5); if (rc_s) return rc_s; }
  CU(cudaMemcpyAsync(c->d_push, c->h_push, b_xy + b_val, cudaMemcpyHostToDevice, c->stream));
  CU(cudaEventRecord(c->push_done, c->stream));
  const int* d_xy = reinterpret_cast<const int*>(c->d_push);
  const float* d_val = reinterpret_cast<const float*>(c->d_push + b_xy);
  c->root_counts_valid = false;
  if (absolute) set_rootdensity_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(c->m, d_xy, d_val, n);
  else add_rootdensity_kernel<<<1, 1, 0, c->stream>>>(c->m, d_xy, d_val, n);
  c->launches++;
  CU(cudaGetLastError());
  return span_end(c);  // stream-ordered: the next erode / download on this context sees the values
}

int shx_synth_terrain(shx_ctx* c, uint32_t seed) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  CU(cudaSetDevice(c->cfg.device));
  const unsigned init[2] = {0xffffffffu, 0u};
  CU(cudaMemcpyAsync(c->d_u32, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
  synth_minmax_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->size, seed, c->d_u32);
  synth_fill_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->m, sequential(c) ? 1 : 0, seed, c->d_u32);
  c->launches += 2;
  c->tracks_clean = true;
  c->root_counts_valid = false;
  CU(cudaGetLastError());
  int rc = refresh_halo_ref(c);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

int shx_init_terrain(shx_ctx* c, int seed) {  // map::init, cellpool.h:349-409
  if (!c) return fail(SHX_ERR_ARG, "null context");
  CU(cudaSetDevice(c->cfg.device));
  const unsigned zero = 0x80000000u;  // f2ord(0.0f): both extremes start at 0 (cellpool.h:382-383)
  const unsigned init[2] = {zero, zero};
  CU(cudaMemcpyAsync(c->d_u32, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
  terrain_raw_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->m, c->p.tilesize, seed, c->d_u32);
  terrain_fill_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->m, sequential(c) ? 1 : 0, c->d_u32);
  c->launches += 2;
  c->tracks_clean = true;
  c->root_counts_valid = false;
  CU(cudaGetLastError());
  int rc = refresh_halo_ref(c);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

int shx_measure_read_bandwidth(shx_ctx* c, size_t bytes, int passes, double* gbs) {
  if (!c || !gbs || bytes < 4096 || passes < 1) return fail(SHX_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->cfg.device));
  int4* buf = nullptr;
  const size_t n = bytes / sizeof(int4);
  if (cudaMalloc((void**)&buf, n * sizeof(int4)) != cudaSuccess) {
    cudaGetLastError();
    return fail(SHX_ERR_NOMEM, "cudaMalloc failed for the bandwidth probe");
  }
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaMemsetAsync(buf, 0, n * sizeof(int4), c->stream));
  const int grid = c->sm_count * 8;
  read_bandwidth_kernel<<<grid, 256, 0, c->stream>>>(buf, n, 2, c->d_flags + 3);  // warm: the buffer is in L2 if it fits
  float best = 0.0f;
  for (int rep = 0; rep < 5; rep++) {
    CU(cudaEventRecord(e0, c->stream));
    read_bandwidth_kernel<<<grid, 256, 0, c->stream>>>(buf, n, passes, c->d_flags + 3);
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep == 0 || ms < best) best = ms;
  }
  CU(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *gbs = (double)n * sizeof(int4) * passes / (best * 1e-3) / 1e9;
  return SHX_OK;
}

// ------------------------------------------------------------------------------- row strips

static int strip_check(shx_ctx* c) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  if (sequential(c)) return fail(SHX_ERR_MODE, "strips need the batched mode");
  CU(cudaSetDevice(c->cfg.device));
  return SHX_OK;
}

// offsets (in cells) of the halo band and of the owned edge band on each side of a strip
static size_t band_halo(const shx_ctx* c, int side) { return side == 0 ? 0 : (size_t)(c->m.row1 - c->m.xlo) * c->size; }
static size_t band_edge(const shx_ctx* c, int side, int rows) {
  return side == 0 ? (size_t)(c->m.row0 - c->m.xlo) * c->size : (size_t)(c->m.row1 - rows - c->m.xlo) * c->size;
}

int shx_strip_pack_halo_delta(shx_ctx* c, int32_t* dev_lo, int32_t* dev_hi) {
  int rc = strip_check(c);
  if (rc) return rc;
  for (int side = 0; side < 2; side++) {
    const int rows = side == 0 ? c->halo_lo : c->halo_hi;
    int32_t* out = side == 0 ? dev_lo : dev_hi;
    if (!rows || !out) continue;
    const size_t n = (size_t)rows * c->size;
    strip_halo_delta_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(c->m.hq + band_halo(c, side), c->d_halo_ref[side], out, n);
    c->launches++;
  }
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_strip_apply_halo_delta(shx_ctx* c, const int32_t* from_lo, const int32_t* from_hi) {
  int rc = strip_check(c);
  if (rc) return rc;
  // the lower neighbour's hi-halo covers our first rows; the upper neighbour's lo-halo our last rows
  for (int side = 0; side < 2; side++) {
    const int rows = side == 0 ? c->halo_lo : c->halo_hi;  // symmetric halos
    const int32_t* in = side == 0 ? from_lo : from_hi;
    if (!rows || !in) continue;
    const size_t n = (size_t)rows * c->size;
    strip_add_rows_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(c->m.hq + band_edge(c, side, rows), in, n);
    c->launches++;
  }
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_strip_pack_boundary(shx_ctx* c, int32_t* dev_lo, int32_t* dev_hi) {
  int rc = strip_check(c);
  if (rc) return rc;
  for (int side = 0; side < 2; side++) {
    const int rows = side == 0 ? c->halo_lo : c->halo_hi;
    int32_t* out = side == 0 ? dev_lo : dev_hi;
    if (!rows || !out) continue;
    const size_t n = (size_t)rows * c->size;
    strip_get_rows_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(c->m.hq + band_edge(c, side, rows), out, n);
    c->launches++;
  }
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_strip_set_halo(shx_ctx* c, const int32_t* dev_lo, const int32_t* dev_hi) {
  int rc = strip_check(c);
  if (rc) return rc;
  for (int side = 0; side < 2; side++) {
    const int rows = side == 0 ? c->halo_lo : c->halo_hi;
    const int32_t* in = side == 0 ? dev_lo : dev_hi;
    if (!rows || !in) continue;
    const size_t n = (size_t)rows * c->size;
    strip_set_rows_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(c->m.hq + band_halo(c, side), c->d_halo_ref[side], in, n);
    c->launches++;
  }
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_strip_pack_migrants(shx_ctx* c, shx_drop* dev_lo, shx_drop* dev_hi, size_t cap, int* n_lo, int* n_hi) {
  int rc = strip_check(c);
  if (rc) return rc;
  if (!dev_lo || !dev_hi || !n_lo || !n_hi) return fail(SHX_ERR_ARG, "null argument");
  CU(cudaMemsetAsync(c->d_u32 + 2, 0, 2 * sizeof(unsigned), c->stream));
  if (c->last_n) {
    strip_pack_migrants_kernel<<<grid_for(c, c->last_n), 256, 0, c->stream>>>(c->d_drops, (unsigned)c->last_n, dev_lo, dev_hi,
                                                                            (unsigned)cap, c->d_u32 + 2);
    c->launches++;
  }
  CU(cudaGetLastError());
  unsigned counts[2];
  CU(cudaMemcpyAsync(counts, c->d_u32 + 2, sizeof counts, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (counts[0] > cap || counts[1] > cap) return fail(SHX_ERR_CAPACITY, "migrant outbox too small");
  *n_lo = (int)counts[0];
  *n_hi = (int)counts[1];
  return SHX_OK;
}

size_t shx_strip_message_words(const shx_ctx* c, size_t cap) {
  if (!c) return 0;
  return (size_t)kMsgHeader + 8 * cap + 2 * (size_t)c->cfg.halo * c->size;
}

int shx_strip_pack_message(shx_ctx* c, int32_t* dev_lo, int32_t* dev_hi, size_t cap) {
  int rc = strip_check(c);
  if (rc) return rc;
  const size_t band = (size_t)c->cfg.halo * c->size;
  int32_t* msg[2] = {c->halo_lo ? dev_lo : nullptr, c->halo_hi ? dev_hi : nullptr};
  for (int side = 0; side < 2; side++) {
    if ((side == 0 ? c->halo_lo : c->halo_hi) && !msg[side]) return fail(SHX_ERR_ARG, "null message buffer for an existing neighbour");
    if (!msg[side]) continue;
    if ((side == 0 ? c->halo_lo : c->halo_hi) != c->cfg.halo) return fail(SHX_ERR_ARG, "messages need full halos");
    CU(cudaMemsetAsync(msg[side], 0, kMsgHeader * sizeof(int32_t), c->stream));
    int32_t* rows = msg[side] + kMsgHeader + 8 * cap;
    strip_msg_rows_kernel<<<grid_for(c, band), 256, 0, c->stream>>>(c->m.hq + band_halo(c, side), c->d_halo_ref[side],
                                                                   c->m.hq + band_edge(c, side, c->cfg.halo), rows, rows + band, band);
    c->launches++;
  }
  if (c->last_n && (msg[0] || msg[1])) {
    // a drop can only have left towards an existing neighbour; the other pointer is never written
    strip_msg_migrants_kernel<<<grid_for(c, c->last_n), 256, 0, c->stream>>>(c->d_drops, (unsigned)c->last_n, msg[0] ? msg[0] : msg[1],
                                                                           msg[1] ? msg[1] : msg[0], (unsigned)cap);
    c->launches++;
  }
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_strip_apply_message(shx_ctx* c, const int32_t* from_lo, const int32_t* from_hi, size_t cap) {
  int rc = strip_check(c);
  if (rc) return rc;
  const size_t band = (size_t)c->cfg.halo * c->size;
  const int32_t* msg[2] = {c->halo_lo ? from_lo : nullptr, c->halo_hi ? from_hi : nullptr};
  for (int side = 0; side < 2; side++) {
    if ((side == 0 ? c->halo_lo : c->halo_hi) && !msg[side]) return fail(SHX_ERR_ARG, "null message buffer for an existing neighbour");
    if (!msg[side]) continue;
    const int32_t* rows = msg[side] + kMsgHeader + 8 * cap;
    strip_msg_apply_kernel<<<grid_for(c, band), 256, 0, c->stream>>>(c->m.hq + band_halo(c, side), c->d_halo_ref[side],
                                                                    c->m.hq + band_edge(c, side, c->cfg.halo), rows, rows + band, band);
    c->launches++;
  }
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_strip_run_device_drops(shx_ctx* c, const shx_drop* dev_drops, size_t n, shx_stats* out) {
  int rc = strip_check(c);
  if (rc) return rc;
  if (n > c->max_drops) return fail(SHX_ERR_CAPACITY, "more drops than max_drops");
  if (!c->strip_open && (rc = begin_call(c))) return rc;  // inside begin..end the counters keep accumulating
  if (n) CU(cudaMemcpyAsync(c->d_drops, dev_drops, n * sizeof(shx_drop), cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = span_begin(c, 1))) return rc;
  if ((rc = run_device_drops(c, n, false))) return rc;
  if ((rc = span_end(c))) return rc;
  if (out) return fetch_stats(c, out);
  return SHX_OK;
}

int shx_strip_erode_begin_with(shx_ctx* c, int cycles, uint64_t seed, const shx_drop* dev_carried, size_t n_carried) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  if (n_carried && !dev_carried) return fail(SHX_ERR_ARG, "null carried drops");
  if (cycles > batch_cycles(c)) return fail(SHX_ERR_ARG, "a strip call takes at most max_cycles_per_launch cycles (migrants are packed once per call)");
  int rc = begin_call(c);
  if (rc) return rc;
  c->strip_open = true;
  if (!c->tracks_clean && (rc = shx_reset_tracks(c))) return rc;  // world.h:56-61
  size_t n = 0;
  if ((rc = span_begin(c, 0))) return rc;
  if ((rc = spawn_device(c, cycles, seed, c->epoch, &n))) return rc;  // world.h:64-74
  if (n + n_carried > c->max_drops) return fail(SHX_ERR_CAPACITY, "spawned + carried drops exceed max_drops");
  if (n_carried)  // drops handed over by the neighbours at the end of the previous call march with this call's batch
    CU(cudaMemcpyAsync(c->d_drops + n, dev_carried, n_carried * sizeof(shx_drop), cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = span_end(c))) return rc;
  c->epoch++;
  if ((rc = span_begin(c, 1))) return rc;
  if ((rc = run_device_drops(c, n + n_carried, false, n_carried > 0))) return rc;  // world.h:76
  return span_end(c);
}

int shx_strip_erode_begin(shx_ctx* c, int cycles, uint64_t seed) { return shx_strip_erode_begin_with(c, cycles, seed, nullptr, 0); }

int shx_strip_erode_end(shx_ctx* c) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  CU(cudaSetDevice(c->cfg.device));
  c->strip_open = false;
  int rc = span_begin(c, 2);
  if (rc) return rc;
  if ((rc = ema_launch(c, !c->cfg.keep_tracks))) return rc;  // world.h:81-86
  return span_end(c);
}

// ------------------------------------------------------------------------------- peer mode

int shx_peer_export(shx_ctx* c, shx_peer_handles* out) {
  if (!c || !out) return fail(SHX_ERR_ARG, "null argument");
  if (!c->peer) return fail(SHX_ERR_MODE, "context was not created with peer_world > 1");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU(cudaSetDevice(c->cfg.device));
  // An IPC handle names the whole allocation a pointer lives in and opens at that allocation's
  // base; small cudaMalloc blocks are sub-allocated, so the offset has to travel with the handle.
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CU(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
  if (!fn) return fail(SHX_ERR_CUDA, "cuMemGetAddressRange is not available");
  const void* ptr[3] = {c->m.hq, c->m.rec, c->d_inbox};
  unsigned char* dst[3] = {out->hq, out->rec, out->inbox};
  uint64_t* off[3] = {&out->off_hq, &out->off_rec, &out->off_inbox};
  for (int k = 0; k < 3; k++) {
    unsigned long long base = 0;
    size_t size = 0;
    if (reinterpret_cast<range_fn>(fn)(&base, &size, (unsigned long long)ptr[k]) != 0)
      return fail(SHX_ERR_CUDA, "cuMemGetAddressRange failed");
    *off[k] = (unsigned long long)ptr[k] - base;
    CU(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(dst[k]), const_cast<void*>(ptr[k])));
  }
  return SHX_OK;
}

int shx_peer_attach(shx_ctx* c, const shx_peer_handles* all) {
  if (!c || !all) return fail(SHX_ERR_ARG, "null argument");
  if (!c->peer) return fail(SHX_ERR_MODE, "context was not created with peer_world > 1");
  if (c->peer_attached) return fail(SHX_ERR_ARG, "already attached");
  CU(cudaSetDevice(c->cfg.device));
  for (int r = 0; r < c->pv.nranks; r++) {
    if (r == c->pv.rank) continue;
    char* p[3] = {nullptr, nullptr, nullptr};
    const unsigned char* h[3] = {all[r].hq, all[r].rec, all[r].inbox};
    const uint64_t off[3] = {all[r].off_hq, all[r].off_rec, all[r].off_inbox};
    for (int k = 0; k < 3; k++) {
      cudaIpcMemHandle_t handle;
      memcpy(&handle, h[k], sizeof handle);
      // the same allocation may back several buffers: it can only be opened once per process
      for (int j = 0; j < k && !p[k]; j++)
        if (memcmp(h[j], h[k], sizeof handle) == 0) p[k] = p[j];
      if (!p[k]) {
        void* base = nullptr;
        CU(cudaIpcOpenMemHandle(&base, handle, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened[3 * r + k] = base;
        p[k] = static_cast<char*>(base);
      }
    }
    c->pv.hq[r] = reinterpret_cast<int4*>(p[0] + off[0]);
    c->pv.rec[r] = reinterpret_cast<CellRec*>(p[1] + off[1]);
    c->pv.inbox[r] = reinterpret_cast<unsigned long long*>(p[2] + off[2]);
  }
  c->peer_attached = true;
  return SHX_OK;
}

// ------------------------------------------------------------------------------- N3: vegetation on the device

void shx_default_plant_params(shx_plant_params* pp) {
  if (!pp) return;
  pp->maxSize = 1.5f;        // vegetation.h:40
  pp->growRate = 0.05f;      // :41
  pp->maxSteep = 0.8f;       // :42
  pp->maxDischarge = 0.3f;   // :43
  pp->maxTreeHeight = 0.8f;  // :44
}

static int veg_ready(shx_ctx* c) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  if (!c->veg) return fail(SHX_ERR_MODE, "shx_veg_create has not been called on this context");
  return SHX_OK;
}

int shx_veg_set_params(shx_ctx* c, const shx_plant_params* pp) {
  if (!c || !pp) return fail(SHX_ERR_ARG, "null argument");
  c->plant = PlantParams{pp->maxSize, pp->growRate, pp->maxSteep, pp->maxDischarge, pp->maxTreeHeight};
  return SHX_OK;
}

int shx_veg_create(shx_ctx* c, size_t max_plants, const shx_plant_params* pp) {
  if (!c) return fail(SHX_ERR_ARG, "null context");
  if (c->veg) return fail(SHX_ERR_ARG, "the plant store exists already");
  if (c->peer || c->m.row0 != 0 || c->m.row1 != c->size) return fail(SHX_ERR_MODE, "vegetation on the device needs a whole-map context");
  CU(cudaSetDevice(c->cfg.device));
  shx_plant_params def;
  shx_default_plant_params(&def);
  shx_veg_set_params(c, pp ? pp : &def);
  const size_t cap = max_plants ? max_plants : std::max<size_t>(1024, c->stored_cells / 4);
  if (cap > 0x7fffffffu) return fail(SHX_ERR_ARG, "max_plants too large");
  const size_t slots = cap + 1, blocks = (slots + kVegBlock - 1) / kVegBlock;
  bool ok = true;
  for (int i = 0; i < 2; i++)
    ok = ok && cudaMalloc((void**)&c->d_plant_pos[i], cap * sizeof(int2)) == cudaSuccess &&
         cudaMalloc((void**)&c->d_plant_size[i], cap * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&c->d_veg_flags, slots * sizeof(unsigned)) == cudaSuccess &&
       cudaMalloc((void**)&c->d_veg_child, slots * sizeof(int2)) == cudaSuccess &&
       cudaMalloc((void**)&c->d_veg_grown, slots * sizeof(float)) == cudaSuccess &&
       cudaMalloc((void**)&c->d_veg_blocks, blocks * sizeof(uint2)) == cudaSuccess &&
       cudaMalloc((void**)&c->d_veg_totals, 8 * sizeof(unsigned)) == cudaSuccess &&
       cudaMallocHost((void**)&c->h_veg_totals, 8 * sizeof(unsigned)) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    return fail(SHX_ERR_NOMEM, "allocation of the plant store failed");
  }
  c->veg_cap = cap;
  c->veg_n = 0;
  c->veg_cur = 0;
  c->veg = true;
  return SHX_OK;
}

// the integer root counts follow the fp32 rootdensity after an upload / a host push / a new terrain
static int veg_sync_counts(shx_ctx* c) {
  if (c->root_counts_valid) return SHX_OK;
  veg_count_from_density_kernel<<<grid_for(c, c->stored_cells), 256, 0, c->stream>>>(c->m.rec, c->stored_cells);
  c->launches++;
  CU(cudaGetLastError());
  c->root_counts_valid = true;
  return SHX_OK;
}

static VegArgs veg_args(shx_ctx* c, uint64_t key) {
  VegArgs a;
  a.v = view_args(c);
  a.pp = c->plant;
  a.key = key;
  a.pos = c->d_plant_pos[c->veg_cur];
  a.size = c->d_plant_size[c->veg_cur];
  a.n = (unsigned)c->veg_n;
  a.pos_out = c->d_plant_pos[c->veg_cur ^ 1];
  a.size_out = c->d_plant_size[c->veg_cur ^ 1];
  a.cap = (unsigned)c->veg_cap;
  a.flags = c->d_veg_flags;
  a.child = c->d_veg_child;
  a.grown = c->d_veg_grown;
  a.block_counts = c->d_veg_blocks;
  a.totals = c->d_veg_totals;
  return a;
}

int shx_veg_grow(shx_ctx* c, uint64_t seed, uint64_t frame, shx_veg_stats* out) {
  int rc = veg_ready(c);
  if (rc) return rc;
  CU(cudaSetDevice(c->cfg.device));
  rc = veg_sync_counts(c);
  if (rc) return rc;
  const VegArgs a = veg_args(c, mix64(mix64(seed) + frame));
  const unsigned slots = a.n + 1u, blocks = (slots + kVegBlock - 1) / kVegBlock;
  veg_decide_kernel<<<blocks, kVegBlock, 0, c->stream>>>(a);
  veg_scan_kernel<<<1, 1024, 0, c->stream>>>(a, blocks);
  veg_apply_kernel<<<blocks, kVegBlock, 0, c->stream>>>(a);
  veg_refresh_kernel<<<blocks, kVegBlock, 0, c->stream>>>(a);
  c->launches += 4;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_veg_totals, c->d_veg_totals, 5 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const unsigned* t = c->h_veg_totals;  // [0] survivors [1] born [2] died [3] refused [4] new count
  c->veg_n = t[4];
  c->veg_cur ^= 1;
  if (out) {
    out->plants = t[4];
    out->born = t[1];
    out->died = t[2];
    out->refused = t[3];
  }
  if (t[3]) return fail(SHX_ERR_CAPACITY, "the plant list is full (max_plants of shx_veg_create): children were refused");
  return SHX_OK;
}

int shx_veg_count(shx_ctx* c, size_t* n) {
  int rc = veg_ready(c);
  if (rc) return rc;
  if (!n) return fail(SHX_ERR_ARG, "null argument");
  *n = c->veg_n;
  return SHX_OK;
}

int shx_veg_download(shx_ctx* c, float* xys3, size_t cap, size_t* n) {
  int rc = veg_ready(c);
  if (rc) return rc;
  if (n) *n = c->veg_n;
  if (c->veg_n > cap) return fail(SHX_ERR_CAPACITY, "buffer smaller than the plant list");
  if (!c->veg_n) return SHX_OK;
  if (!xys3) return fail(SHX_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->cfg.device));
  rc = view_staging(c, c->veg_n * 3 * sizeof(float));
  if (rc) return rc;
  veg_export_kernel<<<grid_for(c, c->veg_n), 256, 0, c->stream>>>(c->d_plant_pos[c->veg_cur], c->d_plant_size[c->veg_cur], (unsigned)c->veg_n, c->d_view);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(xys3, c->d_view, c->veg_n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

int shx_veg_upload(shx_ctx* c, const float* xys3, size_t n, int stamp_roots) {
  int rc = veg_ready(c);
  if (rc) return rc;
  if (n > c->veg_cap) return fail(SHX_ERR_CAPACITY, "more plants than max_plants");
  if (n && !xys3) return fail(SHX_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->cfg.device));
  std::vector<int2> pos(n);
  std::vector<float> size(n);
  for (size_t i = 0; i < n; i++) {
    const int x = (int)xys3[3 * i], y = (int)xys3[3 * i + 1];
    if (x < 0 || y < 0 || x >= c->size || y >= c->size) return fail(SHX_ERR_ARG, "a plant lies outside the map");
    pos[i] = make_int2(x, y);
    size[i] = xys3[3 * i + 2];
  }
  c->veg_cur = 0;
  c->veg_n = n;
  if (n) {
    CU(cudaMemcpyAsync(c->d_plant_pos[0], pos.data(), n * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_plant_size[0], size.data(), n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  if (stamp_roots && n) {
    rc = veg_sync_counts(c);
    if (rc) return rc;
    // every uploaded plant as a "child" of an otherwise empty frame: apply stamps it, refresh writes the fp32 values
    VegArgs a = veg_args(c, 0);
    std::vector<unsigned> fl(n + 1, 2u);
    fl[n] = 0u;
    pos.push_back(make_int2(0, 0));
    CU(cudaMemcpyAsync(c->d_veg_flags, fl.data(), (n + 1) * sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_veg_child, pos.data(), (n + 1) * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    const unsigned blocks = ((unsigned)n + 1u + kVegBlock - 1) / kVegBlock;
    veg_stamp_list_kernel<<<blocks, kVegBlock, 0, c->stream>>>(a);
    veg_refresh_kernel<<<blocks, kVegBlock, 0, c->stream>>>(a);
    c->launches += 2;
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(c->stream));  // the host vectors go out of scope
  return SHX_OK;
}

int shx_veg_tree_models(shx_ctx* c, float* dev_out16, size_t cap, size_t* n) {
  int rc = veg_ready(c);
  if (rc) return rc;
  if (n) *n = c->veg_n;
  if (c->veg_n > cap) return fail(SHX_ERR_CAPACITY, "buffer smaller than the plant list");
  if (!c->veg_n) return SHX_OK;
  if (!dev_out16) return fail(SHX_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->cfg.device));
  veg_models_kernel<<<grid_for(c, c->veg_n), 256, 0, c->stream>>>(view_args(c), c->d_plant_pos[c->veg_cur], c->d_plant_size[c->veg_cur],
                                                                   (unsigned)c->veg_n, reinterpret_cast<float4*>(dev_out16));
  c->launches++;
  CU(cudaGetLastError());
  return SHX_OK;
}

int shx_veg_tree_models_download(shx_ctx* c, float* host_out16, size_t cap, size_t* n) {
  int rc = veg_ready(c);
  if (rc) return rc;
  if (n) *n = c->veg_n;
  if (c->veg_n > cap) return fail(SHX_ERR_CAPACITY, "buffer smaller than the plant list");
  if (!c->veg_n) return SHX_OK;
  if (!host_out16) return fail(SHX_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->cfg.device));
  rc = view_staging(c, c->veg_n * 16 * sizeof(float));
  if (rc) return rc;
  rc = shx_veg_tree_models(c, c->d_view, cap, nullptr);
  if (rc) return rc;
  CU(cudaMemcpyAsync(host_out16, c->d_view, c->veg_n * 16 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SHX_OK;
}

#ifdef SHX_PHASE_TIMING
int shx_debug_starts(unsigned long long* ns8192) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(ns8192, g_start, 8192 * sizeof(unsigned long long)));
  return SHX_OK;
}
int shx_debug_arrivals(unsigned long long* ns8192, unsigned* sm8192) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(ns8192, g_arrival, 8192 * sizeof(unsigned long long)));
  CU(cudaMemcpyFromSymbol(sm8192, g_arrival_sm, 8192 * sizeof(unsigned)));
  return SHX_OK;
}
int shx_debug_set_exp(unsigned flags) {
  CU(cudaMemcpyToSymbol(g_exp, &flags, sizeof flags));
  return SHX_OK;
}
int shx_debug_phase_timing(unsigned long long* out8, int reset) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(out8, g_phase_timing, 8 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyToSymbol(g_phase_timing, z, sizeof z));
  }
  return SHX_OK;
}
#endif

}  // extern "C"
