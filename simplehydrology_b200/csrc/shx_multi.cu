// shx_multi: one world over the GPUs of a box, driven by ONE host thread (include/shx.h, "multi-GPU").
//
// The reference's caller is one process with one frame loop (SimpleHydrology.cpp:314-324); this is the form of the
// row-strip decomposition it can use: a strip context per device (shx_create with row0/row1), every call of
// World::erode runs on all strips at once, and the strips meet ONCE per call.  Nothing is staged and nothing goes
// through a collective: each strip's pack kernels write its message -- migrating drops, the integer height deltas it
// accumulated in its halo rows, its current edge rows -- straight into the neighbour's inbox over NVLink (peer stores,
// cudaDeviceEnablePeerAccess), events order pack -> apply across the devices, and the host never waits inside a
// call (the migrant counts come back through pinned memory and are read at the start of the NEXT call).
// Pairs of devices without peer access fall back to cudaMemcpyPeerAsync from a local outbox.
//
// Same protocol, same kernels and therefore the same bits as simplehydrology_b200/strips.py's erode_cycle (the
// one-process-per-GPU form bench.py uses under torchrun); tests/test_gpu_multi.py checks that.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/shx.h"

namespace {

constexpr int kHeader = 8;  // int32 words before the drop records of a strip message (kMsgHeader)

struct Strip {
  int device = 0;
  shx_ctx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  bool has[2] = {false, false};         // neighbour below (smaller x) / above
  int32_t* inbox[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [call parity][from lo / from hi], on this device
  int32_t* outbox[2] = {nullptr, nullptr};  // only for neighbours without peer access
  bool direct[2] = {true, true};        // pack straight into the neighbour's inbox
  shx_drop* carried = nullptr;          // the records taken out of the inboxes, contiguous
  int32_t* h_counts = nullptr;          // pinned: [parity][side]
  cudaEvent_t ev_packed[2] = {nullptr, nullptr};   // this strip's messages of parity q are complete (in the neighbours' inboxes)
  cudaEvent_t ev_counts[2] = {nullptr, nullptr};   // the counts of inbox parity q have reached the host
  cudaEvent_t ev_taken[2] = {nullptr, nullptr};    // inbox parity q has been read for the last time (records copied out)
  bool taken_valid[2] = {false, false};
};

thread_local std::string g_merr;

int mfail(int code, const std::string& what) {
  g_merr = what;
  return code;
}

}  // namespace

struct shx_multi {
  shx_params p;
  int n = 0;
  std::vector<Strip> s;
  size_t cap = 0, words = 0;
  uint64_t calls = 0;
  bool have_inbox = false;  // at least one exchange has happened
};

#define MCU(call)                                                                                       \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) {                                                                           \
      char b__[512];                                                                                    \
      snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return mfail(SHX_ERR_CUDA, b__);                                                                  \
    }                                                                                                   \
  } while (0)
#define MSHX(call)                                                    \
  do {                                                                \
    const int rc__ = (call);                                          \
    if (rc__ != SHX_OK) return mfail(rc__, shx_last_error());         \
  } while (0)

extern "C" {

const char* shx_multi_last_error(void) { return g_merr.c_str(); }

void shx_multi_destroy(shx_multi* m) {
  if (!m) return;
  for (Strip& st : m->s) {
    cudaSetDevice(st.device);
    if (st.stream) cudaStreamSynchronize(st.stream);
  }
  for (Strip& st : m->s) {
    cudaSetDevice(st.device);
    if (st.ctx) shx_destroy(st.ctx);
    for (int q = 0; q < 2; q++) {
      for (int side = 0; side < 2; side++) cudaFree(st.inbox[q][side]);
      if (st.ev_packed[q]) cudaEventDestroy(st.ev_packed[q]);
      if (st.ev_counts[q]) cudaEventDestroy(st.ev_counts[q]);
      if (st.ev_taken[q]) cudaEventDestroy(st.ev_taken[q]);
    }
    cudaFree(st.outbox[0]);
    cudaFree(st.outbox[1]);
    cudaFree(st.carried);
    if (st.h_counts) cudaFreeHost(st.h_counts);
    if (st.stream) cudaStreamDestroy(st.stream);
  }
  delete m;
}

int shx_multi_create(shx_multi** out, const shx_params* p, int ngpu, const int* devices, const shx_config* base) {
  if (!out || !p || ngpu < 1 || ngpu > 64) return mfail(SHX_ERR_ARG, "shx_multi_create: bad argument");
  *out = nullptr;
  if (p->mapsize % ngpu) return mfail(SHX_ERR_ARG, "the number of tile rows (mapsize) must be divisible by the number of strips");
  shx_multi* m = new (std::nothrow) shx_multi();
  if (!m) return mfail(SHX_ERR_NOMEM, "host allocation failed");
  m->p = *p;
  m->n = ngpu;
  m->s.resize(ngpu);
  const int size = p->mapsize * p->tilesize;
  const int rows = (p->mapsize / ngpu) * p->tilesize;
  const size_t nodes = (size_t)(p->mapsize / ngpu) * p->mapsize;
  // room for the drops a strip spawns plus those handed over (narrow strips carry more than they spawn)
  m->cap = std::max<size_t>(4096, 4 * nodes * 512);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    delete m;
    return mfail(SHX_ERR_CUDA, "no CUDA device: shx has no CPU fallback");
  }
  for (int i = 0; i < ngpu; i++) {
    Strip& st = m->s[i];
    st.device = devices ? devices[i] : i % ndev;
    st.has[0] = i > 0;
    st.has[1] = i < ngpu - 1;
    shx_config cfg;
    if (base) cfg = *base; else shx_default_config(&cfg);
    cfg.device = st.device;
    cfg.mode = SHX_MODE_BATCHED;
    cfg.peer_world = 0;
    if (ngpu > 1) {
      cfg.row0 = i * rows;
      cfg.row1 = (i + 1) * rows;
      if (cfg.halo < 2) cfg.halo = 2;
    } else {
      cfg.row0 = cfg.row1 = 0;
    }
    if (!cfg.max_drops) cfg.max_drops = m->cap;
    int rc = shx_create(&st.ctx, p, &cfg);
    if (rc != SHX_OK) {
      const std::string why = shx_last_error();
      shx_multi_destroy(m);
      return mfail(rc, why);
    }
    if (cudaSetDevice(st.device) != cudaSuccess || cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking) != cudaSuccess) {
      shx_multi_destroy(m);
      return mfail(SHX_ERR_CUDA, "cudaStreamCreate failed");
    }
    shx_set_stream(st.ctx, st.stream);
  }
  if (ngpu > 1) {
    m->words = shx_strip_message_words(m->s[0].ctx, m->cap);
    for (int i = 0; i < ngpu; i++) {
      Strip& st = m->s[i];
      cudaSetDevice(st.device);
      bool ok = true;
      for (int q = 0; q < 2 && ok; q++) {
        for (int side = 0; side < 2 && ok; side++)
          if (st.has[side]) {
            ok = cudaMalloc((void**)&st.inbox[q][side], m->words * sizeof(int32_t)) == cudaSuccess &&
                 cudaMemset(st.inbox[q][side], 0, kHeader * sizeof(int32_t)) == cudaSuccess;
          }
        ok = ok && cudaEventCreateWithFlags(&st.ev_packed[q], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&st.ev_counts[q], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&st.ev_taken[q], cudaEventDisableTiming) == cudaSuccess;
      }
      ok = ok && cudaMalloc((void**)&st.carried, m->cap * sizeof(shx_drop)) == cudaSuccess &&
           cudaMallocHost((void**)&st.h_counts, 4 * sizeof(int32_t)) == cudaSuccess;
      if (ok) memset(st.h_counts, 0, 4 * sizeof(int32_t));
      // peer access towards both neighbours (the pack kernels store into their inboxes)
      for (int side = 0; side < 2 && ok; side++) {
        if (!st.has[side]) continue;
        const int other = m->s[side == 0 ? i - 1 : i + 1].device;
        if (other == st.device) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, st.device, other);
        if (can) {
          const cudaError_t e = cudaDeviceEnablePeerAccess(other, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
          cudaGetLastError();
        }
        if (!can) {
          st.direct[side] = false;
          ok = cudaMalloc((void**)&st.outbox[side], m->words * sizeof(int32_t)) == cudaSuccess;
        }
      }
      if (!ok) {
        cudaGetLastError();
        shx_multi_destroy(m);
        return mfail(SHX_ERR_NOMEM, "allocation of the strip exchange buffers failed");
      }
    }
  }
  for (Strip& st : m->s) {
    cudaSetDevice(st.device);
    cudaDeviceSynchronize();
  }
  *out = m;
  return SHX_OK;
}

int shx_multi_strips(const shx_multi* m) { return m ? m->n : 0; }
shx_ctx* shx_multi_strip(shx_multi* m, int i) { return (m && i >= 0 && i < m->n) ? m->s[i].ctx : nullptr; }

static int sync_all(shx_multi* m) {
  for (Strip& st : m->s) {
    MCU(cudaSetDevice(st.device));
    MCU(cudaStreamSynchronize(st.stream));
  }
  return SHX_OK;
}

int shx_multi_sync(shx_multi* m) {
  if (!m) return mfail(SHX_ERR_ARG, "null argument");
  return sync_all(m);
}

// Every strip uploads its own rows plus halo from the caller's whole pool; the in-flight drops of earlier calls
// belong to the old map and are dropped.
int shx_multi_upload(shx_multi* m, const shx_cell* pool, size_t ncells) {
  if (!m || !pool) return mfail(SHX_ERR_ARG, "null argument");
  for (Strip& st : m->s) MSHX(shx_upload(st.ctx, pool, ncells));
  m->have_inbox = false;
  return SHX_OK;
}

int shx_multi_download(shx_multi* m, shx_cell* pool, size_t ncells, unsigned mask) {
  if (!m || !pool) return mfail(SHX_ERR_ARG, "null argument");
  for (Strip& st : m->s) MSHX(shx_download_async(st.ctx, pool, ncells, mask));  // all devices copy at once
  return sync_all(m);
}

int shx_multi_init_terrain(shx_multi* m, int seed) {
  if (!m) return mfail(SHX_ERR_ARG, "null argument");
  for (Strip& st : m->s) MSHX(shx_init_terrain(st.ctx, seed));
  m->have_inbox = false;
  return SHX_OK;
}

int shx_multi_synth_terrain(shx_multi* m, uint32_t seed) {
  if (!m) return mfail(SHX_ERR_ARG, "null argument");
  for (Strip& st : m->s) MSHX(shx_synth_terrain(st.ctx, seed));
  m->have_inbox = false;
  return SHX_OK;
}

int shx_multi_set_params(shx_multi* m, const shx_params* p) {
  if (!m || !p) return mfail(SHX_ERR_ARG, "null argument");
  for (Strip& st : m->s) MSHX(shx_set_params(st.ctx, p));
  m->p = *p;
  return SHX_OK;
}

int shx_multi_set_rootdensity(shx_multi* m, const int* xy, const float* value, size_t n) {
  if (!m) return mfail(SHX_ERR_ARG, "null argument");
  for (Strip& st : m->s) MSHX(shx_set_rootdensity(st.ctx, xy, value, n));  // a strip ignores cells outside its stored rows
  return SHX_OK;
}

// == World::erode(cycles) on all strips; no host synchronisation (counters: shx_multi_read_stats)
int shx_multi_erode_async(shx_multi* m, int cycles, uint64_t seed) {
  if (!m) return mfail(SHX_ERR_ARG, "null argument");
  if (m->n == 1) {
    MSHX(shx_erode_async(m->s[0].ctx, cycles, seed));
    m->calls++;
    return SHX_OK;
  }
  const int q = (int)(m->calls & 1u), prev = q ^ 1;
  const size_t cap = m->cap;
  // 1. every strip: take over the drops its neighbours handed in at the end of the previous call, march, EMA
  for (int i = 0; i < m->n; i++) {
    Strip& st = m->s[i];
    MCU(cudaSetDevice(st.device));
    size_t n_carried = 0;
    if (m->have_inbox) {
      MCU(cudaEventSynchronize(st.ev_counts[prev]));  // long done: the caller has read the previous call's stats
      for (int side = 0; side < 2; side++) {
        if (!st.has[side]) continue;
        const int32_t cnt = st.h_counts[2 * prev + side];
        if (cnt < 0 || (size_t)cnt > cap) return mfail(SHX_ERR_CAPACITY, "a neighbour handed over more drops than the message capacity");
        if (n_carried + (size_t)cnt > cap) return mfail(SHX_ERR_CAPACITY, "more drops in flight than the strip was sized for");
        if (cnt)
          MCU(cudaMemcpyAsync(st.carried + n_carried, st.inbox[prev][side] + kHeader, (size_t)cnt * sizeof(shx_drop),
                              cudaMemcpyDeviceToDevice, st.stream));
        n_carried += (size_t)cnt;
      }
      MCU(cudaEventRecord(st.ev_taken[prev], st.stream));
      st.taken_valid[prev] = true;
    }
    MSHX(shx_strip_erode_begin_with(st.ctx, cycles, seed, n_carried ? st.carried : nullptr, n_carried));
    MSHX(shx_strip_erode_end(st.ctx));
  }
  // 2. every strip packs ONE message per neighbour, straight into the neighbour's inbox of this call's parity
  for (int i = 0; i < m->n; i++) {
    Strip& st = m->s[i];
    MCU(cudaSetDevice(st.device));
    int32_t* dst[2] = {nullptr, nullptr};
    for (int side = 0; side < 2; side++) {
      if (!st.has[side]) continue;
      Strip& nb = m->s[side == 0 ? i - 1 : i + 1];
      // the neighbour must be done with the message that used this inbox two calls ago
      if (nb.taken_valid[q]) MCU(cudaStreamWaitEvent(st.stream, nb.ev_taken[q], 0));
      dst[side] = st.direct[side] ? nb.inbox[q][side ^ 1] : st.outbox[side];
    }
    MSHX(shx_strip_pack_message(st.ctx, dst[0], dst[1], cap));
    for (int side = 0; side < 2; side++)
      if (st.has[side] && !st.direct[side]) {
        Strip& nb = m->s[side == 0 ? i - 1 : i + 1];
        MCU(cudaMemcpyPeerAsync(nb.inbox[q][side ^ 1], nb.device, st.outbox[side], st.device, m->words * sizeof(int32_t), st.stream));
      }
    MCU(cudaEventRecord(st.ev_packed[q], st.stream));
  }
  // 3. every strip applies its neighbours' messages once they are complete; counts go to the host asynchronously
  for (int i = 0; i < m->n; i++) {
    Strip& st = m->s[i];
    MCU(cudaSetDevice(st.device));
    for (int side = 0; side < 2; side++)
      if (st.has[side]) MCU(cudaStreamWaitEvent(st.stream, m->s[side == 0 ? i - 1 : i + 1].ev_packed[q], 0));
    MSHX(shx_strip_apply_message(st.ctx, st.inbox[q][0], st.inbox[q][1], cap));
    for (int side = 0; side < 2; side++)
      if (st.has[side])
        MCU(cudaMemcpyAsync(st.h_counts + 2 * q + side, st.inbox[q][side], sizeof(int32_t), cudaMemcpyDeviceToHost, st.stream));
    MCU(cudaEventRecord(st.ev_counts[q], st.stream));
  }
  m->have_inbox = true;
  m->calls++;
  return SHX_OK;
}

// counters of the last call summed over the strips (phases: the longest strip); synchronises every strip
int shx_multi_read_stats(shx_multi* m, shx_stats* out) {
  if (!m) return mfail(SHX_ERR_ARG, "null argument");
  shx_stats sum;
  memset(&sum, 0, sizeof sum);
  for (Strip& st : m->s) {
    shx_stats one;
    MSHX(shx_read_stats(st.ctx, &one));
    sum.spawned += one.spawned; sum.rejected += one.rejected; sum.steps += one.steps;
    sum.term_age += one.term_age; sum.term_vol += one.term_vol; sum.term_oob += one.term_oob;
    sum.cascade_transfers += one.cascade_transfers;
    sum.phases = std::max(sum.phases, one.phases);
    sum.fx_eroded += one.fx_eroded; sum.fx_deposited += one.fx_deposited;
    sum.fx_sed_oob_lost += one.fx_sed_oob_lost; sum.fx_sed_deposited += one.fx_sed_deposited;
    sum.fx_sed_inflation += one.fx_sed_inflation;
    sum.migrated_lo += one.migrated_lo; sum.migrated_hi += one.migrated_hi;
    sum.launches += one.launches;
  }
  if (out) *out = sum;
  return SHX_OK;
}

int shx_multi_erode(shx_multi* m, int cycles, uint64_t seed, shx_stats* out) {
  const int rc = shx_multi_erode_async(m, cycles, seed);
  if (rc != SHX_OK) return rc;
  return shx_multi_read_stats(m, out);
}

// drops handed over at the end of the last call, waiting for the next one (sum over the strips); synchronises
int shx_multi_in_flight(shx_multi* m, size_t* n) {
  if (!m || !n) return mfail(SHX_ERR_ARG, "null argument");
  *n = 0;
  if (m->n == 1 || !m->have_inbox) return SHX_OK;
  const int q = (int)((m->calls - 1) & 1u);
  for (Strip& st : m->s) {
    MCU(cudaSetDevice(st.device));
    MCU(cudaEventSynchronize(st.ev_counts[q]));
    for (int side = 0; side < 2; side++)
      if (st.has[side]) *n += (size_t)st.h_counts[2 * q + side];
  }
  return SHX_OK;
}

}  // extern "C"
