// shx host adaptor (C++17, header only): the drop-in for the reference's World::erode(cycles).
//
// The reference's seam is the static call `World::erode(int cycles)` (reference
// source/world.h:33,54-88), made once per frame from SimpleHydrology.cpp:319 on the global cell
// pool `cellpool.root.start` (SimpleHydrology.cpp:11,36) with the static parameter sets
// Drop::* (water.h:43-50) and World::* (world.h:42-44).  shx::Bridge keeps that contract:
//
//   shx::Bridge bridge(cellpool.root.start, quad::mapsize, quad::tilesize);   // after World::map.init
//   ...
//   bridge.erode<Drop, World>(quad::tilesize, &Vegetation::plants);   // instead of world.erode(quad::tilesize)
//   Vegetation::grow();                          // unchanged, edits rootdensity in the host pool
//   updatenode(...);                             // unchanged, reads height from the host pool
//
// Per call it (1) copies the current Drop::/World:: statics, (2) pushes the rootdensity cells the
// host changed since the last call (vegetation.h:87-118 writes them in place), (3) runs the batched
// CUDA erode and (4) brings the cell records back into the same pool, so the
// renderer, the texture builders (SimpleHydrology.cpp:341-354) and vegetation.h read the result
// unchanged.  Errors: the reference reports none; the bridge throws std::runtime_error for misuse
// of the boundary or CUDA failures (there is no CPU fallback to degrade to).
#pragma once
#include <chrono>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/shx.h"

namespace shx {

inline void check(int rc, const char* what) {
  if (rc != SHX_OK) throw std::runtime_error(std::string(what) + ": " + shx_last_error());
}

// reads the reference's static tunables by name; works with the reference's own Drop / World types
template <class DropT, class WorldT>
inline void params_from_statics(shx_params& p) {
  p.maxAge = DropT::maxAge; p.minVol = DropT::minVol; p.evapRate = DropT::evapRate;
  p.depositionRate = DropT::depositionRate; p.entrainment = DropT::entrainment; p.gravity = DropT::gravity;
  p.momentumTransfer = DropT::momentumTransfer;
  p.lrate = WorldT::lrate; p.maxdiff = WorldT::maxdiff; p.settling = WorldT::settling;
}

// Plant:: statics (vegetation.h:40-44) by name
template <class PlantT>
inline shx_plant_params plant_params_from_statics() {
  shx_plant_params pp;
  pp.maxSize = PlantT::maxSize; pp.growRate = PlantT::growRate; pp.maxSteep = PlantT::maxSteep;
  pp.maxDischarge = PlantT::maxDischarge; pp.maxTreeHeight = PlantT::maxTreeHeight;
  return pp;
}

class Bridge {
 public:
  // `pool` is the reference's tiled AoS cell pool (quad::cell == shx_cell, 32 bytes); it stays owned
  // by the caller and must outlive the bridge.  ngpu > 1 spreads the map over that many GPUs as row strips
  // (shx_multi: one host thread, peer stores over NVLink once per call; mapsize must be divisible by ngpu).
  // shadow_roots: keep a copy of the pool's rootdensity so that the erode overloads WITHOUT a plant list can find
  // what the host changed (4 bytes per cell and one pass per frame); pass false when every call hands over the plants.
  Bridge(void* pool, int mapsize, int tilesize, int ngpu = 1, const int* devices = nullptr, bool pin_host_pool = true,
         bool shadow_roots = true)
      : pool_(static_cast<shx_cell*>(pool)), shadow_roots_(shadow_roots) {
    static_assert(sizeof(shx_cell) == 32, "quad::cell layout (cellpool.h:207-220)");
    shx_default_params(&params_, mapsize);
    params_.tilesize = tilesize;
    ncells_ = (size_t)mapsize * mapsize * (size_t)tilesize * tilesize;
    if (shx_multi_create(&multi_, &params_, ngpu, devices, nullptr) != SHX_OK)
      throw std::runtime_error(std::string("shx_multi_create: ") + shx_multi_last_error());
    if (pin_host_pool && shx_host_register(pool_, ncells_ * sizeof(shx_cell)) == SHX_OK) pinned_ = true;
    mcheck(shx_multi_upload(multi_, pool_, ncells_), "shx_multi_upload");
    snapshot_roots();
  }
  ~Bridge() {
    shx_multi_destroy(multi_);
    if (pinned_) shx_host_unregister(pool_);
  }
  Bridge(const Bridge&) = delete;
  Bridge& operator=(const Bridge&) = delete;

  // == World::erode(cycles) with the statics of the given Drop / World types.
  // `plants` (optional): the reference's Vegetation::plants (vegetation.h:48-53), or any container of objects with
  // a `pos` of two floats.  Plant::root (vegetation.h:87-118) writes rootdensity straight into the host pool, always
  // into the 3x3 cells around a plant; with the plant list the bridge pushes exactly those cells (of the plants
  // alive now and of those alive at the previous call: a plant that died was un-rooted first) instead of scanning
  // the whole pool for changes -- 36 000 cells instead of 67 million at 8192^2.  Without it the whole pool is
  // compared with a shadow copy (any writer is caught; O(cells) per frame).
  template <class DropT, class WorldT, class PlantVec>
  shx_stats erode(int cycles, const PlantVec* plants) {
    params_from_statics<DropT, WorldT>(params_);
    push_roots_near(*plants);
    return run(cycles, params_, (uint64_t)WorldT::SEED);
  }
  template <class DropT, class WorldT>
  shx_stats erode(int cycles) {
    params_from_statics<DropT, WorldT>(params_);
    push_roots_by_scan();
    return run(cycles, params_, (uint64_t)WorldT::SEED);
  }
  shx_stats erode(int cycles, const shx_params& p, uint64_t seed) {
    push_roots_by_scan();
    return run(cycles, p, seed);
  }

  // the host edited heights or fields wholesale (e.g. regenerated the world): send everything again
  void reupload() {
    mcheck(shx_multi_upload(multi_, pool_, ncells_), "shx_multi_upload");
    snapshot_roots();
  }

  // == `for (auto& node : world.map.nodes) updatenode(vertexpool, node);` (SimpleHydrology.cpp:322-324)
  // on the device.  `vertices` receives the 48-byte Vertex records (vertexpool.h:6-26) of every
  // cell in pool order, i.e. exactly what the reference's Vertexpool<Vertex>::fill calls leave in
  // the pool's mapped buffer (sections are reserved node by node, cellpool.h:327-336).  Strips are whole tile rows,
  // so strip i's records are a contiguous slice of the pool order.
  void update_vertices(float* vertices) {
    size_t off = 0;
    for (int i = 0; i < shx_multi_strips(multi_); i++) {
      const size_t n = ncells_ / (size_t)shx_multi_strips(multi_);
      check(shx_vertex_download(shx_multi_strip(multi_, i), vertices + 12 * off, n), "shx_vertex_download");
      off += n;
    }
  }
  // the same straight into device memory of the (single) GPU, e.g. the vertex pool's VBO registered with
  // cudaGraphicsGLRegisterBuffer and mapped (see shx_gl.hpp): no vertex data crosses PCIe
  void update_vertices_device(float* dev_vertices) { check(shx_vertex_fill(context(), dev_vertices), "shx_vertex_fill"); }
  // == the dischargeMap / momentumMap lambdas (SimpleHydrology.cpp:341-354): 4 floats per cell in
  // map order {erf(0.4*discharge), 0.5*(1+erf(momentumx)), 0.5*(1+erf(momentumy)), height}
  void view_maps(float* rgba) {
    size_t off = 0;
    for (int i = 0; i < shx_multi_strips(multi_); i++) {
      const size_t n = ncells_ / (size_t)shx_multi_strips(multi_);
      check(shx_view_maps_download(shx_multi_strip(multi_, i), rgba + 4 * off, n), "shx_view_maps_download");
      off += n;
    }
  }

  // ---- the frame loop with the vegetation on the device too (N3; single GPU): nothing but statistics crosses PCIe.
  //   bridge.enable_device_vegetation<Plant>();                 // once, instead of keeping Vegetation::plants on the host
  //   bridge.erode_resident<Drop, World>(quad::tilesize);       // == world.erode(quad::tilesize), SimpleHydrology.cpp:319
  //   bridge.grow<Plant>(World::SEED, frame);                   // == Vegetation::grow(), :320 (vegetation.h:122-188)
  //   bridge.update_vertices_device(vbo_ptr);                   // :322-324
  //   bridge.tree_models_device(instance_ptr, capacity);        // :329-335
  // PlantT supplies the Plant:: statics (vegetation.h:40-44).  The host pool is NOT updated by these calls;
  // sync_pool() brings it up to date when host code wants to look at it.
  template <class PlantT>
  void enable_device_vegetation(size_t max_plants = 0) {
    const shx_plant_params pp = plant_params_from_statics<PlantT>();
    check(shx_veg_create(context(), max_plants, &pp), "shx_veg_create");
  }
  template <class DropT, class WorldT>
  shx_stats erode_resident(int cycles) {
    params_from_statics<DropT, WorldT>(params_);
    mcheck(shx_multi_set_params(multi_, &params_), "shx_multi_set_params");
    shx_stats st;
    mcheck(shx_multi_erode(multi_, cycles, (uint64_t)WorldT::SEED, &st), "shx_multi_erode");
    return st;
  }
  // plants_out (optional): any vector of objects constructible from a two-float position with a float `size`
  // (Vegetation::plants) receives the list, e.g. for host code that still walks it
  template <class PlantT, class PlantVec = std::vector<PlantT>>
  shx_veg_stats grow(uint64_t seed, uint64_t frame, PlantVec* plants_out = nullptr) {
    const shx_plant_params pp = plant_params_from_statics<PlantT>();
    check(shx_veg_set_params(context(), &pp), "shx_veg_set_params");
    shx_veg_stats st;
    check(shx_veg_grow(context(), seed, frame, &st), "shx_veg_grow");
    if (plants_out) {
      std::vector<float> xys((size_t)st.plants * 3);
      size_t n = 0;
      check(shx_veg_download(context(), xys.data(), (size_t)st.plants, &n), "shx_veg_download");
      plants_out->clear();
      for (size_t i = 0; i < n; i++) {
        plants_out->emplace_back(decltype(PlantT::pos)(xys[3 * i], xys[3 * i + 1]));
        plants_out->back().size = xys[3 * i + 2];
      }
    }
    return st;
  }
  size_t tree_models_device(float* dev_mat4, size_t capacity) {
    size_t n = 0;
    check(shx_veg_tree_models(context(), dev_mat4, capacity, &n), "shx_veg_tree_models");
    return n;
  }
  void sync_pool() { mcheck(shx_multi_download(multi_, pool_, ncells_, SHX_F_ALL), "shx_multi_download"); }

  // sparse read-back instead of the pool download (single GPU): the records (and World::map.normal) of n cells {x, y}
  void gather(const int* xy, size_t n, shx_cell* out, float* normals3 = nullptr) {
    check(shx_gather_cells(context(), xy, n, out, normals3), "shx_gather_cells");
  }

  shx_ctx* context() { return shx_multi_strip(multi_, 0); }  // the first (or only) strip
  shx_multi* multi() { return multi_; }
  const shx_params& params() const { return params_; }
  size_t last_push() const { return val_.size(); }  // rootdensity cells sent by the last erode call

 private:
  static void mcheck(int rc, const char* what) {
    if (rc != SHX_OK) throw std::runtime_error(std::string(what) + ": " + shx_multi_last_error());
  }

  shx_stats run(int cycles, const shx_params& p, uint64_t seed) {
    params_ = p;
    mcheck(shx_multi_set_params(multi_, &params_), "shx_multi_set_params");
    if (!val_.empty()) mcheck(shx_multi_set_rootdensity(multi_, xy_.data(), val_.data(), val_.size()), "shx_multi_set_rootdensity");
    shx_stats st;
    mcheck(shx_multi_erode(multi_, cycles, seed, &st), "shx_multi_erode");
    download();
    return st;
  }

  // What comes back per frame.  kRecords: the whole 32-byte records, one contiguous DMA per tile (a strided
  // 16-of-32-byte DMA measured 2x SLOWER at 8192^2).  kCompact (one GPU): {height, discharge, momentumx, momentumy} --
  // everything host code reads after World::erode; rootdensity is the host's own, the *_track fields are scratch of
  // the call (world.h:56-61 zeroes them first thing) -- as a dense 16-byte stream scattered into the pool by host
  // threads (shx_download_compact): half the PCIe bytes, but every record costs the host a read-for-ownership.
  // Measured on the B200 boxes of this project (16 host cores): 52.3 ms per 8192^2 frame compact against 51.7 with
  // records -- the host scatters no faster than PCIe saves, and its memory traffic slows the DMA down -- so kRecords
  // is the default.  kAuto: frame 0 downloads records, frame 1 the compact stream, the faster one serves from then on
  // (the compact probe frame also pays the one-time allocation of its buffers).
 public:
  enum DownloadMode { kAuto = 0, kRecords = 1, kCompact = 2 };
  void set_download_mode(DownloadMode m) { mode_ = m; }
  DownloadMode download_mode() const { return mode_; }  // kAuto until the second frame has decided
  double last_download_ms() const { return last_download_ms_; }

 private:
  void download() {
    const auto t0 = std::chrono::steady_clock::now();
    DownloadMode use = mode_;
    if (shx_multi_strips(multi_) > 1) use = kRecords;
    else if (use == kAuto) use = frames_ == 0 ? kRecords : kCompact;
    if (use == kCompact) check(shx_download_compact(context(), pool_, ncells_, 0), "shx_download_compact");
    else mcheck(shx_multi_download(multi_, pool_, ncells_, SHX_F_ALL), "shx_multi_download");
    last_download_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (mode_ == kAuto && shx_multi_strips(multi_) == 1) {
      if (frames_ == 0) records_ms_ = last_download_ms_;
      else if (frames_ == 1) mode_ = last_download_ms_ < records_ms_ ? kCompact : kRecords;
    }
    frames_++;
  }

  size_t pool_index(int x, int y) const {  // cellpool.h:327-336, math.h:11-14
    const int ts = params_.tilesize;
    return ((size_t)(x / ts) * params_.mapsize + (size_t)(y / ts)) * ts * ts + (size_t)(x % ts) * ts + (size_t)(y % ts);
  }

  // the 3x3 cells around every plant of this and of the previous call (Plant::root, vegetation.h:87-118)
  template <class PlantVec>
  void push_roots_near(const PlantVec& plants) {
    const int size = params_.mapsize * params_.tilesize;
    std::vector<int> now;
    now.reserve(plants.size() * 2);
    for (const auto& pl : plants) {
      now.push_back((int)pl.pos.x);
      now.push_back((int)pl.pos.y);
    }
    xy_.clear();
    val_.clear();
    auto add = [&](const std::vector<int>& at) {
      for (size_t k = 0; k + 1 < at.size(); k += 2)
        for (int dx = -1; dx <= 1; dx++)
          for (int dy = -1; dy <= 1; dy++) {
            const int x = at[k] + dx, y = at[k + 1] + dy;
            if (x < 0 || y < 0 || x >= size || y >= size) continue;  // getCell() == NULL
            xy_.push_back(x);
            xy_.push_back(y);
            val_.push_back(pool_[pool_index(x, y)].rootdensity);
          }
    };
    add(prev_plants_);  // first, so that a cell listed twice ends with the same (current) value either way
    add(now);
    prev_plants_.swap(now);
  }

  void snapshot_roots() {
    if (!shadow_roots_) return;
    root_shadow_.resize(ncells_);
    for (size_t i = 0; i < ncells_; i++) root_shadow_[i] = pool_[i].rootdensity;
  }

  // any writer: compare the whole pool with the shadow copy
  void push_roots_by_scan() {
    if (!shadow_roots_) throw std::runtime_error("shx::Bridge: erode without a plant list needs shadow_roots = true");
    xy_.clear();
    val_.clear();
    const int ts = params_.tilesize, ms = params_.mapsize;
    const size_t tile = (size_t)ts * ts;
    for (size_t i = 0; i < ncells_; i++) {
      const float r = pool_[i].rootdensity;
      if (std::memcmp(&r, &root_shadow_[i], sizeof r) == 0) continue;
      root_shadow_[i] = r;
      const size_t node = i / tile, in = i % tile;
      xy_.push_back((int)(node / ms) * ts + (int)(in / ts));
      xy_.push_back((int)(node % ms) * ts + (int)(in % ts));
      val_.push_back(r);
    }
  }

  DownloadMode mode_ = kRecords;
  unsigned long long frames_ = 0;
  double records_ms_ = 0.0, last_download_ms_ = 0.0;
  shx_cell* pool_;
  size_t ncells_ = 0;
  shx_multi* multi_ = nullptr;
  shx_params params_;
  bool pinned_ = false;
  bool shadow_roots_ = true;
  std::vector<float> root_shadow_;
  std::vector<int> prev_plants_;
  std::vector<int> xy_;
  std::vector<float> val_;
};

}  // namespace shx
