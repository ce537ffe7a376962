// shx host adaptor (C++17, header only): the drop-in for the reference's World::erode(cycles).
//
// The reference's seam is the static call `World::erode(int cycles)` (reference
// source/world.h:33,54-88), made once per frame from SimpleHydrology.cpp:319 on the global cell
// pool `cellpool.root.start` (SimpleHydrology.cpp:11,36) with the static parameter sets
// Drop::* (water.h:43-50) and World::* (world.h:42-44).  shx::Bridge keeps that contract:
//
//   shx::Bridge bridge(cellpool.root.start, quad::mapsize, quad::tilesize);   // after World::map.init
//   ...
//   bridge.erode<Drop, World>(quad::tilesize);   // instead of world.erode(quad::tilesize)
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

class Bridge {
 public:
  // `pool` is the reference's tiled AoS cell pool (quad::cell == shx_cell, 32 bytes); it stays owned
  // by the caller and must outlive the bridge.
  Bridge(void* pool, int mapsize, int tilesize, int device = 0, bool pin_host_pool = true)
      : pool_(static_cast<shx_cell*>(pool)) {
    static_assert(sizeof(shx_cell) == 32, "quad::cell layout (cellpool.h:207-220)");
    shx_default_params(&params_, mapsize);
    params_.tilesize = tilesize;
    ncells_ = (size_t)mapsize * mapsize * (size_t)tilesize * tilesize;
    shx_config cfg;
    shx_default_config(&cfg);
    cfg.device = device;
    check(shx_create(&ctx_, &params_, &cfg), "shx_create");
    if (pin_host_pool && shx_host_register(pool_, ncells_ * sizeof(shx_cell)) == SHX_OK) pinned_ = true;
    check(shx_upload(ctx_, pool_, ncells_), "shx_upload");
    root_shadow_.resize(ncells_);
    for (size_t i = 0; i < ncells_; i++) root_shadow_[i] = pool_[i].rootdensity;
  }
  ~Bridge() {
    if (pinned_) shx_host_unregister(pool_);
    shx_destroy(ctx_);
  }
  Bridge(const Bridge&) = delete;
  Bridge& operator=(const Bridge&) = delete;

  // == World::erode(cycles) with the statics of the given Drop / World types
  template <class DropT, class WorldT>
  shx_stats erode(int cycles) {
    params_from_statics<DropT, WorldT>(params_);
    return erode(cycles, params_, (uint64_t)WorldT::SEED);
  }

  shx_stats erode(int cycles, const shx_params& p, uint64_t seed) {
    params_ = p;
    check(shx_set_params(ctx_, &params_), "shx_set_params");
    push_rootdensity();
    shx_stats st;
    check(shx_erode(ctx_, cycles, seed, &st), "shx_erode");
    // Whole records: one contiguous DMA per tile.  (Selecting height/discharge/momentum only makes it a
    // strided 16-of-32-byte copy, measured 2x SLOWER at 8192^2.)  rootdensity comes back as pushed above;
    // the *_track fields are scratch of the erode call (world.h:56-61 zeroes them first thing).
    check(shx_download(ctx_, pool_, ncells_, SHX_F_ALL), "shx_download");
    return st;
  }

  // the host edited heights or fields wholesale (e.g. regenerated the world): send everything again
  void reupload() {
    check(shx_upload(ctx_, pool_, ncells_), "shx_upload");
    for (size_t i = 0; i < ncells_; i++) root_shadow_[i] = pool_[i].rootdensity;
  }

  // == `for (auto& node : world.map.nodes) updatenode(vertexpool, node);` (SimpleHydrology.cpp:322-324)
  // on the device.  `vertices` receives the 48-byte Vertex records (vertexpool.h:6-26) of every
  // cell in pool order, i.e. exactly what the reference's Vertexpool<Vertex>::fill calls leave in
  // the pool's mapped buffer (sections are reserved node by node, cellpool.h:327-336).
  void update_vertices(float* vertices) { check(shx_vertex_download(ctx_, vertices, ncells_), "shx_vertex_download"); }
  // the same straight into device memory, e.g. the vertex pool's VBO registered with
  // cudaGraphicsGLRegisterBuffer and mapped: no vertex data crosses PCIe
  void update_vertices_device(float* dev_vertices) { check(shx_vertex_fill(ctx_, dev_vertices), "shx_vertex_fill"); }
  // == the dischargeMap / momentumMap lambdas (SimpleHydrology.cpp:341-354): 4 floats per cell in
  // map order {erf(0.4*discharge), 0.5*(1+erf(momentumx)), 0.5*(1+erf(momentumy)), height}
  void view_maps(float* rgba) { check(shx_view_maps_download(ctx_, rgba, ncells_), "shx_view_maps_download"); }

  // sparse read-back instead of the pool download: the records (and World::map.normal) of n cells {x, y}
  void gather(const int* xy, size_t n, shx_cell* out, float* normals3 = nullptr) {
    check(shx_gather_cells(ctx_, xy, n, out, normals3), "shx_gather_cells");
  }

  shx_ctx* context() { return ctx_; }
  const shx_params& params() const { return params_; }

 private:
  // Plant::root (vegetation.h:87-118) writes rootdensity straight into the pool; find what moved
  void push_rootdensity() {
    xy_.clear();
    val_.clear();
    const int ts = params_.tilesize, ms = params_.mapsize;
    const size_t tile = (size_t)ts * ts;
    for (size_t i = 0; i < ncells_; i++) {
      const float r = pool_[i].rootdensity;
      if (std::memcmp(&r, &root_shadow_[i], sizeof r) == 0) continue;
      root_shadow_[i] = r;
      const size_t node = i / tile, in = i % tile;  // cellpool.h:327-336, math.h:11-14
      xy_.push_back((int)(node / ms) * ts + (int)(in / ts));
      xy_.push_back((int)(node % ms) * ts + (int)(in % ts));
      val_.push_back(r);
    }
    if (!val_.empty()) check(shx_set_rootdensity(ctx_, xy_.data(), val_.data(), val_.size()), "shx_set_rootdensity");
  }

  shx_cell* pool_;
  size_t ncells_ = 0;
  shx_ctx* ctx_ = nullptr;
  shx_params params_;
  bool pinned_ = false;
  std::vector<float> root_shadow_;
  std::vector<int> xy_;
  std::vector<float> val_;
};

}  // namespace shx
