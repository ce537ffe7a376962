// TEST INFRASTRUCTURE (oracle/_ref/bridge_real_m<N>).  The reference's own World / Drop / Vegetation / quad::cell --
// its headers #included UNMODIFIED from $REF_DIR, exactly like oracle/ref_driver.cpp -- driven through the shipped
// host adaptor simplehydrology_b200/host/shx_world.hpp:
//   * compile-time proof that shx::Bridge::erode<Drop, World>(cycles, &Vegetation::plants) instantiates against the
//     real types, and that quad::cell is the 32-byte record include/shx.h declares (cellpool.h:207-220);
//   * the coupled run of BASELINE configs[4]: SimpleHydrology.cpp:314-324's frame loop with bridge.erode in the
//     place of world.erode (line 319), followed by the UNCHANGED Vegetation::grow() (line 320) on the host pool.
//   * with device_veg = 1: the same loop with Vegetation::grow on the device as well (bridge.erode_resident +
//     bridge.grow<Plant>, N3), instantiated against the reference's real Plant; the host pool and Vegetation::plants
//     are filled once at the end.
// usage: bridge_real <seed> <frames> <out.bin> [ngpu] [device_veg]   (out: the cell pool, then uint64 n, then n x {x, y, size})
// Runs on the GPU box (the binary travels with the snapshot; /root/reference is only needed to build it).
#include <glm/glm.hpp>
#include "vertexpool_stub.h"

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <sstream>

#include "include/FastNoiseLite.h"
#include "include/math.h"
#include SHX_REF_CELLPOOL
#include "world.h"

#include "../simplehydrology_b200/host/shx_world.hpp"

static_assert(sizeof(quad::cell) == sizeof(shx_cell), "quad::cell is the 32-byte record of include/shx.h");
static_assert(offsetof(quad::cell, height) == offsetof(shx_cell, height) && offsetof(quad::cell, discharge) == offsetof(shx_cell, discharge) &&
              offsetof(quad::cell, momentumx) == offsetof(shx_cell, momentumx) && offsetof(quad::cell, momentumy) == offsetof(shx_cell, momentumy) &&
              offsetof(quad::cell, discharge_track) == offsetof(shx_cell, discharge_track) &&
              offsetof(quad::cell, momentumx_track) == offsetof(shx_cell, momentumx_track) &&
              offsetof(quad::cell, momentumy_track) == offsetof(shx_cell, momentumy_track) &&
              offsetof(quad::cell, rootdensity) == offsetof(shx_cell, rootdensity),
              "field order of quad::cell (cellpool.h:207-220)");

mappool::pool<quad::cell> cellpool;  // SimpleHydrology.cpp:11
Vertexpool<Vertex> vertexpool;       // SimpleHydrology.cpp:12

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s seed frames out.bin [ngpu]\n", argv[0]);
    return 2;
  }
  const int seed = atoi(argv[1]), frames = atoi(argv[2]);
  const int ngpu = argc > 4 ? atoi(argv[4]) : 1;
  const bool device_veg = argc > 5 && atoi(argv[5]) != 0;
  World::SEED = seed;  // SimpleHydrology.cpp:27-38
  srand(seed);
  cellpool.reserve(quad::area);
  vertexpool.reserve(quad::tilearea, quad::maparea);
  {
    std::stringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    World::map.init(vertexpool, cellpool, World::SEED);
    std::cout.rdbuf(old);
  }
  for (size_t i = 0; i < (size_t)quad::area; i++) {  // cellpool.h:122 leaves these uninitialised
    quad::cell& c = cellpool.root.start[i];
    c.discharge = c.momentumx = c.momentumy = c.discharge_track = c.momentumx_track = c.momentumy_track = c.rootdensity = 0.0f;
  }
  std::vector<float> vertices((size_t)quad::area * 12);
  unsigned long long steps = 0;
  size_t pushed = 0;
  try {
    shx::Bridge bridge(cellpool.root.start, quad::mapsize, quad::tilesize, ngpu, nullptr, true, false);
    if (device_veg) bridge.enable_device_vegetation<Plant>();
    for (int f = 0; f < frames; f++) {
      if (device_veg) {
        steps += bridge.erode_resident<Drop, World>(quad::tilesize).steps;                                // :319
        bridge.grow<Plant>((uint64_t)World::SEED, (uint64_t)f, f == frames - 1 ? &Vegetation::plants : nullptr);  // :320 on the device
        continue;
      }
      const shx_stats st = bridge.erode<Drop, World>(quad::tilesize, &Vegetation::plants);  // SimpleHydrology.cpp:319
      Vegetation::grow();                                                                  // :320, unchanged
      steps += st.steps;
      pushed += bridge.last_push();
    }
    if (device_veg) bridge.sync_pool();
    bridge.update_vertices(vertices.data());  // :322-324 on the device
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  FILE* o = fopen(argv[3], "wb");
  if (!o) return 2;
  fwrite(cellpool.root.start, sizeof(quad::cell), (size_t)quad::area, o);
  const uint64_t n = Vegetation::plants.size();
  fwrite(&n, sizeof n, 1, o);
  for (const Plant& p : Vegetation::plants) {
    const float rec[3] = {p.pos.x, p.pos.y, p.size};
    fwrite(rec, sizeof(float), 3, o);
  }
  fwrite(vertices.data(), sizeof(float), 12, o);  // first Vertex record (smoke check of the device fill)
  fclose(o);
  printf("frames %d particle steps %llu plants %llu rootdensity cells pushed %zu\n", frames, steps, (unsigned long long)n, pushed);
  return 0;
}
