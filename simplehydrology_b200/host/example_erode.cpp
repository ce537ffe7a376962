// Host-side example / test driver for shx::Bridge: a stand-in for the reference's main loop
// (SimpleHydrology.cpp:27-38,314-324) without the renderer.  It owns a tiled AoS cell pool like
// `cellpool`, defines Drop / World statics with the reference's names, and calls
// bridge.erode<Drop, World>(cycles) where the reference calls world.erode(quad::tilesize).
//
// usage: example_erode <heights.f32 (tiled pool order)> <mapsize> <frames> <cycles> <out_cells.bin>
// Also exercises the vegetation coupling: before each frame it stamps rootdensity the way
// Plant::root does (vegetation.h:87-118) at a few seeded places.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "shx_world.hpp"

struct Drop {  // water.h:27-33,43-50
  static float maxAge, minVol, evapRate, depositionRate, entrainment, gravity, momentumTransfer;
};
float Drop::evapRate = 0.001f;
float Drop::depositionRate = 0.1f;
float Drop::minVol = 0.01f;
float Drop::maxAge = 500;
float Drop::entrainment = 10.0f;
float Drop::gravity = 1.0f;
float Drop::momentumTransfer = 1.0f;

struct World {  // world.h:21-44
  static unsigned int SEED;
  static float lrate, maxdiff, settling;
};
unsigned int World::SEED = 1;
float World::lrate = 0.1f;
float World::maxdiff = 0.01f;
float World::settling = 0.8f;

int main(int argc, char** argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s heights.f32 mapsize frames cycles out.bin\n", argv[0]);
    return 2;
  }
  const int mapsize = atoi(argv[2]), frames = atoi(argv[3]), cycles = atoi(argv[4]);
  const int ts = 512;
  const size_t n = (size_t)mapsize * mapsize * ts * ts;
  std::vector<shx_cell> pool(n);  // == cellpool.reserve(quad::area), zeroed
  {
    std::vector<float> h(n);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(h.data(), sizeof(float), n, f) != n) { fprintf(stderr, "cannot read heights\n"); return 2; }
    fclose(f);
    for (size_t i = 0; i < n; i++) { pool[i] = shx_cell{}; pool[i].height = h[i]; }
  }
  try {
    shx::Bridge bridge(pool.data(), mapsize, ts);
    unsigned long long steps = 0;
    for (int fr = 0; fr < frames; fr++) {
      // a Plant::root(1.0)-like stamp in the host pool (centre 1.0, axis 0.6, diagonal 0.4)
      const int size = mapsize * ts;
      const int px = 50 + 37 * fr % (size - 100), py = 60 + 91 * fr % (size - 100);
      for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++) {
          const int x = px + dx, y = py + dy;
          const size_t node = (size_t)(x / ts) * mapsize + (y / ts);
          const size_t i = node * ts * ts + (size_t)(x % ts) * ts + (y % ts);
          pool[i].rootdensity += (dx == 0 && dy == 0) ? 1.0f : ((dx == 0 || dy == 0) ? 0.6f : 0.4f);
        }
      const shx_stats st = bridge.erode<Drop, World>(cycles);
      steps += st.steps;
      printf("frame %d spawned %llu steps %llu phases %llu launches %llu\n", fr, (unsigned long long)st.spawned,
             (unsigned long long)st.steps, (unsigned long long)st.phases, (unsigned long long)st.launches);
    }
    printf("total steps %llu\n", steps);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  FILE* o = fopen(argv[5], "wb");
  fwrite(pool.data(), sizeof(shx_cell), n, o);
  fclose(o);
  return 0;
}
