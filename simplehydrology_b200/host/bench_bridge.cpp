// End-to-end timing of the reference-facing call, in the reference's own host language: the frame loop of
// SimpleHydrology.cpp:314-324 with shx::Bridge::erode in the place of world.erode (line 319), on HOST buffers.
// Per frame: the rootdensity cells around the plants go to the device (what Vegetation::grow left in the host pool),
// the batched erode runs, and the 32-byte cell records come back into the caller's pool (what the renderer,
// vegetation.h and the texture builders read).  bench.py runs this binary for its `e2e` figure and copies the JSON.
//
// usage: bench_bridge <mapsize> <frames> <warmup> [ngpu] [seed] [download: 1 records (default), 2 compact 16-byte stream
//        scattered by host threads, 0 auto = probe both in the first two frames (warmup >= 2)]  -> one JSON object
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

struct Vec2 { float x, y; };
struct Plant { Vec2 pos; float size; };  // vegetation.h:12-20

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s mapsize frames warmup [ngpu] [seed]\n", argv[0]);
    return 2;
  }
  const int mapsize = atoi(argv[1]), frames = atoi(argv[2]), warmup = atoi(argv[3]);
  const int ngpu = argc > 4 ? atoi(argv[4]) : 1;
  const int seed = argc > 5 ? atoi(argv[5]) : 1;
  const int dl_mode = argc > 6 ? atoi(argv[6]) : 1;
  const int ts = 512, size = mapsize * ts;
  const size_t n = (size_t)size * size;
  World::SEED = (unsigned)seed;
  shx_cell* pool = static_cast<shx_cell*>(aligned_alloc(4096, n * sizeof(shx_cell)));  // == cellpool.reserve(quad::area)
  if (!pool) return 2;
  memset(pool, 0, n * sizeof(shx_cell));
  try {
    {  // the reference's terrain (map::init), generated on the device and brought into the host pool
      shx_params p;
      shx_default_params(&p, mapsize);
      shx_multi* m = nullptr;
      if (shx_multi_create(&m, &p, ngpu, nullptr, nullptr) != SHX_OK) throw std::runtime_error(shx_multi_last_error());
      if (shx_multi_init_terrain(m, seed) != SHX_OK || shx_multi_download(m, pool, n, SHX_F_ALL) != SHX_OK)
        throw std::runtime_error(shx_multi_last_error());
      shx_multi_destroy(m);
    }
    const double t_reg0 = now_s();
    shx::Bridge bridge(pool, mapsize, ts, ngpu, nullptr, true, false);
    const double t_reg = now_s() - t_reg0;
    bridge.set_download_mode((shx::Bridge::DownloadMode)dl_mode);
    double dl_first[2] = {0.0, 0.0};
    // a plant population like the reference's after a few hundred frames at this size (~4 000 per 512^2 tile would be
    // 10^6 here; Vegetation::grow adds one candidate per frame plus offspring: use 4 096 plants), moved every frame
    std::vector<Plant> plants(4096);
    unsigned long long lcg = 12345;
    auto rnd = [&]() { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(lcg >> 33); };
    for (int i = 0; i < ngpu; i++) shx_timing_enable(shx_multi_strip(bridge.multi(), i), 1);
    double t0 = 0.0;
    unsigned long long steps = 0, pushed = 0;
    double host_push_s = 0.0;
    for (int f = 0; f < warmup + frames; f++) {
      if (f == warmup) {
        for (int i = 0; i < ngpu; i++) {
          shx_timing tm;
          shx_timing_read(shx_multi_strip(bridge.multi(), i), &tm);
        }
        steps = pushed = 0;
        host_push_s = 0.0;
        t0 = now_s();
      }
      const double h0 = now_s();
      for (auto& pl : plants) {  // stand-in for Vegetation::grow(): plants die and seed, Plant::root edits the pool
        if (rnd() % 16 == 0 || f == 0) {
          pl.pos.x = (float)(1 + rnd() % (size - 2));
          pl.pos.y = (float)(1 + rnd() % (size - 2));
          const int x = (int)pl.pos.x, y = (int)pl.pos.y;
          for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++) {
              const int cx = x + dx, cy = y + dy;
              const size_t i = ((size_t)(cx / ts) * mapsize + (cy / ts)) * ts * ts + (size_t)(cx % ts) * ts + (cy % ts);
              pool[i].rootdensity += (dx == 0 && dy == 0) ? 1.0f : ((dx == 0 || dy == 0) ? 0.6f : 0.4f);
            }
        }
      }
      host_push_s += now_s() - h0;
      const shx_stats st = bridge.erode<Drop, World>(ts, &plants);  // SimpleHydrology.cpp:319
      steps += st.steps;
      pushed += bridge.last_push();
      if (f < 2) dl_first[f] = bridge.last_download_ms();
    }
    const double dt = now_s() - t0;
    const bool compact = bridge.download_mode() == shx::Bridge::kCompact;
    shx_timing sum;
    memset(&sum, 0, sizeof sum);
    double descend_max = 0.0;
    for (int i = 0; i < ngpu; i++) {
      shx_timing tm;
      shx_timing_read(shx_multi_strip(bridge.multi(), i), &tm);
      sum.push_ms += tm.push_ms; sum.spawn_ms += tm.spawn_ms; sum.ema_ms += tm.ema_ms; sum.pack_ms += tm.pack_ms; sum.d2h_ms += tm.d2h_ms;
      sum.scatter_ms += tm.scatter_ms;
      descend_max = tm.descend_ms > descend_max ? tm.descend_ms : descend_max;
      sum.descend_ms += tm.descend_ms;
    }
    const double k = frames > 0 ? 1.0 / frames : 0.0;
    const size_t d2h = n * (compact ? sizeof(shx_cell) / 2 : sizeof(shx_cell)), h2d = (size_t)(pushed * k) * 12;
    printf("{\"value\": %.6e, \"unit\": \"particle-steps/s\", \"ms_per_step\": %.4f, \"steps\": %d, \"warmup\": %d, \"n_gpus\": %d, "
           "\"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu, "
           "\"breakdown_ms_per_step\": {\"host_ms\": %.4f, \"h2d_ms\": %.4f, \"spawn_ms\": %.4f, \"erode_ms\": %.4f, \"ema_ms\": %.4f, "
           "\"pack_ms\": %.4f, \"d2h_ms\": %.4f, \"scatter_wall_ms\": %.4f, \"note\": \"device spans from CUDA events (%s); pack and d2h overlap; "
           "scatter_wall_ms = host wall time of the compact download's copy-and-scatter loop (includes waiting for the DMA); "
           "host_ms = the plant loop standing in for Vegetation::grow\"}, "
           "\"download\": \"%s\", \"download_probe_ms\": {\"records\": %.3f, \"compact\": %.3f}, "
           "\"pool_register_s\": %.3f, \"rootdensity_cells_per_step\": %.0f, "
           "\"api\": \"shx::Bridge::erode<Drop, World>(512, &plants) (simplehydrology_b200/host/shx_world.hpp): sparse rootdensity push + shx_multi_erode + "
           "download into the caller's host pool\"}\n",
           steps / dt, 1e3 * dt * k, frames, warmup, ngpu, h2d, d2h, 1e3 * host_push_s * k, sum.push_ms * k / ngpu, sum.spawn_ms * k / ngpu,
           descend_max * k, sum.ema_ms * k / ngpu, sum.pack_ms * k / ngpu, sum.d2h_ms * k / ngpu, sum.scatter_ms * k / ngpu,
           ngpu > 1 ? "mean over the strips; erode_ms = the slowest strip" : "one GPU",
           compact ? "compact: {height, discharge, momentumx, momentumy} as a dense 16-byte stream scattered into the pool by host threads"
                   : "records: the whole 32-byte records, one DMA per tile",
           dl_mode == 0 && ngpu == 1 ? dl_first[0] : 0.0, dl_mode == 0 && ngpu == 1 ? dl_first[1] : 0.0, t_reg, pushed * k);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    free(pool);
    return 1;
  }
  free(pool);
  return 0;
}
