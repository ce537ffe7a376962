// TEST INFRASTRUCTURE (oracle/_ref).  C-ABI driver around the reference's own
// hot-path headers, which are #included UNMODIFIED from $REF_DIR (default
// /root/reference) at build time -- see oracle/Makefile.  Nothing here is part
// of the shipped library; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may load the resulting .so.
//
// Include order mirrors SimpleHydrology.cpp:5-6 (vertexpool.h before world.h).
// The only generated input is a copy of cellpool.h whose `mapsize` constant
// (cellpool.h:171) is rewritten for the 2048^2 / 8192^2 builds; it is written
// to oracle/_ref/gen_m<N>/ by the Makefile and pulled in first, so the include
// guard turns world.h's own `#include "cellpool.h"` into a no-op.
#include <glm/glm.hpp>
#include "vertexpool_stub.h"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>

#include "include/FastNoiseLite.h"
#include "include/math.h"
#include SHX_REF_CELLPOOL  // "<ref>/source/cellpool.h" or the generated mapsize variant
#include "world.h"

mappool::pool<quad::cell> cellpool;  // SimpleHydrology.cpp:11
Vertexpool<Vertex> vertexpool;       // SimpleHydrology.cpp:12

static bool g_ready = false;

static void reserve_once() {
  if (cellpool.root.start == NULL) {
    cellpool.reserve(quad::area);  // SimpleHydrology.cpp:36
    vertexpool.reserve(quad::tilearea, quad::maparea);
  }
}

static void zero_non_height() {
  // cellpool.h:122 leaves the pool uninitialised; the reference relies on fresh
  // pages being zero.  Make that explicit.
  for (size_t i = 0; i < (size_t)quad::area; i++) {
    quad::cell& c = cellpool.root.start[i];
    c.discharge = c.momentumx = c.momentumy = 0.0f;
    c.discharge_track = c.momentumx_track = c.momentumy_track = 0.0f;
    c.rootdensity = 0.0f;
  }
}

extern "C" {

int ref_mapsize() { return quad::mapsize; }
int ref_tilesize() { return quad::tilesize; }
int ref_size() { return quad::size; }
size_t ref_ncells() { return (size_t)quad::area; }
size_t ref_cell_bytes() { return sizeof(quad::cell); }
size_t ref_drop_bytes() { return sizeof(Drop); }

// Full reference start-up path: SimpleHydrology.cpp:27-38 (seed, reserve, map.init).
// map.init keeps function-static noise state (cellpool.h:349,393) so it is only
// reference-faithful the first time it runs in a process.
int ref_init(int seed) {
  if (g_ready) return -1;
  reserve_once();
  World::SEED = seed;
  srand(seed);
  std::stringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  World::map.init(vertexpool, cellpool, World::SEED);
  std::cout.rdbuf(old);
  zero_non_height();
  g_ready = true;
  return 0;
}

// Node table only (cellpool.h:327-336), heights left for the caller to fill
// through ref_cells().  Used for synthetic terrains and hand-made test maps.
int ref_init_blank() {
  if (g_ready) return -1;
  reserve_once();
  for (int i = 0; i < quad::mapsize; i++)
    for (int j = 0; j < quad::mapsize; j++) {
      int ind = i * quad::mapsize + j;
      World::map.nodes[ind] = {quad::tileres * ivec2(i, j),
                               vertexpool.section(quad::tilearea / quad::lodarea),
                               {cellpool.get(quad::tilearea / quad::lodarea), quad::tileres / quad::lodsize}};
    }
  memset(cellpool.root.start, 0, sizeof(quad::cell) * (size_t)quad::area);
  g_ready = true;
  return 0;
}

quad::cell* ref_cells() { return cellpool.root.start; }

// order: maxAge minVol evapRate depositionRate entrainment gravity momentumTransfer | lrate maxdiff settling
void ref_get_params(float* p) {
  p[0] = Drop::maxAge; p[1] = Drop::minVol; p[2] = Drop::evapRate; p[3] = Drop::depositionRate;
  p[4] = Drop::entrainment; p[5] = Drop::gravity; p[6] = Drop::momentumTransfer;
  p[7] = World::lrate; p[8] = World::maxdiff; p[9] = World::settling;
}
void ref_set_params(const float* p) {
  Drop::maxAge = p[0]; Drop::minVol = p[1]; Drop::evapRate = p[2]; Drop::depositionRate = p[3];
  Drop::entrainment = p[4]; Drop::gravity = p[5]; Drop::momentumTransfer = p[6];
  World::lrate = p[7]; World::maxdiff = p[8]; World::settling = p[9];
}

void ref_srand(unsigned s) { srand(s); }

// The reference call itself: World::erode(cycles), world.h:54-88 (spawns with rand()).
void ref_erode(int cycles) { World::erode(cycles); }

double ref_time_erode(int cycles, int reps) {
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; r++) World::erode(cycles);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// World::erode with the rand() spawn (world.h:69) replaced by an explicit spawn
// list so that RNG order never matters: reset tracks (world.h:56-61), per spawn
// the height<0.1 rejection (world.h:71-72) and `while(drop.descend());`
// (world.h:74-76) using the reference's own Drop, then the EMA (world.h:81-86).
// stats[0]=spawned stats[1]=rejected stats[2]=descend calls.
void ref_erode_spawnlist(const float* xy, size_t n, int do_reset, int do_ema, uint64_t* stats) {
  uint64_t spawned = 0, rejected = 0, calls = 0;
  if (do_reset)
    for (auto& node : World::map.nodes)
      for (auto [cell, pos] : node.s) {
        cell.discharge_track = 0;
        cell.momentumx_track = 0;
        cell.momentumy_track = 0;
      }
  for (size_t i = 0; i < n; i++) {
    glm::vec2 newpos(xy[2 * i], xy[2 * i + 1]);
    if (World::map.height(newpos) < 0.1) { rejected++; continue; }
    Drop drop(newpos);
    spawned++;
    bool alive;
    do { alive = drop.descend(); calls++; } while (alive);
  }
  if (do_ema)
    for (auto& node : World::map.nodes)
      for (auto [cell, pos] : node.s) {
        cell.discharge = (1.0f - World::lrate) * cell.discharge + World::lrate * cell.discharge_track;
        cell.momentumx = (1.0f - World::lrate) * cell.momentumx + World::lrate * cell.momentumx_track;
        cell.momentumy = (1.0f - World::lrate) * cell.momentumy + World::lrate * cell.momentumy_track;
      }
  if (stats) { stats[0] = spawned; stats[1] = rejected; stats[2] = calls; }
}

// One drop marched with the reference's Drop::descend; after every call the
// state {age,pos.x,pos.y,speed.x,speed.y,volume,sediment} is appended to trace
// (7 floats per call).  Returns the number of descend calls made.
int ref_trace_drop(float x, float y, float* trace, int max_calls) {
  Drop d(glm::vec2(x, y));
  int n = 0;
  bool alive = true;
  while (alive && n < max_calls) {
    alive = d.descend();
    float* t = trace + 7 * (size_t)n;
    t[0] = (float)d.age; t[1] = d.pos.x; t[2] = d.pos.y; t[3] = d.speed.x; t[4] = d.speed.y;
    t[5] = d.volume; t[6] = d.sediment;
    n++;
  }
  return n;
}

// A single Drop::descend call on an explicit state (same 7-float layout).
int ref_descend_once(float* s) {
  Drop d(glm::vec2(s[1], s[2]));
  d.age = (int)s[0]; d.speed = glm::vec2(s[3], s[4]); d.volume = s[5]; d.sediment = s[6];
  bool alive = d.descend();
  s[0] = (float)d.age; s[1] = d.pos.x; s[2] = d.pos.y; s[3] = d.speed.x; s[4] = d.speed.y;
  s[5] = d.volume; s[6] = d.sediment;
  return alive ? 1 : 0;
}

void ref_normal(int x, int y, float* out3) {
  glm::vec3 n = World::map.normal(ivec2(x, y));
  out3[0] = n.x; out3[1] = n.y; out3[2] = n.z;
}

void ref_cascade(float x, float y) { World::cascade(glm::vec2(x, y)); }

float ref_height(int x, int y) { return World::map.height(ivec2(x, y)); }
float ref_discharge(int x, int y) { return World::map.discharge(ivec2(x, y)); }
int ref_oob(float x, float y) { return World::map.oob(glm::vec2(x, y)) ? 1 : 0; }

// updatenode (cellpool.h:286-305) over every node; out receives 12 floats per
// cell in pool order (node-major, then x*tilesize+y), i.e. the Vertex records.
// The reference's own frame (SimpleHydrology.cpp:319-320): World::erode(cycles) then Vegetation::grow(), both
// drawing from the global rand() stream like the application.
void ref_frame(int cycles) {
  World::erode(cycles);
  Vegetation::grow();
}
void ref_vegetation_grow() { Vegetation::grow(); }  // vegetation.h:122-188
size_t ref_plant_count() { return Vegetation::plants.size(); }
void ref_plants(float* out3) {  // {pos.x, pos.y, size} per plant
  for (size_t i = 0; i < Vegetation::plants.size(); i++) {
    out3[3 * i] = Vegetation::plants[i].pos.x;
    out3[3 * i + 1] = Vegetation::plants[i].pos.y;
    out3[3 * i + 2] = Vegetation::plants[i].size;
  }
}

void ref_update_vertices(float* out) {
  for (auto& node : World::map.nodes) quad::updatenode(vertexpool, node);
  memcpy(out, vertexpool.store.data(), sizeof(Vertex) * (size_t)quad::area);
}

}  // extern "C"
