/* TEST INFRASTRUCTURE -- CPU oracle for the erosion hot path (see shx_oracle.h).
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no FMA contraction: the reference's
 * trajectories change under contraction).
 *
 * Every function cites the reference lines it restates.  glm is not vendored by
 * the reference; its scalar definitions (length = sqrt(dot), normalize =
 * v*(1/sqrt(dot)), dot summed left to right, converting constructors truncate)
 * are restated inline where used.
 */
#include "shx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* float -> fixed point, round to nearest even, SATURATING, NaN -> 0: the semantics of the GPU's
 * cvt.rni.s32.f32 / cvt.rni.s64.f32, so that even a world that has blown up numerically (heights
 * beyond the Q5.26 range) stays comparable bit for bit.  Sums then wrap (two's complement). */
static int32_t sat32(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.0f) return INT32_MAX;
  if (v <= -2147483648.0f) return INT32_MIN;
  return (int32_t)lrintf(v);
}
static int64_t sat64(float v) {
  if (v != v) return 0;
  if (v >= 9223372036854775808.0f) return INT64_MAX;
  if (v <= -9223372036854775808.0f) return INT64_MIN;
  return (int64_t)llrintf(v);
}
/* sediment ledger: Q31.32 (power-of-two scaling is exact in fp32) */
static int64_t tq(float v) { return sat64(v * 4294967296.0f); }
/* track accumulators: Q13.18 in an int32 */
#define TRACK_SCALE 262144.0f
#define TRACK_INV 3.814697265625e-6f
static int32_t trq(float v) { return sat32(v * TRACK_SCALE); }
static float track_f(int32_t v) { return (float)v * TRACK_INV; }
static int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

/* ------------------------------------------------------------------ params */

void orc_default_params(orc_params* p, int mapsize) {
  p->evapRate = 0.001f;        /* water.h:43 */
  p->depositionRate = 0.1f;    /* water.h:44 */
  p->minVol = 0.01f;           /* water.h:45 */
  p->maxAge = 500.0f;          /* water.h:46 */
  p->entrainment = 10.0f;      /* water.h:48 */
  p->gravity = 1.0f;           /* water.h:49 */
  p->momentumTransfer = 1.0f;  /* water.h:50 */
  p->lrate = 0.1f;             /* world.h:42 */
  p->maxdiff = 0.01f;          /* world.h:43 */
  p->settling = 0.8f;          /* world.h:44 */
  p->mapscale = 80;            /* cellpool.h:165 */
  p->tilesize = 512;           /* cellpool.h:167 */
  p->mapsize = mapsize;        /* cellpool.h:171 */
  p->lodsize = 1;              /* cellpool.h:178 (only lodsize 1 is supported) */
}

size_t orc_tiled_index(const orc_params* p, int x, int y) {
  const int ts = p->tilesize;
  const size_t node = (size_t)(x / ts) * (size_t)p->mapsize + (size_t)(y / ts); /* cellpool.h:421-426 */
  return node * (size_t)ts * (size_t)ts + (size_t)(x % ts) * (size_t)ts + (size_t)(y % ts); /* math.h:11-14 */
}

/* ------------------------------------------------------------------ erf */

float orc_erff_libm(float x) { return erff(x); }

/* Restatement of shx_erff (product: simplehydrology_b200/csrc/shx_math.cuh).
 * Plain fp32 multiplies and adds in this exact order, no FMA.
 *   |x| < 0.875 : x + x*q(x^2)
 *   |x| < 4     : 1 - exp(-x^2) * g(1/(1+|x|))
 *   else        : 1
 * Coefficients: tools/fit_erf.py (max error 1.47 ulp against the exact erf). */
static float poly_exp_neg(float y) {
  const float k = rintf(y * 0x1.715476p+0f);
  float r = y - k * 0x1.62e400p-1f;
  r = r - k * 0x1.7f7d1cp-20f;
  float p = 0x1.6c16c2p-10f;
  p = p * r + 0x1.111112p-7f;
  p = p * r + 0x1.555556p-5f;
  p = p * r + 0x1.555556p-3f;
  p = p * r + 0.5f;
  p = p * r + 1.0f;
  p = p * r + 1.0f;
  union { int32_t i; float f; } s;
  s.i = ((int32_t)k + 127) << 23;
  return p * s.f;
}

float orc_erff_poly(float x) {
  const float ax = fabsf(x);
  const float t = ax * ax;
  float r;
  if (ax < 0.875f) {
    float q = 0x1.6cae4ap-14f;
    q = q * t + -0x1.aebb60p-11f;
    q = q * t + 0x1.554138p-8f;
    q = q * t + -0x1.b81a4ep-6f;
    q = q * t + 0x1.ce2e8cp-4f;
    q = q * t + -0x1.812744p-2f;
    q = q * t + 0x1.06eba8p-3f;
    r = ax + ax * q;
  } else if (ax < 4.0f) {
    const float z = 1.0f / (ax + 1.0f);
    float g = 0x1.e2ce40p-1f;
    g = g * z + -0x1.9f1d12p+1f;
    g = g * z + 0x1.1f8fc4p+2f;
    g = g * z + -0x1.65195ep+1f;
    g = g * z + 0x1.1f65a8p-2f;
    g = g * z + 0x1.7e0a80p-3f;
    g = g * z + 0x1.25f36ep-1f;
    g = g * z + 0x1.2092dap-1f;
    g = g * z + 0x1.bfc2c6p-17f;
    r = 1.0f - poly_exp_neg(-t) * g;
  } else {
    r = 1.0f; /* also NaN -> 1, never reached with finite discharge */
  }
  return copysignf(r, x);
}

/* ------------------------------------------------------------------ small glm restatements */

typedef struct { float x, y, z; } v3;

static v3 v3mul(v3 a, v3 b) { v3 r = {a.x * b.x, a.y * b.y, a.z * b.z}; return r; }
static v3 v3cross(v3 a, v3 b) {
  v3 r = {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
  return r;
}
static float v3dot(v3 a, v3 b) { v3 t = v3mul(a, b); return t.x + t.y + t.z; }
static float v2dot(float ax, float ay, float bx, float by) { return ax * bx + ay * by; }
static float v2len(float x, float y) { return sqrtf(v2dot(x, y, x, y)); }

/* ivec2(vec2): truncation toward zero, water.h:60 / world.h:114.  Positions whose
 * truncation does not fit an int cannot occur (|pos| stays within the map +- 2). */
static int trunc_i(float v) { return (int)v; }

/* ================================================================== sequential */

int orc_seq_oob(const orc_seq_world* w, int x, int y) { /* cellpool.h:413-419 */
  const int size = w->p.mapsize * w->p.tilesize;
  return x < 0 || y < 0 || x >= size || y >= size;
}

static orc_cell* seq_cell(const orc_seq_world* w, int x, int y) { /* cellpool.h:428-431 */
  if (orc_seq_oob(w, x, y)) return NULL;
  return w->cells + orc_tiled_index(&w->p, x, y);
}

float orc_seq_height(const orc_seq_world* w, int x, int y) { /* cellpool.h:433-437, :236-240 */
  const orc_cell* c = seq_cell(w, x, y);
  return c ? c->height : 0.0f;
}

void orc_seq_normal(const orc_seq_world* w, int x, int y, float* out) { /* cellpool.h:181-204 */
  v3 n = {0.0f, 0.0f, 0.0f};
  const v3 s = {1.0f, (float)w->p.mapscale, 1.0f}; /* :185 */
  const float hc = orc_seq_height(w, x, y);
  const float dxp = orc_seq_height(w, x + 1, y) - hc;
  const float dxm = orc_seq_height(w, x - 1, y) - hc;
  const float dyp = orc_seq_height(w, x, y + 1) - hc;
  const float dym = orc_seq_height(w, x, y - 1) - hc;
  v3 a, b, c;
  if (!orc_seq_oob(w, x + 1, y + 1)) { /* :187-188 */
    a = (v3){0.0f, dyp, 1.0f}; b = (v3){1.0f, dxp, 0.0f};
    c = v3cross(v3mul(s, a), v3mul(s, b)); n.x += c.x; n.y += c.y; n.z += c.z;
  }
  if (!orc_seq_oob(w, x - 1, y - 1)) { /* :190-191 */
    a = (v3){0.0f, dym, -1.0f}; b = (v3){-1.0f, dxm, 0.0f};
    c = v3cross(v3mul(s, a), v3mul(s, b)); n.x += c.x; n.y += c.y; n.z += c.z;
  }
  if (!orc_seq_oob(w, x + 1, y - 1)) { /* :194-195 */
    a = (v3){1.0f, dxp, 0.0f}; b = (v3){0.0f, dym, -1.0f};
    c = v3cross(v3mul(s, a), v3mul(s, b)); n.x += c.x; n.y += c.y; n.z += c.z;
  }
  if (!orc_seq_oob(w, x - 1, y + 1)) { /* :197-198 */
    a = (v3){-1.0f, dxm, 0.0f}; b = (v3){0.0f, dyp, 1.0f};
    c = v3cross(v3mul(s, a), v3mul(s, b)); n.x += c.x; n.y += c.y; n.z += c.z;
  }
  if (sqrtf(v3dot(n, n)) > 0) { /* :200-201, normalize = v * (1/sqrt(dot)) */
    const float inv = 1.0f / sqrtf(v3dot(n, n));
    n.x *= inv; n.y *= inv; n.z *= inv;
  }
  out[0] = n.x; out[1] = n.y; out[2] = n.z;
}

uint32_t orc_seq_cascade(orc_seq_world* w, float px, float py) { /* world.h:90-168 */
  static const int off[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}}; /* :94-103 */
  struct { int x, y; float h, d; } sn[8], tmp;
  int num = 0;
  uint32_t transfers = 0;
  const int ix = trunc_i(px), iy = trunc_i(py); /* :114 */
  for (int k = 0; k < 8; k++) {                 /* :116-125 */
    const int nx = ix + off[k][0], ny = iy + off[k][1];
    if (orc_seq_oob(w, nx, ny)) continue;
    sn[num].x = nx; sn[num].y = ny;
    sn[num].h = seq_cell(w, nx, ny)->height;
    sn[num].d = sqrtf((float)(off[k][0] * off[k][0] + off[k][1] * off[k][1])); /* length(vec2(nn)) */
    num++;
  }
  /* :129-131 std::sort ascending by h; for n <= 16 libstdc++ runs a plain insertion
   * sort, which keeps equal keys in collection order. */
  for (int i = 1; i < num; i++) {
    tmp = sn[i];
    int j = i - 1;
    while (j >= 0 && tmp.h < sn[j].h) { sn[j + 1] = sn[j]; j--; }
    sn[j + 1] = tmp;
  }
  orc_cell* c = seq_cell(w, ix, iy);
  for (int i = 0; i < num; i++) { /* :133-166 */
    const float diff = c->height - sn[i].h; /* :138, centre re-read, neighbour snapshot */
    if (diff == 0) continue;
    float excess;
    if ((double)sn[i].h > 0.1) /* :144, double compare */
      excess = fabsf(diff) - sn[i].d * w->p.maxdiff * (float)w->p.lodsize;
    else
      excess = fabsf(diff);
    if (excess <= 0) continue;
    const float transfer = w->p.settling * excess / 2.0f; /* :154 */
    orc_cell* nc = seq_cell(w, sn[i].x, sn[i].y);
    if (diff > 0) { c->height -= transfer; nc->height += transfer; } /* :157-164 */
    else { c->height += transfer; nc->height -= transfer; }
    transfers++;
  }
  return transfers;
}

int orc_seq_descend(orc_seq_world* w, orc_drop* d, orc_stats* st) { /* water.h:58-156 */
  const orc_params* P = &w->p;
  const float lod = (float)P->lodsize;
  const int ix = trunc_i(d->px), iy = trunc_i(d->py); /* :60 */
  orc_cell* cell = seq_cell(w, ix, iy);               /* :62-68 */
  if (st) st->steps++;
  if (cell == NULL) { d->flags = ORC_DROP_DONE_NULL; return 0; }

  float n[3];
  orc_seq_normal(w, ix, iy, n); /* :70 */

  if ((float)d->age > P->maxAge) { /* :74-77 */
    cell->height += d->sediment;
    if (st) { st->term_age++; st->fx_sed_deposited += tq(d->sediment); }
    d->flags = ORC_DROP_DONE_AGE;
    return 0;
  }
  if (d->volume < P->minVol) { /* :79-82 */
    cell->height += d->sediment;
    if (st) { st->term_vol++; st->fx_sed_deposited += tq(d->sediment); }
    d->flags = ORC_DROP_DONE_VOL;
    return 0;
  }

  float effD = P->depositionRate * (1.0f - cell->rootdensity); /* :86-87 */
  if (effD < 0) effD = 0;

  { /* :95  speed += lodsize*gravity*vec2(n.x, n.z)/volume */
    const float g = lod * P->gravity;
    d->sx += (g * n[0]) / d->volume;
    d->sy += (g * n[2]) / d->volume;
  }
  const float fx = cell->momentumx, fy = cell->momentumy; /* :97 */
  if (v2len(fx, fy) > 0 && v2len(d->sx, d->sy) > 0) {      /* :98-99 */
    const float fi = 1.0f / sqrtf(v2dot(fx, fy, fx, fy));
    const float si = 1.0f / sqrtf(v2dot(d->sx, d->sy, d->sx, d->sy));
    const float dp = v2dot(fx * fi, fy * fi, d->sx * si, d->sy * si);
    const float k = lod * P->momentumTransfer * dp / (d->volume + cell->discharge);
    d->sx += k * fx;
    d->sy += k * fy;
  }
  if (v2len(d->sx, d->sy) > 0) { /* :108-109 */
    const float si = 1.0f / sqrtf(v2dot(d->sx, d->sy, d->sx, d->sy));
    const float m = lod * sqrtf(2.0f);
    d->sx = m * (d->sx * si);
    d->sy = m * (d->sy * si);
  }
  d->px += d->sx; /* :111 */
  d->py += d->sy;

  cell->discharge_track += d->volume;         /* :115-117, old cell, new speed */
  cell->momentumx_track += d->volume * d->sx;
  cell->momentumy_track += d->volume * d->sy;

  const int nix = trunc_i(d->px), niy = trunc_i(d->py);
  const int out = orc_seq_oob(w, nix, niy);
  float h2;
  if (out) h2 = (float)((double)cell->height - 0.002); /* :121-122, double arithmetic */
  else h2 = orc_seq_height(w, nix, niy);               /* :124 nearest cell, truncated */

  const float er = w->erf_poly ? orc_erff_poly(0.4f * cell->discharge) : erff(0.4f * cell->discharge); /* cellpool.h:242-244 */
  float c_eq = (1.0f + P->entrainment * er) * (cell->height - h2); /* :127-128 */
  if (c_eq < 0) c_eq = 0;
  const float cdiff = c_eq - d->sediment; /* :129 */
  const float before = d->sediment;
  d->sediment += effD * cdiff;            /* :131 */
  cell->height -= effD * cdiff;           /* :132 */

  const float carried = d->sediment;
  d->sediment = (float)((double)d->sediment / (1.0 - (double)P->evapRate)); /* :135 */
  d->volume = (float)((double)d->volume * (1.0 - (double)P->evapRate));     /* :136 */
  if (st) st->fx_sed_inflation += tq(d->sediment) - tq(carried);
  (void)before;

  if (out) { /* :139-142 */
    if (st) { st->term_oob++; st->fx_sed_oob_lost += tq(d->sediment); }
    d->volume = 0.0f;
    d->flags = ORC_DROP_DONE_OOB;
    return 0;
  }
  const uint32_t t = orc_seq_cascade(w, d->px, d->py); /* :151 */
  if (st) st->cascade_transfers += t;
  d->age++; /* :153 */
  return 1;
}

void orc_seq_reset_tracks(orc_seq_world* w) { /* world.h:56-61 */
  const size_t n = (size_t)w->p.mapsize * w->p.mapsize * (size_t)w->p.tilesize * w->p.tilesize;
  for (size_t i = 0; i < n; i++) {
    w->cells[i].discharge_track = 0;
    w->cells[i].momentumx_track = 0;
    w->cells[i].momentumy_track = 0;
  }
}

void orc_seq_ema(orc_seq_world* w) { /* world.h:81-86 */
  const size_t n = (size_t)w->p.mapsize * w->p.mapsize * (size_t)w->p.tilesize * w->p.tilesize;
  const float lr = w->p.lrate;
  for (size_t i = 0; i < n; i++) {
    orc_cell* c = w->cells + i;
    c->discharge = (1.0f - lr) * c->discharge + lr * c->discharge_track;
    c->momentumx = (1.0f - lr) * c->momentumx + lr * c->momentumx_track;
    c->momentumy = (1.0f - lr) * c->momentumy + lr * c->momentumy_track;
  }
}

void orc_seq_erode_spawnlist(orc_seq_world* w, const float* xy, size_t n, int do_reset, int do_ema, orc_stats* st) {
  if (do_reset) orc_seq_reset_tracks(w);
  for (size_t i = 0; i < n; i++) { /* world.h:64-78 */
    const float x = xy[2 * i], y = xy[2 * i + 1];
    if ((double)orc_seq_height(w, trunc_i(x), trunc_i(y)) < 0.1) { /* :71-72 */
      if (st) st->rejected++;
      continue;
    }
    orc_drop d = {x, y, 0.0f, 0.0f, 1.0f, 0.0f, 0, ORC_DROP_ALIVE}; /* water.h:14-23 */
    if (st) st->spawned++;
    while (orc_seq_descend(w, &d, st)) {} /* :76 */
  }
  if (do_ema) orc_seq_ema(w);
}

int orc_seq_trace_drop(orc_seq_world* w, float x, float y, float* trace, int max_calls) {
  orc_drop d = {x, y, 0.0f, 0.0f, 1.0f, 0.0f, 0, ORC_DROP_ALIVE};
  int n = 0, alive = 1;
  while (alive && n < max_calls) {
    alive = orc_seq_descend(w, &d, NULL);
    float* t = trace + 7 * (size_t)n;
    t[0] = (float)d.age; t[1] = d.px; t[2] = d.py; t[3] = d.sx; t[4] = d.sy; t[5] = d.volume; t[6] = d.sediment;
    n++;
  }
  return n;
}

/* ================================================================== lock-step */

#define HSCALE 67108864.0f           /* 2^26 */
#define HINV 1.490116119384765625e-8f /* 2^-26 */

int32_t orc_ls_quantize_height(float h) { return sat32(h * HSCALE); }
static float hf(int32_t v) { return (float)v * HINV; }
orc_ls_world* orc_ls_create(const orc_params* p) {
  orc_ls_world* w = (orc_ls_world*)calloc(1, sizeof(*w));
  w->p = *p;
  w->size = p->mapsize * p->tilesize;
  const size_t n = (size_t)w->size * (size_t)w->size;
  w->h[0] = (int32_t*)calloc(n, sizeof(int32_t));
  w->h[1] = (int32_t*)calloc(n, sizeof(int32_t));
  w->field = (float*)calloc(n * 4, sizeof(float));
  w->track = (orc_track*)calloc(n, sizeof(orc_track));
  w->row0 = 0;
  w->row1 = w->size;
  w->exclusive_cells = 3;
  w->cur_damp = 1.0f;
  w->free_waits = 8;  /* shx_config.free_waits default */
  w->recip_evap = 1;  /* the batched kernels multiply by 1/(1-evapRate) instead of dividing */
  return w;
}

void orc_ls_destroy(orc_ls_world* w) {
  if (!w) return;
  free(w->h[0]); free(w->h[1]); free(w->field); free(w->track); free(w);
}

void orc_ls_upload(orc_ls_world* w, const orc_cell* tiled) {
  for (int x = 0; x < w->size; x++)
    for (int y = 0; y < w->size; y++) {
      const orc_cell* c = tiled + orc_tiled_index(&w->p, x, y);
      const size_t i = (size_t)x * w->size + y;
      w->h[0][i] = w->h[1][i] = orc_ls_quantize_height(c->height);
      w->field[4 * i + 0] = c->discharge;
      w->field[4 * i + 1] = c->momentumx;
      w->field[4 * i + 2] = c->momentumy;
      w->field[4 * i + 3] = c->rootdensity;
      w->track[i].discharge = trq(c->discharge_track);
      w->track[i].momentumx = trq(c->momentumx_track);
      w->track[i].momentumy = trq(c->momentumy_track);
      w->track[i].pad = 0;
    }
}


void orc_ls_download(const orc_ls_world* w, orc_cell* tiled) {
  for (int x = 0; x < w->size; x++)
    for (int y = 0; y < w->size; y++) {
      orc_cell* c = tiled + orc_tiled_index(&w->p, x, y);
      const size_t i = (size_t)x * w->size + y;
      c->height = hf(w->h[0][i]);
      c->discharge = w->field[4 * i + 0];
      c->momentumx = w->field[4 * i + 1];
      c->momentumy = w->field[4 * i + 2];
      c->rootdensity = w->field[4 * i + 3];
      c->discharge_track = track_f(w->track[i].discharge);
      c->momentumx_track = track_f(w->track[i].momentumx);
      c->momentumy_track = track_f(w->track[i].momentumy);
    }
}

static uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void orc_ls_spawn(const orc_params* p, uint64_t seed, uint64_t epoch, int cycles, float* xy) {
  /* world.h:64-69: for every node, `cycles` drops at node.pos + (r % tileres.x, r' % tileres.y);
   * rand() is replaced by a counter-based hash keyed (seed, epoch, node, i). */
  const uint64_t key = mix64(mix64(seed) + epoch);
  const int ts = p->tilesize;
  size_t k = 0;
  for (int node = 0; node < p->mapsize * p->mapsize; node++)
    for (int i = 0; i < cycles; i++, k++) {
      const uint64_t r = mix64(key + (((uint64_t)node << 32) | (uint32_t)i));
      const int nx = (node / p->mapsize) * ts, ny = (node % p->mapsize) * ts; /* cellpool.h:327-333 */
      xy[2 * k + 0] = (float)(nx + (int)((uint32_t)r % (uint32_t)ts));
      xy[2 * k + 1] = (float)(ny + (int)((uint32_t)(r >> 32) % (uint32_t)ts));
    }
}

static int ls_oob(const orc_ls_world* w, int x, int y) { return x < 0 || y < 0 || x >= w->size || y >= w->size; }

/* Twins: drops created on bit-identical positions in one batch have identical state, hence identical claim keys,
 * and would step together for their whole lives (each applying the full erosion to the same cells).  The k-th
 * duplicate of a position starts with `waited` = min(k, 7): the copies take their first turns one after another and
 * are different drops from then on.  Which copy gets which rank does not matter (they are identical), so the result
 * stays independent of the order of the list. */
typedef struct { uint32_t x, y; size_t i; } twin_key;
static int twin_cmp(const void* a, const void* b) {
  const twin_key* p = (const twin_key*)a; const twin_key* q = (const twin_key*)b;
  if (p->x != q->x) return p->x < q->x ? -1 : 1;
  if (p->y != q->y) return p->y < q->y ? -1 : 1;
  return p->i < q->i ? -1 : (p->i > q->i);
}

void orc_ls_make_drops(const orc_ls_world* w, const float* xy, size_t n, orc_drop* drops, orc_stats* st) {
  for (size_t i = 0; i < n; i++) {
    orc_drop d = {xy[2 * i], xy[2 * i + 1], 0.0f, 0.0f, 1.0f, 0.0f, 0, ORC_DROP_ALIVE};
    const int ix = trunc_i(d.px), iy = trunc_i(d.py);
    const float h = ls_oob(w, ix, iy) ? 0.0f : hf(w->h[0][(size_t)ix * w->size + iy]);
    if ((double)h < 0.1) { d.flags = ORC_DROP_REJECTED; if (st) st->rejected++; } /* world.h:71-72 */
    else if (ls_oob(w, ix, iy)) { d.flags = ORC_DROP_DONE_NULL; }
    else if (st) st->spawned++;
    drops[i] = d;
  }
  if (n > 1) {
    twin_key* k = (twin_key*)malloc(sizeof(twin_key) * n);
    for (size_t i = 0; i < n; i++) { memcpy(&k[i].x, &xy[2 * i], 4); memcpy(&k[i].y, &xy[2 * i + 1], 4); k[i].i = i; }
    qsort(k, n, sizeof(twin_key), twin_cmp);
    for (size_t a = 0; a < n;) {
      size_t b = a + 1;
      while (b < n && k[b].x == k[a].x && k[b].y == k[a].y) b++;
      for (size_t r = a + 1; r < b; r++) {
        const int rank = (int)(r - a) < 7 ? (int)(r - a) : 7;
        if (drops[k[r].i].flags & ORC_DROP_ALIVE) drops[k[r].i].flags |= rank << ORC_DROP_WAITED_SHIFT;
      }
      a = b;
    }
    free(k);
  }
}

/* World::cascade (world.h:90-168) on a private 3x3 integer block B (centre index 4,
 * k = (dx+1)*3 + (dy+1)); inb[k] tells which cells exist.  Transfers are quantised
 * to the height fixed point and applied to B and to the delta block D. */
static uint32_t ls_cascade_block(const orc_params* P, int32_t* B, const int* inb, int32_t* D, float damp) {
  static const int order[8] = {0, 1, 2, 3, 5, 6, 7, 8}; /* world.h:94-103 in block indices */
  struct { int k; float h, d; } sn[8], tmp;
  int num = 0;
  uint32_t transfers = 0;
  for (int j = 0; j < 8; j++) {
    const int k = order[j];
    if (!inb[k]) continue;
    sn[num].k = k;
    sn[num].h = hf(B[k]);
    sn[num].d = (k == 0 || k == 2 || k == 6 || k == 8) ? sqrtf(2.0f) : 1.0f;
    num++;
  }
  for (int i = 1; i < num; i++) {
    tmp = sn[i];
    int j = i - 1;
    while (j >= 0 && tmp.h < sn[j].h) { sn[j + 1] = sn[j]; j--; }
    sn[j + 1] = tmp;
  }
  for (int i = 0; i < num; i++) {
    const float diff = hf(B[4]) - sn[i].h;
    if (diff == 0) continue;
    float excess;
    if ((double)sn[i].h > 0.1) excess = fabsf(diff) - sn[i].d * P->maxdiff * (float)P->lodsize;
    else excess = fabsf(diff);
    if (excess <= 0) continue;
    const float transfer = (P->settling * damp) * excess / 2.0f; /* damp: 1, or 2^-n in a crowd (see cur_damp) */
    const int32_t t = orc_ls_quantize_height(transfer);
    if (diff > 0) { B[4] -= t; D[4] -= t; B[sn[i].k] += t; D[sn[i].k] += t; }
    else { B[4] += t; D[4] += t; B[sn[i].k] -= t; D[sn[i].k] -= t; }
    transfers++;
  }
  return transfers;
}

/* One phase of one drop: [cascade owed from the previous step] + one Drop::descend
 * (water.h:58-156) reading plane R only; all height changes go to the 3x3 delta block
 * D around (ix,iy).  Returns 1 if the drop was processed (D/ix/iy valid). */
/* Same-cell exclusion of the lock step: of the drops that stand on one cell in a phase only the holder of the
 * highest key steps, the others wait a phase (a drop reads frozen heights, so two drops eroding one cell in the
 * same phase would both take the full amount: over-erosion that feeds on itself in busy river cells).
 * exclusive_cells == 2 (default) widens the rule to the 3x3 block: a drop also waits while a drop with a higher key
 * stands on one of the eight cells around it, so that neighbouring cells never change in the same phase (a train of
 * drops along a river updating all its cells at once is an explicit scheme with a factor above 1: it oscillates).
 * key = {phase tag | phases waited so far, saturating at 7 : 3 | hash of the drop's state : 13}: longest waiting
 * first, then an order-independent pseudo-random choice; drops with equal keys all step. */
static uint32_t ls_claim_key(uint32_t tag, const orc_drop* d) {
  uint32_t px, py;
  memcpy(&px, &d->px, 4);
  memcpy(&py, &d->py, 4);
  uint32_t h = px * 0x9E3779B1u ^ py * 0x85EBCA77u ^ (uint32_t)d->age * 0xC2B2AE3Du;
  h ^= h >> 15;
  const uint32_t waited = ((uint32_t)d->flags >> ORC_DROP_WAITED_SHIFT) & 7u;
  return (tag << 16) | (waited << 13) | (h & 0x1FFFu);
}

/* what the drop itself added to cell (x, y) in the earlier steps of the current phase (S > 1) */
typedef struct { const int32_t* D; const int* pos; int n; } ls_own;
static int32_t ls_own_delta(const ls_own* own, int x, int y) {
  int32_t v = 0;
  if (!own) return 0;
  for (int s = 0; s < own->n; s++) {
    const int ox = x - own->pos[2 * s], oy = y - own->pos[2 * s + 1];
    if (ox >= -1 && ox <= 1 && oy >= -1 && oy <= 1) v += own->D[9 * s + (ox + 1) * 3 + (oy + 1)];
  }
  return v;
}

static int ls_step(orc_ls_world* w, const int32_t* R, orc_drop* d, int32_t* D, int* pix, int* piy, orc_stats* st, const ls_own* own) {
  const orc_params* P = &w->p;
  const int size = w->size;
  const float lod = (float)P->lodsize;
  const int ix = trunc_i(d->px), iy = trunc_i(d->py);
  *pix = ix; *piy = iy;
  int32_t B[9];
  int inb[9];
  for (int dx = -1; dx <= 1; dx++)
    for (int dy = -1; dy <= 1; dy++) {
      const int k = (dx + 1) * 3 + (dy + 1);
      inb[k] = !ls_oob(w, ix + dx, iy + dy);
      B[k] = inb[k] ? R[(size_t)(ix + dx) * size + (iy + dy)] + ls_own_delta(own, ix + dx, iy + dy) : 0;
      D[k] = 0;
    }
  st->steps++;
  if (d->flags & ORC_DROP_CASCADE) { /* water.h:151 of the previous call */
    st->cascade_transfers += ls_cascade_block(P, B, inb, D, w->cur_damp);
    d->flags &= ~ORC_DROP_CASCADE;
  }
  /* cellpool.h:181-204 on the block; height() of a missing cell is 0 (cellpool.h:433-437) */
  const float hc = hf(B[4]);
  const float hxp = inb[7] ? hf(B[7]) : 0.0f, hxm = inb[1] ? hf(B[1]) : 0.0f;
  const float hyp = inb[5] ? hf(B[5]) : 0.0f, hym = inb[3] ? hf(B[3]) : 0.0f;
  const float sc = (float)P->mapscale;
  const float Bp = sc * (hxp - hc), Bm = sc * (hxm - hc), Ap = sc * (hyp - hc), Am = sc * (hym - hc);
  float nx = 0.0f, ny = 0.0f, nz = 0.0f;
  /* cross products of cellpool.h:188,191,195,198 written out: each plane adds
   * (-+80*dhx, 1, -+80*dhy); the zero products of the generic formula only
   * influence the sign of a zero, which nothing downstream observes. */
  if (inb[8]) { nx += -Bp; ny += 1.0f; nz += -Ap; }
  if (inb[0]) { nx += Bm; ny += 1.0f; nz += Am; }
  if (inb[6]) { nx += -Bp; ny += 1.0f; nz += Am; }
  if (inb[2]) { nx += Bm; ny += 1.0f; nz += -Ap; }
  {
    const float l2 = nx * nx + ny * ny + nz * nz;
    if (sqrtf(l2) > 0) { const float inv = 1.0f / sqrtf(l2); nx *= inv; ny *= inv; nz *= inv; }
  }
  const size_t ci = (size_t)ix * size + iy;
  const float discharge = w->field[4 * ci + 0], fx = w->field[4 * ci + 1], fy = w->field[4 * ci + 2];
  const float root = w->field[4 * ci + 3];

  if ((float)d->age > P->maxAge || d->volume < P->minVol) { /* water.h:74-82 */
    const int32_t q = orc_ls_quantize_height(d->sediment);
    D[4] += q;
    st->fx_deposited += q;
    st->fx_sed_deposited += tq(d->sediment);
    if ((float)d->age > P->maxAge) { st->term_age++; d->flags = ORC_DROP_DONE_AGE; }
    else { st->term_vol++; d->flags = ORC_DROP_DONE_VOL; }
    return 1;
  }
  float effD = P->depositionRate * (1.0f - root); /* :86-87 */
  if (effD < 0) effD = 0;
  effD = effD * w->cur_damp; /* 1, or 2^-n next to n cells holding a higher key (exclusive_cells == 3) */
  {
    const float g = lod * P->gravity; /* :95 */
    d->sx += (g * nx) / d->volume;
    d->sy += (g * nz) / d->volume;
  }
  if (v2len(fx, fy) > 0 && v2len(d->sx, d->sy) > 0) { /* :97-99 */
    const float fi = 1.0f / sqrtf(v2dot(fx, fy, fx, fy));
    const float si = 1.0f / sqrtf(v2dot(d->sx, d->sy, d->sx, d->sy));
    const float dp = v2dot(fx * fi, fy * fi, d->sx * si, d->sy * si);
    const float k = lod * P->momentumTransfer * dp / (d->volume + discharge);
    d->sx += k * fx;
    d->sy += k * fy;
  }
  if (v2len(d->sx, d->sy) > 0) { /* :108-109 */
    const float si = 1.0f / sqrtf(v2dot(d->sx, d->sy, d->sx, d->sy));
    const float m = lod * sqrtf(2.0f);
    d->sx = m * (d->sx * si);
    d->sy = m * (d->sy * si);
  }
  d->px += d->sx; /* :111 */
  d->py += d->sy;

  w->track[ci].discharge = wrap_add(w->track[ci].discharge, trq(d->volume)); /* :115-117 */
  w->track[ci].momentumx = wrap_add(w->track[ci].momentumx, trq(d->volume * d->sx));
  w->track[ci].momentumy = wrap_add(w->track[ci].momentumy, trq(d->volume * d->sy));

  const int nix = trunc_i(d->px), niy = trunc_i(d->py);
  const int out = ls_oob(w, nix, niy);
  float h2;
  if (out) h2 = (float)((double)hc - 0.002); /* :121-122 */
  else {
    const int ddx = nix - ix, ddy = niy - iy;
    if (ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1) h2 = hf(B[(ddx + 1) * 3 + (ddy + 1)]);
    else h2 = hf(R[(size_t)nix * size + niy] + ls_own_delta(own, nix, niy)); /* :124 */
  }
  float c_eq = (1.0f + P->entrainment * orc_erff_poly(0.4f * discharge)) * (hc - h2); /* :127-128 */
  if (c_eq < 0) c_eq = 0;
  const float cdiff = c_eq - d->sediment;
  const float e = effD * cdiff;
  d->sediment += e; /* :131 */
  {
    const int32_t q = orc_ls_quantize_height(-e); /* :132: height += -(effD*cdiff), quantised once */
    D[4] += q;
    st->fx_eroded -= q;
  }
  const float carried = d->sediment;
  if (w->recip_evap) d->sediment = (float)((double)d->sediment * (1.0 / (1.0 - (double)P->evapRate))); /* :135 as a multiplication */
  else d->sediment = (float)((double)d->sediment / (1.0 - (double)P->evapRate)); /* :135 */
  d->volume = (float)((double)d->volume * (1.0 - (double)P->evapRate));     /* :136 */
  st->fx_sed_inflation += tq(d->sediment) - tq(carried);
  if (out) { /* :139-142 */
    st->term_oob++;
    st->fx_sed_oob_lost += tq(d->sediment);
    d->volume = 0.0f;
    d->flags = ORC_DROP_DONE_OOB;
    return 1;
  }
  d->age++; /* :153 */
  d->flags |= ORC_DROP_CASCADE; /* :151, performed at the start of the next phase */
  if (nix < w->row0) d->flags = (d->flags & ~ORC_DROP_ALIVE) | ORC_DROP_MIGRATE_LO;
  else if (nix >= w->row1) d->flags = (d->flags & ~ORC_DROP_ALIVE) | ORC_DROP_MIGRATE_HI;
  return 1;
}

void orc_ls_run(orc_ls_world* w, orc_drop* drops, size_t n, orc_stats* st, float* trace0, int trace_cap, int* trace_n) {
  const int size = w->size;
  const int S = w->steps_per_phase > 1 ? w->steps_per_phase : 1;
  int32_t* deltas = (int32_t*)calloc(n * 9 * S, sizeof(int32_t)); /* this phase: S blocks per drop */
  int* dpos = (int*)calloc(n * 2 * S, sizeof(int));
  unsigned char* has = (unsigned char*)calloc(n * S, 1);
  int tn = 0;
  orc_stats local;
  memset(&local, 0, sizeof(local));
  /* exclusive_cells: claim[c] = highest key {phase tag, phases waited, state hash} of the awake drops on cell c */
  uint32_t* claim = w->exclusive_cells ? (uint32_t*)calloc((size_t)size * size, sizeof(uint32_t)) : NULL;
  for (uint64_t phase = 0;; phase++) {
    const int32_t* R = w->h[phase & 1];
    size_t active = 0;
    if (claim)
      for (size_t i = 0; i < n; i++) {
        if (!(drops[i].flags & ORC_DROP_ALIVE)) continue;
        if (w->align_age && (uint64_t)drops[i].age > phase * S) continue;
        const size_t c = (size_t)trunc_i(drops[i].px) * size + trunc_i(drops[i].py);
        const uint32_t key = ls_claim_key((uint32_t)phase + 1u, &drops[i]);
        if (claim[c] < key) claim[c] = key;
      }
    for (size_t i = 0; i < n; i++) {
      int done = 0; /* steps this drop made in this phase */
      for (int s = 0; s < S; s++) has[i * S + s] = 0;
      if (!(drops[i].flags & ORC_DROP_ALIVE)) continue;
      active++;
      if (claim && !(w->align_age && (uint64_t)drops[i].age > phase * S)) {
        const size_t c = (size_t)trunc_i(drops[i].px) * size + trunc_i(drops[i].py);
        const uint32_t key = ls_claim_key((uint32_t)phase + 1u, &drops[i]);
        int blocked = claim[c] != key; /* another drop has the cell this phase */
        w->cur_damp = 1.0f;
        if (!blocked && w->exclusive_cells >= 2) { /* a drop with a higher key stands on one of the 8 cells around */
          const int x = trunc_i(drops[i].px), y = trunc_i(drops[i].py);
          int crowded = 0; /* cells around that hold a higher key */
          for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
              if ((dx || dy) && !ls_oob(w, x + dx, y + dy) && claim[(size_t)(x + dx) * size + (y + dy)] > key) crowded++;
          if (crowded && w->exclusive_cells == 2) blocked = 1;       /* 2: wait */
          else if (crowded) w->cur_damp = ldexpf(1.0f, -crowded);   /* 3: step, the sediment exchange halved per such cell */
        }
        if (blocked) {
          const int waited = (drops[i].flags >> ORC_DROP_WAITED_SHIFT) & 7;
          drops[i].flags = (drops[i].flags & ~(7 << ORC_DROP_WAITED_SHIFT)) | ((waited < 7 ? waited + 1 : 7) << ORC_DROP_WAITED_SHIFT);
          /* A phase spent waiting is a step of the drop's life not taken: the age advances, so a call never needs
           * more phases than maxAge + 2 however long the queues.  A drop that expires in the queue leaves its
           * sediment where it stands (water.h:74-77). */
          { /* the first free_waits waits of a drop's life do not cost it a step (counter in the flag word) */
            const int fw = (drops[i].flags >> ORC_DROP_FREEW_SHIFT) & 15;
            if (fw < w->free_waits) {
              drops[i].flags = (drops[i].flags & ~(15 << ORC_DROP_FREEW_SHIFT)) | ((fw + 1) << ORC_DROP_FREEW_SHIFT);
              continue;
            }
          }
          drops[i].age++;
          if ((float)drops[i].age > w->p.maxAge) {
            int32_t* D = deltas + 9 * (i * S);
            for (int k = 0; k < 9; k++) D[k] = 0;
            const int32_t q = orc_ls_quantize_height(drops[i].sediment);
            D[4] = q;
            dpos[2 * (i * S)] = trunc_i(drops[i].px);
            dpos[2 * (i * S) + 1] = trunc_i(drops[i].py);
            has[i * S] = 1;
            local.fx_deposited += q;
            local.fx_sed_deposited += tq(drops[i].sediment);
            local.term_age++;
            drops[i].flags = ORC_DROP_DONE_AGE;
          }
          continue;
        }
        drops[i].flags &= ~(7 << ORC_DROP_WAITED_SHIFT);
      }
      for (int s = 0; s < S; s++) {
        if (!(drops[i].flags & ORC_DROP_ALIVE)) break;
        if (w->align_age && (uint64_t)drops[i].age > phase * S + s) continue; /* still asleep: counts as active, does nothing */
        const ls_own own = {deltas + 9 * (i * S), dpos + 2 * (i * S), done};
        has[i * S + done] = (unsigned char)ls_step(w, R, &drops[i], deltas + 9 * (i * S + done), &dpos[2 * (i * S + done)],
                                                   &dpos[2 * (i * S + done) + 1], &local, done ? &own : NULL);
        done++;
        if (drops[i].flags & (ORC_DROP_DONE_AGE | ORC_DROP_DONE_VOL | ORC_DROP_DONE_OOB)) drops[i].flags &= ~ORC_DROP_ALIVE;
        if (i == 0 && trace0 && tn < trace_cap) {
          float* t = trace0 + 7 * (size_t)tn++;
          t[0] = (float)drops[0].age; t[1] = drops[0].px; t[2] = drops[0].py; t[3] = drops[0].sx; t[4] = drops[0].sy;
          t[5] = drops[0].volume; t[6] = drops[0].sediment;
        }
      }
    }
    if (active == 0) break;
    /* The kernel adds every delta to BOTH planes (the one not being read now, and the
     * other one a phase later); sequentially that is just: apply to both. */
    for (size_t i = 0; i < n * S; i++) {
      if (!has[i]) continue;
      for (int k = 0; k < 9; k++) {
        const int32_t v = deltas[9 * i + k];
        if (!v) continue;
        const int x = dpos[2 * i] + k / 3 - 1, y = dpos[2 * i + 1] + k % 3 - 1;
        w->h[0][(size_t)x * size + y] += v;
        w->h[1][(size_t)x * size + y] += v;
      }
    }
    local.phases++;
  }
  if (trace_n) *trace_n = tn;
  if (st) {
    st->steps += local.steps; st->term_age += local.term_age; st->term_vol += local.term_vol;
    st->term_oob += local.term_oob; st->cascade_transfers += local.cascade_transfers; st->phases += local.phases;
    st->fx_eroded += local.fx_eroded; st->fx_deposited += local.fx_deposited;
    st->fx_sed_oob_lost += local.fx_sed_oob_lost; st->fx_sed_deposited += local.fx_sed_deposited; st->fx_sed_inflation += local.fx_sed_inflation;
  }
  free(deltas); free(dpos); free(has); free(claim);
}

void orc_ls_reset_tracks(orc_ls_world* w) { /* world.h:56-61 */
  const size_t n = (size_t)w->size * w->size;
  for (size_t i = 0; i < n; i++) w->track[i].discharge = w->track[i].momentumx = w->track[i].momentumy = 0; /* pad = root count */
}

/* world.h:81-86; returns 1 if a discharge accumulator left the Q13.18 range (the product then
 * reports SHX_ERR_RANGE).  reset != 0 also zeroes the tracks (world.h:56-61 of the next call). */
int orc_ls_ema(orc_ls_world* w, int reset) {
  int overflow = 0;
  const size_t n = (size_t)w->size * w->size;
  const float lr = w->p.lrate;
  for (size_t i = 0; i < n; i++) {
    float* f = w->field + 4 * i;
    f[0] = (1.0f - lr) * f[0] + lr * track_f(w->track[i].discharge);
    f[1] = (1.0f - lr) * f[1] + lr * track_f(w->track[i].momentumx);
    f[2] = (1.0f - lr) * f[2] + lr * track_f(w->track[i].momentumy);
    if (w->track[i].discharge < 0 || w->track[i].discharge > (1 << 30)) overflow = 1;
    if (reset) w->track[i].discharge = w->track[i].momentumx = w->track[i].momentumy = 0;
  }
  return overflow;
}

void orc_ls_erode_spawnlist(orc_ls_world* w, const float* xy, size_t n, orc_stats* st) {
  orc_drop* drops = (orc_drop*)malloc(sizeof(orc_drop) * (n ? n : 1));
  orc_ls_reset_tracks(w);
  orc_ls_make_drops(w, xy, n, drops, st);
  orc_ls_run(w, drops, n, st, NULL, 0, NULL);
  orc_ls_ema(w, 1);
  free(drops);
}

/* World::erode (world.h:54-88) with hash spawns.  At most max_cycles_per_launch (0 = 512) drops per
 * node march together; more cycles run as consecutive batches between ONE reset and ONE EMA
 * (shx_config.max_cycles_per_launch). */
void orc_ls_erode(orc_ls_world* w, int cycles, uint64_t seed, uint64_t epoch, orc_stats* st) {
  const int nodes = w->p.mapsize * w->p.mapsize;
  const int cap = w->max_cycles_per_launch > 0 ? w->max_cycles_per_launch : 512;
  const size_t n = (size_t)nodes * (size_t)cycles;
  float* xy = (float*)malloc(sizeof(float) * 2 * (n ? n : 1));
  float* sub = (float*)malloc(sizeof(float) * 2 * ((size_t)nodes * cap));
  orc_drop* drops = (orc_drop*)malloc(sizeof(orc_drop) * ((size_t)nodes * cap));
  orc_ls_spawn(&w->p, seed, epoch, cycles, xy); /* node-major: drop i of node at node*cycles + i */
  orc_ls_reset_tracks(w);
  for (int i0 = 0; i0 < cycles; i0 += cap) {
    const int m = cycles - i0 < cap ? cycles - i0 : cap;
    for (int node = 0; node < nodes; node++)
      memcpy(sub + 2 * ((size_t)node * m), xy + 2 * ((size_t)node * cycles + i0), sizeof(float) * 2 * m);
    orc_ls_make_drops(w, sub, (size_t)nodes * m, drops, st); /* the 0.1 rejection sees the heights the earlier batches left */
    orc_ls_run(w, drops, (size_t)nodes * m, st, NULL, 0, NULL);
  }
  orc_ls_ema(w, 1);
  free(xy); free(sub); free(drops);
}

/* ================================================================== synthetic terrain */

/* ------------------------------------------------------------------------------------------------
 * Vegetation::grow (vegetation.h:122-188) under the parallel schedule of the device path
 * (simplehydrology_b200/csrc/shx_veg_kernels.cuh): every plant decides from the maps as they are at the start of the
 * frame, rand() is a counter hash keyed (seed, frame, cell, size bits), rootdensity is a count of fifths
 * (orc_track.pad) whose fp32 value is count / 5.  List order: survivors, the seeded plant, children by parent. */
void orc_default_plant_params(orc_plant_params* pp) { /* vegetation.h:40-44 */
  pp->maxSize = 1.5f;
  pp->growRate = 0.05f;
  pp->maxSteep = 0.8f;
  pp->maxDischarge = 0.3f;
  pp->maxTreeHeight = 0.8f;
}

static float veg_height(const orc_ls_world* w, int x, int y) { return hf(w->h[0][(size_t)x * w->size + y]); }

static float veg_normal_y(const orc_ls_world* w, int x, int y) { /* cellpool.h:181-204, map-level oob :413-419 */
  const int size = w->size;
  const float scale = (float)w->p.mapscale;
  const float hc = veg_height(w, x, y);
  const int xm = x > 0, xp = x < size - 1, ym = y > 0, yp = y < size - 1;
  const float hxp = xp ? veg_height(w, x + 1, y) : 0.0f, hxm = xm ? veg_height(w, x - 1, y) : 0.0f;
  const float hyp = yp ? veg_height(w, x, y + 1) : 0.0f, hym = ym ? veg_height(w, x, y - 1) : 0.0f;
  const float Bp = scale * (hxp - hc), Bm = scale * (hxm - hc);
  const float Ap = scale * (hyp - hc), Am = scale * (hym - hc);
  float nx = 0.0f, ny = 0.0f, nz = 0.0f;
  if (xp && yp) { nx += -Bp; ny += 1.0f; nz += -Ap; }
  if (xm && ym) { nx += Bm; ny += 1.0f; nz += Am; }
  if (xp && ym) { nx += -Bp; ny += 1.0f; nz += Am; }
  if (xm && yp) { nx += Bm; ny += 1.0f; nz += -Ap; }
  const float l2 = nx * nx + ny * ny + nz * nz;
  if (l2 > 0.0f) ny *= 1.0f / sqrtf(l2);
  return ny;
}

static float veg_discharge(const orc_ls_world* w, int x, int y) { /* cellpool.h:242-244 with the kernels' erf */
  return orc_erff_poly(0.4f * w->field[4 * ((size_t)x * w->size + y)]);
}

static void veg_stamp(orc_ls_world* w, int x, int y, int sign) { /* Plant::root, vegetation.h:91-120, in fifths */
  for (int dx = -1; dx <= 1; dx++)
    for (int dy = -1; dy <= 1; dy++) {
      const int cx = x + dx, cy = y + dy;
      if (ls_oob(w, cx, cy)) continue;
      w->track[(size_t)cx * w->size + cy].pad += sign * ((dx == 0 && dy == 0) ? 5 : ((dx == 0 || dy == 0) ? 3 : 2));
    }
}

static void veg_refresh(orc_ls_world* w, int x, int y) {
  for (int dx = -1; dx <= 1; dx++)
    for (int dy = -1; dy <= 1; dy++) {
      const int cx = x + dx, cy = y + dy;
      if (ls_oob(w, cx, cy)) continue;
      const size_t i = (size_t)cx * w->size + cy;
      w->field[4 * i + 3] = (float)w->track[i].pad / 5.0f;
    }
}

void orc_veg_sync_counts(orc_ls_world* w) { /* count = round(5 * rootdensity) */
  const size_t n = (size_t)w->size * w->size;
  for (size_t i = 0; i < n; i++) w->track[i].pad = (int32_t)lrintf(w->field[4 * i + 3] * 5.0f);
}

void orc_veg_stamp_list(orc_ls_world* w, const orc_plant* plants, size_t n) { /* root(+1) of every listed plant */
  for (size_t i = 0; i < n; i++) veg_stamp(w, plants[i].x, plants[i].y, +1);
  for (size_t i = 0; i < n; i++) veg_refresh(w, plants[i].x, plants[i].y);
}

size_t orc_veg_grow(orc_ls_world* w, const orc_plant_params* pp, uint64_t seed, uint64_t frame, orc_plant* plants, size_t n,
                    size_t cap, orc_veg_stats* st) {
  const uint64_t key = mix64(mix64(seed) + frame);
  const int size = w->size;
  unsigned* fl = (unsigned*)calloc(n + 1, sizeof(unsigned));
  int* child = (int*)calloc(2 * (n + 1), sizeof(int));
  float* grown = (float*)calloc(n + 1, sizeof(float));
  size_t surv = 0, kids = 0;
  for (size_t i = 0; i < n; i++) { /* decisions from the frozen maps */
    const int x = plants[i].x, y = plants[i].y;
    const float s0 = plants[i].size;
    uint32_t sbits;
    memcpy(&sbits, &s0, 4);
    grown[i] = s0 + pp->growRate * (pp->maxSize - s0); /* vegetation.h:67-69 */
    const uint64_t r = mix64(key + (((uint64_t)(uint32_t)x << 32) | (uint32_t)y) + (uint64_t)sbits * 0x9E3779B97F4A7C15ull);
    const int die = veg_discharge(w, x, y) >= pp->maxDischarge || veg_height(w, x, y) >= pp->maxTreeHeight ||
                    (uint32_t)r % 1000u == 0u; /* :71-78 */
    if (die) {
      fl[i] = 4u;
      continue;
    }
    fl[i] = 1u;
    surv++;
    if ((uint32_t)(r >> 32) % 20u != 0u) continue; /* :157 */
    const uint64_t q = mix64(r);
    const int nx = x + (int)((uint32_t)q % 9u) - 4, ny = y + (int)((uint32_t)(q >> 32) % 9u) - 4; /* :161 */
    if (ls_oob(w, nx, ny)) continue;                              /* :164 */
    if (veg_discharge(w, nx, ny) >= pp->maxDischarge) continue;   /* :167 */
    const uint32_t r5 = (uint32_t)mix64(q) % 1000u;
    if ((double)(float)r5 / 1000.0 <= (double)w->field[4 * ((size_t)nx * size + ny) + 3]) continue; /* :170 */
    if (!(veg_normal_y(w, nx, ny) > pp->maxSteep)) continue;      /* :175 */
    fl[i] |= 2u;
    child[2 * i] = nx;
    child[2 * i + 1] = ny;
    kids++;
  }
  { /* :126-137: one seeding attempt anywhere, Plant::spawn :80-89 */
    const uint64_t r = mix64(key ^ 0x5EED5EED5EED5EEDull);
    const int x = (int)((uint32_t)r % (uint32_t)size), y = (int)((uint32_t)(r >> 32) % (uint32_t)size);
    if (veg_discharge(w, x, y) < pp->maxDischarge && !(veg_normal_y(w, x, y) < pp->maxSteep) && veg_height(w, x, y) < pp->maxTreeHeight) {
      fl[n] = 2u;
      child[2 * n] = x;
      child[2 * n + 1] = y;
      kids++;
    }
  }
  const size_t room = cap > surv ? cap - surv : 0, kept = kids < room ? kids : room;
  orc_plant* out = (orc_plant*)malloc(sizeof(orc_plant) * (surv + kept + 1));
  size_t at = 0;
  for (size_t i = 0; i < n; i++)
    if (fl[i] & 1u) {
      out[at].x = plants[i].x;
      out[at].y = plants[i].y;
      out[at].size = grown[i];
      at++;
    }
  size_t rank = 0;
  for (size_t k = 0; k <= n; k++) { /* newcomers: the seeded plant first, then the children by parent */
    const size_t i = k == 0 ? n : k - 1;
    if (!(fl[i] & 2u)) continue;
    if (rank < kept) {
      out[at].x = child[2 * i];
      out[at].y = child[2 * i + 1];
      out[at].size = 0.0f;
      at++;
      veg_stamp(w, child[2 * i], child[2 * i + 1], +1);
    } else {
      fl[i] &= ~2u; /* refused: no stamp */
    }
    rank++;
  }
  for (size_t i = 0; i < n; i++)
    if (fl[i] & 4u) veg_stamp(w, plants[i].x, plants[i].y, -1);
  for (size_t i = 0; i <= n; i++) {
    if (i < n && (fl[i] & 4u)) veg_refresh(w, plants[i].x, plants[i].y);
    if (fl[i] & 2u) veg_refresh(w, child[2 * i], child[2 * i + 1]);
  }
  memcpy(plants, out, sizeof(orc_plant) * at);
  if (st) {
    st->plants = at;
    st->born = kept;
    st->died = n - surv;
    st->refused = kids - kept;
  }
  free(out);
  free(fl);
  free(child);
  free(grown);
  return at;
}


static uint32_t hash2(uint32_t x, uint32_t y, uint32_t s) {
  uint32_t h = x * 0x9E3779B1u ^ y * 0x85EBCA77u ^ s * 0xC2B2AE3Du;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
static float lattice(uint32_t x, uint32_t y, uint32_t s) { return (float)(hash2(x, y, s) >> 8) * (1.0f / 8388608.0f) - 1.0f; }

static float synth_raw(int x, int y, uint32_t seed) {
  /* 8 octaves, wavelength 256,128,...,2 cells, amplitude 0.6^o (the reference's
   * layer weights, cellpool.h:361-376, on a different -- hash-lattice -- noise). */
  float sum = 0.0f, amp = 0.6f;
  int cell = 256;
  for (int o = 0; o < 8; o++) {
    const int gx = x / cell, gy = y / cell;
    const float fx = (float)(x % cell) / (float)cell, fy = (float)(y % cell) / (float)cell;
    const float ux = fx * fx * (3.0f - 2.0f * fx), uy = fy * fy * (3.0f - 2.0f * fy);
    const uint32_t s = seed * 8u + (uint32_t)o;
    const float v00 = lattice((uint32_t)gx, (uint32_t)gy, s), v01 = lattice((uint32_t)gx, (uint32_t)gy + 1u, s);
    const float v10 = lattice((uint32_t)gx + 1u, (uint32_t)gy, s), v11 = lattice((uint32_t)gx + 1u, (uint32_t)gy + 1u, s);
    const float a = v00 + (v01 - v00) * uy, b = v10 + (v11 - v10) * uy;
    sum = sum + amp * (a + (b - a) * ux);
    amp = amp * 0.6f;
    cell >>= 1;
  }
  return sum;
}

void orc_synth_terrain(float* height, int size, uint32_t seed) {
  const size_t n = (size_t)size * size;
  float mn = INFINITY, mx = -INFINITY;
#pragma omp parallel for reduction(min : mn) reduction(max : mx) schedule(static)
  for (int x = 0; x < size; x++)
    for (int y = 0; y < size; y++) {
      const float v = synth_raw(x, y, seed);
      height[(size_t)x * size + y] = v;
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
    }
  const float range = mx - mn;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) height[i] = (height[i] - mn) / range; /* cellpool.h:408 */
}

void orc_fill_tiled_from_planar(const orc_params* p, const float* planar, orc_cell* tiled) {
  const int size = p->mapsize * p->tilesize;
#pragma omp parallel for schedule(static)
  for (int x = 0; x < size; x++)
    for (int y = 0; y < size; y++) {
      orc_cell* c = tiled + orc_tiled_index(p, x, y);
      memset(c, 0, sizeof(*c));
      c->height = planar[(size_t)x * size + y];
    }
}


/* ================================================================== terrain init (map::init, cellpool.h:349-409)
 *
 * Restates the noise the reference pulls from its vendored FastNoiseLite.h (source/include/FastNoiseLite.h, a
 * header inside /root/reference, so it is pinned by the compiled reference itself: tests/test_oracle_golden.py
 * compares with golden["init_height"], tests/test_oracle_vs_ref.py with a live 2048^2 init):
 *   noise type OpenSimplex2, 3-D overload (z = SEED % 10000), default rotation  FastNoiseLite.h:322-338,687-716
 *   fractal fBm with the constructor's defaults: 3 octaves, lacunarity 2, gain 0.5, bounding 1/1.75  :113-129,866-886
 *   SingleOpenSimplex2 (3-D)  :1054-1148, GradCoord :541-552, Hash :487-504, FastRound :449-450
 * Eight layers, frequency 1,2,..,128, weight 0.6^(o+1) accumulated in fp32 with the reference's double multiply
 * (`scale *= 0.6`), then the min/max normalisation with both extremes starting at 0 (cellpool.h:382-408).
 * All fp32, operation for operation, no contraction. */
static float grad3(int g, int c) { /* component c of gradient g of FastNoiseLite's Gradients3D (64 x {x,y,z,0}) */
  static const int special[4] = {8, 1, 9, 3};
  const int b = g < 60 ? g % 12 : special[g - 60];
  const int grp = b / 4, s = b % 4;
  const float s1 = (s & 1) ? -1.0f : 1.0f, s2 = (s & 2) ? -1.0f : 1.0f;
  if (grp == 0) return c == 0 ? 0.0f : (c == 1 ? s1 : s2);
  if (grp == 1) return c == 1 ? 0.0f : (c == 0 ? s1 : s2);
  return c == 2 ? 0.0f : (c == 0 ? s1 : s2);
}

#define FNL_PX 501125321
#define FNL_PY 1136930381
#define FNL_PZ 1720413743

static int32_t imul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static int32_t iadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t isub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static int fast_round(float f) { return f >= 0 ? (int)(f + 0.5f) : (int)(f - 0.5f); } /* :449-450 */

static float grad_coord(int32_t seed, int32_t xp, int32_t yp, int32_t zp, float xd, float yd, float zd) { /* :541-552 */
  int32_t hash = imul(seed ^ xp ^ yp ^ zp, 0x27d4eb2d); /* :496-504 */
  hash ^= hash >> 15;
  hash &= 63 << 2;
  const int g = hash >> 2;
  return xd * grad3(g, 0) + yd * grad3(g, 1) + zd * grad3(g, 2);
}

static float open_simplex2_3d(int32_t seed, float x, float y, float z) { /* :1054-1148 */
  int32_t i = fast_round(x), j = fast_round(y), k = fast_round(z);
  float x0 = (float)(x - (float)i), y0 = (float)(y - (float)j), z0 = (float)(z - (float)k);
  int xs = (int)(-1.0f - x0) | 1, ys = (int)(-1.0f - y0) | 1, zs = (int)(-1.0f - z0) | 1;
  float ax0 = (float)xs * -x0, ay0 = (float)ys * -y0, az0 = (float)zs * -z0;
  i = imul(i, FNL_PX); j = imul(j, FNL_PY); k = imul(k, FNL_PZ);
  float value = 0;
  float a = (0.6f - x0 * x0) - (y0 * y0 + z0 * z0);
  for (int l = 0;; l++) {
    if (a > 0) value += (a * a) * (a * a) * grad_coord(seed, i, j, k, x0, y0, z0);
    float b = a + 1;
    int32_t i1 = i, j1 = j, k1 = k;
    float x1 = x0, y1 = y0, z1 = z0;
    if (ax0 >= ay0 && ax0 >= az0) { x1 += (float)xs; b -= (float)(xs * 2) * x1; i1 = isub(i1, imul(xs, FNL_PX)); }
    else if (ay0 > ax0 && ay0 >= az0) { y1 += (float)ys; b -= (float)(ys * 2) * y1; j1 = isub(j1, imul(ys, FNL_PY)); }
    else { z1 += (float)zs; b -= (float)(zs * 2) * z1; k1 = isub(k1, imul(zs, FNL_PZ)); }
    if (b > 0) value += (b * b) * (b * b) * grad_coord(seed, i1, j1, k1, x1, y1, z1);
    if (l == 1) break;
    ax0 = 0.5f - ax0; ay0 = 0.5f - ay0; az0 = 0.5f - az0;
    x0 = (float)xs * ax0; y0 = (float)ys * ay0; z0 = (float)zs * az0;
    a += (0.75f - ax0) - (ay0 + az0);
    i = iadd(i, (xs >> 1) & FNL_PX); j = iadd(j, (ys >> 1) & FNL_PY); k = iadd(k, (zs >> 1) & FNL_PZ);
    xs = -xs; ys = -ys; zs = -zs;
    seed = ~seed;
  }
  return value * 32.69428253173828125f;
}

/* GetNoise(x, y, z) of a noise object in the state map::init leaves it in (:322-338) */
static float fnl_get_noise(float frequency, float x, float y, float z) {
  x *= frequency; y *= frequency; z *= frequency; /* :689-691 */
  {
    const float R3 = (float)(2.0 / 3.0); /* :708-715 */
    const float r = (x + y + z) * R3;
    x = r - x; y = r - y; z = r - z;
  }
  int32_t seed = 1337; /* the constructor's default seed (:114), never changed by map::init */
  float sum = 0;
  float amp = 1 / 1.75f; /* mFractalBounding of the constructor (:129); SetFractalType does not recompute it */
  for (int o = 0; o < 3; o++) { /* :872-882 with mWeightedStrength == 0: Lerp(1, ., 0) == 1 exactly */
    const float noise = open_simplex2_3d(seed++, x, y, z);
    sum += noise * amp;
    amp *= 1.0f + 0.0f * ((noise + 1) * 0.5f - 1.0f);
    x *= 2.0f; y *= 2.0f; z *= 2.0f;
    amp *= 0.5f;
  }
  return sum;
}

float orc_init_raw_height(int x, int y, int tilesize, int seed) { /* cellpool.h:354-380 for one cell */
  const float px = (float)x / (float)tilesize, py = (float)y / (float)tilesize; /* :370 */
  const float z = (float)(seed % 10000);
  float h = 0.0f, frequency = 1.0f, scale = 0.6f;
  for (int o = 0; o < 8; o++) {
    h += scale * fnl_get_noise(frequency, px, py, z); /* :371 */
    frequency *= 2;                                   /* :375 */
    scale = (float)((double)scale * 0.6);             /* :376 */
  }
  return h;
}

void orc_init_terrain(float* height, int mapsize, int tilesize, int seed) { /* planar x*size+y */
  const int size = mapsize * tilesize;
  float mn = 0.0f, mx = 0.0f; /* cellpool.h:382-383: both extremes start at 0 */
#pragma omp parallel for reduction(min : mn) reduction(max : mx) schedule(static)
  for (int x = 0; x < size; x++)
    for (int y = 0; y < size; y++) {
      const float v = orc_init_raw_height(x, y, tilesize, seed);
      height[(size_t)x * size + y] = v;
      mn = mn < v ? mn : v;
      mx = mx > v ? mx : v;
    }
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < (size_t)size * size; i++) height[i] = (height[i] - mn) / (mx - mn); /* :408 */
}

/* ---- per-frame views ------------------------------------------------------------------------ */

/* node::height (cellpool.h:237-241): 0 outside the node */
static float node_height(const orc_cell* node, int ts, int lx, int ly) {
  if (lx < 0 || ly < 0 || lx >= ts || ly >= ts) return 0.0f;
  return node[(size_t)lx * ts + ly].height;
}

/* cellpool.h:286-305.  glm semantics as in orc_seq_normal: cross products written out, normalize =
 * v * (1/sqrt(dot(v,v))). */
void orc_vertex_fill(const orc_params* p, const orc_cell* tiled, float* out) {
  const int ts = p->tilesize, ms = p->mapsize;
  const float sc = (float)p->mapscale;
  for (int node = 0; node < ms * ms; node++) {
    const orc_cell* nd = tiled + (size_t)node * ts * ts;
    const int x0 = (node / ms) * ts, y0 = (node % ms) * ts; /* cellpool.h:327-336 */
    for (int lx = 0; lx < ts; lx++)
      for (int ly = 0; ly < ts; ly++) {
        float* v = out + 12 * ((size_t)node * ts * ts + (size_t)lx * ts + ly);
        const float hc = node_height(nd, ts, lx, ly);
        const float hxp = node_height(nd, ts, lx + 1, ly), hxm = node_height(nd, ts, lx - 1, ly);
        const float hyp = node_height(nd, ts, lx, ly + 1), hym = node_height(nd, ts, lx, ly - 1);
        const int xm = lx > 0, xp = lx < ts - 1, ym = ly > 0, yp = ly < ts - 1;
        const float Bp = sc * (hxp - hc), Bm = sc * (hxm - hc), Ap = sc * (hyp - hc), Am = sc * (hym - hc);
        float nx = 0.0f, ny = 0.0f, nz = 0.0f;
        if (xp && yp) { nx += -Bp; ny += 1.0f; nz += -Ap; } /* cellpool.h:187-188 */
        if (xm && ym) { nx += Bm; ny += 1.0f; nz += Am; }   /* :190-191 */
        if (xp && ym) { nx += -Bp; ny += 1.0f; nz += Am; }  /* :194-195 */
        if (xm && yp) { nx += Bm; ny += 1.0f; nz += -Ap; }  /* :197-198 */
        const float l2 = nx * nx + ny * ny + nz * nz;
        if (sqrtf(l2) > 0.0f) { /* :200-201 */
          const float inv = 1.0f / sqrtf(l2);
          nx *= inv; ny *= inv; nz *= inv;
        }
        const float px = (float)(x0 + lx), pz = (float)(y0 + ly), py = sc * hc; /* :290-294 */
        v[0] = px; v[1] = py; v[2] = pz;
        v[3] = nx; v[4] = ny; v[5] = nz;
        v[6] = (float)(x0 + lx + 1) - px; v[7] = sc * hxp - py; v[8] = pz - pz; /* T - P, :295,300 */
        v[9] = px - px; v[10] = sc * hyp - py; v[11] = (float)(y0 + ly + 1) - pz; /* B - P, :296,301 */
      }
  }
}

/* SimpleHydrology.cpp:341-354 (values before TinyEngine packs them into bytes) */
void orc_view_maps(const orc_params* p, const orc_cell* tiled, int erf_poly, float* out) {
  const int size = p->tilesize * p->mapsize;
  for (int x = 0; x < size; x++)
    for (int y = 0; y < size; y++) {
      const orc_cell* c = tiled + orc_tiled_index(p, x, y);
      float* o = out + 4 * ((size_t)x * size + y);
      o[0] = erf_poly ? orc_erff_poly(0.4f * c->discharge) : orc_erff_libm(0.4f * c->discharge); /* cellpool.h:242-244 */
      o[1] = 0.5f * (1.0f + (erf_poly ? orc_erff_poly(c->momentumx) : orc_erff_libm(c->momentumx)));
      o[2] = 0.5f * (1.0f + (erf_poly ? orc_erff_poly(c->momentumy) : orc_erff_libm(c->momentumy)));
      o[3] = c->height;
    }
}
