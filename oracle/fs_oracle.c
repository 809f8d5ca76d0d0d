/*
 * fs_oracle.c -- CPU restatement of futspace's render hot path (TEST INFRASTRUCTURE ONLY).
 * See fs_oracle.h for the parity status ("parity unpinned" for the matte colour maths).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fno-fast-math -fopenmp (oracle/Makefile).
 * -ffp-contract=off matters: every f32 product and sum below is rounded separately, which is
 * the float order north_star defines as "the reference's" (nvcc side: -fmad=false).
 *
 * All file:line citations are into /root/reference.
 */
#include "fs_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * i32.f32 / u32.f32 on values C leaves undefined (SURVEY.md fact 8, 8c last row).
 * ---------------------------------------------------------------------------------------- */
static inline int32_t f2i(float x, int mode) {
  if (mode == FSO_F2I_SATURATE) { /* PTX cvt.rzi.s32.f32: what futhark opencl/cuda emit on NVIDIA */
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
  }
  if (mode == FSO_F2I_X86) { /* cvttss2si: "integer indefinite" for every invalid input */
    if (x != x || x >= 2147483648.0f || x < -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
  }
  /* FSO_F2I_MODERN: futhark >= 0.19 C runtime: isnan||isinf -> 0, else a plain C cast */
  if (x != x || isinf(x)) return 0;
  if (x >= 2147483648.0f || x < -2147483648.0f) return INT32_MIN;
  return (int32_t)x;
}

/* u32.f32 of a value already clamped to [0,255] or NaN (NaN -> 0 on every backend we model). */
static inline uint32_t f2u_channel(float x) {
  if (x != x) return 0u;
  return (uint32_t)x;
}

/* Futhark's integer `%` rounds toward negative infinity: result in [0, n) for n > 0. */
static inline int32_t floored_mod(int32_t a, int32_t n) {
  int32_t m = a % n;
  return m < 0 ? m + n : m;
}

/* ------------------------------------------------------------------------------------------
 * matte 0.1.2 `argb` (restated from the published algorithm; NOT in /root/reference).
 * Call sites that define how it is used: fut/render_functions.fut:101-103 (mix),
 * fut/interactive.fut:163 (scale).
 * ---------------------------------------------------------------------------------------- */
static inline void to_rgba(uint32_t c, float *r, float *g, float *b, float *a) {
  *r = (float)((c >> 16) & 0xFFu) / 255.0f;
  *g = (float)((c >> 8) & 0xFFu) / 255.0f;
  *b = (float)(c & 0xFFu) / 255.0f;
  *a = (float)((c >> 24) & 0xFFu) / 255.0f;
}

static inline float clamp_channel(float x) { /* NaN falls through both comparisons */
  return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
}

static inline uint32_t from_rgba(float r, float g, float b, float a) {
  return (f2u_channel(clamp_channel(a) * 255.0f) << 24) | (f2u_channel(clamp_channel(r) * 255.0f) << 16) |
         (f2u_channel(clamp_channel(g) * 255.0f) << 8) | f2u_channel(clamp_channel(b) * 255.0f);
}

uint32_t fso_mix(float m1, uint32_t c1, float m2, uint32_t c2) {
  float r1, g1, b1, a1, r2, g2, b2, a2;
  to_rgba(c1, &r1, &g1, &b1, &a1);
  to_rgba(c2, &r2, &g2, &b2, &a2);
  float m12 = m1 + m2;
  float m1n = m1 / m12;
  float m2n = m2 / m12;
  float r1s = r1 * r1, r2s = r2 * r2;
  float g1s = g1 * g1, g2s = g2 * g2;
  float b1s = b1 * b1, b2s = b2 * b2;
  float t, u;
  t = m1n * r1s; u = m2n * r2s; float r = sqrtf(t + u);
  t = m1n * g1s; u = m2n * g2s; float g = sqrtf(t + u);
  t = m1n * b1s; u = m2n * b2s; float b = sqrtf(t + u);
  t = m1 * a1;   u = m2 * a2;   float a = (t + u) / m12;
  return from_rgba(r, g, b, a);
}

uint32_t fso_scale(uint32_t c, float s) {
  float r, g, b, a;
  to_rgba(c, &r, &g, &b, &a);
  return from_rgba(r * s, g * s, b * s, a * s);
}

/* ------------------------------------------------------------------------------------------
 * Samplers.  fut/render_functions.fut
 * ---------------------------------------------------------------------------------------- */
/* :63-64  heights[(i32.f32 y)%h, (i32.f32 x)%w]  -- truncation toward zero, then floored mod */
float fso_height_nearest(const int32_t *hm, int q, int r, float x, float y, int m) {
  int32_t iy = floored_mod(f2i(y, m), q), ix = floored_mod(f2i(x, m), r);
  return (float)hm[(size_t)iy * r + ix];
}
/* :91-92 */
uint32_t fso_color_nearest(const uint32_t *cm, int q, int r, float x, float y, int m) {
  int32_t iy = floored_mod(f2i(y, m), q), ix = floored_mod(f2i(x, m), r);
  return cm[(size_t)iy * r + ix];
}
/* :67-77 */
float fso_height_bilinear(const int32_t *hm, int q, int r, float x, float y, int m) {
  float fx = floorf(x), cx = ceilf(x), fy = floorf(y), cy = ceilf(y);
  int32_t x0 = floored_mod(f2i(fx, m), r), x1 = floored_mod(f2i(cx, m), r);
  int32_t y0 = floored_mod(f2i(fy, m), q), y1 = floored_mod(f2i(cy, m), q);
  float wx0 = cx - x, wx1 = x - fx, wy0 = cy - y, wy1 = y - fy;
  float a, b;
  a = wx0 * (float)hm[(size_t)y0 * r + x0];
  b = wx1 * (float)hm[(size_t)y0 * r + x1];
  float xi1 = a + b;
  a = wx0 * (float)hm[(size_t)y1 * r + x0];
  b = wx1 * (float)hm[(size_t)y1 * r + x1];
  float xi2 = a + b;
  a = wy0 * xi1;
  b = wy1 * xi2;
  return a + b;
}
/* :95-105 */
uint32_t fso_color_bilinear(const uint32_t *cm, int q, int r, float x, float y, int m) {
  float fx = floorf(x), cx = ceilf(x), fy = floorf(y), cy = ceilf(y);
  int32_t x0 = floored_mod(f2i(fx, m), r), x1 = floored_mod(f2i(cx, m), r);
  int32_t y0 = floored_mod(f2i(fy, m), q), y1 = floored_mod(f2i(cy, m), q);
  uint32_t i1 = fso_mix(cx - x, cm[(size_t)y0 * r + x0], x - fx, cm[(size_t)y0 * r + x1]);
  uint32_t i2 = fso_mix(cx - x, cm[(size_t)y1 * r + x0], x - fx, cm[(size_t)y1 * r + x1]);
  return fso_mix(cy - y, i1, y - fy, i2);
}

/* ------------------------------------------------------------------------------------------
 * Renderer constants
 * ---------------------------------------------------------------------------------------- */
void fso_params_default(fso_params *p) {
  p->z0 = 0.0f;            /* fut/voxel_renderer.fut:103 */
  p->delta = 0.001f;       /* :104 */
  p->invz_param1 = 1.0f;   /* :217 */
  p->invz_param2 = 0.0f;   /* => f32(w/2), :217 */
  p->filter = FSO_FILTER_BILINEAR;  /* fut/interactive.fut:180-181 */
  p->sentinel = FSO_SENTINEL_ZERO;  /* fut/voxel_renderer.fut:244-248 */
  p->f2i_mode = FSO_F2I_SATURATE;   /* default backend opencl, Makefile:4 */
  p->smoothing = 0;
}
void fso_params_tests_variant(fso_params *p) {
  p->z0 = 1.0f;            /* tests/futspace.fut:84 */
  p->delta = 0.005f;       /* :85 */
  p->invz_param1 = 1.0f;
  p->invz_param2 = 240.0f; /* :91 */
  p->filter = FSO_FILTER_NEAREST;   /* :76-79, :95-96 */
  p->sentinel = FSO_SENTINEL_SKY;   /* :110-121 */
  p->f2i_mode = FSO_F2I_SATURATE;
  p->smoothing = 0;
}

/* ------------------------------------------------------------------------------------------
 * get_zs  fut/voxel_renderer.fut:28-34 (called :108 as get_zs c.distance 0.001 0.0)
 * same series: tests/futspace.fut:47-53, tests/solve_arithm.fut:1-8
 * ---------------------------------------------------------------------------------------- */
static int zs_count(float delta, float dist, float z0) {
  float t = delta - 2.0f * z0;
  float e = 8.0f * delta;
  e = e * dist;
  float s = sqrtf(powf(t, 2.0f) + e);
  float num = s - 2.0f * z0;
  num = num + delta;
  float div = 2.0f * delta;
  float nf = floorf(num / div);
  if (!(nf == nf) || nf > 1.0e8f) return -1;
  if (nf < 0.0f) return -1; /* futhark: `1...n` with n < 0 is a range error */
  return (int)nf;
}
static inline float zs_value(int i1, float delta, float z0) { /* i1 = 1..n */
  float i = (float)i1;
  float a = i / 2.0f;
  float b = 2.0f * z0;
  float c = (i - 1.0f) * delta;
  return a * (b + c);
}
int fso_get_zs(float delta, float dist, float z0, float *out, int cap) {
  int n = zs_count(delta, dist, z0);
  if (n < 0) return -1;
  for (int i = 1; i <= n && i <= cap; ++i) out[i - 1] = zs_value(i, delta, z0);
  return n;
}

/* ------------------------------------------------------------------------------------------
 * Per-depth line set-up: get_h_line fut/voxel_renderer.fut:43-60 and inv_z :217
 * ---------------------------------------------------------------------------------------- */
typedef struct { float sx, sy, dx, dy, inv_z; } depth_line;

static void make_line(const fso_camera *c, const fso_params *p, float z, int w, depth_line *l) {
  float sin_ang = sinf(c->angle), cos_ang = cosf(c->angle);
  float view = c->fov;
  float sv = sin_ang * view, cv = cos_ang * view;
  float left_x = (-cos_ang - sv) * z;
  float left_y = (sin_ang - cv) * z;
  float right_x = (cos_ang - sv) * z;
  float right_y = (-sin_ang - cv) * z;
  l->dx = (right_x - left_x) / (float)w;
  l->dy = (right_y - left_y) / (float)w;
  l->sx = left_x + c->x;
  l->sy = left_y + c->y;
  float mul = p->invz_param2 > 0.0f ? p->invz_param2 : (float)(w / 2);
  l->inv_z = (p->invz_param1 / z) * mul;
}

static depth_line *make_lines(const fso_camera *c, const fso_params *p, int w, int *n_out) {
  int n = zs_count(p->delta, c->distance, p->z0);
  if (n < 0) return NULL;
  depth_line *L = (depth_line *)malloc(sizeof(depth_line) * (size_t)(n > 0 ? n : 1));
  if (!L) return NULL;
  for (int k = 0; k < n; ++k) make_line(c, p, zs_value(k + 1, p->delta, p->z0), w, &L[k]);
  *n_out = n;
  return L;
}

/* one (colour, y) sample: fut/voxel_renderer.fut:219-226 */
static inline int32_t project(const fso_camera *c, const fso_params *p, const depth_line *l, float hgt) {
  float height_diff = c->height - hgt;
  float t = height_diff * l->inv_z;
  float rel = t + c->horizon;
  int32_t y = f2i(rel, p->f2i_mode);
  return y > 0 ? y : 0;
}
static inline void seg_point(const depth_line *l, int i, float *x, float *y) { /* :63-66 */
  float fi = (float)i;
  float a = fi * l->dx, b = fi * l->dy;
  *x = l->sx + a;
  *y = l->sy + b;
}

static int check_args(const fso_camera *cam, const fso_params *prm, const void *a, const void *b, int q,
                      int r, int h, int w, const void *out) {
  if (!cam || !prm || !a || !b || !out) return 1;
  if (q <= 0 || r <= 0 || h <= 0 || w <= 0) return 2;
  if (prm->smoothing && prm->sentinel != FSO_SENTINEL_ZERO) return 2; /* #on exists only in fut/voxel_renderer.fut */
  return 0;
}

/* The pixel of row `row` inside the span of a smoothing tuple (col, prev col, y, prev y, idx, prev idx),
 * fut/voxel_renderer.fut:200-210. */
static inline uint32_t smooth_pixel(uint32_t c, uint32_t cprev, int32_t y, int32_t yprev, int32_t k, int32_t kprev,
                                    int32_t row) {
  if (k - kprev != 1) return c;
  const float range = fmaxf(1.0f, (float)(yprev - y));
  const float delta1 = fabsf((float)row - (float)yprev) / range;
  const float delta2 = fabsf((float)y - (float)row) / range;
  return fso_mix(delta2, cprev, delta1, c);
}

/* Smoothing #on, sequential statement (see fs_oracle.h): per column the list of samples that lower the y-buffer,
 * then rows top-down. */
static int render_smooth(const fso_camera *cam, const fso_params *prm, const uint32_t *color, int cq, int cr,
                         const int32_t *height, int q, int r, int h, int w, uint32_t *out, int eval_all, int nthreads,
                         const depth_line *L, int n) {
  const int bil = prm->filter == FSO_FILTER_BILINEAR, m = prm->f2i_mode;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  (void)nthreads;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    int32_t *rk = (int32_t *)malloc(sizeof(int32_t) * (size_t)(h + 1));
    int32_t *ry = (int32_t *)malloc(sizeof(int32_t) * (size_t)(h + 1));
    uint32_t *rcol = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(h + 1));
#pragma omp for schedule(dynamic, 4)
    for (int j = 0; j < w; ++j) {
      int cnt = 0;
      int32_t ybuf = h;
      for (int k = 0; k < n; ++k) {
        float x, y;
        seg_point(&L[k], j, &x, &y);
        float hgt = bil ? fso_height_bilinear(height, q, r, x, y, m) : fso_height_nearest(height, q, r, x, y, m);
        uint32_t c = 0;
        if (eval_all) c = bil ? fso_color_bilinear(color, cq, cr, x, y, m) : fso_color_nearest(color, cq, cr, x, y, m);
        int32_t yy = project(cam, prm, &L[k], hgt);
        if (yy < ybuf) { /* occlude2 :87-90 keeps the earlier sample on ties */
          if (!eval_all) c = bil ? fso_color_bilinear(color, cq, cr, x, y, m) : fso_color_nearest(color, cq, cr, x, y, m);
          rk[cnt] = k; ry[cnt] = yy; rcol[cnt] = c;
          ++cnt;
          ybuf = yy;
        }
      }
      /* rows above the last record stay at the sentinel tuple: colour 0 -> sky */
      int row = 0;
      const int top = cnt ? ry[cnt - 1] : h;
      for (; row < top; ++row) out[(size_t)row * w + j] = cam->sky_color;
      for (int i = cnt - 1; i >= 0; --i) {
        /* previous running-minimum state: record i-1 or the neutral element (0, h, 0) of :188 */
        const uint32_t cprev = i ? rcol[i - 1] : 0u;
        const int32_t yprev = i ? ry[i - 1] : h, kprev = i ? rk[i - 1] : 0;
        /* the lowering tuple survives the scatter only if no later sample rewrites the row */
        const int keep = (i + 1 < cnt && rk[i + 1] == rk[i] + 1) || rk[i] == n - 1;
        const int end = yprev < h ? yprev : h;
        for (; row < end; ++row) {
          uint32_t px = keep ? smooth_pixel(rcol[i], cprev, ry[i], yprev, rk[i], kprev, row) : rcol[i];
          out[(size_t)row * w + j] = px == 0u ? cam->sky_color : px;
        }
      }
    }
    free(rk); free(ry); free(rcol);
  }
  return 0;
}

int fso_render(const fso_camera *cam, const fso_params *prm, const uint32_t *color, const int32_t *height,
               int q, int r, int h, int w, uint32_t *out, int eval_all, int nthreads) {
  return fso_render_split(cam, prm, color, q, r, height, q, r, h, w, out, eval_all, nthreads);
}

int fso_render_split(const fso_camera *cam, const fso_params *prm, const uint32_t *color, int cq, int cr,
                     const int32_t *height, int q, int r, int h, int w, uint32_t *out, int eval_all, int nthreads) {
  int rc = check_args(cam, prm, color, height, q, r, h, w, out);
  if (rc) return rc;
  if (cq <= 0 || cr <= 0) return 2;
  int n = 0;
  depth_line *L = make_lines(cam, prm, w, &n);
  if (!L) return 3;
  const int bil = prm->filter == FSO_FILTER_BILINEAR, m = prm->f2i_mode;
  const uint32_t empty = prm->sentinel == FSO_SENTINEL_SKY ? cam->sky_color : 0u;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  (void)nthreads;
#endif
  if (prm->smoothing) {
    rc = render_smooth(cam, prm, color, cq, cr, height, q, r, h, w, out, eval_all, nthreads, L, n);
    free(L);
    return rc;
  }
#pragma omp parallel num_threads(nthreads)
  {
    uint32_t *col = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)h);
#pragma omp for schedule(dynamic, 4)
    for (int j = 0; j < w; ++j) {
      for (int i = 0; i < h; ++i) col[i] = empty;
      int32_t ybuf = h;
      for (int k = 0; k < n; ++k) {
        float x, y;
        seg_point(&L[k], j, &x, &y);
        float hgt = bil ? fso_height_bilinear(height, q, r, x, y, m) : fso_height_nearest(height, q, r, x, y, m);
        uint32_t c = 0;
        if (eval_all) c = bil ? fso_color_bilinear(color, cq, cr, x, y, m) : fso_color_nearest(color, cq, cr, x, y, m);
        int32_t yy = project(cam, prm, &L[k], hgt);
        if (yy < ybuf) {
          if (!eval_all) c = bil ? fso_color_bilinear(color, cq, cr, x, y, m) : fso_color_nearest(color, cq, cr, x, y, m);
          col[yy] = c;
          ybuf = yy;
        }
      }
      uint32_t carry = empty;
      for (int i = 0; i < h; ++i) {
        if (col[i] != empty) carry = col[i];
        out[(size_t)i * w + j] = carry == empty ? cam->sky_color : carry;
      }
    }
    free(col);
  }
  free(L);
  return 0;
}

int fso_render_literal(const fso_camera *cam, const fso_params *prm, const uint32_t *color,
                       const int32_t *height, int q, int r, int h, int w, uint32_t *out) {
  int rc = check_args(cam, prm, color, height, q, r, h, w, out);
  if (rc) return rc;
  int n = 0;
  depth_line *L = make_lines(cam, prm, w, &n);
  if (!L) return 3;
  const int bil = prm->filter == FSO_FILTER_BILINEAR, m = prm->f2i_mode;
  const uint32_t empty = prm->sentinel == FSO_SENTINEL_SKY ? cam->sky_color : 0u;
  size_t cells = (size_t)(n > 0 ? n : 1) * (size_t)w;
  uint32_t *cs = (uint32_t *)malloc(cells * 4);
  int32_t *hs = (int32_t *)malloc(cells * 4);
  uint32_t *col = (uint32_t *)malloc((size_t)h * 4);
  if (!cs || !hs || !col) { free(cs); free(hs); free(col); free(L); return 3; }
  /* height_color_map, :215-228 : [n_z][w] */
  for (int k = 0; k < n; ++k)
    for (int i = 0; i < w; ++i) {
      float x, y;
      seg_point(&L[k], i, &x, &y);
      uint32_t c = bil ? fso_color_bilinear(color, q, r, x, y, m) : fso_color_nearest(color, q, r, x, y, m);
      float hgt = bil ? fso_height_bilinear(height, q, r, x, y, m) : fso_height_nearest(height, q, r, x, y, m);
      cs[(size_t)k * w + i] = c;
      hs[(size_t)k * w + i] = project(cam, prm, &L[k], hgt);
    }
  if (prm->smoothing) {
    /* rendered_image2, :186-212, with `futhark c` semantics for the unspecified parts (see fs_oracle.h) */
    typedef struct { uint32_t c, cp; int32_t y, yp, k, kp; } tup;
    const tup sentinel = {0u, 0u, 0, 0, 1000, 1000};
    uint32_t *sc = (uint32_t *)malloc((size_t)(n > 0 ? n : 1) * 4);
    int32_t *sy = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * 4), *sk = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * 4);
    tup *line = (tup *)malloc((size_t)h * sizeof(tup));
    if (!sc || !sy || !sk || !line) { free(sc); free(sy); free(sk); free(line); free(cs); free(hs); free(col); free(L); return 3; }
    for (int j = 0; j < w; ++j) {
      uint32_t ac = 0; int32_t ah = h, ak = 0; /* scan occlude2 (0, h, 0), :188 */
      for (int k = 0; k < n; ++k) {
        const uint32_t c2 = cs[(size_t)k * w + j]; const int32_t h2 = hs[(size_t)k * w + j];
        if (!(ah <= h2)) { ac = c2; ah = h2; ak = k; }
        sc[k] = ac; sy[k] = ah; sk[k] = ak;
      }
      for (int i = 0; i < h; ++i) line[i] = sentinel; /* replicate h (0,0,0,0,1000,1000), :192 */
      for (int k = 0; k < n; ++k) {                   /* rotate (-1): element k pairs with element k-1, wrapping */
        const int p = k ? k - 1 : n - 1;
        const tup t = {sc[k], sc[p], sy[k], sy[p], sk[k], sk[p]};
        if (sy[k] >= 0 && sy[k] < h) line[sy[k]] = t;   /* scatter in index order: the last write wins */
      }
      tup acc = sentinel; /* scan fill_vline3, :193 */
      for (int i = 0; i < h; ++i) {
        const tup t = line[i];
        if (!(t.c == 0u && t.cp == 0u && t.y == 0 && t.yp == 0 && t.k == 1000 && t.kp == 1000)) acc = t;
        const uint32_t px = smooth_pixel(acc.c, acc.cp, acc.y, acc.yp, acc.k, acc.kp, i);
        out[(size_t)i * w + j] = px == 0u ? cam->sky_color : px; /* :211 */
      }
    }
    free(sc); free(sy); free(sk); free(line); free(cs); free(hs); free(col); free(L);
    return 0;
  }
  /* rendered_image, :229-250, over (transpose height_color_map) */
  for (int j = 0; j < w; ++j) {
    /* scan occlude (0,h) : inclusive; res[0] = xs[0] */
    uint32_t ac = 0; int32_t ah = 0;
    for (int i = 0; i < h; ++i) col[i] = empty;  /* replicate h 0 | replicate l sky */
    for (int k = 0; k < n; ++k) {
      uint32_t c2 = cs[(size_t)k * w + j]; int32_t h2 = hs[(size_t)k * w + j];
      if (k == 0 || !(ah <= h2)) { ac = c2; ah = h2; }  /* occlude :69-72 */
      if (ah >= 0 && ah < h) col[ah] = ac;              /* scatter :244 ignores out-of-range */
    }
    /* scan fill_vline 0 :246 ; fill :110-113 in the sentinel variant */
    uint32_t acc = 0;
    for (int i = 0; i < h; ++i) {
      acc = (i == 0) ? col[0] : (col[i] == empty ? acc : col[i]);
      /* :248 sky map (identity in the sky-sentinel variant) */
      out[(size_t)i * w + j] = (prm->sentinel == FSO_SENTINEL_ZERO && acc == 0u) ? cam->sky_color : acc;
    }
  }
  free(cs); free(hs); free(col); free(L);
  return 0;
}

/* fut/effects.fut:6-25 */
static void rot(char axis, float ang, const float v[3], float o[3]) {
  float c = cosf(ang), s = sinf(ang);
  float R[3][3];
  if (axis == 'x') {
    float t[3][3] = {{1, 0, 0}, {0, c, -s}, {0, s, c}};
    memcpy(R, t, sizeof R);
  } else if (axis == 'y') {
    float t[3][3] = {{c, 0, s}, {0, 1, 0}, {-s, 0, c}};
    memcpy(R, t, sizeof R);
  } else {
    float t[3][3] = {{c, -s, 0}, {s, c, 0}, {0, 0, 1}};
    memcpy(R, t, sizeof R);
  }
  for (int i = 0; i < 3; ++i) { /* matvecmul_row: reduce (+) 0 (map2 (*) row v) */
    float a = R[i][0] * v[0], b = R[i][1] * v[1], d = R[i][2] * v[2];
    float acc = 0.0f + a;
    acc = acc + b;
    o[i] = acc + d;
  }
}
void fso_sun_vector(float sun_height, float sun_ang, float out[3]) {
  const float sun0[3] = {0.0f, 1.0f, 0.0f}; /* fut/interactive.fut:56 */
  float t[3];
  rot('z', sun_height, sun0, t);
  rot('y', sun_ang, t, out);
}

/* fut/effects.fut:108-125 */
void fso_bake_shadows(const uint32_t *color, const int32_t *height, int q, int r, const float sun[3], int out_q,
                      int out_r, uint32_t *out, int nthreads) {
  const float step_size = (float)(1024 / 256); /* f32.i32 (max_dist / steps_per_ray), :109-111 */
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int y = 0; y < out_q; ++y)
    for (int x = 0; x < out_r; ++x) {
      float fx = (float)x, fy = (float)y;
      float h0 = fso_height_nearest(height, q, r, fx, fy, FSO_F2I_SATURATE);
      int count = 0;
      for (int dist = 1; dist < 256; ++dist) { /* (1..<steps_per_ray), :118-121 */
        float fd = (float)dist;
        float t = fd * step_size;
        float lift = t * sun[1];
        float sx = t * sun[0], sy = t * sun[2];
        float hh = fso_height_nearest(height, q, r, fx + sx, fy + sy, FSO_F2I_SATURATE);
        float lhs = h0 + lift;
        if (lhs - hh < -0.5f) ++count;
      }
      float amount = (float)count;
      out[(size_t)y * out_r + x] =
          fso_mix(step_size * amount, 0xFF000000u, 1.0f, fso_color_nearest(color, q, r, fx, fy, FSO_F2I_SATURATE)); /* :123 */
    }
}

/* fut/effects.fut:27-45 */
int fso_interpolate(int pd, const uint32_t *img, int h, int w, uint32_t *out) {
  if (h > w) return 1; /* `(x +- pd) % h` (:36-41) would index past the row */
#define PX(yy, xx) img[(size_t)(yy) * w + (xx)]
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int yu = floored_mod(y - pd, h), yd = floored_mod(y + pd, h);
      int xl = floored_mod(x - pd, h), xr = floored_mod(x + pd, h); /* sic: % h */
      uint32_t c = PX(y, x), u = PX(yu, x), d = PX(yd, x), l = PX(y, xl), r = PX(y, xr);
      uint32_t ul = PX(yu, xl), ur = PX(yu, xr), dl = PX(yd, xl), dr = PX(yd, xr);
      uint32_t acc = fso_mix(1.0f, dl, 1.0f, dr);
      acc = fso_mix(1.0f, ur, 1.0f, acc);
      acc = fso_mix(1.0f, ul, 1.0f, acc);
      acc = fso_mix(1.0f, l, 1.0f, acc);
      acc = fso_mix(1.0f, r, 1.0f, acc);
      acc = fso_mix(1.0f, d, 1.0f, acc);
      acc = fso_mix(1.0f, u, 1.0f, acc);
      out[(size_t)y * w + x] = fso_mix(1.0f, c, 1.0f, acc);
    }
#undef PX
  return 0;
}
/* fut/effects.fut:47-52: rotate (-1) xs puts xs[i-1] at i */
void fso_interpolate2(const uint32_t *img, int h, int w, uint32_t *out) {
  for (int y = 0; y < h; ++y) {
    const uint32_t *mids = img + (size_t)y * w;
    const uint32_t *highs = img + (size_t)floored_mod(y - 1, h) * w, *lows = img + (size_t)floored_mod(y + 1, h) * w;
    for (int x = 0; x < w; ++x) {
      uint32_t l = mids[floored_mod(x - 1, w)], c = mids[x], r = mids[floored_mod(x + 1, w)], u = highs[x], d = lows[x];
      out[(size_t)y * w + x] = fso_mix(1.0f, l, 1.0f, fso_mix(1.0f, c, 1.0f, fso_mix(1.0f, r, 1.0f, fso_mix(1.0f, u, 1.0f, d))));
    }
  }
}

void fso_mask_heights(int32_t *hm, long n) { /* fut/interactive.fut:189 */
  for (long i = 0; i < n; ++i) hm[i] &= 0xFF;
}
