/*
 * fsb_futhark_shim.c -- the generated-API names of libfutspace.h on top of the native C-ABI, plus the
 * host-side state machine of fut/interactive.fut (a few floats and key flags; no throughput).
 * Compiled with -ffp-contract=off: the camera arithmetic of process_inputs (fut/interactive.fut:89-125)
 * rounds like the reference's scalar f32 code.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/futspace_b200.h"
#include "../../include/libfutspace.h"

struct futhark_context_config { int device; };
#define POOL_SLOTS 4
struct futhark_context {
  fsb_context *fsb;
  char *error;
  /* frames are returned to a small pool instead of cudaFree (which synchronises the device): the lys loop
   * allocates and frees one frame per displayed image */
  struct { void *dev; size_t bytes; } pool[POOL_SLOTS];
  /* futhark_values_u32_2d of a device frame: asynchronous copy into pinned staging, delivered to the caller's
   * (pageable) buffer by futhark_context_sync -- the generated API's contract is "valid after sync" */
  void *staging;
  size_t staging_bytes;
  uint32_t *pending_dst;
  size_t pending_bytes;
  fsb_map *dummy;           /* init's placeholder landscape [[0],[0]] (fut/interactive.fut:38-43), created on first use */
};
struct futhark_u32_2d { int64_t shape[2]; uint32_t *host; uint32_t *dev; };
struct futhark_i32_2d { int64_t shape[2]; int32_t *host; };

/* landscape, fut/voxel_renderer.fut:14-18: shared (reference-counted) between successive states */
struct landscape {
  int refs;
  int q, r;
  uint32_t *color;     /* lsc.color                     */
  int32_t *altitude;   /* lsc.altitude (masked & 0xFF)  */
  fsb_map *plain;      /* colour + altitude on the device: source of the shadow bake */
  fsb_map *shadowed;   /* lsc.shadowed_color + altitude: what `render` draws (fut/interactive.fut:180-182) */
};

enum { K_A, K_B, K_C, K_D, K_E, K_F, K_G, K_H, K_I, K_J, K_K, K_L, K_M, K_N, K_O, K_P, K_Q, K_R, K_S, K_T, K_U, K_V,
       K_W, K_X, K_Y, K_Z, K_UP, K_LEFT, K_DOWN, K_RIGHT, K_1, K_2, K_3, K_4, K_5, K_6, K_7, K_8, K_9, K_0, K_COUNT };

struct futhark_opaque_state { /* sized_state, fut/interactive.fut:9-22 */
  int mode_math, smoothing_on;
  fsb_camera cam;
  struct landscape *lsc;
  int32_t height, width;
  int8_t inputs[K_COUNT];   /* fut/interactive_input.fut:6-48 */
  float random, sun_height, sun_ang, sun[3];
};

static int fail(struct futhark_context *ctx, const char *msg) {
  if (ctx) {
    free(ctx->error);
    ctx->error = msg ? strdup(msg) : NULL;
  }
  return 1;
}
static int fail_fsb(struct futhark_context *ctx) { return fail(ctx, fsb_context_get_error(ctx->fsb)); }

/* ---------------------------------------------------------------- context */
struct futhark_context_config *futhark_context_config_new(void) { return (struct futhark_context_config *)calloc(1, sizeof(struct futhark_context_config)); }
void futhark_context_config_free(struct futhark_context_config *cfg) { free(cfg); }
void futhark_context_config_set_device(struct futhark_context_config *cfg, const char *s) {
  if (cfg && s) cfg->device = atoi(s[0] == '#' ? s + 1 : s);
}
void futhark_context_config_set_debugging(struct futhark_context_config *cfg, int flag) { (void)cfg; (void)flag; }
void futhark_context_config_set_profiling(struct futhark_context_config *cfg, int flag) { (void)cfg; (void)flag; }
void futhark_context_config_set_logging(struct futhark_context_config *cfg, int flag) { (void)cfg; (void)flag; }

struct futhark_context *futhark_context_new(struct futhark_context_config *cfg) {
  struct futhark_context *ctx = (struct futhark_context *)calloc(1, sizeof *ctx);
  if (!ctx) return NULL;
  int rc = fsb_context_new(cfg ? cfg->device : 0, &ctx->fsb);
  if (rc) { /* like the generated API: the context exists, the error is reported by futhark_context_get_error */
    char msg[96];
    snprintf(msg, sizeof msg, "futspace_b200: no usable sm_100 GPU (fsb_context_new returned %d); there is no CPU fallback", rc);
    fail(ctx, msg);
  }
  return ctx;
}
static void *pool_get(struct futhark_context *ctx, size_t bytes) {
  for (int i = 0; i < POOL_SLOTS; ++i)
    if (ctx->pool[i].dev && ctx->pool[i].bytes == bytes) {
      void *p = ctx->pool[i].dev;
      ctx->pool[i].dev = NULL;
      return p;
    }
  void *p = NULL;
  return fsb_device_malloc(ctx->fsb, bytes, &p) ? NULL : p;
}
static void pool_put(struct futhark_context *ctx, void *dev, size_t bytes) {
  for (int i = 0; i < POOL_SLOTS; ++i)
    if (!ctx->pool[i].dev) {
      ctx->pool[i].dev = dev;
      ctx->pool[i].bytes = bytes;
      return;
    }
  fsb_device_free(ctx->fsb, dev);
}
static int flush_pending(struct futhark_context *ctx) { /* after a stream sync: staging -> caller's buffer */
  if (ctx->pending_dst) {
    memcpy(ctx->pending_dst, ctx->staging, ctx->pending_bytes);
    ctx->pending_dst = NULL;
  }
  return 0;
}

void futhark_context_free(struct futhark_context *ctx) {
  if (!ctx) return;
  if (ctx->fsb) {
    fsb_context_sync(ctx->fsb);
    flush_pending(ctx);
    for (int i = 0; i < POOL_SLOTS; ++i)
      if (ctx->pool[i].dev) fsb_device_free(ctx->fsb, ctx->pool[i].dev);
    if (ctx->staging) fsb_host_free(ctx->fsb, ctx->staging);
    if (ctx->dummy) fsb_map_free(ctx->fsb, ctx->dummy);
  }
  fsb_context_free(ctx->fsb);
  free(ctx->error);
  free(ctx);
}
int futhark_context_clear_caches(struct futhark_context *ctx) {
  if (!ctx || !ctx->fsb) return fail(ctx, "no device context");
  if (fsb_context_sync(ctx->fsb)) return fail_fsb(ctx);
  flush_pending(ctx);
  for (int i = 0; i < POOL_SLOTS; ++i)
    if (ctx->pool[i].dev) {
      fsb_device_free(ctx->fsb, ctx->pool[i].dev);
      ctx->pool[i].dev = NULL;
      ctx->pool[i].bytes = 0;
    }
  if (ctx->staging) fsb_host_free(ctx->fsb, ctx->staging);
  ctx->staging = NULL;
  ctx->staging_bytes = 0;
  return 0;
}
char *futhark_context_report(struct futhark_context *ctx) {
  char buf[256], name[128] = "no device";
  if (ctx && ctx->fsb) fsb_context_device_name(ctx->fsb, name, sizeof name);
  snprintf(buf, sizeof buf, "futspace_b200 on %s: %lld kernel launches\n", name,
           (long long)(ctx && ctx->fsb ? fsb_context_launch_count(ctx->fsb) : 0));
  return strdup(buf);
}
void futhark_context_pause_profiling(struct futhark_context *ctx) { (void)ctx; }
void futhark_context_unpause_profiling(struct futhark_context *ctx) { (void)ctx; }

int futhark_context_sync(struct futhark_context *ctx) {
  if (!ctx || !ctx->fsb) return fail(ctx, "no device context");
  if (fsb_context_sync(ctx->fsb)) return fail_fsb(ctx);
  return flush_pending(ctx);
}
char *futhark_context_get_error(struct futhark_context *ctx) {
  if (!ctx) return NULL;
  char *e = ctx->error;
  ctx->error = NULL;
  return e;
}

/* ---------------------------------------------------------------- arrays */
struct futhark_u32_2d *futhark_new_u32_2d(struct futhark_context *ctx, const uint32_t *data, int64_t d0, int64_t d1) {
  (void)ctx;
  if (!data || d0 <= 0 || d1 <= 0) return NULL;
  struct futhark_u32_2d *a = (struct futhark_u32_2d *)calloc(1, sizeof *a);
  if (!a) return NULL;
  a->shape[0] = d0; a->shape[1] = d1;
  a->host = (uint32_t *)malloc((size_t)d0 * d1 * 4);
  if (!a->host) { free(a); return NULL; }
  memcpy(a->host, data, (size_t)d0 * d1 * 4);
  return a;
}
int futhark_free_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *a) {
  if (!a) return 0;
  if (a->dev && ctx && ctx->fsb) pool_put(ctx, a->dev, (size_t)a->shape[0] * a->shape[1] * 4);
  free(a->host);
  free(a);
  return 0;
}
int futhark_values_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *a, uint32_t *data) {
  if (!a || !data) return fail(ctx, "futhark_values_u32_2d: NULL argument");
  const size_t bytes = (size_t)a->shape[0] * a->shape[1] * 4;
  if (a->host) { memcpy(data, a->host, bytes); return 0; }
  if (!ctx || !ctx->fsb) return fail(ctx, "no device context");
  if (ctx->pending_dst) { /* an earlier values() not yet synced: deliver it first */
    if (fsb_context_sync(ctx->fsb)) return fail_fsb(ctx);
    flush_pending(ctx);
  }
  if (fsb_host_is_registered(ctx->fsb, data)) { /* page-locked destination (fsb_host_register): DMA straight into it */
    if (fsb_copy_to_host(ctx->fsb, data, a->dev, bytes)) return fail_fsb(ctx); /* asynchronous until futhark_context_sync */
    return 0;
  }
  if (bytes > ctx->staging_bytes) {
    if (ctx->staging) fsb_host_free(ctx->fsb, ctx->staging);
    ctx->staging = NULL;
    ctx->staging_bytes = 0;
    if (fsb_host_malloc(ctx->fsb, bytes, &ctx->staging)) return fail_fsb(ctx);
    ctx->staging_bytes = bytes;
  }
  if (fsb_copy_to_host(ctx->fsb, ctx->staging, a->dev, bytes)) return fail_fsb(ctx); /* asynchronous until futhark_context_sync */
  ctx->pending_dst = data;
  ctx->pending_bytes = bytes;
  return 0;
}
const int64_t *futhark_shape_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *a) { (void)ctx; return a ? a->shape : NULL; }

struct futhark_i32_2d *futhark_new_i32_2d(struct futhark_context *ctx, const int32_t *data, int64_t d0, int64_t d1) {
  (void)ctx;
  if (!data || d0 <= 0 || d1 <= 0) return NULL;
  struct futhark_i32_2d *a = (struct futhark_i32_2d *)calloc(1, sizeof *a);
  if (!a) return NULL;
  a->shape[0] = d0; a->shape[1] = d1;
  a->host = (int32_t *)malloc((size_t)d0 * d1 * 4);
  if (!a->host) { free(a); return NULL; }
  memcpy(a->host, data, (size_t)d0 * d1 * 4);
  return a;
}
int futhark_free_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *a) { (void)ctx; if (a) { free(a->host); free(a); } return 0; }
int futhark_values_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *a, int32_t *data) {
  if (!a || !data) return fail(ctx, "futhark_values_i32_2d: NULL argument");
  memcpy(data, a->host, (size_t)a->shape[0] * a->shape[1] * 4);
  return 0;
}
const int64_t *futhark_shape_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *a) { (void)ctx; return a ? a->shape : NULL; }

/* ---------------------------------------------------------------- state */
static void lsc_release(struct futhark_context *ctx, struct landscape *l) {
  if (!l || --l->refs > 0) return;
  if (ctx && ctx->fsb) {
    fsb_map_free(ctx->fsb, l->plain);
    fsb_map_free(ctx->fsb, l->shadowed);
  }
  free(l->color);
  free(l->altitude);
  free(l);
}
static struct futhark_opaque_state *clone_state(const struct futhark_opaque_state *s) {
  struct futhark_opaque_state *n = (struct futhark_opaque_state *)malloc(sizeof *n);
  if (!n) return NULL;
  *n = *s;
  if (n->lsc) ++n->lsc->refs;
  return n;
}
int futhark_free_opaque_state(struct futhark_context *ctx, struct futhark_opaque_state *s) {
  if (!s) return 0;
  lsc_release(ctx, s->lsc);
  free(s);
  return 0;
}

/* matte argb.scale (restated; see oracle/fs_oracle.c fso_scale): from_rgba (r*s) (g*s) (b*s) (a*s) */
static uint32_t argb_scale(uint32_t c, float s) {
  uint32_t out = 0;
  const int shifts[4] = {24, 16, 8, 0};
  for (int i = 0; i < 4; ++i) {
    float v = ((float)((c >> shifts[i]) & 0xFFu) / 255.0f) * s;
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    v = v * 255.0f;
    out |= (v != v ? 0u : (uint32_t)v) << shifts[i];
  }
  return out;
}

/* i32.f32 as the reference's default GPU back-end converts (saturating, NaN -> 0) */
static int32_t f2i_sat(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return INT32_MAX;
  if (x <= -2147483648.0f) return INT32_MIN;
  return (int32_t)x;
}
static int32_t floored_mod(int32_t a, int32_t n) { int32_t m = a % n; return m < 0 ? m + n : m; }

/* shadowed_color = generate_shadowmap_accumulated ... (vec3_rotate #y sun_ang (vec3_rotate #z sun_height sun)),
 * fut/interactive.fut:126-145,194-196; s.sun stays [0,1,0] (:56), so fsb_sun_vector applies. */
static int rebake(struct futhark_context *ctx, struct landscape *l, float sun_height, float sun_ang) {
  if (!ctx->fsb) return fail(ctx, "no device context");
  float sun[3];
  fsb_sun_vector(sun_height, sun_ang, sun);
  uint32_t *sh = (uint32_t *)malloc((size_t)1024 * 1024 * 4); /* the reference bakes 1024 x 1024, fut/effects.fut:124-125 */
  if (!sh) return fail(ctx, "out of memory");
  int rc = fsb_map_bake_shadows(ctx->fsb, l->plain, sun, 1024, 1024, sh);
  fsb_map *m = NULL;
  if (!rc) {
    /* lsc.shadowed_color is 1024 x 1024 whatever the map size; lsc.altitude keeps the map's size, and `render` wraps
     * each by its own (fut/interactive.fut:180-181, fut/render_functions.fut:67-77,95-105) */
    rc = fsb_map_new_split(ctx->fsb, sh, 1024, 1024, l->altitude, l->q, l->r, 0, &m);
  }
  free(sh);
  if (rc) return fail_fsb(ctx);
  if (l->shadowed) fsb_map_free(ctx->fsb, l->shadowed);
  l->shadowed = m;
  return 0;
}

int futhark_entry_init(struct futhark_context *ctx, struct futhark_opaque_state **out0, const uint32_t seed) {
  (void)seed;
  struct futhark_opaque_state *s = (struct futhark_opaque_state *)calloc(1, sizeof *s);
  if (!s) return fail(ctx, "out of memory");
  s->cam.x = 0.98f; s->cam.y = 0.6f; s->cam.height = 58.0f; s->cam.angle = 2.2f; /* fut/interactive.fut:29-36 */
  s->cam.horizon = 200.0f; s->cam.distance = 800.0f; s->cam.fov = 1.2f; s->cam.sky_color = 0;
  s->height = 1024; s->width = 1024;                                              /* :50-51 */
  s->random = 1.0f; s->sun_height = 0.1f; s->sun_ang = 0.1f;                      /* :53-55 */
  s->sun[0] = 0.0f; s->sun[1] = 1.0f / sqrtf(1.0f); s->sun[2] = 0.0f;             /* :56 */
  *out0 = s;
  return 0;
}

int futhark_entry_resize(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t h,
                         const int32_t w, const struct futhark_opaque_state *in) {
  struct futhark_opaque_state *s = clone_state(in);
  if (!s) return fail(ctx, "out of memory");
  s->height = h; s->width = w;                                                    /* :59-61 */
  *out0 = s;
  return 0;
}

static int key_index(int32_t key) { /* SDL keycodes as lys exports them (SDLK_*) */
  if (key >= 'a' && key <= 'z') return K_A + (key - 'a');
  if (key >= '1' && key <= '9') return K_1 + (key - '1');
  if (key == '0') return K_0;
  switch (key) {
    case 0x40000052: return K_UP;
    case 0x40000050: return K_LEFT;
    case 0x40000051: return K_DOWN;
    case 0x4000004F: return K_RIGHT;
    default: return -1;
  }
}

int futhark_entry_key(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t e,
                      const int32_t key, const struct futhark_opaque_state *in) {
  struct futhark_opaque_state *s = clone_state(in);
  if (!s) return fail(ctx, "out of memory");
  const int k = key_index(key);
  if (k >= 0) s->inputs[k] = e == 0 ? 1 : 0; /* e == 0 is keydown, fut/interactive_entrypoints.fut:9-11 */
  *out0 = s;
  return 0;
}
int futhark_entry_mouse(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t b, const int32_t x,
                        const int32_t y, const struct futhark_opaque_state *in) {
  (void)b; (void)x; (void)y;
  return (*out0 = clone_state(in)) ? 0 : fail(ctx, "out of memory"); /* event `_ -> s`, fut/interactive.fut:171 */
}
int futhark_entry_wheel(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t dx, const int32_t dy,
                        const struct futhark_opaque_state *in) {
  (void)dx; (void)dy;
  return (*out0 = clone_state(in)) ? 0 : fail(ctx, "out of memory");
}

int futhark_entry_step(struct futhark_context *ctx, struct futhark_opaque_state **out0, const float td,
                       const struct futhark_opaque_state *in) {
  (void)td;
  struct futhark_opaque_state *s = clone_state(in);
  if (!s) return fail(ctx, "out of memory");
  /* step, fut/interactive.fut:161-164 */
  s->random = in->random + 0.005f;
  s->cam.sky_color = argb_scale(0xFF9090e0u, in->sun_height);
  /* process_inputs, :89-159 -- every right-hand side reads the state before this step */
  const int8_t *k = in->inputs;
  const float sa = sinf(in->cam.angle), ca = cosf(in->cam.angle);
  if (k[K_W] == 1) { s->cam.x = in->cam.x - 3.0f * sa; s->cam.y = in->cam.y - 3.0f * ca; }
  else if (k[K_S] == 1) { s->cam.x = in->cam.x + 3.0f * sa; s->cam.y = in->cam.y + 3.0f * ca; }
  if (k[K_D] == 1) s->cam.angle = in->cam.angle - 0.10f; else if (k[K_A] == 1) s->cam.angle = in->cam.angle + 0.10f;
  if (k[K_E] == 1) s->cam.horizon = in->cam.horizon - 20.0f; else if (k[K_Q] == 1) s->cam.horizon = in->cam.horizon + 20.0f;
  if (k[K_R] == 1) s->cam.height = in->cam.height + 10.0f; else if (k[K_F] == 1) s->cam.height = in->cam.height - 10.0f;
  if (k[K_UP] == 1) s->cam.distance = in->cam.distance + 30.0f; else if (k[K_DOWN] == 1) s->cam.distance = in->cam.distance - 30.0f;
  if (k[K_O] == 1) s->cam.fov = in->cam.fov + 0.1f; else if (k[K_L] == 1) s->cam.fov = in->cam.fov - 0.1f;
  if (k[K_U] == 1) s->sun_height = in->sun_height + 0.005f; else if (k[K_J] == 1) s->sun_height = in->sun_height - 0.005f;
  if (k[K_N] == 1) s->sun_ang = in->sun_ang + 0.005f; else if (k[K_M] == 1) s->sun_ang = in->sun_ang - 0.005f;
  if (k[K_1] == 1) s->mode_math = !in->mode_math;
  if (k[K_2] == 1) s->smoothing_on = !in->smoothing_on;
  if ((k[K_U] == 1 || k[K_J] == 1 || k[K_N] == 1 || k[K_M] == 1) && in->lsc) {
    /* :126-145: a fresh shadowed_color from the OLD sun angles; the landscape object is copied on write */
    struct landscape *l = (struct landscape *)calloc(1, sizeof *l);
    if (!l) { futhark_free_opaque_state(ctx, s); return fail(ctx, "out of memory"); }
    const size_t n = (size_t)in->lsc->q * in->lsc->r;
    l->refs = 1; l->q = in->lsc->q; l->r = in->lsc->r;
    l->color = (uint32_t *)malloc(n * 4); l->altitude = (int32_t *)malloc(n * 4);
    int rc = (!l->color || !l->altitude) ? fail(ctx, "out of memory") : 0;
    if (!rc) {
      memcpy(l->color, in->lsc->color, n * 4);
      memcpy(l->altitude, in->lsc->altitude, n * 4);
      if (!ctx->fsb) rc = fail(ctx, "no device context");
      else if (fsb_map_new(ctx->fsb, l->color, l->altitude, l->q, l->r, 0, &l->plain)) rc = fail_fsb(ctx);
      else rc = rebake(ctx, l, in->sun_height, in->sun_ang);
    }
    lsc_release(ctx, s->lsc);
    s->lsc = l;
    if (rc) { futhark_free_opaque_state(ctx, s); return rc; }
  }
  /* terrain_collision, :67-87 (#png): the camera never sinks below the texel under it */
  {
    float terrain = 0.0f;
    const int32_t x = f2i_sat(s->cam.x), y = f2i_sat(s->cam.y);
    if (s->lsc) terrain = (float)s->lsc->altitude[(size_t)floored_mod(y, s->lsc->q) * s->lsc->r + floored_mod(x, s->lsc->r)];
    if (s->cam.height <= terrain) s->cam.height = terrain;
  }
  *out0 = s;
  return 0;
}

int futhark_entry_text_content(struct futhark_context *ctx, float *o0, float *o1, float *o2, float *o3, float *o4,
                               float *o5, float *o6, float *o7, float *o8, const struct futhark_opaque_state *s) {
  (void)ctx; /* fut/interactive.fut:185-186 */
  *o0 = s->cam.x; *o1 = s->cam.y; *o2 = s->cam.angle; *o3 = s->cam.height; *o4 = s->cam.horizon;
  *o5 = s->cam.distance; *o6 = s->sun_height; *o7 = s->sun_ang; *o8 = s->cam.fov;
  return 0;
}

int futhark_entry_update_map(struct futhark_context *ctx, struct futhark_opaque_state **out0,
                             const struct futhark_u32_2d *color, const struct futhark_i32_2d *height,
                             const struct futhark_opaque_state *in) {
  if (!ctx || !color || !height || !in || !color->host) return fail(ctx, "update_map: NULL argument");
  if (color->shape[0] != height->shape[0] || color->shape[1] != height->shape[1]) return fail(ctx, "update_map: shape mismatch");
  if (!ctx->fsb) return fail(ctx, "no device context");
  struct landscape *l = (struct landscape *)calloc(1, sizeof *l);
  if (!l) return fail(ctx, "out of memory");
  const size_t n = (size_t)color->shape[0] * color->shape[1];
  l->refs = 1; l->q = (int)color->shape[0]; l->r = (int)color->shape[1];
  l->color = (uint32_t *)malloc(n * 4); l->altitude = (int32_t *)malloc(n * 4);
  int rc = (!l->color || !l->altitude) ? fail(ctx, "out of memory") : 0;
  if (!rc) {
    memcpy(l->color, color->host, n * 4);
    for (size_t i = 0; i < n; ++i) l->altitude[i] = height->host[i] & 0xFF; /* fut/interactive.fut:189 */
    if (fsb_map_new(ctx->fsb, l->color, l->altitude, l->q, l->r, 0, &l->plain)) rc = fail_fsb(ctx);
    else rc = rebake(ctx, l, in->sun_height, in->sun_ang);                  /* :194-196 */
  }
  if (rc) { lsc_release(ctx, l); return rc; }
  struct futhark_opaque_state *s = clone_state(in);
  if (!s) { lsc_release(ctx, l); return fail(ctx, "out of memory"); }
  lsc_release(ctx, s->lsc);
  s->lsc = l;
  *out0 = s;
  return 0;
}

int futhark_entry_render(struct futhark_context *ctx, struct futhark_u32_2d **out0, const struct futhark_opaque_state *s) {
  if (!ctx || !s) return fail(ctx, "render: NULL argument");
  if (!ctx->fsb) return fail(ctx, "no device context");
  const fsb_map *map = s->lsc ? s->lsc->shadowed : NULL;
  if (!map) {
    /* before the first update_map the state holds init's dummy landscape, altitude = color = shadowed_color =
     * [[0],[0]] (fut/interactive.fut:38-43): every sample has height 0 and the empty colour 0, so the frame is the sky
     * colour -- rendered like any other map rather than special-cased */
    if (!ctx->dummy) {
      const uint32_t c[2] = {0u, 0u};
      const int32_t z[2] = {0, 0};
      if (fsb_map_new(ctx->fsb, c, z, 2, 1, 0, &ctx->dummy)) return fail_fsb(ctx);
    }
    map = ctx->dummy;
  }
  struct futhark_u32_2d *a = (struct futhark_u32_2d *)calloc(1, sizeof *a);
  if (!a) return fail(ctx, "out of memory");
  a->shape[0] = s->height; a->shape[1] = s->width;
  void *dev = pool_get(ctx, (size_t)s->height * s->width * 4);
  if (!dev) { free(a); return fail_fsb(ctx); }
  a->dev = (uint32_t *)dev;
  fsb_params prm;
  fsb_params_default(&prm); /* #png: fut/interactive.fut:179-183 */
  if (s->smoothing_on) prm.flags |= FSB_FLAG_SMOOTHING; /* s.smoothing_mode, :182 (key `2`, :153-159) */
  if (fsb_render_device(ctx->fsb, &s->cam, &prm, map, s->height, s->width, a->dev, 0)) {
    futhark_free_u32_2d(ctx, a);
    return fail_fsb(ctx);
  }
  *out0 = a;
  return 0;
}
