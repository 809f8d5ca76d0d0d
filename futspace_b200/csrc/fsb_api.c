/*
 * fsb_api.c -- host side (C, like the reference's c/interactive.c) of the C-ABI declared in
 * include/futspace_b200.h: context, map upload, per-pose constants, launch sequencing, copies.
 *
 * Compiled with -ffp-contract=off: the handful of f32 operations done here (the z-series length
 * of get_zs, fut/voxel_renderer.fut:28-32, and the rotated view vectors of get_h_line, :44-50)
 * must round exactly like the reference's scalar code.
 *
 * There is no CPU fallback: without an sm_100 device fsb_context_new fails.
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fsb_internal.h"

#define FSB_MAX_NZ (1 << 22)
#define FSB_MAX_POSES_PER_LAUNCH 32768

/* Per-launch-group device scratch. */
typedef struct fsb_scratch {
  fsb_frame_consts *fc_dev, *fc_host;
  int fc_cap;
  cudaEvent_t fc_free;            /* pinned pose-constant staging may be rewritten */
  float *table;                   /* per-pose depth tables (blocked by chunk, fsb_kernels.cu) */
  size_t tab_cap;                 /* floats */
  void *recs;                     /* march -> expand record lists */
  uint32_t *sidx;
  size_t recs_cap, sidx_cap;      /* bytes */
  uint32_t *cand, *cand_cnt;      /* column-parallel march: candidate lists and their lengths */
  size_t cand_cap, cand_cnt_cap;  /* bytes */
} fsb_scratch;

struct fsb_context {
  int device;
  cudaStream_t stream, copy_stream;
  fsb_scratch sc;
  cudaEvent_t rendered[2], copied[2];
  char err[512];
  int64_t launches;
  int max_smem_optin, sm_count;
  uint32_t *frame_dev[2];
  size_t frame_cap[2];            /* pixels */
  int profiling;
  cudaEvent_t pev[5];             /* profiling: before set-up, after set-up, after march, after colour, after expand */
  double prof_ms[4];
  int64_t prof_n[4];
  int prof_pending;
  unsigned long long *stats_dev;  /* profiling counters: chunks evaluated, records emitted */
  int force_rec8;                 /* env FSB_REC8: always 8-byte records */
  const float *lut;               /* device address of the colour look-up table */
  int force_march_z;              /* env FSB_MARCH_Z: always the lanes-over-depth march */
  int no_pdl;                     /* env FSB_PDL=0: no programmatic dependent launch for single frames */
  uint64_t paint_trips;           /* colour trips of the paint kernel seen by the last fsb_context_get_counters */
  char name[128];
};

struct fsb_map {
  cudaArray_t array, array_h;     /* RGBA8 packed texels / R16F heights for the texture path */
  cudaTextureObject_t tex, tex_h, tex_f; /* tex_f: the RGBA8 array read as normalised floats (c/255, exact) */
  uint32_t *packed, *color;
  int32_t *height;
  int q, r;
  int cq, cr;                     /* colour plane size (= q, r unless fsb_map_new_split) */
  int pow2, log2r;
  uint32_t alpha_bits;
  int32_t hmax;                   /* highest (masked) terrain height */
  uint8_t *hpyr;                  /* device: pyramid of local height maxima for the local occlusion bound (build_height_pyramid), or NULL */
  int pyr_levels;
};

static int set_err(fsb_context *ctx, int code, const char *fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU(ctx, call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      (void)cudaGetLastError(); /* a reported (non-sticky) error must not resurface at the next launch check */ \
      return set_err((ctx), FSB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    }                                                                                              \
  } while (0)

/* ------------------------------------------------------------------------------------------ */
void fsb_params_default(fsb_params *p) {
  if (!p) return;
  p->z0 = 0.0f;          /* fut/voxel_renderer.fut:103 */
  p->delta = 0.001f;     /* :104 */
  p->invz_param1 = 1.0f; /* :217 */
  p->invz_param2 = 0.0f; /* f32 (w / 2), :217 */
  p->filter = FSB_FILTER_BILINEAR;
  p->sentinel = FSB_SENTINEL_ZERO;
  p->f2i_mode = FSB_F2I_SATURATE;
  p->flags = 0;
}

void fsb_params_tests_variant(fsb_params *p) {
  if (!p) return;
  p->z0 = 1.0f;            /* tests/futspace.fut:84 */
  p->delta = 0.005f;       /* :85 */
  p->invz_param1 = 1.0f;
  p->invz_param2 = 240.0f; /* :91 */
  p->filter = FSB_FILTER_NEAREST;
  p->sentinel = FSB_SENTINEL_SKY;
  p->f2i_mode = FSB_F2I_SATURATE;
  p->flags = 0;
}

/* n of get_zs: floor((sqrt((d - 2 z0)**2 + 8 d c) - 2 z0 + d) / (2 d)), all f32; `**` is powf. */
static int zs_length(float delta, float dist, float z0) {
  float two_z0 = 2.0f * z0;
  float root = sqrtf(powf(delta - two_z0, 2.0f) + (8.0f * delta) * dist);
  float n = floorf(((root - two_z0) + delta) / (2.0f * delta));
  if (!(n >= 0.0f) || n > (float)FSB_MAX_NZ) return -1;
  return (int)n;
}

int fsb_get_zs(float delta, float distance, float z0, float *out, int cap) {
  int n = zs_length(delta, distance, z0);
  if (n < 0) return -1;
  for (int k = 1; k <= n && k <= cap && out; ++k) {
    float i = (float)k;
    out[k - 1] = (i / 2.0f) * (2.0f * z0 + (i - 1.0f) * delta);
  }
  return n;
}

static int make_consts(const fsb_camera *cam, const fsb_params *prm, const fsb_map *map, int w, fsb_frame_consts *fc) {
  float s = sinf(cam->angle), c = cosf(cam->angle), view = cam->fov;
  float sv = s * view, cv = c * view;
  fc->a_lx = -c - sv;
  fc->a_ly = s - cv;
  fc->a_rx = c - sv;
  fc->a_ry = -s - cv;
  fc->cam_x = cam->x;
  fc->cam_y = cam->y;
  fc->cam_h = cam->height;
  fc->horizon = cam->horizon;
  fc->fw = (float)w;
  fc->invz_num = prm->invz_param1;
  fc->invz_mul = prm->invz_param2 > 0.0f ? prm->invz_param2 : (float)(w / 2);
  fc->z0 = prm->z0;
  fc->delta = prm->delta;
  fc->n_z = zs_length(prm->delta, cam->distance, prm->z0);
  fc->sky = cam->sky_color;
  fc->empty = prm->sentinel == FSB_SENTINEL_SKY ? cam->sky_color : 0u;
  /* occlusion bound (fsb_kernels.cu): an interpolated height never exceeds the highest texel by more than the
   * rounding of its seven f32 operations -- 0.5 plus a relative 2e-6 covers any i32 height; rounded upward */
  /* the bilinear sampler returns exactly 0 at integer coordinates (both weights 0), which lies above
   * an all-negative (unmasked) terrain: the bound never drops below 0 */
  const double hm = map->hmax > 0 ? (double)map->hmax : 0.0;
  const double hb = hm + 0.5 + hm * 2.0e-6;
  const float hbound = nextafterf((float)hb, INFINITY);
  /* Only with the saturating conversion (the x86 / modern ones wrap huge rows to 0 and are not monotone) and only
   * while inv_z is positive and decreasing along the series: invz_param1 > 0 and z0 >= 0 (a negative z0 makes the
   * first depths negative, fut/voxel_renderer.fut:33). */
  /* ... and the multiplier too: a one-pixel-wide frame has f32(w/2) = 0, so inv_z is 0 everywhere except NaN
   * (inf * 0) at z = 0 -- that one sample converts to row 0 and fills the column, whatever the bound says. */
  const int monotone = prm->f2i_mode == FSB_F2I_SATURATE && prm->invz_param1 > 0.0f && fc->invz_mul > 0.0f &&
                       prm->z0 >= 0.0f && !(prm->flags & FSB_FLAG_NO_CULL);
  fc->cull_d = monotone ? cam->height - hbound : -INFINITY;
  fc->reserved = 0;
  return fc->n_z < 0 ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------ */
int fsb_context_new(int device, fsb_context **out) {
  if (!out) return FSB_ERR_ARG;
  *out = NULL;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return FSB_ERR_NO_DEVICE;
  if (device < 0 || device >= count) return FSB_ERR_ARG;
  struct cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FSB_ERR_CUDA;
  if (prop.major != 10 || prop.minor != 0) return FSB_ERR_NO_DEVICE; /* the library ships sm_100a SASS only (no PTX): not even sm_103 can run it */
  fsb_context *ctx = (fsb_context *)calloc(1, sizeof *ctx);
  if (!ctx) return FSB_ERR_NOMEM;
  ctx->device = device;
  ctx->force_rec8 = getenv("FSB_REC8") != NULL;
  ctx->force_march_z = getenv("FSB_MARCH_Z") != NULL;
  ctx->no_pdl = getenv("FSB_PDL") != NULL && atoi(getenv("FSB_PDL")) == 0;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  ctx->sm_count = prop.multiProcessorCount;
  snprintf(ctx->name, sizeof ctx->name, "%.127s", prop.name);
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    free(ctx);
    return FSB_ERR_CUDA;
  }
  /* the march's colour table lives in the module's global memory of this device; (re)filling it with the same
   * values is harmless, and everything this context launches is ordered after it on ctx->stream */
  if (fsb_launch_lut_init(ctx->stream) != 0 || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    free(ctx);
    return FSB_ERR_CUDA;
  }
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&ctx->rendered[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->copied[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&ctx->sc.fc_free, cudaEventDisableTiming);
  ctx->lut = fsb_lut_device_address();
  *out = ctx;
  return FSB_OK;
}

void fsb_context_free(fsb_context *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  {
    fsb_scratch *sc = &ctx->sc;
    cudaFree(sc->fc_dev);
    cudaFreeHost(sc->fc_host);
    cudaFree(sc->table);
    cudaFree(sc->recs);
    cudaFree(sc->sidx);
    cudaFree(sc->cand);
    cudaFree(sc->cand_cnt);
    cudaEventDestroy(sc->fc_free);
  }
  cudaFree(ctx->frame_dev[0]);
  cudaFree(ctx->frame_dev[1]);
  cudaFree(ctx->stats_dev);
  for (int i = 0; i < 5; ++i)
    if (ctx->pev[i]) cudaEventDestroy(ctx->pev[i]);
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(ctx->rendered[i]);
    cudaEventDestroy(ctx->copied[i]);
  }
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  free(ctx);
}

const char *fsb_context_get_error(fsb_context *ctx) { return ctx ? ctx->err : "no context"; }

int fsb_context_sync(fsb_context *ctx) {
  if (!ctx) return FSB_ERR_ARG;
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
  return FSB_OK;
}

void *fsb_context_stream(fsb_context *ctx) { return ctx ? (void *)ctx->stream : NULL; }
int64_t fsb_context_launch_count(const fsb_context *ctx) { return ctx ? ctx->launches : 0; }
int fsb_context_device_name(fsb_context *ctx, char *buf, size_t n) {
  if (!ctx || !buf || !n) return FSB_ERR_ARG;
  snprintf(buf, n, "%s", ctx->name);
  return FSB_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* IEEE binary16 bits of an integer 0..255 (exact) */
static uint16_t half_bits_of_byte(uint32_t v) {
  if (v == 0) return 0;
  int e = 0;
  while ((v >> e) > 1) ++e;                       /* v = 1.m x 2^e, e <= 7 */
  return (uint16_t)(((e + 15) << 10) | (((v << (10 - e)) & 0x3FF)));
}

/* Pyramid of local height maxima for the march's local occlusion bound (fsb_march_cols.cu, chunk_hidden): level L
 * (1..levels) holds, for every 2^L x 2^L block of texels, the highest texel of the 2 x 2 blocks starting at it (wrapped
 * like the samplers wrap) -- any rectangle of at most 2^L texels a side lies inside one such 2 x 2 neighbourhood, so ONE
 * byte bounds every height a group of samples can read.  Power-of-two maps with heights 0..255; level L at byte offset
 * (q r - (q r >> (2L - 2))) / 3, row-major (q >> L) x (r >> L). */
static uint8_t *build_height_pyramid(const int32_t *hm, int q, int r, int *levels_out, size_t *bytes_out) {
  int levels = 0;
  while ((2 << levels) <= q && (2 << levels) <= r) ++levels;
  if (levels < 1) return NULL;
  const size_t n = (size_t)q * r;
  uint8_t *prev = (uint8_t *)malloc(n), *out = (uint8_t *)malloc(n / 3 + 16);
  if (!prev || !out) {
    free(prev); free(out);
    return NULL;
  }
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) prev[i] = (uint8_t)hm[i];
  int pq = q, pr = r;
  for (int L = 1; L <= levels; ++L) {
    const int cq = pq >> 1, cr = pr >> 1;
    uint8_t *cur = (uint8_t *)malloc((size_t)cq * cr);
    if (!cur) {
      free(prev); free(out);
      return NULL;
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < cq; ++y)
      for (int x = 0; x < cr; ++x) {
        const uint8_t *a = prev + (size_t)(2 * y) * pr + 2 * x, *b = a + pr;
        uint8_t v = a[0] > a[1] ? a[0] : a[1], w = b[0] > b[1] ? b[0] : b[1];
        cur[(size_t)y * cr + x] = v > w ? v : w;
      }
    uint8_t *dst = out + (n - (n >> (2 * L - 2))) / 3;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < cq; ++y) {
      const uint8_t *a = cur + (size_t)y * cr, *b = cur + (size_t)((y + 1) & (cq - 1)) * cr;
      for (int x = 0; x < cr; ++x) {
        const int x1 = (x + 1) & (cr - 1);
        uint8_t v = a[x] > a[x1] ? a[x] : a[x1], w = b[x] > b[x1] ? b[x] : b[x1];
        dst[(size_t)y * cr + x] = v > w ? v : w;
      }
    }
    free(prev);
    prev = cur;
    pq = cq;
    pr = cr;
  }
  free(prev);
  *levels_out = levels;
  *bytes_out = (n - (n >> (2 * levels))) / 3;
  return out;
}

/* test hook (tests/test_local_bound_cpu.py; declared in fsb_internal.h only): the pyramid as uploaded */
int fsb_debug_height_pyramid(const int32_t *hm, int q, int r, uint8_t *out, size_t cap) {
  int levels = 0;
  size_t bytes = 0;
  if (!hm || q <= 0 || r <= 0 || (q & (q - 1)) || (r & (r - 1))) return -1;
  uint8_t *p = build_height_pyramid(hm, q, r, &levels, &bytes);
  if (!p) return 0;
  if (out && bytes <= cap) memcpy(out, p, bytes);
  free(p);
  return levels;
}

int fsb_map_new(fsb_context *ctx, const uint32_t *color, const int32_t *height, int q, int r, int mask_heights,
                fsb_map **out) {
  if (!ctx) return FSB_ERR_ARG;
  if (!color || !height || !out) return set_err(ctx, FSB_ERR_ARG, "fsb_map_new: NULL argument");
  *out = NULL;
  if (q <= 0 || r <= 0 || (int64_t)q * r >= (1ll << 31))
    return set_err(ctx, FSB_ERR_ARG, "fsb_map_new: bad map size %d x %d", q, r);
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)q * r;
  fsb_map *m = (fsb_map *)calloc(1, sizeof *m);
  int32_t *hm = (int32_t *)malloc(n * 4);
  uint32_t *pk = (uint32_t *)malloc(n * 4);
  if (!m || !hm || !pk) {
    free(m); free(hm); free(pk);
    return set_err(ctx, FSB_ERR_NOMEM, "fsb_map_new: out of host memory");
  }
  m->q = q;
  m->r = r;
  m->cq = q;
  m->cr = r;
  m->pow2 = ((q & (q - 1)) == 0) && ((r & (r - 1)) == 0);
  while ((1 << m->log2r) < r) ++m->log2r;
  /* update_map, fut/interactive.fut:189: altitude = height & 0xFF */
  /* packed texel = height byte + rgb + map-uniform alpha; the tiled __ldg layout also needs power-of-two
   * sizes holding at least one 8x4 tile (fsb_kernels.cu texel_x/texel_y) */
  int packable = 1;
  const int tileable = m->pow2 && r >= 8 && q >= 4;
  uint32_t *pk_rm = (uint32_t *)malloc(n * 4); /* row-major packed texels for the texture */
  if (!pk_rm) {
    free(m); free(hm); free(pk);
    return set_err(ctx, FSB_ERR_NOMEM, "fsb_map_new: out of host memory");
  }
  const uint32_t alpha = color[0] & 0xFF000000u;
  int32_t hmax = mask_heights ? (height[0] & 0xFF) : height[0];
  /* rows in parallel: a 16384 x 16384 map is 268 M texels */
#pragma omp parallel for schedule(static) reduction(max : hmax) reduction(&& : packable)
  for (int y = 0; y < q; ++y) {
    for (int x = 0; x < r; ++x) {
      const size_t i = (size_t)y * (size_t)r + (size_t)x;
      const int32_t hv = mask_heights ? (height[i] & 0xFF) : height[i];
      hm[i] = hv;
      if (hv > hmax) hmax = hv;
      if (hv < 0 || hv > 255 || (color[i] & 0xFF000000u) != alpha) packable = 0;
      pk_rm[i] = ((uint32_t)hv << 24) | (color[i] & 0x00FFFFFFu);
      if (tileable) {
        const size_t t = (((size_t)y >> 2) * (size_t)(r >> 3) + ((size_t)x >> 3)) * 32 + (((size_t)y & 3) << 3) + ((size_t)x & 7);
        pk[t] = pk_rm[i];
      }
    }
  }
  m->hmax = hmax;
  m->alpha_bits = alpha;
  cudaError_t e = cudaMalloc((void **)&m->color, n * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->height, n * 4);
  if (e == cudaSuccess && packable && tileable) e = cudaMalloc((void **)&m->packed, n * 4);
  struct cudaDeviceProp prop;
  int gather_ok = cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && r <= prop.maxTexture2DGather[0] &&
                  q <= prop.maxTexture2DGather[1];
  cudaError_t e_planes = e;
  if (e == cudaSuccess && packable && gather_ok) {
    struct cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
    e = cudaMallocArray(&m->array, &cd, (size_t)r, (size_t)q, cudaArrayTextureGather);
    if (e == cudaSuccess)
      e = cudaMemcpy2DToArrayAsync(m->array, 0, 0, pk_rm, (size_t)r * 4, (size_t)r * 4, (size_t)q, cudaMemcpyHostToDevice,
                                   ctx->stream);
    if (e == cudaSuccess) {
      struct cudaResourceDesc rd;
      struct cudaTextureDesc td;
      memset(&rd, 0, sizeof rd);
      memset(&td, 0, sizeof td);
      rd.resType = cudaResourceTypeArray;
      rd.res.array.array = m->array;
      td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap; /* floored modulo of the samplers */
      td.filterMode = cudaFilterModePoint;
      td.readMode = cudaReadModeElementType;
      td.normalizedCoords = 1;
      e = cudaCreateTextureObject(&m->tex, &rd, &td, NULL);
      if (e == cudaSuccess) {
        td.readMode = cudaReadModeNormalizedFloat;
        e = cudaCreateTextureObject(&m->tex_f, &rd, &td, NULL);
        td.readMode = cudaReadModeElementType;
      }
      /* heights alone as IEEE half (0..255 are exact): the march gathers 2-byte texels and gets floats */
      uint16_t *hh = (uint16_t *)malloc(n * 2);
      if (!hh) e = cudaErrorMemoryAllocation;
      if (e == cudaSuccess) {
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)n; ++i) hh[i] = half_bits_of_byte((uint32_t)hm[i]);
        struct cudaChannelFormatDesc ch = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindFloat);
        e = cudaMallocArray(&m->array_h, &ch, (size_t)r, (size_t)q, cudaArrayTextureGather);
        if (e == cudaSuccess)
          e = cudaMemcpy2DToArray(m->array_h, 0, 0, hh, (size_t)r * 2, (size_t)r * 2, (size_t)q, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
          rd.res.array.array = m->array_h;
          e = cudaCreateTextureObject(&m->tex_h, &rd, &td, NULL);
        }
      }
      free(hh);
    }
    if (e != cudaSuccess) { /* no texture path for this map (size limits, array memory): the tiled / two-plane kernels need none */
      (void)cudaGetLastError();
      cudaStreamSynchronize(ctx->stream);
      if (m->tex) cudaDestroyTextureObject(m->tex);
      if (m->tex_h) cudaDestroyTextureObject(m->tex_h);
      if (m->tex_f) cudaDestroyTextureObject(m->tex_f);
      if (m->array) cudaFreeArray(m->array);
      if (m->array_h) cudaFreeArray(m->array_h);
      m->tex = m->tex_h = m->tex_f = 0;
      m->array = m->array_h = NULL;
      (void)cudaGetLastError();
      e = e_planes;
    }
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->color, color, n * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->height, hm, n * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && m->packed) e = cudaMemcpyAsync(m->packed, pk, n * 4, cudaMemcpyHostToDevice, ctx->stream);
  uint8_t *pyr_host = NULL;
  if (e == cudaSuccess && m->tex_h && m->pow2 && n <= ((size_t)1 << 30)) {
    /* optional: without it the march only has the map-wide bound */
    size_t pyr_bytes = 0;
    pyr_host = build_height_pyramid(hm, q, r, &m->pyr_levels, &pyr_bytes);
    if (pyr_host && cudaMalloc((void **)&m->hpyr, pyr_bytes) == cudaSuccess)
      e = cudaMemcpyAsync(m->hpyr, pyr_host, pyr_bytes, cudaMemcpyHostToDevice, ctx->stream);
    else
      (void)cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  free(pyr_host);
  free(hm);
  free(pk);
  free(pk_rm);
  if (e != cudaSuccess) {
    if (m->tex) cudaDestroyTextureObject(m->tex);
    if (m->tex_h) cudaDestroyTextureObject(m->tex_h);
    if (m->tex_f) cudaDestroyTextureObject(m->tex_f);
    if (m->array) cudaFreeArray(m->array);
    if (m->array_h) cudaFreeArray(m->array_h);
    cudaFree(m->color); cudaFree(m->height); cudaFree(m->packed); cudaFree(m->hpyr);
    free(m);
    (void)cudaGetLastError();
    return set_err(ctx, FSB_ERR_CUDA, "fsb_map_new: %s", cudaGetErrorString(e));
  }
  *out = m;
  return FSB_OK;
}

/* Colour and height maps of different sizes, each wrapped by its own size: what `render` sees after the reference's
 * update_map on a map that is not 1024 x 1024 (lsc.shadowed_color is baked at 1024 x 1024 whatever the input,
 * fut/effects.fut:124-125, while lsc.altitude keeps the map's size, fut/interactive.fut:188-198).  Two-plane generic
 * kernel only. */
int fsb_map_new_split(fsb_context *ctx, const uint32_t *color, int cq, int cr, const int32_t *height, int q, int r,
                      int mask_heights, fsb_map **out) {
  if (!ctx) return FSB_ERR_ARG;
  if (cq == q && cr == r) return fsb_map_new(ctx, color, height, q, r, mask_heights, out);
  if (!color || !height || !out) return set_err(ctx, FSB_ERR_ARG, "fsb_map_new_split: NULL argument");
  *out = NULL;
  if (q <= 0 || r <= 0 || cq <= 0 || cr <= 0 || (int64_t)q * r >= (1ll << 31) || (int64_t)cq * cr >= (1ll << 31))
    return set_err(ctx, FSB_ERR_ARG, "fsb_map_new_split: bad map size %d x %d / %d x %d", cq, cr, q, r);
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)q * r, nc = (size_t)cq * cr;
  fsb_map *m = (fsb_map *)calloc(1, sizeof *m);
  int32_t *hm = (int32_t *)malloc(n * 4);
  if (!m || !hm) {
    free(m); free(hm);
    return set_err(ctx, FSB_ERR_NOMEM, "fsb_map_new_split: out of host memory");
  }
  m->q = q; m->r = r; m->cq = cq; m->cr = cr;
  int32_t hmax = mask_heights ? (height[0] & 0xFF) : height[0];
  for (size_t i = 0; i < n; ++i) {
    hm[i] = mask_heights ? (height[i] & 0xFF) : height[i];
    if (hm[i] > hmax) hmax = hm[i];
  }
  m->hmax = hmax;
  m->alpha_bits = color[0] & 0xFF000000u;
  cudaError_t e = cudaMalloc((void **)&m->color, nc * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->height, n * 4);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->color, color, nc * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->height, hm, n * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  free(hm);
  if (e != cudaSuccess) {
    cudaFree(m->color); cudaFree(m->height);
    free(m);
    (void)cudaGetLastError();
    return set_err(ctx, FSB_ERR_CUDA, "fsb_map_new_split: %s", cudaGetErrorString(e));
  }
  *out = m;
  return FSB_OK;
}

int fsb_map_free(fsb_context *ctx, fsb_map *m) {
  if (!ctx) return FSB_ERR_ARG;
  if (!m) return FSB_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (m->tex) cudaDestroyTextureObject(m->tex);
  if (m->tex_h) cudaDestroyTextureObject(m->tex_h);
  if (m->tex_f) cudaDestroyTextureObject(m->tex_f);
  if (m->array) cudaFreeArray(m->array);
  if (m->array_h) cudaFreeArray(m->array_h);
  cudaFree(m->color);
  cudaFree(m->height);
  cudaFree(m->packed);
  cudaFree(m->hpyr);
  free(m);
  return FSB_OK;
}

/* update_map's second half, fut/interactive.fut:194-196 -> fut/effects.fut:108-125 */
int fsb_map_bake_shadows(fsb_context *ctx, const fsb_map *m, const float sun[3], int out_q, int out_r,
                         uint32_t *out_host) {
  if (!ctx) return FSB_ERR_ARG;
  if (!m || !sun || !out_host || out_q <= 0 || out_r <= 0) return set_err(ctx, FSB_ERR_ARG, "fsb_map_bake_shadows: bad argument");
  if (m->cq != m->q || m->cr != m->r)
    return set_err(ctx, FSB_ERR_ARG, "fsb_map_bake_shadows: the map's colour and height planes differ in size");
  CU(ctx, cudaSetDevice(ctx->device));
  uint32_t *d = NULL;
  const size_t bytes = (size_t)out_q * out_r * 4;
  CU(ctx, cudaMalloc((void **)&d, bytes));
  cudaError_t e = (cudaError_t)fsb_launch_shadow(m->color, m->height, m->q, m->r, sun, out_q, out_r, d, ctx->stream,
                                                 &ctx->launches);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return set_err(ctx, FSB_ERR_CUDA, "fsb_map_bake_shadows: %s", cudaGetErrorString(e));
  return FSB_OK;
}

/* vec3_rotate #y sun_ang (vec3_rotate #z sun_height [0,1,0]), fut/effects.fut:6-25, fut/interactive.fut:56,196 */
void fsb_sun_vector(float sun_height, float sun_ang, float out[3]) {
  const float cz = cosf(sun_height), sz = sinf(sun_height), cy = cosf(sun_ang), sy = sinf(sun_ang);
  /* R_z * [0,1,0]: rows summed left to right (linalg matvecmul_row) */
  const float v0 = ((0.0f + cz * 0.0f) + (-sz) * 1.0f) + 0.0f * 0.0f;
  const float v1 = ((0.0f + sz * 0.0f) + cz * 1.0f) + 0.0f * 0.0f;
  const float v2 = ((0.0f + 0.0f * 0.0f) + 0.0f * 1.0f) + 1.0f * 0.0f;
  out[0] = ((0.0f + cy * v0) + 0.0f * v1) + sy * v2;
  out[1] = ((0.0f + 0.0f * v0) + 1.0f * v1) + 0.0f * v2;
  out[2] = ((0.0f + (-sy) * v0) + 0.0f * v1) + cy * v2;
}

int fsb_map_is_packed(const fsb_map *m) { return m && (m->packed != NULL || m->tex != 0); }

/* ------------------------------------------------------------------------------------------ */
static int ensure_tables(fsb_context *ctx, fsb_scratch *sc, int n_poses, int tab_stride) {
  if (n_poses > sc->fc_cap) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(sc->fc_dev);
    cudaFreeHost(sc->fc_host);
    sc->fc_dev = NULL; sc->fc_host = NULL; sc->fc_cap = 0;
    CU(ctx, cudaMalloc((void **)&sc->fc_dev, sizeof(fsb_frame_consts) * (size_t)n_poses));
    CU(ctx, cudaMallocHost((void **)&sc->fc_host, sizeof(fsb_frame_consts) * (size_t)n_poses));
    sc->fc_cap = n_poses;
  }
  const size_t need = (size_t)n_poses * (size_t)tab_stride;
  if (need > sc->tab_cap) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(sc->table);
    sc->table = NULL; sc->tab_cap = 0;
    CU(ctx, cudaMalloc((void **)&sc->table, need * 4));
    sc->tab_cap = need;
  }
  return FSB_OK;
}

#define FSB_RB_SHIFT 5                   /* band of the per-column record index = 32 rows (fsb_expand_kernel) */
#define FSB_MAX_H 32768
#define FSB_SCRATCH_BUDGET ((size_t)8192 << 20)
#define FSB_CAND_PAD 1024                /* words in front of the candidate lists (the paint kernel's prefetch reaches below a list's start) */
#define FSB_COLS_MAX_NZ (1 << 17)        /* candidate word of the column-parallel march: row (15 bits) | sample index (17 bits) */

static int grow(fsb_context *ctx, void **ptr, size_t *cap, size_t need) {
  if (need <= *cap) return FSB_OK;
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(*ptr);
  *ptr = NULL;
  *cap = 0;
  CU(ctx, cudaMalloc(ptr, need));
  *cap = need;
  return FSB_OK;
}

/* How one launch group is rendered: which march, which record format, how the depth series is split. */
typedef struct {
  int mem;        /* FSB_MEM_* */
  int cols;       /* column-parallel lists: fsb_march_cols.cu, or fsb_march_split.cu when `split` */
  int split;      /* depth-parallel cluster march (fsb_march_split.cu): warps per group of 32 columns (32 / 64), 0: off */
  int frame;      /* one CTA per column (fsb_march_frame.cu): single frames, small batches; = warps per column, 0: off */
  int rec4;       /* 4-byte records */
  int cand_cap;   /* candidate words per column */
  int slice_len;  /* colour pass: records per warp (0: whole lists) */
  int paint;      /* colour pass and expand as one kernel (fsb_paint.cu) */
  int seg_bands;  /* paint: bands of 32 rows per warp (0: whole columns) */
  int ncols_pad;
} render_plan;

static int ensure_scratch(fsb_context *ctx, fsb_scratch *sc, int n_poses, int ncols, int h, const render_plan *pl) {
  const int n_bands = (h + (1 << FSB_RB_SHIFT) - 1) >> FSB_RB_SHIFT;
  const size_t np = (size_t)n_poses;
  int rc;
  const size_t lc = pl->cols ? (size_t)pl->ncols_pad : (size_t)ncols; /* lists: interleaved by groups of 32 columns, or one per column */
  if (!pl->paint) { /* the paint kernel keeps the records in shared memory and needs no band index */
    if ((rc = grow(ctx, &sc->recs, &sc->recs_cap, np * lc * (h + 1) * (pl->rec4 ? 4 : 8)))) return rc;
    if ((rc = grow(ctx, (void **)&sc->sidx, &sc->sidx_cap, np * lc * (n_bands + 1) * 4))) return rc;
  }
  if (pl->cols) {
    const size_t lists = np * pl->ncols_pad;
    /* + a pad in front: the paint kernel prefetches a few entries past the start of a list (fsb_paint.cu) */
    if ((rc = grow(ctx, (void **)&sc->cand, &sc->cand_cap, lists * pl->cand_cap * 4 + FSB_CAND_PAD * 4))) return rc;
    if ((rc = grow(ctx, (void **)&sc->cand_cnt, &sc->cand_cnt_cap, lists * 4))) return rc;
  }
  return FSB_OK;
}

/* poses per launch group: bounded by the record-list scratch budget (8 bytes per row and column: 4-byte candidates +
 * 4-byte records, or 8-byte records alone) */
static int group_size(int n, int ncols, int h) {
  size_t per = (size_t)ncols * (h + 1) * 8;
  size_t g = FSB_SCRATCH_BUDGET / (per ? per : 1);
  const char *env = getenv("FSB_GROUP_POSES"); /* tuning aid: poses per launch group */
  if (env && atoi(env) > 0) g = (size_t)atoi(env);
  if (g < 1) g = 1;
  if (g > FSB_MAX_POSES_PER_LAUNCH) g = FSB_MAX_POSES_PER_LAUNCH;
  return (size_t)n < g ? n : (int)g;
}

static int ensure_frame(fsb_context *ctx, int slot, size_t pixels);

static int check_common(fsb_context *ctx, const fsb_camera *cams, int n, const fsb_params *prm, const fsb_map *map,
                        int h, int w, int col_begin, int col_end, const void *out) {
  if (!ctx) return FSB_ERR_ARG;
  if (!cams || !prm || !map || !out) return set_err(ctx, FSB_ERR_ARG, "render: NULL argument");
  if (n <= 0 || h <= 0 || w <= 0 || col_begin < 0 || col_end > w || col_begin >= col_end)
    return set_err(ctx, FSB_ERR_ARG, "render: bad sizes n=%d h=%d w=%d cols=[%d,%d)", n, h, w, col_begin, col_end);
  if ((unsigned)prm->filter > 1u || (unsigned)prm->sentinel > 1u || (unsigned)prm->f2i_mode > 2u)
    return set_err(ctx, FSB_ERR_ARG, "render: unknown filter/sentinel/f2i_mode");
  if (h > FSB_MAX_H) return set_err(ctx, FSB_ERR_RANGE, "render: frame height %d exceeds the limit %d", h, FSB_MAX_H);
  return FSB_OK;
}

/* profiling: fold the event deltas of the previous launch group into the accumulators */
static int prof_collect(fsb_context *ctx) {
  if (!ctx->prof_pending) return FSB_OK;
  CU(ctx, cudaEventSynchronize(ctx->pev[4]));
  for (int i = 0; i < 4; ++i) {
    float ms = 0.f;
    CU(ctx, cudaEventElapsedTime(&ms, ctx->pev[i], ctx->pev[i + 1]));
    ctx->prof_ms[i] += ms;
    ctx->prof_n[i] += 1;
  }
  ctx->prof_pending = 0;
  return FSB_OK;
}

int fsb_context_set_profiling(fsb_context *ctx, int enable) {
  if (!ctx) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  if (enable && !ctx->pev[0])
    for (int i = 0; i < 5; ++i) CU(ctx, cudaEventCreate(&ctx->pev[i]));
  if (enable && !ctx->stats_dev) {
    CU(ctx, cudaMalloc((void **)&ctx->stats_dev, 32));
    CU(ctx, cudaMemsetAsync(ctx->stats_dev, 0, 32, ctx->stream));
  }
  int rc = prof_collect(ctx);
  ctx->profiling = enable != 0;
  return rc;
}

int fsb_context_get_counters(fsb_context *ctx, uint64_t *chunks_evaluated, uint64_t *records) {
  if (!ctx || !chunks_evaluated || !records) return FSB_ERR_ARG;
  if (!ctx->stats_dev) return set_err(ctx, FSB_ERR_ARG, "fsb_context_get_counters: profiling was never enabled");
  unsigned long long h[4] = {0, 0, 0, 0};
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemcpyAsync(h, ctx->stats_dev, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaMemsetAsync(ctx->stats_dev, 0, 32, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  *chunks_evaluated = h[0];
  *records = h[1];
  ctx->paint_trips = h[2];
  return FSB_OK;
}

uint64_t fsb_context_paint_trips(const fsb_context *ctx) { return ctx ? ctx->paint_trips : 0; }

int fsb_context_get_profile(fsb_context *ctx, double *ms, int64_t *launches) {
  if (!ctx || !ms || !launches) return FSB_ERR_ARG;
  int rc = prof_collect(ctx);
  if (rc) return rc;
  for (int i = 0; i < 4; ++i) {
    ms[i] = ctx->prof_ms[i];
    launches[i] = ctx->prof_n[i];
    ctx->prof_ms[i] = 0.0;
    ctx->prof_n[i] = 0;
  }
  return FSB_OK;
}

/* Queue set-up + render kernels for n poses (n <= FSB_MAX_POSES_PER_LAUNCH) on ctx->stream. */
static int render_poses(fsb_context *ctx, const fsb_camera *cams, int n, const fsb_params *prm, const fsb_map *map,
                        int h, int w, int col_begin, int col_end, uint32_t *out_dev, int64_t row_stride,
                        int64_t pose_stride) {
  fsb_scratch *sc = &ctx->sc;
  fsb_frame_consts single;
  int max_nz = 0;
  if (row_stride > (int64_t)(INT32_MAX / 4)) /* fsb_expand_kernel forms row offsets from a 32-bit byte stride */
    return set_err(ctx, FSB_ERR_ARG, "render: row_stride %lld too large", (long long)row_stride);
  if (n == 1) {
    if (make_consts(&cams[0], prm, map, w, &single))
      return set_err(ctx, FSB_ERR_RANGE, "render: z-series undefined for distance=%g delta=%g z0=%g",
                     (double)cams[0].distance, (double)prm->delta, (double)prm->z0);
    max_nz = single.n_z;
  }
  int rc;
  if (n > 1) {
    if ((rc = ensure_tables(ctx, sc, n, 160 * 6))) return rc;
    CU(ctx, cudaEventSynchronize(sc->fc_free));
    for (int i = 0; i < n; ++i) {
      if (make_consts(&cams[i], prm, map, w, &sc->fc_host[i]))
        return set_err(ctx, FSB_ERR_RANGE, "render: z-series undefined for pose %d (distance=%g)", i,
                       (double)cams[i].distance);
      if (sc->fc_host[i].n_z > max_nz) max_nz = sc->fc_host[i].n_z;
    }
  }
  if (prm->flags & FSB_FLAG_SMOOTHING) {
    if (prm->sentinel != FSB_SENTINEL_ZERO)
      return set_err(ctx, FSB_ERR_ARG, "render: smoothing exists only with the zero sentinel (fut/voxel_renderer.fut:175-213)");
    if (max_nz > (1 << 17)) /* the record word holds row (15 bits) | sample index (17 bits) */
      return set_err(ctx, FSB_ERR_RANGE, "render: smoothing supports at most %d depth samples, got %d", 1 << 17, max_nz);
  }
  /* depth table: one 160-float block per chunk of 32 samples, plus the blocks the march loop prefetches
   * past the last chunk (fsb_kernels.cu) */
  const int tab_stride = 160 * ((max_nz + 31) / 32 + 6);
  if ((rc = ensure_tables(ctx, sc, n, tab_stride))) return rc;
  /* ---- plan: which march, which record format, how the depth series is split ---- */
  render_plan pl;
  memset(&pl, 0, sizeof pl);
  pl.mem = FSB_MEM_PLANES;
  const int ncols = col_end - col_begin;
  const int smooth = (prm->flags & FSB_FLAG_SMOOTHING) ? 1 : 0;
  /* The texture path addresses texels with normalised coordinates (floor(x) + 1) / size and relies on that point
   * staying within half a texel of the footprint centre.  For power-of-two sizes the division is exact; otherwise
   * the two roundings cost up to 2^-23 * |coordinate| texels, so the coordinate range is kept where that is < 1/8.
   * Beyond the range the generic kernel (integer addressing) renders the frame. */
  const double limit = map->pow2 ? 4.0e6 : 1.0e6;
  int in_range = 1;
  for (int i = 0; i < n; ++i) {
    const double reach = (double)fabsf(cams[i].distance) * (1.0 + (double)fabsf(cams[i].fov)) * 1.01 + 4.0;
    if (!((double)fabsf(cams[i].x) + reach < limit && (double)fabsf(cams[i].y) + reach < limit)) in_range = 0;
  }
  if (in_range && prm->f2i_mode == FSB_F2I_SATURATE && !(prm->flags & FSB_FLAG_FORCE_GENERIC)) {
    if (map->tex && map->tex_h && map->tex_f && !(prm->flags & FSB_FLAG_NO_TEXTURE)) pl.mem = FSB_MEM_TEX;
    else if (map->packed) pl.mem = FSB_MEM_TILED;
  }
  /* 4-byte records where every emitted colour has alpha 0xFF or is 0 (fsb_expand4_kernel); FSB_REC8=1 keeps the
   * 8-byte format for A/B measurements */
  pl.rec4 = pl.mem != FSB_MEM_PLANES && !smooth && (map->alpha_bits == 0u || map->alpha_bits == 0xFF000000u) &&
            !ctx->force_rec8;
  /* The column-parallel march serves the texture path (the default for packable maps); the lanes-over-depth march of
   * round 1 keeps the tiled / generic paths, very long series, and FSB_FLAG_MARCH_Z (A/B). */
  pl.cols = pl.mem == FSB_MEM_TEX && max_nz <= FSB_COLS_MAX_NZ && !(prm->flags & FSB_FLAG_MARCH_Z) && !ctx->force_march_z;
  pl.ncols_pad = (ncols + 31) & ~31;
  if (pl.cols) {
    /* one march warp per 32 columns: a launch group must offer enough of them to fill the device */
    const long long warps = (long long)(pl.ncols_pad / 32) * n;
    /* measured crossover with the paint kernel behind the march (round 2, GPU session 33, profiles/r2_paint_variants.txt):
     * 32 poses at 1920 columns (13 warps per SM), 16-32 at 3840 (13-26 per SM) */
    long long min_warps = (long long)ctx->sm_count * 16;
    const char *env = getenv("FSB_COLS_MIN_WARPS"); /* tuning aid */
    if (env && atoi(env) >= 0) min_warps = atoi(env);
    if (warps < min_warps) pl.cols = 0;
  }
  /* FSB_SPLIT=1 (A/B, tests): measured slower than the marches below as three launches (profiles/r2_split_march_v1.jsonl:
   * 37.9 against 29.5 us at 1080p -- the colour pass as a launch of its own costs what the split saves) */
  const char *env_split = getenv("FSB_SPLIT");
  if (!pl.cols && pl.mem == FSB_MEM_TEX && max_nz <= FSB_COLS_MAX_NZ && !(prm->flags & FSB_FLAG_MARCH_Z) && !ctx->force_march_z &&
      env_split && atoi(env_split) != 0) {
    /* Too few groups of 32 columns to fill the device with one warp each (single frames, small batches): the depth series
     * of every group is split over the 32 or 64 warps of a thread-block cluster (fsb_march_split.cu), 64 while the whole
     * launch still fits the device at 32 warps per SM. */
    const long long groups = (long long)(pl.ncols_pad / 32) * n;
    const int n_chunks = (max_nz + 31) / 32;
    int wpg = groups * 64 <= (long long)ctx->sm_count * 32 ? 64 : 32;
    const char *env = getenv("FSB_SPLIT_WARPS"); /* tuning aid */
    if (env && (atoi(env) == 32 || atoi(env) == 64)) wpg = atoi(env);
    long long max_groups = (long long)ctx->sm_count * 4;
    env = getenv("FSB_SPLIT_MAX_GROUPS"); /* tuning aid */
    if (env && atoi(env) >= 0) max_groups = atoi(env);
    if (n_chunks > fsb_march_split_max_chunks(wpg)) wpg = 64;
    if (n_chunks <= fsb_march_split_max_chunks(wpg) && groups <= max_groups) {
      pl.split = wpg;
      pl.cols = 1;
    }
  }
  if (!pl.cols && pl.mem == FSB_MEM_TEX && !(prm->flags & FSB_FLAG_MARCH_Z) && !ctx->force_march_z) {
    /* too few columns to fill the device with one warp each: four warps per column.  Measured (round 2,
     * profiles/r2_single_frame_march_variants.jsonl): 37 -> 31 us for a lone 1920-column frame, a loss from 3840 columns on */
    long long max_cols = (long long)ctx->sm_count * 16;
    const char *env = getenv("FSB_FRAME_MAX_COLS"); /* tuning aid */
    if (env && atoi(env) >= 0) max_cols = atoi(env);
    pl.frame = (long long)ncols * n < max_cols ? 4 : 0;
    env = getenv("FSB_FRAME_WARPS"); /* tuning aid: warps per column (2, 3, 4) */
    if (pl.frame && env && atoi(env) >= 2 && atoi(env) <= 8) pl.frame = atoi(env);
  }
  if (pl.cols) {
    pl.cand_cap = max_nz < h ? max_nz : h; /* one candidate per depth sample at most, and rows strictly decrease */
    if (pl.cand_cap < 1) pl.cand_cap = 1;
    /* colour pass: whole lists per warp when there are plenty of them, otherwise slices of 32 records */
    const long long lists = (long long)(pl.ncols_pad / 32) * n;
    pl.slice_len = lists >= (long long)ctx->sm_count * 160 ? 0 : lists >= (long long)ctx->sm_count * 80 ? 64 : 32;
    if (pl.split) pl.slice_len = 8; /* a lone frame has 60 lists of 32 columns: short slices spread the filter over the device */
    const char *env = getenv("FSB_COLOUR_SLICE");
    if (env && atoi(env) >= 0) pl.slice_len = atoi(env);
    /* 4-byte record case: colour pass and expand run as one kernel, the records stay in shared memory (fsb_paint.cu);
     * FSB_PAINT=0 (A/B) keeps the two launches */
    env = getenv("FSB_PAINT");
    pl.paint = pl.rec4 && !pl.split && !(env && atoi(env) == 0);
    if (pl.paint) {
      const int n_bands = (h + (1 << FSB_RB_SHIFT) - 1) >> FSB_RB_SHIFT;
      /* whole columns per warp when there are plenty of lists, otherwise segments of bands: about 270 paint warps per SM
       * (measured, profiles/r2_paint_variants.txt session 25: 3 segments at 4K x 128 poses, 5-7 at 1080p x 128, 1 at 1080p x 512) */
      const long long segs = ((long long)ctx->sm_count * 270 + lists / 2) / (lists > 0 ? lists : 1);
      pl.seg_bands = segs <= 1 ? 0 : (int)((n_bands + segs - 1) / segs);
      if (segs > 1 && pl.seg_bands < 2) pl.seg_bands = 2; /* (a tiny batch: the lanes-over-depth march renders it anyway) */
      env = getenv("FSB_PAINT_SEG"); /* tuning aid: bands per paint warp */
      if (env && atoi(env) >= 0) pl.seg_bands = atoi(env);
    }
  }
  if ((rc = ensure_scratch(ctx, sc, n, ncols, h, &pl))) return rc;
  if (n > 1) {
    CU(ctx, cudaMemcpyAsync(sc->fc_dev, sc->fc_host, sizeof(fsb_frame_consts) * (size_t)n, cudaMemcpyHostToDevice,
                            ctx->stream));
    CU(ctx, cudaEventRecord(sc->fc_free, ctx->stream));
  }
  if (ctx->profiling) {
    int prc = prof_collect(ctx);
    if (prc) return prc;
    CU(ctx, cudaEventRecord(ctx->pev[0], ctx->stream));
  }
  if (!pl.split) /* the depth-parallel march builds its own table entries */
    CU(ctx, (cudaError_t)fsb_launch_setup(sc->fc_dev, n == 1 ? &single : NULL, n, sc->table, tab_stride,
                                          ctx->stream, &ctx->launches));
  if (ctx->profiling) CU(ctx, cudaEventRecord(ctx->pev[1], ctx->stream));
  fsb_render_args a;
  memset(&a, 0, sizeof a);
  a.packed = map->packed;
  a.color = map->color;
  a.height = map->height;
  a.q = map->q;
  a.r = map->r;
  a.cq = map->cq;
  a.cr = map->cr;
  a.fc = sc->fc_dev;
  a.table = sc->table;
  a.tab_stride = tab_stride;
  a.xmask_hi = (map->r - 1) & ~7;
  a.ymask_hi = (map->q - 1) & ~3;
  a.log2r = map->log2r;
  a.out = out_dev;
  a.row_stride = row_stride;
  a.pose_stride = pose_stride;
  a.h = h;
  a.w = w;
  a.col_begin = col_begin;
  a.col_end = col_end;
  a.n_poses = n;
  a.filter = prm->filter;
  a.f2i_mode = prm->f2i_mode;
  a.alpha_bits = map->alpha_bits;
  a.stats = ctx->profiling ? ctx->stats_dev : NULL;
  a.recs = (uint2_fsb *)sc->recs;
  a.sidx = sc->sidx;
  a.rec_cap = h + 1; /* + the guard record */
  a.rec_stride = pl.cols ? 32 : 1;
  a.rb_shift = FSB_RB_SHIFT;
  a.smooth = smooth;
  a.n_bands = (h + (1 << FSB_RB_SHIFT) - 1) >> FSB_RB_SHIFT;
  a.tex = map->tex;
  a.tex_h = map->tex_h;
  a.tex_f = map->tex_f;
  a.inv_r = 1.0f / (float)map->r;
  a.inv_q = 1.0f / (float)map->q;
  a.rec4 = pl.rec4;
  a.full_eval = (prm->flags & FSB_FLAG_NO_CULL) ? 1 : 0;
  a.lut = ctx->lut;
  /* single frames: the three dependent launches overlap their scheduling (programmatic dependent launch); FSB_PDL=0: off */
  a.pdl = (pl.split || (n == 1 && !pl.cols)) && !ctx->no_pdl;
  /* small batches on the lanes-over-depth path: the same chaining (2: the expand keeps the plain list walk) */
  if (!a.pdl && !pl.cols && !ctx->no_pdl && getenv("FSB_PDL_BATCH") != NULL && atoi(getenv("FSB_PDL_BATCH")) != 0) a.pdl = 2;
  {
    const char *env = getenv("FSB_LOCAL_CULL"); /* A/B: FSB_LOCAL_CULL=0 keeps only the map-wide occlusion bound */
    if (map->hpyr && !(env && atoi(env) == 0)) {
      a.hpyr = map->hpyr;
      a.pyr_levels = map->pyr_levels;
    }
  }
  a.cand = sc->cand ? sc->cand + FSB_CAND_PAD : NULL;
  a.cand_cnt = sc->cand_cnt;
  a.cand_cap = pl.cand_cap;
  a.ncols_pad = pl.ncols_pad;
  if (pl.cols) {
    if (pl.split)
      CU(ctx, (cudaError_t)fsb_launch_march_split(&a, n == 1 ? &single : NULL, pl.split, (max_nz + 31) / 32, ctx->stream,
                                                  &ctx->launches));
    else
      CU(ctx, (cudaError_t)fsb_launch_march_cols(&a, ctx->stream, &ctx->launches));
    if (ctx->profiling) CU(ctx, cudaEventRecord(ctx->pev[2], ctx->stream));
    if (!pl.paint) CU(ctx, (cudaError_t)fsb_launch_colour(&a, pl.slice_len, ctx->stream, &ctx->launches));
  } else {
    if (pl.frame) CU(ctx, (cudaError_t)fsb_launch_march_frame(&a, pl.frame, ctx->stream, &ctx->launches));
    else CU(ctx, (cudaError_t)fsb_launch_march(&a, pl.mem, ctx->stream, &ctx->launches));
    if (ctx->profiling) CU(ctx, cudaEventRecord(ctx->pev[2], ctx->stream));
  }
  if (ctx->profiling) CU(ctx, cudaEventRecord(ctx->pev[3], ctx->stream));
  if (pl.paint) CU(ctx, (cudaError_t)fsb_launch_paint(&a, pl.seg_bands, ctx->stream, &ctx->launches));
  else CU(ctx, (cudaError_t)fsb_launch_expand(&a, ctx->stream, &ctx->launches));
  if (ctx->profiling) {
    CU(ctx, cudaEventRecord(ctx->pev[4], ctx->stream));
    ctx->prof_pending = 1;
  }
  return FSB_OK;
}

int fsb_render_columns_device(fsb_context *ctx, const fsb_camera *cam, const fsb_params *prm, const fsb_map *map,
                              int h, int w, int col_begin, int col_end, uint32_t *out_dev, int64_t row_stride) {
  int rc = check_common(ctx, cam, 1, prm, map, h, w, col_begin, col_end, out_dev);
  if (rc) return rc;
  if (row_stride == 0) row_stride = col_end - col_begin;
  if (row_stride < col_end - col_begin) return set_err(ctx, FSB_ERR_ARG, "render: row_stride too small");
  CU(ctx, cudaSetDevice(ctx->device));
  return render_poses(ctx, cam, 1, prm, map, h, w, col_begin, col_end, out_dev, row_stride, 0);
}

/* Host-output variant of the column-split mode: the slab is rendered into the context's staging buffer and copied
 * row by row (one 2-D DMA) into the caller's frame at its column offset.  With one process per GPU writing into one
 * shared, page-locked host frame every GPU's PCIe link carries only its own slab. */
int fsb_render_columns(fsb_context *ctx, const fsb_camera *cam, const fsb_params *prm, const fsb_map *map, int h, int w,
                       int col_begin, int col_end, uint32_t *out_host, int64_t row_stride) {
  int rc = check_common(ctx, cam, 1, prm, map, h, w, col_begin, col_end, out_host);
  if (rc) return rc;
  const int ncols = col_end - col_begin;
  if (row_stride == 0) row_stride = ncols;
  if (row_stride < ncols) return set_err(ctx, FSB_ERR_ARG, "render: row_stride too small");
  CU(ctx, cudaSetDevice(ctx->device));
  if ((rc = ensure_frame(ctx, 0, (size_t)h * ncols))) return rc;
  if ((rc = render_poses(ctx, cam, 1, prm, map, h, w, col_begin, col_end, ctx->frame_dev[0], ncols, 0))) return rc;
  CU(ctx, cudaMemcpy2DAsync(out_host, (size_t)row_stride * 4, ctx->frame_dev[0], (size_t)ncols * 4, (size_t)ncols * 4,
                            (size_t)h, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

int fsb_render_device(fsb_context *ctx, const fsb_camera *cam, const fsb_params *prm, const fsb_map *map, int h,
                      int w, uint32_t *out_dev, int64_t row_stride) {
  if (row_stride == 0) row_stride = w;
  return fsb_render_columns_device(ctx, cam, prm, map, h, w, 0, w, out_dev, row_stride);
}

int fsb_render_batch_device(fsb_context *ctx, const fsb_camera *cams, int n, const fsb_params *prm,
                            const fsb_map *map, int h, int w, uint32_t *out_dev) {
  int rc = check_common(ctx, cams, n, prm, map, h, w, 0, w, out_dev);
  if (rc) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  const int64_t frame = (int64_t)h * w;
  const int g = group_size(n, w, h);
  for (int i = 0; i < n; i += g) {
    int c = n - i < g ? n - i : g;
    if ((rc = render_poses(ctx, cams + i, c, prm, map, h, w, 0, w, out_dev + (size_t)i * frame, w, frame))) return rc;
  }
  return FSB_OK;
}

static int ensure_frame(fsb_context *ctx, int slot, size_t pixels) {
  if (pixels > ctx->frame_cap[slot]) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
    cudaFree(ctx->frame_dev[slot]);
    ctx->frame_dev[slot] = NULL;
    ctx->frame_cap[slot] = 0;
    CU(ctx, cudaMalloc((void **)&ctx->frame_dev[slot], pixels * 4));
    ctx->frame_cap[slot] = pixels;
  }
  return FSB_OK;
}

int fsb_render(fsb_context *ctx, const fsb_camera *cam, const fsb_params *prm, const fsb_map *map, int h, int w,
               uint32_t *out_host) {
  return fsb_render_batch(ctx, cam, 1, prm, map, h, w, out_host);
}

/* Host-output batch: frames are rendered in chunks into two device buffers; the D2H copy of
 * chunk i (copy stream) overlaps the rendering of chunk i+1 (render stream). */
int fsb_render_batch(fsb_context *ctx, const fsb_camera *cams, int n, const fsb_params *prm, const fsb_map *map,
                     int h, int w, uint32_t *out_host) {
  int rc = check_common(ctx, cams, n, prm, map, h, w, 0, w, out_host);
  if (rc) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  for (int i = 0; i < n; ++i) /* every pose is validated before anything is queued: no DMA into out_host is pending on an error return */
    if (zs_length(prm->delta, cams[i].distance, prm->z0) < 0)
      return set_err(ctx, FSB_ERR_RANGE, "render: z-series undefined for pose %d (distance=%g delta=%g z0=%g)", i,
                     (double)cams[i].distance, (double)prm->delta, (double)prm->z0);
  const size_t frame = (size_t)h * w;
  size_t chunk = (64u << 20) / (frame * 4);
  if (chunk < 1) chunk = 1;
  if (chunk > (size_t)n) chunk = (size_t)n;
  if (chunk > (size_t)group_size(n, w, h)) chunk = (size_t)group_size(n, w, h);
  const int nbuf = (size_t)n > chunk ? 2 : 1;
  for (int s = 0; s < nbuf; ++s)
    if ((rc = ensure_frame(ctx, s, chunk * frame))) return rc;
  int it = 0;
  for (size_t i = 0; i < (size_t)n; i += chunk, ++it) {
    const int s = it & 1;
    const int c = (int)((size_t)n - i < chunk ? (size_t)n - i : chunk);
    if (it >= 2) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copied[s], 0));
    if ((rc = render_poses(ctx, cams + i, c, prm, map, h, w, 0, w, ctx->frame_dev[s], w, (int64_t)frame))) {
      cudaStreamSynchronize(ctx->stream); /* earlier chunks may still be on their way into out_host */
      cudaStreamSynchronize(ctx->copy_stream);
      return rc;
    }
    CU(ctx, cudaEventRecord(ctx->rendered[s], ctx->stream));
    CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->rendered[s], 0));
    CU(ctx, cudaMemcpyAsync(out_host + i * frame, ctx->frame_dev[s], (size_t)c * frame * 4, cudaMemcpyDeviceToHost,
                            ctx->copy_stream));
    CU(ctx, cudaEventRecord(ctx->copied[s], ctx->copy_stream));
  }
  CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

/* ------------------------------------------------------------------------------------------ */
int fsb_device_malloc(fsb_context *ctx, size_t bytes, void **ptr) {
  if (!ctx || !ptr) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMalloc(ptr, bytes));
  return FSB_OK;
}
int fsb_device_free(fsb_context *ctx, void *ptr) {
  if (!ctx) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaFree(ptr));
  return FSB_OK;
}
int fsb_host_malloc(fsb_context *ctx, size_t bytes, void **ptr) {
  if (!ctx || !ptr) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMallocHost(ptr, bytes));
  return FSB_OK;
}
int fsb_host_free(fsb_context *ctx, void *ptr) {
  if (!ctx) return FSB_ERR_ARG;
  CU(ctx, cudaFreeHost(ptr));
  return FSB_OK;
}
/* Page-lock memory the caller already owns (the host's frame buffer, c/interactive.c:115) so that fsb_render /
 * fsb_render_batch / fsb_copy_to_host move frames into it by DMA at full PCIe rate instead of through the driver's
 * pageable staging path. */
int fsb_host_register(fsb_context *ctx, void *ptr, size_t bytes) {
  if (!ctx || !ptr || !bytes) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return FSB_OK;
}
/* 1 when ptr lies in page-locked host memory this process registered or allocated (a DMA can target it directly) */
int fsb_host_is_registered(fsb_context *ctx, const void *ptr) {
  if (!ctx || !ptr) return 0;
  struct cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return at.type == cudaMemoryTypeHost;
}
int fsb_host_unregister(fsb_context *ctx, void *ptr) {
  if (!ctx || !ptr) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
  CU(ctx, cudaHostUnregister(ptr));
  return FSB_OK;
}
int fsb_copy_to_host(fsb_context *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx || !dst || !src) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return FSB_OK;
}
int fsb_copy_to_device(fsb_context *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx || !dst || !src) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return FSB_OK;
}

/* ------------------------------------------------------------------------------------------ */
int fsb_effect_interpolate_device(fsb_context *ctx, int pd, const uint32_t *in_dev, int h, int w, uint32_t *out_dev) {
  if (!ctx) return FSB_ERR_ARG;
  if (!in_dev || !out_dev || in_dev == out_dev || h <= 0 || w <= 0) return set_err(ctx, FSB_ERR_ARG, "interpolate: bad argument");
  if (h > w) return set_err(ctx, FSB_ERR_RANGE, "interpolate: needs h <= w (fut/effects.fut:36-41 wraps the x taps with %% h)");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, (cudaError_t)fsb_launch_interpolate(in_dev, h, w, 1, pd, out_dev, ctx->stream, &ctx->launches));
  return FSB_OK;
}
int fsb_effect_interpolate2_device(fsb_context *ctx, const uint32_t *in_dev, int h, int w, uint32_t *out_dev) {
  if (!ctx) return FSB_ERR_ARG;
  if (!in_dev || !out_dev || in_dev == out_dev || h <= 0 || w <= 0) return set_err(ctx, FSB_ERR_ARG, "interpolate2: bad argument");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, (cudaError_t)fsb_launch_interpolate(in_dev, h, w, 0, 0, out_dev, ctx->stream, &ctx->launches));
  return FSB_OK;
}

/* ------------------------------------------------------------------------------------------ */
int fsb_ipc_export(fsb_context *ctx, void *dev_ptr, unsigned char handle[64]) {
  if (!ctx || !dev_ptr || !handle) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  CU(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle, &h, sizeof h);
  return FSB_OK;
}
int fsb_ipc_import(fsb_context *ctx, const unsigned char handle[64], void **dev_ptr) {
  if (!ctx || !dev_ptr || !handle) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CU(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FSB_OK;
}
int fsb_ipc_close(fsb_context *ctx, void *dev_ptr) {
  if (!ctx || !dev_ptr) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaIpcCloseMemHandle(dev_ptr));
  return FSB_OK;
}

/* ------------------------------------------------------------------------------------------ */
int fsb_selftest_sqrt(fsb_context *ctx, uint32_t lo_bits, uint32_t hi_bits, uint64_t *mismatches) {
  if (!ctx || !mismatches || lo_bits > hi_bits) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  unsigned long long *d = NULL, h = 0;
  CU(ctx, cudaMalloc((void **)&d, 8));
  CU(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
  CU(ctx, (cudaError_t)fsb_launch_selftest_sqrt(lo_bits, hi_bits, d, ctx->stream));
  CU(ctx, cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d);
  *mismatches = h;
  return FSB_OK;
}

/* ------------------------------------------------------------------------------------------ */
static int time_launches(fsb_context *ctx, int iters, int kind, const uint32_t *buf, size_t n, uint32_t *sink,
                         int per_thread, float *ms) {
  cudaEvent_t t0, t1;
  CU(ctx, cudaEventCreate(&t0));
  CU(ctx, cudaEventCreate(&t1));
  const int blocks = ctx->sm_count * 8;
  for (int i = -3; i < iters; ++i) {
    if (i == 0) CU(ctx, cudaEventRecord(t0, ctx->stream));
    cudaError_t e = kind == 0 ? (cudaError_t)fsb_launch_l2_stream(buf, n, sink, blocks, ctx->stream)
                              : (cudaError_t)fsb_launch_l2_gather(buf, n, sink, blocks, per_thread, ctx->stream);
    CU(ctx, e);
  }
  CU(ctx, cudaEventRecord(t1, ctx->stream));
  CU(ctx, cudaEventSynchronize(t1));
  CU(ctx, cudaEventElapsedTime(ms, t0, t1));
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  return FSB_OK;
}

int fsb_bench_l2_stream(fsb_context *ctx, size_t bytes, int iters, double *gb_per_s) {
  if (!ctx || !gb_per_s || bytes < 4096 || iters <= 0) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  bytes &= ~(size_t)15;
  uint32_t *buf = NULL, *sink = NULL;
  CU(ctx, cudaMalloc((void **)&buf, bytes));
  CU(ctx, cudaMalloc((void **)&sink, 4));
  CU(ctx, cudaMemsetAsync(buf, 1, bytes, ctx->stream));
  float ms = 0;
  int rc = time_launches(ctx, iters, 0, buf, bytes / 4, sink, 0, &ms);
  cudaFree(buf);
  cudaFree(sink);
  if (rc) return rc;
  *gb_per_s = (double)bytes * iters / (ms * 1e-3) / 1e9;
  return FSB_OK;
}

int fsb_bench_l2_gather(fsb_context *ctx, size_t bytes, int iters, double *gsector_per_s) {
  if (!ctx || !gsector_per_s || bytes < 4096 || iters <= 0) return FSB_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  uint32_t *buf = NULL, *sink = NULL;
  CU(ctx, cudaMalloc((void **)&buf, bytes));
  CU(ctx, cudaMalloc((void **)&sink, 4));
  CU(ctx, cudaMemsetAsync(buf, 1, bytes, ctx->stream));
  const int per_thread = 64;
  float ms = 0;
  int rc = time_launches(ctx, iters, 1, buf, bytes / 32, sink, per_thread, &ms);
  cudaFree(buf);
  cudaFree(sink);
  if (rc) return rc;
  const double gathers = (double)ctx->sm_count * 8 * 256 * per_thread * iters;
  *gsector_per_s = gathers / (ms * 1e-3) / 1e9;
  return FSB_OK;
}
