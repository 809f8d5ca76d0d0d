/* fsb_internal.h -- types shared by the C host layer (fsb_api.c) and the CUDA translation unit. */
#ifndef FSB_INTERNAL_H
#define FSB_INTERNAL_H

#include <stdint.h>
#include "../../include/futspace_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-pose constants, computed on the host (sinf/cosf are frame-uniform, fut/voxel_renderer.fut:44-45)
 * and consumed by the set-up kernel that builds the per-depth line table. */
typedef struct fsb_frame_consts {
  float a_lx, a_ly, a_rx, a_ry; /* (-cos - sin*fov), (sin - cos*fov), (cos - sin*fov), (-sin - cos*fov)  :47-50 */
  float cam_x, cam_y, cam_h, horizon;
  float fw;                     /* f32 w  :52-53 */
  float invz_num, invz_mul;     /* (num / z) * mul  :217 */
  float z0, delta;
  int32_t n_z;
  uint32_t sky, empty;          /* empty = 0 (zero sentinel) or sky (sky sentinel) */
  float cull_d;                 /* cam_h - (highest terrain + 0.5): no sample can project above this bound's row */
  int32_t reserved;
} fsb_frame_consts;

#ifdef __CUDACC__
typedef uint2 uint2_fsb;
#else
typedef struct { uint32_t x, y; } uint2_fsb;
#endif

typedef struct fsb_render_args {
  const uint32_t *packed;   /* height<<24 | rgb in 8x4-texel tiles (fsb_kernels.cu texel_x/texel_y), or NULL */
  int32_t xmask_hi, ymask_hi, log2r; /* tiled addressing: (r-1)&~7, (q-1)&~3, log2(r) */
  unsigned long long tex;   /* cudaTextureObject_t over an RGBA8 array of the packed texels (bytes B,G,R,height), or 0 */
  unsigned long long tex_f; /* the same array read as normalised floats: a channel arrives as c/255, correctly rounded */
  unsigned long long tex_h; /* cudaTextureObject_t over an R16F array of the heights (exact for 0..255)                */
  float inv_r, inv_q;       /* 1/r, 1/q for normalised texture coordinates */
  const uint32_t *color;    /* [q][r] argb  */
  const int32_t *height;    /* [q][r]       */
  int32_t q, r;             /* height plane (and every packed form): rows, columns */
  int32_t cq, cr;           /* colour plane of the generic path (= q, r unless fsb_map_new_split) */
  const fsb_frame_consts *fc; /* device, [n_poses] */
  const float *table;       /* device, [n_poses][tab_stride]: per pose kcap x {sx,sy,dx,dy}, then kcap x inv_z    */
  int32_t tab_stride;       /* floats per pose = 5 * kcap, kcap = 32 * (n_chunks + 6): the march prefetches past the last chunk */
  uint32_t *out;            /* device; pixel (pose 0, row 0, column col_begin)              */
  int64_t row_stride;       /* pixels */
  int64_t pose_stride;      /* pixels */
  int32_t h, w;             /* frame size (w = full frame width: column index base)         */
  int32_t col_begin, col_end;
  int32_t n_poses;
  int32_t filter, f2i_mode;
  uint32_t alpha_bits;      /* packed maps: the map-uniform alpha byte << 24 */
  /* march -> expand hand-off (device scratch, L2-resident in steady state) */
  uint2_fsb *recs;          /* [n_poses][ncols][rec_cap] {row, colour}: visible samples front to back  */
  uint32_t *sidx;           /* [n_poses][ncols][n_bands+1]: sidx[b] = #records with row >= b<<rb_shift */
  int32_t rec_cap;          /* = h + 1: the guard slot, then at most h records (rows strictly decrease along a list) */
  int32_t rec_stride;       /* 1: a contiguous list per column; 32: lists of 32 adjacent columns interleaved (fsb_kernels.cu list_view) */
  int32_t n_bands, rb_shift;
  int32_t rec4;             /* 1: 4-byte records (row & 31) << 24 | alpha flag << 31 | rgb (packed maps, alpha 0x00 / 0xFF) */
  int32_t smooth;           /* smoothing #on: record .x = row | sample index << 15 (fsb_expand_smooth_kernel) */
  unsigned long long *stats; /* optional (profiling): [0] += chunks of 32 samples evaluated, [1] += records emitted */
  /* column-parallel march (fsb_march_cols.cu): per (pose, group of 32 columns) the interleaved candidate lists of
   * row | sample index << 15 words, and per column the list length */
  uint32_t *cand;           /* [n_poses][ncols_pad / 32][cand_cap][32]                                        */
  uint32_t *cand_cnt;       /* [n_poses][ncols_pad]                                                           */
  int32_t cand_cap;         /* >= min(h, depth samples)                                                       */
  int32_t ncols_pad;        /* columns rounded up to a multiple of 32                                         */
  int32_t full_eval;        /* FSB_FLAG_NO_CULL: also no early exit when the y-buffer reaches row 0           */
  const float *lut;         /* c/255 [0..255] and its square [256..511] (device, filled once per device)      */
  const uint8_t *hpyr;      /* pyramid of local height maxima (fsb_api.c build_height_pyramid), or NULL: no local occlusion bound */
  int32_t pyr_levels;
  int32_t pdl;              /* single frames: march / colour / expand are launched with programmatic stream serialization */
} fsb_render_args;

/* launchers implemented in fsb_kernels.cu; stream is a cudaStream_t passed as void*.
 * Return a cudaError_t value (0 = success). *launches is incremented per kernel launch. */
/* single != NULL: one pose whose constants travel as a kernel argument (and are stored to fc_dev[0]
 * by the kernel); otherwise fc_dev[n_poses] must already be in device memory. */
int fsb_launch_setup(const fsb_frame_consts *fc_dev, const fsb_frame_consts *single, int n_poses, float *table,
                     int tab_stride, void *stream, int64_t *launches);
#define FSB_MEM_PLANES 0
#define FSB_MEM_TILED 1
#define FSB_MEM_TEX 2
int fsb_launch_march(const fsb_render_args *a, int mem, void *stream, int64_t *launches);
int fsb_launch_expand(const fsb_render_args *a, void *stream, int64_t *launches);
/* expand of 4-byte records with TMA tile stores (fsb_expand_tma.cu); applicable: base and strides 16-byte aligned */
int fsb_expand_tma_applicable(const fsb_render_args *a);
int fsb_launch_expand_tma(const fsb_render_args *a, void *stream, int64_t *launches);
/* march of single frames and small batches on the texture path: one CTA per column, warps_per_column (2, 3 or 4) warps over its chunks (fsb_march_frame.cu) */
int fsb_launch_march_frame(const fsb_render_args *a, int warps_per_column, void *stream, int64_t *launches);
/* depth-parallel march of single frames and small batches on the texture path (fsb_march_split.cu): one cluster of 8 CTAs per
 * group of 32 columns, warps_per_group (32 or 64) depth segments; no set-up launch (single != NULL: the one pose's constants) */
int fsb_march_split_max_chunks(int warps_per_group);
int fsb_launch_march_split(const fsb_render_args *a, const fsb_frame_consts *single, int warps_per_group, int n_chunks_max,
                           void *stream, int64_t *launches);
/* column-parallel march of the texture path (fsb_march_cols.cu) */
int fsb_launch_march_cols(const fsb_render_args *a, void *stream, int64_t *launches);
/* colour pass: candidate lists -> the records fsb_launch_expand consumes; slice_len: records per warp (0: whole lists) */
int fsb_launch_colour(const fsb_render_args *a, int slice_len, void *stream, int64_t *launches);
/* colour pass + expand as one kernel (fsb_paint.cu): 4-byte record case of the column-parallel path; seg_bands: bands of 32 rows
 * per warp (0: the whole column) */
int fsb_launch_paint(const fsb_render_args *a, int seg_bands, void *stream, int64_t *launches);
const float *fsb_lut_device_address(void); /* after fsb_launch_lut_init */
int fsb_launch_lut_init(void *stream); /* fills the colour look-up table of the march (once per device, before any render) */
int fsb_launch_shadow(const uint32_t *color, const int32_t *height, int q, int r, const float *sun, int out_q, int out_r,
                      uint32_t *out, void *stream, int64_t *launches);
int fsb_launch_interpolate(const uint32_t *img, int h, int w, int mode, int pd, uint32_t *out, void *stream, int64_t *launches);
int fsb_launch_selftest_sqrt(uint32_t lo, uint32_t hi, unsigned long long *mismatches_dev, void *stream);
int fsb_launch_l2_stream(const uint32_t *buf, size_t n_words, uint32_t *sink, int blocks, void *stream);
int fsb_launch_l2_gather(const uint32_t *buf, size_t n_sectors, uint32_t *sink, int blocks, int per_thread,
                         void *stream);

/* test hook: the pyramid of local height maxima fsb_map_new uploads (fsb_api.c build_height_pyramid); -> levels */
int fsb_debug_height_pyramid(const int32_t *hm, int q, int r, uint8_t *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
