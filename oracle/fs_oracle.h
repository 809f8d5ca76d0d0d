/*
 * fs_oracle.h -- CPU restatement of futspace's render hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The shipped library (futspace_b200/libfutspace_b200.so) never links, loads or calls it.
 *
 * PARITY STATUS: "parity unpinned" for the colour arithmetic.  The reference cannot be
 * compiled here (no futhark; lib/ not vendored) and ships no expected outputs.  argb.mix /
 * argb.scale live in the un-vendored package github.com/athas/matte 0.1.2
 * (#2ce1a39b29c329504cb2d849e305cbcc17f701d5, futhark.pkg:3) and are restated from its
 * published algorithm.  Everything else follows /root/reference line by line (citations at
 * each function) and is pinned by the reference's own known-answer relation
 * (tests/solve_arithm.fut vs tests/javascript_zs.py) and by cross-checking two independent
 * formulations (sequential march vs the reference's literal scan/scatter/scan pipeline).
 */
#ifndef FS_ORACLE_H
#define FS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* fut/voxel_renderer.fut:5-12 */
typedef struct {
  float x, y, height, angle, horizon, distance, fov;
  uint32_t sky_color;
} fso_camera;

enum { FSO_FILTER_NEAREST = 0, FSO_FILTER_BILINEAR = 1 };
enum { FSO_SENTINEL_ZERO = 0, FSO_SENTINEL_SKY = 1 };
/* float->int semantics of `i32.f32` on inf/NaN/out-of-range. */
enum { FSO_F2I_SATURATE = 0, FSO_F2I_X86 = 1, FSO_F2I_MODERN = 2 };

typedef struct {
  float z0;          /* fut/voxel_renderer.fut:103  (0.0)   ; tests/futspace.fut:84 (1.0)   */
  float delta;       /* fut/voxel_renderer.fut:104  (0.001) ; tests/futspace.fut:85 (0.005) */
  float invz_param1; /* numerator of 1.0/z, fut/voxel_renderer.fut:217 (1.0)                 */
  float invz_param2; /* multiplier f32(w/2), :217 ; 240.0 in tests/futspace.fut:91 ; <=0 => f32(w/2) */
  int32_t filter;    /* png_*_filtered (1, fut/interactive.fut:180-181) or png_* (0)         */
  int32_t sentinel;  /* 0: colour 0 is "empty" (voxel_renderer.fut:244-248) ; 1: sky (voxel_renderer_new.fut:182-194) */
  int32_t f2i_mode;
  int32_t smoothing; /* 0: #off (fut/voxel_renderer.fut:214-251) ; 1: #on (:175-213), see fso_render */
} fso_params;

void fso_params_default(fso_params *p);       /* live renderer constants */
void fso_params_tests_variant(fso_params *p); /* tests/futspace.fut constants */

/* fut/voxel_renderer.fut:28-34.  Returns n (>=0) and writes min(n,cap) values; -1 on invalid input. */
int fso_get_zs(float delta, float dist, float z0, float *out, int cap);

/* matte argb (restated; see header comment). */
uint32_t fso_mix(float m1, uint32_t c1, float m2, uint32_t c2);
uint32_t fso_scale(uint32_t c, float s);

/* samplers, fut/render_functions.fut:63-105 */
float fso_height_nearest(const int32_t *hm, int q, int r, float x, float y, int f2i_mode);
float fso_height_bilinear(const int32_t *hm, int q, int r, float x, float y, int f2i_mode);
uint32_t fso_color_nearest(const uint32_t *cm, int q, int r, float x, float y, int f2i_mode);
uint32_t fso_color_bilinear(const uint32_t *cm, int q, int r, float x, float y, int f2i_mode);

/* Smoothing #on (fut/voxel_renderer.fut:175-213) scatters DIFFERENT tuples to the same row: the sample that lowers
 * the running minimum writes (colour, previous colour, row, previous row, idx, previous idx) and every later sample
 * that keeps that minimum rewrites the row with (colour, colour, row, row, idx, idx) (:196-198).  Futhark leaves the
 * winner unspecified; the oracle (and the CUDA path) take the semantics of the sequential `futhark c` backend that
 * BASELINE config 1 names: scans fold left to right starting from the neutral element, scatter writes in index order
 * (the last write wins).  Net effect: the span of sample k is blended towards the previous visible colour
 * (argb.mix by row distance, :204-210) exactly when samples k-1, k and k+1 all lower the y-buffer (or k is the last
 * sample); colour 0 is not transparent in this mode (fill_vline3 only skips the (0,0,0,0,1000,1000) sentinel).
 * Requires the zero sentinel. */

/* Sequential front-to-back march per column (SURVEY.md 8a "equivalent sequential statement").
 * eval_all_colors != 0 evaluates the colour sampler for every sample as the reference does
 * (used for the CPU baseline); 0 evaluates it only for visible samples (same output).
 * nthreads: OpenMP threads over columns (<=0: all).  Returns 0 on success. */
int fso_render(const fso_camera *cam, const fso_params *prm, const uint32_t *color,
               const int32_t *height, int q, int r, int h, int w, uint32_t *out,
               int eval_all_colors, int nthreads);

/* The same with a colour map whose size differs from the height map's; each sampler wraps by the size of the array it
 * reads (fut/render_functions.fut:63-105 take the array's own shape).  This is the state update_map leaves for maps that
 * are not 1024 x 1024: shadowed_color is always baked at 1024 x 1024 (fut/effects.fut:124-125, fut/interactive.fut:194-198). */
int fso_render_split(const fso_camera *cam, const fso_params *prm, const uint32_t *color, int cq, int cr,
                     const int32_t *height, int q, int r, int h, int w, uint32_t *out, int eval_all, int nthreads);

/* The reference's pipeline taken literally: materialise [n_z][w] (colour, y) pairs
 * (voxel_renderer.fut:215-228), per column inclusive scan with `occlude` and neutral (0,h)
 * (:231), scatter into replicate h 0 (:244), inclusive scan with `fill_vline` (:246), sky map
 * (:248), transpose (:251).  O(n_z*w) memory; for small cases. */
int fso_render_literal(const fso_camera *cam, const fso_params *prm, const uint32_t *color,
                       const int32_t *height, int q, int r, int h, int w, uint32_t *out);

/* generate_shadowmap_accumulated, fut/effects.fut:108-125, with the nearest samplers png_color / png_height
 * as update_map passes them (fut/interactive.fut:194-196).  The reference hard-codes a 1024 x 1024 output
 * (:124-125); `out_q x out_r` generalises that (pass 1024, 1024 for the reference behaviour).
 * sun = vec3_rotate #y sun_ang (vec3_rotate #z sun_height [0,1,0]) -- see fso_sun_vector. */
void fso_bake_shadows(const uint32_t *color, const int32_t *height, int q, int r, const float sun[3], int out_q,
                      int out_r, uint32_t *out, int nthreads);
/* fut/effects.fut:6-25 (vec3_rotate via linalg matvecmul_row, restated: row-major dot products summed left to
 * right) applied as in fut/interactive.fut:196 to the initial sun [0,1,0] (:56). */
void fso_sun_vector(float sun_height, float sun_ang, float out[3]);

/* Image-space post-passes of fut/effects.fut (never called by the reference; restated for completeness):
 * interpolate pd (:27-45), 9 taps at distance pd nested in argb.mix 1 _ 1 _, x taps wrap with `% h` as written
 * (:36-41; requires h <= w to stay in bounds); interpolate2 (:47-52), 5 taps via rotate. */
int fso_interpolate(int pd, const uint32_t *img, int h, int w, uint32_t *out);
void fso_interpolate2(const uint32_t *img, int h, int w, uint32_t *out);

/* fut/interactive.fut:189 : height & 0xFF */
void fso_mask_heights(int32_t *hm, long n);

#ifdef __cplusplus
}
#endif
#endif
