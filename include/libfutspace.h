/*
 * libfutspace.h -- the C API `futhark --library` generates from fut/interactive_entrypoints.fut
 * (Makefile:24-25 of the reference; the generated header is a build product there), re-implemented on
 * top of the native C-ABI (futspace_b200.h) so c/interactive.c and lys's liblys.c link unmodified.
 *
 * Same names, argument order and conventions as the generated API: entry points return 0 on success;
 * outputs come first, inputs last; every output is a fresh object the caller frees; inputs are not
 * consumed; futhark_context_get_error returns a malloc'ed string (or NULL) that the caller frees.
 *
 * The nine entry points are those of fut/interactive_entrypoints.fut:6-32; the state machine behind them
 * follows fut/interactive.fut:28-198 and fut/interactive_input.fut.  Key `2` toggles the smoothing
 * variant (fut/voxel_renderer.fut:175-213, rendered with FSB_FLAG_SMOOTHING: see futspace_b200.h for the semantics
 * adopted where the reference is unspecified).  Not implemented (out of the hot-path scope): the #math
 * render mode (key `1` toggles the flag, rendering stays #png).
 */
#ifndef LIBFUTSPACE_H
#define LIBFUTSPACE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct futhark_context_config;
struct futhark_context;
struct futhark_u32_2d;
struct futhark_i32_2d;
struct futhark_opaque_state;

/* context (c/interactive.c:144-148 via lys_setup_futhark_context) */
struct futhark_context_config *futhark_context_config_new(void);
void futhark_context_config_free(struct futhark_context_config *cfg);
void futhark_context_config_set_device(struct futhark_context_config *cfg, const char *s); /* "#k" or "k": CUDA device index */
void futhark_context_config_set_debugging(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_profiling(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_logging(struct futhark_context_config *cfg, int flag);
struct futhark_context *futhark_context_new(struct futhark_context_config *cfg);
void futhark_context_free(struct futhark_context *ctx);
int futhark_context_sync(struct futhark_context *ctx);
char *futhark_context_get_error(struct futhark_context *ctx);
/* the housekeeping entries every generated library has (unused by c/interactive.c and liblys.c, provided so that
 * other hosts written against libfutspace.h link): release cached device buffers; a malloc'ed one-line report;
 * profiling pause / unpause (no-ops: this library has no background profiling) */
int futhark_context_clear_caches(struct futhark_context *ctx);
char *futhark_context_report(struct futhark_context *ctx);
void futhark_context_pause_profiling(struct futhark_context *ctx);
void futhark_context_unpause_profiling(struct futhark_context *ctx);

/* arrays (c/interactive.c:50-56; outputs of `render`) */
struct futhark_u32_2d *futhark_new_u32_2d(struct futhark_context *ctx, const uint32_t *data, int64_t dim0, int64_t dim1);
int futhark_free_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr);
int futhark_values_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr, uint32_t *data);
const int64_t *futhark_shape_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr);
struct futhark_i32_2d *futhark_new_i32_2d(struct futhark_context *ctx, const int32_t *data, int64_t dim0, int64_t dim1);
int futhark_free_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);
int futhark_values_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr, int32_t *data);
const int64_t *futhark_shape_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);
int futhark_free_opaque_state(struct futhark_context *ctx, struct futhark_opaque_state *obj);

/* entry points, fut/interactive_entrypoints.fut:6-32 */
int futhark_entry_init(struct futhark_context *ctx, struct futhark_opaque_state **out0, const uint32_t in0);
int futhark_entry_resize(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t in0,
                         const int32_t in1, const struct futhark_opaque_state *in2);
int futhark_entry_key(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t in0,
                      const int32_t in1, const struct futhark_opaque_state *in2);
int futhark_entry_mouse(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t in0,
                        const int32_t in1, const int32_t in2, const struct futhark_opaque_state *in3);
int futhark_entry_wheel(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t in0,
                        const int32_t in1, const struct futhark_opaque_state *in2);
int futhark_entry_step(struct futhark_context *ctx, struct futhark_opaque_state **out0, const float in0,
                       const struct futhark_opaque_state *in1);
int futhark_entry_render(struct futhark_context *ctx, struct futhark_u32_2d **out0, const struct futhark_opaque_state *in0);
int futhark_entry_text_content(struct futhark_context *ctx, float *out0, float *out1, float *out2, float *out3,
                               float *out4, float *out5, float *out6, float *out7, float *out8,
                               const struct futhark_opaque_state *in0);
int futhark_entry_update_map(struct futhark_context *ctx, struct futhark_opaque_state **out0,
                             const struct futhark_u32_2d *in0, const struct futhark_i32_2d *in1,
                             const struct futhark_opaque_state *in2);

#ifdef __cplusplus
}
#endif
#endif
