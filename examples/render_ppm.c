/* Minimal C client of the native C-ABI (include/futspace_b200.h): synthetic terrain -> one frame -> frame.ppm.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/render_ppm.c -Lfutspace_b200 -lfutspace_b200 -Wl,-rpath,$PWD/futspace_b200 -o render_ppm
 *   ./render_ppm [width height distance]          (needs an sm_100 GPU: the library has no CPU fallback)
 *
 * It is the call sequence INTEGRATION.md section A gives for c/interactive.c: context, map (masking the heights as
 * update_map does), a page-locked frame buffer the host owns, fsb_render per frame, error text from the context. */
#include <stdio.h>
#include <stdlib.h>

#include "futspace_b200.h"

static int fail(fsb_context *ctx, const char *what) {
  fprintf(stderr, "%s: %s\n", what, ctx ? fsb_context_get_error(ctx) : "no context");
  return EXIT_FAILURE;
}

int main(int argc, char **argv) {
  const int w = argc > 2 ? atoi(argv[1]) : 1024, h = argc > 2 ? atoi(argv[2]) : 768;
  const float distance = argc > 3 ? (float)atof(argv[3]) : 1000.0f;
  const int m = 1024;
  fsb_context *ctx = NULL;
  if (fsb_context_new(0, &ctx) != FSB_OK) {
    fprintf(stderr, "no usable sm_100 GPU\n");
    return EXIT_FAILURE;
  }
  uint32_t *color = (uint32_t *)malloc((size_t)m * m * 4);
  int32_t *height = (int32_t *)malloc((size_t)m * m * 4);
  uint32_t *frame = (uint32_t *)malloc((size_t)w * h * 4);
  if (!color || !height || !frame) return EXIT_FAILURE;
  if (fsb_terrain_fbm(m, 0x5EED5EEDu, color, height)) return fail(ctx, "fsb_terrain_fbm");
  fsb_map *map = NULL;
  if (fsb_map_new(ctx, color, height, m, m, /*mask_heights=*/1, &map)) return fail(ctx, "fsb_map_new");
  if (fsb_host_register(ctx, frame, (size_t)w * h * 4)) return fail(ctx, "fsb_host_register");

  fsb_params prm;
  fsb_params_default(&prm); /* z_0 = 0, d = 0.001, bilinear samplers, inv_z = (1 / z) * f32(w / 2) */
  fsb_camera cam = {m / 2 + 0.37f, m / 2 + 0.73f, 200.0f, 2.2f, 0.3f * (float)h, distance, 1.2f, 0xFF9090E0u};
  if (fsb_render(ctx, &cam, &prm, map, h, w, frame)) return fail(ctx, "fsb_render");

  FILE *f = fopen("frame.ppm", "wb");
  if (!f) return EXIT_FAILURE;
  fprintf(f, "P6\n%d %d\n255\n", w, h);
  for (long i = 0; i < (long)w * h; ++i) {
    const unsigned char rgb[3] = {(unsigned char)(frame[i] >> 16), (unsigned char)(frame[i] >> 8), (unsigned char)frame[i]};
    fwrite(rgb, 1, 3, f);
  }
  fclose(f);
  char name[128];
  fsb_context_device_name(ctx, name, sizeof name);
  printf("%dx%d frame, distance %g, on %s: frame.ppm (%lld kernel launches)\n", w, h, (double)distance, name,
         (long long)fsb_context_launch_count(ctx));

  fsb_host_unregister(ctx, frame);
  fsb_map_free(ctx, map);
  fsb_context_free(ctx);
  free(color); free(height); free(frame);
  return EXIT_SUCCESS;
}
