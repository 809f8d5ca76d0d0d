/*
 * futspace_cli.c -- stand-in for the executable `futhark opencl tests/futspace.fut` produces, so the
 * reference's own test pipeline runs against this library (tests/Makefile:1-5, tools/runfutspace.sh):
 *
 *     cat C1W.in D1.in | ./futspace -b -D > img_map.data ; python tools/data2png.py img_map.data success.png
 *
 * stdin : two values of type [q][r]i32 in the Futhark binary data format (tools/png2data.py:50-57:
 *         'b', version 2, rank, 4-char type, rank x u64 dims, little-endian data): colour map, height map.
 * stdout: main's result [400][800]i32 (tests/futspace.fut:126-145), binary with -b, Futhark text otherwise.
 * -D    : report device and timing on stderr;  -t FILE: write the runtime in microseconds;  -r N: N timed runs.
 * Exit status 1 with a message on stderr when no sm_100 GPU is present (there is no CPU fallback).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/futspace_b200.h"

static int read_i32_2d(FILE *f, int32_t **data, int64_t *d0, int64_t *d1) {
  int c;
  do c = fgetc(f); while (c == ' ' || c == '\n' || c == '\t' || c == '\r');
  if (c != 'b') { fprintf(stderr, "futspace: expected a value in the Futhark binary data format on stdin\n"); return 1; }
  unsigned char hdr[6];
  if (fread(hdr, 1, 6, f) != 6 || hdr[0] != 2) { fprintf(stderr, "futspace: unsupported binary format version\n"); return 1; }
  if (hdr[1] != 2 || (memcmp(hdr + 2, " i32", 4) != 0 && memcmp(hdr + 2, " u32", 4) != 0)) {
    fprintf(stderr, "futspace: expected [][]i32, got rank %d type '%.4s'\n", hdr[1], (char *)hdr + 2);
    return 1;
  }
  uint64_t dims[2];
  if (fread(dims, 8, 2, f) != 2) return 1;
  const size_t n = (size_t)dims[0] * dims[1];
  *data = (int32_t *)malloc(n * 4 + 4);
  if (!*data || fread(*data, 4, n, f) != n) { fprintf(stderr, "futspace: truncated input\n"); return 1; }
  *d0 = (int64_t)dims[0];
  *d1 = (int64_t)dims[1];
  return 0;
}

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(int argc, char **argv) {
  int binary = 0, debug = 0, runs = 1;
  const char *tfile = NULL;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-b")) binary = 1;
    else if (!strcmp(argv[i], "-D")) debug = 1;
    else if (!strcmp(argv[i], "-t") && i + 1 < argc) tfile = argv[++i];
    else if (!strcmp(argv[i], "-r") && i + 1 < argc) runs = atoi(argv[++i]);
    else { fprintf(stderr, "usage: futspace [-b] [-D] [-t FILE] [-r RUNS] < colour.in+height.in\n"); return 2; }
  }
  fsb_context *ctx = NULL;
  int rc = fsb_context_new(0, &ctx);
  if (rc) { fprintf(stderr, "futspace: no sm_100 GPU available (fsb_context_new = %d); there is no CPU fallback\n", rc); return 1; }
  int32_t *color = NULL, *height = NULL;
  int64_t q, r, q2, r2;
  if (read_i32_2d(stdin, &color, &q, &r) || read_i32_2d(stdin, &height, &q2, &r2)) return 1;
  if (q != q2 || r != r2) { fprintf(stderr, "futspace: colour map %ldx%ld and height map %ldx%ld differ\n", (long)q, (long)r, (long)q2, (long)r2); return 1; }
  fsb_map *map = NULL;
  /* tests/futspace.fut uses the height unmasked (:96-97) */
  if (fsb_map_new(ctx, (const uint32_t *)color, height, (int)q, (int)r, 0, &map)) { fprintf(stderr, "futspace: %s\n", fsb_context_get_error(ctx)); return 1; }
  const fsb_camera cam = {512.0f, 800.0f, 78.0f, 0.0f, 100.0f, 800.0f, 1.0f, 0xFF9090e0u}; /* tests/futspace.fut:129-143 */
  fsb_params prm;
  fsb_params_tests_variant(&prm);
  const int l = 400, m = 800; /* :145 */
  uint32_t *frame = (uint32_t *)malloc((size_t)l * m * 4);
  if (!frame) return 1;
  double best = 1e300;
  for (int i = 0; i < (runs > 0 ? runs : 1); ++i) {
    const double t0 = now_us();
    if (fsb_render(ctx, &cam, &prm, map, l, m, frame)) { fprintf(stderr, "futspace: %s\n", fsb_context_get_error(ctx)); return 1; }
    const double dt = now_us() - t0;
    if (dt < best) best = dt;
  }
  if (debug) {
    char name[128];
    fsb_context_device_name(ctx, name, sizeof name);
    fprintf(stderr, "futspace: device %s, map %ldx%ld, frame %dx%d, %.1f us (render + copy to host)\n", name, (long)q, (long)r, m, l, best);
  }
  if (tfile) {
    FILE *t = fopen(tfile, "w");
    if (t) { fprintf(t, "%ld\n", (long)best); fclose(t); }
  }
  if (binary) {
    const unsigned char hdr[7] = {'b', 2, 2, ' ', 'i', '3', '2'};
    const uint64_t dims[2] = {(uint64_t)l, (uint64_t)m};
    fwrite(hdr, 1, 7, stdout);
    fwrite(dims, 8, 2, stdout);
    fwrite(frame, 4, (size_t)l * m, stdout);
  } else {
    putchar('[');
    for (int y = 0; y < l; ++y) {
      fputs(y ? ", [" : "[", stdout);
      for (int x = 0; x < m; ++x) printf(x ? ", %di32" : "%di32", (int32_t)frame[(size_t)y * m + x]);
      putchar(']');
    }
    puts("]");
  }
  fsb_map_free(ctx, map);
  fsb_context_free(ctx);
  free(frame); free(color); free(height);
  return 0;
}
