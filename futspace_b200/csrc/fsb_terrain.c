/*
 * fsb_terrain.c -- deterministic synthetic terrain for benchmarks and parity tests.
 *
 * Periodic (tileable: the renderer wraps map addresses, fut/render_functions.fut:73-76) fBm value
 * noise: 8 octaves, base cell 256 texels, lacunarity 2, gain 0.5, SplitMix64 lattice hash,
 * smoothstep interpolation.  Output has the format update_map expects (c/interactive.c:50-54):
 * height [m][m] i32 in 0..255, colour [m][m] u32 0xFFrrggbb.  Runs on the host so the oracle
 * and the GPU see identical bytes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/futspace_b200.h"

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

static inline float lattice(uint64_t seed, int oct, int ix, int iy) {
  uint64_t k = seed ^ ((uint64_t)(oct + 1) * 0xD6E8FEB86659FD93ull) ^ ((uint64_t)(uint32_t)ix << 32) ^ (uint64_t)(uint32_t)iy;
  return (float)(splitmix64(k) >> 40) * (1.0f / 16777216.0f);
}

static inline float smooth(float t) { return t * t * (3.0f - 2.0f * t); }

static float fbm(uint64_t seed, int m, int x, int y) {
  float sum = 0.0f, amp = 1.0f, norm = 0.0f;
  for (int o = 0; o < 8; ++o) {
    int cell = 256 >> o;         /* texels per lattice cell */
    int period = m / cell;       /* lattice points per side */
    if (period < 1) period = 1;
    int ix = x / cell, iy = y / cell;
    float tx = smooth((float)(x % cell) / (float)cell), ty = smooth((float)(y % cell) / (float)cell);
    int ix1 = (ix + 1) % period, iy1 = (iy + 1) % period;
    ix %= period;
    iy %= period;
    float a = lattice(seed, o, ix, iy), b = lattice(seed, o, ix1, iy);
    float c = lattice(seed, o, ix, iy1), d = lattice(seed, o, ix1, iy1);
    float top = a + (b - a) * tx, bot = c + (d - c) * tx;
    sum += amp * (top + (bot - top) * ty);
    norm += amp;
    amp *= 0.5f;
  }
  return sum / norm;
}

static inline uint32_t ramp(int h) { /* water / sand / grass / rock / snow */
  static const int stops[6] = {0, 64, 80, 140, 200, 255};
  static const int rgb[6][3] = {{24, 60, 140}, {40, 90, 170}, {196, 180, 120}, {60, 130, 50}, {120, 110, 100}, {245, 245, 250}};
  int s = 0;
  while (s < 4 && h > stops[s + 1]) ++s;
  float t = (float)(h - stops[s]) / (float)(stops[s + 1] - stops[s]);
  if (t < 0.0f) t = 0.0f;
  if (t > 1.0f) t = 1.0f;
  int r = (int)((float)rgb[s][0] + t * (float)(rgb[s + 1][0] - rgb[s][0]));
  int g = (int)((float)rgb[s][1] + t * (float)(rgb[s + 1][1] - rgb[s][1]));
  int b = (int)((float)rgb[s][2] + t * (float)(rgb[s + 1][2] - rgb[s][2]));
  return ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
}

int fsb_terrain_fbm(int m, uint64_t seed, uint32_t *color, int32_t *height) {
  if (m < 256 || (m % 256) != 0 || !color || !height) return FSB_ERR_ARG;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < m; ++y)
    for (int x = 0; x < m; ++x) {
      float v = (fbm(seed, m, x, y) - 0.5f) * 2.4f + 0.5f;
      if (v < 0.0f) v = 0.0f;
      if (v > 1.0f) v = 1.0f;
      int h = (int)(255.0f * v);
      if (h < 64) h = 64; /* sea level */
      height[(size_t)y * m + x] = h;
    }
#pragma omp parallel for schedule(static)
  for (int y = 0; y < m; ++y)
    for (int x = 0; x < m; ++x) {
      int h = height[(size_t)y * m + x];
      int hl = height[(size_t)y * m + (x + m - 1) % m], hr = height[(size_t)y * m + (x + 1) % m];
      int hu = height[(size_t)((y + m - 1) % m) * m + x], hd = height[(size_t)((y + 1) % m) * m + x];
      float shade = 0.75f + 0.06f * (float)((hl - hr) + (hu - hd)); /* light from the upper left */
      if (shade < 0.35f) shade = 0.35f;
      if (shade > 1.15f) shade = 1.15f;
      uint32_t c = ramp(h);
      int dither = (int)(splitmix64(seed ^ ((uint64_t)y << 32) ^ (uint64_t)x) & 15u) - 8;
      int ch[3] = {(int)((c >> 16) & 255u), (int)((c >> 8) & 255u), (int)(c & 255u)};
      for (int k = 0; k < 3; ++k) {
        int t = (int)((float)ch[k] * shade) + dither;
        ch[k] = t < 0 ? 0 : (t > 255 ? 255 : t);
      }
      color[(size_t)y * m + x] = 0xFF000000u | ((uint32_t)ch[0] << 16) | ((uint32_t)ch[1] << 8) | (uint32_t)ch[2];
    }
  return FSB_OK;
}
