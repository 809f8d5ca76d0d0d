/*
 * fsb_colour.cuh -- png_color / png_color_filtered (fut/render_functions.fut:91-105) as the colour pass
 * (fsb_march_cols.cu fsb_colour_kernel) and the fused paint kernel (fsb_paint.cu) evaluate it for one visible sample.
 *
 * The bilinear filter is the unit-weight, alpha 0x00/0xFF form of fsb_device.cuh (sample_color) spelled for these passes:
 *   - ceil x = floor x + (x > floor x), as in the march: two FRND instead of four;
 *   - u32.f32 (v * 255) for v * 255 in [0, 256) is the low byte of round-toward-zero (v * 255 + 2^23): an FADD.RZ on the
 *     FP32 pipe instead of an F2I on the quarter-rate XU pipe, which the nine square roots of a sample already load.
 * Anything else (integer coordinate, |coordinate| < 1, other alpha) takes sample_color unchanged.
 */
#ifndef FSB_COLOUR_CUH
#define FSB_COLOUR_CUH
#include "fsb_device.cuh"

__device__ __forceinline__ uint32_t byte_bits(float v) { /* 0x4B0000nn with nn = u32.f32 (v * 255), 0 <= v * 255 < 256 */
  return __float_as_uint(__fadd_rz(__fmul_rn(v, 255.0f), 8388608.0f));
}
__device__ __forceinline__ float mix_unit_sqrt(float m1, float s1, float m2, float s2) {
  return sqrt_rn_unit(__fadd_rn(__fmul_rn(m1, s1), __fmul_rn(m2, s2)));
}
/* (nn/255)^2 from the shared-memory table for bits = 0x4B0000nn: the address is one multiply-add,
 * bits * 4 + (table - 0x4B000000 * 4) in wrap-around 32-bit arithmetic, instead of a mask and a shift */
__device__ __forceinline__ float sq_of_bits(uint32_t bits, uint32_t sq_sm_biased) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(bits * 4u + sq_sm_biased));
  return v;
}
/* one channel of png_color_filtered from the four normalised texels: -> 0x4B0000nn */
__device__ __forceinline__ uint32_t colour_channel(float v00, float v01, float v10, float v11, float wx0, float wx1, float wy0,
                                                   float wy1, uint32_t sq_sm_biased) {
  const uint32_t i1 = byte_bits(mix_unit_sqrt(wx0, __fmul_rn(v00, v00), wx1, __fmul_rn(v01, v01)));
  const uint32_t i2 = byte_bits(mix_unit_sqrt(wx0, __fmul_rn(v10, v10), wx1, __fmul_rn(v11, v11)));
  return byte_bits(mix_unit_sqrt(wy0, sq_of_bits(i1, sq_sm_biased), wy1, sq_of_bits(i2, sq_sm_biased)));
}

/* REC4: the launch writes 4-byte records, which the host only selects when the map's alpha byte is 0x00 or 0xFF */
template <bool BIL, bool REC4>
__device__ __forceinline__ uint32_t colour_of(const fsb_render_args &a, float x, float y, const float *un, const float *sq,
                                              uint32_t sq_sm) {
  if (!BIL) return sample_color<MEM_TEX, false, FSB_F2I_SATURATE>(a, x, y, un, sq);
  const float fx = floorf(x), fy = floorf(y);
  const float wx1 = __fsub_rn(x, fx), wy1 = __fsub_rn(y, fy);
  const float wx0 = __fsub_rn(__fadd_rn(fx, wx1 > 0.0f ? 1.0f : 0.0f), x), wy0 = __fsub_rn(__fadd_rn(fy, wy1 > 0.0f ? 1.0f : 0.0f), y);
  const uint32_t al = a.alpha_bits;
  if ((REC4 || al == 0xFF000000u || al == 0u) && __fadd_rn(wx0, wx1) == 1.0f && __fadd_rn(wy0, wy1) == 1.0f) {
    const float u = __fmul_rn(__fadd_rn(fx, 1.0f), a.inv_r), v = __fmul_rn(__fadd_rn(fy, 1.0f), a.inv_q);
    float r00, r01, r10, r11, g00, g01, g10, g11, b00, b01, b10, b11;
    FSB_TLD4_F32C("b", a.tex_f, u, v, r10, r11, r01, r00); /* channel order of the RGBA8 texel is {B, G, R, height} */
    FSB_TLD4_F32C("g", a.tex_f, u, v, g10, g11, g01, g00);
    FSB_TLD4_F32C("r", a.tex_f, u, v, b10, b11, b01, b00);
    const uint32_t r = colour_channel(r00, r01, r10, r11, wx0, wx1, wy0, wy1, sq_sm);
    const uint32_t g = colour_channel(g00, g01, g10, g11, wx0, wx1, wy0, wy1, sq_sm);
    const uint32_t b = colour_channel(b00, b01, b10, b11, wx0, wx1, wy0, wy1, sq_sm);
    /* low bytes of b, g, r under the alpha byte */
    return (__byte_perm(__byte_perm(b, g, 0x0040), r, 0x7410) & 0x00FFFFFFu) | al;
  }
  return sample_color<MEM_TEX, true, FSB_F2I_SATURATE>(a, x, y, un, sq);
}


#endif
