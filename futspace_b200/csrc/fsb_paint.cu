/*
 * fsb_paint.cu -- colour pass and expand as ONE kernel for batches on the texture path (sm_100a).
 *
 *   fsb_paint_kernel   one warp per (pose, group of 32 columns[, segment of bands]), lane = column.  The warp walks its
 *                      columns' candidate lists (fsb_marchc_kernel: row | sample index << 15, rows strictly decreasing)
 *                      BACKWARD, i.e. down the screen from row 0: png_color / png_color_filtered of a candidate
 *                      (fut/render_functions.fut:91-105) goes into a per-lane ring of 32 colours in shared memory, and band
 *                      by band (32 rows) the ring is drained into pixels: replicate + scatter (fut/voxel_renderer.fut:244),
 *                      the fill_vline scan (:246), sky (:248) and the transpose (:251) exactly as fsb_expand4_kernel does
 *                      them -- the running colour is a register that lives across the whole column.
 *
 * Why (round 2, profiles/r2_paint_*): after the column-parallel march the 1080p step was march 0.92 + colour 1.05 + expand
 * 0.89 ms per 512 poses.  The colour pass is bound by instruction issue (86 % of the slots, 1.9 % of DRAM), the expand kernel
 * by HBM stores (0.73 of the copy peak, 20 % of the issue slots): run one after the other each leaves the other's resource
 * idle, and running them as two grids on two streams does not mix them (the block scheduler drains one grid first:
 * profiles/r2_overlap_experiment.jsonl).  Inside one kernel every SM holds warps of both phases at all times, the records
 * never travel through DRAM (1.0 MB written + 1.6 MB read per 1080p pose before, none now), the band index and its
 * dependent global loads are gone, and the walk reads its next record from shared memory.
 *
 * Ring and masks: a ring entry is the finished 32-bit colour; WHERE the entries start is kept per lane as two 32-bit row
 * masks, of the band being painted and of the next one (a lane colours ahead of the band being painted while its ring has
 * room and the candidate lies in the next band at most).  The row loop tests one mask bit per row; an entry is popped
 * when the bit is set.  Lanes colour ahead independently, so a trip of 32 lanes is full as long as the lists are not
 * exhausted: 93 % of the lanes of a trip filter a record (fsb_context_paint_trips).
 *
 * Only launched for the 4-byte record case (packed map whose alpha byte is 0x00 or 0xFF, no smoothing); everything else keeps
 * fsb_colour_kernel + fsb_expand*_kernel.
 *
 * Float discipline as in fsb_kernels.cu: every parity-relevant operation uses the round-to-nearest intrinsics.
 */
#include <stdlib.h>

#include "fsb_colour.cuh"
#include "fsb_device.cuh"

#define FSB_PAINT_WARPS 4
#define FSB_RING 32      /* entries per lane (power of two, >= 32: a band can hold one record per row) */
#define FSB_PAINT_SMEM ((FSB_PAINT_WARPS + 1) * 4096)
#define FSB_PAINT_PF 4   /* L2 prefetch distance of the candidate words, in trips (measured: 3-5 alike, 8+ worse) */
#define FSB_RUNAHEAD 32  /* rows past the end of the band a lane may colour ahead: the next band (mask1) */

/* seg_bands > 0: blockIdx.z selects a segment of seg_bands bands of the frame (medium batches: more, shorter warps); the
 * warp finds its first candidate by a four-way search of the list (rows strictly decrease along it) and colours the one record above
 * its first band a second time for the running colour that enters it.  seg_bands == 0: the whole column. */
template <bool BIL, int V>
__global__ void __launch_bounds__(FSB_PAINT_WARPS * 32, (V & 2) ? 10 : (V & 1) ? 9 : 8) fsb_paint_kernel(const fsb_render_args a, int seg_bands, int pf_dist) {
  __shared__ float sq_sm[256]; /* (c/255)^2: the second-stage operands of the three mixes */
  __shared__ uint32_t bias_slot;
  extern __shared__ uint32_t ring_dyn[]; /* FSB_PAINT_WARPS x 4096 bytes of rings + 4096 of alignment slack */
  const float *un = a.lut, *sq = a.lut + 256;
  sq_sm[threadIdx.x] = sq[threadIdx.x];
  sq_sm[threadIdx.x + 128] = sq[threadIdx.x + 128];
  if (threadIdx.x == 0) bias_slot = (uint32_t)__cvta_generic_to_shared(sq_sm) - 0x4B000000u * 4u; /* see fsb_colour.cuh sq_of_bits */
  __syncthreads();
  const uint32_t sq_sm_biased = *reinterpret_cast<volatile uint32_t *>(&bias_slot);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pose = blockIdx.y;
  const int ncols = a.col_end - a.col_begin;
  const int group = blockIdx.x * FSB_PAINT_WARPS + warp;
  if (group * 32 >= ncols) return;
  const int jrel = group * 32 + lane;
  const bool col_ok = jrel < ncols; /* padding lanes of a ragged group: no candidates, no stores */
  const fsb_frame_consts *fcp = a.fc + pose;
  const uint32_t empty = fcp->empty;
  uint32_t cur = fcp->sky; /* running colour of fill_vline: sky until the first non-transparent record (:246-248) */
  const int n = col_ok ? (int)a.cand_cnt[(size_t)pose * a.ncols_pad + jrel] : 0;
  const uint32_t *src = a.cand + cand_group_base(a, pose, group) + lane; /* entry q at src[q * 32] */
  const float4 *line = reinterpret_cast<const float4 *>(a.table + (size_t)pose * a.tab_stride);
  const float fj = (float)(a.col_begin + jrel);

  int b_first = 0, b_end = a.n_bands;
  int p = n - 1; /* next candidate to colour (the list is walked backward: rows increase) */
  if (seg_bands > 0) {
    b_first = (int)blockIdx.z * seg_bands;
    b_end = min(a.n_bands, b_first + seg_bands);
    if (b_first >= b_end) return;
    if (b_first > 0) {
      /* number of candidates with row >= the segment's first row (rows strictly decrease along the list) */
      const uint32_t r_first = (uint32_t)b_first << 5;
      /* four-way search: three independent probes per step (the lists of a large batch are in DRAM, every step of a
       * bisection would wait for one load) */
      int lo = 0, hi = n;
      while (__any_sync(FSB_FULL, lo < hi)) {
        if (lo < hi) {
          const int d = hi - lo;
          const int m1 = lo + (d >> 2), m2 = lo + (d >> 1), m3 = lo + ((3 * d) >> 2);
          const uint32_t w1 = src[(size_t)m1 * 32], w2 = src[(size_t)m2 * 32], w3 = src[(size_t)m3 * 32];
          if ((w1 & FSB_ROW_MASK) < r_first) hi = m1;
          else if ((w2 & FSB_ROW_MASK) < r_first) { lo = m1 + 1; hi = m2; }
          else if ((w3 & FSB_ROW_MASK) < r_first) { lo = m2 + 1; hi = m3; }
          else lo = m3 + 1;
        }
      }
      p = lo - 1;
      /* the running colour that enters the segment: the nearest non-transparent record above it (one look, rarely more) */
      int q = lo;
      bool look = q < n;
      while (__any_sync(FSB_FULL, look)) {
        if (look) {
          const float4 l = __ldg(line + (src[(size_t)q * 32] >> FSB_ROW_BITS));
          const uint32_t c = colour_of<BIL, true>(a, __fadd_rn(l.x, __fmul_rn(fj, l.z)), __fadd_rn(l.y, __fmul_rn(fj, l.w)), un, sq,
                                                  sq_sm_biased);
          const uint32_t col = (c & 0x00FFFFFFu) | ((uint32_t)((int32_t)c >> 31) & 0xFF000000u);
          ++q;
          if (col != empty) {
            cur = col;
            look = false;
          } else look = q < n;
        }
      }
    }
  }

  /* software pipeline of the colour trips: the candidate word of the trip after next and the depth-table entry of the next
   * trip are in flight while this trip's record is filtered (the table address depends on the word) */
  /* wp = address of candidate p, stepped by -128 bytes per trip: the word load and the L2 prefetch address it with
   * immediate offsets (no index arithmetic per load) */
  const uint32_t *wp = src + (ptrdiff_t)p * 32;
  uint32_t word_1 = 0, word_2 = 0;
  float4 l_1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p >= 0) word_1 = src[(size_t)p * 32];
  if (p >= 1) word_2 = src[(size_t)(p - 1) * 32];
  if (p >= 0) l_1 = __ldg(line + (word_1 >> FSB_ROW_BITS));
  /* ring: shared byte addresses of this lane's head and tail slots (slot stride 128 bytes, 4096 bytes per warp: the wrap
   * is (address + 128) & 0xFFF | base), the number of entries, and the rows at which the entries start as bit masks of the
   * band being painted (mask0) and of the next one (mask1) */
  /* aligned by hand in the shared window: __align__ on a static array is relative to the CTA's allocation, whose offset
   * in the window is not a multiple of 4096 (the first kilobyte is reserved) */
  const uint32_t ring_base = (((uint32_t)__cvta_generic_to_shared(ring_dyn) + 4095u) & ~4095u) + (uint32_t)warp * 4096u;
  uint32_t hs = ring_base + (uint32_t)lane * 4u, ts = hs;
  int cnt = 0;
  uint32_t mask0 = 0, mask1 = 0;
  unsigned long long n_trips = 0;

  uint32_t *o = a.out + (size_t)pose * a.pose_stride + (size_t)(b_first << 5) * a.row_stride + jrel;
  const int stride_bytes = (int)a.row_stride * 4; /* 32 rows x stride fits 64 bits through mul.wide */
  const uint32_t row_lim = (uint32_t)b_end << 5;  /* nothing at or past the segment's end is coloured */
  for (int b = b_first; b < b_end; ++b) {
    const uint32_t r_end = (uint32_t)(b + 1) << 5;
    const uint32_t r_ahead = min(r_end + FSB_RUNAHEAD, row_lim);
    /* ---- colour: until no lane has an uncoloured candidate inside this band ---- */
    for (;;) {
      const uint32_t row1 = word_1 & FSB_ROW_MASK;
      const bool have = p >= 0;
      const bool need = have && row1 < r_end;
      if (!__any_sync(FSB_FULL, need)) break;
      /* a lane that needs the trip always has room: its ring holds records of this band only, with rows below row1 */
      const bool can = have && cnt < FSB_RING && row1 < r_ahead;
      ++n_trips;
      if (can) {
        const float4 l = l_1;
        word_1 = word_2;
        --p;
        wp -= 32;
        /* Loads of the next trips, in this order and as volatile asm so that the order survives: the table entry of the next
         * record (its address comes from a word that arrived a trip ago), then the word after it, then the L2 prefetch.
         * (With the word load first, the shift that forms the table address waited for it -- the two loads share a
         * scoreboard -- and exposed the whole L2 latency in every trip: 17 % of the stall samples, ncu r2m.)
         * The prefetch brings the candidate word of FSB_PAINT_PF trips ahead from DRAM to L2 (no register, no scoreboard; the
         * lists of a large batch do not stay in L2 between the march and this kernel, and a register prefetch cannot reach
         * further than one trip).  Unclamped: below the start of a list lies the previous group's list or the pad in front
         * of the buffer (fsb_api.c).
         * Measured and dropped (profiles/r2_paint_variants.txt): this record's gathers issued before these loads and the
         * ring bookkeeping (the texture wait falls from 25 to 16 % of the stall samples, but the split form costs 24
         * instructions more per trip: 1.64 against 1.55 ms), and the gathers issued a whole trip ahead, channel by channel
         * into the registers the record before has just consumed (no texture wait left, +20 % instructions: 1.72 against
         * 1.52 ms; git history of this file, ncu profiles/r2_paint_e_*). */
        if (p >= 0) {
          const float4 *tp = line + (word_1 >> FSB_ROW_BITS);
          asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(l_1.x), "=f"(l_1.y), "=f"(l_1.z), "=f"(l_1.w) : "l"(tp));
        }
        if (p >= 1) asm volatile("ld.global.u32 %0, [%1 + -128];" : "=r"(word_2) : "l"(wp));
        if (pf_dist) asm volatile("prefetch.global.L2 [%0 + %1];" ::"l"(wp), "n"(-128 * FSB_PAINT_PF));
        const float x = __fadd_rn(l.x, __fmul_rn(fj, l.z));
        const float y = __fadd_rn(l.y, __fmul_rn(fj, l.w));
        const uint32_t colour = colour_of<BIL, true>(a, x, y, un, sq, sq_sm_biased);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ts), "r"(colour) : "memory");
        ts = ((ts + 128u) & 0xFFFu) | ring_base;
        ++cnt;
        const uint32_t bit = 1u << (row1 & 31u);
        if (row1 < r_end) mask0 |= bit;
        else mask1 |= bit;
      }
    }
    /* ---- paint the band's rows ---- */
    const int nrows = min(32, a.h - (b << 5));
    const bool full = nrows == 32;
    if (!__any_sync(FSB_FULL, mask0 != 0u)) { /* no record starts in this band in any of the 32 columns */
      if (col_ok) {
        if (full) {
#pragma unroll
          for (int r = 0; r < 32; ++r)
            asm volatile("{\n\t.reg .u64 oa;\n\tmul.wide.s32 oa, %0, %1;\n\tadd.s64 oa, oa, %2;\n\tst.global.u32 [oa], %3;\n\t}" ::"r"(r),
                         "r"(stride_bytes), "l"(o), "r"(cur)
                         : "memory");
        } else {
          for (int r = 0; r < nrows; ++r) o[(size_t)r * a.row_stride] = cur;
        }
      }
    } else if (col_ok) {
      /* One row: m = (a record starts here); if so take its colour unless transparent, step the ring and fetch the entry after
       * it; store the running colour.  Spelled in PTX so that it stays eight instructions. */
      uint32_t e;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(hs) : "memory");
#define FSB_PAINT_ROW(r)                                                                   \
  asm volatile(                                                                            \
      "{\n\t.reg .pred m, c;\n\t.reg .u32 t, hn;\n\t.reg .u64 oa;\n\t"                    \
      "and.b32 t, %3, %4;\n\t"                                                             \
      "setp.ne.u32 m, t, 0;\n\t"                                                           \
      "setp.ne.and.u32 c, %0, %5, m;\n\t"                                                  \
      "@c mov.u32 %1, %0;\n\t"                                                             \
      "add.u32 hn, %2, 128;\n\t"                                                           \
      "@m lop3.b32 %2, hn, 0xFFF, %6, 0xEA;\n\t"                                           \
      "@m ld.shared.u32 %0, [%2];\n\t"                                                     \
      "mul.wide.s32 oa, %7, %8;\n\t"                                                       \
      "add.s64 oa, oa, %9;\n\t"                                                            \
      "st.global.u32 [oa], %1;\n\t}"                                                       \
      : "+r"(e), "+r"(cur), "+r"(hs)                                                       \
      : "r"(mask0), "n"(1u << (r)), "r"(empty), "r"(ring_base), "r"((int)(r)), "r"(stride_bytes), "l"(o) \
      : "memory");
      if (full) {
#define R4(r) FSB_PAINT_ROW(r) FSB_PAINT_ROW(r + 1) FSB_PAINT_ROW(r + 2) FSB_PAINT_ROW(r + 3)
        R4(0) R4(4) R4(8) R4(12) R4(16) R4(20) R4(24) R4(28)
#undef R4
      } else {
        for (int r = 0; r < nrows; ++r) {
          if ((mask0 >> r) & 1u) {
            if (e != empty) cur = e;
            hs = ((hs + 128u) & 0xFFFu) | ring_base;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(hs) : "memory");
          }
          o[(size_t)r * a.row_stride] = cur;
        }
      }
#undef FSB_PAINT_ROW
    } else { /* padding lane of a ragged group: nothing to store (and no candidates: its masks are 0) */
    }
    cnt -= __popc(mask0);
    mask0 = mask1;
    mask1 = 0u;
    o += (size_t)32 * a.row_stride;
  }
  if (a.stats && lane == 0) atomicAdd(a.stats + 2, n_trips); /* colour trips of 32 lanes: utilisation = records / (32 x trips) */
}

extern "C" int fsb_launch_paint(const fsb_render_args *a, int seg_bands, void *stream, int64_t *launches) {
  cudaStream_t s = (cudaStream_t)stream;
  const int groups = a->ncols_pad >> 5;
  const int segs = seg_bands > 0 ? (a->n_bands + seg_bands - 1) / seg_bands : 1;
  dim3 grid((groups + FSB_PAINT_WARPS - 1) / FSB_PAINT_WARPS, a->n_poses, segs);
  /* tuning aids: FSB_PAINT_VARIANT: 1 = 9 CTAs per SM (56 registers), 2 = 10 CTAs per SM (48 registers); default: 8 CTAs per SM
   * (64 registers).  FSB_PAINT_PF=0 switches the L2 prefetch of the candidate words off. */
  static int variant = -1, pf_dist = 1;
  if (variant < 0) {
    const char *e = getenv("FSB_PAINT_VARIANT");
    variant = e ? atoi(e) : 0;
    e = getenv("FSB_PAINT_PF");
    if (e && atoi(e) >= 0) pf_dist = atoi(e);
  }
  /* the rings want the large shared-memory carve-out (8-10 CTAs x 21.5 KB); the attribute is kept per kernel and device */
  int dev = 0;
  cudaGetDevice(&dev);
#define FSB_PAINT_LAUNCH(B, V)                                                                                              \
  do {                                                                                                                      \
    static unsigned long long done = 0;                                                                                     \
    const unsigned long long bit = 1ull << (dev & 63);                                                                      \
    if (!(done & bit)) {                                                                                                    \
      cudaFuncSetAttribute(fsb_paint_kernel<B, V>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
      done |= bit;                                                                                                          \
    }                                                                                                                       \
    fsb_paint_kernel<B, V><<<grid, FSB_PAINT_WARPS * 32, FSB_PAINT_SMEM, s>>>(*a, seg_bands, pf_dist);                     \
  } while (0)
  if (a->filter == FSB_FILTER_BILINEAR) {
    switch (variant) {
      case 2: FSB_PAINT_LAUNCH(true, 2); break;
      case 1: FSB_PAINT_LAUNCH(true, 1); break;
      default: FSB_PAINT_LAUNCH(true, 0); break;
    }
  } else FSB_PAINT_LAUNCH(false, 0);
#undef FSB_PAINT_LAUNCH
  if (launches) ++*launches;
  return (int)cudaGetLastError();
}
