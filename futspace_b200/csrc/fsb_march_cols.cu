/*
 * fsb_march_cols.cu -- the column-parallel march of the texture fast path (sm_100a).
 *
 *   fsb_marchc_kernel   lane = screen column, the warp walks the depth series; per-lane y-buffer in a register;
 *                       emits, per column, the candidate list (row | sample index << 15)          fut/voxel_renderer.fut:215-231
 *   fsb_colour_kernel   one thread per visible record: png_color / png_color_filtered (3 x argb.mix) and the band index
 *                       of the list -> the (row, colour) records fsb_expand*_kernel consumes      fut/render_functions.fut:91-105
 *
 * Why this shape (round 2; the round-1 march, lanes over 32 consecutive depth samples of ONE column, stays in
 * fsb_kernels.cu for the tiled / generic paths and as the A/B reference, FSB_FLAG_MARCH_Z):
 *   - ncu on the round-1 march: 84 warp instructions per 32 samples, 77 % issue-active -- the occlusion scan
 *     (`scan occlude`, :231) cost a REDUX, five SHFL+IMNMX, a ballot and a shared-memory queue per chunk, and every lane
 *     loaded its own depth-table entry.  With lane = column the depth-table entry is warp-uniform (two broadcast loads per
 *     step), `occlude` is one compare against the lane's own running minimum, and a sample that does not lower it
 *     costs nothing more.
 *   - The colour filter (33 divides + 9 square roots per sample in the reference) only runs for samples that are
 *     visible, on full warps, in its own pass: the march keeps no colour state, which is what lets it run at 40+
 *     warps per SM.
 *   - Batches only: a frame offers one warp per 32 columns (60 at 1920), so single frames and small batches stay on the
 *     lanes-over-depth march (one warp per column).  Splitting a column's depth series over several warps with a merge
 *     pass was built and measured in round 2: the candidates of the far segments, the merge and the extra
 *     launches cost more than the split saved.
 *
 * Float discipline as in fsb_kernels.cu: every parity-relevant operation uses the round-to-nearest intrinsics.
 */
#include <stdlib.h>

#include "fsb_colour.cuh"
#include "fsb_device.cuh"

#define FSB_MC_WARPS 4 /* warps (= groups of 32 columns) per march CTA */

/* Per-lane march state: the y-buffer of :231 as a float (exact: rows <= 32768; -inf once it reached row 0 -- y >= 0
 * always, :225, so nothing can pass `y < 0` any more) and the append pointer of the column's candidate list. */
struct col_state {
  float ybuf_f;
  uint32_t *list; /* entry p of this lane's column at list[p * 32] */
  int n;
};

/* Projection and occlusion test of one sample (fut/voxel_renderer.fut:223-225, `occlude` :69-72).
 * With the saturating i32.f32 and an integer y-buffer Y >= 1:  max(0, i32.f32 rel) < Y  <=>  !(rel >= Y)  (a NaN converts
 * to row 0), so the conversion only runs -- behind a warp-uniform branch -- when some column of the group sees the sample.
 * A finished column (-inf) is passed only by a NaN; the exact test in the visible path rejects it. */
template <bool BIL>
__device__ __forceinline__ void cstep_resolve(const col_step<BIL> &t, float iz, float cam_h, float horizon, uint32_t kword,
                                              col_state &st) {
  const float rel = __fadd_rn(__fmul_rn(__fsub_rn(cam_h, cstep_height<BIL>(t)), iz), horizon);
  const bool cand = !(rel >= st.ybuf_f);
  if (__any_sync(FSB_FULL, cand)) {
    const int yy = max(0, __float2int_rz(rel));
    const float yf = (float)yy;
    if (cand && yf < st.ybuf_f) {
      st.list[(size_t)st.n * 32] = (uint32_t)yy | kword;
      ++st.n;
      st.ybuf_f = yy > 0 ? yf : -INFINITY;
    }
  }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

/* Local occlusion bound.  The map-wide bound (fsb_kernels.cu) ends a column once NO terrain could rise above its
 * y-buffer; most of what a low camera does not see is hidden behind a nearby ridge long before that.  Per chunk of 32 depth
 * steps and group of 32 columns the warp asks whether the terrain under THIS block of samples can reach any of its 32
 * y-buffers: lane (s, b) takes the sub-block of steps 4s..4s+3 x columns 8b..8b+7, whose sample positions lie inside the
 * bounding box of its four corner samples (positions are monotone in the column index along a line and move along a ray
 * from the camera with depth); the box, widened by the bilinear footprint and two texels of rounding slack, spans at most
 * 2^L texels a side, and ONE byte of the level-L pyramid (fsb_api.c build_height_pyramid) bounds every texel it can
 * touch.  No sample of the sub-block is higher than that + 0.5 (see the map-wide bound), so none projects above
 * row(bound) -- every operation of :223-225 is monotone, and inv_z decreases along the series (bound at the sub-block's last
 * step for a camera above the local maximum, at its first step below it).  The chunk is skipped when every column's
 * y-buffer is at or above the lowest such row of its eight columns: nothing in it can pass `occlude` (strict <), the
 * frame is unchanged.  About 70 instructions against the 1250 of an evaluated chunk. */
__device__ __forceinline__ bool chunk_hidden(const fsb_render_args &a, const float4 *tl, const float *tz, float fj0, int lane,
                                             float cam_h, float horizon, float ybuf_f) {
  const int s4 = (lane >> 2) * 4;
  const float4 l0 = tl[s4], l1 = tl[s4 + 3];
  const float ja = fj0 + (float)((lane & 3) * 8), jb = ja + 7.0f;
  const float xa0 = l0.x + ja * l0.z, xb0 = l0.x + jb * l0.z, xa1 = l1.x + ja * l1.z, xb1 = l1.x + jb * l1.z;
  const float ya0 = l0.y + ja * l0.w, yb0 = l0.y + jb * l0.w, ya1 = l1.y + ja * l1.w, yb1 = l1.y + jb * l1.w;
  const float mag = fabsf(xa0) + fabsf(xb0) + fabsf(xa1) + fabsf(xb1) + fabsf(ya0) + fabsf(yb0) + fabsf(ya1) + fabsf(yb1);
  int bound = -1; /* hides nothing (a finished column's y-buffer is -inf) */
  if (mag < 3.2e7f) { /* finite and within the texture path's coordinate range (a NaN fails the test) */
    const int x0 = __float2int_rd(fminf(fminf(xa0, xb0), fminf(xa1, xb1))) - 2;
    const int x1 = __float2int_rd(fmaxf(fmaxf(xa0, xb0), fmaxf(xa1, xb1))) + 3;
    const int y0 = __float2int_rd(fminf(fminf(ya0, yb0), fminf(ya1, yb1))) - 2;
    const int y1 = __float2int_rd(fmaxf(fmaxf(ya0, yb0), fmaxf(ya1, yb1))) + 3;
    const int ext = max(x1 - x0, y1 - y0) + 1; /* texels a side, >= 6 */
    const int L = 32 - __clz(ext - 1);         /* 2^L >= ext */
    if (L <= a.pyr_levels) {
      const int rl = a.r >> L, ql = a.q >> L;
      const uint32_t qr = (uint32_t)a.q * (uint32_t)a.r;
      const uint32_t off = (qr - (qr >> (2 * L - 2))) / 3u + (uint32_t)((y0 >> L) & (ql - 1)) * (uint32_t)rl +
                           (uint32_t)((x0 >> L) & (rl - 1));
      const float hb = __fadd_ru((float)__ldg(a.hpyr + off), 0.501f);
      const float d = __fsub_rd(cam_h, hb);
      const float iz = d >= 0.0f ? tz[s4 + 3] : tz[s4];
      bound = max(0, __float2int_rz(__fadd_rn(__fmul_rn(d, iz), horizon)));
    }
  }
  bound = min(bound, __shfl_xor_sync(FSB_FULL, bound, 4));
  bound = min(bound, __shfl_xor_sync(FSB_FULL, bound, 8));
  bound = min(bound, __shfl_xor_sync(FSB_FULL, bound, 16));
  const int mine = __shfl_sync(FSB_FULL, bound, lane >> 3); /* lanes 0..3 hold the bounds of columns 0-7, 8-15, 16-23, 24-31 */
  return __all_sync(FSB_FULL, ybuf_f <= (float)mine);
}

/* U consecutive inv_z of the staged chunk with the widest shared-memory loads (p is U * 4-byte aligned) */
template <int U>
__device__ __forceinline__ void load_inv_z(const float *p, float (&iz)[U]) {
  if (U % 4 == 0) {
#pragma unroll
    for (int v = 0; v < U / 4; ++v) {
      const float4 t = reinterpret_cast<const float4 *>(p)[v];
      iz[4 * v] = t.x; iz[4 * v + 1] = t.y; iz[4 * v + 2] = t.z; iz[4 * v + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int v = 0; v < U / 2; ++v) {
      const float2 t = reinterpret_cast<const float2 *>(p)[v];
      iz[2 * v] = t.x; iz[2 * v + 1] = t.y;
    }
  }
}

/* U = depth steps per register set; two sets are in flight (the gathers of one are issued before the other is resolved).
 *
 * Depth table: the 640 bytes of a chunk of 32 steps ({sx,sy,dx,dy} x 32, inv_z x 32) are the same for every column, so
 * the CTA (4 warps = 128 adjacent columns of one pose) stages them in shared memory with cp.async, two
 * chunks ahead of the one being marched: the per-step operands are LDS broadcasts (29 cycles) instead of global loads
 * queued behind the texture gathers in the same L1 pipe (round-2 ncu: the largest stall site of the first version). */
template <bool BIL, int U, int MINB, bool LC, bool PW>
__global__ void __launch_bounds__(FSB_MC_WARPS * 32, MINB) fsb_marchc_kernel(const fsb_render_args a) {
  static_assert(32 % (2 * U) == 0, "a pair of register sets must tile a 32-step chunk");
  /* PW: every warp stages the table chunks for itself (no CTA barrier: with the local occlusion bound the four groups of a
   * CTA skip different chunks, and the ones that skip would wait at the barrier for the one that marches) */
  __shared__ __align__(16) float sm_all[PW ? FSB_MC_WARPS : 1][3][FSB_TAB_BLOCK];
  float (*sm)[FSB_TAB_BLOCK] = sm_all[PW ? (threadIdx.x >> 5) : 0];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pose = blockIdx.y;
  const int ncols = a.col_end - a.col_begin;
  const int group = blockIdx.x * FSB_MC_WARPS + warp;
  const int jrel = group * 32 + lane;
  /* A warp with no column stays in the loop for the barriers.  Lanes past the last column of a ragged group march a
   * column of their own into the padding of the scratch lists (ncols_pad): no masking in the loop, nobody reads them.
   * (`active` only ever holds vote results, so the compiler knows it is warp-uniform.) */
  bool active = __any_sync(FSB_FULL, group * 32 < ncols);
  const fsb_frame_consts *fcp = a.fc + pose;
  const float cam_h = fcp->cam_h, horizon = fcp->horizon, cull_d = fcp->cull_d;
  const int n_chunks = (fcp->n_z + 31) >> 5;
  const int kcap = a.tab_stride / 5;
  const float4 *line = reinterpret_cast<const float4 *>(a.table + (size_t)pose * a.tab_stride); /* {sx,sy,dx,dy}[k] */
  const float *invz = a.table + (size_t)pose * a.tab_stride + 4 * (size_t)kcap;                 /* inv_z[k] (:217)  */
  col_state st;
  st.ybuf_f = (float)a.h;
  st.list = a.cand + (active ? cand_group_base(a, pose, group) : 0) + lane;
  st.n = 0;
  const float fj = (float)(a.col_begin + jrel);

  int c_first = 0;
  const int c_end = n_chunks; /* the depth series in chunks of 32 samples */
  /* Occlusion bound (see fsb_kernels.cu): camera above the highest terrain -> a prefix of the series projects below the
   * bottom row and is skipped (lane = chunk, bound at the chunk's last sample) ... */
  if (cull_d >= 0.0f && cull_d < INFINITY) {
    int first = c_end;
    for (int base = c_first; base < c_end; base += 32) {
      const int ci = min(base + lane, c_end - 1);
      const float izl = __ldg(invz + ci * 32 + 31);
      const bool below = max(0, __float2int_rz(__fadd_rn(__fmul_rn(cull_d, izl), horizon))) >= a.h;
      const unsigned alive = __ballot_sync(FSB_FULL, !below);
      if (alive) {
        first = base + __ffs(alive) - 1;
        break;
      }
    }
    c_first = first;
  }
  /* ... camera below it -> the bound grows with depth: a column whose y-buffer it has reached is finished */
  const bool can_stop = cull_d < 0.0f && cull_d > -INFINITY;
  int c_done = 0;
  if (c_first < c_end) {
    /* staging: threads 0..31 copy one {sx,sy,dx,dy} each, threads 32..39 four inv_z each; chunks c_first and c_first + 1
     * first (the table is padded with entries that repeat the last sample) */
    const bool stager = PW ? true : tid < FSB_TAB_BLOCK / 4;
    const char *gsrc = (PW || tid < 32) ? reinterpret_cast<const char *>(line + c_first * 32 + (PW ? lane : tid))
                                        : reinterpret_cast<const char *>(invz + c_first * 32 + (tid - 32) * 4);
    const int gstep = (PW || tid < 32) ? 32 * 16 : 32 * 4; /* bytes per chunk in the source array */
    /* PW: lanes 0..7 copy the chunk's 32 inv_z as well */
    const char *gsrc2 = reinterpret_cast<const char *>(invz + c_first * 32 + (lane & 7) * 4);
    const bool stager2 = PW && lane < 8;
    if (stager) {
      cp_async16(&sm[0][(PW ? lane : tid) * 4], gsrc);
      cp_async16(&sm[1][(PW ? lane : tid) * 4], gsrc + gstep);
    }
    if (stager2) {
      cp_async16(&sm[0][128 + lane * 4], gsrc2);
      cp_async16(&sm[1][128 + lane * 4], gsrc2 + 128);
    }
    gsrc += 2 * gstep;
    gsrc2 += 2 * 128;
    cp_async_wait_all();
    if (PW) __syncwarp();
    else __syncthreads();
    col_step<BIL> sa[U], sb[U];
    /* have_sa: sa holds the gathers of the head of the chunk about to be marched (issued at the tail of the chunk before it) */
    bool have_sa = false;
    /* LC: the launch has the pyramid of local height maxima (a.hpyr); the bound needs what the map-wide one needs */
    const bool local_cull = LC && cull_d > -INFINITY && cull_d < INFINITY && !a.full_eval;
    const float fj0 = (float)(a.col_begin + group * 32);
    int slot = 0;
    for (int c = c_first; c < c_end; ++c) {
      const int slot1 = slot == 2 ? 0 : slot + 1, slot2 = slot1 == 2 ? 0 : slot1 + 1;
      /* chunk c + 2 -> the slot chunk c - 1 was read from (every warp has passed the barrier that ended it) */
      if (stager) cp_async16(&sm[slot2][(PW ? lane : tid) * 4], gsrc);
      gsrc += gstep;
      if (stager2) cp_async16(&sm[slot2][128 + lane * 4], gsrc2);
      gsrc2 += 128;
      const float4 *tl = reinterpret_cast<const float4 *>(sm[slot]);
      const float *tz = sm[slot] + 128;
      if (active && !a.full_eval) {
        float bound = 0.0f;
        if (can_stop) bound = (float)max(0, __float2int_rz(__fadd_rn(__fmul_rn(cull_d, tz[0]), horizon)));
        active = __any_sync(FSB_FULL, st.ybuf_f > bound); /* false: every column of the group is finished, for good */
      }
      bool eval = active;
      if (active && !have_sa) { /* first chunk, or the chunk before this one was skipped */
        if (local_cull && chunk_hidden(a, tl, tz, fj0, lane, cam_h, horizon, st.ybuf_f)) eval = false;
        else {
#pragma unroll
          for (int u = 0; u < U; ++u) cstep_issue<BIL>(sa[u], a, tl[u], fj);
        }
      }
      have_sa = !LC;
      if (eval) {
        ++c_done;
        const float4 *tl_next = reinterpret_cast<const float4 *>(sm[slot1]);
        const uint32_t kw = (uint32_t)(c << 5) << FSB_ROW_BITS;
#pragma unroll
        for (int i = 0; i < 32; i += 2 * U) {
#pragma unroll
          for (int u = 0; u < U; ++u) cstep_issue<BIL>(sb[u], a, tl[i + U + u], fj);
          float iza[U], izb[U];
          load_inv_z<U>(tz + i, iza);
          load_inv_z<U>(tz + i + U, izb);
#pragma unroll
          for (int u = 0; u < U; ++u)
            cstep_resolve<BIL>(sa[u], iza[u], cam_h, horizon, kw + ((uint32_t)(i + u) << FSB_ROW_BITS), st);
          /* next set: the following steps of this chunk; without the local bound also the head of the next chunk (a
           * repeated last sample in the padding projects to the same row and `occlude` keeps the earlier one).  With it,
           * the next chunk is tested first, at the top of its turn, with nothing in flight (measured: testing here, with
           * a register set live, spills and is 8 % slower, profiles/r2_local_cull_ab.jsonl) */
          if (i + 2 * U < 32 || !LC) {
#pragma unroll
            for (int u = 0; u < U; ++u) cstep_issue<BIL>(sa[u], a, i + 2 * U < 32 ? tl[i + 2 * U + u] : tl_next[u], fj);
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
            cstep_resolve<BIL>(sb[u], izb[u], cam_h, horizon, kw + ((uint32_t)(i + U + u) << FSB_ROW_BITS), st);
        }
      }
      cp_async_wait_all();
      if (PW) {
        __syncwarp();
        if (!active) break;
      } else if (!__syncthreads_or(active)) break; /* chunk c + 2 visible; everyone is done with chunk c */
      slot = slot1;
    }
  }
  if (group * 32 < ncols) {
    a.cand_cnt[(size_t)pose * a.ncols_pad + jrel] = (uint32_t)st.n;
    if (a.stats) {
      /* in chunks of 32 samples of one column, like the lanes-over-depth march counts them */
      if (lane == 0) atomicAdd(a.stats, (unsigned long long)c_done * (unsigned long long)min(32, ncols - group * 32));
      unsigned long long tot = (jrel < ncols) ? (unsigned long long)st.n : 0ull;
#pragma unroll
      for (int d = 16; d; d >>= 1) tot += __shfl_xor_sync(FSB_FULL, tot, d);
      if (lane == 0) atomicAdd(a.stats + 1, tot);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Colour: one warp per (pose, group of 32 columns[, segment]), lane = column, one record per lane and trip.  Reads the
 * sample index, rebuilds the sample position (get_segment, :63-66), runs png_color / png_color_filtered and writes the
 * record the expand kernels consume, plus the per-band index of the list (sidx[b] = number of records with
 * row >= b * 32; rows strictly decrease along the list).  The filter itself is in fsb_colour.cuh (shared with fsb_paint.cu). */
/* slice_len > 0: blockIdx.y selects a slice of the record index range (medium batches: more, shorter warps); the last
 * slice runs to the end of the list.  slice_len == 0: the whole list. */
template <bool BIL, bool REC4>
__global__ void __launch_bounds__(128) fsb_colour_kernel(const fsb_render_args a, int slice_len) {
  __shared__ float sq_sm[256]; /* (c/255)^2: the second-stage operands of the three mixes */
  const float *un = a.lut, *sq = a.lut + 256;
  __shared__ uint32_t bias_slot;
  pdl_trigger(); /* single frames (programmatic dependent launch): the expand kernel may be scheduled */
  sq_sm[threadIdx.x] = sq[threadIdx.x];
  sq_sm[threadIdx.x + 128] = sq[threadIdx.x + 128];
  /* table address minus 0x4B000000 * 4 (see sq_of_bits), passed through shared memory so that ptxas keeps it in one
   * register instead of re-deriving the difference at each of the six look-ups of a record */
  if (threadIdx.x == 0) bias_slot = (uint32_t)__cvta_generic_to_shared(sq_sm) - 0x4B000000u * 4u;
  __syncthreads();
  pdl_wait(); /* candidate lists, counts and the depth table come from the march */
  const uint32_t sq_sm_biased = *reinterpret_cast<volatile uint32_t *>(&bias_slot);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pose = blockIdx.z;
  const int ncols = a.col_end - a.col_begin;
  const int group = blockIdx.x * 4 + warp;
  if (group * 32 >= ncols) return;
  const int jrel = group * 32 + lane;
  int n = jrel < ncols ? (int)a.cand_cnt[(size_t)pose * a.ncols_pad + jrel] : 0; /* padding lanes of a ragged group: none */
  int p = 0, p_end = n;
  bool closes = jrel < ncols; /* this warp handles the column's last record (or its empty list): it closes the band index */
  if (slice_len > 0) {
    p = (int)blockIdx.y * slice_len;
    if (blockIdx.y + 1 < gridDim.y) p_end = min(n, p + slice_len);
    closes = closes && (int)blockIdx.y == (n == 0 ? 0 : min((n - 1) / slice_len, (int)gridDim.y - 1));
  }
  if (!__any_sync(FSB_FULL, p < p_end || closes)) return;
  const uint32_t *src = a.cand + cand_group_base(a, pose, group) + lane; /* entry q at src[q * 32] */
  const size_t gid = (size_t)pose * (a.ncols_pad >> 5) + group;
  uint32_t *sidx = a.sidx + gid * (a.n_bands + 1) * 32 + lane;                 /* sidx[b] at sidx[b * 32]         */
  const size_t rec0 = (gid * a.rec_cap + 1) * 32 + lane;                        /* record q at rec[q * 32]         */
  uint32_t *rec4 = reinterpret_cast<uint32_t *>(a.recs) + rec0;
  uint2 *rec8 = a.recs + rec0;
  if (blockIdx.y == 0) { /* slot 0: the guard record the 8-byte walks stop at (see fsb_kernels.cu) */
    if (REC4) rec4[-32] = 0u;
    else rec8[-32] = make_uint2(0xffffffffu, 0u);
  }
  const float4 *line = reinterpret_cast<const float4 *>(a.table + (size_t)pose * a.tab_stride);
  const float fj = (float)(a.col_begin + jrel);
  /* band of the record before the slice (n_bands before the first record of the list) */
  int pb = a.n_bands;
  if (p > 0 && p < p_end) pb = (int)((src[(size_t)(p - 1) * 32] & FSB_ROW_MASK) >> a.rb_shift);
  /* Software pipeline: the candidate word two trips ahead and the depth-table entry one trip ahead are in flight while
   * this trip's record is filtered (the table address depends on the word). */
  uint32_t word_1 = 0, word_2 = 0;
  float4 l_1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p < p_end) word_1 = src[(size_t)p * 32];
  if (p + 1 < p_end) word_2 = src[(size_t)(p + 1) * 32];
  if (p < p_end) l_1 = __ldg(line + (word_1 >> FSB_ROW_BITS));
  for (; __any_sync(FSB_FULL, p < p_end); ++p) {
    const bool valid = p < p_end;
    const uint32_t word = word_1;
    const float4 l = l_1;
    word_1 = word_2;
    if (p + 2 < p_end) word_2 = src[(size_t)(p + 2) * 32];
    if (p + 1 < p_end) l_1 = __ldg(line + (word_1 >> FSB_ROW_BITS));
    if (valid) {
      const uint32_t row = word & FSB_ROW_MASK;
      const float x = __fadd_rn(l.x, __fmul_rn(fj, l.z));
      const float y = __fadd_rn(l.y, __fmul_rn(fj, l.w));
      const uint32_t colour = colour_of<BIL, REC4>(a, x, y, un, sq, sq_sm_biased);
      if (REC4) rec4[(size_t)p * 32] = (colour & 0x80FFFFFFu) | ((row & 31u) << 24);
      else rec8[(size_t)p * 32] = make_uint2(a.smooth ? word : row, colour);
      /* band index: the record that opens a new band writes its list position for that band (one predicated store in
       * the common case) and, rarely, for the bands it skipped */
      const int band = (int)(row >> a.rb_shift);
      if (band < pb) {
        sidx[(band + 1) * 32] = (uint32_t)p;
        if (pb - band > 1) {
#pragma unroll 1
          for (int b = band + 2; b <= pb; ++b) sidx[b * 32] = (uint32_t)p;
        }
        pb = band;
      }
    }
  }
  /* bands at or above the last record: every record of the list has a row >= theirs */
  if (closes) {
#pragma unroll 1
    for (int b = pb; b >= 0; --b) sidx[b * 32] = (uint32_t)n;
  }
}

/* ------------------------------------------------------------------------------------------ */
template <bool BIL, int U, int MINB>
static int launch_marchc_t(const fsb_render_args &a, cudaStream_t s) {
  const int ncols = a.col_end - a.col_begin;
  dim3 grid((ncols + FSB_MC_WARPS * 32 - 1) / (FSB_MC_WARPS * 32), a.n_poses);
  static int per_warp = -1; /* tuning aid: FSB_MARCHC_PW=0 -> the CTA stages the table chunks once for its four warps */
  if (per_warp < 0) {
    const char *e = getenv("FSB_MARCHC_PW");
    per_warp = e ? atoi(e) != 0 : 1;
  }
  if (a.hpyr && per_warp) fsb_marchc_kernel<BIL, U, MINB, true, true><<<grid, FSB_MC_WARPS * 32, 0, s>>>(a);
  else if (a.hpyr) fsb_marchc_kernel<BIL, U, MINB, true, false><<<grid, FSB_MC_WARPS * 32, 0, s>>>(a);
  else fsb_marchc_kernel<BIL, U, MINB, false, false><<<grid, FSB_MC_WARPS * 32, 0, s>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int fsb_launch_march_cols(const fsb_render_args *a, void *stream, int64_t *launches) {
  cudaStream_t s = (cudaStream_t)stream;
  const bool bil = a->filter == FSB_FILTER_BILINEAR;
  /* tuning aid: FSB_MARCHC_VARIANT = steps per register set / CTAs per SM of the bilinear march */
  static int variant = -1;
  if (variant < 0) {
    const char *e = getenv("FSB_MARCHC_VARIANT");
    variant = e ? atoi(e) : 0;
  }
  int rc;
  if (bil && variant == 1) rc = launch_marchc_t<true, 4, 5>(*a, s);
  else if (bil && variant == 2) rc = launch_marchc_t<true, 2, 8>(*a, s);
  else if (bil && variant == 3) rc = launch_marchc_t<true, 2, 10>(*a, s);
  else if (bil && variant == 4) rc = launch_marchc_t<true, 8, 4>(*a, s);
  else
    rc = bil ? launch_marchc_t<true, 4, 6>(*a, s) : launch_marchc_t<false, 4, 6>(*a, s);
  if (launches) ++*launches;
  return rc;
}

/* slice_len: records per colour warp (0: a whole list), see fsb_colour_kernel */
extern "C" int fsb_launch_colour(const fsb_render_args *a, int slice_len, void *stream, int64_t *launches) {
  cudaStream_t s = (cudaStream_t)stream;
  const int groups = a->ncols_pad >> 5;
  const int slices = slice_len > 0 ? (a->cand_cap + slice_len - 1) / slice_len : 1;
  dim3 grid((groups + 3) / 4, slices, a->n_poses);
  const bool pdl = a->pdl != 0;
  int rc;
  if (a->filter == FSB_FILTER_BILINEAR) {
    if (a->rec4) rc = (int)fsb_launch_pdl(fsb_colour_kernel<true, true>, grid, dim3(128), s, pdl, *a, slice_len);
    else rc = (int)fsb_launch_pdl(fsb_colour_kernel<true, false>, grid, dim3(128), s, pdl, *a, slice_len);
  } else {
    if (a->rec4) rc = (int)fsb_launch_pdl(fsb_colour_kernel<false, true>, grid, dim3(128), s, pdl, *a, slice_len);
    else rc = (int)fsb_launch_pdl(fsb_colour_kernel<false, false>, grid, dim3(128), s, pdl, *a, slice_len);
  }
  if (launches) ++*launches;
  return rc;
}
