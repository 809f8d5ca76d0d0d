/*
 * fsb_march_frame.cu -- the march of single frames and small batches on the texture path (sm_100a).
 *
 *   fsb_march4_kernel   one CTA per screen column, lanes over 32 consecutive depth samples as in fsb_march_kernel,
 *                       but the column's chunks of 32 samples go round-robin to the CTA's four warps; the running
 *                       minimum of `scan occlude` (fut/voxel_renderer.fut:231, :69-72) is carried from chunk to chunk
 *                       through shared memory.
 *
 * Why: a lone frame offers one warp per column to fsb_march_kernel (1920 warps, 13 per SM), and each of them walks its
 * 40-60 chunks one after the other -- ncu (profiles/r2_marchz_1080p_single_ncu.txt): 17 % issue-active, 9 warps per SM,
 * 38 us.  `occlude` is associative, so the chunks of a round can be sampled and projected by four warps at once; what stays
 * sequential is a four-element minimum per round.  The records come out in the same order and with the same contents
 * as fsb_march_kernel's: the expand kernels and the frames are unchanged.
 *
 * Saturating i32.f32 only (the texture path); float discipline as in fsb_kernels.cu.
 */
#include "fsb_device.cuh"


/* queue words per entry: bilinear {x, y, row word, list position}; nearest {packed texel, row word, list position} */
template <bool BIL>
struct m4_queue_words {
  static const int value = BIL ? 4 : 3;
};

/* colour filter for `count` queued samples (png_color / png_color_filtered, fut/render_functions.fut:91-105) and their
 * records; every entry carries the list position the round assigned to it */
template <bool BIL>
__device__ __forceinline__ void m4_drain(const fsb_render_args &a, const uint32_t *q, int count, int lane, int &qhead, int &qn,
                                         uint2 *__restrict__ rec, const float *un, const float *sq) {
  constexpr int NQ = m4_queue_words<BIL>::value;
  const int slot = (qhead + lane) & (FSB_QCAP - 1);
  if (lane < count) {
    uint32_t colour;
    if (BIL) {
      const float x = __uint_as_float(q[slot]), y = __uint_as_float(q[FSB_QCAP + slot]);
      colour = sample_color<MEM_TEX, true, FSB_F2I_SATURATE>(a, x, y, un, sq);
    } else { /* nearest, packed: the texel is the colour */
      colour = (q[slot] & 0x00FFFFFFu) | a.alpha_bits;
    }
    const uint32_t row = q[(NQ - 2) * FSB_QCAP + slot], pos = q[(NQ - 1) * FSB_QCAP + slot];
    if (a.rec4) reinterpret_cast<uint32_t *>(rec)[pos] = (colour & 0x80FFFFFFu) | ((row & 31u) << 24);
    else rec[pos] = make_uint2(row, colour);
  }
  qhead = (qhead + count) & (FSB_QCAP - 1);
  qn -= count;
}

/* M4_WARPS: warps per column.  Four give the shortest column; three let all 1920 columns of a 1080p frame be resident at once
 * (14 CTAs of 96 threads per SM at 48 registers: 2072 slots, against 1480 slots = 1.3 waves with four). */
template <bool BIL, int M4_WARPS>
__global__ void __launch_bounds__(M4_WARPS * 32) fsb_march4_kernel(const fsb_render_args a) {
  constexpr int NQ = m4_queue_words<BIL>::value;
  __shared__ uint32_t queues[M4_WARPS][NQ * FSB_QCAP];
  __shared__ int sh_min[M4_WARPS], sh_cnt[M4_WARPS];
  const float *un = a.lut, *sq = a.lut + 256;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pose = blockIdx.y, jrel = blockIdx.x;
  pdl_trigger();
  pdl_wait(); /* depth table and pose constants come from the set-up kernel */
  const int ncols = a.col_end - a.col_begin;
  const fsb_frame_consts *fcp = a.fc + pose;
  const float cam_h = fcp->cam_h, horizon = fcp->horizon, cull_d = fcp->cull_d;
  const int n_chunks = (fcp->n_z + 31) >> 5;
  const float4 *line = reinterpret_cast<const float4 *>(a.table + (size_t)pose * a.tab_stride); /* {sx,sy,dx,dy}[k] */
  const float *invz = a.table + (size_t)pose * a.tab_stride + 4 * (size_t)(a.tab_stride / 5);   /* inv_z[k] (:217)  */
  const size_t colid = (size_t)pose * ncols + jrel;
  /* slot 0 of a column's list: the guard record (see fsb_march_kernel) */
  uint2 *rec = a.rec4 ? reinterpret_cast<uint2 *>(reinterpret_cast<uint32_t *>(a.recs) + colid * a.rec_cap + 1)
                      : a.recs + colid * a.rec_cap + 1;
  if (tid == 0) {
    if (a.rec4) reinterpret_cast<uint32_t *>(rec)[-1] = 0u;
    else rec[-1] = make_uint2(0xffffffffu, 0u);
  }
  uint32_t *sidx = a.sidx + colid * (a.n_bands + 1);
  uint32_t *q = queues[warp];
  const float fj = (float)(a.col_begin + jrel);

  /* occlusion bound: see fsb_march_kernel.  Every warp derives the same first chunk. */
  int c_first = 0;
  if (cull_d >= 0.0f && cull_d < INFINITY) {
    for (int base = 0; base < n_chunks; base += 32) {
      const int ci = min(base + lane, n_chunks - 1);
      const float izl = __ldg(invz + ci * 32 + 31);
      const bool below = max(0, __float2int_rz(__fadd_rn(__fmul_rn(cull_d, izl), horizon))) >= a.h;
      const unsigned alive = __ballot_sync(FSB_FULL, !below);
      if (alive) {
        c_first = base + __ffs(alive) - 1;
        break;
      }
      c_first = min(base + 32, n_chunks);
    }
  }
  const bool can_stop = cull_d < 0.0f && cull_d > -INFINITY;

  /* CTA-uniform state, kept identically by every thread */
  int ybuf = a.h; /* running minimum of projected rows = y-buffer; starts at h, the neutral (0,h) of :231 */
  int nrec = 0;   /* records of the column so far */
  int rounds = 0;
  int qhead = 0, qn = 0;
  height_taps<MEM_TEX, BIL, FSB_F2I_SATURATE> cur, nxt;
  if (c_first + warp < n_chunks) {
    const int k = (c_first + warp) * 32 + lane;
    cur.issue(a, __ldg(line + k), __ldg(invz + k), fj);
  }
  /* the depth-table entry of this lane's sample in the NEXT round's chunk, loaded a round before the gathers that need it
   * (loaded and used in the same round, it was 15 % of the stall samples of a lone frame: profiles/r2_march4_1080p_single_ncu.txt) */
  float4 l_n = make_float4(0.f, 0.f, 0.f, 0.f);
  float iz_n = 0.f;
  if (c_first + warp + M4_WARPS < n_chunks) {
    const int k = (c_first + warp + M4_WARPS) * 32 + lane;
    l_n = __ldg(line + k);
    iz_n = __ldg(invz + k);
  }
  /* inv_z of the first sample of the round, loaded a round ahead as well (the map-wide bound below tests it first thing) */
  float iz_round = c_first < n_chunks ? __ldg(invz + c_first * 32) : 0.f;
  for (int c = c_first; c < n_chunks; c += M4_WARPS) {
    /* camera below the highest terrain: the bound grows with depth; once it has reached the y-buffer at the first
     * sample of a round, nothing from there on can be visible */
    if (can_stop && max(0, __float2int_rz(__fadd_rn(__fmul_rn(cull_d, iz_round), horizon))) >= ybuf) break;
    if (c + M4_WARPS < n_chunks) iz_round = __ldg(invz + (c + M4_WARPS) * 32);
    ++rounds;
    const int cc = c + warp;
    const bool have = cc < n_chunks;
    if (cc + M4_WARPS < n_chunks) nxt.issue(a, l_n, iz_n, fj); /* next round's gathers fly while this round is resolved */
    if (cc + 2 * M4_WARPS < n_chunks) {
      const int k = (cc + 2 * M4_WARPS) * 32 + lane;
      l_n = __ldg(line + k);
      iz_n = __ldg(invz + k);
    }
    /* Lanes past n_z read table padding that repeats the last depth sample: a repeated sample projects to the same
     * row and `occlude` (:70) keeps the earlier one, so no masking is needed. */
    int yy = INT_MAX;
    if (have) {
      const float rel = __fadd_rn(__fmul_rn(__fsub_rn(cam_h, cur.finish()), cur.iz), horizon); /* :223-224 */
      yy = max(0, __float2int_rz(rel));                                                          /* :225 */
    }
    const int m = __reduce_min_sync(FSB_FULL, yy);
    if (lane == 0) sh_min[warp] = m;
    __syncthreads();
    int carry = ybuf, lowest = ybuf; /* running minimum entering this warp's chunk / leaving the round */
#pragma unroll
    for (int s = 0; s < M4_WARPS; ++s) {
      const int v = sh_min[s];
      if (s < warp) carry = min(carry, v);
      lowest = min(lowest, v);
    }
    unsigned mask = 0;
    bool vis = false;
    int excl = carry;
    if (m < carry) { /* warp-uniform: the chunk lowers the y-buffer */
      int incl = yy;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) incl = min(incl, __shfl_up_sync(FSB_FULL, incl, d)); /* lanes < d get their own value back */
      const int e = __shfl_up_sync(FSB_FULL, incl, 1);
      excl = lane == 0 ? carry : min(e, carry);
      vis = yy < excl; /* strict: `occlude` keeps the earlier sample on ties, :70 */
      mask = __ballot_sync(FSB_FULL, vis);
    }
    const int cnt = __popc(mask);
    if (lane == 0) sh_cnt[warp] = cnt;
    __syncthreads();
    int pos = nrec, total = 0;
#pragma unroll
    for (int s = 0; s < M4_WARPS; ++s) {
      const int v = sh_cnt[s];
      if (s < warp) pos += v;
      total += v;
    }
    if (vis) {
      pos += __popc(mask & ((1u << lane) - 1u));
      const int slot = (qhead + qn + __popc(mask & ((1u << lane) - 1u))) & (FSB_QCAP - 1);
      const uint32_t roww = (uint32_t)yy | (a.smooth ? (uint32_t)(cc * 32 + lane) << FSB_ROW_BITS : 0u); /* smoothing needs the sample index */
      if (BIL) {
        q[slot] = __float_as_uint(cur.x);
        q[FSB_QCAP + slot] = __float_as_uint(cur.y);
      } else {
        q[slot] = cur.t00;
      }
      q[(NQ - 2) * FSB_QCAP + slot] = roww;
      q[(NQ - 1) * FSB_QCAP + slot] = (uint32_t)pos;
      /* band index, sidx[b] = number of records with row >= b * 32: the sample before this one in the list is the one
       * that set the running minimum `excl` (none: excl == h), so every band between the two rows starts here */
      const int band = yy >> a.rb_shift, pb = excl >= a.h ? a.n_bands : excl >> a.rb_shift;
      for (int b = band + 1; b <= pb; ++b) sidx[b] = (uint32_t)pos;
    }
    qn += cnt;
    __syncwarp();
    if (qn >= 32) {
      m4_drain<BIL>(a, q, 32, lane, qhead, qn, rec, un, sq);
      __syncwarp();
    }
    nrec += total;
    ybuf = lowest;
    cur = nxt;
    if (ybuf == 0 && !a.full_eval) break; /* y >= 0 always (:225): nothing can pass `yy < 0` any more */
  }
  if (qn > 0) m4_drain<BIL>(a, q, qn, lane, qhead, qn, rec, un, sq);
  /* bands at or above the last record (its row is the final running minimum): every record has a row >= theirs */
  if (warp == 0) {
    const int last_band = nrec ? (ybuf >> a.rb_shift) : a.n_bands;
    for (int b = lane; b <= last_band; b += 32) sidx[b] = (uint32_t)nrec;
  }
  if (a.stats && tid == 0) {
    atomicAdd(a.stats, (unsigned long long)min(rounds * M4_WARPS, n_chunks - c_first));
    atomicAdd(a.stats + 1, (unsigned long long)nrec);
  }
}

extern "C" int fsb_launch_march_frame(const fsb_render_args *a, int warps_per_column, void *stream, int64_t *launches) {
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(a->col_end - a->col_begin, a->n_poses);
  if (launches) ++*launches;
  const bool pdl = a->pdl != 0;
  if (warps_per_column == 8) {
    if (a->filter == FSB_FILTER_BILINEAR) return (int)fsb_launch_pdl(fsb_march4_kernel<true, 8>, grid, dim3(256), s, pdl, *a);
    return (int)fsb_launch_pdl(fsb_march4_kernel<false, 8>, grid, dim3(256), s, pdl, *a);
  }
  if (warps_per_column == 6) {
    if (a->filter == FSB_FILTER_BILINEAR) return (int)fsb_launch_pdl(fsb_march4_kernel<true, 6>, grid, dim3(192), s, pdl, *a);
    return (int)fsb_launch_pdl(fsb_march4_kernel<false, 6>, grid, dim3(192), s, pdl, *a);
  }
  if (warps_per_column == 3) {
    if (a->filter == FSB_FILTER_BILINEAR) return (int)fsb_launch_pdl(fsb_march4_kernel<true, 3>, grid, dim3(96), s, pdl, *a);
    return (int)fsb_launch_pdl(fsb_march4_kernel<false, 3>, grid, dim3(96), s, pdl, *a);
  }
  if (warps_per_column == 2) {
    if (a->filter == FSB_FILTER_BILINEAR) return (int)fsb_launch_pdl(fsb_march4_kernel<true, 2>, grid, dim3(64), s, pdl, *a);
    return (int)fsb_launch_pdl(fsb_march4_kernel<false, 2>, grid, dim3(64), s, pdl, *a);
  }
  if (a->filter == FSB_FILTER_BILINEAR) return (int)fsb_launch_pdl(fsb_march4_kernel<true, 4>, grid, dim3(128), s, pdl, *a);
  return (int)fsb_launch_pdl(fsb_march4_kernel<false, 4>, grid, dim3(128), s, pdl, *a);
}
