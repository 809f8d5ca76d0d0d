/*
 * fsb_march_split.cu -- the march of single frames and small batches on the texture path (sm_100a): depth-parallel,
 * one thread-block cluster per group of 32 screen columns.
 *
 *   fsb_marchs_kernel   lane = screen column (as in fsb_marchc_kernel), the depth series of the group is cut into
 *                       segments of 1-4 chunks of 32 samples, one segment per warp of the cluster (8 CTAs x 4 or 8 warps).
 *                       Every warp builds the depth-table entries of its own samples (get_zs / get_h_line / inv_z,
 *                       fut/voxel_renderer.fut:28-34,43-60,217 -- no set-up launch), samples and projects them at once
 *                       (:215-225) and keeps the projected rows in shared memory.  `scan occlude` (:231, :69-72) is
 *                       associative: the segments exchange their minima through distributed shared memory, every
 *                       segment re-reads its rows against the running minimum that enters it, the visible counts are
 *                       exchanged the same way, and each segment appends its visible samples to the column's candidate
 *                       list at its final position -- the list fsb_colour_kernel and the expand kernels consume, word
 *                       for word what fsb_marchc_kernel writes.
 *
 * Why: a lone frame is the reference's actual use (one render per SDL iteration, c/interactive.c:111).  It offers 60
 * warps of 32 columns (1920 columns); walking 63 chunks one after the other per warp is a chain of dependent texture
 * round trips (fsb_march_kernel 23 us, fsb_march4_kernel 22 us of a 29.5 us frame, profiles/r2_pdl_single_frame.jsonl).
 * Here every sample of the frame is in flight at once: 3840 warps each fetch 32-64 samples, and what stays sequential
 * is two cluster barriers and two short loops over 16-bit minima / counts in shared memory.
 *
 * Saturating i32.f32 only (the texture path); float discipline as in fsb_kernels.cu.
 */
#include <cooperative_groups.h>
#include <stdlib.h>

#include "fsb_device.cuh"

namespace cg = cooperative_groups;

#define MS_CLUSTER 8  /* CTAs per cluster (the portable maximum) */
#define MS_MAX_CPW 4  /* chunks of 32 samples per warp at most */

/* rows of a warp's samples: u16, two consecutive depth steps of a lane share a 32-bit word (conflict-free both ways) */
__device__ __forceinline__ int ms_row_index(int step, int lane) { return (((step >> 1) * 32 + lane) << 1) + (step & 1); }

template <bool BIL>
__device__ __forceinline__ int ms_resolve(const col_step<BIL> &t, float iz, float cam_h, float horizon) {
  const float rel = __fadd_rn(__fmul_rn(__fsub_rn(cam_h, cstep_height<BIL>(t)), iz), horizon); /* :223-224 */
  return min(max(0, __float2int_rz(rel)), 0xFFFF);                                              /* :225; rows >= h are never visible */
}

template <bool BIL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 32 / WARPS) fsb_marchs_kernel(const fsb_render_args a, const fsb_frame_consts single,
                                                                              int use_single, int cpw_max) {
  constexpr int G = MS_CLUSTER * WARPS; /* segments (= warps) per group of 32 columns */
  extern __shared__ __align__(16) unsigned char ms_smem[];
  /* layout: rows u16 [WARPS][cpw_max * 32 * 32] | table float [WARPS][160] | segmin u16 [G][32] | segcnt u16 [G][32] */
  uint16_t *rows_all = reinterpret_cast<uint16_t *>(ms_smem);
  float *tab_all = reinterpret_cast<float *>(ms_smem + (size_t)WARPS * cpw_max * 2048);
  uint16_t *segmin = reinterpret_cast<uint16_t *>(tab_all + WARPS * FSB_TAB_BLOCK);
  uint16_t *segcnt = segmin + G * 32;

  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int crank = (int)cluster.block_rank();
  const int group = blockIdx.x / MS_CLUSTER, pose = blockIdx.y;
  const int seg = crank * WARPS + warp;
  pdl_trigger(); /* the colour pass may be scheduled; it waits for this grid's writes */

  if (use_single && blockIdx.x == 0 && tid == 0) const_cast<fsb_frame_consts *>(a.fc)[0] = single; /* the expand kernels read sky / empty */
  const fsb_frame_consts fc = use_single ? single : a.fc[pose];
  const int ncols = a.col_end - a.col_begin;
  const int n_chunks = (fc.n_z + 31) >> 5;
  const float cam_h = fc.cam_h, horizon = fc.horizon;
  uint16_t *rows = rows_all + (size_t)warp * cpw_max * 1024;
  float *tab = tab_all + warp * FSB_TAB_BLOCK;
  const int kcap = a.tab_stride / 5;
  float4 *g_line = reinterpret_cast<float4 *>(const_cast<float *>(a.table) + (size_t)pose * a.tab_stride);
  float *g_invz = const_cast<float *>(a.table) + (size_t)pose * a.tab_stride + 4 * (size_t)kcap;

  /* Occlusion bound (see fsb_kernels.cu): camera above the highest terrain -> a prefix of the series projects below the
   * bottom row and is skipped (lane = chunk, bound at the chunk's last sample).  The other half of the bound (camera
   * below: a column ends once the bound reaches its y-buffer) needs the running minimum and has no place here. */
  int c_first = 0;
  if (fc.cull_d >= 0.0f && fc.cull_d < INFINITY) {
    c_first = n_chunks;
    for (int base = 0; base < n_chunks; base += 32) {
      const int ci = min(base + lane, n_chunks - 1);
      float4 l;
      float izl;
      depth_entry(fc, ci * 32 + 31, l, izl);
      const bool below = max(0, __float2int_rz(__fadd_rn(__fmul_rn(fc.cull_d, izl), horizon))) >= a.h;
      const unsigned alive = __ballot_sync(FSB_FULL, !below);
      if (alive) {
        c_first = base + __ffs(alive) - 1;
        break;
      }
    }
  }
  const int n_act = n_chunks - c_first;
  const int cpw = (n_act + G - 1) / G; /* <= cpw_max (the host sized the rows for the whole series) */
  const int c0 = c_first + seg * cpw, c1 = min(c0 + cpw, n_chunks);
  const float fj = (float)(a.col_begin + group * 32 + lane);

  /* ---- phase 1: sample and project this segment; rows -> shared memory, minimum -> every CTA of the cluster ---- */
  int m = 0xFFFF;
  for (int c = c0, cl = 0; c < c1; ++c, ++cl) {
    {
      float4 l;
      float iz;
      const int k = c * 32 + lane;
      depth_entry(fc, k, l, iz);
      reinterpret_cast<float4 *>(tab)[lane] = l;
      tab[128 + lane] = iz;
      if (group == 0) { /* the colour pass rebuilds sample positions from the table in global memory */
        g_line[k] = l;
        g_invz[k] = iz;
      }
    }
    __syncwarp();
    const float4 *tl = reinterpret_cast<const float4 *>(tab);
    const float *tz = tab + 128;
    uint16_t *rw = rows + cl * 1024;
    col_step<BIL> sa[4], sb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) cstep_issue<BIL>(sa[u], a, tl[u], fj);
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
#pragma unroll
      for (int u = 0; u < 4; ++u) cstep_issue<BIL>(sb[u], a, tl[i + 4 + u], fj);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int y = ms_resolve<BIL>(sa[u], tz[i + u], cam_h, horizon);
        rw[ms_row_index(i + u, lane)] = (uint16_t)y;
        m = min(m, y);
      }
      if (i + 8 < 32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) cstep_issue<BIL>(sa[u], a, tl[i + 8 + u], fj);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int y = ms_resolve<BIL>(sb[u], tz[i + 4 + u], cam_h, horizon);
        rw[ms_row_index(i + 4 + u, lane)] = (uint16_t)y;
        m = min(m, y);
      }
    }
    __syncwarp(); /* the table slot is rewritten for the next chunk */
  }
#pragma unroll
  for (int r = 0; r < MS_CLUSTER; ++r) cluster.map_shared_rank(segmin, r)[seg * 32 + lane] = (uint16_t)m;
  cluster.sync();

  /* ---- phase 2: the running minimum entering this segment; which of its samples lower it (`occlude`, strict <) ---- */
  int carry = a.h; /* the neutral (0, h) of :231 */
  for (int s = 0; s < seg; ++s) carry = min(carry, (int)segmin[s * 32 + lane]);
  uint32_t vis[MS_MAX_CPW];
  int cnt = 0;
#pragma unroll
  for (int cl = 0; cl < MS_MAX_CPW; ++cl) vis[cl] = 0u;
  if (__any_sync(FSB_FULL, m < carry)) {
    int ybuf = carry;
#pragma unroll
    for (int cl = 0; cl < MS_MAX_CPW; ++cl) {
      if (c0 + cl < c1) {
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(rows + cl * 1024) + lane;
        uint32_t mask = 0u;
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
          const uint32_t w = rw[i2 * 32];
          const int ya = (int)(w & 0xFFFFu), yb = (int)(w >> 16);
          if (ya < ybuf) {
            mask |= 1u << (2 * i2);
            ybuf = ya;
          }
          if (yb < ybuf) {
            mask |= 2u << (2 * i2);
            ybuf = yb;
          }
        }
        vis[cl] = mask;
        cnt += __popc(mask);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < MS_CLUSTER; ++r) cluster.map_shared_rank(segcnt, r)[seg * 32 + lane] = (uint16_t)cnt;
  cluster.sync();

  /* ---- phase 3: list positions; append the visible samples (row | sample index << 15) ---- */
  int pos = 0, total = 0;
  for (int s = 0; s < G; ++s) {
    const int v = (int)segcnt[s * 32 + lane];
    if (s < seg) pos += v;
    total += v;
  }
  uint32_t *list = a.cand + cand_group_base(a, pose, group) + lane; /* entry p of this lane's column at list[p * 32] */
#pragma unroll
  for (int cl = 0; cl < MS_MAX_CPW; ++cl) {
    uint32_t mask = vis[cl];
    const uint16_t *rw = rows + cl * 1024;
    const uint32_t kbase = (uint32_t)((c0 + cl) << 5);
    while (mask) {
      const int i = __ffs(mask) - 1;
      mask &= mask - 1u;
      list[(size_t)pos * 32] = (uint32_t)rw[ms_row_index(i, lane)] | ((kbase + (uint32_t)i) << FSB_ROW_BITS);
      ++pos;
    }
  }
  if (seg == 0) {
    const int jrel = group * 32 + lane;
    a.cand_cnt[(size_t)pose * a.ncols_pad + jrel] = (uint32_t)total;
    if (a.stats) {
      /* in chunks of 32 samples of one column, like the other marches count them */
      if (lane == 0) atomicAdd(a.stats, (unsigned long long)n_act * (unsigned long long)min(32, ncols - group * 32));
      unsigned long long tot = (jrel < ncols) ? (unsigned long long)total : 0ull;
#pragma unroll
      for (int d = 16; d; d >>= 1) tot += __shfl_xor_sync(FSB_FULL, tot, d);
      if (lane == 0) atomicAdd(a.stats + 1, tot);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
static size_t ms_smem_bytes(int warps, int cpw_max) {
  return (size_t)warps * cpw_max * 2048 + (size_t)warps * FSB_TAB_BLOCK * 4 + 2 * (size_t)MS_CLUSTER * warps * 32 * 2;
}

template <bool BIL, int WARPS>
static int launch_marchs_t(const fsb_render_args &a, const fsb_frame_consts *single, int cpw_max, cudaStream_t s) {
  const size_t smem = ms_smem_bytes(WARPS, cpw_max);
  static bool attr_set = false; /* per instantiation */
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fsb_marchs_kernel<BIL, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ms_smem_bytes(WARPS, MS_MAX_CPW));
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.ncols_pad / 32) * MS_CLUSTER, (unsigned)a.n_poses);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MS_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  fsb_frame_consts dummy = {};
  return (int)cudaLaunchKernelEx(&cfg, fsb_marchs_kernel<BIL, WARPS>, a, single ? *single : dummy, single ? 1 : 0, cpw_max);
}

extern "C" int fsb_march_split_max_chunks(int warps_per_group) { return warps_per_group * MS_MAX_CPW; }

/* warps_per_group: 32 or 64 (8 CTAs of 4 or 8 warps); n_chunks_max: chunks of 32 samples of the longest series of the launch.
 * single != NULL: one pose whose constants travel as a kernel argument (and are stored to a->fc[0] by the kernel). */
extern "C" int fsb_launch_march_split(const fsb_render_args *a, const fsb_frame_consts *single, int warps_per_group,
                                      int n_chunks_max, void *stream, int64_t *launches) {
  cudaStream_t s = (cudaStream_t)stream;
  const bool bil = a->filter == FSB_FILTER_BILINEAR;
  int cpw_max = (n_chunks_max + warps_per_group - 1) / warps_per_group;
  if (cpw_max < 1) cpw_max = 1;
  if (cpw_max > MS_MAX_CPW || (warps_per_group != 32 && warps_per_group != 64)) return (int)cudaErrorInvalidValue;
  int rc;
  if (warps_per_group == 64) rc = bil ? launch_marchs_t<true, 8>(*a, single, cpw_max, s) : launch_marchs_t<false, 8>(*a, single, cpw_max, s);
  else rc = bil ? launch_marchs_t<true, 4>(*a, single, cpw_max, s) : launch_marchs_t<false, 4>(*a, single, cpw_max, s);
  if (launches) ++*launches;
  return rc;
}
