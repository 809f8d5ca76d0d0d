/*
 * fsb_expand_tma.cu -- expand with TMA tile stores (sm_100a): lists -> pixels for 4-byte records, the frame written by
 * cp.async.bulk.tensor instead of per-lane stores.
 *
 *   fsb_expand4_tma_kernel   the walk of fsb_expand4_kernel (scatter, fill scan, sky: fut/voxel_renderer.fut:244-248);
 *                            a warp stages its 32-row x 32-column tile in shared memory and one lane hands it to the
 *                            TMA unit as one 3-D box {32 columns, 32 rows, 1 pose} of the frame tensor (the final
 *                            `transpose`, :251, is the tile's row-major layout).
 *
 * The frame tensor map (CUtensorMap: u32 [n_poses][h][w] with the caller's row and pose strides) is encoded on the host
 * for every launch -- the destination is the caller's buffer -- through the driver entry point cuTensorMapEncodeTiled,
 * obtained with cudaGetDriverEntryPoint so that the library does not link libcuda.  TMA needs a 16-byte aligned base and
 * strides that are multiples of 16 bytes; fsb_launch_expand falls back to fsb_expand4_kernel otherwise.  Boxes that hang
 * over the right or bottom edge of the frame are clipped by the hardware.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "fsb_device.cuh"

#define FSB_XR 32

/* see fsb_kernels.cu */
struct list_view_t {
  size_t rec0, sidx0;
  int stride;
  __device__ __forceinline__ list_view_t(const fsb_render_args &a, int pose, int jrel) {
    if (a.rec_stride == 1) {
      const size_t colid = (size_t)pose * (a.col_end - a.col_begin) + jrel;
      rec0 = colid * a.rec_cap + 1;
      sidx0 = colid * (a.n_bands + 1);
      stride = 1;
    } else {
      const size_t gid = (size_t)pose * (a.ncols_pad >> 5) + (jrel >> 5);
      rec0 = (gid * a.rec_cap + 1) * 32 + (jrel & 31);
      sidx0 = gid * (a.n_bands + 1) * 32 + (jrel & 31);
      stride = 32;
    }
  }
};

__global__ void __launch_bounds__(256) fsb_expand4_tma_kernel(const fsb_render_args a, const __grid_constant__ CUtensorMap frame_map) {
  __shared__ __align__(128) uint32_t tiles[8][FSB_XR][32]; /* one 4 KB tile per warp, rows of 128 bytes */
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pose = blockIdx.z;
  const int band = blockIdx.y * 8 + warp;
  const int ncols = a.col_end - a.col_begin;
  const int jrel = blockIdx.x * 32 + lane;
  if (band >= a.n_bands) return; /* warp-uniform */
  const int nrows = min(FSB_XR, a.h - band * FSB_XR);
  uint32_t *tile = &tiles[warp][0][0];
  if (jrel < ncols) {
    const fsb_frame_consts fc = a.fc[pose];
    const uint32_t empty = fc.empty;
    const list_view_t lv(a, pose, jrel);
    const uint32_t *rec = reinterpret_cast<const uint32_t *>(a.recs) + lv.rec0;
    const uint32_t *sidx = a.sidx + lv.sidx0;
    const int rs = lv.stride;
    const int lo = (int)__ldg(sidx + (band + 1) * rs), hi = (int)__ldg(sidx + band * rs), n = (int)__ldg(sidx);
#define FSB_REC4_COLOUR(w) (((w) & 0x00FFFFFFu) | ((uint32_t)((int32_t)(w) >> 31) & 0xFF000000u))
    int idx = hi - 1;
    uint32_t cur = empty; /* running colour entering the band: the first non-transparent record above it */
    if (hi < n) {
      const uint32_t w = rec[hi * rs];
      cur = FSB_REC4_COLOUR(w);
    }
    uint32_t nxt = idx >= lo ? rec[idx * rs] : 0u;
    for (int i = hi + 1; cur == empty && i < n; ++i) {
      const uint32_t w = rec[i * rs];
      cur = FSB_REC4_COLOUR(w);
    }
    if (cur == empty) cur = fc.sky;
    const int rsb = rs * 4;
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(tile + lane);
    /* one row: see fsb_expand4_kernel; the pixel goes to the tile instead of the frame */
#define FSB_EXPAND4_ROW_S(r)                                                               \
  asm volatile(                                                                            \
      "{\n\t.reg .pred m, c;\n\t.reg .u64 ra;\n\t.reg .u32 t, col;\n\t.reg .s32 sg;\n\t"  \
      "bfe.u32 t, %0, 24, 5;\n\t"                                                          \
      "setp.ge.s32 m, %2, %3;\n\t"                                                         \
      "setp.eq.and.u32 m, t, %4, m;\n\t"                                                   \
      "shr.s32 sg, %0, 31;\n\t"                                                            \
      "lop3.b32 col, %0, sg, 0xFF000000, 0xD8;\n\t"                                        \
      "setp.ne.and.u32 c, col, %5, m;\n\t"                                                 \
      "@c mov.u32 %1, col;\n\t"                                                            \
      "@m add.s32 %2, %2, -1;\n\t"                                                         \
      "mul.wide.s32 ra, %2, %8;\n\t"                                                       \
      "add.s64 ra, ra, %6;\n\t"                                                            \
      "@m ld.global.u32 %0, [ra];\n\t"                                                     \
      "st.shared.u32 [%7], %1;\n\t}"                                                       \
      : "+r"(nxt), "+r"(cur), "+r"(idx)                                                    \
      : "r"(lo), "r"((int)(r)), "r"(empty), "l"(rec), "r"(sa + (uint32_t)(r) * 128u), "r"(rsb) \
      : "memory");
    if (nrows == FSB_XR) {
#pragma unroll
      for (int r = 0; r < FSB_XR; ++r) FSB_EXPAND4_ROW_S(r)
    } else {
      for (int r = 0; r < nrows; ++r) FSB_EXPAND4_ROW_S(r)
    }
#undef FSB_EXPAND4_ROW_S
#undef FSB_REC4_COLOUR
  }
  /* the tile is handed from the generic proxy (the st.shared above) to the async proxy (TMA) */
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&frame_map),
                 "r"(blockIdx.x * 32), "r"(band * FSB_XR), "r"(pose), "r"(src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); /* the tile may be released once the TMA has read it */
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encoder() {
  static encode_tiled_fn fn = nullptr;
  static int tried = 0;
  if (!tried) {
    tried = 1;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = (encode_tiled_fn)p;
    else
      (void)cudaGetLastError();
  }
  return fn;
}

/* 1 if this launch can use the TMA store (alignment rules of the tensor map), 0 otherwise */
extern "C" int fsb_expand_tma_applicable(const fsb_render_args *a) {
  if (!a->rec4 || a->smooth) return 0;
  if (((uintptr_t)a->out & 15u) || (a->row_stride & 3) || (a->n_poses > 1 && (a->pose_stride & 3))) return 0;
  if (a->row_stride * 4 >= (1ll << 40) || a->pose_stride * 4 >= (1ll << 40)) return 0;
  return get_encoder() != nullptr;
}

extern "C" int fsb_launch_expand_tma(const fsb_render_args *a, void *stream, int64_t *launches) {
  encode_tiled_fn enc = get_encoder();
  if (!enc) return (int)cudaErrorNotSupported;
  const int ncols = a->col_end - a->col_begin;
  CUtensorMap map;
  /* u32 [n_poses][h][ncols] at a->out with the caller's strides; box = one warp's tile */
  const cuuint64_t dims[3] = {(cuuint64_t)ncols, (cuuint64_t)a->h, (cuuint64_t)a->n_poses};
  const cuuint64_t strides[2] = {(cuuint64_t)a->row_stride * 4, (cuuint64_t)(a->n_poses > 1 ? a->pose_stride : (int64_t)a->row_stride * a->h) * 4};
  const cuuint32_t box[3] = {32, FSB_XR, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)a->out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  dim3 grid((ncols + 31) / 32, (a->n_bands + 7) / 8, a->n_poses);
  fsb_expand4_tma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a, map);
  if (launches) ++*launches;
  return (int)cudaGetLastError();
}
