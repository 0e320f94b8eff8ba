// monotonic_align.maximum_path on the device (mb_monotonic_path).
//   replaces  monotonic_align/core.pyx:7-42 (maximum_path_each / maximum_path_c) and the host round trip of
//             monotonic_align/__init__.py:6-19 (neg_cent -> CPU numpy -> Cython -> back to the device).
// One CTA per batch item.  The DP  value[y,x] += max(value[y-1,x-1], value[y-1,x])  over the band
// max(0, t_x + y - t_y) <= x < min(t_x, y+1) has a row-to-row dependency only: the T_y rows are processed in order, the
// x positions of a row in parallel from a shared-memory copy of the previous row (one add per cell, the same single
// float32 operation as the reference -> bit-identical values), then one thread back-tracks the path exactly like the
// reference (the comparison `value[y-1,index] < value[y-1,index-1]` on the stored float32 values).  HBM-bound integer /
// float traffic of 8 bytes per cell; no tensor cores, no reshaping into a GEMM.
#include "mb_common.h"

namespace {

__global__ void k_monotonic_path(float* __restrict__ values, int* __restrict__ paths, const int* __restrict__ t_ys,
                                 const int* __restrict__ t_xs, int T_y, int T_x, float max_neg) {
  extern __shared__ float prev[];  // previous row [T_x]
  const int b = blockIdx.x;
  float* v = values + (size_t)b * T_y * T_x;
  int* p = paths + (size_t)b * T_y * T_x;
  int t_y = t_ys[b], t_x = t_xs[b];
  t_y = t_y < 0 ? 0 : (t_y > T_y ? T_y : t_y);
  t_x = t_x < 0 ? 0 : (t_x > T_x ? T_x : t_x);
  for (int i = threadIdx.x; i < T_y * T_x; i += blockDim.x) p[i] = 0;
  for (int y = 0; y < t_y; ++y) {
    const int x0 = max(0, t_x + y - t_y), x1 = min(t_x, y + 1);
    if (y > 0)
      for (int x = threadIdx.x; x < t_x; x += blockDim.x) prev[x] = v[(size_t)(y - 1) * T_x + x];
    __syncthreads();
    for (int x = x0 + threadIdx.x; x < x1; x += blockDim.x) {
      const float v_cur = (x == y) ? max_neg : prev[x];
      const float v_prev = (x == 0) ? (y == 0 ? 0.f : max_neg) : prev[x - 1];
      v[(size_t)y * T_x + x] += fmaxf(v_prev, v_cur);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && t_x > 0) {
    int index = t_x - 1;
    for (int y = t_y - 1; y >= 0; --y) {
      p[(size_t)y * T_x + index] = 1;
      if (index != 0 && (index == y || v[(size_t)(y - 1) * T_x + index] < v[(size_t)(y - 1) * T_x + index - 1])) index = index - 1;
    }
  }
}

}  // namespace

extern "C" {

/* values: float32 [b, T_y, T_x] (device; updated IN PLACE like the reference's `value`), paths: int32 [b, T_y, T_x] (device,
 * overwritten), t_ys / t_xs: int32 [b] (device) valid lengths. */
int mb_monotonic_path(float* values, int32_t* paths, const int32_t* t_ys, const int32_t* t_xs, int32_t batch, int32_t T_y,
                      int32_t T_x, void* stream) {
  if (!values || !paths || !t_ys || !t_xs) return mb::fail(MB_ERR_INVALID, "mb_monotonic_path: null argument");
  if (batch <= 0 || T_y <= 0 || T_x <= 0) return MB_OK;
  if ((size_t)T_x * sizeof(float) > 200 * 1024) return mb::fail(MB_ERR_INVALID, "mb_monotonic_path: T_x = %d too wide", T_x);
  const int threads = T_x >= 512 ? 512 : (T_x >= 128 ? 256 : 128);
  const size_t smem = (size_t)T_x * sizeof(float);
  if (smem > 48 * 1024) MB_CUDA_CHECK(cudaFuncSetAttribute(k_monotonic_path, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_monotonic_path<<<batch, threads, smem, (cudaStream_t)stream>>>(values, paths, t_ys, t_xs, T_y, T_x, -1e9f);
  MB_LAUNCH_CHECK("k_monotonic_path");
  return MB_OK;
}

}  // extern "C"
