// FP32 FFMA implementation of the tap-conv op (MB_PREC_FP32): the parity anchor of the GAN path.
// Register-tiled direct convolution, 64(co) x 64(l) output tile per CTA, 4x4 per thread, input
// window and weight slabs staged through shared memory in chunks of 8 input channels.
// Reference semantics: hifigan/models.py:35-42, :134-150; fregan/generator.py:137-166.
#include "gan_kernels.h"

namespace mb {

namespace {

constexpr int CO_T = 64;
constexpr int L_T = 64;
constexpr int CI_T = 8;
constexpr int XW_MAX = L_T + 64;  // window: tile + tap span (<= 50 for k=11,d=5)

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__global__ void __launch_bounds__(256) tapconv_f32_kernel(TapConv p, const float* __restrict__ x,
                                                          const float* __restrict__ w,
                                                          const float* __restrict__ bias,
                                                          const float* res, float* y, float* y2) {
  __shared__ float xs[CI_T][XW_MAX];
  __shared__ __align__(16) float ws[kMaxTaps][CI_T][CO_T];

  const int tid = threadIdx.x;
  const int tx = tid & 15;   // l lane
  const int ty = tid >> 4;   // co group (4 channels)
  const int q0 = blockIdx.x * L_T;
  const int co0 = blockIdx.y * CO_T;
  const int b = blockIdx.z / p.stride;
  const int r = blockIdx.z % p.stride;

  const int nt = p.ntaps[r];
  int omin = 0x7fffffff, omax = -0x7fffffff;
  for (int t = 0; t < nt; ++t) {
    omin = min(omin, p.off[r][t]);
    omax = max(omax, p.off[r][t]);
  }
  const int XW = L_T + (omax - omin);
  const int valid_in = p.lengths ? min(p.Lin, p.lengths[b] * p.len_mul_in) : p.Lin;
  const int valid_out = p.lengths ? min(p.Lout, p.lengths[b] * p.len_mul_out) : p.Lout;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float* xb = x + (size_t)b * p.Cin * p.Lin;
  for (int ci0 = 0; ci0 < p.Cin; ci0 += CI_T) {
    // stage input window (activation applied once here)
    for (int i = tid; i < CI_T * XW; i += 256) {
      const int ci = i / XW, j = i - ci * XW;
      const int l = q0 + omin + j;
      float v = 0.f;
      if (ci0 + ci < p.Cin && l >= 0 && l < valid_in) v = lrelu(xb[(size_t)(ci0 + ci) * p.Lin + l], p.in_slope);
      xs[ci][j] = v;
    }
    // stage weight slabs
    for (int i = tid; i < nt * CI_T * CO_T; i += 256) {
      const int t = i / (CI_T * CO_T);
      const int rem = i - t * (CI_T * CO_T);
      const int ci = rem / CO_T, co = rem - ci * CO_T;
      float v = 0.f;
      if (ci0 + ci < p.Cin && co0 + co < p.Cout)
        v = w[((size_t)p.slab[r][t] * p.Cin + (ci0 + ci)) * p.Cout + co0 + co];
      ws[t][ci][co] = v;
    }
    __syncthreads();
    for (int t = 0; t < nt; ++t) {
      const int o = p.off[r][t] - omin;
#pragma unroll
      for (int ci = 0; ci < CI_T; ++ci) {
        const float4 wv = *reinterpret_cast<const float4*>(&ws[t][ci][ty * 4]);
        float xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = xs[ci][tx + 16 * j + o];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0][j] = fmaf(wv.x, xv[j], acc[0][j]);
          acc[1][j] = fmaf(wv.y, xv[j], acc[1][j]);
          acc[2][j] = fmaf(wv.z, xv[j], acc[2][j]);
          acc[3][j] = fmaf(wv.w, xv[j], acc[3][j]);
        }
      }
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.Cout) continue;
    const float bv = bias ? bias[co] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + tx + 16 * j;
      if (q >= p.Lin) continue;
      const int lo = q * p.stride + r;
      const size_t idx = ((size_t)b * p.Cout + co) * p.Lout + lo;
      float v = acc[i][j] + bv;
      if (res) v += res[idx];
      if (p.mode == EPI_ADD) v = y[idx] + v;
      else if (p.mode == EPI_ADD_DIV) v = (y[idx] + v) / p.div;
      if (p.act_tanh) v = tanhf(v);
      if (lo >= valid_out) v = 0.f;
      y[idx] = v;
      if (y2) y2[idx] = (lo >= valid_out) ? 0.f : y2[idx] + v;
    }
  }
}

__global__ void add_inplace_kernel(float* dst, const float* __restrict__ src, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] += src[i];
}

__global__ void pack_slabs_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin,
                                  int K, int transposed) {
  const size_t n = (size_t)Cout * Cin * K;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int co = (int)(i % Cout);
  const int ci = (int)((i / Cout) % Cin);
  const int k = (int)(i / ((size_t)Cout * Cin));
  const size_t src = transposed ? ((size_t)ci * Cout + co) * K + k : ((size_t)co * Cin + ci) * K + k;
  dst[i] = w[src];
}

}  // namespace

cudaError_t launch_tapconv_f32(const TapConv& p, const float* x, const float* w, const float* bias,
                               const float* res, float* y, float* y2, cudaStream_t stream) {
  dim3 grid((p.Lin + L_T - 1) / L_T, (p.Cout + CO_T - 1) / CO_T, p.B * p.stride);
  tapconv_f32_kernel<<<grid, 256, 0, stream>>>(p, x, w, bias, res, y, y2);
  return cudaGetLastError();
}

cudaError_t launch_add_inplace_f32(float* dst, const float* src, size_t n, cudaStream_t stream) {
  const int threads = 256;
  size_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  add_inplace_kernel<<<(unsigned)blocks, threads, 0, stream>>>(dst, src, n);
  return cudaGetLastError();
}

cudaError_t launch_pack_slabs_f32(const float* w, float* dst, int Cout, int Cin, int K, bool transposed,
                                  cudaStream_t stream) {
  const size_t n = (size_t)Cout * Cin * K;
  pack_slabs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, dst, Cout, Cin, K, transposed ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace mb
