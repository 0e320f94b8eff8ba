// FP32 kernels of the Tacotron path (see tacotron_kernels.cuh).
#include "tacotron_kernels.cuh"

namespace mb {
namespace taco {

namespace {

constexpr int BM = 64;
constexpr int BK = 16;

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

// 64 x BN output tile, 256 threads, thread tile 4 x (BN/16); operands staged k-major in shared memory
template <int BN>
__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs a) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int s = 0; s < a.nseg; ++s) {
    const Seg sg = a.seg[s];
    for (int k0 = 0; k0 < sg.K; k0 += BK) {
      // A tile: 64 rows x 16 k (k fastest -> coalesced along the feature axis)
      for (int i = tid; i < BM * BK; i += 256) {
        const int mm = i >> 4, kk = i & 15;
        const int m = m0 + mm;
        float v = 0.f;
        if (m < a.M && k0 + kk < sg.K) {
          int row = m;
          bool ok = true;
          if (sg.shift != 0) {
            const int t = m % a.T + sg.shift;
            ok = (t >= 0 && t < a.T);
            row = m + sg.shift;
          }
          if (ok) v = sg.x[(size_t)row * sg.ld + k0 + kk];
        }
        As[kk][mm] = v;
      }
      for (int i = tid; i < BN * BK; i += 256) {
        const int nn = i >> 4, kk = i & 15;
        const int n = n0 + nn;
        float v = 0.f;
        if (n < a.N && k0 + kk < sg.K) v = a.W[(size_t)n * a.ldw + sg.w_off + (size_t)(k0 + kk) * sg.w_stride];
        Bs[kk][nn] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float ar[4] = {av.x, av.y, av.z, av.w};
        float br[TN];
#pragma unroll
        for (int j = 0; j < TN; j += 2) {
          const float2 bv = *reinterpret_cast<const float2*>(&Bs[kk][tx * TN + j]);
          br[j] = bv.x;
          br[j + 1] = bv.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[n];
      v = act_fn(v, a.act);
      if (a.bn_scale) v = fmaf(v, a.bn_scale[n], a.bn_shift[n]);
      if (a.mask) v = a.mask[(size_t)m * a.N + n] ? v * 2.f : 0.f;
      if (a.res) v += a.res[(size_t)m * a.ldres + n];
      a.Y[(size_t)m * a.ldy + n] = v;
    }
  }
}

__global__ void gru_cell_kernel(const float* __restrict__ gi, int ldgi, const float* __restrict__ gh, float* h, int ldh,
                                float* out2, int ldout2, int M, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * H) return;
  const int m = i / H, u = i - m * H;
  const float* a = gi + (size_t)m * ldgi;
  const float* b = gh + (size_t)m * 3 * H;
  const float r = 1.f / (1.f + expf(-(b[u] + a[u])));
  const float z = 1.f / (1.f + expf(-(b[H + u] + a[H + u])));
  const float n = tanhf(a[2 * H + u] + b[2 * H + u] * r);
  const float hv = h[(size_t)m * ldh + u];
  const float hn = (hv - n) * z + n;
  h[(size_t)m * ldh + u] = hn;
  if (out2) out2[(size_t)m * ldout2 + u] = hn;
}

__global__ void lstm_cell_kernel(const float* __restrict__ g, float* c, float* h, float* x, int M, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * H) return;
  const int m = i / H, u = i - m * H;
  const float* gg = g + (size_t)m * 4 * H;
  const float ig = 1.f / (1.f + expf(-gg[u]));
  const float fg = 1.f / (1.f + expf(-gg[H + u]));
  const float cg = tanhf(gg[2 * H + u]);
  const float og = 1.f / (1.f + expf(-gg[3 * H + u]));
  const float cn = fg * c[i] + ig * cg;
  const float hn = og * tanhf(cn);
  c[i] = cn;
  h[i] = hn;
  if (x) x[i] = x[i] + hn;
}

__global__ void highway_kernel(const float* __restrict__ x12, float* y, int M, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const int m = i / C, c = i - m * C;
  const float x1 = x12[(size_t)m * 2 * C + c];
  const float x2 = x12[(size_t)m * 2 * C + C + c];
  const float g = 1.f / (1.f + expf(-x2));
  y[i] = g * (x1 > 0.f ? x1 : 0.f) + (1.f - g) * y[i];
}

__global__ void maxpool2_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * T * C) return;
  const int t = (int)((i / C) % T);
  const float v = x[i];
  y[i] = t > 0 ? fmaxf(v, x[i - C]) : v;
}

__global__ void embedding_kernel(const int32_t* __restrict__ ids, const float* __restrict__ table, float* __restrict__ y,
                                 int M, int D) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * D) return;
  const int m = (int)(i / D), d = (int)(i - (size_t)m * D);
  y[i] = table[(size_t)ids[m] * D + d];
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int ldsrc, int rows_per_src, float* __restrict__ dst,
                                 int lddst, int off, int M, int n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * n) return;
  const int m = (int)(i / n), j = (int)(i - (size_t)m * n);
  dst[(size_t)m * lddst + off + j] = src[(size_t)(m / rows_per_src) * ldsrc + j];
}

}  // namespace

cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) return cudaSuccess;
  // small-M (decoder step) problems: narrower N tiles -> more CTAs
  if (a.M <= 128 && a.N >= 512) {
    dim3 grid((a.N + 31) / 32, (a.M + BM - 1) / BM);
    gemm_kernel<32><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid((a.N + 63) / 64, (a.M + BM - 1) / BM);
    gemm_kernel<64><<<grid, 256, 0, st>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_gru_cell(const float* gi, int ldgi, const float* gh, float* h, int ldh, float* out2, int ldout2,
                            int M, int H, cudaStream_t st) {
  gru_cell_kernel<<<(M * H + 255) / 256, 256, 0, st>>>(gi, ldgi, gh, h, ldh, out2, ldout2, M, H);
  return cudaGetLastError();
}

cudaError_t launch_lstm_cell(const float* g, float* c, float* h, float* x, int M, int H, cudaStream_t st) {
  lstm_cell_kernel<<<(M * H + 255) / 256, 256, 0, st>>>(g, c, h, x, M, H);
  return cudaGetLastError();
}

cudaError_t launch_highway(const float* x12, float* y, int M, int C, cudaStream_t st) {
  highway_kernel<<<(unsigned)(((size_t)M * C + 255) / 256), 256, 0, st>>>(x12, y, M, C);
  return cudaGetLastError();
}

cudaError_t launch_maxpool2(const float* x, float* y, int B, int T, int C, cudaStream_t st) {
  maxpool2_kernel<<<(unsigned)(((size_t)B * T * C + 255) / 256), 256, 0, st>>>(x, y, B, T, C);
  return cudaGetLastError();
}

cudaError_t launch_embedding(const int32_t* ids, const float* table, float* y, int M, int D, cudaStream_t st) {
  embedding_kernel<<<(unsigned)(((size_t)M * D + 255) / 256), 256, 0, st>>>(ids, table, y, M, D);
  return cudaGetLastError();
}

cudaError_t launch_copy_cols(const float* src, int ldsrc, int rows_per_src, float* dst, int lddst, int off, int M,
                             int n, cudaStream_t st) {
  copy_cols_kernel<<<(unsigned)(((size_t)M * n + 255) / 256), 256, 0, st>>>(src, ldsrc, rows_per_src, dst, lddst, off, M, n);
  return cudaGetLastError();
}

}  // namespace taco
}  // namespace mb
