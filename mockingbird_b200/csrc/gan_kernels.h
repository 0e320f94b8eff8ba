// Kernel-launch interface of the GAN generator path (HiFi-GAN / Fre-GAN).
//
// Every channel-mixing layer of the generators (Conv1d, dilated Conv1d, ConvTranspose1d as a
// polyphase bank, nearest-upsample + 1x1) is expressed as ONE op shape, the "tap conv":
//
//   v[b, co, q*stride + r] = bias[co] + sum_{t < ntaps[r]} sum_ci  W[slab[r][t]][ci][co] *
//                                        act_in( x[b, ci, q + off[r][t]] )          r < stride
//   act_in = leaky_relu(., in_slope)   (in_slope == 1 -> identity); x outside [0, valid_in) == 0
//
// followed by a fused epilogue (residual add, MRF accumulate / mean, tanh, length mask).
// reference: hifigan/models.py:35-42 (ResBlock1), :134-150 (Generator.forward).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mb {

constexpr int kMaxTaps = 11;    // resblock kernel sizes are (3,7,11)
constexpr int kMaxPhases = 8;   // upsample rates are (5,5,4,2) / (5,5,2,2,2)

enum EpiMode : int {
  EPI_STORE = 0,     // y = v
  EPI_ADD = 1,       // y = y + v          (MRF: xs += resblock_j(x), models.py:143)
  EPI_ADD_DIV = 2,   // y = (y + v) / div  (MRF: x = xs / num_kernels, models.py:144)
};

struct TapConv {
  // problem
  int B, Cin, Cout, Lin, Lout;   // Lout == Lin * stride
  int stride;                    // number of output phases
  int ntaps[kMaxPhases];
  int off[kMaxPhases][kMaxTaps];   // input offset of tap t in phase r
  int slab[kMaxPhases][kMaxTaps];  // weight slab (kernel index) of tap t in phase r
  float in_slope;                // leaky-relu slope applied to x on load (1 = none)
  // epilogue
  int mode;                      // EpiMode
  float div;                     // for EPI_ADD_DIV
  int act_tanh;                  // y = tanh(v)
  // valid lengths: utterance b has lengths[b]*len_mul_in valid input rows and
  // lengths[b]*len_mul_out valid output rows (lengths == nullptr -> all valid)
  const int32_t* lengths;
  int len_mul_in, len_mul_out;
};

// ---- FP32 path: tensors are [B][C][L] fp32 (reference layout) ------------------------------
// w: [kernel_size slabs][Cin][Cout] fp32 ; bias [Cout] or nullptr ; res [B][Cout][Lout] or nullptr
// y2 (optional): second destination receiving y2 += v_final (Fre-GAN "x += cond_up(mel)")
cudaError_t launch_tapconv_f32(const TapConv& p, const float* x, const float* w, const float* bias,
                               const float* res, float* y, float* y2, cudaStream_t stream);

// dst[i] += src[i]
cudaError_t launch_add_inplace_f32(float* dst, const float* src, size_t n, cudaStream_t stream);

// weight repack: Conv1d weight [Cout][Cin][K] or ConvTranspose1d weight [Cin][Cout][K]
//   -> slabs [K][Cin][Cout]
cudaError_t launch_pack_slabs_f32(const float* w, float* dst, int Cout, int Cin, int K, bool transposed,
                                  cudaStream_t stream);

}  // namespace mb
