// Tensor-core execution of the GAN generator plan (MB_PREC_F16TC), sm_100a only.
//
// tc_conv_kernel: the tap conv (gan_kernels.h) as an implicit GEMM on tcgen05 tensor cores.
//   D[128 rows x Cout] (fp32, TMEM) += A[128 rows x 16 ci] (fp16, smem) * B[Cout x 16 ci]^T (fp16, smem)
//   * M = 128 consecutive time rows, N = Cout, K = (tap, input channel).
//   * activations live in HBM as "F16B" planes [B][C/64][Lp][64] (already leaky-relu'd by the
//     producer's epilogue) whose rows are stored PRE-SWIZZLED: the 16-byte chunks of every 128-byte
//     row are XOR-permuted by (row & 7) exactly like the UMMA SWIZZLE_128B shared-memory layout
//     (64-byte rows / SWIZZLE_64B when C = 32).  One bulk-TMA copy (cp.async.bulk) per 64-channel
//     chunk fetches the rows [floor8(m0 + omin), ... + W) of a work item; because the copy starts at
//     a row that is a multiple of 8 and lands 1024-byte aligned, it arrives in shared memory already
//     in the canonical K-major swizzled operand layout (SBO = 1024 B) - no tensor map, no repack.
//     The operand of tap t is the SAME buffer with the descriptor start address advanced by off_t
//     rows (the swizzle is a function of absolute smem address bits, so no base-offset fix-up is needed): every tap, every dilation and every phase of a
//     transposed conv reuse one window; zero padding comes from the zero pad rows of the plane.
//     (A first version used the SWIZZLE_NONE 8x16 B core-matrix layout: correct, but the tensor core
//     fetches such operands at 32 B/clk - 128 cycles per 128x16 A tile, profiles/r01_*swizzle_none*.)
//   * weights: per (kernel index, 64-channel chunk) an fp16 image [Cout][64] with the same swizzle,
//     streamed through a ring of shared-memory stages by bulk copies, or kept resident when the
//     layer's whole weight set fits (all C<=64 layers).
//   * warp roles: warp 0 = copy producer, warp 1 = MMA issuer (warp-uniform loop, one elected lane issues) + TMEM allocator,
//     warps 2-9 = epilogue (TMEM -> registers -> bias/residual/MRF/leaky-relu -> fp32 F32B plane
//     and/or fp16 F16B plane, fully coalesced 16 B per thread per 8 channels).
//   * accumulators double-buffered in TMEM (2 x MT x Cout columns <= 512) so the epilogue of work
//     item i overlaps the MMAs of item i+1; persistent CTAs, one per SM, static round-robin.
// reference semantics: hifigan/models.py:35-42 (ResBlock1), :134-150 (Generator.forward)
#include "gan_tc.h"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

#include "gan_tc_dev.cuh"
#include "mb_common.h"

namespace mb {

namespace {

__host__ __device__ constexpr int tc_threads(int ew) { return 64 + 32 * ew; }
constexpr int kAStages = 2;
constexpr int kAccStages = 2;
constexpr int kMaxWStages = 16;
constexpr uint32_t kSmemMax = 227 * 1024;
constexpr size_t kPlaneSlack = 128 * 1024;

struct TcParams {
  int B, Lin, Lout, Cin, Cout;
  int stride;
  int ntaps[kMaxPhases];
  int off[kMaxPhases][kMaxTaps];
  int slab[kMaxPhases][kMaxTaps];
  int omin, W, MT;
  int cw, row_bytes, nk16, n_cchunks;   // channels per K-chunk (64 or 32), bytes per operand row
  int layout_type;                     // UMMA layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
  int baseoff_mode;                    // 1: descriptor base_offset = start address bits 7..9
  int slab_bytes, wstages, resident;
  int tiles_per_utt, n_work;
  uint32_t a_stage_bytes, a_off, w_off, bias_off, bar_off;
  const __half* x16;
  int x_Lp;
  const __half* w16;
  const float* bias;
  const float* res32;
  const __half* res16;                 // residual taken from an (activated) fp16 plane instead: x = y >= 0 ? y : y * res_inv
  int res_Lp;
  int res_hilo;                        // 1: the residual plane is a hi/lo plane (2 x Cout channels, 64-channel row chunks): x = inv_lrelu(hi + lo)
  float res_inv;
  float* y32;
  __half* y16;
  int y_Lp;
  int y_nchunks, y_cw, y_c0;           // fp16 destination plane: row chunks per utterance, channels per row chunk, first channel
                                       // this launch writes
  int y_lo_c;                          // >= 0: hi/lo destination plane - lo = fp16(v - fp16(v)) goes to channel + y_lo_c
  int x_pchunks;                       // 64-channel chunks of the INPUT plane per utterance (K-chunk c reads chunk c % x_pchunks)
  float acc_scale;                     // accumulator -> value (1, or 1/kX3WScale for 3-term-split layers)
  float out_slope;
  int mode;
  int red_add;                         // EPI_ADD without fp16 output: S += v by red.global.add.v4.f32
  float div;
  const int32_t* lengths;
  int len_mul_out;
};

using namespace tcdev;

struct WorkItem {
  int b, m0, r;
};
__device__ __forceinline__ WorkItem decode_work(const TcParams& p, int work) {
  WorkItem w;
  w.r = work % p.stride;
  const int t = work / p.stride;
  w.m0 = (t % p.tiles_per_utt) * (p.MT * 128);
  w.b = t / p.tiles_per_utt;
  return w;
}

// N = Cout, MT = row tiles per work item, CW = channels per operand row (64: SWIZZLE_128B, 32: SWIZZLE_64B).
// They are compile-time so that every MMA's descriptor is "base + immediate" (the single issuing
// thread is otherwise the bottleneck: ~190 cycles per MMA with run-time address arithmetic).
// EW epilogue warps (8, or 16 working on UC = 16-column units so that 576 threads fit the register file): see gan_tc_pair.cu.
template <int N, int MT, int CW, int EW = 8, int UC = 32>
__global__ void __launch_bounds__(tc_threads(EW), 1) tc_conv_kernel(const __grid_constant__ TcParams p) {
  constexpr int kEpiWarps = EW;
  constexpr int kTcThreads = tc_threads(EW);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms need 1024 B alignment
  uint8_t* a_base = smem + p.a_off;
  uint8_t* w_base = smem + p.w_off;
  float* bias_s = reinterpret_cast<float*>(smem + p.bias_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kAStages;
  uint64_t* w_full = a_empty + kAStages;
  uint64_t* w_empty = w_full + kMaxWStages;
  uint64_t* acc_full = w_empty + kMaxWStages;
  uint64_t* acc_empty = acc_full + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kAStages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kMaxWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 32 * kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.Cout; i += kTcThreads) bias_s[i] = p.bias ? p.bias[i] : 0.f;  // weights: never written by a kernel
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, bias) overlapped the
  // tail of the previous layer's kernel; activations written by it may only be touched after this wait.
  // Dependents of THIS kernel may begin their own prologue as soon as SMs free up.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");


  if (warp == 0) {
    // ===================== copy producer =====================
    if (lane == 0) {
      int a_stage = 0, a_phase = 0, w_stage = 0, w_phase = 0;
      uint32_t resident_loaded = 0;
      for (int work = blockIdx.x; work < p.n_work; work += gridDim.x) {
        const WorkItem wi = decode_work(p, work);
        for (int c = 0; c < p.n_cchunks; ++c) {
          mbar_wait(&a_empty[a_stage], a_phase ^ 1);
          const uint32_t bytes = (uint32_t)(p.W * p.row_bytes);
          mbar_expect_tx(&a_full[a_stage], bytes);
          const int row0 = (kPadRows + wi.m0 + p.omin) & ~7;  // multiple of 8: swizzle phases line up
          const __half* src = p.x16 + (((size_t)wi.b * p.x_pchunks + (c % p.x_pchunks)) * p.x_Lp + (size_t)row0) * p.cw;
          bulk_g2s(smem_u32(a_base + (size_t)a_stage * p.a_stage_bytes), src, bytes, &a_full[a_stage]);
          if (++a_stage == kAStages) { a_stage = 0; a_phase ^= 1; }
          for (int t = 0; t < p.ntaps[wi.r]; ++t) {
            const int sid = p.slab[wi.r][t] * p.n_cchunks + c;
            const __half* src = p.w16 + (size_t)sid * (p.slab_bytes >> 1);
            if (p.resident) {
              if (!(resident_loaded & (1u << sid))) {
                resident_loaded |= (1u << sid);
                mbar_expect_tx(&w_full[sid], (uint32_t)p.slab_bytes);
                bulk_g2s(smem_u32(w_base + (size_t)sid * p.slab_bytes), src, (uint32_t)p.slab_bytes, &w_full[sid]);
              }
            } else {
              mbar_wait(&w_empty[w_stage], w_phase ^ 1);
              mbar_expect_tx(&w_full[w_stage], (uint32_t)p.slab_bytes);
              bulk_g2s(smem_u32(w_base + (size_t)w_stage * p.slab_bytes), src, (uint32_t)p.slab_bytes,
                       &w_full[w_stage]);
              if (++w_stage == p.wstages) { w_stage = 0; w_phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop in lock-step; one elected lane issues the tcgen05 instructions.
    // Descriptors differ only in their low word (start address >> 4) and all per-MMA offsets are
    // immediates.
    {
      constexpr uint32_t ROWB = CW * 2;
      constexpr int NK16 = CW / 16;
      constexpr uint32_t MT_STEP = (128u * ROWB) >> 4;
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int a_stage = 0, a_phase = 0, w_stage = 0, w_phase = 0, acc_stage = 0, acc_phase = 0;
      const uint64_t desc_hi = make_desc(0, 8u * ROWB, CW == 64 ? 2u : 4u, 0);  // everything but the address
      const bool leader = elect_one();
      uint32_t resident_seen = 0;  // resident weight images whose arrival has already been observed
      for (int work = blockIdx.x; work < p.n_work; work += gridDim.x) {
        const WorkItem wi = decode_work(p, work);
        mbar_wait(&acc_empty[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const int delta = (kPadRows + wi.m0 + p.omin) & 7;  // rows the window start was rounded down by
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc_stage * MT * N);
        const int nt = p.ntaps[wi.r];
        bool first = true;
        for (int c = 0; c < p.n_cchunks; ++c) {
          mbar_wait(&a_full[a_stage], a_phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_base + (size_t)a_stage * p.a_stage_bytes);
          for (int t = 0; t < nt; ++t) {
            uint32_t w_addr;
            if (p.resident) {
              const int sid = p.slab[wi.r][t] * p.n_cchunks + c;
              if (!(resident_seen & (1u << sid))) {
                mbar_wait(&w_full[sid], 0);
                tc_fence_after();
                resident_seen |= 1u << sid;
              }
              w_addr = smem_u32(w_base + (size_t)sid * p.slab_bytes);
            } else {
              mbar_wait(&w_full[w_stage], w_phase);
              tc_fence_after();
              w_addr = smem_u32(w_base + (size_t)w_stage * p.slab_bytes);
            }
            const uint64_t a0 = desc_hi + (uint64_t)((a_addr + (uint32_t)(delta + p.off[wi.r][t] - p.omin) * ROWB) >> 4);
            const uint64_t b0 = desc_hi + (uint64_t)(w_addr >> 4);
            if (leader) {
              if (first) {
#pragma unroll
                for (int s = 0; s < NK16; ++s)
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt)
                    tc_mma_f16(d_tmem + (uint32_t)(mt * N), a0 + (uint64_t)(2 * s + mt * MT_STEP), b0 + (uint64_t)(2 * s),
                               idesc, s > 0 ? 1u : 0u);
              } else {
#pragma unroll
                for (int s = 0; s < NK16; ++s)
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt)
                    tc_mma_f16(d_tmem + (uint32_t)(mt * N), a0 + (uint64_t)(2 * s + mt * MT_STEP), b0 + (uint64_t)(2 * s),
                               idesc, 1u);
              }
              if (!p.resident) tc_commit(&w_empty[w_stage]);
            }
            first = false;
            if (!p.resident) {
              if (++w_stage == p.wstages) { w_stage = 0; w_phase ^= 1; }
            }
          }
          if (leader) tc_commit(&a_empty[a_stage]);
          if (++a_stage == kAStages) { a_stage = 0; a_phase ^= 1; }
        }
        if (leader) tc_commit(&acc_full[acc_stage]);
        if (++acc_stage == kAccStages) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two groups of four warps; a group covers all four TMEM lane quarters and takes every other
    // (row tile, 32-column) unit.  Per unit the residual / running-sum loads are issued FIRST (they do
    // not depend on the accumulator, so for the first unit they overlap the MMAs of this work item),
    // then the accumulator is read from TMEM, then everything is stored: 16 independent 16-byte
    // loads in flight per thread instead of one.
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
    const int grp = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    int acc_stage = 0, acc_phase = 0;
    const int C4 = p.Cout >> 2;
    const int ocw = f16_cw(p.Cout);  // output plane: channels per row chunk
    const int units_per_tile = p.Cout / UC;
    const int n_units = p.MT * units_per_tile;
    constexpr int G = EW / 4;
    for (int work = blockIdx.x; work < p.n_work; work += gridDim.x) {
      const WorkItem wi = decode_work(p, work);
      const int valid_out = p.lengths ? min(p.Lout, p.lengths[wi.b] * p.len_mul_out) : p.Lout;
      bool waited = false;
      for (int u = grp; u < n_units; u += G) {
        const int mt = u / units_per_tile;
        const int col0 = (u - mt * units_per_tile) * UC;
        const int q = wi.m0 + mt * 128 + row_in_tile;
        const int lo = q * p.stride + wi.r;
        const bool inb = q < p.Lin;
        const bool live = inb && lo < valid_out;
        const size_t i32 = ((size_t)wi.b * C4 + (col0 >> 2)) * p.Lout + lo;  // + g * Lout per 4 channels
        float4 rv[UC / 4], ov[UC / 4];
        uint4 rh[UC / 8];
        if (inb && p.res32) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g) rv[g] = reinterpret_cast<const float4*>(p.res32)[i32 + (size_t)g * p.Lout];
        }
        if (inb && p.res16) {  // (the lo halves of a hi/lo residual share rv's registers: res32 and res16 are exclusive)
          const int rr = kPadRows + lo;
          if (!p.res_hilo) {
            const size_t rbase = (((size_t)wi.b * (p.Cout / ocw) + col0 / ocw) * p.res_Lp + rr) * (size_t)(ocw >> 3);
            const int c0 = (col0 & (ocw - 1)) >> 3, sw = f16_swz(ocw, rr);
#pragma unroll
            for (int g = 0; g < UC / 8; ++g) rh[g] = reinterpret_cast<const uint4*>(p.res16)[rbase + (size_t)((c0 + g) ^ sw)];
          } else {
            // hi/lo plane: rows of 64 channels; hi block = channels [0, Cout), lo block = [Cout, 2 Cout)
            const int nch = (2 * p.Cout) >> 6, sw = f16_swz(64, rr);
#pragma unroll
            for (int g = 0; g < UC / 8; ++g) {
              const int ch = col0 + 8 * g, cl = ch + p.Cout;
              rh[g] = reinterpret_cast<const uint4*>(p.res16)[(((size_t)wi.b * nch + (ch >> 6)) * p.res_Lp + rr) * 8 + (size_t)(((ch & 63) >> 3) ^ sw)];
              rv[g] = reinterpret_cast<const float4*>(p.res16)[(((size_t)wi.b * nch + (cl >> 6)) * p.res_Lp + rr) * 8 + (size_t)(((cl & 63) >> 3) ^ sw)];
            }
          }
        }
        if (inb && p.mode != EPI_STORE && !p.red_add) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g) ov[g] = reinterpret_cast<const float4*>(p.y32)[i32 + (size_t)g * p.Lout];
        }
        if (!waited) {
          mbar_wait(&acc_full[acc_stage], acc_phase);
          tc_fence_after();
          waited = true;
        }
        uint32_t raw[UC];
        tmem_ld<UC>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc_stage * p.MT + mt) * p.Cout + col0), raw);
        if (!inb) continue;
        float v[UC];
#pragma unroll
        for (int i = 0; i < UC; ++i) v[i] = __uint_as_float(raw[i]) * p.acc_scale + bias_s[col0 + i];
        if (p.res32) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g) {
            v[4 * g + 0] += rv[g].x; v[4 * g + 1] += rv[g].y; v[4 * g + 2] += rv[g].z; v[4 * g + 3] += rv[g].w;
          }
        }
        if (p.res16) {
#pragma unroll
          for (int g = 0; g < UC / 8; ++g) {
            if (p.res_hilo) add_res16_hilo(&v[8 * g], rh[g], *reinterpret_cast<const uint4*>(&rv[g]), p.res_inv);
            else add_res16(&v[8 * g], rh[g], p.res_inv);
          }
        }
        if (p.mode != EPI_STORE && !p.red_add) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g) {
            v[4 * g + 0] += ov[g].x; v[4 * g + 1] += ov[g].y; v[4 * g + 2] += ov[g].z; v[4 * g + 3] += ov[g].w;
          }
          if (p.mode == EPI_ADD_DIV) {
#pragma unroll
            for (int i = 0; i < UC; ++i) v[i] /= p.div;
          }
        }
        if (!live) {
#pragma unroll
          for (int i = 0; i < UC; ++i) v[i] = 0.f;
        }
        if (p.red_add) {  // MRF sum of a middle resblock: S += v by vector reductions in L2 (rows past the length add nothing)
          if (live) {
#pragma unroll
            for (int g = 0; g < UC / 4; ++g)
              red_add_f32x4(p.y32 + (i32 + (size_t)g * p.Lout) * 4, v[4 * g + 0], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          }
        } else if (p.y32) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g)
            reinterpret_cast<float4*>(p.y32)[i32 + (size_t)g * p.Lout] =
                make_float4(v[4 * g + 0], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
        if (p.y16) {
          const int rr = kPadRows + lo;
          const int ysw = f16_swz(p.y_cw, rr);
          const int ycw8 = p.y_cw >> 3;
#pragma unroll
          for (int g = 0; g < UC / 8; ++g) {
            float a[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = lrelu(v[8 * g + e], p.out_slope);
            __half2 h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(a[2 * e], a[2 * e + 1]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h[0]);
            pk.y = *reinterpret_cast<uint32_t*>(&h[1]);
            pk.z = *reinterpret_cast<uint32_t*>(&h[2]);
            pk.w = *reinterpret_cast<uint32_t*>(&h[3]);
            const int ct = p.y_c0 + col0 + 8 * g;  // channel of the (hi) block
            const int chunk = ct / p.y_cw, cc = ct - chunk * p.y_cw;
            const size_t i16 = (((size_t)wi.b * p.y_nchunks + chunk) * p.y_Lp + rr) * (size_t)ycw8 + (size_t)((cc >> 3) ^ ysw);
            reinterpret_cast<uint4*>(p.y16)[i16] = pk;
            if (p.y_lo_c >= 0) {
              __half2 l2[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 hf = __half22float2(h[e]);
                l2[e] = __floats2half2_rn(a[2 * e] - hf.x, a[2 * e + 1] - hf.y);
              }
              uint4 pl;
              pl.x = *reinterpret_cast<uint32_t*>(&l2[0]);
              pl.y = *reinterpret_cast<uint32_t*>(&l2[1]);
              pl.z = *reinterpret_cast<uint32_t*>(&l2[2]);
              pl.w = *reinterpret_cast<uint32_t*>(&l2[3]);
              const int ctl = ct + p.y_lo_c;
              const int chl = ctl / p.y_cw, ccl = ctl - chl * p.y_cw;
              const size_t j16 = (((size_t)wi.b * p.y_nchunks + chl) * p.y_Lp + rr) * (size_t)ycw8 + (size_t)((ccl >> 3) ^ ysw);
              reinterpret_cast<uint4*>(p.y16)[j16] = pl;
            }
          }
        }
      }
      if (!waited) {
        mbar_wait(&acc_full[acc_stage], acc_phase);
        tc_fence_after();
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[acc_stage]);
      if (++acc_stage == kAccStages) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// fp32 slabs [K][Cin][Cout] -> fp16 images [K][n_cchunks][Cout][cw], rows swizzled like the operand planes
__global__ void pack_w16_kernel(const float* __restrict__ w32, __half* __restrict__ dst, int K, int Cin, int Cout,
                                int cw) {
  const size_t n = (size_t)K * Cin * Cout;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // source-linear index -> (k, ci, co)
  const int co = (int)(i % Cout);
  const int ci = (int)((i / Cout) % Cin);
  const int k = (int)(i / ((size_t)Cout * Cin));
  const int nch = Cin / cw;
  const int c = ci / cw, cc = ci - c * cw;
  const size_t img = ((size_t)k * nch + c) * (size_t)Cout * cw;
  const size_t off = (size_t)co * cw + (size_t)((((cc >> 3) ^ f16_swz(cw, co)) << 3) + (cc & 7));
  dst[img + off] = __float2half_rn(w32[i]);
}

// 3-term-split layer over an internal hi/lo plane: K-chunks [hi | lo | hi] x images [hi(w) | hi(w) | lo(w)] (Cin >= 64), or for
// Cin == 32 the single 64-channel plane chunk [hi32 | lo32] twice: images [hi(w) | hi(w)] and [lo(w) | 0].  w is pre-scaled by
// kX3WScale (power of two; the epilogue multiplies the accumulator by its inverse) so that lo(w) is a normal fp16.
__global__ void pack_w16_x3_kernel(const float* __restrict__ w32, __half* __restrict__ dst, int K, int Cin, int Cout, float wscale) {
  const int nK = Cin >= 64 ? 3 * (Cin / 64) : 2;
  const size_t n = (size_t)K * nK * Cout * 64;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cc = (int)(i & 63);
  const int co = (int)((i >> 6) % Cout);
  const int j = (int)((i / ((size_t)64 * Cout)) % nK);
  const int k = (int)(i / ((size_t)64 * Cout * nK));
  int ci, part;  // part 0: hi(w), 1: lo(w), 2: zero
  if (Cin >= 64) {
    const int nc = Cin / 64;
    part = j < 2 * nc ? 0 : 1;
    ci = (j % nc) * 64 + cc;
  } else {
    ci = cc & 31;
    part = j == 0 ? 0 : (cc < 32 ? 1 : 2);
  }
  __half v = __float2half_rn(0.f);
  if (part < 2) {
    const float w = w32[((size_t)k * Cin + ci) * Cout + co] * wscale;
    const __half hi = __float2half_rn(w);
    v = part == 0 ? hi : __float2half_rn(w - __half2float(hi));
  }
  const size_t img = ((size_t)k * nK + j) * (size_t)Cout * 64;
  const size_t off = (size_t)co * 64 + (size_t)((((cc >> 3) ^ f16_swz(64, co)) << 3) + (cc & 7));
  dst[img + off] = v;
}

// ---- 3-term split of an fp32 layer (conv_pre) --------------------------------------------------------
// input plane channels: [hi(x) (Cin) | lo(x) (Cin) | hi(x) (Cin) | 0 ...] (256 channels, no activation),
// weight images:        [hi(w)       | hi(w)       | lo(w)       | 0 ...]  -> sum = x*w up to ~2^-22 relative.
__device__ __forceinline__ __half split_hi(float v) { return __float2half_rn(v); }
__device__ __forceinline__ __half split_lo(float v) { return __float2half_rn(v - __half2float(__float2half_rn(v))); }

__global__ void split3_input_kernel(const float* __restrict__ x /* NCL [B][Cin][L] */, __half* __restrict__ dst, int B,
                                    int Cin, int L, const int32_t* __restrict__ lengths) {
  // one thread per (b, 16-byte chunk of the 256-channel row, l); l fastest -> coalesced reads of x
  const size_t n = (size_t)B * 32 * L;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = (int)(i % L);
  const int ch8 = (int)((i / L) % 32);
  const int b = (int)(i / ((size_t)L * 32));
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = ch8 * 8 + e;
    const int part = c / Cin, ci = c - part * Cin;
    float v = 0.f;
    if (part < 3 && (!lengths || l < lengths[b])) v = x[((size_t)b * Cin + ci) * L + l];  // frames past the length read as zero padding
    h[e] = part == 1 ? split_lo(v) : (part < 3 ? split_hi(v) : __float2half_rn(0.f));
  }
  const int Lp = f16_lp(L);
  const int r = kPadRows + l;
  const int chunk = ch8 >> 3, cc8 = ch8 & 7;
  const size_t o = (((size_t)b * 4 + chunk) * Lp + r) * 8 + (size_t)(cc8 ^ f16_swz(64, r));
  reinterpret_cast<uint4*>(dst)[o] = *reinterpret_cast<const uint4*>(h);
}

// fp32 slabs [K][Cin][Cout] -> per 256-output half: images [K][4 chunks][256][64] (swizzled rows)
__global__ void pack_w16_split3_kernel(const float* __restrict__ w32, __half* __restrict__ dst, int K, int Cin, int Cout) {
  const size_t n = (size_t)K * 256 * Cout;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int co = (int)(i % Cout);
  const int c = (int)((i / Cout) % 256);
  const int k = (int)(i / ((size_t)Cout * 256));
  const int part = c / Cin, ci = c - part * Cin;
  __half v = __float2half_rn(0.f);
  if (part < 3) {
    const float w = w32[((size_t)k * Cin + ci) * Cout + co];
    v = part == 2 ? split_lo(w) : split_hi(w);
  }
  const int half_idx = co >> 8, col = co & 255;
  const int chunk = c >> 6, cc = c & 63;
  const size_t img = (((size_t)half_idx * K + k) * 4 + chunk) * (size_t)(256 * 64);
  const size_t off = (size_t)col * 64 + (size_t)((((cc >> 3) ^ f16_swz(64, col)) << 3) + (cc & 7));
  dst[img + off] = v;
}

bool tc_split3_enabled();
bool split3_capable(const TapConv& t) {
  return tc_split3_enabled() && t.stride == 1 && !t.act_tanh && t.Cin * 3 <= 256 && t.Cout % 256 == 0 && t.Cout >= 256 &&
         t.mode == EPI_STORE;
}

int pick_kc(int Cin) {
  if (Cin % 64 == 0) return 64;
  if (Cin == 32) return 32;
  return 0;
}

// shapes the 3-term-split variant of tc_conv_kernel covers (operand rows are always 64 channels wide)
bool x3_capable(const TapConv& t) {
  if (t.Cout != 32 && t.Cout != 64 && t.Cout != 128 && t.Cout != 256) return false;
  if (t.Cin != 32 && t.Cin % 64 != 0) return false;
  return !t.act_tanh;
}

bool tc_capable(const TapConv& t) {
  if (t.Cout != 32 && t.Cout != 64 && t.Cout != 128 && t.Cout != 256) return false;  // kernel instances
  if (t.Cout != 32 && pick_kc(t.Cin) != 64) return false;
  if (pick_kc(t.Cin) == 0) return false;
  if (t.act_tanh) return false;
  return true;
}

// row tiles per work item: as many as TMEM (2 x MT x Cout <= 512 columns) and shared memory allow
int pick_mt_max(int Cout) { return Cout >= 256 ? 1 : (Cout >= 128 ? 2 : (Cout >= 64 ? 4 : 8)); }

struct SmemPlan {
  uint32_t a_stage_bytes, a_off, w_off, bias_off, bar_off, total;
  int wstages, resident, W, omin;
};

bool plan_smem(const TapConv& t, const TcLayer& tc, int total_slabs, SmemPlan* sp) {
  int omin = 0x7fffffff, omax = -0x7fffffff;
  for (int r = 0; r < t.stride; ++r)
    for (int i = 0; i < t.ntaps[r]; ++i) {
      omin = std::min(omin, t.off[r][i]);
      omax = std::max(omax, t.off[r][i]);
    }
  if (-omin > kPadRows || omax > kPadRows) return false;
  const int row_bytes = tc.kc * 2;
  sp->omin = omin;
  sp->W = (tc.mt * 128 + (omax - omin) + 7 + 7) & ~7;  // + up to 7 rows of start rounding, multiple of 8
  sp->a_stage_bytes = (uint32_t)align_up((size_t)sp->W * row_bytes, 1024);
  sp->a_off = 0;
  sp->w_off = sp->a_off + kAStages * sp->a_stage_bytes;
  const uint32_t tail = (uint32_t)align_up((size_t)t.Cout * 4, 128) + 1024;
  const uint32_t usable = kSmemMax - 1024;  // the kernel aligns its base to 1024 B
  if (sp->w_off + tail + 2 * tc.slab_bytes > usable) return false;
  int ws = (int)((usable - sp->w_off - tail) / tc.slab_bytes);
  ws = std::min(ws, kMaxWStages);
  sp->resident = (total_slabs <= ws) ? 1 : 0;
  sp->wstages = sp->resident ? total_slabs : ws;
  sp->bias_off = sp->w_off + (uint32_t)sp->wstages * (uint32_t)tc.slab_bytes;
  sp->bar_off = sp->bias_off + (uint32_t)align_up((size_t)t.Cout * 4, 128);
  sp->total = sp->bar_off + 1024;
  return sp->total <= usable;
}

int kernel_count(const TapConv& t) {
  int k = 0;
  for (int r = 0; r < t.stride; ++r)
    for (int i = 0; i < t.ntaps[r]; ++i) k = std::max(k, t.slab[r][i] + 1);
  return k;
}

// MB_TC_FUSE=0 disables the fused resblock-pair kernel (gan_tc_pair.cu) for A/B measurements
bool tc_fuse_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MB_TC_FUSE");
    on = e ? atoi(e) : 1;
  }
  return on != 0;
}

// MB_TC_RES16=0 keeps an fp32 residual plane in every stage (default: only in the full-rate stage; the
// other stages carry the residual stream as ONE activated fp16 plane that is both the next conv's operand
// and - through the exact inverse of the leaky-relu - the residual; error budget in DESIGN.md 3.4)
bool tc_res16_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MB_TC_RES16");
    on = e ? atoi(e) : 1;
  }
  return on != 0;
}

// MB_TC_X3_RES16=0: 3-term-split layers keep an fp32 residual plane (default: residual = inv_lrelu(hi + lo) of the hi/lo plane)
bool tc_x3_res16_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MB_TC_X3_RES16");
    on = e ? atoi(e) : 1;
  }
  return on != 0;
}

// MB_TC_PAIR32=0 disables the fp32-input pair kernel of the full-rate stage (falls back to the fp16-plane pair)
bool tc_pair32_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MB_TC_PAIR32");
    on = e ? atoi(e) : 1;
  }
  return on != 0;
}

// MB_TC_SPLIT3=0 keeps conv_pre on the FP32 FFMA kernel
bool tc_split3_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MB_TC_SPLIT3");
    on = e ? atoi(e) : 1;
  }
  return on != 0;
}

// how the A descriptor encodes a start address that is not 1024-byte aligned (row-shifted taps).
// Measured on B200 (tests/test_gan_tc_layers.py under MB_TC_BASEOFF=0/1): the operand fetch applies the
// swizzle XOR on absolute shared-memory address bits, so the matrix-base-offset field must stay 0
// (mode 0, default); mode 1 (field = address bits 7..9) double-counts the phase and is wrong.
int tc_baseoff_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("MB_TC_BASEOFF");
    mode = e ? atoi(e) : 0;
  }
  return mode;
}

size_t f16_plane_bytes(size_t B, size_t T, size_t cr) {
  // rows are padded to Lp = ceil8(L + 2*kPadRows) <= L + 2*kPadRows + 7 for at most 512 channels
  return align_up(2 * B * T * cr + 2 * B * 1024 * (2 * kPadRows + 7) + kPlaneSlack, 1024);  // pad rows of <= 1024 channels (hi/lo of 512)
}
size_t f32_plane_bytes(size_t B, size_t T, size_t cr) { return align_up(4 * B * T * cr + 256, 1024); }

int launch_tc(const TcOp& op, const char* tc_arena, const TRef& x16, const TRef& res32, const TRef& res16, float res_slope,
              const TRef& y32, const TRef& y16, float out_slope, const int32_t* lengths, int B, int Lin, cudaStream_t st,
              int y_c0 = 0, const float* bias_override = nullptr) {
  const TapConv& t = op.taps;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.B = B;
  p.Lin = Lin;
  p.Lout = Lin * t.stride;
  p.Cin = t.Cin;
  p.Cout = t.Cout;
  p.stride = t.stride;
  memcpy(p.ntaps, t.ntaps, sizeof(p.ntaps));
  memcpy(p.off, t.off, sizeof(p.off));
  memcpy(p.slab, t.slab, sizeof(p.slab));
  SmemPlan sp;
  const int total_slabs = kernel_count(t) * op.tc.n_cchunks;
  if (!plan_smem(t, op.tc, total_slabs, &sp)) return fail(MB_ERR_INVALID, "tc_conv(%s): shared memory plan failed", op.name);
  p.omin = sp.omin;
  p.W = sp.W;
  p.MT = op.tc.mt;
  p.cw = op.tc.kc;
  p.row_bytes = op.tc.kc * 2;
  p.layout_type = (op.tc.kc == 64) ? 2 : 4;
  p.baseoff_mode = tc_baseoff_mode();
  p.nk16 = op.tc.kc / 16;
  p.n_cchunks = op.tc.n_cchunks;
  p.slab_bytes = (int)op.tc.slab_bytes;
  p.wstages = sp.wstages;
  p.resident = sp.resident;
  p.tiles_per_utt = (Lin + p.MT * 128 - 1) / (p.MT * 128);
  p.n_work = B * p.tiles_per_utt * p.stride;
  p.a_stage_bytes = sp.a_stage_bytes;
  p.a_off = sp.a_off;
  p.w_off = sp.w_off;
  p.bias_off = sp.bias_off;
  p.bar_off = sp.bar_off;
  p.x16 = reinterpret_cast<const __half*>(x16.p);
  p.x_Lp = f16_lp(x16.L);
  p.x_pchunks = op.tc.x3 ? op.tc.x_pchunks : op.tc.n_cchunks;
  p.acc_scale = op.tc.x3 ? 1.f / kX3WScale : 1.f;
  if (op.tc.x3 && (!x16.hilo || x16.C != 64 * op.tc.x_pchunks))
    return fail(MB_ERR_INVALID, "tc_conv(%s): 3-term-split layer needs a hi/lo input plane", op.name);
  if (!op.tc.x3 && !op.tc.split3 && x16.hilo) return fail(MB_ERR_INVALID, "tc_conv(%s): plain layer fed a hi/lo plane", op.name);
  p.w16 = reinterpret_cast<const __half*>(tc_arena + op.tc.w16_off);
  p.bias = bias_override ? bias_override : op.b32;
  p.res32 = reinterpret_cast<const float*>(res32.p);
  p.res16 = reinterpret_cast<const __half*>(res16.p);
  p.res_Lp = f16_lp(res16.L);
  p.res_hilo = res16.p ? res16.hilo : 0;
  p.res_inv = 1.f / res_slope;
  p.y32 = reinterpret_cast<float*>(y32.p);
  p.y16 = reinterpret_cast<__half*>(y16.p);
  p.y_Lp = f16_lp(y16.L);
  p.y_cw = f16_cw(y16.p ? y16.C : t.Cout);
  p.y_nchunks = (y16.p ? y16.C : t.Cout) / p.y_cw;
  p.y_c0 = y_c0;
  p.y_lo_c = (y16.p && y16.hilo) ? (y16.C >> 1) : -1;
  p.out_slope = out_slope;
  p.mode = t.mode;
  p.red_add = (p.mode == EPI_ADD && !y16.p && y32.p && tc_red_add_enabled()) ? 1 : 0;
  p.div = t.div;
  p.lengths = lengths;
  p.len_mul_out = t.len_mul_out;
  if (res16.p && !res16.hilo && res16.C != t.Cout) return fail(MB_ERR_INVALID, "tc_conv(%s): fp16 residual plane geometry", op.name);
  if (res16.p && res16.hilo && res16.C != (t.Cout >= 64 ? 2 * t.Cout : 64)) return fail(MB_ERR_INVALID, "tc_conv(%s): hi/lo residual plane geometry", op.name);
  if (y_c0 != 0 && (p.y32 || p.res32 || p.res16)) return fail(MB_ERR_INVALID, "tc_conv(%s): channel-offset launch supports the fp16 plane only", op.name);
  if (p.mode != EPI_STORE && !p.y32) return fail(MB_ERR_INVALID, "tc_conv(%s): accumulate mode without fp32 plane", op.name);
  void (*kern)(const TcParams) = nullptr;
  static const int ew16 = [] {
    const char* e = getenv("MB_TC_CONV_EW16");  // A/B switch: 1 = sixteen epilogue warps on 16-column units
    return e ? atoi(e) : 0;
  }();
  int threads = tc_threads(8);
  if (ew16) {
    threads = tc_threads(16);
    if (p.Cout == 256 && p.MT == 1 && p.cw == 64) kern = tc_conv_kernel<256, 1, 64, 16, 16>;
    else if (p.Cout == 128 && p.MT == 2 && p.cw == 64) kern = tc_conv_kernel<128, 2, 64, 16, 16>;
    else if (p.Cout == 128 && p.MT == 1 && p.cw == 64) kern = tc_conv_kernel<128, 1, 64, 16, 16>;
    else if (p.Cout == 64 && p.MT == 4 && p.cw == 64) kern = tc_conv_kernel<64, 4, 64, 16, 16>;
    else if (p.Cout == 32 && p.MT == 4 && p.cw == 64) kern = tc_conv_kernel<32, 4, 64, 16, 16>;
    else threads = tc_threads(8);
  }
  if (kern) {
  } else if (p.Cout == 256 && p.MT == 1 && p.cw == 64) kern = tc_conv_kernel<256, 1, 64>;
  else if (p.Cout == 128 && p.MT == 2 && p.cw == 64) kern = tc_conv_kernel<128, 2, 64>;
  else if (p.Cout == 128 && p.MT == 1 && p.cw == 64) kern = tc_conv_kernel<128, 1, 64>;
  else if (p.Cout == 64 && p.MT == 4 && p.cw == 64) kern = tc_conv_kernel<64, 4, 64>;
  else if (p.Cout == 64 && p.MT == 2 && p.cw == 64) kern = tc_conv_kernel<64, 2, 64>;
  else if (p.Cout == 64 && p.MT == 1 && p.cw == 64) kern = tc_conv_kernel<64, 1, 64>;
  else if (p.Cout == 32 && p.MT == 8 && p.cw == 32) kern = tc_conv_kernel<32, 8, 32>;
  else if (p.Cout == 32 && p.MT == 2 && p.cw == 64) kern = tc_conv_kernel<32, 2, 64>;
  else if (p.Cout == 32 && p.MT == 4 && p.cw == 64) kern = tc_conv_kernel<32, 4, 64>;
  else if (p.Cout == 32 && p.MT == 4 && p.cw == 32) kern = tc_conv_kernel<32, 4, 32>;
  else return fail(MB_ERR_INVALID, "tc_conv(%s): no kernel instance for Cout=%d MT=%d cw=%d", op.name, p.Cout, p.MT, p.cw);
  MB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(p.n_work, sms);
  if (grid <= 0) return MB_OK;
  // always claim the whole shared memory: one CTA per SM, so the 512-column TMEM allocation never contends
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = kSmemMax;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  MB_LAUNCH_CHECK("tc_conv_kernel");
  return MB_OK;
}

TRef make_ref(void* p, int layout, int C, int L) {
  TRef t;
  t.p = p;
  t.layout = p ? layout : LAYOUT_NONE;
  t.C = C;
  t.L = L;
  return t;
}

}  // namespace

// MB_TC_RED_ADD=0: accumulate-mode epilogues read, add and store the running sum themselves (round 2)
bool tc_red_add_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MB_TC_RED_ADD");
    on = e ? atoi(e) : 1;
  }
  return on != 0;
}


int tc_plan_layers(std::vector<TcLayerDesc>& layers, size_t* tc_arena_bytes) {
  size_t off = 0;
  for (TcLayerDesc& d : layers) {
    if (!d.is_conv) continue;
    TcLayer& tc = *d.tc;
    const TapConv& t = *d.taps;
    tc = TcLayer{};
    if (!d.force_f32 && !tc_capable(t) && split3_capable(t)) {
      // conv_pre: runs as Cout/256 launches of the <256,1,64> instance over a 256-channel split plane
      tc.split3 = 1;
      tc.kc = 64;
      tc.n_cchunks = 4;
      tc.mt = 1;
      tc.slab_bytes = (size_t)64 * 256 * 2;
      tc.w16_off = off;
      off += align_up((size_t)(t.Cout / 256) * d.k * 4 * tc.slab_bytes, 256);
      continue;  // use_tc stays 0: every generic decision treats the layer as an FP32-input layer
    }
    const bool x3 = d.want_x3 && !d.force_f32 && x3_capable(t);
    if (d.force_f32 || (!x3 && !tc_capable(t))) continue;
    if (x3) {
      tc.x3 = 1;
      tc.kc = 64;
      tc.n_cchunks = t.Cin >= 64 ? 3 * (t.Cin / 64) : 2;
      tc.x_pchunks = t.Cin >= 64 ? 2 * (t.Cin / 64) : 1;
    } else {
      tc.kc = pick_kc(t.Cin);
      tc.n_cchunks = t.Cin / tc.kc;
      tc.x_pchunks = tc.n_cchunks;
    }
    tc.slab_bytes = (size_t)tc.kc * t.Cout * 2;
    SmemPlan sp;
    bool ok = false;
    for (tc.mt = pick_mt_max(t.Cout); tc.mt >= 1; tc.mt >>= 1) {
      // keep the weights resident if a smaller tile count allows it; otherwise take the largest that fits
      if (plan_smem(t, tc, d.k * tc.n_cchunks, &sp)) {
        ok = true;
        TcLayer half = tc;
        half.mt = tc.mt >> 1;
        SmemPlan sp2;
        if (!sp.resident && half.mt >= 1 && plan_smem(t, half, d.k * tc.n_cchunks, &sp2) && sp2.resident) tc.mt = half.mt;
        break;
      }
    }
    if (!ok) continue;  // falls back to the FP32 kernel
    if (t.Cout == 32 && tc.kc == 64 && tc.mt > 4) tc.mt = 4;  // instance list below
    plan_smem(t, tc, d.k * tc.n_cchunks, &sp);
    if (sp.resident && d.k * tc.n_cchunks > 32) continue;
    tc.use_tc = 1;
    tc.w16_off = off;
    off += align_up((size_t)d.k * tc.n_cchunks * tc.slab_bytes, 256);
  }
  *tc_arena_bytes = off;
  return MB_OK;
}

int tc_pack_weights(const TcLayer& tc, const TapConv& taps, const float* w32_slabs, char* tc_arena,
                    cudaStream_t stream) {
  if (tc.split3) {
    const int K = kernel_count(taps);
    const size_t n = (size_t)K * 256 * taps.Cout;
    pack_w16_split3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        w32_slabs, reinterpret_cast<__half*>(tc_arena + tc.w16_off), K, taps.Cin, taps.Cout);
    MB_LAUNCH_CHECK("pack_w16_split3_kernel");
    return MB_OK;
  }
  if (!tc.use_tc) return MB_OK;
  if (tc.x3) {
    const int K = kernel_count(taps);
    const size_t n = (size_t)K * tc.n_cchunks * taps.Cout * 64;
    pack_w16_x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w32_slabs, reinterpret_cast<__half*>(tc_arena + tc.w16_off), K,
                                                                      taps.Cin, taps.Cout, kX3WScale);
    MB_LAUNCH_CHECK("pack_w16_x3_kernel");
    return MB_OK;
  }
  const int K = kernel_count(taps);
  const size_t n = (size_t)K * taps.Cin * taps.Cout;
  pack_w16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
      w32_slabs, reinterpret_cast<__half*>(tc_arena + tc.w16_off), K, taps.Cin, taps.Cout, tc.kc);
  MB_LAUNCH_CHECK("pack_w16_kernel");
  return MB_OK;
}

size_t tc_workspace_bytes(const std::vector<TcBufReq>& bufs, int B, int T, int num_mels, int hop) {
  (void)num_mels;
  (void)hop;
  size_t total = 2048;
  for (const TcBufReq& b : bufs) total += f16_plane_bytes(B, T, 2 * b.cr) + f32_plane_bytes(B, T, b.cr);  // 2x: hi/lo planes
  return total;
}

int tc_forward(const std::vector<TcOp>& ops, const std::vector<TcBufReq>& bufs, const char* tc_arena,
               const float* mel, const int32_t* lengths, int B, int T, int num_mels, int hop, float* wav,
               void* workspace, cudaStream_t st, cudaEvent_t* events) {
  (void)hop;
  constexpr int BUF_IN = 100, BUF_OUT = 101;
  const int nb = (int)bufs.size();
  // carve planes
  std::vector<char*> p16(nb), p32(nb);
  std::vector<TRef> cur16(nb), cur32(nb);  // geometry each plane currently holds
  char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  for (int i = 0; i < nb; ++i) {
    p16[i] = ws;
    ws += f16_plane_bytes(B, T, 2 * bufs[i].cr);
    p32[i] = ws;
    ws += f32_plane_bytes(B, T, bufs[i].cr);
  }
  // buffer -> fp16 plane storage.  The fused pair kernel must not write the fp16 plane it is still reading
  // halos from (other CTAs), so it ping-pongs between the buffer's storage and the unused storage of the
  // pair's intermediate buffer; map16 tracks where each buffer's current fp16 plane lives.
  std::vector<int> map16(nb);
  for (int i = 0; i < nb; ++i) map16[i] = i;
  std::vector<int> map32(nb);  // same indirection for the fp32 planes (fp32-input pair kernel, in-place pairs)
  for (int i = 0; i < nb; ++i) map32[i] = i;
  std::vector<float> plane_slope(nb, 1.f);  // leaky-relu slope each fp16 storage was written with
  const int n = (int)ops.size();
  int full_rate = 1;
  for (const TcOp& o : ops) full_rate = std::max(full_rate, o.rate_out);
  // residual taken from the activated fp16 plane (no fp32 residual plane) in every stage but the full-rate one
  auto res16_ok = [&](const TcOp& c) {
    // 3-term-split layers take the residual from the hi/lo plane their pair's first conv reads anyway (hi + lo carries 22 bits:
    // FP32-equivalent), in every stage; plain fp16 layers from the activated fp16 plane in every stage but the full-rate one
    if (!(tc_res16_enabled() && c.is_conv && c.tc.use_tc && c.res >= 0 && c.res < nb && c.taps.stride == 1)) return false;
    return c.tc.x3 ? tc_x3_res16_enabled() : c.rate_out < full_rate;
  };
  // does a 3-term-split layer consume buffer `buf` (scanning forward from op `from` until the buffer is overwritten)?
  auto wants_hilo = [&](int buf, int from) {
    for (int j = from; j < n; ++j) {
      const TcOp& c = ops[j];
      if (c.is_conv && c.src == buf && c.tc.use_tc && c.tc.x3) return true;
      if (c.is_conv && c.dst == buf) break;  // the next writer (store or accumulate) produces the plane its own consumers need
    }
    return false;
  };
  // ---- pre-pass: which (c1, c2) op pairs run as ONE fused kernel (gan_tc_pair.cu)?
  //   kind 1: fp16-plane input;  kind 2: fp32-plane input (full-rate stage, residual = the pair's input)
  std::vector<int> fuse_kind(n, 0);
  std::vector<TcPairParams> pair_plan(n);
  for (int i = 0; i + 1 < n; ++i) {
    memset(&pair_plan[i], 0, sizeof(TcPairParams));
    const TcOp& op = ops[i];
    if (!tc_fuse_enabled() || !op.is_conv || !op.tc.use_tc || !ops[i + 1].is_conv || !ops[i + 1].tc.use_tc) continue;
    if (op.tc.x3 || ops[i + 1].tc.x3) continue;                 // 3-term-split layers run unfused (tc_conv_kernel)
    if (wants_hilo(ops[i + 1].dst, i + 2) && op.cout != 64) continue;  // the pair kernel writes hi/lo planes for C = 64 only
    const TcOp& c2 = ops[i + 1];
    const TapConv& t1 = op.taps;
    const TapConv& t2 = c2.taps;
    const int k = t1.ntaps[0];
    const int d1 = k > 1 ? t1.off[0][1] - t1.off[0][0] : 1;
    bool ok = op.cin == op.cout && c2.cin == c2.cout && op.cin == c2.cin && t1.stride == 1 && t2.stride == 1 &&
              t2.ntaps[0] == k && (k & 1) && op.res < 0 && op.dst2 < 0 && c2.dst2 < 0 && t1.mode == EPI_STORE &&
              op.dst == c2.src && op.dst >= 0 && op.dst < nb && op.src >= 0 && op.src < nb && c2.res != op.dst &&
              op.rate_in == c2.rate_in;
    for (int t = 0; ok && t < k; ++t)
      ok = (t1.off[0][t] == t * d1 - d1 * (k - 1) / 2) && (t2.off[0][t] == t - (k - 1) / 2) && t1.slab[0][t] == t &&
           t2.slab[0][t] == t;
    // the intermediate buffer must not be read by anything but c2 before it is overwritten
    for (int j = i + 2; ok && j < n; ++j) {
      const TcOp& c = ops[j];
      if (c.src == op.dst || (c.is_conv && (c.res == op.dst || c.dst2 == op.dst))) ok = false;
      if (c.is_conv && c.dst == op.dst && c.taps.mode == EPI_STORE) break;
    }
    if (!ok) continue;
    const bool want32 = tc_pair32_enabled() && c2.res == op.src && !res16_ok(c2);
    if (want32 && tc_pair_plan(op.cin, k, d1, true, &pair_plan[i])) fuse_kind[i] = 2;
    else if (tc_pair_plan(op.cin, k, d1, false, &pair_plan[i])) fuse_kind[i] = 1;
    if (fuse_kind[i]) ++i;  // c2 belongs to this pair
  }
  auto reads_f32 = [&](int j) { return fuse_kind[j] == 2; };  // op j (a c1) takes its operand from the fp32 plane
  for (int i = 0; i < n; ++i) {
    const TcOp& op = ops[i];
    if (events) MB_CUDA_CHECK(cudaEventRecord(events[i], st));
    const int Lin = T * op.rate_in, Lout = T * op.rate_out;
    if (!op.is_conv) {
      // dst32 += src32 ; refresh dst16 if a tensor-core consumer follows
      if (op.dst < 0 || op.dst >= nb || op.src < 0 || op.src >= nb) return fail(MB_ERR_INVALID, "tc_forward: bad add op");
      bool need16 = false;
      float slope16 = 1.f;
      for (int j = i + 1; j < n; ++j) {
        if (ops[j].is_conv && ops[j].src == op.dst && ops[j].tc.use_tc && !reads_f32(j)) { need16 = true; slope16 = ops[j].taps.in_slope; }
        if (ops[j].is_conv && ops[j].dst == op.dst) break;
      }
      TRef d32 = make_ref(p32[map32[op.dst]], LAYOUT_F32B, op.cout, Lout);
      TRef s32 = make_ref(p32[map32[op.src]], LAYOUT_F32B, op.cout, Lout);
      const bool add_hilo = need16 && wants_hilo(op.dst, i + 1);
      TRef d16 = need16 ? make_ref(p16[map16[op.dst]], LAYOUT_F16B, op.cout * (add_hilo ? 2 : 1), Lout) : TRef{};
      d16.hilo = add_hilo ? 1 : 0;
      if (need16) {
        plane_slope[map16[op.dst]] = slope16;
        if (cur16[map16[op.dst]].C != d16.C || cur16[map16[op.dst]].L != Lout || cur16[map16[op.dst]].hilo != d16.hilo) {
          cudaError_t ez = launch_zero_pads_f16(d16, B, st);
          if (ez != cudaSuccess) return fail(MB_ERR_CUDA, "zero_pads: %s", cudaGetErrorString(ez));
          count_launch();
          cur16[map16[op.dst]] = d16;
        }
      }
      cudaError_t e = launch_add_inplace_f32(d32, s32, d16, slope16, B, st);
      if (e != cudaSuccess) return fail(MB_ERR_CUDA, "add kernel: %s", cudaGetErrorString(e));
      count_launch();
      continue;
    }
    // ---- fused resblock pair (c1 -> T -> c2), decided in the pre-pass above
    TcPairParams pp = pair_plan[i];
    const bool fuse = fuse_kind[i] != 0;
    const TcOp& oop = fuse ? ops[i + 1] : op;  // the op whose outputs this launch produces
    const char* fused_x16 = fuse ? p16[map16[op.src]] : nullptr;
    const bool use_res16 = res16_ok(oop);
    TRef res16 = use_res16 ? make_ref(p16[map16[oop.res]], LAYOUT_F16B, oop.cout, Lout) : TRef{};
    if (use_res16 && cur16[map16[oop.res]].hilo) {  // the residual buffer currently holds a hi/lo plane (3-term-split consumers)
      res16.C = cur16[map16[oop.res]].C;
      res16.hilo = 1;
    }
    const float res_slope = use_res16 ? plane_slope[map16[oop.res]] : 1.f;
    const TRef res32 = (oop.res >= 0 && !use_res16) ? make_ref(p32[map32[oop.res]], LAYOUT_F32B, oop.cout, Lout) : TRef{};
    const char* fused_x32 = (fuse_kind[i] == 2) ? p32[map32[op.src]] : nullptr;
    if (fuse_kind[i] == 2 && oop.dst == op.src) std::swap(map32[oop.dst], map32[op.dst]);  // write the other fp32 storage
    if (fuse && oop.dst == op.src) std::swap(map16[oop.dst], map16[op.dst]);  // write the other storage
    const int scan_from = fuse ? i + 2 : i + 1;
    // ---- which planes must this op produce? (scan the consumers of dst until it is overwritten)
    bool need16 = false, need32 = false;
    float slope16 = 1.f;
    bool have_slope = false;
    auto scan = [&](int buf, bool& n16, bool& n32, float& s16, bool& hs) -> int {
      for (int j = scan_from; j < n; ++j) {
        const TcOp& c = ops[j];
        if (c.is_conv) {
          if (c.src == buf) {
            if (reads_f32(j)) {
              n32 = true;
            } else if (c.tc.use_tc) {
              if (hs && s16 != c.taps.in_slope) return fail(MB_ERR_INVALID, "tc_forward: consumers of one buffer disagree on slope");
              n16 = true;
              s16 = c.taps.in_slope;
              hs = true;
            } else {
              n32 = true;
            }
          }
          if (c.res == buf) {
            if (res16_ok(c)) n16 = true;
            else n32 = true;
          }
          if (c.dst2 == buf) n32 = true;
          if (c.dst == buf) {
            if (c.taps.mode != EPI_STORE) n32 = true;
            break;
          }
        } else {
          if (c.src == buf || c.dst == buf) n32 = true;
        }
      }
      return MB_OK;
    };
    TRef y32, y16, y2_32, y2_16;
    float y2_slope = 1.f;
    if (oop.dst == BUF_OUT) {
      y32 = make_ref(wav, LAYOUT_NCL, oop.cout, Lout);
    } else {
      if (oop.dst < 0 || oop.dst >= nb) return fail(MB_ERR_INVALID, "tc_forward: bad dst");
      int rc = scan(oop.dst, need16, need32, slope16, have_slope);
      if (rc != MB_OK) return rc;
      if (oop.taps.mode != EPI_STORE) need32 = true;
      if (need32) y32 = make_ref(p32[map32[oop.dst]], LAYOUT_F32B, oop.cout, Lout);
      if (need16) {
        const bool hl = wants_hilo(oop.dst, scan_from);
        y16 = make_ref(p16[map16[oop.dst]], LAYOUT_F16B, oop.cout * (hl ? 2 : 1), Lout);
        y16.hilo = hl ? 1 : 0;
        if (cur16[map16[oop.dst]].C != y16.C || cur16[map16[oop.dst]].L != Lout || cur16[map16[oop.dst]].hilo != y16.hilo) {
          cudaError_t e = launch_zero_pads_f16(y16, B, st);
          if (e != cudaSuccess) return fail(MB_ERR_CUDA, "zero_pads: %s", cudaGetErrorString(e));
          count_launch();
          cur16[map16[oop.dst]] = y16;
        }
        plane_slope[map16[oop.dst]] = slope16;
      }
    }
    if (oop.dst2 >= 0) {
      if (oop.dst2 >= nb) return fail(MB_ERR_INVALID, "tc_forward: bad dst2");
      bool n16 = false, n32 = true, hs = false;
      int rc = scan(oop.dst2, n16, n32, y2_slope, hs);
      if (rc != MB_OK) return rc;
      y2_32 = make_ref(p32[map32[oop.dst2]], LAYOUT_F32B, oop.cout, Lout);
      if (n16) {
        const bool hl = wants_hilo(oop.dst2, scan_from);
        y2_16 = make_ref(p16[map16[oop.dst2]], LAYOUT_F16B, oop.cout * (hl ? 2 : 1), Lout);
        y2_16.hilo = hl ? 1 : 0;
        plane_slope[map16[oop.dst2]] = y2_slope;
      }
    }
    if (fuse) {
      const TcOp& c2 = ops[i + 1];
      pp.L = Lin;
      pp.x16 = reinterpret_cast<const __half*>(fused_x16);
      pp.x32 = reinterpret_cast<const float*>(fused_x32);
      pp.slope_in = op.taps.in_slope;
      pp.x_Lp = f16_lp(Lin);
      pp.w1 = reinterpret_cast<const __half*>(tc_arena + op.tc.w16_off);
      pp.w2 = reinterpret_cast<const __half*>(tc_arena + c2.tc.w16_off);
      pp.bias1 = op.b32;
      pp.bias2 = c2.b32;
      pp.slope_mid = c2.taps.in_slope;
      pp.res32 = reinterpret_cast<const float*>(res32.p);
      pp.res16 = reinterpret_cast<const __half*>(res16.p);
      pp.res_Lp = f16_lp(Lout);
      pp.res_inv = 1.f / res_slope;
      pp.y32 = reinterpret_cast<float*>(y32.p);
      pp.y16 = reinterpret_cast<__half*>(y16.p);
      pp.y_Lp = f16_lp(Lout);
      pp.y_hilo = y16.hilo;
      if (y16.hilo && c2.cout != 64) return fail(MB_ERR_INVALID, "tc_pair(%s): hi/lo output needs C = 64", c2.name);
      pp.out_slope = slope16;
      pp.mode = c2.taps.mode;
      pp.div = c2.taps.div;
      pp.lengths = lengths;
      pp.len_mul = c2.taps.len_mul_out;
      if (pp.mode != EPI_STORE && !pp.y32) return fail(MB_ERR_INVALID, "tc_pair(%s): accumulate mode without fp32 plane", c2.name);
      int rc = launch_tc_pair(pp, B, st);
      if (rc != MB_OK) return rc;
      if (events) MB_CUDA_CHECK(cudaEventRecord(events[i + 1], st));
      ++i;  // c2 is done as well
      continue;
    }
    if (op.tc.split3 && op.src == BUF_IN && y16.p && !y32.p && op.res < 0 && op.dst2 < 0) {
      // conv_pre on the tensor cores, fp32-accurate: split the fp32 input into fp16 hi/lo planes (scratch = any
      // other buffer's fp16 storage; all planes are dead at this point of the forward)
      int sidx = -1;
      const size_t need = (size_t)B * 256 * f16_lp(Lin) * 2;
      for (int j = 0; j < nb && sidx < 0; ++j)
        if (j != map16[op.dst] && f16_plane_bytes(B, T, 2 * bufs[j].cr) >= need + kPlaneSlack) sidx = j;
      if (sidx < 0) return fail(MB_ERR_WORKSPACE, "tc_forward: no scratch plane for %s", op.name);
      TRef xs = make_ref(p16[sidx], LAYOUT_F16B, 256, Lin);
      if (cur16[sidx].C != 256 || cur16[sidx].L != Lin) {
        cudaError_t e = launch_zero_pads_f16(xs, B, st);
        if (e != cudaSuccess) return fail(MB_ERR_CUDA, "zero_pads: %s", cudaGetErrorString(e));
        count_launch();
        cur16[sidx] = xs;
      }
      {
        const size_t nthr = (size_t)B * 32 * Lin;
        split3_input_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(mel, reinterpret_cast<__half*>(p16[sidx]), B, op.cin, Lin, lengths);
        MB_LAUNCH_CHECK("split3_input_kernel");
      }
      for (int hf = 0; hf < op.cout / 256; ++hf) {
        TcOp half = op;
        half.taps.Cin = 256;
        half.taps.Cout = 256;
        half.tc.w16_off = op.tc.w16_off + (size_t)hf * kernel_count(op.taps) * 4 * op.tc.slab_bytes;
        int rc = launch_tc(half, tc_arena, xs, TRef{}, TRef{}, 1.f, TRef{}, y16, slope16, lengths, B, Lin, st, hf * 256,
                           op.b32 ? op.b32 + hf * 256 : nullptr);
        if (rc != MB_OK) return rc;
      }
    } else if (op.tc.use_tc) {
      if (op.src < 0 || op.src >= nb) return fail(MB_ERR_INVALID, "tc_forward: tensor-core layer %s reads an external buffer", op.name);
      TRef x16 = make_ref(p16[map16[op.src]], LAYOUT_F16B, op.tc.x3 ? 64 * op.tc.x_pchunks : op.cin, Lin);
      x16.hilo = op.tc.x3;
      if (cur16[map16[op.src]].hilo != x16.hilo)
        return fail(MB_ERR_INVALID, "tc_forward: %s expects a %s input plane", op.name, x16.hilo ? "hi/lo" : "plain");
      int rc = launch_tc(op, tc_arena, x16, res32, res16, res_slope, y32, y16, slope16, lengths, B, Lin, st);
      if (rc != MB_OK) return rc;
    } else {
      TapConv p = op.taps;
      p.B = B;
      p.Lin = Lin;
      p.Lout = Lout;
      p.lengths = lengths;
      TapConvIO io;
      if (op.src == BUF_IN) io.x = make_ref(const_cast<float*>(mel), LAYOUT_NCL, num_mels, Lin);
      else if (op.src >= 0 && op.src < nb) io.x = make_ref(p32[map32[op.src]], LAYOUT_F32B, op.cin, Lin);
      else return fail(MB_ERR_INVALID, "tc_forward: bad src");
      io.res = res32;
      io.y32 = y32;
      io.y16 = y16;
      io.out16_slope = slope16;
      io.y2_32 = y2_32;
      io.y2_16 = y2_16;
      io.y2_16_slope = y2_slope;
      cudaError_t e = (op.cout == 1 && p.stride == 1) ? launch_tapconv_cout1_f32(p, io, op.w32, op.b32, st)
                                                      : launch_tapconv_f32(p, io, op.w32, op.b32, st);
      if (e != cudaSuccess) return fail(MB_ERR_CUDA, "tapconv_f32 (%s): %s", op.name, cudaGetErrorString(e));
      count_launch();
    }
  }
  if (events) MB_CUDA_CHECK(cudaEventRecord(events[n], st));
  return MB_OK;
}

int tc_debug_layer(const TcOp& op, const char* tc_arena, const float* x, const float* residual, int B, int Lin,
                   float* y, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const TapConv& t = op.taps;
  const int Lout = Lin * t.stride;
  TRef xn = make_ref(const_cast<float*>(x), LAYOUT_NCL, t.Cin, Lin);
  TRef yn = make_ref(y, LAYOUT_NCL, t.Cout, Lout);
  if (!op.tc.use_tc) {
    TapConv p = t;
    p.B = B;
    p.Lin = Lin;
    p.Lout = Lout;
    p.lengths = nullptr;
    TapConvIO io;
    io.x = xn;
    io.res = make_ref(const_cast<float*>(residual), LAYOUT_NCL, t.Cout, Lout);
    io.y32 = yn;
    cudaError_t e = (t.Cout == 1 && p.stride == 1) ? launch_tapconv_cout1_f32(p, io, op.w32, op.b32, st)
                                                   : launch_tapconv_f32(p, io, op.w32, op.b32, st);
    if (e != cudaSuccess) return fail(MB_ERR_CUDA, "tapconv_f32: %s", cudaGetErrorString(e));
    count_launch();
    return MB_OK;
  }
  // planes: x16 (activated), res32, y32
  const int xC = op.tc.x3 ? 64 * op.tc.x_pchunks : t.Cin;  // hi/lo plane: twice the channels
  const size_t b_x16 = align_up((size_t)B * xC * f16_lp(Lin) * 2 + kPlaneSlack, 1024);
  const size_t b_r32 = align_up((size_t)B * t.Cout * Lout * 4, 1024);
  if (workspace_bytes < b_x16 + 2 * b_r32 + 2048) return fail(MB_ERR_WORKSPACE, "tc_debug_layer: workspace too small");
  char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  TRef x16 = make_ref(ws, LAYOUT_F16B, xC, Lin);
  x16.hilo = op.tc.x3;
  TRef r32 = residual ? make_ref(ws + b_x16, LAYOUT_F32B, t.Cout, Lout) : TRef{};
  TRef y32 = make_ref(ws + b_x16 + b_r32, LAYOUT_F32B, t.Cout, Lout);
  MB_CUDA_CHECK(cudaMemsetAsync(ws, 0, b_x16, st));
  cudaError_t e = launch_convert_layout(xn, x16, B, t.in_slope, st);
  if (e != cudaSuccess) return fail(MB_ERR_CUDA, "convert: %s", cudaGetErrorString(e));
  if (residual) {
    e = launch_convert_layout(make_ref(const_cast<float*>(residual), LAYOUT_NCL, t.Cout, Lout), r32, B, 1.f, st);
    if (e != cudaSuccess) return fail(MB_ERR_CUDA, "convert: %s", cudaGetErrorString(e));
  }
  int rc = launch_tc(op, tc_arena, x16, r32, TRef{}, 1.f, y32, TRef{}, 1.f, nullptr, B, Lin, st);
  if (rc != MB_OK) return rc;
  e = launch_convert_layout(y32, yn, B, 1.f, st);
  if (e != cudaSuccess) return fail(MB_ERR_CUDA, "convert: %s", cudaGetErrorString(e));
  count_launch(3);
  return MB_OK;
}

}  // namespace mb
