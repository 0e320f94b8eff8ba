// Fused resblock pair on tcgen05: one kernel computes   x_new = c2(lrelu(c1(a) + b1)) + b2 + x_old
// (hifigan/models.py:36-41: xt = c1(lrelu(x)); xt = c2(lrelu(xt)); x = xt + x) for the layers whose
// channel count is <= 64.  The intermediate activation xt never goes to HBM: the first accumulator is
// read from TMEM by the epilogue warps, biased, leaky-relu'd, rounded to fp16 and written straight into
// a swizzled shared-memory operand buffer that the second convolution's MMAs consume.
//
//   work item  = (utterance, M_out = MT*128 - (k-1) output rows)
//   MMA1(i)    : acc1[i&1] = c1 over xt rows [m0 - h2, m0 - h2 + MT*128)        (A1 = bulk-TMA window)
//   E1(i)      : acc1 -> +b1 -> lrelu -> 0 outside [0, valid) -> fp16 -> A2[i]   (smem, SWIZZLE_128B/64B)
//   MMA2(i)    : acc2[i&1] = c2 over output rows [m0, m0 + MT*128)              (A2 rows i + t)
//   E2(i)      : acc2 -> +b2 + residual (fp32) -> fp32 plane / MRF sum and fp16 (activated) plane
// issue order MMA1(i+1), MMA2(i) and E1(i+1), E2(i): the tensor core always has the other phase's tile
// to chew on while the epilogue warps drain one (all four accumulators live in TMEM: 4*MT*N <= 512).
// Both weight sets are shared-memory resident for the whole kernel.
//
// F32IN variant (full-rate stage, where the residual stream must stay fp32): the pair reads ONLY the fp32
// F32B plane.  The producer stages the window's fp32 rows in shared memory (one bulk copy per channel
// quad, clamped to the utterance), four converter warps apply the leaky-relu, round to fp16 and write
// the swizzled A1 operand; the residual add re-reads the same fp32 rows (L2 hits).  No fp16 plane is
// read or written: 420 MB instead of 630 MB of HBM traffic per pair at C = 32, L = 51 200, B = 32.
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <vector>

#include "gan_tc.h"
#include "gan_tc_dev.cuh"
#include "mb_common.h"

namespace mb {

namespace {

using namespace tcdev;

// converter warps of the F32IN variants: two, or four (one per TMEM lane quarter) when they also preload the residual into TMEM
__host__ __device__ constexpr int pair_threads(int ew, bool f32in, int cv = 2) { return 64 + 32 * ew + (f32in ? 32 * cv : 0); }

// EW epilogue warps (8 or 16: EW/4 groups, each covering the four TMEM lane quarters), UC accumulator columns per epilogue
// unit (32, or 16 so that sixteen warps fit the register file: 640 threads x <= 102 registers).  More epilogue warps keep
// more residual loads / stores in flight: the pair kernels are bound by epilogue memory-level parallelism
// (profiles/r02_layers_split{0,1}.tsv: halving the warps per epilogue role cost 30-50 %).
//
// RT (F32IN only, "residual through TMEM"): the second epilogue of the F32IN kernel used to re-read the fp32 residual rows from L2 and
// stalled on that load (profiles/r02_pair32_k3_summary.txt).  With RT the four converter warps, which have the fp32 window in
// shared memory anyway, write the residual rows into the second accumulator with tcgen05.st before MMA2 runs, and MMA2 accumulates
// on top (enable-input-d = 1 from its first instruction): acc2 = x + c2(...).  E2 then only adds the bias and stores.
// CV = converter warps (F32IN).  The role timelines (tools/trace_pair.sh, profiles/r02_pair32_trace_*.txt) show the two converter
// warps busy ~9 000 of the ~10 000 cycles an item takes, but four warps do not shorten the item (see MB_TC_PAIR_CV below).
template <int N, int MT, int CW, bool F32IN, int EW = 8, int UC = 32, bool RT = false, int CV = 2>
__global__ void __launch_bounds__(pair_threads(EW, F32IN, CV), 1) tc_pair_kernel(const __grid_constant__ TcPairParams p) {
  static_assert(!RT || (F32IN && N == 32 && CV == 4), "RT: fp32-input pairs at 32 channels, one converter warp per TMEM lane quarter");
  constexpr int kEpiWarps = EW;
  constexpr int kCvtWarps = CV;
  constexpr uint32_t ROWB = CW * 2;
  constexpr int NK16 = CW / 16;
  constexpr uint32_t MT_STEP = (128u * ROWB) >> 4;
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t SLAB = (uint32_t)N * ROWB;  // one tap's weight image [N][CW]

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a1_base = smem + p.a1_off;
  uint8_t* a2_base = smem + p.a2_off;
  uint8_t* w1_base = smem + p.w1_off;
  uint8_t* w2_base = smem + p.w2_off;
  float* bias1_s = reinterpret_cast<float*>(smem + p.bias_off);
  float* bias2_s = bias1_s + N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* a1_full = bars;            // [2]
  uint64_t* a1_empty = bars + 2;       // [2]
  uint64_t* w_full = bars + 4;         // [2] : 0 = c1 weights, 1 = c2 weights
  uint64_t* acc1_full = bars + 6;      // [2]
  uint64_t* acc1_empty = bars + 8;     // [2]
  uint64_t* a2_full = bars + 10;       // [2]
  uint64_t* a2_empty = bars + 12;      // [2]
  uint64_t* acc2_full = bars + 14;     // [2]
  uint64_t* acc2_empty = bars + 16;    // [2]
  uint64_t* s32_full = bars + 18;      // [4]  (F32IN: fp32 staging of the input window, p.s32_pieces row pieces)
  uint64_t* s32_empty = bars + 22;     // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  uint64_t* res_full = bars + 28;      // [2]  (RT: residual rows stored into acc2[i])
  uint64_t* wr_full = bars + 32;       // [8]  (wstream: weight slab ring)
  uint64_t* wr_empty = bars + 40;      // [8]
  uint8_t* s32_base = smem + p.s32_off;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // debug timeline of CTA 0 (p.trace != nullptr only under MB_TC_PAIR_TRACE): one stamping thread per role
  auto stamp = [&](int role, int it, int ev) {
    if (p.trace && blockIdx.x == 0 && it < 64) p.trace[(role * 64 + it) * 8 + ev] = clock64();
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a1_full[i], F32IN ? 32 * kCvtWarps : 1);
      mbar_init(&a1_empty[i], 1);
      mbar_init(&s32_full[i], 1);
      mbar_init(&s32_empty[i], 32 * kCvtWarps);
      mbar_init(&s32_full[i + 2], 1);
      mbar_init(&s32_empty[i + 2], 32 * kCvtWarps);
      mbar_init(&w_full[i], 1);
      mbar_init(&acc1_full[i], 1);
      // epi_split: warps 2-5 run E1 for every unit, warps 6-9 run E2 (two decoupled pipelines); else all eight run E1 then E2
      mbar_init(&acc1_empty[i], p.epi_split ? 16 * kEpiWarps : 32 * kEpiWarps);
      mbar_init(&a2_full[i], p.epi_split ? 16 * kEpiWarps : 32 * kEpiWarps);
      mbar_init(&a2_empty[i], 1);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], p.epi_split ? 16 * kEpiWarps : 32 * kEpiWarps);
      mbar_init(&res_full[i], 32 * kCvtWarps);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&wr_full[i], 1);
      mbar_init(&wr_empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < N; i += (int)blockDim.x) {
    bias1_s[i] = p.bias1[i];
    bias2_s[i] = p.bias2[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int n_items = (p.n_work > (int)blockIdx.x) ? (p.n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int halo = p.h1 + p.h2;
  // fp32 staging pieces (F32IN): piece hf holds window rows [piece_lo(hf), piece_lo(hf) + piece_rows(hf)) as [N/4][rows][16 B]
  auto piece_lo = [&](int hf) { return p.s32_r0 > 0 ? (hf ? p.s32_r0 : 0) : hf * (p.W1 / p.s32_pieces); };
  auto piece_rows = [&](int hf) { return p.s32_r0 > 0 ? (hf ? p.W1 - p.s32_r0 : p.s32_r0) : p.W1 / p.s32_pieces; };

  if (warp == 0) {
    // ===================== copy producer =====================
    if (lane == 0) {
      if (!p.wstream) {
        mbar_expect_tx(&w_full[0], (uint32_t)p.k * SLAB);
        bulk_g2s(smem_u32(w1_base), p.w1, (uint32_t)p.k * SLAB, &w_full[0]);
        mbar_expect_tx(&w_full[1], (uint32_t)p.k * SLAB);
        bulk_g2s(smem_u32(w2_base), p.w2, (uint32_t)p.k * SLAB, &w_full[1]);
      }
      // wstream: the k slabs of a weight set go through the ring in exactly the order the MMA warp consumes them:
      // W1 for MMA1(it), then W2 for MMA2(it - 1)
      int ws = 0, wph = 0;
      auto stream = [&](const __half* w) {
        for (int t = 0; t < p.k; ++t) {
          mbar_wait(&wr_empty[ws], wph ^ 1);
          mbar_expect_tx(&wr_full[ws], SLAB);
          bulk_g2s(smem_u32(w1_base + (size_t)ws * SLAB), w + (size_t)t * (SLAB >> 1), SLAB, &wr_full[ws]);
          if (++ws == p.wstages) { ws = 0; wph ^= 1; }
        }
      };
      for (int it = 0; it < n_items; ++it) {
        const int work = blockIdx.x + it * gridDim.x;
        const int b = work / p.tiles_per_utt;
        const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
        if constexpr (F32IN) {
          // the fp32 window is staged in p.s32_pieces row pieces (one barrier each): while the converter drains one piece the
          // copies of the others are in flight
          const int g0 = m0 - halo;
          for (int hf = 0; hf < p.s32_pieces; ++hf) {
            const int HW = piece_rows(hf);
            stamp(0, it, 2 * hf);
            mbar_wait(&s32_empty[hf], (it & 1) ^ 1);
            stamp(0, it, 2 * hf + 1);
            const int r0 = g0 + piece_lo(hf);
            const int lo = r0 < 0 ? 0 : r0;
            const int hi = (r0 + HW < p.L) ? r0 + HW : p.L;
            const uint32_t qbytes = hi > lo ? (uint32_t)(hi - lo) * 16u : 0u;
            if (qbytes == 0) {
              mbar_arrive(&s32_full[hf]);  // nothing to copy (window past the end): complete the phase by hand
              continue;
            }
            mbar_expect_tx(&s32_full[hf], qbytes * (uint32_t)(N / 4));
            uint8_t* dst = s32_base + (size_t)hf * p.s32_stage_bytes + (size_t)(lo - r0) * 16;
            const float* src = p.x32 + ((size_t)b * (N / 4) * p.L + (size_t)lo) * 4;
            for (int q = 0; q < N / 4; ++q)
              bulk_g2s(smem_u32(dst + (size_t)q * HW * 16), src + (size_t)q * p.L * 4, qbytes, &s32_full[hf]);
          }
        } else {
          const int slot = it % p.a1_stages, ph = (it / p.a1_stages) & 1;
          stamp(0, it, 0);
          mbar_wait(&a1_empty[slot], ph ^ 1);
          stamp(0, it, 1);
          const uint32_t bytes = (uint32_t)p.W1 * ROWB;
          mbar_expect_tx(&a1_full[slot], bytes);
          const int row0 = (kPadRows + m0 - halo) & ~7;
          const __half* src = p.x16 + ((size_t)b * p.x_Lp + (size_t)row0) * CW;
          bulk_g2s(smem_u32(a1_base + (size_t)slot * p.a1_stage_bytes), src, bytes, &a1_full[slot]);
        }
        if (p.wstream) {
          stream(p.w1);
          if (it > 0) stream(p.w2);
        }
      }
      if (p.wstream && n_items > 0) stream(p.w2);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint64_t desc_hi = make_desc(0, 8u * ROWB, CW == 64 ? 2u : 4u, 0);
    const bool leader = elect_one();
    bool w1_seen = p.wstream != 0, w2_seen = p.wstream != 0;
    int ws = 0, wph = 0;  // weight ring position (wstream)
    auto mma2 = [&](int j) {
      const int aslot = j % p.a2_stages, aph = (j / p.a2_stages) & 1;
      const int cslot = j & 1, cph = (j >> 1) & 1;
      if (!w2_seen) {
        mbar_wait(&w_full[1], 0);
        w2_seen = true;
      }
      if (leader) stamp(2, j, 3);
      mbar_wait(&a2_full[aslot], aph);
      if constexpr (RT) mbar_wait(&res_full[cslot], cph);  // (the converter warps waited for acc2_empty before writing the residual)
      else mbar_wait(&acc2_empty[cslot], cph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)((2 + cslot) * MT * N);
      const uint32_t a_addr = smem_u32(a2_base + (size_t)aslot * p.a2_bytes);
      if (leader) stamp(2, j, 4);
      for (int t = 0; t < p.k; ++t) {
        uint32_t w_addr = smem_u32(w2_base) + (uint32_t)t * SLAB;
        if (p.wstream) {
          mbar_wait(&wr_full[ws], wph);
          tc_fence_after();
          w_addr = smem_u32(w1_base) + (uint32_t)ws * SLAB;
        }
        if (leader) {
          const uint64_t a0 = desc_hi + (uint64_t)((a_addr + (uint32_t)t * ROWB) >> 4);
          const uint64_t b0 = desc_hi + (uint64_t)(w_addr >> 4);
#pragma unroll
          for (int s = 0; s < NK16; ++s)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              tc_mma_f16(d_tmem + (uint32_t)(mt * N), a0 + (uint64_t)(2 * s + mt * MT_STEP), b0 + (uint64_t)(2 * s), idesc,
                         (RT || (t | s)) ? 1u : 0u);
          if (p.wstream) tc_commit(&wr_empty[ws]);
        }
        if (p.wstream && ++ws == p.wstages) { ws = 0; wph ^= 1; }
      }
      if (leader) {
        tc_commit(&a2_empty[aslot]);
        tc_commit(&acc2_full[cslot]);
        stamp(2, j, 5);
      }
    };
    for (int it = 0; it < n_items; ++it) {
      const int work = blockIdx.x + it * gridDim.x;
      const int b = work / p.tiles_per_utt;
      const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
      const int slot = it % p.a1_stages, ph = (it / p.a1_stages) & 1;
      const int cslot = it & 1, cph = (it >> 1) & 1;
      if (!w1_seen) {
        mbar_wait(&w_full[0], 0);
        w1_seen = true;
      }
      if (leader) stamp(2, it, 0);
      mbar_wait(&a1_full[slot], ph);
      mbar_wait(&acc1_empty[cslot], cph ^ 1);
      tc_fence_after();
      if (leader) stamp(2, it, 1);
      const int delta = F32IN ? 0 : ((kPadRows + m0 - halo) & 7);  // the converter writes window row 0 at smem row 0
      const uint32_t d_tmem = tmem_base + (uint32_t)(cslot * MT * N);
      const uint32_t a_addr = smem_u32(a1_base + (size_t)slot * p.a1_stage_bytes);
      for (int t = 0; t < p.k; ++t) {
        uint32_t w_addr = smem_u32(w1_base) + (uint32_t)t * SLAB;
        if (p.wstream) {
          mbar_wait(&wr_full[ws], wph);
          tc_fence_after();
          w_addr = smem_u32(w1_base) + (uint32_t)ws * SLAB;
        }
        if (leader) {
          const uint64_t a0 = desc_hi + (uint64_t)((a_addr + (uint32_t)(delta + t * p.d1) * ROWB) >> 4);
          const uint64_t b0 = desc_hi + (uint64_t)(w_addr >> 4);
#pragma unroll
          for (int s = 0; s < NK16; ++s)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              tc_mma_f16(d_tmem + (uint32_t)(mt * N), a0 + (uint64_t)(2 * s + mt * MT_STEP), b0 + (uint64_t)(2 * s), idesc,
                         (t | s) ? 1u : 0u);
          if (p.wstream) tc_commit(&wr_empty[ws]);
        }
        if (p.wstream && ++ws == p.wstages) { ws = 0; wph ^= 1; }
      }
      if (leader) {
        tc_commit(&a1_empty[slot]);
        tc_commit(&acc1_full[cslot]);
        stamp(2, it, 2);
      }
      if (it > 0) mma2(it - 1);
    }
    if (n_items > 0) mma2(n_items - 1);
  } else if (F32IN && warp >= 2 + kEpiWarps) {
    // ===================== converter warps (F32IN): fp32 staging -> lrelu -> fp16 swizzled A1 =====================
    const int tid = threadIdx.x - 32 * (2 + kEpiWarps);
    for (int it = 0; it < n_items; ++it) {
      const int work = blockIdx.x + it * gridDim.x;
      const int b = work / p.tiles_per_utt;
      const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
      const int valid = p.lengths ? min(p.L, p.lengths[b] * p.len_mul) : p.L;
      const int aslot = it % p.a1_stages, aph = (it / p.a1_stages) & 1;
      if (tid == 0) stamp(1, it, 0);
      mbar_wait(&a1_empty[aslot], aph ^ 1);
      if (tid == 0) stamp(1, it, 1);
      uint8_t* a1 = a1_base + (size_t)aslot * p.a1_stage_bytes;
      const int g0 = m0 - halo;
      [[maybe_unused]] const int cslot = it & 1, cph = (it >> 1) & 1;
      [[maybe_unused]] const int quarter = warp & 3;
      for (int hf = 0; hf < p.s32_pieces; ++hf) {
        const int HW = piece_rows(hf), j0 = piece_lo(hf);
        mbar_wait(&s32_full[hf], it & 1);
        if (tid == 0 && hf < 2) stamp(1, it, 2 + 2 * hf);
        const uint8_t* s32 = s32_base + (size_t)hf * p.s32_stage_bytes;
        for (int jj = tid; jj < HW; jj += 32 * kCvtWarps) {
          const int j = j0 + jj;
          const int g = g0 + j;
          const bool live = (g >= 0 && g < valid);
          const int sw = f16_swz(CW, j);
#pragma unroll
          for (int c = 0; c < N / 8; ++c) {
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (live) {
              const float4 lo4 = *reinterpret_cast<const float4*>(s32 + ((size_t)(2 * c) * HW + jj) * 16);
              const float4 hi4 = *reinterpret_cast<const float4*>(s32 + ((size_t)(2 * c + 1) * HW + jj) * 16);
              __half2 h0 = __floats2half2_rn(lrelu(lo4.x, p.slope_in), lrelu(lo4.y, p.slope_in));
              __half2 h1 = __floats2half2_rn(lrelu(lo4.z, p.slope_in), lrelu(lo4.w, p.slope_in));
              __half2 h2 = __floats2half2_rn(lrelu(hi4.x, p.slope_in), lrelu(hi4.y, p.slope_in));
              __half2 h3 = __floats2half2_rn(lrelu(hi4.z, p.slope_in), lrelu(hi4.w, p.slope_in));
              pk.x = *reinterpret_cast<uint32_t*>(&h0);
              pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2);
              pk.w = *reinterpret_cast<uint32_t*>(&h3);
            }
            *reinterpret_cast<uint4*>(a1 + (size_t)j * ROWB + ((c ^ sw) << 4)) = pk;
          }
        }
        if (hf == p.s32_pieces - 1) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> visible to the UMMA operand fetch
          mbar_arrive(&a1_full[aslot]);
        }
        if constexpr (RT) {
          // residual rows of the row tiles this piece covers (window rows halo + i) -> acc2[it & 1]; one TMEM lane quarter per
          // converter warp.  The piece boundary is halo + a whole number of row tiles (tc_pair_plan), so a tile never straddles.
          if (hf == 0) {
            mbar_wait(&acc2_empty[cslot], cph ^ 1);  // E2(it - 2) has drained this accumulator
            tc_fence_after();
            if (tid == 0) stamp(1, it, 6);
          }
          const int mt_lo = hf * (MT / p.s32_pieces), mt_hi = (hf + 1) * (MT / p.s32_pieces);
#pragma unroll 1
          for (int mt = mt_lo; mt < mt_hi; ++mt) {
            const int i = mt * 128 + quarter * 32 + lane;
            const int jj = halo + i - j0;
            const bool inb = (i < p.M_out) && (m0 + i < p.L);
            uint32_t r[32];
            if (inb) {
#pragma unroll
              for (int q = 0; q < N / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(s32 + ((size_t)q * HW + jj) * 16);
                r[4 * q + 0] = __float_as_uint(v.x);
                r[4 * q + 1] = __float_as_uint(v.y);
                r[4 * q + 2] = __float_as_uint(v.z);
                r[4 * q + 3] = __float_as_uint(v.w);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e) r[e] = 0u;
            }
            tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(((2 + cslot) * MT + mt) * N), r);
          }
        }
        mbar_arrive(&s32_empty[hf]);
        if (tid == 0 && hf < 2) stamp(1, it, 3 + 2 * hf);
      }
      if constexpr (RT) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(&res_full[cslot]);
        if (tid == 0) stamp(1, it, 7);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    constexpr int UPT = N / UC;          // UC-column units per row tile
    constexpr int NUNITS = MT * UPT;
    constexpr int G = EW / 4;            // epilogue groups
    // unit schedule of this group: all groups share E1 then E2, or (epi_split) the lower half of the groups runs E1, the upper E2
    const int ustride = p.epi_split ? G / 2 : G;
    const int ufirst = p.epi_split ? grp % (G / 2) : grp;
    const int C4 = N >> 2;
    constexpr int OCW = (N >= 64) ? 64 : N;  // output plane row chunk (f16_cw)

    auto e1 = [&](int it) {
      const int work = blockIdx.x + it * gridDim.x;
      const int b = work / p.tiles_per_utt;
      const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
      const int valid = p.lengths ? min(p.L, p.lengths[b] * p.len_mul) : p.L;
      const int cslot = it & 1, cph = (it >> 1) & 1;
      const int aslot = it % p.a2_stages, aph = (it / p.a2_stages) & 1;
      if (warp == 2 && lane == 0) stamp(3, it, 0);
      mbar_wait(&acc1_full[cslot], cph);
      mbar_wait(&a2_empty[aslot], aph ^ 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) stamp(3, it, 1);
      uint8_t* a2 = a2_base + (size_t)aslot * p.a2_bytes;
      for (int u = ufirst; u < NUNITS; u += ustride) {
        const int mt = u / UPT;
        const int col0 = (u - mt * UPT) * UC;
        const int j = mt * 128 + row_in_tile;   // xt row within the item
        const int g = m0 - p.h2 + j;            // global xt row
        const bool live = (g >= 0 && g < valid);
        uint32_t raw[UC];
        tmem_ld<UC>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((cslot * MT + mt) * N + col0), raw);
        uint8_t* rowp = a2 + (size_t)j * ROWB;
        const int sw = f16_swz(CW, j);
#pragma unroll
        for (int c = 0; c < UC / 8; ++c) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float x = __uint_as_float(raw[8 * c + e]) + bias1_s[col0 + 8 * c + e];
            v[e] = live ? lrelu(x, p.slope_mid) : 0.f;
          }
          __half2 h0 = __floats2half2_rn(v[0], v[1]);
          __half2 h1 = __floats2half2_rn(v[2], v[3]);
          __half2 h2 = __floats2half2_rn(v[4], v[5]);
          __half2 h3 = __floats2half2_rn(v[6], v[7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2);
          pk.w = *reinterpret_cast<uint32_t*>(&h3);
          const int chunk = (col0 >> 3) + c;   // 16-byte chunk within the row (C <= 64: one row chunk)
          *reinterpret_cast<uint4*>(rowp + ((chunk ^ sw) << 4)) = pk;
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> visible to the UMMA operand fetch
      mbar_arrive(&a2_full[aslot]);
      mbar_arrive(&acc1_empty[cslot]);
      if (warp == 2 && lane == 0) stamp(3, it, 2);
    };

    constexpr int UPG = (NUNITS + 1) / 2;  // units per epilogue group
    auto e2 = [&](int it) {
      const int work = blockIdx.x + it * gridDim.x;
      const int b = work / p.tiles_per_utt;
      const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
      const int valid = p.lengths ? min(p.L, p.lengths[b] * p.len_mul) : p.L;
      const int cslot = it & 1, cph = (it >> 1) & 1;
      bool waited = false;
      if (warp == 2 && lane == 0) stamp(3, it, 3);
#pragma unroll
      for (int ui = 0; ui < NUNITS; ++ui) {
        const int u = ufirst + ustride * ui;
        if (u >= NUNITS) break;
        const int mt = u / UPT;
        const int col0 = (u - mt * UPT) * UC;
        const int i = mt * 128 + row_in_tile;
        const int lo = m0 + i;
        const bool inb = (i < p.M_out) && (lo < p.L);
        const bool live = inb && lo < valid;
        const size_t i32 = ((size_t)b * C4 + (col0 >> 2)) * p.L + lo;
        float4 rv[UC / 4], ov[UC / 4];
        uint4 rh[UC / 8];
        if (!RT && inb && p.res32) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g) rv[g] = reinterpret_cast<const float4*>(p.res32)[i32 + (size_t)g * p.L];
        }
        if (inb && p.res16) {
          const int rr = kPadRows + lo;
          const size_t rbase = ((size_t)b * p.res_Lp + rr) * (size_t)(OCW >> 3);
          const int sw = f16_swz(OCW, rr);
#pragma unroll
          for (int g = 0; g < UC / 8; ++g) rh[g] = reinterpret_cast<const uint4*>(p.res16)[rbase + (size_t)(((col0 >> 3) + g) ^ sw)];
        }
        if (inb && p.mode != EPI_STORE && !p.red_add) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g) ov[g] = reinterpret_cast<const float4*>(p.y32)[i32 + (size_t)g * p.L];
        }
        if (!waited) {
          mbar_wait(&acc2_full[cslot], cph);
          tc_fence_after();
          waited = true;
          if (warp == 2 && lane == 0) stamp(3, it, 4);
        }
        uint32_t raw[UC];
        tmem_ld<UC>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(((2 + cslot) * MT + mt) * N + col0), raw);
        if (!inb) continue;
        float v[UC];
#pragma unroll
        for (int e = 0; e < UC; ++e) v[e] = __uint_as_float(raw[e]) + bias2_s[col0 + e];
        if (!RT && p.res32) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g) {
            v[4 * g + 0] += rv[g].x; v[4 * g + 1] += rv[g].y; v[4 * g + 2] += rv[g].z; v[4 * g + 3] += rv[g].w;
          }
        }
        if (p.res16) {
#pragma unroll
          for (int g = 0; g < UC / 8; ++g) add_res16(&v[8 * g], rh[g], p.res_inv);
        }
        if (p.mode != EPI_STORE && !p.red_add) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g) {
            v[4 * g + 0] += ov[g].x; v[4 * g + 1] += ov[g].y; v[4 * g + 2] += ov[g].z; v[4 * g + 3] += ov[g].w;
          }
          if (p.mode == EPI_ADD_DIV) {
#pragma unroll
            for (int e = 0; e < UC; ++e) v[e] /= p.div;
          }
        }
        if (!live) {
#pragma unroll
          for (int e = 0; e < UC; ++e) v[e] = 0.f;
        }
        if (p.red_add) {  // MRF sum of a middle resblock: S += v by vector reductions in L2 (rows past the length add nothing)
          if (live) {
#pragma unroll
            for (int g = 0; g < UC / 4; ++g)
              red_add_f32x4(p.y32 + (i32 + (size_t)g * p.L) * 4, v[4 * g + 0], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          }
        } else if (p.y32) {
#pragma unroll
          for (int g = 0; g < UC / 4; ++g)
            reinterpret_cast<float4*>(p.y32)[i32 + (size_t)g * p.L] =
                make_float4(v[4 * g + 0], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
        if (p.y16) {
          const int rr = kPadRows + lo;
          const int sw = f16_swz(OCW, rr);
#pragma unroll
          for (int g = 0; g < UC / 8; ++g) {
            __half2 h0 = __floats2half2_rn(lrelu(v[8 * g + 0], p.out_slope), lrelu(v[8 * g + 1], p.out_slope));
            __half2 h1 = __floats2half2_rn(lrelu(v[8 * g + 2], p.out_slope), lrelu(v[8 * g + 3], p.out_slope));
            __half2 h2 = __floats2half2_rn(lrelu(v[8 * g + 4], p.out_slope), lrelu(v[8 * g + 5], p.out_slope));
            __half2 h3 = __floats2half2_rn(lrelu(v[8 * g + 6], p.out_slope), lrelu(v[8 * g + 7], p.out_slope));
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2);
            pk.w = *reinterpret_cast<uint32_t*>(&h3);
            const int cc = (col0 >> 3) + g;
            // plain plane: one row chunk per utterance; hi/lo plane (N == 64 only): chunk 0 = hi, chunk 1 = lo
            const size_t cstride = (size_t)p.y_Lp * (size_t)(OCW >> 3);
            const size_t i16 = (size_t)b * (p.y_hilo ? 2 : 1) * cstride + (size_t)rr * (size_t)(OCW >> 3) + (size_t)(cc ^ sw);
            reinterpret_cast<uint4*>(p.y16)[i16] = pk;
            if (p.y_hilo) {
              const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
              __half2 l0 = __floats2half2_rn(lrelu(v[8 * g + 0], p.out_slope) - f0.x, lrelu(v[8 * g + 1], p.out_slope) - f0.y);
              __half2 l1 = __floats2half2_rn(lrelu(v[8 * g + 2], p.out_slope) - f1.x, lrelu(v[8 * g + 3], p.out_slope) - f1.y);
              __half2 l2 = __floats2half2_rn(lrelu(v[8 * g + 4], p.out_slope) - f2.x, lrelu(v[8 * g + 5], p.out_slope) - f2.y);
              __half2 l3 = __floats2half2_rn(lrelu(v[8 * g + 6], p.out_slope) - f3.x, lrelu(v[8 * g + 7], p.out_slope) - f3.y);
              uint4 pl;
              pl.x = *reinterpret_cast<uint32_t*>(&l0);
              pl.y = *reinterpret_cast<uint32_t*>(&l1);
              pl.z = *reinterpret_cast<uint32_t*>(&l2);
              pl.w = *reinterpret_cast<uint32_t*>(&l3);
              reinterpret_cast<uint4*>(p.y16)[i16 + cstride] = pl;
            }
          }
        }
      }
      if (!waited) {
        mbar_wait(&acc2_full[cslot], cph);
        tc_fence_after();
      }
      tc_fence_before();
      mbar_arrive(&acc2_empty[cslot]);
      if (warp == 2 && lane == 0) stamp(3, it, 5);
    };

    if (p.epi_split) {
      // decoupled: E2(it) no longer waits behind E1(it+1) (whose accumulator may be late), E1 never waits behind E2's global traffic
      if (grp < G / 2) {
        for (int it = 0; it < n_items; ++it) e1(it);
      } else {
        for (int it = 0; it < n_items; ++it) e2(it);
      }
    } else {
      for (int it = 0; it < n_items; ++it) {
        e1(it);
        if (it > 0) e2(it - 1);
      }
      if (n_items > 0) e2(n_items - 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

constexpr uint32_t kSmemMax = 227 * 1024;

}  // namespace

// MB_TC_PAIR_RT=0: fp32-input pairs never preload their residual into TMEM (the second epilogue re-reads it from L2); the unequal
// staging pieces the RT kernels need are planned whenever the switch is not 0 (the other kernels take any piece geometry)
bool tc_pair_rt_enabled() {
  static const int on = [] {
    const char* e = getenv("MB_TC_PAIR_RT");
    return e ? atoi(e) : 2;
  }();
  return on != 0;
}

// MB_TC_PAIR_WSTREAM=1: the C = 64 pairs with k >= 7 stream both weight sets through an 8-slab ring (frees the shared memory for
// MT = 2 and double-buffered operands at k = 11).  Default 0: MEASURED SLOWER (profiles/r02_layers_wstream_{0,1}.tsv: k = 7
// 0.124 -> 0.147 ms, k = 11 0.190 -> 0.203 ms per pair; only k = 11 dilation 5 gains, 0.247 -> 0.231).  The role timeline
// (profiles/r02_pair16_trace_wstream_k7.txt) shows the MMA warp at 87 instead of 54 cycles per MMA: a slab is re-requested only when
// the MMAs that read it have COMPLETED, and eight slabs (~3 000 cycles of MMA work) do not cover the loaded L2 latency of the copy
// engine, which also carries the operand windows.  A deeper ring does not fit; the real fix is a 2-CTA pair sharing one weight copy.
bool tc_pair_wstream_enabled() {
  static const int on = [] {
    const char* e = getenv("MB_TC_PAIR_WSTREAM");
    return e ? atoi(e) : 0;
  }();
  return on != 0;
}

bool tc_pair_plan(int C, int k, int d1, bool f32in, TcPairParams* p) {
  if (C != 64 && C != 32) return false;
  if (f32in && C != 32) return false;  // instance list below
  const int row_bytes = C * 2;
  const int h1 = d1 * (k - 1) / 2, h2 = (k - 1) / 2;
  if (h1 + h2 > kPadRows) return false;
  const uint32_t slab = (uint32_t)C * row_bytes;
  const uint32_t wbytes = (uint32_t)align_up((size_t)k * slab, 1024);
  const uint32_t usable = kSmemMax - 1024;
  const int mt_max = (C == 64) ? 2 : 4;  // 4 accumulators of MT*N columns in 512 TMEM columns
  // stage counts tried per tile count, best first: (fp32 staging, A1, A2).  Large tiles matter more than
  // deep rings: the per-item dependency chain (load -> [convert] -> MMA1 -> E1 -> MMA2 -> E2) is a fixed
  // latency, so fewer, bigger items amortise it.
  static const int kTry16[][3] = {{0, 2, 2}, {0, 2, 1}, {0, 1, 1}};
  static const int kTry32[][3] = {{1, 2, 2}, {1, 2, 1}, {1, 1, 1}, {1, 1, 1}};  // fp32 staging: one window, two row halves
  for (int mt = mt_max; mt >= 1; mt >>= 1) {
    for (int v = 0; v < (f32in ? 4 : 3); ++v) {
      const int* st = f32in ? kTry32[v] : kTry16[v];
      const int W1 = f32in ? ((mt * 128 + 2 * h1 + 15) & ~15) : ((mt * 128 + 2 * h1 + 7 + 7) & ~7);
      const int W2 = (mt * 128 + 16 + 7) & ~7;
      const uint32_t a1b = (uint32_t)align_up((size_t)W1 * row_bytes, 1024);
      const uint32_t a2b = (uint32_t)align_up((size_t)W2 * row_bytes, 1024);
      static const int pieces = [] {
        // A/B switch: row pieces of the fp32 staging window, 2 (default) or 4.  Measured: 4 is slightly SLOWER (5.09 vs 5.00 ms per
        // profiled forward) - the bytes in flight are set by the window's size, not by how it is cut
        const char* e = getenv("MB_TC_PAIR_S32P");
        const int v = e ? atoi(e) : 2;
        return v == 4 ? 4 : 2;
      }();
      // RT kernels: two unequal pieces cut at halo + half the row tiles (so that a tile's residual rows live in one piece)
      const bool rt = f32in && tc_pair_rt_enabled();
      const int rt_pieces = mt >= 2 ? 2 : 1;
      const int r0 = (rt && rt_pieces == 2) ? (h1 + h2 + 128 * (mt / 2)) : 0;
      const int np = rt ? rt_pieces : pieces;
      const int piece_max = r0 > 0 ? (r0 > W1 - r0 ? r0 : W1 - r0) : W1 / np;
      const uint32_t s32b = f32in ? (uint32_t)align_up((size_t)piece_max * C * 4, 1024) : 0u;  // one row piece (W1 % 16 == 0)
      // C = 64, k >= 7: both weight sets (2 x k x 8 KB) leave no room for double-buffered operands - stream them through a ring
      const bool wstream = !f32in && C == 64 && k >= 7 && tc_pair_wstream_enabled();
      const int wstages = 8;
      const uint32_t wtotal = wstream ? (uint32_t)align_up((size_t)wstages * slab, 1024) : 2 * wbytes;
      const uint32_t total = np * s32b + st[1] * a1b + st[2] * a2b + wtotal + 1024 + 1024;
      p->wstream = wstream ? 1 : 0;
      p->wstages = wstages;
      p->s32_pieces = np;
      p->s32_r0 = r0;
      p->rt = rt ? 1 : 0;
      if (total > usable) continue;
      p->f32in = f32in ? 1 : 0;
      p->s32_stage_bytes = s32b;
      p->s32_stages = st[0] > 0 ? st[0] : 1;
      p->C = C;
      p->k = k;
      p->d1 = d1;
      p->h1 = h1;
      p->h2 = h2;
      p->MT = mt;
      p->M_out = mt * 128 - 2 * h2;
      p->W1 = W1;
      p->W2 = W2;
      p->a1_stages = st[1];
      p->a2_stages = st[2];
      p->a1_stage_bytes = a1b;
      p->a2_bytes = a2b;
      p->a1_off = 0;
      p->a2_off = st[1] * a1b;
      p->w1_off = p->a2_off + st[2] * a2b;
      p->w2_off = wstream ? p->w1_off : p->w1_off + wbytes;
      p->bias_off = p->w1_off + wtotal;
      p->bar_off = p->bias_off + 1024;
      p->s32_off = p->bar_off + 1024;
      return true;
    }
  }
  return false;
}

int launch_tc_pair(TcPairParams& p, int B, cudaStream_t st) {
  if (tc_pair32s_eligible(p)) {
    bool done = false;
    int rc = launch_tc_pair32s(p, B, st, &done);
    if (rc != MB_OK || done) return rc;
  }
  p.tiles_per_utt = (p.L + p.M_out - 1) / p.M_out;
  p.n_work = B * p.tiles_per_utt;
  static const int split = [] {
    // A/B switch: 1 = E1 and E2 on separate warp groups.  Measured (profiles/r02_layers_split{0,1}.tsv): SLOWER, 5.51 vs 5.01 ms per
    // step - four warps per role keep fewer loads in flight than eight warps doing both in turn; the default stays 0.
    const char* e = getenv("MB_TC_PAIR_SPLIT");
    return e ? atoi(e) : 0;
  }();
  p.epi_split = split ? 1 : 0;
  p.red_add = (p.mode == EPI_ADD && !p.y16 && p.y32 && tc_red_add_enabled()) ? 1 : 0;
  void (*kern)(const TcPairParams) = nullptr;
  static const int ew16 = [] {
    // A/B switch: 1 (default) = sixteen epilogue warps working on 16-column units.  Measured (profiles/r02_layers_ew16_{0,1}.tsv):
    // 5.125 -> 5.012 ms per profiled forward, C = 64 k = 3 pairs -20 %, accumulate-mode pairs of the full-rate stage -8..-17 %.
    const char* e = getenv("MB_TC_PAIR_EW16");
    return e ? atoi(e) : 1;
  }();
  static const int rt_mode = [] {
    // MB_TC_PAIR_RT: 0 = never, 1 = every eligible fp32-input pair, 2 (default) = only k <= 3.  Measured per pair (same file):
    // k = 3: 0.137 -> 0.122 ms; k = 7: 0.143 -> 0.174; k = 11: 0.149 -> 0.204 - with more taps the converter's wait for
    // acc2_empty (E2 of item i - 2) sits on the critical path and the pipeline serialises.
    const char* e = getenv("MB_TC_PAIR_RT");
    return e ? atoi(e) : 2;
  }();
  static const int cv4 = [] {
    // MB_TC_PAIR_CV: converter warps of the fp32-input kernels that do not preload the residual: 2 (default) or 4.  Measured
    // (profiles/r02_ab_pair_{A..E}.tsv): 4 is SLOWER (stage sum 1.477 vs 1.391 ms) although the two converter warps are the busiest
    // role of the item pipeline - every role slows down when the others run (shared LSU / shared-memory pipe), so more converter
    // threads only take bandwidth from the MMA operand fetch and the epilogues.
    const char* e = getenv("MB_TC_PAIR_CV");
    return (e ? atoi(e) : 2) == 4;
  }();
  const bool rt = p.rt && (rt_mode == 1 || p.k <= 3) && p.f32in && p.C == 32 && p.res32 && p.res32 == p.x32 && !p.res16;
  int threads = pair_threads(8, p.f32in != 0);
  if (rt) {
    threads = pair_threads(8, true, 4);
    if (p.MT == 4) kern = tc_pair_kernel<32, 4, 32, true, 8, 32, true, 4>;
    else if (p.MT == 2) kern = tc_pair_kernel<32, 2, 32, true, 8, 32, true, 4>;
    else if (p.MT == 1) kern = tc_pair_kernel<32, 1, 32, true, 8, 32, true, 4>;
  } else if (p.f32in && cv4 && p.C == 32 && p.MT >= 2) {
    if (ew16) {
      threads = pair_threads(16, true, 4);
      if (p.MT == 4) kern = tc_pair_kernel<32, 4, 32, true, 16, 16, false, 4>;
      else kern = tc_pair_kernel<32, 2, 32, true, 16, 16, false, 4>;
    } else {
      threads = pair_threads(8, true, 4);
      if (p.MT == 4) kern = tc_pair_kernel<32, 4, 32, true, 8, 16, false, 4>;
      else kern = tc_pair_kernel<32, 2, 32, true, 8, 16, false, 4>;
    }
  } else if (ew16 && p.MT >= 2) {
    threads = pair_threads(16, p.f32in != 0);
    if (p.f32in && p.C == 32 && p.MT == 4) kern = tc_pair_kernel<32, 4, 32, true, 16, 16>;
    else if (p.f32in && p.C == 32 && p.MT == 2) kern = tc_pair_kernel<32, 2, 32, true, 16, 16>;
    else if (!p.f32in && p.C == 64 && p.MT == 2) kern = tc_pair_kernel<64, 2, 64, false, 16, 16>;
    else if (!p.f32in && p.C == 32 && p.MT == 4) kern = tc_pair_kernel<32, 4, 32, false, 16, 16>;
    else if (!p.f32in && p.C == 32 && p.MT == 2) kern = tc_pair_kernel<32, 2, 32, false, 16, 16>;
    else threads = pair_threads(8, p.f32in != 0);
  }
  if (kern) {
  } else if (p.f32in && p.C == 32 && p.MT == 4) kern = tc_pair_kernel<32, 4, 32, true>;
  else if (p.f32in && p.C == 32 && p.MT == 2) kern = tc_pair_kernel<32, 2, 32, true>;
  else if (p.f32in && p.C == 32 && p.MT == 1) kern = tc_pair_kernel<32, 1, 32, true>;
  else if (p.f32in) kern = nullptr;
  else if (p.C == 64 && p.MT == 2) kern = tc_pair_kernel<64, 2, 64, false>;
  else if (p.C == 64 && p.MT == 1) kern = tc_pair_kernel<64, 1, 64, false>;
  else if (p.C == 32 && p.MT == 4) kern = tc_pair_kernel<32, 4, 32, false>;
  else if (p.C == 32 && p.MT == 2) kern = tc_pair_kernel<32, 2, 32, false>;
  else if (p.C == 32 && p.MT == 1) kern = tc_pair_kernel<32, 1, 32, false>;
  if (!kern) return fail(MB_ERR_INVALID, "tc_pair: no kernel instance for C=%d MT=%d f32in=%d", p.C, p.MT, p.f32in);
  MB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.n_work < sms ? p.n_work : sms;
  if (grid <= 0) return MB_OK;
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
  // debug: MB_TC_PAIR_TRACE=<file> MB_TC_PAIR_TRACE_K=<k> dumps CTA 0's role timeline of the first fp32-input launch with k taps
  static const char* trace_path = getenv("MB_TC_PAIR_TRACE");
  static bool traced = false;
  long long* trace_dev = nullptr;
  p.trace = nullptr;
  static const int trace_f32 = getenv("MB_TC_PAIR_TRACE_F32") ? atoi(getenv("MB_TC_PAIR_TRACE_F32")) : 1;  // 0: trace an fp16-plane pair
  if (trace_path && !traced && (p.f32in != 0) == (trace_f32 != 0)) {
    const char* ek = getenv("MB_TC_PAIR_TRACE_K");
    static int skip = getenv("MB_TC_PAIR_TRACE_SKIP") ? atoi(getenv("MB_TC_PAIR_TRACE_SKIP")) : 0;  // matching launches to let pass
    if ((!ek || atoi(ek) == p.k) && skip-- <= 0) {
      traced = true;
      MB_CUDA_CHECK(cudaMalloc(&trace_dev, 4 * 64 * 8 * sizeof(long long)));
      MB_CUDA_CHECK(cudaMemsetAsync(trace_dev, 0, 4 * 64 * 8 * sizeof(long long), st));
      p.trace = trace_dev;
    }
  }
  MB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
  if (trace_dev) {
    std::vector<long long> host(4 * 64 * 8);
    MB_CUDA_CHECK(cudaStreamSynchronize(st));
    MB_CUDA_CHECK(cudaMemcpy(host.data(), trace_dev, host.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(trace_dev);
    p.trace = nullptr;
    if (FILE* f = fopen(trace_path, "w")) {
      fprintf(f, "# k=%d d1=%d MT=%d M_out=%d W1=%d a1_stages=%d a2_stages=%d threads=%d n_work=%d grid=%d\n", p.k, p.d1, p.MT, p.M_out,
              p.W1, p.a1_stages, p.a2_stages, threads, p.n_work, grid);
      for (int r = 0; r < 4; ++r)
        for (int it = 0; it < 64; ++it) {
          fprintf(f, "%d %d", r, it);
          for (int e = 0; e < 8; ++e) fprintf(f, " %lld", host[(r * 64 + it) * 8 + e]);
          fprintf(f, "\n");
        }
      fclose(f);
    }
  }
  return MB_OK;
}

}  // namespace mb
