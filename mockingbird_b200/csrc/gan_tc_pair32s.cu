// Full-rate-stage resblock pair with the fp32 residual stream kept in shared memory (round 2 follow-up of
// gan_tc_pair.cu's F32IN variant; hifigan/models.py:36-41: xt = c1(lrelu(x)); xt = c2(lrelu(xt)); x = xt + x).
//
// The F32IN kernel of gan_tc_pair.cu stages the fp32 input window in shared memory for the converter warps, frees it, and has
// the second epilogue re-read the same rows from L2 for the residual add and write the result with per-thread stores.  ncu
// (profiles/r02_pair32_k3_summary.txt) shows those epilogue warps waiting on exactly that load: the kernel is bound by the
// memory-level parallelism of its epilogue, at 39 % of the HBM peak.  Here no thread of the CTA touches global memory:
//
//   producer    : bulk-TMA copies the fp32 window of item i into one of THREE staging slots (8 channel-quad rows of W1 x 16 B)
//   converter   : slot -> lrelu -> fp16 -> swizzled A1 operand                          (4 warps)
//   MMA1 / E1   : as in gan_tc_pair.cu (acc1 -> +b1 -> lrelu -> fp16 -> A2 operand in shared memory)
//   MMA2 / E2   : acc2 + b2 + residual, the residual read from the slot and the result written back IN PLACE (st.shared)
//   store warp  : one thread bulk-TMA stores (or, for the MRF sum, reduce-adds: cp.reduce.async.bulk .add.f32) the slot's
//                 M_out result rows per channel quad to the fp32 plane, waits until the copy engine has read them, frees the slot
//
// Three slots cover load(i + 2), MMA / E1 of i + 1 and E2 / store of i.  MT = 2 row tiles per item (40 KB per slot at k = 11,
// dilation 5) is what fits next to both weight sets; the halo rows cost 4-28 % more fp32 reads than MT = 4, mostly L2 hits.
// Eligible launches: C = 32, fp32 input = residual, plain store or MRF accumulate (EPI_ADD), no fp16 output plane; everything
// else (EPI_ADD_DIV needs the running sum in registers for the fp16 plane of the next stage) keeps the gan_tc_pair.cu kernel.
#include <cstdlib>
#include <cstring>

#include "gan_tc.h"
#include "gan_tc_dev.cuh"
#include "mb_common.h"

namespace mb {

namespace {

using namespace tcdev;

constexpr int kEpi = 8;     // epilogue warps (two groups x four TMEM lane quarters)
constexpr int kCvt = 4;     // converter warps
constexpr int kSlots = 3;   // fp32 staging slots
constexpr int kThreads = 32 * (2 + kEpi + kCvt + 1);
constexpr uint32_t kSmemMax = 227 * 1024;

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g_add_f32(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
               : "memory");
}

template <int MT>
__global__ void __launch_bounds__(kThreads, 1) tc_pair32s_kernel(const __grid_constant__ TcPairParams p) {
  constexpr int N = 32;
  constexpr uint32_t ROWB = 64;                   // one fp16 operand row (32 channels)
  constexpr int NK16 = 2;
  constexpr uint32_t MT_STEP = (128u * ROWB) >> 4;
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t SLAB = (uint32_t)N * ROWB;   // one tap's weight image [N][32]
  constexpr uint32_t TMEM_COLS = 4 * MT * N <= 256 ? 256 : 512;

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
  uint64_t* s32_full = bars + 18;      // [kSlots]  window copied in
  uint64_t* s32_empty = bars + 22;     // [kSlots]  result rows read out by the copy engine
  uint64_t* out_ready = bars + 26;     // [kSlots]  E2 has written the result rows in place
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  uint8_t* s32_base = smem + p.s32_off;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a1_full[i], 32 * kCvt);
      mbar_init(&a1_empty[i], 1);
      mbar_init(&w_full[i], 1);
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], 32 * kEpi);
      mbar_init(&a2_full[i], 32 * kEpi);
      mbar_init(&a2_empty[i], 1);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], 32 * kEpi);
    }
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&s32_full[i], 1);
      mbar_init(&s32_empty[i], 1);
      mbar_init(&out_ready[i], 32 * kEpi);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
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
  const int W1 = p.W1;

  if (warp == 0) {
    // ===================== copy producer =====================
    if (lane == 0) {
      mbar_expect_tx(&w_full[0], (uint32_t)p.k * SLAB);
      bulk_g2s(smem_u32(w1_base), p.w1, (uint32_t)p.k * SLAB, &w_full[0]);
      mbar_expect_tx(&w_full[1], (uint32_t)p.k * SLAB);
      bulk_g2s(smem_u32(w2_base), p.w2, (uint32_t)p.k * SLAB, &w_full[1]);
      for (int it = 0; it < n_items; ++it) {
        const int work = blockIdx.x + it * gridDim.x;
        const int b = work / p.tiles_per_utt;
        const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
        const int slot = it % kSlots, sph = (it / kSlots) & 1;
        mbar_wait(&s32_empty[slot], sph ^ 1);
        const int r0 = m0 - halo;
        const int lo = r0 < 0 ? 0 : r0;
        const int hi = (r0 + W1 < p.L) ? r0 + W1 : p.L;
        const uint32_t qbytes = hi > lo ? (uint32_t)(hi - lo) * 16u : 0u;
        if (qbytes == 0) {
          mbar_arrive(&s32_full[slot]);
          continue;
        }
        mbar_expect_tx(&s32_full[slot], qbytes * (uint32_t)(N / 4));
        uint8_t* dst = s32_base + (size_t)slot * p.s32_stage_bytes + (size_t)(lo - r0) * 16;
        const float* src = p.x32 + ((size_t)b * (N / 4) * p.L + (size_t)lo) * 4;
        for (int q = 0; q < N / 4; ++q)
          bulk_g2s(smem_u32(dst + (size_t)q * W1 * 16), src + (size_t)q * p.L * 4, qbytes, &s32_full[slot]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint64_t desc_hi = make_desc(0, 8u * ROWB, 4u, 0);
    const bool leader = elect_one();
    bool w1_seen = false, w2_seen = false;
    auto mma2 = [&](int j) {
      const int aslot = j % p.a2_stages, aph = (j / p.a2_stages) & 1;
      const int cslot = j & 1, cph = (j >> 1) & 1;
      if (!w2_seen) {
        mbar_wait(&w_full[1], 0);
        w2_seen = true;
      }
      mbar_wait(&a2_full[aslot], aph);
      mbar_wait(&acc2_empty[cslot], cph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)((2 + cslot) * MT * N);
      const uint32_t a_addr = smem_u32(a2_base + (size_t)aslot * p.a2_bytes);
      if (leader) {
        for (int t = 0; t < p.k; ++t) {
          const uint64_t a0 = desc_hi + (uint64_t)((a_addr + (uint32_t)t * ROWB) >> 4);
          const uint64_t b0 = desc_hi + (uint64_t)((smem_u32(w2_base) + (uint32_t)t * SLAB) >> 4);
#pragma unroll
          for (int s = 0; s < NK16; ++s)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              tc_mma_f16(d_tmem + (uint32_t)(mt * N), a0 + (uint64_t)(2 * s + mt * MT_STEP), b0 + (uint64_t)(2 * s), idesc,
                         (t | s) ? 1u : 0u);
        }
        tc_commit(&a2_empty[aslot]);
        tc_commit(&acc2_full[cslot]);
      }
    };
    for (int it = 0; it < n_items; ++it) {
      const int slot = it % p.a1_stages, ph = (it / p.a1_stages) & 1;
      const int cslot = it & 1, cph = (it >> 1) & 1;
      if (!w1_seen) {
        mbar_wait(&w_full[0], 0);
        w1_seen = true;
      }
      mbar_wait(&a1_full[slot], ph);
      mbar_wait(&acc1_empty[cslot], cph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(cslot * MT * N);
      const uint32_t a_addr = smem_u32(a1_base + (size_t)slot * p.a1_stage_bytes);
      if (leader) {
        for (int t = 0; t < p.k; ++t) {
          const uint64_t a0 = desc_hi + (uint64_t)((a_addr + (uint32_t)(t * p.d1) * ROWB) >> 4);
          const uint64_t b0 = desc_hi + (uint64_t)((smem_u32(w1_base) + (uint32_t)t * SLAB) >> 4);
#pragma unroll
          for (int s = 0; s < NK16; ++s)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              tc_mma_f16(d_tmem + (uint32_t)(mt * N), a0 + (uint64_t)(2 * s + mt * MT_STEP), b0 + (uint64_t)(2 * s), idesc,
                         (t | s) ? 1u : 0u);
        }
        tc_commit(&a1_empty[slot]);
        tc_commit(&acc1_full[cslot]);
      }
      if (it > 0) mma2(it - 1);
    }
    if (n_items > 0) mma2(n_items - 1);
  } else if (warp >= 2 && warp < 2 + kEpi) {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    constexpr int G = kEpi / 4;

    auto e1 = [&](int it) {
      const int work = blockIdx.x + it * gridDim.x;
      const int b = work / p.tiles_per_utt;
      const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
      const int valid = p.lengths ? min(p.L, p.lengths[b] * p.len_mul) : p.L;
      const int cslot = it & 1, cph = (it >> 1) & 1;
      const int aslot = it % p.a2_stages, aph = (it / p.a2_stages) & 1;
      mbar_wait(&acc1_full[cslot], cph);
      mbar_wait(&a2_empty[aslot], aph ^ 1);
      tc_fence_after();
      uint8_t* a2 = a2_base + (size_t)aslot * p.a2_bytes;
      for (int mt = grp; mt < MT; mt += G) {
        const int j = mt * 128 + row_in_tile;   // xt row within the item
        const int g = m0 - p.h2 + j;            // global xt row
        const bool live = (g >= 0 && g < valid);
        uint32_t raw[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((cslot * MT + mt) * N), raw);
        uint8_t* rowp = a2 + (size_t)j * ROWB;
        const int sw = f16_swz(32, j);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float x = __uint_as_float(raw[8 * c + e]) + bias1_s[8 * c + e];
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
          *reinterpret_cast<uint4*>(rowp + ((c ^ sw) << 4)) = pk;
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> visible to the UMMA operand fetch
      mbar_arrive(&a2_full[aslot]);
      mbar_arrive(&acc1_empty[cslot]);
    };

    auto e2 = [&](int it) {
      const int work = blockIdx.x + it * gridDim.x;
      const int b = work / p.tiles_per_utt;
      const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
      const int valid = p.lengths ? min(p.L, p.lengths[b] * p.len_mul) : p.L;
      const int cslot = it & 1, cph = (it >> 1) & 1;
      const int slot = it % kSlots, sph = (it / kSlots) & 1;
      mbar_wait(&s32_full[slot], sph);  // (completed long ago: makes the copy engine's writes of the residual rows visible here)
      mbar_wait(&acc2_full[cslot], cph);
      tc_fence_after();
      uint8_t* sl = s32_base + (size_t)slot * p.s32_stage_bytes;
      for (int mt = grp; mt < MT; mt += G) {
        const int i = mt * 128 + row_in_tile;
        const int lo = m0 + i;
        const bool inb = (i < p.M_out) && (lo < p.L);
        const bool live = inb && lo < valid;
        uint32_t raw[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(((2 + cslot) * MT + mt) * N), raw);
        if (!inb) continue;
        float4* rp = reinterpret_cast<float4*>(sl + (size_t)(halo + i) * 16);  // + q * W1 per channel quad
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
          const float4 r = rp[(size_t)q * W1];
          float4 v;
          v.x = __uint_as_float(raw[4 * q + 0]) + bias2_s[4 * q + 0] + r.x;
          v.y = __uint_as_float(raw[4 * q + 1]) + bias2_s[4 * q + 1] + r.y;
          v.z = __uint_as_float(raw[4 * q + 2]) + bias2_s[4 * q + 2] + r.z;
          v.w = __uint_as_float(raw[4 * q + 3]) + bias2_s[4 * q + 3] + r.w;
          if (!live) v = make_float4(0.f, 0.f, 0.f, 0.f);
          rp[(size_t)q * W1] = v;
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> visible to the bulk store
      mbar_arrive(&acc2_empty[cslot]);
      mbar_arrive(&out_ready[slot]);
    };

    for (int it = 0; it < n_items; ++it) {
      e1(it);
      if (it > 0) e2(it - 1);
    }
    if (n_items > 0) e2(n_items - 1);
  } else if (warp < 2 + kEpi + kCvt) {
    // ===================== converter warps: fp32 staging -> lrelu -> fp16 swizzled A1 =====================
    const int tid = threadIdx.x - 32 * (2 + kEpi);
    for (int it = 0; it < n_items; ++it) {
      const int work = blockIdx.x + it * gridDim.x;
      const int b = work / p.tiles_per_utt;
      const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
      const int valid = p.lengths ? min(p.L, p.lengths[b] * p.len_mul) : p.L;
      const int aslot = it % p.a1_stages, aph = (it / p.a1_stages) & 1;
      const int slot = it % kSlots, sph = (it / kSlots) & 1;
      mbar_wait(&s32_full[slot], sph);
      mbar_wait(&a1_empty[aslot], aph ^ 1);
      uint8_t* a1 = a1_base + (size_t)aslot * p.a1_stage_bytes;
      const uint8_t* s32 = s32_base + (size_t)slot * p.s32_stage_bytes;
      const int g0 = m0 - halo;
      for (int j = tid; j < W1; j += 32 * kCvt) {
        const int g = g0 + j;
        const bool live = (g >= 0 && g < valid);
        const int sw = f16_swz(32, j);
#pragma unroll
        for (int c = 0; c < N / 8; ++c) {
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (live) {
            const float4 lo4 = *reinterpret_cast<const float4*>(s32 + ((size_t)(2 * c) * W1 + j) * 16);
            const float4 hi4 = *reinterpret_cast<const float4*>(s32 + ((size_t)(2 * c + 1) * W1 + j) * 16);
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
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> visible to the UMMA operand fetch
      mbar_arrive(&a1_full[aslot]);
    }
  } else {
    // ===================== store warp: result rows of a slot -> fp32 plane (bulk store / reduce-add), then free the slot ==========
    if (lane == 0) {
      for (int it = 0; it < n_items; ++it) {
        const int work = blockIdx.x + it * gridDim.x;
        const int b = work / p.tiles_per_utt;
        const int m0 = (work - b * p.tiles_per_utt) * p.M_out;
        const int slot = it % kSlots, sph = (it / kSlots) & 1;
        mbar_wait(&out_ready[slot], sph);
        const int rows = (p.L - m0 < p.M_out) ? p.L - m0 : p.M_out;
        const uint32_t src = smem_u32(s32_base + (size_t)slot * p.s32_stage_bytes + (size_t)halo * 16);
        float* dst = p.y32 + ((size_t)b * (N / 4) * p.L + (size_t)m0) * 4;
        if (rows > 0) {
          for (int q = 0; q < N / 4; ++q) {
            if (p.mode == EPI_ADD) bulk_s2g_add_f32(dst + (size_t)q * p.L * 4, src + (uint32_t)(q * W1 * 16), (uint32_t)rows * 16u);
            else bulk_s2g(dst + (size_t)q * p.L * 4, src + (uint32_t)(q * W1 * 16), (uint32_t)rows * 16u);
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&s32_empty[slot]);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

bool plan32s(TcPairParams* p) {
  const int k = p->k, d1 = p->d1;
  const int h1 = d1 * (k - 1) / 2, h2 = (k - 1) / 2;
  if (h1 + h2 > kPadRows) return false;
  const int mt = 2;
  const uint32_t wbytes = (uint32_t)align_up((size_t)k * 32 * 64, 1024);
  const int W1 = (mt * 128 + 2 * h1 + 15) & ~15;
  const int W2 = (mt * 128 + 16 + 7) & ~7;
  const uint32_t a1b = (uint32_t)align_up((size_t)W1 * 64, 1024);
  const uint32_t a2b = (uint32_t)align_up((size_t)W2 * 64, 1024);
  const uint32_t s32b = (uint32_t)align_up((size_t)W1 * 128, 1024);
  for (int a1s = 2; a1s >= 1; --a1s) {
    const uint32_t total = kSlots * s32b + a1s * a1b + 2 * a2b + 2 * wbytes + 1024 + 1024;
    if (total > kSmemMax - 1024) continue;
    p->h1 = h1;
    p->h2 = h2;
    p->MT = mt;
    p->M_out = mt * 128 - 2 * h2;
    p->W1 = W1;
    p->W2 = W2;
    p->a1_stages = a1s;
    p->a2_stages = 2;
    p->a1_stage_bytes = a1b;
    p->a2_bytes = a2b;
    p->s32_stage_bytes = s32b;
    p->s32_stages = kSlots;
    p->s32_pieces = 1;
    p->a1_off = 0;
    p->a2_off = a1s * a1b;
    p->w1_off = p->a2_off + 2 * a2b;
    p->w2_off = p->w1_off + wbytes;
    p->bias_off = p->w2_off + wbytes;
    p->bar_off = p->bias_off + 1024;
    p->s32_off = p->bar_off + 1024;
    return true;
  }
  return false;
}

}  // namespace

// MB_TC_PAIR32S=1 selects this kernel for the eligible fp32-input pairs.  Default 0: MEASURED SLOWER than the gan_tc_pair.cu kernel
// (profiles/r02_layers_pair32s_{0,1}.tsv: 0.131 / 0.154 / 0.179 ms vs 0.119 / 0.123 / 0.131 ms per pair at k = 3 / 7 / 11).  A slot
// is occupied from the issue of its load to the end of its store, i.e. for the whole load -> convert -> MMA1 -> E1 -> MMA2 -> E2 ->
// store chain (~8 us); three slots of MT = 2 items then cap the CTA at one item per ~2.7 us, and more slots do not fit.  Kept as an
// A/B switch (parity-tested in tests/test_gan_gpu_toggles.py); the residual now goes through TMEM instead (gan_tc_pair.cu).
bool tc_pair32s_enabled() {
  static const int on = [] {
    const char* e = getenv("MB_TC_PAIR32S");
    return e ? atoi(e) : 0;
  }();
  return on != 0;
}

bool tc_pair32s_eligible(const TcPairParams& p) {
  return tc_pair32s_enabled() && p.f32in && p.C == 32 && (p.mode == EPI_STORE || p.mode == EPI_ADD) && !p.y16 && p.y32 &&
         p.res32 == p.x32 && !p.res16;
}

// launches the shared-memory-residual kernel for an eligible pair; returns MB_OK and sets *done = false when the geometry does not fit
int launch_tc_pair32s(const TcPairParams& p_in, int B, cudaStream_t st, bool* done) {
  *done = false;
  TcPairParams p = p_in;
  if (!plan32s(&p)) return MB_OK;
  p.tiles_per_utt = (p.L + p.M_out - 1) / p.M_out;
  p.n_work = B * p.tiles_per_utt;
  void (*kern)(const TcPairParams) = tc_pair32s_kernel<2>;
  MB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.n_work < sms ? p.n_work : sms;
  *done = true;
  if (grid <= 0) return MB_OK;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemMax;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
  return MB_OK;
}

}  // namespace mb
