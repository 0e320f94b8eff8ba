/*
 * CPU twin of the WaveRNN path (TEST INFRASTRUCTURE - the oracle of the sample loop; never linked
 * into or called by the product).
 *
 * A plain sequential C restatement of
 *   UpsampleNetwork.forward / MelResNet      models/vocoder/wavernn/models/fatchord_version.py:27-85
 *   the sample loop of WaveRNN.generate      fatchord_version.py:176-234  (zero init :178-185,
 *                                            GRUCell maths = ATen gru_cell, RNN.cpp; Categorical
 *                                            sample == argmax(p/q), q ~ Exp(1): SURVEY.md fact 5)
 * written in the CANONICAL floating-point order that the CUDA kernel uses (k-slices, pair sums,
 * binary trees - see DESIGN.md "WaveRNN arithmetic contract"), with the scalar functions of
 * include/mb_wavernn_math.h.  Compile with -O2 -ffp-contract=off: then twin == kernel bit for bit,
 * free running, which tests/test_wavernn_gpu.py asserts; twin vs the unmodified reference is
 * pinned by tests/test_wavernn_twin.py against golden indices generated from the live reference.
 *
 * build: oracle/build_oracle.py  ->  oracle/libwavernn_twin.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mb_wavernn_math.h"

#define RNN 512
#define NCLS 512
#define AUXD 32
#define FEAT 80
#define CDIM 128
#define HOP 200 /* prod(upsample_factors) */

typedef struct {
  /* upsample.resnet */
  const float* conv_in_w;          /* [128][80][5] */
  const float* bn0[4];             /* weight, bias, running_mean, running_var [128] */
  const float* res_conv1[10];      /* [128][128] */
  const float* res_bn1[10][4];
  const float* res_conv2[10];
  const float* res_bn2[10][4];
  const float* conv_out_w;         /* [128][128] */
  const float* conv_out_b;         /* [128] */
  const float* up_w[3];            /* FIR taps, 11 / 11 / 17 */
  /* recurrent core */
  const float* I_w;                /* [512][113] */
  const float* I_b;
  const float* rnn1_wih;           /* [1536][512] */
  const float* rnn1_whh;           /* [1536][512] */
  const float* rnn1_bih;
  const float* rnn1_bhh;
  const float* rnn2_wih;           /* [1536][544] */
  const float* rnn2_whh;           /* [1536][512] */
  const float* rnn2_bih;
  const float* rnn2_bhh;
  const float* fc1_w;              /* [512][544] */
  const float* fc1_b;
  const float* fc2_w;              /* [512][544] */
  const float* fc2_b;
  const float* fc3_w;              /* [512][512] */
  const float* fc3_b;
  int32_t n_res;                   /* 10 */
  int32_t up_scales[3];            /* 5,5,8 */
} twin_weights;

static float dot_seq(const float* w, const float* a, int n) {
  float acc = 0.0f;
  for (int k = 0; k < n; ++k) acc = fmaf(w[k], a[k], acc);
  return acc;
}

/* K = 512 contraction in the kernel's order: `nsub` contiguous sub-slices accumulated sequentially,
 * adjacent pairs added (the two half-warps), then a binary tree over the pair sums. */
static float dot512(const float* w, const float* a, int nsub) {
  float p[32];
  const int len = 512 / nsub;
  for (int j = 0; j < nsub; ++j) p[j] = 0.0f;
  /* the nsub chains are independent: interleaving them does not change any chain's order */
  for (int k = 0; k < len; ++k)
    for (int j = 0; j < nsub; ++j) p[j] = fmaf(w[j * len + k], a[j * len + k], p[j]);
  int n = nsub;
  while (n > 1) {
    for (int i = 0; i < n / 2; ++i) p[i] = p[2 * i] + p[2 * i + 1];
    n /= 2;
  }
  return p[0];
}

static void batch_norm(const float* const bn[4], float* x /*[128][T]*/, int T, int relu) {
  for (int c = 0; c < CDIM; ++c) {
    const float invstd = 1.0f / sqrtf(bn[3][c] + 1e-5f);
    const float alpha = bn[0][c] * invstd;
    const float beta = bn[1][c] - bn[2][c] * alpha;
    for (int t = 0; t < T; ++t) {
      float v = fmaf(x[c * T + t], alpha, beta);
      if (relu && v < 0.0f) v = 0.0f;
      x[c * T + t] = v;
    }
  }
}

static void conv1x1(const float* w /*[128][128]*/, const float* x /*[128][T]*/, float* y, int T) {
  float col[CDIM];
  for (int t = 0; t < T; ++t) {
    for (int c = 0; c < CDIM; ++c) col[c] = x[c * T + t];
    for (int o = 0; o < CDIM; ++o) y[o * T + t] = dot_seq(w + o * CDIM, col, CDIM);
  }
}

/* mel: [80][T] (already divided by mel_max_abs_value).  aux: [T][128].  melup: [200*T][80]. */
void twin_condition(const twin_weights* W, const float* mel, int32_t T, float* aux, float* melup) {
  const int Tp = T + 4;
  float* P = (float*)calloc((size_t)FEAT * Tp, sizeof(float)); /* pad_tensor, fatchord_version.py:169 */
  for (int c = 0; c < FEAT; ++c) memcpy(P + c * Tp + 2, mel + c * T, sizeof(float) * T);

  /* ---- MelResNet (:27-44) */
  float* x = (float*)malloc(sizeof(float) * CDIM * T);
  float* y = (float*)malloc(sizeof(float) * CDIM * T);
  float* r = (float*)malloc(sizeof(float) * CDIM * T);
  for (int o = 0; o < CDIM; ++o)
    for (int t = 0; t < T; ++t) {
      float acc = 0.0f;
      for (int ci = 0; ci < FEAT; ++ci)
        for (int k = 0; k < 5; ++k) acc = fmaf(W->conv_in_w[(o * FEAT + ci) * 5 + k], P[ci * Tp + t + k], acc);
      x[o * T + t] = acc;
    }
  batch_norm(W->bn0, x, T, 1);
  for (int i = 0; i < W->n_res; ++i) {
    memcpy(r, x, sizeof(float) * CDIM * T);
    conv1x1(W->res_conv1[i], x, y, T);
    batch_norm(W->res_bn1[i], y, T, 1);
    conv1x1(W->res_conv2[i], y, x, T);
    batch_norm(W->res_bn2[i], x, T, 0);
    for (int j = 0; j < CDIM * T; ++j) x[j] = x[j] + r[j];
  }
  conv1x1(W->conv_out_w, x, y, T);
  for (int t = 0; t < T; ++t)
    for (int o = 0; o < CDIM; ++o) aux[t * CDIM + o] = y[o * T + t] + W->conv_out_b[o];

  /* ---- mel upsampling ladder: 3 x (repeat s, (2s+1)-tap FIR with zero pad s) (:60-85) */
  int len = Tp;
  float* cur = (float*)malloc(sizeof(float) * FEAT * len);
  memcpy(cur, P, sizeof(float) * FEAT * len);
  for (int st = 0; st < 3; ++st) {
    const int s = W->up_scales[st];
    const int nl = len * s;
    float* nxt = (float*)malloc(sizeof(float) * FEAT * nl);
    for (int c = 0; c < FEAT; ++c)
      for (int i = 0; i < nl; ++i) {
        float acc = 0.0f;
        for (int j = 0; j <= 2 * s; ++j) {
          const int src = i + j - s;
          const float v = (src >= 0 && src < nl) ? cur[c * len + src / s] : 0.0f;
          acc = fmaf(W->up_w[st][j], v, acc);
        }
        nxt[c * nl + i] = acc;
      }
    free(cur);
    cur = nxt;
    len = nl;
  }
  const int indent = 2 * HOP; /* pad * total_scale */
  for (int t = 0; t < HOP * T; ++t)
    for (int c = 0; c < FEAT; ++c) melup[(size_t)t * FEAT + c] = cur[c * len + t + indent];
  free(cur);
  free(P);
  free(x);
  free(y);
  free(r);
}

static void gru_cell(const float* gi, const float* gh, float* h /* in/out [512] */) {
  for (int u = 0; u < RNN; ++u) {
    const float rg = mb_sigmoidf(gh[u] + gi[u]);
    const float zg = mb_sigmoidf(gh[RNN + u] + gi[RNN + u]);
    const float ng = mb_tanhf(gi[2 * RNN + u] + gh[2 * RNN + u] * rg);
    h[u] = (h[u] - ng) * zg + ng;
  }
}

/* noise: [steps][B][512] Exp(1) draws in the reference's order, or NULL -> built-in generator(seed).
 * out_idx: [B][steps].  logits_out (optional): [B][512] logits of the LAST step. */
static void generate_impl(const twin_weights* W, const float* aux, const float* melup, int32_t T,
                          const int32_t* fold_starts, int32_t B, int32_t steps, const float* noise, uint64_t seed,
                          int16_t* out_idx, float* logits_out, const int16_t* force_idx, float* margin) {
  const int total = HOP * T;
  float* h1 = (float*)calloc((size_t)B * RNN, sizeof(float));
  float* h2 = (float*)calloc((size_t)B * RNN, sizeof(float));
  float* xprev = (float*)calloc((size_t)B, sizeof(float));
  const float zeros[AUXD] = {0};
  for (int i = 0; i < steps; ++i) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < B; ++r) {
      float c[FEAT + AUXD], x[RNN], gi[3 * RNN], gh[3 * RNN], f1[RNN], f2[RNN], lg[NCLS], e[NCLS], pr[NCLS];
      const int t = fold_starts[r] + i;
      const int live = t < total; /* fold_with_overlap zero-pads past the end (:321-325) */
      const float* a = live ? aux + (size_t)(t / HOP) * CDIM : NULL;
      for (int j = 0; j < FEAT; ++j) c[j] = live ? melup[(size_t)t * FEAT + j] : 0.0f;
      for (int j = 0; j < AUXD; ++j) c[FEAT + j] = live ? a[j] : 0.0f;
      const float* a2 = live ? a + AUXD : zeros;
      const float* a3 = live ? a + 2 * AUXD : zeros;
      const float* a4 = live ? a + 3 * AUXD : zeros;
      /* I: x = I([x_prev, m_t, a1_t]) */
      for (int u = 0; u < RNN; ++u) {
        const float cond = dot_seq(W->I_w + u * 113 + 1, c, FEAT + AUXD) + W->I_b[u];
        x[u] = fmaf(W->I_w[u * 113], xprev[r], cond);
      }
      /* rnn1 */
      float* hr = h1 + (size_t)r * RNN;
      for (int g = 0; g < 3 * RNN; ++g) {
        gi[g] = dot512(W->rnn1_wih + (size_t)g * RNN, x, 16) + W->rnn1_bih[g];
        gh[g] = dot512(W->rnn1_whh + (size_t)g * RNN, hr, 16) + W->rnn1_bhh[g];
      }
      gru_cell(gi, gh, hr);
      for (int u = 0; u < RNN; ++u) x[u] = x[u] + hr[u];
      /* rnn2 on [x, a2] */
      hr = h2 + (size_t)r * RNN;
      for (int g = 0; g < 3 * RNN; ++g) {
        const float* wr = W->rnn2_wih + (size_t)g * (RNN + AUXD);
        const float auxp = dot_seq(wr + RNN, a2, AUXD) + W->rnn2_bih[g];
        gi[g] = dot512(wr, x, 16) + auxp;
        gh[g] = dot512(W->rnn2_whh + (size_t)g * RNN, hr, 16) + W->rnn2_bhh[g];
      }
      gru_cell(gi, gh, hr);
      for (int u = 0; u < RNN; ++u) x[u] = x[u] + hr[u];
      /* fc1, fc2, fc3 */
      for (int u = 0; u < RNN; ++u) {
        const float* wr = W->fc1_w + (size_t)u * (RNN + AUXD);
        const float v = dot512(wr, x, 32) + (dot_seq(wr + RNN, a3, AUXD) + W->fc1_b[u]);
        f1[u] = v > 0.0f ? v : 0.0f;
      }
      for (int u = 0; u < RNN; ++u) {
        const float* wr = W->fc2_w + (size_t)u * (RNN + AUXD);
        const float v = dot512(wr, f1, 32) + (dot_seq(wr + RNN, a4, AUXD) + W->fc2_b[u]);
        f2[u] = v > 0.0f ? v : 0.0f;
      }
      for (int u = 0; u < NCLS; ++u) lg[u] = dot512(W->fc3_w + (size_t)u * RNN, f2, 32) + W->fc3_b[u];
      /* softmax -> Categorical (renormalise) -> argmax(p / q) */
      float m = lg[0];
      for (int u = 1; u < NCLS; ++u) m = lg[u] > m ? lg[u] : m;
      float part[32];
      for (int l = 0; l < 32; ++l) {
        float s = 0.0f;
        for (int j = 0; j < 16; ++j) {
          e[l + 32 * j] = mb_expf(lg[l + 32 * j] - m);
          s = s + e[l + 32 * j];
        }
        part[l] = s;
      }
      for (int off = 16; off >= 1; off >>= 1) {
        float tmp[32];
        for (int l = 0; l < 32; ++l) tmp[l] = part[l] + part[l ^ off];
        memcpy(part, tmp, sizeof(tmp));
      }
      const float S = part[0];
      for (int l = 0; l < 32; ++l) {
        float s = 0.0f;
        for (int j = 0; j < 16; ++j) {
          pr[l + 32 * j] = e[l + 32 * j] / S;
          s = s + pr[l + 32 * j];
        }
        part[l] = s;
      }
      for (int off = 16; off >= 1; off >>= 1) {
        float tmp[32];
        for (int l = 0; l < 32; ++l) tmp[l] = part[l] + part[l ^ off];
        memcpy(part, tmp, sizeof(tmp));
      }
      const float S2 = part[0];
      int best = 0;
      float bestv = -1.0f, second = -1.0f;
      for (int u = 0; u < NCLS; ++u) {
        const float q = noise ? noise[((size_t)i * B + r) * NCLS + u] : mb_exp1_noise(seed, (uint32_t)i, (uint32_t)r, (uint32_t)u);
        const float v = (pr[u] / S2) / q;
        if (v > bestv) {
          second = bestv;
          bestv = v;
          best = u;
        } else if (v > second) {
          second = v;
        }
      }
      out_idx[(size_t)r * steps + i] = (int16_t)best;
      if (margin) margin[(size_t)r * steps + i] = (bestv - second) / bestv;
      const int fed = force_idx ? (int)force_idx[(size_t)r * steps + i] : best;
      xprev[r] = (2.0f * (float)fed) / 511.0f - 1.0f;
      if (logits_out && i == steps - 1) memcpy(logits_out + (size_t)r * NCLS, lg, sizeof(lg));
    }
  }
  free(h1);
  free(h2);
  free(xprev);
}

void twin_generate(const twin_weights* W, const float* aux, const float* melup, int32_t T,
                   const int32_t* fold_starts, int32_t B, int32_t steps, const float* noise, uint64_t seed,
                   int16_t* out_idx, float* logits_out) {
  generate_impl(W, aux, melup, T, fold_starts, B, steps, noise, seed, out_idx, logits_out, NULL, NULL);
}

/* teacher-forced variant used to pin the twin against the reference: feeds the reference's own
 * sample indices back (x_prev) and reports this twin's argmax and the relative top-2 margin. */
void twin_teacher_forced(const twin_weights* W, const float* aux, const float* melup, int32_t T,
                         const int32_t* fold_starts, int32_t B, int32_t steps, const float* noise,
                         const int16_t* ref_idx /*[B][steps]*/, int16_t* out_idx, float* margin /*[B][steps]*/) {
  generate_impl(W, aux, melup, T, fold_starts, B, steps, noise, 0, out_idx, NULL, ref_idx, margin);
}

/* scalar functions exported for tests/test_wavernn_math.py */
float twin_expf(float x) { return mb_expf(x); }
float twin_logf(float x) { return mb_logf(x); }
float twin_sigmoidf(float x) { return mb_sigmoidf(x); }
float twin_tanhf(float x) { return mb_tanhf(x); }
float twin_noise(uint64_t seed, uint32_t step, uint32_t row, uint32_t cls) { return mb_exp1_noise(seed, step, row, cls); }
