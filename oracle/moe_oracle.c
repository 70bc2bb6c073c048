/*
 * moe_oracle.c -- plain-C restatement of the fast_moe expert layer.  TEST INFRASTRUCTURE, NOT PRODUCT CODE: only
 * tests/ may load the library built from this file; the product (3m-asr-inference_b200/) never links or calls it.
 *
 * It follows the same reference lines as oracle/moe_oracle.py (paths relative to the upstream tree):
 *   oracle_gate_3m     trainer_3m_fix/model/dfsmn_base_fmoe_localComm_catEmbed.py:166-181,210-211;
 *                      TRTAPI++/plugin/softmax_topk_plugin/softmax_topk_kernel.cu:26-89
 *   oracle_gate_naive  trainer_3m_fix/fmoe/gates.py:51-66
 *   oracle_prepare     trainer_3m_fix/fmoe/functions.py:29-35; TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_kernel.cu:25-73
 *   oracle_expert_ffn  TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_plugin.cpp:82-128; fmoe/functions.py:142-148
 *   oracle_combine     fmoe_expert_kernel.cu:191-227; layer/positionwise_feed_forward.py:257-258; fmoe/layers.py:204-206;
 *                      layer/fmoe_transformer.py:155-158
 * Parity: pinned through tests/test_c_oracle.py against oracle/moe_oracle.py, which is itself pinned against vectors
 * produced by the reference's own Python (tests/golden/make_golden.py).  The reference's C++/CUDA sources for this
 * path cannot be compiled here (they include NvInfer.h; TensorRT is not in the image), so there is no oracle/_ref.
 *
 * Scalar loops on purpose (double accumulation for the router, float for the experts); the omp pragmas are inert
 * unless the file is built with -fopenmp (libgomp is not in this image).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float act_apply(float v, int act) {
  if (act == 0) return v / (1.0f + expf(-v));            /* SiLU / Swish: x * sigmoid(x) */
  if (act == 1) return v > 0.0f ? v : 0.0f;              /* ReLU */
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); /* GELU (erf form) */
}

/* logits[e] = sum_k cat(embed, x)[k] * Wr[k, e] + br[e], in double */
static void router_logits(const float* x, const float* embed, const float* Wr, const float* br, int D, int Demb, int E,
                          double* logits) {
  for (int e = 0; e < E; ++e) logits[e] = br ? (double)br[e] : 0.0;
  for (int k = 0; k < Demb; ++k)
    for (int e = 0; e < E; ++e) logits[e] += (double)embed[k] * (double)Wr[(size_t)k * E + e];
  for (int k = 0; k < D; ++k)
    for (int e = 0; e < E; ++e) logits[e] += (double)x[k] * (double)Wr[(size_t)(Demb + k) * E + e];
}

/* softmax over all experts then max; ties -> lowest index. idx [S], value [S] */
void oracle_gate_3m(const float* x, const float* embed, const float* Wr, const float* br, int S, int D, int Demb, int E,
                    int32_t* idx, float* value) {
#pragma omp parallel for
  for (int s = 0; s < S; ++s) {
    double* l = (double*)malloc(sizeof(double) * E);
    router_logits(x + (size_t)s * D, embed ? embed + (size_t)s * Demb : NULL, Wr, br, D, embed ? Demb : 0, E, l);
    int best = 0;
    for (int e = 1; e < E; ++e)
      if (l[e] > l[best]) best = e;
    double sum = 0.0;
    for (int e = 0; e < E; ++e) sum += exp(l[e] - l[best]);
    idx[s] = best;
    value[s] = (float)(1.0 / sum);
    free(l);
  }
}

/* top-k of the logits (descending, ties -> lowest index), softmax over the k selected. idx/score [S, k] */
void oracle_gate_naive(const float* x, const float* W, const float* b, int S, int D, int E, int k, int32_t* idx,
                       float* score) {
#pragma omp parallel for
  for (int s = 0; s < S; ++s) {
    double* l = (double*)malloc(sizeof(double) * E);
    char* used = (char*)calloc(E, 1);
    double sel[16];
    router_logits(x + (size_t)s * D, NULL, W, b, D, 0, E, l);
    for (int j = 0; j < k; ++j) {
      int best = -1;
      for (int e = 0; e < E; ++e)
        if (!used[e] && (best < 0 || l[e] > l[best])) best = e;
      used[best] = 1;
      idx[(size_t)s * k + j] = best;
      sel[j] = l[best];
    }
    double sum = 0.0;
    for (int j = 0; j < k; ++j) sum += exp(sel[j] - sel[0]);
    for (int j = 0; j < k; ++j) score[(size_t)s * k + j] = (float)(exp(sel[j] - sel[0]) / sum);
    free(l);
    free(used);
  }
}

/* stable counting sort of n entries by expert; entries with idx outside [0, E) are dropped (mapping -1).
 * counts [E], offsets [E+1], mapping [n], pos [n] (first n_valid entries used). returns n_valid */
int oracle_prepare(const int32_t* idx, int n, int E, int32_t* counts, int32_t* offsets, int32_t* mapping, int32_t* pos) {
  memset(counts, 0, sizeof(int32_t) * E);
  for (int i = 0; i < n; ++i)
    if (idx[i] >= 0 && idx[i] < E) counts[idx[i]]++;
  offsets[0] = 0;
  for (int e = 0; e < E; ++e) offsets[e + 1] = offsets[e] + counts[e];
  int32_t* cursor = (int32_t*)malloc(sizeof(int32_t) * E);
  memcpy(cursor, offsets, sizeof(int32_t) * E);
  for (int i = 0; i < n; ++i) {
    if (idx[i] >= 0 && idx[i] < E) {
      mapping[i] = cursor[idx[i]]++;
      pos[mapping[i]] = i;
    } else {
      mapping[i] = -1;
    }
  }
  free(cursor);
  return offsets[E];
}

/* ybuf[rows of e] = act(xbuf_e . W1[e]^T + b1[e]) . W2[e]^T + b2[e]; W1 [E,H,D], W2 [E,D,H] */
void oracle_expert_ffn(const float* xbuf, const int32_t* offsets, const float* W1, const float* b1, const float* W2,
                       const float* b2, int E, int D, int H, int act, float* ybuf) {
  for (int e = 0; e < E; ++e) {
#pragma omp parallel for
    for (int r = offsets[e]; r < offsets[e + 1]; ++r) {
      float* h = (float*)malloc(sizeof(float) * H);
      const float* xr = xbuf + (size_t)r * D;
      for (int j = 0; j < H; ++j) {
        const float* w = W1 + ((size_t)e * H + j) * D;
        float acc = 0.0f;
        for (int k = 0; k < D; ++k) acc += xr[k] * w[k];
        h[j] = act_apply(acc + (b1 ? b1[(size_t)e * H + j] : 0.0f), act);
      }
      for (int d = 0; d < D; ++d) {
        const float* w = W2 + ((size_t)e * D + d) * H;
        float acc = 0.0f;
        for (int j = 0; j < H; ++j) acc += h[j] * w[j];
        ybuf[(size_t)r * D + d] = acc + (b2 ? b2[(size_t)e * D + d] : 0.0f);
      }
      free(h);
    }
  }
}

/* out[s] = (residual ? residual[s] : 0) + ff_scale * sum_j (score ? score[s,j] : 1) * ybuf[mapping[s*k+j]] */
void oracle_combine(const float* ybuf, const int32_t* mapping, const float* score, const float* residual, float ff_scale,
                    int S, int D, int k, float* out) {
#pragma omp parallel for
  for (int s = 0; s < S; ++s) {
    for (int d = 0; d < D; ++d) {
      float acc = 0.0f;
      for (int j = 0; j < k; ++j) {
        const int32_t row = mapping[(size_t)s * k + j];
        if (row < 0) continue;
        acc += (score ? score[(size_t)s * k + j] : 1.0f) * ybuf[(size_t)row * D + d];
      }
      out[(size_t)s * D + d] = (residual ? residual[(size_t)s * D + d] : 0.0f) + ff_scale * acc;
    }
  }
}

/* nn.LayerNorm(D, eps) over every row (trainer_3m_fix/layer/fmoe_transformer.py:54-65; the TensorRT build uses
 * TRTAPI++/plugin/layer_norm_plugin/layer_norm_kernel.cu): biased variance, eps inside the root. in == out allowed. */
void oracle_layer_norm(const float* in, const float* gamma, const float* beta, float eps, int S, int D, float* out) {
  for (int s = 0; s < S; ++s) {
    const float* x = in + (size_t)s * D;
    float* y = out + (size_t)s * D;
    double mean = 0.0, var = 0.0;
    for (int d = 0; d < D; ++d) mean += x[d];
    mean /= D;
    for (int d = 0; d < D; ++d) var += (x[d] - mean) * (x[d] - mean);
    var /= D;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    for (int d = 0; d < D; ++d) y[d] = (float)((x[d] - mean) * rstd * gamma[d] + beta[d]);
  }
}

/* The whole 3M layer (top-1): gate -> prepare -> scatter -> FFN -> combine. Scratch is allocated here. */
void oracle_moe_forward_3m(const float* x, const float* embed, const float* Wr, const float* br, const float* W1,
                           const float* b1, const float* W2, const float* b2, const float* residual, float ff_scale,
                           int S, int D, int Demb, int E, int H, int act, int32_t* idx, float* value, int32_t* counts,
                           int32_t* mapping, float* out) {
  int32_t* offsets = (int32_t*)malloc(sizeof(int32_t) * (E + 1));
  int32_t* pos = (int32_t*)malloc(sizeof(int32_t) * (S > 0 ? S : 1));
  float* xbuf = (float*)malloc(sizeof(float) * (size_t)(S > 0 ? S : 1) * D);
  float* ybuf = (float*)malloc(sizeof(float) * (size_t)(S > 0 ? S : 1) * D);
  oracle_gate_3m(x, embed, Wr, br, S, D, Demb, E, idx, value);
  const int nv = oracle_prepare(idx, S, E, counts, offsets, mapping, pos);
  for (int r = 0; r < nv; ++r) memcpy(xbuf + (size_t)r * D, x + (size_t)pos[r] * D, sizeof(float) * D);
  oracle_expert_ffn(xbuf, offsets, W1, b1, W2, b2, E, D, H, act, ybuf);
  oracle_combine(ybuf, mapping, value, residual, ff_scale, S, D, 1, out);
  free(offsets);
  free(pos);
  free(xbuf);
  free(ybuf);
}

/* The feed-forward part of a Conformer block (layer/fmoe_transformer.py:144-166):
 * out = norm_final(x + ff_scale * MoE(norm_ff(x), embed)); a NULL gamma skips that norm. */
void oracle_moe_block_forward_3m(const float* x, const float* embed, const float* Wr, const float* br, const float* W1,
                                 const float* b1, const float* W2, const float* b2, const float* ff_gamma,
                                 const float* ff_beta, const float* final_gamma, const float* final_beta, float eps,
                                 float ff_scale, int S, int D, int Demb, int E, int H, int act, int32_t* idx,
                                 float* value, int32_t* counts, int32_t* mapping, float* out) {
  float* xn = (float*)malloc(sizeof(float) * (size_t)(S > 0 ? S : 1) * D);
  float* moe = (float*)malloc(sizeof(float) * (size_t)(S > 0 ? S : 1) * D);
  if (ff_gamma) oracle_layer_norm(x, ff_gamma, ff_beta, eps, S, D, xn);
  else memcpy(xn, x, sizeof(float) * (size_t)S * D);
  oracle_moe_forward_3m(xn, embed, Wr, br, W1, b1, W2, b2, NULL, 1.0f, S, D, Demb, E, H, act, idx, value, counts,
                        mapping, moe);
  for (size_t i = 0; i < (size_t)S * D; ++i) out[i] = x[i] + ff_scale * moe[i];
  if (final_gamma) oracle_layer_norm(out, final_gamma, final_beta, eps, S, D, out);
  free(xn);
  free(moe);
}
