/*
 * kdcc_oracle.c -- CPU restatement of the distillation hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path
 * (libkdcc.so, CUDA, sm_100a) never links or calls anything in here.
 *
 * Parity status: PINNED.  Every function below is checked against outputs of the
 * reference's own modules executed in the build container (oracle/make_golden.py imports
 * them from /root/reference and freezes tests/golden/*.npz; tests/test_oracle_golden.py
 * replays them).  The reference ships no tests or golden vectors of its own.
 *
 * The reference keeps its arithmetic in PyTorch/ATen calls; each function cites the
 * reference call site whose semantics it restates:
 *   orc_dw_fwd / orc_dw_bwd    models/students/transform_blocks/depthwise_separable_conv.py:7-8,12
 *                              (nn.Conv2d, groups=C, stride 1, zero padding, dilation d)
 *   orc_pw_fwd / orc_pw_bwd    models/students/transform_blocks/depthwise_separable_conv.py:9,13
 *                              (nn.Conv2d 1x1)
 *   orc_kd_loss                losses/KLDiv.py:19-23 and losses/EnsembleKLDiv.py:18-22
 *   orc_hint_loss              losses/WeightedHintMSELoss.py:12-16 and losses/MSELoss.py:14-16
 *
 * Layout is the reference's: NCHW, contiguous, fp32.  Accumulation is in double so the
 * oracle is at least as exact as the fp32 ATen kernels it stands in for.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---------------------------------------------------------------------------------------
 * Depthwise k x k cross-correlation (depthwise_separable_conv.py:7-8,12).
 *   y[n,c,i,j] = b[c] + sum_{u,v<k} w[c,u,v] * x[n,c,i+u*d-p, j+v*d-p]     (zero outside)
 * Output size Ho = H + 2p - d(k-1), Wo likewise (stride 1).
 * ------------------------------------------------------------------------------------- */
ORC_API void orc_dw_fwd(const float *x, const float *w, const float *bias, float *y,
                        int N, int C, int H, int W, int k, int d, int p) {
  const int Ho = H + 2 * p - d * (k - 1), Wo = W + 2 * p - d * (k - 1);
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      const float *xp = x + ((size_t)n * C + c) * H * W;
      const float *wp = w + (size_t)c * k * k;
      float *yp = y + ((size_t)n * C + c) * Ho * Wo;
      for (int i = 0; i < Ho; ++i)
        for (int j = 0; j < Wo; ++j) {
          double acc = bias ? (double)bias[c] : 0.0;
          for (int u = 0; u < k; ++u) {
            const int ii = i + u * d - p;
            if (ii < 0 || ii >= H) continue;
            for (int v = 0; v < k; ++v) {
              const int jj = j + v * d - p;
              if (jj < 0 || jj >= W) continue;
              acc += (double)wp[u * k + v] * (double)xp[(size_t)ii * W + jj];
            }
          }
          yp[(size_t)i * Wo + j] = (float)acc;
        }
    }
}

/* Backward of orc_dw_fwd (autograd of the call at depthwise_separable_conv.py:12).
 *   dx[n,c,a,b]  = sum_{u,v} w[c,u,v] * dy[n,c,a-u*d+p, b-v*d+p]
 *   dw[c,u,v]    = sum_{n,i,j} dy[n,c,i,j] * x[n,c,i+u*d-p, j+v*d-p]
 *   dbias[c]     = sum_{n,i,j} dy[n,c,i,j]
 * dx / dw / dbias may each be NULL. */
ORC_API void orc_dw_bwd(const float *x, const float *w, const float *dy, float *dx, float *dw,
                        float *dbias, int N, int C, int H, int W, int k, int d, int p) {
  const int Ho = H + 2 * p - d * (k - 1), Wo = W + 2 * p - d * (k - 1);
  if (dx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c) {
        const float *gp = dy + ((size_t)n * C + c) * Ho * Wo;
        const float *wp = w + (size_t)c * k * k;
        float *dxp = dx + ((size_t)n * C + c) * H * W;
        for (int a = 0; a < H; ++a)
          for (int b = 0; b < W; ++b) {
            double acc = 0.0;
            for (int u = 0; u < k; ++u) {
              const int i = a - u * d + p;
              if (i < 0 || i >= Ho) continue;
              for (int v = 0; v < k; ++v) {
                const int j = b - v * d + p;
                if (j < 0 || j >= Wo) continue;
                acc += (double)wp[u * k + v] * (double)gp[(size_t)i * Wo + j];
              }
            }
            dxp[(size_t)a * W + b] = (float)acc;
          }
      }
  }
  if (dw || dbias) {
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; ++c) {
      double *acc = (double *)calloc((size_t)k * k + 1, sizeof(double));
      for (int n = 0; n < N; ++n) {
        const float *xp = x + ((size_t)n * C + c) * H * W;
        const float *gp = dy + ((size_t)n * C + c) * Ho * Wo;
        for (int i = 0; i < Ho; ++i)
          for (int j = 0; j < Wo; ++j) {
            const double g = gp[(size_t)i * Wo + j];
            acc[k * k] += g;
            for (int u = 0; u < k; ++u) {
              const int ii = i + u * d - p;
              if (ii < 0 || ii >= H) continue;
              for (int v = 0; v < k; ++v) {
                const int jj = j + v * d - p;
                if (jj < 0 || jj >= W) continue;
                acc[u * k + v] += g * (double)xp[(size_t)ii * W + jj];
              }
            }
          }
      }
      if (dw)
        for (int t = 0; t < k * k; ++t) dw[(size_t)c * k * k + t] = (float)acc[t];
      if (dbias) dbias[c] = (float)acc[k * k];
      free(acc);
    }
  }
}

/* ---------------------------------------------------------------------------------------
 * Pointwise 1x1 convolution (depthwise_separable_conv.py:9,13).
 *   y[n,o,q] = b[o] + sum_c w[o,c] * x[n,c,q],  q over H*W pixels
 * ------------------------------------------------------------------------------------- */
ORC_API void orc_pw_fwd(const float *x, const float *w, const float *bias, float *y, int N, int Ci,
                        int Co, long HW) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int o = 0; o < Co; ++o) {
      double *row = (double *)malloc(sizeof(double) * (size_t)HW);
      const double b = bias ? (double)bias[o] : 0.0;
      for (long q = 0; q < HW; ++q) row[q] = b;
      for (int c = 0; c < Ci; ++c) {
        const double wv = w[(size_t)o * Ci + c];
        const float *xp = x + ((size_t)n * Ci + c) * HW;
        for (long q = 0; q < HW; ++q) row[q] += wv * (double)xp[q];
      }
      float *yp = y + ((size_t)n * Co + o) * HW;
      for (long q = 0; q < HW; ++q) yp[q] = (float)row[q];
      free(row);
    }
}

/* Backward of orc_pw_fwd:  dx[n,c,q] = sum_o w[o,c] dy[n,o,q];  dw[o,c] = sum_{n,q} dy[n,o,q] x[n,c,q];
 * dbias[o] = sum_{n,q} dy[n,o,q].  Any output may be NULL. */
ORC_API void orc_pw_bwd(const float *x, const float *w, const float *dy, float *dx, float *dw,
                        float *dbias, int N, int Ci, int Co, long HW) {
  if (dx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < Ci; ++c) {
        double *row = (double *)calloc((size_t)HW, sizeof(double));
        for (int o = 0; o < Co; ++o) {
          const double wv = w[(size_t)o * Ci + c];
          const float *gp = dy + ((size_t)n * Co + o) * HW;
          for (long q = 0; q < HW; ++q) row[q] += wv * (double)gp[q];
        }
        float *dxp = dx + ((size_t)n * Ci + c) * HW;
        for (long q = 0; q < HW; ++q) dxp[q] = (float)row[q];
        free(row);
      }
  }
  if (dw) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int o = 0; o < Co; ++o)
      for (int c = 0; c < Ci; ++c) {
        double acc = 0.0;
        for (int n = 0; n < N; ++n) {
          const float *gp = dy + ((size_t)n * Co + o) * HW;
          const float *xp = x + ((size_t)n * Ci + c) * HW;
          for (long q = 0; q < HW; ++q) acc += (double)gp[q] * (double)xp[q];
        }
        dw[(size_t)o * Ci + c] = (float)acc;
      }
  }
  if (dbias) {
#pragma omp parallel for schedule(static)
    for (int o = 0; o < Co; ++o) {
      double acc = 0.0;
      for (int n = 0; n < N; ++n) {
        const float *gp = dy + ((size_t)n * Co + o) * HW;
        for (long q = 0; q < HW; ++q) acc += (double)gp[q];
      }
      dbias[o] = (float)acc;
    }
  }
}

/* ---------------------------------------------------------------------------------------
 * Temperature-scaled KL distillation loss (losses/KLDiv.py:19-23).
 *   p_s = log_softmax(s/T, dim=1); p_t = softmax(t/T, dim=1)
 *   loss = kl_div(p_s, p_t, reduction='mean' over ALL N*C*HW elements) * T^2 * C
 *        = T^2/(N*HW) * sum_{n,q} sum_c p_t * (log p_t - log p_s)          [xlogy: p_t==0 -> 0]
 *   dloss/ds[n,c,q] = T/(N*HW) * (softmax(s/T)[c] - p_t[c])
 * target_is_prob != 0 restates losses/EnsembleKLDiv.py:18-22: targets are already
 * probabilities (no teacher softmax; the reference fixes T=1 there); the gradient is then
 *   T/(N*HW) * (softmax(s/T)[c] * sum_c' p_t[c'] - p_t[c])
 * which is exact also when the ensemble target does not sum to one.
 * Layout (N, C, HW) contiguous.  ds may be NULL.  Returns the loss.
 * ------------------------------------------------------------------------------------- */
ORC_API double orc_kd_loss(const float *s, const float *t, float *ds, int N, int C, long HW, double T,
                           int target_is_prob) {
  double total = 0.0;
  const double invT = 1.0 / T;
  const double gscale = T / ((double)N * (double)HW);
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : total)
  for (int n = 0; n < N; ++n)
    for (long q = 0; q < HW; ++q) {
      const float *sp = s + (size_t)n * C * HW + q;
      const float *tp = t + (size_t)n * C * HW + q;
      double smax = -INFINITY, tmax = -INFINITY;
      for (int c = 0; c < C; ++c) {
        const double sv = sp[(size_t)c * HW] * invT, tv = tp[(size_t)c * HW] * invT;
        if (sv > smax) smax = sv;
        if (tv > tmax) tmax = tv;
      }
      double ssum = 0.0, tsum = 0.0;
      for (int c = 0; c < C; ++c) {
        ssum += exp(sp[(size_t)c * HW] * invT - smax);
        if (!target_is_prob) tsum += exp(tp[(size_t)c * HW] * invT - tmax);
      }
      const double slse = smax + log(ssum);
      const double tlse = target_is_prob ? 0.0 : tmax + log(tsum);
      double kl = 0.0, psum = 0.0;
      for (int c = 0; c < C; ++c) {
        const double logps = sp[(size_t)c * HW] * invT - slse;
        double pt, logpt;
        if (target_is_prob) {
          pt = tp[(size_t)c * HW];
          logpt = pt > 0.0 ? log(pt) : 0.0;
        } else {
          logpt = tp[(size_t)c * HW] * invT - tlse;
          pt = exp(logpt);
        }
        if (pt > 0.0) kl += pt * (logpt - logps);
        psum += pt;
      }
      total += kl;
      if (ds) {
        float *dp = ds + (size_t)n * C * HW + q;
        for (int c = 0; c < C; ++c) {
          const double ps = exp(sp[(size_t)c * HW] * invT - slse);
          const double pt = target_is_prob ? (double)tp[(size_t)c * HW]
                                           : exp(tp[(size_t)c * HW] * invT - tlse);
          dp[(size_t)c * HW] = (float)(gscale * (ps * (target_is_prob ? psum : 1.0) - pt));
        }
      }
    }
  return total * T * T / ((double)N * (double)HW);
}

/* ---------------------------------------------------------------------------------------
 * Channel-weighted hint loss (losses/WeightedHintMSELoss.py:12-16):
 *   m[n,c] = mean_{hw} (s-t)^2 ;  L = mean_n( sum_c w[n,c] m[n,c] / sum_c w[n,c] ) * scale
 *   dL/ds[n,c,q] = scale * 2 (s-t) w[n,c] / (sum_c w[n,c] * HW * N)
 * w == NULL means uniform weights, which with scale = num_classes is exactly
 * losses/MSELoss.py:14-16 (nn.MSELoss('mean') * num_classes).  w_per_sample selects a
 * (N,C) weight table instead of a broadcast (C,) vector.  Layout (N, C, HW) contiguous.
 * ------------------------------------------------------------------------------------- */
ORC_API double orc_hint_loss(const float *s, const float *t, const float *w, int w_per_sample,
                             float *ds, int N, int C, long HW, double scale) {
  double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int n = 0; n < N; ++n) {
    double wsum = 0.0;
    for (int c = 0; c < C; ++c) wsum += w ? (double)w[w_per_sample ? (size_t)n * C + c : c] : 1.0;
    double acc = 0.0;
    for (int c = 0; c < C; ++c) {
      const double wc = w ? (double)w[w_per_sample ? (size_t)n * C + c : c] : 1.0;
      const float *sp = s + ((size_t)n * C + c) * HW;
      const float *tp = t + ((size_t)n * C + c) * HW;
      double sq = 0.0;
      for (long q = 0; q < HW; ++q) {
        const double df = (double)sp[q] - (double)tp[q];
        sq += df * df;
      }
      acc += wc * sq / (double)HW;
      if (ds) {
        float *dp = ds + ((size_t)n * C + c) * HW;
        const double g = scale * 2.0 * wc / (wsum * (double)HW * (double)N);
        for (long q = 0; q < HW; ++q) dp[q] = (float)(g * ((double)sp[q] - (double)tp[q]));
      }
    }
    total += acc / wsum;
  }
  return scale * total / (double)N;
}

/* NCHW <-> NHWC repack helpers so tests can feed the product (channels-last) layout. */
ORC_API void orc_nchw_to_nhwc(const float *src, float *dst, int N, int C, long HW) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (long q = 0; q < HW; ++q)
      for (int c = 0; c < C; ++c) dst[((size_t)n * HW + q) * C + c] = src[((size_t)n * C + c) * HW + q];
}

ORC_API void orc_nhwc_to_nchw(const float *src, float *dst, int N, int C, long HW) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (long q = 0; q < HW; ++q) dst[((size_t)n * C + c) * HW + q] = src[((size_t)n * HW + q) * C + c];
}
