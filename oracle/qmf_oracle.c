/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.  See qmf_oracle.h for scope and pinning.
 *
 * Plain-C (scalar, single-threaded, no FMA contraction: build with -ffp-contract=off)
 * restatement of the reference's hot path.  Loop orders and floating-point association follow
 * the cited reference lines so that results agree with the reference compiled for x86-64
 * (which has no FMA in its baseline ISA) to the last bit wherever the reference itself is
 * deterministic; the only exception is the linear solve, where the reference calls an external
 * LAPACK (dsysv_) whose BLAS kernels may associate differently (agreement ~1e-14 relative).
 */
#include "qmf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ========================================================================== WALS ========= */

typedef struct {
  int64_t row, col, pos;
} coo_key;

static int cmp_coo(const void* a, const void* b) {
  const coo_key* x = (const coo_key*)a;
  const coo_key* y = (const coo_key*)b;
  /* WALSEngine::sortDataset comparator, WALSEngine.cpp:156-163 */
  if (x->row != y->row) return x->row < y->row ? -1 : 1;
  if (x->col != y->col) return x->col < y->col ? -1 : 1;
  /* std::sort is not stable; exact duplicates (same row AND col) may come out in either order
   * in the reference.  We break ties by input position; tests compare duplicates as multisets. */
  return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

int64_t qmfo_group_signals(const int64_t* row_id, const int64_t* col_id, int64_t n, int64_t* perm,
                           int64_t* row_ids, int64_t* row_ptr) {
  coo_key* keys = (coo_key*)malloc(sizeof(coo_key) * (size_t)(n > 0 ? n : 1));
  for (int64_t p = 0; p < n; ++p) {
    keys[p].row = row_id[p];
    keys[p].col = col_id[p];
    keys[p].pos = p;
  }
  qsort(keys, (size_t)n, sizeof(coo_key), cmp_coo);
  /* WALSEngine::groupSignals, WALSEngine.cpp:130-154: a new group whenever the row id changes;
   * index idx == rank of the id (CHECK_EQ(idx, i) at :150-153) */
  int64_t nrows = 0;
  for (int64_t p = 0; p < n; ++p) {
    if (p == 0 || keys[p].row != keys[p - 1].row) {
      row_ids[nrows] = keys[p].row;
      row_ptr[nrows] = p;
      ++nrows;
    }
    perm[p] = keys[p].pos;
  }
  row_ptr[nrows] = n;
  free(keys);
  return nrows;
}

void qmfo_gram(const double* Y, int64_t n, int64_t k, double* G) {
  /* WALSEngine.cpp:251 out->clear(); :256-263 row-major triple loop, r outermost */
  for (int64_t e = 0; e < k * k; ++e) G[e] = 0.0;
  for (int64_t r = 0; r < n; ++r) {
    const double* y = Y + r * k;
    for (int64_t i = 0; i < k; ++i) {
      for (int64_t j = 0; j < k; ++j) {
        G[i * k + j] += y[i] * y[j];
      }
    }
  }
}

/* ---- LAPACK dsysv('U') = dsytf2 (lwork = n forces the unblocked path inside dsytrf) + dsytrs.
 * Column-major accessor on a symmetric matrix stored as a full square: element (i,j) with i<=j
 * (upper triangle) lives at A[i + j*n].  Since the caller's matrix is symmetric and the
 * reference transposes before the call (Matrix.cpp:87), row-major input is valid input here. */
#define AU(i, j) A[(i) + (j) * n]

static int64_t idamax_col(const double* A, int64_t n, int64_t j, int64_t lo, int64_t hi) {
  /* argmax_{lo<=i<hi} |A(i,j)|, first maximum (BLAS idamax semantics) */
  int64_t best = lo;
  double bv = -1.0;
  for (int64_t i = lo; i < hi; ++i) {
    const double v = fabs(AU(i, j));
    if (v > bv) {
      bv = v;
      best = i;
    }
  }
  return best;
}

int qmfo_sysv_upper(double* A, int64_t n, double* b, int32_t* ipiv) {
  const double alpha = (1.0 + sqrt(17.0)) / 8.0;
  int info = 0;
  /* ---------------- dsytf2, UPLO = 'U': K runs from N down to 1 (0-based: k = n-1 .. 0) ---- */
  int64_t k = n - 1;
  while (k >= 0) {
    int64_t kstep = 1, kp = k;
    const double absakk = fabs(AU(k, k));
    int64_t imax = 0;
    double colmax = 0.0;
    if (k > 0) {
      imax = idamax_col(A, n, k, 0, k);
      colmax = fabs(AU(imax, k));
    }
    if (fmax(absakk, colmax) == 0.0 || isnan(absakk)) {
      if (info == 0) info = (int)(k + 1);
      kp = k;
    } else {
      if (absakk >= alpha * colmax) {
        kp = k; /* 1x1 pivot, no interchange */
      } else {
        /* largest off-diagonal element in row imax */
        int64_t jmax = imax + 1;
        double rowmax = 0.0;
        {
          double bv = -1.0;
          for (int64_t j = imax + 1; j <= k; ++j) {
            const double v = fabs(AU(imax, j));
            if (v > bv) {
              bv = v;
              jmax = j;
            }
          }
          rowmax = fabs(AU(imax, jmax));
        }
        if (imax > 0) {
          jmax = idamax_col(A, n, imax, 0, imax);
          rowmax = fmax(rowmax, fabs(AU(jmax, imax)));
        }
        if (absakk >= alpha * colmax * (colmax / rowmax)) {
          kp = k;
        } else if (fabs(AU(imax, imax)) >= alpha * rowmax) {
          kp = imax;
        } else {
          kp = imax;
          kstep = 2;
        }
      }
      const int64_t kk = k - kstep + 1;
      if (kp != kk) {
        /* interchange rows and columns kk and kp in the leading submatrix A(0:k,0:k) */
        for (int64_t i = 0; i < kp; ++i) {
          const double t = AU(i, kk);
          AU(i, kk) = AU(i, kp);
          AU(i, kp) = t;
        }
        for (int64_t j = kp + 1; j < kk; ++j) {
          const double t = AU(j, kk);
          AU(j, kk) = AU(kp, j);
          AU(kp, j) = t;
        }
        {
          const double t = AU(kk, kk);
          AU(kk, kk) = AU(kp, kp);
          AU(kp, kp) = t;
        }
        if (kstep == 2) {
          const double t = AU(k - 1, k);
          AU(k - 1, k) = AU(kp, k);
          AU(kp, k) = t;
        }
      }
      if (kstep == 1) {
        /* A := A - U(k) D(k) U(k)^T = A - (1/d) x x^T (dsyr, upper), then x := x/d (dscal) */
        const double r1 = 1.0 / AU(k, k);
        for (int64_t j = 0; j < k; ++j) {
          if (AU(j, k) != 0.0) {
            const double temp = -r1 * AU(j, k);
            for (int64_t i = 0; i <= j; ++i) AU(i, j) += AU(i, k) * temp;
          }
        }
        for (int64_t i = 0; i < k; ++i) AU(i, k) *= r1;
      } else if (k > 1) {
        /* 2x2 pivot block D(k) in rows/cols k-1, k */
        double d12 = AU(k - 1, k);
        const double d22 = AU(k - 1, k - 1) / d12;
        const double d11 = AU(k, k) / d12;
        const double t = 1.0 / (d11 * d22 - 1.0);
        d12 = t / d12;
        for (int64_t j = k - 2; j >= 0; --j) {
          const double wkm1 = d12 * (d11 * AU(j, k - 1) - AU(j, k));
          const double wk = d12 * (d22 * AU(j, k) - AU(j, k - 1));
          for (int64_t i = j; i >= 0; --i) {
            AU(i, j) = AU(i, j) - AU(i, k) * wk - AU(i, k - 1) * wkm1;
          }
          AU(j, k) = wk;
          AU(j, k - 1) = wkm1;
        }
      }
    }
    if (kstep == 1) {
      ipiv[k] = (int32_t)(kp + 1);
    } else {
      ipiv[k] = -(int32_t)(kp + 1);
      ipiv[k - 1] = -(int32_t)(kp + 1);
    }
    k -= kstep;
  }
  if (info != 0) return info;

  /* ---------------- dsytrs, UPLO = 'U', one right-hand side -------------------------------- */
  /* solve U*D*x = b */
  k = n - 1;
  while (k >= 0) {
    if (ipiv[k] > 0) {
      const int64_t kp = ipiv[k] - 1;
      if (kp != k) {
        const double t = b[k];
        b[k] = b[kp];
        b[kp] = t;
      }
      for (int64_t i = 0; i < k; ++i) b[i] -= AU(i, k) * b[k];
      b[k] *= 1.0 / AU(k, k);
      k -= 1;
    } else {
      const int64_t kp = -ipiv[k] - 1;
      if (kp != k - 1) {
        const double t = b[k - 1];
        b[k - 1] = b[kp];
        b[kp] = t;
      }
      for (int64_t i = 0; i < k - 1; ++i) b[i] -= AU(i, k) * b[k];
      for (int64_t i = 0; i < k - 1; ++i) b[i] -= AU(i, k - 1) * b[k - 1];
      const double akm1k = AU(k - 1, k);
      const double akm1 = AU(k - 1, k - 1) / akm1k;
      const double ak = AU(k, k) / akm1k;
      const double denom = akm1 * ak - 1.0;
      const double bkm1 = b[k - 1] / akm1k;
      const double bk = b[k] / akm1k;
      b[k - 1] = (ak * bkm1 - bk) / denom;
      b[k] = (akm1 * bk - bkm1) / denom;
      k -= 2;
    }
  }
  /* solve U^T x = b */
  k = 0;
  while (k < n) {
    if (ipiv[k] > 0) {
      double s = 0.0;
      for (int64_t i = 0; i < k; ++i) s += AU(i, k) * b[i];
      b[k] -= s;
      const int64_t kp = ipiv[k] - 1;
      if (kp != k) {
        const double t = b[k];
        b[k] = b[kp];
        b[kp] = t;
      }
      k += 1;
    } else {
      double s0 = 0.0, s1 = 0.0;
      for (int64_t i = 0; i < k; ++i) s0 += AU(i, k) * b[i];
      for (int64_t i = 0; i < k; ++i) s1 += AU(i, k + 1) * b[i];
      b[k] -= s0;
      b[k + 1] -= s1;
      const int64_t kp = -ipiv[k] - 1;
      if (kp != k) {
        const double t = b[k];
        b[k] = b[kp];
        b[kp] = t;
      }
      k += 2;
    }
  }
  return 0;
}
#undef AU

double qmfo_wals_update_row(const double* Y, int64_t k, const int32_t* cols, const double* vals, int64_t nnz,
                            const double* YtY, double alpha, double lambda, double* x) {
  double loss = 0.0;
  double* A = (double*)malloc(sizeof(double) * (size_t)(k * k));
  double* B = (double*)malloc(sizeof(double) * (size_t)(k * k));
  double* b = (double*)calloc((size_t)k, sizeof(double));
  int32_t* ipiv = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
  memcpy(A, YtY, sizeof(double) * (size_t)(k * k)); /* Matrix A passed by value, WALSEngine.cpp:271 */
  for (int64_t s = 0; s < nnz; ++s) {
    const double* y = Y + (int64_t)cols[s] * k;
    const double v = vals[s];
    for (int64_t i = 0; i < k; ++i) {
      b[i] += y[i] * (1.0 + alpha * v); /* :280 */
      for (int64_t j = 0; j < k; ++j) {
        A[i * k + j] += y[i] * alpha * v * y[j]; /* :282, left-to-right product */
      }
    }
    loss += 1.0 + alpha * v; /* :286 */
  }
  memcpy(B, A, sizeof(double) * (size_t)(k * k)); /* :289 */
  for (int64_t i = 0; i < k; ++i) A[i * k + i] += lambda; /* :290-292 */
  for (int64_t i = 0; i < k; ++i) x[i] = b[i];
  const int info = qmfo_sysv_upper(A, k, x, ipiv); /* :294 -> Matrix.cpp:81-96 */
  if (info != 0) {
    fprintf(stderr, "qmf_oracle: dsysv failed, code %d\n", info); /* CHECK_EQ(result, 0), Matrix.cpp:94 */
    abort();
  }
  for (int64_t i = 0; i < k; ++i) {
    for (int64_t j = 0; j < k; ++j) {
      loss += B[i * k + j] * x[i] * x[j]; /* :296-300 */
    }
  }
  for (int64_t i = 0; i < k; ++i) loss -= 2 * x[i] * b[i]; /* :302-304 */
  free(A);
  free(B);
  free(b);
  free(ipiv);
  return loss;
}

double qmfo_wals_half_step(double* X, int64_t nleft, const double* Y, int64_t nright, int64_t k,
                           const int64_t* row_ptr, const int32_t* cols, const double* vals, double alpha,
                           double lambda, int64_t nusers, int64_t nitems, int64_t nthreads) {
  double* G = (double*)malloc(sizeof(double) * (size_t)(k * k));
  memset(X, 0, sizeof(double) * (size_t)(nleft * k)); /* :170-171 */
  qmfo_gram(Y, nright, k, G);                           /* :177-178 */
  if (nthreads < 1) nthreads = 1;
  /* mapReduce(ntasks): thread t folds tasks t, t+T, ... from 0.0; the partials are then folded
   * in thread order from 0.0 (ParallelExecutor-inl.h:45-57) */
  double total = 0.0;
  for (int64_t t = 0; t < nthreads; ++t) {
    double part = 0.0;
    for (int64_t r = t; r < nleft; r += nthreads) {
      part = part + qmfo_wals_update_row(Y, k, cols + row_ptr[r], vals + row_ptr[r], row_ptr[r + 1] - row_ptr[r], G,
                                         alpha, lambda, X + r * k);
    }
    total = total + part;
  }
  free(G);
  return total / (double)nusers / (double)nitems; /* :215 */
}

/* =========================================================================== BPR ========= */

double qmfo_bpr_predict_difference(const double* P, const double* Q, const double* bias, int64_t k, int64_t u,
                                   int64_t i, int64_t j) {
  double pred = 0.0;
  if (bias != NULL) pred += bias[i] - bias[j]; /* BPREngine.cpp:227-229 */
  for (int64_t f = 0; f < k; ++f) {
    pred += P[u * k + f] * (Q[i * k + f] - Q[j * k + f]); /* :230-233 */
  }
  return pred;
}

double qmfo_bpr_update(double* P, double* Q, double* bias, int64_t k, int64_t u, int64_t i, int64_t j, double lr,
                       double user_lambda, double item_lambda, double bias_lambda) {
  /* e = 1 / (1 + exp(x)), BPREngine.cpp:241-244 */
  const double e = 1.0 / (1.0 + exp(qmfo_bpr_predict_difference(P, Q, bias, k, u, i, j)));
  if (bias != NULL) { /* :189-196 */
    double step = lr * (e - bias_lambda * bias[i]);
    bias[i] += step;
    step = lr * (-e - bias_lambda * bias[j]);
    bias[j] += step;
  }
  for (int64_t f = 0; f < k; ++f) { /* :200-205, old q_i, q_j */
    const double step = lr * (e * (Q[i * k + f] - Q[j * k + f]) - user_lambda * P[u * k + f]);
    P[u * k + f] += step;
  }
  for (int64_t f = 0; f < k; ++f) { /* :208-212, NEW p_u */
    const double step = lr * (e * P[u * k + f] - item_lambda * Q[i * k + f]);
    Q[i * k + f] += step;
  }
  for (int64_t f = 0; f < k; ++f) { /* :215-219 */
    const double step = lr * (-e * P[u * k + f] - item_lambda * Q[j * k + f]);
    Q[j * k + f] += step;
  }
  return e;
}

double qmfo_bpr_eval_loss(const double* P, const double* Q, const double* bias, int64_t k, const int64_t* u,
                          const int64_t* i, const int64_t* j, int64_t n, int64_t nthreads) {
  if (n == 0) return -1.0; /* BPREngine.cpp:254 */
  if (nthreads < 1) nthreads = 1;
  const int64_t block = n / nthreads; /* ParallelExecutor-inl.h:72: tail n % nthreads is dropped */
  double total = 0.0;
  for (int64_t t = 0; t < nthreads; ++t) {
    double part = 0.0;
    const int64_t lo = t * block;
    const int64_t hi = (t + 1) * block < n ? (t + 1) * block : n;
    for (int64_t p = lo; p < hi; ++p) {
      /* loss = log(1 + exp(-x)), BPREngine.cpp:237-239 */
      part = part + log(1.0 + exp(-qmfo_bpr_predict_difference(P, Q, bias, k, u[p], i[p], j[p])));
    }
    total = total + part;
  }
  return total / (double)n;
}

/* ---- std::mt19937 (32-bit Mersenne twister, ISO C++ [rand.eng.mers] parameters) ------------ */
typedef struct {
  uint32_t mt[624];
  int idx;
} mt19937_t;

static void mt_seed(mt19937_t* g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

static uint32_t mt_next(mt19937_t* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* libstdc++ (GCC >= 11) uniform_int_distribution<int>(0, range-1) on a 32-bit URBG: Lemire's
 * nearly-divisionless method (bits/uniform_int_dist.h, _S_nd) */
static uint32_t uniform_below(mt19937_t* g, uint32_t range) {
  uint64_t product = (uint64_t)mt_next(g) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    const uint32_t threshold = (uint32_t)(-range) % range;
    while (low < threshold) {
      product = (uint64_t)mt_next(g) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (uint32_t)(product >> 32);
}

void qmfo_bpr_sample_negatives(const int64_t* u, int64_t npairs, int64_t num_neg, int64_t nitems,
                               const int64_t* pos_ptr, const int64_t* pos_items, uint32_t seed, int64_t* neg_out) {
  mt19937_t g;
  mt_seed(&g, seed);
  for (int64_t p = 0; p < npairs; ++p) { /* BPREngine::iterate, BPREngine-inl.h:19-29 */
    const int64_t lo = pos_ptr[u[p]], hi = pos_ptr[u[p] + 1];
    for (int64_t q = 0; q < num_neg; ++q) {
      int64_t neg;
      for (;;) { /* sampleRandomNegative, BPREngine-inl.h:52-59 */
        neg = (int64_t)uniform_below(&g, (uint32_t)nitems);
        int found = 0; /* userPosSet.count(negIdx) > 0: binary search over the sorted set */
        int64_t a = lo, b = hi;
        while (a < b) {
          const int64_t m = (a + b) / 2;
          if (pos_items[m] < neg) {
            a = m + 1;
          } else {
            b = m;
          }
        }
        found = (a < hi && pos_items[a] == neg);
        if (!found) break;
      }
      neg_out[p * num_neg + q] = neg;
    }
  }
}

/* ==================================================================== evaluation ========= */

void qmfo_compute_test_scores(const double* U, const double* V, const double* bias, int64_t ni, int64_t k,
                              const int64_t* test_users, int64_t nT, double* out) {
  for (int64_t t = 0; t < nT; ++t) {
    const double* pu = U + test_users[t] * k;
    for (int64_t idx = 0; idx < ni; ++idx) {
      double s = bias != NULL ? bias[idx] : 0.0; /* Engine.cpp:86-87 */
      for (int64_t f = 0; f < k; ++f) s += pu[f] * V[idx * k + f]; /* :88-91 */
      out[t * ni + idx] = s;
    }
  }
}

typedef struct {
  double score;
  int label;
} scored_t;

static int cmp_scored_desc(const void* a, const void* b) {
  /* std::greater<std::pair<double,bool>>: score descending, then label true before false */
  const scored_t* x = (const scored_t*)a;
  const scored_t* y = (const scored_t*)b;
  if (x->score != y->score) return x->score > y->score ? -1 : 1;
  return y->label - x->label;
}

static scored_t* sorted_pairs(const double* labels, const double* scores, int64_t n, int32_t* npos) {
  scored_t* s = (scored_t*)malloc(sizeof(scored_t) * (size_t)(n > 0 ? n : 1));
  int32_t pos = 0;
  for (int64_t i = 0; i < n; ++i) {
    s[i].score = scores[i];
    s[i].label = labels[i] > 0.0;
    pos += s[i].label;
  }
  qsort(s, (size_t)n, sizeof(scored_t), cmp_scored_desc);
  *npos = pos;
  return s;
}

double qmfo_metric_one(int kind, int64_t k, const double* labels, const double* scores, int64_t n) {
  if (kind == 0) { /* MeanSquaredError, Metrics.cpp:54-63 */
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) sum += pow(labels[i] - scores[i], 2);
    return sum / (double)n;
  }
  int32_t pos = 0;
  scored_t* s = sorted_pairs(labels, scores, n, &pos);
  double r = 0.0;
  if (kind == 1) { /* AUC, Metrics.cpp:65-99 */
    const int32_t neg = (int32_t)n - pos;
    if (pos == 0 || neg == 0) {
      r = 1.0; /* :80-83 */
    } else {
      int tp = 0;
      double auc = 0;
      for (int64_t i = 0; i < n; ++i) {
        if (s[i].label) {
          ++tp;
        } else {
          auc += (double)tp / pos / neg; /* :93 */
        }
      }
      r = auc;
    }
  } else if (kind == 2) { /* AveragePrecision, Metrics.cpp:139-164 */
    double ap = 0.0;
    int32_t p = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (s[i].label) {
        ++p;
        ap += (double)p / (double)(i + 1);
      }
    }
    r = ap / pos;
  } else { /* Precision / Recall @k, Metrics.cpp:101-137 (nth_element == top-k set of the order) */
    int64_t c = 0;
    for (int64_t i = 0; i < k && i < n; ++i) c += s[i].label;
    r = kind == 3 ? (double)c / (double)k : (double)c / (double)pos;
  }
  free(s);
  return r;
}

double qmfo_metric_avg(int kind, int64_t k, const double* labels, const double* scores, int64_t nT, int64_t ni,
                       int64_t nthreads) {
  if (nthreads <= 0) { /* serial overload, Metrics.cpp:27-36 */
    double sum = 0.0;
    for (int64_t t = 0; t < nT; ++t) sum += qmfo_metric_one(kind, k, labels + t * ni, scores + t * ni, ni);
    return sum / (double)nT;
  }
  double total = 0.0; /* Metrics.cpp:43-51 via mapReduce(ntasks), strided */
  for (int64_t th = 0; th < nthreads; ++th) {
    double part = 0.0;
    for (int64_t t = th; t < nT; t += nthreads) {
      part = part + qmfo_metric_one(kind, k, labels + t * ni, scores + t * ni, ni);
    }
    total = total + part;
  }
  return total / (double)nT;
}

void qmfo_rank_stats(const double* labels, const double* scores, int64_t n, const int64_t* ks, int64_t nks,
                     int64_t* stats, double* ap_out) {
  int32_t pos = 0;
  scored_t* s = sorted_pairs(labels, scores, n, &pos);
  int64_t tp = 0, aucnum = 0;
  double ap = 0.0;
  for (int64_t q = 0; q < nks; ++q) stats[2 + q] = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (s[i].label) {
      ++tp;
      ap += (double)tp / (double)(i + 1);
      for (int64_t q = 0; q < nks; ++q) {
        if (i < ks[q]) stats[2 + q] += 1;
      }
    } else {
      aucnum += tp;
    }
  }
  stats[0] = pos;
  stats[1] = aucnum;
  if (ap_out != NULL) *ap_out = pos > 0 ? ap / pos : 0.0;
  free(s);
}

int64_t qmfo_save_factors(const double* F, const double* bias, const int64_t* ids, int64_t n, int64_t k, char* out,
                          int64_t cap) {
  /* std::fixed << setprecision(9) == printf("%.9f") (both round-to-nearest on the exact binary
   * value in glibc/libstdc++); Engine.cpp:108-121 */
  int64_t used = 0;
  char tmp[512];
  for (int64_t r = 0; r < n; ++r) {
    int len = snprintf(tmp, sizeof tmp, "%lld", (long long)ids[r]);
    if (used + len <= cap && out) memcpy(out + used, tmp, (size_t)len);
    used += len;
    if (bias != NULL) {
      len = snprintf(tmp, sizeof tmp, " %.9f", bias[r]);
      if (used + len <= cap && out) memcpy(out + used, tmp, (size_t)len);
      used += len;
    }
    for (int64_t f = 0; f < k; ++f) {
      len = snprintf(tmp, sizeof tmp, " %.9f", F[r * k + f]);
      if (used + len <= cap && out) memcpy(out + used, tmp, (size_t)len);
      used += len;
    }
    if (used + 1 <= cap && out) out[used] = '\n';
    used += 1;
  }
  return used;
}
