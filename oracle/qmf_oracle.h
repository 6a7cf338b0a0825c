/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * Plain-C restatement of the reference's (taozhijiang/qmf) training hot path, used ONLY as the
 * checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing under
 * qmf_b200/ may include, link or call this.  Each function cites the reference file:line it
 * restates.  Parity is PINNED: tests/test_oracle_*.py check this file against (a) the golden
 * vectors of the reference's own gtests (qmf/test/MetricsTest.cpp:35-88,
 * EngineTest.cpp:75-139, WALSEngineTest.cpp:112-205, MatrixTest.cpp:92-116) and (b) outputs of
 * the unmodified reference compiled here (oracle/_ref/libqmf_ref.so) on seeded inputs, with
 * fixtures committed under tests/golden/.
 */
#ifndef QMF_ORACLE_H
#define QMF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- WALS ------------------------------------------------------------------------------- */
/* WALSEngine::groupSignals + sortDataset (qmf/wals/WALSEngine.cpp:130-163): sort COO by
 * (row id, col id), group into rows in ascending raw-id order, duplicates kept.  perm[] receives
 * the sorted order (indices into the input), row_ids[] the distinct row ids, row_ptr[] (nrows+1)
 * the CSR offsets.  Returns nrows. */
int64_t qmfo_group_signals(const int64_t* row_id, const int64_t* col_id, int64_t n, int64_t* perm,
                           int64_t* row_ids, int64_t* row_ptr);
/* WALSEngine::computeXtX(const Matrix&, Matrix*) run with OMP_NUM_THREADS=1
 * (WALSEngine.cpp:246-264): G(i,j) = sum_r Y(r,i)*Y(r,j), r ascending. */
void qmfo_gram(const double* Y, int64_t n, int64_t k, double* G);
/* LAPACK dsysv_("Upper", n, 1, A, n, ipiv, b, n, work, lwork=n) as called from
 * linearSymmetricSolve (qmf/Matrix.cpp:81-96): unblocked Bunch-Kaufman U*D*U^T (dsytf2) +
 * dsytrs, restated from the published LAPACK 3.x algorithm.  A is n x n symmetric (row-major ==
 * column-major), only the upper triangle is read; A and b are overwritten.  Returns info. */
int qmfo_sysv_upper(double* A, int64_t n, double* b, int32_t* ipiv);
/* WALSEngine::updateFactorsForOne (WALSEngine.cpp:266-310): one row.  x (k) receives the solved
 * factors, return value is the row's loss term. */
double qmfo_wals_update_row(const double* Y, int64_t k, const int32_t* cols, const double* vals, int64_t nnz,
                            const double* YtY, double alpha, double lambda, double* x);
/* WALSEngine::iterate (WALSEngine.cpp:165-218): zero X, Gram of Y, every row, loss summed in
 * the strided per-thread order of ParallelExecutor::mapReduce (ParallelExecutor-inl.h:37-58)
 * for `nthreads` pool threads, divided by nusers and nitems.  X is nleft x k. */
double qmfo_wals_half_step(double* X, int64_t nleft, const double* Y, int64_t nright, int64_t k,
                           const int64_t* row_ptr, const int32_t* cols, const double* vals, double alpha,
                           double lambda, int64_t nusers, int64_t nitems, int64_t nthreads);

/* ---- BPR -------------------------------------------------------------------------------- */
/* BPREngine::predictDifference (qmf/bpr/BPREngine.cpp:222-235); biases may be NULL */
double qmfo_bpr_predict_difference(const double* P, const double* Q, const double* bias, int64_t k, int64_t u,
                                   int64_t i, int64_t j);
/* BPREngine::update (BPREngine.cpp:178-220): one SGD step in place.  Returns e. */
double qmfo_bpr_update(double* P, double* Q, double* bias, int64_t k, int64_t u, int64_t i, int64_t j, double lr,
                       double user_lambda, double item_lambda, double bias_lambda);
/* loss half of BPREngine::evaluate (BPREngine.cpp:246-261) with the block/tail-drop summation
 * of ParallelExecutor::mapReduce(elems) (ParallelExecutor-inl.h:60-85); returns mean (sum / n) */
double qmfo_bpr_eval_loss(const double* P, const double* Q, const double* bias, int64_t k, const int64_t* u,
                          const int64_t* i, const int64_t* j, int64_t n, int64_t nthreads);
/* std::mt19937 + libstdc++ std::uniform_int_distribution<int>(0, n-1) rejection sampling as in
 * BPREngine::sampleRandomNegative (BPREngine-inl.h:48-60) driving BPREngine::iterate
 * (BPREngine-inl.h:19-29): the fixed evaluation triplets.  pos_ptr/pos_items: per-user sorted
 * positive item sets (CSR).  Writes num_neg negatives per (u,i) pair into neg_out. */
void qmfo_bpr_sample_negatives(const int64_t* u, int64_t npairs, int64_t num_neg, int64_t nitems,
                               const int64_t* pos_ptr, const int64_t* pos_items, uint32_t seed, int64_t* neg_out);

/* ---- evaluation ------------------------------------------------------------------------- */
/* Engine::computeTestScores (qmf/Engine.cpp:73-96); out is nT x ni; bias may be NULL */
void qmfo_compute_test_scores(const double* U, const double* V, const double* bias, int64_t ni, int64_t k,
                              const int64_t* test_users, int64_t nT, double* out);
/* Metrics.cpp:54-164.  kind: 0 mse, 1 auc, 2 ap, 3 p@k, 4 r@k */
double qmfo_metric_one(int kind, int64_t k, const double* labels, const double* scores, int64_t n);
/* Metric::compute(labels, scores, parallel) (Metrics.cpp:38-52) — strided per-thread partial
 * sums (nthreads >= 1) or the serial overload (Metrics.cpp:27-36, nthreads == 0) */
double qmfo_metric_avg(int kind, int64_t k, const double* labels, const double* scores, int64_t nT, int64_t ni,
                       int64_t nthreads);
/* integer rank statistics of one user (closed forms of Metrics.cpp:65-164):
 * stats[0]=npos, stats[1]=sum over negatives of (#positives ranked before it) [AUC numerator],
 * stats[2+..] = #positives in the top-ks[q] for each q.  ap_out receives AP. */
void qmfo_rank_stats(const double* labels, const double* scores, int64_t n, const int64_t* ks, int64_t nks,
                     int64_t* stats, double* ap_out);

/* Engine::saveFactors (Engine.cpp:98-122): "<id>[ <bias>] <f0> ... <fk-1>\n", fixed, 9 decimals.
 * Returns bytes needed; writes at most cap bytes. */
int64_t qmfo_save_factors(const double* F, const double* bias, const int64_t* ids, int64_t n, int64_t k, char* out,
                          int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
