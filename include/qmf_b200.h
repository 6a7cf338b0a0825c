/* qmf_b200 — C ABI of the B200-native (sm_100a) replacement for the training hot path of
 * taozhijiang/qmf.  Plain pointers and sizes only; no C++/torch types.  All functions return 0 on
 * success or a negative QMFB_ERR_* code; qmfb_last_error() describes the last failure of the
 * calling thread.  The reference has no FFI of its own (single C++ process); each entry point
 * below cites the reference function whose inner loop it replaces, and INTEGRATION.md shows the
 * call a maintainer would add at that site.
 *
 * Two levels:
 *   qmfb_*_dev   kernel level: the caller owns device memory and the CUDA stream (used by the
 *                one-process-per-GPU driver that does its exchange steps with NCCL).
 *   qmfb_wals_* / qmfb_bpr_* / qmfb_eval_*   engine level: host buffers in, host buffers out;
 *                the library owns device memory.  This is what the C++ engines
 *                (qmf_b200/host/qmf/...) bind.
 *
 * Device layout: a factor matrix with k factors is stored row-major with row stride
 * KP = qmfb_padded_k(k) doubles (k rounded up to a multiple of 32, k <= 256), pad columns are zero.
 */
#ifndef QMF_B200_H
#define QMF_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMFB_OK 0
#define QMFB_ERR_INVALID (-1)    /* bad argument */
#define QMFB_ERR_CUDA (-2)       /* CUDA runtime failure (no device, launch failure, ...) */
#define QMFB_ERR_NOT_SPD (-3)    /* a row's normal equations were not positive definite; mirrors
                                    CHECK_EQ(result, 0) after dsysv_, qmf/Matrix.cpp:94 */
#define QMFB_ERR_NOT_FINITE (-4) /* non-finite BPR gradient; mirrors CHECK(std::isfinite(e)),
                                    qmf/bpr/BPREngine.cpp:184-185 */
#define QMFB_ERR_UNSUPPORTED (-5)

#define QMFB_SIDE_USER 0
#define QMFB_SIDE_ITEM 1

const char* qmfb_last_error(void);
int qmfb_version(void);
/* number of CUDA devices visible, or a negative error */
int qmfb_device_count(void);

/* ---------------------------------------------------------------- layout helpers ---------- */
int qmfb_padded_k(int k);                 /* KP; negative if k is unsupported (k < 1 or k > 256) */
int64_t qmfb_gram_packed_len(int k);      /* doubles in the packed upper-tile Gram of KP x KP */
int64_t qmfb_gram_workspace_len(int k);   /* doubles of scratch qmfb_gram_dev needs */

/* ---------------------------------------------------------------- kernel level (device) ---- */
/* Partial Gram G = sum_{r in [row_begin,row_end)} y_r y_r^T in packed upper-tile form
 * (replaces WALSEngine::computeXtX, qmf/wals/WALSEngine.cpp:246-264).  Deterministic.  Partial
 * results of several ranks are combined by summing gram_packed element-wise (NCCL allreduce). */
int qmfb_gram_dev(void* stream, const double* Y, int64_t ldy, int64_t row_begin, int64_t row_end, int k,
                  double* workspace, double* gram_packed);
/* The same Gram as a sum over a FIXED set of row parts, so that several devices can share the work and
 * still produce bit-identical results: the rows [row_begin, row_end) are cut into
 * P = qmfb_gram_parts_count(row_end - row_begin, k) parts that depend only on the row count;
 * qmfb_gram_parts_dev writes the partial Gram of parts [part_begin, part_end) into
 * workspace[part * qmfb_gram_packed_len(k) ...]; qmfb_gram_reduce_parts_dev sums all P parts in part
 * order, part p being read from workspaces[d] for part_end[d-1] <= p < part_end[d] (host arrays of
 * `nsrc` <= 16 entries; the pointers may be peer memory of other GPUs).  qmfb_gram_dev is the
 * nsrc == 1 case.  Y must be 16-byte aligned, ldy even; the pad columns [k, ldy) of Y must be zero. */
int qmfb_gram_parts_count(int64_t nrows, int k);
int qmfb_gram_parts_dev(void* stream, const double* Y, int64_t ldy, int64_t row_begin, int64_t row_end, int k, int part_begin,
                        int part_end, double* workspace);
int qmfb_gram_reduce_parts_dev(void* stream, const double* const* workspaces, const int* part_end, int nsrc, int k,
                               double* gram_packed);
/* packed upper tiles -> dense symmetric k x k row-major */
int qmfb_gram_unpack_dev(void* stream, const double* gram_packed, int k, double* out);
/* Solve `nrows` rows of one half-step (replaces the per-row body of WALSEngine::iterate,
 * qmf/wals/WALSEngine.cpp:205-214 -> updateFactorsForOne :266-310 -> dsysv_ Matrix.cpp:92):
 *   A = G + sum_s alpha r_s y_s y_s^T + lambda I,  b = sum_s (1 + alpha r_s) y_s,  x = A^{-1} b,
 *   row_loss = sum_s (1 + alpha r_s) + x^T (A - lambda I) x - 2 x^T b.
 * CSR arrays are local to the shard (row_ptr[0] == 0); X row written = row_offset + local row.
 * `order` lists local rows longest-first (any permutation is valid; the rows at its head with >= 32768
 * signals are cut into segments that all CTAs build ahead of the solve - a device-planned work list, so a
 * blockbuster row is not the tail of the half-step).  `nnz` = row_ptr[nrows] (the host's copy; -1 if
 * unknown; a hint only).  Results do not depend on the kernel choice below.  `scratch` is 2 ints
 * (scheduler counter, error flag; the call resets both).  loss_sum (device, 1 double) receives
 * the deterministic sum of row_loss.  Asynchronous on `stream`. */
int qmfb_wals_solve_dev(void* stream, double* X, int64_t ldx, int64_t row_offset, const double* Y, int64_t ldy, int k,
                        const int64_t* row_ptr, const int32_t* col, const double* val, const int32_t* order,
                        int64_t nrows, int64_t nnz, const double* gram_packed, double alpha, double lambda, double* row_loss,
                        double* loss_sum, int32_t* scratch);
/* Process-wide choice of the row-solve kernel for k <= 128: 0 = default (the plain kernel: CTA per row, two to
 * twelve resident CTAs per SM by k), 1 = the plain kernel, 2 = the warp-specialised kernel (one CTA per SM: builder
 * warps + two solver groups; measured no faster).  Same results either way; for tests and
 * measurements (environment: QMFB_SOLVE=classic|ws). */
int qmfb_wals_set_solve_kernel(int mode);
/* Same, with the all-gather of the solved shard FUSED into the kernel: every solved row is also
 * stored into row (row_offset + local row) of the `npeers` replicas peer_X[0..npeers) (device
 * pointers into the other ranks' factor matrices, same ldx; peer memory over NVLink, mapped with
 * qmfb_ipc_open).  The caller orders the next reader after all ranks' kernels (any collective that
 * follows the kernel on every rank's stream does, e.g. the allreduce of the loss). */
int qmfb_wals_solve_peers_dev(void* stream, double* X, int64_t ldx, int64_t row_offset, const double* Y, int64_t ldy, int k,
                              const int64_t* row_ptr, const int32_t* col_idx, const double* val, const int32_t* order,
                              int64_t nrows, int64_t nnz, const double* gram_packed, double alpha, double lambda, double* row_loss,
                              double* loss_sum, int32_t* scratch, double* const* peer_X, int npeers);
/* Device buffers shareable between the ranks of one box (one process per GPU): allocate (zeroed) +
 * export a 64-byte CUDA IPC handle; map another rank's buffer; unmap; free. */
int qmfb_ipc_alloc(int device, int64_t bytes, void** ptr, void* handle64);
int qmfb_ipc_open(int device, const void* handle64, void** ptr);
int qmfb_ipc_close(void* ptr);
int qmfb_ipc_free(void* ptr);

/* ---------------------------------------------------------------- WALS engine (host) ------- */
typedef struct qmfb_wals qmfb_wals_t;
/* One handle per process/GPU.  nusers x nitems problem with `nfactors` factors; both factor
 * matrices are resident (zero-initialised) on `device`. */
int qmfb_wals_create(int device, int64_t nusers, int64_t nitems, int nfactors, qmfb_wals_t** out);
int qmfb_wals_destroy(qmfb_wals_t* h);
/* Upload the CSR of one orientation (side USER: user rows over item idx; side ITEM: item rows
 * over user idx), as built by WALSEngine::groupSignals (qmf/wals/WALSEngine.cpp:130-154): rows in
 * ascending raw-id order, entries in ascending raw-id order, duplicates kept.  The shard
 * [row_begin, row_begin + nrows) is what this handle solves; row_ptr has nrows+1 entries
 * starting at 0. */
int qmfb_wals_set_csr(qmfb_wals_t* h, int side, int64_t row_begin, int64_t nrows, const int64_t* row_ptr,
                      const int32_t* col_idx, const double* val);
/* host (n x nfactors row-major, the layout of qmf::Matrix, qmf/Matrix.h:79-81) <-> device */
int qmfb_wals_set_factors(qmfb_wals_t* h, int side, const double* host);
int qmfb_wals_get_factors(qmfb_wals_t* h, int side, double* host);
/* Gram of one side's factors, dense nfactors x nfactors row-major on the host */
int qmfb_wals_gram(qmfb_wals_t* h, int side, double* host_out);
/* One half-step over this handle's shard of `update_side` (WALSEngine::iterate,
 * qmf/wals/WALSEngine.cpp:165-218): zero the side, Gram of the other side, solve every row.
 * *loss_sum receives sum of row losses (NOT yet divided by nusers*nitems). */
int qmfb_wals_half_step(qmfb_wals_t* h, int update_side, double alpha, double lambda, double* loss_sum);
/* One epoch through HOST buffers (WALSEngine::optimize loop body, WALSEngine.cpp:86-92):
 * uploads item factors, runs the user then the item half-step, downloads both factor matrices
 * and returns the item-step loss divided by nusers*nitems, exactly what the reference logs. */
int qmfb_wals_epoch_host(qmfb_wals_t* h, double alpha, double lambda, const double* item_factors_in,
                         double* user_factors_out, double* item_factors_out, double* loss_out);
/* device-side views for multi-GPU drivers (row stride = qmfb_padded_k(nfactors)) */
double* qmfb_wals_factors_device(qmfb_wals_t* h, int side);
void* qmfb_wals_stream(qmfb_wals_t* h);
/* number of kernels launched by this handle so far */
int64_t qmfb_wals_launch_count(qmfb_wals_t* h);
/* device milliseconds (CUDA events on the handle's stream) spent in the last half-step's Gram
 * and row-solve kernels */
int qmfb_wals_last_timing(qmfb_wals_t* h, float* gram_ms, float* solve_ms);

/* ---------------------------------------------------------------- WALS across the GPUs of one box --
 * One process, one handle, `ndev` GPUs (dev_ids; a device may be listed more than once - its shards then
 * share that GPU).  Users and items are each cut into ndev contiguous row ranges balanced by nnz; every
 * device keeps full replicas of both factor matrices.  A half-step is: partial Gram parts on every
 * device -> each device sums all parts over NVLink peer memory -> solve of the device's rows with the
 * all-gather fused into the solve kernel as peer stores -> row losses summed on the first device.  No
 * library collective; results (factors AND loss) are bit-identical for every ndev, including
 * qmfb_wals_* on one GPU.  This is what `wals --ngpus N` binds (WALSEngine::iterate,
 * qmf/wals/WALSEngine.cpp:165-218; the reference's own multi-worker mode is distributed/, out of scope). */
typedef struct qmfb_wals_sharded qmfb_wals_sharded_t;
typedef struct qmfb_signals qmfb_signals_t;   /* GPU-built ingest handle, see "dataset ingest" below */
int qmfb_wals_sharded_create(int ndev, const int* dev_ids, int64_t nusers, int64_t nitems, int nfactors,
                             qmfb_wals_sharded_t** out);
int qmfb_wals_sharded_destroy(qmfb_wals_sharded_t* h);
int qmfb_wals_sharded_ndev(const qmfb_wals_sharded_t* h);
/* FULL CSR of one orientation on the host (row_ptr has n[side]+1 entries); the library cuts and uploads */
int qmfb_wals_sharded_set_csr(qmfb_wals_sharded_t* h, int side, const int64_t* row_ptr, const int32_t* col_idx, const double* val);
/* both orientations from a GPU-built ingest handle, device to device (peer copies of each shard) */
int qmfb_wals_sharded_set_signals(qmfb_wals_sharded_t* h, const qmfb_signals_t* s);
/* device / row range / nnz of shard `slot` for one side (after set_csr / set_signals) */
int qmfb_wals_sharded_shard(const qmfb_wals_sharded_t* h, int slot, int side, int* device, int64_t* row_begin, int64_t* nrows,
                            int64_t* nnz);
/* host n x nfactors row-major -> every replica;  replica of shard `slot` -> host */
int qmfb_wals_sharded_set_factors(qmfb_wals_sharded_t* h, int side, const double* host);
int qmfb_wals_sharded_get_factors(qmfb_wals_sharded_t* h, int side, int slot, double* host);
/* as qmfb_wals_half_step / qmfb_wals_epoch_host; in epoch_host every device moves the rows it solved
 * over its own PCIe link (pinned host buffers keep the copies asynchronous) */
int qmfb_wals_sharded_half_step(qmfb_wals_sharded_t* h, int update_side, double alpha, double lambda, double* loss_sum);
int qmfb_wals_sharded_epoch_host(qmfb_wals_sharded_t* h, double alpha, double lambda, const double* item_factors_in,
                                 double* user_factors_out, double* item_factors_out, double* loss_out);
double* qmfb_wals_sharded_factors_device(qmfb_wals_sharded_t* h, int side, int slot);
int64_t qmfb_wals_sharded_launch_count(qmfb_wals_sharded_t* h);
/* device ms of the last half-step's Gram (parts + reduce) and solve on the first device */
int qmfb_wals_sharded_last_timing(qmfb_wals_sharded_t* h, float* gram_ms, float* solve_ms);

/* ---------------------------------------------------------------- BPR engine (host) -------- */
typedef struct qmfb_bpr qmfb_bpr_t;
/* Factors (and optional item biases, BPRConfig::useBiases) resident on `device`, zero-initialised. */
int qmfb_bpr_create(int device, int64_t nusers, int64_t nitems, int nfactors, int use_biases, qmfb_bpr_t** out);
int qmfb_bpr_destroy(qmfb_bpr_t* h);
/* The positive pairs in data_ order (first-appearance dense idx, qmf/bpr/BPREngine.cpp:65-77).
 * The per-user sorted positive sets used to reject sampled negatives (itemMap_, :79-82) are
 * derived here. */
int qmfb_bpr_set_data(qmfb_bpr_t* h, const int32_t* user_idx, const int32_t* item_idx, int64_t npairs);
/* host row-major n x nfactors (qmf::Matrix layout) <-> device; side = QMFB_SIDE_USER / ITEM */
int qmfb_bpr_set_factors(qmfb_bpr_t* h, int side, const double* host);
int qmfb_bpr_get_factors(qmfb_bpr_t* h, int side, double* host);
int qmfb_bpr_set_biases(qmfb_bpr_t* h, const double* host);
int qmfb_bpr_get_biases(qmfb_bpr_t* h, double* host);
/* One Hogwild SGD pass over all pairs, num_neg sampled negatives each (the SGD half of one
 * iteration of BPREngine::optimize, qmf/bpr/BPREngine.cpp:151-164 -> iterate/iterateBlock ->
 * sampleRandomNegative -> update :178-220).  Negatives come from Philox4x32-10 keyed by
 * (seed, epoch); `shuffle` != 0 visits the pairs in a fresh uniformly random order (the distribution of
 * BPREngine::shuffle's std::shuffle, :276-278; the reference shuffles AFTER an epoch, so its first epoch runs
 * in file order: pass shuffle = 0 there), shuffle == 0 in data_ order.  Only the first
 * hogwild_blocks * floor(npairs / hogwild_blocks) positions are visited (qmfb_bpr_set_hogwild_blocks).
 * *n_updates receives (pairs visited) * num_neg.  Returns QMFB_ERR_NOT_FINITE if a gradient was not finite
 * (CHECK at :184-185). */
int qmfb_bpr_epoch(qmfb_bpr_t* h, double lr, double user_lambda, double item_lambda, double bias_lambda, int num_neg,
                   uint64_t seed, uint64_t epoch, int shuffle, int64_t* n_updates);
/* Apply explicit triplets one after the other in the given order (BPREngine::update, :178-220);
 * the sequential semantics of num_hogwild_threads <= 1, used for exact parity checks. */
int qmfb_bpr_update_triplets(qmfb_bpr_t* h, const int32_t* u, const int32_t* i, const int32_t* j, int64_t n, double lr,
                             double user_lambda, double item_lambda, double bias_lambda);
/* sum over the first n triplets of log(1 + exp(-x)) (loss half of BPREngine::evaluate,
 * :246-261; the caller applies the reference's tail-drop by passing
 * n = nthreads * floor(size / nthreads) and divides by size). */
int qmfb_bpr_eval_loss(qmfb_bpr_t* h, const int32_t* u, const int32_t* i, const int32_t* j, int64_t n, double* sum_out);
/* Upper bound on the number of pairs processed concurrently (the GPU analogue of
 * --num_hogwild_threads, qmf/bpr.cpp:39): 0 = automatic, min(nusers, nitems) / 2 capped by what
 * the device can hold; more pairs in flight means staler gradients per row. */
int qmfb_bpr_set_concurrency(qmfb_bpr_t* h, int64_t max_pairs_in_flight);
/* --num_hogwild_threads (qmf/bpr.cpp:39): the reference cuts data_ into that many blocks of
 * floor(ndata / numHogwildThreads) pairs and never visits the tail (BPREngine.cpp:156-160); the same
 * tail is dropped here.  The parallelism itself is qmfb_bpr_set_concurrency's. */
int qmfb_bpr_set_hogwild_blocks(qmfb_bpr_t* h, int64_t num_hogwild_threads);
/* device time (ms, CUDA events on the handle's stream) of the last qmfb_bpr_epoch kernel */
int qmfb_bpr_last_epoch_ms(qmfb_bpr_t* h, float* ms);
double* qmfb_bpr_factors_device(qmfb_bpr_t* h, int side);
double* qmfb_bpr_biases_device(qmfb_bpr_t* h);
int64_t qmfb_bpr_launch_count(qmfb_bpr_t* h);

/* ---------------------------------------------------------------- ranking evaluation ------- */
/* All-item scoring of nT test users fused with the rank statistics of AUC / AP / P@k / R@k
 * (replaces Engine::computeTestScores, qmf/Engine.cpp:73-96, and the sorts in
 * qmf/metrics/Metrics.cpp:65-164).  Host pointers.  U is nusers x k, V is nitems x k row-major,
 * biases may be NULL.  label_ptr (nT+1) / label_items: per test user the ascending, distinct item
 * idx whose test label is > 0 (Engine::initAvgTestData, Engine.cpp:58-69).
 * Outputs, per test user t with nP = label_ptr[t+1]-label_ptr[t] positives:
 *   cnt[label_ptr[t] + t + i], i = 0..nP : number of NEGATIVE items x such that exactly i
 *       positives score strictly less than x (scores are bit-identical to the reference's);
 *   pos_scores[label_ptr[t] + i]        : the positives' scores in ascending order.
 * A negative in bucket i is preceded, in the reference's order (score descending, positives
 * first on ties), by exactly nP - i positives; every metric follows from that. */
int qmfb_eval_rank(int device, const double* U, int64_t nusers, const double* V, int64_t nitems, int k,
                   const double* biases, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                   const int32_t* label_items, int32_t* cnt, double* pos_scores);
/* The same on device pointers, asynchronous on `stream` (the kernel level): rows of U and V must be 16-byte
 * aligned with an even stride >= qmfb_padded_k(k) and ZERO pad columns (the layout of the WALS engine's
 * factor matrices); max_positives >= the largest label_ptr[t+1]-label_ptr[t] (pass nlabels if unknown);
 * cnt holds nlabels + nT ints, pos_scores nlabels doubles.  Three steps: the positives' scores in the
 * reference's exact order, sorted; all items by DMMA (a users x items GEMM tile by tile, fused with the
 * bucket counting); only pairs whose DMMA score lies within a proven rounding bound of one of the user's
 * positive scores are re-scored in the exact order - the counts are the reference's, bit for bit. */
int qmfb_eval_rank_dev(void* stream, const double* U, int64_t ldu, const double* V, int64_t ldv, int64_t nitems, int k,
                       const double* biases, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                       const int32_t* label_items, int64_t nlabels, int64_t max_positives, int32_t* cnt, double* pos_scores);
/* The engines' evaluation on their RESIDENT factors (no host round trip of the factor matrices): host
 * test-user arrays in, host counters out, as qmfb_eval_rank.  The sharded engine cuts the test users into
 * one contiguous slice per GPU (test users are independent, every GPU holds full replicas); the counts
 * do not depend on the cut. */
int qmfb_wals_eval_rank(qmfb_wals_t* h, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                        const int32_t* label_items, int32_t* cnt, double* pos_scores);
int qmfb_wals_sharded_eval_rank(qmfb_wals_sharded_t* h, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                                const int32_t* label_items, int32_t* cnt, double* pos_scores);
int qmfb_bpr_eval_rank(qmfb_bpr_t* h, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                       const int32_t* label_items, int32_t* cnt, double* pos_scores);

/* HOST side of the evaluation (no device work): the reference's metric arithmetic on the bucket counts.
 * qmfb_repeated_add: s <- fl(s + t), `count` times, with the exact result of the one-by-one loop (the
 * reference adds one equal term per negative when it accumulates AUC, qmf/metrics/Metrics.cpp:87-95) in
 * O(binades crossed) steps.  qmfb_rank_metrics: per_user[t] = "auc" | "ap" | "p@K" | "r@K" of test user t
 * from cnt (layout of qmfb_eval_rank), on `host_threads` threads (0 = all); QMFB_ERR_INVALID where the
 * reference CHECK-fails (ap / r@K without a positive, K > nitems). */
double qmfb_repeated_add(double s, double t, int64_t count);
int qmfb_rank_metrics(const char* metric, const int32_t* cnt, const int64_t* label_ptr, int64_t nT, int64_t nitems, int host_threads,
                      double* per_user);

/* ---- dataset ingest on the GPU ------------------------------------------
 * Replaces IdIndex (qmf/utils/IdIndex.h) + WALSEngine::groupSignals / sortDataset
 * (qmf/wals/WALSEngine.cpp:130-163): raw (user id, item id, value) cells in file order ->
 * dense indices (idx = rank of the id among the distinct ids, the reference's assignment) and both
 * CSR orientations (rows by row id, cells by column id, duplicates kept in file order) plus the
 * longest-first row order the solve kernel deals rows in.  Inputs are host arrays. */
int qmfb_signals_build(int device, int64_t nnz, const int64_t* user_ids, const int64_t* item_ids, const double* values,
                       qmfb_signals_t** out);
int qmfb_signals_destroy(qmfb_signals_t* s);
int qmfb_signals_device_ordinal(const qmfb_signals_t* s);   /* the CUDA device the handle lives on */
int qmfb_signals_dims(const qmfb_signals_t* s, int64_t* nusers, int64_t* nitems, int64_t* nnz);
/* ids_host[idx] = raw id of dense index idx (n[side] entries, ascending) */
int qmfb_signals_ids(const qmfb_signals_t* s, int side, int64_t* ids_host);
/* host copies of one orientation (side = QMFB_SIDE_USER: rows are users); any pointer may be NULL */
int qmfb_signals_csr(const qmfb_signals_t* s, int side, int64_t* row_ptr, int32_t* col_idx, double* val, int32_t* order);
/* device pointers of one orientation (owned by the handle) */
int qmfb_signals_device(const qmfb_signals_t* s, int side, const int64_t** row_ptr, const int32_t** col_idx, const double** val,
                        const int32_t** order);
/* give a WALS engine (created with the handle's nusers / nitems) both orientations, device to device */
int qmfb_wals_set_signals(qmfb_wals_t* h, const qmfb_signals_t* s);

#ifdef __cplusplus
}
#endif
#endif /* QMF_B200_H */
