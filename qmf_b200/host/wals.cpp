// `wals` binary: same flags, defaults, log lines and file formats as the reference's
// qmf/wals.cpp:26-107; the training loop runs on the GPU.  Additive flags: --seed, --device, --ngpus.
#include <memory>

#include <qmf/DatasetReader.h>
#include <qmf/metrics/Metrics.h>
#include <qmf/utils/Flags.h>
#include <qmf/utils/Util.h>
#include <qmf/wals/WALSEngine.h>

DEFINE_uint64(nepochs, 10, "number of epochs for ALS");
DEFINE_uint64(nfactors, 30, "dimension of learned factors");
DEFINE_double(regularization_lambda, 0.05, "regularization param");
DEFINE_double(confidence_weight, 40, "confidence weight");
DEFINE_double(init_distribution_bound, 0.01, "init distirbution bound");
DEFINE_string(distribution_file, "", "uniform distribution file, for repeatable result");
DEFINE_int32(nthreads, 16, "number of threads for parallel execution");
DEFINE_string(train_dataset, "", "filename of training dataset");
DEFINE_string(test_dataset, "", "filename of test dataset");
DEFINE_string(test_avg_metrics, "", "comma-separated list of test metrics (averaged per-user)");
DEFINE_int32(eval_seed, 42, "random seed for picking test users");
DEFINE_uint64(num_test_users, 0, "# users to use for computing test avg metrics (0 = all users)");
DEFINE_bool(test_always, false, "whether to compute test avg metrics after each epoch (if false, only computes at the end)");
DEFINE_string(user_factors, "", "filename of user factors");
DEFINE_string(item_factors, "", "filename of item factors");
DEFINE_int64(seed, -1, "seed of the initial item factors when no distribution file is given (-1: random_device)");
DEFINE_int32(device, 0, "CUDA device ordinal (the first one when --ngpus > 1)");
DEFINE_int32(ngpus, 1, "row-partition every half-step over this many GPUs of the box (devices device .. device+ngpus-1)");

int main(int argc, char** argv) {
  qmf::flags::parse(argc, argv);
  if (FLAGS_user_factors.empty() || FLAGS_item_factors.empty()) {
    LOG(WARNING) << "warning: missing model output filenames! (use options --{user,item}_factors)";
  }
  qmf::WALSConfig config{FLAGS_nepochs, FLAGS_nfactors, FLAGS_regularization_lambda, FLAGS_confidence_weight,
                         FLAGS_init_distribution_bound, FLAGS_distribution_file, FLAGS_seed, FLAGS_device, FLAGS_ngpus};
  CHECK_GE(FLAGS_ngpus, 1) << "--ngpus must be >= 1";
  const auto metricsEngine = std::make_unique<qmf::MetricsEngine>(
    qmf::MetricsConfig{FLAGS_num_test_users, FLAGS_test_always, FLAGS_eval_seed});
  for (const auto& metric : qmf::split(FLAGS_test_avg_metrics, ',')) {
    CHECK(metricsEngine->addTestAvgMetric(metric)) << "metric " << metric << " is not available";
  }
  qmf::WALSEngine engine(config, metricsEngine, size_t(FLAGS_nthreads));

  LOG(INFO) << "loading training data";
  engine.init(qmf::DatasetReader(FLAGS_train_dataset).readAll());
  if (!FLAGS_test_dataset.empty()) {
    LOG(INFO) << "loading test data";
    engine.initTest(qmf::DatasetReader(FLAGS_test_dataset).readAll());
  }
  LOG(INFO) << "training";
  engine.optimize();
  if (!FLAGS_user_factors.empty() && !FLAGS_item_factors.empty()) {
    LOG(INFO) << "saving model output";
    engine.saveUserFactors(FLAGS_user_factors);
    engine.saveItemFactors(FLAGS_item_factors);
  }
  return 0;
}
