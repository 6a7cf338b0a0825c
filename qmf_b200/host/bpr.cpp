// `bpr` binary: same flags, defaults, log lines and file formats as the reference's
// qmf/bpr.cpp:28-125; SGD and evaluation run on the GPU.  Additive flags: --seed, --device.
#include <memory>

#include <qmf/DatasetReader.h>
#include <qmf/bpr/BPREngine.h>
#include <qmf/metrics/Metrics.h>
#include <qmf/utils/Flags.h>
#include <qmf/utils/Util.h>

DEFINE_uint64(nepochs, 10, "number of epochs for SGD");
DEFINE_uint64(nfactors, 30, "dimension of learned factors");
DEFINE_double(init_learning_rate, 0.05, "initial learning rate");
DEFINE_double(bias_lambda, 1.0, "regularization on biases");
DEFINE_double(user_lambda, 0.025, "regularization on user factors");
DEFINE_double(item_lambda, 0.0025, "regularization on item factors");
DEFINE_double(decay_rate, 0.9, "decay rate on learning rate");
DEFINE_bool(use_biases, false, "use bias term");
DEFINE_double(init_distribution_bound, 0.01, "init distirbution bound");
DEFINE_uint64(num_negative_samples, 3, "number of negative items to sample for each positive item");
DEFINE_uint64(num_hogwild_threads, 1, "number of parallel threads for hogwild");
DEFINE_bool(shuffle_training_set, true, "shuffle training set after each epoch");
DEFINE_uint64(eval_num_neg, 3, "number of negatives generated per positive in evaluation");
DEFINE_int32(eval_seed, 42, "random seed for generating evaluation set and test users");
DEFINE_uint64(nthreads, 16, "number of threads for parallel execution");
DEFINE_string(train_dataset, "", "filename of training dataset");
DEFINE_string(test_dataset, "", "filename of test dataset");
DEFINE_string(test_avg_metrics, "", "comma-separated list of test metrics (averaged per-user)");
DEFINE_uint64(num_test_users, 0, "# users to use for computing test avg metrics (0 = all users)");
DEFINE_bool(test_always, false, "whether to compute test avg metrics after each epoch (if false, only computes at the end)");
DEFINE_string(user_factors, "", "filename of user factors");
DEFINE_string(item_factors, "", "filename of item factors");
DEFINE_int64(seed, -1, "seed of the initial factors and of the device sampler (-1: random_device)");
DEFINE_int32(device, 0, "CUDA device ordinal");

int main(int argc, char** argv) {
  qmf::flags::parse(argc, argv);
  if (FLAGS_user_factors.empty() || FLAGS_item_factors.empty()) {
    LOG(WARNING) << "warning: missing model output filenames! (use options --{user,item}_factors)";
  }
  qmf::BPRConfig config{FLAGS_nepochs, FLAGS_nfactors, FLAGS_init_learning_rate, FLAGS_bias_lambda, FLAGS_user_lambda,
                        FLAGS_item_lambda, FLAGS_decay_rate, FLAGS_use_biases, FLAGS_init_distribution_bound,
                        FLAGS_num_negative_samples, FLAGS_num_hogwild_threads, FLAGS_shuffle_training_set, FLAGS_seed,
                        FLAGS_device};
  const auto metricsEngine = std::make_unique<qmf::MetricsEngine>(
    qmf::MetricsConfig{FLAGS_num_test_users, FLAGS_test_always, FLAGS_eval_seed});
  for (const auto& metric : qmf::split(FLAGS_test_avg_metrics, ',')) {
    CHECK(metricsEngine->addTestAvgMetric(metric)) << "metric " << metric << " is not available";
  }
  qmf::BPREngine engine(config, metricsEngine, FLAGS_eval_num_neg, FLAGS_eval_seed, FLAGS_nthreads);

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
