// Fixed-size pool of host threads with the reference's surface (qmf/utils/ThreadPool.h:33-68:
// ThreadPool(nthreads), nthreads(), addTask(f, args...) -> std::future).  The training hot path does
// not use it (kernel launches replace the std::thread loops); it serves host-side callers and the
// summation orders of ParallelExecutor that define results (tail-drop, strided partial sums).
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace qmf {

class ThreadPool {
 public:
  explicit ThreadPool(const size_t nthreads) {
    workers_.reserve(nthreads);
    for (size_t t = 0; t < nthreads; ++t) workers_.emplace_back([this] { work(); });
  }
  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> lock(mu_);
      stopping_ = true;
    }
    wake_.notify_all();
    for (auto& w : workers_) w.join();
  }
  ThreadPool(const ThreadPool&) = delete;
  ThreadPool& operator=(const ThreadPool&) = delete;

  size_t nthreads() const { return workers_.size(); }

  template <typename FuncT, typename... Args>
  auto addTask(FuncT&& func, Args&&... args) -> std::future<std::invoke_result_t<FuncT, Args...>> {
    using R = std::invoke_result_t<FuncT, Args...>;
    auto job = std::make_shared<std::packaged_task<R()>>(std::bind(std::forward<FuncT>(func), std::forward<Args>(args)...));
    std::future<R> result = job->get_future();
    {
      std::lock_guard<std::mutex> lock(mu_);
      queue_.emplace_back([job] { (*job)(); });
    }
    wake_.notify_one();
    return result;
  }

 private:
  void work() {
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> lock(mu_);
        wake_.wait(lock, [this] { return stopping_ || !queue_.empty(); });
        if (queue_.empty()) return;  // stopping and drained
        job = std::move(queue_.front());
        queue_.pop_front();
      }
      job();
    }
  }

  std::vector<std::thread> workers_;
  std::deque<std::function<void()>> queue_;
  std::mutex mu_;
  std::condition_variable wake_;
  bool stopping_ = false;
};

}  // namespace qmf
