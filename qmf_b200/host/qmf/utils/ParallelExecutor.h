// Parallel primitives with the reference's surface and - where they define results - its exact
// work split and reduction order (qmf/utils/ParallelExecutor.h:27-62, ParallelExecutor-inl.h:21-85):
//   execute(ntasks, f)               thread t runs tasks t, t + T, t + 2T, ...
//   mapReduce(ntasks, map, red, e)   per-thread partials over the same strided tasks, folded in thread order
//   mapReduce(elems, map, red, e)    thread t folds the block [t B, min((t+1) B, n)) with B = floor(n / T):
//                                    the last n mod T elements are NOT visited (the reference's tail-drop,
//                                    which BPREngine's evaluation loss inherits, BPREngine.cpp:246-263)
#pragma once
#include <algorithm>
#include <future>
#include <memory>
#include <vector>

#include <qmf/utils/ThreadPool.h>

namespace qmf {

class ParallelExecutor {
 public:
  explicit ParallelExecutor(const size_t nthreads) : pool_(std::make_unique<ThreadPool>(nthreads)) {}

  size_t nthreads() const { return pool_->nthreads(); }

  template <typename FuncT>
  void execute(const size_t ntasks, FuncT&& func) {
    const size_t T = nthreads();
    std::vector<std::future<void>> done;
    for (size_t t = 0; t < T; ++t) {
      done.emplace_back(pool_->addTask([t, T, ntasks, func]() {
        for (size_t task = t; task < ntasks; task += T) func(task);
      }));
    }
    for (auto& d : done) d.get();
  }

  template <typename T, typename MapperT, typename ReducerT>
  T mapReduce(const size_t ntasks, MapperT&& mapper, ReducerT&& reducer, T neutral) {
    const size_t nt = nthreads();
    std::vector<std::future<T>> parts;
    for (size_t t = 0; t < nt; ++t) {
      parts.emplace_back(pool_->addTask([t, nt, ntasks, mapper, reducer, neutral]() {
        T acc = neutral;
        for (size_t task = t; task < ntasks; task += nt) acc = reducer(acc, mapper(task));
        return acc;
      }));
    }
    T total = neutral;
    for (auto& p : parts) total = reducer(total, p.get());
    return total;
  }

  template <typename T, typename ElemT, typename MapperT, typename ReducerT>
  T mapReduce(const std::vector<ElemT>& elems, MapperT&& mapper, ReducerT&& reducer, T neutral) {
    const size_t nt = nthreads(), n = elems.size(), block = n / nt;
    std::vector<std::future<T>> parts;
    for (size_t t = 0; t < nt; ++t) {
      parts.emplace_back(pool_->addTask([&elems, t, n, block, mapper, reducer, neutral]() {
        T acc = neutral;
        const size_t end = std::min((t + 1) * block, n);
        for (size_t p = t * block; p < end; ++p) acc = reducer(acc, mapper(elems[p]));
        return acc;
      }));
    }
    T total = neutral;
    for (auto& p : parts) total = reducer(total, p.get());
    return total;
  }

 private:
  std::unique_ptr<ThreadPool> pool_;
};

}  // namespace qmf
