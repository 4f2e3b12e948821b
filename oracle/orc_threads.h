// TEST INFRASTRUCTURE ONLY — restatement of the reference's CPU "scheduler"
// util/IndexThreadReduce.h:L39-217: NUM_THREADS persistent workers, one mutex + two condition
// variables, dynamic chunk hand-out (nextIndex += stepSize under the lock), stepSize==0 => static
// ceil(n/NUM_THREADS) split, every worker is called at least once per reduce (with an empty range)
// and a per-reduce `stats` sum.  Used only for the CPU-baseline timing of the oracle.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace orc {

class ThreadPool {
 public:
  typedef std::function<void(int, int, double*, int)> Fn;  // (min, max, stats[10], tid)
  explicit ThreadPool(int n) : n_(n), isDone_(n, false), gotOne_(n, true) {
    for (int i = 0; i < n_; i++) workers_.emplace_back(&ThreadPool::workerLoop, this, i);
  }
  ~ThreadPool() {
    {
      std::unique_lock<std::mutex> lock(m_);
      running_ = false;
      todo_.notify_all();
    }
    for (auto& t : workers_) t.join();
  }
  int size() const { return n_; }
  double stats[10];

  void reduce(Fn fn, int first, int end, int stepSize = 0) {
    for (int i = 0; i < 10; i++) stats[i] = 0;
    if (stepSize == 0) stepSize = ((end - first) + n_ - 1) / n_;
    std::unique_lock<std::mutex> lock(m_);
    fn_ = fn;
    nextIndex_ = first;
    maxIndex_ = end;
    stepSize_ = stepSize;
    for (int i = 0; i < n_; i++) { isDone_[i] = false; gotOne_[i] = false; }
    todo_.notify_all();
    while (true) {
      done_.wait(lock);
      bool allDone = true;
      for (int i = 0; i < n_; i++) allDone = allDone && isDone_[i];
      if (allDone) break;
    }
    nextIndex_ = 0;
    maxIndex_ = 0;
  }

 private:
  void workerLoop(int idx) {
    std::unique_lock<std::mutex> lock(m_);
    while (running_) {
      int todo = 0;
      bool got = false;
      if (nextIndex_ < maxIndex_) { todo = nextIndex_; nextIndex_ += stepSize_; got = true; }
      if (got) {
        int hi = std::min(todo + stepSize_, maxIndex_);
        lock.unlock();
        double s[10] = {0};
        fn_(todo, hi, s, idx);
        gotOne_[idx] = true;
        lock.lock();
        for (int i = 0; i < 10; i++) stats[i] += s[i];
      } else {
        if (!gotOne_[idx]) {
          lock.unlock();
          double s[10] = {0};
          fn_(0, 0, s, idx);
          gotOne_[idx] = true;
          lock.lock();
          for (int i = 0; i < 10; i++) stats[i] += s[i];
        }
        isDone_[idx] = true;
        done_.notify_all();
        todo_.wait(lock);
      }
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::vector<bool> isDone_, gotOne_;
  std::mutex m_;
  std::condition_variable todo_, done_;
  int nextIndex_ = 0, maxIndex_ = 0, stepSize_ = 1;
  bool running_ = true;
  Fn fn_;
};

}  // namespace orc
