// TEST INFRASTRUCTURE ONLY — the handful of Boost.Thread / Boost.Bind names util/IndexThreadReduce.h uses, mapped onto <thread>/<functional>.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
namespace boost {
using std::thread;
using std::mutex;
using std::condition_variable;
template <class M> using unique_lock = std::unique_lock<M>;
template <class M> using lock_guard = std::lock_guard<M>;
template <class S> using function = std::function<S>;
using std::bind;
namespace this_thread { using namespace std::this_thread; }
namespace placeholders { using namespace std::placeholders; }
}  // namespace boost
