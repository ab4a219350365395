// Workers.h — a few persistent host threads for the parallel passes of the host mirror (BundleAdjusterCuda::Marshal).
// Run(n, f) executes f on n - 1 pooled threads and on the caller, and returns when all are done.  The threads are created
// on first use and live for the process; between two passes of one call they spin briefly, otherwise they sleep on a
// condition variable (creating 8 std::threads per pass cost more than the passes: ~0.2 ms each on the build box).
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mcp_host {

class Workers {
 public:
  static Workers& Get() { static Workers w; return w; }
  void Run(int n, const std::function<void()>& f)
  {
    if (n <= 1) { f(); return; }
    std::lock_guard<std::mutex> serial(mRunMutex);            // one Run at a time (the adapters are per-thread objects, the pool is shared)
    // (a thread born here starts from the current generation, so it takes part in this very Run)
    const unsigned long long g = mnGen.load(std::memory_order_acquire);
    while ((int)mvThreads.size() < n - 1) { const int id = (int)mvThreads.size(); mvThreads.emplace_back([this, id, g] { Loop(id, g); }); }
    mpJob = &f;
    mnWanted = n - 1;
    mnRemaining.store(n - 1, std::memory_order_relaxed);
    { std::lock_guard<std::mutex> lk(mMutex); mnGen.fetch_add(1, std::memory_order_release); }
    mCv.notify_all();
    f();
    while (mnRemaining.load(std::memory_order_acquire) > 0) std::this_thread::yield();
  }
  ~Workers()
  {
    { std::lock_guard<std::mutex> lk(mMutex); mbStop = true; mnGen.fetch_add(1, std::memory_order_release); }
    mCv.notify_all();
    for (std::thread& t : mvThreads) t.join();
  }

 private:
  void Loop(int id, unsigned long long last)
  {
    for (;;) {
      bool got = false;
      const auto t0 = std::chrono::steady_clock::now();
      for (int spin = 0;; spin++) {
        if (mnGen.load(std::memory_order_acquire) != last) { got = true; break; }
        if ((spin & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) break;
      }
      if (!got) {
        std::unique_lock<std::mutex> lk(mMutex);
        mCv.wait(lk, [&] { return mnGen.load(std::memory_order_acquire) != last; });
      }
      if (mbStop) return;
      last = mnGen.load(std::memory_order_acquire);
      if (id < mnWanted) {
        (*mpJob)();
        mnRemaining.fetch_sub(1, std::memory_order_release);
      }
    }
  }
  std::vector<std::thread> mvThreads;
  std::mutex mMutex, mRunMutex;
  std::condition_variable mCv;
  std::atomic<unsigned long long> mnGen{ 0 };
  std::atomic<int> mnRemaining{ 0 };
  const std::function<void()>* mpJob = nullptr;
  int mnWanted = 0;
  bool mbStop = false;
};

}  // namespace mcp_host
