// Host -> device upload of PAGEABLE arrays: the reference API speaks float64 NumPy (SURVEY 8b "inputs/outputs at this level
// are host NumPy float64 arrays"), and a pageable cudaMemcpy is staged by the driver on one thread (~10 GB/s; the 88 MB
// of a cfg3 batch in float64 took 8 ms, ten times the device step).  Here a small pool of host threads narrows (float64 ->
// float32, round to nearest even exactly like the device cast) or copies the source into two pinned staging buffers,
// chunk by chunk, while the previous chunk is on its way over the host link at full rate.  The source has been read
// completely when the call returns (the caller may reuse it); the copies themselves are ordered on `stream`.
#include "common.cuh"
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <string.h>

namespace bf {

namespace {

constexpr size_t UP_CHUNK = 4u << 20;      // elements per staging buffer (16 MB of float32)

class Pool {
 public:
  // never destroyed: at process exit the workers still wait on the condition variable, and destroying that would block
  static Pool& get() { static Pool* p = new Pool; return *p; }
  // runs fn(part, nparts) on every worker and waits
  template <typename F>
  void run(F fn) {
    std::unique_lock<std::mutex> lk(mu_);
    job_ = [&](int i, int n) { fn(i, n); };
    pending_ = (int)workers_.size();
    ++gen_;
    cv_.notify_all();
    done_.wait(lk, [&] { return pending_ == 0; });
  }
  int size() const { return (int)workers_.size(); }

 private:
  Pool() {
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 4;
    if (n > 16) n = 16;
    if (const char* e = getenv("BF_UPLOAD_THREADS")) { int v = atoi(e); if (v >= 1 && v <= 64) n = (unsigned)v; }
    for (unsigned i = 0; i < n; ++i) workers_.emplace_back([this, i, n] { loop((int)i, (int)n); });
    for (auto& t : workers_) t.detach();      // the pool lives as long as the process
  }
  void loop(int idx, int n) {
    unsigned long long seen = 0;
    for (;;) {
      std::function<void(int, int)> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        job = job_;
      }
      job(idx, n);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::function<void(int, int)> job_;
  std::vector<std::thread> workers_;
  unsigned long long gen_ = 0;
  int pending_ = 0;
};

struct Stage {
  float* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
};
Stage g_stage;
std::mutex g_upload_mu;

}  // namespace

}  // namespace bf

using namespace bf;

extern "C" {

int bf_upload_host(float* d_dst, const void* h_src, size_t n, int src_is_f64, void* stream) {
  if (n == 0) return BF_OK;
  BF_REQUIRE(d_dst && h_src, BF_ERR_ARG, "bf_upload_host: null pointer");
  std::lock_guard<std::mutex> guard(g_upload_mu);
  cudaStream_t s = as_stream(stream);
  for (int i = 0; i < 2; ++i) {
    if (!g_stage.buf[i]) {
      BF_CUDA(cudaHostAlloc((void**)&g_stage.buf[i], UP_CHUNK * sizeof(float), cudaHostAllocDefault));
      BF_CUDA(cudaEventCreateWithFlags(&g_stage.ev[i], cudaEventDisableTiming));
    }
  }
  Pool& pool = Pool::get();
  int slot = 0;
  for (size_t off = 0; off < n; off += UP_CHUNK, slot ^= 1) {
    const size_t cnt = n - off < UP_CHUNK ? n - off : UP_CHUNK;
    if (g_stage.used[slot]) BF_CUDA(cudaEventSynchronize(g_stage.ev[slot]));      // its previous copy has left the buffer
    float* dst = g_stage.buf[slot];
    if (src_is_f64) {
      const double* src = (const double*)h_src + off;
      pool.run([&](int part, int nparts) {
        const size_t a = cnt * (size_t)part / nparts, b = cnt * (size_t)(part + 1) / nparts;
        for (size_t i = a; i < b; ++i) dst[i] = (float)src[i];
      });
    } else {
      const float* src = (const float*)h_src + off;
      pool.run([&](int part, int nparts) {
        const size_t a = cnt * (size_t)part / nparts, b = cnt * (size_t)(part + 1) / nparts;
        memcpy(dst + a, src + a, (b - a) * sizeof(float));
      });
    }
    BF_CUDA(cudaMemcpyAsync(d_dst + off, dst, cnt * sizeof(float), cudaMemcpyHostToDevice, s));
    BF_CUDA(cudaEventRecord(g_stage.ev[slot], s));
    g_stage.used[slot] = true;
  }
  return BF_OK;
}

}  // extern "C"
