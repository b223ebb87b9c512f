// Host side of the predict() pipeline: frames in ordinary (pageable) memory -> the engine's page-locked staging ring.
//
// The reference hands `model.predict` one `cap.read()` ndarray at a time (yolo_seg/app.py:85-91): pageable memory.
// The DMA engine only reads page-locked memory at full PCIe rate, so the frames are copied once, by a persistent pool
// of host threads (no thread creation per call), in 256 KB pieces (balanced whatever the frame count), with
// non-temporal stores where the CPU has AVX2 (the destination is written once and read by the DMA engine next: going
// around the cache saves the read-for-ownership of every destination line).  The Python caller releases the GIL for
// the duration of ypb_stage_frames() (ctypes).
#include <atomic>
#include <cstdlib>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

constexpr size_t kPiece = 256 * 1024;

#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_nt_avx2(uint8_t* d, const uint8_t* s, size_t n) {
  // head: up to the first 32-byte boundary of the destination
  size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
  if (head > n) head = n;
  memcpy(d, s, head);
  d += head; s += head; n -= head;
  size_t blocks = n / 128;
  for (size_t i = 0; i < blocks; ++i) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 64));
    const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 96), e);
    s += 128; d += 128;
  }
  _mm_sfence();
  memcpy(d, s, n - blocks * 128);
}
bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}
#endif

inline void copy_piece(uint8_t* d, const uint8_t* s, size_t n, int mode) {
#if defined(__x86_64__)
  if (mode == 1 && have_avx2()) {
    copy_nt_avx2(d, s, n);
    return;
  }
#endif
  memcpy(d, s, n);
}

struct Job {
  void* const* dst;
  const void* const* src;
  const size_t* bytes;
  int n = 0;
  int mode = 0;
  std::vector<long> first_piece;  // prefix sum of pieces per frame
  long total = 0;
  std::atomic<long> next{0};
  std::atomic<long> done{0};
};

class StagePool {
 public:
  static StagePool& get() {
    static StagePool p;
    return p;
  }
  void run(Job* job, int nthreads) {
    std::unique_lock<std::mutex> run_lock(run_mutex_);  // one job at a time (an engine is single-threaded anyway)
    grow(nthreads - 1);
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = job;
      want_ = nthreads - 1;
      ++epoch_;
      static const int hc = (int)std::thread::hardware_concurrency();
      int lim = nthreads - 1;
      if (lim > hc - 2) lim = hc - 2;
      spin_limit_.store(lim < 0 ? 0 : lim, std::memory_order_relaxed);
      epoch_hint_.store(epoch_, std::memory_order_release);
    }
    cv_.notify_all();
    work(job);
    // wait for the pieces other threads are still copying
    while (job->done.load(std::memory_order_acquire) < job->total) std::this_thread::yield();
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = nullptr;
    }
    // no worker may still hold the job pointer when we return
    while (active_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
  }

 private:
  StagePool() = default;
  ~StagePool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void grow(int n) {
    while ((int)threads_.size() < n) {
      const int id = (int)threads_.size();
      threads_.emplace_back([this, id] { loop(id); });
    }
  }
  static void work(Job* job) {
    for (;;) {
      const long p = job->next.fetch_add(1, std::memory_order_relaxed);
      if (p >= job->total) return;
      // frame of piece p: the few frames of a call make a linear scan cheap
      int f = 0;
      while (f + 1 < job->n && job->first_piece[f + 1] <= p) ++f;
      const size_t off = (size_t)(p - job->first_piece[f]) * kPiece;
      const size_t len = job->bytes[f] - off < kPiece ? job->bytes[f] - off : kPiece;
      copy_piece(static_cast<uint8_t*>(job->dst[f]) + off, static_cast<const uint8_t*>(job->src[f]) + off, len, job->mode);
      job->done.fetch_add(1, std::memory_order_release);
    }
  }
  void loop(int id) {
    unsigned long seen = 0;
    for (;;) {
      Job* job = nullptr;
      // stay hot between the chunk calls of one predict(): spin on the epoch for a short while before sleeping (a thread
      // parked on the condition variable needs 50-100 us to come back and its core has clocked down by then; measured
      // in situ the chunk calls ran at 26 GB/s against 71 GB/s back to back)
      // (only the threads the last job used, and never more spinners than cores minus two: a spinning thread that
      // shares a core with a working one halves it)
      if (id < spin_limit_.load(std::memory_order_relaxed))
        for (int spin = 0; spin < spin_iters() && epoch_hint_.load(std::memory_order_acquire) == seen; ++spin) cpu_relax();
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
        if (id < want_ && job_ != nullptr) {
          job = job_;
          active_.fetch_add(1, std::memory_order_acq_rel);
        }
      }
      if (job) {
        work(job);
        active_.fetch_sub(1, std::memory_order_acq_rel);
      }
    }
  }
  // ~0.4 ms of spinning by default.  YPB_STAGE_SPIN=<iterations> (0: park at once), YPB_STAGE_SPIN_NOPAUSE=1: spin on the
  // load alone - under a hypervisor a long PAUSE loop is what pause-loop exiting deschedules a vCPU for.
  static int spin_iters() {
    static const int n = getenv("YPB_STAGE_SPIN") ? atoi(getenv("YPB_STAGE_SPIN")) : 40000;
    return n;
  }
  static void cpu_relax() {
    static const bool nopause = getenv("YPB_STAGE_SPIN_NOPAUSE") != nullptr;
    if (nopause) return;
#if defined(__x86_64__)
    _mm_pause();
#endif
  }
  std::atomic<unsigned long> epoch_hint_{0};
  std::atomic<int> spin_limit_{0};
  std::mutex m_, run_mutex_;
  std::condition_variable cv_;
  std::vector<std::thread> threads_;
  Job* job_ = nullptr;
  int want_ = 0;
  unsigned long epoch_ = 0;
  bool stop_ = false;
  std::atomic<int> active_{0};
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// Asynchronous flavour: the caller queues staging jobs and goes on enqueuing GPU work; a host function placed in the copy
// stream in front of each chunk's H2D copy (ypb_stage_gate, ypb200.cu) holds that stream until the chunk's ticket is
// complete.  predict() then never blocks on staging: frames are staged by the pool while Python enqueues the engine
// passes, and the copy engine picks every chunk up the moment it is ready.
// ------------------------------------------------------------------------------------------------
namespace {

struct AsyncJob {
  std::vector<void*> dst;
  std::vector<const void*> src;
  std::vector<size_t> bytes;
  std::vector<long> first_piece;
  int mode = 0;
  long total = 0;
  std::atomic<long> next{0};
  std::atomic<long> done{0};
  std::mutex m;
  std::condition_variable cv;
  bool finished = false;
};

class AsyncStagePool {
 public:
  static AsyncStagePool& get() {
    static AsyncStagePool p;
    return p;
  }
  void submit(std::shared_ptr<AsyncJob> job, int nthreads) {
    {
      std::lock_guard<std::mutex> lk(m_);
      while ((int)threads_.size() < nthreads) threads_.emplace_back([this] { loop(); });
      queue_.push_back(std::move(job));
    }
    cv_.notify_all();
  }

 private:
  AsyncStagePool() = default;
  ~AsyncStagePool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void loop() {
    for (;;) {
      std::shared_ptr<AsyncJob> job;  // shared ownership: the waiter may drop its reference while a late worker still polls
      {
        std::unique_lock<std::mutex> lk(m_);
        // the head job stays at the front until its last piece has been HANDED OUT: all threads work on the oldest chunk
        for (;;) {
          while (!queue_.empty() && queue_.front()->next.load(std::memory_order_relaxed) >= queue_.front()->total) queue_.erase(queue_.begin());
          if (stop_ || !queue_.empty()) break;
          cv_.wait(lk);
        }
        if (stop_) return;
        job = queue_.front();
      }
      for (;;) {
        const long p = job->next.fetch_add(1, std::memory_order_relaxed);
        if (p >= job->total) break;
        int f = 0;
        const int n = (int)job->bytes.size();
        while (f + 1 < n && job->first_piece[f + 1] <= p) ++f;
        const size_t off = (size_t)(p - job->first_piece[f]) * kPiece;
        const size_t len = job->bytes[f] - off < kPiece ? job->bytes[f] - off : kPiece;
        copy_piece(static_cast<uint8_t*>(job->dst[f]) + off, static_cast<const uint8_t*>(job->src[f]) + off, len, job->mode);
        if (job->done.fetch_add(1, std::memory_order_acq_rel) + 1 == job->total) {
          {
            std::lock_guard<std::mutex> lk(job->m);
            job->finished = true;
          }
          job->cv.notify_all();
          break;
        }
      }
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::vector<std::thread> threads_;
  std::vector<std::shared_ptr<AsyncJob>> queue_;
  bool stop_ = false;
};

}  // namespace

// Queue a staging job; returns an opaque ticket that MUST be passed to exactly one ypb_host_stage_wait().
extern "C" void* ypb_host_stage_submit(void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads, int mode) {
  auto job = std::make_shared<AsyncJob>();
  job->mode = mode;
  job->dst.assign(dst, dst + n);
  job->src.assign(src, src + n);
  job->bytes.assign(bytes, bytes + n);
  job->first_piece.resize(n + 1);
  long acc = 0;
  for (int i = 0; i < n; ++i) {
    job->first_piece[i] = acc;
    acc += (long)((bytes[i] + kPiece - 1) / kPiece);
  }
  job->first_piece[n] = acc;
  job->total = acc;
  auto* ticket = new std::shared_ptr<AsyncJob>(job);
  if (acc == 0) {
    job->finished = true;
    return ticket;
  }
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  AsyncStagePool::get().submit(job, nthreads);
  return ticket;
}

// Block until the ticket's frames are staged, then free it (called from the CUDA host function, or directly).
extern "C" void ypb_host_stage_wait(void* ticket) {
  auto* holder = static_cast<std::shared_ptr<AsyncJob>*>(ticket);
  if (!holder) return;
  {
    AsyncJob* job = holder->get();
    std::unique_lock<std::mutex> lk(job->m);
    job->cv.wait(lk, [&] { return job->finished; });
  }
  delete holder;
}

// mode 0: memcpy; 1: non-temporal stores (AVX2) when available.  Returns 0.
extern "C" int ypb_host_stage_frames(void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads, int mode) {
  if (n <= 0) return 0;
  Job job;
  job.dst = dst; job.src = src; job.bytes = bytes; job.n = n; job.mode = mode;
  job.first_piece.resize(n + 1);
  long acc = 0;
  for (int i = 0; i < n; ++i) {
    job.first_piece[i] = acc;
    acc += (long)((bytes[i] + kPiece - 1) / kPiece);
  }
  job.first_piece[n] = acc;
  job.total = acc;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  if ((long)nthreads > acc) nthreads = (int)acc;
  if (nthreads <= 1) {
    for (int i = 0; i < n; ++i) copy_piece(static_cast<uint8_t*>(dst[i]), static_cast<const uint8_t*>(src[i]), bytes[i], mode);
    return 0;
  }
  StagePool::get().run(&job, nthreads);
  return 0;
}
