// Host-side narrowing of the counts' values before they cross PCIe.
//
// R hands the counts over as a dgCMatrix whose `x` slot is double (8 bytes per stored count,
// R/helper_functions.R:50-88 sums them as such); single-cell counts are tiny integers.  The streamed
// call (cv_deviations_csc) therefore converts `x` to bytes on a few host threads, slice by slice and
// ahead of the uploads, into page-locked staging memory: 1 byte instead of 8 per nonzero travels
// (config 2: 205 MB instead of 425 MB per call), which matters once several GPUs share the
// host's PCIe uplinks, and for pageable caller memory (R's) the large copy starts from pinned
// memory.  A slice holding anything but an integer in 0..255 (a large count, a fraction, NaN) is
// reported and that chunk travels in its original type; validation stays on the device (K1).
//
// The same worker threads stage PAGEABLE caller memory (what R passes): the row indices of a slice are
// copied into page-locked memory next to the narrowed values, so every upload is a true asynchronous
// DMA from pinned memory instead of the driver's single-threaded bounce copy, and finished slabs of
// the results are copied from a page-locked landing buffer into the caller's matrices in parallel
// pieces (cv_narrower_copy).
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "chromvar_b200.h"

namespace {

const int64_t SLICE = 1 << 18;

template <typename T>
bool narrow_scalar(const T* x, uint8_t* out, int64_t n) {
  bool bad = false;
  for (int64_t i = 0; i < n; ++i) {
    const T v = x[i];
    const bool in = v >= (T)0 && v <= (T)255;          // NaN compares false
    const int t = in ? (int)v : 0;
    bad |= !in || (T)t != v;
    out[i] = (uint8_t)t;
  }
  return !bad;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) bool narrow_f64_avx2(const double* x, uint8_t* out, int64_t n) {
  int neq = 0;
  __m128i hi = _mm_setzero_si128();
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m256d a = _mm256_loadu_pd(x + i), b = _mm256_loadu_pd(x + i + 4);
    const __m128i ia = _mm256_cvttpd_epi32(a), ib = _mm256_cvttpd_epi32(b);
    neq |= _mm256_movemask_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(ia), a, _CMP_NEQ_UQ)) |
           _mm256_movemask_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(ib), b, _CMP_NEQ_UQ));
    hi = _mm_or_si128(hi, _mm_or_si128(ia, ib));       // any bit above the low byte: out of range / negative
    const __m128i w = _mm_packus_epi32(ia, ib);
    _mm_storel_epi64(reinterpret_cast<__m128i*>(out + i), _mm_packus_epi16(w, w));
  }
  bool ok = neq == 0 && _mm_testz_si128(hi, _mm_set1_epi32(~0xFF));
  if (i < n) ok &= narrow_scalar(x + i, out + i, n - i);
  return ok;
}

// 16 values per iteration; VPMOVUSDB packs with unsigned saturation (out-of-range values are flagged through
// `hi` / `neq` as in the AVX2 form, what lands in the bytes then does not matter)
__attribute__((target("avx512f,avx512bw,avx512vl"))) bool narrow_f64_avx512(const double* x, uint8_t* out, int64_t n) {
  unsigned neq = 0;
  __m256i hi = _mm256_setzero_si256();
  int64_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m512d a = _mm512_loadu_pd(x + i), b = _mm512_loadu_pd(x + i + 8);
    const __m256i ia = _mm512_cvttpd_epi32(a), ib = _mm512_cvttpd_epi32(b);
    neq |= (unsigned)_mm512_cmp_pd_mask(_mm512_cvtepi32_pd(ia), a, _CMP_NEQ_UQ) |
           (unsigned)_mm512_cmp_pd_mask(_mm512_cvtepi32_pd(ib), b, _CMP_NEQ_UQ);
    hi = _mm256_or_si256(hi, _mm256_or_si256(ia, ib));
    const __m512i w = _mm512_inserti64x4(_mm512_castsi256_si512(ia), ib, 1);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i), _mm512_cvtusepi32_epi8(w));
  }
  bool ok = neq == 0 && _mm256_testz_si256(hi, _mm256_set1_epi32(~0xFF));
  if (i < n) ok &= narrow_scalar(x + i, out + i, n - i);
  return ok;
}
#endif

bool narrow_any(const void* x, int x_type, int64_t off, uint8_t* out, int64_t n) {
  switch (x_type) {
    case CV_F64:
#if defined(__x86_64__)
      if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl"))
        return narrow_f64_avx512((const double*)x + off, out, n);
      if (__builtin_cpu_supports("avx2")) return narrow_f64_avx2((const double*)x + off, out, n);
#endif
      return narrow_scalar((const double*)x + off, out, n);
    case CV_F32: return narrow_scalar((const float*)x + off, out, n);
    case CV_I32: return narrow_scalar((const int32_t*)x + off, out, n);
    default: return false;
  }
}

}  // namespace

struct cv_narrower {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_idle;
  uint64_t generation = 0;
  bool quit = false;
  int active = 0;
  // current job
  const void* x = nullptr;
  int x_type = 0;
  int64_t n = 0, n_slices = 0;
  uint8_t* out = nullptr;
  const int32_t* ri = nullptr;                 // optional: row indices staged slice by slice next to the values
  int32_t* ri_out = nullptr;
  std::atomic<int64_t> next{0};
  // plain copies (results leaving through page-locked memory): pieces handed out under the mutex
  struct Copy { char* dst; const char* src; size_t n; };
  std::deque<Copy> copies;
  int64_t copies_open = 0;                     // queued + running
  std::vector<std::atomic<uint8_t>> flag;      // 0 pending, 1 all values narrowed, 2 slice holds something else

  void do_slice(int64_t s) {
    const int64_t a = s * SLICE, b = std::min(n, a + SLICE);
    if (ri) memcpy(ri_out + a, ri + a, (size_t)(b - a) * 4);
    const bool ok = narrow_any(x, x_type, a, out + a, b - a);
    flag[(size_t)s].store(ok ? 1 : 2, std::memory_order_release);
  }

  void run() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return quit || generation != seen || !copies.empty(); });
        if (quit) return;
        seen = generation;
        ++active;
      }
      for (;;) {
        Copy c{nullptr, nullptr, 0};
        {
          std::lock_guard<std::mutex> lk(mu);
          if (!copies.empty()) { c = copies.front(); copies.pop_front(); }
        }
        if (c.dst) {                            // results first: they free the landing buffer's consumers
          memcpy(c.dst, c.src, c.n);
          std::lock_guard<std::mutex> lk(mu);
          --copies_open;
          cv_idle.notify_all();
          continue;
        }
        const int64_t s = next.fetch_add(1, std::memory_order_relaxed);
        if (s >= n_slices) break;
        do_slice(s);
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        --active;
      }
      cv_idle.notify_all();
    }
  }
};

cv_narrower* cv_narrower_create(int threads) {
  cv_narrower* nw = new cv_narrower();
  if (threads < 1) threads = 1;
  for (int i = 0; i < threads; ++i) nw->workers.emplace_back([nw] { nw->run(); });
  return nw;
}

int cv_narrower_threads(const cv_narrower* nw) { return nw ? (int)nw->workers.size() : 0; }

// No worker touches the caller's buffers after this returns (pending slices are dropped, queued
// copies are carried out).
void cv_narrower_finish(cv_narrower* nw) {
  if (!nw) return;
  nw->next.store(nw->n_slices, std::memory_order_relaxed);
  std::unique_lock<std::mutex> lk(nw->mu);
  nw->cv_idle.wait(lk, [&] { return nw->active == 0 && nw->copies_open == 0; });
}

// memcpy(dst, src, n) on the worker threads, in pieces; returns at once.  May be called from a CUDA
// host callback (no CUDA call inside).  cv_narrower_copies_wait / cv_narrower_finish wait for them.
void cv_narrower_copy(cv_narrower* nw, void* dst, const void* src, size_t n) {
  const size_t piece = (size_t)4 << 20;
  {
    std::lock_guard<std::mutex> lk(nw->mu);
    for (size_t o = 0; o < n; o += piece) {
      nw->copies.push_back(cv_narrower::Copy{(char*)dst + o, (const char*)src + o, std::min(piece, n - o)});
      ++nw->copies_open;
    }
  }
  nw->cv_job.notify_all();
}

void cv_narrower_copies_wait(cv_narrower* nw) {
  std::unique_lock<std::mutex> lk(nw->mu);
  nw->cv_idle.wait(lk, [&] { return nw->copies_open == 0; });
}

void cv_narrower_destroy(cv_narrower* nw) {
  if (!nw) return;
  cv_narrower_finish(nw);
  {
    std::lock_guard<std::mutex> lk(nw->mu);
    nw->quit = true;
  }
  nw->cv_job.notify_all();
  for (std::thread& t : nw->workers) t.join();
  delete nw;
}

// Start converting x[0, n) (x_type CV_F64 / CV_F32 / CV_I32) into out[0, n); returns at once.
// ri / ri_out (optional): the row indices of every slice are copied to ri_out alongside.
void cv_narrower_start(cv_narrower* nw, const void* x, int x_type, int64_t n, uint8_t* out, const int32_t* ri,
                       int32_t* ri_out) {
  cv_narrower_finish(nw);
  {
    std::lock_guard<std::mutex> lk(nw->mu);
    nw->x = x; nw->x_type = x_type; nw->n = n; nw->out = out;
    nw->ri = ri; nw->ri_out = ri_out;
    nw->n_slices = (n + SLICE - 1) / SLICE;
    if ((int64_t)nw->flag.size() < nw->n_slices) {
      std::vector<std::atomic<uint8_t>> f((size_t)nw->n_slices);
      nw->flag.swap(f);
    }
    for (int64_t s = 0; s < nw->n_slices; ++s) nw->flag[(size_t)s].store(0, std::memory_order_relaxed);
    nw->next.store(0, std::memory_order_relaxed);
    ++nw->generation;
  }
  nw->cv_job.notify_all();
}

// Blocks until elements [e0, e1) are converted.  1: every value in the range is an integer in
// 0..255 and out[e0, e1) holds it; 0: the range has to travel in its original type.
int cv_narrower_wait(cv_narrower* nw, int64_t e0, int64_t e1) {
  if (e1 <= e0) return 1;
  int ok = 1;
  for (int64_t s = e0 / SLICE; s <= (e1 - 1) / SLICE; ++s) {
    uint8_t f;
    int spins = 0;
    while ((f = nw->flag[(size_t)s].load(std::memory_order_acquire)) == 0) {
      if (nw->next.load(std::memory_order_relaxed) >= nw->n_slices) {
        // every slice has been handed out or dropped (cv_narrower_finish): when no worker is busy
        // nobody will ever set this flag, so the waiting thread converts the slice itself
        std::unique_lock<std::mutex> lk(nw->mu);
        if (nw->active == 0 && nw->flag[(size_t)s].load(std::memory_order_acquire) == 0) {
          nw->do_slice(s);
          continue;
        }
      }
      if (++spins > 64) std::this_thread::yield();
    }
    if (f != 1) ok = 0;
  }
  return ok;
}

// Host-only utility / test hook: narrow x[0, n) to bytes with `threads` host threads.  Returns 1
// when every value is an integer in 0..255 (out holds them), 0 otherwise, negative on bad arguments.
extern "C" int cv_host_narrow(const void* x, int x_type, int64_t n, uint8_t* out, int threads) {
  if (!x || !out || n < 0 || (x_type != CV_F64 && x_type != CV_F32 && x_type != CV_I32)) return -CV_ERR_INVALID;
  cv_narrower* nw = cv_narrower_create(threads);
  cv_narrower_start(nw, x, x_type, n, out, nullptr, nullptr);
  const int ok = cv_narrower_wait(nw, 0, n);
  cv_narrower_destroy(nw);
  return ok;
}
