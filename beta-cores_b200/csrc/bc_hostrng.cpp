// beta-cores B200, HOST code: numpy's legacy global random stream (RandomState: MT19937 + polar Box-Muller), natively.
//
// Every optimiser step of the reference calls the user's sampler, which ends in `np.random.randn(S, D)`
// (examples/zellner_gaussian/main.py:87-92, zellner_logreg/main.py:139-144), and the sub-sampled modes then draw
// `np.random.randint(N, size=n)` (bayesiancoresets/coreset/bcores.py:53).  Reproducing the reference's coresets means
// consuming that one global stream in exactly its order -- and numpy's legacy generator needs 25-30 ns per normal: at
// S x D = 200 x 100 that is 0.3-0.6 ms per step on one host core, more than all the kernels of a step of the Gaussian example
// take together.  The functions below produce the SAME numbers (bit for bit: same integer stream, same floating-point
// operations in the same order, the process's own libm `log`) several times faster:
//   * the Mersenne-Twister words are generated and tempered in bulk (vectorisable loops);
//   * the polar method's accept/reject pass runs over the word buffer without a function call per attempt;
//   * the expensive part, f = sqrt(-2 log(r2) / r2) of every accepted pair, has no dependence between pairs: it is
//     split over a few worker threads.
// The state is numpy's own (`np.random.get_state()`: key[624], pos, has_gauss, cached_gaussian), so the caller can check
// the stream out of numpy, draw natively, and hand it back at any point.
//
// numpy sources restated (numpy/random/src/mt19937/mt19937.c, src/legacy/legacy-distributions.c, _bounded_integers.pyx):
//   mt19937_gen / mt19937_next (tempering), mt19937_next_double = ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0,
//   legacy_gauss (polar method; returns f * x2 and caches f * x1), RandomState.randint -> masked rejection on uint32.
// This unit must be compiled WITHOUT floating-point contraction (no FMA: numpy's x86-64 baseline build has none); the
// CPU tests compare against np.random on long streams.
#include <stdint.h>
#include <math.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/betacores.h"

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

// The three bulk loops (state regeneration, tempering, word -> attempt conversion) are compiled twice, for the x86-64
// baseline and for AVX2, and picked at load time (GCC function multi-versioning).  Neither version may fuse a multiply
// with an add: the unit is built with -ffp-contract=off, and "avx2" does not include FMA.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define BC_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define BC_CLONES
#endif

#define BC_MT_TWIST(u, v) ((((u) & kUpper) | ((v) & kLower)) >> 1) ^ ((uint32_t)(-(int32_t)((v) & 1)) & kMatrixA)
// mt19937_gen: key[i] = key[(i + M) mod N] ^ twist(key[i], key[(i + 1) mod N]).  Written as ranges inside which no
// iteration reads a word another iteration of the same range writes (word i + M - N was finished a whole range earlier),
// so that each range vectorises.
BC_CLONES void mt_gen(uint32_t* key) {
  int i;
#pragma GCC ivdep
  for (i = 0; i < kN - kM; ++i) key[i] = key[i + kM] ^ BC_MT_TWIST(key[i], key[i + 1]);                 // 0 .. 226 reads 397 .. 623 (old)
#pragma GCC ivdep
  for (; i < 2 * (kN - kM); ++i) key[i] = key[i + (kM - kN)] ^ BC_MT_TWIST(key[i], key[i + 1]);         // 227 .. 453 reads 0 .. 226 (new)
#pragma GCC ivdep
  for (; i < kN - 1; ++i) key[i] = key[i + (kM - kN)] ^ BC_MT_TWIST(key[i], key[i + 1]);                // 454 .. 622 reads 227 .. 395 (new)
  key[kN - 1] = key[kM - 1] ^ BC_MT_TWIST(key[kN - 1], key[0]);
}

inline uint32_t temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
BC_CLONES void temper_block(const uint32_t* __restrict__ k, uint32_t* __restrict__ o, int n) {
  for (int i = 0; i < n; ++i) o[i] = temper(k[i]);
}
// `na` attempts of the polar method from 4 na words: x1, x2 and r2 = x1^2 + x2^2 (legacy_gauss with mt19937_next_double)
BC_CLONES void attempts(const uint32_t* __restrict__ b, int na, double* __restrict__ tx1, double* __restrict__ tx2, double* __restrict__ tr2) {
  for (int a = 0; a < na; ++a) {
    const double d1 = ((double)(int32_t)(b[4 * a] >> 5) * 67108864.0 + (double)(int32_t)(b[4 * a + 1] >> 6)) / 9007199254740992.0;
    const double d2 = ((double)(int32_t)(b[4 * a + 2] >> 5) * 67108864.0 + (double)(int32_t)(b[4 * a + 3] >> 6)) / 9007199254740992.0;
    const double x1 = 2.0 * d1 - 1.0, x2 = 2.0 * d2 - 1.0;
    tx1[a] = x1;
    tx2[a] = x2;
    tr2[a] = x1 * x1 + x2 * x2;
  }
}

// A window of tempered words in stream order, refilled one state block at a time (at most 3 unconsumed words are carried
// over a refill, so the window never holds more than one block plus those).  release() puts the state's `pos` back to the
// first unconsumed word: after at least one 4-word attempt past a refill every unconsumed word lies in the current block.
struct Words {
  bc_mt_state* st;
  uint32_t buf[kN + 4];
  int at = 0, end = 0;
  explicit Words(bc_mt_state* s) : st(s) {}
  void refill() {
    const int left = end - at;
    for (int i = 0; i < left; ++i) buf[i] = buf[at + i];
    if (st->pos >= kN) {
      mt_gen(st->key);
      st->pos = 0;
    }
    const int take = kN - st->pos;
    const uint32_t* k = st->key + st->pos;
    uint32_t* o = buf + left;
    temper_block(k, o, take);
    st->pos = kN;
    at = 0;
    end = left + take;
  }
  void release() { st->pos = kN - (end - at); }
};

// ---- a small persistent worker pool for the per-pair transcendental part ----
// The calling thread walks the word stream (sequential by nature) and hands every kChunk accepted pairs to the pool as
// they become available, so the log / sqrt of earlier pairs runs while later words are still being generated; at the end
// the caller helps to drain the queue.  Workers sleep on a condition variable between calls.
constexpr int64_t kChunk = 1024;
struct Pool {
  struct Task { int64_t lo, hi; };
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  std::vector<std::thread> th;
  std::vector<Task> q;
  int inflight = 0;
  const double* x1 = nullptr;
  const double* x2 = nullptr;
  const double* r2 = nullptr;
  double* out = nullptr;
  bool stop = false;

  static void work(const double* x1, const double* x2, const double* r2, double* out, int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      const double f = sqrt(-2.0 * log(r2[i]) / r2[i]);
      out[2 * i] = f * x2[i];       // legacy_gauss returns f * x2 first and keeps f * x1 for the next call
      out[2 * i + 1] = f * x1[i];
    }
  }
  void loop() {
    std::unique_lock<std::mutex> lk(m);
    for (;;) {
      cv_go.wait(lk, [&] { return stop || !q.empty(); });
      if (stop) return;
      const Task t = q.back();
      q.pop_back();
      const double *a = x1, *b = x2, *c = r2;
      double* o = out;
      lk.unlock();
      work(a, b, c, o, t.lo, t.hi);
      lk.lock();
      if (--inflight == 0) cv_done.notify_all();
    }
  }
  void begin(const double* a, const double* b, const double* c, double* o, int threads) {
    std::unique_lock<std::mutex> lk(m);
    while ((int)th.size() < threads - 1) th.emplace_back([this] { loop(); });
    x1 = a; x2 = b; r2 = c; out = o;
  }
  void submit(int64_t lo, int64_t hi) {
    {
      std::unique_lock<std::mutex> lk(m);
      q.push_back({lo, hi});
      ++inflight;
    }
    cv_go.notify_one();
  }
  void finish() {   // the caller takes queued chunks itself, then waits for the ones in flight
    std::unique_lock<std::mutex> lk(m);
    while (!q.empty()) {
      const Task t = q.back();
      q.pop_back();
      lk.unlock();
      work(x1, x2, r2, out, t.lo, t.hi);
      lk.lock();
      --inflight;
    }
    cv_done.wait(lk, [&] { return inflight == 0; });
  }
};

Pool& pool() {
  static Pool* p = new Pool();   // never destroyed: worker threads must not be joined from a static destructor at exit
  return *p;
}
std::mutex g_call;               // one native draw at a time (the pool serves one request)

}  // namespace

extern "C" int bc_mt_randn(bc_mt_state* st, double* h_out, int64_t n, int threads) {
  if (!st || (!h_out && n > 0) || n < 0 || st->pos < 0 || st->pos > kN) return BC_ERR_ARG;
  if (n == 0) return BC_OK;
  std::lock_guard<std::mutex> call(g_call);
  int64_t done = 0;
  if (st->has_gauss) {
    h_out[done++] = st->gauss;
    st->has_gauss = 0;
    st->gauss = 0.0;
  }
  const int64_t need = n - done;            // normals still to produce
  const int64_t npairs = (need + 1) / 2;    // accepted attempts needed
  if (npairs == 0) return BC_OK;
  static thread_local std::vector<double> X1, X2, R2, OUT;
  X1.resize(npairs + 1); X2.resize(npairs + 1); R2.resize(npairs + 1);
  // an odd count leaves the last pair's second value cached, exactly like legacy_gauss: such a request goes through a
  // scratch array of whole pairs
  const bool whole = need == 2 * npairs;
  if (!whole) OUT.resize(2 * npairs);
  double* dst = whole ? h_out + done : OUT.data();
  if (threads > 16) threads = 16;
  const bool par = threads > 1 && npairs >= 2 * kChunk;
  Pool& P = pool();
  if (par) P.begin(X1.data(), X2.data(), R2.data(), dst, threads);
  int64_t handed = 0;
  Words w(st);
  int64_t got = 0;
  while (got < npairs) {
    if (w.end - w.at < 4) w.refill();
    // every attempt of the polar method consumes exactly four words, accepted or not: the window is a run of whole
    // attempts.  First all of them at once (a loop without branches: it vectorises), then the accepted ones are compacted in
    // stream order, stopping at the attempt that completes the request.
    const uint32_t* b = w.buf + w.at;
    const int na = (w.end - w.at) / 4;
    double tx1[(kN + 4) / 4], tx2[(kN + 4) / 4], tr2[(kN + 4) / 4];
    attempts(b, na, tx1, tx2, tr2);
    int a = 0;
    for (; a < na && got < npairs; ++a) {
      const double r2 = tr2[a];
      X1[got] = tx1[a]; X2[got] = tx2[a]; R2[got] = r2;         // (slot npairs exists: the arrays hold one spare entry)
      got += (r2 < 1.0 && r2 != 0.0) ? 1 : 0;                    // legacy_gauss: retry while r2 >= 1.0 || r2 == 0.0
    }
    w.at += 4 * a;
    if (par) {
      while (got - handed >= kChunk) {
        P.submit(handed, handed + kChunk);
        handed += kChunk;
      }
    }
  }
  w.release();
  if (par) {
    if (handed < npairs) P.submit(handed, npairs);
    P.finish();
  } else {
    Pool::work(X1.data(), X2.data(), R2.data(), dst, 0, npairs);
  }
  if (!whole) {
    memcpy(h_out + done, OUT.data(), sizeof(double) * (size_t)need);
    st->gauss = OUT[2 * npairs - 1];
    st->has_gauss = 1;
  }
  return BC_OK;
}

extern "C" int bc_mt_randint(bc_mt_state* st, int64_t high, int64_t* h_out, int64_t n) {
  // RandomState.randint(high, size=n) with the default dtype: _rand_int64(0, high - 1, masked rejection); ranges below 2^32
  // draw one 32-bit word per attempt (buffered_bounded_masked_uint32), which is every use on this path
  if (!st || (!h_out && n > 0) || n < 0 || high < 1 || st->pos < 0 || st->pos > kN) return BC_ERR_ARG;
  const uint64_t rng = (uint64_t)high - 1;
  if (rng > 0xFFFFFFFFull) return BC_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> call(g_call);
  if (rng == 0) {
    for (int64_t i = 0; i < n; ++i) h_out[i] = 0;
    return BC_OK;
  }
  uint32_t mask = (uint32_t)rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  for (int64_t i = 0; i < n; ++i) {
    uint32_t v;
    do {
      if (st->pos >= kN) {
        mt_gen(st->key);
        st->pos = 0;
      }
      v = temper(st->key[st->pos++]) & mask;
    } while (v > (uint32_t)rng);
    h_out[i] = (int64_t)v;
  }
  return BC_OK;
}
