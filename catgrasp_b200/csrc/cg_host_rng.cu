// cg_host_rng.cu -- host-side helper: the reference's per-candidate point-subset draw, bit-identical and fast.
//
// The reference draws every candidate's subset with ONE call of the GLOBAL legacy numpy generator
//     ids = np.random.choice(np.arange(M), size=n_pts, replace=(M < n_pts))        (dataset_grasp.py:72-73)
// i.e. for M >= n_pts a full Fisher-Yates shuffle of arange(M) (RandomState.permutation -> _shuffle_raw, one
// masked-rejection random_interval(i) per element, i = M-1 .. 1) and for M < n_pts n_pts masked-rejection draws
// (RandomState.randint -> random_bounded_uint64_fill, 32-bit path).  A drop-in must consume exactly the same
// MT19937 words, otherwise every later np.random call of the host program diverges from the reference run.
// Doing that through numpy costs a Python-level call + an arange + a copy per candidate (~0.3 ms at M = 20 000);
// this file restates the generator (MT19937 genrand_int32, numpy/random/src/mt19937) and the two draw loops in C:
// the caller passes numpy's state in (np.random.get_state()), gets ids for `count` candidates and the advanced
// state back (np.random.set_state()).  tests/test_abi_and_host.py checks ids AND the post-state against numpy itself.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/catgrasp_b200.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;

struct Mt {
  uint32_t *key;
  int pos;
  inline void refill() {
    uint32_t *mt = key;
    int kk;
    uint32_t y;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < MT_N - 1; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    pos = 0;
  }
  inline uint32_t next() {
    if (pos == MT_N) refill();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
};

inline uint32_t mask_of(uint32_t max) {
  uint32_t mask = max;
  mask |= mask >> 1;
  mask |= mask >> 2;
  mask |= mask >> 4;
  mask |= mask >> 8;
  mask |= mask >> 16;
  return mask;
}

}  // namespace

// One candidate's shuffle from generator state g (advanced in place); writes the first n_pts entries to out.
static void shuffle_one(Mt &g, int64_t M, int32_t n_pts, int32_t *p, int32_t *out) {
  uint32_t *key = g.key;
  for (int32_t i = 0; i < (int32_t)M; i++) p[i] = i;
  uint32_t mask = mask_of((uint32_t)(M - 1));
  // Branch-free form of "do j = next & mask while (j > i); swap(p[i], p[j]); i--": a rejected draw swaps p[i]
  // with itself and leaves i alone, so the ~50 % unpredictable rejections cost no pipeline flush.
  uint32_t i = (uint32_t)(M - 1);
  while (i >= 1) {
    if (g.pos == MT_N) g.refill();
    const int avail = MT_N - g.pos;
    const uint32_t *kp = key + g.pos;
    int k = 0;
    for (; k < avail && i >= 1; k++) {
      uint32_t y = kp[k];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      const uint32_t j = y & mask;
      const uint32_t acc = (j <= i) ? 1u : 0u;
      const uint32_t jj = acc ? j : i;
      const int32_t t = p[jj];
      p[jj] = p[i];
      p[i] = t;
      i -= acc;
      if ((mask >> 1) >= i) mask >>= 1;   // smallest all-ones mask >= i
    }
    g.pos += k;
  }
  memcpy(out, p, (size_t)n_pts * sizeof(int32_t));
}

// Advances g exactly as shuffle_one would, without touching a permutation (which words are accepted does not depend
// on the permutation): the sequential part of the multi-threaded draw.  The tempering + masking of a block of raw words
// has no loop-carried dependency (the compiler vectorises it); only the accept scan `i -= (j <= i)` is serial, and the mask
// is constant while i stays above mask >> 1.
static void shuffle_skip(Mt &g, int64_t M) {
  uint32_t mask = mask_of((uint32_t)(M - 1));
  uint32_t i = (uint32_t)(M - 1);
  uint32_t tmp[MT_N];
  while (i >= 1) {
    if (g.pos == MT_N) g.refill();
    const int avail = MT_N - g.pos;
    const uint32_t *kp = g.key + g.pos;
    for (int k = 0; k < avail; k++) {
      uint32_t y = kp[k];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      tmp[k] = y;
    }
    int k = 0;
    while (k < avail && i >= 1) {
      // segment with a constant mask: i in (mask >> 1, mask]
      const uint32_t lim = mask >> 1;
      const uint32_t m = mask;
      while (k < avail && i > lim) {
        i -= ((tmp[k] & m) <= i) ? 1u : 0u;
        k++;
      }
      if (i <= lim) mask >>= 1;   // i crossed below the next power of two (i >= 1 keeps mask >= 1)
    }
    g.pos += k;
  }
}

// key: 624 words, *pos in [0, 624] (numpy's state tuple fields 1 and 2), both updated in place.
// out: (count, n_pts) int32.  M < 2^31.  nthreads <= 0: one worker per host core (at most 32).
// The generator is inherently sequential, but which words a candidate consumes does not depend on its permutation:
// the calling thread walks the stream (snapshotting the state at every candidate boundary) while worker threads
// replay the real shuffles from those snapshots.
extern "C" int cg_host_legacy_choice(uint32_t *key, int32_t *pos, int64_t M, int32_t n_pts, int32_t count, int32_t *out,
                                     int32_t nthreads) {
  if (!key || !pos || !out || M <= 0 || M >= (1ll << 31) || n_pts <= 0 || count < 0 || *pos < 0 || *pos > MT_N)
    return CG_EINVAL;
  Mt g{key, *pos};
  if (M < n_pts) {
    // replace=True: randint(0, M, size=n_pts): value = next_uint32 & mask until value <= M-1 (masked rejection)
    const uint32_t rng = (uint32_t)(M - 1);
    const uint32_t mask = mask_of(rng);
    for (int64_t c = 0; c < count; c++) {
      int32_t *o = out + c * (int64_t)n_pts;
      for (int i = 0; i < n_pts; i++) {
        uint32_t v;
        if (rng == 0) v = 0;   // random_bounded_uint64_fill: rng == 0 consumes nothing
        else
          while ((v = (g.next() & mask)) > rng) {
          }
        o[i] = (int32_t)v;
      }
    }
    *pos = g.pos;
    return CG_OK;
  }
  // replace=False: permutation(M)[:n_pts]: shuffle arange(M) from the top, j = random_interval(i)
  if (nthreads <= 0) {
    nthreads = (int32_t)std::thread::hardware_concurrency();
    if (nthreads > 32) nthreads = 32;
  }
  if (nthreads < 2 || count < 4 || M < 2048) {
    std::vector<int32_t> perm((size_t)M);
    for (int64_t c = 0; c < count; c++) shuffle_one(g, M, n_pts, perm.data(), out + c * (int64_t)n_pts);
    *pos = g.pos;
    return CG_OK;
  }
  struct Snap { uint32_t key[MT_N]; int pos; };
  std::vector<Snap> snaps((size_t)count);
  std::atomic<int> ready(0), next(0);
  auto worker = [&]() {
    std::vector<int32_t> perm((size_t)M);
    for (;;) {
      const int c = next.fetch_add(1);
      if (c >= count) return;
      while (ready.load(std::memory_order_acquire) <= c) std::this_thread::yield();
      Mt lg{snaps[(size_t)c].key, snaps[(size_t)c].pos};
      shuffle_one(lg, M, n_pts, perm.data(), out + (int64_t)c * n_pts);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads - 1; t++) pool.emplace_back(worker);
  for (int c = 0; c < count; c++) {
    memcpy(snaps[(size_t)c].key, key, sizeof(uint32_t) * MT_N);
    snaps[(size_t)c].pos = g.pos;
    ready.store(c + 1, std::memory_order_release);
    shuffle_skip(g, M);
  }
  worker();   // the walking thread helps with what is left
  for (auto &t : pool) t.join();
  *pos = g.pos;
  return CG_OK;
}

// Advance the generator over `count` candidates without producing indices (a rank that scores candidates [lo, hi) of a
// list still has to leave numpy's generator where the reference's full loop leaves it).
extern "C" int cg_host_legacy_skip(uint32_t *key, int32_t *pos, int64_t M, int32_t n_pts, int32_t count) {
  if (!key || !pos || M <= 0 || M >= (1ll << 31) || n_pts <= 0 || count < 0 || *pos < 0 || *pos > MT_N) return CG_EINVAL;
  Mt g{key, *pos};
  if (M < n_pts) {
    const uint32_t rng = (uint32_t)(M - 1);
    const uint32_t mask = mask_of(rng);
    if (rng != 0)
      for (int64_t i = 0; i < (int64_t)count * n_pts; i++)
        while ((g.next() & mask) > rng) {
        }
  } else {
    for (int c = 0; c < count; c++) shuffle_skip(g, M);
  }
  *pos = g.pos;
  return CG_OK;
}
