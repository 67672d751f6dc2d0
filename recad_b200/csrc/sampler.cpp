// Host-side bit-exact replay of the reference samplers' consumption of numpy's legacy global
// MT19937 stream (reference: recad/dataset/implicit.py:18-35 shuffle, 50-74 pairwise_sample,
// 77-91 pointwise_sample; seeded by recad/__init__.py:11-14).
//
// numpy rule restated (numpy 2.3 RandomState.randint / choice / shuffle -> masked rejection on
// 32-bit words): for a closed range [0, r]: r == 0 consumes nothing; otherwise mask = smallest
// 2^k - 1 >= r and words are drawn until (word & mask) <= r.
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "../../include/recad_b200.h"

namespace recad { void set_error(const char* fmt, ...); }

namespace {

struct MT {
  uint32_t* key;
  int pos;
  void twist() {
    constexpr int N = 624, M = 397;
    constexpr uint32_t UP = 0x80000000u, LOW = 0x7fffffffu, MAG = 0x9908b0dfu;
    int k = 0;
    for (; k < N - M; ++k) {
      uint32_t y = (key[k] & UP) | (key[k + 1] & LOW);
      key[k] = key[k + M] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
    }
    for (; k < N - 1; ++k) {
      uint32_t y = (key[k] & UP) | (key[k + 1] & LOW);
      key[k] = key[k + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
    }
    uint32_t y = (key[N - 1] & UP) | (key[0] & LOW);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
    pos = 0;
  }
  inline uint32_t next() {
    if (pos >= 624) twist();
    uint32_t y = key[pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  inline uint64_t masked(uint64_t r) {
    if (r == 0) return 0;
    uint64_t mask = r;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = next() & (uint32_t)mask; } while (v > r);
    return v;
  }
};

}  // namespace

extern "C" {

int recad_mt19937_pairwise(uint32_t* key, int32_t* pos, int64_t n_users, int64_t n_items, int64_t train_size,
                           const int64_t* allpos_rowptr, const int32_t* allpos_col, int64_t* out, int64_t* n_out) {
  if (!key || !pos || !allpos_rowptr || !out || !n_out || n_users <= 0 || n_items <= 0 || train_size < 0 ||
      n_users > 0xffffffffLL || n_items > 0xffffffffLL) {
    recad::set_error("mt19937_pairwise: bad argument");
    return RECAD_ERR_ARG;
  }
  MT mt{key, *pos};
  // np.random.randint(0, n_users, train_size): ONE vector call, all users drawn first (implicit.py:57)
  int64_t* users = new int64_t[train_size > 0 ? train_size : 1];
  for (int64_t k = 0; k < train_size; ++k) users[k] = (int64_t)mt.masked((uint64_t)n_users - 1);
  int64_t w = 0;
  // The users are known up front, so the two dependent random accesses per sample (row pointer,
  // then the row's items) are software-prefetched: the loop is otherwise DRAM-latency bound.
  constexpr int64_t kAheadPtr = 24, kAheadRow = 12;
  for (int64_t k = 0; k < train_size; ++k) {
    if (k + kAheadPtr < train_size) __builtin_prefetch(allpos_rowptr + users[k + kAheadPtr]);
    if (k + kAheadRow < train_size) {
      const int64_t u1 = users[k + kAheadRow];
      const int64_t a = allpos_rowptr[u1], b = allpos_rowptr[u1 + 1];
      for (int64_t q = a; q < b && q < a + 192; q += 16) __builtin_prefetch(allpos_col + q);
      if (b > a + 192) __builtin_prefetch(allpos_col + (a + b) / 2);
    }
    const int64_t u = users[k];
    const int64_t lo = allpos_rowptr[u], hi = allpos_rowptr[u + 1];
    if (hi == lo) continue;  // implicit.py:63-64
    if (hi - lo >= n_items) {
      delete[] users;
      recad::set_error("mt19937_pairwise: user %lld interacted with every item; negative sampling cannot terminate",
                       (long long)u);
      return RECAD_ERR_ARG;
    }
    const int64_t p = allpos_col[lo + (int64_t)mt.masked((uint64_t)(hi - lo) - 1)];
    int64_t neg;
    do { neg = (int64_t)mt.masked((uint64_t)n_items - 1); } while (std::binary_search(allpos_col + lo, allpos_col + hi, (int32_t)neg));
    out[3 * w] = u; out[3 * w + 1] = p; out[3 * w + 2] = neg;
    ++w;
  }
  delete[] users;
  *n_out = w;
  *pos = mt.pos;
  return RECAD_OK;
}

int recad_mt19937_pointwise(uint32_t* key, int32_t* pos, int64_t n_dict_users, const int64_t* user_ids,
                            const int64_t* pos_rowptr, const int64_t* pos_items, const int64_t* pos_sorted,
                            int64_t n_items, int32_t ratio, int64_t* out) {
  if (!key || !pos || !user_ids || !pos_rowptr || !out || n_items <= 0 || ratio < 0 || n_dict_users < 0) {
    recad::set_error("mt19937_pointwise: bad argument");
    return RECAD_ERR_ARG;
  }
  MT mt{key, *pos};
  int64_t w = 0;
  for (int64_t k = 0; k < n_dict_users; ++k) {
    const int64_t lo = pos_rowptr[k], hi = pos_rowptr[k + 1], n = hi - lo, uid = user_ids[k];
    for (int64_t j = lo; j < hi; ++j) { out[3 * w] = uid; out[3 * w + 1] = pos_items[j]; out[3 * w + 2] = 1; ++w; }
    const int64_t n_neg = n * ratio;
    if (n_neg == 0) continue;
    // distinct sorted positives of this user
    const int64_t* sp = pos_sorted + lo;
    int64_t n_distinct = 0;
    for (int64_t j = 0; j < n; ++j) if (j == 0 || sp[j] != sp[j - 1]) ++n_distinct;
    const int64_t n_left = n_items - n_distinct;
    if (n_left <= 0) {
      recad::set_error("mt19937_pointwise: user %lld has no negative item to draw", (long long)uid);
      return RECAD_ERR_ARG;
    }
    for (int64_t c = 0; c < n_neg; ++c) {
      const int64_t r = (int64_t)mt.masked((uint64_t)n_left - 1);
      // r-th element of the ascending complement: r + #{distinct positives p_j with p_j - j <= r}
      int64_t t = 0, j_distinct = 0;
      if (n_distinct == n) {
        int64_t a = 0, b = n;  // binary search on the non-decreasing sequence sp[j] - j
        while (a < b) { int64_t mid = (a + b) >> 1; if (sp[mid] - mid <= r) a = mid + 1; else b = mid; }
        t = a;
      } else {
        for (int64_t j = 0; j < n; ++j) {
          if (j > 0 && sp[j] == sp[j - 1]) continue;
          if (sp[j] - j_distinct <= r) ++t; else break;
          ++j_distinct;
        }
      }
      out[3 * w] = uid; out[3 * w + 1] = r + t; out[3 * w + 2] = 0; ++w;
    }
  }
  *pos = mt.pos;
  return RECAD_OK;
}

int recad_mt19937_permutation(uint32_t* key, int32_t* pos, int64_t n, int64_t* perm) {
  if (!key || !pos || !perm || n < 0 || n > 0xffffffffLL) {
    recad::set_error("mt19937_permutation: bad argument");
    return RECAD_ERR_ARG;
  }
  MT mt{key, *pos};
  for (int64_t i = 0; i < n; ++i) perm[i] = i;
  // j = random_interval(i) depends on i only, not on the data: draw kAhead swaps ahead (same stream
  // order) and prefetch perm[j], which is otherwise one DRAM miss per element
  constexpr int kAhead = 32;
  int64_t ring[kAhead];
  int64_t drawn = n - 1;  // next i whose j has not been drawn yet
  for (int q = 0; q < kAhead && drawn > 0; ++q, --drawn) {
    ring[(n - 1 - drawn) % kAhead] = (int64_t)mt.masked((uint64_t)drawn);
    __builtin_prefetch(perm + ring[(n - 1 - drawn) % kAhead], 1);
  }
  for (int64_t i = n - 1; i > 0; --i) {
    const int slot = (int)((n - 1 - i) % kAhead);
    const int64_t j = ring[slot];
    if (drawn > 0) {
      ring[slot] = (int64_t)mt.masked((uint64_t)drawn);
      __builtin_prefetch(perm + ring[slot], 1);
      --drawn;
    }
    std::swap(perm[i], perm[j]);
  }
  *pos = mt.pos;
  return RECAD_OK;
}

}  // extern "C"
