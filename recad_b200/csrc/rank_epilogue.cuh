// Per-row ranking state shared by the two full-rank kernels (eval.cu: exact CUDA-core, eval_tc.cu: tcgen05).
// A thread owns one user; it is handed the user's scores for NC consecutive items at a time (registers) and
// must (a) count, per target, the items that outrank it and (b) keep the K best (score desc, id asc).
//
// The per-score cost decides the kernel's speed (2e11 scores on the synthetic workload), so:
//   * masked / out-of-range scores are set to -inf up front (only when the chunk has any), after which no
//     per-score branch is needed: -inf never counts and never enters the list;
//   * the tie rule "outranks iff s > s_t, or s == s_t and id < target" becomes ONE compare per score: for a
//     chunk entirely before the target the threshold is nextafter(s_t, -inf) (s >= s_t  <=>  s > that), for
//     a chunk entirely after it s_t; only the single chunk that contains the target takes the exact path;
//   * the top-K list is touched only when the chunk maximum beats the current K-th best.
#pragma once
#include <math.h>
#include <stdint.h>

namespace recad {

template <int TMAX>
struct RankState {
  static constexpr int N = TMAX > 0 ? TMAX : 1;
  float st[N];      // target score
  float st_lo[N];   // nextafter(st, -inf)
  int tg[N];        // target item id
  int rk[N];        // running count
  unsigned in_train = 0;  // bit t: the target is one of the user's train items (reported rank -1)
};

// Insert (v, item) after every entry >= v (items arrive in ascending id, so ties keep id order); returns the
// new K-th best.  Rare (about K ln(I / K) times per user), hence out of line.
static __device__ __noinline__ float topk_insert(float v, int32_t item, float* __restrict__ topv, int32_t* __restrict__ topi, int K,
                                          int stride, int slot) {
  int p = K - 1;
  while (p > 0 && topv[(p - 1) * stride + slot] < v) {
    topv[p * stride + slot] = topv[(p - 1) * stride + slot];
    topi[p * stride + slot] = topi[(p - 1) * stride + slot];
    --p;
  }
  topv[p * stride + slot] = v;
  topi[p * stride + slot] = item;
  return topv[(K - 1) * stride + slot];
}

template <int NC, int TMAX>
__device__ __forceinline__ void rank_topk_chunk(float (&s)[NC], int64_t base, int T, RankState<TMAX>& rs, float& tau,
                                                float* __restrict__ topv, int32_t* __restrict__ topi, int K, int stride,
                                                int slot) {
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < T) {
      const int tg = rs.tg[t];
      if (base + NC <= tg || base > tg) {          // warp-uniform: targets are the same for every user
        const float thr = base > tg ? rs.st[t] : rs.st_lo[t];
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;          // four independent chains: one warp per scheduler has
#pragma unroll                                     // nothing else to hide the add latency behind
        for (int e = 0; e < NC; e += 4) {
          c0 += s[e] > thr ? 1 : 0;
          c1 += s[e + 1] > thr ? 1 : 0;
          c2 += s[e + 2] > thr ? 1 : 0;
          c3 += s[e + 3] > thr ? 1 : 0;
        }
        rs.rk[t] += (c0 + c1) + (c2 + c3);
      } else {                                     // the one chunk that holds the target itself
#pragma unroll
        for (int e = 0; e < NC; ++e) {
          const int item = (int)base + e;
          if (item == tg) continue;
          rs.rk[t] += (item < tg ? s[e] >= rs.st[t] : s[e] > rs.st[t]) ? 1 : 0;
        }
      }
    }
  }
  float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
  for (int e = 4; e < NC; e += 4) {
    m0 = fmaxf(m0, s[e]);
    m1 = fmaxf(m1, s[e + 1]);
    m2 = fmaxf(m2, s[e + 2]);
    m3 = fmaxf(m3, s[e + 3]);
  }
  const float cmax = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  if (cmax > tau) {
#pragma unroll
    for (int e = 0; e < NC; ++e)
      if (s[e] > tau) tau = topk_insert(s[e], (int32_t)(base + e), topv, topi, K, stride, slot);
  }
}

template <int TMAX>
__device__ __forceinline__ void rank_state_init(RankState<TMAX>& rs, int T, const int32_t* __restrict__ targets,
                                                const int32_t* __restrict__ train_col, int64_t lo0, int64_t hi0) {
#pragma unroll
  for (int t = 0; t < RankState<TMAX>::N; ++t) {
    rs.st[t] = 0.f; rs.st_lo[t] = 0.f; rs.tg[t] = -1; rs.rk[t] = 0;
    if (t < T && t < TMAX) {
      rs.tg[t] = targets[t];
      int64_t lo = lo0, hi = hi0;
      while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (train_col[mid] < rs.tg[t]) lo = mid + 1; else hi = mid; }
      if (lo < hi0 && train_col[lo] == rs.tg[t]) rs.in_train |= 1u << t;
    }
  }
}

template <int TMAX>
__device__ __forceinline__ void rank_state_set_score(RankState<TMAX>& rs, int t, float score) {
  rs.st[t] = score;
  rs.st_lo[t] = nextafterf(score, -INFINITY);
}

}  // namespace recad
