// Full-ranking evaluation without materialising the score matrix
// (reference: recad/workflow/normal.py:57-93 user_item_model_generate + 111-160 normal_evaluate,
//  and lightgcn.py:115-120 getUsersRating for the Recall/NDCG entry).
//
// v1 (this file): exact fp32 CUDA-core contraction.  ONE THREAD OWNS ONE USER: its embedding row
// lives in registers (DP floats), item tiles [DP x 64] stream through shared memory (cp.async,
// double buffered) from the TRANSPOSED item table, every inner-product term is an FMA in ascending
// d, so a score is bit-identical wherever it is recomputed (the target's score is computed once up
// front with the same sequence).  The user's score never leaves registers: train-item mask (a 64-bit
// register built by walking the user's sorted train list), target-rank counters and a threshold-
// filtered top-K insertion list (shared memory, touched only when a score beats the current K-th)
// are all applied in the epilogue of each 16-item register block.
#include <math.h>

#include "common.cuh"
#include "rank_epilogue.cuh"

namespace recad {

constexpr int kEvalThreads = 128;  // users per CTA
constexpr int kTI = 64;            // items per tile
constexpr int kJB = 16;            // items per register block
constexpr int kMaxT = 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int DP, int TMAX>
__global__ void __launch_bounds__(kEvalThreads)
fullrank_kernel(const float* __restrict__ user_emb, const float* __restrict__ item_T, int64_t ld, int64_t n_items,
                int D, const int64_t* __restrict__ user_ids, int64_t n_eval, const int64_t* __restrict__ train_rowptr,
                const int32_t* __restrict__ train_col, const int32_t* __restrict__ targets, int T, int K,
                int32_t* __restrict__ topk_idx, float* __restrict__ topk_val, int32_t* __restrict__ target_rank,
                float* __restrict__ target_score) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);                       // [2][DP][kTI]
  float* topv = tile + 2 * DP * kTI;                                      // [K][kEvalThreads]
  int32_t* topi = reinterpret_cast<int32_t*>(topv + (size_t)K * kEvalThreads);  // [K][kEvalThreads]

  const int tid = threadIdx.x;
  const int64_t g = (int64_t)blockIdx.x * kEvalThreads + tid;
  const bool active = g < n_eval;
  const int64_t uid = active ? user_ids[g] : 0;

  float a[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) a[d] = (active && d < D) ? user_emb[uid * D + d] : 0.f;
  for (int k = 0; k < K; ++k) { topv[k * kEvalThreads + tid] = -INFINITY; topi[k * kEvalThreads + tid] = -1; }
  float tau = -INFINITY;  // current K-th best
  int min_pos = 0;        // its slot in the (unsorted) list

  int64_t cur = 0, end = 0;
  if (active) { cur = train_rowptr[uid]; end = train_rowptr[uid + 1]; }

  // target scores with the SAME fma sequence as the tiles; rank = -1 if the target is a train item
  RankState<TMAX> rs;
  rank_state_init(rs, T, targets, train_col, cur, end);
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < T) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) acc = fmaf(a[d], d < D ? item_T[(int64_t)d * ld + rs.tg[t]] : 0.f, acc);
      rank_state_set_score(rs, t, acc);
    }
  }

  const int64_t n_tiles = (n_items + kTI - 1) / kTI;
  constexpr int kChunks = DP * kTI / 4;  // 16-byte chunks per tile
  auto load_tile = [&](int64_t t, int buf) {
    float* dst = tile + buf * DP * kTI;
    const int64_t j0 = t * kTI;
    for (int c = tid; c < kChunks; c += kEvalThreads) {
      const int d = c / (kTI / 4), q = c % (kTI / 4);
      if (d < D) cp_async16(dst + d * kTI + q * 4, item_T + (int64_t)d * ld + j0 + q * 4);
    }
    cp_async_commit();
  };
  // rows d >= D of the tile are never loaded: zero them once (a[d] is 0 there, but 0 * garbage may be NaN)
  for (int c = tid; c < 2 * DP * kTI; c += kEvalThreads) {
    const int d = (c / kTI) % DP;
    if (d >= D) tile[c] = 0.f;
  }
  load_tile(0, 0);
  for (int64_t t = 0; t < n_tiles; ++t) {
    const int buf = (int)(t & 1);
    if (t + 1 < n_tiles) { load_tile(t + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* bt = tile + buf * DP * kTI;
    const int64_t j0 = t * kTI;
    // train-item mask of this tile for this thread's user
    unsigned long long mask = 0ull;
    while (cur < end) {
      const int c = train_col[cur];
      if (c >= j0 + kTI) break;
      if (c >= j0) mask |= 1ull << (c - j0);
      ++cur;
    }
#pragma unroll 1
    for (int jb = 0; jb < kTI; jb += kJB) {
      float acc[kJB];
#pragma unroll
      for (int j = 0; j < kJB; ++j) acc[j] = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float4* b4 = reinterpret_cast<const float4*>(bt + d * kTI + jb);
#pragma unroll
        for (int q = 0; q < kJB / 4; ++q) {
          const float4 b = b4[q];  // all lanes read the same address: shared-memory broadcast
          acc[4 * q + 0] = fmaf(a[d], b.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(a[d], b.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(a[d], b.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(a[d], b.w, acc[4 * q + 3]);
        }
      }
      {
        const int64_t base = j0 + jb;
        const int lim = (int)min((int64_t)kJB, n_items - base);
        if (lim > 0) {
          uint32_t drop = (uint32_t)((mask >> jb) & 0xffffull) | (lim < kJB ? (0xffffu << lim) & 0xffffu : 0u) |
                          (active ? 0u : 0xffffu);
          if (drop) {
#pragma unroll
            for (int j = 0; j < kJB; ++j) acc[j] = ((drop >> j) & 1u) ? -INFINITY : acc[j];
          }
          rank_topk_chunk<kJB, TMAX>(acc, base, T, rs, tau, topv, topi, K, kEvalThreads, tid, min_pos);
        }
      }
    }
    __syncthreads();  // everyone is done with `buf` before it is refilled two iterations later
  }
  if (active) {
    topk_finalize(topv, topi, K, kEvalThreads, tid);
    for (int k = 0; k < K; ++k) {
      topk_idx[g * K + k] = topi[k * kEvalThreads + tid];
      topk_val[g * K + k] = topv[k * kEvalThreads + tid];
    }
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < T) {
        target_rank[g * T + t] = ((rs.in_train >> t) & 1u) ? -1 : rs.rk[t];
        target_score[g * T + t] = rs.st[t];
      }
  }
}

__global__ void transpose_items_kernel(const float* __restrict__ in, int64_t n_items, int D, float* __restrict__ out,
                                       int64_t ld) {
  __shared__ float t[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r;
    const int d = d0 + threadIdx.x;
    t[r][threadIdx.x] = (i < n_items && d < D) ? in[i * D + d] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r;
    const int64_t i = i0 + threadIdx.x;
    if (d < D && i < ld) out[(int64_t)d * ld + i] = t[threadIdx.x][r];
  }
}

__global__ void recall_ndcg_kernel(const int32_t* __restrict__ topk_idx, int64_t n_eval, int K,
                                   const int64_t* __restrict__ user_ids, const int64_t* __restrict__ gt_rowptr,
                                   const int32_t* __restrict__ gt_col, double* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double rec = 0.0, ndcg = 0.0, one = 0.0;
  if (g < n_eval) {
    const int64_t u = user_ids[g];
    const int64_t lo0 = gt_rowptr[u], hi0 = gt_rowptr[u + 1];
    const int64_t ngt = hi0 - lo0;
    if (ngt > 0) {
      double dcg = 0.0, idcg = 0.0;
      int hits = 0;
      for (int k = 0; k < K; ++k) {
        const double disc = 1.0 / log2((double)(k + 2));
        if (k < ngt) idcg += disc;
        const int item = topk_idx[g * K + k];
        if (item < 0) continue;
        int64_t lo = lo0, hi = hi0;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (gt_col[mid] < item) lo = mid + 1; else hi = mid; }
        if (lo < hi0 && gt_col[lo] == item) { ++hits; dcg += disc; }
      }
      rec = (double)hits / (double)ngt;
      ndcg = dcg / idcg;
      one = 1.0;
    }
  }
  rec = warp_sum(rec); ndcg = warp_sum(ndcg); one = warp_sum(one);
  if ((threadIdx.x & 31) == 0 && one > 0.0) {
    atomicAdd(out + 0, rec);
    atomicAdd(out + 1, ndcg);
    atomicAdd(out + 2, one);
  }
}

// One warp per row of a materialised score block (NCF).  Items are visited in ascending id,
// 32 at a time; a lane's score enters the warp's shared top-K list only when it beats the
// current K-th, lowest lane (= lowest id) first, so ties keep id order.
constexpr int kRankWarps = 4;
__global__ void __launch_bounds__(kRankWarps * 32)
rank_from_scores_kernel(const float* __restrict__ scores, int64_t n_rows, int64_t n_items,
                        const int64_t* __restrict__ user_ids, const int64_t* __restrict__ train_rowptr,
                        const int32_t* __restrict__ train_col, const int32_t* __restrict__ targets, int T, int K,
                        int32_t* __restrict__ topk_idx, float* __restrict__ topk_val,
                        int32_t* __restrict__ target_rank, float* __restrict__ target_score) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* topv = reinterpret_cast<float*>(smem_raw) + (size_t)w * K;
  int32_t* topi = reinterpret_cast<int32_t*>(reinterpret_cast<float*>(smem_raw) + (size_t)kRankWarps * K) + (size_t)w * K;
  const int64_t r = (int64_t)blockIdx.x * kRankWarps + w;
  if (r >= n_rows) return;
  const int64_t uid = user_ids[r];
  const int64_t lo0 = train_rowptr[uid], hi0 = train_rowptr[uid + 1];
  const float* __restrict__ srow = scores + r * n_items;
  for (int k = lane; k < K; k += 32) { topv[k] = -INFINITY; topi[k] = -1; }
  __syncwarp();
  float st[kMaxT];
  int tg[kMaxT], rk[kMaxT];
  bool tmask[kMaxT];
#pragma unroll
  for (int t = 0; t < kMaxT; ++t) {
    st[t] = 0.f; tg[t] = -1; rk[t] = 0; tmask[t] = false;
    if (t < T) {
      tg[t] = targets[t];
      st[t] = srow[tg[t]];
      int64_t lo = lo0, hi = hi0;
      while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (train_col[mid] < tg[t]) lo = mid + 1; else hi = mid; }
      tmask[t] = lo < hi0 && train_col[lo] == tg[t];
    }
  }
  float tau = -INFINITY;
  for (int64_t j0 = 0; j0 < n_items; j0 += 32) {
    const int64_t item = j0 + lane;
    bool ok = item < n_items;
    float s = -INFINITY;
    if (ok) {
      s = srow[item];
      int64_t lo = lo0, hi = hi0;
      while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (train_col[mid] < item) lo = mid + 1; else hi = mid; }
      if (lo < hi0 && train_col[lo] == item) ok = false;
    }
    if (ok) {
#pragma unroll
      for (int q = 0; q < kMaxT; ++q)
        if (q < T && (s > st[q] || (s == st[q] && item < tg[q]))) ++rk[q];
    }
    unsigned cand = __ballot_sync(kFull, ok && s > tau);
    while (cand) {
      const int src = __ffs(cand) - 1;
      const float cs = __shfl_sync(kFull, s, src);
      const int ci = (int)(j0 + src);
      if (lane == 0) {
        int p = K - 1;
        while (p > 0 && topv[p - 1] < cs) { topv[p] = topv[p - 1]; topi[p] = topi[p - 1]; --p; }
        topv[p] = cs; topi[p] = ci;
      }
      __syncwarp();
      tau = topv[K - 1];
      cand &= cand - 1;
      cand &= __ballot_sync(kFull, ok && s > tau);  // the bar moved: drop lanes that no longer qualify
    }
  }
  for (int k = lane; k < K; k += 32) { topk_idx[r * K + k] = topi[k]; topk_val[r * K + k] = topv[k]; }
#pragma unroll
  for (int t = 0; t < kMaxT; ++t) {
    if (t < T) {
      int tot = rk[t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
      if (lane == 0) {
        target_rank[r * T + t] = tmask[t] ? -1 : tot;
        target_score[r * T + t] = st[t];
      }
    }
  }
}

template <int DP, int TMAX>
static int launch_fullrank_t(const float* user_emb, const float* item_T, int64_t ld, int64_t n_items, int D,
                           const int64_t* user_ids, int64_t n_eval, const int64_t* train_rowptr,
                           const int32_t* train_col, const int32_t* targets, int T, int K, int32_t* topk_idx,
                           float* topk_val, int32_t* target_rank, float* target_score, cudaStream_t s) {
  const size_t smem = (size_t)2 * DP * kTI * 4 + (size_t)K * kEvalThreads * 8;
  RECAD_REQUIRE(smem <= 227 * 1024, RECAD_ERR_UNSUPPORTED, "fullrank: K = %d needs %zu B of shared memory", K, smem);
  RECAD_CUDA_CHECK(cudaFuncSetAttribute(fullrank_kernel<DP, TMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((n_eval + kEvalThreads - 1) / kEvalThreads);
  fullrank_kernel<DP, TMAX><<<grid, kEvalThreads, smem, s>>>(user_emb, item_T, ld, n_items, D, user_ids, n_eval, train_rowptr,
                                                      train_col, targets, T, K, topk_idx, topk_val, target_rank,
                                                      target_score);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

template <int DP>
static int launch_fullrank(const float* user_emb, const float* item_T, int64_t ld, int64_t n_items, int D,
                           const int64_t* user_ids, int64_t n_eval, const int64_t* train_rowptr,
                           const int32_t* train_col, const int32_t* targets, int T, int K, int32_t* topk_idx,
                           float* topk_val, int32_t* target_rank, float* target_score, cudaStream_t s) {
#define RECAD_FR_T(TMAX)                                                                                             \
  return launch_fullrank_t<DP, TMAX>(user_emb, item_T, ld, n_items, D, user_ids, n_eval, train_rowptr, train_col, targets, \
                                     T, K, topk_idx, topk_val, target_rank, target_score, s);
  if (T == 0) RECAD_FR_T(0)
  if (T == 1) RECAD_FR_T(1)
  if (T <= 4) RECAD_FR_T(4)
  RECAD_FR_T(8)
#undef RECAD_FR_T
}

}  // namespace recad

using namespace recad;

// ------------------------------------------------------------------------------------------
// Candidate users of the evaluation (normal.py:133-143): train users with NO target item in their train row and at
// least one candidate item left, ascending.  flags -> exclusive scan -> compaction; train rows are sorted, so a target
// is looked up by binary search.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eligible_flag_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_users, int64_t n_items,
                     const int32_t* __restrict__ targets, int n_targets, const uint8_t* __restrict__ is_key, uint32_t* __restrict__ flag) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_users) return;
  const int64_t lo = rowptr[u], hi = rowptr[u + 1];
  bool ok = is_key ? is_key[u] != 0 : hi > lo;                 // every KEY of train_dict is evaluated, even with an empty list
  ok = ok && (hi - lo) < n_items;                              // no candidate left: skipped (normal.py:63-64)
  for (int t = 0; ok && t < n_targets; ++t) {
    const int32_t x = targets[t];
    int64_t a = lo, b = hi;
    while (a < b) {
      const int64_t m = (a + b) >> 1;
      if (col[m] < x) a = m + 1; else b = m;
    }
    if (a < hi && col[a] == x) ok = false;
  }
  flag[u] = ok ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
eligible_compact_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, int64_t n_users, int64_t* __restrict__ out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n_users && flag[u]) out[pos[u]] = u;
}

extern "C" {

int recad_transpose_items(const float* item_emb, int64_t n_items, int32_t D, float* item_T, int64_t ld, void* stream) {
  RECAD_REQUIRE(item_emb && item_T && n_items > 0 && D > 0 && ld >= n_items && ld % kTI == 0, RECAD_ERR_ARG,
                "transpose_items: ld must be a multiple of %d and >= n_items", kTI);
  dim3 grid((unsigned)((ld + 31) / 32), (unsigned)((D + 31) / 32));
  transpose_items_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(item_emb, n_items, D, item_T, ld);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

int recad_fullrank_eval(const float* user_emb, const float* item_T, int64_t ld, int64_t n_items, int32_t D,
                        const int64_t* user_ids, int64_t n_eval, const int64_t* train_rowptr,
                        const int32_t* train_col, const int32_t* targets, int32_t T, int32_t K, int32_t* topk_idx,
                        float* topk_val, int32_t* target_rank, float* target_score, void* stream) {
  cudaStream_t s = as_stream(stream);
  RECAD_REQUIRE(user_emb && item_T && user_ids && train_rowptr && topk_idx && topk_val, RECAD_ERR_ARG,
                "fullrank: null pointer");
  RECAD_REQUIRE(n_eval > 0 && n_items > 0 && ld >= n_items && ld % kTI == 0, RECAD_ERR_ARG,
                "fullrank: bad sizes (ld must be a multiple of %d)", kTI);
  RECAD_REQUIRE(T >= 0 && T <= kMaxT && (T == 0 || (targets && target_rank && target_score)), RECAD_ERR_ARG,
                "fullrank: 0 <= T <= %d targets", kMaxT);
  RECAD_REQUIRE(K >= 1 && K <= 128, RECAD_ERR_UNSUPPORTED, "fullrank: 1 <= K <= 128");
  RECAD_REQUIRE(D >= 1 && D <= 136, RECAD_ERR_UNSUPPORTED, "fullrank: D = %d not in [1, 136]", D);
#define RECAD_FR(DP)                                                                                              \
  if (D <= DP)                                                                                                    \
    return launch_fullrank<DP>(user_emb, item_T, ld, n_items, D, user_ids, n_eval, train_rowptr, train_col, targets, \
                               T, K, topk_idx, topk_val, target_rank, target_score, s);
  RECAD_FR(16) RECAD_FR(32) RECAD_FR(64) RECAD_FR(72) RECAD_FR(128) RECAD_FR(136)
#undef RECAD_FR
  return RECAD_ERR_UNSUPPORTED;
}

int recad_rank_from_scores(const float* scores, int64_t n_rows, int64_t n_items, const int64_t* user_ids,
                           const int64_t* train_rowptr, const int32_t* train_col, const int32_t* targets, int32_t T,
                           int32_t K, int32_t* topk_idx, float* topk_val, int32_t* target_rank, float* target_score,
                           void* stream) {
  RECAD_REQUIRE(scores && user_ids && train_rowptr && topk_idx && topk_val && n_rows > 0 && n_items > 0, RECAD_ERR_ARG,
                "rank_from_scores: bad argument");
  RECAD_REQUIRE(T >= 0 && T <= kMaxT && (T == 0 || (targets && target_rank && target_score)), RECAD_ERR_ARG,
                "rank_from_scores: 0 <= T <= %d targets", kMaxT);
  RECAD_REQUIRE(K >= 1 && K <= 1024, RECAD_ERR_UNSUPPORTED, "rank_from_scores: 1 <= K <= 1024");
  const size_t smem = (size_t)kRankWarps * K * 8;
  rank_from_scores_kernel<<<(unsigned)((n_rows + kRankWarps - 1) / kRankWarps), kRankWarps * 32, smem, as_stream(stream)>>>(
      scores, n_rows, n_items, user_ids, train_rowptr, train_col, targets, T, K, topk_idx, topk_val, target_rank,
      target_score);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

int recad_recall_ndcg(const int32_t* topk_idx, int64_t n_eval, int32_t K, const int64_t* user_ids,
                      const int64_t* gt_rowptr, const int32_t* gt_col, double* out, void* stream) {
  RECAD_REQUIRE(topk_idx && user_ids && gt_rowptr && out && n_eval > 0 && K > 0, RECAD_ERR_ARG, "recall_ndcg: bad argument");
  recall_ndcg_kernel<<<(unsigned)((n_eval + 255) / 256), 256, 0, as_stream(stream)>>>(topk_idx, n_eval, K, user_ids,
                                                                                       gt_rowptr, gt_col, out);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

int64_t recad_eligible_users_scratch_bytes(int64_t n_users) { return 2 * ((n_users + 3) / 4 * 4) * (int64_t)sizeof(uint32_t) + scan_scratch_bytes(n_users) + 64; }

int recad_eligible_users(const int64_t* train_rowptr, const int32_t* train_col, int64_t n_users, int64_t n_items,
                         const int32_t* targets, int32_t n_targets, const uint8_t* is_key, int64_t* users_out, int64_t* n_out,
                         void* scratch, int64_t scratch_bytes, void* stream) {
  RECAD_REQUIRE(train_rowptr && users_out && n_out && scratch && n_users > 0 && n_items > 0 && n_targets >= 0 &&
                    (n_targets == 0 || targets) && scratch_bytes >= recad_eligible_users_scratch_bytes(n_users),
                RECAD_ERR_ARG, "eligible_users: bad argument");
  cudaStream_t s = as_stream(stream);
  const int64_t n4 = (n_users + 3) / 4 * 4;
  uint32_t* flag = static_cast<uint32_t*>(scratch);
  uint32_t* pos = flag + n4;
  unsigned long long* total = reinterpret_cast<unsigned long long*>(pos + n4);
  void* scan_scratch = reinterpret_cast<char*>(total) + 64;
  const unsigned grid = (unsigned)((n_users + 255) / 256);
  eligible_flag_kernel<<<grid, 256, 0, s>>>(train_rowptr, train_col, n_users, n_items, targets, n_targets, is_key, flag);
  RECAD_LAUNCH_CHECK();
  int rc = exclusive_scan_u32(flag, pos, n_users, total, scan_scratch, s);
  if (rc) return rc;
  eligible_compact_kernel<<<grid, 256, 0, s>>>(flag, pos, n_users, users_out);
  RECAD_LAUNCH_CHECK();
  unsigned long long h = 0;
  RECAD_CUDA_CHECK(cudaMemcpyAsync(&h, total, sizeof(h), cudaMemcpyDeviceToHost, s));
  RECAD_CUDA_CHECK(cudaStreamSynchronize(s));
  *n_out = (int64_t)h;
  return RECAD_OK;
}

}  // extern "C"
