// Full-ranking evaluation on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
// (reference: recad/workflow/normal.py:57-93 / 111-160; the score tile is lightgcn.py:115-120's
//  U_b . I^T, never materialised.)
//
// Scores are fp32-accurate: every operand x is split into x_hi = tf32(x) and x_lo = tf32(x - x_hi), and a
// tile is three kind::tf32 MMAs accumulated in fp32 in TMEM:
//       S = A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T            (|error| ~ 2^-21 |a||b|, "3xTF32")
//
// One CTA = 128 users (one TMEM lane each) x ALL items, streamed as 64-item tiles.  The user operand never
// touches shared memory: each epilogue thread loads its user's row, splits it in registers and parks hi / lo in
// TMEM with tcgen05.st; the MMAs read A from TMEM (the ".ts" form).  That leaves 256 TMEM columns and ~85 KB of
// shared memory per CTA, so TWO CTAs -- eight epilogue warps -- share an SM and the tensor pipe:
//   warp 0    : TMA producer   -- item tiles (hi + lo, 32 KB) into a 2- or 3-stage ring
//   warp 1    : MMA issuer     -- one thread issues 24 tcgen05.mma (3 passes x 8 k-steps of 8) per tile into one
//                                 of 2 TMEM accumulators (128 lanes x 64 columns)
//   warps 2-5 : epilogue       -- tcgen05.ld the user's row (32 columns at a time) and, in registers, apply the
//                                 train-item mask, the target-rank counters and a threshold-filtered top-K
//                                 insertion (shared memory is touched only when a score beats the K-th best)
// TMEM columns: [acc0 | acc1 | A_hi | A_lo], 64 each.
// The producer / MMA / epilogue run concurrently on mbarrier pipelines (smem full/empty, TMEM full/empty).
// The targets' scores come from the SAME arithmetic: a first 16-column MMA over the gathered target rows.
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include "rank_epilogue.cuh"
#include "tc_common.cuh"

namespace recad {

constexpr int kTcM = 128;        // users per CTA = TMEM lanes
constexpr int kTcN = 64;         // items per tile = TMEM columns per accumulator
constexpr int kTcK = 64;         // padded embedding width
constexpr int kTcKB = kTcK / 32; // 128-byte k-blocks (32 fp32)
constexpr int kTcMaxStages = 3;
constexpr int kTcAcc = 2;
constexpr int kTcThreads = 192;
constexpr int kTcTgtN = 16;
constexpr int kTcMaxT = 8;
constexpr int kTcCols = 256;                                 // TMEM columns per CTA
constexpr int kTcColA = kTcAcc * kTcN;                       // first column of A_hi; A_lo follows at + kTcK
constexpr uint32_t kKbBytes = kTcN * 128;                    // one k-block of a 64-row item tile: 8 KB
constexpr uint32_t kOperandBytes = kTcKB * kKbBytes;         // 16 KB (hi or lo)
constexpr uint32_t kStageBytes = 2 * kOperandBytes;          // hi + lo: 32 KB
constexpr int kTcBars = 1 + 2 * kTcMaxStages + 2 * kTcAcc;

struct TcMaps {
  CUtensorMap b_hi, b_lo, t_hi, t_lo;
};

// ---------------------------------------------------------------------------------------------- the kernel
template <int TMAX>
__global__ void __launch_bounds__(kTcThreads, 2)
fullrank_tc_kernel(const __grid_constant__ TcMaps maps, const float* __restrict__ user_emb, int D, int64_t n_items,
                   const int64_t* __restrict__ user_ids, int64_t n_eval, const int64_t* __restrict__ train_rowptr,
                   const int32_t* __restrict__ train_col, const int32_t* __restrict__ targets, int T, int K,
                   const float* __restrict__ item_bias, int32_t* __restrict__ topk_idx, float* __restrict__ topk_val,
                   int32_t* __restrict__ target_rank, float* __restrict__ target_score, int stages, int dbg) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte-swizzled operand tiles need a 1024-byte aligned base (the launch reserves the slack)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // [stage s: B_hi B_lo]* | top-K values | top-K ids | barriers | tmem ptr
  float* topv = reinterpret_cast<float*>(smem + (size_t)stages * kStageBytes);
  int32_t* topi = reinterpret_cast<int32_t*>(topv + (size_t)(K | 1) * kTcM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(topi + (size_t)(K | 1) * kTcM);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kTcBars);
  const uint32_t sB = smem_u32(smem);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_a_full = bar0;
  auto bar_b_full = [&](int s) { return bar0 + 8 * (1 + s); };
  auto bar_b_empty = [&](int s) { return bar0 + 8 * (1 + kTcMaxStages + s); };
  auto bar_acc_full = [&](int a) { return bar0 + 8 * (1 + 2 * kTcMaxStages + a); };
  auto bar_acc_empty = [&](int a) { return bar0 + 8 * (1 + 2 * kTcMaxStages + kTcAcc + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (n_items + kTcN - 1) / kTcN;
  const int64_t row0 = (int64_t)blockIdx.x * kTcM;

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, kTcM);
    for (int s = 0; s < kTcMaxStages; ++s) { mbar_init(bar_b_full(s), 1); mbar_init(bar_b_empty(s), 1); }
    for (int a = 0; a < kTcAcc; ++a) { mbar_init(bar_acc_full(a), 1); mbar_init(bar_acc_empty(a), kTcM); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTcCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t use = 0;
      for (int64_t j = 0; j <= n_tiles; ++j) {      // j = 0 is the 16-row target tile, j >= 1 item tile j-1
        mbar_wait_relaxed(bar_b_empty(s), (use & 1) ^ 1);
        const uint32_t dst = sB + s * kStageBytes;
        if (j == 0) {
          mbar_expect_tx(bar_b_full(s), 2 * kTcKB * kTcTgtN * 128);
          for (int kb = 0; kb < kTcKB; ++kb) {
            tma_load_2d(dst + kb * kKbBytes, &maps.t_hi, kb * 32, 0, bar_b_full(s));
            tma_load_2d(dst + kOperandBytes + kb * kKbBytes, &maps.t_lo, kb * 32, 0, bar_b_full(s));
          }
        } else {
          mbar_expect_tx(bar_b_full(s), kStageBytes);
          const int r = (int)((j - 1) * kTcN);
          for (int kb = 0; kb < kTcKB; ++kb) {
            tma_load_2d(dst + kb * kKbBytes, &maps.b_hi, kb * 32, r, bar_b_full(s));
            tma_load_2d(dst + kOperandBytes + kb * kKbBytes, &maps.b_lo, kb * 32, r, bar_b_full(s));
          }
        }
        if (++s == stages) { s = 0; ++use; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(bar_a_full, 0);                     // the epilogue threads have parked A_hi / A_lo in TMEM
      int s = 0;
      uint32_t use = 0;
      for (int64_t j = 0; j <= n_tiles; ++j) {
        const int a = (int)(j % kTcAcc);
        mbar_wait(bar_b_full(s), use & 1);
        mbar_wait(bar_acc_empty(a), ((uint32_t)(j / kTcAcc) & 1) ^ 1);
        tc_fence_after();
        const uint32_t idesc = j == 0 ? umma_idesc(kTcM, kTcTgtN) : umma_idesc(kTcM, kTcN);
        const uint32_t d = tmem_base + a * kTcN;
        const uint32_t bst = sB + s * kStageBytes;
        uint32_t acc = 0;
        // (dbg 3: profiling aid, ranking epilogue + TMA only: only the first of a tile's 24 MMAs is issued, so the accumulator
        //  holds a real -- coarser -- score tile and the epilogue sees realistic data)
#pragma unroll
        for (int pass = 0; pass < (dbg == 3 && j >= 1 ? 1 : 3); ++pass) {      // hi*hi, lo*hi, hi*lo
          const uint32_t ta = tmem_base + kTcColA + (pass == 1 ? kTcK : 0);
          const uint32_t bo = bst + (pass == 2 ? kOperandBytes : 0);
#pragma unroll
          for (int kb = 0; kb < kTcKB; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (dbg == 3 && j >= 1 && (kb | k)) continue;
              tc_mma_tf32_ts(d, ta + kb * 32 + k * 8, umma_desc(bo + kb * kKbBytes + k * 32), idesc, acc);
              acc = 1;
            }
        }
        tc_commit(bar_b_empty(s));      // the smem stage is free once these MMAs have read it
        tc_commit(bar_acc_full(a));     // ... and the accumulator is complete
        if (++s == stages) { s = 0; ++use; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread = user row
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int64_t g = row0 + r;
    const bool active = g < n_eval;
    const int64_t uid = active ? user_ids[g] : 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    // the user operand: row -> registers -> (tf32 hi, tf32 lo) -> TMEM lane r
    {
      const float* urow = user_emb + uid * D;
#pragma unroll 1
      for (int c = 0; c < kTcK / 32; ++c) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int d = c * 32 + e;
          const float x = (active && d < D) ? __ldg(urow + d) : 0.f;
          float h, l;
          split_tf32(x, h, l);
          hi[e] = __float_as_uint(h);
          lo[e] = __float_as_uint(l);
        }
        tmem_st32(lane_addr + kTcColA + c * 32, hi);
        tmem_st32(lane_addr + kTcColA + kTcK + c * 32, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_a_full);
    }
    // this row's K best: contiguous, odd pitch between rows (bank-conflict free when every lane scans its own)
    const int KP = K | 1;
    float* myv = topv + (size_t)r * KP;
    int32_t* myi = topi + (size_t)r * KP;
    for (int k = 0; k < K; ++k) { myv[k] = -INFINITY; myi[k] = -1; }
    float tau = -INFINITY;
    int min_pos = 0;
    int64_t cur = 0, end = 0;
    if (active) { cur = train_rowptr[uid]; end = train_rowptr[uid + 1]; }
    RankState<TMAX> rs;
    rank_state_init(rs, T, targets, train_col, cur, end);

    // j = 0: the targets' scores, same arithmetic as every other score
    mbar_wait(bar_acc_full(0), 0);
    tc_fence_after();
    {
      uint32_t v[16];
      tmem_ld16(lane_addr, v);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < TMAX; ++t)
        if (t < T) rank_state_set_score(rs, t, __uint_as_float(v[t]) + (item_bias ? __ldg(item_bias + rs.tg[t]) : 0.f));
    }
    tc_fence_before();
    mbar_arrive(bar_acc_empty(0));

    // one 32-score chunk, in registers: bias, -inf for train / padding / idle, rank counters, top-K
    auto chunk = [&](const uint32_t (&v)[32], int base, uint32_t train_mask) {
      const int lim = min(32, (int)n_items - base);
      if (lim <= 0) return;                          // warp-uniform
      float s[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) s[e] = __uint_as_float(v[e]);
      if (item_bias) {                               // uniform addresses: one broadcast load per item
#pragma unroll
        for (int e = 0; e < 32; ++e) s[e] += (e < lim) ? __ldg(item_bias + base + e) : 0.f;
      }
      const uint32_t drop = train_mask | (lim < 32 ? ~0u << lim : 0u) | (active ? 0u : ~0u);
      if (drop) {
#pragma unroll
        for (int e = 0; e < 32; ++e) s[e] = ((drop >> e) & 1u) ? -INFINITY : s[e];
      }
      rank_topk_chunk<32, TMAX>(s, base, T, rs, tau, myv, myi, K, 1, 0, min_pos);
    };

    int next_c = cur < end ? train_col[cur] : INT_MAX;   // the user's next train item, kept in a register
    for (int64_t j = 1; j <= n_tiles; ++j) {
      const int a = (int)(j % kTcAcc);
      const int j0 = (int)(j - 1) * kTcN;
      uint32_t m0 = 0u, m1 = 0u;                     // train-item mask of this tile (sorted list, walking pointer)
      while (next_c < j0 + kTcN) {
        const int o = next_c - j0;
        if (o < 32) m0 |= 1u << o; else m1 |= 1u << (o - 32);
        ++cur;
        next_c = cur < end ? train_col[cur] : INT_MAX;
      }
      mbar_wait(bar_acc_full(a), (uint32_t)(j / kTcAcc) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      if (dbg != 2) {                                // (2: profiling aid, MMA/TMA pipeline only)
        tmem_ld32(lane_addr + a * kTcN, v0);
        tmem_ld32(lane_addr + a * kTcN + 32, v1);
        tmem_ld_wait();
      }
      // the tile now lives in registers: hand the accumulator back BEFORE ranking it, so the MMAs of tile j + 2
      // run under this tile's epilogue
      tc_fence_before();
      mbar_arrive(bar_acc_empty(a));
      if (dbg == 2) continue;
      if (dbg == 1) { if (v0[0] == 0x7fc00001u && v1[0] == 0x7fc00001u) tau = 1.f; continue; }   // + TMEM loads
      chunk(v0, j0, m0);
      chunk(v1, j0 + 32, m1);
    }
    if (active) {
      topk_finalize(myv, myi, K, 1, 0);
      for (int k = 0; k < K; ++k) { topk_idx[g * K + k] = myi[k]; topk_val[g * K + k] = myv[k]; }
#pragma unroll
      for (int t = 0; t < TMAX; ++t)
        if (t < T) {
          target_rank[g * T + t] = ((rs.in_train >> t) & 1u) ? -1 : rs.rk[t];
          target_score[g * T + t] = rs.st[t];
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcCols) : "memory");
  }
}

// x -> (tf32(x), tf32(x - tf32(x))), rows gathered through idx, zero padded to [n_pad, 64]
__global__ void split_tf32_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, const int32_t* __restrict__ idx32,
                                  int64_t n_rows, int64_t n_pad, int D, float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_pad * kTcK) return;
  const int64_t r = e / kTcK;
  const int d = (int)(e % kTcK);
  float x = 0.f;
  if (r < n_rows && d < D) {
    const int64_t sr = idx ? idx[r] : (idx32 ? (int64_t)idx32[r] : r);
    x = src[sr * D + d];
  }
  float h, l;
  split_tf32(x, h, l);
  hi[e] = h;
  lo[e] = l;
}

static int make_map(CUtensorMap* m, const float* base, int64_t rows, int box_rows) {
  return make_tensor_map_f32(m, base, rows, kTcK, box_rows);
}

int make_tensor_map_f32(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int box_rows) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    RECAD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    RECAD_REQUIRE(p && q == cudaDriverEntryPointSuccess, RECAD_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RECAD_REQUIRE(r == CUDA_SUCCESS, RECAD_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return RECAD_OK;
}

}  // namespace recad

using namespace recad;

extern "C" {

int64_t recad_fullrank_tc_scratch_floats(int64_t n_eval, int64_t n_items) {
  (void)n_eval;   // the user operand is split in registers and lives in TMEM
  const int64_t ip = (n_items + kTcN - 1) / kTcN * kTcN;
  return 2 * kTcK * (ip + kTcTgtN) + 64;
}

int recad_fullrank_eval_tc(const float* user_emb, const float* item_emb, int64_t n_items, int32_t D,
                           const int64_t* user_ids, int64_t n_eval, const int64_t* train_rowptr,
                           const int32_t* train_col, const int32_t* targets, int32_t T, int32_t K,
                           const float* item_bias, int32_t* topk_idx, float* topk_val, int32_t* target_rank,
                           float* target_score, float* scratch, int64_t scratch_floats, void* stream) {
  cudaStream_t s = as_stream(stream);
  RECAD_REQUIRE(user_emb && item_emb && user_ids && train_rowptr && topk_idx && topk_val && scratch, RECAD_ERR_ARG,
                "fullrank_tc: null pointer");
  RECAD_REQUIRE(n_eval > 0 && n_items > 0 && n_items < (int64_t(1) << 31) - kTcN, RECAD_ERR_ARG, "fullrank_tc: bad sizes");
  RECAD_REQUIRE(D >= 1 && D <= kTcK, RECAD_ERR_UNSUPPORTED, "fullrank_tc: D = %d > %d (use recad_fullrank_eval)", D, kTcK);
  RECAD_REQUIRE(K >= 1 && K <= 32, RECAD_ERR_UNSUPPORTED, "fullrank_tc: K = %d > 32 (use recad_fullrank_eval)", K);
  RECAD_REQUIRE(T >= 0 && T <= kTcMaxT && (T == 0 || (targets && target_rank && target_score)), RECAD_ERR_ARG,
                "fullrank_tc: 0 <= T <= %d targets", kTcMaxT);
  RECAD_REQUIRE(scratch_floats >= recad_fullrank_tc_scratch_floats(n_eval, n_items), RECAD_ERR_SCRATCH,
                "fullrank_tc: scratch too small");
  RECAD_REQUIRE(((uintptr_t)scratch & 255) == 0, RECAD_ERR_ARG, "fullrank_tc: scratch must be 256-byte aligned");
  const int64_t np = (n_eval + kTcM - 1) / kTcM * kTcM, ip = (n_items + kTcN - 1) / kTcN * kTcN;
  float* b_hi = scratch;
  float* b_lo = b_hi + ip * kTcK;
  float* t_hi = b_lo + ip * kTcK;
  float* t_lo = t_hi + kTcTgtN * kTcK;
  const int TB = 256;
  split_tf32_kernel<<<(unsigned)((ip * kTcK + TB - 1) / TB), TB, 0, s>>>(item_emb, nullptr, nullptr, n_items, ip, D, b_hi, b_lo);
  RECAD_LAUNCH_CHECK();
  split_tf32_kernel<<<(unsigned)((kTcTgtN * kTcK + TB - 1) / TB), TB, 0, s>>>(item_emb, nullptr, targets, T, kTcTgtN, D, t_hi, t_lo);
  RECAD_LAUNCH_CHECK();
  TcMaps maps;
  int rc;
  if ((rc = make_map(&maps.b_hi, b_hi, ip, kTcN))) return rc;
  if ((rc = make_map(&maps.b_lo, b_lo, ip, kTcN))) return rc;
  if ((rc = make_map(&maps.t_hi, t_hi, kTcTgtN, kTcTgtN))) return rc;
  if ((rc = make_map(&maps.t_lo, t_lo, kTcTgtN, kTcTgtN))) return rc;
  // two CTAs per SM: take the third ring stage only when both still fit (1 KB per CTA is reserved by the driver)
  const size_t fixed = 1024 + (size_t)(K | 1) * kTcM * 8 + kTcBars * 8 + 16;
  static const int force_stages = getenv("RECAD_TC_STAGES") ? atoi(getenv("RECAD_TC_STAGES")) : 0;
  int stages = 2 * (fixed + 3 * kStageBytes + 1024) <= 228 * 1024 ? 3 : 2;
  if (force_stages >= 1 && force_stages <= kTcMaxStages) stages = force_stages;
  const size_t smem_bytes = fixed + (size_t)stages * kStageBytes;
  RECAD_REQUIRE(smem_bytes <= 227 * 1024, RECAD_ERR_UNSUPPORTED, "fullrank_tc: shared memory %zu B", smem_bytes);
  static const int dbg = getenv("RECAD_TC_DEBUG") ? atoi(getenv("RECAD_TC_DEBUG")) : 0;
#define RECAD_TC_LAUNCH(TMAX)                                                                                          \
  {                                                                                                                    \
    auto kern = fullrank_tc_kernel<TMAX>;                                                                              \
    RECAD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));             \
    RECAD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,                        \
                                          cudaSharedmemCarveoutMaxShared));                                            \
    kern<<<(unsigned)(np / kTcM), kTcThreads, smem_bytes, s>>>(                                                        \
        maps, user_emb, D, n_items, user_ids, n_eval, train_rowptr, train_col, targets, T, K, item_bias, topk_idx,     \
        topk_val, target_rank, target_score, stages, dbg);                                                             \
  }
  if (T == 0) RECAD_TC_LAUNCH(0)
  else if (T == 1) RECAD_TC_LAUNCH(1)
  else if (T == 2) RECAD_TC_LAUNCH(2)
  else if (T <= 4) RECAD_TC_LAUNCH(4)
  else RECAD_TC_LAUNCH(8)
#undef RECAD_TC_LAUNCH
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

}  // extern "C"
