// fp32-accurate GEMM on the 5th-generation tensor cores for the NCF tower
// (reference: the nn.Linear layers of recad/model/victim/ncf.py:41-53, forward and backward).
//
//   C[M, N] = A[M, K] . B[N, K]^T  (+ bias[N]) (ReLU),   A and B row-major with K contiguous ("TN", both K-major)
//
// Operands arrive already split for the 3xTF32 scheme (x = hi + lo, both tf32): three tcgen05.mma kind::tf32
// per k-step (hi.hi + lo.hi + hi.lo) accumulate in fp32 in TMEM, so the result carries ~fp32 accuracy and the
// step stays within the 1e-4 parity bar of the fp32 reference.  One CTA computes one 128 x BN output tile:
//   warp 0: TMA producer (3-stage ring of 32-wide k-blocks of A_hi, A_lo, B_hi, B_lo, 128-byte swizzle)
//   warp 1: MMA issuer   (one thread; 12 MMAs per k-block)
//   warps 2-5: epilogue  (tcgen05.ld, bias / ReLU, fp32 store and -- optionally -- the hi/lo split of the result,
//                         so the next layer's A operand needs no extra pass)
// Transposed operands (weight^T for dgrad, activation^T for wgrad) are produced by small split/transpose kernels;
// TMA zero-fills out-of-range rows / k, so M, N need no padding and K only to a multiple of 4 floats.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <tuple>

#include "tc_common.cuh"

namespace recad {

constexpr int kGM = 128;
constexpr int kGStages = 3;
constexpr int kGThreads = 192;

struct GemmMaps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

template <int BN>
__global__ void __launch_bounds__(kGThreads, 1)
gemm_tc_kernel(const __grid_constant__ GemmMaps maps, int M, int N, int n_kb, float* __restrict__ Cm, int ldc,
               const float* __restrict__ bias, int relu, float* __restrict__ out_hi, float* __restrict__ out_lo, int ld_split) {
  constexpr uint32_t kABytes = kGM * 128, kBBytes = BN * 128;          // one k-block of one operand half
  constexpr uint32_t kStage = 2 * kABytes + 2 * kBBytes;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGStages * kStage);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGStages + 1);
  const uint32_t s0 = smem_u32(smem), bar0 = smem_u32(bars);
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (kGStages + s); };
  const uint32_t bar_acc = bar0 + 8 * (2 * kGStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kGM, n0 = blockIdx.x * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb % kGStages;
        mbar_wait(bar_empty(s), ((uint32_t)(kb / kGStages) & 1) ^ 1);
        const uint32_t dst = s0 + s * kStage;
        mbar_expect_tx(bar_full(s), kStage);
        tma_load_2d(dst, &maps.a_hi, kb * 32, m0, bar_full(s));
        tma_load_2d(dst + kABytes, &maps.a_lo, kb * 32, m0, bar_full(s));
        tma_load_2d(dst + 2 * kABytes, &maps.b_hi, kb * 32, n0, bar_full(s));
        tma_load_2d(dst + 2 * kABytes + kBBytes, &maps.b_lo, kb * 32, n0, bar_full(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(kGM, BN);
      uint32_t acc = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb % kGStages;
        mbar_wait(bar_full(s), (uint32_t)(kb / kGStages) & 1);
        tc_fence_after();
        const uint32_t a_hi = s0 + s * kStage, a_lo = a_hi + kABytes, b_hi = a_hi + 2 * kABytes, b_lo = b_hi + kBBytes;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {        // hi*hi, lo*hi, hi*lo
          const uint32_t ao = pass == 1 ? a_lo : a_hi, bo = pass == 2 ? b_lo : b_hi;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc_mma_tf32(tmem_base, umma_desc(ao + k * 32), umma_desc(bo + k * 32), idesc, acc);
            acc = 1;
          }
        }
        tc_commit(bar_empty(s));
      }
      tc_commit(bar_acc);
    }
  } else {
    const int q = warp & 3;
    const int gm = m0 + q * 32 + lane;
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool vec_ok = ((ldc & 3) == 0) && ((ld_split & 3) == 0) && (((uintptr_t)Cm | (uintptr_t)out_hi | (uintptr_t)out_lo) & 15) == 0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(lane_addr + c * 32, v);
      tmem_ld_wait();
      const int nb = n0 + c * 32;
      if (gm < M && nb < N) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float x[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            x[t] = __uint_as_float(v[e + t]);
            if (bias && nb + e + t < N) x[t] += __ldg(bias + nb + e + t);
            if (relu) x[t] = fmaxf(x[t], 0.f);
          }
          if (vec_ok && nb + e + 3 < N) {
            if (Cm) *reinterpret_cast<float4*>(Cm + (int64_t)gm * ldc + nb + e) = make_float4(x[0], x[1], x[2], x[3]);
            if (out_hi) {
              float h[4], l[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) split_tf32(x[t], h[t], l[t]);
              *reinterpret_cast<float4*>(out_hi + (int64_t)gm * ld_split + nb + e) = make_float4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<float4*>(out_lo + (int64_t)gm * ld_split + nb + e) = make_float4(l[0], l[1], l[2], l[3]);
            }
          } else {
            for (int t = 0; t < 4; ++t)
              if (nb + e + t < N) {
                if (Cm) Cm[(int64_t)gm * ldc + nb + e + t] = x[t];
                if (out_hi) {
                  float h, l;
                  split_tf32(x[t], h, l);
                  out_hi[(int64_t)gm * ld_split + nb + e + t] = h;
                  out_lo[(int64_t)gm * ld_split + nb + e + t] = l;
                }
              }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

// src [R, C] (row stride ld) -> hi / lo [R, ldo] ; columns C..ldo are zeroed
__global__ void split_rows_kernel(const float* __restrict__ src, int R, int Cc, int ld, float* __restrict__ hi,
                                  float* __restrict__ lo, int ldo) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)R * ldo) return;
  const int r = (int)(e / ldo), c = (int)(e % ldo);
  float h = 0.f, l = 0.f;
  if (c < Cc) split_tf32(src[(int64_t)r * ld + c], h, l);
  hi[e] = h;
  lo[e] = l;
}

// src [R, C] -> hi / lo [C, ldo] = src^T ; columns R..ldo are zeroed.  32 x 32 shared-memory tiles.
__global__ void split_transpose_kernel(const float* __restrict__ src, int R, int Cc, int ld, float* __restrict__ hi,
                                       float* __restrict__ lo, int ldo) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < R && c < Cc) ? src[(int64_t)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;       // output row c, output column r
    if (c < Cc && r < ldo) {
      float h, l;
      split_tf32(t[threadIdx.x][i], h, l);
      hi[(int64_t)c * ldo + r] = h;
      lo[(int64_t)c * ldo + r] = l;
    }
  }
}

// The three operand splits of one backward layer in ONE launch (they were three of a layer's six launches; the NCF step is
// launch bound): blockIdx.z = 0: dz [R, C0] -> dz^T hi / lo [C0, ldo] AND the row split dz hi / lo [R, C0];
// blockIdx.z = 1: h [R, C1] -> h^T hi / lo [C1, ldo].  Columns R..ldo of the transposes are zeroed.
__global__ void split_dz_h_kernel(const float* __restrict__ dz, int R, int C0, const float* __restrict__ h, int C1,
                                  float* __restrict__ dzt_hi, float* __restrict__ dzt_lo, float* __restrict__ ht_hi,
                                  float* __restrict__ ht_lo, int ldo, float* __restrict__ dz_hi, float* __restrict__ dz_lo) {
  __shared__ float t[32][33];
  const bool second = blockIdx.z == 1;
  const float* __restrict__ src = second ? h : dz;
  const int Cc = second ? C1 : C0;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  if (c0 >= Cc) return;                                     // (whole block: the grid is sized for the wider matrix)
  float* __restrict__ hi = second ? ht_hi : dzt_hi;
  float* __restrict__ lo = second ? ht_lo : dzt_lo;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    const bool in = r < R && c < Cc;
    const float v = in ? src[(int64_t)r * Cc + c] : 0.f;
    t[i][threadIdx.x] = v;
    if (!second && in) {
      float a, b;
      split_tf32(v, a, b);
      dz_hi[(int64_t)r * Cc + c] = a;
      dz_lo[(int64_t)r * Cc + c] = b;
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;             // output row c, output column r
    if (c < Cc && r < ldo) {
      float a, b;
      split_tf32(t[threadIdx.x][i], a, b);
      hi[(int64_t)c * ldo + r] = a;
      lo[(int64_t)c * ldo + r] = b;
    }
  }
}

int tc_split_dz_h(const float* dz, int R, int C0, const float* h, int C1, float* dzt_hi, float* dzt_lo, float* ht_hi, float* ht_lo,
                  int ldo, float* dz_hi, float* dz_lo, cudaStream_t s) {
  if (R == 0 || C0 == 0 || C1 == 0) return RECAD_OK;
  dim3 grid((std::max(C0, C1) + 31) / 32, (ldo + 31) / 32, 2);
  split_dz_h_kernel<<<grid, dim3(32, 8), 0, s>>>(dz, R, C0, h, C1, dzt_hi, dzt_lo, ht_hi, ht_lo, ldo, dz_hi, dz_lo);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

int tc_split_rows(const float* src, int R, int Cc, int ld, float* hi, float* lo, int ldo, cudaStream_t s) {
  const int64_t n = (int64_t)R * ldo;
  if (n == 0) return RECAD_OK;
  split_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, R, Cc, ld, hi, lo, ldo);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

int tc_split_transpose(const float* src, int R, int Cc, int ld, float* hi, float* lo, int ldo, cudaStream_t s) {
  if (R == 0 || Cc == 0) return RECAD_OK;
  dim3 grid((Cc + 31) / 32, (ldo + 31) / 32);
  split_transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(src, R, Cc, ld, hi, lo, ldo);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

// tensor maps are a pure function of (pointer, shape, box): the NCF buffers never move, so encode once
static const CUtensorMap* cached_map(const float* base, int64_t rows, int64_t cols, int box_rows, int* rc) {
  static thread_local std::map<std::tuple<const float*, int64_t, int64_t, int>, CUtensorMap> cache;
  auto key = std::make_tuple(base, rows, cols, box_rows);
  auto it = cache.find(key);
  if (it == cache.end()) {
    CUtensorMap m;
    *rc = make_tensor_map_f32(&m, base, rows, cols, box_rows);
    if (*rc) return nullptr;
    if (cache.size() > 4096) cache.clear();
    it = cache.emplace(key, m).first;
  }
  *rc = RECAD_OK;
  return &it->second;
}

// C[M, N] = A[M, K] B[N, K]^T (+ bias) (relu); A_*: [M, lda], B_*: [N, ldb] with lda, ldb >= K, multiples of 4 and the
// columns K.. zero (so the tail of the last 32-wide k-block contributes nothing).
int gemm_tc(const float* A_hi, const float* A_lo, int M, int lda, const float* B_hi, const float* B_lo, int N, int ldb, int K,
            float* Cm, int ldc, const float* bias, bool relu, float* out_hi, float* out_lo, int ld_split, cudaStream_t s) {
  RECAD_REQUIRE(M > 0 && N > 0 && K > 0 && lda % 4 == 0 && ldb % 4 == 0 && lda >= K && ldb >= K, RECAD_ERR_ARG,
                "gemm_tc: bad shape M=%d N=%d K=%d lda=%d ldb=%d", M, N, K, lda, ldb);
  // 128-wide tiles only when they fill the chip: at the reference's batch of 1024 a tower GEMM has 8-32 of them, and a CTA's
  // time is set by the operand bytes it streams per k-block (64 KB at BN = 128, 48 KB at BN = 64) -- the narrow tile puts
  // twice as many SMs to work on shorter k-steps
  static const int bn_env = getenv("RECAD_GEMM_BN") ? atoi(getenv("RECAD_GEMM_BN")) : 0;
  const int64_t tiles128 = (int64_t)((N + 127) / 128) * ((M + kGM - 1) / kGM);
  const int BN = bn_env ? (bn_env >= 128 && N > 64 ? 128 : 64) : (N > 64 && tiles128 >= sm_count() ? 128 : 64);
  int rc;
  GemmMaps maps;
  const CUtensorMap* p;
  if (!(p = cached_map(A_hi, M, lda, kGM, &rc))) return rc;
  maps.a_hi = *p;
  if (!(p = cached_map(A_lo, M, lda, kGM, &rc))) return rc;
  maps.a_lo = *p;
  if (!(p = cached_map(B_hi, N, ldb, BN, &rc))) return rc;
  maps.b_hi = *p;
  if (!(p = cached_map(B_lo, N, ldb, BN, &rc))) return rc;
  maps.b_lo = *p;
  const int n_kb = (K + 31) / 32;
  dim3 grid((N + BN - 1) / BN, (M + kGM - 1) / kGM);
  if (BN == 128) {
    const size_t smem = 1024 + kGStages * (2 * kGM * 128 + 2 * 128 * 128) + 128;
    static bool attr = false;
    if (!attr) { RECAD_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    gemm_tc_kernel<128><<<grid, kGThreads, smem, s>>>(maps, M, N, n_kb, Cm, ldc, bias, relu ? 1 : 0, out_hi, out_lo, ld_split);
  } else {
    const size_t smem = 1024 + kGStages * (2 * kGM * 128 + 2 * 64 * 128) + 128;
    static bool attr = false;
    if (!attr) { RECAD_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    gemm_tc_kernel<64><<<grid, kGThreads, smem, s>>>(maps, M, N, n_kb, Cm, ldc, bias, relu ? 1 : 0, out_hi, out_lo, ld_split);
  }
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

}  // namespace recad

using namespace recad;

// Stand-alone entry (tests, other callers): C = A B^T (+ bias) (relu) with the operand split done here.
// scratch [dev] float[2 * (M + N) * ((K + 3) / 4 * 4)].
extern "C" int recad_gemm_tn_tf32x3(const float* A, const float* B, int32_t M, int32_t N, int32_t K, const float* bias,
                                    int32_t relu, float* Cm, float* scratch, void* stream) {
  RECAD_REQUIRE(A && B && Cm && scratch && M > 0 && N > 0 && K > 0, RECAD_ERR_ARG, "gemm_tn_tf32x3: bad argument");
  cudaStream_t s = as_stream(stream);
  const int ld = (K + 3) / 4 * 4;
  float* a_hi = scratch;
  float* a_lo = a_hi + (int64_t)M * ld;
  float* b_hi = a_lo + (int64_t)M * ld;
  float* b_lo = b_hi + (int64_t)N * ld;
  int rc;
  if ((rc = tc_split_rows(A, M, K, K, a_hi, a_lo, ld, s))) return rc;
  if ((rc = tc_split_rows(B, N, K, K, b_hi, b_lo, ld, s))) return rc;
  return gemm_tc(a_hi, a_lo, M, ld, b_hi, b_lo, N, ld, K, Cm, N, bias, relu != 0, nullptr, nullptr, 0, s);
}
