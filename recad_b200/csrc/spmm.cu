// CSR SpMM fused with the LightGCN layer mean:  Y = A X,  Z = alpha * (C + A X).
// (reference: recad/model/victim/lightgcn.py:99-111 `torch.sparse.mm` + stack + mean; the
// autograd backward of the same lines is this kernel again because A_hat is symmetric.)
//
// HBM-bound gather: per stored entry 8 B of (col, val) streamed + one D*4-byte row of X gathered.
// Work unit = one warp per SEGMENT of a row (plan built by recad_spmm_plan): short rows are one
// segment and write their result directly through the fused epilogue; long rows (popular items)
// are cut into <= seg_len pieces whose partial sums are combined in fixed order by a second tiny
// kernel, so the result is deterministic and no warp ever walks a 10^5-entry row alone.
//
// Inside a warp: the segment's (col, val) pairs are staged 32 at a time with one coalesced load
// per lane and broadcast with shuffles; a row of X is covered by D/4 lanes with 128-bit loads, so
// a warp load instruction fetches 32/(D/4) neighbour rows at once and UNR of them are in flight.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace recad {

constexpr int kSpmmWarps = 8;  // warps per CTA

template <int D>
struct RowLanes {
  static constexpr int LPR = D / 4;      // lanes per row (float4 each)
  static constexpr int NPL = 32 / LPR;   // neighbour rows per warp-wide load
};

// Peer-memory epilogue of the sharded path (recad_spmm_scatter): row i of the product is not stored in a local
// Y but sent to the rank that OWNS item i -- dst[i / slice] is that rank's staging block for this sender, mapped
// into this process over NVLink (torch symmetric memory), so the partial sums travel while the SpMM still runs.
constexpr int kMaxPeers = 16;
struct RowScatter {
  float* dst[kMaxPeers];
  int32_t slice;
};

template <int D, bool kScatter = false>
__device__ __forceinline__ void spmm_epilogue(int64_t row, int slot, int l, float4 y, float* Y, const float* C,
                                              float* Z, float alpha, float* partials, const RowScatter* sc = nullptr) {  // Z may alias C
  constexpr int LPR = D / 4;
  if (slot >= 0) {
    reinterpret_cast<float4*>(partials)[(int64_t)slot * LPR + l] = y;
    return;
  }
  if (kScatter) {
    const int owner = (int)(row / sc->slice);
    reinterpret_cast<float4*>(sc->dst[owner])[(row - (int64_t)owner * sc->slice) * LPR + l] = y;
    return;
  }
  const int64_t o = row * LPR + l;
  if (Y) reinterpret_cast<float4*>(Y)[o] = y;
  if (Z) {
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (C) c = reinterpret_cast<const float4*>(C)[o];
    float4 z;
    z.x = alpha * (c.x + y.x);
    z.y = alpha * (c.y + y.y);
    z.z = alpha * (c.z + y.z);
    z.w = alpha * (c.w + y.w);
    reinterpret_cast<float4*>(Z)[o] = z;
  }
}

template <int D, int UNR, int MINB, bool kStreamX, bool kScatter = false>
__global__ void __launch_bounds__(kSpmmWarps * 32, MINB)
spmm_seg_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                int64_t n_seg, int32_t seg_len, const int32_t* __restrict__ seg_row,
                const int64_t* __restrict__ seg_lo, const int32_t* __restrict__ seg_slot,
                const float* __restrict__ X, float* Y, const float* C, float* Z, float alpha, float* partials,
                const __grid_constant__ RowScatter sc) {
  constexpr int LPR = RowLanes<D>::LPR, NPL = RowLanes<D>::NPL;
  static_assert(32 % (NPL * UNR) == 0, "unroll must divide the staged chunk");
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;  // which neighbour of the NPL fetched together
  const int l = lane % LPR;    // which float4 of the row
  // persistent launch: the grid is sized to the resident-CTA capacity of the chip and every warp
  // strides over the segment list, so occupancy stays full until the list is exhausted (a warp
  // per segment leaves a CTA's slots idle while its longest row finishes)
  const int64_t n_warps = (int64_t)gridDim.x * kSpmmWarps;
  for (int64_t seg = (int64_t)blockIdx.x * kSpmmWarps + (threadIdx.x >> 5); seg < n_seg; seg += n_warps) {
  const int64_t row = seg_row[seg];
  const int64_t lo = seg_lo[seg];
  const int64_t hi = min(lo + (int64_t)seg_len, rowptr[row + 1]);
  const int slot = seg_slot[seg];
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // software pipeline on the (col, val) stream: the next 32 pairs are requested before the
  // current 32 are consumed
  int c_next = 0;
  float v_next = 0.f;
  if (lo + lane < hi) {
    c_next = ld_stream(colidx + lo + lane);
    v_next = ld_stream(vals + lo + lane);
  }
  for (int64_t base = lo; base < hi; base += 32) {
    const int c = c_next;
    const float v = v_next;
    const int64_t nb = base + 32 + lane;
    if (nb < hi) {
      c_next = ld_stream(colidx + nb);
      v_next = ld_stream(vals + nb);
    }
    const int cnt = (int)min((int64_t)32, hi - base);
    for (int j = 0; j < cnt; j += NPL * UNR) {
      float4 x[UNR];
      float w[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int k = j + u * NPL + sub;
        const int cc = __shfl_sync(kFull, c, k & 31);
        w[u] = __shfl_sync(kFull, v, k & 31);
        if (k < cnt) {
          const uint32_t off = (uint32_t)cc * (uint32_t)LPR + (uint32_t)l;   // float4 units; checked < 2^32 on the host
          x[u] = kStreamX ? ld_stream4(X4 + off) : __ldg(X4 + off);
        } else {
          x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          w[u] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        acc.x = fmaf(w[u], x[u].x, acc.x);
        acc.y = fmaf(w[u], x[u].y, acc.y);
        acc.z = fmaf(w[u], x[u].z, acc.z);
        acc.w = fmaf(w[u], x[u].w, acc.w);
      }
    }
  }
#pragma unroll
  for (int o = 16; o >= LPR; o >>= 1) {
    acc.x += __shfl_xor_sync(kFull, acc.x, o);
    acc.y += __shfl_xor_sync(kFull, acc.y, o);
    acc.z += __shfl_xor_sync(kFull, acc.z, o);
    acc.w += __shfl_xor_sync(kFull, acc.w, o);
  }
  if (sub == 0) spmm_epilogue<D, kScatter>(row, slot, l, acc, Y, C, Z, alpha, partials, &sc);
  }
}

// rows with more than one segment: sum their partial slots in slot order, then the same epilogue
template <int D, bool kScatter = false>
__global__ void __launch_bounds__(kSpmmWarps * 32)
spmm_fixup_kernel(int64_t n_mrow, const int32_t* __restrict__ mrow, const int32_t* __restrict__ mrow_lo,
                  const float* __restrict__ partials, float* Y, const float* C, float* Z, float alpha,
                  const __grid_constant__ RowScatter sc) {
  constexpr int LPR = RowLanes<D>::LPR, NPL = RowLanes<D>::NPL;
  const int64_t j = (int64_t)blockIdx.x * kSpmmWarps + (threadIdx.x >> 5);
  if (j >= n_mrow) return;
  const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
  const int s0 = mrow_lo[j], s1 = mrow_lo[j + 1];
  const float4* __restrict__ P4 = reinterpret_cast<const float4*>(partials);
  // a hot item row has thousands of slots: keep 4 loads in flight per lane (fixed association => deterministic)
  float4 a4[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) a4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  int s = s0 + sub;
  for (; s + 3 * NPL < s1; s += 4 * NPL) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 p = P4[(int64_t)(s + q * NPL) * LPR + l];
      a4[q].x += p.x; a4[q].y += p.y; a4[q].z += p.z; a4[q].w += p.w;
    }
  }
  for (; s < s1; s += NPL) {
    const float4 p = P4[(int64_t)s * LPR + l];
    a4[0].x += p.x; a4[0].y += p.y; a4[0].z += p.z; a4[0].w += p.w;
  }
  float4 acc = make_float4((a4[0].x + a4[1].x) + (a4[2].x + a4[3].x), (a4[0].y + a4[1].y) + (a4[2].y + a4[3].y),
                           (a4[0].z + a4[1].z) + (a4[2].z + a4[3].z), (a4[0].w + a4[1].w) + (a4[2].w + a4[3].w));
#pragma unroll
  for (int o = 16; o >= LPR; o >>= 1) {
    acc.x += __shfl_xor_sync(kFull, acc.x, o);
    acc.y += __shfl_xor_sync(kFull, acc.y, o);
    acc.z += __shfl_xor_sync(kFull, acc.z, o);
    acc.w += __shfl_xor_sync(kFull, acc.w, o);
  }
  if (sub == 0) spmm_epilogue<D, kScatter>(mrow[j], -1, l, acc, Y, C, Z, alpha, nullptr, &sc);
}

// any D that is a multiple of 4 (<= 1024): a lane owns float4 columns lane, lane+32, ...
constexpr int kGenericMaxVec = 8;
__global__ void __launch_bounds__(kSpmmWarps * 32)
spmm_generic_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                    const float* __restrict__ vals, int64_t n_seg, int32_t seg_len,
                    const int32_t* __restrict__ seg_row, const int64_t* __restrict__ seg_lo,
                    const int32_t* __restrict__ seg_slot, const float* __restrict__ X, float* Y, const float* C,
                    float* Z, float alpha, float* partials, int D) {
  const int64_t seg = (int64_t)blockIdx.x * kSpmmWarps + (threadIdx.x >> 5);
  if (seg >= n_seg) return;
  const int lane = threadIdx.x & 31;
  const int nvec = D / 4;
  const int64_t row = seg_row[seg], lo = seg_lo[seg];
  const int64_t hi = min(lo + (int64_t)seg_len, rowptr[row + 1]);
  const int slot = seg_slot[seg];
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);
  float4 acc[kGenericMaxVec];
#pragma unroll
  for (int q = 0; q < kGenericMaxVec; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t base = lo; base < hi; base += 32) {
    int c = 0;
    float v = 0.f;
    if (base + lane < hi) { c = colidx[base + lane]; v = vals[base + lane]; }
    const int cnt = (int)min((int64_t)32, hi - base);
    for (int k = 0; k < cnt; ++k) {
      const int cc = __shfl_sync(kFull, c, k);
      const float w = __shfl_sync(kFull, v, k);
#pragma unroll
      for (int q = 0; q < kGenericMaxVec; ++q) {
        const int col4 = lane + q * 32;
        if (col4 < nvec) {
          const float4 x = __ldg(X4 + (int64_t)cc * nvec + col4);
          acc[q].x = fmaf(w, x.x, acc[q].x); acc[q].y = fmaf(w, x.y, acc[q].y);
          acc[q].z = fmaf(w, x.z, acc[q].z); acc[q].w = fmaf(w, x.w, acc[q].w);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kGenericMaxVec; ++q) {
    const int col4 = lane + q * 32;
    if (col4 >= nvec) continue;
    const float4 y = acc[q];
    if (slot >= 0) { reinterpret_cast<float4*>(partials)[(int64_t)slot * nvec + col4] = y; continue; }
    const int64_t o = row * nvec + col4;
    if (Y) reinterpret_cast<float4*>(Y)[o] = y;
    if (Z) {
      float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (C) cv = reinterpret_cast<const float4*>(C)[o];
      reinterpret_cast<float4*>(Z)[o] = make_float4(alpha * (cv.x + y.x), alpha * (cv.y + y.y),
                                                    alpha * (cv.z + y.z), alpha * (cv.w + y.w));
    }
  }
}

__global__ void __launch_bounds__(kSpmmWarps * 32)
spmm_generic_fixup_kernel(int64_t n_mrow, const int32_t* __restrict__ mrow, const int32_t* __restrict__ mrow_lo,
                          const float* __restrict__ partials, float* Y, const float* C, float* Z, float alpha,
                          int D) {
  const int64_t j = (int64_t)blockIdx.x * kSpmmWarps + (threadIdx.x >> 5);
  if (j >= n_mrow) return;
  const int lane = threadIdx.x & 31, nvec = D / 4;
  const int64_t row = mrow[j];
  for (int col4 = lane; col4 < nvec; col4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = mrow_lo[j]; s < mrow_lo[j + 1]; ++s) {
      const float4 p = reinterpret_cast<const float4*>(partials)[(int64_t)s * nvec + col4];
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    const int64_t o = row * nvec + col4;
    if (Y) reinterpret_cast<float4*>(Y)[o] = acc;
    if (Z) {
      float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (C) cv = reinterpret_cast<const float4*>(C)[o];
      reinterpret_cast<float4*>(Z)[o] = make_float4(alpha * (cv.x + acc.x), alpha * (cv.y + acc.y),
                                                    alpha * (cv.z + acc.z), alpha * (cv.w + acc.w));
    }
  }
}

// tuning knob (RECAD_SPMM_VARIANT): bit0 = persistent grid, bits 1-2 = register cap (0: none, 1: 6 CTAs/SM,
// 2: 5 CTAs/SM, 3: 4 CTAs/SM), bit3 = deeper unroll.  The default is the variant measured fastest on B200 (profiles/).
static int spmm_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RECAD_SPMM_VARIANT");
    v = e ? atoi(e) : 5;
  }
  return v;
}

template <int D, int UNR, int MINB, bool kStreamX, bool kScatter = false>
static int launch_spmm_v(const recad_csr* A, const float* X, float* Y, const float* C, float* Z, float alpha,
                         bool persistent, cudaStream_t s, const RowScatter& sc = RowScatter{}) {
  auto kern = spmm_seg_kernel<D, UNR, MINB, kStreamX, kScatter>;
  int64_t grid = (A->n_seg + kSpmmWarps - 1) / kSpmmWarps;
  if (persistent) {
    static int per_sm = 0;
    if (!per_sm) RECAD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSpmmWarps * 32, 0));
    grid = std::min<int64_t>(grid, (int64_t)sm_count() * std::max(per_sm, 1));
  }
  kern<<<(unsigned)grid, kSpmmWarps * 32, 0, s>>>(A->rowptr, A->colidx, A->vals, A->n_seg, A->seg_len, A->seg_row,
                                                 A->seg_lo, A->seg_slot, X, Y, C, Z, alpha, A->partials, sc);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

// Y = A X with every finished row sent to its owner (see RowScatter)
template <int D, int UNR>
static int launch_spmm_scatter(const recad_csr* A, const float* X, const RowScatter& sc, cudaStream_t s) {
  int rc = launch_spmm_v<D, UNR, 5, false, true>(A, X, nullptr, nullptr, nullptr, 1.f, true, s, sc);
  if (rc) return rc;
  if (A->n_mrow > 0) {
    const unsigned g2 = (unsigned)((A->n_mrow + kSpmmWarps - 1) / kSpmmWarps);
    spmm_fixup_kernel<D, true><<<g2, kSpmmWarps * 32, 0, s>>>(A->n_mrow, A->mrow, A->mrow_lo, A->partials, nullptr, nullptr,
                                                             nullptr, 1.f, sc);
    RECAD_LAUNCH_CHECK();
  }
  return RECAD_OK;
}

template <int D, int UNR>
static int launch_spmm(const recad_csr* A, const float* X, float* Y, const float* C, float* Z, float alpha,
                       cudaStream_t s) {
  const int v = spmm_variant();
  const bool pers = v & 1;
  int rc;
  switch ((v >> 1) & 3) {
    case 0: rc = launch_spmm_v<D, UNR, 1, false>(A, X, Y, C, Z, alpha, pers, s); break;
    case 1: rc = launch_spmm_v<D, UNR, 6, false>(A, X, Y, C, Z, alpha, pers, s); break;
    case 2: rc = launch_spmm_v<D, UNR, 5, false>(A, X, Y, C, Z, alpha, pers, s); break;
    default: rc = launch_spmm_v<D, UNR, 4, false>(A, X, Y, C, Z, alpha, pers, s); break;
  }
  if (rc) return rc;
  if (A->n_mrow > 0) {
    const unsigned g2 = (unsigned)((A->n_mrow + kSpmmWarps - 1) / kSpmmWarps);
    spmm_fixup_kernel<D><<<g2, kSpmmWarps * 32, 0, s>>>(A->n_mrow, A->mrow, A->mrow_lo, A->partials, Y, C, Z, alpha, RowScatter{});
    RECAD_LAUNCH_CHECK();
  }
  return RECAD_OK;
}

}  // namespace recad

using namespace recad;

extern "C" int recad_spmm(const recad_csr* A, const float* X, float* Y, const float* C, float* Z, float alpha,
                          int32_t D, void* stream) {
  cudaStream_t s = as_stream(stream);
  RECAD_REQUIRE(A && X, RECAD_ERR_ARG, "spmm: null matrix or X");
  RECAD_REQUIRE(Y || Z, RECAD_ERR_ARG, "spmm: neither Y nor Z requested");
  RECAD_REQUIRE(A->n_rows > 0 && A->n_seg >= A->n_rows && A->rowptr && A->seg_row && A->seg_lo && A->seg_slot,
                RECAD_ERR_ARG, "spmm: matrix has no plan (call recad_spmm_plan)");
  RECAD_REQUIRE(A->nnz == 0 || (A->colidx && A->vals), RECAD_ERR_ARG, "spmm: null colidx/vals");
  RECAD_REQUIRE(A->n_mrow == 0 || (A->mrow && A->mrow_lo && A->partials), RECAD_ERR_ARG, "spmm: null multi-row plan");
  RECAD_REQUIRE((int64_t)2147483647 / (D / 4 > 0 ? D / 4 : 1) > 0, RECAD_ERR_ARG, "spmm: bad D");
  RECAD_REQUIRE(D >= 4 && D % 4 == 0 && D <= 4 * 32 * kGenericMaxVec, RECAD_ERR_UNSUPPORTED,
                "spmm: D = %d must be a multiple of 4 in [4, %d]", D, 4 * 32 * kGenericMaxVec);
  RECAD_REQUIRE((((uintptr_t)X | (uintptr_t)Y | (uintptr_t)C | (uintptr_t)Z | (uintptr_t)A->partials) & 15) == 0,
                RECAD_ERR_ARG, "spmm: buffers must be 16-byte aligned");
  switch (D) {
    case 32: return launch_spmm<32, 2>(A, X, Y, C, Z, alpha, s);
    case 64: return (spmm_variant() & 8) ? launch_spmm<64, 8>(A, X, Y, C, Z, alpha, s) : launch_spmm<64, 4>(A, X, Y, C, Z, alpha, s);
    case 128: return launch_spmm<128, 8>(A, X, Y, C, Z, alpha, s);
    default: break;
  }
  const unsigned grid = (unsigned)((A->n_seg + kSpmmWarps - 1) / kSpmmWarps);
  spmm_generic_kernel<<<grid, kSpmmWarps * 32, 0, s>>>(A->rowptr, A->colidx, A->vals, A->n_seg, A->seg_len, A->seg_row,
                                                      A->seg_lo, A->seg_slot, X, Y, C, Z, alpha, A->partials, D);
  RECAD_LAUNCH_CHECK();
  if (A->n_mrow > 0) {
    const unsigned g2 = (unsigned)((A->n_mrow + kSpmmWarps - 1) / kSpmmWarps);
    spmm_generic_fixup_kernel<<<g2, kSpmmWarps * 32, 0, s>>>(A->n_mrow, A->mrow, A->mrow_lo, A->partials, Y, C, Z,
                                                            alpha, D);
    RECAD_LAUNCH_CHECK();
  }
  return RECAD_OK;
}

// ---------------------------------------------------------------------------------------------- peer-memory path
namespace recad {

struct PeerOut {
  float* out[kMaxPeers];
};

// out[r][e] = sum_s stage[s * stride + e] for every peer r: the owner adds the partial rows its peers sent (fixed
// rank order => every replica receives the same bits) and stores the sum straight into each peer's table
// kMulticast: out[0] is an NVSwitch multicast address -- ONE multimem.st is replicated into every rank's buffer by
// the switch instead of world separate peer stores
template <bool kMulticast>
__global__ void __launch_bounds__(256)
peer_reduce_bcast_kernel(const float4* stage, int n_src, int64_t stride4, int64_t n4, const __grid_constant__ PeerOut po,
                         int n_out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    float4 a = __ldcg(stage + e);                 // written by peers over NVLink: read at L2, never through L1 / nc
    for (int s = 1; s < n_src; ++s) {
      const float4 p = __ldcg(stage + s * stride4 + e);
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    if (kMulticast) {
      asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<float4*>(po.out[0]) + e),
                   "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w)
                   : "memory");
    } else {
      for (int r = 0; r < n_out; ++r) reinterpret_cast<float4*>(po.out[r])[e] = a;
    }
  }
}

}  // namespace recad

extern "C" int recad_spmm_scatter(const recad_csr* A, const float* X, float* const* dst, int32_t n_dst, int64_t slice_rows,
                                  int32_t D, void* stream) {
  cudaStream_t s = as_stream(stream);
  RECAD_REQUIRE(A && X && dst, RECAD_ERR_ARG, "spmm_scatter: null argument");
  RECAD_REQUIRE(A->n_rows > 0 && A->n_seg >= A->n_rows && A->rowptr && A->seg_row && A->seg_lo && A->seg_slot,
                RECAD_ERR_ARG, "spmm_scatter: matrix has no plan (call recad_spmm_plan)");
  RECAD_REQUIRE(A->n_mrow == 0 || (A->mrow && A->mrow_lo && A->partials), RECAD_ERR_ARG, "spmm_scatter: null multi-row plan");
  RECAD_REQUIRE(n_dst >= 1 && n_dst <= kMaxPeers && slice_rows > 0 && slice_rows < ((int64_t)1 << 31) &&
                    slice_rows * n_dst >= A->n_rows,
                RECAD_ERR_ARG, "spmm_scatter: %d destinations of %lld rows do not cover %lld rows", n_dst,
                (long long)slice_rows, (long long)A->n_rows);
  RowScatter sc{};
  sc.slice = (int32_t)slice_rows;
  for (int r = 0; r < n_dst; ++r) {
    RECAD_REQUIRE(dst[r] && ((uintptr_t)dst[r] & 15) == 0, RECAD_ERR_ARG, "spmm_scatter: destination %d null or misaligned", r);
    sc.dst[r] = dst[r];
  }
  RECAD_REQUIRE((((uintptr_t)X | (uintptr_t)A->partials) & 15) == 0, RECAD_ERR_ARG, "spmm_scatter: buffers must be 16-byte aligned");
  switch (D) {
    case 32: return launch_spmm_scatter<32, 2>(A, X, sc, s);
    case 64: return launch_spmm_scatter<64, 4>(A, X, sc, s);
    case 128: return launch_spmm_scatter<128, 8>(A, X, sc, s);
    default: break;
  }
  RECAD_REQUIRE(false, RECAD_ERR_UNSUPPORTED, "spmm_scatter: D = %d (supported: 32, 64, 128)", D);
  return RECAD_OK;
}

extern "C" int recad_peer_reduce_bcast(const float* stage, int32_t n_src, int64_t src_stride, int64_t n_floats,
                                       float* const* out, int32_t n_out, int32_t multicast, void* stream) {
  RECAD_REQUIRE(stage && out && n_src >= 1 && n_out >= 1 && n_out <= kMaxPeers && n_floats >= 0 && src_stride >= n_floats,
                RECAD_ERR_ARG, "peer_reduce_bcast: bad argument");
  RECAD_REQUIRE(n_floats % 4 == 0 && src_stride % 4 == 0 && ((uintptr_t)stage & 15) == 0, RECAD_ERR_ARG,
                "peer_reduce_bcast: sizes must be multiples of 4 floats, 16-byte aligned");
  if (n_floats == 0) return RECAD_OK;
  PeerOut po{};
  for (int r = 0; r < n_out; ++r) {
    RECAD_REQUIRE(out[r] && ((uintptr_t)out[r] & 15) == 0, RECAD_ERR_ARG, "peer_reduce_bcast: output %d null or misaligned", r);
    po.out[r] = out[r];
  }
  const int64_t n4 = n_floats / 4;
  const unsigned grid = (unsigned)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 8);
  RECAD_REQUIRE(!multicast || n_out == 1, RECAD_ERR_ARG, "peer_reduce_bcast: a multicast store has ONE output address");
  if (multicast)
    peer_reduce_bcast_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(stage), n_src, src_stride / 4,
                                                                        n4, po, n_out);
  else
    peer_reduce_bcast_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(stage), n_src, src_stride / 4,
                                                                         n4, po, n_out);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}
