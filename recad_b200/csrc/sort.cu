// Stable LSD radix sort of 64-bit keys, 11 bits per pass, written for large HBM:
// every WARP owns a contiguous chunk of the input and a PRIVATE 2048-bin counter
// table in shared memory, so both the histogram and the scatter are
// deterministic and stable without any cross-warp ranking; the price is a
// [2048 x n_chunks] counter matrix in HBM (4 bytes per key per pass at the
// default chunk size / 8192 keys per chunk = 1 KiB... i.e. 0.25 B/key), which a
// 180 GB part does not notice.  Per pass: read keys twice, write once
// (scattered 8-byte stores), scan the counter matrix.
#include "common.cuh"

namespace recad {

constexpr int kRadixBits = 11;
constexpr int kBins = 1 << kRadixBits;
constexpr int kSortWarps = 8;                       // per CTA -> 64 KiB of counters
constexpr int kChunk = 8192;                        // keys per warp
constexpr int kSortSmem = kSortWarps * kBins * 4;

__global__ void __launch_bounds__(kSortWarps * 32) sort_hist_kernel(const uint64_t* __restrict__ keys, int64_t n,
                                                                    int shift, int64_t n_chunks,
                                                                    uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t chunk = (int64_t)blockIdx.x * kSortWarps + w;
  uint32_t* cnt = smem + w * kBins;
  for (int b = lane; b < kBins; b += 32) cnt[b] = 0;
  __syncwarp();
  if (chunk < n_chunks) {
    const int64_t lo = chunk * kChunk;
    const int64_t hi = min(lo + (int64_t)kChunk, n);
    for (int64_t i = lo + lane; i < hi; i += 32) {
      uint32_t d = (uint32_t)(keys[i] >> shift) & (kBins - 1);
      atomicAdd(&cnt[d], 1u);
    }
    __syncwarp();
    for (int b = lane; b < kBins; b += 32) hist[(int64_t)b * n_chunks + chunk] = cnt[b];
  }
}

__global__ void __launch_bounds__(kSortWarps * 32) sort_scatter_kernel(const uint64_t* __restrict__ keys,
                                                                       uint64_t* __restrict__ out, int64_t n,
                                                                       int shift, int64_t n_chunks,
                                                                       const uint32_t* __restrict__ offs) {
  extern __shared__ uint32_t smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t chunk = (int64_t)blockIdx.x * kSortWarps + w;
  if (chunk >= n_chunks) return;
  uint32_t* cnt = smem + w * kBins;
  for (int b = lane; b < kBins; b += 32) cnt[b] = offs[(int64_t)b * n_chunks + chunk];
  __syncwarp();
  const int64_t lo = chunk * kChunk;
  const int64_t hi = min(lo + (int64_t)kChunk, n);
  const unsigned lt = (1u << lane) - 1;
  constexpr int kAhead = 4;
  uint64_t buf[kAhead];
#pragma unroll
  for (int k = 0; k < kAhead; ++k) {
    int64_t i = lo + (int64_t)k * 32 + lane;
    buf[k] = i < hi ? keys[i] : 0;
  }
  for (int64_t t = lo; t < hi; t += 32 * kAhead) {
    uint64_t cur[kAhead];
#pragma unroll
    for (int k = 0; k < kAhead; ++k) cur[k] = buf[k];
#pragma unroll
    for (int k = 0; k < kAhead; ++k) {  // prefetch the next group of tiles
      int64_t i = t + (int64_t)(kAhead + k) * 32 + lane;
      buf[k] = i < hi ? keys[i] : 0;
    }
#pragma unroll
    for (int k = 0; k < kAhead; ++k) {
      const int64_t i = t + (int64_t)k * 32 + lane;
      const bool valid = i < hi;
      // invalid lanes get a digit no valid lane can have, so they form their own group
      const uint32_t d = valid ? ((uint32_t)(cur[k] >> shift) & (kBins - 1)) : 0xffffffffu;
      const unsigned grp = __match_any_sync(kFull, d);
      const int rank = __popc(grp & lt);
      uint32_t base = 0;
      if (valid) base = cnt[d];
      __syncwarp();
      if (valid && rank == 0) cnt[d] = base + __popc(grp);
      __syncwarp();
      if (valid) out[(int64_t)base + rank] = cur[k];
    }
  }
}

int64_t sort_scratch_bytes(int64_t n) {
  const int64_t n_chunks = (n + kChunk - 1) / kChunk;
  const int64_t hist = ((n_chunks * kBins * 4 + 255) / 256) * 256;
  return hist + scan_scratch_bytes(n_chunks * kBins) + 256;
}

int radix_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, int bits, void* scratch, cudaStream_t s) {
  if (n <= 1) return RECAD_OK;
  RECAD_REQUIRE(n < (int64_t)0xffffffffLL, RECAD_ERR_OVERFLOW, "radix sort: %lld keys exceed 2^32", (long long)n);
  static bool attr_set = false;
  if (!attr_set) {
    RECAD_CUDA_CHECK(cudaFuncSetAttribute(sort_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmem));
    RECAD_CUDA_CHECK(cudaFuncSetAttribute(sort_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmem));
    attr_set = true;
  }
  const int64_t n_chunks = (n + kChunk - 1) / kChunk;
  uint32_t* hist = reinterpret_cast<uint32_t*>(scratch);
  void* scan_scr = reinterpret_cast<char*>(scratch) + ((n_chunks * kBins * 4 + 255) / 256) * 256;
  const unsigned grid = (unsigned)((n_chunks + kSortWarps - 1) / kSortWarps);
  int passes = (bits + kRadixBits - 1) / kRadixBits;
  if (passes & 1) ++passes;  // even number of passes: the result lands back in `keys`
  uint64_t* src = keys;
  uint64_t* dst = tmp;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kRadixBits;
    sort_hist_kernel<<<grid, kSortWarps * 32, kSortSmem, s>>>(src, n, shift, n_chunks, hist);
    RECAD_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(hist, hist, n_chunks * kBins, nullptr, scan_scr, s);
    if (rc) return rc;
    sort_scatter_kernel<<<grid, kSortWarps * 32, kSortSmem, s>>>(src, dst, n, shift, n_chunks, hist);
    RECAD_LAUNCH_CHECK();
    uint64_t* t = src;
    src = dst;
    dst = t;
  }
  return RECAD_OK;
}

}  // namespace recad
