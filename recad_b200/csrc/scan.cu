// Device-wide exclusive scan (uint32) used by the CSR builder, the radix sort and the SpMM planner.
// Three-phase reduce / scan-of-sums / downsweep, recursive on the block sums.  HBM-bound: 3 reads +
// 1 write of the array; all accesses coalesced 128-bit.
#include "common.cuh"

namespace recad {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 16;  // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  // exclusive scan of one value per thread across the block; returns this thread's prefix
  __shared__ uint32_t warp_tot[kScanThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t t = lane < kScanThreads / 32 ? warp_tot[lane] : 0;
    uint32_t ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(kFull, ti, o);
      if (lane >= o) ti += u;
    }
    if (lane < kScanThreads / 32) warp_tot[lane] = ti - t;  // exclusive warp offsets
    if (lane == 31 && total) *total = ti;
  }
  __syncthreads();
  uint32_t r = warp_tot[w] + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                   uint32_t* __restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
    if (i < n) acc += in[i];
  }
  __shared__ uint32_t tot;
  block_exclusive_scan(acc, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// each thread owns kScanItems CONSECUTIVE elements (blocked arrangement)
__global__ void __launch_bounds__(kScanThreads) scan_down_kernel(const uint32_t* __restrict__ in,
                                                                 uint32_t* __restrict__ out, int64_t n,
                                                                 const uint32_t* __restrict__ block_off,
                                                                 unsigned long long* __restrict__ total) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    acc += v[k];
  }
  __shared__ uint32_t tot;
  uint32_t pre = block_exclusive_scan(acc, &tot) + (block_off ? block_off[blockIdx.x] : 0);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = pre;
    pre += v[k];
  }
  if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *total = pre;
}

int64_t scan_scratch_bytes(int64_t n) {
  int64_t bytes = 0;
  while (n > kScanTile) {
    n = (n + kScanTile - 1) / kScanTile;
    bytes += ((n * 4 + 255) / 256) * 256;
  }
  return bytes + 256;
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, unsigned long long* total, void* scratch,
                       cudaStream_t s) {
  if (n <= 0) {
    if (total) RECAD_CUDA_CHECK(cudaMemsetAsync(total, 0, sizeof(unsigned long long), s));
    return RECAD_OK;
  }
  const int64_t nb = (n + kScanTile - 1) / kScanTile;
  if (nb == 1) {
    scan_down_kernel<<<1, kScanThreads, 0, s>>>(in, out, n, nullptr, total);
    RECAD_LAUNCH_CHECK();
    return RECAD_OK;
  }
  uint32_t* sums = reinterpret_cast<uint32_t*>(scratch);
  void* next = reinterpret_cast<char*>(scratch) + ((nb * 4 + 255) / 256) * 256;
  scan_reduce_kernel<<<(unsigned)nb, kScanThreads, 0, s>>>(in, n, sums);
  RECAD_LAUNCH_CHECK();
  int rc = exclusive_scan_u32(sums, sums, nb, nullptr, next, s);
  if (rc) return rc;
  scan_down_kernel<<<(unsigned)nb, kScanThreads, 0, s>>>(in, out, n, sums, total);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

}  // namespace recad
