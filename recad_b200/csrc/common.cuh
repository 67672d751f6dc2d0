// Shared helpers for the recad_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/recad_b200.h"

namespace recad {

void set_error(const char* fmt, ...);

#define RECAD_CUDA_CHECK(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::recad::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return RECAD_ERR_CUDA;                                                                \
    }                                                                                       \
  } while (0)

#define RECAD_LAUNCH_CHECK() RECAD_CUDA_CHECK(cudaGetLastError())

#define RECAD_REQUIRE(cond, code, ...)   \
  do {                                   \
    if (!(cond)) {                       \
      ::recad::set_error(__VA_ARGS__);   \
      return (code);                     \
    }                                    \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached SM count of the current device

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// streaming (read-once) loads: do not pollute L1
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// vector fp32 reduction to global memory (sm_90+): one 16-byte RED instead of four
__device__ __forceinline__ void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// device-wide exclusive scan of uint32 (scan.cu).  out may alias in.  total (device, may be null)
// receives the grand total as uint64.  scratch: scan_scratch_bytes(n).
int64_t scan_scratch_bytes(int64_t n);
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, unsigned long long* total,
                       void* scratch, cudaStream_t s);
// stable LSD radix sort of 64-bit keys on their low `bits` bits (sort.cu).
// keys is sorted in place (tmp is the ping-pong buffer, same size).
int64_t sort_scratch_bytes(int64_t n);
int radix_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, int bits, void* scratch, cudaStream_t s);

// dense Adam shared by every victim (bpr.cu)
struct AdamScalars {
  float w1;         // 1 - beta1
  float b2;         // beta2
  float w2;         // 1 - beta2
  float step_size;  // lr / (1 - beta1^t)
  float bc2_sqrt;   // sqrt(1 - beta2^t)
  float eps;
};
// One element's Adam step, with every rounding spelled out: the dense kernels, the scalar tail, the graph-replayed NCF step
// and the lazy embedding rows must produce the SAME bits (left to the compiler, `V * b2 + (w2 * G) * G` was contracted
// around the first product in one kernel and around the second in another: the graph-replayed NCF epoch differed from the
// plain loop in the last bit of a few parameters).
__device__ __forceinline__ void adam_update(float& P, float G, float& M, float& V, const AdamScalars& a) {
  M = __fmaf_rn(a.w1, __fsub_rn(G, M), M);
  V = __fmaf_rn(V, a.b2, __fmul_rn(__fmul_rn(a.w2, G), G));
  P = __fmaf_rn(-a.step_size, __fdiv_rn(M, __fadd_rn(__fdiv_rn(__fsqrt_rn(V), a.bc2_sqrt), a.eps)), P);
}
// optional end-of-batch bookkeeping done by thread 0 of the Adam launch:
//   acc[2] += acc[0] * inv_B + half_lambda * acc[1] * inv_B;  acc[0] = acc[1] = 0
struct LossFold {
  double* acc;
  double inv_B;
  double half_lambda;
};
AdamScalars adam_scalars(float lr, float b1, float b2, float eps, int64_t step);
int launch_adam(float* p, const float* g, const float* cnt, float reg_scale, float* m, float* v, int64_t n, int D,
                const AdamScalars& a, const LossFold& fold, cudaStream_t s);

}  // namespace recad
