// WMF surrogate of the AIA / Leg-UP attack loop (reference: recad/model/attacker/aia.py:222-247 WeightedMF, 283-289
// weighted_mse_loss, 393-489 WMFTrainer.fit_adv): dim-16 weighted-MSE matrix factorisation over the dense rating matrix
// [(n_users + n_fake) x n_items], batch of 16 rows, dense Adam with weight decay on EVERY parameter after every batch,
// 50 epochs per attack step, the last `unroll_steps` epochs differentiable w.r.t. the rating matrix (the reference uses
// `higher`; here the reverse pass through the unrolled Adam steps is written out).
//
// One step is 3 MFLOP and 4 MB of traffic, and a call is ~19 000 dependent steps (ml1m): the cost is launch latency and
// synchronisation, not arithmetic -- tensor cores have nothing to bite on (M = 16, K = 16).  So the whole call is ONE
// persistent kernel on ONE thread-block cluster: the items (rows of Q with their Adam state, columns of the batch's data
// rows) are spread over the threads of the cluster, which makes the Q update thread-private; the batch's 16 x dim
// gradient of P is reduced warp -> CTA -> cluster through distributed shared memory in a fixed order (deterministic);
// every row of P is then updated by the whole cluster.  Two cluster barriers per step (~0.3 us each) replace ~10 kernel
// launches per step of the stock implementation.
#include <cooperative_groups.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace recad {

constexpr int kWmfBatchMax = 32;
constexpr int kWmfThreads = 256;

struct WmfArgs {
  float *P, *Q, *mP, *vP, *mQ, *vQ;
  const float* data;
  const int32_t* orders;       // [n_epochs, n_rows]
  int64_t n_rows, n_items;
  int batch, n_epochs;
  int64_t step0;               // optimiser steps taken before this call
  float lr, b1, b2, eps, wd, wpos, wneg;
  int higher;                  // 0: torch.optim.Adam arithmetic, 1: higher.optim.DifferentiableAdam arithmetic
  float* snap;                 // per step [P | Q | mP' | mQ' | vP' | vQ'] (values BEFORE the step for P / Q, AFTER it for m / v) or null
  // backward only
  float *Pbar, *Qbar, *amP, *avP, *amQ, *avQ, *gbar, *ddata;
};

struct StepScalars {
  float w1, w2, step_size, bc2_sqrt;
};
__device__ __forceinline__ StepScalars step_scalars(const WmfArgs& a, int64_t t) {
  const double bc1 = 1.0 - pow((double)a.b1, (double)t), bc2 = 1.0 - pow((double)a.b2, (double)t);
  StepScalars s;
  s.w1 = (float)(1.0 - (double)a.b1);
  s.w2 = (float)(1.0 - (double)a.b2);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.step_size = a.higher ? (float)((double)a.lr * sqrt(bc2) / bc1) : (float)((double)a.lr / bc1);
  return s;
}
// one Adam element; g already holds weight_decay * p
__device__ __forceinline__ void adam1(const WmfArgs& a, const StepScalars& s, float g, float& p, float& m, float& v) {
  if (a.higher) {
    m = m * a.b1 + s.w1 * g;
    v = v * a.b2 + s.w2 * g * g;
    p = p - s.step_size * m / (sqrtf(v) + a.eps);
  } else {
    m = m + s.w1 * (g - m);
    v = v * a.b2 + (s.w2 * g) * g;
    p = p - s.step_size * (m / (sqrtf(v) / s.bc2_sqrt + a.eps));
  }
}

// shared memory carve-up (floats)
template <int D>
struct WmfSmem {
  static constexpr int kWarps = kWmfThreads / 32;
  static constexpr int kOut = kWmfBatchMax * D;              // outputs of the batch gradient of P
  float* sP;        // [kWmfBatchMax * D]       batch rows of P
  int* sB;          // [kWmfBatchMax]           their row ids
  float* sX;        // [kWarps][2][kWmfBatchMax][32]   per warp: residual-like tiles (b, lane)
  float* sY;        // [kWarps][2][32][D + 1]          per warp: row-like tiles (lane, d)
  float* sG;        // [kWarps][kOut]           warp partials
  float* sGc;       // [kOut]                   CTA partial (read by the whole cluster)
  float* sGt;       // [kOut]                   cluster total
  float* sGb;       // [kOut]                   backward: g-bar of the batch rows
  __device__ explicit WmfSmem(float* base) {
    sP = base; base += kOut;
    sB = reinterpret_cast<int*>(base); base += kWmfBatchMax;
    sX = base; base += kWarps * 2 * kWmfBatchMax * 32;
    sY = base; base += kWarps * 2 * 32 * (D + 1);
    sG = base; base += kWarps * kOut;
    sGc = base; base += kOut;
    sGt = base; base += kOut;
    sGb = base;
  }
  static constexpr size_t bytes() {
    return sizeof(float) * (size_t)(kOut + kWmfBatchMax + kWarps * 2 * kWmfBatchMax * 32 + kWarps * 2 * 32 * (D + 1) + kWarps * kOut + 3 * kOut);
  }
};

// warp partial of  out[b, d] += sum_j X[b][j] * Y[j][d]  over the warp's 32 staged items (outputs o = lane + 32 k)
template <int D>
__device__ __forceinline__ void warp_outer(const float* __restrict__ X, const float* __restrict__ Y, int nb, int lane, float scale, float* acc) {
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const int o = lane + 32 * k;
    if (o < nb * D) {
      const int b = o / D, d = o % D;
      float s = 0.f;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) s = fmaf(X[b * 32 + j], Y[j * (D + 1) + d], s);
      acc[k] = fmaf(scale, s, acc[k]);
    }
  }
}

// warp partials -> CTA partial (fixed warp order) -> cluster total (fixed rank order) in sm.sGt of every CTA
template <int D>
__device__ __forceinline__ void cluster_reduce(cg::cluster_group& cluster, WmfSmem<D>& sm, const float* acc, int nb, int tid, int lane, int warp) {
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const int o = lane + 32 * k;
    if (o < nb * D) sm.sG[warp * WmfSmem<D>::kOut + o] = acc[k];
  }
  __syncthreads();
  for (int o = tid; o < nb * D; o += kWmfThreads) {
    float t = 0.f;
    for (int w = 0; w < WmfSmem<D>::kWarps; ++w) t += sm.sG[w * WmfSmem<D>::kOut + o];
    sm.sGc[o] = t;
  }
  cluster.sync();
  const int C = (int)cluster.num_blocks();
  for (int o = tid; o < nb * D; o += kWmfThreads) {
    float part[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) part[r] = r < C ? cluster.map_shared_rank(sm.sGc, r)[o] : 0.f;   // 16 remote loads in flight
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) t += part[r];                                                   // rank order: deterministic
    sm.sGt[o] = t;
  }
  __syncthreads();
}

template <int D>
__device__ __forceinline__ int batch_pos(const int* sB, int nb, int row) {
  int k = -1;
  for (int q = 0; q < nb; ++q) k = sB[q] == row ? q : k;
  return k;
}

// ------------------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(kWmfThreads, 1) wmf_fit_kernel(const __grid_constant__ WmfArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float smem_raw[];
  WmfSmem<D> sm(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
  const int64_t gthreads = (int64_t)C * kWmfThreads, gtid = (int64_t)cr * kWmfThreads + tid;
  const int64_t gwarp = gtid >> 5, ngwarps = gthreads >> 5;
  const int64_t nP = a.n_rows * D, nQ = a.n_items * D, n_params = nP + nQ;
  const int64_t spe = (a.n_rows + a.batch - 1) / a.batch;       // steps per epoch
  const int64_t n_steps = spe * a.n_epochs;
  float* X = sm.sX + warp * 2 * kWmfBatchMax * 32;
  float* Y = sm.sY + warp * 2 * 32 * (D + 1);
  for (int64_t step = 0; step < n_steps; ++step) {
    const int64_t epoch = step / spe, s0 = (step % spe) * a.batch;
    const int nb = (int)min((int64_t)a.batch, a.n_rows - s0);
    __shared__ StepScalars sc_s;
    if (tid == 0) sc_s = step_scalars(a, a.step0 + step + 1);
    float* snap = a.snap ? a.snap + step * 3 * n_params : nullptr;
    if (tid < nb) sm.sB[tid] = a.orders[epoch * a.n_rows + s0 + tid];
    __syncthreads();
    const StepScalars sc = sc_s;
    for (int e = tid; e < nb * D; e += kWmfThreads) sm.sP[e] = __ldcg(a.P + (int64_t)sm.sB[e / D] * D + e % D);   // written by other CTAs: L2
    __syncthreads();
    float acc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = 0.f;
    // items: residuals, gradient of the Q row (thread private) and the warp's part of the gradient of the batch rows of P
    for (int64_t chunk = gwarp; chunk * 32 < a.n_items; chunk += ngwarps) {
      const int64_t i = chunk * 32 + lane;
      const bool on = i < a.n_items;
      float q[D], gq[D], dv[kWmfBatchMax];
#pragma unroll
      for (int b = 0; b < kWmfBatchMax; ++b)          // all loads of the batch's column issued before any is used
        dv[b] = (on && b < nb) ? __ldg(a.data + (int64_t)sm.sB[b] * a.n_items + i) : 0.f;
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        const float4 t = on ? *reinterpret_cast<const float4*>(a.Q + i * D + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        q[d] = t.x; q[d + 1] = t.y; q[d + 2] = t.z; q[d + 3] = t.w;
      }
      float mq[D], vq[D];
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        const float4 tm = on ? *reinterpret_cast<const float4*>(a.mQ + i * D + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 tv = on ? *reinterpret_cast<const float4*>(a.vQ + i * D + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        mq[d] = tm.x; mq[d + 1] = tm.y; mq[d + 2] = tm.z; mq[d + 3] = tm.w;
        vq[d] = tv.x; vq[d + 1] = tv.y; vq[d + 2] = tv.z; vq[d + 3] = tv.w;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) { gq[d] = 0.f; Y[lane * (D + 1) + d] = q[d]; }
#pragma unroll
      for (int b = 0; b < kWmfBatchMax; ++b) {
        if (b < nb) {
          const float w = dv[b] > 0.f ? a.wpos : a.wneg;
          float logit = 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) logit = fmaf(sm.sP[b * D + d], q[d], logit);
          const float r = on ? w * (dv[b] - logit) : 0.f;
          X[b * 32 + lane] = r;
#pragma unroll
          for (int d = 0; d < D; ++d) gq[d] = fmaf(-2.f * r, sm.sP[b * D + d], gq[d]);
        }
      }
      __syncwarp();
      warp_outer<D>(X, Y, nb, lane, -2.f, acc);
      __syncwarp();
      if (on) {
        float pn[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          pn[d] = q[d];
          adam1(a, sc, gq[d] + a.wd * q[d], pn[d], mq[d], vq[d]);
        }
#pragma unroll
        for (int d = 0; d < D; d += 4) {
          *reinterpret_cast<float4*>(a.Q + i * D + d) = make_float4(pn[d], pn[d + 1], pn[d + 2], pn[d + 3]);
          *reinterpret_cast<float4*>(a.mQ + i * D + d) = make_float4(mq[d], mq[d + 1], mq[d + 2], mq[d + 3]);
          *reinterpret_cast<float4*>(a.vQ + i * D + d) = make_float4(vq[d], vq[d + 1], vq[d + 2], vq[d + 3]);
          if (snap) {
            *reinterpret_cast<float4*>(snap + nP + i * D + d) = make_float4(q[d], q[d + 1], q[d + 2], q[d + 3]);
            *reinterpret_cast<float4*>(snap + n_params + nP + i * D + d) = make_float4(mq[d], mq[d + 1], mq[d + 2], mq[d + 3]);
            *reinterpret_cast<float4*>(snap + 2 * n_params + nP + i * D + d) = make_float4(vq[d], vq[d + 1], vq[d + 2], vq[d + 3]);
          }
        }
      }
    }
    cluster_reduce<D>(cluster, sm, acc, nb, tid, lane, warp);
    // every row of P: dense Adam (rows outside the batch have gradient weight_decay * p only); 128-bit accesses
    if (step + 1 < n_steps) {                                    // pull the next batch's data rows towards L2 while P is updated
      const int64_t s1 = ((step + 1) % spe) * a.batch, ep1 = (step + 1) / spe;
      const int nb1 = (int)min((int64_t)a.batch, a.n_rows - s1);
      for (int64_t c = gtid; c < (int64_t)nb1 * ((a.n_items + 31) / 32); c += gthreads) {
        const int64_t row = a.orders[ep1 * a.n_rows + s1 + c / ((a.n_items + 31) / 32)];
        const float* ptr = a.data + row * a.n_items + (c % ((a.n_items + 31) / 32)) * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
      }
    }
#pragma unroll 4
    for (int64_t e4 = gtid; e4 < nP / 4; e4 += gthreads) {
      const int64_t e = e4 * 4;
      const int row = (int)(e / D), d = (int)(e % D);
      const int k = batch_pos<D>(sm.sB, nb, row);
      const float4 p0 = *reinterpret_cast<const float4*>(a.P + e);
      float4 m = *reinterpret_cast<const float4*>(a.mP + e), v = *reinterpret_cast<const float4*>(a.vP + e);
      const float4 g = k >= 0 ? *reinterpret_cast<const float4*>(sm.sGt + k * D + d) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 p = p0;
      adam1(a, sc, g.x + a.wd * p0.x, p.x, m.x, v.x);
      adam1(a, sc, g.y + a.wd * p0.y, p.y, m.y, v.y);
      adam1(a, sc, g.z + a.wd * p0.z, p.z, m.z, v.z);
      adam1(a, sc, g.w + a.wd * p0.w, p.w, m.w, v.w);
      *reinterpret_cast<float4*>(a.P + e) = p;
      *reinterpret_cast<float4*>(a.mP + e) = m;
      *reinterpret_cast<float4*>(a.vP + e) = v;
      if (snap) {
        *reinterpret_cast<float4*>(snap + e) = p0;
        *reinterpret_cast<float4*>(snap + n_params + e) = m;
        *reinterpret_cast<float4*>(snap + 2 * n_params + e) = v;
      }
    }
    __threadfence();
    cluster.sync();           // the new P rows are visible to every CTA; the shared buffers may be reused
  }
}

// ------------------------------------------------------------------------------------------ backward
// Reverse pass through the higher-form steps recorded in `snap`, last step first.  In: Pbar / Qbar = d loss / d (final P, Q);
// amP .. avQ = 0.  Out: ddata += d loss / d data (rows of the unrolled batches); Pbar / Qbar end as the adjoint of the
// parameters BEFORE the first unrolled step (unused by the caller).
template <int D>
__global__ void __launch_bounds__(kWmfThreads, 1) wmf_backward_kernel(const __grid_constant__ WmfArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float smem_raw[];
  WmfSmem<D> sm(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
  const int64_t gthreads = (int64_t)C * kWmfThreads, gtid = (int64_t)cr * kWmfThreads + tid;
  const int64_t gwarp = gtid >> 5, ngwarps = gthreads >> 5;
  const int64_t nP = a.n_rows * D, nQ = a.n_items * D, n_params = nP + nQ;
  const int64_t spe = (a.n_rows + a.batch - 1) / a.batch;
  const int64_t n_steps = spe * a.n_epochs;
  float* X = sm.sX + warp * 2 * kWmfBatchMax * 32;
  float* Y = sm.sY + warp * 2 * 32 * (D + 1);
  float* X2 = X + kWmfBatchMax * 32;
  float* Y2 = Y + 32 * (D + 1);
  const float c1 = 1.f - a.b1, c2 = 1.f - a.b2;
  for (int64_t step = n_steps - 1; step >= 0; --step) {
    const int64_t epoch = step / spe, s0 = (step % spe) * a.batch;
    const int nb = (int)min((int64_t)a.batch, a.n_rows - s0);
    const StepScalars sc = step_scalars(a, a.step0 + step + 1);
    const float alpha = sc.step_size;
    const float* snap = a.snap + step * 3 * n_params;
    const float* Pold = snap;
    const float* Qold = snap + nP;
    const float* mN = snap + n_params;          // m after the step
    const float* vN = snap + 2 * n_params;      // v after the step
    if (tid < nb) sm.sB[tid] = a.orders[epoch * a.n_rows + s0 + tid];
    __syncthreads();
    for (int e = tid; e < nb * D; e += kWmfThreads) sm.sP[e] = Pold[(int64_t)sm.sB[e / D] * D + e % D];
    __syncthreads();
    // (1) recompute the gradient of the batch rows of P (needed by the elementwise adjoint of those rows)
    float acc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = 0.f;
    for (int64_t chunk = gwarp; chunk * 32 < a.n_items; chunk += ngwarps) {
      const int64_t i = chunk * 32 + lane;
      const bool on = i < a.n_items;
      float q[D];
#pragma unroll
      for (int d = 0; d < D; ++d) { q[d] = on ? Qold[i * D + d] : 0.f; Y[lane * (D + 1) + d] = q[d]; }
      for (int b = 0; b < nb; ++b) {
        const float dv = on ? a.data[(int64_t)sm.sB[b] * a.n_items + i] : 0.f;
        const float w = dv > 0.f ? a.wpos : a.wneg;
        float logit = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) logit = fmaf(sm.sP[b * D + d], q[d], logit);
        X[b * 32 + lane] = on ? w * (dv - logit) : 0.f;
      }
      __syncwarp();
      warp_outer<D>(X, Y, nb, lane, -2.f, acc);
      __syncwarp();
    }
    cluster_reduce<D>(cluster, sm, acc, nb, tid, lane, warp);       // sGt = gradient (without decay) of the batch rows
    // (2) every row of P: adjoint of the Adam update; g-bar of the batch rows goes to a.gbar
    for (int64_t e = gtid; e < nP; e += gthreads) {
      const int row = (int)(e / D), d = (int)(e % D);
      const int k = batch_pos<D>(sm.sB, nb, row);
      const float g = (k >= 0 ? sm.sGt[k * D + d] : 0.f) + a.wd * Pold[e];
      const float v1 = vN[e], m1 = mN[e], sq = sqrtf(v1), den = sq + a.eps;
      const float tb = __ldcg(a.Pbar + e);             // CTA 0 adds the batch rows' bilinear part: read at L2
      const float mt = a.amP[e] - alpha * tb / den;
      const float vt = a.avP[e] + (v1 > 0.f ? tb * alpha * m1 / (den * den * 2.f * sq) : 0.f);
      const float gb = c1 * mt + 2.f * c2 * g * vt;
      a.amP[e] = a.b1 * mt;
      a.avP[e] = a.b2 * vt;
      a.Pbar[e] = tb + a.wd * gb;
      if (k >= 0) a.gbar[k * D + d] = gb;
    }
    __threadfence();
    cluster.sync();
    for (int e = tid; e < nb * D; e += kWmfThreads) sm.sGb[e] = __ldcg(a.gbar + e);
    __syncthreads();
    // (3) items: adjoint of the Q rows, of the residuals (-> ddata) and the items' part of the adjoint of the batch rows
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = 0.f;
    for (int64_t chunk = gwarp; chunk * 32 < a.n_items; chunk += ngwarps) {
      const int64_t i = chunk * 32 + lane;
      const bool on = i < a.n_items;
      float q[D], gq[D], r[kWmfBatchMax], wv[kWmfBatchMax];
#pragma unroll
      for (int d = 0; d < D; ++d) { q[d] = on ? Qold[i * D + d] : 0.f; gq[d] = 0.f; }
#pragma unroll
      for (int b = 0; b < kWmfBatchMax; ++b) {
        if (b < nb) {
          const float dv = on ? a.data[(int64_t)sm.sB[b] * a.n_items + i] : 0.f;
          wv[b] = dv > 0.f ? a.wpos : a.wneg;
          float logit = 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) logit = fmaf(sm.sP[b * D + d], q[d], logit);
          r[b] = on ? wv[b] * (dv - logit) : 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) gq[d] = fmaf(-2.f * r[b], sm.sP[b * D + d], gq[d]);
        } else { r[b] = 0.f; wv[b] = 0.f; }
      }
      float gb[D], qb[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        gb[d] = 0.f; qb[d] = 0.f;
        if (on) {
          const int64_t e = nP + i * D + d;
          const float g = gq[d] + a.wd * q[d];
          const float v1 = vN[e], m1 = mN[e], sq = sqrtf(v1), den = sq + a.eps;
          const float tb = a.Qbar[i * D + d];
          const float mt = a.amQ[i * D + d] - alpha * tb / den;
          const float vt = a.avQ[i * D + d] + (v1 > 0.f ? tb * alpha * m1 / (den * den * 2.f * sq) : 0.f);
          gb[d] = c1 * mt + 2.f * c2 * g * vt;
          a.amQ[i * D + d] = a.b1 * mt;
          a.avQ[i * D + d] = a.b2 * vt;
          qb[d] = tb + a.wd * gb[d];
        }
        Y[lane * (D + 1) + d] = -2.f * gb[d];
        Y2[lane * (D + 1) + d] = q[d];
      }
#pragma unroll
      for (int b = 0; b < kWmfBatchMax; ++b) {
        if (b < nb) {
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) { t1 = fmaf(sm.sGb[b * D + d], q[d], t1); t2 = fmaf(sm.sP[b * D + d], gb[d], t2); }
          const float rbar = -2.f * (t1 + t2);
          const float sv = -wv[b] * rbar;                       // adjoint of the logit
          if (on) a.ddata[(int64_t)sm.sB[b] * a.n_items + i] += wv[b] * rbar;
#pragma unroll
          for (int d = 0; d < D; ++d) qb[d] = fmaf(-2.f * r[b], sm.sGb[b * D + d], fmaf(sv, sm.sP[b * D + d], qb[d]));
          X[b * 32 + lane] = r[b];
          X2[b * 32 + lane] = on ? sv : 0.f;
        }
      }
      if (on) {
#pragma unroll
        for (int d = 0; d < D; ++d) a.Qbar[i * D + d] = qb[d];
      }
      __syncwarp();
      warp_outer<D>(X, Y, nb, lane, 1.f, acc);                  // sum_i r[b, i] * (-2 gbar_Q[i, d])
      warp_outer<D>(X2, Y2, nb, lane, 1.f, acc);                // sum_i s[b, i] * Q[i, d]
      __syncwarp();
    }
    cluster_reduce<D>(cluster, sm, acc, nb, tid, lane, warp);
    if (cr == 0)
      for (int e = tid; e < nb * D; e += kWmfThreads) {
        float* pb = a.Pbar + (int64_t)sm.sB[e / D] * D + e % D;
        *pb = __ldcg(pb) + sm.sGt[e];
      }
    __threadfence();
    cluster.sync();
  }
}

template <int D>
static int launch_wmf(bool backward, const WmfArgs& a, cudaStream_t s) {
  auto kern = backward ? wmf_backward_kernel<D> : wmf_fit_kernel<D>;
  const size_t smem = WmfSmem<D>::bytes();
  RECAD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int C = 16;                                          // B200: 16 CTAs per cluster with the non-portable opt-in
  if (const char* e = getenv("RECAD_WMF_CLUSTER")) C = std::max(1, std::min(16, atoi(e)));
  while (C > 1 && (int64_t)C * kWmfThreads / 2 > a.n_items + a.n_rows) C >>= 1;     // tiny problems: fewer CTAs, cheaper barriers
  if (C > 8) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      C = 8;
    }
  }
  for (;; C >>= 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)C);
    cfg.blockDim = dim3(kWmfThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e == cudaSuccess) return RECAD_OK;
    if (C <= 8) RECAD_CUDA_CHECK(e);
    cudaGetLastError();                                // a 16-CTA cluster could not be placed: fall back to the portable size
  }
}

static int check_wmf(const recad_wmf* st, const float* data, const int32_t* orders, int32_t n_epochs, const char* who) {
  RECAD_REQUIRE(st && data && orders && n_epochs >= 0, RECAD_ERR_ARG, "%s: null argument", who);
  RECAD_REQUIRE(st->n_rows > 0 && st->n_items > 0 && st->n_rows < ((int64_t)1 << 31) && st->n_items < ((int64_t)1 << 31), RECAD_ERR_ARG,
                "%s: bad shape", who);
  RECAD_REQUIRE(st->dim == 8 || st->dim == 16 || st->dim == 32, RECAD_ERR_UNSUPPORTED, "%s: hidden_dim = %d (supported: 8, 16, 32)", who, st->dim);
  RECAD_REQUIRE(st->batch >= 1 && st->batch <= kWmfBatchMax, RECAD_ERR_UNSUPPORTED, "%s: batch_size = %d (supported: 1 .. %d)", who, st->batch,
                kWmfBatchMax);
  RECAD_REQUIRE(st->P && st->Q && st->mP && st->vP && st->mQ && st->vQ, RECAD_ERR_ARG, "%s: null parameter / optimiser buffer", who);
  return RECAD_OK;
}

static WmfArgs make_wmf_args(const recad_wmf* st, const float* data, const int32_t* orders, int32_t n_epochs, int64_t step0, int higher,
                             float* snap) {
  WmfArgs a{};
  a.P = st->P; a.Q = st->Q; a.mP = st->mP; a.vP = st->vP; a.mQ = st->mQ; a.vQ = st->vQ;
  a.data = data; a.orders = orders; a.n_rows = st->n_rows; a.n_items = st->n_items; a.batch = st->batch; a.n_epochs = n_epochs;
  a.step0 = step0; a.lr = st->lr; a.b1 = st->beta1; a.b2 = st->beta2; a.eps = st->eps; a.wd = st->weight_decay;
  a.wpos = st->weight_pos; a.wneg = st->weight_neg; a.higher = higher; a.snap = snap;
  return a;
}

}  // namespace recad

using namespace recad;

extern "C" {

int64_t recad_wmf_snapshot_floats(const recad_wmf* st, int32_t n_epochs) {
  if (!st || n_epochs < 0 || st->batch < 1) return 0;
  const int64_t spe = (st->n_rows + st->batch - 1) / st->batch;
  return spe * n_epochs * 3 * (st->n_rows + st->n_items) * st->dim;
}

int recad_wmf_fit(const recad_wmf* st, const float* data, const int32_t* orders, int32_t n_epochs, int64_t step0, int32_t unrolled,
                  float* snap, void* stream) {
  int rc = check_wmf(st, data, orders, n_epochs, "wmf_fit");
  if (rc) return rc;
  RECAD_REQUIRE(step0 >= 0 && (!snap || unrolled), RECAD_ERR_ARG, "wmf_fit: snapshots are recorded for unrolled epochs only");
  if (n_epochs == 0) return RECAD_OK;
  const WmfArgs a = make_wmf_args(st, data, orders, n_epochs, step0, unrolled ? 1 : 0, snap);
  switch (st->dim) {
    case 8: return launch_wmf<8>(false, a, as_stream(stream));
    case 16: return launch_wmf<16>(false, a, as_stream(stream));
    default: return launch_wmf<32>(false, a, as_stream(stream));
  }
}

int recad_wmf_backward(const recad_wmf* st, const float* data, const int32_t* orders, int32_t n_epochs, int64_t step0, const float* snap,
                       float* Pbar, float* Qbar, float* scratch, float* d_data, void* stream) {
  int rc = check_wmf(st, data, orders, n_epochs, "wmf_backward");
  if (rc) return rc;
  RECAD_REQUIRE(snap && Pbar && Qbar && scratch && d_data && step0 >= 0, RECAD_ERR_ARG, "wmf_backward: null argument");
  if (n_epochs == 0) return RECAD_OK;
  cudaStream_t s = as_stream(stream);
  const int64_t nP = st->n_rows * st->dim, nQ = st->n_items * st->dim;
  RECAD_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (2 * (nP + nQ) + kWmfBatchMax * 32) * sizeof(float), s));
  WmfArgs a = make_wmf_args(st, data, orders, n_epochs, step0, 1, const_cast<float*>(snap));
  a.Pbar = Pbar; a.Qbar = Qbar;
  a.amP = scratch; a.avP = scratch + nP; a.amQ = scratch + 2 * nP; a.avQ = scratch + 2 * nP + nQ; a.gbar = scratch + 2 * (nP + nQ);
  a.ddata = d_data;
  switch (st->dim) {
    case 8: return launch_wmf<8>(true, a, s);
    case 16: return launch_wmf<16>(true, a, s);
    default: return launch_wmf<32>(true, a, s);
  }
}

}  // extern "C"
