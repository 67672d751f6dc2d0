// AUSH generator / discriminator step (reference: recad/model/attacker/aush.py:78-180 train_step, 182-230 generate_fake,
// 254-283 the two networks).
//
// What the reference computes per batch of B user rows over I items, with dense [B, I] tensors:
//   input_template = real * fillers_mask                     (filler_num sampled columns per row)
//   gen            = 5 * sigmoid(W2g sigmoid(W1g template + b1g) + b2g)     -- netG in eval mode, DETACHED (aush.py:126-128)
//   fake           = template + gen * selects_mask + 5 * [column selected]  (aush.py:130-132; train_step patches the
//                                                                            SELECTED columns, generate_fake the targets)
//   D step         : BCE(D(real * (fillers + selects)), 1) / 2 + BCE(D(fake * (fillers + selects)), 0) / 2, backward, Adam
//   reported       : BCE(D_new(fake * ...), 1), MSE(fake * selects, 5 * selects), MSE on the ZR-masked selected columns
// Because `gen` is detached, g_loss never reaches the generator's parameters: G_optimizer.step() has no gradients and
// netG stays at its initial weights (pinned by tests/golden/make_golden_aush.py: "generator moved by 0.0").  The step
// is therefore G forward + D forward/backward/Adam + D forward.
//
// B200 formulation: every network input has at most filler_num + |selected| non-zeros per row, so the first layer of
// both networks is a GATHER of weight rows, not a [B, I] x [I, H] GEMM, and only the selected columns of the
// generator's output layer are ever used.  The discriminator's first-layer gradient is a scatter into the same few
// rows; it is fused with that layer's dense Adam update through a per-batch column index built on the host
// (deterministic: fixed summation order, no atomics).  Kernels per batch:
//   aush_gen_kernel     warp per row: gather-sum of W1g^T rows, sigmoid, |selected| dot products
//   aush_disc_kernel<1> block per 8 rows: gather layer 1, two 150 x 150 layers staged in shared memory (cp.async), BCE,
//                       full backward; per-block partial weight gradients, dz1 rows to global
//   aush_adam_w1_kernel block per item row: gradient from the column index + Adam
//   aush_adam_small_kernel  the other parameters: block partials summed in block order + Adam
//   aush_disc_kernel<0> forward of the updated discriminator on the fake rows (g_loss_gan)
// and one aush_loss_kernel per epoch folding the per-row terms into the four epoch means.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "tc_common.cuh"

namespace recad {

constexpr int kAH = 150;        // discriminator hidden width (aush.py:272-278)
constexpr int kAHP = 152;       // padded row stride (16-byte rows); the two pad columns stay exactly zero
constexpr int kAG = 128;        // generator hidden width (aush.py:258)
constexpr int kAR = 8;          // rows per block pass
constexpr int kAThreads = 160;  // one thread per hidden unit (+ idle pad threads)

struct AushLayout {
  int64_t W1t, b1, W2, b2, W3, b3, w4, b4, total;
};
static AushLayout aush_layout(int64_t I) {
  AushLayout o;
  int64_t p = 0;
  o.W1t = p; p += I * kAHP;
  o.b1 = p; p += kAHP;
  o.W2 = p; p += (int64_t)kAH * kAHP;
  o.b2 = p; p += kAHP;
  o.W3 = p; p += (int64_t)kAH * kAHP;
  o.b3 = p; p += kAHP;
  o.w4 = p; p += kAHP;
  o.b4 = p; p += 4;
  o.total = p;
  return o;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void adam1(float& P, float G, float& M, float& V, const AdamScalars& a) {
  M = M + a.w1 * (G - M);
  V = V * a.b2 + (a.w2 * G) * G;
  P = P - a.step_size * (M / (sqrtf(V) / a.bc2_sqrt + a.eps));
}

// ------------------------------------------------------------------------------------------
// generator forward on the selected columns.  One warp per row; lane l owns hidden units 4l .. 4l+3.  The row's
// (column, value) pairs are fetched 32 at a time with one coalesced load and handed round with shuffles, so that the
// gathers of W1g^T rows are independent of each other (4 in flight).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
aush_gen_kernel(const float* __restrict__ W1t, const float* __restrict__ b1, const float* __restrict__ W2,
                const float* __restrict__ b2, const int32_t* __restrict__ sel, int S, const int32_t* __restrict__ cols,
                const float* __restrict__ tval, int F, const float* __restrict__ tsel, const float* __restrict__ msel,
                const float* __restrict__ zr, int64_t B, float* __restrict__ gen_out, float* __restrict__ fin,
                float* __restrict__ shil, float* __restrict__ rec) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float4 h = reinterpret_cast<const float4*>(b1)[lane];
  float4 h1 = make_float4(0.f, 0.f, 0.f, 0.f), h2 = h1, h3 = h1;
  for (int f0 = 0; f0 < F; f0 += 32) {
    const int my_c = f0 + lane < F ? cols[row * F + f0 + lane] : 0;
    const float my_x = f0 + lane < F ? tval[row * F + f0 + lane] : 0.f;
    const int n = min(32, F - f0);
#define RECAD_AUSH_G1(acc, q)                                                                        \
    {                                                                                                \
      const int c = __shfl_sync(kFull, my_c, q);                                                     \
      const float x = __shfl_sync(kFull, my_x, q);                                                   \
      const float4 w = reinterpret_cast<const float4*>(W1t + (int64_t)c * kAG)[lane];                \
      acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w); \
    }
    int q = 0;
    for (; q + 4 <= n; q += 4) { RECAD_AUSH_G1(h, q) RECAD_AUSH_G1(h1, q + 1) RECAD_AUSH_G1(h2, q + 2) RECAD_AUSH_G1(h3, q + 3) }
    for (; q < n; ++q) RECAD_AUSH_G1(h, q)
#undef RECAD_AUSH_G1
  }
  h.x = sigmoidf_((h.x + h1.x) + (h2.x + h3.x)); h.y = sigmoidf_((h.y + h1.y) + (h2.y + h3.y));
  h.z = sigmoidf_((h.z + h1.z) + (h2.z + h3.z)); h.w = sigmoidf_((h.w + h1.w) + (h2.w + h3.w));
  float sh = 0.f, rc = 0.f;
  for (int s = 0; s < S; ++s) {
    const int c = sel[s];
    const float4 w = reinterpret_cast<const float4*>(W2 + (int64_t)c * kAG)[lane];
    const float z = warp_sum(h.x * w.x + h.y * w.y + h.z * w.z + h.w * w.w) + b2[c];
    const float g = sigmoidf_(z) * 5.f;
    if (gen_out && lane == 0) gen_out[row * S + s] = g;
    if (fin) {
      const float t = tsel[row * S + s];
      const float fake = (t + g) + 5.f;                       // template + selected_patch + target_patch (aush.py:130-132)
      const float d1 = fake - 5.f;                            // mse(fake * selects, 5 * selects)        (aush.py:156)
      const float zm = zr[row * S + s];
      const float d2 = fake * zm - t * zm;                    // mse(fake * selects * ZR, selects * template * ZR) (157-160)
      sh = fmaf(d1, d1, sh);
      rc = fmaf(d2, d2, rc);
      if (lane == 0) fin[row * S + s] = fake * msel[row * S + s];
    }
  }
  if (fin && lane == 0) { shil[row] = sh; rec[row] = rc; }
}

// ------------------------------------------------------------------------------------------
// discriminator forward (+ backward).  Virtual rows v < nv: kTrain: v < B the real row, v >= B the fake row v - B;
// !kTrain: fake rows only.  Block of 160 threads, thread t <-> hidden unit t, 8 rows per pass held in registers;
// activations are exchanged through shared memory as [unit][row] (two 128-bit broadcast loads per unit).  The two
// 150 x 152 weight matrices of a phase arrive in shared memory by bulk async copies (cp.async.bulk + mbarrier, one
// instruction each, issued a phase ahead: the forward pair while the first layer gathers, the backward pair as soon as a
// forward layer has released its buffer) -- forward from the transposed copies Dt (input-major), backward from D
// (output-major), so that thread t always reads word t of a 152-float row: no bank conflicts in either direction.
// ------------------------------------------------------------------------------------------
struct AushBatch {
  const int32_t* cols;   // [B, F]
  const float* dval;     // [B, F]
  const float* rsel;     // [B, S]
  const float* fin;      // [B, S]  fake * mask at the selected columns (aush_gen_kernel)
  const int32_t* sel;    // [S]
  int B, F, S;
};

__device__ __forceinline__ void load8(const float* p, float* o) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* o) {
  reinterpret_cast<float4*>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
}

__device__ __forceinline__ void bulk_g2s(float* dst, const float* src, uint32_t bytes, uint32_t bar) {
  mbar_expect_tx(bar, bytes);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// out[r] = bias + sum_k Ws[k][t] * act[k][r]      (Ws = transposed weight, input-major)
__device__ __forceinline__ void dense_fwd(const float* Ws, const float* act, int t, float bias, float* out) {
#pragma unroll
  for (int r = 0; r < kAR; ++r) out[r] = bias;
#pragma unroll 6
  for (int k = 0; k < kAH; ++k) {
    const float w = Ws[k * kAHP + t];
    float a[kAR];
    load8(act + k * kAR, a);
#pragma unroll
    for (int r = 0; r < kAR; ++r) out[r] = fmaf(w, a[r], out[r]);
  }
}

// thread t = input unit k of the layer (Ws = weight, output-major): gW partial[j][k] (+)= sum_r dz[j][r] a_in[k][r];
// returns d a_in[k][r] = sum_j W[j][k] dz[j][r]
__device__ __forceinline__ void dense_bwd(const float* Ws, const float* dz, const float* a_in_reg, int t, float* gW, bool first,
                                          float* da) {
#pragma unroll
  for (int r = 0; r < kAR; ++r) da[r] = 0.f;
#pragma unroll 1
  for (int j0 = 0; j0 < kAH; j0 += 10) {
    float old[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) old[q] = first ? 0.f : gW[(j0 + q) * kAHP + t];
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      float d[kAR];
      load8(dz + (j0 + q) * kAR, d);
      const float w = Ws[(j0 + q) * kAHP + t];
      float g = 0.f;
#pragma unroll
      for (int r = 0; r < kAR; ++r) {
        g = fmaf(d[r], a_in_reg[r], g);
        da[r] = fmaf(w, d[r], da[r]);
      }
      gW[(j0 + q) * kAHP + t] = old[q] + g;
    }
  }
}

template <bool kTrain>
__global__ void __launch_bounds__(kAThreads)
aush_disc_kernel(const float* __restrict__ Dp, const float* __restrict__ Dt, AushLayout lay, AushBatch bt, int nv, float inv_B,
                 float* __restrict__ bce0, float* __restrict__ bce1, float* __restrict__ dz1g, float* __restrict__ partial,
                 int64_t partial_stride) {
  extern __shared__ __align__(16) float smem[];
  float* WA = smem;                                   // [150][152] each
  float* WB = WA + kAH * kAHP;
  float* a1 = WB + kAH * kAHP;                        // [152][8] each
  float* a2 = a1 + kAHP * kAR;
  float* dzA = a2 + kAHP * kAR;
  float* dzB = dzA + kAHP * kAR;
  float* red = dzB + kAHP * kAR;                      // [5 warps][8] + dz4[8]
  const int E = bt.F + bt.S;
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 56);
  int32_t* ec = reinterpret_cast<int32_t*>(red + 64);  // [8][E]
  float* ev = reinterpret_cast<float*>(ec + kAR * E);  // [8][E]
  const uint32_t barA = smem_u32(bars), barB = smem_u32(bars + 1);
  constexpr uint32_t kWBytes = kAH * kAHP * sizeof(float);
  uint32_t phA = 0, phB = 0;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool unit = t < kAH;
  const float* W1t = Dp + lay.W1t;
  const float* W2t = Dt;
  const float* W3t = Dt + kAH * kAHP;
  float* gp = kTrain ? partial + (int64_t)blockIdx.x * partial_stride : nullptr;   // layout = Dp + lay.b1 onwards, then [S][152]
  const int64_t o_b1 = 0, o_W2 = lay.W2 - lay.b1, o_b2 = lay.b2 - lay.b1, o_W3 = lay.W3 - lay.b1, o_b3 = lay.b3 - lay.b1,
                o_w4 = lay.w4 - lay.b1, o_b4 = lay.b4 - lay.b1, o_sel = lay.total - lay.b1;
  const float b1 = t < kAHP ? Dp[lay.b1 + t] : 0.f, b2 = t < kAHP ? Dp[lay.b2 + t] : 0.f, b3 = t < kAHP ? Dp[lay.b3 + t] : 0.f;
  const float w4 = t < kAHP ? Dp[lay.w4 + t] : 0.f, b4 = Dp[lay.b4];
  const int tc = unit ? t : 0;                        // pad threads read a valid column and discard the result
  bool first = true;
  if (t == 0) {
    mbar_init(barA, 1);
    mbar_init(barB, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  for (int g0 = blockIdx.x * kAR; g0 < nv; g0 += gridDim.x * kAR) {
    __syncthreads();
    if (t == 0 && (kTrain || first)) {                // forward pair (the forward-only kernel keeps it for every pass)
      bulk_g2s(WA, W2t, kWBytes, barA);
      bulk_g2s(WB, W3t, kWBytes, barB);
    }
    // the pass's sparse inputs -> shared
    for (int idx = t; idx < kAR * E; idx += kAThreads) {
      const int r = idx / E, e = idx - r * E, v = g0 + r;
      int c = 0;
      float x = 0.f;
      if (v < nv) {
        const bool fake = !kTrain || v >= bt.B;
        const int b = kTrain && v >= bt.B ? v - bt.B : v;
        if (e < bt.F) { c = bt.cols[(int64_t)b * bt.F + e]; x = bt.dval[(int64_t)b * bt.F + e]; }
        else { c = bt.sel[e - bt.F]; x = fake ? bt.fin[(int64_t)b * bt.S + e - bt.F] : bt.rsel[(int64_t)b * bt.S + e - bt.F]; }
      }
      ec[idx] = c;
      ev[idx] = x;
    }
    __syncthreads();
    // layer 1: gather-sum of W1^T rows (a zero-valued entry points at column 0: the load is harmless, its weight is 0)
    float acc[kAR], av1[kAR], av2[kAR], av3[kAR];
    {
      float s0[kAR], s1[kAR];
#pragma unroll
      for (int r = 0; r < kAR; ++r) s0[r] = s1[r] = 0.f;
      int e = 0;
      for (; e + 4 <= E; e += 4) {                      // 32 independent gathers per step
        float w[kAR][4];
#pragma unroll
        for (int r = 0; r < kAR; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) w[r][q] = __ldg(W1t + (int64_t)ec[r * E + e + q] * kAHP + tc);
#pragma unroll
        for (int r = 0; r < kAR; ++r) {
          s0[r] = fmaf(ev[r * E + e], w[r][0], s0[r]);
          s1[r] = fmaf(ev[r * E + e + 1], w[r][1], s1[r]);
          s0[r] = fmaf(ev[r * E + e + 2], w[r][2], s0[r]);
          s1[r] = fmaf(ev[r * E + e + 3], w[r][3], s1[r]);
        }
      }
      for (; e < E; ++e) {
        float w[kAR];
#pragma unroll
        for (int r = 0; r < kAR; ++r) w[r] = __ldg(W1t + (int64_t)ec[r * E + e] * kAHP + tc);
#pragma unroll
        for (int r = 0; r < kAR; ++r) s0[r] = fmaf(ev[r * E + e], w[r], s0[r]);
      }
#pragma unroll
      for (int r = 0; r < kAR; ++r) acc[r] = b1 + (s0[r] + s1[r]);
    }
#pragma unroll
    for (int r = 0; r < kAR; ++r) av1[r] = unit ? sigmoidf_(acc[r]) : 0.f;
    if (t < kAHP) store8(a1 + t * kAR, av1);
    __syncthreads();
    if (kTrain || first) { mbar_wait(barA, phA); phA ^= 1; }
    dense_fwd(WA, a1, tc, b2, acc);
#pragma unroll
    for (int r = 0; r < kAR; ++r) av2[r] = unit ? sigmoidf_(acc[r]) : 0.f;
    if (t < kAHP) store8(a2 + t * kAR, av2);
    __syncthreads();                                   // a2 visible; WA released
    if (kTrain && t == 0) bulk_g2s(WA, Dp + lay.W3, kWBytes, barA);
    if (kTrain || first) { mbar_wait(barB, phB); phB ^= 1; }
    dense_fwd(WB, a2, tc, b3, acc);
#pragma unroll
    for (int r = 0; r < kAR; ++r) av3[r] = unit ? sigmoidf_(acc[r]) : 0.f;
    // output unit: z4[r] = b4 + sum_j w4[j] a3[j][r]
#pragma unroll
    for (int r = 0; r < kAR; ++r) {
      const float s = warp_sum(w4 * av3[r]);
      if (lane == 0) red[warp * kAR + r] = s;
    }
    __syncthreads();                                   // WB released
    if (kTrain && t == 0) bulk_g2s(WB, Dp + lay.W2, kWBytes, barB);
    if (!kTrain) first = false;
    if (t < kAR) {
      float z = b4;
      for (int w = 0; w < kAThreads / 32; ++w) z += red[w * kAR + t];
      const float p = sigmoidf_(z);
      const int v = g0 + t;
      float dz = 0.f;
      if (v < nv) {
        const float y = kTrain ? (v < bt.B ? 1.f : 0.f) : 1.f;              // valid_labels / fake_labels (aush.py:86-97)
        // nn.BCELoss: -(y log p + (1 - y) log(1 - p)), logs clamped at -100
        const float l = y != 0.f ? -fmaxf(logf(p), -100.f) : -fmaxf(logf(1.f - p), -100.f);
        if (kTrain && v >= bt.B) bce1[v - bt.B] = l; else bce0[v] = l;
        dz = (p - y) * (0.5f * inv_B);                                      // d_loss = (D_real_loss + D_fake_loss) / 2, means over B
      }
      red[48 + t] = dz;
    }
    if (!kTrain) continue;
    __syncthreads();
    // ---- backward
    float dz4[kAR], d3[kAR], da[kAR];
    load8(red + 48, dz4);
    {
      float gw4 = 0.f, gb4 = 0.f, gb3 = 0.f;
#pragma unroll
      for (int r = 0; r < kAR; ++r) {
        gw4 = fmaf(dz4[r], av3[r], gw4);
        gb4 += dz4[r];
        d3[r] = dz4[r] * w4 * (av3[r] * (1.f - av3[r]));
        gb3 += d3[r];
      }
      if (t < kAHP) {
        store8(dzA + t * kAR, d3);
        gp[o_w4 + t] = first ? gw4 : gp[o_w4 + t] + gw4;
        gp[o_b3 + t] = first ? gb3 : gp[o_b3 + t] + gb3;
      }
      if (t == 0) gp[o_b4] = first ? gb4 : gp[o_b4] + gb4;
    }
    __syncthreads();
    // layer 3: weight gradient against a2, d a2
    float d2[kAR], gb = 0.f;
    mbar_wait(barA, phA); phA ^= 1;
    if (unit) dense_bwd(WA, dzA, av2, t, gp + o_W3, first, da);
#pragma unroll
    for (int r = 0; r < kAR; ++r) { d2[r] = unit ? da[r] * (av2[r] * (1.f - av2[r])) : 0.f; gb += d2[r]; }
    if (t < kAHP) {
      store8(dzB + t * kAR, d2);
      gp[o_b2 + t] = first ? gb : gp[o_b2 + t] + gb;
    }
    __syncthreads();
    float d1[kAR];
    gb = 0.f;
    mbar_wait(barB, phB); phB ^= 1;
    if (unit) dense_bwd(WB, dzB, av1, t, gp + o_W2, first, da);
#pragma unroll
    for (int r = 0; r < kAR; ++r) { d1[r] = unit ? da[r] * (av1[r] * (1.f - av1[r])) : 0.f; gb += d1[r]; }
    if (t < kAHP) {
      gp[o_b1 + t] = first ? gb : gp[o_b1 + t] + gb;
#pragma unroll
      for (int r = 0; r < kAR; ++r)
        if (g0 + r < nv) dz1g[(int64_t)(g0 + r) * kAHP + t] = d1[r];
      // the selected columns are dense (every row holds them): their first-layer gradient rows are block partials too
      for (int s = 0; s < bt.S; ++s) {
        float g = 0.f;
#pragma unroll
        for (int r = 0; r < kAR; ++r) g = fmaf(ev[r * E + bt.F + s], d1[r], g);
        gp[o_sel + s * kAHP + t] = first ? g : gp[o_sel + s * kAHP + t] + g;
      }
    }
    first = false;
  }
}

// first-layer gradient from the batch's column index (filler entries) or the block partials (selected columns), fused with
// the dense Adam update of W1^T (one block per item row)
__global__ void __launch_bounds__(kAThreads)
aush_adam_w1_kernel(float* __restrict__ P, float* __restrict__ M, float* __restrict__ V, const int32_t* __restrict__ colptr,
                    const int32_t* __restrict__ ent, AushBatch bt, const float* __restrict__ dz1g, const float* __restrict__ partial,
                    int64_t partial_stride, int64_t o_sel, int n_blocks, AdamScalars a) {
  const int c = blockIdx.x, t = threadIdx.x;
  if (t >= kAHP) return;
  float g0 = 0.f, g1 = 0.f;
  const int beg = colptr[c], end = colptr[c + 1];
  int q = beg;
  for (; q + 2 <= end; q += 2) {
    const int c0 = ent[q], c1 = ent[q + 1], r0 = c0 / bt.F, r1 = c1 / bt.F;
    const float x0 = bt.dval[c0], x1 = bt.dval[c1];                       // ent = row * F + slot indexes dval directly
    const float d0 = dz1g[(int64_t)r0 * kAHP + t] + dz1g[(int64_t)(bt.B + r0) * kAHP + t];
    const float d1 = dz1g[(int64_t)r1 * kAHP + t] + dz1g[(int64_t)(bt.B + r1) * kAHP + t];
    g0 = fmaf(x0, d0, g0);
    g1 = fmaf(x1, d1, g1);
  }
  if (q < end) {
    const int c0 = ent[q], r0 = c0 / bt.F;
    g0 = fmaf(bt.dval[c0], dz1g[(int64_t)r0 * kAHP + t] + dz1g[(int64_t)(bt.B + r0) * kAHP + t], g0);
  }
  float g = g0 + g1;
  for (int s = 0; s < bt.S; ++s)
    if (bt.sel[s] == c) {
      float gs = 0.f;
      for (int b = 0; b < n_blocks; ++b) gs += partial[(int64_t)b * partial_stride + o_sel + s * kAHP + t];
      g += gs;
    }
  const int64_t i = (int64_t)c * kAHP + t;
  float p = P[i], m = M[i], v = V[i];
  adam1(p, g, m, v, a);
  P[i] = p; M[i] = m; V[i] = v;
}

// the other parameters: block partials summed in block order + Adam; the two 150 x 150 matrices also refresh their
// transposed copies (Dt) for the next forward
__global__ void __launch_bounds__(256)
aush_adam_small_kernel(float* __restrict__ P, float* __restrict__ M, float* __restrict__ V, const float* __restrict__ partial,
                       int64_t partial_stride, int n_blocks, int64_t n, int64_t o_W2, int64_t o_W3, int64_t o_b4,
                       float* __restrict__ Dt, AdamScalars a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || i > o_b4) return;                                                  // the floats after main.6.bias are padding
  const int64_t w = i >= o_W3 ? i - o_W3 : i - o_W2;
  const bool mat = i >= o_W2 && w < (int64_t)kAH * kAHP;
  const int j = mat ? (int)(w / kAHP) : 0, k = mat ? (int)(w - (int64_t)j * kAHP) : 0;
  if (k >= kAH) return;                                                            // row padding: stays zero
  float g0 = 0.f, g1 = 0.f;
  int b = 0;
  for (; b + 2 <= n_blocks; b += 2) {
    g0 += partial[(int64_t)b * partial_stride + i];
    g1 += partial[(int64_t)(b + 1) * partial_stride + i];
  }
  if (b < n_blocks) g0 += partial[(int64_t)b * partial_stride + i];
  float p = P[i], m = M[i], v = V[i];
  adam1(p, g0 + g1, m, v, a);
  P[i] = p; M[i] = m; V[i] = v;
  if (mat) Dt[(i >= o_W3 ? kAH * kAHP : 0) + k * kAHP + j] = p;
}

// Dt = transposed copies of main.2.weight and main.4.weight (start of an epoch: the parameters may have been loaded)
__global__ void aush_transpose_kernel(const float* __restrict__ Dp, AushLayout lay, float* __restrict__ Dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * kAH * kAHP) return;
  const int l = i / (kAH * kAHP), w = i - l * kAH * kAHP, k = w / kAHP, j = w - k * kAHP;
  Dt[i] = j < kAH ? Dp[(l ? lay.W3 : lay.W2) + (int64_t)j * kAHP + k] : 0.f;
}

// the four values train_step returns (aush.py:171-176): means over the batches of per-batch means
__global__ void __launch_bounds__(256)
aush_loss_kernel(const float* __restrict__ bce_real, const float* __restrict__ bce_fake, const float* __restrict__ bce_gan,
                 const float* __restrict__ shil, const float* __restrict__ rec, int64_t n_rows, int batch, int64_t n_items,
                 double* __restrict__ out) {
  __shared__ double sh[4][8];
  const int64_t nb = (n_rows + batch - 1) / batch;
  double tot[4] = {0, 0, 0, 0};
  for (int64_t k = 0; k < nb; ++k) {
    const int64_t lo = k * batch, hi = min(n_rows, lo + (int64_t)batch);
    double s[4] = {0, 0, 0, 0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      s[0] += 0.5 * ((double)bce_real[i] + (double)bce_fake[i]);
      s[1] += rec[i];
      s[2] += shil[i];
      s[3] += bce_gan[i];
    }
    for (int q = 0; q < 4; ++q) {
      const double w = warp_sum(s[q]);
      if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const double Bk = (double)(hi - lo);
      for (int q = 0; q < 4; ++q) {
        double w = 0;
        for (int i = 0; i < 8; ++i) w += sh[q][i];
        tot[q] += (q == 1 || q == 2) ? w / (Bk * (double)n_items) : w / Bk;      // MSELoss means over B * I elements
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int q = 0; q < 4; ++q) out[q] = nb ? tot[q] / (double)nb : nan("");
}

static size_t disc_smem(int E) { return sizeof(float) * (size_t)(2 * kAH * kAHP + 4 * kAHP * kAR + 64) + (size_t)kAR * E * 8; }
static int64_t aush_small(const AushLayout& o, int S) { return o.total - o.b1 + (int64_t)S * kAHP; }   // floats of one block partial

}  // namespace recad

using namespace recad;

extern "C" {

int recad_aush_d_layout(int64_t n_items, int64_t* offsets) {
  RECAD_REQUIRE(offsets && n_items > 0, RECAD_ERR_ARG, "aush_d_layout: bad argument");
  const AushLayout o = aush_layout(n_items);
  const int64_t v[9] = {o.W1t, o.b1, o.W2, o.b2, o.W3, o.b3, o.w4, o.b4, o.total};
  memcpy(offsets, v, sizeof(v));
  return RECAD_OK;
}

int64_t recad_aush_work_floats(int64_t n_items, int64_t n_rows, int32_t batch, int32_t n_sel) {
  const AushLayout o = aush_layout(n_items);
  return 5 * n_rows + (int64_t)batch * std::max(n_sel, 1) + 8 + 2 * (int64_t)batch * kAHP + 2 * kAH * kAHP +
         (int64_t)2 * sm_count() * aush_small(o, n_sel);
}

int recad_aush_plan_columns(const int32_t* cols, int64_t n_rows, int32_t batch, int32_t F, int64_t n_items, int32_t* colptr,
                            int32_t* ent) {
  RECAD_REQUIRE(cols && colptr && ent && n_rows >= 0 && batch > 0 && F >= 0 && n_items > 0, RECAD_ERR_ARG,
                "aush_plan_columns: bad argument");
  const int64_t nb = (n_rows + batch - 1) / batch;
  std::vector<int32_t> cur((size_t)n_items + 1);
  for (int64_t k = 0; k < nb; ++k) {
    const int64_t lo = k * batch, Bk = std::min<int64_t>(batch, n_rows - lo);
    int32_t* cp = colptr + k * (n_items + 1);
    std::fill(cp, cp + n_items + 1, 0);
    for (int64_t r = 0; r < Bk; ++r) {
      for (int e = 0; e < F; ++e) {
        const int32_t c = cols[(lo + r) * F + e];
        RECAD_REQUIRE(c >= 0 && c < n_items, RECAD_ERR_ARG, "aush_plan_columns: column %d out of range", c);
        ++cp[c + 1];
      }
    }
    for (int64_t c = 0; c < n_items; ++c) cp[c + 1] += cp[c];
    std::copy(cp, cp + n_items, cur.begin());
    int32_t* out = ent + lo * F;
    for (int64_t r = 0; r < Bk; ++r)
      for (int e = 0; e < F; ++e) out[cur[cols[(lo + r) * F + e]]++] = (int32_t)(r * F + e);
  }
  return RECAD_OK;
}

static int aush_check(const recad_aush* st, const char* who) {
  RECAD_REQUIRE(st && st->n_items > 0 && st->n_sel >= 0 && st->n_sel <= 64 && st->filler_num >= 0 && st->G_W1t && st->G_b1 &&
                    st->G_W2 && st->G_b2 && (st->n_sel == 0 || st->selected),
                RECAD_ERR_ARG, "%s: bad state", who);
  return RECAD_OK;
}

int recad_aush_generate(const recad_aush* st, const int32_t* cols, const float* tval, int64_t n_rows, float* gen_out, void* stream) {
  int rc;
  if ((rc = aush_check(st, "aush_generate"))) return rc;
  RECAD_REQUIRE(cols && tval && n_rows >= 0 && (gen_out || n_rows == 0 || st->n_sel == 0), RECAD_ERR_ARG,
                "aush_generate: bad argument");      // (an empty [n_rows, 0] output has no address)
  if (n_rows == 0 || st->n_sel == 0) return RECAD_OK;
  aush_gen_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, as_stream(stream)>>>(st->G_W1t, st->G_b1, st->G_W2, st->G_b2, st->selected,
                                                                               st->n_sel, cols, tval, st->filler_num, nullptr, nullptr,
                                                                               nullptr, n_rows, gen_out, nullptr, nullptr, nullptr);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

int recad_aush_train_epoch(const recad_aush* st, const recad_aush_epoch* ep, int64_t step0, double* loss_out, void* stream) {
  int rc;
  if ((rc = aush_check(st, "aush_train_epoch"))) return rc;
  RECAD_REQUIRE(st->D && st->Dm && st->Dv && st->work && ep && loss_out && ep->n_rows >= 0 && ep->batch > 0 && step0 >= 0,
                RECAD_ERR_ARG, "aush_train_epoch: bad argument");
  cudaStream_t s = as_stream(stream);
  const int64_t n = ep->n_rows, I = st->n_items;
  const int F = st->filler_num, S = st->n_sel, E = F + S, batch = ep->batch;
  RECAD_REQUIRE(n == 0 || ((F == 0 || (ep->cols && ep->tval && ep->dval && ep->ent)) && ep->colptr &&
                           (S == 0 || (ep->rsel && ep->tsel && ep->msel && ep->zr))),
                RECAD_ERR_ARG, "aush_train_epoch: missing epoch array");
  const AushLayout lay = aush_layout(I);
  const int64_t small = aush_small(lay, S), n_adam = lay.total - lay.b1;
  const int sms = sm_count();
  float* w = st->work;
  float* bce_real = w; w += n;
  float* bce_fake = w; w += n;
  float* bce_gan = w; w += n;
  float* shil = w; w += n;
  float* rec = w; w += n;
  float* fin = w; w += (int64_t)batch * std::max(S, 1);
  w += (4 - ((w - st->work) & 3)) & 3;
  float* dz1 = w; w += 2 * (int64_t)batch * kAHP;
  float* Dt = w; w += 2 * kAH * kAHP;
  float* partial = w;                                 // [<= 2 * sms][small]
  const size_t smem = disc_smem(E);
  RECAD_REQUIRE(smem <= 226 * 1024, RECAD_ERR_ARG, "aush_train_epoch: filler_num + |selected| = %d is too large", E);
  static bool attr = false;
  if (!attr) {
    RECAD_CUDA_CHECK(cudaFuncSetAttribute(aush_disc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    RECAD_CUDA_CHECK(cudaFuncSetAttribute(aush_disc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr = true;
  }
  const int64_t nb = (n + batch - 1) / batch;
  if (nb) {
    aush_transpose_kernel<<<(2 * kAH * kAHP + 255) / 256, 256, 0, s>>>(st->D, lay, Dt);
    RECAD_LAUNCH_CHECK();
  }
  for (int64_t k = 0; k < nb; ++k) {
    const int64_t lo = k * batch;
    const int B = (int)std::min<int64_t>(batch, n - lo);
    AushBatch bt{ep->cols + lo * F, ep->dval + lo * F, ep->rsel + lo * S, fin, st->selected, B, F, S};
    aush_gen_kernel<<<(unsigned)((B + 7) / 8), 256, 0, s>>>(st->G_W1t, st->G_b1, st->G_W2, st->G_b2, st->selected, S, bt.cols,
                                                             ep->tval + lo * F, F, ep->tsel + lo * S, ep->msel + lo * S, ep->zr + lo * S,
                                                             B, nullptr, fin, shil + lo, rec + lo);
    RECAD_LAUNCH_CHECK();
    const int g_train = std::min(sms, (2 * B + kAR - 1) / kAR), g_fwd = std::min(sms, (B + kAR - 1) / kAR);
    aush_disc_kernel<true><<<g_train, kAThreads, smem, s>>>(st->D, Dt, lay, bt, 2 * B, 1.f / (float)B, bce_real + lo, bce_fake + lo,
                                                            dz1, partial, small);
    RECAD_LAUNCH_CHECK();
    const AdamScalars a = adam_scalars(st->lr, st->beta1, st->beta2, st->eps, step0 + k + 1);
    aush_adam_w1_kernel<<<(unsigned)I, kAThreads, 0, s>>>(st->D + lay.W1t, st->Dm + lay.W1t, st->Dv + lay.W1t,
                                                          ep->colptr + k * (I + 1), ep->ent + lo * F, bt, dz1, partial, small,
                                                          n_adam, g_train, a);
    RECAD_LAUNCH_CHECK();
    aush_adam_small_kernel<<<(unsigned)((n_adam + 255) / 256), 256, 0, s>>>(st->D + lay.b1, st->Dm + lay.b1, st->Dv + lay.b1, partial,
                                                                             small, g_train, n_adam, lay.W2 - lay.b1, lay.W3 - lay.b1,
                                                                             lay.b4 - lay.b1, Dt, a);
    RECAD_LAUNCH_CHECK();
    aush_disc_kernel<false><<<g_fwd, kAThreads, smem, s>>>(st->D, Dt, lay, bt, B, 1.f / (float)B, bce_gan + lo, nullptr, nullptr, nullptr, 0);
    RECAD_LAUNCH_CHECK();
  }
  aush_loss_kernel<<<1, 256, 0, s>>>(bce_real, bce_fake, bce_gan, shil, rec, n, batch, I, loss_out);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

}  // extern "C"
