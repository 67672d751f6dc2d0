// LightGCN BPR step: fused gather + dot + softplus + L2-reg + gradient scatter, dense Adam, and the
// per-epoch driver that strings  propagate -> BPR -> Horner backward -> Adam  for every batch on one
// stream with no host synchronisation (reference: recad/model/victim/lightgcn.py:122-172).
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace recad {

// ------------------------------------------------------------------------------------------
// BPR forward + backward.  A group of LPR lanes owns one sample; each lane holds VPL float4 columns
// of the six rows (O_u, O_p, O_n, E_u, E_p, E_n).  Gradients leave through 128-bit vector REDs.
// ------------------------------------------------------------------------------------------
// Where a batch's (user, pos, neg) triples come from.
// RowSamples: rows [*, 3] in sampler order, visited through the epoch permutation (single GPU, reference layout).
template <typename IdxT>
struct RowSamples {
  static constexpr bool kSparse = false;     // every sample of the batch is processed
  const IdxT* samples;
  const IdxT* perm;
  // returns 1 = valid sample, 0 = nothing to do, -1 = id out of range
  __device__ __forceinline__ int fetch(int64_t b, int64_t n_users, int64_t n_items, int64_t& u, int64_t& p, int64_t& n) const {
    const int64_t row = perm ? (int64_t)perm[b] : b;      // the epoch shuffle is this indirection
    u = samples[3 * row]; p = samples[3 * row + 1]; n = samples[3 * row + 2];
    return (u < 0 || u >= n_users || p < 0 || p >= n_items || n < 0 || n >= n_items) ? -1 : 1;
  }
};
// ShardSamples: the GLOBAL epoch as the host sampler emits it (user, index of the positive inside the user's row,
// negative; 32 bit each) + the global permutation.  A rank of the user-sharded path walks the whole global batch and
// keeps the samples whose user it owns -- no routing pass, no per-rank copy of the epoch; the positive ITEM comes out
// of the rank's own rows of the interaction matrix (its users' sorted distinct items, implicit.py:339-343).
struct ShardSamples {
  static constexpr bool kSparse = true;      // only the samples of the rank's own users are processed
  const uint32_t* users;
  const uint32_t* rel;
  const uint32_t* negs;
  const int32_t* perm;
  int64_t user_lo, user_hi;        // global ids [lo, hi) live on this rank as local users 0 .. hi - lo
  const int64_t* pos_rowptr;       // [n_users_local + 1]
  const int32_t* pos_col;          // item ids + pos_col_offset
  int32_t pos_col_offset;
  __device__ __forceinline__ int fetch(int64_t b, int64_t n_users, int64_t n_items, int64_t& u, int64_t& p, int64_t& n) const {
    const int64_t row = perm ? (int64_t)perm[b] : b;
    const int64_t gu = users[row];
    if (gu < user_lo || gu >= user_hi) return 0;
    u = gu - user_lo;
    const int64_t lo = pos_rowptr[u], len = pos_rowptr[u + 1] - lo;
    const int64_t r = rel[row];
    n = negs[row];
    if (u >= n_users || r >= len || n >= n_items) return -1;
    p = (int64_t)pos_col[lo + r] - pos_col_offset;
    return (p < 0 || p >= n_items) ? -1 : 1;
  }
};

// ShardRows: the global epoch as (user, pos, neg) rows; a rank keeps its own users' rows (tests, injected datasets).
struct ShardRows {
  static constexpr bool kSparse = true;
  const int64_t* rows;
  const int64_t* perm;
  int64_t user_lo, user_hi;
  __device__ __forceinline__ int fetch(int64_t b, int64_t n_users, int64_t n_items, int64_t& u, int64_t& p, int64_t& n) const {
    const int64_t row = perm ? perm[b] : b;
    const int64_t gu = rows[3 * row];
    if (gu < user_lo || gu >= user_hi) return 0;
    u = gu - user_lo; p = rows[3 * row + 1]; n = rows[3 * row + 2];
    return (u >= n_users || p < 0 || p >= n_items || n < 0 || n >= n_items) ? -1 : 1;
  }
};

template <int LPR, int VPL, typename Src>
__global__ void __launch_bounds__(256)
bpr_kernel(const float* __restrict__ O, const float* __restrict__ E, int64_t n_users, int64_t n_items,
           const Src src, int64_t b0,
           int64_t B, int64_t B_norm, float grad_scale, float* __restrict__ gO, float* __restrict__ cnt,
           double* __restrict__ loss_acc, int nvec, int* __restrict__ bad) {
  constexpr int GPW = 32 / LPR;  // sample groups per warp
  const int lane = threadIdx.x & 31;
  const int l = lane % LPR;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / LPR;
  const float inv_B = 1.0f / (float)B_norm;   // the GLOBAL batch size: a rank of a sharded run sees only its users' rows
  float sp_sum = 0.f, sq_sum = 0.f;
  const float4* __restrict__ O4 = reinterpret_cast<const float4*>(O);
  const float4* __restrict__ E4 = reinterpret_cast<const float4*>(E);
  auto process = [&](const bool valid, const int64_t u, const int64_t p, const int64_t n) {
    const int64_t ru = u * nvec, rp = (n_users + p) * nvec, rn = (n_users + n) * nvec;
    float4 ou[VPL], op[VPL], on[VPL];
    float dn = 0.f, dp = 0.f, sq = 0.f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int c = l + q * LPR;
      if (valid && c < nvec) {
        ou[q] = __ldg(O4 + ru + c); op[q] = __ldg(O4 + rp + c); on[q] = __ldg(O4 + rn + c);
        const float4 eu = __ldg(E4 + ru + c), ep = __ldg(E4 + rp + c), en = __ldg(E4 + rn + c);
        dn += ou[q].x * on[q].x + ou[q].y * on[q].y + ou[q].z * on[q].z + ou[q].w * on[q].w;
        dp += ou[q].x * op[q].x + ou[q].y * op[q].y + ou[q].z * op[q].z + ou[q].w * op[q].w;
        sq += eu.x * eu.x + eu.y * eu.y + eu.z * eu.z + eu.w * eu.w + ep.x * ep.x + ep.y * ep.y + ep.z * ep.z +
              ep.w * ep.w + en.x * en.x + en.y * en.y + en.z * en.z + en.w * en.w;
      } else {
        ou[q] = op[q] = on[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      dn += __shfl_xor_sync(kFull, dn, o);
      dp += __shfl_xor_sync(kFull, dp, o);
    }
    // x = <u, n> - <u, p>;  softplus with torch's threshold 20;  d softplus = z / (z + 1)
    const float x = dn - dp;
    float sp, sig;
    if (x > 20.f) { sp = x; sig = 1.f; }
    else { const float z = expf(x); sp = log1pf(z); sig = z / (z + 1.f); }
    if (valid) {
      if (l == 0) sp_sum += sp;
      sq_sum += sq;
      const float s = sig * inv_B * grad_scale;
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int c = l + q * LPR;
        if (c < nvec) {
          const float4 a = ou[q], pp = op[q], nn = on[q];
          red_add4(gO + (ru + c) * 4, make_float4(s * (nn.x - pp.x), s * (nn.y - pp.y), s * (nn.z - pp.z), s * (nn.w - pp.w)));
          red_add4(gO + (rp + c) * 4, make_float4(-s * a.x, -s * a.y, -s * a.z, -s * a.w));
          red_add4(gO + (rn + c) * 4, make_float4(s * a.x, s * a.y, s * a.z, s * a.w));
        }
      }
      if (l == 0) {
        atomicAdd(cnt + u, 1.0f);
        atomicAdd(cnt + n_users + p, 1.0f);
        atomicAdd(cnt + n_users + n, 1.0f);
      }
    }
  };
  if constexpr (!Src::kSparse) {
    // all lanes of a warp iterate the same number of times (the shuffles inside are warp-wide)
    const int64_t iters = (B + n_groups - 1) / n_groups;
    for (int64_t it = 0; it < iters; ++it) {
      const int64_t b = it * n_groups + group;
      const bool valid = b < B;
      int64_t u = 0, p = 0, n = 0;
      if (valid && src.fetch(b0 + b, n_users, n_items, u, p, n) < 0) {
        if (l == 0) atomicOr(bad, 1);
        u = 0; p = 0; n = 0;
      }
      process(valid, u, p, n);
    }
  } else {
    // Sharded source: a warp looks at 32 samples of the GLOBAL batch at a time, one per lane (coalesced index loads = the
    // ownership test), then its GPW lane groups work through the samples the rank keeps, GPW at a time.  The ids of the
    // NEXT 32 samples are fetched before the current ones are processed.
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int grp = lane / LPR;
    auto fetch32 = [&](int64_t c0, int64_t& fu, int64_t& fp, int64_t& fn) -> int {
      int st = 0;
      fu = 0; fp = 0; fn = 0;
      if (c0 + lane < B) {
        st = src.fetch(b0 + c0 + lane, n_users, n_items, fu, fp, fn);
        if (st < 0) {
          atomicOr(bad, 1);
          fu = 0; fp = 0; fn = 0;
          st = 1;
        }
      }
      return st;
    };
    int64_t nu, np_, nn_;
    int nst = fetch32(warp_global * 32, nu, np_, nn_);
    for (int64_t c0 = warp_global * 32; c0 < B; c0 += n_warps * 32) {
      const int64_t mu = nu, mp = np_, mn = nn_;
      const int st = nst;
      nst = fetch32(c0 + n_warps * 32, nu, np_, nn_);
      unsigned todo = __ballot_sync(kFull, st != 0);
      while (todo) {
        unsigned m = todo;                       // lane group `grp` takes the grp-th lowest pending sample
        int from = -1;
#pragma unroll
        for (int k = 0; k < GPW; ++k) {
          const int low = m ? __ffs(m) - 1 : -1;
          if (k == grp) from = low;
          m &= m - 1;
        }
        todo = m;
        const bool valid = from >= 0;
        const int sl = valid ? from : 0;
        process(valid, __shfl_sync(kFull, mu, sl), __shfl_sync(kFull, mp, sl), __shfl_sync(kFull, mn, sl));
      }
    }
  }
  // block reduction of the two loss partials -> one double atomic each per block
  __shared__ double red[2][8];
  double a = warp_sum((double)sp_sum), c = warp_sum((double)sq_sum);
  const int w = threadIdx.x >> 5;
  if (lane == 0) { red[0][w] = a; red[1][w] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tc = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { ta += red[0][k]; tc += red[1][k]; }
    atomicAdd(loss_acc + 0, ta);
    atomicAdd(loss_acc + 1, tc);
  }
  (void)GPW;
}

// ------------------------------------------------------------------------------------------
// dense Adam (+ optional L2-reg gradient from batch multiplicities, + optional loss fold)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, const float* __restrict__ cnt, float reg_scale,
            float* __restrict__ m, float* __restrict__ v, int64_t n4, int vec_per_row, AdamScalars a, LossFold fold) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    if (cnt) {
      const float r = reg_scale * cnt[i / vec_per_row];
      G.x = fmaf(r, P.x, G.x); G.y = fmaf(r, P.y, G.y); G.z = fmaf(r, P.z, G.z); G.w = fmaf(r, P.w, G.w);
    }
    adam_update(P.x, G.x, M.x, V.x, a);
    adam_update(P.y, G.y, M.y, V.y, a);
    adam_update(P.z, G.z, M.z, V.z, a);
    adam_update(P.w, G.w, M.w, V.w, a);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  if (fold.acc && blockIdx.x == 0 && threadIdx.x == 0) {
    // final_loss of this batch = mean softplus + lambda * 0.5 * sum sq / B  (lightgcn.py:149-165)
    fold.acc[2] += fold.acc[0] * fold.inv_B + fold.half_lambda * fold.acc[1] * fold.inv_B;
    fold.acc[0] = 0.0;
    fold.acc[1] = 0.0;
  }
}

// scalar tail for n % 4 != 0 (bias vectors)
__global__ void adam_tail_kernel(float* p, const float* g, float* m, float* v, int64_t lo, int64_t n, AdamScalars a) {
  int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float G = g[i], M = m[i], V = v[i], P = p[i];
  adam_update(P, G, M, V, a);
  p[i] = P; m[i] = M; v[i] = V;
}

AdamScalars adam_scalars(float lr, float b1, float b2, float eps, int64_t step) {
  // torch.optim.adam._single_tensor_adam: python-double scalars, cast to fp32 at the tensor op
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  AdamScalars a;
  a.w1 = (float)(1.0 - (double)b1);
  a.b2 = b2;
  a.w2 = (float)(1.0 - (double)b2);
  a.step_size = (float)((double)lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.eps = eps;
  return a;
}

int launch_adam(float* p, const float* g, const float* cnt, float reg_scale, float* m, float* v, int64_t n, int D,
                const AdamScalars& a, const LossFold& fold, cudaStream_t s) {
  const int64_t n4 = n / 4;
  if (n4 > 0) {
    const int64_t want = (n4 + 255) / 256;
    const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sm_count() * 16);
    adam_kernel<<<grid, 256, 0, s>>>(p, g, cnt, reg_scale, m, v, n4, cnt ? D / 4 : 1, a, fold);
    RECAD_LAUNCH_CHECK();
  }
  if (n % 4) {
    adam_tail_kernel<<<1, 32, 0, s>>>(p, g, m, v, n4 * 4, n, a);
    RECAD_LAUNCH_CHECK();
  }
  return RECAD_OK;
}

template <int LPR, int VPL, typename Src>
static int launch_bpr_t(const float* O, const float* E, int64_t U, int64_t I, const Src& src, int64_t b0, int64_t B, int64_t B_work,
                        int64_t B_norm, float gs, float* gO, float* cnt, double* loss, int nvec, int* bad, cudaStream_t s) {
  const int64_t groups_per_block = 256 / LPR;
  const int64_t want = (B_work + groups_per_block - 1) / groups_per_block;       // B_work: samples expected to be valid
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 32));
  bpr_kernel<LPR, VPL, Src><<<grid, 256, 0, s>>>(O, E, U, I, src, b0, B, B_norm, gs, gO, cnt, loss, nvec, bad);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

template <typename Src>
static int launch_bpr_src(const float* O, const float* E, int64_t U, int64_t I, const Src& src, int64_t b0, int64_t B, int64_t B_work,
                          int64_t B_norm, float gs, float* gO, float* cnt, double* loss, int D, int* bad, cudaStream_t s) {
  const int nvec = D / 4;
  if (nvec <= 8) return launch_bpr_t<8, 1, Src>(O, E, U, I, src, b0, B, B_work, B_norm, gs, gO, cnt, loss, nvec, bad, s);
  if (nvec <= 16) return launch_bpr_t<16, 1, Src>(O, E, U, I, src, b0, B, B_work, B_norm, gs, gO, cnt, loss, nvec, bad, s);
  if (nvec <= 32) return launch_bpr_t<32, 1, Src>(O, E, U, I, src, b0, B, B_work, B_norm, gs, gO, cnt, loss, nvec, bad, s);
  if (nvec <= 64) return launch_bpr_t<32, 2, Src>(O, E, U, I, src, b0, B, B_work, B_norm, gs, gO, cnt, loss, nvec, bad, s);
  if (nvec <= 128) return launch_bpr_t<32, 4, Src>(O, E, U, I, src, b0, B, B_work, B_norm, gs, gO, cnt, loss, nvec, bad, s);
  return launch_bpr_t<32, 8, Src>(O, E, U, I, src, b0, B, B_work, B_norm, gs, gO, cnt, loss, nvec, bad, s);
}

template <typename IdxT>
static int launch_bpr(const float* O, const float* E, int64_t U, int64_t I, const IdxT* samples, const IdxT* perm,
               int64_t B, int64_t B_norm, float gs, float* gO, float* cnt, double* loss, int D, int* bad,
               cudaStream_t s) {
  const RowSamples<IdxT> src{samples, perm};
  return launch_bpr_src(O, E, U, I, src, 0, B, B, B_norm, gs, gO, cnt, loss, D, bad, s);
}

// sharded path (shard.cu): samples b0 .. b0 + B of the global epoch, of which about B / world belong to this rank
int launch_bpr_shard(const float* O, const float* E, int64_t U, int64_t I, const uint32_t* users, const uint32_t* rel,
                     const uint32_t* negs, const int32_t* perm, int64_t b0, int64_t B, int64_t B_norm, int64_t user_lo,
                     const int64_t* pos_rowptr, const int32_t* pos_col, int32_t pos_col_offset, int world, float gs, float* gO,
                     float* cnt, double* loss, int D, cudaStream_t s) {
  const ShardSamples src{users, rel, negs, perm, user_lo, user_lo + U, pos_rowptr, pos_col, pos_col_offset};
  // groups sized for ~4 global samples each: the ownership test is 12 bytes per sample, the work 4.6 kB per kept sample
  const int64_t B_work = std::max<int64_t>(B / std::max(1, std::min(world, 4)), 1);
  return launch_bpr_src(O, E, U, I, src, b0, B, B_work, B_norm, gs, gO, cnt, loss, D, reinterpret_cast<int*>(loss + 3), s);
}

int launch_bpr_shard_rows(const float* O, const float* E, int64_t U, int64_t I, const int64_t* rows, const int64_t* perm, int64_t b0,
                          int64_t B, int64_t B_norm, int64_t user_lo, int world, float gs, float* gO, float* cnt, double* loss, int D,
                          cudaStream_t s) {
  const ShardRows src{rows, perm, user_lo, user_lo + U};
  const int64_t B_work = std::max<int64_t>(B / std::max(1, std::min(world, 4)), 1);
  return launch_bpr_src(O, E, U, I, src, b0, B, B_work, B_norm, gs, gO, cnt, loss, D, reinterpret_cast<int*>(loss + 3), s);
}

// z = a * x + b * y (float4 grid-stride); z may alias x or y
__global__ void __launch_bounds__(256) axpby_kernel(float* z, float a, const float* x, float b, const float* y, int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 X = reinterpret_cast<const float4*>(x)[i], Y = reinterpret_cast<const float4*>(y)[i];
    reinterpret_cast<float4*>(z)[i] = make_float4(a * X.x + b * Y.x, a * X.y + b * Y.y, a * X.z + b * Y.z, a * X.w + b * Y.w);
  }
}

__global__ void dot_scores_kernel(const float* __restrict__ O, int64_t n_users, const int64_t* __restrict__ users,
                                  const int64_t* __restrict__ items, int64_t B, int nvec, float* __restrict__ out) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float4* __restrict__ O4 = reinterpret_cast<const float4*>(O);
  const int64_t ru = users[b] * nvec, ri = (n_users + items[b]) * nvec;
  float acc = 0.f;
  for (int c = lane; c < nvec; c += 32) {
    const float4 a = __ldg(O4 + ru + c), q = __ldg(O4 + ri + c);
    acc += a.x * q.x + a.y * q.y + a.z * q.z + a.w * q.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[b] = acc;
}

static int check_lightgcn(const recad_lightgcn* st) {
  RECAD_REQUIRE(st && st->graph, RECAD_ERR_ARG, "lightgcn: null state");
  RECAD_REQUIRE(st->n_users > 0 && st->n_items > 0 && st->graph->n_rows == st->n_users + st->n_items, RECAD_ERR_ARG,
                "lightgcn: graph has %lld rows, expected n_users + n_items = %lld", (long long)st->graph->n_rows,
                (long long)(st->n_users + st->n_items));
  RECAD_REQUIRE(st->D >= 4 && st->D % 4 == 0 && st->n_layers >= 0, RECAD_ERR_UNSUPPORTED, "lightgcn: bad D / layers");
  RECAD_REQUIRE(st->E && st->O && (st->n_layers < 2 || (st->X0 && st->X1)) && (st->n_layers < 1 || st->X0),
                RECAD_ERR_ARG, "lightgcn: null buffer");
  return RECAD_OK;
}

}  // namespace recad

using namespace recad;

extern "C" {

int recad_bpr_fwd_bwd(const float* O, const float* E, int64_t n_users, int64_t n_items, const int64_t* samples,
                      const int64_t* perm, int64_t B, int64_t B_norm, float grad_scale, float* gO, float* cnt,
                      double* loss_acc, int32_t D, void* stream) {
  RECAD_REQUIRE(O && E && samples && gO && cnt && loss_acc, RECAD_ERR_ARG, "bpr: null pointer");
  RECAD_REQUIRE(B >= 0 && B_norm >= B && B_norm > 0 && D >= 4 && D % 4 == 0 && D <= 1024, RECAD_ERR_UNSUPPORTED, "bpr: bad B or D");
  if (B == 0) return RECAD_OK;
  // loss_acc[3] doubles as the out-of-range flag (stays 0.0 when all ids are valid)
  return launch_bpr<int64_t>(O, E, n_users, n_items, samples, perm, B, B_norm, grad_scale, gO, cnt, loss_acc, D,
                             reinterpret_cast<int*>(loss_acc + 3), as_stream(stream));
}

int recad_adam(float* p, const float* g, const float* cnt, float reg_scale, float* m, float* v, int64_t n, int32_t D,
               float lr, float b1, float b2, float eps, int64_t step, void* stream) {
  RECAD_REQUIRE(p && g && m && v && n > 0 && step >= 1, RECAD_ERR_ARG, "adam: bad argument");
  RECAD_REQUIRE(!cnt || (D >= 4 && D % 4 == 0 && n % D == 0), RECAD_ERR_ARG, "adam: cnt needs n %% D == 0, D %% 4 == 0");
  LossFold none{nullptr, 0.0, 0.0};
  return launch_adam(p, g, cnt, reg_scale, m, v, n, D, adam_scalars(lr, b1, b2, eps, step), none, as_stream(stream));
}

int recad_lightgcn_propagate(const recad_lightgcn* st, void* stream) {
  int rc = check_lightgcn(st);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const int64_t N = st->n_users + st->n_items;
  const int L = st->n_layers;
  if (L == 0) {
    RECAD_CUDA_CHECK(cudaMemcpyAsync(st->O, st->E, N * st->D * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return RECAD_OK;
  }
  const float* x = st->E;
  for (int k = 0; k < L; ++k) {
    const bool last = k == L - 1;
    float* y = last ? nullptr : ((k & 1) ? st->X1 : st->X0);
    rc = recad_spmm(st->graph, x, y, k == 0 ? st->E : st->O, st->O, last ? 1.0f / (float)(L + 1) : 1.0f, st->D, stream);
    if (rc) return rc;
    x = y;
  }
  return RECAD_OK;
}

}  // extern "C"

template <typename IdxT>
static int lightgcn_train_epoch_t(const recad_lightgcn* st, const IdxT* samples, const IdxT* perm,
                                  int64_t n_samples, int64_t batch, int64_t step0, void* stream) {
  int rc = check_lightgcn(st);
  if (rc) return rc;
  RECAD_REQUIRE(st->m && st->v && st->g && st->cnt && st->loss_acc && st->X0 && st->X1, RECAD_ERR_ARG,
                "lightgcn_train_epoch: null training buffer");
  RECAD_REQUIRE(samples && n_samples > 0 && batch > 0 && step0 >= 0, RECAD_ERR_ARG, "lightgcn_train_epoch: bad samples");
  cudaStream_t s = as_stream(stream);
  const int64_t N = st->n_users + st->n_items;
  const int L = st->n_layers, D = st->D;
  RECAD_CUDA_CHECK(cudaMemsetAsync(st->loss_acc, 0, 4 * sizeof(double), s));
  int64_t step = step0;
  for (int64_t b0 = 0; b0 < n_samples; b0 += batch) {
    const int64_t B = min(batch, n_samples - b0);
    ++step;
    rc = recad_lightgcn_propagate(st, stream);
    if (rc) return rc;
    RECAD_CUDA_CHECK(cudaMemsetAsync(st->g, 0, N * D * sizeof(float), s));
    RECAD_CUDA_CHECK(cudaMemsetAsync(st->cnt, 0, N * sizeof(float), s));
    rc = launch_bpr(st->O, st->E, st->n_users, st->n_items, perm ? samples : samples + 3 * b0, perm ? perm + b0 : (const IdxT*)nullptr, B, B,
                    1.0f / (float)(L + 1), st->g, st->cnt, st->loss_acc, D, reinterpret_cast<int*>(st->loss_acc + 3), s);
    if (rc) return rc;
    // Horner: t <- g + A t, L times, so that t = (I + A + ... + A^L) g
    const float* t = st->g;
    for (int k = 0; k < L; ++k) {
      float* z = (k & 1) ? st->X1 : st->X0;
      rc = recad_spmm(st->graph_t ? st->graph_t : st->graph, t, nullptr, st->g, z, 1.0f, D, stream);
      if (rc) return rc;
      t = z;
    }
    LossFold fold{st->loss_acc, 1.0 / (double)B, 0.5 * (double)st->lambda};
    rc = launch_adam(st->E, t, st->cnt, st->lambda / (float)B, st->m, st->v, N * D, D,
                     adam_scalars(st->lr, st->beta1, st->beta2, st->eps, step), fold, s);
    if (rc) return rc;
  }
  return RECAD_OK;
}

extern "C" {

int recad_lightgcn_train_epoch(const recad_lightgcn* st, const int64_t* samples, const int64_t* perm,
                               int64_t n_samples, int64_t batch, int64_t step0, void* stream) {
  return lightgcn_train_epoch_t<int64_t>(st, samples, perm, n_samples, batch, step0, stream);
}

int recad_lightgcn_train_epoch_i32(const recad_lightgcn* st, const int32_t* samples, const int32_t* perm,
                                   int64_t n_samples, int64_t batch, int64_t step0, void* stream) {
  return lightgcn_train_epoch_t<int32_t>(st, samples, perm, n_samples, batch, step0, stream);
}

int recad_bpr_fwd_bwd_i32(const float* O, const float* E, int64_t n_users, int64_t n_items, const int32_t* samples,
                          const int32_t* perm, int64_t B, int64_t B_norm, float grad_scale, float* gO, float* cnt,
                          double* loss_acc, int32_t D, void* stream) {
  RECAD_REQUIRE(O && E && samples && gO && cnt && loss_acc, RECAD_ERR_ARG, "bpr: null pointer");
  RECAD_REQUIRE(B >= 0 && B_norm >= B && B_norm > 0 && D >= 4 && D % 4 == 0 && D <= 1024, RECAD_ERR_UNSUPPORTED, "bpr: bad B or D");
  if (B == 0) return RECAD_OK;
  return launch_bpr<int32_t>(O, E, n_users, n_items, samples, perm, B, B_norm, grad_scale, gO, cnt, loss_acc, D,
                             reinterpret_cast<int*>(loss_acc + 3), as_stream(stream));
}

}  // extern "C"

// rows[k] = (users[k], allpos_col[allpos_rowptr[users[k]] + rel[k]], negs[k])
__global__ void __launch_bounds__(256)
samples_expand_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const uint32_t* __restrict__ users,
                      const uint32_t* __restrict__ rel, const uint32_t* __restrict__ negs, int64_t n, int32_t* __restrict__ rows) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t u = users[k];
    rows[3 * k] = (int32_t)u;
    rows[3 * k + 1] = col[rowptr[u] + rel[k]];
    rows[3 * k + 2] = (int32_t)negs[k];
  }
}

extern "C" {

int recad_samples_expand(const int64_t* allpos_rowptr, const int32_t* allpos_col, const uint32_t* users, const uint32_t* rel,
                         const uint32_t* negs, int64_t n, int32_t* rows, void* stream) {
  RECAD_REQUIRE(n >= 0 && (n == 0 || (allpos_rowptr && allpos_col && users && rel && negs && rows)), RECAD_ERR_ARG,
                "samples_expand: bad argument");
  if (n == 0) return RECAD_OK;
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
  samples_expand_kernel<<<grid, 256, 0, as_stream(stream)>>>(allpos_rowptr, allpos_col, users, rel, negs, n, rows);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

int recad_axpby(float* z, float a, const float* x, float b, const float* y, int64_t n, void* stream) {
  RECAD_REQUIRE(z && x && y && n > 0 && n % 4 == 0, RECAD_ERR_ARG, "axpby: n must be a positive multiple of 4");
  const int64_t n4 = n / 4;
  const unsigned grid = (unsigned)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 16);
  axpby_kernel<<<grid, 256, 0, as_stream(stream)>>>(z, a, x, b, y, n4);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

int recad_dot_scores(const float* O, int64_t n_users, const int64_t* users, const int64_t* items, int64_t B, int32_t D,
                     float* scores, void* stream) {
  RECAD_REQUIRE(O && users && items && scores && B > 0 && D >= 4 && D % 4 == 0, RECAD_ERR_ARG, "dot_scores: bad argument");
  const int64_t threads = B * 32;
  dot_scores_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(O, n_users, users, items, B, D / 4, scores);
  RECAD_LAUNCH_CHECK();
  return RECAD_OK;
}

}  // extern "C"
