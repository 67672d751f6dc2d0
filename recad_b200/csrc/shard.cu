// User-sharded LightGCN epoch, driven from C on one stream per rank with NO collective library in the batch loop
// (SURVEY.md 8e scheme B; reference math: recad/model/victim/lightgcn.py:82-172).
//
// Rank g owns users [lo, hi): its rows of the normalised adjacency as two blocks (g_user: own user rows gathering the
// item replica; g_item: all item rows restricted to own users = PARTIAL item sums), its user rows of the table + Adam
// state, and a replica of the item rows.  Everything a peer must read or write lives in ONE NVLink-mapped symmetric
// block per rank (layer buffers X0 / X1, the propagated mean O, the gradient g, the batch multiplicities cnt, the
// staging slots, the barrier pad), so an exchange is plain loads / stores on peer pointers:
//   layer / Horner step:  item-row SpMM whose epilogue PUSHES each finished partial row into the owner's staging slot
//                         | ONE barrier | user-row SpMM with the layer-mean / Horner epilogue fused | owner adds the `world`
//                         partial copies of its slice in rank order, applies the same epilogue (mean / Horner) and stores
//                         the result into EVERY replica (one multimem.st through the NVSwitch multicast mapping, or one
//                         peer store per rank); those stores are certified by the NEXT exchange's barrier (staging slots
//                         are double buffered), a sequence of exchanges ends with one more barrier
//   gradient block:       BPR (each rank walks the global batch and keeps its users' samples) | barrier | owner PULLS
//                         its slice of every rank's partial gradient over NVLink, adds in rank order, stores the sum
//                         into every replica (same for the multiplicities)
// = 2 L + 3 barriers per batch.
// The barrier is a one-block kernel on a monotonically increasing epoch (st.release.sys to every peer's pad, ld.acquire.sys
// on the own pad), so the whole epoch is enqueued without touching the host.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace recad {

constexpr int kMaxShardPeers = 16;

struct PeerPads {
  uint32_t* pad[kMaxShardPeers];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// state[0] = barrier epoch of this rank, state[1] = set to 1 if a wait ever timed out (a peer died): the epoch then
// finishes with wrong numbers instead of hanging the device, and the host raises after the epoch.
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerPads pads, int rank, int world, uint32_t* state) {
  __shared__ uint32_t e_s;
  if (threadIdx.x == 0) {
    e_s = state[0] + 1;
    state[0] = e_s;
  }
  __syncwarp();
  const uint32_t e = e_s;
  if ((int)threadIdx.x < world) {
    __threadfence_system();                                   // everything this stream wrote before (kernel boundary) goes first
    st_release_sys(pads.pad[threadIdx.x] + rank, e);
    const uint32_t* mine = pads.pad[rank] + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > 20000000000LL) {                   // ~10 s
        state[1] = 1;
        break;
      }
    }
  }
}

struct ReduceArgs {
  const float* src[kMaxShardPeers];   // partial copies of the owner's slice: staging slots (push) or peer buffers (pull)
  float* out_y[kMaxShardPeers];       // raw sum goes here (every replica), n_y destinations
  float* out_z[kMaxShardPeers];       // alpha * (C + sum) goes here, n_z destinations
  const float* C;                     // local, same indexing; may be null (= 0)
  float alpha;
  int n_src, n_y, n_z;
  int64_t n;                          // elements (float4 for the vector kernel)
};

__device__ __forceinline__ void multimem_st4(float* p, float4 a) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
}

// kMcY / kMcZ: out_y[0] / out_z[0] is an NVSwitch multicast address (one store lands in every replica)
template <bool kMcY, bool kMcZ>
__global__ void __launch_bounds__(512) shard_reduce_kernel(const __grid_constant__ ReduceArgs a) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (int64_t)gridDim.x * blockDim.x) {
    float4 s = __ldcg(reinterpret_cast<const float4*>(a.src[0]) + e);    // written by peers: read at L2 / over NVLink, never via L1
    for (int r = 1; r < a.n_src; ++r) {
      const float4 p = __ldcg(reinterpret_cast<const float4*>(a.src[r]) + e);
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    if (a.n_y) {
      if (kMcY) multimem_st4(a.out_y[0] + 4 * e, s);
      else for (int r = 0; r < a.n_y; ++r) reinterpret_cast<float4*>(a.out_y[r])[e] = s;
    }
    if (a.n_z) {
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.C) c = reinterpret_cast<const float4*>(a.C)[e];
      const float4 z = make_float4(a.alpha * (c.x + s.x), a.alpha * (c.y + s.y), a.alpha * (c.z + s.z), a.alpha * (c.w + s.w));
      if (kMcZ) multimem_st4(a.out_z[0] + 4 * e, z);
      else for (int r = 0; r < a.n_z; ++r) reinterpret_cast<float4*>(a.out_z[r])[e] = z;
    }
  }
}

// scalar variant (the multiplicities: their item block starts at an arbitrary float offset)
__global__ void __launch_bounds__(256) shard_reduce_scalar_kernel(const __grid_constant__ ReduceArgs a) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (int64_t)gridDim.x * blockDim.x) {
    float s = __ldcg(a.src[0] + e);
    for (int r = 1; r < a.n_src; ++r) s += __ldcg(a.src[r] + e);
    for (int r = 0; r < a.n_y; ++r) a.out_y[r][e] = s;
  }
}

int launch_bpr_shard(const float* O, const float* E, int64_t U, int64_t I, const uint32_t* users, const uint32_t* rel,
                     const uint32_t* negs, const int32_t* perm, int64_t b0, int64_t B, int64_t B_norm, int64_t user_lo,
                     const int64_t* pos_rowptr, const int32_t* pos_col, int32_t pos_col_offset, int world, float gs, float* gO,
                     float* cnt, double* loss, int D, cudaStream_t s);
int launch_bpr_shard_rows(const float* O, const float* E, int64_t U, int64_t I, const int64_t* rows, const int64_t* perm, int64_t b0,
                          int64_t B, int64_t B_norm, int64_t user_lo, int world, float gs, float* gO, float* cnt, double* loss, int D,
                          cudaStream_t s);

namespace {

// per-device side stream (highest priority) + the two events that tie it to the caller's stream; created once
struct SideRes {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_go = nullptr, ev_done = nullptr;
};
SideRes* side_res() {
  static SideRes res[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideRes& r = res[dev];
  if (!r.side) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const bool low = getenv("RECAD_SHARD_SIDE_LOW") && atoi(getenv("RECAD_SHARD_SIDE_LOW"));   // experiment: reduce fills the SpMM's tail
    if (cudaStreamCreateWithPriority(&r.side, cudaStreamNonBlocking, low ? lo : hi) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&r.ev_go, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&r.ev_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &r;
}

struct Shard {
  const recad_lightgcn_shard* st;
  cudaStream_t s;
  int rank, world, D, L;
  int64_t Ug, I, slice, rows_mine;
  bool multicast;
  char* base[kMaxShardPeers];
  PeerPads pads;
  uint32_t* bstate;
  // phase trace
  std::vector<std::pair<int, cudaEvent_t>> marks;
  bool trace = false;
  SideRes* sr = nullptr;
  bool pending = false;                 // a slice reduce is in flight on the side stream

  float* buf(int r, int64_t off) const { return reinterpret_cast<float*>(base[r] + off); }
  float* mc(int64_t off) const { return reinterpret_cast<float*>(reinterpret_cast<char*>(st->mc_base) + off); }
  float* local(int64_t off) const { return buf(rank, off); }
  int64_t item_off(int r) const { return st->peer_users[r] * (int64_t)D; }      // floats: item block of rank r's [users ; items] buffer

  int init(const recad_lightgcn_shard* st_, void* stream) {
    st = st_;
    s = as_stream(stream);
    RECAD_REQUIRE(st && st->g_user && st->g_item && st->peer_base && st->peer_users, RECAD_ERR_ARG, "lightgcn_shard: null state");
    rank = st->rank; world = st->world; D = st->D; L = st->n_layers;
    Ug = st->n_users_local; I = st->n_items; slice = st->slice;
    RECAD_REQUIRE(world >= 1 && world <= kMaxShardPeers && rank >= 0 && rank < world, RECAD_ERR_ARG, "lightgcn_shard: bad rank / world");
    RECAD_REQUIRE(D == 32 || D == 64 || D == 128, RECAD_ERR_UNSUPPORTED, "lightgcn_shard: D = %d (supported: 32, 64, 128)", D);
    RECAD_REQUIRE(L >= 1 && Ug > 0 && I > 0 && slice > 0 && slice * world >= I, RECAD_ERR_ARG, "lightgcn_shard: bad sizes");
    RECAD_REQUIRE(st->g_user->n_rows == Ug && st->g_item->n_rows == I && st->peer_users[rank] == Ug, RECAD_ERR_ARG,
                  "lightgcn_shard: row blocks do not match the shard");
    RECAD_REQUIRE(st->E && st->m && st->v && st->loss_acc, RECAD_ERR_ARG, "lightgcn_shard: null table");
    rows_mine = std::max<int64_t>(0, std::min<int64_t>(slice, I - (int64_t)rank * slice));
    multicast = st->mc_base != nullptr;
    for (int r = 0; r < world; ++r) {
      RECAD_REQUIRE(st->peer_base[r], RECAD_ERR_ARG, "lightgcn_shard: null peer mapping %d", r);
      base[r] = reinterpret_cast<char*>(st->peer_base[r]);
      pads.pad[r] = reinterpret_cast<uint32_t*>(base[r] + st->off_signal);
      if (st->peer_users[r] != Ug) multicast = false;            // the item block must sit at the same offset everywhere
    }
    bstate = reinterpret_cast<uint32_t*>(base[rank] + st->off_signal) + 32;
    sr = side_res();
    RECAD_REQUIRE(sr, RECAD_ERR_CUDA, "lightgcn_shard: cannot create the side stream");
    return RECAD_OK;
  }

  // The owner's reduce + store runs on the side stream, next to the SpMMs of the main stream: it is bound by NVLink
  // ingress (every rank receives the whole item block, 51 MB at I = 200 k), not by SMs, so it gets a small grid.
  int side_begin() {
    RECAD_CUDA_CHECK(cudaEventRecord(sr->ev_go, s));
    RECAD_CUDA_CHECK(cudaStreamWaitEvent(sr->side, sr->ev_go, 0));
    return RECAD_OK;
  }
  int side_end() {
    RECAD_CUDA_CHECK(cudaEventRecord(sr->ev_done, sr->side));
    pending = true;
    return RECAD_OK;
  }
  int join_side() {                     // the main stream continues only after the side stream's reduce
    if (!pending) return RECAD_OK;
    RECAD_CUDA_CHECK(cudaStreamWaitEvent(s, sr->ev_done, 0));
    pending = false;
    mark(3);
    return RECAD_OK;
  }

  void mark(int phase) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    marks.emplace_back(phase, e);
  }

  int barrier() {
    if (world == 1) return RECAD_OK;
    peer_barrier_kernel<<<1, 32, 0, s>>>(pads, rank, world, bstate);
    RECAD_LAUNCH_CHECK();
    mark(2);
    return RECAD_OK;
  }

  // owner's half of an exchange over its slice of the item block (side stream)
  int reduce_rows(const ReduceArgs& a, bool mc_y, bool mc_z) {
    if (a.n <= 0) return RECAD_OK;
    static const int max_blocks = getenv("RECAD_SHARD_REDUCE_BLOCKS") ? std::max(1, atoi(getenv("RECAD_SHARD_REDUCE_BLOCKS"))) : 48;
    const unsigned grid = (unsigned)std::min<int64_t>((a.n + 511) / 512, max_blocks);
    if (mc_y && mc_z) shard_reduce_kernel<true, true><<<grid, 512, 0, sr->side>>>(a);
    else if (mc_y) shard_reduce_kernel<true, false><<<grid, 512, 0, sr->side>>>(a);
    else if (mc_z) shard_reduce_kernel<false, true><<<grid, 512, 0, sr->side>>>(a);
    else shard_reduce_kernel<false, false><<<grid, 512, 0, sr->side>>>(a);
    RECAD_LAUNCH_CHECK();
    return RECAD_OK;
  }

  // y = A x over the sharded matrix.  x: local pointer to [users ; items] rows (a symmetric buffer or the table E);
  // item rows: sum of the ranks' partial rows -> raw into buffer off_y (if >= 0) and alpha * (C + sum) into off_z (if >= 0)
  // user rows: the same two outputs, local.  C: local base pointer of [users ; items] rows or null.
  // ONE barrier per exchange: it sits between the push and the user-row SpMM and certifies two things at once -- every
  // rank's pushes for THIS exchange have landed, and every owner's stores of the PREVIOUS exchange have landed (the
  // user-row SpMM is the first kernel that reads them).  The staging slots are double buffered: a rank may push
  // exchange e + 1 while a slower owner still adds up the slots of exchange e.  The results of the last exchange of a
  // sequence become visible with finish_exchanges().
  int n_exchange = 0;
  // z_all: the epilogue result alpha * (C + sum) goes to EVERY replica (Horner steps, the last forward layer); otherwise only
  // to the owner's own copy (the running mean of the inner forward layers is needed by nobody else until the last layer)
  int spmm_exchange(const float* x, int64_t off_y, const float* C, int64_t off_z, float alpha, bool z_all) {
    const int64_t fD = D;
    const int64_t stage_off = st->off_stage + (int64_t)(n_exchange & 1) * (int64_t)world * slice * fD * (int64_t)sizeof(float);
    ++n_exchange;
    // 1. partial item rows, pushed to their owners from the epilogue (reads only local user rows)
    float* dst[kMaxShardPeers];
    for (int o = 0; o < world; ++o) dst[o] = buf(o, stage_off) + (int64_t)rank * slice * fD;
    int rc = recad_spmm_scatter(st->g_item, x, dst, world, slice, D, s);
    if (rc) return rc;
    mark(0);
    if ((rc = join_side())) return rc;              // the previous exchange's stores were issued before this barrier
    if ((rc = barrier())) return rc;
    if ((rc = side_begin())) return rc;
    // 2. own user rows (complete), epilogue fused; reads the item block the owners stored in the previous exchange
    rc = recad_spmm(st->g_user, x + Ug * fD, off_y >= 0 ? local(off_y) : nullptr, C, off_z >= 0 ? local(off_z) : nullptr, alpha, D, s);
    if (rc) return rc;
    mark(1);
    // 3. (side stream, concurrently) my slice: add the partial copies in rank order, store into every replica
    ReduceArgs a{};
    a.n_src = world;
    for (int r = 0; r < world; ++r) a.src[r] = local(stage_off) + (int64_t)r * slice * fD;
    a.n = rows_mine * fD / 4;
    const int64_t mine = (int64_t)rank * slice * fD;                    // floats into the item block
    a.C = C ? C + Ug * fD + mine : nullptr;
    a.alpha = alpha;
    auto fill = [&](int64_t off, float** out, int& n, bool all, bool& mc_flag) {
      mc_flag = false;
      if (off < 0) { n = 0; return; }
      if (!all) { out[0] = local(off) + item_off(rank) + mine; n = 1; return; }
      if (multicast) { out[0] = mc(off) + item_off(rank) + mine; n = 1; mc_flag = true; return; }
      for (int r = 0; r < world; ++r) out[r] = buf(r, off) + item_off(r) + mine;
      n = world;
    };
    bool mc_y, mc_z;
    fill(off_y, a.out_y, a.n_y, true, mc_y);
    fill(off_z, a.out_z, a.n_z, z_all, mc_z);
    if ((rc = reduce_rows(a, mc_y, mc_z))) return rc;
    return side_end();
  }
  int finish_exchanges() {
    int rc = join_side();
    if (rc) return rc;
    return barrier();
  }

  // O = mean_k A^k E
  int propagate() {
    const float* x = st->E;
    for (int k = 0; k < L; ++k) {
      const bool last = k == L - 1;
      const int64_t off_y = last ? -1 : ((k & 1) ? st->off_X1 : st->off_X0);
      int rc = spmm_exchange(x, off_y, k == 0 ? st->E : local(st->off_O), st->off_O, last ? 1.0f / (float)(L + 1) : 1.0f, last);
      if (rc) return rc;
      if (!last) x = local(off_y);
    }
    return finish_exchanges();       // O's item block is complete on every rank
  }

  // the item block of g and of cnt: every owner pulls its slice of every rank's partial block, sums, stores into every replica
  int gradient_exchange() {
    if (world == 1) return RECAD_OK;
    int rc = barrier();                           // every rank's BPR is done: the partial blocks are final
    if (rc) return rc;
    if ((rc = side_begin())) return rc;
    const int64_t fD = D, mine = (int64_t)rank * slice * fD;
    ReduceArgs a{};
    a.n_src = world;
    for (int r = 0; r < world; ++r) a.src[r] = buf(r, st->off_g) + item_off(r) + mine;
    a.n = rows_mine * fD / 4;
    if (multicast) { a.out_y[0] = mc(st->off_g) + item_off(rank) + mine; a.n_y = 1; }
    else { for (int r = 0; r < world; ++r) a.out_y[r] = buf(r, st->off_g) + item_off(r) + mine; a.n_y = world; }
    if ((rc = reduce_rows(a, multicast, false))) return rc;
    ReduceArgs c{};
    c.n_src = world;
    for (int r = 0; r < world; ++r) {
      c.src[r] = buf(r, st->off_cnt) + st->peer_users[r] + (int64_t)rank * slice;
      c.out_y[r] = buf(r, st->off_cnt) + st->peer_users[r] + (int64_t)rank * slice;
    }
    c.n_y = world;
    c.n = rows_mine;
    if (c.n > 0) {
      shard_reduce_scalar_kernel<<<(unsigned)std::min<int64_t>((c.n + 255) / 256, 32), 256, 0, sr->side>>>(c);
      RECAD_LAUNCH_CHECK();
    }
    mark(6);
    return side_end();               // joined, then certified, by the barrier of the first Horner exchange
  }
};

}  // namespace
}  // namespace recad

using namespace recad;

extern "C" {

int recad_lightgcn_shard_propagate(const recad_lightgcn_shard* st, void* stream) {
  Shard sh;
  int rc = sh.init(st, stream);
  if (rc) return rc;
  return sh.propagate();
}

int recad_lightgcn_shard_train_epoch(const recad_lightgcn_shard* st, const recad_epoch_samples* ep, int64_t batch, int64_t step0,
                                     double* trace_ms, void* stream) {
  Shard sh;
  int rc = sh.init(st, stream);
  if (rc) return rc;
  RECAD_REQUIRE(ep && ep->n_samples > 0 && batch > 0 && step0 >= 0, RECAD_ERR_ARG, "lightgcn_shard_train_epoch: bad samples");
  RECAD_REQUIRE(ep->rows || (ep->users && ep->rel && ep->negs && st->pos_rowptr && st->pos_col), RECAD_ERR_ARG,
                "lightgcn_shard_train_epoch: neither rows nor (users, rel, negs) + the shard's positives");
  sh.trace = trace_ms != nullptr;
  cudaStream_t s = sh.s;
  const int64_t Ug = sh.Ug, I = sh.I, N = Ug + I;
  const int D = sh.D, L = sh.L;
  float* g = sh.local(st->off_g);
  float* cnt = sh.local(st->off_cnt);
  float* O = sh.local(st->off_O);
  RECAD_CUDA_CHECK(cudaMemsetAsync(st->loss_acc, 0, 4 * sizeof(double), s));
  int64_t step = step0;
  sh.mark(-1);
  for (int64_t b0 = 0; b0 < ep->n_samples; b0 += batch) {
    const int64_t B = std::min(batch, ep->n_samples - b0);
    ++step;
    if ((rc = sh.propagate())) return rc;
    RECAD_CUDA_CHECK(cudaMemsetAsync(g, 0, N * D * sizeof(float), s));
    RECAD_CUDA_CHECK(cudaMemsetAsync(cnt, 0, N * sizeof(float), s));
    sh.mark(4);
    if (ep->rows)
      rc = launch_bpr_shard_rows(O, st->E, Ug, I, ep->rows, ep->perm64, b0, B, B, st->user_lo, sh.world, 1.0f / (float)(L + 1), g, cnt,
                                 st->loss_acc, D, s);
    else
      rc = launch_bpr_shard(O, st->E, Ug, I, ep->users, ep->rel, ep->negs, ep->perm32, b0, B, B, st->user_lo, st->pos_rowptr, st->pos_col,
                            (int32_t)st->pos_col_offset, sh.world, 1.0f / (float)(L + 1), g, cnt, st->loss_acc, D, s);
    if (rc) return rc;
    sh.mark(5);
    if ((rc = sh.gradient_exchange())) return rc;
    // Horner: t <- g + A t, L times
    const float* t = g;
    for (int k = 0; k < L; ++k) {
      const int64_t off_z = (k & 1) ? st->off_X1 : st->off_X0;
      if ((rc = sh.spmm_exchange(t, -1, g, off_z, 1.0f, true))) return rc;
      t = sh.local(off_z);
    }
    if ((rc = sh.finish_exchanges())) return rc;
    LossFold fold{st->loss_acc, 1.0 / (double)B, 0.5 * (double)st->lambda};
    rc = launch_adam(st->E, t, cnt, st->lambda / (float)B, st->m, st->v, N * D, D,
                     adam_scalars(st->lr, st->beta1, st->beta2, st->eps, step), fold, s);
    if (rc) return rc;
    sh.mark(7);
  }
  if (sh.trace) {
    RECAD_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int k = 0; k < 8; ++k) trace_ms[k] = 0.0;
    for (size_t k = 1; k < sh.marks.size(); ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, sh.marks[k - 1].second, sh.marks[k].second);
      if (sh.marks[k].first >= 0 && sh.marks[k].first < 8) trace_ms[sh.marks[k].first] += ms;
    }
    for (auto& m : sh.marks) cudaEventDestroy(m.second);
  }
  return RECAD_OK;
}

/* state[0] = barriers passed, state[1] = 1 if a barrier wait timed out (a peer never arrived) */
int recad_lightgcn_shard_barrier_state(const recad_lightgcn_shard* st, uint32_t* state2, void* stream) {
  RECAD_REQUIRE(st && st->peer_base && state2, RECAD_ERR_ARG, "lightgcn_shard_barrier_state: null argument");
  const uint32_t* p = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(st->peer_base[st->rank]) + st->off_signal) + 32;
  RECAD_CUDA_CHECK(cudaMemcpyAsync(state2, p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, as_stream(stream)));
  RECAD_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
  return RECAD_OK;
}

}  // extern "C"
