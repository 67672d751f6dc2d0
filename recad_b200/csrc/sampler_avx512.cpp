// AVX-512 inner loop of the exact pairwise parse (recad_mt19937_pairwise_soa, reference: recad/dataset/implicit.py:50-74).
// Compiled with -mavx512f; entered only after a run-time CPU check (sampler.cpp).
//
// The parse is a chain: sample k + 1 starts where sample k stopped consuming the MT19937 stream.  The scalar form walks
// that chain through four loads, two mask-and-compare steps and three selects per sample, and mispredicts whenever a
// draw needs a third attempt (~14 % of the samples).  Here the stream is looked at in 64-word WINDOWS: one 64-bit mask
// says which words the negative draw would accept (its bound is the same for every sample), and per sample one
// 64-bit mask says which words its positive draw would accept -- 8 vector instructions that do not depend on the
// chain, so they run ahead of it.  What remains on the chain is shift / count-trailing-zeros / add, twice per sample,
// with no loads and no data-dependent branch.
#include <immintrin.h>
#include <stdint.h>

namespace {

inline uint64_t accept_mask64(const uint32_t* w, uint32_t mask, uint32_t r) {
  const __m512i vm = _mm512_set1_epi32((int)mask), vr = _mm512_set1_epi32((int)r);
  const __mmask16 k0 = _mm512_cmple_epu32_mask(_mm512_and_si512(_mm512_loadu_si512(w), vm), vr);
  const __mmask16 k1 = _mm512_cmple_epu32_mask(_mm512_and_si512(_mm512_loadu_si512(w + 16), vm), vr);
  const __mmask16 k2 = _mm512_cmple_epu32_mask(_mm512_and_si512(_mm512_loadu_si512(w + 32), vm), vr);
  const __mmask16 k3 = _mm512_cmple_epu32_mask(_mm512_and_si512(_mm512_loadu_si512(w + 48), vm), vr);
  return (uint64_t)k0 | ((uint64_t)k1 << 16) | ((uint64_t)k2 << 32) | ((uint64_t)k3 << 48);
}

}  // namespace

// Draw samples k0, k0 + 1, .. optimistically (position of the positive, first negative candidate the mask accepts) while
//   * k < k1,  * 64 stream words from the cursor are available (t + 64 <= avail),
//   * the sample is ordinary: 1 <= lens[k] < n_items  (users without positives and full rows are left to the caller).
// ring: the stream ring, `ring_mask` + 1 words with the first 64 words mirrored behind its end (a window never wraps).
// tst[k & tst_mask] receives the low 32 bits of the stream offset in front of sample k.  Returns the first sample NOT drawn;
// *t_io is the stream offset in front of it.
extern "C" int64_t recad_parse_window_avx512(const uint32_t* ring, uint64_t ring_mask, uint64_t* t_io, uint64_t avail,
                                             const uint32_t* lens, int64_t k0, int64_t k1, uint32_t* rel, uint32_t* negs,
                                             uint32_t* tst, int64_t tst_mask, uint32_t neg_r, uint32_t neg_mask, uint32_t n_items) {
  constexpr unsigned kLimit = 48;          // a sample starts at window offset <= kLimit: >= 16 words left for its two draws
  uint64_t T = *t_io;
  int64_t k = k0;
  while (k < k1 && T + 64 <= avail) {
    const uint32_t* w = ring + (T & ring_mask);
    const uint64_t Bm = accept_mask64(w, neg_mask, neg_r);
    unsigned o = 0;
    bool stop = false;
    const int64_t kend = k + 32 < k1 ? k + 32 : k1;
    for (; k < kend && o <= kLimit; ++k) {
      const uint32_t len = lens[k];
      if (__builtin_expect(len == 0 || len >= n_items, 0)) { stop = true; break; }
      const uint32_t r1 = len - 1, m1 = r1 ? 0xffffffffu >> __builtin_clz(r1) : 0u;
      unsigned o1 = o;
      uint32_t rv = 0;
      if (r1) {
        const uint64_t a = accept_mask64(w, m1, r1) >> o;
        if (__builtin_expect(a == 0, 0)) { stop = o == 0; break; }
        const unsigned ca = (unsigned)_tzcnt_u64(a);
        rv = w[o + ca] & m1;
        o1 = o + ca + 1;
        if (__builtin_expect(o1 > 63, 0)) { stop = o == 0; break; }
      }
      const uint64_t b = Bm >> o1;
      if (__builtin_expect(b == 0, 0)) { stop = o == 0; break; }
      const unsigned cb = (unsigned)_tzcnt_u64(b);
      tst[k & tst_mask] = (uint32_t)(T + o);
      rel[k] = rv;
      negs[k] = w[o1 + cb] & neg_mask;
      o = o1 + cb + 1;
    }
    T += o;
    if (stop) break;     // an extraordinary sample, or one whose draws do not fit a whole window: the caller's exact loop takes it
  }
  *t_io = T;
  return k;
}
