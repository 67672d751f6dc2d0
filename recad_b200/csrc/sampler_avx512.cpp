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

// The draws of np.random.shuffle (implicit.py:24-25) inside one power-of-two band: j[i] = first stream word with
// (word & mask) <= i, for i descending.  16 words per step: a word is accepted for certain when it is <= i - 15 (i drops
// by at most 15 inside the step); if any word falls in the 16-wide doubtful zone below i the step is left to the caller's
// scalar loop (probability 16 / (mask + 1) per word).  Accepted words are compressed, reversed and stored at
// j[i - cnt + 1 .. i].  Stops while at least 16 acceptances remain in the band.  Returns the words consumed.
extern "C" uint64_t recad_shuffle_draws_avx512(const uint32_t* src, uint64_t n_words, uint32_t mask, int64_t* i_io,
                                               int64_t band_lo, uint32_t* j_out) {
  int64_t i = *i_io;
  uint64_t q = 0;
  const __m512i vmask = _mm512_set1_epi32((int)mask);
  const __m512i lane = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
  while (q + 16 <= n_words && i - band_lo >= 16 && i >= 32) {
    const __m512i v = _mm512_and_si512(_mm512_loadu_si512(src + q), vmask);
    const __mmask16 sure = _mm512_cmple_epu32_mask(v, _mm512_set1_epi32((int)(uint32_t)(i - 15)));
    const __mmask16 all = _mm512_cmple_epu32_mask(v, _mm512_set1_epi32((int)(uint32_t)i));
    if (__builtin_expect(all != sure, 0)) break;             // a doubtful word: the caller decides it one by one
    const int cnt = __builtin_popcount((unsigned)sure);
    const __m512i packed = _mm512_maskz_compress_epi32(sure, v);
    const __m512i rev = _mm512_permutexvar_epi32(_mm512_sub_epi32(_mm512_set1_epi32(cnt - 1), lane), packed);
    _mm512_mask_storeu_epi32(j_out + (i - cnt + 1), (__mmask16)((1u << cnt) - 1u), rev);
    i -= cnt;
    q += 16;
  }
  *i_io = i;
  return q;
}
