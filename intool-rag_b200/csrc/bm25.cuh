// bm25.cuh — BM25 scoring over a CSR-by-term inverted index (HBM-bound integer/scatter work).
//
// The reference advertises BM25 (README.md:54-58, rag/config.py:43-45) but implements none; the
// definition is oracle/bm25.py.  Layout in HBM (struct of arrays):
//   indptr int64[V+1] (list lengths / df) | pstart int64[V+1] | post_doc int32[nnz_pad] | post_imp fp32[nnz_pad]
// List t occupies [pstart[t], pstart[t] + len_t), doc ids ascending; every list starts at a multiple of 4 and is
// padded to the next multiple of 4 with sentinels (doc = INT_MAX, impact = 0), so 16-byte groups of postings
// never straddle two lists (bm25_sweep.cuh).  post_imp is the length-normalised saturation
// tf(k1+1)/(tf+k1(1-b+b dl/avgdl)) folded at build time, so one posting costs 8 streamed bytes and
// score(q,d) = sum_t mult(t) idf(t) imp(t,d).
//
// A search is three kernels:
//   bm25_plan_terms_kernel   per query: de-duplicate the terms (first-occurrence order, multiplicity
//                            folded into the weight), drop empty / out-of-vocabulary ones.
//   bm25_plan_cursors_kernel per (query, term, slice boundary): lower_bound of the boundary's first
//                            doc id in the term's posting list (two levels: every 16th boundary by a
//                            search over the whole list, the others inside the bracketing pair).
//                            The document axis is cut into slices of 3072 (2688) docs; with the cursor
//                            table every (query, slice) is an independent, exactly-known set of
//                            posting ranges.
//   bm25_slice_kernel        A WARP owns a run of consecutive slices of one query and works alone:
//                            no block barrier, no atomics on the accumulators.  The slice's
//                            accumulators (one fp32 per doc) live in the warp's shared memory.  The
//                            posting ranges of the slice are cut into slots (<= 32 postings: one per
//                            lane; otherwise 128 postings, four per lane with 16-byte loads); the
//                            warp loads kBsBatch slots at once (all loads in flight before the first
//                            use), then applies them in term order with plain shared-memory
//                            read-modify-writes: doc ids inside one posting list are unique, terms
//                            follow each other in program order, so the fp32 summation order is
//                            fixed and results are deterministic.  A doc whose running score reaches
//                            the running threshold is pushed to a small hot list; at the end of a
//                            slice only the listed docs become candidates (warp-private key buffer,
//                            bitonic compaction; thresholds shared through the CTA and, per query,
//                            through global memory) and the accumulators are cleared with one sweep.
#pragma once
#include "common.cuh"
#include "dense_exact.cuh"

namespace hr {

constexpr int kBmMaxTerms = 64;   // distinct scorable terms per query (lane t owns terms t and t + 32)
constexpr int kBmMaxK = 128;      // candidate depth the kernel supports
constexpr int kBsThreads = 256;
constexpr int kBsWarps = kBsThreads / 32;
// Docs per warp slice.  Two CTAs (16 warps) per SM with the largest slice that fits measured best: fewer,
// fuller slots beat more resident warps (3 x 8 warps x 1920 docs: 9.9 ms; 2 x 8 x 3072: 9.3 ms at C1).
// The slice shrinks when the candidate buffers are large (k_c > 64) so that two CTAs still fit.
constexpr int kBsSliceLarge = 24 * 128;   // k_c <= 64: 96 KB of accumulators + 8 KB of key buffers per CTA
constexpr int kBsSliceSmall = 21 * 128;   // k_c <= 128: 84 KB + 16 KB
constexpr int kBsCtasPerSm = 2;
__host__ __device__ constexpr int bs_slice_docs(int kcp) { return kcp <= 64 ? kBsSliceLarge : kBsSliceSmall; }
constexpr int kBsCoarse = 16;        // slice boundaries per coarse boundary in the cursor plan
constexpr int kBsSlotCap = 32;       // slot descriptors per round (512 B per warp)
constexpr int kBsSlotLen = 128;      // postings of a wide slot
constexpr int kBsBatch = 5;          // slots a warp keeps in flight
constexpr int kBsHotCap = 64;        // docs that may reach the threshold in one slice before the full sweep takes over
// dynamic shared memory: accumulators | warp key buffers (2*kcp keys each)
__host__ __device__ constexpr int bs_smem_bytes(int kcp) {
  return kBsWarps * bs_slice_docs(kcp) * 4 + kBsWarps * 2 * kcp * 8;
}

constexpr int32_t kBmSentinelDoc = 0x7FFFFFFF;

// CSR (indptr, post_doc, post_tf) -> padded posting arrays with folded impacts.  Thread = posting i: its term
// by a binary search in indptr, destination pstart[t] + (i - indptr[t]).  Validates what the scoring kernels
// rely on: 0 <= doc < n_docs, doc ids strictly ascending inside a list, tf > 0 (bad[0] counts violations).
__global__ void bm25_pad_impact_kernel(const int64_t* __restrict__ indptr, const int64_t* __restrict__ pstart,
                                       const int32_t* __restrict__ post_doc, const int32_t* __restrict__ post_tf,
                                       const int32_t* __restrict__ doc_len, int64_t nnz, int64_t vocab,
                                       int64_t n_docs, double k1, double b, double avgdl,
                                       int32_t* __restrict__ out_doc, float* __restrict__ out_imp,
                                       unsigned long long* __restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = vocab;   // last t with indptr[t] <= i
    while (lo < hi) {
      const int64_t mid = (lo + hi + 1) >> 1;
      if (__ldg(indptr + mid) <= i) lo = mid; else hi = mid - 1;
    }
    const int64_t t0 = __ldg(indptr + lo);
    const int32_t doc = post_doc[i];
    const int32_t tfi = post_tf[i];
    bool ok = doc >= 0 && (int64_t)doc < n_docs && tfi > 0;
    if (ok && i > t0) ok = post_doc[i - 1] < doc;
    if (!ok) {
      atomicAdd(bad, 1ull);
      continue;
    }
    const double tf = (double)tfi;
    const double dl = (double)doc_len[doc];
    const double norm = avgdl > 0.0 ? k1 * (1.0 - b + b * dl / avgdl) : k1;
    const int64_t dst = __ldg(pstart + lo) + (i - t0);
    out_doc[dst] = doc;
    out_imp[dst] = (float)(tf * (k1 + 1.0) / (tf + norm));
  }
}

// sentinels behind every list (up to the next multiple of 4) and behind the last one
__global__ void bm25_pad_sentinels_kernel(const int64_t* __restrict__ indptr, const int64_t* __restrict__ pstart,
                                          int64_t vocab, int64_t tail, int32_t* __restrict__ out_doc,
                                          float* __restrict__ out_imp) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; t < vocab; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = pstart[t] + (indptr[t + 1] - indptr[t]);
    for (int64_t j = e; j < pstart[t + 1]; ++j) {
      out_doc[j] = kBmSentinelDoc;
      out_imp[j] = 0.f;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < tail) {
    out_doc[pstart[vocab] + threadIdx.x] = kBmSentinelDoc;
    out_imp[pstart[vocab] + threadIdx.x] = 0.f;
  }
}

// load-time validation of a padded index read from a file: lists ascending, docs in range, padding = sentinels
__global__ void bm25_check_padded_kernel(const int64_t* __restrict__ indptr, const int64_t* __restrict__ pstart,
                                         const int32_t* __restrict__ post_doc, int64_t vocab, int64_t n_docs,
                                         unsigned long long* __restrict__ bad) {
  // a warp per term, lanes over its padded list
  const int lane = threadIdx.x & 31;
  int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; t < vocab; t += nw) {
    const int64_t p0 = pstart[t], len = indptr[t + 1] - indptr[t], pe = pstart[t + 1];
    unsigned long long nb = 0;
    for (int64_t j = p0 + lane; j < pe; j += 32) {
      const int32_t d = post_doc[j];
      if (j < p0 + len) {
        if (d < 0 || (int64_t)d >= n_docs || (j > p0 && post_doc[j - 1] >= d)) nb++;
      } else if (d != kBmSentinelDoc) nb++;
    }
    if (nb) atomicAdd(bad, nb);
  }
}

// sort n (power of two, >= 64) keys in shared memory, descending, by one warp
__device__ __forceinline__ void warp_bitonic_desc(uint64_t* keys, int n, int lane) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int i = lane; i < (n >> 1); i += 32) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncwarp();
}

// ---- plan, step 1: unique terms of each query ------------------------------------------------------
// One warp per query.  Slot u of query q lives at index q_indptr[q] + u of the plan arrays (u < nt[q] <=
// raw term count), in order of first occurrence.
__global__ void __launch_bounds__(256)
bm25_plan_terms_kernel(const int64_t* __restrict__ indptr, const int64_t* __restrict__ pstart,
                       const float* __restrict__ idf, int64_t V,
                       const int32_t* __restrict__ q_indptr, const int32_t* __restrict__ q_terms, int nq,
                       int* __restrict__ plan_nt, int64_t* __restrict__ plan_start, uint32_t* __restrict__ plan_len,
                       float* __restrict__ plan_wgt, unsigned long long* __restrict__ postings_touched,
                       int* __restrict__ too_many) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const int qa = q_indptr[q];
  const int nraw = q_indptr[q + 1] - qa;   // any length; at most kBmMaxTerms DISTINCT scorable terms
  int base = 0;
  unsigned long long touched = 0;
  for (int r0 = 0; r0 < nraw; r0 += 32) {
    const int i = r0 + lane;
    bool keep = false;
    int mult = 0;
    int64_t a = 0, e = 0;
    int t = -1;
    if (i < nraw) {
      t = q_terms[qa + i];
      keep = (t >= 0 && t < V);
      if (keep) {
        for (int j = 0; j < nraw; ++j) {
          const int u = q_terms[qa + j];
          if (u == t) {
            if (j < i) keep = false;
            mult++;
          }
        }
      }
      if (keep) {
        a = indptr[t];
        e = indptr[t + 1];
        keep = e > a;
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (base + __popc(m) > kBmMaxTerms) {
      // more distinct terms than a warp owns: report (the host turns it into HR_ERR_INVALID), keep the first 64
      if (lane == 0 && too_many) atomicAdd(too_many, 1);
      keep = keep && (base + __popc(m & ((1u << lane) - 1u)) < kBmMaxTerms);
      m = __ballot_sync(0xffffffffu, keep);
    }
    if (keep) {
      const int slot = qa + base + __popc(m & ((1u << lane) - 1u));
      plan_start[slot] = pstart[t];
      plan_len[slot] = (uint32_t)(e - a);
      plan_wgt[slot] = (float)mult * idf[t];
      touched += (unsigned long long)(e - a);
    }
    base += __popc(m);
  }
  if (lane == 0) plan_nt[q] = base;
  if (postings_touched) {
    for (int o = 16; o > 0; o >>= 1) touched += __shfl_xor_sync(0xffffffffu, touched, o);
    if (lane == 0 && touched) atomicAdd(postings_touched, touched);
  }
}

// ---- plan, step 2: cursor table --------------------------------------------------------------------
// cur[(size_t)q_indptr[q] * nb + j * nt + u] = number of postings of term u with doc < j * bdocs, for the nb
// boundaries j = 0..nb-1 (the last one is the end of the list).  Thread = boundary j (consecutive threads
// search consecutive boundaries of the same list, so the searches share cache lines), looping over the
// query's terms.  With `coarse` != nullptr the search runs inside [coarse[j / ratio], coarse[j / ratio + 1]]
// (a table of the same layout with nbc boundaries every ratio * bdocs docs).
__global__ void __launch_bounds__(256)
bm25_plan_cursors_kernel(const int32_t* __restrict__ post_doc, const int32_t* __restrict__ q_indptr,
                         const int* __restrict__ plan_nt, const int64_t* __restrict__ plan_start,
                         const uint32_t* __restrict__ plan_len, int64_t nb, int64_t bdocs,
                         const uint32_t* __restrict__ coarse, int64_t nbc, int ratio, uint32_t* __restrict__ cur) {
  const int q = blockIdx.x;
  const int nt = plan_nt[q];
  const int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= nb || nt == 0) return;
  const int qa = q_indptr[q];
  uint32_t* out = cur + (size_t)qa * (size_t)nb + (size_t)j * nt;
  const int64_t bound64 = j * bdocs;
  const uint32_t* cq = coarse ? coarse + (size_t)qa * (size_t)nbc : nullptr;
  const int64_t jc = j / ratio;
  for (int u = 0; u < nt; ++u) {
    const uint32_t len = plan_len[qa + u];
    uint32_t lo = 0, hi = len;
    if (cq) {
      lo = cq[(size_t)jc * nt + u];
      hi = (jc + 1 < nbc) ? cq[(size_t)(jc + 1) * nt + u] : len;
    }
    if (j == 0) hi = lo = 0;
    else if (j == nb - 1 || bound64 > 0x7FFFFFFFll) lo = hi = len;
    else {
      const int32_t* p = post_doc + plan_start[qa + u];
      const int32_t bound = (int32_t)bound64;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(p + mid) < bound) lo = mid + 1; else hi = mid;
      }
    }
    out[u] = lo;
  }
}

// ---- scoring ---------------------------------------------------------------------------------------
// One slot = consecutive postings of one term inside the current slice.
//   narrow: e <= 32 postings starting at p, lane l takes posting l (4-byte loads)
//   wide  : kBsSlotLen postings starting at the 4-aligned index p, lane l takes 4l..4l+3 (16-byte loads);
//           [f, e) of them lie inside the slice's range of the term
struct __align__(16) BsSlot {
  uint32_t p_lo, p_hi;   // global posting index of the slot's first posting
  uint32_t meta;         // f | e << 8 | narrow << 16 | first slot of its term << 17
  float w;               // term weight (multiplicity * idf)
};

// Per-term state of a query in the registers of the warp: lane l owns terms l and l + 32.
struct BsTerms {
  int64_t start[2];        // first posting of the list
  float wgt[2];
  uint32_t c0[2], c1[2];   // cursors at the two boundaries of the current slice
};

// Warp-collective: describe slots [r0, r0 + kBsSlotCap) of the slice with cursors [c0, c1) per term.
// Returns the number of slots of the slice.  Slot numbers come from warp prefix sums; each lane then
// takes one slot of the round and finds the owning term by a binary search over the lanes' prefix sums.
template <bool TWO_HALVES>
__device__ __forceinline__ int bs_build_slots(int lane, const BsTerms& T, int r0, BsSlot* slots) {
  int64_t a[2] = {0, 0};
  uint32_t n[2] = {0, 0};
  int cnt[2] = {0, 0};
#pragma unroll
  for (int half = 0; half < (TWO_HALVES ? 2 : 1); ++half) {
    n[half] = T.c1[half] - T.c0[half];
    a[half] = T.start[half] + T.c0[half];
    if (n[half] > 32u) cnt[half] = (int)((a[half] + n[half] - (a[half] & ~(int64_t)3) + kBsSlotLen - 1) / kBsSlotLen);
    else cnt[half] = n[half] ? 1 : 0;
  }
  int iA = cnt[0], iB = cnt[1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int va = __shfl_up_sync(0xffffffffu, iA, o);
    if (lane >= o) iA += va;
    if (TWO_HALVES) {
      const int vb = __shfl_up_sync(0xffffffffu, iB, o);
      if (lane >= o) iB += vb;
    }
  }
  const int totA = __shfl_sync(0xffffffffu, iA, 31);
  const int totB = TWO_HALVES ? __shfl_sync(0xffffffffu, iB, 31) : 0;
  const int nslots = totA + totB;
  const int incl[2] = {iA, totA + iB};
  const int gg = r0 + lane;
  const int half = (TWO_HALVES && gg >= totA) ? 1 : 0;
  int lo_l = 0, hi_l = 31;   // smallest lane whose inclusive prefix exceeds gg
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    const int mid = (lo_l + hi_l) >> 1;
    const int v0 = __shfl_sync(0xffffffffu, incl[0], mid);
    int v = v0;
    if (TWO_HALVES) {
      const int v1 = __shfl_sync(0xffffffffu, incl[1], mid);
      v = half ? v1 : v0;
    }
    if (v > gg) hi_l = mid; else lo_l = mid + 1;
  }
  const int src = lo_l;
  int64_t at = __shfl_sync(0xffffffffu, a[0], src);
  uint32_t nn = __shfl_sync(0xffffffffu, n[0], src);
  int inc = __shfl_sync(0xffffffffu, incl[0], src);
  int cn = __shfl_sync(0xffffffffu, cnt[0], src);
  float wt = __shfl_sync(0xffffffffu, T.wgt[0], src);
  if (TWO_HALVES) {
    const int64_t a1 = __shfl_sync(0xffffffffu, a[1], src);
    const uint32_t n1 = __shfl_sync(0xffffffffu, n[1], src);
    const int i1 = __shfl_sync(0xffffffffu, incl[1], src);
    const int c1 = __shfl_sync(0xffffffffu, cnt[1], src);
    const float w1 = __shfl_sync(0xffffffffu, T.wgt[1], src);
    if (half) {
      at = a1;
      nn = n1;
      inc = i1;
      cn = c1;
      wt = w1;
    }
  }
  if (gg < nslots) {
    const int sl = gg - (inc - cn);   // slot index inside its term
    BsSlot d;
    int64_t ps;
    uint32_t f, e, narrow;
    if (nn <= 32u) {
      ps = at;
      f = 0;
      e = nn;
      narrow = 1;
    } else {
      const int64_t s0 = at & ~(int64_t)3;
      ps = s0 + (int64_t)sl * kBsSlotLen;
      f = sl == 0 ? (uint32_t)(at - s0) : 0u;
      e = (uint32_t)min((int64_t)kBsSlotLen, at + nn - ps);
      narrow = 0;
    }
    d.p_lo = (uint32_t)ps;
    d.p_hi = (uint32_t)(ps >> 32);
    d.meta = f | (e << 8) | (narrow << 16) | ((sl == 0 ? 1u : 0u) << 17);
    d.w = wt;
    slots[lane] = d;
  }
  return nslots;
}

// Warp-collective append of a candidate key to the warp's key buffer (compaction by bitonic sort keeps the
// best kc and raises the warp / CTA / query thresholds).
__device__ __forceinline__ void bs_append(bool take, unsigned long long key, uint64_t* cb, int& cbn, int cbcap, int kc,
                                          unsigned long long& tau, float& tau_f, int lane,
                                          unsigned long long* s_tau, unsigned long long* tau_gq) {
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (!m) return;
  if (cbn + 32 > cbcap) {   // make room: keep the best kc (warp-uniform branch)
    for (int i = cbn + lane; i < cbcap; i += 32) cb[i] = 0;
    warp_bitonic_desc(cb, cbcap, lane);
    cbn = min(cbn, kc);
    if (cbn == kc) {
      const unsigned long long kth = cb[kc - 1];
      if (kth > tau) {
        tau = kth;
        tau_f = key_score(tau);
        if (lane == 0) {
          atomicMax(s_tau, tau);
          atomicMax(tau_gq, tau);
        }
      }
    }
    take = take && key > tau;
  }
  const unsigned m2 = __ballot_sync(0xffffffffu, take);
  if (take) cb[cbn + __popc(m2 & ((1u << lane) - 1u))] = key;
  cbn += __popc(m2);
  __syncwarp();
}

// grid = (nq, S): query fastest, so the first wave holds span 0 of many queries and later spans start
// from the thresholds earlier spans published in tau_g.  CTA (q, g) covers slices [g*spc, (g+1)*spc), its
// warp w the sub-run [w*spw, (w+1)*spw) of that.  out_keys [nq][S][kc], out_n [nq][S].
// kcp = power of two >= max(kc, 32); a warp's key buffer holds 2*kcp keys.
template <bool TWO_HALVES, int kBsSlice>
__device__ __forceinline__ void bs_warp_run(const int32_t* __restrict__ post_doc, const float* __restrict__ post_imp,
                                            const uint32_t* __restrict__ curq, int nt, int64_t nsl, int64_t s_begin,
                                            int64_t s_end, BsTerms& T, float* acc, BsSlot* slots, uint16_t* hotl,
                                            uint64_t* cb, int& cbn, int cbcap, int kc, unsigned long long& tau,
                                            float& tau_f, unsigned long long* s_tau, unsigned long long* tau_gq,
                                            int lane) {
  constexpr int NH = TWO_HALVES ? 2 : 1;
  uint32_t nxt[2] = {0, 0};
  for (int64_t sidx = s_begin; sidx < s_end; ++sidx) {
    const int32_t docbase = (int32_t)(sidx * kBsSlice);
    // cursors of the boundary after the next slice; postings of the next slice towards L2
#pragma unroll
    for (int half = 0; half < NH; ++half) {
      const int term = lane + 32 * half;
      nxt[half] = T.c1[half];
      if (term < nt && sidx + 2 <= nsl) nxt[half] = __ldg(curq + (size_t)(sidx + 2) * nt + term);
    }
    unsigned long long gt = 0;
    if ((sidx & 7) == 0 && lane == 0) gt = *((volatile unsigned long long*)tau_gq);
    const float tau_pos = tau_f > 0.f ? tau_f : 1.4e-45f;   // x >= tau_pos <=> x > 0 && x >= tau_f
    int nhot = 0;   // warp-uniform
    __syncwarp();
    const int nslots = bs_build_slots<TWO_HALVES>(lane, T, 0, slots);
    __syncwarp();
    for (int r0 = 0; r0 < nslots; r0 += kBsSlotCap) {
      if (r0 > 0) {
        __syncwarp();
        bs_build_slots<TWO_HALVES>(lane, T, r0, slots);
        __syncwarp();
      }
      const int nr = min(kBsSlotCap, nslots - r0);
      for (int b0 = 0; b0 < nr; b0 += kBsBatch) {
        int4 dd[kBsBatch];
        float4 vv[kBsBatch];
        uint32_t meta[kBsBatch];
        float ww[kBsBatch];
#pragma unroll
        for (int j = 0; j < kBsBatch; ++j) {
          meta[j] = 0;
          ww[j] = 0.f;
          dd[j] = make_int4(0, 0, 0, 0);
          vv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (b0 + j < nr) {
            const BsSlot e = slots[b0 + j];
            meta[j] = e.meta;
            ww[j] = e.w;
            const int64_t p = (int64_t)(((uint64_t)e.p_hi << 32) | e.p_lo);
            const uint32_t en = (e.meta >> 8) & 0xFFu;
            if (e.meta & 0x10000u) {
              if ((uint32_t)lane < en) {
                dd[j].x = __ldg(post_doc + p + lane);
                vv[j].x = __ldg(post_imp + p + lane);
              }
            } else if ((uint32_t)(4 * lane) < en) {
              dd[j] = __ldg(reinterpret_cast<const int4*>(post_doc + p + 4 * lane));
              vv[j] = __ldg(reinterpret_cast<const float4*>(post_imp + p + 4 * lane));
            }
          }
        }
#pragma unroll
        for (int j = 0; j < kBsBatch; ++j) {
          if (b0 + j < nr) {   // warp-uniform
            const uint32_t f = meta[j] & 0xFFu, e = (meta[j] >> 8) & 0xFFu;
            if (meta[j] & 0x20000u) __syncwarp();   // a new term may touch docs of the previous one
            float xs[4] = {0.f, 0.f, 0.f, 0.f};
            int offs[4] = {0, 0, 0, 0};
            if (meta[j] & 0x10000u) {
              // narrow: one posting per lane
              const bool valid = (uint32_t)lane < e;
              offs[0] = valid ? dd[j].x - docbase : 0;
              if (valid) {
                xs[0] = fmaf(ww[j], vv[j].x, acc[offs[0]]);
                acc[offs[0]] = xs[0];
              }
            } else {
              const int dv[4] = {dd[j].x, dd[j].y, dd[j].z, dd[j].w};
              const float iv[4] = {vv[j].x, vv[j].y, vv[j].z, vv[j].w};
              bool valid[4];
              // the four docs of a lane are distinct (one posting list): read all, then add, then write
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const uint32_t idx = 4u * lane + e4;
                valid[e4] = idx >= f && idx < e;
                offs[e4] = valid[e4] ? dv[e4] - docbase : 0;
                xs[e4] = acc[offs[e4]];
              }
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                xs[e4] = valid[e4] ? fmaf(ww[j], iv[e4], xs[e4]) : 0.f;
                if (valid[e4]) acc[offs[e4]] = xs[e4];
              }
            }
            // docs whose running score reached the threshold go to the hot list (the lane applying a doc's
            // last posting sees its final score, so every candidate is listed at least once)
            const bool hot = fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3])) >= tau_pos;
            if (__any_sync(0xffffffffu, hot)) {
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const bool h = xs[e4] >= tau_pos;
                const unsigned hm = __ballot_sync(0xffffffffu, h);
                if (hm) {
                  if (h) {
                    const int pos = nhot + __popc(hm & ((1u << lane) - 1u));
                    if (pos < kBsHotCap) hotl[pos] = (uint16_t)offs[e4];
                  }
                  nhot += __popc(hm);
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
    // ---- end of slice ----
    {
      unsigned long long ct = *((volatile unsigned long long*)s_tau);
      const unsigned long long g0 = __shfl_sync(0xffffffffu, gt, 0);
      if (g0 > ct) ct = g0;
      if (ct > tau) {
        tau = ct;
        tau_f = key_score(tau);
      }
    }
    if (nhot > kBsHotCap) {
      // cold thresholds: sweep the slice, extract and clear
      if (tau == 0) {
        // No threshold at all yet (first slice of a cold warp): appending every scored doc would cost a
        // bitonic compaction per 64 docs.  Take each lane's best m = ceil(kc/32) scores first; the kc-th
        // largest of those 32m scores (distinct docs of this slice) is a valid lower bound of the kc-th best.
        float top[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = lane * 4; j < kBsSlice; j += 128) {
          const float4 v = *reinterpret_cast<const float4*>(acc + j);
          const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            float x = ve[e4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {   // insertion into the descending top-4
              const float hi = fmaxf(top[t], x);
              x = fminf(top[t], x);
              top[t] = hi;
            }
          }
        }
        const int m = (kc + 31) >> 5;   // 1..4
        float seed = 0.f;
        bool found = false;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (t < m) {
            int rank = 0;   // values above top[t] under (value desc, lane asc, slot asc)
            for (int l = 0; l < 32; ++l) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float o = __shfl_sync(0xffffffffu, top[u], l);
                if (u < m) rank += (o > top[t]) || (o == top[t] && (l < lane || (l == lane && u < t)));
              }
            }
            if (rank == kc - 1) {
              seed = top[t];
              found = true;
            }
          }
        }
        const unsigned fm = __ballot_sync(0xffffffffu, found);
        if (fm) seed = __shfl_sync(0xffffffffu, seed, __ffs(fm) - 1);
        if (fm && seed > 0.f) {
          tau = make_key(seed, 0xFFFFFFFFu) - 1;   // every doc scoring >= seed still passes `key > tau`
          tau_f = seed;
          if (lane == 0) {
            atomicMax(s_tau, tau);
            atomicMax(tau_gq, tau);
          }
        }
      }
#pragma unroll 2
      for (int j = lane * 4; j < kBsSlice; j += 128) {
        float4 v = *reinterpret_cast<float4*>(acc + j);
        const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        *reinterpret_cast<float4*>(acc + j) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (__any_sync(0xffffffffu, mx > 0.f && mx >= tau_f)) {
          const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            unsigned long long key = 0;
            bool take = false;
            if (ve[e4] > 0.f && ve[e4] >= tau_f) {
              key = make_key(ve[e4], (uint32_t)(docbase + j + e4));
              take = key > tau;
            }
            bs_append(take, key, cb, cbn, cbcap, kc, tau, tau_f, lane, s_tau, tau_gq);
          }
        }
      }
    } else {
      // the listed docs only: the first reader of a doc takes its score (exchange with 0), duplicates see 0
      for (int i0 = 0; i0 < nhot; i0 += 32) {
        const int i = i0 + lane;
        unsigned long long key = 0;
        bool take = false;
        if (i < nhot) {
          const int off = hotl[i];
          const float v = atomicExch(acc + off, 0.f);
          if (v > 0.f && v >= tau_f) {
            key = make_key(v, (uint32_t)(docbase + off));
            take = key > tau;
          }
        }
        bs_append(take, key, cb, cbn, cbcap, kc, tau, tau_f, lane, s_tau, tau_gq);
      }
      __syncwarp();
#pragma unroll
      for (int j = lane * 4; j < kBsSlice; j += 128) *reinterpret_cast<float4*>(acc + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // advance the cursors by one slice
#pragma unroll
    for (int half = 0; half < NH; ++half) {
      T.c0[half] = T.c1[half];
      T.c1[half] = nxt[half];
    }
  }
}

template <int kBsSlice>
__global__ void __launch_bounds__(kBsThreads, kBsCtasPerSm)
bm25_slice_kernel(const int32_t* __restrict__ post_doc, const float* __restrict__ post_imp,
                  const int32_t* __restrict__ q_indptr, const int* __restrict__ plan_nt,
                  const int64_t* __restrict__ plan_start, const float* __restrict__ plan_wgt,
                  const uint32_t* __restrict__ plan_cur, int64_t nsl, int spc, int S, int kc, int kcp,
                  uint64_t* __restrict__ out_keys, int* __restrict__ out_n, unsigned long long* __restrict__ tau_g) {
  extern __shared__ __align__(16) uint8_t bsm[];
  float* acc_all = (float*)bsm;
  uint64_t* cb_all = (uint64_t*)(bsm + kBsWarps * kBsSlice * 4);
  __shared__ BsSlot s_slots[kBsWarps][kBsSlotCap];
  __shared__ uint16_t s_hot[kBsWarps][kBsHotCap];
  __shared__ unsigned long long s_tau;
  __shared__ int s_wn[kBsWarps];

  const int q = blockIdx.x;
  const int g = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int w = tid >> 5;
  uint64_t* cb = cb_all + (size_t)w * 2 * kcp;
  const int cbcap = 2 * kcp;
  const int nt = plan_nt[q];
  const int64_t cta0 = (int64_t)g * spc;
  const int64_t cta1 = min(nsl, cta0 + spc);
  if (nt == 0 || cta0 >= cta1) {   // uniform
    if (tid == 0) out_n[(size_t)q * S + g] = 0;
    return;
  }
  const int64_t spw = (cta1 - cta0 + kBsWarps - 1) / kBsWarps;
  const int64_t s_begin = min(cta1, cta0 + (int64_t)w * spw);
  const int64_t s_end = min(cta1, s_begin + spw);
  const int qa = q_indptr[q];
  const uint32_t* curq = plan_cur + (size_t)qa * (size_t)(nsl + 1);
  unsigned long long* tau_gq = tau_g + q;
  float* acc = acc_all + w * kBsSlice;
  BsTerms T;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int term = lane + 32 * half;
    T.start[half] = 0;
    T.wgt[half] = 0.f;
    T.c0[half] = T.c1[half] = 0;
    if (term < nt && s_begin < s_end) {
      T.start[half] = plan_start[qa + term];
      T.wgt[half] = plan_wgt[qa + term];
      T.c0[half] = curq[(size_t)s_begin * nt + term];
      T.c1[half] = curq[(size_t)(s_begin + 1) * nt + term];
    }
  }
  if (tid == 0) s_tau = *((volatile unsigned long long*)tau_gq);
#pragma unroll
  for (int j = lane * 4; j < kBsSlice; j += 128) *reinterpret_cast<float4*>(acc + j) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  int cbn = 0;                 // keys in this warp's buffer (warp-uniform)
  unsigned long long tau = s_tau;   // this warp's threshold key: a lower bound of the query's kc-th best
  float tau_f = tau ? key_score(tau) : 0.f;
  if (nt <= 32)
    bs_warp_run<false, kBsSlice>(post_doc, post_imp, curq, nt, nsl, s_begin, s_end, T, acc, s_slots[w], s_hot[w], cb, cbn, cbcap,
                       kc, tau, tau_f, &s_tau, tau_gq, lane);
  else
    bs_warp_run<true, kBsSlice>(post_doc, post_imp, curq, nt, nsl, s_begin, s_end, T, acc, s_slots[w], s_hot[w], cb, cbn, cbcap,
                      kc, tau, tau_f, &s_tau, tau_gq, lane);
  // ---- warp list -> sorted top-kc ----
  for (int i = cbn + lane; i < cbcap; i += 32) cb[i] = 0;
  warp_bitonic_desc(cb, cbcap, lane);
  cbn = min(cbn, kc);
  if (lane == 0) {
    s_wn[w] = cbn;
    if (cbn == kc && cb[kc - 1] > tau) atomicMax(tau_gq, cb[kc - 1]);
  }
  __syncthreads();
  // ---- CTA merge of the warp lists (the accumulator area is the sort buffer) ----
  uint64_t* mbuf = (uint64_t*)acc_all;
  const int total = kBsWarps * kcp;
  for (int i = tid; i < total; i += kBsThreads) {
    const int ww = i / kcp, j = i - ww * kcp;
    mbuf[i] = (j < s_wn[ww]) ? cb_all[(size_t)ww * 2 * kcp + j] : 0ull;
  }
  block_bitonic_desc(mbuf, total);
  uint64_t* o = out_keys + ((size_t)q * S + g) * kc;
  int m = 0;
  for (int ww = 0; ww < kBsWarps; ++ww) m += s_wn[ww];
  m = min(m, kc);
  for (int j = tid; j < kc; j += kBsThreads) o[j] = (j < m) ? mbuf[j] : 0ull;
  if (tid == 0) out_n[(size_t)q * S + g] = m;
}

// merge S lists of kc keys per query -> S_out/I_out [nq][k]
constexpr int kBmMergeCap = 4096;
__global__ void __launch_bounds__(256)
bm25_merge_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ ns, int S, int kc, int k,
                  int64_t id_base, float* __restrict__ So, int64_t* __restrict__ Io) {
  __shared__ uint64_t buf[kBmMergeCap];
  const int q = blockIdx.x;
  const int total = S * kc;
  int pw = 1;
  while (pw < total) pw <<= 1;
  for (int i = threadIdx.x; i < pw; i += blockDim.x) {
    uint64_t v = 0;
    if (i < total) {
      int g = i / kc, j = i - g * kc;
      if (j < ns[(size_t)q * S + g]) v = keys[((size_t)q * S + g) * kc + j];
    }
    buf[i] = v;
  }
  block_bitonic_desc(buf, pw);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    uint64_t key = (j < total) ? buf[j] : 0ull;
    if (key) {
      So[(size_t)q * k + j] = key_score(key);
      Io[(size_t)q * k + j] = (int64_t)key_row(key) + id_base;
    } else {
      So[(size_t)q * k + j] = 0.f;
      Io[(size_t)q * k + j] = -1;
    }
  }
}

}  // namespace hr
