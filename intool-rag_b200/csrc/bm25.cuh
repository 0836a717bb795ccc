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
// A search is: bm25_plan_terms_kernel (per query: de-duplicate the terms in first-occurrence order, multiplicity
// folded into the weight, empty / out-of-vocabulary ones dropped), bm25_plan_cursors_kernel (per (query, term,
// slice boundary): lower_bound of the boundary's first doc id in the term's posting list, two levels), then the
// scoring kernel of bm25_sweep.cuh over slices of 3072 (2688) docs.  This file holds the build / validation
// kernels, the plan and the helpers the scoring kernel shares.
#pragma once
#include "common.cuh"
#include "dense_exact.cuh"

namespace hr {

constexpr int kBmMaxTerms = 64;   // distinct scorable terms per query (lane t owns terms t and t + 32)
constexpr int kBmMaxK = 128;      // candidate depth the kernel supports
constexpr int kBsCoarse = 16;     // slice boundaries per coarse boundary in the cursor plan
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
// boundaries j = 0..nb-1 (the last one is the end of the list).  Thread = (term group g, boundary j): consecutive
// threads search consecutive boundaries of the same list, so the searches share cache lines; a thread handles the
// terms g, g + G, ... of the query.  G = 1 for big tables (enough threads anyway); small tables (a shard, the coarse
// level) use up to 8 groups, or the chain of nt dependent binary searches per thread is all the kernel's time.
// With `coarse` != nullptr the search runs inside [coarse[j / ratio], coarse[j / ratio + 1]] (a table of the same
// layout with nbc boundaries every ratio * bdocs docs).
__global__ void __launch_bounds__(256)
bm25_plan_cursors_kernel(const int32_t* __restrict__ post_doc, const int32_t* __restrict__ q_indptr,
                         const int* __restrict__ plan_nt, const int64_t* __restrict__ plan_start,
                         const uint32_t* __restrict__ plan_len, int64_t nb, int64_t bdocs,
                         const uint32_t* __restrict__ coarse, int64_t nbc, int ratio, uint32_t* __restrict__ cur,
                         int G) {
  const int q = blockIdx.x;
  const int nt = plan_nt[q];
  const int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  const int g = (int)(idx / nb);
  const int64_t j = idx - (int64_t)g * nb;
  if (g >= G || g >= nt) return;
  const int qa = q_indptr[q];
  uint32_t* out = cur + (size_t)qa * (size_t)nb + (size_t)j * nt;
  const int64_t bound64 = j * bdocs;
  const uint32_t* cq = coarse ? coarse + (size_t)qa * (size_t)nbc : nullptr;
  const int64_t jc = j / ratio;
  for (int u = g; u < nt; u += G) {
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
// term groups for a table of nq * nb boundaries: enough threads to fill the GPU a few times over
inline int bm25_plan_groups(int64_t nq, int64_t nb) {
  const int64_t want = 2400000;
  const int64_t g = (want + nq * nb - 1) / std::max<int64_t>(1, nq * nb);
  return (int)std::min<int64_t>(8, std::max<int64_t>(1, g));
}

// sort buffer of the merge kernels (keys per query)
constexpr int kBmMergeCap = 4096;

}  // namespace hr
