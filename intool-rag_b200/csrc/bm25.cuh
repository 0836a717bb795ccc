// bm25.cuh — BM25 scoring over a CSR-by-term inverted index (HBM-bound integer/scatter work).
//
// The reference advertises BM25 (README.md:54-58, rag/config.py:43-45) but implements none; the
// definition is oracle/bm25.py.  Layout in HBM (struct of arrays):
//   indptr int64[V+1] | post_doc int32[nnz] (ascending per term) | post_imp fp32[nnz]
// post_imp is the length-normalised saturation tf(k1+1)/(tf+k1(1-b+b dl/avgdl)) folded at build
// time, so one posting costs 8 streamed bytes and score(q,d) = sum_t mult(t) idf(t) imp(t,d).
//
// Search kernel: CTA (query q, range group g) sweeps its document ranges in ascending order.  For a
// range of R docs the accumulators live in shared memory; each query term keeps a cursor into its
// posting list, the CTA streams the postings below the range end with coalesced loads and adds
// them with plain (non-atomic) shared read-modify-writes: doc ids inside ONE posting list are
// unique, and terms are separated by a block barrier, so no two threads ever touch the same
// accumulator concurrently.  One read-and-clear pass then extracts candidates above the CTA's
// running threshold into a key buffer that is compacted by a bitonic sort when half full.
#pragma once
#include "common.cuh"
#include "dense_exact.cuh"

namespace hr {

constexpr int kBmThreads = 256;
constexpr int kBmUnroll = 4;
constexpr int kBmRange = 16384;  // docs per shared-memory accumulator window (64 KB)
constexpr int kBmCB = 1024;      // candidate key buffer
constexpr int kBmMaxTerms = 256; // unique terms per query
constexpr int kBmSmemBytes = kBmRange * 4 + kBmCB * 8 + kBmMaxTerms * (8 + 8 + 4) + 64;

__global__ void bm25_impact_kernel(const int32_t* __restrict__ post_doc, const int32_t* __restrict__ post_tf,
                                   const int32_t* __restrict__ doc_len, int64_t nnz, double k1, double b,
                                   double avgdl, float* __restrict__ imp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    double tf = (double)post_tf[i];
    double dl = (double)doc_len[post_doc[i]];
    double norm = avgdl > 0.0 ? k1 * (1.0 - b + b * dl / avgdl) : k1;
    imp[i] = (float)(tf * (k1 + 1.0) / (tf + norm));
  }
}

__device__ __forceinline__ int64_t lower_bound_doc(const int32_t* __restrict__ a, int64_t lo, int64_t hi,
                                                   int32_t v) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// out_keys [nq][S][kc], out_n [nq][S]
__global__ void __launch_bounds__(kBmThreads)
bm25_score_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ post_doc,
                  const float* __restrict__ post_imp, const float* __restrict__ idf, int64_t N, int64_t V,
                  const int32_t* __restrict__ q_indptr, const int32_t* __restrict__ q_terms, int S, int kc,
                  uint64_t* __restrict__ out_keys, int* __restrict__ out_n,
                  unsigned long long* __restrict__ postings_touched) {
  extern __shared__ __align__(16) uint8_t bsm[];
  float* acc = (float*)bsm;
  uint64_t* cb = (uint64_t*)(bsm + kBmRange * 4);
  int64_t* cur = (int64_t*)(bsm + kBmRange * 4 + kBmCB * 8);
  int64_t* endp = cur + kBmMaxTerms;
  float* wgt = (float*)(endp + kBmMaxTerms);
  __shared__ int s_nt, s_cbn, s_overflow;
  __shared__ int s_cnt3[3];
  __shared__ unsigned long long s_tau;

  const int q = blockIdx.y;
  const int g = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t NR = (N + kBmRange - 1) / kBmRange;
  const int64_t per = (NR + S - 1) / S;
  const int64_t r_first = (int64_t)g * per;
  const int64_t r_last = min(NR, r_first + per);

  // ---- query terms: dedup (multiplicity folds into the weight), cursors at the group's first doc
  const int qa = q_indptr[q], qb = q_indptr[q + 1];
  const int nraw = min(qb - qa, kBmMaxTerms);
  if (tid == 0) {
    s_nt = 0;
    s_cbn = 0;
    s_tau = 0;
    s_cnt3[0] = s_cnt3[1] = s_cnt3[2] = 0;
  }
  int par = 0;
  __syncthreads();
  if (tid < nraw) {
    const int t = q_terms[qa + tid];
    bool first = (t >= 0 && t < V);
    int mult = 0;
    for (int j = 0; j < nraw; ++j) {
      int u = q_terms[qa + j];
      if (u == t) {
        if (j < tid) first = false;
        mult++;
      }
    }
    if (first) {
      int64_t a = indptr[t], e = indptr[t + 1];
      if (e > a) {
        int slot = atomicAdd(&s_nt, 1);
        int64_t c0 = (r_first == 0) ? a : lower_bound_doc(post_doc, a, e, (int32_t)(r_first * kBmRange));
        cur[slot] = c0;
        endp[slot] = e;
        wgt[slot] = (float)mult * idf[t];
        if (g == 0 && postings_touched) atomicAdd(postings_touched, (unsigned long long)(e - a));
      }
    }
  }
  for (int i = tid; i < kBmRange; i += kBmThreads) acc[i] = 0.f;
  __syncthreads();
  const int nt = s_nt;

  for (int64_t r = r_first; r < r_last; ++r) {
    const int32_t r0 = (int32_t)(r * kBmRange);
    const int64_t rend64 = min((int64_t)N, (int64_t)r0 + kBmRange);
    const int32_t r_end = (int32_t)rend64;
    bool any = false;
    // ---- scatter-accumulate the postings of this range, one term at a time
    for (int t = 0; t < nt; ++t) {
      int64_t c = cur[t];
      const int64_t e = endp[t];
      const float w = wgt[t];
      if (c >= e) continue;  // uniform: cur/endp are shared
      for (;;) {
        int inr = 0;
#pragma unroll
        for (int u = 0; u < kBmUnroll; ++u) {
          int64_t idx = c + tid + u * kBmThreads;
          if (idx < e) {
            int32_t doc = __ldg(post_doc + idx);
            if (doc < r_end) {
              float im = __ldg(post_imp + idx);
              acc[doc - r0] = fmaf(w, im, acc[doc - r0]);
              inr++;
            }
          }
        }
        // postings are sorted by doc, so the in-range ones are a prefix of the chunk: the block
        // sum of `inr` is exactly how far the cursor moves.  Rotating 3-slot counter = 1 barrier.
        const int slot = par;
        par = (par == 2) ? 0 : par + 1;
        if (tid == 0) s_cnt3[par] = 0;  // slot of the NEXT iteration (last read two barriers ago)
        int wsum = __reduce_add_sync(0xffffffffu, inr);
        if ((tid & 31) == 0 && wsum) atomicAdd(&s_cnt3[slot], wsum);
        __syncthreads();
        const int tot = s_cnt3[slot];
        c += tot;
        if (tot) any = true;
        if (tot < kBmThreads * kBmUnroll) break;
      }
      if (tid == 0) cur[t] = c;
    }
    __syncthreads();
    if (!any) continue;  // uniform
    // ---- read-and-clear pass: candidates above the running threshold go to the key buffer
    for (;;) {
      if (tid == 0) s_overflow = 0;
      __syncthreads();
      const unsigned long long tau = s_tau;
      const float tau_f = tau ? key_score(tau) : 0.f;
      for (int j = tid; j < kBmRange; j += kBmThreads) {
        float v = acc[j];
        if (v > 0.f) {
          bool keep = false;
          if (v >= tau_f) {
            uint64_t key = make_key(v, (uint32_t)(r0 + j));
            if (key > tau) {
              int slot = atomicAdd(&s_cbn, 1);
              if (slot < kBmCB) cb[slot] = key; else { keep = true; s_overflow = 1; }
            }
          }
          if (!keep) acc[j] = 0.f;
        }
      }
      __syncthreads();
      const int ovf = s_overflow;
      int n = min(s_cbn, kBmCB);
      if (ovf || n > kBmCB / 2) {
        for (int i = n + tid; i < kBmCB; i += kBmThreads) cb[i] = 0;
        block_bitonic_desc(cb, kBmCB);
        if (tid == 0) {
          int m = n < kc ? n : kc;
          s_cbn = m;
          s_tau = (m == kc) ? cb[kc - 1] : 0ull;
        }
        __syncthreads();
      } else if (tid == 0) {
        s_cbn = n;
      }
      __syncthreads();
      if (!ovf) break;
    }
  }
  // ---- final compaction and write-out
  __syncthreads();
  {
    int n = min(s_cbn, kBmCB);
    for (int i = n + tid; i < kBmCB; i += kBmThreads) cb[i] = 0;
    block_bitonic_desc(cb, kBmCB);
    int m = n < kc ? n : kc;
    uint64_t* o = out_keys + ((size_t)q * S + g) * kc;
    for (int j = tid; j < kc; j += kBmThreads) o[j] = (j < m) ? cb[j] : 0ull;
    if (tid == 0) out_n[(size_t)q * S + g] = m;
  }
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
