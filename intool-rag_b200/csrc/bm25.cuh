// bm25.cuh — BM25 scoring over a CSR-by-term inverted index (HBM-bound integer/scatter work).
//
// The reference advertises BM25 (README.md:54-58, rag/config.py:43-45) but implements none; the
// definition is oracle/bm25.py.  Layout in HBM (struct of arrays):
//   indptr int64[V+1] | post_doc int32[nnz] (ascending per term) | post_imp fp32[nnz]
// post_imp is the length-normalised saturation tf(k1+1)/(tf+k1(1-b+b dl/avgdl)) folded at build
// time, so one posting costs 8 streamed bytes and score(q,d) = sum_t mult(t) idf(t) imp(t,d).
//
// Search kernel: CTA (query q, document span g).  Every WARP owns a contiguous sub-span of the
// documents and sweeps it in windows of 1024 docs whose accumulators live in shared memory
// (warp-private, so the scatter needs no atomics and no block barrier: doc ids inside one posting
// list are unique, and terms are applied one after the other inside the warp).  Per window the
// warp first issues the posting loads of up to 8 query terms (16 independent 128-byte requests in
// flight per warp), then adds them, then makes one vectorised read-and-clear pass that extracts
// the (rare) scores above its running threshold into a warp-private key buffer, compacted by a
// warp bitonic sort.  Warps publish their k-th best key to a CTA-wide threshold.
#pragma once
#include "common.cuh"
#include "dense_exact.cuh"

namespace hr {

constexpr int kBmWarps = 8;
constexpr int kBmThreads = kBmWarps * 32;
constexpr int kBmWin = 1024;      // docs per warp window (4 KB of accumulators)
constexpr int kBmMaxTerms = 64;   // unique terms per query
constexpr int kBmTermBatch = 8;   // posting streams a warp keeps in flight
constexpr int kBmMaxK = 128;      // candidate depth the kernel supports
// dynamic shared memory: accumulators | key buffers (2*kcp per warp) | cursors | ends | weights
__host__ __device__ constexpr int bm_smem_bytes(int kcp) {
  return kBmWarps * kBmWin * 4 + kBmWarps * 2 * kcp * 8 + kBmWarps * kBmMaxTerms * 8 + kBmMaxTerms * 8 +
         kBmMaxTerms * 4 + 64;
}

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

// out_keys [nq][S][kc], out_n [nq][S].  kcp = power of two >= max(kc, 32); key buffer holds 2*kcp keys.
__global__ void __launch_bounds__(kBmThreads)
bm25_score_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ post_doc,
                  const float* __restrict__ post_imp, const float* __restrict__ idf, int64_t N, int64_t V,
                  const int32_t* __restrict__ q_indptr, const int32_t* __restrict__ q_terms, int S, int kc, int kcp,
                  uint64_t* __restrict__ out_keys, int* __restrict__ out_n,
                  unsigned long long* __restrict__ postings_touched) {
  extern __shared__ __align__(16) uint8_t bsm[];
  float* acc_all = (float*)bsm;
  uint64_t* cb_all = (uint64_t*)(bsm + kBmWarps * kBmWin * 4);
  int64_t* cur_all = (int64_t*)(cb_all + kBmWarps * 2 * kcp);
  int64_t* endp = cur_all + kBmWarps * kBmMaxTerms;
  float* wgt = (float*)(endp + kBmMaxTerms);
  __shared__ int s_nt;
  __shared__ int s_term[kBmMaxTerms];
  __shared__ int64_t s_begin[kBmMaxTerms];
  __shared__ unsigned long long s_tau;
  __shared__ int s_wn[kBmWarps];

  const int q = blockIdx.y;
  const int g = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int w = tid >> 5;
  float* acc = acc_all + w * kBmWin;
  uint64_t* cb = cb_all + (size_t)w * 2 * kcp;
  int64_t* cur = cur_all + w * kBmMaxTerms;
  const int cbcap = 2 * kcp;

  // ---- document span of this CTA, sub-span of this warp (multiples of the window size) ----
  const int64_t nwin = (N + kBmWin - 1) / kBmWin;
  const int64_t per_cta = (nwin + S - 1) / S;
  const int64_t cta_w0 = (int64_t)g * per_cta;
  const int64_t cta_w1 = min(nwin, cta_w0 + per_cta);
  const int64_t span = max((int64_t)0, cta_w1 - cta_w0);
  const int64_t per_warp = (span + kBmWarps - 1) / kBmWarps;
  const int64_t w0 = cta_w0 + (int64_t)w * per_warp;
  const int64_t w1 = min(cta_w1, w0 + per_warp);

  // ---- query terms: dedup (multiplicity folds into the weight) ----
  const int qa = q_indptr[q], qb = q_indptr[q + 1];
  const int nraw = min(qb - qa, kBmMaxTerms);
  if (tid == 0) {
    s_nt = 0;
    s_tau = 0;
  }
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
        s_term[slot] = t;
        s_begin[slot] = a;
        endp[slot] = e;
        wgt[slot] = (float)mult * idf[t];
        if (g == 0 && postings_touched) atomicAdd(postings_touched, (unsigned long long)(e - a));
      }
    }
  }
  for (int i = tid; i < kBmWarps * kBmWin; i += kBmThreads) acc_all[i] = 0.f;
  __syncthreads();
  const int nt = s_nt;
  // cursors: first posting of each term at or after this warp's first doc (one binary search per
  // (warp, term); lanes take terms)
  for (int t = lane; t < nt; t += 32) {
    int64_t a = s_begin[t], e = endp[t];
    int64_t first_doc = w0 * kBmWin;
    cur[t] = (first_doc == 0 || w0 >= w1) ? a : lower_bound_doc(post_doc, a, e, (int32_t)min(first_doc, (int64_t)0x7FFFFFFF));
  }
  __syncwarp();

  int cbn = 0;                      // keys in this warp's buffer (warp-uniform)
  unsigned long long tau = 0;       // this warp's running threshold key (k-th best seen), 0 = none
  for (int64_t win = w0; win < w1; ++win) {
    const int32_t s0 = (int32_t)(win * kBmWin);
    const int32_t s_end = (int32_t)min((int64_t)N, (int64_t)s0 + kBmWin);
    bool any = false;
    for (int tb = 0; tb < nt; tb += kBmTermBatch) {
      int64_t c[kBmTermBatch];
      int32_t doc[kBmTermBatch];
      float imp[kBmTermBatch];
      // issue: first 32 postings of every term of the batch (independent loads, all in flight)
#pragma unroll
      for (int u = 0; u < kBmTermBatch; ++u) {
        const int t = tb + u;
        doc[u] = 0x7FFFFFFF;
        imp[u] = 0.f;
        c[u] = 0;
        if (t < nt) {
          c[u] = cur[t];
          const int64_t idx = c[u] + lane;
          if (idx < endp[t]) {
            doc[u] = __ldg(post_doc + idx);
            imp[u] = __ldg(post_imp + idx);
          }
        }
      }
      // apply: one term after the other (a doc can appear in several terms, never twice in one)
#pragma unroll
      for (int u = 0; u < kBmTermBatch; ++u) {
        const int t = tb + u;
        if (t >= nt) break;
        const float wt = wgt[t];
        const int64_t e = endp[t];
        for (;;) {
          const bool inr = doc[u] < s_end;
          if (inr) acc[doc[u] - s0] = fmaf(wt, imp[u], acc[doc[u] - s0]);
          __syncwarp();
          const int n = __popc(__ballot_sync(0xffffffffu, inr));
          c[u] += n;
          if (n) any = true;
          if (n < 32) break;
          const int64_t idx = c[u] + lane;   // list continues inside this window (long posting list)
          doc[u] = 0x7FFFFFFF;
          if (idx < e) {
            doc[u] = __ldg(post_doc + idx);
            imp[u] = __ldg(post_imp + idx);
          }
        }
        if (lane == 0) cur[t] = c[u];
      }
      __syncwarp();
    }
    if (!any) continue;
    // ---- read-and-clear pass over the window: 128 accumulators per warp instruction ----
    const unsigned long long cta_tau = *((volatile unsigned long long*)&s_tau);
    if (cta_tau > tau) tau = cta_tau;
    float tau_f = tau ? key_score(tau) : 0.f;
#pragma unroll 2
    for (int j = lane * 4; j < kBmWin; j += 128) {
      float4 v = *reinterpret_cast<float4*>(acc + j);
      const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      if (mx != 0.f) *reinterpret_cast<float4*>(acc + j) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (__any_sync(0xffffffffu, mx > 0.f && mx >= tau_f)) {
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          unsigned long long key = 0;
          bool take = false;
          if (vv[e4] > 0.f && vv[e4] >= tau_f) {
            key = make_key(vv[e4], (uint32_t)(s0 + j + e4));
            take = key > tau;
          }
          const unsigned m = __ballot_sync(0xffffffffu, take);
          if (m) {
            if (cbn + 32 > cbcap) {   // make room: keep the best kc (warp-uniform branch)
              for (int i = cbn + lane; i < cbcap; i += 32) cb[i] = 0;
              warp_bitonic_desc(cb, cbcap, lane);
              cbn = min(cbn, kc);
              if (cbn == kc) {
                tau = cb[kc - 1];
                tau_f = key_score(tau);
                if (lane == 0) atomicMax(&s_tau, tau);
              }
              take = take && key > tau;
            }
            const unsigned m2 = __ballot_sync(0xffffffffu, take);
            if (take) cb[cbn + __popc(m2 & ((1u << lane) - 1u))] = key;
            cbn += __popc(m2);
            __syncwarp();
          }
        }
      }
    }
  }
  // ---- warp list -> sorted top-kc ----
  for (int i = cbn + lane; i < cbcap; i += 32) cb[i] = 0;
  warp_bitonic_desc(cb, cbcap, lane);
  cbn = min(cbn, kc);
  if (lane == 0) s_wn[w] = cbn;
  __syncthreads();
  // ---- CTA merge of the 8 warp lists (reuse the accumulator area as the sort buffer) ----
  uint64_t* mbuf = (uint64_t*)acc_all;   // 8 * kcp keys <= 8 KB * ... fits in 32 KB
  const int total = kBmWarps * kcp;
  for (int i = tid; i < total; i += kBmThreads) {
    const int ww = i / kcp, j = i - ww * kcp;
    mbuf[i] = (j < s_wn[ww]) ? cb_all[(size_t)ww * 2 * kcp + j] : 0ull;
  }
  block_bitonic_desc(mbuf, total);
  uint64_t* o = out_keys + ((size_t)q * S + g) * kc;
  int m = 0;
  for (int ww = 0; ww < kBmWarps; ++ww) m += s_wn[ww];
  m = min(m, kc);
  for (int j = tid; j < kc; j += kBmThreads) o[j] = (j < m) ? mbuf[j] : 0ull;
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
