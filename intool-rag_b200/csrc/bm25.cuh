// bm25.cuh — BM25 scoring over a CSR-by-term inverted index (HBM-bound integer/scatter work).
//
// The reference advertises BM25 (README.md:54-58, rag/config.py:43-45) but implements none; the
// definition is oracle/bm25.py.  Layout in HBM (struct of arrays):
//   indptr int64[V+1] | post_doc int32[nnz] (ascending per term) | post_imp fp32[nnz]
// post_imp is the length-normalised saturation tf(k1+1)/(tf+k1(1-b+b dl/avgdl)) folded at build
// time, so one posting costs 8 streamed bytes and score(q,d) = sum_t mult(t) idf(t) imp(t,d).
//
// A search is three kernels:
//   bm25_plan_terms_kernel   per query: de-duplicate the terms (first-occurrence order, multiplicity
//                            folded into the weight), drop empty / out-of-vocabulary ones.
//   bm25_plan_cursors_kernel per (query, term, window boundary): lower_bound of the boundary's first
//                            doc id in the term's posting list.  The document axis is cut into
//                            windows of kBwWin docs; with the cursor table every (query, window) is
//                            an independent, exactly-known set of posting ranges.
//   bm25_window_kernel       CTA = (query, span of consecutive windows).  The window's accumulators
//                            (kBwWin fp32) live in shared memory.  The posting ranges of the window
//                            are cut into slots of 256 postings; a thread loads its posting of up to
//                            kBwBatch slots at once (coalesced, all loads in flight before the first
//                            use), then applies them with plain shared-memory read-modify-writes:
//                            doc ids inside one posting list are unique, and a block barrier
//                            separates different terms, so no atomics are needed and the fp32
//                            summation order is fixed (term order) - results are deterministic.
//                            At the end of a window the warps clear their slice of the accumulators
//                            and, only if some score reached the running threshold, extract
//                            candidates into warp-private key buffers (bitonic compaction).
#pragma once
#include "common.cuh"
#include "dense_exact.cuh"

namespace hr {

constexpr int kBmMaxTerms = 64;   // raw terms per query
constexpr int kBmMaxK = 128;      // candidate depth the kernel supports
constexpr int kBwThreads = 256;
constexpr int kBwWarps = kBwThreads / 32;
constexpr int kBwWin = 16384;     // docs per window (64 KB of accumulators)
constexpr int kBwSlice = kBwWin / kBwWarps;
constexpr int kBwBatch = 8;       // posting slots a thread keeps in flight
// dynamic shared memory: accumulators | warp key buffers (2*kcp keys each)
__host__ __device__ constexpr int bw_smem_bytes(int kcp) { return kBwWin * 4 + kBwWarps * 2 * kcp * 8; }

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
bm25_plan_terms_kernel(const int64_t* __restrict__ indptr, const float* __restrict__ idf, int64_t V,
                       const int32_t* __restrict__ q_indptr, const int32_t* __restrict__ q_terms, int nq,
                       int* __restrict__ plan_nt, int64_t* __restrict__ plan_start, uint32_t* __restrict__ plan_len,
                       float* __restrict__ plan_wgt, unsigned long long* __restrict__ postings_touched) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const int qa = q_indptr[q];
  const int nraw = min(q_indptr[q + 1] - qa, kBmMaxTerms);
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
      for (int j = 0; j < nraw; ++j) {
        const int u = q_terms[qa + j];
        if (u == t) {
          if (j < i) keep = false;
          mult++;
        }
      }
      if (keep) {
        a = indptr[t];
        e = indptr[t + 1];
        keep = e > a;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int slot = qa + base + __popc(m & ((1u << lane) - 1u));
      plan_start[slot] = a;
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
// cur[(size_t)q_indptr[q] * (nwin + 1) + j * nt + u] = number of postings of term u with doc < j * kBwWin.
// Thread = boundary j (consecutive threads search consecutive boundaries of the same list, so the upper
// levels of the binary searches share cache lines), looping over the query's terms.
__global__ void __launch_bounds__(256)
bm25_plan_cursors_kernel(const int32_t* __restrict__ post_doc, const int32_t* __restrict__ q_indptr,
                         const int* __restrict__ plan_nt, const int64_t* __restrict__ plan_start,
                         const uint32_t* __restrict__ plan_len, int64_t nwin, uint32_t* __restrict__ cur) {
  const int q = blockIdx.x;
  const int nt = plan_nt[q];
  const int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (j > nwin || nt == 0) return;
  const int qa = q_indptr[q];
  uint32_t* out = cur + (size_t)qa * (size_t)(nwin + 1) + (size_t)j * nt;
  const int64_t bound64 = j * (int64_t)kBwWin;
  for (int u = 0; u < nt; ++u) {
    const uint32_t len = plan_len[qa + u];
    uint32_t lo = 0, hi = len;
    if (j == 0) hi = 0;
    else if (j == nwin || bound64 > 0x7FFFFFFFll) lo = len;
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
// grid = (nq, S): query fastest, so the first wave holds span 0 of many queries and later spans start
// from the thresholds earlier spans published in tau_g.  out_keys [nq][S][kc], out_n [nq][S].
// kcp = power of two >= max(kc, 32); a warp's key buffer holds 2*kcp keys.
__global__ void __launch_bounds__(kBwThreads, 3)
bm25_window_kernel(const int32_t* __restrict__ post_doc, const float* __restrict__ post_imp,
                   const int32_t* __restrict__ q_indptr, const int* __restrict__ plan_nt,
                   const int64_t* __restrict__ plan_start, const float* __restrict__ plan_wgt,
                   const uint32_t* __restrict__ plan_cur, int64_t nwin, int wpc, int S, int kc, int kcp,
                   uint64_t* __restrict__ out_keys, int* __restrict__ out_n, unsigned long long* __restrict__ tau_g) {
  extern __shared__ __align__(16) uint8_t bsm[];
  float* acc = (float*)bsm;
  uint64_t* cb_all = (uint64_t*)(bsm + kBwWin * 4);
  __shared__ int64_t s_start[kBmMaxTerms];
  __shared__ float s_w[kBmMaxTerms];
  __shared__ uint32_t s_lo[kBmMaxTerms], s_hi[kBmMaxTerms];
  __shared__ unsigned long long s_tau;
  __shared__ int s_wn[kBwWarps];

  const int q = blockIdx.x;
  const int g = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int w = tid >> 5;
  uint64_t* cb = cb_all + (size_t)w * 2 * kcp;
  const int cbcap = 2 * kcp;
  const int nt = plan_nt[q];
  const int64_t win0 = (int64_t)g * wpc;
  const int64_t win1 = min(nwin, win0 + wpc);
  if (nt == 0 || win0 >= win1) {   // uniform
    if (tid == 0) out_n[(size_t)q * S + g] = 0;
    return;
  }
  const int qa = q_indptr[q];
  const uint32_t* curq = plan_cur + (size_t)qa * (size_t)(nwin + 1);
  if (tid < nt) {
    s_start[tid] = plan_start[qa + tid];
    s_w[tid] = plan_wgt[qa + tid];
    s_lo[tid] = curq[(size_t)win0 * nt + tid];
    s_hi[tid] = curq[(size_t)(win0 + 1) * nt + tid];
  }
  if (tid == 0) s_tau = *((volatile unsigned long long*)(tau_g + q));
  for (int i = tid * 4; i < kBwWin; i += kBwThreads * 4) *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  int cbn = 0;                 // keys in this warp's buffer (warp-uniform)
  unsigned long long tau = 0;  // this warp's threshold key: a lower bound of the query's kc-th best
  float* slice = acc + w * kBwSlice;
  for (int64_t win = win0; win < win1; ++win) {
    const int32_t docbase = (int32_t)(win * kBwWin);
    // cursors of the next window (consumed at the end of this one)
    uint32_t nxt = 0;
    if (tid < nt && win + 2 <= nwin) nxt = __ldg(curq + (size_t)(win + 2) * nt + tid);
    unsigned long long gt = 0;
    if (tid == 0) gt = *((volatile unsigned long long*)(tau_g + q));
    float wmax = 0.f;
    // ---- apply: slots of 256 postings, kBwBatch of them in flight per thread ----
    int t = 0;
    uint32_t c = 0;
    bool first = true;
    while (t < nt && s_hi[t] == s_lo[t]) ++t;
    while (t < nt) {
      int32_t dd[kBwBatch];
      float vv[kBwBatch], ww[kBwBatch];
      unsigned newterm = 0;
#pragma unroll
      for (int j = 0; j < kBwBatch; ++j) {
        dd[j] = -1;
        vv[j] = 0.f;
        ww[j] = 0.f;
        if (t < nt) {
          const uint32_t lo = s_lo[t], hi = s_hi[t];
          const uint32_t idx = lo + c + (uint32_t)tid;
          if (c == 0 && !first) newterm |= 1u << j;
          first = false;
          ww[j] = s_w[t];
          if (idx < hi) {
            const int64_t p = s_start[t] + idx;
            dd[j] = __ldg(post_doc + p);
            vv[j] = __ldg(post_imp + p);
          }
          c += kBwThreads;
          if (lo + c >= hi) {
            c = 0;
            do { ++t; } while (t < nt && s_hi[t] == s_lo[t]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kBwBatch; ++j) {
        if ((newterm >> j) & 1u) __syncthreads();   // uniform: a different term may touch the same docs
        if (dd[j] >= 0) {
          float* a = acc + (dd[j] - docbase);
          const float x = fmaf(ww[j], vv[j], *a);
          *a = x;
          wmax = fmaxf(wmax, x);
        }
      }
    }
    // ---- end of window ----
    if (tid == 0 && gt > s_tau) s_tau = gt;   // only thread 0 and compacting warps write s_tau (atomicMax below)
    const float tau_f0 = tau ? key_score(tau) : 0.f;
    const int any = __syncthreads_or(wmax > 0.f && wmax >= tau_f0);
    if (tid < nt) {   // rotate the cursors; ordered before the next window's reads by the closing barrier
      s_lo[tid] = s_hi[tid];
      s_hi[tid] = nxt;
    }
    {
      const unsigned long long ct = *((volatile unsigned long long*)&s_tau);
      if (ct > tau) tau = ct;
    }
    if (!any) {
#pragma unroll
      for (int j = lane * 4; j < kBwSlice; j += 128) *reinterpret_cast<float4*>(slice + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      float tau_f = tau ? key_score(tau) : 0.f;
      const int32_t s0 = docbase + w * kBwSlice;
#pragma unroll 2
      for (int j = lane * 4; j < kBwSlice; j += 128) {
        float4 v = *reinterpret_cast<float4*>(slice + j);
        const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        *reinterpret_cast<float4*>(slice + j) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (__any_sync(0xffffffffu, mx > 0.f && mx >= tau_f)) {
          const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            unsigned long long key = 0;
            bool take = false;
            if (ve[e4] > 0.f && ve[e4] >= tau_f) {
              key = make_key(ve[e4], (uint32_t)(s0 + j + e4));
              take = key > tau;
            }
            const unsigned m = __ballot_sync(0xffffffffu, take);
            if (m) {
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
                      atomicMax(&s_tau, tau);
                      atomicMax(tau_g + q, tau);
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
          }
        }
      }
    }
    __syncthreads();   // clears and cursor rotation visible before the next window
  }
  // ---- warp list -> sorted top-kc ----
  for (int i = cbn + lane; i < cbcap; i += 32) cb[i] = 0;
  warp_bitonic_desc(cb, cbcap, lane);
  cbn = min(cbn, kc);
  if (lane == 0) {
    s_wn[w] = cbn;
    if (cbn == kc && cb[kc - 1] > tau) atomicMax(tau_g + q, cb[kc - 1]);
  }
  __syncthreads();
  // ---- CTA merge of the warp lists (the accumulator area is the sort buffer) ----
  uint64_t* mbuf = (uint64_t*)acc;
  const int total = kBwWarps * kcp;
  for (int i = tid; i < total; i += kBwThreads) {
    const int ww = i / kcp, j = i - ww * kcp;
    mbuf[i] = (j < s_wn[ww]) ? cb_all[(size_t)ww * 2 * kcp + j] : 0ull;
  }
  block_bitonic_desc(mbuf, total);
  uint64_t* o = out_keys + ((size_t)q * S + g) * kc;
  int m = 0;
  for (int ww = 0; ww < kBwWarps; ++ww) m += s_wn[ww];
  m = min(m, kc);
  for (int j = tid; j < kc; j += kBwThreads) o[j] = (j < m) ? mbuf[j] : 0ull;
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
