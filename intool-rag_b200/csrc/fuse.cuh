// fuse.cuh — hybrid fusion (weighted-score / RRF), final top-k, and the cross-shard k-way merge.
// Latency-bound warp/block-primitive kernels on (nq x a few hundred) candidates.
//
// Definitions: oracle/fusion.py.  Only reference-side pins: the weights
// 0.7 / 0.3 (rag/config.py:44-45) and the dense transform clamp(1 - d/2, 0, 1)
// (rag/storage/faiss_index.py:86-88).  Order everywhere: (score desc, id asc).
#pragma once
#include "common.cuh"

namespace hr {

constexpr int kFuseMaxKc = 256;

// better(a,b): a ranks strictly before b
__device__ __forceinline__ bool ranks_before(float sa, int64_t ia, float sb, int64_t ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}
// total order for rank-by-counting: entries with an equal (score, id) — duplicate ids from overlapping shards —
// are ordered by their position, so the ranks are a permutation and every output slot is written
__device__ __forceinline__ bool ranks_before_pos(float sa, int64_t ia, int pa, float sb, int64_t ib, int pb) {
  return (sa > sb) || (sa == sb && (ia < ib || (ia == ib && pa < pb)));
}

// One block per query.  Entries [0,kc) = dense list, [kc,2kc) = sparse list (only sparse-only docs valid).
__global__ void __launch_bounds__(256)
fuse_kernel(const float* __restrict__ dD, const int64_t* __restrict__ dI, const float* __restrict__ bS,
            const int64_t* __restrict__ bI, const float* __restrict__ bmax, int kc, int top_k, int metric,
            int mode, float w_vec, float w_bm25, float* __restrict__ oS, int64_t* __restrict__ oI) {
  __shared__ float fs[2 * kFuseMaxKc];
  __shared__ int64_t fid[2 * kFuseMaxKc];
  __shared__ int s_valid;
  const int q = blockIdx.x;
  const float* dDq = dD + (size_t)q * kc;
  const int64_t* dIq = dI + (size_t)q * kc;
  const float* bSq = bS + (size_t)q * kc;
  const int64_t* bIq = bI + (size_t)q * kc;
  // both id lists once into shared memory: every entry is looked up in the other list
  __shared__ int64_t sdi[kFuseMaxKc], sbi[kFuseMaxKc];
  if (threadIdx.x == 0) s_valid = 0;
  for (int e = threadIdx.x; e < kc; e += blockDim.x) {
    sdi[e] = dIq[e];
    sbi[e] = bIq[e];
  }
  __syncthreads();
  float mx = 0.f;
  if (bmax) mx = bmax[q];
  else if (sbi[0] >= 0) mx = bSq[0];  // lists are best-first
  const int n2 = 2 * kc;
  for (int e = threadIdx.x; e < n2; e += blockDim.x) {
    float f = 0.f;
    int64_t id = -1;
    if (e < kc) {
      id = sdi[e];
      if (id >= 0) {
        int sp = -1;
        for (int j = 0; j < kc; ++j)
          if (sbi[j] == id) { sp = j; break; }
        if (mode == 0) {
          float sim = (metric == 1) ? (1.0f - dDq[e] * 0.5f) : dDq[e];
          sim = fminf(1.0f, fmaxf(0.0f, sim));
          f = w_vec * sim;
          if (sp >= 0 && mx > 0.f) f = fmaf(w_bm25, bSq[sp] / mx, f);
        } else {
          f = 1.0f / (60.0f + (float)(e + 1));
          if (sp >= 0) f += 1.0f / (60.0f + (float)(sp + 1));
        }
      }
    } else {
      const int j = e - kc;
      id = sbi[j];
      if (id >= 0) {
        bool in_dense = false;
        for (int i = 0; i < kc; ++i)
          if (sdi[i] == id) { in_dense = true; break; }
        if (in_dense) id = -1;
        else if (mode == 0) f = (mx > 0.f) ? w_bm25 * (bSq[j] / mx) : 0.f;
        else f = 1.0f / (60.0f + (float)(j + 1));
      }
    }
    fs[e] = f;
    fid[e] = id;
    if (id >= 0) atomicAdd(&s_valid, 1);
  }
  __syncthreads();
  const int nv = s_valid;
  for (int e = threadIdx.x; e < n2; e += blockDim.x) {
    const int64_t id = fid[e];
    if (id < 0) continue;
    const float f = fs[e];
    int rank = 0;
    for (int j = 0; j < n2; ++j) {
      const int64_t oj = fid[j];
      if (oj >= 0 && ranks_before_pos(fs[j], oj, j, f, id, e)) rank++;
    }
    if (rank < top_k) {
      oS[(size_t)q * top_k + rank] = f;
      oI[(size_t)q * top_k + rank] = id;
    }
  }
  for (int j = nv + threadIdx.x; j < top_k; j += blockDim.x) {
    oS[(size_t)q * top_k + j] = 0.f;
    oI[(size_t)q * top_k + j] = -1;
  }
}

// k-way merge of candidate lists -> best k (largest or smallest first), id asc ties.
// Candidate e of query q is entry c = e % kc of list l = e / kc, stored at l * list_stride + q * kc + c
// (one list with kc = n_cand: the plain [nq][n_cand] layout; several lists: the all-gathered per-rank
// buffers, read in place).
constexpr int kMergeTopkCap = 2048;
// sorted_lists != 0: every list is already best-first with its padding at the end (what hr_index_search /
// hr_bm25_search produce): an entry's rank is its position in its own list plus, for every other list, the number
// of entries that precede it there (one binary search each) — O(n L log kc) instead of the O(n^2) counting that an
// arbitrary candidate set needs.  Equal (score, id) entries are ordered by list, then position.
__global__ void __launch_bounds__(256)
merge_topk_kernel(const float* __restrict__ S, const int64_t* __restrict__ I, int n_cand, int kc, int64_t s_stride,
                  int64_t i_stride, int k, int largest, float pad_score, float* __restrict__ oS,
                  int64_t* __restrict__ oI, int sorted_lists) {
  __shared__ float ss[kMergeTopkCap];
  __shared__ int64_t si[kMergeTopkCap];
  __shared__ int s_valid;
  const int q = blockIdx.x;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  for (int e = threadIdx.x; e < n_cand; e += blockDim.x) {
    const int l = e / kc, c = e - l * kc;
    float s = S[(size_t)l * s_stride + (size_t)q * kc + c];
    int64_t id = I[(size_t)l * i_stride + (size_t)q * kc + c];
    if (s != s) s = pad_score;   // NaN has no place in a total order: it ranks like padding
    ss[e] = largest ? s : -s;
    si[e] = id;
    if (id >= 0) atomicAdd(&s_valid, 1);
  }
  __syncthreads();
  const int nv = s_valid;
  const int n_lists = n_cand / kc;
  for (int e = threadIdx.x; e < n_cand; e += blockDim.x) {
    const int64_t id = si[e];
    if (id < 0) continue;
    const float f = ss[e];
    int rank = 0;
    if (sorted_lists) {
      const int l = e / kc;
      rank = e - l * kc;
      for (int m = 0; m < n_lists; ++m) {
        if (m == l) continue;
        // first entry of list m that does NOT precede (f, id): padding never precedes
        int lo = 0, hi = kc;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const int64_t oj = si[m * kc + mid];
          const bool before = oj >= 0 && (ranks_before(ss[m * kc + mid], oj, f, id) ||
                                          (m < l && ss[m * kc + mid] == f && oj == id));
          if (before) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
    } else {
      for (int j = 0; j < n_cand; ++j) {
        const int64_t oj = si[j];
        if (oj >= 0 && ranks_before_pos(ss[j], oj, j, f, id, e)) rank++;
      }
    }
    if (rank < k) {
      oS[(size_t)q * k + rank] = largest ? f : -f;
      oI[(size_t)q * k + rank] = id;
    }
  }
  for (int j = nv + threadIdx.x; j < k; j += blockDim.x) {
    oS[(size_t)q * k + j] = pad_score;
    oI[(size_t)q * k + j] = -1;
  }
}

// ---- page-level ranking of a batch of hit lists --------------------------------------
// What /root/reference/rag/query/page_retriever.py:145-236 does per request on the host, for every query of a
// batch on the device: group the k hits by page (first appearance order), page score = mean(hit scores) +
// min(0.05 n, 0.15), stable sort by score descending, first top_pages.  Double precision with the reference's
// summation order (hits best first), so the values equal the Python floats of the reference bit for bit.
// score_kind 0: S is the hit score (fused / inner product); 1: S is a squared L2 distance and the hit score is
// the wrapper's clamp(1 - d/2, 0, 1) (rag/storage/faiss_index.py:86-88).  One block per query.
constexpr int kPageMaxHits = 256;
__global__ void __launch_bounds__(128)
page_rank_kernel(const float* __restrict__ S, const int64_t* __restrict__ I, int k, int score_kind,
                 const int32_t* __restrict__ page_of_row, int64_t n_rows, int64_t id_base, int top_pages,
                 int32_t* __restrict__ oP, double* __restrict__ oS, int32_t* __restrict__ oC) {
  __shared__ int pg[kPageMaxHits];
  __shared__ double sc[kPageMaxHits];
  __shared__ double ps[kPageMaxHits];
  __shared__ int pc[kPageMaxHits];     // hits of the page led by entry e, 0 = e is not a page's first hit
  const int q = blockIdx.x;
  for (int e = threadIdx.x; e < k; e += blockDim.x) {
    const int64_t id = I[(size_t)q * k + e];
    const int64_t row = id - id_base;
    const bool valid = id >= 0 && row >= 0 && row < n_rows;
    pg[e] = valid ? page_of_row[row] : 0;
    double v = (double)S[(size_t)q * k + e];
    if (score_kind == 1) v = fmax(0.0, fmin(1.0, 1.0 - v / 2.0));
    sc[e] = v;
    pc[e] = valid ? -1 : -2;            // -1: valid hit, leader unknown; -2: padding
  }
  __syncthreads();
  for (int e = threadIdx.x; e < k; e += blockDim.x) {
    if (pc[e] == -2) continue;
    bool leader = true;
    for (int j = 0; j < e; ++j)
      if (pc[j] != -2 && pg[j] == pg[e]) { leader = false; break; }
    if (!leader) continue;
    double sum = 0.0;
    int n = 0;
    for (int j = e; j < k; ++j)
      if (pc[j] != -2 && pg[j] == pg[e]) { sum += sc[j]; ++n; }
    ps[e] = sum / (double)n + fmin((double)n * 0.05, 0.15);
    pc[e] = -(n + 2);                   // leaders: -(n + 2) <= -3 until every thread has finished reading pc
  }
  __syncthreads();
  for (int e = threadIdx.x; e < k; e += blockDim.x) pc[e] = pc[e] <= -3 ? -(pc[e] + 2) : 0;
  __syncthreads();
  int npages = 0;
  for (int j = 0; j < k; ++j) npages += pc[j] > 0;
  for (int e = threadIdx.x; e < k; e += blockDim.x) {
    if (pc[e] <= 0) continue;
    int rank = 0;                        // stable: equal scores keep the order of first appearance
    for (int j = 0; j < k; ++j)
      if (pc[j] > 0 && (ps[j] > ps[e] || (ps[j] == ps[e] && j < e))) rank++;
    if (rank < top_pages) {
      oP[(size_t)q * top_pages + rank] = pg[e];
      oS[(size_t)q * top_pages + rank] = ps[e];
      oC[(size_t)q * top_pages + rank] = pc[e];
    }
  }
  for (int j = npages + threadIdx.x; j < top_pages; j += blockDim.x) {
    oP[(size_t)q * top_pages + j] = -1;
    oS[(size_t)q * top_pages + j] = 0.0;
    oC[(size_t)q * top_pages + j] = 0;
  }
}

}  // namespace hr
