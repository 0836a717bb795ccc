// bm25_build.cuh — ingest side of the BM25 index: flat (term, doc) token
// occurrences -> CSR by term, on the device.  The reference tokenises with `text.lower().split()`
// (rag/agent/query_processor.py:26) and builds no sparse index at all (rag/ingest/ingestion_pipeline.py:79-94).
//
// key = term << 32 | doc;  radix sort (cub, a library primitive outside the query hot path);  run-length
// encode -> unique (term, doc) pairs with tf;  indptr by one lower_bound per term.  Integer work: the result
// is bit-identical to the host builder (numpy `unique` of term * N + doc, intool-rag_b200/bm25.py:build_csr).
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"

namespace hr {

__global__ void tokens_to_keys_kernel(const int32_t* __restrict__ term_ids, const int32_t* __restrict__ doc_ids,
                                      int64_t n, int64_t vocab, int64_t n_docs, uint64_t* __restrict__ keys,
                                      int32_t* __restrict__ doc_len, unsigned long long* __restrict__ n_bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t t = term_ids[i], d = doc_ids[i];
    if (t < 0 || t >= vocab || d < 0 || d >= n_docs) {
      atomicAdd(n_bad, 1ull);
      keys[i] = ~0ull;   // sorts last; never reaches the CSR because the build fails
    } else {
      keys[i] = ((uint64_t)(uint32_t)t << 32) | (uint32_t)d;
      atomicAdd(doc_len + d, 1);
    }
  }
}

__global__ void unique_to_postings_kernel(const uint64_t* __restrict__ uniq, const int32_t* __restrict__ counts,
                                          int64_t nnz, int32_t* __restrict__ post_doc, int32_t* __restrict__ post_tf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    post_doc[i] = (int32_t)(uint32_t)uniq[i];
    post_tf[i] = counts[i];
  }
}

// indptr[t] = first unique key with term >= t, t = 0..V
__global__ void term_offsets_kernel(const uint64_t* __restrict__ uniq, int64_t nnz, int64_t vocab,
                                    int64_t* __restrict__ indptr) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > vocab) return;
  const uint64_t bound = (uint64_t)t << 32;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (uniq[mid] < bound) lo = mid + 1; else hi = mid;
  }
  indptr[t] = lo;
}

}  // namespace hr
