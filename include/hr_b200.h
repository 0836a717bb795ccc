/*
 * hr_b200.h — C ABI of the B200-native hybrid retrieval backend (libhr_b200.so).
 *
 * This is the drop-in boundary for intool-rag's query hot path.  Every entry point is
 * `extern "C"`, takes plain pointers / sizes (no torch or C++ types) and returns an int
 * status: 0 = ok, < 0 = error (text via hr_last_error(), thread-local).  The reference has
 * no native FFI of its own: its hot path is the Python call sequence into the third-party
 * faiss-cpu 1.7.4 wheel.  Each group below cites the reference call site it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes binding.
 *
 * Ownership: the caller owns every input/output buffer; a handle owns its corpus /
 * postings device memory and its scratch.  A handle is bound to one CUDA device.  Calls on
 * one handle are serialised by a mutex inside the handle (the scratch is per handle), so
 * concurrent searches from several threads are safe, like faiss-cpu's IndexFlat::search; use
 * several handles for concurrency.  (The reference adds at ingest into a new index,
 * rag/ingest/ingestion_pipeline.py:88, and searches from one event-loop thread.)  There is no CPU fallback: without a CUDA device
 * every compute entry point fails with HR_ERR_CUDA.
 */
#ifndef HR_B200_H
#define HR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HR_OK 0
#define HR_ERR_INVALID (-1) /* bad argument (shape, dtype, k, null pointer)            */
#define HR_ERR_CUDA (-2)    /* CUDA runtime / driver error, or no device               */
#define HR_ERR_IO (-3)      /* file could not be read / written / parsed               */
#define HR_ERR_NOMEM (-4)   /* device or host allocation failed                        */

/* metric_type values are faiss's (faiss/MetricType.h): the reference builds IndexFlatL2
 * (rag/storage/faiss_index.py:123); BASELINE.json names IndexFlatIP.                      */
#define HR_METRIC_INNER_PRODUCT 0
#define HR_METRIC_L2 1

#define HR_STORAGE_F32 0  /* fp32 rows, TF32 tensor-core filter + exact fp32 re-score      */
#define HR_STORAGE_BF16 1 /* bf16 rows (fp32 accumulate), bf16 filter + exact re-score    */
/* fp32 rows (results, reconstruct and save identical to HR_STORAGE_F32) plus a bf16 shadow copy
 * that the tensor-core filter streams instead: half the scan bytes, twice the MMA rate,
 * 1.5x the memory.  The exact fp32 re-score + certificate make the answers the same. */
#define HR_STORAGE_F32_SHADOW16 2

/* dense search strategy (hr_index_set_mode) */
#define HR_MODE_AUTO 0      /* tcgen05 filter scan + exact re-score + certified fallback  */
#define HR_MODE_EXACT_SIMT 1 /* exhaustive exact fp32 CUDA-core scan (the fallback path)   */

#define HR_FUSE_WEIGHTED 0
#define HR_FUSE_RRF 1

#define HR_IDF_LUCENE 0
#define HR_IDF_OKAPI 1

#define HR_MAX_K 2048

typedef struct hr_index hr_index; /* flat dense index                                      */
typedef struct hr_bm25 hr_bm25;   /* BM25 inverted index (CSR by term)                     */
typedef struct hr_comm hr_comm;   /* one rank of a row-sharded corpus (NCCL communicator)  */

typedef struct hr_scan_stats {
  int64_t launches;         /* kernels launched by the last search on this handle        */
  int64_t flagged;          /* queries the certificate sent to the exact fallback        */
  int64_t overflow;         /* queries whose candidate merge overflowed (also fallback)  */
  float scan_ms;            /* CUDA-event time of the filter-scan kernel (0 if not timed)*/
  float total_ms;           /* CUDA-event time of the whole search                       */
  int mode_used;            /* HR_MODE_* actually taken                                  */
  int list_len;             /* per-CTA candidate list length used by the filter          */
  int grid;                 /* CTAs of the scan kernel                                   */
  int deeper;               /* queries certified only after re-scoring all their candidates */
} hr_scan_stats;

/* ---- library ----------------------------------------------------------------------- */
const char* hr_last_error(void);
int hr_version(void);
/* first 16 hex digits of the sha256 over the library's sources (csrc/Makefile: HASH), compiled in at build time */
const char* hr_source_hash(void);
int hr_device_count(int* out);
/* kernels launched by this library in this process (monotonic; bench.py's gpu_launches). */
int64_t hr_launch_count(void);
/* tuning / diagnostic knobs by name = the HR_* environment variables without the prefix, lower case
 * ("bm25_spans", "bm25_wide", "pre_tiles", ...).  Benchmarks and tests only. */
int hr_set_option(const char* name, int value);

/* ---- flat dense index: replaces faiss.IndexFlatL2/IP ---------------------------------
 * create  <- faiss.IndexFlatL2(d)            rag/storage/faiss_index.py:123
 * add     <- index.add(float32[n,d])         rag/storage/faiss_index.py:124
 * search  <- index.search(float32[nq,d], k)  rag/storage/faiss_index.py:83,
 *                                            rag/agent/search_engine.py:45
 * d/ntotal<- index.d / index.ntotal          rag/storage/faiss_index.py:58-59,97,103,126
 * save    <- faiss.write_index(index, path)  rag/storage/faiss_index.py:133
 * load    <- faiss.read_index(path)          rag/storage/faiss_index.py:54,
 *                                            rag/agent/search_engine.py:30
 */
int hr_index_create(int d, int metric, int storage_dtype, int device, hr_index** out);
int hr_index_destroy(hr_index* h);
int hr_index_reserve(hr_index* h, int64_t n_rows);
/* x: float32 [n, d] C-contiguous; is_device != 0 means x is a device pointer on h's device. */
int hr_index_add(hr_index* h, const float* x, int64_t n, int is_device, void* stream);
int hr_index_reset(hr_index* h);
int64_t hr_index_ntotal(const hr_index* h);
int hr_index_d(const hr_index* h);
int hr_index_metric(const hr_index* h);
int hr_index_storage(const hr_index* h);
/* ids returned by search are row + id_base (row sharding: rank r sets its first global row). */
int hr_index_set_id_base(hr_index* h, int64_t id_base);
int hr_index_set_mode(hr_index* h, int mode);
/* copy rows [i0, i0+n) back as float32 [n, d] into host memory. */
int hr_index_reconstruct(hr_index* h, int64_t i0, int64_t n, float* out_host);
/* D: float32 [nq,k]; I: int64 [nq,k].  L2: squared distances ascending; IP: inner products
 * descending; fewer than k rows -> I=-1, D=+FLT_MAX (L2) / -FLT_MAX (IP).  io_on_device != 0:
 * q, D, I are device pointers and the call is ordered on `stream`; otherwise they are host
 * pointers and the call copies in/out itself.  Returns after the results are complete. */
int hr_index_search(hr_index* h, const float* q, int64_t nq, int k, float* D, int64_t* I,
                    int io_on_device, void* stream);
int hr_index_last_stats(const hr_index* h, hr_scan_stats* out);
/* test/diagnostic hook: copies the filter's internals of the LAST search batch to host memory:
 * lists (score fp32,row u32) [grid][nq][list_len], cnts int32 [grid][nq], tau uint32 [nq] (ordered
 * float, 0 = unset), short_rows uint32 [nq][list_len], tprime fp32 [nq].  Any pointer may be NULL. */
int hr_index_debug_dump(hr_index* h, int64_t nq, void* lists, int32_t* cnts, uint32_t* tau,
                        uint32_t* short_rows, float* tprime);
/* faiss flat-index bytes ("IxF2"/"IxFI" header + raw fp32 rows), SURVEY.md Appendix A. */
int hr_index_save(hr_index* h, const char* path);
int hr_index_load(const char* path, int device, int storage_dtype, hr_index** out);

/* ---- BM25 (README.md:54-58 advertises it; the reference has no implementation) ---------
 * CSR by term: indptr int64[V+1], post_doc int32[nnz] (ascending inside a list),
 * post_tf int32[nnz], doc_len int32[N].  Arrays are host or device (is_device).  Global
 * statistics (n_docs_global, avgdl_global, df_global int64[V] or NULL = local) let a row
 * shard score with corpus-wide idf / avgdl.  */
int hr_bm25_create(const int64_t* indptr, const int32_t* post_doc, const int32_t* post_tf,
                   const int32_t* doc_len, int64_t n_docs, int64_t vocab, float k1, float b,
                   int idf_variant, int64_t n_docs_global, double avgdl_global,
                   const int64_t* df_global, int is_device, int device, void* stream,
                   hr_bm25** out);
/* Ingest side: build the CSR on the device from flat token occurrences term_ids /
 * doc_ids int32[n_tokens] (host or device; the reference tokenises with text.lower().split(),
 * rag/agent/query_processor.py:26, and builds no sparse index: rag/ingest/ingestion_pipeline.py:79-94).
 * doc_len is the number of occurrences per doc.  Same statistics arguments as hr_bm25_create. */
int hr_bm25_create_from_tokens(const int32_t* term_ids, const int32_t* doc_ids, int64_t n_tokens,
                               int64_t n_docs, int64_t vocab, float k1, float b, int idf_variant,
                               int64_t n_docs_global, double avgdl_global, const int64_t* df_global,
                               int is_device, int device, void* stream, hr_bm25** out);
/* Persistence of the BM25 index (the sidecar next to the faiss file; the reference persists no sparse
 * index): "HRBM25" v1 = header + indptr + idf + postings + folded impacts. */
int hr_bm25_save(hr_bm25* h, const char* path);
int hr_bm25_load(const char* path, int device, hr_bm25** out);
int hr_bm25_destroy(hr_bm25* h);
int64_t hr_bm25_ndocs(const hr_bm25* h);
int64_t hr_bm25_vocab(const hr_bm25* h);
int64_t hr_bm25_nnz(const hr_bm25* h);
int hr_bm25_set_id_base(hr_bm25* h, int64_t id_base);
/* queries as CSR: q_indptr int32[nq+1], q_terms int32[n_terms] with n_terms >= q_indptr[nq]
 * (duplicates count, ids outside [0,V) ignored; any length, but at most 64 DISTINCT scorable terms per
 * query: more is HR_ERR_INVALID, for host and device queries alike).  n_terms < 0 = unknown
 * (device io then reads q_indptr[nq] back, one small synchronous copy).  S float32[nq,k]
 * descending, I int64[nq,k], padding I=-1,S=0.  postings_touched (optional, host int64)
 * receives the number of postings scored. */
int hr_bm25_search(hr_bm25* h, const int32_t* q_indptr, const int32_t* q_terms, int64_t nq,
                   int64_t n_terms, int k, float* S, int64_t* I, int io_on_device, void* stream,
                   int64_t* postings_touched);

/* ---- fusion + merges (all pointers are DEVICE pointers, ordered on `stream`) ----------- */
/* k-way merge of per-shard candidate lists S,I [nq, n_lists*kc] -> best k by (score best
 * first, id asc); ids < 0 are padding.  largest != 0: higher score is better. */
int hr_merge_topk(const float* S, const int64_t* I, int64_t nq, int n_cand, int k, int largest,
                  float pad_score, float* out_S, int64_t* out_I, int device, void* stream);
/* dense_D/dense_I [nq,kc] as returned by hr_index_search (metric says how to map D to a
 * similarity: IP -> D, L2 -> 1 - D/2, rag/storage/faiss_index.py:86-88); bm25_S/bm25_I
 * [nq,kc]; bm25_max float32[nq] or NULL (= best of each list).  out [nq, top_k]. */
int hr_fuse(const float* dense_D, const int64_t* dense_I, const float* bm25_S,
            const int64_t* bm25_I, const float* bm25_max, int64_t nq, int kc, int top_k,
            int metric, int mode, float w_vec, float w_bm25, float* out_S, int64_t* out_I,
            int device, void* stream);

/* ---- page-level ranking of a batch of hit lists on the device ---------------------
 * group_chunks_by_page + rank_pages + select_top_pages of rag/query/page_retriever.py:145-236 for every query
 * of a batch: the k hits (S, I [nq,k], ids < 0 = padding) are grouped by page_of_row[id - id_base] in order of
 * first appearance, a page scores mean(hit score) + min(0.05 n, 0.15) in double precision with the reference's
 * summation order, pages are sorted by score (stable) and the first top_pages are written:
 * out_page int32[nq,top_pages] (-1 padded), out_score double[nq,top_pages], out_count int32[nq,top_pages].
 * score_kind 0: S is the hit score (fused score, inner product); 1: S is a squared L2 distance and the hit
 * score is the wrapper's clamp(1 - d/2, 0, 1) (rag/storage/faiss_index.py:86-88).  Device pointers, k <= 256. */
int hr_rank_pages(const float* S, const int64_t* I, int64_t nq, int k, int score_kind,
                  const int32_t* page_of_row, int64_t n_rows, int64_t id_base, int top_pages,
                  int32_t* out_page, double* out_score, int32_t* out_count, int device, void* stream);

/* ---- row sharding: a rank's local candidates, then merge + fusion of all ranks' ------
 * hr_candidates: dense top-kc (D,I) and BM25 top-kc (S,J) of THIS shard with global ids (id_base),
 * all device pointers [nq,kc]; bm may be NULL (S=0, J=-1).  Returns after the results are complete.
 * hr_merge_fuse_lists: n_lists candidate sets laid out list_stride_bytes apart (the per-rank blocks of
 * one all-gather, read in place; list order = rank order = id order), merged per modality under
 * (score best first, id asc) and fused to [nq, top_k].  Asynchronous on `stream`. */
int hr_candidates(hr_index* ix, hr_bm25* bm, const float* q, const int32_t* q_indptr,
                  const int32_t* q_terms, int64_t nq, int64_t n_terms, int kc, float* D, int64_t* I,
                  float* S, int64_t* J, void* stream);
int hr_merge_fuse_lists(hr_index* ix, const float* D, const int64_t* I, const float* S, const int64_t* J,
                        int n_lists, int64_t list_stride_bytes, int64_t nq, int kc, int top_k, int mode,
                        float w_vec, float w_bm25, float* out_S, int64_t* out_I, void* stream);

/* ---- the whole query hot path on one device: dense + BM25 + fusion --------------------
 * retrieve(query_embeddings, query_tokens, top_k) of rag/query/retriever.py (path
 * advertised at README.md:90).  bm may be NULL (dense only).  Host or device io. */
int hr_retrieve(hr_index* ix, hr_bm25* bm, const float* q, const int32_t* q_indptr,
                const int32_t* q_terms, int64_t nq, int64_t n_terms, int top_k, int kc, int mode, float w_vec,
                float w_bm25, float* out_S, int64_t* out_I, int io_on_device, void* stream);

/* ---- row-sharded corpus, one process per GPU (the reference is single-process) -----------
 * The corpus is split by row: rank r holds rows [lo_r, hi_r) in `ix` / `bm` (id_base = lo_r, BM25 built with the
 * corpus-wide n_docs_global / avgdl_global / df_global).  hr_retrieve_sharded = hr_retrieve on every rank's
 * shard + ONE ncclAllGather of the (world x kc) candidates per query and modality + merge under (score best
 * first, id asc) + fusion, all enqueued on `stream` by one call; every rank returns the same answer, equal to
 * hr_retrieve on the unsplit corpus.  NCCL is bound at run time (dlopen "libnccl.so.2", or $HR_NCCL_LIB); a
 * world of 1 needs no NCCL.  Rank 0 calls hr_comm_unique_id and ships the 128 bytes to the other ranks over
 * any channel (torch.distributed broadcast, a file, a socket); then every rank calls hr_comm_init.
 * Collective: every rank must call hr_retrieve_sharded with the same nq, kc, top_k and mode. */
int hr_comm_unique_id(void* out_id_128_bytes);
int hr_comm_init(const void* unique_id_128_bytes, int rank, int world, int device, hr_comm** out);
int hr_comm_destroy(hr_comm* c);
int hr_comm_rank(const hr_comm* c);
int hr_comm_world(const hr_comm* c);
int hr_retrieve_sharded(hr_comm* c, hr_index* ix, hr_bm25* bm, const float* q, const int32_t* q_indptr,
                        const int32_t* q_terms, int64_t nq, int64_t n_terms, int top_k, int kc, int mode,
                        float w_vec, float w_bm25, float* out_S, int64_t* out_I, int io_on_device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HR_B200_H */
