// hr_api.cu — C ABI of libhr_b200.so (declared in include/hr_b200.h).
// Host-side orchestration only: device memory, tensor maps, launches, error codes.
// There is NO CPU compute path in this file: without a CUDA device every entry point fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hr_b200.h"
#include "common.cuh"
#include "dense_exact.cuh"
#include "dense_scan_tc.cuh"
#include "bm25.cuh"
#include "bm25_sweep.cuh"
#include "bm25_build.cuh"
#include "fuse.cuh"

using namespace hr;

// -------------------------------------------------------------------------------------------------
// errors, launch accounting
// -------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int set_err(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define HR_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      char _b[512];                                                                                \
      snprintf(_b, sizeof _b, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,        \
               __LINE__, cudaGetErrorString(_e));                                                  \
      return set_err(_e == cudaErrorMemoryAllocation ? HR_ERR_NOMEM : HR_ERR_CUDA, _b);            \
    }                                                                                              \
  } while (0)
#define HR_TRY(expr)             \
  do {                           \
    int _r = (expr);             \
    if (_r != HR_OK) return _r;  \
  } while (0)
#define HR_LAUNCHED()                 \
  do {                                \
    g_launches.fetch_add(1);          \
    HR_CUDA(cudaGetLastError());      \
  } while (0)

// Tuning knobs, read once from the environment (defaults are the measured best on B200).
struct Tuning {
  int q_rows = 128;     // HR_QROWS: query rows materialised for the TMA box of small batches
  int pre_tiles = 4;    // HR_PRE_TILES: corpus tiles per CTA pair sampled by the threshold pre-pass (batches > 128)
  int pre_tiles_small = 1;  // HR_PRE_TILES_SMALL: the same for batches of up to 128 queries (HBM-bound scan: a
                            // sparser sample costs less than the few extra list insertions it causes)
  int pre_rank = 2;     // HR_PRE_RANK: target corpus rank of the seeded threshold, in units of KL
  bool no_pair = false; // HR_NO_PAIR: never use the cta_group::2 scan kernel
  int bm25_wide = 0;    // HR_BM25_WIDE: 1 = one CTA per SM with twice the slice (6144 docs) instead of two CTAs
  int bm25_spans = 0;   // HR_BM25_SPANS: doc windows per query (0 = automatic)
  int bm25_batch = 3;   // HR_BM25_BATCH: 3 = deeper load batch (3 iterations in flight; 4 on the wide variant), 2 = one less
};
static Tuning& tuning_mut() {
  static Tuning t = [] {
    Tuning x;
    auto geti = [](const char* name, int dflt, int lo, int hi) {
      const char* v = getenv(name);
      if (!v || !*v) return dflt;
      const int i = atoi(v);
      return i < lo ? lo : (i > hi ? hi : i);
    };
    x.q_rows = geti("HR_QROWS", x.q_rows, 1, 128);
    x.pre_tiles = geti("HR_PRE_TILES", x.pre_tiles, 1, 64);
    x.pre_tiles_small = geti("HR_PRE_TILES_SMALL", x.pre_tiles_small, 1, 64);
    x.pre_rank = geti("HR_PRE_RANK", x.pre_rank, 1, 16);
    x.no_pair = getenv("HR_NO_PAIR") != nullptr;
    x.bm25_wide = geti("HR_BM25_WIDE", 0, 0, 1);
    x.bm25_spans = geti("HR_BM25_SPANS", 0, 0, 65535);
    x.bm25_batch = geti("HR_BM25_BATCH", 3, 2, 3);
    return x;
  }();
  return t;
}
static const Tuning& tuning() { return tuning_mut(); }

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define HR_DEVICE(dev)                                                        \
  DeviceGuard _guard(dev);                                                    \
  if (!_guard.ok) return set_err(HR_ERR_CUDA, "no usable CUDA device (hr_b200 has no CPU fallback)")

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t n) {
    if (n <= bytes) return HR_OK;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return set_err(HR_ERR_NOMEM, "cudaMalloc of scratch failed");
    }
    bytes = want;
    return HR_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const { return (T*)p; }
};

// -------------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point fetched at run time: no link-time libcuda dependency)
// -------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}
// rows x (kblocks * 128 bytes) K-major matrix, box = {128 bytes, box_rows}, 128-byte swizzle
static int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int ld_elems, int elem_bytes, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return set_err(HR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)ld_elems, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t gstride[1] = {(cuuint64_t)ld_elems * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(m, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[128];
    snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return set_err(HR_ERR_CUDA, b);
  }
  return HR_OK;
}

// -------------------------------------------------------------------------------------------------
// flat dense index
// -------------------------------------------------------------------------------------------------
struct RetrieveScratch {
  DevBuf q, qi, qt, dD, dI, bS, bI, oS, oI;
  // host-io calls upload the query embeddings on this stream while the BM25 kernels (which only need the token
  // ids) already run on the caller's stream
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_copied = nullptr, ev_free = nullptr;
};

struct hr_index {
  RetrieveScratch rs;
  int d = 0, ld = 0, metric = 0, storage = 0, device = 0, mode = HR_MODE_AUTO;
  int elem = 4;
  int num_sms = 148;
  int64_t ntotal = 0, capacity = 0, id_base = 0;
  void* x = nullptr;           // exact rows: fp32 (F32, F32_SHADOW16) or bf16 (BF16), [capacity][ld]
  __nv_bfloat16* xs = nullptr;  // F32_SHADOW16 only: bf16 copy of the rows for the tensor-core filter, [capacity][ld]
  float* norms = nullptr;
  unsigned int* max_norm2 = nullptr;  // ordered-uints: [0] max |x|^2, [1] max |x - filter's view of x|^2
  DevBuf qpad, qh, lists, cnts, tau_g, short_rows, short_s, short_n, short_tot, tprime, tprime_tot, flagged, deeper,
      counters, pre_max;
  DevBuf ex_lists, ex_cnts, ex_tau, ex_sel, io_q, io_D, io_I, stage;
  int* h_counters = nullptr;  // pinned: [0]=nflag [1]=overflow [2]=ndeeper [3]=max |x|^2 (ordered uint)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  hr_scan_stats stats{};
  // a chunk whose device-side counters have not been read back yet (chunk_enqueue / chunk_finish)
  bool pend = false;
  int pend_cap = 0, pend_k = 0;
  float* pend_D = nullptr;
  int64_t* pend_I = nullptr;
  long long launches0 = 0;
  std::mutex mu;   // one search / add at a time per handle: the scratch above is per handle
};

static size_t row_bytes(const hr_index* h) { return (size_t)h->ld * h->elem; }
// the rows the tensor-core filter streams: the exact rows, or their bf16 shadow
static bool filter_is_bf16(const hr_index* h) { return h->storage != HR_STORAGE_F32; }
static const void* filter_rows(const hr_index* h) { return h->storage == HR_STORAGE_F32_SHADOW16 ? (const void*)h->xs : (const void*)h->x; }
static int filter_elem(const hr_index* h) { return filter_is_bf16(h) ? 2 : 4; }

extern "C" const char* hr_last_error(void) { return g_err.c_str(); }
extern "C" int hr_version(void) { return 200; }
#ifndef HR_SOURCE_HASH
#define HR_SOURCE_HASH "unknown"
#endif
// tagged so that the hash can also be read from the file without loading it (__graft_entry__.build)
static const char kSourceHashTag[] = "HR_SOURCE_HASH=" HR_SOURCE_HASH;
extern "C" const char* hr_source_hash(void) { return kSourceHashTag + 15; }
extern "C" int64_t hr_launch_count(void) { return (int64_t)g_launches.load(); }
// Tuning / diagnostic knobs by name (the HR_* environment variables without the prefix, lower case).  Not
// thread-safe against running searches; meant for benchmarks and tests.
extern "C" int hr_set_option(const char* name, int value) {
  if (!name) return set_err(HR_ERR_INVALID, "null option name");
  Tuning& t = tuning_mut();
  const std::string n(name);
  auto clamp = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
  if (n == "qrows") t.q_rows = clamp(value, 1, 128);
  else if (n == "pre_tiles") t.pre_tiles = clamp(value, 1, 64);
  else if (n == "pre_tiles_small") t.pre_tiles_small = clamp(value, 1, 64);
  else if (n == "pre_rank") t.pre_rank = clamp(value, 1, 16);
  else if (n == "no_pair") t.no_pair = value != 0;
  else if (n == "bm25_wide") t.bm25_wide = clamp(value, 0, 1);
  else if (n == "bm25_spans") t.bm25_spans = clamp(value, 0, 65535);
  else if (n == "bm25_batch") t.bm25_batch = clamp(value, 2, 3);
  else return set_err(HR_ERR_INVALID, "unknown option: " + n);
  return HR_OK;
}
extern "C" int hr_device_count(int* out) {
  if (!out) return set_err(HR_ERR_INVALID, "null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    *out = 0;
    return set_err(HR_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *out = n;
  return HR_OK;
}

extern "C" int hr_index_create(int d, int metric, int storage_dtype, int device, hr_index** out) {
  if (!out) return set_err(HR_ERR_INVALID, "null out");
  *out = nullptr;
  // the exact re-score / fallback kernels keep kExactF padded query rows in shared memory (200 KB): d <= 6400
  if (d <= 0 || d > 6400) return set_err(HR_ERR_INVALID, "d must be in [1, 6400]");
  if (metric != HR_METRIC_INNER_PRODUCT && metric != HR_METRIC_L2)
    return set_err(HR_ERR_INVALID, "metric must be METRIC_INNER_PRODUCT (0) or METRIC_L2 (1)");
  if (storage_dtype != HR_STORAGE_F32 && storage_dtype != HR_STORAGE_BF16 && storage_dtype != HR_STORAGE_F32_SHADOW16)
    return set_err(HR_ERR_INVALID, "storage dtype must be HR_STORAGE_F32, HR_STORAGE_BF16 or HR_STORAGE_F32_SHADOW16");
  int ndev = 0;
  HR_TRY(hr_device_count(&ndev));
  if (ndev <= 0) return set_err(HR_ERR_CUDA, "no CUDA device (hr_b200 has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_err(HR_ERR_INVALID, "device ordinal out of range");
  HR_DEVICE(device);
  cudaDeviceProp prop;
  HR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    char b[160];
    snprintf(b, sizeof b, "device %d is sm_%d%d; libhr_b200 is built for sm_100a (B200) only", device, prop.major,
             prop.minor);
    return set_err(HR_ERR_CUDA, b);
  }
  hr_index* h = new hr_index();
  h->d = d;
  h->metric = metric;
  h->storage = storage_dtype;
  h->device = device;
  h->elem = storage_dtype == HR_STORAGE_BF16 ? 2 : 4;
  // elements per 128-byte K block of the filter's rows (the shadow shares ld with the fp32 rows)
  const int kelems = storage_dtype == HR_STORAGE_F32 ? 32 : 64;
  h->ld = ((d + kelems - 1) / kelems) * kelems;
  h->num_sms = prop.multiProcessorCount;
  if (cudaMalloc((void**)&h->max_norm2, 2 * sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(h->max_norm2, 0, 2 * sizeof(unsigned int)) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_counters, 4 * sizeof(int)) != cudaSuccess) {
    (void)cudaGetLastError();
    delete h;
    return set_err(HR_ERR_NOMEM, "allocation failed in hr_index_create");
  }
  for (int i = 0; i < 4; ++i)
    if (cudaEventCreate(&h->ev[i]) != cudaSuccess) {
      (void)cudaGetLastError();
      hr_index_destroy(h);
      return set_err(HR_ERR_CUDA, "cudaEventCreate failed in hr_index_create");
    }
  *out = h;
  return HR_OK;
}

extern "C" int hr_index_destroy(hr_index* h) {
  if (!h) return HR_OK;
  DeviceGuard g(h->device);
  if (h->x) cudaFree(h->x);
  if (h->xs) cudaFree(h->xs);
  if (h->norms) cudaFree(h->norms);
  if (h->max_norm2) cudaFree(h->max_norm2);
  if (h->h_counters) cudaFreeHost(h->h_counters);
  DevBuf* bufs[] = {&h->qpad, &h->qh, &h->lists, &h->cnts, &h->tau_g, &h->short_rows, &h->short_n, &h->tprime,
                    &h->flagged, &h->counters, &h->pre_max, &h->short_tot, &h->tprime_tot, &h->deeper, &h->short_s, &h->ex_lists, &h->ex_cnts, &h->ex_tau, &h->ex_sel, &h->io_q,
                    &h->io_D, &h->io_I, &h->stage, &h->rs.q, &h->rs.qi, &h->rs.qt, &h->rs.dD, &h->rs.dI,
                    &h->rs.bS, &h->rs.bI, &h->rs.oS, &h->rs.oI};
  for (DevBuf* b : bufs) b->release();
  for (int i = 0; i < 4; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->rs.copy) cudaStreamDestroy(h->rs.copy);
  if (h->rs.ev_copied) cudaEventDestroy(h->rs.ev_copied);
  if (h->rs.ev_free) cudaEventDestroy(h->rs.ev_free);
  delete h;
  return HR_OK;
}

static int index_grow(hr_index* h, int64_t need, cudaStream_t st) {
  if (need <= h->capacity) return HR_OK;
  void* nx = nullptr;
  void* ns = nullptr;
  float* nn = nullptr;
  const bool shadow = h->storage == HR_STORAGE_F32_SHADOW16;
  cudaError_t e = cudaMalloc(&nx, (size_t)need * row_bytes(h));
  if (e == cudaSuccess) e = cudaMalloc((void**)&nn, (size_t)need * sizeof(float));
  if (e == cudaSuccess && shadow) e = cudaMalloc(&ns, (size_t)need * h->ld * 2);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    if (nx) cudaFree(nx);
    if (nn) cudaFree(nn);
    if (ns) cudaFree(ns);
    return set_err(HR_ERR_NOMEM, "cudaMalloc failed while growing the index (corpus does not fit in HBM?)");
  }
  if (h->ntotal > 0) {
    cudaError_t c = cudaMemcpyAsync(nx, h->x, (size_t)h->ntotal * row_bytes(h), cudaMemcpyDeviceToDevice, st);
    if (c == cudaSuccess) c = cudaMemcpyAsync(nn, h->norms, (size_t)h->ntotal * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (c == cudaSuccess && shadow) c = cudaMemcpyAsync(ns, h->xs, (size_t)h->ntotal * h->ld * 2, cudaMemcpyDeviceToDevice, st);
    if (c == cudaSuccess) c = cudaStreamSynchronize(st);
    if (c != cudaSuccess) {
      (void)cudaGetLastError();
      cudaFree(nx);
      cudaFree(nn);
      if (ns) cudaFree(ns);
      return set_err(HR_ERR_CUDA, std::string("copy failed while growing the index: ") + cudaGetErrorString(c));
    }
  }
  if (h->x) cudaFree(h->x);
  if (h->xs) cudaFree(h->xs);
  if (h->norms) cudaFree(h->norms);
  h->x = nx;
  h->xs = (__nv_bfloat16*)ns;
  h->norms = nn;
  h->capacity = need;
  return HR_OK;
}

extern "C" int hr_index_reserve(hr_index* h, int64_t n_rows) {
  if (!h) return set_err(HR_ERR_INVALID, "null index");
  if (n_rows < 0 || n_rows >= (int64_t)0xFFFFFFF0ll) return set_err(HR_ERR_INVALID, "n_rows out of range");
  HR_DEVICE(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  return index_grow(h, n_rows, 0);
}

extern "C" int hr_index_add(hr_index* h, const float* x, int64_t n, int is_device, void* stream) {
  if (!h) return set_err(HR_ERR_INVALID, "null index");
  if (n < 0) return set_err(HR_ERR_INVALID, "n < 0");
  if (n == 0) return HR_OK;
  if (!x) return set_err(HR_ERR_INVALID, "null x");
  if (h->ntotal + n >= (int64_t)0xFFFFFFF0ll) return set_err(HR_ERR_INVALID, "too many rows for one shard");
  HR_DEVICE(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->ntotal + n > h->capacity) {
    int64_t want = h->ntotal + n;
    if (h->capacity > 0) want = std::max<int64_t>(want, h->capacity + h->capacity / 2);
    HR_TRY(index_grow(h, want, st));
  }
  const int64_t chunk_rows = std::max<int64_t>(1, ((int64_t)64 << 20) / ((int64_t)h->d * 4));
  for (int64_t r0 = 0; r0 < n; r0 += chunk_rows) {
    const int64_t nr = std::min(chunk_rows, n - r0);
    const float* src = x + r0 * (int64_t)h->d;
    if (!is_device) {
      HR_TRY(h->stage.ensure((size_t)nr * h->d * 4));
      HR_CUDA(cudaMemcpyAsync(h->stage.p, src, (size_t)nr * h->d * 4, cudaMemcpyHostToDevice, st));
      src = h->stage.as<float>();
    }
    const int64_t row0 = h->ntotal + r0;
    const int blocks = (int)std::min<int64_t>((nr + 7) / 8, (int64_t)h->num_sms * 8);
    if (h->storage != HR_STORAGE_BF16) {
      convert_pad_norm_kernel<float><<<blocks, 256, 0, st>>>(src, nr, h->d, (float*)h->x + row0 * h->ld, h->ld,
                                                            h->norms + row0, h->max_norm2,
                                                            h->storage == HR_STORAGE_F32 ? 1 : 2);
      if (h->storage == HR_STORAGE_F32_SHADOW16) {
        HR_LAUNCHED();
        const int64_t tot = nr * (int64_t)h->ld;
        shadow_bf16_kernel<<<(int)std::min<int64_t>((tot / 4 + 255) / 256, (int64_t)h->num_sms * 16), 256, 0, st>>>(
            (const float*)h->x + row0 * h->ld, tot, h->xs + row0 * h->ld);
      }
    } else
      convert_pad_norm_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
          src, nr, h->d, (__nv_bfloat16*)h->x + row0 * h->ld, h->ld, h->norms + row0, h->max_norm2, 0);
    HR_LAUNCHED();
    if (!is_device) HR_CUDA(cudaStreamSynchronize(st));  // staging buffer is reused
  }
  HR_CUDA(cudaStreamSynchronize(st));
  h->ntotal += n;
  return HR_OK;
}

extern "C" int hr_index_reset(hr_index* h) {
  if (!h) return set_err(HR_ERR_INVALID, "null index");
  HR_DEVICE(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  h->ntotal = 0;
  HR_CUDA(cudaMemset(h->max_norm2, 0, 2 * sizeof(unsigned int)));
  return HR_OK;
}
extern "C" int64_t hr_index_ntotal(const hr_index* h) { return h ? h->ntotal : -1; }
extern "C" int hr_index_d(const hr_index* h) { return h ? h->d : -1; }
extern "C" int hr_index_metric(const hr_index* h) { return h ? h->metric : -1; }
extern "C" int hr_index_storage(const hr_index* h) { return h ? h->storage : -1; }
extern "C" int hr_index_set_id_base(hr_index* h, int64_t id_base) {
  if (!h) return set_err(HR_ERR_INVALID, "null index");
  h->id_base = id_base;
  return HR_OK;
}
extern "C" int hr_index_set_mode(hr_index* h, int mode) {
  if (!h) return set_err(HR_ERR_INVALID, "null index");
  if (mode != HR_MODE_AUTO && mode != HR_MODE_EXACT_SIMT) return set_err(HR_ERR_INVALID, "unknown mode");
  h->mode = mode;
  return HR_OK;
}
extern "C" int hr_index_last_stats(const hr_index* h, hr_scan_stats* out) {
  if (!h || !out) return set_err(HR_ERR_INVALID, "null argument");
  *out = h->stats;
  return HR_OK;
}

extern "C" int hr_index_debug_dump(hr_index* h, int64_t nq, void* lists, int32_t* cnts, uint32_t* tau,
                                   uint32_t* short_rows, float* tprime) {
  if (!h) return set_err(HR_ERR_INVALID, "null index");
  HR_DEVICE(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  const int64_t G = h->stats.grid, KL = h->stats.list_len;
  if (h->stats.mode_used != HR_MODE_AUTO || G <= 0 || nq <= 0 || nq > kScanNqMax)
    return set_err(HR_ERR_INVALID, "debug_dump: last search did not run the filter scan");
  HR_CUDA(cudaDeviceSynchronize());
  if (lists) HR_CUDA(cudaMemcpy(lists, h->lists.p, (size_t)G * nq * KL * sizeof(Cand), cudaMemcpyDeviceToHost));
  if (cnts) HR_CUDA(cudaMemcpy(cnts, h->cnts.p, (size_t)G * nq * 4, cudaMemcpyDeviceToHost));
  if (tau) HR_CUDA(cudaMemcpy(tau, h->tau_g.p, (size_t)nq * 4, cudaMemcpyDeviceToHost));
  if (short_rows)
    HR_CUDA(cudaMemcpy2D(short_rows, (size_t)KL * 4, h->short_rows.p, (size_t)kShortCap * 4, (size_t)KL * 4, (size_t)nq,
                         cudaMemcpyDeviceToHost));
  if (tprime) HR_CUDA(cudaMemcpy(tprime, h->tprime.p, (size_t)nq * 4, cudaMemcpyDeviceToHost));
  return HR_OK;
}

extern "C" int hr_index_reconstruct(hr_index* h, int64_t i0, int64_t n, float* out_host) {
  if (!h || !out_host) return set_err(HR_ERR_INVALID, "null argument");
  if (i0 < 0 || n < 0 || i0 + n > h->ntotal) return set_err(HR_ERR_INVALID, "reconstruct: row range out of bounds");
  if (n == 0) return HR_OK;
  HR_DEVICE(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  const int64_t chunk = std::max<int64_t>(1, ((int64_t)64 << 20) / ((int64_t)h->d * 4));
  for (int64_t r0 = 0; r0 < n; r0 += chunk) {
    const int64_t nr = std::min(chunk, n - r0);
    HR_TRY(h->stage.ensure((size_t)nr * h->d * 4));
    const int blocks = (int)std::min<int64_t>((nr * h->d + 255) / 256, 4096);
    if (h->elem == 4)
      reconstruct_kernel<float><<<blocks, 256>>>((const float*)h->x, h->ld, h->d, i0 + r0, nr, h->stage.as<float>());
    else
      reconstruct_kernel<__nv_bfloat16><<<blocks, 256>>>((const __nv_bfloat16*)h->x, h->ld, h->d, i0 + r0, nr,
                                                         h->stage.as<float>());
    HR_LAUNCHED();
    HR_CUDA(cudaMemcpy(out_host + r0 * (int64_t)h->d, h->stage.p, (size_t)nr * h->d * 4, cudaMemcpyDeviceToHost));
  }
  return HR_OK;
}

// ---- search -------------------------------------------------------------------------------------
__global__ void fill_pad_kernel(float* D, int64_t* I, int64_t n, float pad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    D[i] = pad;
    I[i] = -1;
  }
}
__global__ void iota_kernel(int* a, int n, int base) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = base + i;
}

constexpr int kSampleStrideMax = 1024;  // the threshold pre-pass visits at most every 2nd, at least every 1024th corpus tile

static int list_len_for_k(int k) {
  if (k <= 16) return 32;
  if (k <= 32) return 64;
  if (k <= 64) return 128;
  return 256;
}

// queries the device-driven fallback (enqueued without knowing the count) handles per chunk; more than that
// (pathological ties) are finished on the host-driven path after the caller's synchronisation
constexpr int kExactDevCap = 64;

static int exact_group(int k) { return std::max(kExactF, std::min(kExactDevCap, (64 * 128 / std::max(k, 1)) / kExactF * kExactF)); }

// Exhaustive exact scan of the queries listed in qsel_dev (device), results written to D/I rows qsel[i].
// Host-driven (nsel_dev == nullptr): nsel queries, in groups.  Device-driven: the count is read on the device
// from *nsel_dev (at most `nsel` = capacity are handled); both kernel variants are launched and exit at once
// unless the count is in their range, so nothing here waits for the host.
template <typename T>
static int launch_exact(hr_index* h, const int* qsel_dev, int nsel, int k, float* D, int64_t* I, cudaStream_t st,
                        const int* nsel_dev = nullptr) {
  // warps in flight set the memory-level parallelism of this row-per-warp scan: as many CTAs per SM as the
  // query tile in shared memory allows, up to 2 (measured at 10M x 1024: 39 / 28 / 35 ms per 8 queries with 1 / 2 / 4);
  // k is up to 2048, so the per-warp lists bound it as well
  const size_t smem_q = (size_t)kExactF * h->ld * 4;
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (200 * 1024) / std::max<size_t>(smem_q, 1)));
  if (k > 256) per_sm = 1;
  const int grid = h->num_sms * per_sm;
  const int64_t W = (int64_t)grid * 8;
  const int fsel = exact_group(k);
  HR_TRY(h->ex_lists.ensure((size_t)fsel * W * k * 8));
  HR_TRY(h->ex_cnts.ensure((size_t)fsel * W * 4));
  HR_TRY(h->ex_tau.ensure((size_t)fsel * 8));
  if (smem_q > 200 * 1024) return set_err(HR_ERR_INVALID, "d too large for the exact scan kernel");
  auto run_scan = [&](auto scan, int F, const int* qs, int ns, int lo, int hi) -> int {
    const size_t smem = (size_t)F * h->ld * 4;
    HR_CUDA(cudaFuncSetAttribute(scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scan<<<grid, 256, smem, st>>>((const T*)h->x, h->ntotal, h->ld, h->qpad.as<float>(), qs, ns, k,
                                  h->ex_lists.as<uint64_t>(), h->ex_cnts.as<int>(), h->ex_tau.as<unsigned long long>(),
                                  nsel_dev, lo, hi);
    HR_LAUNCHED();
    return HR_OK;
  };
  auto run_merge = [&](auto merge, const int* qs, int ns) -> int {
    merge<<<ns, 256, 0, st>>>(h->ex_lists.as<uint64_t>(), h->ex_cnts.as<int>(), h->ex_tau.as<unsigned long long>(), qs, W,
                              k, h->id_base, D, I, nsel_dev);
    HR_LAUNCHED();
    return HR_OK;
  };
  const bool ip = h->metric == HR_METRIC_INNER_PRODUCT;
  if (nsel_dev) {
    const int cap = std::min(nsel, fsel);
    HR_CUDA(cudaMemsetAsync(h->ex_tau.p, 0, (size_t)cap * 8, st));
    if (ip) {
      HR_TRY(run_scan(exact_scan_kernel<T, kMetricIP, 2>, 2, qsel_dev, cap, 1, 2));
      if (cap > 2) HR_TRY(run_scan(exact_scan_kernel<T, kMetricIP, kExactF>, kExactF, qsel_dev, cap, 3, 1 << 30));
      HR_TRY(run_merge(exact_merge_kernel<kMetricIP>, qsel_dev, cap));
    } else {
      HR_TRY(run_scan(exact_scan_kernel<T, kMetricL2, 2>, 2, qsel_dev, cap, 1, 2));
      if (cap > 2) HR_TRY(run_scan(exact_scan_kernel<T, kMetricL2, kExactF>, kExactF, qsel_dev, cap, 3, 1 << 30));
      HR_TRY(run_merge(exact_merge_kernel<kMetricL2>, qsel_dev, cap));
    }
    return HR_OK;
  }
  for (int g0 = 0; g0 < nsel; g0 += fsel) {
    const int ns = std::min(fsel, nsel - g0);
    HR_CUDA(cudaMemsetAsync(h->ex_tau.p, 0, (size_t)ns * 8, st));
    const bool few = ns <= 2;
    if (ip) {
      if (few) HR_TRY(run_scan(exact_scan_kernel<T, kMetricIP, 2>, 2, qsel_dev + g0, ns, 0, 0));
      else HR_TRY(run_scan(exact_scan_kernel<T, kMetricIP, kExactF>, kExactF, qsel_dev + g0, ns, 0, 0));
      HR_TRY(run_merge(exact_merge_kernel<kMetricIP>, qsel_dev + g0, ns));
    } else {
      if (few) HR_TRY(run_scan(exact_scan_kernel<T, kMetricL2, 2>, 2, qsel_dev + g0, ns, 0, 0));
      else HR_TRY(run_scan(exact_scan_kernel<T, kMetricL2, kExactF>, kExactF, qsel_dev + g0, ns, 0, 0));
      HR_TRY(run_merge(exact_merge_kernel<kMetricL2>, qsel_dev + g0, ns));
    }
  }
  return HR_OK;
}

template <int KIND, int METRIC>
static int launch_scan(hr_index* h, const CUtensorMap& tq, const CUtensorMap& tx, const ScanParams& p, int grid,
                       cudaStream_t st) {
  HR_CUDA(cudaFuncSetAttribute(scan_tc_kernel<KIND, METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kScanSmemBytes));
  scan_tc_kernel<KIND, METRIC><<<grid, kScanThreads, kScanSmemBytes, st>>>(tq, tx, p);
  HR_LAUNCHED();
  return HR_OK;
}

template <int KIND, int METRIC>
static int launch_scan2(hr_index* h, const CUtensorMap& tq, const CUtensorMap& tx, const ScanParams& p, int grid,
                        cudaStream_t st) {
  HR_CUDA(cudaFuncSetAttribute(scan_tc2_kernel<KIND, METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kScan2SmemBytes));
  scan_tc2_kernel<KIND, METRIC><<<grid, kScanThreads, kScan2SmemBytes, st>>>(tq, tx, p);
  HR_LAUNCHED();
  return HR_OK;
}

// stage 1: every query of the batch, the best KL candidates; stage 2 (launched for every query, runs for the
// `*ndeeper` queries stage 1 listed in h->deeper): all their candidates (up to kShortCap) against the bound of the
// thresholds alone
template <typename T>
static int launch_rescore(hr_index* h, int nb, int KL, int k, float c_acc, float* D, int64_t* I, cudaStream_t st,
                          bool stage2) {
  const int fk = filter_is_bf16(h) ? 2 : 1;   // how the filter rounded the query
  const int depth = stage2 ? kShortCap : KL;
  const size_t smem = (size_t)depth * 8;
  const int* n_in = stage2 ? h->short_tot.as<int>() : h->short_n.as<int>();
  const float* tp = stage2 ? h->tprime_tot.as<float>() : h->tprime.as<float>();
  const int* qsel = stage2 ? h->deeper.as<int>() : nullptr;
  int* deeper = stage2 ? nullptr : h->deeper.as<int>();
  const int* nsel_dev = stage2 ? h->counters.as<int>() + 2 : nullptr;
  const int threads = nb <= 32 ? 1024 : 256;
  if (h->metric == HR_METRIC_INNER_PRODUCT)
    rescore_finalize_kernel<T, kMetricIP><<<nb, threads, smem, st>>>(
        (const T*)h->x, h->ld, h->qpad.as<float>(), h->short_rows.as<uint32_t>(), kShortCap, n_in, tp, depth, k, c_acc,
        fk, h->max_norm2, h->id_base, D, I, h->flagged.as<int>(), h->counters.as<int>(), qsel,
        h->short_tot.as<int>(), deeper, h->counters.as<int>() + 2, h->short_s.as<float>(), nsel_dev);
  else
    rescore_finalize_kernel<T, kMetricL2><<<nb, threads, smem, st>>>(
        (const T*)h->x, h->ld, h->qpad.as<float>(), h->short_rows.as<uint32_t>(), kShortCap, n_in, tp, depth, k, c_acc,
        fk, h->max_norm2, h->id_base, D, I, h->flagged.as<int>(), h->counters.as<int>(), qsel,
        h->short_tot.as<int>(), deeper, h->counters.as<int>() + 2, h->short_s.as<float>(), nsel_dev);
  HR_LAUNCHED();
  return HR_OK;
}

// ---- one chunk (<= kScanNqMax queries) of a dense search: everything is enqueued on `st`, nothing waits for the
//      host.  The certificate's decisions are taken on the device: the second re-score stage and the exact
//      fallback are always launched and run only for the queries the previous kernel listed.  chunk_finish()
//      (after the caller synchronised `st`) reads the counters back, and completes the rare case of more than
//      kExactDevCap fallback queries. ----
static int chunk_enqueue(hr_index* h, const float* q_dev, int nb, int k, float* Db, int64_t* Ib, cudaStream_t st) {
  h->pend = false;
  const float pad = h->metric == HR_METRIC_INNER_PRODUCT ? HR_NEG_INF : -HR_NEG_INF;
  if (h->ntotal == 0) {
    fill_pad_kernel<<<(int)std::min<int64_t>(((int64_t)nb * k + 255) / 256, 1024), 256, 0, st>>>(Db, Ib, (int64_t)nb * k, pad);
    HR_LAUNCHED();
    return HR_OK;
  }
  const bool use_tc = (h->mode == HR_MODE_AUTO) && k <= 128;
  if (h->mode == HR_MODE_AUTO && k > 128) {
    static std::atomic<bool> warned{false};
    if (!warned.exchange(true))
      fprintf(stderr, "hr_b200: k = %d > 128: the tensor-core filter keeps at most 256 candidates per query, this search "
                      "runs the exhaustive exact CUDA-core scan (same answers, ~10x slower per query)\n", k);
  }
  // the bf16 shadow doubles the filter's error bound: a deeper shortlist keeps the certificate's margin
  const int KL = h->storage == HR_STORAGE_F32_SHADOW16 ? std::min(256, 2 * list_len_for_k(k)) : list_len_for_k(k);
  const int num_ctiles = (int)((h->ntotal + kScanBN - 1) / kScanBN);
  const int grid = std::min(num_ctiles, h->num_sms);
  h->stats.mode_used = use_tc ? HR_MODE_AUTO : HR_MODE_EXACT_SIMT;
  h->stats.list_len = use_tc ? KL : k;
  h->stats.grid = use_tc ? grid : h->num_sms;
  // query rows materialised for the TMA (zero rows beyond nb): a full 128-row box, so that the SMs'
  // re-reads of a small batch spread over 128 rows instead of hammering a few L2 lines
  const int nbp = std::max(nb, tuning().q_rows);
  HR_TRY(h->qpad.ensure((size_t)nbp * h->ld * 4));
  if (filter_is_bf16(h)) HR_TRY(h->qh.ensure((size_t)nbp * h->ld * 2));
  {
    const int64_t tot = (int64_t)nbp * h->ld;
    pad_queries_kernel<<<(int)std::min<int64_t>((tot + 255) / 256, 2048), 256, 0, st>>>(
        q_dev, nb, nbp, h->d, h->ld, h->qpad.as<float>(), filter_is_bf16(h) ? h->qh.as<__nv_bfloat16>() : nullptr);
    HR_LAUNCHED();
  }
  HR_TRY(h->ex_sel.ensure((size_t)nb * 4));
  if (!use_tc) {
    iota_kernel<<<(nb + 255) / 256, 256, 0, st>>>(h->ex_sel.as<int>(), nb, 0);
    HR_LAUNCHED();
    if (h->elem == 4) return launch_exact<float>(h, h->ex_sel.as<int>(), nb, k, Db, Ib, st);
    return launch_exact<__nv_bfloat16>(h, h->ex_sel.as<int>(), nb, k, Db, Ib, st);
  }
  // ---- tensor-core filter scan ----
  HR_TRY(h->lists.ensure((size_t)h->num_sms * nb * KL * sizeof(Cand)));
  HR_TRY(h->cnts.ensure((size_t)h->num_sms * nb * 4));
  HR_TRY(h->tau_g.ensure((size_t)nb * 4));
  HR_TRY(h->short_rows.ensure((size_t)nb * kShortCap * 4));
  HR_TRY(h->short_s.ensure((size_t)nb * kShortCap * 4));
  HR_TRY(h->short_tot.ensure((size_t)nb * 4));
  HR_TRY(h->tprime_tot.ensure((size_t)nb * 4));
  HR_TRY(h->deeper.ensure((size_t)nb * 4));
  HR_TRY(h->short_n.ensure((size_t)nb * 4));
  HR_TRY(h->tprime.ensure((size_t)nb * 4));
  HR_TRY(h->flagged.ensure((size_t)nb * 4));
  HR_TRY(h->counters.ensure(16));
  HR_CUDA(cudaMemsetAsync(h->tau_g.p, 0, (size_t)nb * 4, st));
  HR_CUDA(cudaMemsetAsync(h->counters.p, 0, 16, st));
  CUtensorMap tq, tx;
  const void* qsrc = filter_is_bf16(h) ? (const void*)h->qh.p : (const void*)h->qpad.p;
  HR_TRY(make_tmap(&tq, qsrc, nbp, h->ld, filter_elem(h), kScanBM));
  HR_TRY(make_tmap(&tx, filter_rows(h), h->ntotal, h->ld, filter_elem(h), kScanBN));
  ScanParams p;
  p.N = h->ntotal;
  p.nq = nb;
  p.kblocks = (int)((size_t)h->ld * filter_elem(h) / 128);
  p.KL = KL;
  p.num_qtiles = (nb + kScanBM - 1) / kScanBM;
  p.norms = h->norms;
  p.lists = h->lists.as<Cand>();
  p.cnts = h->cnts.as<int>();
  p.tau_g = h->tau_g.as<unsigned int>();
  p.pre_max = nullptr;
  // batches of more than 128 queries run on CTA pairs (cta_group::2, 128-row corpus halves per CTA)
  const bool use_pair = nb > kScanBM && h->num_sms >= 2 && !tuning().no_pair;
  CUtensorMap tx2;
  if (use_pair) HR_TRY(make_tmap(&tx2, filter_rows(h), h->ntotal, h->ld, filter_elem(h), 128));
  auto run_scan = [&](int g) -> int {
    if (use_pair) {
      const int g2 = std::max(2, std::min(2 * p.tile_count, h->num_sms) & ~1);
      if (!filter_is_bf16(h)) {
        if (h->metric == HR_METRIC_INNER_PRODUCT) return launch_scan2<0, 0>(h, tq, tx2, p, g2, st);
        return launch_scan2<0, 1>(h, tq, tx2, p, g2, st);
      }
      if (h->metric == HR_METRIC_INNER_PRODUCT) return launch_scan2<1, 0>(h, tq, tx2, p, g2, st);
      return launch_scan2<1, 1>(h, tq, tx2, p, g2, st);
    }
    if (!filter_is_bf16(h)) {
      if (h->metric == HR_METRIC_INNER_PRODUCT) return launch_scan<0, 0>(h, tq, tx, p, g, st);
      return launch_scan<0, 1>(h, tq, tx, p, g, st);
    }
    if (h->metric == HR_METRIC_INNER_PRODUCT) return launch_scan<1, 0>(h, tq, tx, p, g, st);
    return launch_scan<1, 1>(h, tq, tx, p, g, st);
  };
  // ---- threshold pre-pass over a strided sample of the corpus tiles.  The sample's KLs-th best score
  //      seeds tau_g: about KLs*stride (= 4*KL) corpus rows beat it, so in the main pass only a handful
  //      of scores per CTA pass the threshold and no per-CTA list ever fills.  Correctness never depends
  //      on the seed: the certificate in rescore_finalize compares against the final threshold. ----
  // Sample: at least pre_tiles tiles per scheduling unit (CTA or CTA pair) and at least every 32nd tile, at most
  // kSeedCap tiles.  The seed is the j-th largest tile maximum, j >= 32: its corpus rank is Gamma(j, stride)
  // distributed, mean about pre_rank * KL (or 32 * stride if that is larger), spread 1/sqrt(j).  A seed taken
  // from fewer maxima (j = 8) landed inside the top 100 rows of a 5M-row shard for about 1 query in 10^4 and sent
  // it to the exact scan: 16 ms for one query.
  const int units = use_pair ? std::max(1, h->num_sms / 2) : h->num_sms;
  const int pre_tiles = use_pair ? tuning().pre_tiles : tuning().pre_tiles_small;
  int stride = std::max(1, std::min(kSampleStrideMax, num_ctiles / (pre_tiles * units)));
  stride = std::min(stride, 32);
  stride = std::max(stride, (num_ctiles + kSeedCap - 1) / kSeedCap);
  if (stride > 1) {
    p.tile_stride = stride;
    p.tile_count = (num_ctiles + stride - 1) / stride;
    const int rk = tuning().pre_rank;   // target rank of the seed, in KL
    const int jth = std::min(p.tile_count, std::max(32, (rk * KL + stride - 1) / stride));
    HR_TRY(h->pre_max.ensure((size_t)p.tile_count * nb * 4));
    p.pre_max = h->pre_max.as<float>();
    const int gs = use_pair ? std::max(2, std::min(2 * p.tile_count, h->num_sms) & ~1)
                            : std::min(p.tile_count, h->num_sms);
    HR_TRY(run_scan(gs));
    scan_seed_kernel<<<nb, 256, 0, st>>>(h->pre_max.as<float>(), p.tile_count, nb, jth, h->tau_g.as<unsigned int>());
    HR_LAUNCHED();
    p.pre_max = nullptr;
  }
  p.tile_stride = 1;
  p.tile_count = num_ctiles;
  cudaEventRecord(h->ev[2], st);
  const int gmain = use_pair ? std::max(2, std::min(2 * p.tile_count, h->num_sms) & ~1) : grid;
  HR_TRY(run_scan(gmain));
  cudaEventRecord(h->ev[3], st);
  h->stats.grid = gmain;
  scan_merge_kernel<<<nb, 256, 0, st>>>(h->lists.as<Cand>(), h->cnts.as<int>(), h->tau_g.as<unsigned int>(), gmain,
                                        nb, KL, h->short_rows.as<uint32_t>(), h->short_n.as<int>(),
                                        h->tprime.as<float>(), h->counters.as<int>() + 1, nullptr,
                                        h->short_tot.as<int>(), h->tprime_tot.as<float>(), h->short_s.as<float>());
  HR_LAUNCHED();
  // filter error bound = measured rounding residuals of both operands (rescore_finalize_kernel) + this
  // relative allowance for the fp32 accumulation over ld terms
  const float c_acc = 2.4e-7f * (float)h->ld;
  const int cap = std::min(nb, exact_group(k));
  if (h->elem == 4) {
    HR_TRY(launch_rescore<float>(h, nb, KL, k, c_acc, Db, Ib, st, false));
    HR_TRY(launch_rescore<float>(h, nb, KL, k, c_acc, Db, Ib, st, true));
    HR_TRY(launch_exact<float>(h, h->flagged.as<int>(), cap, k, Db, Ib, st, h->counters.as<int>()));
  } else {
    HR_TRY(launch_rescore<__nv_bfloat16>(h, nb, KL, k, c_acc, Db, Ib, st, false));
    HR_TRY(launch_rescore<__nv_bfloat16>(h, nb, KL, k, c_acc, Db, Ib, st, true));
    HR_TRY(launch_exact<__nv_bfloat16>(h, h->flagged.as<int>(), cap, k, Db, Ib, st, h->counters.as<int>()));
  }
  HR_CUDA(cudaMemcpyAsync(h->h_counters, h->counters.p, 12, cudaMemcpyDeviceToHost, st));
  HR_CUDA(cudaMemcpyAsync(h->h_counters + 3, h->max_norm2, 4, cudaMemcpyDeviceToHost, st));
  h->pend = true;
  h->pend_cap = cap;
  h->pend_k = k;
  h->pend_D = Db;
  h->pend_I = Ib;
  return HR_OK;
}

// After the caller synchronised the stream of chunk_enqueue: statistics, and the fallback queries beyond the
// device-driven capacity (then *changed = true: D/I were modified by more work, which has been synchronised).
static int chunk_finish(hr_index* h, cudaStream_t st, bool* changed) {
  if (changed) *changed = false;
  if (!h->pend) return HR_OK;
  h->pend = false;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) h->stats.scan_ms += ms;
  const int nflag = h->h_counters[0];
  if (nflag > 0 && !std::isfinite(ord2f((uint32_t)h->h_counters[3]))) {
    // a row with a NaN / Inf component (or |x| > 1.8e19): the filter's error bound is unbounded, no query can be
    // certified, every search takes the exhaustive exact scan (same answers; such rows are never returned)
    static std::atomic<bool> warned{false};
    if (!warned.exchange(true))
      fprintf(stderr, "hr_b200: the index holds rows whose squared norm is not finite; every query runs the exhaustive "
                      "exact scan (remove NaN / Inf rows to get the tensor-core path back)\n");
  }
  h->stats.flagged += nflag;
  h->stats.overflow += h->h_counters[1];
  h->stats.deeper += h->h_counters[2];
  if (nflag > h->pend_cap) {
    const int rest = nflag - h->pend_cap;
    if (h->elem == 4)
      HR_TRY(launch_exact<float>(h, h->flagged.as<int>() + h->pend_cap, rest, h->pend_k, h->pend_D, h->pend_I, st));
    else
      HR_TRY(launch_exact<__nv_bfloat16>(h, h->flagged.as<int>() + h->pend_cap, rest, h->pend_k, h->pend_D, h->pend_I, st));
    HR_CUDA(cudaStreamSynchronize(st));
    if (changed) *changed = true;
  }
  return HR_OK;
}

// q_dev: fp32 [nq, d] on device; D_dev/I_dev: [nq, k] on device.  Enqueues the search; the caller synchronises
// `st` and then calls index_search_finish.  Batches of more than kScanNqMax queries are processed in chunks, each
// synchronised here (the per-query scratch is per chunk).
static int index_search_enqueue(hr_index* h, const float* q_dev, int64_t nq, int k, float* D_dev, int64_t* I_dev,
                                cudaStream_t st) {
  h->launches0 = g_launches.load();
  h->stats = hr_scan_stats{};
  h->pend = false;
  if (nq == 0) return HR_OK;
  cudaEventRecord(h->ev[0], st);
  for (int64_t q0 = 0; q0 < nq; q0 += kScanNqMax) {
    const int nb = (int)std::min<int64_t>(kScanNqMax, nq - q0);
    HR_TRY(chunk_enqueue(h, q_dev + q0 * h->d, nb, k, D_dev + q0 * k, I_dev + q0 * k, st));
    if (q0 + kScanNqMax < nq) {
      HR_CUDA(cudaStreamSynchronize(st));
      HR_TRY(chunk_finish(h, st, nullptr));
    }
  }
  cudaEventRecord(h->ev[1], st);
  return HR_OK;
}
static int index_search_finish(hr_index* h, cudaStream_t st, bool* changed) {
  HR_TRY(chunk_finish(h, st, changed));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->stats.total_ms = ms;
  h->stats.launches = g_launches.load() - h->launches0;
  return HR_OK;
}

extern "C" int hr_index_search(hr_index* h, const float* q, int64_t nq, int k, float* D, int64_t* I,
                               int io_on_device, void* stream) {
  if (!h) return set_err(HR_ERR_INVALID, "null index");
  if (nq < 0) return set_err(HR_ERR_INVALID, "nq < 0");
  if (k <= 0 || k > HR_MAX_K) return set_err(HR_ERR_INVALID, "k must be in [1, 2048]");
  if (nq == 0) return HR_OK;
  if (!q || !D || !I) return set_err(HR_ERR_INVALID, "null q / D / I");
  HR_DEVICE(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  cudaStream_t st = (cudaStream_t)stream;
  if (io_on_device) {
    HR_TRY(index_search_enqueue(h, q, nq, k, D, I, st));
    HR_CUDA(cudaStreamSynchronize(st));
    return index_search_finish(h, st, nullptr);
  }
  HR_TRY(h->io_q.ensure((size_t)nq * h->d * 4));
  HR_TRY(h->io_D.ensure((size_t)nq * k * 4));
  HR_TRY(h->io_I.ensure((size_t)nq * k * 8));
  HR_CUDA(cudaMemcpyAsync(h->io_q.p, q, (size_t)nq * h->d * 4, cudaMemcpyHostToDevice, st));
  HR_TRY(index_search_enqueue(h, h->io_q.as<float>(), nq, k, h->io_D.as<float>(), h->io_I.as<int64_t>(), st));
  HR_CUDA(cudaMemcpyAsync(D, h->io_D.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
  HR_CUDA(cudaMemcpyAsync(I, h->io_I.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
  HR_CUDA(cudaStreamSynchronize(st));
  bool changed = false;
  HR_TRY(index_search_finish(h, st, &changed));
  if (changed) {
    HR_CUDA(cudaMemcpyAsync(D, h->io_D.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    HR_CUDA(cudaMemcpyAsync(I, h->io_I.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    HR_CUDA(cudaStreamSynchronize(st));
  }
  return HR_OK;
}

// ---- persistence: faiss flat-index bytes -----------------------------
extern "C" int hr_index_save(hr_index* h, const char* path) {
  if (!h || !path) return set_err(HR_ERR_INVALID, "null argument");
  FILE* f = fopen(path, "wb");
  if (!f) return set_err(HR_ERR_IO, std::string("cannot open for writing: ") + path);
  const char* fourcc = h->metric == HR_METRIC_INNER_PRODUCT ? "IxFI" : "IxF2";
  int32_t d = h->d, metric = h->metric;
  int64_t nt = h->ntotal, dummy = 1 << 20;
  uint8_t trained = 1;
  uint64_t count = (uint64_t)h->ntotal * h->d;
  bool ok = fwrite(fourcc, 1, 4, f) == 4 && fwrite(&d, 4, 1, f) == 1 && fwrite(&nt, 8, 1, f) == 1 &&
            fwrite(&dummy, 8, 1, f) == 1 && fwrite(&dummy, 8, 1, f) == 1 && fwrite(&trained, 1, 1, f) == 1 &&
            fwrite(&metric, 4, 1, f) == 1 && fwrite(&count, 8, 1, f) == 1;
  const int64_t chunk = std::max<int64_t>(1, ((int64_t)64 << 20) / ((int64_t)h->d * 4));
  std::vector<float> buf;
  for (int64_t r0 = 0; ok && r0 < h->ntotal; r0 += chunk) {
    const int64_t nr = std::min(chunk, h->ntotal - r0);
    buf.resize((size_t)nr * h->d);
    int rc = hr_index_reconstruct(h, r0, nr, buf.data());
    if (rc != HR_OK) {
      fclose(f);
      return rc;
    }
    ok = fwrite(buf.data(), 4, buf.size(), f) == buf.size();
  }
  if (fclose(f) != 0) ok = false;
  if (!ok) return set_err(HR_ERR_IO, std::string("write failed: ") + path);
  return HR_OK;
}

extern "C" int hr_index_load(const char* path, int device, int storage_dtype, hr_index** out) {
  if (!path || !out) return set_err(HR_ERR_INVALID, "null argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return set_err(HR_ERR_IO, std::string("cannot open index file: ") + path);
  char fourcc[4];
  int32_t d = 0, metric = 0;
  int64_t nt = 0, dummy = 0;
  uint8_t trained = 0;
  uint64_t count = 0;
  bool ok = fread(fourcc, 1, 4, f) == 4 && fread(&d, 4, 1, f) == 1 && fread(&nt, 8, 1, f) == 1 &&
            fread(&dummy, 8, 1, f) == 1 && fread(&dummy, 8, 1, f) == 1 && fread(&trained, 1, 1, f) == 1 &&
            fread(&metric, 4, 1, f) == 1;
  if (ok && metric > 1) {
    float arg;
    ok = fread(&arg, 4, 1, f) == 1;
  }
  ok = ok && fread(&count, 8, 1, f) == 1;
  const bool known = !memcmp(fourcc, "IxF2", 4) || !memcmp(fourcc, "IxFI", 4) || !memcmp(fourcc, "IxFl", 4);
  if (!ok || !known || d <= 0 || nt < 0 || count != (uint64_t)nt * (uint64_t)d || metric < 0 || metric > 1) {
    fclose(f);
    return set_err(HR_ERR_IO, std::string("not a faiss flat index (IxF2/IxFI) or corrupt header: ") + path);
  }
  hr_index* h = nullptr;
  int rc = hr_index_create(d, metric, storage_dtype, device, &h);
  if (rc != HR_OK) {
    fclose(f);
    return rc;
  }
  rc = hr_index_reserve(h, nt);
  const int64_t chunk = std::max<int64_t>(1, ((int64_t)64 << 20) / ((int64_t)d * 4));
  std::vector<float> buf;
  for (int64_t r0 = 0; rc == HR_OK && r0 < nt; r0 += chunk) {
    const int64_t nr = std::min(chunk, nt - r0);
    buf.resize((size_t)nr * d);
    if (fread(buf.data(), 4, buf.size(), f) != buf.size()) {
      rc = set_err(HR_ERR_IO, std::string("truncated index file: ") + path);
      break;
    }
    rc = hr_index_add(h, buf.data(), nr, 0, nullptr);
  }
  fclose(f);
  if (rc != HR_OK) {
    std::string keep = g_err;
    hr_index_destroy(h);
    g_err = keep;
    return rc;
  }
  *out = h;
  return HR_OK;
}

// -------------------------------------------------------------------------------------------------
// BM25
// -------------------------------------------------------------------------------------------------
struct hr_bm25 {
  int device = 0;
  int num_sms = 148;
  int64_t N = 0, V = 0, nnz = 0, nnz_pad = 0, id_base = 0;
  int64_t* indptr = nullptr;   // [V+1] unpadded offsets: list lengths / df
  int64_t* pstart = nullptr;   // [V+1] padded list starts (multiples of 4)
  int32_t* post_doc = nullptr; // [nnz_pad + 4]
  float* post_imp = nullptr;   // [nnz_pad + 4]
  float* idf = nullptr;
  DevBuf keys, ns, io_qi, io_qt, io_S, io_I, touched, plan_nt, plan_start, plan_len, plan_wgt, plan_cur, plan_coarse, tau,
      jobctr, plan_err;
  int* h_plan_err = nullptr;   // pinned copy of plan_err (queries with too many distinct terms in the last search)
  // batch-1 / batch-2 searches run next to the dense scan on this stream (candidates_enqueue)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::mutex mu;               // one search at a time per handle (the scratch above is per handle)
};

extern "C" int hr_bm25_destroy(hr_bm25* h) {
  if (!h) return HR_OK;
  DeviceGuard g(h->device);
  if (h->indptr) cudaFree(h->indptr);
  if (h->pstart) cudaFree(h->pstart);
  if (h->post_doc) cudaFree(h->post_doc);
  if (h->post_imp) cudaFree(h->post_imp);
  if (h->idf) cudaFree(h->idf);
  if (h->h_plan_err) cudaFreeHost(h->h_plan_err);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  DevBuf* bufs[] = {&h->keys,    &h->ns,         &h->io_qi,    &h->io_qt,    &h->io_S,     &h->io_I, &h->touched,
                    &h->plan_nt, &h->plan_start, &h->plan_len, &h->plan_wgt, &h->plan_cur, &h->plan_coarse, &h->tau,
                    &h->jobctr,  &h->plan_err};
  for (DevBuf* b : bufs) b->release();
  delete h;
  return HR_OK;
}
extern "C" int64_t hr_bm25_ndocs(const hr_bm25* h) { return h ? h->N : -1; }
extern "C" int64_t hr_bm25_vocab(const hr_bm25* h) { return h ? h->V : -1; }
extern "C" int64_t hr_bm25_nnz(const hr_bm25* h) { return h ? h->nnz : -1; }
extern "C" int hr_bm25_set_id_base(hr_bm25* h, int64_t id_base) {
  if (!h) return set_err(HR_ERR_INVALID, "null bm25");
  h->id_base = id_base;
  return HR_OK;
}

// padded list starts from the (host) CSR offsets: every list begins at a multiple of 4
static bool padded_starts(const std::vector<int64_t>& indptr, std::vector<int64_t>& pstart) {
  const size_t V = indptr.size() - 1;
  pstart.resize(V + 1);
  int64_t p = 0;
  for (size_t t = 0; t < V; ++t) {
    const int64_t len = indptr[t + 1] - indptr[t];
    if (len < 0 || len > (int64_t)0xFFFFFFF0ll) return false;
    pstart[t] = p;
    p += (len + 3) & ~(int64_t)3;
  }
  pstart[V] = p;
  return true;
}

// handle with device arrays for (N, V, nnz) allocated; postings not yet filled
static int bm25_alloc(hr_bm25** out, int device, int64_t n_docs, int64_t vocab, const std::vector<int64_t>& h_indptr,
                      const std::vector<int64_t>& h_pstart) {
  hr_bm25* h = new hr_bm25();
  h->device = device;
  h->N = n_docs;
  h->V = vocab;
  h->nnz = h_indptr[vocab];
  h->nnz_pad = h_pstart[vocab];
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->num_sms = prop.multiProcessorCount;
  const size_t np = (size_t)h->nnz_pad + 4;   // +4: sentinels behind the last list
  if (cudaMalloc((void**)&h->indptr, (size_t)(vocab + 1) * 8) != cudaSuccess ||
      cudaMalloc((void**)&h->pstart, (size_t)(vocab + 1) * 8) != cudaSuccess ||
      cudaMalloc((void**)&h->post_doc, np * 4) != cudaSuccess || cudaMalloc((void**)&h->post_imp, np * 4) != cudaSuccess ||
      cudaMalloc((void**)&h->idf, (size_t)vocab * 4) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_plan_err, 4) != cudaSuccess) {
    (void)cudaGetLastError();
    hr_bm25_destroy(h);
    return set_err(HR_ERR_NOMEM, "cudaMalloc failed for the BM25 index");
  }
  *h->h_plan_err = 0;
  if (cudaMemcpy(h->indptr, h_indptr.data(), (size_t)(vocab + 1) * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->pstart, h_pstart.data(), (size_t)(vocab + 1) * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
    (void)cudaGetLastError();
    hr_bm25_destroy(h);
    return set_err(HR_ERR_CUDA, "copy of BM25 tables failed");
  }
  *out = h;
  return HR_OK;
}

extern "C" int hr_bm25_create(const int64_t* indptr, const int32_t* post_doc, const int32_t* post_tf,
                              const int32_t* doc_len, int64_t n_docs, int64_t vocab, float k1, float b,
                              int idf_variant, int64_t n_docs_global, double avgdl_global,
                              const int64_t* df_global, int is_device, int device, void* stream, hr_bm25** out) {
  if (!out) return set_err(HR_ERR_INVALID, "null out");
  *out = nullptr;
  if (!indptr || n_docs < 0 || vocab <= 0) return set_err(HR_ERR_INVALID, "bad BM25 index arguments");
  if (n_docs >= (int64_t)0x7FFFFFF0ll) return set_err(HR_ERR_INVALID, "too many docs for one shard");
  if (idf_variant != HR_IDF_LUCENE && idf_variant != HR_IDF_OKAPI) return set_err(HR_ERR_INVALID, "idf variant");
  int ndev = 0;
  HR_TRY(hr_device_count(&ndev));
  if (ndev <= 0) return set_err(HR_ERR_CUDA, "no CUDA device (hr_b200 has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_err(HR_ERR_INVALID, "device ordinal out of range");
  HR_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  // host copy of indptr (df, nnz)
  std::vector<int64_t> h_indptr((size_t)vocab + 1);
  if (is_device) {
    HR_CUDA(cudaMemcpyAsync(h_indptr.data(), indptr, (size_t)(vocab + 1) * 8, cudaMemcpyDeviceToHost, st));
    HR_CUDA(cudaStreamSynchronize(st));
  } else {
    memcpy(h_indptr.data(), indptr, (size_t)(vocab + 1) * 8);
  }
  const int64_t nnz = h_indptr[vocab];
  std::vector<int64_t> h_pstart;
  if (h_indptr[0] != 0 || nnz < 0 || !padded_starts(h_indptr, h_pstart))
    return set_err(HR_ERR_INVALID, "indptr must start at 0 and be non-decreasing");
  if (nnz > 0 && (!post_doc || !post_tf || !doc_len)) return set_err(HR_ERR_INVALID, "null postings");
  std::vector<int64_t> h_df;
  if (df_global) {
    h_df.resize((size_t)vocab);
    if (is_device) {
      HR_CUDA(cudaMemcpyAsync(h_df.data(), df_global, (size_t)vocab * 8, cudaMemcpyDeviceToHost, st));
      HR_CUDA(cudaStreamSynchronize(st));
    } else {
      memcpy(h_df.data(), df_global, (size_t)vocab * 8);
    }
  }
  const double Ng = (double)(n_docs_global > 0 ? n_docs_global : n_docs);
  std::vector<float> h_idf((size_t)vocab);
  {
    std::vector<double> raw((size_t)vocab);
    double sum = 0.0;
    int64_t present = 0;
    for (int64_t t = 0; t < vocab; ++t) {
      const double df = (double)(df_global ? h_df[t] : (h_indptr[t + 1] - h_indptr[t]));
      if (idf_variant == HR_IDF_LUCENE) raw[t] = std::log((Ng - df + 0.5) / (df + 0.5) + 1.0);
      else {
        raw[t] = std::log((Ng - df + 0.5) / (df + 0.5));
        if (df > 0) { sum += raw[t]; present++; }
      }
    }
    const double eps = present ? 0.25 * (sum / (double)present) : 0.0;
    for (int64_t t = 0; t < vocab; ++t) {
      double v = raw[t];
      if (idf_variant == HR_IDF_OKAPI && v < 0) v = eps;
      h_idf[t] = (float)v;
    }
  }
  hr_bm25* h = nullptr;
  HR_TRY(bm25_alloc(&h, device, n_docs, vocab, h_indptr, h_pstart));
  DevBuf d_doc, d_tf, d_dl, d_bad;
  auto fail = [&](int code, const char* msg) {
    (void)cudaGetLastError();
    d_doc.release();
    d_tf.release();
    d_dl.release();
    d_bad.release();
    hr_bm25_destroy(h);
    return set_err(code, msg);
  };
  if (cudaMemcpyAsync(h->idf, h_idf.data(), (size_t)vocab * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)
    return fail(HR_ERR_CUDA, "copy of BM25 tables failed");
  double avgdl = avgdl_global;
  unsigned long long n_bad = 0;
  if (nnz > 0) {
    const int32_t* docp = post_doc;
    const int32_t* tfp = post_tf;
    const int32_t* dlp = doc_len;
    if (!is_device) {
      if (d_doc.ensure((size_t)nnz * 4) || d_tf.ensure((size_t)nnz * 4) || d_dl.ensure((size_t)std::max<int64_t>(n_docs, 1) * 4))
        return fail(HR_ERR_NOMEM, "cudaMalloc failed for BM25 build scratch");
      if (cudaMemcpyAsync(d_doc.p, post_doc, (size_t)nnz * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
          cudaMemcpyAsync(d_tf.p, post_tf, (size_t)nnz * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
          cudaMemcpyAsync(d_dl.p, doc_len, (size_t)n_docs * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return fail(HR_ERR_CUDA, "copy of BM25 build inputs failed");
      docp = d_doc.as<int32_t>();
      tfp = d_tf.as<int32_t>();
      dlp = d_dl.as<int32_t>();
    }
    if (!(avgdl > 0.0)) {
      // local average document length (fp64 on the host, like the oracle)
      std::vector<int32_t> h_dl((size_t)n_docs);
      if (cudaMemcpyAsync(h_dl.data(), dlp, (size_t)n_docs * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaStreamSynchronize(st) != cudaSuccess)
        return fail(HR_ERR_CUDA, "copy of doc_len failed");
      double s = 0.0;
      for (int64_t i = 0; i < n_docs; ++i) s += (double)h_dl[i];
      avgdl = n_docs ? s / (double)n_docs : 0.0;
    }
    if (d_bad.ensure(8) || cudaMemsetAsync(d_bad.p, 0, 8, st) != cudaSuccess)
      return fail(HR_ERR_NOMEM, "cudaMalloc failed for BM25 build scratch");
    const int blocks = (int)std::min<int64_t>((nnz + 255) / 256, (int64_t)h->num_sms * 16);
    bm25_pad_impact_kernel<<<blocks, 256, 0, st>>>(h->indptr, h->pstart, docp, tfp, dlp, nnz, vocab, n_docs, (double)k1,
                                                   (double)b, avgdl, h->post_doc, h->post_imp,
                                                   d_bad.as<unsigned long long>());
    g_launches.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess) return fail(HR_ERR_CUDA, "bm25_pad_impact_kernel launch failed");
    if (cudaMemcpyAsync(&n_bad, d_bad.p, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess)
      return fail(HR_ERR_CUDA, "BM25 build failed");
  }
  bm25_pad_sentinels_kernel<<<(unsigned)std::min<int64_t>((vocab + 255) / 256, (int64_t)h->num_sms * 16), 256, 0, st>>>(
      h->indptr, h->pstart, vocab, 4, h->post_doc, h->post_imp);
  g_launches.fetch_add(1);
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    return fail(HR_ERR_CUDA, "BM25 build failed");
  if (n_bad)
    return fail(HR_ERR_INVALID, "bad CSR: doc ids must lie in [0, n_docs) and ascend strictly inside a posting list, tf > 0");
  d_doc.release();
  d_tf.release();
  d_dl.release();
  d_bad.release();
  *out = h;
  return HR_OK;
}

// ---- ingest: CSR build from token occurrences on the device ------------------------------------------
extern "C" int hr_bm25_create_from_tokens(const int32_t* term_ids, const int32_t* doc_ids, int64_t n_tokens,
                                          int64_t n_docs, int64_t vocab, float k1, float b, int idf_variant,
                                          int64_t n_docs_global, double avgdl_global, const int64_t* df_global,
                                          int is_device, int device, void* stream, hr_bm25** out) {
  if (!out) return set_err(HR_ERR_INVALID, "null out");
  *out = nullptr;
  if (n_tokens < 0 || n_docs < 0 || vocab <= 0 || vocab > (int64_t)0x7FFFFFFFll)
    return set_err(HR_ERR_INVALID, "bad BM25 build arguments");
  if (n_tokens > 0 && (!term_ids || !doc_ids)) return set_err(HR_ERR_INVALID, "null token arrays");
  if (n_docs >= (int64_t)0x7FFFFFF0ll) return set_err(HR_ERR_INVALID, "too many docs for one shard");
  int ndev = 0;
  HR_TRY(hr_device_count(&ndev));
  if (ndev <= 0) return set_err(HR_ERR_CUDA, "no CUDA device (hr_b200 has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_err(HR_ERR_INVALID, "device ordinal out of range");
  HR_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf d_t, d_d, keys_a, keys_b, counts, nruns, tmp, indptr, pdoc, ptf, dlen, bad;
  DevBuf* all[] = {&d_t, &d_d, &keys_a, &keys_b, &counts, &nruns, &tmp, &indptr, &pdoc, &ptf, &dlen, &bad};
  auto cleanup = [&]() { for (DevBuf* x : all) x->release(); };
  auto fail = [&](int rc) { cleanup(); return rc; };
  const size_t nt = (size_t)std::max<int64_t>(n_tokens, 1);
  const size_t nd = (size_t)std::max<int64_t>(n_docs, 1);
  int rc = HR_OK;
  if ((rc = keys_a.ensure(nt * 8)) || (rc = keys_b.ensure(nt * 8)) || (rc = counts.ensure(nt * 4)) ||
      (rc = nruns.ensure(8)) || (rc = indptr.ensure((size_t)(vocab + 1) * 8)) || (rc = dlen.ensure(nd * 4)) ||
      (rc = bad.ensure(8)))
    return fail(rc);
  const int32_t* tp = term_ids;
  const int32_t* dp = doc_ids;
  if (!is_device && n_tokens > 0) {
    if ((rc = d_t.ensure(nt * 4)) || (rc = d_d.ensure(nt * 4))) return fail(rc);
    if (cudaMemcpyAsync(d_t.p, term_ids, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d_d.p, doc_ids, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)
      return fail(set_err(HR_ERR_CUDA, "copy of token arrays failed"));
    tp = d_t.as<int32_t>();
    dp = d_d.as<int32_t>();
  }
  if (cudaMemsetAsync(dlen.p, 0, nd * 4, st) != cudaSuccess || cudaMemsetAsync(bad.p, 0, 8, st) != cudaSuccess ||
      cudaMemsetAsync(nruns.p, 0, 8, st) != cudaSuccess)
    return fail(set_err(HR_ERR_CUDA, "memset failed in the BM25 build"));
  int64_t nnz = 0;
  if (n_tokens > 0) {
    const int blocks = (int)std::min<int64_t>((n_tokens + 255) / 256, 148 * 16);
    tokens_to_keys_kernel<<<blocks, 256, 0, st>>>(tp, dp, n_tokens, vocab, n_docs, keys_a.as<uint64_t>(),
                                                  dlen.as<int32_t>(), bad.as<unsigned long long>());
    g_launches.fetch_add(1);
    unsigned long long n_bad = 0;
    if (cudaMemcpyAsync(&n_bad, bad.p, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return fail(set_err(HR_ERR_CUDA, "BM25 build: key kernel failed"));
    if (n_bad) return fail(set_err(HR_ERR_INVALID, "token id outside [0, vocab) or doc id outside [0, n_docs)"));
    int vbits = 1;
    while (((int64_t)1 << vbits) < vocab) ++vbits;
    size_t tb = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), (int64_t)n_tokens, 0,
                                   32 + vbits, st);
    size_t tb2 = 0;
    cub::DeviceRunLengthEncode::Encode(nullptr, tb2, keys_b.as<uint64_t>(), keys_a.as<uint64_t>(), counts.as<int32_t>(),
                                       nruns.as<int64_t>(), (int64_t)n_tokens, st);
    if ((rc = tmp.ensure(std::max(tb, tb2) + 256))) return fail(rc);
    if (cub::DeviceRadixSort::SortKeys(tmp.p, tb, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), (int64_t)n_tokens, 0,
                                       32 + vbits, st) != cudaSuccess ||
        cub::DeviceRunLengthEncode::Encode(tmp.p, tb2, keys_b.as<uint64_t>(), keys_a.as<uint64_t>(),
                                           counts.as<int32_t>(), nruns.as<int64_t>(), (int64_t)n_tokens,
                                           st) != cudaSuccess)
      return fail(set_err(HR_ERR_CUDA, "BM25 build: sort / run-length encode failed"));
    g_launches.fetch_add(2);
    if (cudaMemcpyAsync(&nnz, nruns.p, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return fail(set_err(HR_ERR_CUDA, "BM25 build: run count copy failed"));
  }
  if ((rc = pdoc.ensure((size_t)std::max<int64_t>(nnz, 1) * 4)) || (rc = ptf.ensure((size_t)std::max<int64_t>(nnz, 1) * 4)))
    return fail(rc);
  if (nnz > 0) {
    const int blocks = (int)std::min<int64_t>((nnz + 255) / 256, 148 * 16);
    unique_to_postings_kernel<<<blocks, 256, 0, st>>>(keys_a.as<uint64_t>(), counts.as<int32_t>(), nnz,
                                                      pdoc.as<int32_t>(), ptf.as<int32_t>());
    g_launches.fetch_add(1);
  }
  term_offsets_kernel<<<(unsigned)((vocab + 1 + 255) / 256), 256, 0, st>>>(keys_a.as<uint64_t>(), nnz, vocab,
                                                                          indptr.as<int64_t>());
  g_launches.fetch_add(1);
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    return fail(set_err(HR_ERR_CUDA, "BM25 build: CSR kernels failed"));
  keys_b.release();
  counts.release();
  tmp.release();
  d_t.release();
  d_d.release();
  // the CSR now lives on the device: a host df_global goes there too before the device-side create
  DevBuf dfg;
  const int64_t* dfp = df_global;
  if (df_global && !is_device) {
    if ((rc = dfg.ensure((size_t)vocab * 8))) return fail(rc);
    if (cudaMemcpyAsync(dfg.p, df_global, (size_t)vocab * 8, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      dfg.release();
      return fail(set_err(HR_ERR_CUDA, "copy of df_global failed"));
    }
    dfp = dfg.as<int64_t>();
  }
  rc = hr_bm25_create(indptr.as<int64_t>(), pdoc.as<int32_t>(), ptf.as<int32_t>(), dlen.as<int32_t>(), n_docs, vocab,
                      k1, b, idf_variant, n_docs_global, avgdl_global, dfp, 1, device, stream, out);
  dfg.release();
  cleanup();
  return rc;
}

// ---- persistence: the BM25 sidecar the reference never wrote ---------------------
// "HRBM25\0\2" | int64 N, V, nnz, id_base, nnz_pad | indptr int64[V+1] | idf fp32[V] |
// post_doc int32[nnz_pad] | post_imp fp32[nnz_pad]   (padded lists: bm25.cuh; pstart is derived from indptr)
static const char kBm25Magic[8] = {'H', 'R', 'B', 'M', '2', '5', 0, 2};

static bool dev_to_file(FILE* f, const void* dev, size_t bytes, std::vector<char>& buf) {
  const size_t chunk = (size_t)64 << 20;
  for (size_t o = 0; o < bytes; o += chunk) {
    const size_t n = std::min(chunk, bytes - o);
    buf.resize(n);
    if (cudaMemcpy(buf.data(), (const char*)dev + o, n, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    if (fwrite(buf.data(), 1, n, f) != n) return false;
  }
  return true;
}
static bool file_to_dev(FILE* f, void* dev, size_t bytes, std::vector<char>& buf) {
  const size_t chunk = (size_t)64 << 20;
  for (size_t o = 0; o < bytes; o += chunk) {
    const size_t n = std::min(chunk, bytes - o);
    buf.resize(n);
    if (fread(buf.data(), 1, n, f) != n) return false;
    if (cudaMemcpy((char*)dev + o, buf.data(), n, cudaMemcpyHostToDevice) != cudaSuccess) return false;
  }
  return true;
}

extern "C" int hr_bm25_save(hr_bm25* h, const char* path) {
  if (!h || !path) return set_err(HR_ERR_INVALID, "null argument");
  HR_DEVICE(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  FILE* f = fopen(path, "wb");
  if (!f) return set_err(HR_ERR_IO, std::string("cannot open for writing: ") + path);
  int64_t hdr[5] = {h->N, h->V, h->nnz, h->id_base, h->nnz_pad};
  std::vector<char> buf;
  bool ok = fwrite(kBm25Magic, 1, 8, f) == 8 && fwrite(hdr, 8, 5, f) == 5 &&
            dev_to_file(f, h->indptr, (size_t)(h->V + 1) * 8, buf) && dev_to_file(f, h->idf, (size_t)h->V * 4, buf) &&
            dev_to_file(f, h->post_doc, (size_t)h->nnz_pad * 4, buf) &&
            dev_to_file(f, h->post_imp, (size_t)h->nnz_pad * 4, buf);
  if (fclose(f) != 0) ok = false;
  if (!ok) return set_err(HR_ERR_IO, std::string("write failed: ") + path);
  return HR_OK;
}

extern "C" int hr_bm25_load(const char* path, int device, hr_bm25** out) {
  if (!path || !out) return set_err(HR_ERR_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  HR_TRY(hr_device_count(&ndev));
  if (ndev <= 0) return set_err(HR_ERR_CUDA, "no CUDA device (hr_b200 has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_err(HR_ERR_INVALID, "device ordinal out of range");
  FILE* f = fopen(path, "rb");
  if (!f) return set_err(HR_ERR_IO, std::string("cannot open BM25 index file: ") + path);
  char magic[8];
  int64_t hdr[5] = {0, 0, 0, 0, 0};
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kBm25Magic, 8) != 0 || fread(hdr, 8, 5, f) != 5 || hdr[0] < 0 ||
      hdr[0] >= (int64_t)0x7FFFFFF0ll || hdr[1] <= 0 || hdr[1] > (int64_t)0x7FFFFFFFll || hdr[2] < 0 || hdr[4] < hdr[2]) {
    fclose(f);
    return set_err(HR_ERR_IO, std::string("not a BM25 index file (HRBM25 v2) or corrupt header: ") + path);
  }
  const int64_t V = hdr[1];
  std::vector<int64_t> h_indptr((size_t)V + 1), h_pstart;
  bool ok = fread(h_indptr.data(), 8, (size_t)V + 1, f) == (size_t)V + 1 && h_indptr[0] == 0 && h_indptr[V] == hdr[2] &&
            padded_starts(h_indptr, h_pstart) && h_pstart[V] == hdr[4];
  if (!ok) {
    fclose(f);
    return set_err(HR_ERR_IO, std::string("corrupt BM25 index file (term offsets): ") + path);
  }
  HR_DEVICE(device);
  hr_bm25* h = nullptr;
  int rc = bm25_alloc(&h, device, hdr[0], V, h_indptr, h_pstart);
  if (rc != HR_OK) {
    fclose(f);
    return rc;
  }
  h->id_base = hdr[3];
  std::vector<char> buf;
  ok = file_to_dev(f, h->idf, (size_t)V * 4, buf) && file_to_dev(f, h->post_doc, (size_t)h->nnz_pad * 4, buf) &&
       file_to_dev(f, h->post_imp, (size_t)h->nnz_pad * 4, buf);
  fclose(f);
  if (!ok) {
    hr_bm25_destroy(h);
    return set_err(HR_ERR_IO, std::string("truncated BM25 index file: ") + path);
  }
  // what the scoring kernels rely on: ascending in-range doc ids, sentinel padding
  DevBuf bad;
  unsigned long long n_bad = 1;
  if (bad.ensure(8) == HR_OK && cudaMemset(bad.p, 0, 8) == cudaSuccess) {
    bm25_check_padded_kernel<<<h->num_sms * 8, 256>>>(h->indptr, h->pstart, h->post_doc, V, h->N,
                                                     bad.as<unsigned long long>());
    // the 4 sentinels behind the last list are not part of the file
    bm25_pad_sentinels_kernel<<<(unsigned)std::min<int64_t>((V + 255) / 256, (int64_t)h->num_sms * 16), 256>>>(
        h->indptr, h->pstart, V, 4, h->post_doc, h->post_imp);
    g_launches.fetch_add(2);
    if (cudaMemcpy(&n_bad, bad.p, 8, cudaMemcpyDeviceToHost) != cudaSuccess) n_bad = 1;
  }
  bad.release();
  if (n_bad) {
    (void)cudaGetLastError();
    hr_bm25_destroy(h);
    return set_err(HR_ERR_IO, std::string("corrupt BM25 index file (posting lists): ") + path);
  }
  *out = h;
  return HR_OK;
}

// device-pointer search; does not synchronise.  n_terms = length of q_terms (sizes the plan).
// narrow: one-warp CTAs (14 KB of shared memory each) that fit on an SM NEXT TO a persistent dense-scan CTA: the
// batch-1 search then runs concurrently with the scan instead of in front of it.
static int bm25_search_dev(hr_bm25* h, const int32_t* qi_dev, const int32_t* qt_dev, int64_t nq, int64_t n_terms,
                           int k, float* S_dev, int64_t* I_dev, cudaStream_t st, unsigned long long* touched_dev,
                           bool narrow = false) {
  if (nq == 0) return HR_OK;
  if (k > kBmMaxK) return set_err(HR_ERR_INVALID, "bm25: k must be <= 128");
  if (nq >= (int64_t)0x7FFFFFF0ll) return set_err(HR_ERR_INVALID, "bm25: too many queries in one call");
  if (n_terms < 0) {  // unknown to the caller: read q_indptr[nq] back (one small synchronous copy)
    int32_t last = 0;
    HR_CUDA(cudaMemcpyAsync(&last, qi_dev + nq, 4, cudaMemcpyDeviceToHost, st));
    HR_CUDA(cudaStreamSynchronize(st));
    n_terms = last;
  }
  int kcp = 32;
  while (kcp < k) kcp <<= 1;
  const bool wide = tuning().bm25_wide != 0;
  const int slice_docs = kcp <= 64 ? (wide ? kSwSliceWideA : kSwSliceA) : (wide ? kSwSliceWideB : kSwSliceB);
  const int64_t nsl = std::max<int64_t>(1, (h->N + slice_docs - 1) / slice_docs);   // slices; nsl + 1 boundaries
  const int64_t nbc = (nsl + kBsCoarse - 1) / kBsCoarse + 1;                    // coarse boundaries
  const size_t nterm_slots = (size_t)std::max<int64_t>(n_terms, 1);
  const size_t cur_bytes = nterm_slots * (size_t)(nsl + 1) * 4;
  if (cur_bytes > ((size_t)16 << 30))
    return set_err(HR_ERR_INVALID, "bm25: query batch too large for the cursor table (split the batch)");
  int64_t S;   // doc windows (spans) per query
  int spc;     // slices per window
  bool use_lock = false;
  {
    // jobs = (window, query), one warp each, drawn window-major by the resident warps.  Enough jobs for a
    // balanced tail (~24 per resident warp), windows small enough that the posting ranges all queries of the
    // batch read inside one window stay in L2 (~192k docs), bounded by the merge capacity.
    const int64_t resident = narrow ? (int64_t)h->num_sms : (int64_t)h->num_sms * (wide ? 1 : 2) * kSwWarps;
    S = std::max<int64_t>((24 * resident + nq - 1) / nq, (h->N + 196607) / 196608);
    if (tuning().bm25_spans > 0) S = tuning().bm25_spans;
    // Large batches keep one running top-k list per query under a lock: at most ~4 jobs of a query run at the
    // same time.  In a smaller batch many jobs of one query finish together and would serialise there, so they
    // write per-job slots that one merge kernel combines (S * k bounded by its sort buffer).  A job is at least 8
    // slices long, or its epilogue (a 128-key sort) weighs as much as its work.
    use_lock = nq * 4 >= resident;
    S = std::max<int64_t>(1, std::min<int64_t>({S, std::max<int64_t>(1, nsl / 8),
                                                use_lock ? (int64_t)1024 : (int64_t)(kBmMergeCap / k)}));
    if ((uint64_t)nq * (uint64_t)S >= 0xFFFF0000ull) return set_err(HR_ERR_INVALID, "bm25: too many (query, window) jobs");
  }
  spc = (int)((nsl + S - 1) / S);
  S = (nsl + spc - 1) / spc;
  HR_TRY(h->plan_nt.ensure((size_t)nq * 4));
  HR_TRY(h->plan_start.ensure(nterm_slots * 8));
  HR_TRY(h->plan_len.ensure(nterm_slots * 4));
  HR_TRY(h->plan_wgt.ensure(nterm_slots * 4));
  HR_TRY(h->plan_cur.ensure(cur_bytes));
  HR_TRY(h->plan_coarse.ensure(nterm_slots * (size_t)nbc * 4));
  HR_TRY(h->tau.ensure((size_t)nq * 8));
  // large batches: running top-k list per query + count + lock; small ones: per-job slots [nq][S][k]
  HR_TRY(h->keys.ensure(use_lock ? (size_t)nq * k * 8 : (size_t)nq * S * k * 8));
  HR_TRY(h->ns.ensure(use_lock ? (size_t)nq * 8 : (size_t)nq * S * 4));
  HR_TRY(h->jobctr.ensure(4));
  HR_CUDA(cudaMemsetAsync(h->tau.p, 0, (size_t)nq * 8, st));
  HR_TRY(h->plan_err.ensure(4));
  HR_CUDA(cudaMemsetAsync(h->plan_err.p, 0, 4, st));
  bm25_plan_terms_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(
      h->indptr, h->pstart, h->idf, h->V, qi_dev, qt_dev, (int)nq, h->plan_nt.as<int>(), h->plan_start.as<int64_t>(),
      h->plan_len.as<uint32_t>(), h->plan_wgt.as<float>(), touched_dev, h->plan_err.as<int>());
  HR_LAUNCHED();
  HR_CUDA(cudaMemcpyAsync(h->h_plan_err, h->plan_err.p, 4, cudaMemcpyDeviceToHost, st));
  {
    // coarse boundaries (every kBsCoarse slices) by a search over the whole list, then every slice boundary
    // inside its bracketing coarse pair
    const int gc = bm25_plan_groups(nq, nbc);
    dim3 gridc((unsigned)nq, (unsigned)((nbc * gc + 255) / 256));
    bm25_plan_cursors_kernel<<<gridc, 256, 0, st>>>(h->post_doc, qi_dev, h->plan_nt.as<int>(),
                                                    h->plan_start.as<int64_t>(), h->plan_len.as<uint32_t>(), nbc,
                                                    (int64_t)slice_docs * kBsCoarse, nullptr, 0, 1,
                                                    h->plan_coarse.as<uint32_t>(), gc);
    HR_LAUNCHED();
    const int gf = bm25_plan_groups(nq, nsl + 1);
    dim3 grid((unsigned)nq, (unsigned)(((nsl + 1) * gf + 255) / 256));
    bm25_plan_cursors_kernel<<<grid, 256, 0, st>>>(h->post_doc, qi_dev, h->plan_nt.as<int>(),
                                                   h->plan_start.as<int64_t>(), h->plan_len.as<uint32_t>(), nsl + 1,
                                                   (int64_t)slice_docs, h->plan_coarse.as<uint32_t>(), nbc, kBsCoarse,
                                                   h->plan_cur.as<uint32_t>(), gf);
    HR_LAUNCHED();
  }
  {
    HR_CUDA(cudaMemsetAsync(h->jobctr.p, 0, 4, st));
    if (use_lock) HR_CUDA(cudaMemsetAsync(h->ns.p, 0, (size_t)nq * 8, st));   // gcount [nq] | glock [nq]
    const int warps = narrow ? 1 : kSwWarps;
    const int smem = warps * sw_warp_bytes(slice_docs, kcp);
    const int64_t njobs = nq * S;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->num_sms * (narrow || wide ? 1 : 2), (njobs + warps - 1) / warps));
    auto launch = [&](auto kern) -> int {
      HR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwWarps * sw_warp_bytes(slice_docs, kcp)));
      kern<<<grid, 32 * warps, smem, st>>>(h->post_doc, h->post_imp, qi_dev, h->plan_nt.as<int>(),
                                           h->plan_start.as<int64_t>(), h->plan_wgt.as<float>(),
                                           h->plan_cur.as<uint32_t>(), nsl, spc, (int)S, (int)nq, k, kcp,
                                           h->keys.as<uint64_t>(), h->ns.as<int>(), h->ns.as<int>() + nq,
                                           h->tau.as<unsigned long long>(), h->jobctr.as<unsigned int>(),
                                           use_lock ? 1 : 0);
      HR_LAUNCHED();
      return HR_OK;
    };
    const int batch = tuning().bm25_batch;
    if (slice_docs == kSwSliceA) {
      if (batch == 2) HR_TRY(launch(bm25_sweep_kernel<kSwSliceA, 2>));
      else HR_TRY(launch(bm25_sweep_kernel<kSwSliceA, 3>));
    } else if (slice_docs == kSwSliceB) {
      if (batch == 2) HR_TRY(launch(bm25_sweep_kernel<kSwSliceB, 2>));
      else HR_TRY(launch(bm25_sweep_kernel<kSwSliceB, 3>));
    } else if (slice_docs == kSwSliceWideA) {
      if (batch == 2) HR_TRY(launch(bm25_sweep_kernel<kSwSliceWideA, 3>));
      else HR_TRY(launch(bm25_sweep_kernel<kSwSliceWideA, 4>));
    } else {
      if (batch == 2) HR_TRY(launch(bm25_sweep_kernel<kSwSliceWideB, 3>));
      else HR_TRY(launch(bm25_sweep_kernel<kSwSliceWideB, 4>));
    }
    if (use_lock)
      bm25_sweep_finish_kernel<<<(unsigned)((nq * k + 255) / 256), 256, 0, st>>>(h->keys.as<uint64_t>(), h->ns.as<int>(),
                                                                                 (int)nq, k, k, h->id_base, S_dev, I_dev);
    else
      bm25_sweep_merge_kernel<<<(unsigned)nq, 256, 0, st>>>(h->keys.as<uint64_t>(), h->ns.as<int>(), (int)S, k, k,
                                                            h->tau.as<unsigned long long>(), h->id_base, S_dev, I_dev);
    HR_LAUNCHED();
    return HR_OK;
  }
}

static int bm25_check_plan(hr_bm25* h) {   // after a synchronisation of the stream bm25_search_dev ran on
  if (*h->h_plan_err) {
    *h->h_plan_err = 0;
    return set_err(HR_ERR_INVALID, "a query has more than 64 distinct scorable terms");
  }
  return HR_OK;
}
static int check_query_csr_host(const int32_t* q_indptr, int64_t nq) {
  if (q_indptr[0] < 0) return set_err(HR_ERR_INVALID, "q_indptr must start at a non-negative offset");
  for (int64_t i = 0; i < nq; ++i)
    if (q_indptr[i + 1] < q_indptr[i]) return set_err(HR_ERR_INVALID, "q_indptr must be non-decreasing");
  return HR_OK;
}

extern "C" int hr_bm25_search(hr_bm25* h, const int32_t* q_indptr, const int32_t* q_terms, int64_t nq,
                              int64_t n_terms, int k, float* S, int64_t* I, int io_on_device, void* stream,
                              int64_t* postings_touched) {
  if (!h) return set_err(HR_ERR_INVALID, "null bm25");
  if (nq < 0) return set_err(HR_ERR_INVALID, "nq < 0");
  if (k <= 0 || k > kBmMaxK) return set_err(HR_ERR_INVALID, "bm25 k must be in [1, 128]");
  if (nq == 0) return HR_OK;
  if (!q_indptr || !S || !I) return set_err(HR_ERR_INVALID, "null argument");
  HR_DEVICE(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  cudaStream_t st = (cudaStream_t)stream;
  HR_TRY(h->touched.ensure(8));
  HR_CUDA(cudaMemsetAsync(h->touched.p, 0, 8, st));
  if (io_on_device) {
    HR_TRY(bm25_search_dev(h, q_indptr, q_terms, nq, n_terms, k, S, I, st, h->touched.as<unsigned long long>()));
  } else {
    HR_TRY(check_query_csr_host(q_indptr, nq));
    const int64_t nterms = q_indptr[nq];
    if (n_terms >= 0 && n_terms < nterms) return set_err(HR_ERR_INVALID, "n_terms is smaller than q_indptr[nq]");
    HR_TRY(h->io_qi.ensure((size_t)(nq + 1) * 4));
    HR_TRY(h->io_qt.ensure((size_t)std::max<int64_t>(nterms, 1) * 4));
    HR_TRY(h->io_S.ensure((size_t)nq * k * 4));
    HR_TRY(h->io_I.ensure((size_t)nq * k * 8));
    HR_CUDA(cudaMemcpyAsync(h->io_qi.p, q_indptr, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice, st));
    if (nterms > 0) HR_CUDA(cudaMemcpyAsync(h->io_qt.p, q_terms, (size_t)nterms * 4, cudaMemcpyHostToDevice, st));
    HR_TRY(bm25_search_dev(h, h->io_qi.as<int32_t>(), h->io_qt.as<int32_t>(), nq, nterms, k, h->io_S.as<float>(),
                           h->io_I.as<int64_t>(), st, h->touched.as<unsigned long long>()));
    HR_CUDA(cudaMemcpyAsync(S, h->io_S.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    HR_CUDA(cudaMemcpyAsync(I, h->io_I.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
  }
  unsigned long long t = 0;
  HR_CUDA(cudaMemcpyAsync(&t, h->touched.p, 8, cudaMemcpyDeviceToHost, st));
  HR_CUDA(cudaStreamSynchronize(st));
  if (postings_touched) *postings_touched = (int64_t)t;
  return bm25_check_plan(h);
}

// -------------------------------------------------------------------------------------------------
// fusion / merge
// -------------------------------------------------------------------------------------------------
extern "C" int hr_merge_topk(const float* S, const int64_t* I, int64_t nq, int n_cand, int k, int largest,
                             float pad_score, float* out_S, int64_t* out_I, int device, void* stream) {
  if (nq < 0 || n_cand <= 0 || k <= 0) return set_err(HR_ERR_INVALID, "bad merge arguments");
  if (n_cand > kMergeTopkCap) return set_err(HR_ERR_INVALID, "merge: more than 2048 candidates per query");
  if (nq == 0) return HR_OK;
  if (!S || !I || !out_S || !out_I) return set_err(HR_ERR_INVALID, "null argument");
  HR_DEVICE(device);
  merge_topk_kernel<<<(unsigned)nq, 256, 0, (cudaStream_t)stream>>>(S, I, n_cand, n_cand, 0, 0, k, largest, pad_score,
                                                                    out_S, out_I, 0);
  HR_LAUNCHED();
  return HR_OK;
}

extern "C" int hr_fuse(const float* dense_D, const int64_t* dense_I, const float* bm25_S, const int64_t* bm25_I,
                       const float* bm25_max, int64_t nq, int kc, int top_k, int metric, int mode, float w_vec,
                       float w_bm25, float* out_S, int64_t* out_I, int device, void* stream) {
  if (nq < 0 || kc <= 0 || top_k <= 0) return set_err(HR_ERR_INVALID, "bad fuse arguments");
  if (kc > kFuseMaxKc) return set_err(HR_ERR_INVALID, "fuse: candidate depth kc must be <= 256");
  if (mode != HR_FUSE_WEIGHTED && mode != HR_FUSE_RRF) return set_err(HR_ERR_INVALID, "unknown fusion mode");
  if (nq == 0) return HR_OK;
  if (!dense_D || !dense_I || !bm25_S || !bm25_I || !out_S || !out_I) return set_err(HR_ERR_INVALID, "null argument");
  HR_DEVICE(device);
  fuse_kernel<<<(unsigned)nq, 256, 0, (cudaStream_t)stream>>>(dense_D, dense_I, bm25_S, bm25_I, bm25_max, kc, top_k,
                                                             metric, mode, w_vec, w_bm25, out_S, out_I);
  HR_LAUNCHED();
  return HR_OK;
}

extern "C" int hr_rank_pages(const float* S, const int64_t* I, int64_t nq, int k, int score_kind,
                             const int32_t* page_of_row, int64_t n_rows, int64_t id_base, int top_pages,
                             int32_t* out_page, double* out_score, int32_t* out_count, int device, void* stream) {
  if (nq < 0 || k <= 0 || k > kPageMaxHits) return set_err(HR_ERR_INVALID, "rank_pages: k must be in [1, 256]");
  if (top_pages <= 0 || top_pages > kPageMaxHits) return set_err(HR_ERR_INVALID, "rank_pages: top_pages must be in [1, 256]");
  if (score_kind != 0 && score_kind != 1) return set_err(HR_ERR_INVALID, "rank_pages: score_kind must be 0 or 1");
  if (nq == 0) return HR_OK;
  if (!S || !I || !page_of_row || !out_page || !out_score || !out_count || n_rows < 0)
    return set_err(HR_ERR_INVALID, "null argument");
  HR_DEVICE(device);
  page_rank_kernel<<<(unsigned)nq, 128, 0, (cudaStream_t)stream>>>(S, I, k, score_kind, page_of_row, n_rows, id_base,
                                                                  top_pages, out_page, out_score, out_count);
  HR_LAUNCHED();
  return HR_OK;
}

// -------------------------------------------------------------------------------------------------
// row-sharded retrieval: local candidates of one shard, and the merge + fusion of the gathered shards
// -------------------------------------------------------------------------------------------------
// BM25 then dense for one shard, all enqueued on `st` (no synchronisation): S,J then D,I [nq,kc] on the device
static int candidates_enqueue(hr_index* ix, hr_bm25* bm, const float* q, const int32_t* qi, const int32_t* qt,
                              int64_t nq, int64_t n_terms, int kc, float* D, int64_t* I, float* S, int64_t* J,
                              cudaStream_t st, cudaEvent_t q_ready = nullptr) {
  if (bm && nq <= 16 && ix->ntotal > 0) {
    // Latency mode (the reference's real operating point: one query, rag/storage/faiss_index.py:81): BM25 on a
    // side stream in one-warp CTAs that share the SMs with the persistent scan CTAs (16 KB of shared memory are
    // free next to one); the step then takes what the dense search takes.  One warp per SM scores ~10 queries in
    // the time the HBM-bound scan needs, hence the limit; bigger batches are serialised (their scan is
    // tensor- and power-bound and nothing overlaps: DESIGN.md section 4).
    if (!bm->side) {
      if (cudaStreamCreateWithFlags(&bm->side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&bm->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&bm->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        (void)cudaGetLastError();
        return set_err(HR_ERR_CUDA, "could not create the BM25 side stream");
      }
    }
    HR_CUDA(cudaEventRecord(bm->ev_fork, st));
    HR_CUDA(cudaStreamWaitEvent(bm->side, bm->ev_fork, 0));
    HR_TRY(bm25_search_dev(bm, qi, qt, nq, n_terms, kc, S, J, bm->side, nullptr, true));
    HR_CUDA(cudaEventRecord(bm->ev_join, bm->side));
    HR_TRY(index_search_enqueue(ix, q, nq, kc, D, I, st));
    HR_CUDA(cudaStreamWaitEvent(st, bm->ev_join, 0));
    return HR_OK;
  }
  if (bm) {
    HR_TRY(bm25_search_dev(bm, qi, qt, nq, n_terms, kc, S, J, st, nullptr));
  } else {
    fill_pad_kernel<<<(int)std::min<int64_t>((nq * kc + 255) / 256, 1024), 256, 0, st>>>(S, J, nq * kc, 0.f);
    HR_LAUNCHED();
  }
  if (q_ready) HR_CUDA(cudaStreamWaitEvent(st, q_ready, 0));   // embeddings uploaded on another stream meanwhile
  return index_search_enqueue(ix, q, nq, kc, D, I, st);
}

static int merge_fuse_enqueue(hr_index* ix, const float* D, const int64_t* I, const float* S, const int64_t* J,
                              int n_lists, int64_t list_stride_bytes, int64_t nq, int kc, int top_k, int mode,
                              float w_vec, float w_bm25, float* out_S, int64_t* out_I, cudaStream_t st) {
  const float* dD = D;
  const int64_t* dI = I;
  const float* bS = S;
  const int64_t* bI = J;
  if (n_lists > 1) {
    RetrieveScratch& rs = ix->rs;
    HR_TRY(rs.dD.ensure((size_t)nq * kc * 4));
    HR_TRY(rs.dI.ensure((size_t)nq * kc * 8));
    HR_TRY(rs.bS.ensure((size_t)nq * kc * 4));
    HR_TRY(rs.bI.ensure((size_t)nq * kc * 8));
    const int largest = ix->metric == HR_METRIC_INNER_PRODUCT;
    merge_topk_kernel<<<(unsigned)nq, 256, 0, st>>>(D, I, n_lists * kc, kc, list_stride_bytes / 4, list_stride_bytes / 8,
                                                    kc, largest, largest ? HR_NEG_INF : -HR_NEG_INF, rs.dD.as<float>(),
                                                    rs.dI.as<int64_t>(), 1);
    HR_LAUNCHED();
    merge_topk_kernel<<<(unsigned)nq, 256, 0, st>>>(S, J, n_lists * kc, kc, list_stride_bytes / 4, list_stride_bytes / 8,
                                                    kc, 1, 0.f, rs.bS.as<float>(), rs.bI.as<int64_t>(), 1);
    HR_LAUNCHED();
    dD = rs.dD.as<float>();
    dI = rs.dI.as<int64_t>();
    bS = rs.bS.as<float>();
    bI = rs.bI.as<int64_t>();
  }
  fuse_kernel<<<(unsigned)nq, 256, 0, st>>>(dD, dI, bS, bI, nullptr, kc, top_k, ix->metric, mode, w_vec, w_bm25, out_S,
                                            out_I);
  HR_LAUNCHED();
  return HR_OK;
}

extern "C" int hr_candidates(hr_index* ix, hr_bm25* bm, const float* q, const int32_t* q_indptr,
                             const int32_t* q_terms, int64_t nq, int64_t n_terms, int kc, float* D, int64_t* I,
                             float* S, int64_t* J, void* stream) {
  if (!ix) return set_err(HR_ERR_INVALID, "null index");
  if (nq < 0 || kc <= 0 || kc > kBmMaxK) return set_err(HR_ERR_INVALID, "candidates: kc must be in [1, 128]");
  if (nq == 0) return HR_OK;
  if (!q || !D || !I || !S || !J) return set_err(HR_ERR_INVALID, "null argument");
  if (bm && !q_indptr) return set_err(HR_ERR_INVALID, "null query tokens");
  if (bm && bm->device != ix->device) return set_err(HR_ERR_INVALID, "index and bm25 live on different devices");
  HR_DEVICE(ix->device);
  std::unique_lock<std::mutex> l1(ix->mu), l2;
  if (bm) l2 = std::unique_lock<std::mutex>(bm->mu);
  cudaStream_t st = (cudaStream_t)stream;
  HR_TRY(candidates_enqueue(ix, bm, q, q_indptr, q_terms, nq, n_terms, kc, D, I, S, J, st));
  HR_CUDA(cudaStreamSynchronize(st));
  HR_TRY(index_search_finish(ix, st, nullptr));
  return bm ? bm25_check_plan(bm) : HR_OK;
}

extern "C" int hr_merge_fuse_lists(hr_index* ix, const float* D, const int64_t* I, const float* S, const int64_t* J,
                                   int n_lists, int64_t list_stride_bytes, int64_t nq, int kc, int top_k, int mode,
                                   float w_vec, float w_bm25, float* out_S, int64_t* out_I, void* stream) {
  if (!ix) return set_err(HR_ERR_INVALID, "null index");
  if (nq < 0 || kc <= 0 || top_k <= 0 || n_lists <= 0) return set_err(HR_ERR_INVALID, "bad merge_fuse arguments");
  if (kc > kFuseMaxKc) return set_err(HR_ERR_INVALID, "merge_fuse: candidate depth kc must be <= 256");
  if ((int64_t)n_lists * kc > kMergeTopkCap) return set_err(HR_ERR_INVALID, "merge_fuse: more than 2048 candidates per query");
  if (list_stride_bytes % 8 != 0) return set_err(HR_ERR_INVALID, "merge_fuse: list stride must be a multiple of 8 bytes");
  if (mode != HR_FUSE_WEIGHTED && mode != HR_FUSE_RRF) return set_err(HR_ERR_INVALID, "unknown fusion mode");
  if (nq == 0) return HR_OK;
  if (!D || !I || !S || !J || !out_S || !out_I) return set_err(HR_ERR_INVALID, "null argument");
  HR_DEVICE(ix->device);
  std::lock_guard<std::mutex> lock(ix->mu);
  return merge_fuse_enqueue(ix, D, I, S, J, n_lists, list_stride_bytes, nq, kc, top_k, mode, w_vec, w_bm25, out_S, out_I,
                            (cudaStream_t)stream);
}

// -------------------------------------------------------------------------------------------------
// whole hot path: one device (hr_retrieve) or one row shard per rank with the exchange inside (hr_retrieve_sharded)
// -------------------------------------------------------------------------------------------------
// NCCL is bound at run time (dlopen of the library the host process already uses, e.g. torch's), so libhr_b200.so
// has no link-time NCCL dependency and a world of 1 needs no NCCL at all.
typedef struct { char internal[128]; } hr_nccl_id;   // = ncclUniqueId (NCCL_UNIQUE_ID_BYTES)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(hr_nccl_id*) = nullptr;
  int (*CommInitRank)(void**, int, hr_nccl_id, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("HR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      void* l = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (!l) continue;
      api.GetUniqueId = (int (*)(hr_nccl_id*))dlsym(l, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void**, int, hr_nccl_id, int))dlsym(l, "ncclCommInitRank");
      api.CommDestroy = (int (*)(void*))dlsym(l, "ncclCommDestroy");
      api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(l, "ncclAllGather");
      api.GetErrorString = (const char* (*)(int))dlsym(l, "ncclGetErrorString");
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather) {
        api.lib = l;
        return;
      }
      dlclose(l);
    }
  });
  return api.lib ? &api : nullptr;
}
static int nccl_err(NcclApi* a, int rc, const char* what) {
  std::string m = std::string(what) + " failed";
  if (a && a->GetErrorString) m += std::string(": ") + a->GetErrorString(rc);
  return set_err(HR_ERR_CUDA, m);
}

struct hr_comm {
  int rank = 0, world = 1, device = 0;
  void* nccl = nullptr;        // ncclComm_t (world > 1)
  DevBuf local, gathered, q, qi, qt, oS, oI, flag;
  int* h_flag = nullptr;       // pinned: some rank has fallback queries beyond the device-driven capacity
  std::mutex mu;
};

extern "C" int hr_comm_unique_id(void* out_id_128_bytes) {
  if (!out_id_128_bytes) return set_err(HR_ERR_INVALID, "null id buffer");
  NcclApi* a = nccl_api();
  if (!a) return set_err(HR_ERR_CUDA, "libnccl.so.2 not found (set HR_NCCL_LIB)");
  hr_nccl_id id;
  const int rc = a->GetUniqueId(&id);
  if (rc != 0) return nccl_err(a, rc, "ncclGetUniqueId");
  memcpy(out_id_128_bytes, &id, sizeof id);
  return HR_OK;
}

extern "C" int hr_comm_init(const void* unique_id_128_bytes, int rank, int world, int device, hr_comm** out) {
  if (!out) return set_err(HR_ERR_INVALID, "null out");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return set_err(HR_ERR_INVALID, "bad rank / world");
  if (world > kMergeTopkCap) return set_err(HR_ERR_INVALID, "world too large (the merge handles world * kc <= 2048 candidates per query)");
  int ndev = 0;
  HR_TRY(hr_device_count(&ndev));
  if (ndev <= 0) return set_err(HR_ERR_CUDA, "no CUDA device (hr_b200 has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_err(HR_ERR_INVALID, "device ordinal out of range");
  HR_DEVICE(device);
  hr_comm* c = new hr_comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
  if (cudaMallocHost((void**)&c->h_flag, 4) != cudaSuccess) {
    (void)cudaGetLastError();
    delete c;
    return set_err(HR_ERR_NOMEM, "pinned allocation failed in hr_comm_init");
  }
  *c->h_flag = 0;
  if (world > 1) {
    if (!unique_id_128_bytes) {
      hr_comm_destroy(c);
      return set_err(HR_ERR_INVALID, "null unique id");
    }
    NcclApi* a = nccl_api();
    if (!a) {
      hr_comm_destroy(c);
      return set_err(HR_ERR_CUDA, "libnccl.so.2 not found (set HR_NCCL_LIB)");
    }
    hr_nccl_id id;
    memcpy(&id, unique_id_128_bytes, sizeof id);
    const int rc = a->CommInitRank(&c->nccl, world, id, rank);
    if (rc != 0) {
      c->nccl = nullptr;
      hr_comm_destroy(c);
      return nccl_err(a, rc, "ncclCommInitRank");
    }
  }
  *out = c;
  return HR_OK;
}

extern "C" int hr_comm_destroy(hr_comm* c) {
  if (!c) return HR_OK;
  DeviceGuard g(c->device);
  if (c->nccl) {
    NcclApi* a = nccl_api();
    if (a) a->CommDestroy(c->nccl);
  }
  if (c->h_flag) cudaFreeHost(c->h_flag);
  DevBuf* bufs[] = {&c->local, &c->gathered, &c->q, &c->qi, &c->qt, &c->oS, &c->oI, &c->flag};
  for (DevBuf* b : bufs) b->release();
  delete c;
  return HR_OK;
}
extern "C" int hr_comm_rank(const hr_comm* c) { return c ? c->rank : -1; }
extern "C" int hr_comm_world(const hr_comm* c) { return c ? c->world : -1; }

// trailer of a rank's packed block: [0] = this rank has more fallback queries than the device-driven capacity
__global__ void shard_trailer_kernel(const int* __restrict__ counters, int cap, int* __restrict__ trailer) {
  if (threadIdx.x == 0) {
    trailer[0] = (counters && counters[0] > cap) ? 1 : 0;
    trailer[1] = 0;
  }
}
__global__ void shard_flag_kernel(const uint8_t* __restrict__ gathered, int64_t block_bytes, int64_t trailer_off,
                                  int world, int* __restrict__ flag) {
  if (threadIdx.x == 0) {
    int f = 0;
    for (int r = 0; r < world; ++r) f |= *(const int*)(gathered + (size_t)r * block_bytes + trailer_off);
    *flag = f;
  }
}

static int retrieve_impl(hr_comm* c, hr_index* ix, hr_bm25* bm, const float* q, const int32_t* q_indptr,
                         const int32_t* q_terms, int64_t nq, int64_t n_terms, int top_k, int kc, int mode, float w_vec,
                         float w_bm25, float* out_S, int64_t* out_I, int io_on_device, cudaStream_t st) {
  const int world = c ? c->world : 1;
  // one packed block per rank: D fp32 | S fp32 | I int64 | J int64 ([nq,kc] each) | trailer (16 bytes)
  const int64_t n = nq * kc;
  const int64_t block = 24 * n + 16;
  RetrieveScratch& rs = ix->rs;
  DevBuf& local = c ? c->local : rs.dD;
  HR_TRY(local.ensure((size_t)block));
  uint8_t* lb = local.as<uint8_t>();
  float* lD = (float*)lb;
  float* lS = lD + n;
  int64_t* lI = (int64_t*)(lb + 8 * n);
  int64_t* lJ = lI + n;
  int* trailer = (int*)(lb + 24 * n);
  uint8_t* gb = lb;
  if (world > 1) {
    HR_TRY(c->gathered.ensure((size_t)block * world));
    HR_TRY(c->flag.ensure(4));
    gb = c->gathered.as<uint8_t>();
  }
  const float* qd = q;
  const int32_t* qid = q_indptr;
  const int32_t* qtd = q_terms;
  float* oS = out_S;
  int64_t* oI = out_I;
  bool q_on_copy_stream = false;
  if (!io_on_device) {
    HR_TRY(rs.q.ensure((size_t)nq * ix->d * 4));
    if (bm && nq > 16) {
      // a batch: BM25 runs first and only needs the token ids, so the embeddings (nq * d * 4 bytes) travel on a
      // second stream meanwhile; the dense search waits for them (candidates_enqueue)
      if (!rs.copy) {
        if (cudaStreamCreateWithFlags(&rs.copy, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&rs.ev_copied, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&rs.ev_free, cudaEventDisableTiming) != cudaSuccess) {
          (void)cudaGetLastError();
          return set_err(HR_ERR_CUDA, "could not create the upload stream");
        }
      }
      HR_CUDA(cudaEventRecord(rs.ev_free, st));              // whatever used rs.q before on `st` is ordered first
      HR_CUDA(cudaStreamWaitEvent(rs.copy, rs.ev_free, 0));
      HR_CUDA(cudaMemcpyAsync(rs.q.p, q, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, rs.copy));
      HR_CUDA(cudaEventRecord(rs.ev_copied, rs.copy));
      q_on_copy_stream = true;
    } else {
      HR_CUDA(cudaMemcpyAsync(rs.q.p, q, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, st));
    }
    qd = rs.q.as<float>();
    if (bm) {
      HR_TRY(check_query_csr_host(q_indptr, nq));
      const int64_t nterms = q_indptr[nq];
      n_terms = nterms;
      HR_TRY(rs.qi.ensure((size_t)(nq + 1) * 4));
      HR_TRY(rs.qt.ensure((size_t)std::max<int64_t>(nterms, 1) * 4));
      HR_CUDA(cudaMemcpyAsync(rs.qi.p, q_indptr, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice, st));
      if (nterms > 0) HR_CUDA(cudaMemcpyAsync(rs.qt.p, q_terms, (size_t)nterms * 4, cudaMemcpyHostToDevice, st));
      qid = rs.qi.as<int32_t>();
      qtd = rs.qt.as<int32_t>();
    }
    HR_TRY(rs.oS.ensure((size_t)nq * top_k * 4));
    HR_TRY(rs.oI.ensure((size_t)nq * top_k * 8));
    oS = rs.oS.as<float>();
    oI = rs.oI.as<int64_t>();
  }
  // BM25, dense (decisions on the device), exchange, merge + fusion: one stream, no host round-trip in between
  HR_TRY(candidates_enqueue(ix, bm, qd, qid, qtd, nq, n_terms, kc, lD, lI, lS, lJ, st,
                            q_on_copy_stream ? rs.ev_copied : nullptr));
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (world > 1) {
      NcclApi* a = nccl_api();
      shard_trailer_kernel<<<1, 32, 0, st>>>(ix->pend ? ix->counters.as<int>() : nullptr, ix->pend_cap, trailer);
      HR_LAUNCHED();
      const int rc = a->AllGather(lb, gb, (size_t)block, 1 /* ncclUint8 */, c->nccl, st);
      if (rc != 0) return nccl_err(a, rc, "ncclAllGather");
      shard_flag_kernel<<<1, 32, 0, st>>>(gb, block, 24 * n, world, c->flag.as<int>());
      HR_LAUNCHED();
      HR_CUDA(cudaMemcpyAsync(c->h_flag, c->flag.p, 4, cudaMemcpyDeviceToHost, st));
    }
    HR_TRY(merge_fuse_enqueue(ix, (const float*)gb, (const int64_t*)(gb + 8 * n), (const float*)gb + n,
                              (const int64_t*)(gb + 8 * n) + n, world, block, nq, kc, top_k, mode, w_vec, w_bm25, oS, oI,
                              st));
    if (!io_on_device) {
      HR_CUDA(cudaMemcpyAsync(out_S, oS, (size_t)nq * top_k * 4, cudaMemcpyDeviceToHost, st));
      HR_CUDA(cudaMemcpyAsync(out_I, oI, (size_t)nq * top_k * 8, cudaMemcpyDeviceToHost, st));
    }
    HR_CUDA(cudaStreamSynchronize(st));
    bool changed = false;
    HR_TRY(index_search_finish(ix, st, &changed));
    // A rank with more fallback queries than the device-driven capacity has now finished them on the host path;
    // every rank saw the same trailers, so all of them repeat the exchange and the merge (never in the benchmarks).
    const bool again = world > 1 ? (*c->h_flag != 0) : changed;
    if (!again) break;
  }
  return bm ? bm25_check_plan(bm) : HR_OK;
}

static int retrieve_args(hr_index* ix, hr_bm25* bm, const float* q, const int32_t* q_indptr, int64_t nq, int top_k,
                         int& kc, int mode, float* out_S, int64_t* out_I) {
  if (!ix) return set_err(HR_ERR_INVALID, "null index");
  if (nq < 0 || top_k <= 0) return set_err(HR_ERR_INVALID, "bad retrieve arguments");
  if (mode != HR_FUSE_WEIGHTED && mode != HR_FUSE_RRF) return set_err(HR_ERR_INVALID, "unknown fusion mode");
  if (kc <= 0) kc = top_k > 50 ? top_k : 50;  // live path depth, rag/query/page_retriever.py:81
  if (kc < top_k) kc = top_k;
  if (kc > kBmMaxK) return set_err(HR_ERR_INVALID, "retrieve: candidate depth must be <= 128");
  if (nq > 0 && (!q || !out_S || !out_I)) return set_err(HR_ERR_INVALID, "null argument");
  if (bm && (!q_indptr)) return set_err(HR_ERR_INVALID, "null query tokens");
  if (bm && bm->device != ix->device) return set_err(HR_ERR_INVALID, "index and bm25 live on different devices");
  return HR_OK;
}

extern "C" int hr_retrieve(hr_index* ix, hr_bm25* bm, const float* q, const int32_t* q_indptr,
                           const int32_t* q_terms, int64_t nq, int64_t n_terms, int top_k, int kc, int mode, float w_vec,
                           float w_bm25, float* out_S, int64_t* out_I, int io_on_device, void* stream) {
  HR_TRY(retrieve_args(ix, bm, q, q_indptr, nq, top_k, kc, mode, out_S, out_I));
  if (nq == 0) return HR_OK;
  HR_DEVICE(ix->device);
  std::unique_lock<std::mutex> l1(ix->mu), l2;
  if (bm) l2 = std::unique_lock<std::mutex>(bm->mu);
  return retrieve_impl(nullptr, ix, bm, q, q_indptr, q_terms, nq, n_terms, top_k, kc, mode, w_vec, w_bm25, out_S, out_I,
                       io_on_device, (cudaStream_t)stream);
}

extern "C" int hr_retrieve_sharded(hr_comm* c, hr_index* ix, hr_bm25* bm, const float* q, const int32_t* q_indptr,
                                   const int32_t* q_terms, int64_t nq, int64_t n_terms, int top_k, int kc, int mode,
                                   float w_vec, float w_bm25, float* out_S, int64_t* out_I, int io_on_device,
                                   void* stream) {
  if (!c) return set_err(HR_ERR_INVALID, "null comm");
  HR_TRY(retrieve_args(ix, bm, q, q_indptr, nq, top_k, kc, mode, out_S, out_I));
  if (c->device != ix->device) return set_err(HR_ERR_INVALID, "comm and index live on different devices");
  if ((int64_t)c->world * kc > kMergeTopkCap) return set_err(HR_ERR_INVALID, "retrieve_sharded: world * kc exceeds 2048");
  if (nq == 0) return HR_OK;   // every rank must call with the same nq
  HR_DEVICE(ix->device);
  std::unique_lock<std::mutex> l0(c->mu), l1(ix->mu), l2;
  if (bm) l2 = std::unique_lock<std::mutex>(bm->mu);
  return retrieve_impl(c, ix, bm, q, q_indptr, q_terms, nq, n_terms, top_k, kc, mode, w_vec, w_bm25, out_S, out_I,
                       io_on_device, (cudaStream_t)stream);
}
