// dense_exact.cuh — exact fp32 kernels of the flat index (CUDA cores).
//
//  * convert_pad / row_norms : `add` path (faiss IndexFlat::add, rag/storage/faiss_index.py:124)
//  * exact_scan + exact_merge: exhaustive exact search under the total order (score best first,
//    row asc).  This is the certified *fallback* of the tensor-core filter and HR_MODE_EXACT_SIMT.
//  * rescore_finalize        : exact fp32 re-score of the filter's shortlist, final top-k, and the
//    certificate that decides whether a query needs the fallback.
//
// "Exact score" has ONE definition, used by every kernel here (so fallback and re-score agree
// bit for bit): lane l of a warp accumulates elements i with (i/4)%32 == l in increasing i with a
// single fp32 fma chain, then a 16/8/4/2/1 xor-butterfly sums the 32 lanes.
#pragma once
#include "common.cuh"

namespace hr {

constexpr int kMetricIP = 0;
constexpr int kMetricL2 = 1;

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float4 r;
  r.x = __uint_as_float(u.x << 16);
  r.y = __uint_as_float(u.x & 0xFFFF0000u);
  r.z = __uint_as_float(u.y << 16);
  r.w = __uint_as_float(u.y & 0xFFFF0000u);
  return r;
}

// F queries (rows of q, stride q_stride floats) against one stored row; every lane gets the sums.
template <typename T, int METRIC, int F>
__device__ __forceinline__ void warp_exact_scores(const T* __restrict__ xrow, const float* q, int q_stride,
                                                  int ld, int lane, float (&out)[F]) {
  float acc[F];
#pragma unroll
  for (int f = 0; f < F; ++f) acc[f] = 0.f;
  // up to eight 16-byte loads of the row in flight per lane before the first use (the chain order per lane is unchanged:
  // elements in increasing c)
  constexpr int U = F <= 2 ? 8 : 1;   // batch variant (F = 8): registers go to the accumulators instead
  for (int c0 = lane * 4; c0 < ld; c0 += 128 * U) {
    float4 xv4[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + 128 * u;
      xv4[u] = c < ld ? load4<T>(xrow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + 128 * u;
      if (c < ld) {
        const float4 xv = xv4[u];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float4 qv = *reinterpret_cast<const float4*>(q + (size_t)f * q_stride + c);
          if (METRIC == kMetricIP) {
            acc[f] = fmaf(xv.x, qv.x, acc[f]);
            acc[f] = fmaf(xv.y, qv.y, acc[f]);
            acc[f] = fmaf(xv.z, qv.z, acc[f]);
            acc[f] = fmaf(xv.w, qv.w, acc[f]);
          } else {
            float d0 = xv.x - qv.x, d1 = xv.y - qv.y, d2 = xv.z - qv.z, d3 = xv.w - qv.w;
            acc[f] = fmaf(d0, d0, acc[f]);
            acc[f] = fmaf(d1, d1, acc[f]);
            acc[f] = fmaf(d2, d2, acc[f]);
            acc[f] = fmaf(d3, d3, acc[f]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int f = 0; f < F; ++f) out[f] = warp_sum_xor(acc[f]);
}

// What the tensor-core filter does to an operand: kind::tf32 ignores the low 13 mantissa bits (truncation),
// the bf16 paths round to nearest even.  filter_kind: 0 = operand used as stored, 1 = tf32, 2 = bf16.
__device__ __forceinline__ float filter_view(float v, int filter_kind) {
  if (filter_kind == 1) return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  if (filter_kind == 2) return __bfloat162float(__float2bfloat16_rn(v));
  return v;
}

// ---- add path ---------------------------------------------------------------------------------
// src fp32 [n, d] -> dst T [n, ld] (zero padded), plus |x|^2 per row, the running max norm
// (max_norm2_ord[0]) and the running max of |x - filter_view(x)|^2 (max_norm2_ord[1]): the certificate's
// bound on what the filter can get wrong (dense_exact.cuh: rescore_finalize_kernel).
template <typename T>
__global__ void convert_pad_norm_kernel(const float* __restrict__ src, int64_t n, int d, T* __restrict__ dst,
                                        int ld, float* __restrict__ norms, unsigned int* __restrict__ max_norm2_ord,
                                        int filter_kind) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += nwarps) {
    const float* s = src + r * (int64_t)d;
    T* o = dst + r * (int64_t)ld;
    float acc = 0.f, res = 0.f;
    for (int c = lane; c < ld; c += 32) {
      float v = (c < d) ? s[c] : 0.f;
      T t = (T)v;
      o[c] = t;
      float back = (float)t;
      acc = fmaf(back, back, acc);
      const float dv = back - filter_view(back, filter_kind);
      res = fmaf(dv, dv, res);
    }
    acc = warp_sum_xor(acc);
    res = warp_sum_xor(res);
    if (lane == 0) {
      norms[r] = acc;
      atomicMax(max_norm2_ord, f2ord(acc));
      atomicMax(max_norm2_ord + 1, f2ord(res));
    }
  }
}

// fp32 rows (already padded) -> bf16 shadow rows for the tensor-core filter (HR_STORAGE_F32_SHADOW16)
__global__ void shadow_bf16_kernel(const float* __restrict__ src, int64_t n_elems, __nv_bfloat16* __restrict__ dst) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  for (; i < n_elems; i += (int64_t)gridDim.x * blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i) = o;
  }
}

// queries fp32 [nq, d] -> padded fp32 [nq, ld] (+ optional bf16 copy for the bf16 filter)
// rows [nq, nq_rows) are written as zeros: the TMA box of a small batch then still reads >= 32 distinct
// cache lines per K block instead of every SM hammering the same line of a 1-row query matrix.
__global__ void pad_queries_kernel(const float* __restrict__ q, int64_t nq, int64_t nq_rows, int d, int ld,
                                   float* __restrict__ qp, __nv_bfloat16* __restrict__ qh) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = nq_rows * (int64_t)ld;
  for (; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / ld;
    int c = (int)(i - r * ld);
    float v = (r < nq && c < d) ? q[r * (int64_t)d + c] : 0.f;
    qp[i] = v;
    if (qh) qh[i] = __float2bfloat16_rn(v);
  }
}

template <typename T>
__global__ void reconstruct_kernel(const T* __restrict__ x, int ld, int d, int64_t i0, int64_t n,
                                   float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n * (int64_t)d;
  for (; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / d;
    int c = (int)(i - r * d);
    out[i] = (float)x[(i0 + r) * (int64_t)ld + c];
  }
}

// ---- exact exhaustive scan ----------------------------------------------------------------------
// Each warp owns a private top-k list (replace-min) per selected query; a global per-query key
// threshold (tau_g, atomicMax) lets every warp skip rows that already lost somewhere else.
constexpr int kExactF = 8;  // queries scored per corpus pass

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

__device__ __forceinline__ void warp_list_insert(volatile uint64_t* list, int k, int& cnt, uint64_t& minkey,
                                                 uint64_t key, int lane) {
  if (cnt < k) {
    if (lane == 0) list[cnt] = key;
    cnt++;
  } else {
    for (int i = lane; i < k; i += 32)
      if (list[i] == minkey) list[i] = key;
  }
  __syncwarp();
  if (cnt == k) {
    uint64_t m = ~0ull;
    for (int i = lane; i < k; i += 32) {
      uint64_t v = list[i];
      m = v < m ? v : m;
    }
    minkey = warp_min_u64(m);
  }
}

// F = queries scored per corpus pass: 8 for a batch (HR_MODE_EXACT_SIMT), 2 when only one or two queries
// need the fallback (a pass is then bound by HBM, not by shared-memory reads of six idle query slots).
template <typename T, int METRIC, int F>
__global__ void __launch_bounds__(256)
exact_scan_kernel(const T* __restrict__ x, int64_t N, int ld, const float* __restrict__ qpad,
                  const int* __restrict__ qsel, int nsel, int k, uint64_t* __restrict__ lists,
                  int* __restrict__ cnts, unsigned long long* __restrict__ tau_g,
                  const int* __restrict__ nsel_dev, int nsel_lo, int nsel_hi) {
  extern __shared__ __align__(16) float qs[];  // [F][ld]
  // device-driven launch (the certificate's fallback, enqueued without a host round-trip): the number of
  // selected queries is read here; the kernel runs only when it lies in [nsel_lo, nsel_hi] (nsel = capacity)
  if (nsel_dev) {
    const int nd = *nsel_dev;
    if (nd < nsel_lo || nd > nsel_hi) return;
    nsel = min(nd, nsel);
  }
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int64_t W = (int64_t)gridDim.x * wpb;
  const int64_t gw = (int64_t)blockIdx.x * wpb + wib;

  for (int g0 = 0; g0 < nsel; g0 += F) {
    const int nf = min(F, nsel - g0);
    __syncthreads();
    for (int i = threadIdx.x; i < F * ld; i += blockDim.x) {
      int f = i / ld, c = i - f * ld;
      qs[i] = (f < nf) ? qpad[(int64_t)qsel[g0 + f] * ld + c] : 0.f;
    }
    __syncthreads();

    int cnt[F];
    uint64_t minkey[F], tg[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      cnt[f] = 0;
      minkey[f] = 0;
      tg[f] = 0;
    }
    int it = 0;
    for (int64_t r = gw; r < N; r += W, ++it) {
      if ((it & 15) == 0) {
#pragma unroll
        for (int f = 0; f < F; ++f)
          if (f < nf) tg[f] = *((volatile unsigned long long*)&tau_g[g0 + f]);
      }
      float sc[F];
      warp_exact_scores<T, METRIC, F>(x + r * (int64_t)ld, qs, ld, ld, lane, sc);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        if (f < nf) {
          float s = (METRIC == kMetricIP) ? sc[f] : -sc[f];
          uint64_t key = make_key(s, (uint32_t)r);
          uint64_t thr = (cnt[f] == k) ? (minkey[f] > tg[f] ? minkey[f] : tg[f]) : tg[f];
          if (s == s && key > thr) {   // a NaN score is never a candidate (faiss' heap test is false for NaN)
            volatile uint64_t* lst = lists + ((int64_t)(g0 + f) * W + gw) * k;
            warp_list_insert(lst, k, cnt[f], minkey[f], key, lane);
            if (cnt[f] == k && minkey[f] > tg[f]) {
              if (lane == 0) atomicMax(&tau_g[g0 + f], (unsigned long long)minkey[f]);
              tg[f] = minkey[f];
            }
          }
        }
      }
    }
#pragma unroll
    for (int f = 0; f < F; ++f)
      if (f < nf && lane == 0) cnts[(int64_t)(g0 + f) * W + gw] = cnt[f];
  }
}

// One block per selected query: gather the warps' lists, keep keys >= tau_g, sort, write top-k.
constexpr int kMergeCap = 4096;

template <int METRIC>
__global__ void __launch_bounds__(256)
exact_merge_kernel(const uint64_t* __restrict__ lists, const int* __restrict__ cnts,
                   const unsigned long long* __restrict__ tau_g, const int* __restrict__ qsel, int64_t W, int k,
                   int64_t id_base, float* __restrict__ D, int64_t* __restrict__ I,
                   const int* __restrict__ nsel_dev) {
  __shared__ uint64_t buf[kMergeCap];
  __shared__ int s_n;
  __shared__ unsigned long long s_best;
  const int f = blockIdx.x;
  if (nsel_dev && f >= *nsel_dev) return;   // device-driven launch with the capacity as grid
  const int q = qsel[f];
  const uint64_t thr = tau_g[f];
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const int64_t total = W * k;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    int64_t w = i / k;
    int j = (int)(i - w * k);
    if (j < cnts[(int64_t)f * W + w]) {
      uint64_t key = lists[((int64_t)f * W + w) * k + j];
      if (key >= thr) {
        int slot = atomicAdd(&s_n, 1);
        if (slot < kMergeCap) buf[slot] = key;
      }
    }
  }
  __syncthreads();
  const int n = s_n;
  float* Dq = D + (int64_t)q * k;
  int64_t* Iq = I + (int64_t)q * k;
  const float pad = (METRIC == kMetricIP) ? HR_NEG_INF : -HR_NEG_INF;
  if (n <= kMergeCap) {
    int p = 1;
    while (p < n) p <<= 1;
    for (int i = n + threadIdx.x; i < p; i += blockDim.x) buf[i] = 0;
    block_bitonic_desc(buf, p);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      if (j < n) {
        float s = key_score(buf[j]);
        Dq[j] = (METRIC == kMetricIP) ? s : -s;
        Iq[j] = (int64_t)key_row(buf[j]) + id_base;
      } else {
        Dq[j] = pad;
        Iq[j] = -1;
      }
    }
  } else {
    // rare: more surviving keys than the sort buffer holds -> k rounds of "largest key below the last"
    unsigned long long last = ~0ull;
    for (int j = 0; j < k; ++j) {
      if (threadIdx.x == 0) s_best = 0;
      __syncthreads();
      unsigned long long best = 0;
      for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
        int64_t w = i / k;
        int jj = (int)(i - w * k);
        if (jj < cnts[(int64_t)f * W + w]) {
          unsigned long long key = lists[((int64_t)f * W + w) * k + jj];
          if (key < last && key > best) best = key;
        }
      }
      atomicMax(&s_best, best);
      __syncthreads();
      unsigned long long b = s_best;
      if (threadIdx.x == 0) {
        if (b != 0) {
          float s = key_score(b);
          Dq[j] = (METRIC == kMetricIP) ? s : -s;
          Iq[j] = (int64_t)key_row(b) + id_base;
        } else {
          Dq[j] = pad;
          Iq[j] = -1;
        }
      }
      last = b ? b : 0;
      __syncthreads();
    }
  }
}

// ---- exact re-score of the filter shortlist + final top-k + certificate ---------------------------
// Two stages.  Stage 1 (qsel == nullptr, one block per query): the best short_n <= KL candidates.  A query
// whose certificate fails while more candidates exist (short_tot > short_n) goes to the `deeper` list; stage 2
// (qsel = that list, KL = row_stride) re-scores ALL its candidates against the bound of the thresholds alone.
// Only what still fails is flagged for the exhaustive exact scan.
// One block (256 threads) per query.  short_rows [nq][row_stride] (0xFFFFFFFF = empty), short_n [nq],
// tprime [nq] = upper bound on the APPROXIMATE score of every row that is not in the shortlist
// (HR_NEG_INF when nothing was dropped).  A query is certified when its exact k-th best, mapped to
// the filter's score domain, beats tprime by more than the filter's worst-case error eps.
// Block size: 256 threads, or 1024 for small batches (a warp re-scores one row at a time: with few blocks in
// flight the rows of a query are a latency chain, and 32 warps make it 4x shorter).
template <typename T, int METRIC>
__global__ void __launch_bounds__(1024)
rescore_finalize_kernel(const T* __restrict__ x, int ld, const float* __restrict__ qpad,
                        const uint32_t* __restrict__ short_rows, int row_stride, const int* __restrict__ short_n,
                        const float* __restrict__ tprime, int KL, int k, float c_acc, int filter_kind,
                        const unsigned int* __restrict__ max_norm2_ord, int64_t id_base, float* __restrict__ D,
                        int64_t* __restrict__ I, int* __restrict__ flagged, int* __restrict__ nflag,
                        const int* __restrict__ qsel, const int* __restrict__ short_tot, int* __restrict__ deeper,
                        int* __restrict__ ndeeper, const float* __restrict__ short_s,
                        const int* __restrict__ nsel_dev) {
  extern __shared__ uint64_t skeys[];  // [KL]
  if (nsel_dev && (int)blockIdx.x >= *nsel_dev) return;   // stage 2 is launched for every query, runs for the listed ones
  __shared__ float s_qn2, s_dq2, s_qt2;
  __shared__ float s_ek;
  __shared__ int s_have_k;
  __shared__ int s_need;
  const int q = qsel ? qsel[blockIdx.x] : blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarp = blockDim.x >> 5;
  const int n = min(short_n[q], KL);
  const float* qv = qpad + (int64_t)q * ld;
  if (threadIdx.x == 0) s_have_k = 0;
  if (warp == 0) {
    // |q|^2, |q - q~|^2 and |q~|^2 with q~ = the query as the filter saw it
    float a = 0.f, dq = 0.f, qt = 0.f;
    for (int c = lane; c < ld; c += 32) {
      const float v = qv[c], f = filter_view(v, filter_kind);
      a = fmaf(v, v, a);
      dq = fmaf(v - f, v - f, dq);
      qt = fmaf(f, f, qt);
    }
    a = warp_sum_xor(a);
    dq = warp_sum_xor(dq);
    qt = warp_sum_xor(qt);
    if (lane == 0) {
      s_qn2 = a;
      s_dq2 = dq;
      s_qt2 = qt;
    }
  }
  __syncthreads();
  // the filter's error bound (see the certificate below)
  const float xn = sqrtf(ord2f(max_norm2_ord[0]));
  const float dxn = sqrtf(fmaxf(ord2f(max_norm2_ord[1]), 0.f));
  const float eps = 1.0001f * (sqrtf(s_dq2) * xn + sqrtf(s_qt2) * dxn) + c_acc * sqrtf(s_qn2) * xn +
                    1e-6f * (s_qn2 + xn * xn) + 1e-30f;
  // Only rows whose filter score is within 2*eps of the k-th filter score can be in the exact top k: the k rows
  // in front have exact scores >= s_k - eps, a row behind s_k - 2*eps has an exact score < s_k - eps.  The
  // candidates are sorted by filter score, so those rows are a prefix.
  if (threadIdx.x == 0) s_need = n;
  __syncthreads();
  if (n > k && short_s) {
    const float* ss = short_s + (int64_t)q * row_stride;
    const float cut = ss[k - 1] - 2.f * eps;
    for (int i = k + threadIdx.x; i < n; i += blockDim.x)
      if (ss[i] < cut && ss[i - 1] >= cut) s_need = i;   // sorted descending: exactly one boundary (or none)
    __syncthreads();
  }
  const int nn = s_need;
  for (int i = warp; i < nn; i += nwarp) {
    uint32_t row = short_rows[(int64_t)q * row_stride + i];
    float sc[1];
    warp_exact_scores<T, METRIC, 1>(x + (int64_t)row * ld, qv, ld, ld, lane, sc);
    if (lane == 0) skeys[i] = make_key(METRIC == kMetricIP ? sc[0] : -sc[0], row);
  }
  __syncthreads();
  float* Dq = D + (int64_t)q * k;
  int64_t* Iq = I + (int64_t)q * k;
  const float pad = (METRIC == kMetricIP) ? HR_NEG_INF : -HR_NEG_INF;
  for (int i = threadIdx.x; i < nn; i += blockDim.x) {
    uint64_t me = skeys[i];
    int rank = 0;
    for (int j = 0; j < nn; ++j) rank += (skeys[j] > me);
    if (rank < k) {
      float s = key_score(me);
      Dq[rank] = (METRIC == kMetricIP) ? s : -s;
      Iq[rank] = (int64_t)key_row(me) + id_base;
      if (rank == k - 1) {
        s_ek = s;
        s_have_k = 1;
      }
    }
  }
  for (int j = nn + threadIdx.x; j < k; j += blockDim.x) {
    Dq[j] = pad;
    Iq[j] = -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float tp = tprime[q];
    bool ok;
    if (tp <= HR_NEG_INF) {
      ok = true;  // nothing was ever dropped: the shortlist is the whole index
    } else if (!s_have_k) {
      ok = false;  // rows were dropped but fewer than k survived: cannot certify
    } else {
      // eps: |q.x - q~.x~| = |dq.x + q~.dx| <= |dq| max|x| + |q~| max|dx| (Cauchy-Schwarz; dq, dx are the operands'
      // actual rounding residuals, measured, not their worst case), plus the fp32 accumulation of d terms
      // map the exact k-th best into the filter's score domain
      float shat = (METRIC == kMetricIP) ? s_ek : 0.5f * (s_qn2 + s_ek);  // s_ek = -dist for L2
      ok = shat > tp + eps;
    }
    if (!ok) {
      if (deeper && short_tot[q] > n) {
        int slot = atomicAdd(ndeeper, 1);
        deeper[slot] = q;
      } else {
        int slot = atomicAdd(nflag, 1);
        flagged[slot] = q;
      }
    }
  }
}

}  // namespace hr
