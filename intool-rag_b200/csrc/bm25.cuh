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
//                            are cut into slots of 128 postings (one warp, 16-byte loads, 4 postings
//                            per lane); a warp loads kBwBatch slots at once (all loads in flight
//                            before the first use), then applies them with plain shared-memory
//                            read-modify-writes: doc ids inside one posting list are unique, and a
//                            block barrier separates different terms (each slot carries the number of
//                            barriers a warp must have passed before applying it), so no atomics are
//                            needed and the fp32 summation order is fixed (term order): results are
//                            deterministic.  A doc whose running score reaches the running threshold
//                            is pushed to a small hot list; at the end of a window only the listed
//                            docs are turned into candidates (warp-private key buffers, bitonic
//                            compaction) and the accumulators are cleared with one vectorised sweep.
#pragma once
#include "common.cuh"
#include "dense_exact.cuh"

namespace hr {

constexpr int kBmMaxTerms = 64;   // raw terms per query
constexpr int kBmMaxK = 128;      // candidate depth the kernel supports
constexpr int kBwThreads = 256;
constexpr int kBwWarps = kBwThreads / 32;
constexpr int kBwWin = 15360;     // docs per window (60 KB of accumulators: three CTAs per SM)
constexpr int kBwSlice = kBwWin / kBwWarps;
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
constexpr int kBwSlotCap = 128;   // slot descriptors staged per round (2 KB)
constexpr int kBwSlotLen = 128;   // postings per slot: one warp, four consecutive postings per lane (16-byte loads)
constexpr int kBwBatch = 4;       // slots a warp keeps in flight (4 x 2 x 512 B)
constexpr int kBwHotCap = 1024;   // docs that may reach the threshold in one window before the full sweep takes over

// One slot = kBwSlotLen consecutive postings of one term, starting at a 4-aligned global posting index;
// [f, e) of them lie inside the current window's range of that term.  Slot g of a round is applied by
// warp g % 8.  Slots are ordered by term; `need` counts the term boundaries before the slot inside its
// round: a warp passes that many block barriers before it applies the slot, so postings of different
// terms never race on a document (inside one term doc ids are unique).
struct __align__(16) BwSlot {
  uint32_t p_lo, p_hi;   // global posting index of the slot's first posting (multiple of 4)
  uint32_t meta;         // f | e << 8 | need << 16
  float w;               // term weight (multiplicity * idf)
};

// Warp 0: describe slots [r0, r0 + kBwSlotCap) of the window whose cursors are in s_lo/s_hi (r0 < number of
// slots, or the window is empty).  Lane l owns terms l and l + 32; slot numbers come from warp prefix sums.
// s_meta[0] = slots in the window, s_meta[1] = barriers the round needs in total.
__device__ __forceinline__ void bw_build_slots(int lane, int nt, const int64_t* s_start, const float* s_w,
                                               const uint32_t* s_lo, const uint32_t* s_hi, int r0, BwSlot* slots,
                                               int* s_meta) {
  int64_t a[2] = {0, 0};
  uint32_t n[2] = {0, 0};
  int cnt[2] = {0, 0};
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int term = lane + 32 * half;
    if (term < nt) {
      n[half] = s_hi[term] - s_lo[term];
      a[half] = s_start[term] + s_lo[term];
      if (n[half]) cnt[half] = (int)((a[half] + n[half] - (a[half] & ~(int64_t)3) + kBwSlotLen - 1) / kBwSlotLen);
    }
  }
  int iA = cnt[0], iB = cnt[1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int va = __shfl_up_sync(0xffffffffu, iA, o), vb = __shfl_up_sync(0xffffffffu, iB, o);
    if (lane >= o) {
      iA += va;
      iB += vb;
    }
  }
  const int totA = __shfl_sync(0xffffffffu, iA, 31), totB = __shfl_sync(0xffffffffu, iB, 31);
  const int nslots = totA + totB;
  const int off[2] = {iA - cnt[0], totA + iB - cnt[1]};
  const unsigned neA = __ballot_sync(0xffffffffu, cnt[0] > 0), neB = __ballot_sync(0xffffffffu, cnt[1] > 0);
  const unsigned lt = (1u << lane) - 1u;
  const int rank[2] = {__popc(neA & lt), __popc(neA) + __popc(neB & lt)};
  // rank of the term that owns slot r0
  const unsigned inA = __ballot_sync(0xffffffffu, cnt[0] > 0 && off[0] <= r0 && r0 < off[0] + cnt[0]);
  const unsigned inB = __ballot_sync(0xffffffffu, cnt[1] > 0 && off[1] <= r0 && r0 < off[1] + cnt[1]);
  int rank0 = 0;
  if (inA) rank0 = __shfl_sync(0xffffffffu, rank[0], __ffs(inA) - 1);
  else if (inB) rank0 = __shfl_sync(0xffffffffu, rank[1], __ffs(inB) - 1);
  const int last = min(nslots, r0 + kBwSlotCap) - 1;
  if (lane == 0) {
    s_meta[0] = nslots;
    if (nslots == 0) s_meta[1] = 0;
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if (cnt[half] > 0) {
      const int term = lane + 32 * half;
      const int64_t s0 = a[half] & ~(int64_t)3;
      const int64_t end = a[half] + n[half];
      const float wt = s_w[term];
      const uint32_t need = (uint32_t)max(rank[half] - rank0, 0);
      const int sl0 = max(0, r0 - off[half]), sl1 = min(cnt[half], r0 + kBwSlotCap - off[half]);
      for (int sl = sl0; sl < sl1; ++sl) {
        const int64_t ps = s0 + (int64_t)sl * kBwSlotLen;
        const uint32_t f = sl == 0 ? (uint32_t)(a[half] - s0) : 0u;
        const uint32_t e = (uint32_t)min((int64_t)kBwSlotLen, end - ps);
        BwSlot d;
        d.p_lo = (uint32_t)ps;
        d.p_hi = (uint32_t)(ps >> 32);
        d.meta = f | (e << 8) | (need << 16);
        d.w = wt;
        slots[off[half] + sl - r0] = d;
        if (off[half] + sl == last) s_meta[1] = (int)need;
      }
    }
  }
}

__device__ __forceinline__ void bw_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// Warp-collective append of a candidate key to the warp's key buffer (compaction by bitonic sort keeps the
// best kc and raises the warp / CTA / query thresholds).
__device__ __forceinline__ void bw_append(bool take, unsigned long long key, uint64_t* cb, int& cbn, int cbcap, int kc,
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
  __shared__ BwSlot s_slots[kBwSlotCap];
  __shared__ int64_t s_start[kBmMaxTerms];
  __shared__ float s_w[kBmMaxTerms];
  __shared__ uint32_t s_lo[kBmMaxTerms], s_hi[kBmMaxTerms];
  __shared__ uint16_t s_hot[kBwHotCap];
  __shared__ unsigned long long s_tau;
  __shared__ unsigned int s_nhot[2];
  __shared__ int s_meta[2];
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
  unsigned long long* tau_gq = tau_g + q;
  if (tid < nt) {
    s_start[tid] = plan_start[qa + tid];
    s_w[tid] = plan_wgt[qa + tid];
    s_lo[tid] = curq[(size_t)win0 * nt + tid];
    s_hi[tid] = curq[(size_t)(win0 + 1) * nt + tid];
  }
  if (tid == 0) {
    s_tau = *((volatile unsigned long long*)tau_gq);
    s_nhot[0] = 0;
    s_nhot[1] = 0;
  }
  for (int i = tid * 4; i < kBwWin; i += kBwThreads * 4) *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  if (w == 0) bw_build_slots(lane, nt, s_start, s_w, s_lo, s_hi, 0, s_slots, s_meta);
  __syncthreads();

  int cbn = 0;                 // keys in this warp's buffer (warp-uniform)
  unsigned long long tau = 0;  // this warp's threshold key: a lower bound of the query's kc-th best
  float tau_f = 0.f;
  float* slice = acc + w * kBwSlice;
  for (int64_t win = win0; win < win1; ++win) {
    const int32_t docbase = (int32_t)(win * kBwWin);
    unsigned int* nhot_p = &s_nhot[win & 1];
    // cursors of the next window (consumed at the end of this one)
    uint32_t nxtA = 0, nxtB = 0;   // warp 0: lane l keeps terms l and l + 32
    if (w == 0 && win + 2 <= nwin) {
      if (lane < nt) nxtA = __ldg(curq + (size_t)(win + 2) * nt + lane);
      if (lane + 32 < nt) nxtB = __ldg(curq + (size_t)(win + 2) * nt + lane + 32);
    }
    unsigned long long gt = 0;
    if (tid == 0) gt = *((volatile unsigned long long*)tau_gq);
    // ---- apply: slot g of a round belongs to warp g % 8; kBwBatch slots in flight per warp ----
    const float tau_pos = tau_f > 0.f ? tau_f : 1.4e-45f;   // x >= tau_pos <=> x > 0 && x >= tau_f
    const int nslots = s_meta[0];
    for (int r0 = 0; r0 < nslots; r0 += kBwSlotCap) {
      if (r0 > 0) {   // more slots than the staging table holds: describe the next round
        __syncthreads();
        if (w == 0) bw_build_slots(lane, nt, s_start, s_w, s_lo, s_hi, r0, s_slots, s_meta);
        __syncthreads();
      }
      const int nr = min(kBwSlotCap, nslots - r0);
      const int total_need = s_meta[1];
      int done = 0;   // barriers this warp has passed in this round
      for (int b0 = 0; b0 < nr; b0 += kBwWarps * kBwBatch) {
        int4 dd[kBwBatch];
        float4 vv[kBwBatch];
        uint32_t meta[kBwBatch];
        float ww[kBwBatch];
#pragma unroll
        for (int j = 0; j < kBwBatch; ++j) {
          const int gr = b0 + w + kBwWarps * j;
          meta[j] = 0;
          ww[j] = 0.f;
          dd[j] = make_int4(0, 0, 0, 0);
          vv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gr < nr) {
            const BwSlot e = s_slots[gr];
            meta[j] = e.meta;
            ww[j] = e.w;
            if ((uint32_t)(4 * lane) < ((e.meta >> 8) & 0xFFu)) {
              const int64_t p = (int64_t)(((uint64_t)e.p_hi << 32) | e.p_lo) + 4 * lane;
              dd[j] = __ldg(reinterpret_cast<const int4*>(post_doc + p));
              vv[j] = __ldg(reinterpret_cast<const float4*>(post_imp + p));
            }
          }
        }
#pragma unroll
        for (int j = 0; j < kBwBatch; ++j) {
          const int gr = b0 + w + kBwWarps * j;
          if (gr < nr) {   // warp-uniform
            const int need = (int)(meta[j] >> 16);
            while (done < need) {
              bw_barrier();
              ++done;
            }
            const uint32_t f = meta[j] & 0xFFu, e = (meta[j] >> 8) & 0xFFu;
            const int dv[4] = {dd[j].x, dd[j].y, dd[j].z, dd[j].w};
            const float iv[4] = {vv[j].x, vv[j].y, vv[j].z, vv[j].w};
            float xs[4];
            int offs[4];
            bool hot = false;
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const uint32_t idx = 4u * lane + e4;
              const bool valid = idx >= f && idx < e;
              offs[e4] = valid ? dv[e4] - docbase : 0;
              xs[e4] = 0.f;
              if (valid) {
                xs[e4] = fmaf(ww[j], iv[e4], acc[offs[e4]]);
                acc[offs[e4]] = xs[e4];
              }
              hot = hot || xs[e4] >= tau_pos;
            }
            // docs whose running score reached the threshold go to the hot list (the thread applying a doc's
            // last posting sees its final score, so every candidate is listed at least once)
            if (__any_sync(0xffffffffu, hot)) {
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const bool h = xs[e4] >= tau_pos;
                const unsigned hm = __ballot_sync(0xffffffffu, h);
                if (hm) {
                  unsigned base = 0;
                  const int leader = __ffs(hm) - 1;
                  if (lane == leader) base = atomicAdd(nhot_p, (unsigned)__popc(hm));
                  base = __shfl_sync(0xffffffffu, base, leader);
                  if (h) {
                    const unsigned pos = base + __popc(hm & ((1u << lane) - 1u));
                    if (pos < kBwHotCap) s_hot[pos] = (uint16_t)offs[e4];
                  }
                }
              }
            }
          }
        }
      }
      while (done < total_need) {   // warps without slots behind the last term boundary catch up
        bw_barrier();
        ++done;
      }
    }
    // ---- end of window ----
    if (tid == 0 && gt > s_tau) s_tau = gt;   // elsewhere s_tau only changes by atomicMax after the next barrier
    __syncthreads();
    const unsigned nhot = *((volatile unsigned int*)nhot_p);
    {
      const unsigned long long ct = *((volatile unsigned long long*)&s_tau);
      if (ct > tau) {
        tau = ct;
        tau_f = key_score(tau);
      }
    }
    if (nhot > kBwHotCap) {
      // cold thresholds: sweep the warp's slice, extract and clear
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
            bw_append(take, key, cb, cbn, cbcap, kc, tau, tau_f, lane, &s_tau, tau_gq);
          }
        }
      }
    } else {
      if (nhot > 0) {
        // the listed docs only: the first reader of a doc takes its score (exchange with 0), duplicates see 0
        for (unsigned i0 = 0; i0 < nhot; i0 += kBwThreads) {
          const unsigned i = i0 + tid;
          unsigned long long key = 0;
          bool take = false;
          if (i < nhot) {
            const int off = s_hot[i];
            const float v = atomicExch(acc + off, 0.f);
            if (v > 0.f && v >= tau_f) {
              key = make_key(v, (uint32_t)(docbase + off));
              take = key > tau;
            }
          }
          bw_append(take, key, cb, cbn, cbcap, kc, tau, tau_f, lane, &s_tau, tau_gq);
        }
        __syncthreads();   // uniform (nhot is): the sweep below must not clear a listed doc before it is read
      }
#pragma unroll
      for (int j = lane * 4; j < kBwSlice; j += 128) *reinterpret_cast<float4*>(slice + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (w == 0 && win + 1 < win1) {   // next window: rotate the cursors, describe its slots, reset its hot counter
      if (lane < nt) {
        s_lo[lane] = s_hi[lane];
        s_hi[lane] = nxtA;
      }
      if (lane + 32 < nt) {
        s_lo[lane + 32] = s_hi[lane + 32];
        s_hi[lane + 32] = nxtB;
      }
      if (lane == 0) s_nhot[(win + 1) & 1] = 0;
      __syncwarp();
      bw_build_slots(lane, nt, s_start, s_w, s_lo, s_hi, 0, s_slots, s_meta);
    }
    __syncthreads();   // clears, cursors and slots visible before the next window
  }
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
