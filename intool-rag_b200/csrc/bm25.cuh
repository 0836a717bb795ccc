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
constexpr int kBwCons = kBwWarps - 1;   // consumer warps; the last warp plans the next window
constexpr int kBwSlice = 17 * 128;       // docs per consumer warp in the clear / sweep phases
constexpr int kBwWin = kBwCons * kBwSlice;   // 15232 docs per window (59.5 KB of accumulators: three CTAs per SM)
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
// consumer warp g % 7.  Slots are ordered by term; `need` counts the term boundaries before the slot inside its
// round: a warp passes that many consumer barriers before it applies the slot, so postings of different
// terms never race on a document (inside one term doc ids are unique).
struct __align__(16) BwSlot {
  uint32_t p_lo, p_hi;   // global posting index of the slot's first posting (multiple of 4)
  uint32_t meta;         // f | e << 8 | need << 16
  float w;               // term weight (multiplicity * idf)
};

// Per-term state of a query, kept in the registers of warp 0: lane l owns terms l and l + 32.
struct BwTerms {
  int64_t start[2];        // first posting of the list
  float wgt[2];
  uint32_t b0[2], b1[2], b2[2];   // cursors at the boundaries of window k, k+1, k+2 (k = current window)
};

// Warp 0: describe slots [r0, r0 + kBwSlotCap) of the window with cursors [lo, hi) per term (r0 < number of
// slots, or the window is empty).  Slot numbers come from warp prefix sums.  meta[0] = slots in the window,
// meta[1] = barriers the round needs in total.  With prefetch != 0 the posting ranges are also requested
// into L2 (one bulk prefetch per array and term).
__device__ __forceinline__ void bw_build_slots(int lane, const BwTerms& T, const uint32_t (&lo)[2], const uint32_t (&hi)[2],
                                               int r0, BwSlot* slots, int* meta, const int32_t* post_doc,
                                               const float* post_imp, bool prefetch) {
  int64_t a[2];
  uint32_t n[2];
  int cnt[2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    n[half] = hi[half] - lo[half];
    a[half] = T.start[half] + lo[half];
    cnt[half] = n[half] ? (int)((a[half] + n[half] - (a[half] & ~(int64_t)3) + kBwSlotLen - 1) / kBwSlotLen) : 0;
    if (prefetch && n[half]) {
      const int64_t s0 = a[half] & ~(int64_t)3;
      const uint32_t bytes = (uint32_t)(((a[half] + n[half] - s0 + 3) & ~(int64_t)3) * 4);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(post_doc + s0), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(post_imp + s0), "r"(bytes) : "memory");
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
    meta[0] = nslots;
    if (nslots == 0) meta[1] = 0;
  }
  // slot-parallel: lane takes slots r0 + lane, r0 + lane + 32, ...; the owning term is found by a binary
  // search over the lanes' inclusive slot prefix sums (shuffles), then its data is fetched from that lane
  const int incl[2] = {iA, totA + iB};
  for (int g0 = r0; g0 <= last; g0 += 32) {
    const int gg = g0 + lane;
    const int half = gg >= totA ? 1 : 0;
    int lo_l = 0, hi_l = 31;   // smallest lane whose inclusive prefix exceeds gg
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int mid = (lo_l + hi_l) >> 1;
      const int v0 = __shfl_sync(0xffffffffu, incl[0], mid), v1 = __shfl_sync(0xffffffffu, incl[1], mid);
      if ((half ? v1 : v0) > gg) hi_l = mid; else lo_l = mid + 1;
    }
    const int src = lo_l;
    const int64_t a0 = __shfl_sync(0xffffffffu, a[0], src), a1 = __shfl_sync(0xffffffffu, a[1], src);
    const uint32_t n0 = __shfl_sync(0xffffffffu, n[0], src), n1 = __shfl_sync(0xffffffffu, n[1], src);
    const int o0 = __shfl_sync(0xffffffffu, off[0], src), o1 = __shfl_sync(0xffffffffu, off[1], src);
    const int r_0 = __shfl_sync(0xffffffffu, rank[0], src), r_1 = __shfl_sync(0xffffffffu, rank[1], src);
    const float w0 = __shfl_sync(0xffffffffu, T.wgt[0], src), w1 = __shfl_sync(0xffffffffu, T.wgt[1], src);
    if (gg <= last) {
      const int64_t at = half ? a1 : a0;
      const int64_t s0 = at & ~(int64_t)3;
      const int64_t end = at + (half ? n1 : n0);
      const int sl = gg - (half ? o1 : o0);
      const int64_t ps = s0 + (int64_t)sl * kBwSlotLen;
      const uint32_t need = (uint32_t)max((half ? r_1 : r_0) - rank0, 0);
      const uint32_t f = sl == 0 ? (uint32_t)(at - s0) : 0u;
      const uint32_t e = (uint32_t)min((int64_t)kBwSlotLen, end - ps);
      BwSlot d;
      d.p_lo = (uint32_t)ps;
      d.p_hi = (uint32_t)(ps >> 32);
      d.meta = f | (e << 8) | (need << 16);
      d.w = half ? w1 : w0;
      slots[gg - r0] = d;
      if (gg == last) meta[1] = (int)need;
    }
  }
}

// barrier 1: the consumer warps only (the planner warp never takes part in a term boundary)
__device__ __forceinline__ void bw_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kBwCons * 32) : "memory"); }

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
                   uint64_t* __restrict__ out_keys, int* __restrict__ out_n, unsigned long long* __restrict__ tau_g, int flags) {
  extern __shared__ __align__(16) uint8_t bsm[];
  float* acc = (float*)bsm;
  uint64_t* cb_all = (uint64_t*)(bsm + kBwWin * 4);
  __shared__ BwSlot s_slots[2][kBwSlotCap];   // double buffered: warp 0 describes window k+1 during window k
  __shared__ int s_meta[2][2];
  __shared__ uint16_t s_hot[kBwHotCap];
  __shared__ unsigned long long s_tau;
  __shared__ unsigned int s_nhot[2];
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
  BwTerms T;
  uint32_t nxt[2] = {0, 0};   // cursors at boundary k+3, in flight during window k
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int term = lane + 32 * half;
    T.start[half] = 0;
    T.wgt[half] = 0.f;
    T.b0[half] = T.b1[half] = T.b2[half] = 0;
    if (w == kBwCons && term < nt) {
      T.start[half] = plan_start[qa + term];
      T.wgt[half] = plan_wgt[qa + term];
      T.b0[half] = curq[(size_t)win0 * nt + term];
      T.b1[half] = curq[(size_t)(win0 + 1) * nt + term];
      T.b2[half] = win0 + 2 <= nwin ? curq[(size_t)(win0 + 2) * nt + term] : T.b1[half];
    }
  }
  if (tid == 0) {
    s_tau = *((volatile unsigned long long*)tau_gq);
    s_nhot[0] = 0;
    s_nhot[1] = 0;
  }
  for (int i = tid * 4; i < kBwWin; i += kBwThreads * 4) *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool planner = w == kBwCons;
  if (planner) bw_build_slots(lane, T, T.b0, T.b1, 0, s_slots[win0 & 1], s_meta[win0 & 1], post_doc, post_imp, false);
  __syncthreads();

  int cbn = 0;                 // keys in this warp's buffer (warp-uniform)
  unsigned long long tau = 0;  // this warp's threshold key: a lower bound of the query's kc-th best
  float tau_f = 0.f;
  float* slice = acc + (planner ? 0 : w) * kBwSlice;
  // one batch of slots per warp in registers
  int4 dd[kBwBatch];
  float4 vv[kBwBatch];
  uint32_t meta[kBwBatch];
  float ww[kBwBatch];
  auto load_batch = [&](const BwSlot* tab, int nr, int b0) {
#pragma unroll
    for (int j = 0; j < kBwBatch; ++j) {
      const int gr = b0 + w + kBwCons * j;
      meta[j] = 0;
      ww[j] = 0.f;
      dd[j] = make_int4(0, 0, 0, 0);
      vv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < nr) {
        const BwSlot e = tab[gr];
        meta[j] = e.meta;
        ww[j] = e.w;
        if ((uint32_t)(4 * lane) < ((e.meta >> 8) & 0xFFu)) {
          const int64_t p = (int64_t)(((uint64_t)e.p_hi << 32) | e.p_lo) + 4 * lane;
          dd[j] = __ldg(reinterpret_cast<const int4*>(post_doc + p));
          vv[j] = __ldg(reinterpret_cast<const float4*>(post_imp + p));
        }
      }
    }
  };
  if (!planner) load_batch(s_slots[win0 & 1], min(kBwSlotCap, s_meta[win0 & 1][0]), 0);
  for (int64_t win = win0; win < win1; ++win) {
    const int32_t docbase = (int32_t)(win * kBwWin);
    const int buf = (int)(win & 1);
    unsigned int* nhot_p = &s_nhot[buf];
    unsigned long long gt = 0;
    if (planner) {
      if (lane == 0) gt = *((volatile unsigned long long*)tau_gq);
      // cursors three boundaries ahead; slots of the next window (its postings go to L2 meanwhile)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int term = lane + 32 * half;
        nxt[half] = T.b2[half];
        if (term < nt && win + 3 <= nwin) nxt[half] = __ldg(curq + (size_t)(win + 3) * nt + term);
      }
      if (win + 1 < win1)
        bw_build_slots(lane, T, T.b1, T.b2, 0, s_slots[buf ^ 1], s_meta[buf ^ 1], post_doc, post_imp, !(flags & 1));
    }
    // ---- apply: slot g of a round belongs to warp g % 8; kBwBatch slots in flight per warp ----
    const float tau_pos = tau_f > 0.f ? tau_f : 1.4e-45f;   // x >= tau_pos <=> x > 0 && x >= tau_f
    const int nslots = s_meta[buf][0];
    for (int r0 = 0; r0 < nslots; r0 += kBwSlotCap) {
      if (r0 > 0) {   // more slots than the staging table holds: describe the next round
        __syncthreads();
        if (planner) bw_build_slots(lane, T, T.b0, T.b1, r0, s_slots[buf], s_meta[buf], post_doc, post_imp, false);
        __syncthreads();
      }
      const int nr = min(kBwSlotCap, nslots - r0);
      const int total_need = s_meta[buf][1];
      if (planner) continue;
      int done = 0;   // barriers this warp has passed in this round
      for (int b0 = 0; b0 < nr; b0 += kBwCons * kBwBatch) {
        if (r0 + b0 > 0 || ((flags & 2) && win > win0)) load_batch(s_slots[buf], nr, b0);   // the window's first batch is already in flight
#pragma unroll
        for (int j = 0; j < kBwBatch; ++j) {
          const int gr = b0 + w + kBwCons * j;
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
            bool valid[4];
            bool hot = false;
            // the four docs of a lane are distinct (one posting list): read all, then add, then write
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const uint32_t idx = 4u * lane + e4;
              valid[e4] = idx >= f && idx < e;
              offs[e4] = valid[e4] ? dv[e4] - docbase : 0;
              xs[e4] = acc[offs[e4]];
            }
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              xs[e4] = valid[e4] ? fmaf(ww[j], iv[e4], xs[e4]) : 0.f;
              if (valid[e4]) acc[offs[e4]] = xs[e4];
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
    if (planner && lane == 0 && gt > s_tau) s_tau = gt;   // elsewhere s_tau only changes by atomicMax after the next barrier
    __syncthreads();
    // the next window's first batch of postings starts its trip now (its slots were described during this window)
    if (!planner && win + 1 < win1 && !(flags & 2)) load_batch(s_slots[buf ^ 1], min(kBwSlotCap, s_meta[buf ^ 1][0]), 0);
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
      for (int j = lane * 4; !planner && j < kBwSlice; j += 128) {
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
        for (unsigned i0 = 0; !planner && i0 < nhot; i0 += kBwCons * 32) {
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
      if (!planner) {
#pragma unroll
        for (int j = lane * 4; j < kBwSlice; j += 128) *reinterpret_cast<float4*>(slice + j) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (planner) {   // advance the cursors by one window
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        T.b0[half] = T.b1[half];
        T.b1[half] = T.b2[half];
        T.b2[half] = nxt[half];
      }
      if (lane == 0) s_nhot[buf ^ 1] = 0;
    }
    __syncthreads();   // clears visible before the next window's first slot is applied
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
