// bm25_sweep.cuh — BM25 scoring, round 2: a flat sweep over the posting ranges of a (query, slice).
//
// Same definition, same plan (bm25.cuh: unique terms + cursor table) and same data layout as the round-1
// slice kernel; what changed is how a warp walks the postings and how the work is scheduled.
//
//  * Flat sweep.  Inside one slice of SLICE docs the posting ranges [c0_t, c1_t) of the query's terms are
//    cut into groups of 4 postings (16-byte aligned in the posting arrays) and the groups of ALL terms are
//    numbered consecutively in term order.  Lane l of iteration i takes group 32 i + l: two 16-byte loads
//    (4 doc ids, 4 impacts), whatever term the group belongs to.  A warp instruction therefore carries up to
//    128 postings of several terms; the round-1 kernel spent one ~100-instruction slot per term.
//    The owner of a group is found without a search: the lanes owning the terms (lane t = term t) mark the
//    first group of their term in a 32-bit mask (one REDUX.OR per iteration), a popc gives every lane the
//    index of its term in a small shared-memory table (posting base, valid range, weight).
//  * Term order is kept.  The accumulators are updated with plain shared-memory read-modify-writes in one
//    pass per term present in the iteration (lanes of other terms idle; doc ids inside one posting list are
//    unique, so a pass has no conflicts).  Every doc's score is the fp32 fma chain over the query's terms in
//    plan order, whatever the slicing: bit-identical between shards and the single index, run to run.
//  * Persistent warps, window-major jobs.  A job is (window g, query q): one warp scores the slices
//    [g spj, (g+1) spj) of one query, alone (no block barrier, no atomics on accumulators).  Jobs are drawn
//    from a global counter in window-major order, so at any time all resident warps work on the same one or
//    two doc windows of different queries: the posting ranges heavy terms share across queries are read from
//    HBM once and hit L2 afterwards.  At the end of a job the warp merges its candidates into the query's
//    running top-k list in global memory (sorted, one lock per query); the list's last key is the exact
//    k-th best of all windows scored so far and becomes the threshold (tau_g) of the query's later windows.
//  * Candidate selection is the round-1 scheme: a doc whose running score reaches the threshold goes to a
//    small hot list; at the end of the slice only the listed docs become candidate keys.
#pragma once
#include "bm25.cuh"

namespace hr {

constexpr int kSwThreads = 256;
constexpr int kSwWarps = kSwThreads / 32;
constexpr int kSwHotCap = 64;

// one term of the current slice in the warp's table: flat coordinate x (4 * group + element) of this sweep maps
// to posting base + x; w = term weight
struct __align__(16) SwTerm {
  uint32_t base_lo, base_hi;
  float w;
  uint32_t pad;
};

// per warp: accumulators (+ one trash slot, padded to 16 bytes) | key buffer | term table | hot list
__host__ __device__ constexpr int sw_warp_bytes(int slice, int kcp) {
  return slice * 4 + 16 + 2 * kcp * 8 + 32 * (int)sizeof(SwTerm) + kSwHotCap * 2;
}
// docs per slice: two CTAs of 8 warps per SM (kSwSliceA/B) or one (kSwSliceWideA/B)
constexpr int kSwSliceA = 24 * 128;       // k_c <= 64
constexpr int kSwSliceB = 21 * 128;       // k_c <= 128
constexpr int kSwSliceWideA = 48 * 128;
constexpr int kSwSliceWideB = 44 * 128;

// warp-collective append of candidate keys (compaction by bitonic sort keeps the best kc and raises the
// warp's and the query's threshold)
__device__ __forceinline__ void sw_append(bool take, unsigned long long key, uint64_t* cb, int& cbn, int cbcap, int kc,
                                          unsigned long long& tau, float& tau_f, int lane,
                                          unsigned long long* tau_gq) {
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (!m) return;
  if (cbn + 32 > cbcap) {
    for (int i = cbn + lane; i < cbcap; i += 32) cb[i] = 0;
    warp_bitonic_desc(cb, cbcap, lane);
    cbn = min(cbn, kc);
    if (cbn == kc) {
      const unsigned long long kth = cb[kc - 1];
      if (kth > tau) {
        tau = kth;
        tau_f = key_score(tau);
        if (lane == 0) atomicMax(tau_gq, tau);
      }
    }
    take = take && key > tau;
  }
  const unsigned m2 = __ballot_sync(0xffffffffu, take);
  if (take) cb[cbn + __popc(m2 & ((1u << lane) - 1u))] = key;
  cbn += __popc(m2);
  __syncwarp();
}

// One slice, up to 32 terms (lane t owns term t: padded list start `pstart` (a multiple of 4), weight `wgt`,
// cursors [c0, c1)).  Posting lists are stored 4-aligned and padded with sentinels (doc = INT_MAX, impact = 0),
// so a group of 4 postings never leaves its list and a posting belongs to the slice iff docbase <= doc <
// docbase + SLICE: no per-element range bookkeeping.
// kSwBatch = iterations (128 postings each) a warp keeps in flight.
template <int SLICE, int kSwBatch>
__device__ __forceinline__ void sw_sweep_terms(int lane, int64_t pstart, float wgt, uint32_t c0, uint32_t c1,
                                               int32_t docbase, const int32_t* __restrict__ post_doc,
                                               const float* __restrict__ post_imp, float* acc, SwTerm* tab,
                                               uint16_t* hotl, int& nhot, float tau_pos) {
  const uint32_t n = c1 - c0, head = c0 & 3u;
  const int ng = n ? (int)((head + n + 3u) >> 2) : 0;
  int incl = ng;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int gtot = __shfl_sync(0xffffffffu, incl, 31);
  if (gtot == 0) return;   // warp-uniform
  const int excl = incl - ng;
  const unsigned present = __ballot_sync(0xffffffffu, ng > 0);
  __syncwarp();            // the previous sweep's table reads are complete
  if (ng > 0) {
    const int64_t base = pstart + (int64_t)(c0 - head) - 4 * (int64_t)excl;
    SwTerm t;
    t.base_lo = (uint32_t)base;
    t.base_hi = (uint32_t)((uint64_t)base >> 32);
    t.w = wgt;
    t.pad = 0;
    tab[__popc(present & ((1u << lane) - 1u))] = t;
  }
  // the iteration and the lane in which this lane's term starts
  const int myit = ng > 0 ? (excl >> 5) : -1;
  const unsigned mybit = 1u << (excl & 31);
  const unsigned le_mask = 0xFFFFFFFFu >> (31 - lane);
  __syncwarp();
  const int niter = (gtot + 31) >> 5;
  int cbefore = 0;   // non-empty terms whose first group lies before the current iteration
  for (int it0 = 0; it0 < niter; it0 += kSwBatch) {
    int4 dd[kSwBatch];
    float4 vv[kSwBatch];
    float ww[kSwBatch];
    int ci[kSwBatch], tfirst[kSwBatch], tlast[kSwBatch];
    // ---- all loads of the batch in flight before the first use ----
#pragma unroll
    for (int j = 0; j < kSwBatch; ++j) {
      const int it = it0 + j;
      if (it < niter) {   // warp-uniform
        const unsigned bits = __reduce_or_sync(0xffffffffu, myit == it ? mybit : 0u);
        const int c = cbefore + __popc(bits & le_mask) - 1;
        tfirst[j] = cbefore + (int)(bits & 1u) - 1;
        cbefore += __popc(bits);
        tlast[j] = cbefore - 1;
        const int g = it * 32 + lane;
        const int gc = min(g, gtot - 1);   // lanes past the end re-read the last group and apply nothing
        const int cc = g < gtot ? c : tlast[j];
        const SwTerm t = tab[cc];
        const int64_t p = (int64_t)(((uint64_t)t.base_hi << 32) | t.base_lo) + 4 * (int64_t)gc;
        dd[j] = __ldg(reinterpret_cast<const int4*>(post_doc + p));
        vv[j] = __ldg(reinterpret_cast<const float4*>(post_imp + p));
        ww[j] = t.w;
        ci[j] = g < gtot ? c : -1;
      }
    }
    // ---- apply: one pass per term present in the iteration, in term order ----
#pragma unroll
    for (int j = 0; j < kSwBatch; ++j) {
      if (it0 + j < niter) {   // warp-uniform
        // a posting outside the slice (the neighbours in a partial first / last group, sentinels) is redirected
        // to the trash slot behind the accumulators with a zero impact: the passes need no predicates
        const uint32_t o0 = min((uint32_t)(dd[j].x - docbase), (uint32_t)SLICE),
                       o1 = min((uint32_t)(dd[j].y - docbase), (uint32_t)SLICE),
                       o2 = min((uint32_t)(dd[j].z - docbase), (uint32_t)SLICE),
                       o3 = min((uint32_t)(dd[j].w - docbase), (uint32_t)SLICE);
        const float i0 = o0 < (uint32_t)SLICE ? vv[j].x : 0.f, i1 = o1 < (uint32_t)SLICE ? vv[j].y : 0.f,
                    i2 = o2 < (uint32_t)SLICE ? vv[j].z : 0.f, i3 = o3 < (uint32_t)SLICE ? vv[j].w : 0.f;
        float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
        for (int t = tfirst[j]; t <= tlast[j]; ++t) {
          __syncwarp();   // a later term may touch docs of an earlier one
          if (ci[j] == t) {
            // the four docs of a lane are distinct (one posting list): read all, then add, then write
            x0 = acc[o0];
            x1 = acc[o1];
            x2 = acc[o2];
            x3 = acc[o3];
            x0 = fmaf(ww[j], i0, x0);
            x1 = fmaf(ww[j], i1, x1);
            x2 = fmaf(ww[j], i2, x2);
            x3 = fmaf(ww[j], i3, x3);
            acc[o0] = x0;
            acc[o1] = x1;
            acc[o2] = x2;
            acc[o3] = x3;
          }
        }
        // docs whose running score reached the threshold go to the hot list (the lane applying a doc's last
        // posting sees its final score, so every candidate is listed at least once)
        const bool hot = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3)) >= tau_pos;
        if (__any_sync(0xffffffffu, hot)) {
          const float xs[4] = {x0, x1, x2, x3};
          const uint32_t offs[4] = {o0, o1, o2, o3};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool h = xs[e] >= tau_pos;   // the trash slot holds 0: never hot
            const unsigned hm = __ballot_sync(0xffffffffu, h);
            if (hm) {
              if (h) {
                const int pos = nhot + __popc(hm & ((1u << lane) - 1u));
                if (pos < kSwHotCap) hotl[pos] = (uint16_t)offs[e];
              }
              nhot += __popc(hm);
            }
          }
        }
      }
    }
  }
}

// use_lock: glist [nq][kc] (sorted best first), gcount [nq], glock [nq]; otherwise glist [nq][S][kc] (unsorted
// slots), gcount [nq][S].  Job j = (window j / nq, query j % nq).
template <int SLICE, int kSwBatch>
__global__ void __launch_bounds__(kSwThreads, (SLICE <= kSwSliceA ? 2 : 1))
bm25_sweep_kernel(const int32_t* __restrict__ post_doc, const float* __restrict__ post_imp,
                  const int32_t* __restrict__ q_indptr, const int* __restrict__ plan_nt,
                  const int64_t* __restrict__ plan_start, const float* __restrict__ plan_wgt,
                  const uint32_t* __restrict__ plan_cur, int64_t nsl, int spj, int S, int nq, int kc, int kcp,
                  uint64_t* glist, int* gcount, int* glock, unsigned long long* __restrict__ tau_g,
                  unsigned int* __restrict__ job_counter, int use_lock) {
  static_assert(SLICE % 128 == 0 && SLICE <= 65536, "hot list entries are 16-bit doc offsets");
  extern __shared__ __align__(16) uint8_t swm[];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  uint8_t* wb = swm + (size_t)w * sw_warp_bytes(SLICE, kcp);
  float* acc = (float*)wb;
  uint64_t* cb = (uint64_t*)(wb + SLICE * 4 + 16);
  SwTerm* tab = (SwTerm*)(wb + SLICE * 4 + 16 + 2 * kcp * 8);
  uint16_t* hotl = (uint16_t*)(wb + SLICE * 4 + 16 + 2 * kcp * 8 + 32 * sizeof(SwTerm));
  const int cbcap = 2 * kcp;
  const unsigned njobs = (unsigned)nq * (unsigned)S;

#pragma unroll
  for (int j = lane * 4; j < SLICE + 4; j += 128) *reinterpret_cast<float4*>(acc + j) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();

  for (;;) {
    unsigned job = 0;
    if (lane == 0) job = atomicAdd(job_counter, 1u);
    job = __shfl_sync(0xffffffffu, job, 0);
    if (job >= njobs) break;
    const int g = (int)(job / (unsigned)nq);
    const int q = (int)(job - (unsigned)g * (unsigned)nq);
    const int nt = plan_nt[q];
    const int64_t s_begin = (int64_t)g * spj;
    const int64_t s_end = min(nsl, s_begin + spj);
    if (nt == 0 || s_begin >= s_end) {
      if (!use_lock && lane == 0) gcount[(size_t)q * S + g] = 0;
      continue;
    }
    const int qa = q_indptr[q];
    const uint32_t* curq = plan_cur + (size_t)qa * (size_t)(nsl + 1);
    unsigned long long* tau_gq = tau_g + q;
    // lane t owns term t (and term t + 32 of a query with more than 32 unique terms: second sweep per slice)
    int64_t start0 = 0;
    float wgt0 = 0.f;
    uint32_t c0 = 0, c1 = 0;
    if (lane < nt) {
      start0 = plan_start[qa + lane];
      wgt0 = plan_wgt[qa + lane];
      c0 = curq[(size_t)s_begin * nt + lane];
      c1 = curq[(size_t)(s_begin + 1) * nt + lane];
    }
    int cbn = 0;
    unsigned long long tau = 0;
    if (lane == 0) tau = *((volatile unsigned long long*)tau_gq);
    tau = __shfl_sync(0xffffffffu, tau, 0);
    float tau_f = tau ? key_score(tau) : 0.f;

    for (int64_t sidx = s_begin; sidx < s_end; ++sidx) {
      const int32_t docbase = (int32_t)(sidx * SLICE);
      uint32_t nxt = c1;
      if (lane < nt && sidx + 2 <= nsl) nxt = __ldg(curq + (size_t)(sidx + 2) * nt + lane);
      unsigned long long gt = 0;
      if ((sidx & 7) == 0 && lane == 0) gt = *((volatile unsigned long long*)tau_gq);
      const float tau_pos = tau_f > 0.f ? tau_f : 1.4e-45f;   // x >= tau_pos <=> x > 0 && x >= tau_f
      int nhot = 0;   // warp-uniform
      sw_sweep_terms<SLICE, kSwBatch>(lane, start0, wgt0, c0, c1, docbase, post_doc, post_imp, acc, tab, hotl, nhot, tau_pos);
      if (nt > 32) {   // rare: terms 32..63, state re-read per slice
        int64_t start1 = 0;
        float wgt1 = 0.f;
        uint32_t d0 = 0, d1 = 0;
        if (lane + 32 < nt) {
          start1 = plan_start[qa + lane + 32];
          wgt1 = plan_wgt[qa + lane + 32];
          d0 = __ldg(curq + (size_t)sidx * nt + lane + 32);
          d1 = __ldg(curq + (size_t)(sidx + 1) * nt + lane + 32);
        }
        sw_sweep_terms<SLICE, kSwBatch>(lane, start1, wgt1, d0, d1, docbase, post_doc, post_imp, acc, tab, hotl, nhot, tau_pos);
      }
      __syncwarp();
      // ---- end of slice ----
      {
        const unsigned long long g0 = __shfl_sync(0xffffffffu, gt, 0);
        if (g0 > tau) {
          tau = g0;
          tau_f = key_score(tau);
        }
      }
      if (nhot > kSwHotCap) {
        // cold threshold: sweep the slice, extract and clear
        if (tau == 0) {
          // No threshold at all yet: appending every scored doc would cost a bitonic compaction per 64 docs.
          // Take each lane's best m = ceil(kc/32) scores first; the kc-th largest of those 32m scores (distinct
          // docs of this slice) is a valid lower bound of the kc-th best.
          float top[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = lane * 4; j < SLICE; j += 128) {
            const float4 v = *reinterpret_cast<const float4*>(acc + j);
            const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              float x = ve[e4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {   // insertion into the descending top-4
                const float hi = fmaxf(top[t], x);
                x = fminf(top[t], x);
                top[t] = hi;
              }
            }
          }
          const int m = (kc + 31) >> 5;   // 1..4
          float seed = 0.f;
          bool found = false;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (t < m) {
              int rank = 0;   // values above top[t] under (value desc, lane asc, slot asc)
              for (int l = 0; l < 32; ++l) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float ov = __shfl_sync(0xffffffffu, top[u], l);
                  if (u < m) rank += (ov > top[t]) || (ov == top[t] && (l < lane || (l == lane && u < t)));
                }
              }
              if (rank == kc - 1) {
                seed = top[t];
                found = true;
              }
            }
          }
          const unsigned fm = __ballot_sync(0xffffffffu, found);
          if (fm) seed = __shfl_sync(0xffffffffu, seed, __ffs(fm) - 1);
          if (fm && seed > 0.f) {
            tau = make_key(seed, 0xFFFFFFFFu) - 1;   // every doc scoring >= seed still passes `key > tau`
            tau_f = seed;
            if (lane == 0) atomicMax(tau_gq, tau);
          }
        }
#pragma unroll 2
        for (int j = lane * 4; j < SLICE; j += 128) {
          float4 v = *reinterpret_cast<float4*>(acc + j);
          const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
          *reinterpret_cast<float4*>(acc + j) = make_float4(0.f, 0.f, 0.f, 0.f);
          if (__any_sync(0xffffffffu, mx > 0.f && mx >= tau_f)) {
            const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              unsigned long long key = 0;
              bool take = false;
              if (ve[e4] > 0.f && ve[e4] >= tau_f) {
                key = make_key(ve[e4], (uint32_t)(docbase + j + e4));
                take = key > tau;
              }
              sw_append(take, key, cb, cbn, cbcap, kc, tau, tau_f, lane, tau_gq);
            }
          }
        }
      } else {
        // the listed docs only: the first reader of a doc takes its score (exchange with 0), duplicates see 0
        for (int i0 = 0; i0 < nhot; i0 += 32) {
          const int i = i0 + lane;
          unsigned long long key = 0;
          bool take = false;
          if (i < nhot) {
            const int off = hotl[i];
            const float v = atomicExch(acc + off, 0.f);
            if (v > 0.f && v >= tau_f) {
              key = make_key(v, (uint32_t)(docbase + off));
              take = key > tau;
            }
          }
          sw_append(take, key, cb, cbn, cbcap, kc, tau, tau_f, lane, tau_gq);
        }
        __syncwarp();
#pragma unroll
        for (int j = lane * 4; j < SLICE; j += 128)
          *reinterpret_cast<float4*>(acc + j) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      c0 = c1;
      c1 = nxt;
    }
    // ---- end of job ----
    __syncwarp();
    if (!use_lock) {
      // small batches (few queries, many simultaneous jobs per query): the job's best kc keys go to its own slot
      // glist[q][g][kc], gcount[q][g]; bm25_sweep_merge_kernel merges the slots of a query
      if (cbn > kc) {
        for (int i = cbn + lane; i < cbcap; i += 32) cb[i] = 0;
        warp_bitonic_desc(cb, cbcap, lane);
        cbn = kc;
        if (lane == 0 && cb[kc - 1] > tau) atomicMax(tau_gq, cb[kc - 1]);
      }
      uint64_t* o = glist + ((size_t)q * S + g) * kc;
      for (int j = lane; j < cbn; j += 32) o[j] = cb[j];
      if (lane == 0) gcount[(size_t)q * S + g] = cbn;
      __syncwarp();
      continue;
    }
    // Large batches: merge the warp's candidates into the query's running top-kc (global, sorted, guarded by a
    // per-query lock; at most a few jobs of one query run at the same time).  Its kc-th key is the exact kc-th
    // best of every window scored so far: the threshold the later windows of this query start from.  Most warm
    // jobs have nothing to merge.
    if (cbn > 0) {   // warp-uniform
      if (cbn > kcp) {   // make room for the global list
        for (int i = cbn + lane; i < cbcap; i += 32) cb[i] = 0;
        warp_bitonic_desc(cb, cbcap, lane);
        cbn = min(cbn, kc);
      }
      uint64_t* gl = glist + (size_t)q * kc;
      if (lane == 0) {
        while (atomicCAS(glock + q, 0, 1) != 0) __nanosleep(100);
        __threadfence();
      }
      __syncwarp();
      const int gn = *((volatile int*)(gcount + q));
      for (int i = lane; i < gn; i += 32) cb[cbn + i] = __ldcg(reinterpret_cast<const unsigned long long*>(gl + i));
      const int total = cbn + gn;
      for (int i = total + lane; i < cbcap; i += 32) cb[i] = 0;
      warp_bitonic_desc(cb, cbcap, lane);
      const int m = min(total, kc);
      for (int i = lane; i < m; i += 32) __stcg(reinterpret_cast<unsigned long long*>(gl + i), cb[i]);
      __syncwarp();
      if (lane == 0) {
        *((volatile int*)(gcount + q)) = m;
        if (m == kc) atomicMax(tau_gq, cb[kc - 1]);
        __threadfence();
        atomicExch(glock + q, 0);
      }
      __syncwarp();
    }
  }
}

// small batches: merge the S unsorted slots of a query: keep keys >= the query's final threshold (a lower bound
// of the kc-th best key, so nothing that belongs to the top k is dropped), sort, write S_out / I_out [nq][k]
__global__ void __launch_bounds__(256)
bm25_sweep_merge_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ ns, int S, int kc, int k,
                        const unsigned long long* __restrict__ tau_g, int64_t id_base, float* __restrict__ So,
                        int64_t* __restrict__ Io) {
  __shared__ uint64_t buf[kBmMergeCap];
  __shared__ int s_n;
  const int q = blockIdx.x;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const unsigned long long thr = tau_g[q];
  const int total = S * kc;   // <= kBmMergeCap (host)
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int g = i / kc, j = i - g * kc;
    if (j < ns[(size_t)q * S + g]) {
      const uint64_t v = keys[((size_t)q * S + g) * kc + j];
      if (v >= thr && v != 0) buf[atomicAdd(&s_n, 1)] = v;
    }
  }
  __syncthreads();
  const int n = s_n;
  int pw = 1;
  while (pw < n) pw <<= 1;
  for (int i = n + threadIdx.x; i < pw; i += blockDim.x) buf[i] = 0;
  block_bitonic_desc(buf, pw);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t key = (j < n) ? buf[j] : 0ull;
    if (key) {
      So[(size_t)q * k + j] = key_score(key);
      Io[(size_t)q * k + j] = (int64_t)key_row(key) + id_base;
    } else {
      So[(size_t)q * k + j] = 0.f;
      Io[(size_t)q * k + j] = -1;
    }
  }
}

// the query's running top-kc list (sorted best first) -> S_out / I_out [nq][k]
__global__ void __launch_bounds__(256)
bm25_sweep_finish_kernel(const uint64_t* __restrict__ glist, const int* __restrict__ gcount, int nq, int kc, int k,
                         int64_t id_base, float* __restrict__ So, int64_t* __restrict__ Io) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nq * k) return;
  const int q = (int)(i / k), j = (int)(i - (int64_t)q * k);
  const uint64_t key = (j < gcount[q]) ? glist[(size_t)q * kc + j] : 0ull;
  if (key) {
    So[i] = key_score(key);
    Io[i] = (int64_t)key_row(key) + id_base;
  } else {
    So[i] = 0.f;
    Io[i] = -1;
  }
}

}  // namespace hr
