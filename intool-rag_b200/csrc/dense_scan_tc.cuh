// dense_scan_tc.cuh — the dense filter scan: query x corpus inner products on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA, with the per-query top-KL
// selection fused into the epilogue so the score matrix never reaches HBM.
//
// Replaces the hot loop of faiss::IndexFlat::search called at
// /root/reference/rag/storage/faiss_index.py:83 (N*d FMAs + N heap tests per query).
//
// GEMM shape:  D[128 queries, 256 corpus rows] = Q_tile[128, K] * X_tile[256, K]^T, both K-major.
//   M (TMEM lanes)   = queries      -> one epilogue thread owns one query: thresholds in registers
//   N (TMEM columns) = corpus rows  -> 2 accumulator stages x 256 columns = all 512 TMEM columns
//   K                = d, streamed in 128-byte blocks (32 tf32 / 64 bf16 elements), 4-stage ring
// Persistent CTAs (one per SM): CTA b owns corpus tiles b, b+G, ... and, for each, loops over the
// query tiles, so a corpus tile is fetched from HBM once and re-read from L2 for nq > 128.
//
// Warp roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (warp w reads TMEM lanes 32*(w%4)..+31).
//
// Epilogue selection: thread (query) keeps a private candidate list of KL (score,row) entries in
// global memory (L2 resident) with replace-min insertion, a local threshold tau_l (its KL-th best)
// and a cross-CTA threshold tau_g[q] (atomicMax of every CTA's tau_l).  A score is inserted only if
// it beats max(tau_l, tau_g); the common case is one FMNMX per score and one compare per 32.
// Invariant used by the certificate (dense_exact.cuh): every row that is NOT in some list at the
// end has approximate score <= final tau_g[q].
#pragma once
#include "common.cuh"

namespace hr {

constexpr int kScanBM = 128;
constexpr int kScanBN = 256;
constexpr int kScanStages = 4;
constexpr int kScanABytes = kScanBM * 128;
constexpr int kScanBBytes = kScanBN * 128;
constexpr int kScanStageBytes = kScanABytes + kScanBBytes;
constexpr int kScanNqMax = 2048;  // queries per launch (per-query state lives in shared memory)
constexpr int kScanThreads = 256;

struct ScanSmemTail {
  uint64_t full_bar[kScanStages];
  uint64_t empty_bar[kScanStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_ptr;
  uint32_t pad_[3];
  float half_norms[2][kScanBN];
  float tau_l[kScanNqMax];
  uint16_t cnt[kScanNqMax];
  uint16_t minpos[kScanNqMax];
};
constexpr int kScanSmemBytes = kScanStages * kScanStageBytes + (int)sizeof(ScanSmemTail) + 1024;

struct ScanParams {
  int64_t N;        // corpus rows
  int nq;           // queries in this launch (<= kScanNqMax)
  int kblocks;      // 128-byte K blocks per row (ld * elem_size / 128)
  int KL;           // candidate list length
  int tile_count;   // corpus tiles this launch visits: tiles t*tile_stride, t in [0, tile_count)
  int tile_stride;  // 1 = every tile (main pass), >1 = strided sample (threshold pre-pass)
  int num_qtiles;   // ceil(nq / 128)
  const float* norms;  // |x|^2 per row (L2 only)
  Cand* lists;         // [gridDim.x][nq][KL]
  int* cnts;           // [gridDim.x][nq]
  unsigned int* tau_g; // [nq] ordered-uint threshold, 0 = unset
  float* pre_max;      // threshold pre-pass: [tile_count][nq] best score of each sampled tile (nullptr = main pass)
};

__device__ __noinline__ void cand_insert(Cand* list, int KL, float v, uint32_t row, float& tau_l, uint32_t& cnt,
                                         uint32_t& minpos, float& thr, float tg, unsigned int* tau_g_slot) {
  Cand c;
  c.s = v;
  c.row = row;
  if (cnt < (uint32_t)KL) {
    list[cnt] = c;
    cnt++;
    if (cnt < (uint32_t)KL) return;
  } else {
    list[minpos] = c;
  }
  float m = list[0].s;
  uint32_t mp = 0;
  for (int i = 1; i < KL; ++i) {
    float s = list[i].s;
    if (s < m) {
      m = s;
      mp = i;
    }
  }
  tau_l = m;
  minpos = mp;
  if (m > tg) atomicMax(tau_g_slot, f2ord(m));
  thr = fmaxf(m, tg);
}

// One (corpus tile, query tile) accumulator: thread = query q (TMEM lane), 256 columns = corpus rows.
// Reads the accumulator in 32-column chunks and feeds scores above the query's threshold to its list.
// Threshold pre-pass: no lists, only the best score of the tile for this query (31 FMNMX per 32 scores).
template <int METRIC>
__device__ __forceinline__ void scan_epilogue_max(ScanSmemTail* st, const ScanParams& p, uint32_t taddr, int q, int nb,
                                                  int col_limit, int tile_slot) {
  float best = HR_NEG_INF;
#pragma unroll 1
  for (int ch = 0; ch < kScanBN / 32; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + ch * 32, r);
    tmem_ld_wait();
    const int cbase = ch * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float v = __uint_as_float(r[j]);
      if (METRIC == 1) v -= st->half_norms[nb][cbase + j];
      if (cbase + j < col_limit) best = fmaxf(best, v);
    }
  }
  if (q < p.nq) p.pre_max[(size_t)tile_slot * p.nq + q] = best;
}

template <int METRIC>
__device__ __forceinline__ void scan_epilogue_item(ScanSmemTail* st, const ScanParams& p, uint32_t taddr, int q, int nb,
                                                   int col_limit, uint32_t row0) {
        const bool active = q < p.nq;
        float tau_l = HR_NEG_INF, tg = HR_NEG_INF, thr = HR_NEG_INF;
        uint32_t cnt = 0, minpos = 0;
        Cand* list = nullptr;
        if (active) {
          tau_l = st->tau_l[q];
          cnt = st->cnt[q];
          minpos = st->minpos[q];
          unsigned int o = *((volatile unsigned int*)&p.tau_g[q]);
          tg = o ? ord2f(o) : HR_NEG_INF;
          thr = (cnt == (uint32_t)p.KL) ? fmaxf(tau_l, tg) : tg;
          list = p.lists + ((size_t)blockIdx.x * p.nq + q) * p.KL;
        }
#pragma unroll 1
        for (int ch = 0; ch < kScanBN / 32; ++ch) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + ch * 32, r);
          tmem_ld_wait();
          if (active) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = __uint_as_float(r[j]);
              if (METRIC == 1) v[j] -= st->half_norms[nb][ch * 32 + j];
            }
            const int cbase = ch * 32;
            if (cbase + 32 <= col_limit) {
              float mx = v[0];
#pragma unroll
              for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
              if (mx > thr) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (v[j] > thr)
                    cand_insert(list, p.KL, v[j], row0 + cbase + j, tau_l, cnt, minpos, thr, tg, &p.tau_g[q]);
              }
            } else if (cbase < col_limit) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cbase + j < col_limit && v[j] > thr)
                  cand_insert(list, p.KL, v[j], row0 + cbase + j, tau_l, cnt, minpos, thr, tg, &p.tau_g[q]);
            }
          }
        }
        if (active) {
          st->tau_l[q] = tau_l;
          st->cnt[q] = (uint16_t)cnt;
          st->minpos[q] = (uint16_t)minpos;
        }
}

// KIND 0: fp32 storage, kind::tf32 (32 elements per K block).  KIND 1: bf16 storage, kind::f16.
template <int KIND, int METRIC>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
               const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  ScanSmemTail* st = (ScanSmemTail*)(smem + kScanStages * kScanStageBytes);
  constexpr int kKElems = (KIND == 0) ? 32 : 64;
  constexpr uint32_t kIdesc = umma_idesc(KIND == 0 ? 2u : 1u, kScanBM, kScanBN);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kScanStages; ++s) {
      mbar_init(&st->full_bar[s], 1);
      mbar_init(&st->empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&st->tmem_full[a], 1);
      mbar_init(&st->tmem_empty[a], 128);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(&st->tmem_ptr, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kScanNqMax; i += blockDim.x) {
    st->tau_l[i] = HR_NEG_INF;
    st->cnt[i] = 0;
    st->minpos[i] = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = st->tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.tile_count; t += gridDim.x) {
        const int c = t * p.tile_stride;
        for (int m = 0; m < p.num_qtiles; ++m) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&st->empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&st->full_bar[stage], kScanStageBytes);
            uint8_t* sa = smem + stage * kScanStageBytes;
            tma_load_2d(sa, &tmap_q, &st->full_bar[stage], kb * kKElems, m * kScanBM);
            tma_load_2d(sa + kScanABytes, &tmap_x, &st->full_bar[stage], kb * kKElems, c * kScanBN);
            if (++stage == kScanStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.tile_count; t += gridDim.x) {
        for (int m = 0; m < p.num_qtiles; ++m, ++it) {
          const uint32_t as = it & 1u;
          const uint32_t aphase = (it >> 1) & 1u;
          mbar_wait(&st->tmem_empty[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * kScanBN;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&st->full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kScanStageBytes);
            const uint64_t adesc = umma_desc_sw128(sa);
            const uint64_t bdesc = umma_desc_sw128(sa + kScanABytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // advance 32 bytes of K inside the 128-byte swizzle span: +2 in the (addr >> 4) field
              if (KIND == 0)
                tc_mma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (uint32_t)((kb | k) != 0));
              else
                tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (uint32_t)((kb | k) != 0));
            }
            tc_commit(&st->empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (++stage == kScanStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          tc_commit(&st->tmem_full[as]);  // accumulator ready for the epilogue
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: fused per-query top-KL =====================
    const int ew = warp & 3;
    const int et = threadIdx.x - 128;
    uint32_t it = 0;
    uint32_t cit = 0;
    for (int t = blockIdx.x; t < p.tile_count; t += gridDim.x, ++cit) {
      const int c = t * p.tile_stride;
      const int nb = cit & 1;
      if (METRIC == 1) {
        for (int j = et; j < kScanBN; j += 128) {
          int64_t row = (int64_t)c * kScanBN + j;
          st->half_norms[nb][j] = row < p.N ? 0.5f * p.norms[row] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const int64_t rem = p.N - (int64_t)c * kScanBN;
      const int col_limit = rem < kScanBN ? (int)rem : kScanBN;
      const uint32_t row0 = (uint32_t)c * kScanBN;
      for (int m = 0; m < p.num_qtiles; ++m, ++it) {
        const uint32_t as = it & 1u;
        const uint32_t aphase = (it >> 1) & 1u;
        const int q = m * kScanBM + ew * 32 + lane;
        mbar_wait(&st->tmem_full[as], aphase);
        tc_fence_after();
        if (p.pre_max)
          scan_epilogue_max<METRIC>(st, p, tmem_base + as * kScanBN + ((uint32_t)(ew * 32) << 16), q, nb, col_limit, t);
        else
          scan_epilogue_item<METRIC>(st, p, tmem_base + as * kScanBN + ((uint32_t)(ew * 32) << 16), q, nb, col_limit,
                                     row0);
        tc_fence_before();
        mbar_arrive(&st->tmem_empty[as]);
      }
    }
    for (int q = et; q < p.nq; q += 128) p.cnts[(size_t)blockIdx.x * p.nq + q] = st->cnt[q];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================
// CTA-pair variant (tcgen05 cta_group::2) for batches of more than 128 queries.
//   One MMA spans two SMs: M = 256 queries (CTA r of the pair owns queries [r*128, r*128+128) of the
//   256-query pair tile and their accumulators in its own TMEM), N = 256 corpus rows of which each CTA
//   stages only its half (128 rows) in shared memory.  Per SM and K block that is 16 KB of queries +
//   16 KB of corpus for 128x256x32 MACs: 2/3 of the single-CTA kernel's operand traffic, and only one
//   1 MB corpus tile in flight per PAIR, so the tiles of all pairs (74 MB) stay L2 resident while the
//   pair sweeps its query tiles over them.
//   Barriers: TMA of both CTAs completes on the LEADER's full barrier; the leader's MMA thread commits
//   (multicast) to both CTAs' empty / tmem_full barriers; both epilogues arrive on the leader's tmem_empty.
// =====================================================================================================
constexpr int kScan2Stages = 6;
constexpr int kScan2HalfBytes = 128 * 128;                 // 128 rows x 128 B
constexpr int kScan2StageBytes = 2 * kScan2HalfBytes;      // A half + B half
constexpr int kScan2SmemBytes = kScan2Stages * kScan2StageBytes + (int)sizeof(ScanSmemTail) + 1024;
static_assert(kScan2Stages <= 8, "barrier arrays");

struct Scan2Bars {
  uint64_t full_bar[kScan2Stages];
  uint64_t empty_bar[kScan2Stages];
};

template <int KIND, int METRIC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kScanThreads, 1)
scan_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  ScanSmemTail* st = (ScanSmemTail*)(smem + kScan2Stages * kScan2StageBytes);
  __shared__ Scan2Bars bars;
  constexpr int kKElems = (KIND == 0) ? 32 : 64;
  constexpr uint32_t kIdesc = umma_idesc(KIND == 0 ? 2u : 1u, 256, kScanBN);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int num_ptiles = (p.nq + 255) / 256;   // 256-query pair tiles

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kScan2Stages; ++s) {
      mbar_init(&bars.full_bar[s], 1);
      mbar_init(&bars.empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&st->tmem_full[a], 1);
      mbar_init(&st->tmem_empty[a], 256);   // 128 epilogue threads of each CTA
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc2(&st->tmem_ptr, 512);
    tmem_relinquish2();
  }
  for (int i = threadIdx.x; i < kScanNqMax; i += blockDim.x) {
    st->tau_l[i] = HR_NEG_INF;
    st->cnt[i] = 0;
    st->minpos[i] = 0;
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = st->tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < p.tile_count; t += npairs) {
        const int c = t * p.tile_stride;
        for (int m = 0; m < num_ptiles; ++m) {
          // the corpus tile is re-read by the following query tiles: keep it in L2 until the last one
          const uint64_t xpol = (m == num_ptiles - 1) ? kEvictFirst : kEvictLast;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&bars.empty_bar[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&bars.full_bar[stage], 2 * kScan2StageBytes);
            uint8_t* sa = smem + stage * kScan2StageBytes;
            tma_load_2d_pair(sa, &tmap_q, &bars.full_bar[stage], kb * kKElems, m * 256 + (int)rank * 128,
                             kEvictLast);
            tma_load_2d_pair(sa + kScan2HalfBytes, &tmap_x, &bars.full_bar[stage], kb * kKElems,
                             c * kScanBN + (int)rank * 128, xpol);
            if (++stage == kScan2Stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int t = pair; t < p.tile_count; t += npairs) {
        for (int m = 0; m < num_ptiles; ++m, ++it) {
          const uint32_t as = it & 1u;
          const uint32_t aphase = (it >> 1) & 1u;
          mbar_wait(&st->tmem_empty[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * kScanBN;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&bars.full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kScan2StageBytes);
            const uint64_t adesc = umma_desc_sw128(sa);
            const uint64_t bdesc = umma_desc_sw128(sa + kScan2HalfBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (KIND == 0)
                tc_mma2_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (uint32_t)((kb | k) != 0));
              else
                tc_mma2_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (uint32_t)((kb | k) != 0));
            }
            tc_commit2(&bars.empty_bar[stage]);
            if (++stage == kScan2Stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          tc_commit2(&st->tmem_full[as]);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs; CTA r owns queries r*128.. of each pair tile) =========
    const int ew = warp & 3;
    const int et = threadIdx.x - 128;
    uint32_t it = 0;
    uint32_t cit = 0;
    for (int t = pair; t < p.tile_count; t += npairs, ++cit) {
      const int c = t * p.tile_stride;
      const int nb = cit & 1;
      if (METRIC == 1) {
        for (int j = et; j < kScanBN; j += 128) {
          int64_t row = (int64_t)c * kScanBN + j;
          st->half_norms[nb][j] = row < p.N ? 0.5f * p.norms[row] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const int64_t rem = p.N - (int64_t)c * kScanBN;
      const int col_limit = rem < kScanBN ? (int)rem : kScanBN;
      const uint32_t row0 = (uint32_t)c * kScanBN;
      for (int m = 0; m < num_ptiles; ++m, ++it) {
        const uint32_t as = it & 1u;
        const uint32_t aphase = (it >> 1) & 1u;
        const int q = m * 256 + (int)rank * 128 + ew * 32 + lane;
        mbar_wait(&st->tmem_full[as], aphase);
        tc_fence_after();
        if (p.pre_max)
          scan_epilogue_max<METRIC>(st, p, tmem_base + as * kScanBN + ((uint32_t)(ew * 32) << 16), q, nb, col_limit, t);
        else
          scan_epilogue_item<METRIC>(st, p, tmem_base + as * kScanBN + ((uint32_t)(ew * 32) << 16), q, nb, col_limit,
                                     row0);
        tc_fence_before();
        mbar_arrive_leader(&st->tmem_empty[as]);
      }
    }
    for (int q = et; q < p.nq; q += 128) p.cnts[(size_t)blockIdx.x * p.nq + q] = st->cnt[q];
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still touch this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---- merge of the per-CTA candidate lists -> shortlist for the exact re-score -------------------
// One block per query.  Keeps entries with score >= T* (= final tau_g), sorts them by key, emits
// the best KL rows and tprime = upper bound on the approximate score of everything left out.
// When more than kShortCap entries survive T* (small corpora: few tiles per CTA, so the shared
// threshold never tightens) an exact 4-pass radix select finds the KL-th largest score first.
constexpr int kShortCap = 2048;

__device__ __forceinline__ int merge_compact(const Cand* __restrict__ lists, const int* __restrict__ cnts, int G,
                                             int nq, int KL, int q, float tstar, uint32_t ord_min, uint64_t* buf,
                                             int* s_n) {
  if (threadIdx.x == 0) *s_n = 0;
  __syncthreads();
  // a warp per list, lanes over its entries: only the filled part of each list is read
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int g = warp; g < G; g += nwarp) {
    const int cnt = min(cnts[(size_t)g * nq + q], KL);
    const Cand* l = lists + ((size_t)g * nq + q) * KL;
    for (int j = lane; j < cnt; j += 32) {
      const Cand c = l[j];
      if (c.s >= tstar && f2ord(c.s) >= ord_min) {
        int slot = atomicAdd(s_n, 1);
        if (slot < kShortCap) buf[slot] = make_key(c.s, c.row);
      }
    }
  }
  __syncthreads();
  return *s_n;
}

__global__ void __launch_bounds__(256)
scan_merge_kernel(const Cand* __restrict__ lists, const int* __restrict__ cnts,
                  const unsigned int* __restrict__ tau_g, int G, int nq, int KL,
                  uint32_t* __restrict__ short_rows, int* __restrict__ short_n, float* __restrict__ tprime,
                  int* __restrict__ overflow_count, unsigned int* __restrict__ tau_seed,
                  int* __restrict__ short_tot, float* __restrict__ tprime_tot, float* __restrict__ short_s) {
  __shared__ uint64_t buf[kShortCap];
  __shared__ int s_n;
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining;
  const int q = blockIdx.x;
  const unsigned int o = tau_g[q];
  const float tstar = o ? ord2f(o) : HR_NEG_INF;
  int n = merge_compact(lists, cnts, G, nq, KL, q, tstar, 0u, buf, &s_n);
  bool radix = false;
  if (n > kShortCap) {
    // exact m-th largest ordered score among the entries >= T*, m = max(KL, 1024)  (MSB-first radix select)
    radix = true;
    if (threadIdx.x == 0) {
      s_prefix = 0;
      s_remaining = max(KL, kShortCap / 2);   // keep a deep candidate set for the second re-score stage
    }
    const int total = G * KL;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
      __syncthreads();
      const uint32_t mask = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
      const uint32_t prefix = s_prefix;
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int g = i / KL, j = i - g * KL;
        if (j < cnts[(size_t)g * nq + q]) {
          float sc = lists[((size_t)g * nq + q) * KL + j].s;
          uint32_t od = f2ord(sc);
          if (sc >= tstar && (od & mask) == (prefix & mask)) atomicAdd(&hist[(od >> shift) & 255u], 1);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int cum = 0, rem = s_remaining;
        for (int b = 255; b >= 0; --b) {
          if (cum + hist[b] >= rem) {
            s_prefix = prefix | ((uint32_t)b << shift);
            s_remaining = rem - cum;
            break;
          }
          cum += hist[b];
        }
      }
      __syncthreads();
    }
    n = merge_compact(lists, cnts, G, nq, KL, q, tstar, s_prefix, buf, &s_n);
  }
  // rows are written best first with stride kShortCap: the first KL are the stage-1 shortlist, the rest (every
  // candidate that survived the thresholds) is only re-scored if the certificate fails on the first KL
  uint32_t* out = short_rows + (size_t)q * kShortCap;
  if (n > kShortCap) {  // massive exact ties at the KL-th score: cannot rank -> exact fallback
    if (threadIdx.x == 0) {
      short_n[q] = 0;
      short_tot[q] = 0;
      tprime_tot[q] = -HR_NEG_INF;
      tprime[q] = -HR_NEG_INF;
      if (tau_seed) tau_seed[q] = o;
      else atomicAdd(overflow_count, 1);
    }
    for (int j = threadIdx.x; j < KL; j += blockDim.x) out[j] = 0xFFFFFFFFu;
    return;
  }
  int pw = 1;
  while (pw < n) pw <<= 1;
  for (int i = n + threadIdx.x; i < pw; i += blockDim.x) buf[i] = 0;
  block_bitonic_desc(buf, pw);
  for (int j = threadIdx.x; j < max(KL, n); j += blockDim.x) {
    out[j] = (j < n) ? key_row(buf[j]) : 0xFFFFFFFFu;
    if (j < n) short_s[(size_t)q * kShortCap + j] = key_score(buf[j]);   // filter scores, best first
  }
  if (threadIdx.x == 0) {
    short_n[q] = n < KL ? n : KL;
    short_tot[q] = n;
    // bound on the approximate score of every row outside the n candidates: the final threshold (or the radix cut)
    float tt = o ? tstar : HR_NEG_INF;
    if (radix && n > 0) tt = fmaxf(tt, key_score(buf[n - 1]));
    tprime_tot[q] = tt;
    float tp = HR_NEG_INF;
    if (o) tp = tstar;                                  // rows dropped by a threshold are <= T*
    if (n > KL) tp = fmaxf(tp, key_score(buf[KL]));     // best compacted entry that was left out
    else if (radix && n > 0) tp = fmaxf(tp, key_score(buf[n - 1]));  // entries below the radix threshold
    tprime[q] = tp;
    // threshold pre-pass: the KL-th best score of the SAMPLE is a valid lower bound of the KL-th best
    // of the whole corpus (the sample rows are corpus rows) -> seed of tau_g for the main pass
    if (tau_seed) tau_seed[q] = (n >= KL) ? (uint32_t)(buf[KL - 1] >> 32) : o;
  }
}

// ---- threshold seed from the pre-pass ---------------------------------------------------------------
// pre_max [T][nq]: the best filter score of each of T <= kSeedCap sampled tiles (each the score of a distinct
// corpus row).  The j-th largest of them is <= the j-th best score of the sample; its rank R in the corpus is
// Gamma(j, stride) distributed (mean j * stride, relative spread 1/sqrt(j)): with j >= 32 the seed is both close
// to its target rank and, for all practical purposes, never inside the true top KL.  One block per query.
constexpr int kSeedCap = 2048;
__global__ void __launch_bounds__(256)
scan_seed_kernel(const float* __restrict__ pre_max, int T, int nq, int j, unsigned int* __restrict__ tau_g) {
  __shared__ uint64_t buf[kSeedCap];
  const int q = blockIdx.x;
  int pw = 1;
  while (pw < T) pw <<= 1;
  for (int i = threadIdx.x; i < pw; i += blockDim.x)
    buf[i] = i < T ? make_key(pre_max[(size_t)i * nq + q], (uint32_t)i) : 0ull;
  block_bitonic_desc(buf, pw);
  if (threadIdx.x == 0) {
    const uint64_t key = buf[min(j, T) - 1];
    if (key != 0ull && key_score(key) > HR_NEG_INF) tau_g[q] = (uint32_t)(key >> 32);
  }
}

}  // namespace hr
