// vb_attn.cu — tcgen05 / TMEM / TMA flash-attention forward for sm_100a, driven by key/value *run lists*.
//
// One kernel serves all three VORTA branches (reference: vorta/attention/wan.py:103-149 SDPA, :243-270 coreset,
// :272-294 sliding tile).  A CTA owns up to two 128-row query tiles that attend to the same list of contiguous
// key ranges ("runs", in kernel order):
//     full     : one run [0, N)
//     coreset  : one run [0, S_c (+ text))   over the pooled sequence
//     sliding  : the <= 9 (+1 text) runs of the 3-D tile window, tile-major order
// so the block-sparse schedule is just a different table, and masked 3-D tiles are never visited.
//
// Warp roles (608 threads):
//     warps 0-15 : softmax + correction + epilogue; warp = 8 t + 4 h + q: query tile t, key / channel half h, row
//                  quarter q (TMEM lanes 32 q .. 32 q + 31) - two threads per query row
//     warp  16   : TMA producer (Q tiles once, then K/V blocks through a 5-slot ring of 32 KB)
//     warps 17,18: tcgen05.mma issuers, one per query tile (17 also allocates TMEM)
// TMEM (512 columns): S [0,128) shared by the two tiles | P0 [128,192) P1 [192,256) (bf16) | O0 [256,384) O1 [384,512).
// S = Q K^T is an SS UMMA (both operands K-major, SWIZZLE_128B as written by TMA), O += P V is a TS UMMA (P from
// TMEM, V from shared memory, MN-major).
//
// Decoupled pipeline.  P does not alias S, so QK_t(j+1) does not have to wait for softmax_t(j) -> PV_t(j): it is issued
// as soon as the softmax warps of tile t have their row maximum of block j (bar_qk_go) and the S region has been
// loaded into registers by its previous user (bar_s_free; the region alternates strictly between the tiles).  The
// softmax warps find S(j+1) complete when they finish block j and do not wait in steady state.  (With P aliased on S
// and one S per tile the chain softmax -> PV -> QK was serial per tile: 1600 + 300 + 1024 cycles per block, tensor pipe
// 61 % busy; profiles/r1g_timeline_dense16k.log vs r1l_timeline_decoupled.log.)
#include "vb_common.cuh"
#include "vb_ptx.cuh"

namespace vb {

constexpr int kNumSlots = 5;                       // K/V ring
constexpr int kTileBytes = kBlockN * kHeadDim * 2; // 32 KB: one 128x128 bf16 block = two 128x64 swizzled halves
constexpr int kHalfBytes = kTileBytes / 2;
constexpr int kSoftmaxWarps = 16;                  // 2 tiles x 2 column halves x 4 row quarters
constexpr int kTmaWarp = kSoftmaxWarps;
constexpr int kMmaWarp = kSoftmaxWarps + 1;
constexpr int kMmaWarps = 2;                       // one issuer per query tile, on different schedulers
constexpr int kAttnThreads = (kSoftmaxWarps + 1 + kMmaWarps) * 32;
#ifndef VB_POLY_GROUPS
#define VB_POLY_GROUPS 0x02        // group 1 of every 8 groups of 4 keys: 1/8 of the exps on the FMA pipe (measured best of 0, 1/8, 2/8, 3/8)
#endif
constexpr unsigned kPolyGroups = VB_POLY_GROUPS;
constexpr float kRescaleThreshold = 8.0f;          // lazy rescale: tolerate 2^8 growth before touching O

struct SmemLayout {
  uint8_t q[2 * kTileBytes];                 // two query tiles
  uint8_t kv[kNumSlots * kTileBytes];        // K/V ring
  uint64_t bar_q_full[2];
  uint64_t bar_slot_full[kNumSlots];
  uint64_t bar_slot_empty[kNumSlots];
  uint64_t bar_s_full[2];
  uint64_t bar_p_half[2];
  uint64_t bar_p_ready[2];
  uint64_t bar_o_full[2];
  uint64_t bar_s_free;          // the shared S region has been loaded into registers by its tile's softmax warps
  uint64_t bar_qk_go[2];        // softmax of tile t passed the trigger point of its block: QK of the next block may issue
  uint64_t bar_pv_done[2];      // PV of tile t completed: P_t is free again and O_t is quiescent
  uint32_t tmem_base_slot;
  KvRun runs[32];
  float xchg[2][2][kBlockM];                 // row max / row sum exchange between the two threads of a row
};
constexpr int kAttnSmemBytes = sizeof(SmemLayout);
static_assert(kAttnSmemBytes <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");

struct BlockWalker {
  const KvRun* runs;
  int n_runs, r, off;
  __device__ __forceinline__ BlockWalker(const KvRun* rr, int n) : runs(rr), n_runs(n), r(0), off(0) {}
  // next 128-key block: first key row and number of valid keys; false when exhausted
  __device__ __forceinline__ bool next(int& row0, int& valid) {
    while (r < n_runs) {
      const int len = runs[r].len;
      if (off < len) {
        row0 = runs[r].start + off;
        valid = min(kBlockN, len - off);
        off += kBlockN;
        return true;
      }
      ++r;
      off = 0;
    }
    return false;
  }
};

#ifdef VB_TIMELINE               // perf experiment: clock stamps of CTA (0,0,0) -> p.dbg as int64[who][block][8]
#define VB_STAMP(who, j, slot)                                                                         \
  do {                                                                                                 \
    if (p.dbg != nullptr && blockIdx.x == 0 && (j) < 64)         \
      reinterpret_cast<long long*>(p.dbg)[((who) * 64 + (j)) * 8 + (slot)] = clock64();                \
  } while (0)
#else
#define VB_STAMP(who, j, slot) do {} while (0)
#endif

#define VB_EXP2(x) fast_exp2(x)

__device__ __forceinline__ int count_blocks(const KvRun* runs, int n_runs) {
  int n = 0;
  for (int r = 0; r < n_runs; ++r) n += (runs[r].len + kBlockN - 1) / kBlockN;
  return n;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
vb_attn_fwd_kernel(const __grid_constant__ AttnTmaps tmaps, const __grid_constant__ AttnParams p) {
  // Everything lives in dynamic shared memory (no static __shared__), so the operand area starts at the CTA's
  // shared window base, which is 1024-byte aligned as SWIZZLE_128B needs; checked below.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  uint64_t* bar_q_full = sm.bar_q_full;
  uint64_t* bar_slot_full = sm.bar_slot_full;
  uint64_t* bar_slot_empty = sm.bar_slot_empty;
  uint64_t* bar_s_full = sm.bar_s_full;
  uint64_t* bar_p_half = sm.bar_p_half;
  uint64_t* bar_p_ready = sm.bar_p_ready;
  uint64_t* bar_o_full = sm.bar_o_full;
  uint32_t& tmem_base_slot = sm.tmem_base_slot;
  KvRun* s_runs = sm.runs;
  float (*s_xchg)[2][kBlockM] = sm.xchg;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform: role / address math in uniform registers
  const int lane = threadIdx.x & 31;
  // linear CTA index -> (segment, pair, head, batch); segments are laid out longest CTAs first by the host
  int seg_idx = 0;
  if (p.n_seg > 1 && static_cast<int>(blockIdx.x) >= p.seg[1].cta_begin) seg_idx = 1;
  if (p.n_seg > 2 && static_cast<int>(blockIdx.x) >= p.seg[2].cta_begin) seg_idx = 2;
  const AttnSeg& seg = p.seg[seg_idx];
  const int cta_local = static_cast<int>(blockIdx.x) - seg.cta_begin;
  const QPair pair = seg.pairs[cta_local % seg.n_pairs];
  const AttnHead head = p.heads[seg.head0 + (cta_local / seg.n_pairs) % seg.n_heads];
  const int batch = cta_local / (seg.n_pairs * seg.n_heads) + p.batch0;
  const int nq = pair.nq;
  const CUtensorMap& tmap_q = tmaps.m[seg_idx][0];
  const CUtensorMap& tmap_k = tmaps.m[seg_idx][1];
  const CUtensorMap& tmap_v = tmaps.m[seg_idx][2];

  if ((smem_u32(smem_raw) & 1023u) != 0) {      // SWIZZLE_128B atoms are 8 rows x 128 B
    if (threadIdx.x == 0) printf("vb_attn_fwd_kernel: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* smem_q = sm.q;                       // 2 x 32 KB
  uint8_t* smem_kv = sm.kv;                     // kNumSlots x 32 KB

  const int n_runs = min(pair.run_count, 32);
  if (threadIdx.x < n_runs) s_runs[threadIdx.x] = seg.runs[pair.run_begin + threadIdx.x];

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_q_full[i], 1);
      mbar_init(&bar_s_full[i], 1);
      mbar_init(&bar_p_half[i], kBlockM);    // 4 warps: column half 0 of P stored
      mbar_init(&bar_p_ready[i], kBlockM);   // 4 warps: column half 1 of P stored
      mbar_init(&bar_o_full[i], 1);
    }
    for (int i = 0; i < kNumSlots; ++i) {
      mbar_init(&bar_slot_full[i], 1);
      mbar_init(&bar_slot_empty[i], nq);     // one commit per query tile that consumed the slot
    }
    mbar_init(&sm.bar_s_free, kSoftmaxWarps / 2);        // lane 0 of the 8 warps of the loading tile
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.bar_qk_go[i], kSoftmaxWarps / 2);
      mbar_init(&sm.bar_pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == kMmaWarp) {
    tmem_alloc(&tmem_base_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int n_blocks = count_blocks(s_runs, n_runs);

  if (warp == kTmaWarp) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      for (int t = 0; t < nq; ++t) {
        mbar_arrive_expect_tx(&bar_q_full[t], kTileBytes);
        tma_load_4d(smem_q + t * kTileBytes, &tmap_q, &bar_q_full[t], 0, pair.q_row0[t], head.hk, batch);
        tma_load_4d(smem_q + t * kTileBytes + kHalfBytes, &tmap_q, &bar_q_full[t], 64, pair.q_row0[t], head.hk,
                    batch);
      }
      BlockWalker w(s_runs, n_runs);
      int row0, valid;
      uint32_t load_idx = 0;
      while (w.next(row0, valid)) {
#pragma unroll
        for (int which = 0; which < 2; ++which, ++load_idx) {
          const uint32_t slot = load_idx % kNumSlots;
          const uint32_t phase = (load_idx / kNumSlots) & 1u;
          mbar_wait(&bar_slot_empty[slot], phase ^ 1u);
          mbar_arrive_expect_tx(&bar_slot_full[slot], kTileBytes);
          const CUtensorMap* m = which == 0 ? &tmap_k : &tmap_v;
          uint8_t* dst = smem_kv + slot * kTileBytes;
          tma_load_4d(dst, m, &bar_slot_full[slot], 0, row0, head.hk, batch);
          tma_load_4d(dst + kHalfBytes, m, &bar_slot_full[slot], 64, row0, head.hk, batch);
        }
      }
    }
  } else if (warp >= kMmaWarp) {
    // ======================================= MMA issuer =========================================
    // The whole warp runs this role converged and every address below is made warp-uniform, so descriptors live
    // in uniform registers and one elected lane issues; a single-lane role pays a register->uniform "waterfall"
    // per operand (measured: ~250 SASS instructions per 16 MMAs, the round-1 bottleneck).
    const int nblk = __shfl_sync(0xffffffffu, n_blocks, 0);
    const int nq_u = __shfl_sync(0xffffffffu, nq, 0);
    if (nblk > 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kBlockM, kBlockN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kBlockM, kHeadDim, 0, 1);
      constexpr uint64_t desc_hi = static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
      constexpr uint32_t lbo_k = 1u << 16;
      constexpr uint32_t lbo_v = (kHalfBytes >> 4) << 16;
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t q_lo = __shfl_sync(0xffffffffu, (smem_u32(smem_q) & 0x3FFFFu) >> 4, 0);
      const uint32_t kv_lo = __shfl_sync(0xffffffffu, (smem_u32(smem_kv) & 0x3FFFFu) >> 4, 0);
      auto issue_qk = [&](int t, uint32_t slot) {
        const uint32_t a0 = q_lo + t * (kTileBytes >> 4) + lbo_k;
        const uint32_t b0 = kv_lo + slot * (kTileBytes >> 4) + lbo_k;
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          const uint32_t off = (k >> 2) * (kHalfBytes >> 4) + (k & 3) * 2;
          umma_ss(tb, desc_hi | (a0 + off), desc_hi | (b0 + off), idesc_qk, k > 0);
        }
      };
      auto issue_pv = [&](int t, uint32_t slot, uint32_t accumulate, int half) {
        const uint32_t b0 = kv_lo + slot * (kTileBytes >> 4) + lbo_v;
        const uint32_t d = tb + 2 * kBlockN + t * kHeadDim;
        const uint32_t a = tb + kBlockN + t * 64;
#pragma unroll
        for (int kk = 0; kk < kBlockN / 32; ++kk) {
          const int k = half * (kBlockN / 32) + kk;
          umma_ts(d, a + k * 8, desc_hi | (b0 + k * (2048 >> 4)), idesc_pv, accumulate | (k > 0));
        }
      };
      // One issuer warp per query tile, on different schedulers: an issuer shares its scheduler with four busy softmax
      // warps and crawls through the ~40 instructions between two MMA batches (measured 220-600 cycles); with two
      // issuers one tile's gap overlaps the other tile's batch (one issuer for both tiles: 1230 TFLOP/s dense, two:
      // 1415; a polling event loop instead of in-order suspended waits starved completely: 1167).  Each issues, in
      // order, QK_t(0), then per block QK_t(j+1) and PV_t(j) in two halves of 64 keys.
      // The S region alternates strictly between the tiles through bar_s_free: use number u = nq * j + t waits for
      // completion u - 1.
      const int t = warp - kMmaWarp;
      if (t < nq_u) {
        mbar_wait(&bar_q_full[t], 0);
        auto do_qk = [&](int j) {
          const uint32_t k_idx = 2 * j, k_slot = k_idx % kNumSlots, k_phase = (k_idx / kNumSlots) & 1u;
          const int use = nq_u * j + t;
          if (j > 0) mbar_wait(&sm.bar_qk_go[t], (j - 1) & 1);
          if (use > 0) mbar_wait(&sm.bar_s_free, (use - 1) & 1);
          mbar_wait(&bar_slot_full[k_slot], k_phase);
          tc_fence_after();
          if (lane == 0) VB_STAMP(2 + t, j, 0);
          if (elect_one()) {
            issue_qk(t, k_slot);
            umma_commit(&bar_s_full[t]);
            umma_commit(&bar_slot_empty[k_slot]);
          }
          __syncwarp();
          if (lane == 0) VB_STAMP(2 + t, j, 3);
        };
        do_qk(0);
        for (int j = 0; j < nblk; ++j) {
          const uint32_t v_idx = 2 * j + 1, v_slot = v_idx % kNumSlots, v_phase = (v_idx / kNumSlots) & 1u;
          if (j + 1 < nblk) do_qk(j + 1);
          mbar_wait(&bar_slot_full[v_slot], v_phase);
          mbar_wait(&bar_p_half[t], j & 1);
          tc_fence_after();
          if (lane == 0) VB_STAMP(2 + t, j, 1);
          if (elect_one()) issue_pv(t, v_slot, j > 0, 0);
          __syncwarp();
          mbar_wait(&bar_p_ready[t], j & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(t, v_slot, j > 0, 1);
            umma_commit(&sm.bar_pv_done[t]);
            umma_commit(&bar_slot_empty[v_slot]);
            if (j == nblk - 1) umma_commit(&bar_o_full[t]);
          }
          __syncwarp();
          if (lane == 0) VB_STAMP(2 + t, j, 2);
        }
      }
    }
  } else {
    // ============================ softmax / correction / epilogue ===============================
    // TWO threads per query row: warp = 8 t + 4 h + q handles rows [32 q, 32 q + 32) of tile t, keys / channels
    // [64 h, 64 h + 64).  All warps with the same q sit on scheduler q and may touch TMEM lanes 32 q..32 q + 31, so
    // the pair (h = 0, 1) interleaves on one scheduler: a single warp per tile left the issue slot idle ~60 % of the
    // time on fixed-latency dependencies (round-1 timeline: 1550-1700 cycles per block for ~600 instructions).
    const int t = warp >> 3;
    const int h = (warp >> 2) & 1;
    const int q4 = warp & 3;
    const int row = (q4 << 5) + lane;               // query row inside the tile == TMEM lane
    const int pair_bar = 1 + t * 4 + q4;            // named barrier of the two warps sharing these rows
    if (t < nq && n_blocks > 0) {
      const uint32_t lane_addr = static_cast<uint32_t>(q4 << 5) << 16;
      const uint32_t s_addr = tmem_base + lane_addr;                              // S region shared by both tiles
      const uint32_t p_addr = tmem_base + lane_addr + kBlockN + t * 64 + h * 32;  // P_t: its own 64 columns
      const uint32_t o_addr = tmem_base + lane_addr + 2 * kBlockN + t * kHeadDim + h * 64;
      const float scale = p.scale_log2;
      float m_ref = 0.f, l_sum = 0.f;

      BlockWalker w(s_runs, n_runs);
      int row0, valid;
      for (int j = 0; w.next(row0, valid); ++j) {
        if (row == 0 && h == 0) VB_STAMP(t, j, 0);
        mbar_wait(&bar_s_full[t], j & 1);
        tc_fence_after();
        if (row == 0 && h == 0) VB_STAMP(t, j, 1);
        uint32_t s[2][32];
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld32(s_addr + h * 64 + c * 32, s[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.bar_s_free);      // the S region may be overwritten by the next QK
        if (row == 0 && h == 0) VB_STAMP(t, j, 2);
#if defined(VB_DEBUG_DUMP) && !defined(VB_TIMELINE)   // bring-up builds only: costs ~4 % in the product kernel
        if (p.dbg != nullptr && j == 0 && blockIdx.x == 0) {
          float* d = p.dbg + (static_cast<size_t>(t) * kBlockM + row) * kBlockN + h * 64;   // raw scores of block 0
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) d[c * 32 + i] = __uint_as_float(s[c][i]);
        }
#endif
        if (valid < kBlockN) {   // run tail: keys beyond the run do not exist for this query
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (h * 64 + c * 32 + i >= valid) s[c][i] = 0xff800000u;   // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(s[c][i + 0]));
            mx1 = fmaxf(mx1, __uint_as_float(s[c][i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(s[c][i + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(s[c][i + 3]));
          }
        // row maximum over both halves (exchange through shared memory + a 64-thread named barrier)
        s_xchg[t][h][row] = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        named_bar_sync(pair_bar, 64);
        const float m_new = fmaxf(s_xchg[t][0][row], s_xchg[t][1][row]) * scale;   // block has >= 1 valid key: finite
        if (j + 1 < n_blocks && lane == 0) mbar_arrive(&sm.bar_qk_go[t]);          // QK_t(j+1) may issue now
        if (row == 0 && h == 0) VB_STAMP(t, j, 6);

        if (j == 0) {
          m_ref = m_new;
        } else {
          const bool need = m_new > m_ref + kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {             // identical in both warps of the pair (same rows)
            float alpha = 1.f;
            if (need) {
              alpha = fast_exp2(m_ref - m_new);
              m_ref = m_new;
            }
            // O_t must be quiescent: PV_t(j-1) may still be in flight (QK_t(j) was issued before it), PV_t(j) cannot
            // start before BOTH halves have arrived below.  Each thread rescales its 64 channels.
            mbar_wait(&sm.bar_pv_done[t], (j - 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t o[16];
              tmem_ld16(o_addr + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(o_addr + c * 16, o);
            }
            l_sum *= alpha;
            tmem_st_wait();
            named_bar_sync(pair_bar, 64);    // the half-0 arrive below releases PV over ALL channels
          }
        }

        // p = exp2(s * scale - m_ref): packed fp32x2 FMA / ADD (one issue slot per two keys), MUFU ex2 per key
        const float neg_m = -m_ref;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float x0, x1, x2, x3;
            ffma2(x0, x1, __uint_as_float(s[c][i + 0]), __uint_as_float(s[c][i + 1]), scale, scale, neg_m, neg_m);
            ffma2(x2, x3, __uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3]), scale, scale, neg_m, neg_m);
            float e0, e1, e2, e3;
            if ((kPolyGroups >> (i >> 2)) & 1) {      // compile-time pattern: this group of 4 keys skips the MUFU
              exp2_poly2(x0, x1, e0, e1);
              exp2_poly2(x2, x3, e2, e3);
            } else {
              e0 = VB_EXP2(x0); e1 = VB_EXP2(x1); e2 = VB_EXP2(x2); e3 = VB_EXP2(x3);
            }
            fadd2(sum0, sum1, sum0, sum1, e0, e1);
            fadd2(sum2, sum3, sum2, sum3, e2, e3);
            pk[c * 16 + (i >> 1) + 0] = pack_bf16x2(e0, e1);
            pk[c * 16 + (i >> 1) + 1] = pack_bf16x2(e2, e3);
          }
#ifdef VB_TIMELINE
          if (c == 0) { asm volatile("" ::: "memory"); if (row == 0 && h == 0) VB_STAMP(t, j, 7); }
#endif
        }
        if (j > 0) {                       // P_t(j-1) must have been consumed before it is overwritten
          mbar_wait(&sm.bar_pv_done[t], (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st32(p_addr, pk);            // P_t(j): keys [64 h, 64 h + 64) -> 32 columns of bf16 pairs
        l_sum += (sum0 + sum1) + (sum2 + sum3);
        if (row == 0 && h == 0) VB_STAMP(t, j, 3);
        tmem_st_wait();
        if (row == 0 && h == 0) VB_STAMP(t, j, 4);
        tc_fence_before();
        mbar_arrive(h == 0 ? &bar_p_half[t] : &bar_p_ready[t]);   // PV over keys [0,64) / [64,128) may start
        if (row == 0 && h == 0) VB_STAMP(t, j, 5);
      }

      // ------------------------------------ epilogue ------------------------------------
      s_xchg[t][h][row] = l_sum;
      named_bar_sync(pair_bar, 64);
      l_sum = s_xchg[t][0][row] + s_xchg[t][1][row];
      mbar_wait(&bar_o_full[t], 0);
      tc_fence_after();
      const bool row_ok = row < pair.q_rows[t];
      const int krow = pair.q_row0[t] + row;      // row in kernel order
      const float inv = head.weight / l_sum;
      const bool accumulate = (head.flags & 1) != 0;
      int n_dst = 0;
      int64_t dst_tok = 0;
      const int32_t* bc = nullptr;
      if (row_ok) {
        dst_tok = seg.out_map ? seg.out_map[batch * seg.out_map_stride_b + head.hk * seg.out_map_stride_h + krow] : krow;
        n_dst = 1;
        if (seg.bcast_map != nullptr && krow < seg.bcast_rows) {
          bc = seg.bcast_map + batch * seg.bcast_stride_b + head.hk * seg.bcast_stride_h +
               static_cast<int64_t>(krow) * seg.bcast_n;
          n_dst += seg.bcast_n;
        }
      }
      const int64_t head_off = batch * p.out_stride_b + head.ho * p.out_stride_h + h * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(o_addr + c * 32, o);
        tmem_ld_wait();
#if defined(VB_DEBUG_DUMP) && !defined(VB_TIMELINE)   // bring-up builds only: costs ~4 % in the product kernel
        if (p.dbg != nullptr && blockIdx.x == 0) {
          float* d = p.dbg + 2 * kBlockM * kBlockN + (static_cast<size_t>(t) * kBlockM + row) * kHeadDim + h * 64;
#pragma unroll
          for (int i = 0; i < 32; ++i) d[c * 32 + i] = __uint_as_float(o[i]);   // un-normalised O
          if (c == 0 && h == 0) p.dbg[4 * kBlockM * kBlockN + t * kBlockM + row] = l_sum;
        }
#endif
        for (int dsti = 0; dsti < n_dst; ++dsti) {
          int64_t tok = dsti == 0 ? dst_tok : static_cast<int64_t>(bc[dsti - 1]);
          __nv_bfloat16* obase = p.out;
          if (p.out_peer_count > 0) {      // fused Ulysses "out" exchange: the owner rank of this token gets the row
            const int peer = static_cast<int>(tok / p.out_peer_rows);
            obase = p.out_peers[peer];
            tok -= static_cast<int64_t>(peer) * p.out_peer_rows;
          }
          uint4* dst = reinterpret_cast<uint4*>(obase + head_off + tok * p.out_stride_s + c * 32);
#pragma unroll
          for (int q4i = 0; q4i < 4; ++q4i) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[q4i * 8 + i]) * inv;
            if (accumulate) {
              const uint4 prev = dst[q4i];
              const uint32_t pw[4] = {prev.x, prev.y, prev.z, prev.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                f[2 * i + 0] += __uint_as_float(pw[i] << 16);
                f[2 * i + 1] += __uint_as_float(pw[i] & 0xffff0000u);
              }
            }
            uint4 v;
            v.x = pack_bf16x2(f[0], f[1]);
            v.y = pack_bf16x2(f[2], f[3]);
            v.z = pack_bf16x2(f[4], f[5]);
            v.w = pack_bf16x2(f[6], f[7]);
            dst[q4i] = v;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
  }
  return fn;
}

// 4-D bf16 tensor map over (channel=128, token=n_rows, head, batch) with a (64, 128, 1, 1) box, SWIZZLE_128B.
int make_qkv_tensor_map(CUtensorMap* map, const void* base, int64_t n_rows, int64_t heads, int64_t batch,
                        int64_t stride_b, int64_t stride_h, int64_t stride_s) {
  PFN_encodeTiled enc = get_encode_fn();
  VB_REQUIRE(enc != nullptr, VB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, VB_ERR_INVALID, "tensor base must be 16-byte aligned");
  VB_REQUIRE(stride_s % 8 == 0 && stride_h % 8 == 0 && stride_b % 8 == 0, VB_ERR_INVALID,
             "token/head/batch strides must be multiples of 8 elements (16 bytes)");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(kHeadDim), static_cast<cuuint64_t>(n_rows),
                        static_cast<cuuint64_t>(heads), static_cast<cuuint64_t>(batch)};
  // a size-1 dimension may carry a zero stride from the caller; TMA needs a positive multiple of 16 bytes
  auto fix = [](int64_t s) { return static_cast<cuuint64_t>((s > 0 ? s : 8) * 2); };
  cuuint64_t strides[3] = {fix(stride_s), fix(stride_h), fix(stride_b)};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(kBlockN), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VB_REQUIRE(r == CUDA_SUCCESS, VB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VB_OK;
}

// `params.seg[i].cta_begin` must already hold the prefix sums; n_ctas = total over the segments.
int launch_attn(const AttnTmaps& tmaps, const AttnParams& params, int n_ctas, cudaStream_t stream) {
  static bool configured[64] = {false};      // the attribute is per device
  int dev = 0;
  VB_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    VB_CUDA_OK(cudaFuncSetAttribute(vb_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kAttnSmemBytes));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  if (n_ctas == 0) return VB_OK;
  vb_attn_fwd_kernel<<<static_cast<unsigned>(n_ctas), kAttnThreads, kAttnSmemBytes, stream>>>(tmaps, params);
  VB_CUDA_OK(cudaGetLastError());
  return VB_OK;
}

}  // namespace vb
