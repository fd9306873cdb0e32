// Epilogue policies of the X-stationary kernel (umma_xw.cuh) for the ArcFace head.
//
// Operands are the fp16 rows K1 emits: x_hat * S and w_hat * S (already L2-normalised, S a power of two), so
// an accumulator is  acc = S^2 * cos(theta)  and the epilogues need no per-row / per-column scale.  Every
// policy has a fast path for whole 32-column slices (no clamp, no target column, no NaN scrub -- validated
// after the fact from the slice's min / max / sum) and a careful path that applies the reference's
// element-wise sequence (src/face_models.py:363-427) to the same registers.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma_epilogues.cuh"
#include "umma_xw.cuh"

namespace b200f {
namespace umma {

constexpr int PART_COLS = 6;   // sumexp, sumexp2, ztarget, sumz, best, bestidx (int32 bits)

// phi out of line: its acosf / cosf code is large, and it runs for one element of one slice in a hundred
__device__ __noinline__ float head_phi_once(HeadMath hm, float c) { return hm.phi(c); }

// -------------------------------------------------------------------------------------------------
// Raw accumulators -> out[row, class] (self-test of the kernel itself).
struct XwStore {
  struct Params { float* out; int64_t ld; };
  struct State { bool row_ok; };
  static __device__ __forceinline__ void item_begin(State& st, const Params&, const XwParams& p, const XwItem& it) {
    st.row_ok = it.row < p.B;
  }
  static __device__ __forceinline__ void tile_begin(State&, const Params&, const XwParams&, const XwItem&, int, int, int) {}
  static __device__ __forceinline__ void slice(State& st, const Params& ep, const XwParams& p, const XwItem& it,
                                               float (&v)[32], int cls0) {
    if (!st.row_ok) return;
    const int cc = min(32, p.C - cls0);
    float* dst = ep.out + it.row * ep.ld + cls0;
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j < cc) dst[j] = v[j];
  }
  static __device__ __forceinline__ void item_end(State&, const Params&, const XwParams&, const XwItem&, float*) {}
};

// Pipeline probe: the epilogue only folds every accumulator into one checksum per thread (no math, no stores
// but one per item) -- what is left is the TMA / MMA / TMEM pipeline itself.
struct XwNull {
  struct Params { float* out; };
  struct State { float acc; };
  static __device__ __forceinline__ void item_begin(State& st, const Params&, const XwParams&, const XwItem&) { st.acc = 0.f; }
  static __device__ __forceinline__ void tile_begin(State&, const Params&, const XwParams&, const XwItem&, int, int, int) {}
  static __device__ __forceinline__ void slice(State& st, const Params&, const XwParams&, const XwItem&, float (&v)[32], int) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) { a += v[j]; b += v[j + 1]; }
    st.acc += a + b;
  }
  static __device__ __forceinline__ void item_end(State& st, const Params& ep, const XwParams& p, const XwItem& it, float*) {
    if (it.row < p.B) atomicAdd(ep.out + it.row, st.acc);
  }
};

// -------------------------------------------------------------------------------------------------
// K2: cosine logits -> margin -> scale -> softmax / cross-entropy statistics (src/face_models.py:355-427,
// training.py:515).  Nothing B x C is stored: ONE partial record per (row, class chunk).
template <int EG, int SC, int PW = 0>
struct XwFwdT {
  static constexpr int kEpiGroups = EG;
  static constexpr int kSliceCols = SC;
  static constexpr int kPrepWarps = PW;        // warps that run K1 of the class weights inside the kernel (XwParams::prep_*)
  struct Params {
    const int64_t* label;
    int64_t class_offset;       // global id of this launch's class 0
    HeadMath hm;
    float inv_scale;            // 1 / S^2 : cos = acc * inv_scale
    float* part;                // [B, n_chunks * EG, PART_COLS]
    float* cos_part;            // [items * PAIR * 8 * EG, 2]
    int32_t* nan_flag;
    int pair;
    unsigned int* zero_word;    // a word of the workspace the kernel clears (block counter of the fused loss finalize)
    int whole_slice_targets;    // tunable "target_patch" = 0: a slice with a target column goes the element-wise way (round 1)
  };
  struct State {
    float sumexp, sumexp2, sumz, ztgt, best, cmin, cmax;
    int bestidx, tgt;
    bool row_ok, saw_nan;
  };

  static __device__ __forceinline__ void item_begin(State& st, const Params& ep, const XwParams& p, const XwItem& it) {
    st.sumexp = st.sumexp2 = st.sumz = st.ztgt = 0.f;
    st.best = -INFINITY; st.bestidx = -1;
    st.cmin = INFINITY; st.cmax = -INFINITY;
    st.saw_nan = false;
    st.row_ok = it.row < p.B;
    st.tgt = -1;
    if (it.item == 0 && it.rank == 0 && it.grp == 0 && it.ew == 0 && it.lane == 0 && ep.zero_word != nullptr) *ep.zero_word = 0u;
    if (st.row_ok) {
      const int64_t tg = __ldg(ep.label + it.row) - ep.class_offset;
      if (tg >= 0 && tg < p.C) st.tgt = (int)tg;
    }
  }

  static __device__ __forceinline__ void tile_begin(State&, const Params&, const XwParams&, const XwItem&, int, int, int) {}
  static __device__ __forceinline__ void slice(State& st, const Params& ep, const XwParams& p, const XwItem& it,
                                               float (&v)[SC], int cls0) {
    const int cc = min(SC, p.C - cls0);
    const float s_eff = ep.hm.s_eff;
    const float isc = ep.inv_scale;
    const float zs = isc * s_eff;                             // z = acc * zs off the target column
    const float a = zs * LOG2E, b = -s_eff * LOG2E;           // exp(z - s_eff) = 2^(acc*a + b)
    const float lo = cos_lo(), hi = cos_hi();
    // four independent accumulator lanes: the dependent chains are SC / 4 long, not SC (latency, not issue, bounds
    // a lone epilogue warp); the summation order is still fixed, so results are bitwise reproducible
    float ce4[4] = {0.f, 0.f, 0.f, 0.f}, cq4[4] = {0.f, 0.f, 0.f, 0.f}, ct4[4] = {0.f, 0.f, 0.f, 0.f};
    float mn4[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < SC; j += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float tt = v[j + u];
        const float e = ex2_approx(fmaf(tt, a, b));
        ce4[u] += e;
        cq4[u] = fmaf(e, e, cq4[u]);
        ct4[u] += tt;
        mn4[u] = fminf(mn4[u], tt);
        mx4[u] = fmaxf(mx4[u], tt);
      }
    }
    const float ce = (ce4[0] + ce4[1]) + (ce4[2] + ce4[3]);
    const float ce2 = (cq4[0] + cq4[1]) + (cq4[2] + cq4[3]);
    const float ct = (ct4[0] + ct4[1]) + (ct4[2] + ct4[3]);
    const float tmn = fminf(fminf(mn4[0], mn4[1]), fminf(mn4[2], mn4[3]));
    const float tmx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    const bool has_t = (st.tgt >= cls0) && (st.tgt < cls0 + SC);
    // The whole slice goes the element-wise way only when something in it needs the clamp or the NaN scrub (or the slice is
    // ragged).  A target column alone does not: the slice is summed as if it had no margin and the ONE element is then
    // exchanged by the lane that owns it.  (Round 1 sent every slice with a target down the element-wise path -- 1 % of
    // the slices at cfg3, but every second tile then waited for a warp in it: K2 61.9 us against 49.9 us with the labels
    // moved out of range, tools/careful_probe.py.)
    bool careful = !(s_eff > 0.f) || (cc < SC) || !(tmx * isc <= hi) || !(tmn * isc >= lo) ||
                   !isfinite(ct) || !isfinite(ce) || (has_t && ep.whole_slice_targets);
    careful = __any_sync(0xffffffffu, careful);               // warp stays convergent for the next tcgen05.ld
    if (!careful) {
      st.sumexp += ce; st.sumexp2 += ce2; st.sumz = fmaf(ct, zs, st.sumz);
      st.cmin = fminf(st.cmin, tmn * isc); st.cmax = fmaxf(st.cmax, tmx * isc);
      if (!has_t) {
        const float bz = tmx * zs;
        if (bz > st.best) {                                   // rare after the first few slices
          st.best = bz;
          int tix = SC - 1;
#pragma unroll
          for (int j = SC - 2; j >= 0; --j) if (v[j] == tmx) tix = j;   // first index of the maximum
          st.bestidx = cls0 + tix;
        }
      } else {
        // my target column sits in this slice: take its margin-free contribution out, put phi(cos) in (:363-427)
        const int jt = st.tgt - cls0;
        float at = 0.f;
#pragma unroll
        for (int j = 0; j < SC; ++j) if (j == jt) at = v[j];
        const float e_old = ex2_approx(fmaf(at, a, b));       // the bits the sums above hold for this element
        const float z_old = at * zs;
        float z_new = head_phi_once(ep.hm, at * isc) * s_eff;  // inside [lo, hi]: no clamp to apply
        if (!isfinite(z_new)) { z_new = 0.f; st.saw_nan = true; }
        const float e_new = exp2f((z_new - s_eff) * LOG2E);
        st.sumexp += e_new - e_old;
        st.sumexp2 += fmaf(e_new, e_new, -(e_old * e_old));
        st.sumz += z_new - z_old;
        st.ztgt = z_new;
        // the row maximum with the exchanged element, first index wins (the element-wise order)
#pragma unroll
        for (int j = 0; j < SC; ++j) {
          const float z = (j == jt) ? z_new : v[j] * zs;
          if (z > st.best) { st.best = z; st.bestidx = cls0 + j; }
        }
      }
    } else {
      // the margin touches ONE element of the slice: evaluate phi once (its acos / cos code is large; 32 inlined
      // copies per slice made the kernel instruction-fetch bound)
      float tphi = 0.f;
      if (has_t) {
        float ct = 0.f;
#pragma unroll
        for (int j = 0; j < SC; ++j) if (cls0 + j == st.tgt) ct = v[j] * isc;
        tphi = ep.hm.phi((ct != ct) ? ct : fminf(fmaxf(ct, lo), hi));
      }
#pragma unroll
      for (int j = 0; j < SC; ++j) {
        if (j < cc) {
          const float cosv = v[j] * isc;
          st.cmin = fminf(st.cmin, cosv); st.cmax = fmaxf(st.cmax, cosv);
          const float c = (cosv != cosv) ? cosv : fminf(fmaxf(cosv, lo), hi);
          const bool is_t = (cls0 + j == st.tgt);
          const float tv = is_t ? tphi : c;
          float z = tv * s_eff;
          if (!isfinite(z)) { z = 0.f; st.saw_nan = true; }
          if (is_t) st.ztgt = z;
          const float e = exp2f((z - s_eff) * LOG2E);
          st.sumexp += e; st.sumexp2 = fmaf(e, e, st.sumexp2); st.sumz += z;
          if (z > st.best) { st.best = z; st.bestidx = cls0 + j; }
        }
      }
    }
    (void)it;
  }

  static __device__ __forceinline__ void item_end(State& st, const Params& ep, const XwParams& p, const XwItem& it,
                                                  float* scratch) {
    // the two column halves of a row live in two warps: the upper half parks its record in shared memory
    float* slot = scratch + (it.quad * 32 + it.lane) * 8;
    if (it.half == 1) {
      slot[0] = st.sumexp; slot[1] = st.sumexp2; slot[2] = st.ztgt; slot[3] = st.sumz; slot[4] = st.best;
      reinterpret_cast<int*>(slot)[5] = st.bestidx;
    }
    epi_bar_sync(it.grp);
    if (it.half == 0 && st.row_ok) {
      const float ob = slot[4];
      const int oi = reinterpret_cast<const int*>(slot)[5];
      float best = st.best; int bi = st.bestidx;
      if (oi >= 0 && (bi < 0 || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
      float* dst = ep.part + (it.row * (p.n_chunks * EG) + it.chunk * EG + it.grp) * PART_COLS;
      dst[0] = st.sumexp + slot[0]; dst[1] = st.sumexp2 + slot[1]; dst[2] = st.ztgt + slot[2];
      dst[3] = st.sumz + slot[3]; dst[4] = best;
      reinterpret_cast<int32_t*>(dst)[5] = (bi < 0) ? -1 : (int32_t)(ep.class_offset + bi);
    }
    float cmin = st.row_ok ? st.cmin : INFINITY, cmax = st.row_ok ? st.cmax : -INFINITY;
    cmin = warp_min(cmin); cmax = warp_max(cmax);
    if (it.lane == 0) {
      float* cp = ep.cos_part + 2 * (((int64_t)it.item * ep.pair + it.rank) * (XW_EPI_WARPS * EG) + it.grp * XW_EPI_WARPS + it.ew);
      cp[0] = cmin; cp[1] = cmax;
    }
    if (__any_sync(0xffffffffu, st.saw_nan) && it.lane == 0) atomicExch(ep.nan_flag, 1);
    epi_bar_sync(it.grp);
  }
};
using XwFwd = XwFwdT<1, 32>;        // one epilogue group, 32-column slices (round 1)
using XwFwd2 = XwFwdT<2, 16>;       // two epilogue groups (16 warps), 16-column slices
// K2 with K1(W) inside: one epilogue group on 16-column slices (<= 96 registers) + TEN prep warps = 20 warps.  The prep
// code is latency-bound scalar work (a 4-row trip takes 3-5 us under this kernel's memory traffic), so it needs warps: two
// beside the 32-column group gave K2 227 us, six at 128 registers 140 us.
using XwFwdP = XwFwdT<1, 16, 10>;

// -------------------------------------------------------------------------------------------------
// K3a, class-major (xw_kernel SWAP mode): the thread owns ONE class of the tile, the 32 columns of a slice are batch
// rows of the resident group.  Per-column quantities (-lse_b log2 e, the row's target class) sit in shared memory
// (item_begin); r_j = sum_i G_ij cos_ij is a private running sum; G leaves class-major, G^T[class][batch row], 64
// contiguous bytes per thread and slice.  Fast / careful slices as in XwFwd; phi / dphi are ONE out-of-line copy.
// (The batch-major predecessor needed a 31-shuffle butterfly per 32 classes for r and twice the instructions.)
// grad4 = {grad_scale, n, kappa, g_scale} from b200f_arcface_hook_scale; g_scale is the power of two that puts
// |grad_scale| * g_scale in (512, 1024], so a target-column entry (|p-q| <= 1, dphi <~ 30) stays below fp16 max and
// entries down to p ~ 1e-7 stay normal; the consumers divide it out again.
__device__ __noinline__ void head_phi_dphi(HeadMath hm, float c, float* phi, float* dphi) {
  *phi = hm.phi(c);
  *dphi = hm.dphi(c);
}

// (Measured and rejected: G^T through per-warp TMA tensor stores -- 2 KB staged slices, one ring stage given up --
//  ran within noise of the 32-byte global stores kept here, 2225 vs 2186 us on one rank's cfg4 step.  Removing the
//  stores altogether (k3a_ablate = 2) takes 313 us off that step although the kernel moves only 1.6 TB/s.)
// Probes that skip memory traffic (WRONG results) exist only in -DB200F_PROBES builds (tools/): the shipped library
// carries neither the branch nor the ABI to switch it on.
#ifdef B200F_PROBES
#define B200F_PROBE_FIELD int ablate;
#define B200F_PROBE_ON(ep, bit) (((ep).ablate & (bit)) != 0)
#else
#define B200F_PROBE_FIELD
#define B200F_PROBE_ON(ep, bit) false
#endif

// TMAST: the slice's packed words leave through a 2 KB staging buffer per warp (one ring stage given up: 4 instead of 5) and
// ONE cp.async.bulk.tensor store per slice instead of two 32-byte stores per lane.  Why: a 32-byte-per-lane store touches 32
// lines; ncu had 14 % of the kernel's stall samples on the instructions that reuse such a store's registers -- it sits in the
// load/store queue for ~500 cycles before it has read them.  st.shared.v4 into the 64-byte swizzle is 4 wavefronts.
template <int EG, int SC, int SPLIT = 2, bool TMAST = false>
struct XwBwdGTT {
  static_assert(!TMAST || (EG == 1 && SC == 32 && SPLIT == 2), "the TMA-store form: one group of 8 warps, 32-column slices");
  static constexpr int kRingStages = TMAST ? XW_STAGES - 1 : XW_STAGES;
  static constexpr int kEpiGroups = EG;
  static constexpr int kSliceCols = SC;
  static constexpr int kEpiSplit = SPLIT;     // 4: sixteen warps of ONE group on every tile, a quarter of its columns each
  static constexpr int kGroupThreads = 32 * 4 * SPLIT;
  struct Params {
    const int64_t* label; const float* lse; const float* grad4;
    int64_t class_offset;       // global id of this launch's class 0
    HeadMath hm;
    float ls_eps, inv_Ctot, inv_scale;
    alignas(64) CUtensorMap tm_gt;   // TMAST: G^T [classes of this launch, B] fp16, box 32 batch rows x 32 classes, 64-byte swizzle
    uint16_t* GT; int64_t ldgt; // G^T[class of this launch, batch row]: element (c, b) at GT[(b / tn) * gstride + c * ldgt + b % tn] with tn
    int blocked_nb;             // > 0: G^T in [32 classes x 64 batch rows] blocks of 4 KB, blocked_nb = batch blocks per class block:
                                // element (c, b) at GT[(((c >> 5) * blocked_nb + (b >> 6)) << 11) + ((c & 31) << 6) + (b & 63)] -- what a
                                // warp writes in two consecutive slices is 4 KB contiguous, 128 bytes per lane.  Class rows beyond this
                                // launch inside its last block are written as ZEROS (the dx GEMM reads whole blocks).  0: the
                                // (gstride, ldgt) form below
    int64_t gstride;            // the rows of a resident group.  Plain row-major [c, b]: gstride = tn, ldgt = row pitch.  Per-group
                                // storage [group][c][tn] (batch > 512): gstride = classes * tn, ldgt = tn
    float* r_part; int64_t ldr; // [SPLIT * m_groups, ldr]: one partial per (row group, column half / quarter)
    int gt_hint;                // L2 policy of the G^T stores (K3b and K3c read them next): 0 none, 2 evict_last
    int whole_slice_targets;    // tunable "target_patch" = 0: a slice with a target element goes the element-wise way (round 1)
    int defer_targets;          // tunable "target_patch" = 2: target elements are queued and patched at the end of the item
    B200F_PROBE_FIELD           // probe builds only: 2 = no G^T stores (WRONG results)
  };
  // Deferred target patches: up to kQueue entries {accumulator, -lse log2 e, batch row, class row} per warp and item in the
  // group's scratch behind the two column tables.  Why: a patch costs the warp ~1.5 us (one TMEM column, the out-of-line
  // phi / dphi, exp2, a scalar store) while its tile's accumulator stage is held; 3.5 of them per CTA and item sat on the
  // critical path (K3a 65 us against 60 us with the labels out of range).  At the end of the item they run lane-parallel.
  static constexpr int kQueue = 8;
  // a / gq / lim / tab_s: per-item constants of the fast path, formed ONCE (item_begin).  Re-forming them in every slice
  // put an LDC -> FMUL -> FMUL chain in front of the slice's first FFMA: with two epilogue warps per scheduler nothing hid it
  // (ncu, B = 4096 x 125 k: 5 % of the kernel's stall samples on that one FMUL).
  struct State { float gs, r; int cls; bool row_ok; uint64_t pol; int nq; uint32_t q_s; float a, gq, lim; uint32_t tab_s; };

  static __device__ __forceinline__ void item_begin(State& st, const Params& ep, const XwParams& p, const XwItem& it,
                                                    float* scratch, int TN) {
    epi_bar_sync_n<kGroupThreads>(it.grp);                    // the previous item's readers (of this group) are done
    const int e = it.ew * 32 + it.lane;
    if (e < TN) {
      const int64_t b = (int64_t)it.group * TN + e;
      float bneg = -INFINITY; int lab = -1;                    // beyond the batch: p = 0, no target
      if (b < p.B) {
        bneg = -__ldg(ep.lse + b) * LOG2E;
        const int64_t tg = __ldg(ep.label + b) - ep.class_offset;
        if (tg >= 0 && tg < p.C) lab = (int)tg;
      }
      scratch[e] = bneg;
      reinterpret_cast<int*>(scratch)[TN + e] = lab;
    }
    epi_bar_sync_n<kGroupThreads>(it.grp);
    st.gs = __ldg(ep.grad4) * __ldg(ep.grad4 + 3);
    st.r = 0.f; st.cls = 0; st.row_ok = false;
    st.pol = l2_policy(ep.gt_hint);
    st.nq = 0;
    st.a = ep.inv_scale * ep.hm.s_eff * LOG2E;                // p = 2^(acc * a - lse log2 e)
    st.gq = st.gs * (ep.ls_eps * ep.inv_Ctot);
    st.lim = (ep.hm.s_eff > 0.f) ? cos_hi() / ep.inv_scale : -1.0f;   // |acc| above it (or NaN, or s_eff <= 0): element-wise path
    st.tab_s = smem_u32(scratch);
    st.q_s = smem_u32(scratch + 2 * TN) + (uint32_t)it.ew * (kQueue * 16);   // 2 * TN + 16 warps * 32 floats <= 1024 floats
  }
  static __device__ __forceinline__ void tile_begin(State& st, const Params&, const XwParams& p, const XwItem& it) {
    st.cls = (int)it.row; st.row_ok = it.row < p.C; st.r = 0.f;
  }
  // TMAST: 16 columns (two 16-byte chunks) of this lane's 64-byte row of the warp's staging box, in the 64-byte swizzle
  // (chunk c of row r sits at chunk c ^ ((r >> 1) & 3); the box starts on a 512-byte boundary)
  static __device__ __forceinline__ void stage_block(const XwItem& it, int h, const uint32_t (&w)[8]) {
    const uint32_t row_s = smem_u32(it.aux) + (uint32_t)it.lane * 64u;
    const uint32_t sw = ((uint32_t)it.lane >> 1) & 3u;
#pragma unroll
    for (int c = 0; c < 2; ++c)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                   ::"r"(row_s + ((((uint32_t)(h >> 3) + c) ^ sw) << 4)), "r"(w[4 * c]), "r"(w[4 * c + 1]), "r"(w[4 * c + 2]), "r"(w[4 * c + 3])
                   : "memory");
  }

  static __device__ __forceinline__ void slice(State& st, const Params& ep, const XwParams& p, const XwItem& it,
                                               float (&v)[SC], int col0, float* scratch) {
    // Fast path first, for every slice, in blocks of 16 columns: table loads -> exponentials -> fp16 pack -> ONE 32-byte store
    // per block, the r sum and the |max| test riding along.  Whether the slice needed the element-wise path is judged AFTER
    // its stores; if so (rare) that path runs over the same registers and overwrites them.  Why this order: with the test
    // in front of the stores all SC gradients were live next to the SC accumulators, ptxas gave the NEXT slice's
    // tcgen05.ld the same registers and could issue it only at the end of this slice -- its latency sat at the top of every
    // slice (ncu: 13 % of the slice's stall samples on the first instruction behind tcgen05.wait::ld).
    if constexpr (TMAST) {                                    // the previous slice's tensor store has READ the staging buffer
      if (it.lane == 0) tma_store_wait_read();
      __syncwarp();
    }
    const float a = st.a, gs = st.gs, gq = st.gq;
    const uint32_t tb_s = st.tab_s + (uint32_t)col0 * 4u;
    const int64_t b0 = (int64_t)it.group * p_tn(p) + col0;
#ifdef B200F_GT_BLOCKED_PROBE   // tools/ probe builds only (WRONG results for the consumers): a warp's slice = 2 KB contiguous
    uint16_t* const gdst = ep.GT + ((((int64_t)(st.cls >> 5) * (ep.ldgt >> 5)) + (b0 >> 5)) << 10) + (st.cls & 31) * 32;
#else
    uint16_t* const gdst = ep.blocked_nb > 0
        ? ep.GT + ((((int64_t)(st.cls >> 5) * ep.blocked_nb) + (b0 >> 6)) << 11) + ((st.cls & 31) << 6) + (b0 & 63)
        : ep.GT + (int64_t)it.group * ep.gstride + (int64_t)st.cls * ep.ldgt + col0;
#endif
    const bool cols_full = b0 + SC <= p.B;                     // warp-uniform
    // blocked layout: a lane whose class lies beyond the launch but inside its last 32-class block stores zeros
    const bool zero_row = ep.blocked_nb > 0 && !st.row_ok && (st.cls - it.lane) < p.C;
    float am4[4] = {0.f, 0.f, 0.f, 0.f};
    float r4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int h = 0; h < SC; h += 16) {
      uint32_t w1[8];
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float bb[4];                                           // -lse_b log2 e of the four batch rows (columns)
        // explicit ld.shared: through the generic pointer these compile to LD.E, which queues in the global load/store
        // path behind this kernel's own G^T stores
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(bb[0]), "=f"(bb[1]), "=f"(bb[2]), "=f"(bb[3]) : "r"(tb_s + (uint32_t)(h + j) * 4u));
        float g4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float tt = v[h + j + u];
          g4[u] = fmaf(gs, ex2_approx(fmaf(tt, a, bb[u])), -gq);
          r4[u] = fmaf(g4[u], tt, r4[u]);
        }
        // |max| of the four accumulators, NaN-propagating, two per instruction (FMNMX3)
        asm("max.NaN.abs.f32 %0, %0, %1, %2;" : "+f"(am4[(j >> 2) & 1]) : "f"(v[h + j]), "f"(v[h + j + 1]));
        asm("max.NaN.abs.f32 %0, %0, %1, %2;" : "+f"(am4[2 + ((j >> 2) & 1)]) : "f"(v[h + j + 2]), "f"(v[h + j + 3]));
        w1[j / 2] = pack_f16(g4[0], g4[1]);
        w1[j / 2 + 1] = pack_f16(g4[2], g4[3]);
      }
      if (zero_row) {
#pragma unroll
        for (int j = 0; j < 8; ++j) w1[j] = 0u;
      }
      if constexpr (TMAST) {
        stage_block(it, h, w1);
      } else
      if (cols_full && (st.row_ok || zero_row) && !(B200F_PROBE_ON(ep, 2) && w1[0] != 0x12345678u)) {
        if (ep.gt_hint) st_global_256_hint(gdst + h, w1[0], w1[1], w1[2], w1[3], w1[4], w1[5], w1[6], w1[7], st.pol);
        else st_global_256(gdst + h, w1[0], w1[1], w1[2], w1[3], w1[4], w1[5], w1[6], w1[7]);
      }
    }
    float amax = 0.f;
    asm("max.NaN.abs.f32 %0, %1, %2, %3;" : "=f"(amax) : "f"(am4[0]), "f"(am4[1]), "f"(am4[2]));
    asm("max.NaN.abs.f32 %0, %0, %1, %1;" : "+f"(amax) : "f"(am4[3]));
    // a target element sits in this slice iff one of its SC batch rows is labelled with one of the warp's 32 classes
    const int* tl = reinterpret_cast<const int*>(scratch) + p_tn(p) + col0;
    int lab_l;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(lab_l) : "r"(st.tab_s + (uint32_t)(p_tn(p) + col0 + (it.lane & (SC - 1))) * 4u));
    const int c_w0 = st.cls - it.lane;
    // Element-wise path only for slices that need the clamp (or whose columns run past the batch).  Target elements (a batch
    // row of this slice labelled with one of the warp's 32 classes: 0.5 % of the slices at cfg3, but 64 % of the tiles then
    // waited for a warp in that path -- K3a 71.5 us against 62.7 us with the labels out of range, tools/careful_probe.py) go
    // through the fast path like any other element and are PATCHED behind the slice's stores (below): the lane that checked
    // the column hands (column, class) to the lane that owns the class, which overwrites its one fp16 element and corrects
    // its r sum.  The patch sits after the hot code so that the phi / dphi call costs the slices without a target nothing.
    const bool lab_hit = it.lane < SC && lab_l >= c_w0 && lab_l < c_w0 + 32;
    bool careful = !(amax <= st.lim) || (lab_hit && ep.whole_slice_targets) || !cols_full;
    careful = __any_sync(0xffffffffu, careful);
    // columns of this slice whose label is one of the warp's classes (lane j < SC checked column j): patched after the stores
    unsigned hits = careful ? 0u : __ballot_sync(0xffffffffu, lab_hit);
    float racc = (r4[0] + r4[1]) + (r4[2] + r4[3]);
    if (careful) {                                            // the reference's element-wise sequence; overwrites the block stores
      const float s_eff = ep.hm.s_eff, isc = ep.inv_scale;
      const float lo = cos_lo(), hi = cos_hi();
      const float q_off = ep.ls_eps * ep.inv_Ctot;
      racc = 0.f;
#pragma unroll
      for (int h = 0; h < SC; h += 16) {
        float g[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float vj = v[h + j];
          const float cosv = vj * isc;
          const float c = (cosv != cosv) ? cosv : fminf(fmaxf(cosv, lo), hi);
          const bool is_t = st.row_ok && (tl[h + j] == st.cls);
          float tv = c, f = 1.0f;
          if (is_t) head_phi_dphi(ep.hm, c, &tv, &f);
          float z = tv * s_eff;
          if (!isfinite(z)) { z = 0.f; f = 0.f; }
          if (!(cosv >= lo && cosv <= hi)) f = 0.f;
          const float pr = exp2f(fmaf(z, LOG2E, scratch[col0 + h + j]));
          const float q = is_t ? (1.0f - ep.ls_eps) + q_off : q_off;
          g[j] = fminf(fmaxf(gs * (pr - q) * f, -65504.f), 65504.f);
          const float t = g[j] * vj;
          racc += (g[j] != 0.f && t == t) ? t : 0.f;
        }
        if constexpr (TMAST) {                                // over the fast path's words in the staging buffer
          uint32_t w1[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w1[j] = pack_f16(g[2 * j], g[2 * j + 1]);
          stage_block(it, h, w1);
        } else
        if (st.row_ok && !(B200F_PROBE_ON(ep, 2) && g[0] != 12345.678f)) {
          if (cols_full) {
            uint32_t w1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w1[j] = pack_f16(g[2 * j], g[2 * j + 1]);
            st_global_256(gdst + h, w1[0], w1[1], w1[2], w1[3], w1[4], w1[5], w1[6], w1[7]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (b0 + h + j < p.B) gdst[h + j] = (uint16_t)(pack_f16(g[j], 0.f) & 0xffff);
          }
        }
      }
    }
    st.r += racc;
    if constexpr (TMAST) {
      // one tensor store per slice: [32 batch rows x 32 classes] from the staging buffer; the map clips columns beyond the
      // batch and class rows beyond this launch
      fence_proxy_async();
      __syncwarp();
      const int row_w = st.cls - it.lane;                     // first class row of this warp in the tile
      if (it.lane == 0 && row_w < p.C && b0 < p.B && !B200F_PROBE_ON(ep, 2)) {
        if (ep.gt_hint) tma_store_2d_hint(&ep.tm_gt, it.aux, (int)b0, row_w, st.pol);
        else tma_store_2d(&ep.tm_gt, it.aux, (int)b0, row_w);
      }
      // a patch that is applied right away (not queued) writes behind this store: it has to be in memory first
      if (hits != 0u && !(ep.defer_targets && st.nq < kQueue)) { if (it.lane == 0) tma_store_wait_all(); __syncwarp(); }
    }
    while (hits) {                                            // warp-uniform, usually zero trips
      const int src = __ffs(hits) - 1;
      hits &= hits - 1;
      const int lab_s = __shfl_sync(0xffffffffu, lab_l, src);          // column src of the slice is labelled lab_s
      // the accumulator of (my class, that column) again from TMEM: the stage is still ours, and the cold code then does
      // not index the slice's registers (a select chain over v[] here cost the hot path ~3 us per kernel)
      const float vt = tmem_ld1(it.taddr0 + (uint32_t)(col0 + src));
      if (ep.defer_targets && st.nq < kQueue) {               // warp-uniform: queue it, patch at the end of the item
        if (st.row_ok && st.cls == lab_s) {                   // exactly one lane owns the class
          const int bt = it.group * p_tn(p) + col0 + src;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                       ::"r"(st.q_s + (uint32_t)st.nq * 16u), "r"(__float_as_uint(vt)), "r"(__float_as_uint(scratch[col0 + src])), "r"(bt), "r"(st.cls)
                       : "memory");
        }
        ++st.nq;
        continue;
      }
      if (st.row_ok && st.cls == lab_s) {
        const float s_eff = ep.hm.s_eff, isc = ep.inv_scale;
        const float q_off = ep.ls_eps * ep.inv_Ctot;
        const float bb = scratch[col0 + src];
        const float g_old = fmaf(gs, ex2_approx(fmaf(vt, a, bb)), -gq);    // the bits the fast path stored and summed
        float tv, f;
        head_phi_dphi(ep.hm, vt * isc, &tv, &f);              // inside [lo, hi]: no clamp to apply
        float z = tv * s_eff;
        if (!isfinite(z)) { z = 0.f; f = 0.f; }
        const float pr = exp2f(fmaf(z, LOG2E, bb));
        const float q = (1.0f - ep.ls_eps) + q_off;
        const float gn = fminf(fmaxf(gs * (pr - q) * f, -65504.f), 65504.f);
        st.r = fmaf(gn - g_old, vt, st.r);
        const int64_t bt = (int64_t)it.group * p_tn(p) + col0 + src;
        if (bt < p.B && !(B200F_PROBE_ON(ep, 2) && gn != 12345.678f))
          gdst[src] = (uint16_t)(pack_f16(gn, 0.f) & 0xffff);
      }
    }
  }

  static __device__ __forceinline__ void tile_end(State& st, const Params& ep, const XwParams&, const XwItem& it) {
    if (st.row_ok && ep.r_part != nullptr)
      ep.r_part[(int64_t)(it.group * SPLIT + it.half) * ep.ldr + st.cls] = st.r * ep.inv_scale;
  }

  // rows of the resident group = width of the column space = 128 * PAIR; carried in XwParams.tn
  static __device__ __forceinline__ int p_tn(const XwParams& p) { return p.tn; }
  // The queued target patches of this warp: lane e re-evaluates entry e (phi / dphi once for all of them), overwrites the
  // fp16 element the fast path stored, and lane 0 then corrects the r partials one after the other (two entries may share a
  // class: a fixed order keeps the sum reproducible).  All tile_end stores of the item precede this in program order and
  // __syncwarp orders them for the other lanes.
  static __device__ __forceinline__ void item_end_swap(State& st, const Params& ep, const XwParams& p, const XwItem& it, int) {
    if constexpr (TMAST) {                                    // this warp's tensor stores are in memory: the patches write behind them
      if (it.lane == 0) tma_store_wait_all();
      __syncwarp();
    }
    const int n = st.nq;
    if (n == 0) return;
    __syncwarp();
    float delta = 0.f; int cls = 0;
    if (it.lane < n) {
      uint32_t u0, u1, u2, u3;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "r"(st.q_s + (uint32_t)it.lane * 16u) : "memory");
      const float vt = __uint_as_float(u0), bb = __uint_as_float(u1);
      const int bt = (int)u2; cls = (int)u3;
      const float s_eff = ep.hm.s_eff, isc = ep.inv_scale, gs = st.gs;
      const float q_off = ep.ls_eps * ep.inv_Ctot;
      const float g_old = fmaf(gs, ex2_approx(fmaf(vt, isc * s_eff * LOG2E, bb)), -(gs * q_off));   // the bits the fast path stored and summed
      float tv, f;
      head_phi_dphi(ep.hm, vt * isc, &tv, &f);                // inside [lo, hi]: the slice took the fast path
      float z = tv * s_eff;
      if (!isfinite(z)) { z = 0.f; f = 0.f; }
      const float pr = exp2f(fmaf(z, LOG2E, bb));
      const float q = (1.0f - ep.ls_eps) + q_off;
      const float gn = fminf(fmaxf(gs * (pr - q) * f, -65504.f), 65504.f);
      delta = (gn - g_old) * vt * ep.inv_scale;
      if (bt < p.B && !(B200F_PROBE_ON(ep, 2) && gn != 12345.678f))
        ep.GT[ep.blocked_nb > 0 ? ((((int64_t)(cls >> 5) * ep.blocked_nb) + (bt >> 6)) << 11) + ((cls & 31) << 6) + (bt & 63)
                                : (int64_t)(bt / p_tn(p)) * ep.gstride + (int64_t)cls * ep.ldgt + bt % p_tn(p)] =
            (uint16_t)(pack_f16(gn, 0.f) & 0xffff);
    }
    if (ep.r_part != nullptr) {
      float* const rrow = ep.r_part + (int64_t)(it.group * SPLIT + it.half) * ep.ldr;
      for (int e = 0; e < n; ++e) {                           // warp-uniform
        const float d = __shfl_sync(0xffffffffu, delta, e);
        const int c = __shfl_sync(0xffffffffu, cls, e);
        if (it.lane == 0) {
          float v;
          asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(rrow + c) : "memory");
          rrow[c] = v + d;
        }
      }
    }
    st.nq = 0;
  }
};
using XwBwdGT = XwBwdGTT<1, 32>;
using XwBwdGT2 = XwBwdGTT<2, 16>;
using XwBwdGT4 = XwBwdGTT<1, 16, 4>;   // one group of sixteen warps, column quarters
using XwBwdGTS = XwBwdGTT<1, 32, 2, true>;   // G^T through staging + TMA tensor stores

// -------------------------------------------------------------------------------------------------
// ---- K3b, class-major: dW[c, d] = coef.x * (acc[c, d] - w_hat16[c, d] * coef.y) --------------------------------
// SWAP side of the kernel (G^T rows streamed as the A operand, x_hat resident MN-major as B): the thread owns ONE
// class row per tile and 128 of the group's 256 features; it stores 128 contiguous bytes per 32-feature slice.
//
// w_hat reaches the epilogue through TMA, not through the load/store unit.  Measured on B200
// (tools/microbench/store_bw.cu): row-per-lane 64-byte global loads deliver 2.0-2.2 TB/s however many warps issue
// them -- the 102 MB of w_hat alone would take 46 us -- and the feature-major variant (XwDw, coalesced 64 B requests)
// was bound by the same loads.  Here every epilogue warp runs a private two-deep pipeline of [32 classes x 32
// features] boxes (2 KB, one cp.async.bulk.tensor per slice, issued two slices ahead by lane 0) into staging space
// taken from the operand ring: G^T needs a third of the HBM rate, so 3 ring stages suffice and the other 32 KB are
// 8 warps x 2 buffers x 2 KB.
// TMAST (round 2): dW leaves through shared-memory staging and TMA tensor stores instead of row-per-lane 32-byte stores
// (tools/microbench/store_bw.cu: 4.35 TB/s for that pattern, 5.55 TB/s coalesced).  Per warp and slice: the [32 classes x 32
// features] fp32 block goes to a 4 KB staging buffer in the 128-byte swizzle (lane = class row, its 16-byte chunk c at
// chunk c ^ (row & 7): conflict-free st.shared.v4), fence.proxy.async, one cp.async.bulk.tensor store by lane 0; the
// buffer is reused once that store has READ it.  The 32 KB of staging come out of the ring (G^T gets 2 stages instead of
// 3, its tiles are prefetched into L2 two ahead) and the w_hat boxes move into the third stage, the scratch area (unused by
// this policy) and 8 KB of extra shared memory.
template <int EG, int SC, bool TMAST = false>
struct XwDwTT {
  static_assert(!TMAST || (EG == 1 && SC == 32), "the TMA-store form: one epilogue group, 32-feature slices");
  static constexpr int kRingStages = TMAST ? 2 : 3;
  static constexpr int kExtraSmem = TMAST ? 256 + 8192 + 256 : 0;   // pad to a 512-byte boundary (the boxes' 64-byte swizzle), boxes, barriers
  static constexpr int kEpiGroups = EG;
  static constexpr int kSliceCols = SC;
  static constexpr int kBoxBytes = 32 * SC * 2;       // [32 classes x SC features] fp16: 2 KB (SC = 32) or 1 KB (SC = 16)
  // boxes: warps 0-3 in ring stage 2, warps 4-5 in the scratch area, warps 6-7 in the extra bytes; staging: ring stages 3-4
  static __device__ __forceinline__ void aux_layout(XwItem& it, uint8_t* ring, uint8_t* scratch, uint8_t* extra) {
    const int w = it.ew;
    it.aux = (w < 4) ? ring + 2 * XW_TILE_BYTES + w * 4096 : (w < 6) ? scratch + (w - 4) * 4096 : extra + 256 + (w - 6) * 4096;
    it.stage = ring + 3 * XW_TILE_BYTES + w * 4096;
  }
  struct Params {
    alignas(64) CUtensorMap tm_wh;      // w_hat16 rows of this launch [classes, D], box SC features x 32 classes; 64-byte swizzle (SC = 32) or none (16)
    alignas(64) CUtensorMap tm_dw;      // TMAST: dW [classes of this launch, D] fp32, box 32 features x 32 classes, 128-byte swizzle
    CoefSrc coef; float* dw; int64_t c0; int ld;
    float* sq_part;                     // NULL, or [items * PAIR * EG * 8]: sum of dW^2 per (item, CTA, epilogue warp) -- the
                                        // ||dW||^2 clip_grad_norm_ needs (src/training.py:528-533) without a pass over dW
    int dw_hint, wh_hint;               // L2 policies: dW stores (nobody reads them in this step: 1 evict_first), w_hat boxes
    B200F_PROBE_FIELD                   // probe builds only: 1 = no w_hat loads, 2 = no dW stores (WRONG results)
  };
  struct State { float2 cf; float rp_next[8]; float inw_next; int64_t next_row; float inv_sg; int seq, n_seq, row0, row_step; bool row_ok;
                 uint64_t dw_pol, wh_pol; float sq; };

  // slice n of this warp's item = column slice (n % spt) of its (n / spt)-th OWN tile (with EG groups a group owns every
  // EG-th tile of the walk, starting at it.first_tile); buffer n & 1
  static __device__ __forceinline__ void issue(const State& st, const Params& ep, const XwParams& p, const XwItem& it, int n) {
    const int spt = (p.tn >> 1) / SC;
    if (n >= st.n_seq || B200F_PROBE_ON(ep, 1)) return;
    const int d0 = it.group * p.tn + it.half * (p.tn >> 1) + (n % spt) * SC;
    if (d0 >= p.B) return;                                    // ragged feature count: nothing there (the reader skips too)
    if (it.lane == 0) {
      uint64_t* bar = it.aux_bar + (n & 1);
      mbar_arrive_expect_tx(bar, kBoxBytes);
      const int row = st.row0 + (it.first_tile + EG * (n / spt)) * st.row_step;
      if (ep.wh_hint) tma_load_2d_hint(it.aux + (n & 1) * kBoxBytes, &ep.tm_wh, bar, d0, row, st.wh_pol);
      else tma_load_2d(it.aux + (n & 1) * kBoxBytes, &ep.tm_wh, bar, d0, row);
    }
  }
  static __device__ __forceinline__ void item_begin(State& st, const Params& ep, const XwParams& p, const XwItem& it,
                                                    float*, int) {
    const int t_begin = (int)((int64_t)it.chunk * p.n_tiles / p.n_chunks);
    const int t_end = (int)((int64_t)(it.chunk + 1) * p.n_tiles / p.n_chunks);
    const int own = (t_end - t_begin - it.first_tile + EG - 1) / EG;         // tiles of this item this group takes
    st.n_seq = (own > 0 ? own : 0) * ((p.tn >> 1) / SC);
    st.row0 = (p.reverse ? t_end - 1 : t_begin) * p.tn + it.rank * XW_WROWS + it.quad * 32;
    st.row_step = p.reverse ? -p.tn : p.tn;
    st.seq = 0; st.next_row = -1; st.row_ok = false; st.sq = 0.f;
    st.inv_sg = ep.coef.inv_sg();
    st.dw_pol = l2_policy(ep.dw_hint); st.wh_pol = l2_policy(ep.wh_hint);
    issue(st, ep, p, it, 0);
    issue(st, ep, p, it, 1);
  }
  static __device__ __forceinline__ void tile_begin(State& st, const Params& ep, const XwParams& p, const XwItem& it) {
    st.row_ok = it.row < p.C;
    if (!st.row_ok) return;
    // this tile's coefficients were REQUESTED one (own) tile ago (raw loads, consumed only here: the warp never waits for
    // them in the middle of a tile; forming the sum where the loads are issued blocked it once per partial)
    if (st.next_row == it.row)
      st.cf = make_float2(st.inw_next * st.inv_sg,
                          (ep.coef.n_rb <= 4) ? ((st.rp_next[0] + st.rp_next[1]) + st.rp_next[2]) + st.rp_next[3]
                                              : ((((((st.rp_next[0] + st.rp_next[1]) + st.rp_next[2]) + st.rp_next[3]) + st.rp_next[4]) +
                                                  st.rp_next[5]) + st.rp_next[6]) + st.rp_next[7]);   // CoefSrc::load's order
    else
      st.cf = ep.coef.load(it.row, ep.c0 + it.row, st.inv_sg);
    const int64_t nr = it.row + (int64_t)EG * st.row_step;
    st.next_row = -1;
    if (nr >= 0 && nr < p.C && ep.coef.n_rb <= 8) {
      st.next_row = nr;
#pragma unroll
      for (int rb = 0; rb < 8; ++rb)
        st.rp_next[rb] = (rb < ep.coef.n_rb) ? __ldg(ep.coef.r_part + (int64_t)rb * ep.coef.ldr + nr) : 0.f;
      st.inw_next = __ldg(ep.coef.inv_nw + ep.c0 + nr);
    }
  }
  static __device__ __forceinline__ void slice(State& st, const Params& ep, const XwParams& p, const XwItem& it,
                                               float (&v)[SC], int col0, float*) {
    const int n = st.seq++;
    const int d0 = it.group * p.tn + col0;                   // first of this slice's SC features
    if (d0 >= p.B) return;
    constexpr int NV = SC / 8;                                // 16-byte vectors of fp16 per lane and slice
    uint4 w[NV];
    if (B200F_PROBE_ON(ep, 1)) {
#pragma unroll
      for (int i = 0; i < NV; ++i) w[i] = make_uint4(0, 0, 0, 0);
    } else {
      const uint32_t b = (uint32_t)n & 1u;
      mbar_wait(it.aux_bar + b, (it.aux_phase >> b) & 1u);
      it.aux_phase ^= 1u << b;
      const uint32_t src = smem_u32(it.aux + b * kBoxBytes + it.lane * (SC * 2));
      // SC = 32: the box arrives in the 64-byte swizzle (16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3)): a lane
      // owns a 64-byte row, so eight lanes reading the same chunk of their rows would hit the same four banks four times over
      const uint32_t sw = (SC == 32) ? ((uint32_t)(it.lane >> 1) & 3u) : 0u;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[i].x), "=r"(w[i].y), "=r"(w[i].z), "=r"(w[i].w) : "r"(src + (((uint32_t)i ^ sw) << 4)) : "memory");
    }
    float o[SC];
    if constexpr (TMAST) {
#pragma unroll
      for (int j = 0; j < SC; ++j) o[j] = 0.f;                 // rows beyond C: staged as zeros, clipped by the tensor map
    }
    if (st.row_ok) {
      float* dst = ep.dw + (ep.c0 + it.row) * (int64_t)ep.ld + d0;
      const float cx = st.cf.x, cy = st.cf.y;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint32_t q[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&q[j]));
          o[i * 8 + j * 2] = cx * fmaf(-f.x, cy, v[i * 8 + j * 2]);
          o[i * 8 + j * 2 + 1] = cx * fmaf(-f.y, cy, v[i * 8 + j * 2 + 1]);
        }
      }
      if (ep.sq_part != nullptr) {                            // fixed order: slices in walk order, four chains per slice
        float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < SC; j += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) q4[u] = (d0 + j + u < p.B) ? fmaf(o[j + u], o[j + u], q4[u]) : q4[u];
        }
        st.sq += (q4[0] + q4[1]) + (q4[2] + q4[3]);
      }
      if constexpr (TMAST) {
        // (handled below, by the whole warp)
      } else
      if (B200F_PROBE_ON(ep, 2) && o[0] != 12345.678f) {
      } else if (d0 + SC <= p.B && (ep.ld & 7) == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (ep.dw_hint)
            st_global_256_hint(dst + i * 8, __float_as_uint(o[i * 8]), __float_as_uint(o[i * 8 + 1]), __float_as_uint(o[i * 8 + 2]),
                               __float_as_uint(o[i * 8 + 3]), __float_as_uint(o[i * 8 + 4]), __float_as_uint(o[i * 8 + 5]),
                               __float_as_uint(o[i * 8 + 6]), __float_as_uint(o[i * 8 + 7]), st.dw_pol);
          else
            st_global_256(dst + i * 8, __float_as_uint(o[i * 8]), __float_as_uint(o[i * 8 + 1]), __float_as_uint(o[i * 8 + 2]),
                          __float_as_uint(o[i * 8 + 3]), __float_as_uint(o[i * 8 + 4]), __float_as_uint(o[i * 8 + 5]),
                          __float_as_uint(o[i * 8 + 6]), __float_as_uint(o[i * 8 + 7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < SC; ++j)
          if (d0 + j < p.B) dst[j] = o[j];
      }
    } else {
      // a lane without a row still has to have RECEIVED its shared-memory reads before the buffer is handed back
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) acc ^= w[i].x;
      asm volatile("" :: "r"(acc) : "memory");
    }
    if constexpr (TMAST) {
      if (!(B200F_PROBE_ON(ep, 2) && o[0] != 12345.678f)) {
        if (it.lane == 0) tma_store_wait_read();              // the previous slice's store has read the staging buffer
        __syncwarp();
        const uint32_t sb = smem_u32(it.stage) + (uint32_t)it.lane * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                       ::"r"(sb + (uint32_t)((c ^ (it.lane & 7)) << 4)), "f"(o[4 * c]), "f"(o[4 * c + 1]), "f"(o[4 * c + 2]), "f"(o[4 * c + 3])
                       : "memory");
        fence_proxy_async();
        __syncwarp();
        const int64_t row_w = it.row - it.lane;               // first class row of this warp in the tile (row of the launch)
        if (it.lane == 0 && row_w < p.C) {
          if (ep.dw_hint) tma_store_2d_hint(&ep.tm_dw, it.stage, d0, (int)row_w, st.dw_pol);
          else tma_store_2d(&ep.tm_dw, it.stage, d0, (int)row_w);
        }
      }
    }
    // Refill only now: the stores above consumed w[], so every lane's reads of the buffer have completed.  (Issuing
    // the refill right after the ld.shared instructions let the TMA write overtake reads still queued in the
    // load/store unit behind this kernel's global stores: single 16-byte pieces of the NEXT box showed up.)
    __syncwarp();
    issue(st, ep, p, it, n + 2);
  }
  static __device__ __forceinline__ void tile_end(State&, const Params&, const XwParams&, const XwItem&) {}
  static __device__ __forceinline__ void item_end_swap(State& st, const Params& ep, const XwParams&, const XwItem& it, int pair) {
    if constexpr (TMAST) { if (it.lane == 0) tma_store_wait_all(); }      // this thread's tensor stores are in global memory
    if (ep.sq_part == nullptr) return;
    const float s = warp_sum(st.sq);
    if (it.lane == 0) ep.sq_part[(((int64_t)it.item * pair + it.rank) * EG + it.grp) * XW_EPI_WARPS + it.ew] = s;
  }
};
using XwDwT = XwDwTT<1, 32>;
using XwDwT2 = XwDwTT<2, 16>;
using XwDwTS = XwDwTT<1, 32, true>;   // dW through staging + TMA stores

}  // namespace umma
}  // namespace b200f
