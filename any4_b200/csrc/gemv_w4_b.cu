// B200 (sm_100a) weight-only 4-bit GEMV/small-batch GEMM, weight in the reference's "B" int4
// tensor-core layout (weightOnRight = true, the Any4Linear default).
//
// Replaces the reference path tinygemm_y_f16RM_x_f16RM_w_{int4,any4,mx4}TC ->
// tinygemm_m16n8k16_chunk_kernel<ALayout_RM, BLayout_TC_int4> (TinyGemm_int4.cu:294-548,
// TinyGemmImpl.cuh:23-345, MatrixLayoutB.cuh:686-1101, Dequantization.cuh:55-131).
// It is NOT a port of that gmem->register kernel.  Design (see DESIGN.md for the budget):
//
//  * "Lane per weight row".  A CTA owns 32 consecutive weight rows (4 n-tiles of the packed
//    layout, one contiguous run of bytes) and lane L of EVERY warp works on row L; the warps
//    split k.  The per-row 16-entry LUT is expanded once per CTA into a 256-entry *byte pair*
//    table  pair[b] = (LUT[b & 15], LUT[b >> 4])  stored bank-private (row L only ever touches
//    shared-memory bank L), so the two nibbles of a packed byte are dequantised by ONE
//    conflict-free LDS.32 whose address is ONE PRMT (table at a 64 KiB-aligned shared address,
//    256-byte entry pitch, so `byte << 8 | lane*4 | base` is a byte permute).
//  * Group scale/zero are applied with one fma.rn.{bf16,f16}x2 per pair - the same single
//    rounded FMA as the reference (MatrixLayoutB.cuh:1042-1046), so every dequantised weight
//    is bit-identical to the reference's.
//  * Weights stream HBM -> shared memory with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx, 4 KiB copies)
//    into a CTA-wide ring of 3 x 32 KiB stages fed by a dedicated producer warp; a stage = 128 k per consumer
//    warp, one full barrier per k half, one empty barrier per stage.  The CTA is persistent over row blocks.
//  * Lane L owns row L with bits 2 and 3 swapped and the odd tile of a tile pair is staged a few bytes further,
//    which makes the 16-byte weight loads bank-conflict free (row_of_lane / tile_off_t below).
//  * The dot products go to the tensor pipe (mma.sync m16n8k16, fp32 accumulate) even at m = 1
//    so the FMA pipe stays free for the dequant.  Because all 32 lanes hold DIFFERENT weight
//    rows, the activation operand is block-structured: x sits in k-slots {2q,2q+1,2q+8,2q+9}
//    of operand column/row q only, which makes output entry (g, q) the dot product of lane
//    4g+q's row.  At m = 1 the second half of the A fragment carries a second k-set of the
//    same rows (256 useful MACs per HMMA); for m > 1 the weights are the B fragment and four
//    activation rows ride in one HMMA.
//  * Split-k for small n uses a thread-block cluster and a DSMEM reduction (no workspace).
//  * Programmatic dependent launch: launch_dependents at entry, wait before the activations are read (before
//    anything is read unless the caller declared the weights static).
//  * Epilogue variants: plain store, store into every rank's symmetric buffer (row-sharded multi-GPU), and
//    silu(gate) * up over row-interleaved gate/up weights.
//
// Numerics: dequantised weights bit-identical to the reference; products exact; fp32
// accumulation (order differs from the reference, as allowed by SURVEY.md 3.6); one RN at
// the end.  Non-finite weights (mx4 exponent >= 254, Inf/NaN LUT or scale) would poison the up to three other
// rows that share an mma row with them (0 * NaN); every CTA checks its sums and recomputes its rows one by one
// (slow_rows) when one is non-finite, which confines the NaN to its row as the reference does.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace tg {
namespace {

}  // namespace (reopened below)
}  // namespace tg
#include "w4_common.cuh"
namespace tg {
using namespace w4;
namespace {

// Static description of the packed words one lane owns in a 16-byte "unit" of its row's
// stage slice: word i of unit u covers k-slot q and tile pair tp (k-tiles 2tp, 2tp+1 of the
// stage).  See the B int4 layout [n/8][k/(ik*16)][32][ik/2] (TinyGemmConvertB.cu:252-308).
template <int IK>
struct Geo;
template <>
struct Geo<4> {  // slice = [2 super-tiles][32 lanes][2 words]; row g at +g*32 inside each 256 B
  static constexpr int kRowStride = 32;
  __device__ static constexpr int unit_off(int u) { return (u >> 1) * 256 + (u & 1) * 16; }
  __device__ static constexpr int q(int u, int i) { return (u & 1) * 2 + (i >> 1); }
  __device__ static constexpr int tp(int u, int i) { return (u >> 1) * 2 + (i & 1); }
  // words i and i + 2 use the same tile pair and adjacent k-slots: their activations are 16 contiguous bytes
  static constexpr int kXPartner = 2;
  static constexpr int kStagger = 16;  // see tile_off(): rows 32 B apart, the partner tile fills the other 16-byte halves
  __device__ static int swz(int u, int) { return u; }
};
template <>
struct Geo<2> {  // slice = [4 super-tiles][32 lanes][1 word]; row g at +g*16 inside each 128 B
  static constexpr int kRowStride = 16;
  __device__ static constexpr int unit_off(int u) { return u * 128; }
  __device__ static constexpr int q(int, int i) { return i; }
  __device__ static constexpr int tp(int u, int) { return u; }
  static constexpr int kXPartner = 1;  // words (0,1) and (2,3) are adjacent k-slots of one tile pair
  static constexpr int kStagger = 64;  // rows 16 B apart: four rows fill 64 B, the partner tile the other 64
  __device__ static int swz(int u, int) { return u; }
};
template <>
struct Geo<8> {  // slice = [1 super-tile][32 lanes][4 words]; row g at +g*64
  static constexpr int kRowStride = 64;
  __device__ static constexpr int unit_off(int u) { return u * 16; }
  __device__ static constexpr int q(int u, int) { return u; }
  __device__ static constexpr int tp(int, int i) { return i; }
  static constexpr int kXPartner = 0;  // each word is a different tile pair: no 16-byte pairing
  // rows 64 B apart: lanes with k-slot id (lane & 3) >= 2 take their units one step ahead, so four rows of a tile
  // cover slots {u, u+4, u+1, u+5} (16-byte slots mod 128 B) and the partner tile, 32 B further, the other four.
  // The order may only depend on lane & 3: lanes with equal lane & 3 share the activation operand of an mma.
  static constexpr int kStagger = 32;
  __device__ static int swz(int u, int lane) { return (u + ((lane >> 1) & 1)) & 3; }
};

// Shared-memory bank conflicts of the 16-byte weight loads: a quarter-warp (8 lanes) is served per wavefront, and 8
// rows of ONE n-tile sit 16*IK/2 bytes apart, so rows j and j+4 collide.  Hence (a) lane L owns row
// row_of_lane(L) = L with bits 2 and 3 swapped: a quarter-warp then holds rows 4h..4h+3 of TWO tiles, and (b) the
// odd tile of each pair is staged kStagger bytes further, which puts its rows into the banks the even tile leaves free.
__device__ __forceinline__ int row_of_lane(int l) { return (l & 0x13) | ((l & 4) << 1) | ((l & 8) >> 1); }
template <int IK>
__device__ __forceinline__ uint32_t tile_off_t(int t) {
  return (uint32_t)t * kTileStageBytes + (uint32_t)((t + 1) >> 1) * Geo<IK>::kStagger;
}
constexpr uint32_t kStageBytesB = kStageBytes + 256u;             // room for the staggers, keeps stages 128-B aligned
constexpr uint32_t kDynSmemBytesB = kDynSmemBytes + (kStages + 1) * 256u;
static_assert(kDynSmemBytesB <= 232448u, "exceeds the 227 KiB opt-in shared memory of sm_100");

// ---------------------------------------------------------------------------------------
// Exact per-row fallback, taken only when a CTA produced a non-finite sum.  The block-structured
// mma operand relies on 0 * w == 0; an Inf/NaN weight (mx4 exponents >= 254, non-finite LUT or
// scale) breaks that for the rows sharing its mma row.  Recomputing the CTA's rows one at a time
// (decode -> FFMA, like the tensor core: exact products, fp32 accumulate) confines the non-finite
// value to the row that owns it, which is what the reference does
// (tests/tinygemm/test_tinygemm_mx4.py:443-506).  out[mi][row] fp32.
// ---------------------------------------------------------------------------------------
template <tg_dtype DT, int IK>
__device__ __noinline__ void slow_rows(const Params p, int row0, int rows_valid, int chunk_begin, int chunk_end,
                                       float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= kWarps) return;
  const int kb = chunk_begin * kChunkK, ke = min(chunk_end * kChunkK, p.k);
  const int n_groups = p.k >> p.glog2;
  constexpr int kWordsPerSuper = 2 * IK;  // words of one row per super-tile (16*IK k)
  const uint32_t* wq = reinterpret_cast<const uint32_t*>(p.w);
  for (int rr = warp * 2; rr < warp * 2 + 2; ++rr) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    if (rr < rows_valid && ke > kb) {
      const int row = row0 + rr;
      const int64_t tile_words = p.tile_stride / 4;
      const int ks0 = kb / (16 * IK);
      const int n_words = (ke - kb) / 8;
      for (int wn = lane; wn < n_words; wn += 32) {
        const int ks = ks0 + wn / kWordsPerSuper, rem = wn % kWordsPerSuper;
        const int q = rem / (IK / 2), j = rem % (IK / 2);
        const uint32_t w = wq[(int64_t)(row >> 3) * tile_words + ((int64_t)ks * 32 + 4 * (row & 7) + q) * (IK / 2) + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kk = (ks * IK + 2 * j + (i >> 2)) * 16 + 2 * q + (i & 1) + ((i >> 1) & 1) * 8;
          const uint32_t code = (w >> ((i >> 1) * 4 + (i & 1) * 16)) & 0xfu;
          const int gi = kk >> p.glog2;
          uint32_t szw;
          if (p.sz == nullptr) szw = e8m0_to_dt<DT>((uint32_t)p.exps[(int64_t)row * n_groups + gi]) | 0x80000000u;
          else szw = p.sz[(int64_t)gi * p.w_rows + row];
          const uint32_t v = p.lut[(int64_t)row * p.lut_stride + code];
          const uint32_t wd = fma2<DT>(v, szw & 0xffffu, szw >> 16) & 0xffffu;  // low half: the single-rounded FMA
          float wf;
          if constexpr (DT == TG_BF16) wf = __uint_as_float(wd << 16);
          else wf = __half2float(__ushort_as_half((unsigned short)wd));
          for (int mi = 0; mi < p.m; ++mi) {
            const uint16_t xv = p.x[(int64_t)mi * p.k + kk];
            float xf;
            if constexpr (DT == TG_BF16) xf = __uint_as_float((uint32_t)xv << 16);
            else xf = __half2float(__ushort_as_half(xv));
            a[mi] = fmaf(wf, xf, a[mi]);
          }
        }
      }
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      float v = a[mi];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) out[mi * 32 + rr] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// the kernel
//   M1 = true : exactly one activation row, weights are the mma A operand (two k-sets)
//   M1 = false: 1..4 activation rows, weights are the mma B operand
// grid = (G, splits), cluster = (1, splits, 1); block = 16 consumer warps + 1 producer warp.
// PERSISTENT over row blocks: CTA (s, y) handles row blocks s, s+G, s+2G, ... for its k split y.  The
// activations are staged once; the producer streams the weights of consecutive row blocks back to back
// through one ring, so a block's table / scale staging and the previous block's epilogue hide under the
// stream instead of costing a full kernel prologue + pipeline fill per 32 rows.
// ---------------------------------------------------------------------------------------
template <tg_dtype DT, int IK, bool M1, bool PEERS>
__device__ __forceinline__ void gemv_w4_b_body(const Params& p, const Peers& peers) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t dyn_base = smem_u32(smem_raw);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) trace_stamp(p, 0);
  // Programmatic dependent launch: let the next kernel of the stream get resident as soon as all our CTAs have
  // started; wait for the previous kernel before touching anything it may have produced.  Packed weights, LUT
  // and scales may be declared static by the caller (tg_set_static_weights), in which case only the activations
  // and the output are ordered behind the previous kernel and the weight stream starts immediately.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool static_w = (p.flags & 8) != 0;
  if (!static_w) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int split = blockIdx.y;                 // k split (rank in cluster)
  const int G = (int)gridDim.x;
  const int n_blk = p.blk_q + ((int)blockIdx.x < p.blk_r ? 1 : 0);   // row blocks of this CTA (>= 1)

  // k range of this CTA in 128-wide chunks; a stage is kWarps consecutive chunks (one per consumer warp)
  const int chunks_total = (p.k + kChunkK - 1) >> 7;
  const int chunk_begin = split * p.chunks_per_split;
  const int chunk_end = min(chunks_total, chunk_begin + p.chunks_per_split);
  const int n_stage_iters = (max(chunk_end - chunk_begin, 0) + kWarps - 1) >> 4;
  static_assert(kChunkK == 128 && kWarps == 16, "shifts above");
  const int n_groups = p.k >> p.glog2;

  // ---- shared memory carve-up (window addresses, see the constants above) ----
  // full barriers come in halves: [s][0] covers the first 1024 k of stage s (chunks of warps 0..7), [s][1] the rest
  // (warps 8..15), so that half of the warps start on the first 16 KiB of the stream instead of the first 32
  const uint32_t full_bar = dyn_base;
  const uint32_t empty_bar = dyn_base + 16u * kStages;
  const uint32_t low_base = dyn_base + kCtrlBytes;
  const uint32_t table_base = (low_base + 0xffffu) & ~0xffffu;
  const uint32_t x_base = table_base + 128u;
  const uint32_t high_base = table_base + kTableBytes;
  const int n_low = min(kStages, (int)((table_base - low_base) / kStageBytesB));
  const uint32_t sz_base = high_base + (uint32_t)(kStages - n_low) * kStageBytesB;
  const uint32_t red_base = sz_base + kSzBytes;
  auto stage_addr = [&](int s) -> uint32_t {
    return s < n_low ? low_base + (uint32_t)s * kStageBytesB : high_base + (uint32_t)(s - n_low) * kStageBytesB;
  };
  auto tile_off = [](int t) -> uint32_t { return tile_off_t<IK>(t); };

  // group scale/zero words of this CTA's k range (the host picks `splits` so that they fit kSzBytes)
  const int group_first = (chunk_begin * kChunkK) >> p.glog2;
  const int group_last = chunk_end > chunk_begin ? (min(chunk_end * kChunkK, p.k) - 1) >> p.glog2 : group_first;
  const int n_groups_cta = group_last - group_first + 1;
  const int sz_words = n_groups_cta * 32;
  const bool is_mx4 = (p.sz == nullptr);

  if (warp == kWarps) {
    // =========================== TMA producer warp: starts the weight stream immediately ===========================
    uint64_t pol = 0;
    auto issue_stage = [&](int rb, int j, int jj) {
      const int s = jj % kStages;
      const int tiles_valid = min(kRowsPerCta, p.w_rows - rb * kRowsPerCta) >> 3;
      const uint8_t* wsrc = p.w + (int64_t)(rb * 4) * p.tile_stride;
      const int c0 = chunk_begin + j * kWarps;
      const int k0 = c0 * kChunkK;
      const int kvalid = min(min(kStageK, (chunk_end - c0) * kChunkK), p.k - k0);
      const uint32_t bytes = (uint32_t)kvalid * 4u;  // per n-tile: 8 rows * kvalid / 2
      const uint32_t bytes0 = min(bytes, 4096u), bytes1 = bytes - bytes0;
      const uint32_t bar = full_bar + s * 16;
      mbar_expect_tx(bar, bytes0 * (uint32_t)tiles_valid);
      mbar_expect_tx(bar + 8, bytes1 * (uint32_t)tiles_valid);  // 0 bytes: the phase completes right away
      const uint32_t dst = stage_addr(s);
      // 4 KiB bulk copies: the size class that sustains full HBM rate (scripts/microbench/stream_bw.cu); the
      // first k half of all four tiles is requested before the second
      for (int t = 0; t < tiles_valid; ++t)
        bulk_g2s(dst + tile_off(t), wsrc + t * p.tile_stride + (int64_t)k0 * 4, bytes0, bar, pol);
      if (bytes1)
        for (int t = 0; t < tiles_valid; ++t)
          bulk_g2s(dst + tile_off(t) + 4096u, wsrc + t * p.tile_stride + (int64_t)k0 * 4 + 4096, bytes1, bar + 8, pol);
    };
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar + s * 16, 1);
        mbar_init(full_bar + s * 16 + 8, 1);
        mbar_init(empty_bar + s * 8, kWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      pol = l2_evict_first_policy();
      trace_stamp(p, 2);
      if (n_stage_iters > 0) issue_stage((int)blockIdx.x, 0, 0);
    }
    __syncwarp();
    // tell the consumers the barriers exist (they wait on named barrier 2 before their main loop)
    asm volatile("bar.arrive 2, %0;" ::"n"(kThreads) : "memory");
    if (lane == 0) {
      int jj = 0;
      for (int b = 0; b < n_blk; ++b) {
        const int rb = (int)blockIdx.x + b * G;
        for (int j = 0; j < n_stage_iters; ++j, ++jj) {
          if (jj == 0) continue;  // issued above
          // Pipeline fill: the first stage is requested alone.  All SMs start together, and with every stage of
          // every SM in flight at once the memory system serves them interleaved: the first stage then lands
          // only when (almost) everything has.  Requesting the rest once stage 0 is here gets the consumers
          // going ~1 us earlier, and they need > 1 us for a stage anyway.
          if (jj == 1) {
            if (p.flags & 1) mbar_wait(full_bar + 8, 0u);
            trace_stamp(p, 3);
          }
          if (jj >= kStages) mbar_wait(empty_bar + (jj % kStages) * 8, (uint32_t)(jj / kStages - 1) & 1u);
          issue_stage(rb, j, jj);
        }
      }
    } else {
      // lanes 1..31: warm L2 with the LUT rows and group words of this CTA's LATER row blocks (plain LSU
      // prefetches), so their staging does not pay a DRAM round trip behind the weight stream
      for (int b = 1; b < n_blk; ++b) {
        const int nrow0 = ((int)blockIdx.x + b * G) * kRowsPerCta;
        if (nrow0 + kRowsPerCta > p.w_rows) break;
        if (lane <= 8 && p.lut_stride) l2_prefetch_line(p.lut + (int64_t)nrow0 * p.lut_stride + (lane - 1) * 64);
        if (!is_mx4)
          for (int g = group_first + lane - 1; g <= group_last; g += 31)
            l2_prefetch_line(p.sz + (int64_t)g * p.w_rows + nrow0);
      }
    }
  } else {
    // =========================== consumers ===========================
    // small per-row-block tensors travel global -> registers -> shared memory; the loads for block b+1 are
    // issued during the last stage of block b
    uint4 lut0 = make_uint4(0, 0, 0, 0), lut1 = lut0;
    uint32_t lut_hi = 0;
    uint32_t psz[kPreSz];
    auto load_sz_word = [&](int row0, int i) -> uint32_t {  // word i = (group i / 32, row i % 32) of a row block
      const int gi = group_first + (i >> 5);
      const int row = min(row0 + row_of_lane(i & 31), p.w_rows - 1);  // staged in lane order
      if (is_mx4) return e8m0_to_dt<DT>((uint32_t)p.exps[(int64_t)row * n_groups + gi]) | 0x80000000u;  // zero = -0
      return p.sz[(int64_t)gi * p.w_rows + row];
    };
    auto load_block_regs = [&](int rb) {
      const int row0 = rb * kRowsPerCta;
      const int row = min(row0 + row_of_lane(lane), p.w_rows - 1);
      const uint16_t* lrow = p.lut + (int64_t)row * p.lut_stride;
      lut0 = *reinterpret_cast<const uint4*>(lrow);
      lut1 = *reinterpret_cast<const uint4*>(lrow + 8);
      lut_hi = (uint32_t)lrow[warp];  // T[w]: this warp builds the 16 table entries whose high nibble is w
#pragma unroll
      for (int i = 0; i < kPreSz; ++i) {
        const int w = (int)threadIdx.x + i * kConsumerThreads;
        psz[i] = w < sz_words ? load_sz_word(row0, w) : 0u;
      }
    };
    // pair table: entry e = hi*16 + lo of row L at table_base + e*256 + 4L;  warp w builds hi = w.
    // group words: sz_s[group - group_first][row]
    auto store_block_smem = [&](int rb) {
      const uint32_t tp_[8] = {lut0.x, lut0.y, lut0.z, lut0.w, lut1.x, lut1.y, lut1.z, lut1.w};
      const uint32_t dst = table_base + (uint32_t)(warp * 16) * 256u + 4u * lane;
#pragma unroll
      for (int lo = 0; lo < 16; ++lo) {
        // result = (T[lo], T[hi]): low half from the LUT pair register, high half = lut_hi's low half
        sts32(dst + (uint32_t)lo * 256u, prmt(tp_[lo >> 1], lut_hi, (lo & 1) ? 0x5432u : 0x5410u));
      }
#pragma unroll
      for (int i = 0; i < kPreSz; ++i) {
        const int w = (int)threadIdx.x + i * kConsumerThreads;
        if (w < sz_words) sts32(sz_base + (uint32_t)w * 4u, psz[i]);
      }
      for (int w = (int)threadIdx.x + kPreSz * kConsumerThreads; w < sz_words; w += kConsumerThreads)
        sts32(sz_base + (uint32_t)w * 4u, load_sz_word(rb * kRowsPerCta, w));
    };

    // ---- kernel prologue: issue the small global loads first (first block's LUT / group words, activations)
    load_block_regs((int)blockIdx.x);
    if (static_w) asm volatile("griddepcontrol.wait;" ::: "memory");  // activations come from the previous kernel
    // one x item = 4 k values of one tile = 8 staged bytes; items beyond k (tail of the last chunk) are zero
    const int item_begin = chunk_begin * (kChunkK >> 2);
    const int item_end = chunk_end * (kChunkK >> 2);
    const int item_valid_end = p.k >> 2;
    {
      // activations, permuted so that (x[16t+i], x[16t+i+8]) are adjacent:
      //   xp[16t + 2i] = x[16t + i], xp[16t + 2i + 1] = x[16t + i + 8], i = 0..7
      // linear byte offset o of row r lives at x_base + ((r*x_row_bytes + o) / 128) * 256 + (o % 128)
      auto put_x = [&](int r, int it, uint32_t x1, uint32_t x2) {
        const uint32_t o = (uint32_t)r * p.x_row_bytes + (uint32_t)(it - item_begin) * 8u;
        sts64(x_base + (o >> 7) * 256u + (o & 127u), prmt(x1, x2, 0x5410u), prmt(x1, x2, 0x7632u));
      };
      for (int r = 0; r < p.m; ++r) {
        const uint32_t* xr = reinterpret_cast<const uint32_t*>(p.x + (int64_t)r * p.k);
        for (int it = item_begin + (int)threadIdx.x; it < item_end; it += kConsumerThreads) {
          const int t = it >> 2, pp = it & 3;
          const bool ok = it < item_valid_end;
          put_x(r, it, ok ? xr[t * 8 + pp] : 0u, ok ? xr[t * 8 + 4 + pp] : 0u);
        }
      }
    }
    store_block_smem((int)blockIdx.x);
    if (threadIdx.x == 0) trace_stamp(p, 4);
    asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");  // 512 consumers + the producer warp's arrive
    if (threadIdx.x == 0) trace_stamp(p, 5);

    // ---- loop-invariant lane state ----
    const uint32_t lanebase = table_base | (uint32_t)(lane * 4);
    const int g_ = lane >> 2, q_ = lane & 3;
    // lanes that carry activations in the block-structured operand
    const bool set1 = (g_ == q_);          // lanes 0, 5, 10, 15
    const bool set2 = (g_ == q_ + 4);      // lanes 16, 21, 26, 31
    const uint32_t x_active = (set1 || set2) ? 1u : 0u;
    // M1: set1 lanes read tile t0, set2 lanes tile t0+1 (+32 B).  !M1: set1 rows {0,2}, set2 rows {1,3}
    uint32_t x_lane_off;
    if constexpr (M1) {
      x_lane_off = set2 ? 32u : 0u;
    } else {
      const uint32_t o = set2 ? (uint32_t)p.x_row_bytes : 0u;
      x_lane_off = (o >> 7) * 256u;  // x_row_bytes % 128 == 0
    }
    const uint32_t x_row2 = ((2u * (uint32_t)p.x_row_bytes) >> 7) * 256u;  // rows mi + 2 (!M1)
    const bool has_row01 = M1 ? true : (set1 || p.m > 1);
    const bool has_row23 = M1 ? false : (set1 ? p.m > 2 : p.m > 3);
    const uint32_t xa01 = (x_active && has_row01) ? 1u : 0u;
    const uint32_t xa23 = (x_active && has_row23) ? 1u : 0u;
    const int rlane = row_of_lane(lane);  // the weight row (within the block) this lane owns
    const uint32_t w_lane_off = tile_off(rlane >> 3) + (uint32_t)warp * kTileChunkBytes +
                                (uint32_t)(rlane & 7) * Geo<IK>::kRowStride;
    uint32_t xr0[4] = {0u, 0u, 0u, 0u}, xr1[4] = {0u, 0u, 0u, 0u};  // x fragments (stay zero on inactive lanes)
    // !M1: one set of operand registers for all words (rows mi: xa/xb, rows mi+2: xc/xd; second tile: ya..yd); they are
    // only ever written by predicated loads, so the inactive lanes keep the zeros the block structure needs
    uint32_t xa_ = 0u, xb_ = 0u, xc_ = 0u, xd_ = 0u, ya_ = 0u, yb_ = 0u, yc_ = 0u, yd_ = 0u;
    constexpr int kChains = 2;            // independent HMMA accumulation chains
    const int nj = M1 ? 1 : p.m;
    const int tj = threadIdx.x >> 5, trow = threadIdx.x & 31;  // epilogue thread (tj, trow) owns y[tj][row0 + trow]
    const int tlane = row_of_lane(trow);                       // the lane that owns row trow (the swap is an involution)
    float* const exch = reinterpret_cast<float*>(smem_raw + kExchOff);

    int jj = 0;  // stage counter across row blocks (ring position / parity)
    for (int b = 0; b < n_blk; ++b) {
      const int rb = (int)blockIdx.x + b * G;
      const int row0 = rb * kRowsPerCta;
      const int rows_valid = min(kRowsPerCta, p.w_rows - row0);  // multiple of 8
      float acc[kChains][4];
#pragma unroll
      for (int a = 0; a < kChains; ++a)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[a][i] = 0.f;

      for (int j = 0; j < n_stage_iters; ++j, ++jj) {
        const int s = jj % kStages;
        const int c = chunk_begin + j * kWarps + warp;  // this warp's chunk in stage j
        if (j == n_stage_iters - 1 && b + 1 < n_blk) load_block_regs(rb + G);  // next block's LUT / group words

        // group (scale, zero) of the four tile pairs (32 k each) of this chunk, from the staged words
        uint32_t s2[4], z2[4];
        {
          const int kc = min(c, chunk_end - 1) * kChunkK;
          if (p.glog2 >= 7) {  // one group covers the whole 128-k chunk
            const uint32_t v = lds32(sz_base + (uint32_t)(((kc >> p.glog2) - group_first) * 32 + lane) * 4u);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              s2[t] = prmt(v, v, 0x1010u);  // mx4 words carry zero = -0: fma(v, s, -0) == v * s incl. sign of zero
              z2[t] = prmt(v, v, 0x3232u);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int gi = ((kc + 32 * t) >> p.glog2) - group_first;
              const uint32_t v = lds32(sz_base + (uint32_t)(gi * 32 + lane) * 4u);
              s2[t] = prmt(v, v, 0x1010u);
              z2[t] = prmt(v, v, 0x3232u);
            }
          }
        }

        // one lane per warp polls (512 threads spinning on try_wait would compete with the TMA writes for the
        // shared-memory pipe); after the warp-level sync every lane observes the completed phase itself
        const uint32_t my_full = full_bar + s * 16 + (warp >> 3) * 8;
        if (lane == 0) mbar_wait(my_full, (uint32_t)(jj / kStages) & 1u);
        __syncwarp();
        while (!mbar_try(my_full, (uint32_t)(jj / kStages) & 1u)) {
        }
        if (threadIdx.x == 0 && jj < 4) trace_stamp(p, 6 + jj);
        // A chunk is always processed whole: beyond k the staged activations are zero, so whatever bytes the
        // stage holds there contribute 0 (finite weights; the non-finite case is handled after the loop).
        if (c < chunk_end && !(p.flags & 2)) {
          const uint32_t sbase = stage_addr(s) + w_lane_off;
          // x base for this chunk: tile t0 = 8c + 2tp ; byte offset 32*t0 -> piece (t0/4), within (t0%4)*32
          const uint32_t xc = x_base + (uint32_t)(c - chunk_begin) * 512u + x_lane_off;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int uu = Geo<IK>::swz(u, lane);  // this lane's u-th unit (bank-conflict-free order for ik = 8)
            const uint4 wv = lds128(sbase + Geo<IK>::unit_off(uu));
            const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
            if constexpr (M1) {
              // activations of the unit's four words: 16-byte loads where two words are adjacent in x
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t xo_i = xc + (uint32_t)((Geo<IK>::tp(u, i) >> 1) * 256 + (Geo<IK>::tp(u, i) & 1) * 64 +
                                                      Geo<IK>::q(uu, i) * 8);
                constexpr int P = Geo<IK>::kXPartner;
                if constexpr (P == 0) {
                  lds64_if(xr0[i], xr1[i], xo_i, x_active);
                } else if ((i % (2 * P)) < P) {
                  lds128_if(xr0[i], xr1[i], xr0[i + P], xr1[i + P], xo_i, x_active);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int q = Geo<IK>::q(uu, i);
              const int tp = Geo<IK>::tp(u, i);  // == tp(uu, i) whenever swz is not the identity (ik = 8: tp = i)
              const uint32_t w = ww[i];
              // byte0: tile 2tp (k0, k0+8)   byte2: tile 2tp (k0+1, k0+9)
              // byte1: tile 2tp+1 (k0, k0+8) byte3: tile 2tp+1 (k0+1, k0+9)
              uint32_t p0 = lds32(prmt(w, lanebase, 0x7604u));
              uint32_t p1 = lds32(prmt(w, lanebase, 0x7614u));
              uint32_t p2 = lds32(prmt(w, lanebase, 0x7624u));
              uint32_t p3 = lds32(prmt(w, lanebase, 0x7634u));
              p0 = fma2<DT>(p0, s2[tp], z2[tp]);
              p1 = fma2<DT>(p1, s2[tp], z2[tp]);
              p2 = fma2<DT>(p2, s2[tp], z2[tp]);
              p3 = fma2<DT>(p3, s2[tp], z2[tp]);
              // x for tile 2tp, slot q: bytes (tp/2)*256 + (tp%2)*64 + 8q of this chunk's x
              const uint32_t xo = xc + (uint32_t)((tp >> 1) * 256 + (tp & 1) * 64 + q * 8);
              if constexpr (M1) {
                // A = weights: a0/a2 = k-set 1 (tile 2tp), a1/a3 = k-set 2 (tile 2tp+1)
                mma16816<DT>(acc[i & 1], p0, p1, p2, p3, xr0[i], xr1[i]);
              } else {
                lds64_if(xa_, xb_, xo, xa01);
                lds64_if(xc_, xd_, xo + x_row2, xa23);
                lds64_if(ya_, yb_, xo + 32u, xa01);
                lds64_if(yc_, yd_, xo + 32u + x_row2, xa23);
                // tile 2tp: B = (byte0, byte2); A = x (a0,a2 rows mi, a1,a3 rows mi+2)
                mma16816<DT>(acc[i & 1], xa_, xc_, xb_, xd_, p0, p2);
                mma16816<DT>(acc[i & 1], ya_, yc_, yb_, yd_, p1, p3);
              }
            }
          }
        }

        // hand the stage back to the producer
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_bar + s * 8) : "memory");
      }
      if (threadIdx.x == 0 && b == 0) trace_stamp(p, 10);

      // ---- per-warp partial results -> red[warp][j][row] fp32 ----
#pragma unroll
      for (int a = 1; a < kChains; ++a)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[0][i] += acc[a][i];
      const uint32_t rbase = red_base + (uint32_t)warp * 512u;
      if constexpr (M1) {
        // valid: lanes q_<2: acc[0],acc[1] = rows 4g+2q_, 4g+2q_+1 (k-set 1); lanes q_>=2: acc[2],acc[3] = rows
        // 4g+2(q_-2), +1 (k-set 2)
        const int jq = q_ >> 1;
        const int r = 4 * g_ + 2 * (q_ & 1);
        sts32(rbase + (uint32_t)(jq * 32 + r) * 4u, __float_as_uint(jq ? acc[0][2] : acc[0][0]));
        sts32(rbase + (uint32_t)(jq * 32 + r + 1) * 4u, __float_as_uint(jq ? acc[0][3] : acc[0][1]));
      } else {
        // acc[0],acc[1] = C[g_][2q_, 2q_+1]: mi = g_/4, rows 4*(2q_)+g_%4 and 4*(2q_+1)+g_%4; acc[2],acc[3]: mi + 2
        const int mi = g_ >> 2, qq = g_ & 3;
        sts32(rbase + (uint32_t)(mi * 32 + 8 * q_ + qq) * 4u, __float_as_uint(acc[0][0]));
        sts32(rbase + (uint32_t)(mi * 32 + 8 * q_ + 4 + qq) * 4u, __float_as_uint(acc[0][1]));
        sts32(rbase + (uint32_t)((mi + 2) * 32 + 8 * q_ + qq) * 4u, __float_as_uint(acc[0][2]));
        sts32(rbase + (uint32_t)((mi + 2) * 32 + 8 * q_ + 4 + qq) * 4u, __float_as_uint(acc[0][3]));
      }
      // all warps are done with this block's table / group words, and all partials are visible
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");

      // ---- block sums: thread (tj, trow) adds the warps' partials in warp order ----
      float total = 0.f;
      if (threadIdx.x < 128) {
        if constexpr (M1) {
          if (tj == 0) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
              total += __uint_as_float(lds32(red_base + (uint32_t)w * 512u + (uint32_t)tlane * 4u));
              total += __uint_as_float(lds32(red_base + (uint32_t)w * 512u + (uint32_t)(32 + tlane) * 4u));
            }
          }
        } else {
#pragma unroll
          for (int w = 0; w < kWarps; ++w)
            total += __uint_as_float(lds32(red_base + (uint32_t)w * 512u + (uint32_t)(tj * 32 + tlane) * 4u));
        }
      }
      // next block's table and group words (their loads were issued during the last stage)
      if (b + 1 < n_blk) store_block_smem(rb + G);

      // non-finite sums (Inf/NaN weights) are recomputed row by row so they stay confined to their row;
      // the vote doubles as the barrier that publishes the next block's table
      const bool bad = threadIdx.x < 128 && tj < nj && trow < rows_valid && !(fabsf(total) <= 3.0e38f);
      uint32_t any_bad;
      asm volatile(
          "{ .reg .pred pi, po; setp.ne.u32 pi, %1, 0; barrier.cta.red.or.pred.aligned po, 1, %2, pi; selp.u32 %0, 1, 0, po; }"
          : "=r"(any_bad)
          : "r"(bad ? 1u : 0u), "n"(kConsumerThreads)
          : "memory");
      if (any_bad) {
        float* out = reinterpret_cast<float*>(smem_raw + (red_base - dyn_base));
        slow_rows<DT, IK>(p, row0, rows_valid, chunk_begin, chunk_end, out);
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
        if (threadIdx.x < 128) total = out[tj * 32 + trow];
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
      }

      if (threadIdx.x < 128) {
        if (p.splits == 1) {
          if (p.flags & 16) {  // (gate, up) row pairs -> silu(gate) * up
            const uint32_t mine = f32_to_dt<DT>(total);
            const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
            if (tj < nj && !(trow & 1) && trow < rows_valid)
              store_y<PEERS>(p, peers, (int64_t)tj * p.y_stride + ((row0 + trow) >> 1),
                             silu_mul_dt<DT>((uint16_t)mine, (uint16_t)other));
          } else if (tj < nj && trow < rows_valid) {
            store_y<PEERS>(p, peers, (int64_t)tj * p.y_stride + row0 + trow, f32_to_dt<DT>(total));
          }
        } else {
          exch[tj * 32 + trow] = total;  // one row block per CTA when k is split: exchanged after the loop
        }
      }
    }
    if (threadIdx.x == 0) trace_stamp(p, 11);
  }

  if (p.splits > 1) {
    // split-k: every CTA of the cluster has published its 32 x nj partials; rank 0 adds them in rank order
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    const int nj = M1 ? 1 : p.m;
    const int tj = threadIdx.x >> 5, trow = threadIdx.x & 31;
    const int row0 = (int)blockIdx.x * kRowsPerCta;
    const int rows_valid = min(kRowsPerCta, p.w_rows - row0);
    float* part = reinterpret_cast<float*>(smem_raw + kExchOff);
    if (cluster.block_rank() == 0 && threadIdx.x < 128 && tj < nj) {
      float sum = 0.f;
      for (unsigned r = 0; r < (unsigned)p.splits; ++r) sum += cluster.map_shared_rank(part, r)[tj * 32 + trow];
      if (p.flags & 16) {
        const uint32_t mine = f32_to_dt<DT>(sum);
        const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
        if (!(trow & 1) && trow < rows_valid)
          store_y<PEERS>(p, peers, (int64_t)tj * p.y_stride + ((row0 + trow) >> 1),
                         silu_mul_dt<DT>((uint16_t)mine, (uint16_t)other));
      } else if (trow < rows_valid) {
        store_y<PEERS>(p, peers, (int64_t)tj * p.y_stride + row0 + trow, f32_to_dt<DT>(sum));
      }
    }
    cluster.sync();  // keep remote shared memory alive until rank 0 has read it
  }
#ifdef TG_W4_TRACE
  if (threadIdx.x == 0) {
    trace_stamp(p, 12);
    if (p.trace != nullptr) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 15] = smid;
    }
  }
#endif
}

template <tg_dtype DT, int IK, bool M1>
__global__ void __launch_bounds__(kThreads, 1) gemv_w4_b_kernel(const Params p) {
  gemv_w4_b_body<DT, IK, M1, false>(p, Peers{});
}
// row-sharded variant: the epilogue stores into every rank's symmetric output buffer
template <tg_dtype DT, int IK, bool M1>
__global__ void __launch_bounds__(kThreads, 1) gemv_w4_b_peer_kernel(const Params p, const __grid_constant__ Peers peers) {
  gemv_w4_b_body<DT, IK, M1, true>(p, peers);
}

}  // namespace
namespace w4 {
bool g_pdl = true;             // tg_set_option: programmatic dependent launch
bool g_static_weights = false;  // tg_set_option: packed weights / LUT / scales never written by a preceding kernel
unsigned long long* g_trace_buf = nullptr;  // set by tg_debug_set_trace (not part of the public header)
int g_flags_env = -1;          // TG_W4_FLAGS (tuning / debug), read once
}  // namespace w4
namespace {

template <tg_dtype DT, int IK, bool M1>
int launch_one(const Params& p, const Peers& peers, int row_blocks, cudaStream_t st) {
  auto kern = gemv_w4_b_kernel<DT, IK, M1>;
  auto kern_peer = gemv_w4_b_peer_kernel<DT, IK, M1>;
  static thread_local bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmemBytesB) != cudaSuccess ||
        cudaFuncSetAttribute(kern_peer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmemBytesB) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%u) failed: %s", kDynSmemBytesB, cudaGetErrorString(cudaGetLastError()));
      return TG_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  // persistent: one CTA (or one k-split cluster) per SM, each walking over its share of the row blocks
  static thread_local int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  static const bool persist = getenv("TG_W4_PERSIST") == nullptr || atoi(getenv("TG_W4_PERSIST")) != 0;  // tuning knob
  // k split across a cluster: one row block per cluster (the DSMEM exchange after the block loop belongs to ONE block)
  const int slots = (!persist || p.splits > 1) ? row_blocks : n_sm;
  const int gx = row_blocks < slots ? row_blocks : slots;
  Params pp = p;
  pp.blk_q = row_blocks / gx;
  pp.blk_r = row_blocks % gx;
  cfg.gridDim = dim3((unsigned)gx, (unsigned)p.splits, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kDynSmemBytesB;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (p.splits > 1) {  // split-k: one cluster per row block
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = 1;
    attrs[na].val.clusterDim.y = (unsigned)p.splits;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_pdl) {  // programmatic dependent launch (see the kernel prologue)
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t e = peers.n > 0 ? cudaLaunchKernelEx(&cfg, kern_peer, pp, peers) : cudaLaunchKernelEx(&cfg, kern, pp);
  if (e != cudaSuccess) {
    set_error("gemv_w4_b launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return TG_ERR_CUDA;
  }
  count_launch();
  return TG_OK;
}

template <tg_dtype DT, int IK>
int launch_m(Params p, const Peers& peers0, int row_blocks, int64_t rows_x, const uint16_t* x, uint16_t* y,
             cudaStream_t st) {
  // activation rows are processed in passes of up to 4 (bounded by the staging area)
  const int cap = kMaxXBytes / p.x_row_bytes;  // >= 1 by the choice of `splits`
  const int per_pass = cap < 4 ? cap : 4;
  Peers peers = peers0;
  for (int64_t r0 = 0; r0 < rows_x; r0 += per_pass) {
    p.m = (int)((rows_x - r0) < per_pass ? (rows_x - r0) : per_pass);
    p.x = x + r0 * p.k;
    p.y = y + r0 * p.y_stride;
    for (int r = 0; r < peers0.n; ++r) peers.y[r] = peers0.y[r] + r0 * p.y_stride;
    int rc = (p.m == 1) ? launch_one<DT, IK, true>(p, peers, row_blocks, st) : launch_one<DT, IK, false>(p, peers, row_blocks, st);
    if (rc != TG_OK) return rc;
  }
  return TG_OK;
}

template <tg_dtype DT>
int launch_ik(const Params& p, const Peers& peers, int ik, int row_blocks, int64_t rows_x, const uint16_t* x,
              uint16_t* y, cudaStream_t st) {
  switch (ik) {
    case 2: return launch_m<DT, 2>(p, peers, row_blocks, rows_x, x, y, st);
    case 4: return launch_m<DT, 4>(p, peers, row_blocks, rows_x, x, y, st);
    case 8: return launch_m<DT, 8>(p, peers, row_blocks, rows_x, x, y, st);
  }
  set_error("B-layout int4 innerKTiles must be 2, 4 or 8 (got %d)", ik);
  return TG_ERR_INVALID_ARGUMENT;
}

}  // namespace

int launch_gemm_w4_rm_B(void* y, const void* x, const int32_t* w, const void* sz, const void* lut,
                        const uint8_t* exps, int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik,
                        tg_w4_format fmt, tg_dtype dt, const uint16_t* const_lut, cudaStream_t st,
                        void* const* y_peers, int n_peers, int64_t y_row_stride, int silu_pairs) {
  Params p{};
  Peers peers{};
  peers.n = n_peers;
  for (int r = 0; r < n_peers; ++r) peers.y[r] = static_cast<uint16_t*>(y_peers[r]);
  p.w = reinterpret_cast<const uint8_t*>(w);
  p.sz = (fmt == TG_W4_MX4) ? nullptr : reinterpret_cast<const uint32_t*>(sz);
  p.exps = (fmt == TG_W4_MX4) ? exps : nullptr;
  p.w_rows = (int)w_rows;
  p.k = (int)k;
  p.glog2 = group == 32 ? 5 : group == 64 ? 6 : group == 128 ? 7 : 8;
  p.tile_stride = 4 * k;
  p.y_stride = n_peers > 0 ? y_row_stride : silu_pairs ? w_rows / 2 : w_rows;

  if (fmt == TG_W4_ANY4_GLOBAL || fmt == TG_W4_ANY4_ROWWISE) {
    p.lut = reinterpret_cast<const uint16_t*>(lut);
    p.lut_stride = (fmt == TG_W4_ANY4_ROWWISE) ? 16 : 0;
  } else {
    p.lut = const_lut;  // int4 / mx4: constant table, same for every row
    p.lut_stride = 0;
  }

  const int row_blocks = (int)div_up(w_rows, kRowsPerCta);
  const int chunks = (int)div_up(k, kChunkK);
  // k is split across a thread-block cluster (a) when the row blocks alone cannot fill the machine and
  // (b) as far as needed for one CTA's activations and group words to fit its shared-memory staging areas
  int splits = 1;
  while (splits < 8 && row_blocks * splits * 2 <= 148 && chunks / (splits * 2) >= kWarps) splits *= 2;
  auto fits = [&](int sp) {
    const int64_t cps = div_up(chunks, sp);                         // chunks per split
    // groups touched by one split: exact when a group divides a chunk, else an upper bound
    const int64_t groups = group <= kChunkK ? cps * (kChunkK / group) : cps * kChunkK / group + 2;
    return cps * 256 <= kMaxXBytes && groups * 128 <= (int64_t)kSzBytes;
  };
  while (splits < 8 && !fits(splits)) ++splits;
  if (!fits(splits)) {
    set_error("k = %lld with group %d exceeds what one cluster can stage (k <= %d at this group size)", (long long)k,
              group, (int)(8 * (kSzBytes / 128 - 2) * group));
    return TG_ERR_UNSUPPORTED;
  }
  p.splits = splits;
  p.trace = g_trace_buf;
  if (g_flags_env < 0) g_flags_env = getenv("TG_W4_FLAGS") ? atoi(getenv("TG_W4_FLAGS")) : 0;  // tuning knob
  p.flags = g_flags_env | (g_static_weights ? 8 : 0) | (silu_pairs ? 16 : 0);
  p.chunks_per_split = (int)div_up(chunks, splits);
  p.x_row_bytes = p.chunks_per_split * 256;  // one split's activations, whole 128-k chunks

  if (dt == TG_BF16)
    return launch_ik<TG_BF16>(p, peers, ik, row_blocks, rows_x, (const uint16_t*)x, (uint16_t*)y, st);
  return launch_ik<TG_FP16>(p, peers, ik, row_blocks, rows_x, (const uint16_t*)x, (uint16_t*)y, st);
}

}  // namespace tg

// debug hook (scripts/trace_kernel.py): every following B-layout 4-bit launch stamps %globaltimer at its
// phase boundaries into buf[CTA][16]; pass nullptr to switch it off
namespace tg {
void set_w4_options(int pdl, int static_weights) {
  if (pdl >= 0) g_pdl = pdl != 0;
  if (static_weights >= 0) g_static_weights = static_weights != 0;
}
}  // namespace tg
extern "C" void tg_debug_set_trace(void* buf) { tg::w4::g_trace_buf = static_cast<unsigned long long*>(buf); }
