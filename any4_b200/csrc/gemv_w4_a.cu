// B200 (sm_100a) weight-only 4-bit GEMV, weight in the reference's "A" int4 tensor-core layout
// (weightOnRight = false, the Int4Linear default: y = (W x^T)^T).
//
// Replaces tinygemm_y_f16RM_x_f16RM_w_{int4,any4,mx4}TC with weightOnRight = false ->
// tinygemm_m16n8k16_chunk_kernel<ALayout_TC_int4, BLayout_RM> (TinyGemm_int4.cu:294-548, :530-541,
// MatrixLayoutA.cuh:375-816).  Same machinery as gemv_w4_b.cu (see there and DESIGN.md): persistent CTAs,
// a bulk-TMA ring fed by a producer warp, bank-private byte-pair tables addressed with one PRMT, one
// single-rounded fma.rn per pair, HMMA with a block-structured activation operand.  What differs is the
// packed layout [m/16][k/(ik*16)][32][ik] (TinyGemmConvertA.cu:226-285): one word holds one k-tile of the
// row PAIR (g, g+8) and each of its bytes is (row g, row g+8) at the SAME k:
//     byte0 = k0, byte1 = k0+8, byte2 = k0+1, byte3 = k0+9        (k0 = 16*tile + 2q)
// so
//  * a lane owns a row pair: the table entry of byte b is (LUT_g[b & 15], LUT_{g+8}[b >> 4]), the group
//    scale/zero registers hold (s_g, s_{g+8}) / (z_g, z_{g+8}) - still ONE lookup and ONE fma per byte;
//  * two PRMTs per byte pair transpose (g,g+8)@k0 / (g,g+8)@k0+8 into (g@k0, g@k0+8) / (g+8@k0, g+8@k0+8),
//    which are exactly the a0/a1 (rows r, r+8) registers of an mma A fragment: the lane's two rows ride in
//    the two row halves of the fragment and share the activation operand (256 useful MACs per HMMA);
//  * a CTA owns 32 weight rows = 2 m-tiles = 16 row pairs; lanes L and L+16 share a pair and split every
//    128-k chunk (first / second 64 k); their activations sit in operand columns q and q+4.
// One activation row per launch (rows_x > 1 loops over launches).
#include <cooperative_groups.h>

#include <cstdlib>

#include "w4_common.cuh"

namespace cg = cooperative_groups;

namespace tg {
using namespace w4;
namespace {

constexpr int kATileStageBytes = kStageK * 8;   // one m-tile (16 rows) x 2048 k = 16 KiB
constexpr int kATileChunkBytes = 1024;          // one m-tile per 128 k
static_assert(2 * kATileStageBytes == kStageBytes, "A and B kernels share the ring geometry");

// Where lane (pair g of its m-tile, k-half `sub`) finds its 16-byte units inside the m-tile's 1 KiB chunk, and
// which k-tile / k-slot each of a unit's four words covers.  tile index is relative to the chunk (0..7).
template <int IK>
struct GeoA;
template <>
struct GeoA<4> {  // chunk = [2 super-tiles][32 lanes][4 words]; lane t = 4g+q holds tiles 0..3 of slot q
  __device__ static constexpr int pair_off(int g, int sub) { return sub * 512 + g * 64; }
  __device__ static constexpr int unit_off(int u) { return u * 16; }
  __device__ static constexpr int q(int u, int) { return u; }
  __device__ static constexpr int tile(int sub, int, int i) { return sub * 4 + i; }
};
template <>
struct GeoA<2> {  // chunk = [4 super-tiles][32 lanes][2 words]; a unit = slots (2h, 2h+1) x tiles (0, 1) of one super-tile
  __device__ static constexpr int pair_off(int g, int sub) { return sub * 512 + g * 32; }
  __device__ static constexpr int unit_off(int u) { return (u >> 1) * 256 + (u & 1) * 16; }
  __device__ static constexpr int q(int u, int i) { return (u & 1) * 2 + (i >> 1); }
  __device__ static constexpr int tile(int sub, int u, int i) { return sub * 4 + (u >> 1) * 2 + (i & 1); }
};
template <>
struct GeoA<1> {  // chunk = [8 super-tiles][32 lanes][1 word]; a unit = the four slots of one tile
  __device__ static constexpr int pair_off(int g, int sub) { return sub * 512 + g * 16; }
  __device__ static constexpr int unit_off(int u) { return u * 128; }
  __device__ static constexpr int q(int, int i) { return i; }
  __device__ static constexpr int tile(int sub, int u, int) { return sub * 4 + u; }
};

// exact per-row fallback for non-finite sums (see gemv_w4_b.cu: slow_rows); out[row 0..31] fp32
template <tg_dtype DT, int IK>
__device__ __noinline__ void slow_rows_a(const Params p, int row0, int rows_valid, int chunk_begin, int chunk_end,
                                         float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= kWarps) return;
  const int kb = chunk_begin * kChunkK, ke = min(chunk_end * kChunkK, p.k);
  const int n_groups = p.k >> p.glog2;
  const uint32_t* wq = reinterpret_cast<const uint32_t*>(p.w);
  const int64_t tile_words = p.tile_stride / 4;
  for (int rr = warp * 2; rr < warp * 2 + 2; ++rr) {
    float a = 0.f;
    if (rr < rows_valid && ke > kb) {
      const int row = row0 + rr;
      const int g = row & 7, hi = (row >> 3) & 1;
      const int t0 = kb / 16, n_words = (ke - kb) / 16 * 4;  // 4 words (slots q) per k-tile for this row pair
      for (int wn = lane; wn < n_words; wn += 32) {
        const int tile = t0 + (wn >> 2), q = wn & 3;
        const int ks = tile / IK, i = tile % IK;
        const uint32_t w = wq[(int64_t)(row >> 4) * tile_words + ((int64_t)ks * 32 + 4 * g + q) * IK + i];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          // byte b: k offset 0, 8, 1, 9; low nibble = row g, high nibble = row g+8
          const int kk = tile * 16 + 2 * q + (b >> 1) + (b & 1) * 8;
          const uint32_t code = (w >> (b * 8 + hi * 4)) & 0xfu;
          const int gi = kk >> p.glog2;
          uint32_t szw;
          if (p.sz == nullptr) szw = e8m0_to_dt<DT>((uint32_t)p.exps[(int64_t)row * n_groups + gi]) | 0x80000000u;
          else szw = p.sz[(int64_t)gi * p.w_rows + row];
          const uint32_t v = p.lut[(int64_t)row * p.lut_stride + code];
          const uint32_t wd = fma2<DT>(v, szw & 0xffffu, szw >> 16) & 0xffffu;
          float wf, xf;
          const uint16_t xv = p.x[kk];
          if constexpr (DT == TG_BF16) {
            wf = __uint_as_float(wd << 16);
            xf = __uint_as_float((uint32_t)xv << 16);
          } else {
            wf = __half2float(__ushort_as_half((unsigned short)wd));
            xf = __half2float(__ushort_as_half(xv));
          }
          a = fmaf(wf, xf, a);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) out[rr] = a;
  }
}

// grid = (G, splits), cluster = (1, splits, 1); block = 16 consumer warps + 1 producer warp; persistent over row blocks
template <tg_dtype DT, int IK>
__global__ void __launch_bounds__(kThreads, 1) gemv_w4_a_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t dyn_base = smem_u32(smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool static_w = (p.flags & 8) != 0;
  if (!static_w) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int split = blockIdx.y;
  const int G = (int)gridDim.x;
  const int n_blk = p.blk_q + ((int)blockIdx.x < p.blk_r ? 1 : 0);

  const int chunks_total = (p.k + kChunkK - 1) >> 7;
  const int chunk_begin = split * p.chunks_per_split;
  const int chunk_end = min(chunks_total, chunk_begin + p.chunks_per_split);
  const int n_stage_iters = (max(chunk_end - chunk_begin, 0) + kWarps - 1) >> 4;
  const int n_groups = p.k >> p.glog2;

  // shared-memory carve-up: identical to gemv_w4_b.cu (w4_common.cuh)
  const uint32_t full_bar = dyn_base;
  const uint32_t empty_bar = dyn_base + 8u * kStages;
  const uint32_t low_base = dyn_base + kCtrlBytes;
  const uint32_t table_base = (low_base + 0xffffu) & ~0xffffu;
  const uint32_t x_base = table_base + 128u;
  const uint32_t high_base = table_base + kTableBytes;
  const int n_low = min(kStages, (int)((table_base - low_base) / kStageBytes));
  const uint32_t sz_base = high_base + (uint32_t)(kStages - n_low) * kStageBytes;
  const uint32_t red_base = sz_base + kSzBytes;
  auto stage_addr = [&](int s) -> uint32_t {
    return s < n_low ? low_base + (uint32_t)s * kStageBytes : high_base + (uint32_t)(s - n_low) * kStageBytes;
  };
  const int group_first = (chunk_begin * kChunkK) >> p.glog2;
  const int group_last = chunk_end > chunk_begin ? (min(chunk_end * kChunkK, p.k) - 1) >> p.glog2 : group_first;
  const int n_groups_cta = group_last - group_first + 1;
  const int sz_words = n_groups_cta * 32;
  const bool is_mx4 = (p.sz == nullptr);

  if (warp == kWarps) {
    // =========================== TMA producer warp ===========================
    uint64_t pol = 0;
    auto issue_stage = [&](int rb, int j, int jj) {
      const int s = jj % kStages;
      const int tiles_valid = min(kRowsPerCta, p.w_rows - rb * kRowsPerCta) >> 4;  // m-tiles (16 rows)
      const uint8_t* wsrc = p.w + (int64_t)(rb * 2) * p.tile_stride;
      const int c0 = chunk_begin + j * kWarps;
      const int k0 = c0 * kChunkK;
      const int kvalid = min(min(kStageK, (chunk_end - c0) * kChunkK), p.k - k0);
      const uint32_t bytes = (uint32_t)kvalid * 8u;  // per m-tile: 16 rows * kvalid / 2
      const uint32_t bar = full_bar + s * 8;
      mbar_expect_tx(bar, bytes * (uint32_t)tiles_valid);
      const uint32_t dst = stage_addr(s);
      for (int t = 0; t < tiles_valid; ++t)
        for (uint32_t o = 0; o < bytes; o += 4096u)
          bulk_g2s(dst + t * kATileStageBytes + o, wsrc + t * p.tile_stride + (int64_t)k0 * 8 + o, min(4096u, bytes - o),
                   bar, pol);
    };
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar + s * 8, 1);
        mbar_init(empty_bar + s * 8, kWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      pol = l2_evict_first_policy();
      if (n_stage_iters > 0) issue_stage((int)blockIdx.x, 0, 0);
    }
    __syncwarp();
    asm volatile("bar.arrive 2, %0;" ::"n"(kThreads) : "memory");
    if (lane == 0) {
      int jj = 0;
      for (int b = 0; b < n_blk; ++b) {
        const int rb = (int)blockIdx.x + b * G;
        for (int j = 0; j < n_stage_iters; ++j, ++jj) {
          if (jj == 0) continue;
          if (jj >= kStages) mbar_wait(empty_bar + (jj % kStages) * 8, (uint32_t)(jj / kStages - 1) & 1u);
          issue_stage(rb, j, jj);
        }
      }
    }
  } else {
    // =========================== consumers ===========================
    const int pair = lane & 15, sub = lane >> 4;          // row pair of the CTA, k-half of every chunk
    const int r_lo = (pair >> 3) * 16 + (pair & 7);       // CTA-relative rows of the pair
    const int r_hi = r_lo + 8;
    uint4 lut0 = make_uint4(0, 0, 0, 0), lut1 = lut0;
    uint32_t lut_hi = 0;
    uint32_t psz[kPreSz];
    auto load_sz_word = [&](int row0, int i) -> uint32_t {
      const int gi = group_first + (i >> 5);
      const int row = min(row0 + (i & 31), p.w_rows - 1);
      if (is_mx4) return e8m0_to_dt<DT>((uint32_t)p.exps[(int64_t)row * n_groups + gi]) | 0x80000000u;
      return p.sz[(int64_t)gi * p.w_rows + row];
    };
    auto load_block_regs = [&](int rb) {
      const int row0 = rb * kRowsPerCta;
      const uint16_t* llo = p.lut + (int64_t)min(row0 + r_lo, p.w_rows - 1) * p.lut_stride;
      const uint16_t* lhi = p.lut + (int64_t)min(row0 + r_hi, p.w_rows - 1) * p.lut_stride;
      lut0 = *reinterpret_cast<const uint4*>(llo);
      lut1 = *reinterpret_cast<const uint4*>(llo + 8);
      lut_hi = (uint32_t)lhi[warp];  // LUT_{g+8}[w]: this warp builds the entries whose high nibble is w
#pragma unroll
      for (int i = 0; i < kPreSz; ++i) {
        const int w = (int)threadIdx.x + i * kConsumerThreads;
        psz[i] = w < sz_words ? load_sz_word(row0, w) : 0u;
      }
    };
    auto store_block_smem = [&](int rb) {
      const uint32_t tp_[8] = {lut0.x, lut0.y, lut0.z, lut0.w, lut1.x, lut1.y, lut1.z, lut1.w};
      const uint32_t dst = table_base + (uint32_t)(warp * 16) * 256u + 4u * lane;
#pragma unroll
      for (int lo = 0; lo < 16; ++lo)
        sts32(dst + (uint32_t)lo * 256u, prmt(tp_[lo >> 1], lut_hi, (lo & 1) ? 0x5432u : 0x5410u));
#pragma unroll
      for (int i = 0; i < kPreSz; ++i) {
        const int w = (int)threadIdx.x + i * kConsumerThreads;
        if (w < sz_words) sts32(sz_base + (uint32_t)w * 4u, psz[i]);
      }
      for (int w = (int)threadIdx.x + kPreSz * kConsumerThreads; w < sz_words; w += kConsumerThreads)
        sts32(sz_base + (uint32_t)w * 4u, load_sz_word(rb * kRowsPerCta, w));
    };

    load_block_regs((int)blockIdx.x);
    if (static_w) asm volatile("griddepcontrol.wait;" ::: "memory");
    const int item_begin = chunk_begin * (kChunkK >> 2);
    const int item_end = chunk_end * (kChunkK >> 2);
    const int item_valid_end = p.k >> 2;
    {
      // activations, same permutation as the B kernel: xp[16t + 2i] = x[16t + i], xp[16t + 2i + 1] = x[16t + i + 8]
      const uint32_t* xr = reinterpret_cast<const uint32_t*>(p.x);
      for (int it = item_begin + (int)threadIdx.x; it < item_end; it += kConsumerThreads) {
        const int t = it >> 2, pp = it & 3;
        const bool ok = it < item_valid_end;
        const uint32_t x1 = ok ? xr[t * 8 + pp] : 0u, x2 = ok ? xr[t * 8 + 4 + pp] : 0u;
        const uint32_t o = (uint32_t)(it - item_begin) * 8u;
        sts64(x_base + (o >> 7) * 256u + (o & 127u), prmt(x1, x2, 0x5410u), prmt(x1, x2, 0x7632u));
      }
    }
    store_block_smem((int)blockIdx.x);
    asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");

    const uint32_t lanebase = table_base | (uint32_t)(lane * 4);
    const int g_ = lane >> 2, q_ = lane & 3;
    // activation carriers: column q of the operand for the first k-half (lanes with g == q), column q + 4 for
    // the second (g == q + 4); they read the words of their own half
    const bool set1 = (g_ == q_), set2 = (g_ == q_ + 4);
    const uint32_t x_active = (set1 || set2) ? 1u : 0u;
    const int xsub = set2 ? 1 : 0;
    const uint32_t w_lane_off = (uint32_t)(pair >> 3) * kATileStageBytes + (uint32_t)warp * kATileChunkBytes +
                                (uint32_t)GeoA<IK>::pair_off(pair & 7, 0) + (uint32_t)sub * 512u;
    uint32_t xr0[4] = {0u, 0u, 0u, 0u}, xr1[4] = {0u, 0u, 0u, 0u};
    const int trow = threadIdx.x & 31;
    float* const exch = reinterpret_cast<float*>(smem_raw + kExchOff);

    int jj = 0;
    for (int b = 0; b < n_blk; ++b) {
      const int rb = (int)blockIdx.x + b * G;
      const int row0 = rb * kRowsPerCta;
      const int rows_valid = min(kRowsPerCta, p.w_rows - row0);
      float acc[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[a][i] = 0.f;

      for (int j = 0; j < n_stage_iters; ++j, ++jj) {
        const int s = jj % kStages;
        const int c = chunk_begin + j * kWarps + warp;
        if (j == n_stage_iters - 1 && b + 1 < n_blk) load_block_regs(rb + G);

        // (s_g, s_g8) / (z_g, z_g8) for the two 32-k tile pairs of this lane's half chunk
        uint32_t s2[2], z2[2];
        {
          const int kc = min(c, chunk_end - 1) * kChunkK + sub * 64;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int gi = ((kc + 32 * t) >> p.glog2) - group_first;
            const uint32_t vlo = lds32(sz_base + (uint32_t)(gi * 32 + r_lo) * 4u);
            const uint32_t vhi = lds32(sz_base + (uint32_t)(gi * 32 + r_hi) * 4u);
            s2[t] = prmt(vlo, vhi, 0x5410u);
            z2[t] = prmt(vlo, vhi, 0x7632u);
          }
        }

        if (lane == 0) mbar_wait(full_bar + s * 8, (uint32_t)(jj / kStages) & 1u);
        __syncwarp();
        while (!mbar_try(full_bar + s * 8, (uint32_t)(jj / kStages) & 1u)) {
        }
        if (c < chunk_end && !(p.flags & 2)) {
          const uint32_t sbase = stage_addr(s) + w_lane_off;
          // x of this chunk: tile T (0..7 within the chunk) at byte offset 32*T -> piece T/4, within (T%4)*32
          const uint32_t xc = x_base + (uint32_t)(c - chunk_begin) * 512u;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            // ik = 4: a lane's four units are the four k-slots q of its pair, 64 B apart between pairs, so the 8
            // lanes of a quarter-warp hit only 2 bank groups (4-way conflict).  Lanes that share an operand
            // column (same lane & 3) must walk the units in the same order, but the order may differ between
            // columns: rotating by (lane & 3) >> 1 halves the conflict, and since only the k-slot (hence the x
            // offset) depends on the unit the rotation costs nothing.
            const int ur = (IK == 4) ? ((u + ((lane & 3) >> 1)) & 3) : u;
            const uint4 wv = lds128(sbase + (IK == 4 ? ur * 16 : GeoA<IK>::unit_off(u)));
            const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t w = ww[i];
              const int q = (IK == 4) ? ur : GeoA<IK>::q(u, i);
              const int tl = GeoA<IK>::tile(0, u, i);  // tile within this lane's half (0..3)
              // activation carriers read the word of THEIR half: tile = xsub*4 + tl
              const uint32_t xo = xc + (uint32_t)xsub * 256u + (uint32_t)(tl * 32 + q * 8);
              lds64_if(xr0[i], xr1[i], xo, x_active);
              uint32_t r0 = lds32(prmt(w, lanebase, 0x7604u));  // (g, g+8) @ k0
              uint32_t r1 = lds32(prmt(w, lanebase, 0x7614u));  // @ k0 + 8
              uint32_t r2 = lds32(prmt(w, lanebase, 0x7624u));  // @ k0 + 1
              uint32_t r3 = lds32(prmt(w, lanebase, 0x7634u));  // @ k0 + 9
              const int tp = tl >> 1;  // 32-k tile pair inside the half chunk -> group registers
              r0 = fma2<DT>(r0, s2[tp], z2[tp]);
              r1 = fma2<DT>(r1, s2[tp], z2[tp]);
              r2 = fma2<DT>(r2, s2[tp], z2[tp]);
              r3 = fma2<DT>(r3, s2[tp], z2[tp]);
              // transpose to row-pure registers: a0/a2 = row g, a1/a3 = row g+8
              const uint32_t a0 = prmt(r0, r1, 0x5410u), a1 = prmt(r0, r1, 0x7632u);
              const uint32_t a2 = prmt(r2, r3, 0x5410u), a3 = prmt(r2, r3, 0x7632u);
              mma16816<DT>(acc[i & 1], a0, a1, a2, a3, xr0[i], xr1[i]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_bar + s * 8) : "memory");
      }

      // per-warp partials: lane (g_, q_) holds C[g_][2q_, 2q_+1] (rows lo) and C[g_+8][..] (rows hi); valid columns
      // are n* = q_src (+4 for second-half lanes), i.e. source lane L = 4*g_ + (n* & 3)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[0][i] += acc[1][i];
      const uint32_t rbase = red_base + (uint32_t)warp * 256u;  // [2 halves][32 rows]
      {
        const bool upper = g_ >= 4;                       // source lanes of the second k-half
        const bool mine = upper ? (q_ >= 2) : (q_ < 2);   // this lane holds valid columns
        if (mine) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int n = 2 * q_ + e;                 // operand column
            const int L = 4 * g_ + (n & 3);           // lane that owns the pair
            const int pr = L & 15, hs = L >> 4;
            const int rl = (pr >> 3) * 16 + (pr & 7);
            sts32(rbase + (uint32_t)(hs * 32 + rl) * 4u, __float_as_uint(acc[0][e]));
            sts32(rbase + (uint32_t)(hs * 32 + rl + 8) * 4u, __float_as_uint(acc[0][2 + e]));
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");

      float total = 0.f;
      if (threadIdx.x < 32) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          total += __uint_as_float(lds32(red_base + (uint32_t)w * 256u + (uint32_t)trow * 4u));
          total += __uint_as_float(lds32(red_base + (uint32_t)w * 256u + (uint32_t)(32 + trow) * 4u));
        }
      }
      if (b + 1 < n_blk) store_block_smem(rb + G);
      const bool bad = threadIdx.x < 32 && trow < rows_valid && !(fabsf(total) <= 3.0e38f);
      uint32_t any_bad;
      asm volatile(
          "{ .reg .pred pi, po; setp.ne.u32 pi, %1, 0; barrier.cta.red.or.pred.aligned po, 1, %2, pi; selp.u32 %0, 1, 0, po; }"
          : "=r"(any_bad)
          : "r"(bad ? 1u : 0u), "n"(kConsumerThreads)
          : "memory");
      if (any_bad) {
        float* out = reinterpret_cast<float*>(smem_raw + (red_base - dyn_base));
        slow_rows_a<DT, IK>(p, row0, rows_valid, chunk_begin, chunk_end, out);
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
        if (threadIdx.x < 32) total = out[trow];
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
      }
      if (threadIdx.x < 32) {
        if (p.splits == 1) {
          if (trow < rows_valid) p.y[row0 + trow] = f32_to_dt<DT>(total);
        } else {
          exch[trow] = total;
        }
      }
    }
  }

  if (p.splits > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    const int trow = threadIdx.x & 31;
    const int row0 = (int)blockIdx.x * kRowsPerCta;
    const int rows_valid = min(kRowsPerCta, p.w_rows - row0);
    float* part = reinterpret_cast<float*>(smem_raw + kExchOff);
    if (cluster.block_rank() == 0 && threadIdx.x < 32) {
      float sum = 0.f;
      for (unsigned r = 0; r < (unsigned)p.splits; ++r) sum += cluster.map_shared_rank(part, r)[trow];
      if (trow < rows_valid) p.y[row0 + trow] = f32_to_dt<DT>(sum);
    }
    cluster.sync();
  }
}

template <tg_dtype DT, int IK>
int launch_a(const Params& p0, int row_blocks, int64_t rows_x, const uint16_t* x, uint16_t* y, cudaStream_t st) {
  auto kern = gemv_w4_a_kernel<DT, IK>;
  static thread_local bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmemBytes) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%u) failed: %s", kDynSmemBytes, cudaGetErrorString(cudaGetLastError()));
      return TG_ERR_CUDA;
    }
    attr_set = true;
  }
  static thread_local int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  Params p = p0;
  // k split across a cluster: one row block per cluster (the DSMEM exchange after the block loop belongs to ONE block)
  const int slots = p.splits > 1 ? row_blocks : n_sm;
  const int gx = row_blocks < slots ? row_blocks : slots;
  p.blk_q = row_blocks / gx;
  p.blk_r = row_blocks % gx;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)gx, (unsigned)p.splits, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kDynSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (p.splits > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = 1;
    attrs[na].val.clusterDim.y = (unsigned)p.splits;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  for (int64_t r = 0; r < rows_x; ++r) {  // one activation row per launch
    p.x = x + r * p.k;
    p.y = y + r * p.y_stride;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) {
      set_error("gemv_w4_a launch failed: %s", cudaGetErrorString(e));
      (void)cudaGetLastError();
      return TG_ERR_CUDA;
    }
    count_launch();
  }
  return TG_OK;
}

}  // namespace

int launch_gemm_w4_rm_A_generic(void* y, const void* x, const int32_t* w, const void* sz, const void* lut, const uint8_t* exps,
                                int lut_stride, int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik, bool mx4,
                                tg_dtype dt, cudaStream_t st);  // gemv_generic.cu

int launch_gemm_w4_rm_A(void* y, const void* x, const int32_t* w, const void* sz, const void* lut,
                        const uint8_t* exps, int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik,
                        tg_w4_format fmt, tg_dtype dt, const uint16_t* const_lut, cudaStream_t st) {
  Params p{};
  p.w = reinterpret_cast<const uint8_t*>(w);
  p.sz = (fmt == TG_W4_MX4) ? nullptr : reinterpret_cast<const uint32_t*>(sz);
  p.exps = (fmt == TG_W4_MX4) ? exps : nullptr;
  p.w_rows = (int)w_rows;
  p.k = (int)k;
  p.m = 1;
  p.glog2 = group == 32 ? 5 : group == 64 ? 6 : group == 128 ? 7 : 8;
  p.tile_stride = 8 * k;  // one m-tile: 16 rows * k / 2 bytes
  p.y_stride = w_rows;
  if (fmt == TG_W4_ANY4_GLOBAL || fmt == TG_W4_ANY4_ROWWISE) {
    p.lut = reinterpret_cast<const uint16_t*>(lut);
    p.lut_stride = (fmt == TG_W4_ANY4_ROWWISE) ? 16 : 0;
  } else {
    p.lut = const_lut;
    p.lut_stride = 0;
  }
  const int row_blocks = (int)div_up(w_rows, kRowsPerCta);
  const int chunks = (int)div_up(k, kChunkK);
  int splits = 1;
  while (splits < 8 && row_blocks * splits * 2 <= 148 && chunks / (splits * 2) >= kWarps) splits *= 2;
  auto fits = [&](int sp) {
    const int64_t cps = div_up(chunks, sp);
    const int64_t groups = group <= kChunkK ? cps * (kChunkK / group) : cps * kChunkK / group + 2;  // exact / bound
    return cps * 256 <= kMaxXBytes && groups * 128 <= (int64_t)kSzBytes;
  };
  while (splits < 8 && !fits(splits)) ++splits;
  if (!fits(splits))  // k too long for the staging areas of one cluster: the simple per-k-tile kernel takes any k
    return launch_gemm_w4_rm_A_generic(y, x, w, sz, p.lut, exps, p.lut_stride, rows_x, w_rows, k, group, ik, fmt == TG_W4_MX4, dt,
                                       st);
  p.splits = splits;
  p.chunks_per_split = (int)div_up(chunks, splits);
  p.x_row_bytes = p.chunks_per_split * 256;
  if (g_flags_env < 0) g_flags_env = getenv("TG_W4_FLAGS") ? atoi(getenv("TG_W4_FLAGS")) : 0;
  p.flags = g_flags_env | (g_static_weights ? 8 : 0);
  p.trace = nullptr;
  const uint16_t* xp = (const uint16_t*)x;
  uint16_t* yp = (uint16_t*)y;
#define TG_A(DTV, IKV) return launch_a<DTV, IKV>(p, row_blocks, rows_x, xp, yp, st)
  if (dt == TG_BF16) {
    if (ik == 1) TG_A(TG_BF16, 1);
    if (ik == 2) TG_A(TG_BF16, 2);
    if (ik == 4) TG_A(TG_BF16, 4);
  } else {
    if (ik == 1) TG_A(TG_FP16, 1);
    if (ik == 2) TG_A(TG_FP16, 2);
    if (ik == 4) TG_A(TG_FP16, 4);
  }
#undef TG_A
  set_error("A-layout int4 innerKTiles must be 1, 2 or 4 (got %d)", ik);
  return TG_ERR_INVALID_ARGUMENT;
}

}  // namespace tg
