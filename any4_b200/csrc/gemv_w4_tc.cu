// B200 (sm_100a) weight-only 4-bit GEMV / small-batch GEMM on the 5th-generation tensor cores, weight in the
// reference's "B" int4 tensor-core layout (weightOnRight = true, the Any4Linear default), 1..16 activation rows
// in ONE pass over the weights.
//
// Replaces the reference path tinygemm_y_f16RM_x_f16RM_w_{int4,any4,mx4}TC ->
// tinygemm_m16n8k16_chunk_kernel<ALayout_RM, BLayout_TC_int4> (TinyGemm_int4.cu:294-548, TinyGemmImpl.cuh:23-345,
// MatrixLayoutB.cuh:686-1101, Dequantization.cuh:55-131).  Not a port of that gmem->register mma.sync kernel:
//
//  * "Lane per weight row".  A CTA works on 32 consecutive weight rows (4 n-tiles of the packed layout); lane L of
//    every dequant warp owns row row_of_lane(L).  The row's 16-entry LUT is expanded once per row block into a
//    256-entry byte-PAIR table  pair[b] = (LUT[b & 15], LUT[b >> 4])  stored bank-private in shared memory
//    (entry e of lane L at e*256 + 4L), so both nibbles of a packed byte are dequantised by ONE conflict-free
//    LDS.32 whose address is ONE PRMT (the table base is an immediate).  Group scale/zero: one
//    fma.rn.{bf16,f16}x2 per pair - the same single-rounded FMA as the reference (MatrixLayoutB.cuh:1042-1046),
//    hence bit-identical dequantised weights.
//  * The dequantised pairs never touch shared memory or mma.sync registers: each thread writes its row's k-slice
//    straight into TENSOR MEMORY with tcgen05.st (32x32b: thread = TMEM lane = weight row, 8 columns = one K = 16
//    step), and ONE elected thread issues tcgen05.mma (M = 128, N = 4 * MB, K = 16, kind::f16, fp32 accumulators
//    in TMEM) with A read from TMEM and the activations read from shared memory through a K-major no-swizzle
//    descriptor.  The four warps of a "quad" own the four 32-lane TMEM sub-partitions; all of them work on the
//    SAME 32 weight rows but on different 128-wide k chunks, and the activation operand is block structured:
//    operand column 4*mi + j carries activation row mi for the k chunk of quarter j, so accumulator element
//    (lane 32j + L, column 4*mi + j) is the partial dot product of lane L's row over quarter j's chunks.  The other
//    columns of a lane hold cross terms nobody reads.  One pass therefore handles 16 activation rows (N = 64).
//  * Weights stream HBM -> shared memory with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx, 4 KiB copies,
//    L2 evict-first) into a 3 x 16 KiB ring fed by a producer warp.  The activations are staged by the dequant
//    threads themselves (global -> PRMT k-permutation -> the operand layout the descriptor describes): for one
//    activation row the whole x stays resident (k <= 14336), otherwise it is staged per ring stage, so any k works.
//  * Work units are ring stages (32 rows x 1024 k).  CTA i of G processes a contiguous range of the (row block,
//    stage) sequence (stream-K) or whole row blocks, whichever finishes earlier.  A row block cut by a range
//    boundary is finished by the LAST of its CTAs to arrive (fp32 partials in a device workspace, arrival counter,
//    summed in CTA order: deterministic).
//  * The decode CTA needs 113 KB of shared memory, 256 TMEM columns and 96 registers x 320 threads, so TWO CTAs - of
//    the same GEMV or of two consecutive GEMVs of a stream (programmatic dependent launch) - share an SM: the
//    prologue, first-byte latency and row-block boundaries of one hide under the dequant of the other.  With
//    TG_OPT_STATIC_WEIGHTS the weight stream, the table build and the dequant of the first TMEM slots start before
//    the previous kernel has finished; only the activation staging (and with it every MMA and store) waits.
//
// Numerics: dequantised weights bit-identical to the reference; products exact; fp32 accumulation in the tensor
// core (order differs from the reference, as allowed by SURVEY.md 3.6); one RN at the end.  A non-finite weight
// only reaches its own row's sums (the cross terms it pollutes are never read), as in the reference.
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "w4_common.cuh"

namespace tg {
namespace tc {
using namespace w4;

constexpr int kDqWarps = 8;                 // dequant warps: quad Q = warp / 4, quarter j = warp % 4
constexpr int kDqThreads = kDqWarps * 32;
constexpr int kWStages = 3;                 // weight ring depth
constexpr uint32_t kTileBytes = 4096;       // one n-tile (8 rows) x 1024 k of packed weights = one bulk copy
constexpr uint32_t kWStageStride = 4 * kTileBytes + 128;  // + room for the bank staggers of tiles 1..3
constexpr uint32_t kRegionA = 65536;        // pair table (even 128-B half-lines) + odd half-lines (see odd_hl)
constexpr uint32_t kWRingOff = kRegionA;
constexpr uint32_t kXDenseOff = kWRingOff + kWStages * kWStageStride;  // activation ring of the MB >= 8 kernels
constexpr uint32_t kSmemBase = 0x400;       // shared-window address of the dynamic shared memory (no static smem): see lds_table
constexpr int kCtrlHl = 240;                // odd half-lines 240.. hold the mbarriers and the TMEM base address
constexpr int kXResMaxStages = 14;          // resident activations: 16 half-lines per 1024 k -> k <= 14336
constexpr int kMaxGrid = 512;
constexpr int kWsPools = 4;
// Two scheduling experiments, measured on a B200 and left OFF (profiles/r2/EXPERIMENTS.md): neither moves the m = 1
// kernel (the SM is instruction-issue bound: a stall of one CTA is absorbed by the co-resident one) and the deferred
// epilogue costs the single-accumulator m = 16 kernel 20 %.
#ifndef TG_TC_EARLY_RELEASE
#define TG_TC_EARLY_RELEASE 0
#endif
#ifndef TG_TC_DEFER
#define TG_TC_DEFER 0
#endif
constexpr bool kEarlyRelease = TG_TC_EARLY_RELEASE != 0;  // hand a ring stage back as soon as its bytes are in registers
constexpr bool kDefer = TG_TC_DEFER != 0;                  // read a row block's sums one stage into the next row block

// split fix-up workspace: tagged fp32 partial sums [pool][CTA][slot 0/1][mi][row], one 8-byte word (value, launch tag) each
__device__ unsigned long long g_ws_partial[kWsPools][kMaxGrid * 2 * 16 * 32];

struct ParamsTC {
  const uint8_t* w;      // packed weight
  const uint16_t* x;     // [m][k]
  uint16_t* y;           // [m][y_stride]
  const uint32_t* sz;    // [k/g][w_rows] (scale, zero) pairs (unused for mx4)
  const uint8_t* exps;   // [w_rows][k/g] e8m0, mx4 only
  const uint16_t* lut;   // [16] or [w_rows][16]
  const uint8_t* xperm;  // MB >= 8: the activations pre-permuted into the operand layout, one kXStageBytes block per stage of
                         // the row (x_permute_kernel); the TMA producer brings each stage's block in.  null: register path
  unsigned long long* ws_partial;  // this launch's workspace pool
  uint32_t ws_tag;                  // launch tag carried by every partial of this launch
  int64_t tile_stride;   // bytes between consecutive n-tiles of the packed weight = 4 * k
  int64_t y_stride;
  int lut_stride;        // 0 or 16
  int m;                 // activation rows of this launch (<= MB)
  int w_rows;            // padded weight rows (multiple of 8)
  int k;
  int glog2;             // log2(group)
  int chunks_per_row;    // C = ceil(k / 128)
  int stages_per_row;    // S = ceil(C / 8): work units per row block
  int ug;                // stages per scheduling unit: 1 (stream-K over stages) or S (whole row blocks per CTA)
  int cq, cr;            // scheduling units per CTA: U / G and U % G
  int flags;             // bit 3: static weights; bit 4: silu(gate)*up over interleaved row pairs
#ifdef TG_W4_TRACE
  unsigned long long* trace;  // [CTAs][64] globaltimer stamps (scripts/trace_tc.py)
#endif
};
#ifdef TG_W4_TRACE
#define TC_TRACE(slot)                                                           \
  do {                                                                           \
    if (p.trace != nullptr) {                                                    \
      unsigned long long t__;                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                    \
      p.trace[(size_t)blockIdx.x * 64 + (slot)] = t__;                           \
    }                                                                            \
  } while (0)
#else
#define TC_TRACE(slot) \
  do {                 \
  } while (0)
#endif

// ---- geometry of the packed B int4 layout [n/8][k/(ik*16)][32][ik/2] (TinyGemmConvertB.cu:252-308) inside one
// 128-k chunk of one row: four 16-byte units of four words; word (tp, q) holds k-slot q (k0 = 2q) of tiles 2tp, 2tp+1:
//   byte 0 = tile 2tp (k0 | k0+8), byte 1 = tile 2tp+1 (k0 | k0+8), byte 2 = tile 2tp (k0+1 | k0+9), byte 3 = tile 2tp+1
template <int IK>
struct Geo;
template <>
struct Geo<4> {  // chunk = [2 super-tiles][32 lanes][2 words]; row g at +g*32 inside each 256 B
  static constexpr int kRowStride = 32;
  static constexpr int kStagger = 16;
  __device__ static constexpr int unit_off(int u) { return (u >> 1) * 256 + (u & 1) * 16; }
  __device__ static constexpr int unit_of(int tp, int q) { return (tp >> 1) * 2 + (q >> 1); }
  __device__ static constexpr int word_of(int tp, int q) { return (q & 1) * 2 + (tp & 1); }
};
template <>
struct Geo<2> {  // chunk = [4 super-tiles][32 lanes][1 word]; row g at +g*16 inside each 128 B
  static constexpr int kRowStride = 16;
  static constexpr int kStagger = 64;
  __device__ static constexpr int unit_off(int u) { return u * 128; }
  __device__ static constexpr int unit_of(int tp, int) { return tp; }
  __device__ static constexpr int word_of(int, int q) { return q; }
};
template <>
struct Geo<8> {  // chunk = [1 super-tile][32 lanes][4 words]; row g at +g*64
  static constexpr int kRowStride = 64;
  static constexpr int kStagger = 32;
  __device__ static constexpr int unit_off(int u) { return u * 16; }
  __device__ static constexpr int unit_of(int, int q) { return q; }
  __device__ static constexpr int word_of(int tp, int) { return tp; }
};
// Bank conflicts of the 16-byte weight loads: a quarter-warp (8 lanes) is served per wavefront and the 8 rows of ONE
// n-tile sit 16*IK/2 bytes apart, so rows j and j+4 collide.  Hence lane L owns row L with bits 2 and 3 swapped (a
// quarter-warp holds rows 4h..4h+3 of TWO tiles) and the odd tile of each pair is staged kStagger bytes further.
// ik = 8 additionally rotates the unit order of half the lanes (rows 64 B apart).
__device__ __forceinline__ int row_of_lane(int l) { return (l & 0x13) | ((l & 4) << 1) | ((l & 8) >> 1); }
template <int IK>
__device__ __forceinline__ uint32_t tile_off(int t) {
  return (uint32_t)t * kTileBytes + (uint32_t)((t + 1) >> 1) * Geo<IK>::kStagger;
}

__device__ __forceinline__ uint32_t odd_hl(int hl) { return (uint32_t)hl * 256u + 128u; }
// mbarrier i (8 bytes each, 16 per half-line)
__device__ __forceinline__ uint32_t bar_off(int i) { return odd_hl(kCtrlHl + (i >> 4)) + (uint32_t)(i & 15) * 8u; }
enum : int { B_WFULL = 0, B_WEMPTY = 3, B_AFULL = 6, B_AEMPTY = 9, B_DFULL = 12, B_DEMPTY = 14, B_XEMPTY = 16, B_XFULL = 19, B_COUNT = 22 };
constexpr uint32_t kHolderOff = 242u * 256u + 128u;  // TMEM base address; +4: "this CTA is the last arriver" flag

template <int MB, bool XRES>
struct Cfg {
  static_assert(!XRES || MB == 4, "resident activations: decode kernel only");
  static constexpr int N = 4 * MB;                 // MMA N: operand column 4*mi + j
  static constexpr int NS = 3;                     // TMEM A slots (64 columns = one 128-k chunk per quarter), shared by the quads
  static constexpr int NB = MB == 16 ? 1 : 2;      // accumulator buffers
  static constexpr int NX = 3;                     // activation ring depth (streamed activations)
  static constexpr int XR = MB == 4 ? 8 : 2 * MB;  // register ring of pieces in flight: XR / pieces-per-stage stages ahead
  static constexpr int kWarps = 10;                // 8 dequant, TMA producer, MMA issuer
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kABase = NB * N;            // TMEM: accumulators first, then the A slots
  static constexpr int kRedHl0 = XRES ? 224 : (MB == 4 ? 192 : 0);  // reduction scratch [4][MB] half-lines
  static constexpr int kLutHl0 = XRES ? 243 : (MB == 4 ? 208 : 64);  // next row block's LUT rows (8 half-lines)
  static constexpr uint32_t kXStageBytes = MB == 4 ? 64u * 256u : 32u * (MB / 2) * 128u;
  static constexpr uint32_t kSmem = MB == 4 ? kXDenseOff : kXDenseOff + NX * kXStageBytes;
  static constexpr int kMinBlocks = MB == 4 ? 2 : 1;
};
static_assert(Cfg<4, true>::kSmem <= 115712u, "two CTAs of the decode kernel must fit one SM");
static_assert(Cfg<16, false>::kSmem <= 232448u, "exceeds the 227 KiB opt-in shared memory of sm_100");
static_assert(Cfg<16, false>::kABase + 3 * 64 <= 256 && Cfg<8, false>::kABase + 3 * 64 <= 256, "TMEM budget");

// ---- small PTX wrappers ----
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// non-blocking probe (acquire).  A probe costs ~200 cycles of latency even on a completed phase
// (scripts/microbench/tc_probe2.cu), so the dequant loop probes the NEXT stage's barriers while it still has a stage
// of lookups to issue and only falls back to the blocking wait when the probe said "not yet".
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok;
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t addr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void tc_mma(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(d),
               "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred;
}
// bulk copy of data that other CTAs read too (the permuted activations): default L2 policy
__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// pair-table lookup: the dynamic shared window of a kernel without static shared memory starts at kSmemBase (checked
// at kernel entry), so the table base is an immediate and a lookup is PRMT + LDS.  Not volatile: a pure function of
// the index between two table builds (its index always comes from a weight load behind the build's barrier).
__device__ __forceinline__ uint32_t lds_table(uint32_t idx) {
  uint32_t v;
  asm("ld.shared.b32 %0, [%1+%2];" : "=r"(v) : "r"(idx), "n"(kSmemBase));
  return v;
}
template <bool PEERS>
__device__ __forceinline__ void store_out(uint16_t* y, const Peers& peers, int64_t idx, uint16_t v) {
  if constexpr (PEERS) {
#pragma unroll 1
    for (int r = 0; r < peers.n; ++r) peers.y[r][idx] = v;  // the same location in every rank's symmetric buffer
  } else {
    y[idx] = v;
  }
}

// CTA i of the grid owns stage units [begin, end): the first cr CTAs get cq + 1 scheduling units of ug stages
__device__ __forceinline__ void cta_range(const ParamsTC& p, int i, int& begin, int& end) {
  begin = (i * p.cq + min(i, p.cr)) * p.ug;
  end = begin + (p.cq + (i < p.cr ? 1 : 0)) * p.ug;
}
__device__ __forceinline__ int cta_of_unit(const ParamsTC& p, int u) {  // split mode only (ug == 1)
  const int big = p.cr * (p.cq + 1);
  return u < big ? u / (p.cq + 1) : p.cr + (u - big) / p.cq;
}

// Cursor over the CTA's (row block, stage-in-row) sequence
struct Cursor {
  int u, rb, sir;
  __device__ __forceinline__ void init(int u0, int S) {
    u = u0;
    rb = u0 / S;
    sir = u0 - rb * S;
  }
  __device__ __forceinline__ void next(int S) {
    ++u;
    if (++sir == S) {
      sir = 0;
      ++rb;
    }
  }
  __device__ __forceinline__ void skip(int n, int S) {  // n <= S - sir
    u += n;
    sir += n;
    if (sir == S) {
      sir = 0;
      ++rb;
    }
  }
};

// ---------------------------------------------------------------------------------------
// kernel: grid = G CTAs (1-D), block = 10 warps
//   warps 0..7  dequant: quad Q = w / 4, quarter j = w % 4; warp w dequantises chunk w of every ring stage (8 chunks)
//               into TMEM lanes 32j.. of A slot (2 * stage + Q) % 3, and stages its share of the activations
//   warp 8      bulk-TMA producer of the weight ring
//   warp 9      MMA issuer (one elected lane): per stage and quad 8 x tcgen05.mma + commit
// Ordering without extra barriers (streamed activations): a warp stores its pieces of stage i+1's activations BEFORE
// it signals its A slot of stage i, and the issuer starts stage i+1 only after it has seen all eight A-slot signals
// of stage i; an A slot is refilled only after the commit of its previous use, which (one issuer: commits are
// cumulative) also proves that the MMAs of stage i-2 are done, i.e. that activation ring slot (i+1) % 3 is free.
// ---------------------------------------------------------------------------------------
template <tg_dtype DT, int IK, int MB, bool XRES, bool MX4, bool G128, bool PEERS>
__device__ __forceinline__ void gemv_w4_tc_body(const ParamsTC& p, const Peers& peers) {
  using C = Cfg<MB, XRES>;
  extern __shared__ __align__(1024) uint8_t smem[];
  // the dynamic shared window of a kernel without static shared memory starts at kSmemBase (checked right below):
  // every shared address in this kernel is that constant plus an offset, so most of them fold into immediates
  constexpr uint32_t sbase = kSmemBase;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  // Programmatic dependent launch: the next kernel of the stream may become resident as soon as all our CTAs have
  // started.  We wait for the previous kernel before touching anything it may have produced; weights, LUTs and
  // scales may be declared static by the caller (TG_OPT_STATIC_WEIGHTS), in which case only the activation staging
  // (and with it every MMA and store) is ordered behind it.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool static_w = (p.flags & 8) != 0;
  if (threadIdx.x == 0) TC_TRACE(0);
  if (smem_u32(smem) != kSmemBase) __trap();
  if (!static_w) griddep_wait();

  int u_begin, u_end;
  cta_range(p, (int)blockIdx.x, u_begin, u_end);
  const int Cn = p.chunks_per_row;
  const int S = p.stages_per_row;
  const int n_groups = p.k >> p.glog2;

  // activation operand geometry.  K step s of a ring stage (s = Q*8 + T: quad Q, k-tile T of the chunk) reads operand
  // rows n = 4*mi + j (activation row mi, quarter j -> chunk Q*4 + j of the stage), 16 K values each:
  //   K index 2c + f  <->  k = 16*tile + c + 8f      (the order the packed bytes dequantise in)
  // 16-byte unit (s, h, n) = K indices 8h..8h+7.  Placement (the strides are what the descriptor is told):
  //   m = 1 (MB = 4): s*256 + h*64 + j*16           in the odd half-lines of region A (rows 4.. alias the other half);
  //                   resident: stage t of the row at t*16*256, else ring slot * 16 KiB
  //   MB = 4, m >= 2: ((s*2 + h)*G8 + n/8)*256 + (n%8)*16,  G8 = 1 (m = 2) or 2
  //   MB >= 8       : ((s*2 + h)*G8 + n/8)*128 + (n%8)*16 in the dense ring, G8 = MB / 2
  const bool x_single = MB == 4 && p.m == 1;
  const uint32_t x_g8 = MB == 4 ? (p.m > 2 ? 2u : 1u) : (uint32_t)(MB / 2);
  const uint32_t x_pitch = MB == 4 ? 256u : 128u;
  const uint32_t x0 = sbase + (MB == 4 ? 128u : kXDenseOff);
  // staging: a lane pair (f = lane & 1) holds x[k0 + 4h ..+3] (f = 0) and x[k0 + 8 + 4h ..+3] (f = 1) and turns them
  // into the 16-byte unit (lo0,hi0,lo1,hi1,lo2,hi2,lo3,hi3): lane 0 writes the first 8 bytes (needs the partner's .x),
  // lane 1 the rest
  const int x_f = lane & 1;
  auto x_put = [&](uint32_t dst, uint2 v) {
    const uint32_t got = __shfl_xor_sync(0xffffffffu, x_f ? v.x : v.y, 1);
    const uint32_t a = x_f ? got : v.x, b = x_f ? v.y : got;
    sts64(dst + (uint32_t)x_f * 8u, prmt(a, b, 0x5410u), prmt(a, b, 0x7632u));
  };

  // ------------------------------------------------------------------ setup
  Cursor pc;  // producer cursor (warp 8)
  pc.init(u_begin, S);
  uint32_t pseq = 0;
  uint64_t pol = 0;
  // MB >= 8 with pre-permuted activations: a stage's operand block travels with the stage's weights (same barrier)
  // The activation blocks have their own ring barriers: a weight stage is refilled as soon as the dequant warps have
  // read it, an activation block only once the stage's MMAs are done.
  const bool x_tma = MB >= 8 && p.xperm != nullptr;
  Cursor xpc;  // activation producer cursor
  xpc.init(u_begin, S);
  uint32_t xseq = 0;
  auto produce_x = [&]() {  // lane 0
    const int s = (int)(xseq % C::NX);
    const uint32_t bar = sbase + bar_off(B_XFULL + s);
    mbar_expect_tx(bar, C::kXStageBytes);
    bulk_g2s_plain(x0 + (uint32_t)s * C::kXStageBytes, p.xperm + (size_t)xpc.sir * C::kXStageBytes, C::kXStageBytes, bar);
    ++xseq;
    xpc.next(S);
  };
  // one ring stage: lane 0 arms the barrier, lanes 0..3 issue one 4 KiB n-tile copy each
  auto produce = [&]() {
    const int s = (int)(pseq % kWStages);
    const int tiles_valid = min(32, p.w_rows - pc.rb * 32) >> 3;
    const int k0 = pc.sir * 1024;
    const uint32_t bytes = (uint32_t)min(1024, p.k - k0) * 4u;  // per n-tile: 8 rows * k / 2
    const uint32_t bar = sbase + bar_off(B_WFULL + s);
    if (lane == 0) mbar_expect_tx(bar, bytes * (uint32_t)tiles_valid);
    __syncwarp();
    if (lane < tiles_valid)
      bulk_g2s(sbase + kWRingOff + (uint32_t)s * kWStageStride + tile_off<IK>(lane),
               p.w + (int64_t)(pc.rb * 4 + lane) * p.tile_stride + (int64_t)k0 * 4, bytes, bar, pol);
    if (lane == 0 && pseq < 3) TC_TRACE(17 + pseq);
    ++pseq;
    pc.next(S);
  };
  const int rl = row_of_lane(lane);  // the weight row (within the block) this lane owns
  // The LUT rows of a row block (32 x 32 B) travel global -> shared memory with cp.async (no registers, no stall):
  // requested one row block ahead, read back when the pair table is built.  Threads 0..63 copy 16 bytes each.
  auto request_lut = [&](int rb) {
    if (threadIdx.x < 64) {
      const int r = (int)threadIdx.x >> 1;
      const int row = min(rb * 32 + r, p.w_rows - 1);
      const uint16_t* src = p.lut + (int64_t)row * p.lut_stride + (threadIdx.x & 1) * 8;
      const uint32_t dst = sbase + odd_hl(C::kLutHl0 + (r >> 2)) + (uint32_t)(r & 3) * 32u + (threadIdx.x & 1) * 16u;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (warp == kDqWarps) {
    // the weight stream starts first thing: the ring has kWStages free stages, nobody has to be asked, and only the
    // three "full" barriers have to exist; the other barriers are set up while the first bytes are on their way
    if (lane == 0) {
      for (int i = 0; i < 3; ++i) mbar_init(sbase + bar_off(B_WFULL + i), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      TC_TRACE(16);
    }
    __syncwarp();
    pol = l2_evict_first_policy();
    for (int i = 0; i < kWStages && pc.u < u_end; ++i) produce();
    if (lane == 0) {
      for (int i = 0; i < 3; ++i) {
        mbar_init(sbase + bar_off(B_WEMPTY + i), kDqWarps);
        mbar_init(sbase + bar_off(B_AFULL + i), 4);
        mbar_init(sbase + bar_off(B_AEMPTY + i), 1);
        mbar_init(sbase + bar_off(B_XEMPTY + i), 1);
        mbar_init(sbase + bar_off(B_XFULL + i), 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(sbase + bar_off(B_DFULL + i), 1);
        mbar_init(sbase + bar_off(B_DEMPTY + i), 4);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
  } else if (warp == kDqWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kHolderOff), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    request_lut(u_begin / S);  // in flight across the setup barrier
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<const volatile uint32_t*>(smem + kHolderOff);
  if (threadIdx.x == 0) TC_TRACE(1);

  if (warp < kDqWarps) {
    // =============================================================== dequant warps
    const int Q = warp >> 2, j = warp & 3;
    const uint32_t lane4 = (uint32_t)lane * 4u;
    const uint32_t wl_off = kWRingOff + tile_off<IK>(rl >> 3) + (uint32_t)warp * 512u + (uint32_t)(rl & 7) * Geo<IK>::kRowStride;
    const uint32_t my_tmem = tmem + ((uint32_t)(j * 32) << 16);
    uint32_t dseq = 0;  // row blocks (accumulator buffer dseq % NB)
    uint32_t gst = 0;   // ring stages done by this CTA
    bool first_seg = true;
    // weight ring slot / phase of the current stage; A slot / phase of this quad's current use (= 2 * stage + Q)
    uint32_t ws = 0, wpar = 0, as = (uint32_t)Q, apar = 0;
    int free_uses = Q == 0 ? 2 : 1;     // uses 0, 1, 2 find their A slot untouched
    uint32_t w_ready = 0, a_ready = 0;  // answers of the probes issued one stage ago

    // Small global loads issued while the weight ring is full come back only after everything queued in front of
    // them (~2 us for 96 KB per SM), so the per-stage operands - group (scale, zero) words and streamed activation
    // pieces - travel through REGISTER RINGS filled several stages ahead: a stage consumes the head, shifts, and
    // requests the stage `depth` ahead at the tail.  Depth = ring size / values per stage.

    // ---- group words ----
    // G128 (group >= 128: one word per stage): ring of 4, one load per stage; else nsz = 2 (group 64) or 4 (group 32)
    // words per stage in a ring of 8 (4 / 2 stages ahead).
    const int nsz = (G128 || p.glog2 >= 7) ? 1 : (p.glog2 == 6 ? 2 : 4);
    uint32_t szr[G128 ? 4 : 8];
    Cursor zc;
    zc.init(u_begin, S);
    auto sz_word = [&](int t, int n) -> uint32_t {  // word t of n for this warp's chunk of stage zc
      const int cc = zc.sir * 8 + warp;
      if (zc.u >= u_end || cc >= Cn) return 0u;
      const int row = min(zc.rb * 32 + rl, p.w_rows - 1);
      const int gi = min((cc * 128 + t * (128 / n)) >> p.glog2, n_groups - 1);
      if constexpr (MX4) return e8m0_to_dt<DT>((uint32_t)p.exps[(int64_t)row * n_groups + gi]) | 0x80000000u;  // zero = -0
      else return p.sz[(int64_t)gi * p.w_rows + row];
    };
    auto sz_fill = [&](auto n_) {  // prologue: the first stages
      constexpr int n = decltype(n_)::value;
      constexpr int R = G128 ? 4 : 8;
#pragma unroll
      for (int d = 0; d < R / n; ++d) {
#pragma unroll
        for (int t = 0; t < n; ++t) szr[d * n + t] = sz_word(t, n);
        zc.next(S);
      }
    };
    // consume this stage's words into (s, s) / (z, z) pairs for the four 32-k quarters, shift, request the tail stage
    auto sz_step = [&](auto n_, uint32_t (&s2)[4], uint32_t (&z2)[4]) {
      constexpr int n = decltype(n_)::value;
      constexpr int R = G128 ? 4 : 8;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t v = szr[t * n / 4];
        s2[t] = prmt(v, v, 0x1010u);  // mx4 words carry zero = -0: fma(v, s, -0) == v * s incl. sign of zero
        z2[t] = prmt(v, v, 0x3232u);
      }
#pragma unroll
      for (int i = 0; i + n < R; ++i) szr[i] = szr[i + n];
#pragma unroll
      for (int t = 0; t < n; ++t) szr[R - n + t] = sz_word(t, n);
      zc.next(S);
    };
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    using I4 = std::integral_constant<int, 4>;
    using IM = std::integral_constant<int, MB>;

    // ---- activation staging: thread (warp w, lane) owns the 8-byte pieces f of the units (s = 2w + slo, h, n) ----
    // lane = f | j<<1 | (h or mi&1)<<3 | slo<<4: a half-warp writes one contiguous 128-byte half-line (conflict free)
    // and reads whole 32-byte sectors of x.  Pieces per stage: 1 (m = 1), 2 (m = 2), 4 (m = 3, 4), MB (MB >= 8).
    const int x_j = (lane >> 1) & 3, x_b3 = (lane >> 3) & 1, x_s = 2 * warp + (lane >> 4);
    const int x_ch = (x_s >> 3) * 4 + x_j;                     // chunk of the stage
    const int x_koff = x_ch * 128 + (x_s & 7) * 16 + 8 * x_f;  // + 4h
    // streamed activations
    Cursor xl, xs;
    xl.init(u_begin, S);
    xs = xl;
    uint32_t xs_seq = 0;
    uint2 xr[XRES ? 1 : C::XR];
    const int x_nit = x_single ? 1 : 2 * (int)x_g8;
    auto x_piece = [&](int it) -> uint2 {  // piece `it` of stage xl
      const int h = x_single ? x_b3 : (it & 1);
      const int mi = x_single ? 0 : (it >> 1) * 2 + x_b3;
      const int kk = xl.sir * 1024 + x_koff + 4 * h;
      if (xl.u < u_end && kk < p.k && mi < p.m) return *reinterpret_cast<const uint2*>(p.x + (int64_t)mi * p.k + kk);
      return make_uint2(0u, 0u);
    };
    auto x_fill = [&](auto n_) {
      constexpr int n = decltype(n_)::value;
      if constexpr (!XRES) {
#pragma unroll
        for (int d = 0; d < C::XR / n; ++d) {
#pragma unroll
          for (int it = 0; it < n; ++it) xr[d * n + it] = x_piece(it);
          xl.next(S);
        }
      }
    };
    // store the head stage's pieces into activation ring slot xs_seq % NX, shift, request the tail stage
    auto x_step = [&](auto n_) {
      constexpr int n = decltype(n_)::value;
      if constexpr (!XRES) {
        if (xs.u < u_end) {
          const uint32_t xb = x0 + (xs_seq % C::NX) * C::kXStageBytes;
#pragma unroll
          for (int it = 0; it < n; ++it) {
            uint32_t dst;
            if (x_single) {
              dst = xb + (uint32_t)x_s * 256u + (uint32_t)x_b3 * 64u + (uint32_t)x_j * 16u;
            } else {
              const uint32_t h = it & 1, grp = it >> 1;
              dst = xb + ((uint32_t)(x_s * 2 + h) * x_g8 + grp) * x_pitch + (uint32_t)(x_b3 * 4 + x_j) * 16u;
            }
            x_put(dst, xr[it]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
          ++xs_seq;
          xs.next(S);
        }
#pragma unroll
        for (int i = 0; i + n < C::XR; ++i) xr[i] = xr[i + n];
#pragma unroll
        for (int it = 0; it < n; ++it) xr[C::XR - n + it] = x_piece(it);
        xl.next(S);
      }
    };
    auto x_fill_any = [&]() {
      if constexpr (MB == 4) {
        if (x_nit == 1) x_fill(I1{});
        else if (x_nit == 2) x_fill(I2{});
        else x_fill(I4{});
      } else {
        x_fill(IM{});
      }
    };
    auto x_step_any = [&]() {
      if constexpr (MB == 4) {
        if (x_nit == 1) x_step(I1{});
        else if (x_nit == 2) x_step(I2{});
        else x_step(I4{});
      } else {
        x_step(IM{});
      }
    };
    if (G128 || nsz == 1) sz_fill(I1{});
    else if (nsz == 2) sz_fill(I2{});
    else sz_fill(I4{});

    Cursor cur;
    cur.init(u_begin, S);
    const int rb_first = cur.rb;

    // ---- epilogue of one row-block segment: accumulators -> row sums -> y (or the split fix-up workspace) ----
    // Software pipelined: the dequant warps do NOT drain at a row-block boundary.  They build the next pair table and
    // dequantise the next block's first stage while the tensor core finishes the previous block (the accumulators are
    // double buffered; with one buffer, MB = 16, the issuer simply holds the next block's MMAs until `finish` has
    // released it - by then at most 2 of the 3 A slots are waiting), and only then read the sums, which are long
    // complete.  Only the last segment of a CTA waits for its MMAs.
    bool pend_on = false, pend_mid = false;
    int pend_rb = 0, pend_first = 0, pend_n = 0;
    auto finish = [&](int rb, int seg_first, int seg_n, bool last) {
      const int row0 = rb * 32;
      // accumulators -> red[j][mi][lane] (quad 0's warps cover the four TMEM sub-partitions)
      if (Q == 0) {
        const int buf = (int)(dseq % C::NB);
        mbar_wait(sbase + bar_off(B_DFULL + buf), (dseq / C::NB) & 1u);
        tc_fence_after();
        if (threadIdx.x == 0 && last) TC_TRACE(12);
        const uint32_t dcol = my_tmem + (uint32_t)(buf * C::N + j);
        uint32_t v[MB];
#pragma unroll
        for (int mi = 0; mi < MB; ++mi)
          if (mi < p.m) v[mi] = tmem_ld1(dcol + (uint32_t)(mi * 4));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int mi = 0; mi < MB; ++mi)
          if (mi < p.m) sts32(sbase + odd_hl(C::kRedHl0 + j * MB + mi) + lane4, v[mi]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sbase + bar_off(B_DEMPTY + buf));
      }
      ++dseq;
      bar_sync(1, kDqThreads);

      // row sums: thread idx -> (mi, row); the four quarters' partials added in order
      const bool complete = (seg_first == 0 && seg_n == S);
      const int rows_valid = min(32, p.w_rows - row0);
      auto emit = [&](int mi, int rr, float total) {
        if constexpr (PEERS) {
          if (peers.tag != 0u && (p.flags & 16)) {
            // in-kernel exchange of silu(gate) * up over interleaved (gate, up) rows: rows (rr, rr + 1) make output
            // (row0 + rr) / 2; outputs of rows rr and rr + 2 travel as one tagged 8-byte word.  peers.n_total / col0
            // count OUTPUTS here (half the weight rows).
            const uint32_t mine = f32_to_dt<DT>(total);
            const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
            const uint32_t act = (uint32_t)silu_mul_dt<DT>((uint16_t)mine, (uint16_t)other);  // valid on even rr
            const uint32_t act2 = __shfl_xor_sync(0xffffffffu, act, 2);
            if (!(rr & 3) && rr < rows_valid) {
              const unsigned long long word = ((unsigned long long)peers.tag << 32) | (unsigned long long)(act | (act2 << 16));
              const int64_t widx = (int64_t)mi * (peers.n_total >> 1) + ((peers.col0 + ((row0 + rr) >> 1)) >> 1);
#pragma unroll 1
              for (int r = 0; r < peers.n; ++r)
                asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long*>(peers.y[r]) + widx), "l"(word)
                             : "memory");
            }
            return;
          }
          if (peers.tag != 0u) {  // in-kernel exchange: rows (rr, rr + 1) travel as one tagged 8-byte word to every rank
            const uint32_t mine = f32_to_dt<DT>(total);
            const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
            if (!(rr & 1) && rr < rows_valid) {
              const unsigned long long word = ((unsigned long long)peers.tag << 32) | (unsigned long long)(mine | (other << 16));
              const int64_t widx = (int64_t)mi * (peers.n_total >> 1) + ((peers.col0 + row0 + rr) >> 1);
#pragma unroll 1
              for (int r = 0; r < peers.n; ++r)
                asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long*>(peers.y[r]) + widx), "l"(word)
                             : "memory");
            }
            return;
          }
        }
        if (p.flags & 16) {  // (gate, up) row pairs -> silu(gate) * up
          const uint32_t mine = f32_to_dt<DT>(total);
          const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
          if (!(rr & 1) && rr < rows_valid)
            store_out<PEERS>(p.y, peers, (int64_t)mi * p.y_stride + ((row0 + rr) >> 1), silu_mul_dt<DT>((uint16_t)mine, (uint16_t)other));
        } else if (rr < rows_valid) {
          store_out<PEERS>(p.y, peers, (int64_t)mi * p.y_stride + row0 + rr, f32_to_dt<DT>(total));
        }
      };
      auto block_sum = [&](int mi, int rr) {
        const uint32_t a = sbase + odd_hl(C::kRedHl0 + mi) + (uint32_t)row_of_lane(rr) * 4u;
        float total = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) total += __uint_as_float(lds32(a + (uint32_t)(q * MB) * 256u));
        return total;
      };
      if (complete) {
        for (int idx = (int)threadIdx.x; idx < 32 * p.m; idx += kDqThreads) emit(idx >> 5, idx & 31, block_sum(idx >> 5, idx & 31));
      } else {
        // Row block shared with other CTAs.  The CTA that holds the block's FIRST stages owns the result: that segment
        // is the last thing the CTA does, while the other CTAs meet the block at the very beginning of their ranges,
        // so their partials are long there.  A partial travels as ONE 8-byte word (value, launch tag): whoever sees
        // the tag sees the value, so no fence, no atomic and no counter is needed (a fence would wait ~2 us for this
        // thread's prefetched group words to come back behind the weight stream).  The owner adds the partials in CTA
        // order: deterministic.
        const int me = (int)blockIdx.x;
        const int i_first = cta_of_unit(p, rb * S);
        const int i_last = cta_of_unit(p, rb * S + S - 1);
        constexpr int PER = (32 * MB + kDqThreads - 1) / kDqThreads;  // outputs per thread (2 at 16 rows)
        float tot[PER];
        bool act[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) {
          const int idx = (int)threadIdx.x + e * kDqThreads;
          act[e] = idx < 32 * p.m;
          tot[e] = act[e] ? block_sum(idx >> 5, idx & 31) : 0.f;
        }
        if (me != i_first) {
          const int slot = rb == rb_first ? 0 : 1;
#pragma unroll
          for (int e = 0; e < PER; ++e) {
            if (!act[e]) continue;
            const int idx = (int)threadIdx.x + e * kDqThreads;
            const unsigned long long word = ((unsigned long long)p.ws_tag << 32) | (unsigned long long)__float_as_uint(tot[e]);
            asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p.ws_partial + ((size_t)(me * 2 + slot) * 16) * 32 + idx), "l"(word)
                         : "memory");
          }
        } else {
          for (int ii = i_first + 1; ii <= i_last; ++ii) {
            int b0, e0;
            cta_range(p, ii, b0, e0);
            const int sl = (b0 / S == rb) ? 0 : 1;
            unsigned long long* src = p.ws_partial + ((size_t)(ii * 2 + sl) * 16) * 32 + threadIdx.x;
            // all of this thread's loads first (each one queues behind the weight stream: ~2 us), then the stragglers
            unsigned long long word[PER];
#pragma unroll
            for (int e = 0; e < PER; ++e)
              if (act[e]) asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(word[e]) : "l"(src + e * kDqThreads) : "memory");
#pragma unroll
            for (int e = 0; e < PER; ++e) {
              if (!act[e]) continue;
              while ((uint32_t)(word[e] >> 32) != p.ws_tag)
                asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(word[e]) : "l"(src + e * kDqThreads) : "memory");
              // consumed: clear the word.  A CUDA-graph replay re-launches this kernel with the SAME tag; without the
              // clear its owner could take this launch's partial for its own.  (The next writer of this word is a
              // launch that starts its stores only after this one has completed.)
              asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(src + e * kDqThreads), "l"(0ull) : "memory");
              tot[e] += __uint_as_float((uint32_t)word[e]);
            }
          }
#pragma unroll
          for (int e = 0; e < PER; ++e) {
            const int idx = (int)threadIdx.x + e * kDqThreads;
            if (act[e]) emit(idx >> 5, idx & 31, tot[e]);
          }
        }
      }
      if (threadIdx.x == 0 && last) TC_TRACE(13);
    };

    while (cur.u < u_end) {
      const int rb = cur.rb;
      const int seg_first = cur.sir;
      const int seg_n = min(S - cur.sir, u_end - cur.u);  // stages of this row block done by this CTA

      // ---- pair table of this row block (its LUT rows were requested one row block ago) ----
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      bar_sync(1, kDqThreads);  // LUT rows landed; everybody is done with the previous table
      {
        const uint32_t lrow = sbase + odd_hl(C::kLutHl0 + (rl >> 2)) + (uint32_t)(rl & 3) * 32u;
        const uint4 l0 = lds128(lrow), l1 = lds128(lrow + 16u);
        const uint32_t tp_[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const uint32_t hw = lds32(lrow + (uint32_t)warp * 4u);  // entries 2w, 2w+1: this warp builds the table
                                                                // entries whose high nibble is 2w / 2w+1
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t hi = hh ? hw >> 16 : hw;
          const uint32_t dst = sbase + (uint32_t)((2 * warp + hh) * 16) * 256u + lane4;
#pragma unroll
          for (int lo = 0; lo < 16; ++lo)
            sts32(dst + (uint32_t)lo * 256u, prmt(tp_[lo >> 1], hi, (lo & 1) ? 0x5432u : 0x5410u));
        }
      }
      bar_sync(1, kDqThreads);  // table complete, LUT staging area free again
      if (cur.u + seg_n < u_end) request_lut(rb + 1);
      if constexpr (!XRES) {
        if (first_seg && !x_tma) {
          // the activations are the previous kernel's output: everything up to here overlapped its tail
          if (static_w) griddep_wait();
          x_fill_any();
          x_step_any();                  // stage 0's activations
          bar_sync(2, kDqThreads + 32);  // with the issuer warp: they are complete
        }
      }
      if (threadIdx.x == 0 && first_seg) TC_TRACE(2);

#pragma unroll 1
      for (int i = 0; i < seg_n; ++i, ++gst) {
        const int cc = (seg_first + i) * 8 + warp;  // this warp's chunk of the row
        const int kt_valid = (p.k - cc * 128) >> 4;  // its k-tiles that exist (>= 8: all, <= 0: none)
        uint32_t s2[4], z2[4];
        if (G128 || nsz == 1) sz_step(I1{}, s2, z2);
        else if (nsz == 2) sz_step(I2{}, s2, z2);
        else sz_step(I4{}, s2, z2);

        const uint32_t acol = my_tmem + (uint32_t)C::kABase + as * 64u;
        const uint32_t wb = sbase + wl_off + ws * kWStageStride;
        if (!w_ready) mbar_wait(sbase + bar_off(B_WFULL) + ws * 8u, wpar);
        if (threadIdx.x == 0 && gst < 4) TC_TRACE(21 + gst * 4);

        // one k-tile (K step) of this lane's row: 8 byte lookups -> 8 TMEM columns
        auto k_step = [&](const uint32_t (&W)[4][4], int T, uint32_t (&r)[8]) {
          const int tp = T >> 1, b = T & 1;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t w = W[Geo<IK>::unit_of(tp, q)][Geo<IK>::word_of(tp, q)];
            // column 2q: byte b = (k0 | k0+8), column 2q+1: byte b+2 = (k0+1 | k0+9) of k-tile T
            r[2 * q] = lds_table(prmt(w, lane4, 0x7604u | (uint32_t)(b << 4)));
            r[2 * q + 1] = lds_table(prmt(w, lane4, 0x7604u | (uint32_t)((b + 2) << 4)));
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) r[q] = fma2<DT>(r[q], s2[tp], z2[tp]);
        };
        const bool no_wait = free_uses > 0;
        auto slot_wait = [&]() {  // the slot's previous use has been consumed (and with it all MMAs of stage - 2)
          if (!no_wait) {
            if (!a_ready) mbar_wait(sbase + bar_off(B_AEMPTY) + as * 8u, apar ^ 1u);
            tc_fence_after();
          }
        };
        // next stage's ring slot; next use of this quad: slot (as + 2) % 3, phase of use + 2
        const uint32_t ws_n = ws == 2u ? 0u : ws + 1u, wpar_n = ws == 2u ? wpar ^ 1u : wpar;
        const uint32_t as_n = as == 0u ? 2u : as - 1u, apar_n = as == 0u ? apar : apar ^ 1u;
        const bool more = cur.u + i + 1 < u_end;
        if (kt_valid >= 8) {
          // the common case, straight-line: 4 x LDS.128, then per k-tile 8 x (PRMT, LDS.32, HFMA2) + one tcgen05.st
          uint32_t W[4][4];
          if constexpr (IK == 8) {
            // lanes whose k-slot id bit 1 is set take their units one step ahead (bank-conflict-free 16-byte loads)
            const int sh = (lane >> 1) & 1;
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = lds128(wb + (uint32_t)(((u + sh) & 3) * 16));
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // logical unit u was loaded at step (u - sh) & 3
              const uint4 a = r[u], b = r[(u + 3) & 3];
              W[u][0] = sh ? b.x : a.x, W[u][1] = sh ? b.y : a.y, W[u][2] = sh ? b.z : a.z, W[u][3] = sh ? b.w : a.w;
            }
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 v = lds128(wb + (uint32_t)Geo<IK>::unit_off(u));
              W[u][0] = v.x, W[u][1] = v.y, W[u][2] = v.z, W[u][3] = v.w;
            }
          }
          // The stage's bytes now live in registers: hand the ring slot back to the producer right away (the arrive is
          // a release: it orders the loads above, of all lanes after the __syncwarp, before the refill), so the next
          // bulk copy into this slot is in flight during the whole dequant of the stage.
          if constexpr (kEarlyRelease) {
            __syncwarp();
            if (lane == 0) mbar_arrive(sbase + bar_off(B_WEMPTY) + ws * 8u);
          }
          // probe the next stage's weights now: the answer arrives while this stage's lookups issue
          w_ready = more ? mbar_test(sbase + bar_off(B_WFULL) + ws_n * 8u, wpar_n) : 0u;
          uint32_t r0[8];
          k_step(W, 0, r0);  // overlaps the wait for the TMEM slot
          slot_wait();
          tmem_st8(acol, r0);
#pragma unroll
          for (int T = 1; T < 8; ++T) {
            uint32_t r[8];
            k_step(W, T, r);
            tmem_st8(acol + (uint32_t)(T * 8), r);
          }
        } else {
          // partial / missing chunk (k tail): compact code, exact zeros where nothing exists.  A stage always serves
          // both quads (uniform slot sequence), so a warp without a chunk just writes zeros.
          w_ready = 0u;
          slot_wait();
#pragma unroll 1
          for (int T = 0; T < 8; ++T) {
            uint32_t r[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (T < kt_valid) {
              const int tp = T >> 1, b = T & 1;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                int unit, word;
                if constexpr (IK == 4) unit = (tp >> 1) * 2 + (q >> 1), word = (q & 1) * 2 + (tp & 1);
                else if constexpr (IK == 2) unit = tp, word = q;
                else unit = q, word = tp;
                const uint32_t w = lds32(wb + (uint32_t)(Geo<IK>::unit_off(unit) + 4 * word));
                r[2 * q] = lds_table(prmt(w, lane4, 0x7604u | (uint32_t)(b << 4)));
                r[2 * q + 1] = lds_table(prmt(w, lane4, 0x7604u | (uint32_t)((b + 2) << 4)));
              }
              const uint32_t sv = tp == 0 ? s2[0] : tp == 1 ? s2[1] : tp == 2 ? s2[2] : s2[3];
              const uint32_t zv = tp == 0 ? z2[0] : tp == 1 ? z2[1] : tp == 2 ? z2[2] : z2[3];
#pragma unroll
              for (int q = 0; q < 8; ++q) r[q] = fma2<DT>(r[q], sv, zv);
            }
            tmem_st8(acol + (uint32_t)(T * 8), r);
          }
        }
        if (!x_tma) x_step_any();  // (streamed activations) stage + 1's pieces: their ring slot is free, see the kernel comment
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(sbase + bar_off(B_AFULL) + as * 8u);
          if (!kEarlyRelease || kt_valid < 8) mbar_arrive(sbase + bar_off(B_WEMPTY) + ws * 8u);  // (k tail: word loads all along the stage)
        }
        if (free_uses > 0) --free_uses;
        // probe this quad's next A slot (previous use: the OTHER quad's stage before ours); consumed one stage later
        a_ready = (more && free_uses == 0) ? mbar_test(sbase + bar_off(B_AEMPTY) + as_n * 8u, apar_n ^ 1u) : 0u;
        if (threadIdx.x == 0 && gst < 4) TC_TRACE(22 + gst * 4);
        ws = ws_n, wpar = wpar_n;
        as = as_n, apar = apar_n;
        pend_mid = false;
        if (kDefer && pend_on) {  // (i == 0) the previous row block's sums: complete by now, nobody waits
          finish(pend_rb, pend_first, pend_n, false);
          pend_on = false, pend_mid = true;
        }
      }
      cur.skip(seg_n, S);
      if (threadIdx.x == 0 && first_seg) TC_TRACE(11);
      // this row block's sums are read one stage into the NEXT row block (or after the loop): see `finish`
      pend_on = true, pend_rb = rb, pend_first = seg_first, pend_n = seg_n;
      first_seg = false;
      if (!kDefer) {
        finish(pend_rb, pend_first, pend_n, cur.u >= u_end);
        pend_on = false;
      }
    }
    if (pend_on) {
      if (pend_mid) bar_sync(1, kDqThreads);  // the deferred epilogue of this very stage may still be reading `red`
      finish(pend_rb, pend_first, pend_n, true);
    }
    if (threadIdx.x == 0) TC_TRACE(14);
    if constexpr (PEERS) {
      if (peers.tag != 0u) {
        // collect this CTA's slice of the full output from the local exchange buffer (all ranks store into it, ours
        // included): spin until a word carries this call's tag, then write its two values to the plain output
        const int half = peers.n_total >> 1;
        const int words = p.m * half;
        const int lo = (int)(((int64_t)words * (int)blockIdx.x) / (int)gridDim.x);
        const int hi = (int)(((int64_t)words * ((int)blockIdx.x + 1)) / (int)gridDim.x);
        unsigned long long* ll = reinterpret_cast<unsigned long long*>(peers.y[peers.self]);
        for (int wi = lo + (int)threadIdx.x; wi < hi; wi += kDqThreads) {
          unsigned long long word;
          do {
            asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(word) : "l"(ll + wi) : "memory");
          } while ((uint32_t)(word >> 32) != peers.tag);
          // consumed: clear the word, so that the buffer's next use (>= 2 calls later, possibly a CUDA-graph replay
          // of this very launch with the same tag) starts from words without a tag
          asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(ll + wi), "l"(0ull) : "memory");
          const int mi = wi / half, c = wi - mi * half;
          *reinterpret_cast<uint32_t*>(p.y + (int64_t)mi * p.y_stride + 2 * c) = (uint32_t)word;
        }
      }
    }
  } else if (warp == kDqWarps) {
    // =============================================================== TMA producer: the rest of the stream
    if (x_tma) {
      // the permuted activations are the previous kernel's output (x_permute_kernel); the weights were not - the copies
      // of the first stages are long under way
      if (static_w) griddep_wait();
      // weights and activations advance independently: whichever ring has a free slot is refilled
      while (pc.u < u_end || xpc.u < u_end) {
        auto probe = [&](uint32_t bar, uint32_t parity) {  // lane 0 probes for the warp
          return __shfl_sync(0xffffffffu, lane == 0 ? mbar_try(bar, parity) : 0u, 0) != 0u;
        };
        if (pc.u < u_end) {
          const int s = (int)(pseq % kWStages);
          if (pseq < (uint32_t)kWStages || probe(sbase + bar_off(B_WEMPTY + s), ((pseq / kWStages) - 1u) & 1u)) produce();
        }
        if (xpc.u < u_end) {
          const int s = (int)(xseq % C::NX);
          if (xseq < (uint32_t)C::NX || probe(sbase + bar_off(B_XEMPTY + s), ((xseq / C::NX) - 1u) & 1u)) {
            if (lane == 0) produce_x();
            else ++xseq, xpc.next(S);
            __syncwarp();
          }
        }
      }
    } else {
      while (pc.u < u_end) {
        const int s = (int)(pseq % kWStages);
        mbar_wait(sbase + bar_off(B_WEMPTY + s), ((pseq / kWStages) - 1u) & 1u);
        produce();
      }
    }
  } else {
    // =============================================================== MMA issuer (one warp, one elected lane)
    const uint32_t leader = elect_one();
    constexpr uint32_t fmt = DT == TG_BF16 ? 1u : 0u;
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(C::N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t step = x_single ? 256u : 2u * x_g8 * x_pitch;
    const uint32_t lbo = x_single ? 64u : x_g8 * x_pitch;
    const uint32_t sbo = x_pitch;
    const uint32_t desc_hi = (sbo >> 4) | (1u << 14);  // descriptor version 1 (sm_100), no swizzle
    uint32_t gst = 0, dseq = 0;
    uint32_t x_staged = 0;  // resident activations: bit t = stage t of the row is in shared memory
    // Resident activations (one row, all of k; stage t of the row -> half-lines [16t, 16t+16)) are staged by THIS warp,
    // so the dequant warps never wait for the previous kernel: they fill the TMEM slots while it is still running.
    // Up to 4 stages (32 eight-byte pieces per lane) are requested together; each is stored right before its first
    // MMA, so the first MMA waits for one load latency only.
    uint2 xv[4][8];      // pieces of up to four row stages [xv_t0, xv_t0 + xv_n), requested together
    int xv_t0 = 0, xv_n = 0;
    const int xl_j = (lane >> 1) & 3, xl_b3 = (lane >> 3) & 1, xl_hi = lane >> 4;
    auto x_request = [&](int t0, int n) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int xs_ = 2 * q + xl_hi;
          const int kk = (t0 + i) * 1024 + ((xs_ >> 3) * 4 + xl_j) * 128 + (xs_ & 7) * 16 + 8 * x_f + 4 * xl_b3;
          xv[i][q] = (i < n && kk < p.k) ? *reinterpret_cast<const uint2*>(p.x + kk) : make_uint2(0u, 0u);
        }
      xv_t0 = t0, xv_n = n;
    };
    auto x_store_stage = [&](int t, const uint2 (&v)[8]) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        x_put(x0 + (uint32_t)(t * 16 + 2 * q + xl_hi) * 256u + (uint32_t)xl_b3 * 64u + (uint32_t)xl_j * 16u, v[q]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
      __syncwarp();
    };
    if constexpr (XRES) {
      if (lane == 0) TC_TRACE(50);
      if (static_w) griddep_wait();  // the activations are the previous kernel's output
      if (lane == 0) TC_TRACE(51);
    } else if (!x_tma) {
      bar_sync(2, kDqThreads + 32);  // the first activations are staged (by the dequant warps)
    }
    Cursor cur;
    cur.init(u_begin, S);
    while (cur.u < u_end) {
      const int seg_n = min(S - cur.sir, u_end - cur.u);
      const int buf = (int)(dseq % C::NB);
      if (dseq >= (uint32_t)C::NB) {
        mbar_wait(sbase + bar_off(B_DEMPTY + buf), ((dseq / C::NB) - 1u) & 1u);
        tc_fence_after();
      }
      const uint32_t dcol = tmem + (uint32_t)(buf * C::N);
      uint32_t acc = 0;
      for (int i = 0; i < seg_n; ++i, ++gst) {
        const uint32_t xb = XRES ? x0 + (uint32_t)(cur.sir + i) * (16u * 256u) : x0 + (gst % C::NX) * C::kXStageBytes;
        if constexpr (XRES) {
          const int t = cur.sir + i;
          if (!((x_staged >> t) & 1u)) {
            if (t < xv_t0 || t >= xv_t0 + xv_n) {  // not requested yet: this and the next unstaged stages of the row
              int n = 1;
              while (n < 4 && t + n < S && !((x_staged >> (t + n)) & 1u)) ++n;
              x_request(t, n);
            }
            switch (t - xv_t0) {  // store only the stage needed now; the others' loads keep flying
              case 0: x_store_stage(t, xv[0]); break;
              case 1: x_store_stage(t, xv[1]); break;
              case 2: x_store_stage(t, xv[2]); break;
              default: x_store_stage(t, xv[3]); break;
            }
            if (lane == 0 && x_staged == 0) TC_TRACE(52);
            x_staged |= 1u << t;
          }
        }
        if (x_tma) mbar_wait(sbase + bar_off(B_XFULL) + (gst % C::NX) * 8u, (gst / C::NX) & 1u);  // the stage's activation block
#pragma unroll
        for (int Q = 0; Q < 2; ++Q) {
          const uint32_t use = 2u * gst + (uint32_t)Q;
          const int slot = (int)(use % C::NS);
          mbar_wait(sbase + bar_off(B_AFULL + slot), (use / C::NS) & 1u);
          tc_fence_after();
          if (gst < 3 && Q == 0 && lane == 0) TC_TRACE(37 + gst * 3);
          __syncwarp();
          const uint32_t acol = tmem + (uint32_t)(C::kABase + slot * 64);
          if (leader) {
            // every tcgen05 issue of the warp sits in this one block: ptxas keeps it a real branch (a lone
            // leader-predicated commit gets if-converted into a predicated R2UR + unpredicated UTCBAR, which
            // faults when the warp is not converged)
#pragma unroll
            for (int T = 0; T < 8; ++T) {
              const uint32_t addr = xb + (uint32_t)(Q * 8 + T) * step;
              const uint64_t desc = ((uint64_t)desc_hi << 32) | (uint64_t)(((addr >> 4) & 0x3fffu) | ((lbo >> 4) << 16));
              tc_mma(dcol, acol + (uint32_t)(T * 8), desc, idesc, acc);
              acc = 1u;
            }
            tc_commit(sbase + bar_off(B_AEMPTY + slot));
            if (Q == 1 && x_tma) tc_commit(sbase + bar_off(B_XEMPTY) + (gst % C::NX) * 8u);  // the stage's activation block is free
            if (Q == 1 && i == seg_n - 1) tc_commit(sbase + bar_off(B_DFULL + buf));  // the row block's sums are complete
          }
          acc = 1u;
          __syncwarp();
        }
        if (gst < 3 && lane == 0) TC_TRACE(38 + gst * 3);
      }
      ++dseq;
      cur.skip(seg_n, S);
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
#ifdef TG_W4_TRACE
  if (threadIdx.x == 0) {
    TC_TRACE(48);
    if (p.trace != nullptr) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[(size_t)blockIdx.x * 64 + 63] = smid;
    }
  }
#endif
  if (warp == kDqWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

template <tg_dtype DT, int IK, int MB, bool XRES, bool MX4, bool G128>
__global__ void __launch_bounds__(Cfg<MB, XRES>::kThreads, Cfg<MB, XRES>::kMinBlocks) gemv_w4_tc_kernel(const ParamsTC p) {
  gemv_w4_tc_body<DT, IK, MB, XRES, MX4, G128, false>(p, Peers{});
}
// row-sharded variant (decode kernels only): the epilogue stores into every rank's symmetric output buffer
template <tg_dtype DT, int IK, int MB, bool XRES, bool MX4, bool G128>
__global__ void __launch_bounds__(Cfg<MB, XRES>::kThreads, Cfg<MB, XRES>::kMinBlocks)
    gemv_w4_tc_peer_kernel(const ParamsTC p, const __grid_constant__ Peers peers) {
  gemv_w4_tc_body<DT, IK, MB, XRES, MX4, G128, true>(p, peers);
}

// ---------------------------------------------------------------------------------------
// MB >= 8: the activations, permuted ONCE into the operand layout the MMA descriptor describes (instead of by every CTA
// for every stage through registers: 32 KB per stage at 16 rows, twice the stage's weight bytes).  Block t of the
// output = the operand block of stage t of a row (kXStageBytes): 16-byte unit (s, h, n) at
// ((s*2 + h) * G8 + n/8) * 128 + (n%8) * 16 holds K indices 8h..8h+7 of K step s = Q*8 + T for operand row n = 4*mi + j,
// K index 2c + f <-> k = 1024 t + 128 (4Q + j) + 16 T + c + 8 f; zeros beyond k and beyond the m valid rows.
// One CTA per stage; PDL citizen like the GEMV that follows it.
// ---------------------------------------------------------------------------------------
template <int MB>
__global__ void __launch_bounds__(256) x_permute_kernel(const uint16_t* __restrict__ x, uint4* __restrict__ out, int m, int k) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  constexpr int G8 = MB / 2;
  constexpr int kUnits = 16 * 2 * 4 * MB;  // per stage
  const int t = (int)blockIdx.x;
  for (int u = (int)threadIdx.x; u < kUnits; u += 256) {
    const int n = (u & 7) + 8 * ((u >> 3) % G8);
    const int sh = u / (8 * G8), h = sh & 1, sstep = sh >> 1;
    const int mi = n >> 2, j = n & 3, Q = sstep >> 3, T = sstep & 7;
    const int kb = t * 1024 + (Q * 4 + j) * 128 + T * 16 + 4 * h;  // c = 4h .. 4h+3; f = 0: +0, f = 1: +8
    uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
    if (mi < m) {
      const uint16_t* row = x + (int64_t)mi * k;
      if (kb + 4 <= k) lo = *reinterpret_cast<const uint2*>(row + kb);
      if (kb + 12 <= k) hi = *reinterpret_cast<const uint2*>(row + kb + 8);
    }
    // K indices 8h + i, i = 0..7: (c, f) = (4h + i/2, i & 1) -> lo0 hi0 lo1 hi1 lo2 hi2 lo3 hi3
    uint4 v;
    v.x = prmt(lo.x, hi.x, 0x5410u), v.y = prmt(lo.x, hi.x, 0x7632u);
    v.z = prmt(lo.y, hi.y, 0x5410u), v.w = prmt(lo.y, hi.y, 0x7632u);
    out[(size_t)t * kUnits + u] = v;
  }
}
constexpr size_t kXPermPoolBytes = 1u << 20;  // 16 rows x 32768 k x 2 bytes
__device__ __align__(128) uint8_t g_xperm[kWsPools][kXPermPoolBytes];

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
std::atomic<unsigned> g_launch_seq{0};  // launch tag of the split fix-up (one counter for ALL instantiations)
int g_ctas_per_sm = 0;  // 0 = heuristic (tuning: env TG_TC_CTAS)
int g_split = -1;       // -1 = heuristic, 0 = whole row blocks per CTA, 1 = stream-K over stages (tuning: env TG_TC_SPLIT)
bool g_x_tma = true;    // MB >= 8: activations pre-permuted + TMA-staged (tuning: env TG_TC_XTMA=0 -> register-staged)

struct DeviceInfo {
  int n_sm = 0;
  unsigned long long* ws_partial = nullptr;
  uint8_t* xperm = nullptr;
};
std::atomic<unsigned> g_xperm_seq{0};
static int device_info(DeviceInfo** out) {
  static thread_local DeviceInfo info[kMaxDevices];
  DeviceInfo& d = info[current_device_slot()];
  if (d.n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (cudaGetSymbolAddress((void**)&d.ws_partial, g_ws_partial) != cudaSuccess ||
        cudaGetSymbolAddress((void**)&d.xperm, g_xperm) != cudaSuccess) {
      set_error("cudaGetSymbolAddress failed: %s", cudaGetErrorString(cudaGetLastError()));
      return TG_ERR_CUDA;
    }
    d.n_sm = n;
    static bool env_read = false;
    if (!env_read) {
      if (getenv("TG_TC_CTAS")) g_ctas_per_sm = atoi(getenv("TG_TC_CTAS"));
      if (getenv("TG_TC_SPLIT")) g_split = atoi(getenv("TG_TC_SPLIT"));
      if (getenv("TG_TC_XTMA")) g_x_tma = atoi(getenv("TG_TC_XTMA")) != 0;
      env_read = true;
    }
  }
  *out = &d;
  return TG_OK;
}

template <tg_dtype DT, int IK, int MB, bool XRES, bool MX4, bool G128>
int launch_one(ParamsTC p, const Peers& peers, int row_blocks, cudaStream_t st) {
  using C = Cfg<MB, XRES>;
  auto kern = gemv_w4_tc_kernel<DT, IK, MB, XRES, MX4, G128>;
  const void* kern_peer = nullptr;
  if constexpr (MB == 4) kern_peer = (const void*)gemv_w4_tc_peer_kernel<DT, IK, MB, XRES, MX4, G128>;
  static thread_local bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    for (const void* f : {(const void*)kern, kern_peer}) {
      if (f == nullptr) continue;
      if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem) != cudaSuccess ||
          cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) !=
              cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem=%u) failed: %s", C::kSmem, cudaGetErrorString(cudaGetLastError()));
        return TG_ERR_CUDA;
      }
    }
    attr_set = true;
  }
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc != TG_OK) return rc;

  // Work split.  One CTA per SM while a CTA's share is small AND there are no more row blocks than SMs (the second
  // slot of every SM is then free for the next kernel of the stream, whose prologue overlaps our dequant), two per SM
  // otherwise (they hide each other's row-block boundaries).  Stream-K over ring stages when that shortens the longest CTA by more than the fix-up
  // of a shared row block costs (~kFixup stages), else whole row blocks per CTA.
  const int S = p.stages_per_row;
  const int64_t stages = (int64_t)row_blocks * S;
  int per_sm = g_ctas_per_sm;
  // measured, one activation row, homogeneous chains (us, 1 / 2 CTAs per SM): 6144 x 4096 8.4 / 6.7; 4096 x 6144 8.1 / 7.1;
  // 4096 x 8192 10.3 / 9.2; 4096 x 11008 13.9 / 11.0 - but 4096 x 4096 5.6 / 6.0; 4096 x 5120 6.9 / 7.7; 2048 x 14336 9.2 / 9.7
  if (per_sm <= 0)
    per_sm = (stages >= (int64_t)di->n_sm * 2 * 5 || row_blocks > di->n_sm || (row_blocks >= 128 && stages >= (int64_t)di->n_sm * 5)) ? 2 : 1;
  if (per_sm > C::kMinBlocks) per_sm = C::kMinBlocks;
  int64_t slots = (int64_t)di->n_sm * per_sm;
  if (slots > kMaxGrid) slots = kMaxGrid;
  constexpr int64_t kFixup = MB == 16 ? 3 : MB == 8 ? 2 : 1;  // stages a shared row block costs (32 x m partials through global memory)
  const int64_t g_split_ctas = stages < slots ? stages : slots;
  const int64_t g_whole_ctas = row_blocks < slots ? row_blocks : slots;
  const int64_t t_split = div_up(stages, g_split_ctas) + kFixup;
  const int64_t t_whole = div_up(row_blocks, g_whole_ctas) * S;
  const bool split = g_split >= 0 ? g_split != 0 : t_split < t_whole;
  int64_t G;
  if (split) {
    G = g_split_ctas;
    p.ug = 1;
    p.cq = (int)(stages / G);
    p.cr = (int)(stages % G);
  } else {
    G = g_whole_ctas;
    p.ug = S;
    p.cq = (int)(row_blocks / G);
    p.cr = (int)(row_blocks % G);
  }
  const unsigned seq = g_launch_seq.fetch_add(1u, std::memory_order_relaxed) + 1u;  // process-wide: tags must be unique
  p.ws_partial = di->ws_partial + (size_t)(seq % kWsPools) * (kMaxGrid * 2 * 16 * 32);
  p.ws_tag = seq;

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)G, 1, 1);
  cfg.blockDim = dim3(C::kThreads, 1, 1);
  cfg.dynamicSmemBytes = C::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  int na = 0;
  if (g_pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t e;
  if (peers.n > 0) {
    if constexpr (MB == 4) {
      e = cudaLaunchKernelEx(&cfg, gemv_w4_tc_peer_kernel<DT, IK, MB, XRES, MX4, G128>, p, peers);
    } else {
      set_error("row-sharded epilogue: at most 4 activation rows per pass");
      return TG_ERR_UNSUPPORTED;
    }
  } else {
    e = cudaLaunchKernelEx(&cfg, kern, p);
  }
  if (e != cudaSuccess) {
    set_error("gemv_w4_tc launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return TG_ERR_CUDA;
  }
  count_launch();
  return TG_OK;
}

template <tg_dtype DT, int IK, bool MX4>
int launch_m(ParamsTC p, const Peers& peers0, int row_blocks, int64_t rows_x, const uint16_t* x, uint16_t* y, cudaStream_t st) {
  Peers peers = peers0;
  const int per_pass = peers0.n > 0 ? 4 : 16;  // one pass carries up to 16 activation rows (4 with the peer-store epilogue)
  for (int64_t r0 = 0; r0 < rows_x; r0 += per_pass) {
    p.m = (int)((rows_x - r0) < per_pass ? (rows_x - r0) : per_pass);
    p.x = x + r0 * p.k;
    p.y = y + r0 * p.y_stride;
    for (int r = 0; r < peers0.n; ++r)  // plain peer stores: element offset; exchange: 8-byte words [m][n_total / 2]
      peers.y[r] = peers0.y[r] + (peers0.tag != 0u ? r0 * (peers0.n_total >> 1) * 4 : r0 * p.y_stride);
    int rc;
    const bool res = p.m == 1 && p.stages_per_row <= kXResMaxStages;
    // the decode kernel (one activation row) is specialised for groups >= 128 (one group word per stage)
    if (res && p.glog2 >= 7) rc = launch_one<DT, IK, 4, true, MX4, true>(p, peers, row_blocks, st);
    else if (res) rc = launch_one<DT, IK, 4, true, MX4, false>(p, peers, row_blocks, st);
    // 2..4 rows: the two-CTAs-per-SM kernel pays off once an SM has enough stages to keep both busy; small problems run
    // faster as one 8-row CTA per SM (measured: profiles/r2/kernel_choice.md)
    else if (p.m <= 4 && ((int64_t)row_blocks * p.stages_per_row >= 148 * 10 || peers0.n > 0))
      rc = launch_one<DT, IK, 4, false, MX4, false>(p, peers, row_blocks, st);
    else {
      // 5..16 rows: permute the activations once (a tiny PDL-chained kernel) and let the GEMV's TMA producer bring each
      // stage's operand block in; k too long for the device-global staging pool: the register-staged path
      const int mb = p.m <= 8 ? 8 : 16;
      const size_t need = (size_t)p.stages_per_row * (size_t)(32 * (mb / 2) * 128);
      p.xperm = nullptr;
      DeviceInfo* di = nullptr;
      rc = device_info(&di);  // (also reads the tuning environment on the first call)
      if (rc != TG_OK) return rc;
      if (need <= kXPermPoolBytes && g_x_tma) {
        uint8_t* pool = di->xperm + (size_t)(g_xperm_seq.fetch_add(1u, std::memory_order_relaxed) % kWsPools) * kXPermPoolBytes;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)p.stages_per_row, 1, 1);
        cfg.blockDim = dim3(256, 1, 1);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = g_pdl ? 1 : 0;
        cudaError_t e = mb == 8 ? cudaLaunchKernelEx(&cfg, x_permute_kernel<8>, p.x, (uint4*)pool, p.m, p.k)
                                : cudaLaunchKernelEx(&cfg, x_permute_kernel<16>, p.x, (uint4*)pool, p.m, p.k);
        if (e != cudaSuccess) {
          set_error("x_permute_kernel launch failed: %s", cudaGetErrorString(e));
          (void)cudaGetLastError();
          return TG_ERR_CUDA;
        }
        count_launch();
        p.xperm = pool;
      }
      if (mb == 8) rc = launch_one<DT, IK, 8, false, MX4, false>(p, peers, row_blocks, st);
      else rc = launch_one<DT, IK, 16, false, MX4, false>(p, peers, row_blocks, st);
    }
    if (rc != TG_OK) return rc;
  }
  return TG_OK;
}

template <tg_dtype DT, bool MX4>
int launch_ik(const ParamsTC& p, const Peers& peers, int ik, int row_blocks, int64_t rows_x, const uint16_t* x, uint16_t* y,
              cudaStream_t st) {
  switch (ik) {
    case 2: return launch_m<DT, 2, MX4>(p, peers, row_blocks, rows_x, x, y, st);
    case 4: return launch_m<DT, 4, MX4>(p, peers, row_blocks, rows_x, x, y, st);
    case 8: return launch_m<DT, 8, MX4>(p, peers, row_blocks, rows_x, x, y, st);
  }
  set_error("B-layout int4 innerKTiles must be 2, 4 or 8 (got %d)", ik);
  return TG_ERR_INVALID_ARGUMENT;
}

}  // namespace tc

int launch_gemm_w4_tc_B(void* y, const void* x, const int32_t* w, const void* sz, const void* lut, const uint8_t* exps,
                        int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik, tg_w4_format fmt, tg_dtype dt,
                        const uint16_t* const_lut, cudaStream_t st, void* const* y_peers, int n_peers, int64_t y_row_stride,
                        int silu_pairs, int self_rank, uint32_t exchange_tag) {
  tc::ParamsTC p{};
  w4::Peers peers{};
  peers.n = n_peers;
  for (int r = 0; r < n_peers; ++r) peers.y[r] = static_cast<uint16_t*>(y_peers[r]);
  peers.tag = exchange_tag;  // != 0: y_peers are the ranks' exchange buffers, y is the plain local output
  peers.self = self_rank;
  peers.n_total = (int)(w_rows * n_peers) >> (silu_pairs && exchange_tag ? 1 : 0);  // (outputs, not rows, with the activation)
  peers.col0 = (int)(w_rows * self_rank) >> (silu_pairs && exchange_tag ? 1 : 0);
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
  p.chunks_per_row = (int)div_up(k, 128);
  p.stages_per_row = (int)div_up(k, 1024);
  const int64_t row_blocks = div_up(w_rows, 32);
  if (row_blocks * p.stages_per_row >= (1ll << 31)) {
    set_error("w_rows * k = %lld * %lld exceeds what one launch can index", (long long)w_rows, (long long)k);
    return TG_ERR_UNSUPPORTED;
  }
  p.flags = (w4::g_static_weights ? 8 : 0) | (silu_pairs ? 16 : 0);
#ifdef TG_W4_TRACE
  p.trace = w4::g_trace_buf;
#endif
  const uint16_t* xx = (const uint16_t*)x;
  uint16_t* yy = (uint16_t*)y;
  if (fmt == TG_W4_MX4) return tc::launch_ik<TG_BF16, true>(p, peers, ik, (int)row_blocks, rows_x, xx, yy, st);  // bf16 only
  if (dt == TG_BF16) return tc::launch_ik<TG_BF16, false>(p, peers, ik, (int)row_blocks, rows_x, xx, yy, st);
  return tc::launch_ik<TG_FP16, false>(p, peers, ik, (int)row_blocks, rows_x, xx, yy, st);
}

void set_tc_ctas_per_sm(int v) { tc::g_ctas_per_sm = v; }

}  // namespace tg
