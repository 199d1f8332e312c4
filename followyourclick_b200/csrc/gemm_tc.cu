// tcgen05 (5th-gen tensor core) GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[p, n] = epilogue( alpha * sum_{tap, c} X[pixel(p) + (dy,dx)(tap), c] * W[n, tap, c] )
//
// One persistent CTA per SM, 10 warps, warp-specialised:
//   warp 0 (1 lane)  TMA producer : cp.async.bulk.tensor (4-D box for the activation patch, 3-D box for the weight
//                                   slab) into a 4-stage 128B-swizzled shared-memory ring, mbarrier expect_tx
//   warp 1 (1 lane)  MMA issuer   : tcgen05.mma.cta_group::1.kind::f16  128 x BN x 16, bf16 x bf16 -> fp32 in TMEM,
//                                   tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..9       epilogue     : tcgen05.ld (32x32b) TMEM -> registers, fused bias / time-embedding row bias /
//                                   residual / GEGLU / alpha, vectorised global stores; double-buffered TMEM
//                                   accumulators (2 x 256 columns) overlap the epilogue with the next tile's MMAs.
// The 3x3 convolution never builds an im2col matrix: for each filter tap the producer loads the SAME 4-D box shifted
// by (dy, dx); TMA's out-of-bounds zero fill implements the padding halo.  A plain GEMM is the 1-tap special case.
// Stride-2 convolutions run on a parity-plane split of the input (fyc_space_to_planes), which turns every tap into
// a unit-stride shifted read of one plane (tap_img selects the plane).
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;            // UMMA M (rows = output pixels per tile)
constexpr int BK = 64;             // K per stage = one 128-byte swizzle row of bf16
constexpr int STAGES = 4;
constexpr int MAX_BN = 256;
constexpr int A_BYTES = BM * BK * 2;           // 16 KB
constexpr int B_BYTES = MAX_BN * BK * 2;       // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int MAX_STAGES = 8;                   // barrier slots; the W-resident mode runs up to 8 A-only stages
constexpr int RING_BYTES = STAGES * STAGE_BYTES;                // 192 KB operand ring (or W slab + A ring)
constexpr int STAGING_BYTES = 8 * 4096;         // per epilogue warp: 32 rows x 128 B (64 bf16 columns), 128B-swizzled
constexpr int OFF_STAGING = RING_BYTES + 256;
constexpr int OFF_BIAS = OFF_STAGING + STAGING_BYTES;           // GEGLU: 2 x 256 fp32 bias values of the current / next tile
constexpr int SMEM_BYTES = OFF_BIAS + 2048;                     // dynamic smem is declared 1024-aligned (checked in the kernel)
constexpr int NUM_THREADS = 320;          // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter)

struct TcParams {
  // problem
  int64_t M;                 // valid output rows (pixels)
  int N, N_out;              // accumulator columns / stored columns (N/2 for GEGLU)
  int taps, cin_blocks;      // K loop = taps x cin_blocks
  int BN, n_tiles;
  // output-pixel tiling: tile = bn images x bh rows x bw cols; Wt = ceil(Wo / bw) tiles per row, etc.
  int bw, bh, bn, Wo, Ho, w_tiles, h_tiles;
  int lg_bw, lg_bh;          // log2 of bw, bh (both powers of two)
  int64_t m_tiles;
  int tap_dy[9], tap_dx[9], tap_img[9];
  // epilogue
  const float* bias; const void* residual; const float* rowbias; void* out;
  int64_t ldo, ldr, rows_per_group;
  int64_t ldrb;              // row stride of rowbias (elements; N unless the caller passes a slice of a wider table)
  float alpha;
  int flags;
  // 16-wide k-steps of the LAST k block of a tap that hold data (1..4): K % 64 != 0 leaves TMA-zero-filled columns there - an LN-folded GEMM's
  // 8-column mean block would otherwise cost a full 64-deep block of MMA work
  int last_ksteps;
  // two-segment K (fyc_gemm_args.A2): k blocks [0, cb_split) of a tap come from map_a, the rest from map_a2 (INT_MAX: single source)
  int cb_split;
  // LayerNorm folded into the GEMM (FYC_EPI_LNFOLD): per-row rstd; the mean subtraction is in the (row-centred) weights
  const float* ln_rs;
  // W-resident mode (small K): the CTA keeps its whole BN x K weight slab in shared memory and only streams A
  int resident, a_stages;
  // CTA-pair mode (tcgen05 cta_group::2): a cluster of two CTAs computes a 256 x BN tile; each CTA stages its own 128 rows of A
  // and HALF of the weight tile, the leader issues M = 256 MMAs that read both halves, each CTA drains its own 128 TMEM lanes
  int pair;
  // gridDim.x as mixed-radix digits (n, w, h, image tiles): tile coordinates advance by addition, not by four divisions per tile
  int dg_n, dg_w, dg_h, dg_i;
  long long* debug;          // optional [gridDim.x][8] cycle counters (FYC_TC_DEBUG diagnostics), else nullptr
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// --- CTA-pair (cta_group::2) variants.  A CTA's 32-bit shared address carries its cluster rank in bit 24; clearing it names the
// same offset in the leader (rank 0) CTA (cute::Sm100MmaPeerBitMask).
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the leader's earlier MMAs have retired
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// epilogue warp -> MMA issuer: this warp has drained its part of the accumulator (8 arrivals per CTA)
template <int PAIR>
__device__ __forceinline__ void epi_release(uint64_t* tempty_bar, int lane) {
  tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty_bar); else mbar_arrive(tempty_bar); }
}

// D[tmem] (+)= A[smem] * B[smem]^T   (both operands K-major, 128B swizzle)
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// wait for the TMEM loads; naming the registers as in/out operands keeps their uses below the wait
__device__ __forceinline__ void tmem_ld_wait32(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address        bits [0,14)
  d |= (uint64_t)0 << 16;                            // leading byte offset  bits [16,30)  (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version 1 (Blackwell) bits [46,48)
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B         bits [61,64)
  return d;
}

// ---------------------------------------------------------------------------------------------- tile coordinates
// A CTA (or CTA pair) walks tiles t0, t0 + stride, ... < num.  Plain: tile = m_blk * n_tiles + n_blk.  Pair: the cluster's tile names
// (pair of m blocks, n block) and CTA `rank` owns m block 2 * pm + rank (a phantom block past the end is zero-filled by TMA and never
// stored).
struct TileSched { int64_t t0, stride, num; int rank; };
struct TileCoord { int64_t t; int n, w, h, i; };   // linear index; n block; patch column / row / image-group of the 128-pixel m block
template <int PAIR>
__device__ __forceinline__ TileSched tile_sched(const TcParams& p) {
  TileSched s;
  if (PAIR) {
    uint32_t r; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    s.rank = (int)r; s.t0 = blockIdx.x >> 1; s.stride = gridDim.x >> 1; s.num = ((p.m_tiles + 1) >> 1) * p.n_tiles;
  } else { s.rank = 0; s.t0 = blockIdx.x; s.stride = gridDim.x; s.num = p.m_tiles * p.n_tiles; }
  return s;
}
__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, const TileSched& ts, int64_t t) {
  TileCoord c;
  c.t = t;
  const uint32_t tile = (uint32_t)t;
  c.n = (int)(tile % (uint32_t)p.n_tiles);
  uint32_t m_blk = tile / (uint32_t)p.n_tiles;
  if (p.pair) m_blk = 2u * m_blk + (uint32_t)ts.rank;
  c.w = (int)(m_blk % (uint32_t)p.w_tiles);
  const uint32_t m2 = m_blk / (uint32_t)p.w_tiles;
  c.h = (int)(m2 % (uint32_t)p.h_tiles);
  c.i = (int)(m2 / (uint32_t)p.h_tiles);
  return c;
}
// t += stride
__device__ __forceinline__ void tile_advance(const TcParams& p, const TileSched& ts, TileCoord& c) {
  if (p.pair) { c = tile_coord(p, ts, c.t + ts.stride); return; }
  c.t += ts.stride;
  c.n += p.dg_n; int cy = c.n >= p.n_tiles ? 1 : 0; c.n -= cy ? p.n_tiles : 0;
  c.w += p.dg_w + cy; cy = c.w >= p.w_tiles ? 1 : 0; c.w -= cy ? p.w_tiles : 0;
  c.h += p.dg_h + cy; cy = c.h >= p.h_tiles ? 1 : 0; c.h -= cy ? p.h_tiles : 0;
  c.i += p.dg_i + cy;
}

// ---------------------------------------------------------------------------------------------- plain bf16 epilogue
// Two phases per 32-column group of the accumulator, through a per-warp 4 KB fp32 staging tile (32 rows x 32 columns,
// 16-byte chunks XOR-swizzled by row):
//   P2  thread = row   : one tcgen05.ld.x32 and 8 conflict-free 16-byte shared stores          (pure TMEM -> smem transpose)
//   P3  4 lanes = row  : lane owns 8 fixed columns; alpha and bias are one FFMA, bias / row bias / residual arrive in
//                        registers that were loaded one group AHEAD (across tile boundaries too), the output leaves as
//                        64-byte row segments, 8 rows per instruction, with a single bf16 rounding.
// Why: with 226 KB of shared memory the L1 is nearly gone, so every bias / residual load is an L2 round trip (~700 clk),
// and with two epilogue warps per scheduler there is no one to hide it - issued at their point of use these loads made a
// 32-column group cost ~1800 clk and left the MMA warp waiting on tmem-empty for half of the K = 320 GEMMs.  The phase
// is also instruction-bound (8 warps share the 4 schedulers with nobody else), so addresses are 32-bit offsets in 16-byte
// units computed once per tile, not 64-bit pixel * ld products per store.
struct EpiRows {
  uint32_t oo[4], ro[4];      // (row ps*8+rip, column n0 + q*8) of out / residual, in 16-byte units (< 2^32, checked at launch)
  int rgu;                    // the warp's common row-bias group, or -1 when its 32 rows straddle two groups
  uint32_t ok;                // bit ps: the row exists
  int n0;                     // first column of the tile
};
struct EpiPrefetch { float4 b0, b1; uint4 res[4]; float4 rb0, rb1; };      // rb*: raw row bias, carried only by the LNF + ROWB variant
// Forces every load of `f` to have landed HERE.  The group loop rotates `pf = pn` with register moves, which wait for the loads of `pn` at the
// END of a group; a prefetch that is still in flight on a path INTO the loop (first tile, a warp half without groups) makes ptxas guard the first
// use of `pf` inside the loop with the same scoreboard the in-loop prefetch re-arms - and every group then waits, in its middle, for the loads it
// issued itself (ncu, round 2 call L: 9 % of all samples of the K = 320 residual GEMM sat on that wait).
template <bool LNF>
__device__ __forceinline__ void epi_settle(const EpiPrefetch& f) {
  asm volatile("" ::"f"(f.b0.x), "f"(f.b0.y), "f"(f.b0.z), "f"(f.b0.w), "f"(f.b1.x), "f"(f.b1.y), "f"(f.b1.z), "f"(f.b1.w));
  if constexpr (!LNF)         // (an LN-folded GEMM carries no residual)
    asm volatile("" ::"r"(f.res[0].x), "r"(f.res[0].y), "r"(f.res[0].z), "r"(f.res[0].w), "r"(f.res[1].x), "r"(f.res[1].y), "r"(f.res[1].z),
                 "r"(f.res[1].w), "r"(f.res[2].x), "r"(f.res[2].y), "r"(f.res[2].z), "r"(f.res[2].w), "r"(f.res[3].x), "r"(f.res[3].y),
                 "r"(f.res[3].z), "r"(f.res[3].w));
}
// LayerNorm fold: the 4 rows' rstd, fetched with the row offsets one tile ahead (the mean term is already in the accumulator, see fyc.h)
struct EpiLnRows { float rs[4]; };

template <bool LNF, bool ROWB>
__device__ __forceinline__ void epi_rows(const TcParams& p, const TileCoord& tc, int r0, int q, EpiRows& t, EpiLnRows& ln) {
  const int n_blk = tc.n, wt = tc.w, ht = tc.h, it = tc.i;
  t.n0 = (int)n_blk * p.BN;
  t.ok = 0; t.rgu = -1;
  const uint32_t cq = (uint32_t)(t.n0 + q * 8) >> 3, ldo8 = (uint32_t)(p.ldo >> 3), ldr8 = (uint32_t)(p.ldr >> 3);
  int mn = 0x7fffffff, mx = -1;
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    // (w, h, image) of patch row r0 + 8 ps: bw, bh are powers of two (pick_patch), so three shifts / masks per row per tile - cheaper than
    // four live registers in a loop that is at the register cap (they were being spilled: a local-memory round trip per tile)
    const int r = r0 + ps * 8;
    const int ow = (int)wt * p.bw + (r & (p.bw - 1)), oh = (int)ht * p.bh + ((r >> p.lg_bw) & (p.bh - 1));
    const int img = (int)it * p.bn + (r >> (p.lg_bw + p.lg_bh));
    const int64_t pix = ((int64_t)img * p.Ho + oh) * p.Wo + ow;
    const bool ok = (ow < p.Wo) && (oh < p.Ho) && (pix < p.M);
    t.ok |= (ok ? 1u : 0u) << ps;
    const uint32_t px = ok ? (uint32_t)pix : 0u;
    t.oo[ps] = px * ldo8 + cq;
    t.ro[ps] = px * ldr8 + cq;
    if constexpr (LNF) ln.rs[ps] = ok ? __ldg(p.ln_rs + px) : 0.f;
    if constexpr (ROWB) {
      if (ok) {
        const int rg = (int)(px / (uint32_t)p.rows_per_group);
        mn = min(mn, rg); mx = max(mx, rg);
      }
    }
  }
  if constexpr (ROWB) {
    mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
    t.rgu = (mn == mx) ? mn : -1;
  }
}

// loads for 32-column group g of tile t: bias (+ the warp-uniform row bias) of this lane's 8 columns, residual of its 4 rows.
// ROWB is a template parameter, not a flag test: as predicated-off `@!P FADD bias, bias, rowbias` instructions the fold below still carried the
// scoreboard wait for the bias load that precedes it, i.e. every GEMM WITHOUT a row bias waited for its bias right after issuing the "prefetch"
// (ncu, round 2 call L: another 9 % of the samples of the K = 320 residual GEMM).
template <bool LNF, bool ROWB>
__device__ __forceinline__ void epi_prefetch(const TcParams& p, const EpiRows& t, int g, int q, EpiPrefetch& f) {
  const int c = g * 32 + q * 8, n = t.n0 + c;
  const bool col_ok = (c < p.BN) && (n < p.N);
  f.b0 = f.b1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((p.flags & FYC_EPI_BIAS) && col_ok) {
    f.b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    f.b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
  }
  if constexpr (ROWB) {
    // the warp's 32 rows share one row-bias vector (time embedding of a clip, position-table row of a frame): fetched here, one group AHEAD like
    // the bias - at its point of use this load was an exposed L2 round trip per 32-column group.  LNF (the temporal q/k/v projections, whose
    // epilogue is the bottleneck at K = 320) keeps it raw and folds it when the registers rotate (epi_fold); the other row-bias users (conv1 +
    // time embedding: MMA-bound, and their residual registers are live) fold at once.
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
    if (col_ok && t.rgu >= 0) {
      const float* rbp = p.rowbias + (int64_t)t.rgu * p.ldrb + n;
      r0 = __ldg(reinterpret_cast<const float4*>(rbp)); r1 = __ldg(reinterpret_cast<const float4*>(rbp + 4));
    }
    if constexpr (LNF) { f.rb0 = r0; f.rb1 = r1; }
    else {
      f.b0.x += r0.x; f.b0.y += r0.y; f.b0.z += r0.z; f.b0.w += r0.w;
      f.b1.x += r1.x; f.b1.y += r1.y; f.b1.z += r1.z; f.b1.w += r1.w;
    }
  }
  if constexpr (!LNF) {      // (an LN-folded GEMM never carries a residual: q/k/v and FF1 projections - its registers go to the LN terms)
    if (p.flags & FYC_EPI_RESIDUAL) {
      const uint4* rbase = reinterpret_cast<const uint4*>(p.residual);
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        f.res[ps] = make_uint4(0, 0, 0, 0);
        if (col_ok && ((t.ok >> ps) & 1u)) f.res[ps] = __ldg(rbase + (t.ro[ps] + (uint32_t)g * 4u));
      }
    }
  }
}
// bias += raw row bias (LNF + ROWB only), at the point where the loads must have landed anyway
template <bool LNF, bool ROWB>
__device__ __forceinline__ void epi_fold(EpiPrefetch& f) {
  if constexpr (LNF && ROWB) {
    f.b0.x += f.rb0.x; f.b0.y += f.rb0.y; f.b0.z += f.rb0.z; f.b0.w += f.rb0.w;
    f.b1.x += f.rb1.x; f.b1.y += f.rb1.y; f.b1.z += f.rb1.z; f.b1.w += f.rb1.w;
  }
}

template <int PAIR, bool LNF, bool ROWB>
__device__ __forceinline__ void epilogue_plain(const TcParams& p, const TileSched& ts, uint8_t* stage, uint64_t* tfull, uint64_t* tempty,
                                               uint32_t tmem_base, int warp, int lane) {
  const int64_t num_tiles = ts.num;
  const int quarter = warp & 3, egroup = (warp - 2) >> 2;
  const int rip = lane >> 2, q = lane & 3;
  const int NG = (p.BN + 31) >> 5;
  const float alpha = p.alpha;
  const int r0 = quarter * 32 + rip;                // this lane's rows inside the 128-pixel patch: r0 + 8 ps
  // staging addresses: P2 writes row `lane`, P3 reads rows ps*8 + rip (row & 7 == rip for every ps)
  uint8_t* const srow = stage + lane * 128;
  const int sw = lane & 7;
  const uint8_t* const prow = stage + rip * 128;
  const int px0 = ((2 * q) ^ rip) << 4, px1 = ((2 * q + 1) ^ rip) << 4;
  uint4* const obase = reinterpret_cast<uint4*>(p.out);
  long long dbg_epi = 0;
  int acc = 0; uint32_t aphase = 0;
  int eg = egroup;                                  // the column half alternates per tile: NG is odd for N = 320 (3 + 2 groups)
  EpiRows cur, nxt;
  EpiPrefetch pf;
  EpiLnRows lrc, lrn;          // LNF only (unused otherwise: the compiler drops them)
  int64_t tile = ts.t0;
  TileCoord tcn = tile_coord(p, ts, tile);           // coordinates of the NEXT tile to be decoded
  if (tile < num_tiles) { epi_rows<LNF, ROWB>(p, tcn, r0, q, cur, lrc); epi_prefetch<LNF, ROWB>(p, cur, eg, q, pf); epi_fold<LNF, ROWB>(pf); epi_settle<LNF>(pf); }
  nxt = cur; lrn = lrc;
  for (; tile < num_tiles; tile += ts.stride) {
    mbar_wait(&tfull[acc], aphase);
    const long long te1 = p.debug ? clock64() : 0;
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * MAX_BN;
    const int64_t next_tile = tile + ts.stride;
    tile_advance(p, ts, tcn);
    if (next_tile < num_tiles) {
      epi_rows<LNF, ROWB>(p, tcn, r0, q, nxt, lrn);
      if (!LNF && (p.flags & FYC_EPI_RESIDUAL)) {             // pull the next tile's residual rows into L2 a whole tile ahead: lane q takes its
        const int g = (eg ^ 1) + 2 * q;             // q-th group, so one instruction per row covers all of this warp's groups
        if (g < NG && nxt.n0 + g * 32 < p.N) {
#pragma unroll
          for (int ps = 0; ps < 4; ++ps)
            if ((nxt.ok >> ps) & 1u) {
              const uint4* ptr = reinterpret_cast<const uint4*>(p.residual) + (nxt.ro[ps] - (uint32_t)q + (uint32_t)g * 4u);
              asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
            }
        }
      }
    } else nxt.ok = 0;                               // nothing follows: the prefetch below degenerates to (valid) bias loads
    if (eg >= NG) { epi_prefetch<LNF, ROWB>(p, nxt, eg ^ 1, q, pf); epi_fold<LNF, ROWB>(pf); epi_settle<LNF>(pf); }
    for (int gi = eg; gi < NG; gi += 2) {
      // ---- P2
      uint32_t rr[32];
      tmem_ld32(taddr + gi * 32, rr);
      // ---- prefetch of the group after this one (possibly the next tile's first) while the TMEM load is in flight
      EpiPrefetch pn;
      {
        const bool last = gi + 2 >= NG;
        EpiRows src = cur;
        if (last) src = nxt;
        epi_prefetch<LNF, ROWB>(p, src, last ? (eg ^ 1) : gi + 2, q, pn);
      }
      tmem_ld_wait32(rr);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(srow + ((c ^ sw) << 4)) = make_uint4(rr[4 * c], rr[4 * c + 1], rr[4 * c + 2], rr[4 * c + 3]);
      __syncwarp();
      // ---- P3
      const int c0 = gi * 32 + q * 8;
      const bool col_ok = (c0 < p.BN) && (cur.n0 + c0 < p.N);
      const uint32_t g4 = (uint32_t)gi * 4u;
      float a_eff = alpha;
      if constexpr (ROWB) {
        if (cur.rgu < 0) {                                                  // rare: rows straddle two groups - fold alpha and the row bias into the staged tile
                                                           // (the usual case, one group per warp, is part of the prefetched bias)
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            if (!(((cur.ok >> ps) & 1u) && col_ok)) continue;
            const uint32_t pix = (uint32_t)((((uint64_t)cur.oo[ps] << 3) - (uint32_t)(cur.n0 + q * 8)) / (uint64_t)p.ldo);
            const float* rbp = p.rowbias + (int64_t)(pix / (uint32_t)p.rows_per_group) * p.ldrb + cur.n0 + c0;
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(rbp)), r1 = __ldg(reinterpret_cast<const float4*>(rbp + 4));
            float4* s0 = reinterpret_cast<float4*>(const_cast<uint8_t*>(prow) + ps * 1024 + px0);
            float4* s1 = reinterpret_cast<float4*>(const_cast<uint8_t*>(prow) + ps * 1024 + px1);
            float4 x0 = *s0, x1 = *s1;
            x0.x = fmaf(x0.x, alpha, r0.x); x0.y = fmaf(x0.y, alpha, r0.y); x0.z = fmaf(x0.z, alpha, r0.z); x0.w = fmaf(x0.w, alpha, r0.w);
            x1.x = fmaf(x1.x, alpha, r1.x); x1.y = fmaf(x1.y, alpha, r1.y); x1.z = fmaf(x1.z, alpha, r1.z); x1.w = fmaf(x1.w, alpha, r1.w);
            *s0 = x0; *s1 = x1;
          }
          a_eff = 1.0f;
        }
      }
      float4 xs[4][2];
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        xs[ps][0] = *reinterpret_cast<const float4*>(prow + ps * 1024 + px0);
        xs[ps][1] = *reinterpret_cast<const float4*>(prow + ps * 1024 + px1);
      }
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const float4 x0 = xs[ps][0], x1 = xs[ps][1];
        float v[8];
        if constexpr (LNF) {       // LN(x) W^T + b = rstd * (x W"^T) + (beta W^T + b), W" row-centred: the bracket is the accumulator
          const float rs = lrc.rs[ps];
          v[0] = fmaf(x0.x, rs, pf.b0.x); v[1] = fmaf(x0.y, rs, pf.b0.y); v[2] = fmaf(x0.z, rs, pf.b0.z); v[3] = fmaf(x0.w, rs, pf.b0.w);
          v[4] = fmaf(x1.x, rs, pf.b1.x); v[5] = fmaf(x1.y, rs, pf.b1.y); v[6] = fmaf(x1.z, rs, pf.b1.z); v[7] = fmaf(x1.w, rs, pf.b1.w);
        } else {
          v[0] = fmaf(x0.x, a_eff, pf.b0.x); v[1] = fmaf(x0.y, a_eff, pf.b0.y); v[2] = fmaf(x0.z, a_eff, pf.b0.z); v[3] = fmaf(x0.w, a_eff, pf.b0.w);
          v[4] = fmaf(x1.x, a_eff, pf.b1.x); v[5] = fmaf(x1.y, a_eff, pf.b1.y); v[6] = fmaf(x1.z, a_eff, pf.b1.z); v[7] = fmaf(x1.w, a_eff, pf.b1.w);
        }
        if (!LNF && (p.flags & FYC_EPI_RESIDUAL)) {
          const uint32_t u[4] = {pf.res[ps].x, pf.res[ps].y, pf.res[ps].z, pf.res[ps].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { v[2 * i] += __uint_as_float(u[i] << 16); v[2 * i + 1] += __uint_as_float(u[i] & 0xffff0000u); }
        }
        if (((cur.ok >> ps) & 1u) && col_ok) {
          uint4 o;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
          o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
          obase[cur.oo[ps] + g4] = o;
        }
      }
      __syncwarp();
      pf = pn;
      epi_fold<LNF, ROWB>(pf);
    }
    epi_release<PAIR>(&tempty[acc], lane);
    if (p.debug) dbg_epi += clock64() - te1;
    if (++acc == 2) { acc = 0; aphase ^= 1; }
    cur = nxt;
    if constexpr (LNF) lrc = lrn;
    eg ^= 1;
  }
  if (p.debug && lane == 0 && (warp == 2 || warp == 6)) p.debug[blockIdx.x * 8 + 5 + (warp == 6)] = dbg_epi;
}

// ---------------------------------------------------------------------------------------------- GEGLU epilogue
// Columns [0,128) of the tile are `a`, [128,256) the matching `gate` (weight rows pre-interleaved per 128 outputs).
// Thread = row for the math (bias + a * gelu(gate) in fp32, one bf16 rounding), 32 `a` + 32 gate columns per TMEM round
// trip; the bf16 results go through a 32 x 64 staging tile so that global stores cover full 128-byte row segments.
// The 256 bias values of a tile are fetched one tile ahead (one register per epilogue thread), parked in shared memory
// and read back as broadcast LDS.128 - as global loads at their point of use they were an L2 round trip per chunk.
template <int PAIR, bool LNF>
__device__ __forceinline__ void epilogue_geglu(const TcParams& p, const TileSched& ts, uint8_t* stage, float* sbias, uint64_t* tfull,
                                               uint64_t* tempty, uint32_t tmem_base, int warp, int lane) {
  const int64_t num_tiles = ts.num;
  const int quarter = warp & 3, egroup = (warp - 2) >> 2;
  const int r = quarter * 32 + lane;
  const int wl = r % p.bw, hl = (r / p.bw) % p.bh, il = r / (p.bw * p.bh);
  const int et = (warp - 2) * 32 + lane;                 // 0..255: the bias element this thread carries
  const int sub = lane >> 3, ch8 = lane & 7;
  const int c0 = egroup * 64;                            // this warp's 64 output columns of the tile
  uint8_t* const srow = stage + lane * 128;
  uint4* const obase = reinterpret_cast<uint4*>(p.out);
  const uint32_t ldo8 = (uint32_t)(p.ldo >> 3);
  long long dbg_epi = 0;
  int acc = 0; uint32_t aphase = 0;
  int64_t tile = ts.t0;
  float bnext = 0.f;
  float rnext = 1.0f;                                      // LNF: rstd of this thread's row in the NEXT tile (fetched one tile ahead like the bias)
  TileCoord tc = tile_coord(p, ts, tile);
  auto row_rstd = [&](const TileCoord& c) -> float {
    const int ow = c.w * p.bw + wl, oh = c.h * p.bh + hl, img = c.i * p.bn + il;
    const int64_t px = ((int64_t)img * p.Ho + oh) * p.Wo + ow;
    return ((ow < p.Wo) && (oh < p.Ho) && (px < p.M)) ? __ldg(p.ln_rs + px) : 1.0f;
  };
  if (tile < num_tiles) {
    bnext = __ldg(p.bias + tc.n * 256 + et);
    if constexpr (LNF) rnext = row_rstd(tc);
  }
  for (; tile < num_tiles; tile += ts.stride) {
    const int n_blk = tc.n;
    const int ow = tc.w * p.bw + wl, oh = tc.h * p.bh + hl, img = tc.i * p.bn + il;
    const int64_t pix = ((int64_t)img * p.Ho + oh) * p.Wo + ow;
    const bool row_ok = (ow < p.Wo) && (oh < p.Ho) && (pix < p.M);
    const uint32_t rowoff = row_ok ? (uint32_t)pix * ldo8 : 0u;
    const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
    float* const sb = sbias + acc * 256;
    sb[et] = bnext;
    const float rstd = rnext;
    const int64_t next_tile = tile + ts.stride;
    tile_advance(p, ts, tc);                               // tc now describes next_tile; n_blk / ow / oh / img above are this tile's
    if (next_tile < num_tiles) {
      bnext = __ldg(p.bias + tc.n * 256 + et);
      if constexpr (LNF) rnext = row_rstd(tc);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");       // the 8 epilogue warps: bias of this tile visible, previous reads done
    mbar_wait(&tfull[acc], aphase);
    const long long te1 = p.debug ? clock64() : 0;
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * MAX_BN;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t ar[32], gr[32];
      tmem_ld32(taddr + c0 + hh * 32, ar);
      tmem_ld32(taddr + 128 + c0 + hh * 32, gr);
      tmem_ld_wait32(ar);
      tmem_ld_wait32(gr);
      const float* ba_p = sb + c0 + hh * 32;                // packed (interleaved) bias of the `a` columns; gate = +128
#pragma unroll
      for (int c = 0; c < 4; ++c) {                          // 8 output columns = one 16-byte chunk of bf16
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 ba = *reinterpret_cast<const float4*>(ba_p + c * 8 + i);
          const float4 bg = *reinterpret_cast<const float4*>(ba_p + 128 + c * 8 + i);
          const float av[4] = {ba.x, ba.y, ba.z, ba.w}, gv[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // LNF: the accumulator already holds x W'^T - mean * colsum (row-centred weights); one FMA applies rstd
            if constexpr (LNF) o[i + e] = fmaf(__uint_as_float(ar[c * 8 + i + e]), rstd, av[e]) * gelu_erf_fast(fmaf(__uint_as_float(gr[c * 8 + i + e]), rstd, gv[e]));
            else o[i + e] = (__uint_as_float(ar[c * 8 + i + e]) + av[e]) * gelu_erf_fast(__uint_as_float(gr[c * 8 + i + e]) + gv[e]);
          }
        }
        Vec8<bf16>::store(reinterpret_cast<bf16*>(srow + (((hh * 4 + c) ^ (lane & 7)) << 4)), o);
      }
    }
    __syncwarp();
    const uint32_t ncol8 = (uint32_t)(n_blk * 128 + c0 + ch8 * 8) >> 3;
    const bool col_ok = (int)(n_blk * 128) + c0 + ch8 * 8 < p.N_out;
#pragma unroll
    for (int itp = 0; itp < 8; ++itp) {
      const int rl = itp * 4 + sub;
      const uint32_t pl = __shfl_sync(0xffffffffu, rowoff, rl);
      if (((okmask >> rl) & 1u) && col_ok)
        obase[pl + ncol8] = *reinterpret_cast<const uint4*>(stage + rl * 128 + ((ch8 ^ (rl & 7)) << 4));
    }
    epi_release<PAIR>(&tempty[acc], lane);
    if (p.debug) dbg_epi += clock64() - te1;
    if (++acc == 2) { acc = 0; aphase ^= 1; }
  }
  if (p.debug && lane == 0 && (warp == 2 || warp == 6)) p.debug[blockIdx.x * 8 + 5 + (warp == 6)] = dbg_epi;
}

// ---------------------------------------------------------------------------------------------- kernel
template <int PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_a2,
               const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem) & 1023u) __trap();   // SWIZZLE_128B operands need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES);     // 256 B reserved; staging follows
  uint64_t* full = bars;                       // [MAX_STAGES]
  uint64_t* empty = bars + MAX_STAGES;         // [MAX_STAGES]
  uint64_t* tfull = bars + 2 * MAX_STAGES;     // [2]
  uint64_t* tempty = bars + 2 * MAX_STAGES + 2;
  uint64_t* wbar = bars + 2 * MAX_STAGES + 4;  // W-resident mode: the weight slab has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 5);
  // operand ring, three layouts:
  //   plain     [A 16 KB | W BN x 128 B (<= 32 KB)] x 4 stages
  //   resident  [W slab k_iters x BN x 128 B][A 16 KB x a_stages]
  //   pair      [A 16 KB | W half BN/2 x 128 B (<= 16 KB)] x 6 stages of 32 KB
  constexpr int pair = PAIR;
  const int nstages = p.resident ? p.a_stages : (pair ? 6 : STAGES);
  const uint32_t slab_kb = (uint32_t)p.BN * (BK * 2);
  const uint32_t a_base = p.resident ? (uint32_t)(p.taps * p.cin_blocks) * slab_kb : 0u;
  const uint32_t a_stride = p.resident ? (uint32_t)A_BYTES : (pair ? 2u * A_BYTES : (uint32_t)STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TileSched ts = tile_sched<PAIR>(p);
  const int k_iters = p.taps * p.cin_blocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a2)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    // tmem-empty: one arrival per epilogue warp; in pair mode the leader's barrier collects both CTAs' 8 warps
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], pair ? 16 : 8); }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // TMEM: all 512 columns (2 accumulator stages x 256); this warp also frees them
    if (pair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();           // the peer's barriers must exist before anything is signalled across the pair
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      // pair: the LEADER's full barrier counts both CTAs' bytes (2 x A + the two W halves); each CTA frees / refills its own stage
      const uint32_t tx_bytes = p.resident ? (uint32_t)A_BYTES : (pair ? 2u * A_BYTES + slab_kb : A_BYTES + slab_kb);
      if (p.resident) {   // grid % n_tiles == 0, so this CTA's n block never changes: load its weight slab once
        const int n_blk = (int)(blockIdx.x % p.n_tiles);
        mbar_expect_tx(wbar, (uint32_t)k_iters * slab_kb);
        for (int k = 0; k < k_iters; ++k)
          tma_load_3d(&map_w, wbar, smem + (uint32_t)k * slab_kb, (k % p.cin_blocks) * BK, k / p.cin_blocks, n_blk * p.BN);
      }
      for (int64_t tile = ts.t0; tile < ts.num; tile += ts.stride) {
        const TileCoord tc = tile_coord(p, ts, tile);
        const int ow0 = tc.w * p.bw, oh0 = tc.h * p.bh, img0 = tc.i * p.bn;
        const int wrow0 = tc.n * p.BN + (pair ? ts.rank * (p.BN >> 1) : 0);
        for (int tap = 0; tap < p.taps; ++tap) {
          for (int cb = 0; cb < p.cin_blocks; ++cb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + a_base + (uint32_t)stage * a_stride;
            const bool seg2 = cb >= p.cb_split;                    // second K segment: the other source tensor, its own channel origin
            const CUtensorMap* ma = seg2 ? &map_a2 : &map_a;
            const int ka = (seg2 ? cb - p.cb_split : cb) * BK;
            if (pair) {
              if (ts.rank == 0) mbar_expect_tx(&full[stage], tx_bytes);
              tma2_load_4d(ma, &full[stage], sa, ka, ow0 + p.tap_dx[tap], oh0 + p.tap_dy[tap], img0 + p.tap_img[tap]);
              tma2_load_3d(&map_w, &full[stage], sa + A_BYTES, cb * BK, tap, wrow0);
            } else {
              mbar_expect_tx(&full[stage], tx_bytes);
              tma_load_4d(ma, &full[stage], sa, ka, ow0 + p.tap_dx[tap], oh0 + p.tap_dy[tap], img0 + p.tap_img[tap]);
              if (!p.resident) tma_load_3d(&map_w, &full[stage], sa + A_BYTES, cb * BK, tap, wrow0);
            }
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer (pair mode: the leader CTA only)
    if (lane == 0 && ts.rank == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, K-major both, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((pair ? 2 * BM : BM) >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t aphase = 0;
      long long dbg_wfull = 0, dbg_wtempty = 0; const long long dbg_t0 = clock64();
      if (p.resident) mbar_wait(wbar, 0);
      for (int64_t tile = ts.t0; tile < ts.num; tile += ts.stride) {
        long long tw0 = clock64();
        mbar_wait(&tempty[acc], aphase ^ 1);
        dbg_wtempty += clock64() - tw0;
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * MAX_BN;
        for (int k = 0; k < k_iters; ++k) {
          tw0 = clock64();
          mbar_wait(&full[stage], phase);
          dbg_wfull += clock64() - tw0;
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + a_base + (uint32_t)stage * a_stride);
          const uint64_t a_desc = make_sw128_desc(sa);
          const uint64_t b_desc = make_sw128_desc(p.resident ? smem_u32(smem) + (uint32_t)k * slab_kb : sa + A_BYTES);
          const int nks = ((k + 1) % p.cin_blocks == 0) ? p.last_ksteps : BK / 16;     // last k block of a tap: skip the all-zero k-steps
          if (pair) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              if (kk < nks) umma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (k > 0 || kk > 0) ? 1u : 0u);
            tcgen05_commit_pair(&empty[stage]);                     // frees this stage in BOTH CTAs
            if (k == k_iters - 1) tcgen05_commit_pair(&tfull[acc]); // accumulator complete, both CTAs' epilogues may drain
          } else {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
              if (kk < nks) umma_bf16(d_tmem, a_desc + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (k > 0 || kk > 0) ? 1u : 0u);
            }
            tcgen05_commit(&empty[stage]);                     // frees the smem stage when these MMAs retire
            if (k == k_iters - 1) tcgen05_commit(&tfull[acc]); // accumulator complete
          }
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; aphase ^= 1; }
      }
      if (p.debug) { p.debug[blockIdx.x * 8 + 2] = dbg_wfull; p.debug[blockIdx.x * 8 + 3] = dbg_wtempty; p.debug[blockIdx.x * 8 + 4] = clock64() - dbg_t0; }
    }
  } else {
    // ================================================================== epilogue (warps 2..9)
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may access
    const int egroup = (warp - 2) >> 2;                // 0: warps 2-5, 1: warps 6-9 (same rows, other half of the columns)
    const int r = quarter * 32 + lane;                 // row of the tile handled by this thread
    uint8_t* stage = smem + OFF_STAGING + (warp - 2) * 4096;
    long long dbg_epi = 0;
    int acc = 0; uint32_t aphase = 0;
    const bool geglu = (p.flags & FYC_EPI_GEGLU) != 0;
    const bool out_f32 = (p.flags & FYC_EPI_OUT_F32) != 0;
    const bool lnf = (p.flags & FYC_EPI_LNFOLD) != 0;
    const bool rowb = (p.flags & FYC_EPI_ROWBIAS) != 0;
    if (geglu && lnf) epilogue_geglu<PAIR, true>(p, ts, stage, reinterpret_cast<float*>(smem + OFF_BIAS), tfull, tempty, tmem_base, warp, lane);
    else if (geglu) epilogue_geglu<PAIR, false>(p, ts, stage, reinterpret_cast<float*>(smem + OFF_BIAS), tfull, tempty, tmem_base, warp, lane);
    else if (!out_f32 && lnf && rowb) epilogue_plain<PAIR, true, true>(p, ts, stage, tfull, tempty, tmem_base, warp, lane);
    else if (!out_f32 && lnf) epilogue_plain<PAIR, true, false>(p, ts, stage, tfull, tempty, tmem_base, warp, lane);
    else if (!out_f32 && rowb) epilogue_plain<PAIR, false, true>(p, ts, stage, tfull, tempty, tmem_base, warp, lane);
    else if (!out_f32) epilogue_plain<PAIR, false, false>(p, ts, stage, tfull, tempty, tmem_base, warp, lane);
    else for (int64_t tile = ts.t0; tile < ts.num; tile += ts.stride) {
      const TileCoord tc = tile_coord(p, ts, tile);
      const int n_blk = tc.n, wt = tc.w, ht = tc.h, it = tc.i;
      // row r of the tile = (image il, row hl, col wl) of the patch
      const int wl = r % p.bw, hl = (r / p.bw) % p.bh, il = r / (p.bw * p.bh);
      const int ow = wt * p.bw + wl, oh = ht * p.bh + hl, img = it * p.bn + il;
      const int64_t pix = ((int64_t)img * p.Ho + oh) * p.Wo + ow;
      const bool row_ok = (ow < p.Wo) && (oh < p.Ho) && (pix < p.M);
      const int n0 = n_blk * p.BN;
      mbar_wait(&tfull[acc], aphase);
      const long long te1 = clock64();
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * MAX_BN;
      const float* rb = (p.flags & FYC_EPI_ROWBIAS) ? p.rowbias + (row_ok ? pix / p.rows_per_group : 0) * p.ldrb : nullptr;
      if (out_f32) {
        // fp32 output (attention scores of the VAE mid block): direct per-thread stores, 16-column chunks split between
        // the two warps of a lane quarter
        const int nch = p.BN >> 4, half = (nch + 1) >> 1;
        const int cb = egroup ? half : 0, ce = egroup ? nch : half;
        for (int ch = cb; ch < ce; ++ch) {
          float v[16];
          tmem_ld16(taddr + ch * 16, v);
          const int n = n0 + ch * 16;
          if (row_ok && n < p.N) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= p.alpha;
            if (p.flags & FYC_EPI_BIAS) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
              }
            }
            if (rb) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                float4 b = __ldg(reinterpret_cast<const float4*>(rb + n + i));
                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
              }
            }
            if (p.flags & FYC_EPI_RESIDUAL) {
              float f[16];
              Vec8<float>::load(reinterpret_cast<const float*>(p.residual) + pix * p.ldr + n, f);
              Vec8<float>::load(reinterpret_cast<const float*>(p.residual) + pix * p.ldr + n + 8, f + 8);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += f[i];
            }
            float* o = reinterpret_cast<float*>(p.out) + pix * p.ldo + n;
            Vec8<float>::store(o, v); Vec8<float>::store(o + 8, v + 8);
          }
        }
      }
      epi_release<PAIR>(&tempty[acc], lane);
      dbg_epi += clock64() - te1;
      if (++acc == 2) { acc = 0; aphase ^= 1; }
    }
    if (p.debug && out_f32 && lane == 0 && (warp == 2 || warp == 6)) p.debug[blockIdx.x * 8 + 5 + (warp == 6)] = dbg_epi;
  }
  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();           // the peer may still be signalling this CTA's barriers / reading its operands
  if (warp == 2) {
    tcgen05_fence_after();
    if (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// parity-plane split for stride-2 convolutions: x [NB, H, W, C] -> planes [4, NB, H/2, W/2, C], plane = 2*(h&1) + (w&1)
__global__ void space_to_planes_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int64_t NB, int64_t H, int64_t W, int64_t C) {
  const int64_t cv = C / 8, H2 = H / 2, W2 = W / 2;
  const int64_t total = NB * H * W * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t c = (i % cv) * 8; int64_t t = i / cv;
    int64_t w = t % W; t /= W;
    int64_t h = t % H; int64_t n = t / H;
    int64_t plane = 2 * (h & 1) + (w & 1);
    bf16* dst = out + ((((plane * NB + n) * H2) + h / 2) * W2 + w / 2) * C + c;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(x + i * 8);
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

int32_t encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  FYC_CHECK(fn != nullptr, "tcgen05 path: cuTensorMapEncodeTiled driver entry point unavailable");
  cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FYC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu %llu)", (int)r, rank,
            (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
            (unsigned long long)(rank > 3 ? dims[3] : 0));
  return FYC_OK;
}

int pick_bn(int64_t N, bool geglu) {
  if (geglu) return 256;
  for (int bn = 256; bn >= 64; bn -= 16)
    if (N % bn == 0) return bn;
  if (N <= 256) return (int)N;          // N % 16 == 0 checked by the caller
  return 256;                           // ragged last tile (guarded stores)
}

// Tile width, operand-staging mode and grid for one launch.
//  * default: the largest BN <= 256 dividing N (fewest re-reads of the A tiles through L2);
//  * few rounds of tiles per SM (small M): BN that minimises rounds x (BN + fixed cost) - a 2.2-round launch at BN = 256 runs as
//    3 full rounds, the same problem at BN = 144 as 4 rounds of 56 % the length (ragged last N tile is zero-filled by TMA and
//    guarded in the epilogue);
//  * small K (k_iters x BN x 128 B <= 128 KB) and many tiles per CTA: W-resident - the grid is rounded down to a multiple of
//    n_tiles so a CTA's n block is fixed, its weight slab is loaded once, and the ring streams only A (L2 -> SM traffic per tile
//    drops from (128 + BN) x K to 128 x K elements; the K = 320 GEMMs were bound by it).
void choose_tiles(TcParams& p, int* grid_out) {
  const int sms = fyc_sm_count();
  const bool geglu = (p.flags & FYC_EPI_GEGLU) != 0;
  const int k_iters = p.taps * p.cin_blocks;
  p.resident = 0; p.a_stages = STAGES;
  p.BN = pick_bn(p.N, geglu);
  p.n_tiles = (int)ceil_div64(p.N, p.BN);
  int64_t tiles = p.m_tiles * p.n_tiles;
  if (!geglu) {
    // W-resident candidate
    int rbn = 0;
    for (int min_stages = 4; min_stages >= 3 && !rbn; --min_stages)      // prefer a slab that leaves the full 4-stage A ring; settle for 3
      for (int bn = 256; bn >= 128; bn -= 16)
        if (p.N % bn == 0 && (int64_t)k_iters * bn * 128 <= RING_BYTES - min_stages * A_BYTES) { rbn = bn; break; }
    if (rbn) {
      const int nt = p.N / rbn;
      const int grid = (sms / nt) * nt;
      if (nt <= sms && grid * 16 >= sms * 15 && p.m_tiles * nt >= (int64_t)grid * 4) {
        p.resident = 1; p.BN = rbn; p.n_tiles = nt;
        int st = (int)((RING_BYTES - (int64_t)k_iters * rbn * 128) / A_BYTES);
        p.a_stages = st > MAX_STAGES ? MAX_STAGES : st;
        *grid_out = grid;
        return;
      }
    }
  }
  // CTA pairs (cta_group::2): per 128 x BN of output each SM stages (128 + BN / 2) x K operand elements instead of (128 + BN) x K and
  // reads half the B bytes from shared memory per MMA.  Measured on B200 (tests/diag_conv.py, FYC_TC_PAIR=0/1): +3..26 % for
  // 4096 <= M <= 65536 with K >= 640 (the 32x32 / 16x16 convolutions, the FF2 and K >= 1280 projections); -7..11 % on the M = 131072
  // level-0 shapes (long power-capped launches, where the M = 256 instruction runs at a lower clock for the same watts) and on
  // epilogue-bound small-K tiles (the two CTAs' epilogues gate each other).  FYC_TC_PAIR=0 / 2 forces it off / on where legal.
  const char* pair_e = getenv("FYC_TC_PAIR");
  const int pair_env = pair_e ? atoi(pair_e) : 1;
  const bool pair_ok = p.m_tiles >= 4 && k_iters >= 4 && sms >= 2;
  const bool pair_wins = p.M >= 4096 && p.M <= 65536 && k_iters >= 10;
  p.pair = (pair_ok && (pair_env == 2 || (pair_env == 1 && pair_wins))) ? 1 : 0;
  const int64_t units = p.pair ? (p.m_tiles + 1) / 2 : p.m_tiles;      // m blocks (or pairs of them) to schedule
  const int64_t slots = p.pair ? sms / 2 : sms;                        // CTAs (or clusters) to schedule them on
  if (!geglu) {
    const int64_t rounds0 = ceil_div64(units * p.n_tiles, slots);
    if (rounds0 < 8) {
      // per k block a tile costs max(MMA, operand feed) clocks: 128 x bn x 64 MACs at 4096 MAC/clk, (128 + W rows) x 128 B at
      // ~67 B/clk (W rows are halved per CTA in pair mode), plus a fixed per-tile overhead
      auto cost_of = [&](int bn, int64_t nt) {
        const int64_t mma = 2 * bn, feed = (int64_t)(1.91 * (128 + (p.pair ? bn / 2 : bn)));
        return ceil_div64(units * nt, slots) * ((mma > feed ? mma : feed) + 48);
      };
      int64_t best = cost_of(p.BN, p.n_tiles);
      for (int bn = 256; bn >= 64; bn -= 16) {
        if (bn > p.N) continue;
        const int64_t nt = ceil_div64(p.N, bn);
        const int64_t cost = cost_of(bn, nt);
        if (cost < best) { best = cost; p.BN = bn; p.n_tiles = (int)nt; }
      }
    }
  }
  tiles = units * p.n_tiles;
  const int64_t g = tiles < slots ? tiles : slots;
  *grid_out = (int)(p.pair ? 2 * g : g);
}

long long* g_tc_debug = nullptr;

int32_t launch_tc(const CUtensorMap& ma, const CUtensorMap& mw, TcParams p, int grid, cudaStream_t st, const CUtensorMap* ma2p = nullptr) {
  const CUtensorMap& ma2 = ma2p ? *ma2p : ma;
  if (!ma2p) p.cb_split = 0x7fffffff;
  FYC_CHECK((p.bw & (p.bw - 1)) == 0 && (p.bh & (p.bh - 1)) == 0 && p.bw > 0 && p.bh > 0, "tcgen05 GEMM: patch dims must be powers of two");
  if (p.last_ksteps < 1 || p.last_ksteps > BK / 16) p.last_ksteps = BK / 16;
  p.lg_bw = 0; while ((1 << p.lg_bw) < p.bw) ++p.lg_bw;
  p.lg_bh = 0; while ((1 << p.lg_bh) < p.bh) ++p.lg_bh;
  if (p.pair) {
    p.debug = g_tc_debug;
    static bool attr_set2 = false;
    if (!attr_set2) {
      FYC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set2 = true;
    }
    const int64_t tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
    FYC_CHECK(tiles < (1ll << 30) && p.M < (1ll << 31) && p.rows_per_group < (1ll << 31), "tcgen05 GEMM: problem exceeds the 32-bit tile index range");
    FYC_CHECK(p.M * (p.ldo / 8) + p.ldo / 8 < (1ll << 32) && p.M * (p.ldr / 8) + p.ldr / 8 < (1ll << 32),
              "tcgen05 GEMM: output larger than 64 GB is not addressable by the epilogue");
    FYC_CHECK(grid % 2 == 0 && grid >= 2, "tcgen05 GEMM: pair mode needs an even grid");
    p.dg_n = p.dg_w = p.dg_h = p.dg_i = 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    FYC_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<1>, ma, mw, ma2, p));
    return FYC_OK;
  }
  p.debug = g_tc_debug;
  static bool attr_set = false;
  if (!attr_set) {
    FYC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  int64_t tiles = p.m_tiles * p.n_tiles;
  FYC_CHECK(tiles < (1ll << 31) && p.M < (1ll << 31) && p.rows_per_group < (1ll << 31), "tcgen05 GEMM: problem exceeds the 32-bit tile index range");
  FYC_CHECK(p.M * (p.ldo / 8) + p.ldo / 8 < (1ll << 32) && p.M * (p.ldr / 8) + p.ldr / 8 < (1ll << 32),
            "tcgen05 GEMM: output larger than 64 GB is not addressable by the epilogue");
  if (grid > tiles) grid = (int)tiles;
  {
    int64_t g = grid;
    p.dg_n = (int)(g % p.n_tiles); g /= p.n_tiles;
    p.dg_w = (int)(g % p.w_tiles); g /= p.w_tiles;
    p.dg_h = (int)(g % p.h_tiles); g /= p.h_tiles;
    p.dg_i = (int)g;
  }
  gemm_tc_kernel<0><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mw, ma2, p);
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}

}  // namespace

// diagnostics only (not part of include/fyc.h): per-CTA cycle counters of the next launches are written to buf[grid][8]
extern "C" void fyc_debug_tc_counters(long long* buf) { g_tc_debug = buf; }

extern "C" int32_t fyc_tcgen05_available(void) { return get_encode_fn() != nullptr ? 1 : 0; }

// Is this GEMM eligible for the tensor-core path?
bool fyc_gemm_tc_eligible(const fyc_gemm_args* g) {
  if (g->dtype != FYC_BF16) return false;
  if (g->K % 8 || g->lda % 8 || g->ldw % 8 || g->N % 16 || g->M < 64) return false;
  if (((uintptr_t)g->A | (uintptr_t)g->W) & 15) return false;
  if (g->batch > 1 && ((g->strideA | g->strideW | g->strideO) % 8)) return false;
  const int64_t n_out = (g->epilogue & FYC_EPI_GEGLU) ? g->N / 2 : g->N;
  if (g->ldo % 8 || ((uintptr_t)g->out & 15)) return false;
  if ((g->epilogue & FYC_EPI_RESIDUAL) && (g->ldr % 8 || ((uintptr_t)g->residual & 15))) return false;
  if ((g->epilogue & FYC_EPI_GEGLU) && (g->N % 256 || !(g->epilogue & FYC_EPI_BIAS) || (g->epilogue & ~(FYC_EPI_GEGLU | FYC_EPI_BIAS | FYC_EPI_LNFOLD)))) return false;
  if (g->A2) {
    if (g->batch != 1 || g->K1 <= 0 || g->K1 >= g->K || g->K1 % BK || g->lda2 % 8 || (((uintptr_t)g->A2) & 15)) return false;
  }
  if (g->epilogue & FYC_EPI_LNFOLD) {
    if (!g->ln_rowstats || (((uintptr_t)g->ln_rowstats) & 3)) return false;
    if (g->alpha != 1.0f || (g->epilogue & (FYC_EPI_OUT_F32 | FYC_EPI_RESIDUAL)) || g->N % 8) return false;
    if ((g->epilogue & FYC_EPI_ROWBIAS) && g->rows_per_group % 128) return false;     // a warp's 32 rows never straddle two row-bias groups
  }
  if ((g->epilogue & FYC_EPI_BIAS) && ((uintptr_t)g->bias & 15)) return false;
  if ((g->epilogue & FYC_EPI_ROWBIAS) && ((uintptr_t)g->rowbias & 15)) return false;
  (void)n_out;
  return get_encode_fn() != nullptr;
}

int32_t fyc_gemm_tc(const fyc_gemm_args* g, cudaStream_t st) {
  FYC_CHECK(fyc_gemm_tc_eligible(g), "gemm(tcgen05): shape/alignment not eligible (M=%lld N=%lld K=%lld)", (long long)g->M,
            (long long)g->N, (long long)g->K);
  const bool geglu = (g->epilogue & FYC_EPI_GEGLU) != 0;
  const bool f32 = (g->epilogue & FYC_EPI_OUT_F32) != 0;
  for (int64_t b = 0; b < g->batch; ++b) {
    const bf16* A = (const bf16*)g->A + b * g->strideA;
    const bf16* W = (const bf16*)g->W + b * g->strideW;
    CUtensorMap ma, mw, ma2;
    const int64_t Ka = g->A2 ? g->K1 : g->K;          // columns of the first (or only) source
    {
      uint64_t dims[4] = {(uint64_t)Ka, (uint64_t)g->M, 1, 1};
      uint64_t str[3] = {(uint64_t)g->lda * 2, (uint64_t)g->lda * 2 * (uint64_t)g->M, (uint64_t)g->lda * 2 * (uint64_t)g->M};
      uint32_t box[4] = {BK, BM, 1, 1};
      int32_t rc = encode_map(&ma, A, 4, dims, str, box);
      if (rc) return rc;
    }
    if (g->A2) {
      uint64_t dims[4] = {(uint64_t)(g->K - g->K1), (uint64_t)g->M, 1, 1};
      uint64_t str[3] = {(uint64_t)g->lda2 * 2, (uint64_t)g->lda2 * 2 * (uint64_t)g->M, (uint64_t)g->lda2 * 2 * (uint64_t)g->M};
      uint32_t box[4] = {BK, BM, 1, 1};
      int32_t rc = encode_map(&ma2, g->A2, 4, dims, str, box);
      if (rc) return rc;
    }
    TcParams p{};
    p.M = g->M; p.N = (int)g->N; p.N_out = geglu ? (int)g->N / 2 : (int)g->N;
    p.taps = 1; p.cin_blocks = (int)ceil_div64(g->K, BK);
    p.bw = BM; p.bh = 1; p.bn = 1; p.Wo = (int)g->M; p.Ho = 1;
    p.w_tiles = (int)ceil_div64(g->M, BM); p.h_tiles = 1; p.m_tiles = p.w_tiles;
    p.flags = g->epilogue;
    int grid = 0;
    choose_tiles(p, &grid);
    p.tap_dy[0] = p.tap_dx[0] = p.tap_img[0] = 0;
    {
      uint64_t dims[3] = {(uint64_t)g->K, 1, (uint64_t)g->N};
      uint64_t str[2] = {(uint64_t)g->ldw * 2, (uint64_t)g->ldw * 2};
      uint32_t box[3] = {BK, 1, (uint32_t)(p.pair ? p.BN / 2 : p.BN)};   // pair mode: each CTA stages half of the tile's W rows
      int32_t rc = encode_map(&mw, W, 3, dims, str, box);
      if (rc) return rc;
    }
    FYC_CHECK(g->M < (1ll << 31), "gemm(tcgen05): M too large");
    p.bias = g->bias; p.rowbias = g->rowbias; p.rows_per_group = g->rows_per_group > 0 ? g->rows_per_group : 1; p.ldrb = g->N;
    p.residual = g->residual ? (f32 ? (const void*)((const float*)g->residual + b * g->strideO) : (const void*)((const bf16*)g->residual + b * g->strideO)) : nullptr;
    p.out = f32 ? (void*)((float*)g->out + b * g->strideO) : (void*)((bf16*)g->out + b * g->strideO);
    p.ldo = g->ldo; p.ldr = g->ldr; p.alpha = g->alpha; p.flags = g->epilogue;
    p.ln_rs = g->ln_rowstats;
    p.cb_split = g->A2 ? (int)(g->K1 / BK) : 0x7fffffff;
    p.last_ksteps = (g->K % BK) ? (int)ceil_div64(g->K % BK, 16) : BK / 16;
    int32_t rc = launch_tc(ma, mw, p, grid, st, g->A2 ? &ma2 : nullptr);
    if (rc) return rc;
  }
  return FYC_OK;
}

// Tile shape for a conv output of Ho x Wo over NB images: bw*bh*bn == 128, all dividing evenly.
static bool pick_patch(int64_t NB, int64_t Ho, int64_t Wo, int* bw, int* bh, int* bn) {
  int w = 1; while (w < 128 && Wo % (w * 2) == 0) w *= 2;
  int h = 1; while (w * h < 128 && Ho % (h * 2) == 0) h *= 2;
  int n = 128 / (w * h);
  if (w * h * n != 128 || NB % n != 0) return false;
  *bw = w; *bh = h; *bn = n;
  return true;
}

bool fyc_conv3x3_tc_eligible(const fyc_conv3x3_args* c) {
  if (c->dtype != FYC_BF16 || c->upsample != 1) return false;
  if (c->stride != 1 && c->stride != 2) return false;
  if (c->Cin % 8 || c->Cout % 16) return false;
  if (c->stride == 2 && (c->H % 2 || c->W % 2)) return false;
  if (((uintptr_t)c->x | (uintptr_t)c->w | (uintptr_t)c->out) & 15) return false;
  if ((c->epilogue & FYC_EPI_RESIDUAL) && ((uintptr_t)c->residual & 15)) return false;
  if ((c->epilogue & FYC_EPI_ROWBIAS) && (((uintptr_t)c->rowbias & 15) || (c->ld_rowbias > 0 && c->ld_rowbias % 4))) return false;
  if (c->epilogue & FYC_EPI_GEGLU) return false;
  int bw, bh, bn;
  if (!pick_patch(c->NB, c->H / c->stride, c->W / c->stride, &bw, &bh, &bn)) return false;
  return get_encode_fn() != nullptr;
}

// x for stride 2 must already be the parity-plane split (see fyc_space_to_planes); H, W are the ORIGINAL dims.
int32_t fyc_conv3x3_tc(const fyc_conv3x3_args* c, const void* x_planes, cudaStream_t st) {
  FYC_CHECK(fyc_conv3x3_tc_eligible(c), "conv3x3(tcgen05): shape/alignment not eligible");
  const int s = c->stride;
  const int64_t Ho = c->H / s, Wo = c->W / s;
  const bool f32 = (c->epilogue & FYC_EPI_OUT_F32) != 0;
  TcParams p{};
  pick_patch(c->NB, Ho, Wo, &p.bw, &p.bh, &p.bn);
  p.M = c->NB * Ho * Wo; p.N = (int)c->Cout; p.N_out = p.N;
  p.taps = 9; p.cin_blocks = (int)ceil_div64(c->Cin, BK);
  p.last_ksteps = (c->Cin % BK) ? (int)ceil_div64(c->Cin % BK, 16) : BK / 16;          // e.g. the 16-channel stem: one k-step per tap instead of four
  p.Wo = (int)Wo; p.Ho = (int)Ho; p.w_tiles = (int)(Wo / p.bw); p.h_tiles = (int)(Ho / p.bh);
  p.m_tiles = (int64_t)p.w_tiles * p.h_tiles * (c->NB / p.bn);
  p.flags = c->epilogue;
  int grid = 0;
  choose_tiles(p, &grid);
  CUtensorMap ma, mw;
  const void* xa = c->x;
  uint64_t imgs = (uint64_t)c->NB;
  for (int t = 0; t < 9; ++t) {
    int kh = t / 3, kw = t % 3;
    if (s == 1) { p.tap_dy[t] = kh - 1; p.tap_dx[t] = kw - 1; p.tap_img[t] = 0; }
    else if (c->pad_mode == 1) {   // ih = 2*oh + kh: kh=0 -> even plane, row oh; kh=1 -> odd plane, row oh; kh=2 -> even plane, row oh+1
      int ph = (kh == 1) ? 1 : 0, pw = (kw == 1) ? 1 : 0;       // (row H/2 of a plane is out of bounds: TMA zero fill = the bottom / right pad)
      p.tap_dy[t] = (kh == 2) ? 1 : 0; p.tap_dx[t] = (kw == 2) ? 1 : 0;
      p.tap_img[t] = (2 * ph + pw) * (int)c->NB;
    } else {   // ih = 2*oh + kh - 1: kh=0 -> odd plane, row oh-1; kh=1 -> even plane, row oh; kh=2 -> odd plane, row oh
      int ph = (kh == 1) ? 0 : 1, pw = (kw == 1) ? 0 : 1;
      p.tap_dy[t] = (kh == 0) ? -1 : 0; p.tap_dx[t] = (kw == 0) ? -1 : 0;
      p.tap_img[t] = (2 * ph + pw) * (int)c->NB;
    }
  }
  if (s == 2) { xa = x_planes; imgs = 4ull * c->NB; FYC_CHECK(x_planes != nullptr, "conv3x3(tcgen05): stride 2 needs the plane-split input"); }
  {
    uint64_t dims[4] = {(uint64_t)c->Cin, (uint64_t)Wo, (uint64_t)Ho, imgs};
    uint64_t str[3] = {(uint64_t)c->Cin * 2, (uint64_t)c->Cin * 2 * Wo, (uint64_t)c->Cin * 2 * Wo * Ho};
    uint32_t box[4] = {BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int32_t rc = encode_map(&ma, xa, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)c->Cin, 9, (uint64_t)c->Cout};
    uint64_t str[2] = {(uint64_t)c->Cin * 2, (uint64_t)c->Cin * 2 * 9};
    uint32_t box[3] = {BK, 1, (uint32_t)(p.pair ? p.BN / 2 : p.BN)};   // pair mode: each CTA stages half of the tile's W rows
    int32_t rc = encode_map(&mw, c->w, 3, dims, str, box);
    if (rc) return rc;
  }
  p.bias = c->bias; p.rowbias = c->rowbias; p.residual = c->residual; p.out = c->out;
  p.ldrb = c->ld_rowbias > 0 ? c->ld_rowbias : c->Cout;
  p.rows_per_group = (c->images_per_group > 0 ? c->images_per_group : 1) * Ho * Wo;
  p.ldo = c->Cout; p.ldr = c->Cout; p.alpha = 1.0f; p.flags = c->epilogue;
  (void)f32;
  return launch_tc(ma, mw, p, grid, st);
}

// nearest-x2 upsample + padded 3x3 conv as four 2x2-tap implicit GEMMs on the low-resolution image (fyc.h: w_phases).
// Phase (py, px) produces output pixels (2*oh + py, 2*ow + px).  No kernel change is needed: the A boxes are the usual shifted
// patches of x (TMA zero fill = the conv's zero padding, because an upsampled halo pixel is out of bounds exactly when its source
// pixel is), and the interleaved destination is expressed through the epilogue's own address arithmetic - it computes
// pix = (img * Ho + oh) * Wo + ow and stores at out + pix * ldo: with Wo := 2 W, ldo := 2 Cout and out advanced by
// (py * 2W + px) * Cout that is ((img * 2H + 2 oh + py) * 2W + 2 ow + px) * Cout, the NHWC offset of the upsampled pixel.
bool fyc_conv3x3_up2_tc_eligible(const fyc_conv3x3_args* c) {
  if (c->dtype != FYC_BF16 || c->upsample != 2 || c->stride != 1 || c->pad_mode != 0 || !c->w_phases) return false;
  if (c->Cin % 8 || c->Cout % 16) return false;
  if (((uintptr_t)c->x | (uintptr_t)c->w_phases | (uintptr_t)c->out) & 15) return false;
  if (c->epilogue & ~FYC_EPI_BIAS) return false;          // the upsamplers carry a bias only (resnet.py:168, diffusers resnet.py:139)
  int bw, bh, bn;
  if (!pick_patch(c->NB, c->H, c->W, &bw, &bh, &bn)) return false;
  return get_encode_fn() != nullptr;
}

int32_t fyc_conv3x3_up2_tc(const fyc_conv3x3_args* c, cudaStream_t st) {
  FYC_CHECK(fyc_conv3x3_up2_tc_eligible(c), "conv3x3 up2(tcgen05): shape/alignment not eligible");
  const int64_t H = c->H, W = c->W;
  TcParams p{};
  pick_patch(c->NB, H, W, &p.bw, &p.bh, &p.bn);
  p.M = c->NB * H * W; p.N = (int)c->Cout; p.N_out = p.N;
  p.taps = 4; p.cin_blocks = (int)ceil_div64(c->Cin, BK);
  p.last_ksteps = (c->Cin % BK) ? (int)ceil_div64(c->Cin % BK, 16) : BK / 16;
  p.Wo = (int)W; p.Ho = (int)H; p.w_tiles = (int)(W / p.bw); p.h_tiles = (int)(H / p.bh);
  p.m_tiles = (int64_t)p.w_tiles * p.h_tiles * (c->NB / p.bn);
  p.flags = c->epilogue;
  int grid = 0;
  choose_tiles(p, &grid);                       // tile shape / staging mode from the true (low-resolution) problem size
  p.Wo = (int)(2 * W); p.M = c->NB * H * 2 * W;  // epilogue addressing only (see above); every ow < W < Wo, every pix < M
  p.ldo = 2 * c->Cout; p.ldr = p.ldo; p.alpha = 1.0f;
  p.bias = c->bias; p.rowbias = nullptr; p.residual = nullptr; p.rows_per_group = 1; p.ldrb = p.N;
  CUtensorMap ma;
  {
    uint64_t dims[4] = {(uint64_t)c->Cin, (uint64_t)W, (uint64_t)H, (uint64_t)c->NB};
    uint64_t str[3] = {(uint64_t)c->Cin * 2, (uint64_t)c->Cin * 2 * W, (uint64_t)c->Cin * 2 * W * H};
    uint32_t box[4] = {BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int32_t rc = encode_map(&ma, c->x, 4, dims, str, box);
    if (rc) return rc;
  }
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    for (int t = 0; t < 4; ++t) {               // tap (a, b): input row oh + a - 1 + py, column ow + b - 1 + px
      p.tap_dy[t] = (t >> 1) - 1 + py; p.tap_dx[t] = (t & 1) - 1 + px; p.tap_img[t] = 0;
    }
    CUtensorMap mw;
    const bf16* wp = (const bf16*)c->w_phases + (int64_t)ph * c->Cout * 4 * c->Cin;
    uint64_t dims[3] = {(uint64_t)c->Cin, 4, (uint64_t)c->Cout};
    uint64_t str[2] = {(uint64_t)c->Cin * 2, (uint64_t)c->Cin * 2 * 4};
    uint32_t box[3] = {BK, 1, (uint32_t)(p.pair ? p.BN / 2 : p.BN)};
    int32_t rc = encode_map(&mw, wp, 3, dims, str, box);
    if (rc) return rc;
    p.out = (bf16*)c->out + ((int64_t)py * 2 * W + px) * c->Cout;
    rc = launch_tc(ma, mw, p, grid, st);
    if (rc) return rc;
  }
  return FYC_OK;
}

int32_t fyc_space_to_planes(const void* x, void* out, int64_t NB, int64_t H, int64_t W, int64_t C, cudaStream_t st) {
  FYC_CHECK(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "space_to_planes: H, W must be even and C a multiple of 8");
  int64_t total = NB * H * W * C / 8;
  int64_t blocks = ceil_div64(total, 256), cap = (int64_t)fyc_sm_count() * 16;
  space_to_planes_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, st>>>((const bf16*)x, (bf16*)out, NB, H, W, C);
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}
