// bf16 GEMM on the 5th-generation tensor cores (sm_100a): D[M,N] = epilogue(alpha * sum_k A(m,k) B(n,k)).
//
// This one kernel carries every dense contraction of the policy / value / discriminator / VAE MLPs
// (network_builder.py:105-124, amp_network_builder.py:58-249, amp_network_z_builder.py:341-467):
//   forward   Y  = act(X W^T + b)            A = X  [M,K] K-major,   B = W  [N,K] K-major
//   dgrad     dX = (dY W) * act'(.)          A = dY [M,N] K-major,   B = W  [N,K] read MN-major
//   wgrad     dW = dY^T X                    A = dY [M,N] MN-major,  B = X  [M,K] MN-major   (reduction over the batch rows)
// Operands are read as they sit in memory: the UMMA descriptors' major bits select K-major or MN-major, nothing is transposed.
//
// Structure (persistent: one CTA -- or one CTA PAIR, see GemmSmemT -- per SM loops over output tiles x split-K slices; the
// fp32 accumulator is double-buffered in TMEM so one item's epilogue overlaps the next item's main loop):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads (128B swizzle) into a shared-memory ring that runs
//               continuously across work items, completion on "full" mbarriers;
//   warp 1      allocates the 512 TMEM columns, then one elected lane issues tcgen05.mma (M128/M256 x N256 x K16, fp32
//               accumulate in TMEM) four times per stage and tcgen05.commit's the stage back to the producer ("empty") and,
//               after the last k-block, the accumulator to the epilogue;
//   warps 2..9  epilogue (two warps per TMEM lane quarter, 128 accumulator columns each): tcgen05.ld 32 lanes x 32 columns at
//               a time -> bias / activation / activation-derivative gate / column sums -> bf16 and/or fp32 outputs, or
//               coalesced fp32 atomics into the weight-gradient buffer (split-K).
// What bounds it and why the epilogue looks the way it does: DESIGN.md section 3.3 (measured with the phase-trace build).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_bf16.h>

#include "pulse_common.cuh"

namespace pulse {
namespace {

#ifndef PULSE_GEMM_VARIANT
#define PULSE_GEMM_VARIANT 0   // 0 = product; 3 = phase-trace build for tools/gemm_trace.py (tools/build_variant.sh trace gemm_tcgen05.cu -DPULSE_GEMM_VARIANT=3)
#endif
constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
// Epilogue warps per CTA: kEpi / 4 per TMEM lane quarter, each draining 256 / (kEpi / 4) accumulator columns in 32-column chunks.
// Measured (round 2, profiles/r02_gemm_epilogue_warps.txt): 16 epilogue warps on the CTA-pair kernel are SLOWER than 8 (PPO update set
// 765 vs 687 us; ReLU-dgrad 16384x1024x512 52.8 vs 36 us): the epilogue is not short of warps, it is short of shared-memory pipe -- at
// full MMA rate the UMMA operand reads (64 B/clk/SM) plus the TMA fills (64 B/clk/SM) already take the whole 128 B/clk, so every LDS /
// STS / SHFL of the epilogue queues behind them.  The lever is fewer shared-pipe operations per output, not more warps.
#ifndef PULSE_GEMM_EPI_PAIR
#define PULSE_GEMM_EPI_PAIR 8
#endif
template <int CTAS>
struct EpiCfg {
  static constexpr int kWarps = CTAS == 2 ? PULSE_GEMM_EPI_PAIR : 8;
  static constexpr int kCols = 256 / (kWarps / 4);      // accumulator columns per epilogue warp
  static constexpr int kChunks = kCols / 32;            // 32-column chunks per epilogue warp
  static constexpr int kThreads = 64 + 32 * kWarps;
};
constexpr unsigned kStageBytesA = BM * BK * 2;
constexpr unsigned kTmemCols = 512;  // two 256-column fp32 accumulators (all of TMEM)

// CTAS = 1: one CTA per 128 x 256 tile, 4 x 48 KB ring.  CTAS = 2: a CTA PAIR (cluster of two SMs, tcgen05 cta_group::2) per
// 256 x 256 tile -- each CTA stages its own 128 rows of A and HALF of B (128 of the 256 columns), the leader's MMA reads both
// halves, so the L2 -> SM operand traffic per MAC drops by a third (the single-CTA kernel was measured at the L2 delivery
// limit: 8.9 TB/s of operand reads at 47 % tensor-pipe activity) and the 32 KB stages allow a 6-deep ring.
// WG (weight-gradient specialisation): the epilogue stages each warp's 32 x 32 fp32 block in a 4 KB, 128-byte-swizzled tile that ONE
// bulk tensor reduction (cp.reduce.async.bulk.tensor ... .add) adds into the gradient matrix -- double-buffered per warp (64 KB), paid for
// with ring stages (the weight-gradient main loops are long, 4 / 3 stages cover the L2 latency).
template <int CTAS, bool WG>
struct __align__(1024) GemmSmemT {
  static constexpr int kEpiWarps = EpiCfg<CTAS>::kWarps;
  static constexpr int kStages = WG ? (CTAS == 2 ? 4 : 3) : (CTAS == 2 ? (kEpiWarps > 8 ? 5 : 6) : 4);
  static constexpr unsigned kStageBytesB = (BN / CTAS) * BK * 2;
  static constexpr int kRedFloats = WG ? 2 * 1024 : 16 * 33;
  unsigned char a[kStages][kStageBytesA];
  unsigned char b[kStages][kStageBytesB];
  float red[kEpiWarps][kRedFloats];   // per-epilogue-warp tile: bf16 store staging / fp32 transpose for atomics / (WG) two 4 KB reduction tiles
  float bias[kEpiWarps][EpiCfg<CTAS>::kCols];   // per-epilogue-warp copy of the bias of its columns (broadcast reads in the forward epilogue)
  unsigned long long full[kStages];
  unsigned long long empty[kStages];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  unsigned tmem_base;
};

__device__ __forceinline__ unsigned s_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void g_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ bool g_mbar_try(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(s_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void g_mbar_wait(unsigned long long* bar, unsigned parity) {
  for (int spin = 0; spin < (1 << 26); ++spin)
    if (g_mbar_try(bar, parity)) return;
  __trap();  // a protocol bug must surface as a launch error, never as a hung GPU
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                   s_u32(smem_dst)),
               "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024 B between 8-row groups | version 1 [46,48) | SWIZZLE_128B (2) [61,64)
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= static_cast<unsigned long long>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<unsigned long long>(1u) << 16;
  d |= static_cast<unsigned long long>(1024u >> 4) << 32;
  d |= static_cast<unsigned long long>(1u) << 46;
  d |= static_cast<unsigned long long>(2u) << 61;
  return d;
}
// MN-major operand (the NON-reduction dimension is the contiguous one, e.g. dY[batch, n] or W[n, k_out] read as
// the operand of a reduction over its rows): the stage holds two [64 reduction rows][64 elements] boxes
// (8 KB each, 128-byte rows, 128B swizzle).  Canonical layout (cute make_umma_desc<Major::MN>, B128):
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -> LBO = 8192 B between the 64-element MN chunks,
// SBO = 1024 B between 8-row reduction groups.
__device__ __forceinline__ unsigned long long umma_desc_mn(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= static_cast<unsigned long long>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<unsigned long long>(8192u >> 4) << 16;
  d |= static_cast<unsigned long long>(1024u >> 4) << 32;
  d |= static_cast<unsigned long long>(1u) << 46;
  d |= static_cast<unsigned long long>(2u) << 61;
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10),
// a_major / b_major at bits 15 / 16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr unsigned instr_desc(bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         (static_cast<unsigned>(BN >> 3) << 17) | (static_cast<unsigned>(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s_u32(bar)) : "memory");
}
// ---- CTA-pair (cta_group::2) variants -------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// The pair's barriers live in the LEADER (cluster rank 0): clearing bit 24 of a shared::cluster address turns the peer's
// window address into the leader's (the offset inside the CTA is the same) -- the convention cta_group::2 TMA loads use.
constexpr unsigned kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                   s_u32(smem_dst)),
               "l"(map), "r"(s_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same offset in BOTH CTAs of the pair when the preceding MMAs retire
__device__ __forceinline__ void umma_commit_pair(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(s_u32(bar)),
               "h"(static_cast<unsigned short>(3))
               : "memory");
}
// arrive on the LEADER's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(unsigned long long* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(s_u32(bar))
      : "memory");
}
template <int UM>
__host__ __device__ constexpr unsigned instr_desc_m(bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         (static_cast<unsigned>(BN >> 3) << 17) | (static_cast<unsigned>(UM >> 4) << 24);
}

// issue only: the registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == PULSE_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == PULSE_ACT_SILU) return __fdividef(x, 1.0f + __expf(-x));   // 2 MUFU ops; the IEEE division subroutine dominated the SiLU epilogue
  return x;
}
__device__ __forceinline__ float act_grad(float g, int mode) {
  // g: saved tensor -- ReLU: the layer's OUTPUT (>0 <=> active); SiLU: the layer's PRE-activation z
  if (mode == PULSE_ACT_RELU) return g > 0.0f ? 1.0f : 0.0f;
  if (mode == PULSE_ACT_SILU) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-g));
    return s * (1.0f + g * (1.0f - s));
  }
  return 1.0f;
}


// Coalesced bf16 store of one epilogue warp's 32-row x 32-column block.  A lane holds 32 columns (64 bytes) of ITS row, so
// a direct 16-byte store instruction would touch 32 different rows (32 half-used sectors; measured 45.9 -> 41.6 us on the
// 16384 x 1024 x 934 layer).  The block goes through the warp's private 2 KB shared tile instead (16-byte units,
// XOR-swizzled so both the row-wise writes and the 8-rows-at-a-time reads are bank-conflict free) and leaves as
// 8 rows x 64 contiguous bytes per instruction.  All 32 lanes must call it; rows >= rows_valid are not written.
__device__ __forceinline__ void store_block_bf16(uint4* st, const float (&v)[32], __nv_bfloat16* base, long long ld, int rows_valid, int lane) {
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = 8 * q;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i], v[i + 1]), h1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]), h3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
    uint4 u;
    u.x = *reinterpret_cast<unsigned*>(&h0);
    u.y = *reinterpret_cast<unsigned*>(&h1);
    u.z = *reinterpret_cast<unsigned*>(&h2);
    u.w = *reinterpret_cast<unsigned*>(&h3);
    st[lane * 4 + (q ^ sw)] = u;
  }
  __syncwarp();
  __nv_bfloat16* obase = base + (lane & 3) * 8;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + (lane >> 2);
    const uint4 val = st[rr * 4 + ((lane & 3) ^ ((rr >> 1) & 3))];
    if (rr < rows_valid) *reinterpret_cast<uint4*>(obase + static_cast<long long>(rr) * ld) = val;
  }
  __syncwarp();
}

#if PULSE_GEMM_VARIANT == 3   // development: per-phase clock64 trace of CTA 0 (tools/gemm_trace.py)
__device__ long long g_gemm_trace[32];
#define PULSE_TRACE(slot)                                          \
  do {                                                             \
    if (blockIdx.x == 0) g_gemm_trace[slot] = clock64();           \
  } while (0)
#else
#define PULSE_TRACE(slot) \
  do {                    \
  } while (0)
#endif

// silu(z) = z / (1 + e^-z) on a bf16 pair, evaluated in fp32 (ex2.approx + rcp.approx per element) and rounded ONCE to bf16.
// Round 1 used tanh.approx.bf16x2 (one MUFU per two elements): measured at im_z_fit.yaml widths its error is not zero-mean --
// the encoder / prior heads came out 0.24 % small after four SiLU layers, kin_KLD 0.6 % low (tests/test_gpu_vae.py, full width).
__device__ __forceinline__ __nv_bfloat162 silu_bf16x2(__nv_bfloat162 z) {
  float2 f = __bfloat1622float2(z);
  f.x = __fdividef(f.x, 1.0f + __expf(-f.x));
  f.y = __fdividef(f.y, 1.0f + __expf(-f.y));
  return __floats2bfloat162_rn(f.x, f.y);
}

// A_MN / B_MN: operand is MN-major in global memory ([reduction rows, non-reduction cols] row-major) instead of K-major.
// Persistent: one CTA per SM loops over (tile, split) work items.  The TMA ring runs continuously across
// items; the accumulator is double-buffered in TMEM (2 x 128 columns) so the epilogue of item i overlaps the
// main loop of item i+1.
// MODE selects which epilogue features are COMPILED IN.  The fully general epilogue is ~7.4k SASS instructions per
// instance and the three warp roles run disjoint parts of it: measured, growing it by 260 instructions that never execute
// slowed the forward layers from 39.6 to 54.6 us (instruction-cache misses).  Each specialisation carries only what its
// caller can ask for; the host picks the smallest one that covers the request.
enum : int { kModeGeneric = 0, kModeFwd = 1, kModeDgrad = 2, kModeWgrad = 3, kModeDgradVec = 4 };

// ---- grouped launch: several problems of the same operand majors / epilogue mode in ONE persistent launch ------------------
constexpr int kMaxGroup = 4;
struct GemmProblem {
  CUtensorMap map_a, map_b, map_c;
  pulse_gemm_epilogue_t ep;
  int M, N, K, kb_per_split;
  int item_end;   // cumulative work items (tiles x split-K slices) up to and including this problem
  int use_tma_red;
  int pad[2];
};
struct GemmGroup {
  GemmProblem p[kMaxGroup];
  int count, total_items;
};

#define PULSE_GEMM_GROUPED 0
#include "gemm_kernel.inc"
#undef PULSE_GEMM_GROUPED
#define PULSE_GEMM_GROUPED 1
#include "gemm_kernel.inc"
#undef PULSE_GEMM_GROUPED

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency) ------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rows x cols bf16 matrix, row stride ld elements, box `box_rows` rows x 64 cols, 128B swizzle, OOB reads return 0.
// K-major operand: rows = M (or N), cols = K, box 128 x 64.  MN-major operand: rows = K (reduction), cols = M (or N), box 64 x 64.
bool make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, unsigned box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 [rows, cols] output (row stride ld floats) as a tensor map with 32 x 32 boxes, 128-byte swizzle: the target of the weight-gradient
// epilogue's bulk tensor reductions (out-of-range rows / columns of a box are clipped by the hardware).
bool make_map_c(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// the reduction path needs 16-byte aligned rows; otherwise the kernel falls back to fp32 atomics
bool tma_reduce_ok(const pulse_gemm_epilogue_t& ep) {
  return ep.out_f32 != nullptr && ep.accumulate && (ep.ldf % 4) == 0 && (reinterpret_cast<uintptr_t>(ep.out_f32) % 16) == 0;
}

template <bool A_MN, bool B_MN, int MODE, int CTAS>
int launch_gemm(const CUtensorMap& map_a, const CUtensorMap& map_b, const pulse_gemm_epilogue_t& ep, int m, int n, int k, int splits,
                int kb_per_split, cudaStream_t stream) {
  static bool attr_set = false;
  const size_t smem = sizeof(GemmSmemT<CTAS, MODE == kModeWgrad>) + 1024;  // slack so the kernel can align the ring to 1024 B
  CUtensorMap map_c;
  memset(&map_c, 0, sizeof(map_c));
  int use_tma_red = 0;
  if (MODE == kModeWgrad && tma_reduce_ok(ep)) {
    if (!make_map_c(&map_c, ep.out_f32, m, n, ep.ldf)) {
      set_error("pulse_gemm_bf16: cuTensorMapEncodeTiled failed for the fp32 output");
      return PULSE_ERR_CUDA;
    }
    use_tma_red = 1;
  }
  if (!attr_set) {
    PULSE_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<A_MN, B_MN, MODE, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    PULSE_CUDA_OK(cudaGetDevice(&dev));
    PULSE_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    // PULSE_GEMM_SMS=n: leave SMs free for a concurrent collective (the persistent grid otherwise owns the whole GPU)
    const char* e = getenv("PULSE_GEMM_SMS");
    if (e != nullptr && atoi(e) >= 2 && atoi(e) < num_sms) num_sms = atoi(e) & ~1;
  }
  // persistent: one CTA (or CTA pair) per SM (pair of SMs) loops over the work items
  const long long total = static_cast<long long>((n + BN - 1) / BN) * ((m + BM * CTAS - 1) / (BM * CTAS)) * splits;
  const long long slots = num_sms / CTAS;
  const unsigned grid = static_cast<unsigned>((total < slots ? total : slots) * CTAS);
  static int use_pdl = -1;
  if (use_pdl < 0) {
    const char* e = getenv("PULSE_GEMM_PDL");
    use_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(EpiCfg<CTAS>::kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CTAS == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  PULSE_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<A_MN, B_MN, MODE, CTAS>, map_a, map_b, map_c, ep, m, n, k, kb_per_split, splits, use_tma_red));
  PULSE_LAUNCH_OK("gemm_bf16_kernel");
  return PULSE_OK;
}

template <bool A_MN, bool B_MN, int MODE, int CTAS>
int launch_gemm_grouped(const GemmGroup& grp, cudaStream_t stream) {
  static bool attr_set = false;
  const size_t smem = sizeof(GemmSmemT<CTAS, MODE == kModeWgrad>) + 1024;
  if (!attr_set) {
    PULSE_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_grouped_kernel<A_MN, B_MN, MODE, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    PULSE_CUDA_OK(cudaGetDevice(&dev));
    PULSE_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long slots = num_sms / CTAS;
  const unsigned grid = static_cast<unsigned>((grp.total_items < slots ? grp.total_items : slots) * CTAS);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(EpiCfg<CTAS>::kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CTAS == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  PULSE_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16_grouped_kernel<A_MN, B_MN, MODE, CTAS>, grp));
  PULSE_LAUNCH_OK("gemm_bf16_grouped_kernel");
  return PULSE_OK;
}

int epilogue_mode(const pulse_gemm_epilogue_t* ep) {
  const bool want_fwd = ep->bias || ep->act != PULSE_ACT_NONE || ep->preact || ep->out_t || ep->relu_mask;
  const bool want_dgrad = ep->gate || ep->colsum || ep->sumsq || ep->gate_mask;
  const bool want_accum = ep->accumulate != 0;
  if (!want_dgrad && !want_accum) return kModeFwd;
  if (!want_fwd && !want_accum) return (ep->gate && ep->gate_mode != PULSE_ACT_RELU) ? kModeDgradVec : kModeDgrad;
  if (!want_fwd && !want_dgrad && !ep->out) return kModeWgrad;
  return kModeGeneric;
}

}  // namespace
}  // namespace pulse

#if PULSE_GEMM_VARIANT == 3
extern "C" int pulse_debug_gemm_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, pulse::g_gemm_trace, sizeof(long long) * 32) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int pulse_gemm_num_splits(int64_t k, int32_t split_k) {
  const int num_kb = static_cast<int>((k + 63) / 64);
  int splits = split_k < 1 ? 1 : (split_k > num_kb ? num_kb : split_k);
  const int kb_per_split = (num_kb + splits - 1) / splits;
  return (num_kb + kb_per_split - 1) / kb_per_split;  // every slab gets at least one k-block
}

// flags: bit 0 = A is MN-major ([K, M] row-major in memory), bit 1 = B is MN-major ([K, N] row-major in memory)
extern "C" int pulse_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                               const pulse_gemm_epilogue_t* ep, int32_t split_k, uint32_t flags, void* stream) {
  using namespace pulse;
  PULSE_REQUIRE(a && b && ep, "pulse_gemm_bf16: null argument");
  PULSE_REQUIRE(m > 0 && n > 0 && k > 0, "pulse_gemm_bf16: empty problem %lld x %lld x %lld", (long long)m, (long long)n, (long long)k);
  PULSE_REQUIRE(m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31), "pulse_gemm_bf16: dimension too large");
  PULSE_REQUIRE((flags & ~3u) == 0, "pulse_gemm_bf16: bad flags");
  const bool a_mn = flags & PULSE_GEMM_A_MN, b_mn = flags & PULSE_GEMM_B_MN;
  PULSE_REQUIRE(lda >= (a_mn ? m : k) && ldb >= (b_mn ? n : k) && (lda % 8) == 0 && (ldb % 8) == 0,
                "pulse_gemm_bf16: leading dimensions must cover the contiguous extent and be multiples of 8 (16-byte rows), got %lld %lld",
                (long long)lda, (long long)ldb);
  PULSE_REQUIRE(aligned16(a) && aligned16(b), "pulse_gemm_bf16: operands must be 16-byte aligned");
  PULSE_REQUIRE(ep->out || ep->out_t || ep->out_f32, "pulse_gemm_bf16: no output requested");
  PULSE_REQUIRE(split_k >= 1, "pulse_gemm_bf16: split_k must be >= 1");
  PULSE_REQUIRE(split_k == 1 || (ep->out_f32 && !ep->out && !ep->out_t && !ep->bias && ep->act == PULSE_ACT_NONE && !ep->gate && !ep->preact && !ep->colsum),
                "pulse_gemm_bf16: split-K only supports plain fp32 outputs (slabs or atomic accumulation)");
  PULSE_REQUIRE(ep->gate == nullptr || ep->gate_mode == PULSE_ACT_RELU || ep->gate_mode == PULSE_ACT_SILU, "pulse_gemm_bf16: bad gate_mode");
  PULSE_REQUIRE((ep->relu_mask == nullptr || ep->ld_rmask >= m) && (ep->gate_mask == nullptr || ep->ld_gmask >= m),
                "pulse_gemm_bf16: mask word rows must hold at least M entries");
  PULSE_REQUIRE(!(ep->gate_mask && ep->gate), "pulse_gemm_bf16: give the ReLU gate either as bf16 activations or as bit words, not both");
  // CTA pairs (256 x 256 tiles, cta_group::2) for everything large enough to fill them; PULSE_GEMM_PAIR=0 forces single CTAs
  static int use_pair = -1;
  if (use_pair < 0) {
    const char* e = getenv("PULSE_GEMM_PAIR");
    use_pair = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const bool pair = use_pair && m > 128 && n > 128;
  CUtensorMap map_a, map_b;
  const bool ok_a = a_mn ? make_map(&map_a, a, k, m, lda, 64) : make_map(&map_a, a, m, k, lda, BM);
  const bool ok_b = b_mn ? make_map(&map_b, b, k, n, ldb, 64) : make_map(&map_b, b, n, k, ldb, pair ? BN / 2 : BN);
  if (!ok_a || !ok_b) {
    set_error("pulse_gemm_bf16: cuTensorMapEncodeTiled failed (driver entry point missing or bad strides)");
    return PULSE_ERR_CUDA;
  }
  const int num_kb = static_cast<int>((k + BK - 1) / BK);
  const int splits = pulse_gemm_num_splits(k, split_k);
  const int kb_per_split = (num_kb + splits - 1) / splits;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // smallest epilogue specialisation that covers the request (see the MODE comment on the kernel)
  const bool want_fwd = ep->bias || ep->act != PULSE_ACT_NONE || ep->preact || ep->out_t || ep->relu_mask;
  const bool want_dgrad = ep->gate || ep->colsum || ep->sumsq || ep->gate_mask;
  const bool want_accum = ep->accumulate != 0;
  int mode = kModeGeneric;
  if (!want_dgrad && !want_accum) mode = kModeFwd;
  else if (!want_fwd && !want_accum) mode = (ep->gate && ep->gate_mode != PULSE_ACT_RELU) ? kModeDgradVec : kModeDgrad;
  else if (!want_fwd && !want_dgrad && !ep->out) mode = kModeWgrad;
  const int mi = (int)m, ni = (int)n, ki = (int)k;
#define PULSE_GEMM_DISPATCH(AM, BM_)                                                                                          \
  if (pair) {                                                                                                                 \
    switch (mode) {                                                                                                           \
      case kModeFwd: return launch_gemm<AM, BM_, kModeFwd, 2>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st);           \
      case kModeDgrad: return launch_gemm<AM, BM_, kModeDgrad, 2>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st);       \
      case kModeWgrad: return launch_gemm<AM, BM_, kModeWgrad, 2>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st);       \
      case kModeDgradVec: return launch_gemm<AM, BM_, kModeDgradVec, 2>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st); \
      default: return launch_gemm<AM, BM_, kModeGeneric, 2>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st);             \
    }                                                                                                                         \
  }                                                                                                                           \
  switch (mode) {                                                                                                             \
    case kModeFwd: return launch_gemm<AM, BM_, kModeFwd, 1>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st);             \
    case kModeDgrad: return launch_gemm<AM, BM_, kModeDgrad, 1>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st);         \
    case kModeWgrad: return launch_gemm<AM, BM_, kModeWgrad, 1>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st);         \
    case kModeDgradVec: return launch_gemm<AM, BM_, kModeDgradVec, 1>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st);   \
    default: return launch_gemm<AM, BM_, kModeGeneric, 1>(map_a, map_b, *ep, mi, ni, ki, splits, kb_per_split, st);               \
  }
  if (a_mn && b_mn) { PULSE_GEMM_DISPATCH(true, true) }
  if (a_mn) { PULSE_GEMM_DISPATCH(true, false) }
  if (b_mn) { PULSE_GEMM_DISPATCH(false, true) }
  PULSE_GEMM_DISPATCH(false, false)
#undef PULSE_GEMM_DISPATCH
}

extern "C" int pulse_gemm_bf16_nt(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                                  const pulse_gemm_epilogue_t* ep, int32_t split_k, void* stream) {
  return pulse_gemm_bf16(a, lda, b, ldb, m, n, k, ep, split_k, 0u, stream);
}

// Several GEMMs of the same kind (operand majors `flags`, same epilogue specialisation) in ONE persistent launch: the work items
// of all problems are concatenated, so the tail of one problem fills with tiles of the next and the ~8 us of per-launch
// prologue / drain is paid once (actor + critic layers of a PPO minibatch, or the weight gradients of several layers).
// EXPERIMENTAL in round 1: compiled, not yet run on a device; nothing calls it unless PULSE_GROUPED=1 (pulse_b200/dense.py).
extern "C" int pulse_gemm_bf16_grouped(const pulse_gemm_problem_t* problems, int32_t count, uint32_t flags, void* stream) {
  using namespace pulse;
  PULSE_REQUIRE(problems != nullptr && count >= 1 && count <= kMaxGroup, "pulse_gemm_bf16_grouped: 1..%d problems", kMaxGroup);
  PULSE_REQUIRE((flags & ~3u) == 0, "pulse_gemm_bf16_grouped: bad flags");
  const bool a_mn = flags & PULSE_GEMM_A_MN, b_mn = flags & PULSE_GEMM_B_MN;
  static int use_pair = -1;
  if (use_pair < 0) {
    const char* e = getenv("PULSE_GEMM_PAIR");
    use_pair = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  bool pair = use_pair != 0;
  int mode = -1;
  for (int i = 0; i < count; ++i) {
    const pulse_gemm_problem_t& q = problems[i];
    PULSE_REQUIRE(q.a && q.b, "pulse_gemm_bf16_grouped: null operand in problem %d", i);
    PULSE_REQUIRE(q.m > 0 && q.n > 0 && q.k > 0 && q.m < (1ll << 31) && q.n < (1ll << 31) && q.k < (1ll << 31), "pulse_gemm_bf16_grouped: bad shape in problem %d", i);
    PULSE_REQUIRE(q.lda >= (a_mn ? q.m : q.k) && q.ldb >= (b_mn ? q.n : q.k) && (q.lda % 8) == 0 && (q.ldb % 8) == 0,
                  "pulse_gemm_bf16_grouped: leading dimensions of problem %d must cover the contiguous extent and be multiples of 8", i);
    PULSE_REQUIRE(aligned16(q.a) && aligned16(q.b), "pulse_gemm_bf16_grouped: operands of problem %d must be 16-byte aligned", i);
    PULSE_REQUIRE(q.ep.out || q.ep.out_t || q.ep.out_f32, "pulse_gemm_bf16_grouped: problem %d has no output", i);
    PULSE_REQUIRE(q.split_k >= 1, "pulse_gemm_bf16_grouped: split_k must be >= 1");
    PULSE_REQUIRE(q.split_k == 1 || (q.ep.out_f32 && q.ep.accumulate && !q.ep.out && !q.ep.out_t && !q.ep.bias && q.ep.act == PULSE_ACT_NONE && !q.ep.gate &&
                                     !q.ep.preact && !q.ep.colsum),
                  "pulse_gemm_bf16_grouped: split-K only with fp32 atomic accumulation");
    const int mq = epilogue_mode(&q.ep);
    PULSE_REQUIRE(mode < 0 || mq == mode, "pulse_gemm_bf16_grouped: problem %d needs a different epilogue specialisation than problem 0", i);
    mode = mq;
    pair = pair && q.m > 128 && q.n > 128;
  }
  PULSE_REQUIRE(mode == kModeFwd || mode == kModeDgrad || mode == kModeWgrad, "pulse_gemm_bf16_grouped: forward, ReLU-dgrad and wgrad groups only");
  PULSE_REQUIRE((mode == kModeFwd && !a_mn && !b_mn) || (mode == kModeDgrad && !a_mn && b_mn) || (mode == kModeWgrad && a_mn && b_mn),
                "pulse_gemm_bf16_grouped: operand majors do not match the group kind (fwd K/K, dgrad K/MN, wgrad MN/MN)");
  GemmGroup grp;
  memset(&grp, 0, sizeof(grp));
  grp.count = count;
  long long items = 0;
  for (int i = 0; i < count; ++i) {
    const pulse_gemm_problem_t& q = problems[i];
    GemmProblem& g = grp.p[i];
    const bool ok_a = a_mn ? make_map(&g.map_a, q.a, q.k, q.m, q.lda, 64) : make_map(&g.map_a, q.a, q.m, q.k, q.lda, BM);
    const bool ok_b = b_mn ? make_map(&g.map_b, q.b, q.k, q.n, q.ldb, 64) : make_map(&g.map_b, q.b, q.n, q.k, q.ldb, pair ? BN / 2 : BN);
    if (!ok_a || !ok_b) {
      set_error("pulse_gemm_bf16_grouped: cuTensorMapEncodeTiled failed for problem %d", i);
      return PULSE_ERR_CUDA;
    }
    g.ep = q.ep;
    g.use_tma_red = 0;
    if (mode == kModeWgrad && tma_reduce_ok(q.ep)) {
      if (!make_map_c(&g.map_c, q.ep.out_f32, q.m, q.n, q.ep.ldf)) {
        set_error("pulse_gemm_bf16_grouped: cuTensorMapEncodeTiled failed for the fp32 output of problem %d", i);
        return PULSE_ERR_CUDA;
      }
      g.use_tma_red = 1;
    }
    g.M = static_cast<int>(q.m);
    g.N = static_cast<int>(q.n);
    g.K = static_cast<int>(q.k);
    const int num_kb = static_cast<int>((q.k + BK - 1) / BK);
    const int splits = pulse_gemm_num_splits(q.k, q.split_k);
    g.kb_per_split = (num_kb + splits - 1) / splits;
    const int bmt = pair ? 2 * BM : BM;
    items += static_cast<long long>((q.n + BN - 1) / BN) * ((q.m + bmt - 1) / bmt) * splits;
    PULSE_REQUIRE(items < (1ll << 30), "pulse_gemm_bf16_grouped: too many work items");
    g.item_end = static_cast<int>(items);
  }
  grp.total_items = static_cast<int>(items);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == kModeFwd) return pair ? launch_gemm_grouped<false, false, kModeFwd, 2>(grp, st) : launch_gemm_grouped<false, false, kModeFwd, 1>(grp, st);
  if (mode == kModeDgrad) return pair ? launch_gemm_grouped<false, true, kModeDgrad, 2>(grp, st) : launch_gemm_grouped<false, true, kModeDgrad, 1>(grp, st);
  return pair ? launch_gemm_grouped<true, true, kModeWgrad, 2>(grp, st) : launch_gemm_grouped<true, true, kModeWgrad, 1>(grp, st);
}
