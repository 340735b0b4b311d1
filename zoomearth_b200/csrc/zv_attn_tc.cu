// K4: tcgen05 flash attention for the full-attention layers (whole-image segments; HF modeling_qwen2_5_vl.py
// :207-287 with cu_seqlens, blocks 7/15/23/31).  One CTA = one 128-row q tile of one head; K/V stream through a
// TMA ring in 64-row tiles.  Both MMAs take their A operand from TENSOR memory:
//   Q               loaded once by the softmax threads (one q row each, 160 contiguous bytes from global memory) and
//                   parked in TMEM as 40 columns of 16-bit pairs
//   S_j = Q K_j^T   tcgen05.mma M=128 N=64, A = Q (TMEM), B = K_j (smem: a 128B-swizzled [rows][64] block for k 0..63
//                   and a 32B-swizzled [rows][16] block for k 64..79), fp32 accumulator in TMEM.  One S buffer: the
//                   softmax pulls a tile into registers as soon as it is complete, which frees the buffer for
//                   Q K_{j+1}^T while the exponentials run
//   softmax         four warps, one thread per q row: tcgen05.ld of the row, online max/sum in fp32 with exp2,
//                   P_j rounded to 16 bit and written back to TMEM (tcgen05.st, two values per column)
//   O += P_j V_j    tcgen05.mma M=128 K=64, A = P (TMEM), B = V_j straight from the V rows of qkv as an MN-major
//                   operand (head dim contiguous; the same 64 + 16 column TMA boxes as K, one N = 64 and one N = 16
//                   MMA per K step): no transposed copy of V exists.  O stays in TMEM for the whole segment and is
//                   rescaled lazily (only when the row max grows by more than 2^8).
// Why TMEM operands: with A read from shared memory every K=16 step streams 4 KB of Q or P next to 1-3 KB of K/V and the
// tensor pipe sat at 80 % busy on operand fetch alone (ncu sm__pipe_tc_cycles_active) while doing 40 % of its math rate;
// with A in TMEM only the K and V tiles cross the shared-memory port.
// K/V traffic: every q tile of a head streams the head's whole K and V.  CTAs run as clusters of two neighbouring q
// tiles of the same head: each CTA fetches half of every K / V tile and TMA-multicasts it into both CTAs' shared
// memory, which halves the L2 reads; a stage is recycled when BOTH tensor cores are done with it (multicast
// tcgen05.commit).  Pairs that straddle a segment boundary (or the odd last tile) fall back to private loads.
// Warp roles (192 threads): warp 0 = TMA producer + TMEM allocator, warp 1 = MMA issuer, warps 2-5 = softmax
// (TMEM lane quarters 2,3,0,1).  Two CTAs per SM (216 of 256 TMEM columns each).  Rotary is already applied to q,k by
// the QKV GEMM epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <type_traits>

#include "zv_common.h"
#include "zv_gemm.h"
#include "zv_ptx.cuh"

namespace zv {
using namespace ptx;
namespace {

constexpr int HD = 80, BQ = 128, BKV = 64, STAGES = 4, kCtasPerSm = 2;
constexpr int kK64 = BKV * 64 * 2, kK16 = BKV * 16 * 2;                   // 8192, 2048: [rows][64] and [rows][16] blocks
constexpr int kStage = 2 * (kK64 + kK16);                                // 20480: K and V tiles have the same shape
constexpr int kOffOnes = STAGES * kStage;                                // [16][64] 16-bit ones: B operand of the row-sum MMA
constexpr int kOnesBytes = 16 * 64 * 2;
constexpr int kOffBar = kOffOnes + kOnesBytes;
constexpr int kSmem = kOffBar + 256 + 1024;
// TMEM columns: S [0,64)  O [64,144)  P [144,176) 16-bit pairs  Q [176,216) 16-bit pairs  L [216,232) row sums of P
constexpr int kTmemCols = 256, kOCol = BKV, kPCol = kOCol + HD, kQCol = kPCol + BKV / 2, kLCol = kQCol + HD / 2;
// every kPolyEvery-th pair of exponentials is evaluated on the FMA pipe (Cody-Waite + a cubic) instead of MUFU.EX2: the
// softmax warps are bound by the 16 ex2 / clk / SM of the SFU (512 clk per 128 x 64 tile against 320 clk of MMA)
#ifndef ZV_ATTN_POLY_EVERY
#define ZV_ATTN_POLY_EVERY 4
#endif
constexpr int kPolyEvery = ZV_ATTN_POLY_EVERY;                           // 0 = all on the SFU
constexpr int kThreads = 192;
constexpr uint32_t kSw128 = 2, kSw32 = 6;                                // UMMA descriptor layout types
static_assert(kStage % 1024 == 0 && (kK64 + kK16) % 1024 == 0 && kK64 % 1024 == 0, "swizzle atom alignment");
static_assert(kLCol + 16 <= kTmemCols, "TMEM budget");
static_assert(kOffOnes % 1024 == 0, "swizzle atom alignment of the ones tile");
static_assert(STAGES <= 4, "mbarrier slots");

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (F16) { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load delivered to the same shared-memory offset (data and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// arrive, once all prior MMAs of this thread are done, on the same barrier in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T: A = 128 rows (lanes) x 16 K-values packed two per 32-bit column
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// 2^x for x in [-125, 125] on the FMA / ALU pipes: n = round(x) by the magic-number add, 2^f on [-0.5, 0.5] by a cubic
// (max rel err 7.5e-5, a sixth of the 16-bit rounding P gets anyway), exponent patched in with an integer add.
__device__ __forceinline__ float exp2_poly(float x) {
  const float t = x + 12582912.f;                           // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05517084f, 0.24260935f);              // minimax cubic of 2^f on [-0.5, 0.5] (relative error)
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992818f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

struct AttnArgs {
  void* out;
  const void* qkv;           // (S, 3 * hidden) 16-bit: the q rows are read directly, K goes through TMA
  int64_t S;
  const int4* tiles;
  int n_tiles;
  int heads, hidden, f16;
  float scale_log2;
};

template <bool F16>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) attn_tc_kernel(const __grid_constant__ CUtensorMap tm_k64,
                                                                       const __grid_constant__ CUtensorMap tm_k16, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars;             // Q parked in TMEM (4 softmax warps)
  uint64_t* k_full = bars + 1;         // STAGES (<= 4)
  uint64_t* v_full = bars + 5;
  uint64_t* k_empty = bars + 9;
  uint64_t* v_empty = bars + 13;
  uint64_t* s_full = bars + 17;
  uint64_t* s_empty = bars + 18;
  uint64_t* p_full = bars + 19;
  uint64_t* pv_done = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  // the pair shares K/V when both q tiles exist and belong to the same segment; otherwise each CTA loads privately
  const uint32_t rank = cluster_ctarank();
  const bool valid = (int)blockIdx.x < a.n_tiles;
  if (!valid) {                                  // odd tail of the grid: only keeps the partner's cluster barriers company
    cluster_sync_all();
    cluster_sync_all();
    return;
  }
  const int4 tl = a.tiles[blockIdx.x];
  const int q0 = tl.x, q_len = tl.y, seg_b = tl.z, seg_e = tl.w;
  bool shared_kv = false;
  if ((int)(blockIdx.x ^ 1u) < a.n_tiles) {
    const int4 to = a.tiles[blockIdx.x ^ 1u];
    shared_kv = to.z == seg_b && to.w == seg_e;
  }
  const int head = blockIdx.y;
  // K/V tiles start at the segment start; columns from seg_e on (last tile) are masked in the softmax.
  const int kv_base = seg_b;
  const int n_kv = (seg_e - kv_base + BKV - 1) / BKV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm_k64); prefetch_tensormap(&tm_k16);
    mbar_init(q_full, 4);
    const int users = shared_kv ? 2 : 1;          // tensor cores that must be done with a stage before it is refilled
    for (int s = 0; s < STAGES; ++s) { mbar_init(k_full + s, 1); mbar_init(v_full + s, 1); mbar_init(k_empty + s, users); mbar_init(v_empty + s, users); }
    mbar_init(s_full, 1); mbar_init(s_empty, 4);
    mbar_init(p_full, 4); mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  {
    // B operand of the row-sum MMA: 16 x 64 ones (every element equal, so the swizzle pattern is irrelevant)
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + kOffOnes);
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += kThreads) ones[i] = F16 ? 0x3C003C00u : 0x3F803F80u;
    fence_proxy_async();                           // generic-proxy stores -> visible to the tensor core (async proxy)
  }
  tc_fence_before();
  cluster_sync_all();                             // both CTAs' barriers exist before either multicasts into the other
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                                     // qkv (the previous kernel's output) is complete from here on

  if (warp == 0) {
    if (elect_one()) {
      // ---- TMA producer: the K / V ring.  Boxes are half tiles (32 rows): in a sharing pair CTA `rank` fetches half
      // `rank` and multicasts it to both CTAs; a private CTA fetches both halves itself.
      const int colk = a.hidden + head * HD;
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        uint8_t* sk = smem + st * kStage;
        const int row = kv_base + j * BKV;
        mbar_wait(k_empty + st, ph ^ 1);
        mbar_arrive_expect_tx(k_full + st, kK64 + kK16);
        if (shared_kv) {
          tma_load_2d_mc(sk + rank * (kK64 / 2), &tm_k64, k_full + st, colk, row + rank * (BKV / 2), 3);
          tma_load_2d_mc(sk + kK64 + rank * (kK16 / 2), &tm_k16, k_full + st, colk + 64, row + rank * (BKV / 2), 3);
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tma_load_2d(sk + h * (kK64 / 2), &tm_k64, k_full + st, colk, row + h * (BKV / 2));
            tma_load_2d(sk + kK64 + h * (kK16 / 2), &tm_k16, k_full + st, colk + 64, row + h * (BKV / 2));
          }
        }
        // V rows [kv][hd] exactly as they lie in qkv: the P V product reads them as an MN-major B operand
        const int colv = 2 * a.hidden + head * HD;
        uint8_t* sv = sk + kK64 + kK16;
        mbar_wait(v_empty + st, ph ^ 1);
        mbar_arrive_expect_tx(v_full + st, kK64 + kK16);
        if (shared_kv) {
          tma_load_2d_mc(sv + rank * (kK64 / 2), &tm_k64, v_full + st, colv, row + rank * (BKV / 2), 3);
          tma_load_2d_mc(sv + kK64 + rank * (kK16 / 2), &tm_k16, v_full + st, colv + 64, row + rank * (BKV / 2), 3);
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tma_load_2d(sv + h * (kK64 / 2), &tm_k64, v_full + st, colv, row + h * (BKV / 2));
            tma_load_2d(sv + kK64 + h * (kK16 / 2), &tm_k16, v_full + st, colv + 64, row + h * (BKV / 2));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---- MMA issuer
      const uint32_t idesc_qk = umma_idesc_16bit(BQ, BKV, F16);
      // P V: B = V is MN-major (head dim contiguous): transpose-B bit 16 of the instruction descriptor; canonical layout
      // ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units = rows of one kv, 8-row groups SBO apart (CUTLASS make_umma_desc)
      const uint32_t idesc_pv64 = umma_idesc_16bit(BQ, 64, F16) | (1u << 16);
      const uint32_t idesc_pv16 = umma_idesc_16bit(BQ, 16, F16) | (1u << 16);
      // row sums l = P 1: the same P the P V product consumes, summed exactly in fp32 by the tensor core (K-major B of ones)
      const uint32_t idesc_l = umma_idesc_16bit(BQ, 16, F16);
      const uint64_t desc_ones = umma_desc(smem_u32(smem + kOffOnes), 1024, kSw128);
      auto issue_qk = [&](int t) {
        const int st = t % STAGES;
        const uint32_t sk = smem_u32(smem + st * kStage);
        mbar_wait(k_full + st, (t / STAGES) & 1);
        mbar_wait(s_empty, (t & 1) ^ 1);                     // the softmax holds S_{t-1} in registers
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ts(tmem, tmem + kQCol + 8 * ks, umma_desc(sk, 1024, kSw128) + 2 * ks, idesc_qk, ks != 0);
        umma_ts(tmem, tmem + kQCol + 32, umma_desc(sk + kK64, 256, kSw32), idesc_qk, 1);
        if (shared_kv) umma_commit_mc(k_empty + st, 3); else umma_commit(k_empty + st);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_qk(j + 1);
        const int st = j % STAGES;
        const uint32_t sv = smem_u32(smem + st * kStage + kK64 + kK16);
        mbar_wait(v_full + st, (j / STAGES) & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          // 16 kv rows per step: 2048 B of the [64][64] block, 512 B of the [64][16] block
          umma_ts(tmem + kOCol, tmem + kPCol + 8 * ks, umma_desc(sv, 1024, kSw128) + 128 * ks, idesc_pv64, (j | ks) != 0);
          umma_ts(tmem + kOCol + 64, tmem + kPCol + 8 * ks, umma_desc(sv + kK64, 256, kSw32) + 32 * ks, idesc_pv16, (j | ks) != 0);
          umma_ts(tmem + kLCol, tmem + kPCol + 8 * ks, desc_ones + 2 * ks, idesc_l, (j | ks) != 0);
        }
        if (shared_kv) umma_commit_mc(v_empty + st, 3); else umma_commit(v_empty + st);
        umma_commit(pv_done);
      }
    }
  } else {
    // ---- softmax warps: thread = q row.  O accumulates in TMEM across KV tiles; it is rescaled (tcgen05.ld ->
    // multiply -> tcgen05.st) only when the running row max grew by more than 2^8 since the scale in use was chosen
    // ("lazy rescale": P may then exceed 1 by at most 2^8, harmless in fp32 sums and 16-bit P), so the common tile
    // costs one TMEM row load, 64 exp2 and one TMEM row store per thread.
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = a.scale_log2;
    float m_used = -INFINITY;               // scale in use (raw score units); the row sum lives in TMEM next to O

    // Q row -> TMEM (A operand of every Q K^T): 80 16-bit values = 40 columns.  Rows past the end of the tensor read as 0.
    {
      uint32_t qv[40];
      const int64_t grow_ = (int64_t)q0 + row;
      if (grow_ < a.S) {
        const uint4* qp = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(a.qkv) + grow_ * 3 * a.hidden + head * HD);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const uint4 v = __ldg(qp + i);
          qv[4 * i] = v.x; qv[4 * i + 1] = v.y; qv[4 * i + 2] = v.z; qv[4 * i + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 40; ++i) qv[i] = 0u;
      }
      tmem_st_x16(tmem + lane_addr + kQCol, *reinterpret_cast<const uint32_t(*)[16]>(qv));
      tmem_st_x16(tmem + lane_addr + kQCol + 16, *reinterpret_cast<const uint32_t(*)[16]>(qv + 16));
      tmem_st_x8(tmem + lane_addr + kQCol + 32, qv + 32);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }

    // rescale the O row by `factor` - warp-collective TMEM round trip
    auto rescale_o = [&](const float factor) {
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < HD; c += 16) {
        uint32_t t[16];
        tmem_ld_x16(tmem + lane_addr + kOCol + c, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * factor);
        tmem_st_x16(tmem + lane_addr + kOCol + c, t);
      }
      {                                      // ... and the row sum (16 identical columns; one chunk)
        uint32_t t[16];
        tmem_ld_x16(tmem + lane_addr + kLCol, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * factor);
        tmem_st_x16(tmem + lane_addr + kLCol, t);
      }
      tmem_st_wait();
      tc_fence_before();
    };

    // One KV tile.  MASK = the tile may hold columns outside [seg_b, seg_e) (only the first and the last tile of a
    // segment can): the column mask is compiled into that instantiation alone - the compiler if-converts it into
    // compares and selects on every column, which the interior tiles must not carry.
    auto softmax_tile = [&](const int j, auto mask_tag) {
      constexpr bool MASK = decltype(mask_tag)::value;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_x32(tmem + lane_addr, r0);
      tmem_ld_x32(tmem + lane_addr + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      float s[BKV];
#pragma unroll
      for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(r0[i]); s[32 + i] = __uint_as_float(r1[i]); }
      if constexpr (MASK) {
        const int lo = seg_b - (kv_base + j * BKV), hi = seg_e - (kv_base + j * BKV);   // valid columns: [lo, hi)
#pragma unroll
        for (int i = 0; i < BKV; ++i) if (i < lo || i >= hi) s[i] = -INFINITY;
      }
      // row max with three-input FMNMX3: four independent chains of 2 values per instruction (32 instead of 64)
      float mx4[4] = {fmaxf(s[0], s[1]), fmaxf(s[2], s[3]), fmaxf(s[4], s[5]), fmaxf(s[6], s[7])};
#pragma unroll
      for (int i = 8; i < BKV; i += 8) {
        mx4[0] = fmax3(mx4[0], s[i], s[i + 1]); mx4[1] = fmax3(mx4[1], s[i + 2], s[i + 3]);
        mx4[2] = fmax3(mx4[2], s[i + 4], s[i + 5]); mx4[3] = fmax3(mx4[3], s[i + 6], s[i + 7]);
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // the first tile of a segment always holds a valid column, so m_used is finite from then on; a later tile that
      // is masked entirely (mx = -inf) keeps the scale
      const bool grow = (mx - m_used) * sl2 > 8.0f;          // true on the first tile (m_used = -inf)
      float factor = 1.0f;
      if (grow) { factor = ex2_approx((m_used - mx) * sl2); m_used = mx; }
      // the exponentials only need registers: they run while the tensor core is still busy with P_{j-1} V_{j-1}
      const float ms = m_used * sl2;
      uint32_t pk[BKV / 2];
#pragma unroll
      for (int i = 0; i < BKV / 2; ++i) {
        const float x0 = fmaf(s[2 * i], sl2, -ms), x1 = fmaf(s[2 * i + 1], sl2, -ms);
        float p0, p1;
        // interior tiles only: a masked column is -inf, which the SFU maps to 0 and the polynomial cannot
        if (!MASK && kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) {
          p0 = exp2_poly(fmaxf(x0, -120.f)); p1 = exp2_poly(fmaxf(x1, -120.f));
        } else {
          p0 = ex2_approx(x0); p1 = ex2_approx(x1);
        }
        pk[i] = pack2<F16>(p0, p1);
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);                     // the tensor core is done with P_{j-1}; O_{j-1} accumulated
        if (__any_sync(0xffffffffu, grow)) rescale_o(factor);
      }
      tc_fence_after();
      tmem_st_x16(tmem + lane_addr + kPCol, *reinterpret_cast<const uint32_t(*)[16]>(pk));        // P row -> TMEM
      tmem_st_x16(tmem + lane_addr + kPCol + 16, *reinterpret_cast<const uint32_t(*)[16]>(pk + 16));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    };
    for (int j = 0; j < n_kv; ++j) {
      if (j == 0 || j == n_kv - 1) softmax_tile(j, std::true_type{});
      else softmax_tile(j, std::false_type{});
    }
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    float inv;
    {
      uint32_t t[16];
      tmem_ld_x16(tmem + lane_addr + kLCol, t);
      tmem_ld_wait();
      inv = 1.f / __uint_as_float(t[0]);
    }
    uint16_t* dst = static_cast<uint16_t*>(a.out) + (int64_t)(q0 + row) * a.hidden + head * HD;
#pragma unroll
    for (int c = 0; c < 80; c += 16) {
      uint32_t t[16];
      tmem_ld_x16(tmem + lane_addr + kOCol + c, t);
      tmem_ld_wait();
      if (row < q_len) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
          *reinterpret_cast<uint4*>(dst + c + 8 * h) =
              make_uint4(pack2<F16>(__uint_as_float(t[8 * h]) * inv, __uint_as_float(t[8 * h + 1]) * inv),
                         pack2<F16>(__uint_as_float(t[8 * h + 2]) * inv, __uint_as_float(t[8 * h + 3]) * inv),
                         pack2<F16>(__uint_as_float(t[8 * h + 4]) * inv, __uint_as_float(t[8 * h + 5]) * inv),
                         pack2<F16>(__uint_as_float(t[8 * h + 6]) * inv, __uint_as_float(t[8 * h + 7]) * inv));
      }
    }
    tc_fence_before();
  }
  pdl_trigger();
  tc_fence_before();
  cluster_sync_all();                             // the partner's last multicast arrivals have landed before this CTA exits
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, kTmemCols); }
}

}  // namespace

int attention_tc(const void* qkv, void* out, int64_t S, int heads, int head_dim,
                 const int32_t* tiles_dev, int n_tiles, void* stream_, bool f16) {
  if (head_dim != HD) return fail(ZV_EINVAL, "attention_tc: only head_dim=80 is built (got %d)", head_dim);
  if (n_tiles <= 0) return ZV_OK;
  const int hidden = heads * head_dim;
  CUtensorMap t64, t16;
  int rc = make_tmap_2d(&t64, qkv, S, 3 * hidden, 3 * hidden, 64, BKV / 2, 128, f16);     // half K / V tiles (32 rows)
  if (rc) return rc;
  rc = make_tmap_2d(&t16, qkv, S, 3 * hidden, 3 * hidden, 16, BKV / 2, 32, f16);
  if (rc) return rc;
  static std::atomic<uint64_t> attr_set{0};
  const int dev = current_device();
  if (device_needs_setup(attr_set, dev)) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return fail(ZV_ECUDA, "attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    mark_device(attr_set, dev);
  }
  AttnArgs a{};
  a.out = out; a.qkv = qkv; a.S = S; a.tiles = reinterpret_cast<const int4*>(tiles_dev); a.n_tiles = n_tiles; a.heads = heads; a.hidden = hidden; a.f16 = f16;
  a.scale_log2 = (float)(1.4426950408889634 / std::sqrt((double)head_dim));
  const dim3 grid((unsigned)((n_tiles + 1) & ~1), (unsigned)heads);      // clusters of two neighbouring q tiles
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  cudaError_t le;
  {
    NvtxRange nvtx("zv:K4 full attention (tcgen05)");
    KernelTimer timer(KC_ATTN_FULL, stream_);
    le = f16 ? launch_pdl(attn_tc_kernel<true>, grid, dim3(kThreads), kSmem, stream, 2, t64, t16, a)
             : launch_pdl(attn_tc_kernel<false>, grid, dim3(kThreads), kSmem, stream, 2, t64, t16, a);
  }
  if (le != cudaSuccess) return fail(ZV_ECUDA, "attention_tc: launch: %s", cudaGetErrorString(le));
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ZV_ECUDA, "attention_tc: launch: %s", cudaGetErrorString(e));
  return ZV_OK;
}

}  // namespace zv
