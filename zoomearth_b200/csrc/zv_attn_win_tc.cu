// K3: tcgen05 window attention for the 28 windowed layers (HF modeling_qwen2_5_vl.py:207-287 with cu_window_seqlens:
// segments = windows of <= 64 patches, :498-502).
//
// The layer is bandwidth-bound (10.2 KB per patch: q, k, v in, attention out) but the mma.sync kernel it replaces was
// bound by instruction issue at the power-capped clock of the bench (1 600 warp instructions per (window, head): ldmatrix,
// mma.sync, shuffles).  Here the tensor core does both products from TMA-staged shared memory and a thread only runs the
// softmax of one row:
//   item       (row block, head): a row block = up to 128 CONSECUTIVE rows holding whole windows (the plan packs windows
//              greedily; in window order the windows of an image are contiguous), so one TMA box per operand fetches it
//   loads      warp 0: Q, K, V tiles [128 rows][80] as a 128B-swizzled [128][64] block + a 32B-swizzled [128][16] block
//              each (TMA out-of-bounds rows read as zero), 3-stage ring, one mbarrier per stage
//   S = Q K^T  warp 1: tcgen05.mma M=128 N=128, A and B from shared memory, 4 + 1 K steps, fp32 S in TMEM.  S covers the
//              whole block: entries that pair rows of different windows are masked by the softmax (block-diagonal)
//   softmax    warps 2-5 (even items) and 6-9 (odd items), thread = row: the row's window is columns [cb, ce) of the block (per-row bounds table of the
//              plan); two passes over the 32-column chunks that intersect it (max, then exp2 / sum); P is rounded to 16 bit
//              and written over the consumed S columns (S and P share a TMEM buffer), zeros outside the window
//   O = P V    warp 1: A = P from TMEM, B = V as it lies in qkv (MN-major), 8 K steps of one N=64 + one N=16 MMA
//   epilogue   the same warps: O row * 1/l -> 16 bit -> shared-memory staging -> one TMA store per warp ([32 rows][80])
// Measured (B200): 0.73 ms per layer at the bench shape alone (4.4 TB/s = 0.68 of the copy peak), 1.02 ms inside the
// power-capped bench step (the mma.sync kernel: 0.72 / 1.15-1.25 ms).  Not bound by the TMA row-request rate (dropping the
// 32-byte-row boxes changes nothing), nor by bytes in flight (an L2 prefetch of the next blocks made it 1.8x slower), and a
// cp.async producer variant writing the swizzled layouts by hand was slower (1.07 ms): see profiles/README.md.
// S/P and O are double-buffered in TMEM (2 x 128 + 2 x 80 columns), so the products of item i+1 run under the softmax of
// item i and the epilogue of item i runs after the softmax of item i+1.  Persistent: one CTA per SM walks the items.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>

#include "zv_common.h"
#include "zv_gemm.h"
#include "zv_ptx.cuh"

namespace zv {
using namespace ptx;
namespace {

constexpr int HD = 80, BR = 128, STAGES = 3, kThreads = 320;      // warps: 0 loads, 1 MMA issue, 2-5 and 6-9 softmax + epilogue (two sets)
constexpr int kB64 = BR * 64 * 2, kB16 = BR * 16 * 2;                      // 16384, 4096
constexpr int kTile = kB64 + kB16;                                        // 20480: one of Q, K, V
constexpr int kMetaRows = BR + 2;                                         // per-row window bounds of a block (+ alignment slack)
constexpr int kMetaBytes = 1152;                                          // [tile record 16 B | pad | bounds at +32: kMetaRows x 8 B]
constexpr int kBoundsBytes = kMetaRows * 8;                               // 1040: a multiple of 16 (bulk-copy granularity)
constexpr int kStage = 3 * kTile;                                         // 61440
constexpr int kOffMeta = STAGES * kStage;                                 // [STAGES][kMetaBytes]
constexpr int kOffOut = kOffMeta + STAGES * kMetaBytes;                   // [8 warps][32 rows][80] 16-bit: output staging for the TMA stores
constexpr int kOutBytes = 8 * 32 * HD * 2;                                // 40960
constexpr int kOffBar = kOffOut + kOutBytes;
constexpr int kSmem = kOffBar + 256 + 1024;
constexpr int kTmemCols = 512, kOCol = 256;                               // S/P buffer b at [128 b, 128 b + 128), O buffer b at [256 + 80 b, ..)
constexpr uint32_t kSw128 = 2, kSw32 = 6;                                 // UMMA descriptor layout types
static_assert(kTile % 1024 == 0 && kB64 % 1024 == 0 && kStage % 1024 == 0, "swizzle atom alignment");
static_assert(kBoundsBytes % 16 == 0 && 32 + kBoundsBytes <= kMetaBytes && kMetaBytes % 128 == 0, "meta block");
static_assert(kSmem <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (F16) { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 1-D bulk copy global -> shared, completing on an mbarrier (src, dst 16-byte aligned; bytes a multiple of 16)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// TMA store of a [box rows][box cols] tile from shared memory (dense, no swizzle) to a 2-D tensor; bulk async-group completion
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

#ifdef ZV_WIN_TRACE      // debug build only: per-item clock stamps of CTA 0 (tools/win_trace.py)
__device__ long long g_win_trace[16 * 64];
#define ZV_TRACE(slot, k) do { if (blockIdx.x == 0 && (k) < 64) g_win_trace[(slot) * 64 + (k)] = clock64(); } while (0)
#else
#define ZV_TRACE(slot, k) do { } while (0)
#endif

struct WinArgs {
  void* out;
  const int4* tiles;          // (row0, n_rows, -, -): row blocks of whole windows
  const int2* bounds;         // [S + kMetaRows]: (first row, end row) of the window each patch row belongs to (zero padded)
  int n_tiles, heads, hidden;
  float scale_log2;
};

template <bool F16>
__global__ void __launch_bounds__(kThreads, 1) attn_win_tc_kernel(const __grid_constant__ CUtensorMap tm64,
                                                                  const __grid_constant__ CUtensorMap tm16,
                                                                  const __grid_constant__ CUtensorMap tm_out, const WinArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* load_full = bars;            // STAGES
  uint64_t* load_empty = bars + 3;       // STAGES
  uint64_t* s_full = bars + 6;           // 2
  uint64_t* p_full = bars + 8;           // 2
  uint64_t* o_full = bars + 10;          // 2
  uint64_t* o_empty = bars + 12;         // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm64); prefetch_tensormap(&tm16); prefetch_tensormap(&tm_out);
    for (int s = 0; s < STAGES; ++s) { mbar_init(load_full + s, 1); mbar_init(load_empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(s_full + b, 1); mbar_init(p_full + b, 4); mbar_init(o_full + b, 1); mbar_init(o_empty + b, 4); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                                     // qkv (the QKV GEMM's output) is complete from here on

  const int n_items = a.n_tiles * a.heads;
  const int first = blockIdx.x, step = gridDim.x;
  const int n_mine = first < n_items ? (n_items - first + step - 1) / step : 0;

  if (warp == 0) {
    if (elect_one()) {
      // ---- TMA producer
      for (int k = 0; k < n_mine; ++k) {
        const int item = first + k * step;
        const int st = k % STAGES;
        const uint32_t ph = (k / STAGES) & 1;
        const int4 tl = __ldg(&a.tiles[item / a.heads]);
        const int row0 = tl.x;
        const int head = item % a.heads;
        uint8_t* s = smem + st * kStage;
        mbar_wait(load_empty + st, ph ^ 1);
        // the block's record and per-row window bounds travel with the stage: a dependent L2 round trip per item in the
        // softmax warps would otherwise sit on the critical path (two of them: record, then bounds)
        uint8_t* meta = smem + kOffMeta + st * kMetaBytes;
        *reinterpret_cast<int4*>(meta) = tl;                                  // released by the arrive below
#ifdef ZV_WIN_SKIP16                     // timing experiment only (wrong results): no 32-byte-row boxes
        mbar_arrive_expect_tx(load_full + st, 3 * kB64 + kBoundsBytes);
#else
        mbar_arrive_expect_tx(load_full + st, 3 * kTile + kBoundsBytes);
#endif
        bulk_load_1d(meta + 32, a.bounds + (row0 & ~1), kBoundsBytes, load_full + st);
#pragma unroll
        for (int t = 0; t < 3; ++t) {               // q, k, v column groups of qkv
          const int col = t * a.hidden + head * HD;
          if (t == 0) ZV_TRACE(0, k);                       // stage free, loads go out
          tma_load_2d(s + t * kTile, &tm64, load_full + st, col, row0);
#ifndef ZV_WIN_SKIP16
          tma_load_2d(s + t * kTile + kB64, &tm16, load_full + st, col + 64, row0);
#endif
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---- MMA issuer
      const uint32_t idesc_qk = umma_idesc_16bit(BR, BR, F16);
      const uint32_t idesc_pv64 = umma_idesc_16bit(BR, 64, F16) | (1u << 16);    // B = V is MN-major (transpose-B bit)
      const uint32_t idesc_pv16 = umma_idesc_16bit(BR, 16, F16) | (1u << 16);
      auto issue_qk = [&](int k) {
        const int st = k % STAGES, b = k & 1;
        // S/P buffer b was last read by P V of item k - 2 (its A operand): that product must have retired
        if (k >= 2) mbar_wait(o_full + b, ((k - 2) >> 1) & 1);
        mbar_wait(load_full + st, (k / STAGES) & 1);
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + st * kStage), sk = sq + kTile;
        const uint32_t d = tmem + b * BR;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(d, umma_desc(sq, 1024, kSw128) + 2 * ks, umma_desc(sk, 1024, kSw128) + 2 * ks, idesc_qk, ks != 0);
        umma_bf16(d, umma_desc(sq + kB64, 256, kSw32), umma_desc(sk + kB64, 256, kSw32), idesc_qk, 1);
        umma_commit(s_full + b);
      };
      if (n_mine > 0) issue_qk(0);
      for (int k = 0; k < n_mine; ++k) {
        if (k + 1 < n_mine) issue_qk(k + 1);
        const int st = k % STAGES, b = k & 1;
        const uint32_t sv = smem_u32(smem + st * kStage + 2 * kTile);
        mbar_wait(p_full + b, (k >> 1) & 1);
        mbar_wait(o_empty + b, ((k >> 1) & 1) ^ 1);          // the epilogue of item k - 2 has the old O in registers
        tc_fence_after();
        const uint32_t o = tmem + kOCol + b * HD, pa = tmem + b * BR;
#pragma unroll
        for (int ks = 0; ks < BR / 16; ++ks) {
          // 16 kv rows per step: 2048 B of the [128][64] block, 512 B of the [128][16] block
          umma_ts(o, pa + 8 * ks, umma_desc(sv, 1024, kSw128) + 128 * ks, idesc_pv64, ks != 0);
          umma_ts(o + 64, pa + 8 * ks, umma_desc(sv + kB64, 256, kSw32) + 32 * ks, idesc_pv16, ks != 0);
        }
        umma_commit(load_empty + st);
        umma_commit(o_full + b);
      }
    }
  } else if (warp >= 2) {
    // ---- softmax + epilogue warps: two sets of four (warps 2-5: even items, S/P/O buffer 0; warps 6-9: odd items, buffer 1;
    // a warp reads the TMEM lane quarter warp % 4), thread = row of the block.  A tcgen05.ld round trip is ~400 clk plus
    // the bytes over a 64 B/clk port (72 KB of S and O per item: measured 1 200 clk), then 600 clk of exponentials, the P
    // store and the output staging - 4 000 clk per item when one set did everything in a row (clock stamps, tools/win_trace.py).
    // With two sets the TMEM transfers of one item run under the arithmetic of the other.
    const int set = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = a.scale_log2;
    uint8_t* stage_out = smem + kOffOut + (set * 4 + quarter) * (32 * HD * 2);
    const bool tracer = warp == 2 && lane == 0;
    for (int k = set; k < n_mine; k += 2) {
      const int b = set;                                          // = k & 1
      const uint32_t ph = (uint32_t)(k >> 1) & 1u;
      const uint32_t sbuf = tmem + lane_addr + b * BR;
      // ---- softmax of item k
      if (tracer) ZV_TRACE(6, k);
      mbar_wait(s_full + b, ph);                                  // S is complete, so the stage (and its meta block) has landed
      if (tracer) ZV_TRACE(7, k);
      tc_fence_after();
      const uint8_t* meta = smem + kOffMeta + (k % STAGES) * kMetaBytes;
      const int4 tl = *reinterpret_cast<const int4*>(meta);
      const bool valid = row < tl.y;
      int cb = 0, ce = 0;
      if (valid) { const int2 w = reinterpret_cast<const int2*>(meta + 32)[(tl.x & 1) + row]; cb = w.x - tl.x; ce = w.y - tl.x; }
      // 32-column chunks of S this warp needs: [c_lo, c_hi]; at most two for windows aligned to 32 rows (the usual case)
      const int c_lo = __reduce_min_sync(0xffffffffu, valid ? cb >> 5 : 3);
      const int c_hi = __reduce_max_sync(0xffffffffu, valid ? (ce - 1) >> 5 : 0);
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
      if (c_hi <= c_lo + 1) {
        uint32_t sr0[32], sr1[32];
        if (c_lo <= c_hi) {
          tmem_ld_x32(sbuf + 32 * c_lo, sr0);
          if (c_lo < 3) tmem_ld_x32(sbuf + 32 * (c_lo + 1), sr1);
          tmem_ld_wait();
        }
        if (tracer) ZV_TRACE(8, k);
        const int col0 = 32 * c_lo;
        const bool full = __all_sync(0xffffffffu, cb <= col0 && ce >= col0 + 64);
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            m4[(i >> 1) & 3] = fmax3(m4[(i >> 1) & 3], __uint_as_float(sr0[i]), __uint_as_float(sr0[i + 1]));
            m4[(i >> 1) & 3] = fmax3(m4[(i >> 1) & 3], __uint_as_float(sr1[i]), __uint_as_float(sr1[i + 1]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c0 = col0 + i, c1 = col0 + 32 + i;
            m4[i & 3] = fmaxf(m4[i & 3], (c0 >= cb && c0 < ce) ? __uint_as_float(sr0[i]) : -INFINITY);
            m4[i & 3] = fmaxf(m4[i & 3], (c1 >= cb && c1 < ce) ? __uint_as_float(sr1[i]) : -INFINITY);
          }
        }
        const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float ms = valid ? m * sl2 : 0.f;
        if (tracer) ZV_TRACE(13, k);
        uint32_t pk0[16], pk1[16];
        if (full) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(sr0[2 * i]), sl2, -ms)), p1 = ex2_approx(fmaf(__uint_as_float(sr0[2 * i + 1]), sl2, -ms));
            const float p2 = ex2_approx(fmaf(__uint_as_float(sr1[2 * i]), sl2, -ms)), p3 = ex2_approx(fmaf(__uint_as_float(sr1[2 * i + 1]), sl2, -ms));
            l4[i & 3] += (p0 + p1) + (p2 + p3);
            pk0[i] = pack2<F16>(p0, p1);
            pk1[i] = pack2<F16>(p2, p3);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c0 = col0 + 2 * i, c1 = col0 + 32 + 2 * i;
            const float p0 = (c0 >= cb && c0 < ce) ? ex2_approx(fmaf(__uint_as_float(sr0[2 * i]), sl2, -ms)) : 0.f;
            const float p1 = (c0 + 1 >= cb && c0 + 1 < ce) ? ex2_approx(fmaf(__uint_as_float(sr0[2 * i + 1]), sl2, -ms)) : 0.f;
            const float p2 = (c1 >= cb && c1 < ce) ? ex2_approx(fmaf(__uint_as_float(sr1[2 * i]), sl2, -ms)) : 0.f;
            const float p3 = (c1 + 1 >= cb && c1 + 1 < ce) ? ex2_approx(fmaf(__uint_as_float(sr1[2 * i + 1]), sl2, -ms)) : 0.f;
            l4[i & 3] += (p0 + p1) + (p2 + p3);
            pk0[i] = pack2<F16>(p0, p1);
            pk1[i] = pack2<F16>(p2, p3);
          }
        }
        if (tracer) ZV_TRACE(14, k);
        // P over the S columns (all of this warp's S is in registers by now); zeros for the chunks of other windows
        uint32_t zero[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) zero[i] = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c == c_lo) tmem_st_x16(sbuf + 16 * c, pk0);
          else if (c == c_lo + 1) tmem_st_x16(sbuf + 16 * c, pk1);
          else tmem_st_x16(sbuf + 16 * c, zero);
        }
      } else {
        // general case (a warp's rows span three or four chunks: windows not aligned to 32 rows): two passes over TMEM
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool need = cb < 32 * (c + 1) && ce > 32 * c;
          if (!__any_sync(0xffffffffu, need)) continue;
          uint32_t r[32];
          tmem_ld_x32(sbuf + 32 * c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = 32 * c + i;
            m = fmaxf(m, (col >= cb && col < ce) ? __uint_as_float(r[i]) : -INFINITY);
          }
        }
        const float ms = valid ? m * sl2 : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool need = cb < 32 * (c + 1) && ce > 32 * c;
          uint32_t pk[16];
          if (__any_sync(0xffffffffu, need)) {
            uint32_t r[32];
            tmem_ld_x32(sbuf + 32 * c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = 32 * c + 2 * i;
              const float p0 = (col >= cb && col < ce) ? ex2_approx(fmaf(__uint_as_float(r[2 * i]), sl2, -ms)) : 0.f;
              const float p1 = (col + 1 >= cb && col + 1 < ce) ? ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), sl2, -ms)) : 0.f;
              l4[i & 3] += p0 + p1;
              pk[i] = pack2<F16>(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          }
          tmem_st_x16(sbuf + 16 * c, pk);       // P chunk c only overlaps S chunks <= c / 2: already consumed
        }
      }
      if (tracer) ZV_TRACE(15, k);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + b);
      if (tracer) ZV_TRACE(9, k);
      const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);

      // ---- epilogue of item k: O row * 1/l -> 16 bit -> global
      const int head = (first + k * step) % a.heads;
      if (tracer) ZV_TRACE(10, k);
      mbar_wait(o_full + b, ph);
      if (tracer) ZV_TRACE(11, k);
      tc_fence_after();
      uint32_t o[HD];
#pragma unroll
      for (int c = 0; c < HD; c += 16) tmem_ld_x16(tmem + lane_addr + kOCol + b * HD + c, *reinterpret_cast<uint32_t(*)[16]>(o + c));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty + b);
      if (tracer) ZV_TRACE(1, k);                                 // (trace slots 1-5 are reused by the tracer warp here)
      // A thread's row is 160 contiguous bytes, but 32 lanes x 16 bytes of 32 different rows per store instruction cost the
      // LSU ~300 clk each under load (measured: 3 100 clk per item for the ten of them).  A warp whose 32 rows are all inside
      // the block stages them in shared memory and hands them to the TMA unit as one [32][80] box; warps at the ragged end
      // of a block keep the per-thread stores.
      const float inv = valid ? 1.f / l : 0.f;
      uint4 ov[HD / 8];
#pragma unroll
      for (int j = 0; j < HD / 8; ++j)
        ov[j] = make_uint4(pack2<F16>(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                           pack2<F16>(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                           pack2<F16>(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                           pack2<F16>(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv));
      if (__all_sync(0xffffffffu, valid)) {
        uint4* stg = reinterpret_cast<uint4*>(stage_out + lane * (HD * 2));
        if (tracer) ZV_TRACE(2, k);
        if (lane == 0) bulk_store_wait_read();                   // the previous box has left this staging area
        __syncwarp();
        if (tracer) ZV_TRACE(3, k);
#pragma unroll
        for (int j = 0; j < HD / 8; ++j) stg[j] = ov[j];
        fence_proxy_async_smem();                                // generic-proxy writes -> visible to the TMA unit
        __syncwarp();
        if (tracer) ZV_TRACE(4, k);
        if (lane == 0) tma_store_2d(&tm_out, stage_out, head * HD, tl.x + quarter * 32);
      } else if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(a.out) + (int64_t)(tl.x + row) * a.hidden + head * HD);
#pragma unroll
        for (int j = 0; j < HD / 8; ++j) dst[j] = ov[j];
      }
      if (tracer) ZV_TRACE(12, k);
    }
    if (lane == 0) bulk_store_wait_all();             // this warp's TMA stores have been written before the CTA exits
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, kTmemCols); }
}

}  // namespace

// tiles_dev: (row0, n_rows, -, -) row blocks of whole windows (<= 128 rows); bounds_dev: int32 [S][2] per-row window bounds.
int attention_win_tc(const void* qkv, void* out, int64_t S, int heads, int head_dim, const int32_t* tiles_dev, int n_tiles,
                     const int32_t* bounds_dev, void* stream_, bool f16) {
  if (head_dim != HD) return fail(ZV_EINVAL, "attention_win_tc: only head_dim=80 is built (got %d)", head_dim);
  if (n_tiles <= 0) return ZV_OK;
  const int hidden = heads * head_dim;
  CUtensorMap t64, t16;
  int rc = make_tmap_2d(&t64, qkv, S, 3 * hidden, 3 * hidden, 64, BR, 128, f16);
  if (rc) return rc;
  rc = make_tmap_2d(&t16, qkv, S, 3 * hidden, 3 * hidden, 16, BR, 32, f16);
  if (rc) return rc;
  CUtensorMap tout;                                   // output (S, hidden): [32 rows][80] boxes, dense in shared memory
  rc = make_tmap_2d(&tout, out, S, hidden, hidden, HD, 32, 0, f16);
  if (rc) return rc;
  static std::atomic<uint64_t> attr_set{0};
  const int dev = current_device();
  if (device_needs_setup(attr_set, dev)) {
    cudaError_t e = cudaFuncSetAttribute(attn_win_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_win_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return fail(ZV_ECUDA, "attention_win_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    mark_device(attr_set, dev);
  }
  WinArgs a{};
  a.out = out; a.tiles = reinterpret_cast<const int4*>(tiles_dev); a.bounds = reinterpret_cast<const int2*>(bounds_dev);
  a.n_tiles = n_tiles; a.heads = heads; a.hidden = hidden;
  a.scale_log2 = (float)(1.4426950408889634 / std::sqrt((double)head_dim));
  const int64_t n_items = (int64_t)n_tiles * heads;
  const int grid = (int)std::min<int64_t>(n_items, num_sms());
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  cudaError_t le;
  {
    NvtxRange nvtx("zv:K3 window attention (tcgen05)");
    KernelTimer timer(KC_ATTN_WINDOW, stream_);
    le = f16 ? launch_pdl(attn_win_tc_kernel<true>, dim3((unsigned)grid), dim3(kThreads), kSmem, stream, 1, t64, t16, tout, a)
             : launch_pdl(attn_win_tc_kernel<false>, dim3((unsigned)grid), dim3(kThreads), kSmem, stream, 1, t64, t16, tout, a);
  }
  if (le != cudaSuccess) return fail(ZV_ECUDA, "attention_win_tc: launch: %s", cudaGetErrorString(le));
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ZV_ECUDA, "attention_win_tc: launch: %s", cudaGetErrorString(e));
  return ZV_OK;
}

}  // namespace zv

#ifdef ZV_WIN_TRACE
extern "C" __attribute__((visibility("default"))) int zv_debug_win_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, zv::g_win_trace, sizeof(long long) * 16 * 64);
}
#endif
