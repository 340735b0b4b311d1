// K3/K4: varlen non-causal attention for the vision tower (HF modeling_qwen2_5_vl.py:207-287).
// One kernel serves both the 28 windowed layers (segments = windows of <= 64 patches, HF :498-502 with
// cu_window_seqlens) and the 4 full-attention layers (segments = whole images, cu_seqlens): the host plan
// turns either segment list into q tiles (q0, q_len, seg_begin, seg_end); a CTA owns one (q tile, head).
// Flash-style: K/V stream through a cp.async double buffer in 64-row tiles, S = QK^T and O += PV run on
// mma.sync m16n8k16 (bf16 in, fp32 accumulate), softmax is online in fp32 with exp2.  Rotary embedding is
// already applied to q,k by the QKV GEMM epilogue (zv_gemm.cu EPI_QKV_ROPE).
// Layout: qkv (S, 3, heads, 80) bf16, out (S, heads*80) bf16.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>

#include "zv_common.h"
#include "zv_gemm.h"
#include "zv_ptx.cuh"

namespace zv {
namespace {

constexpr int HD = 80;          // head dim
constexpr int BQ = 64, BKV = 64;
constexpr int LDS = 88;         // smem row pitch in elements (176 B: conflict-free ldmatrix)
constexpr int kTileElems = 64 * LDS;
template <int NW> constexpr int smem_bytes() { return (16 * NW + (NW == 4 ? 2 : 4) * 64) * LDS * 2; }   // Q + (K, V) x 1 (windows) or x 2 (long segments)

using ptx::smem_u32;
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;   // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
template <bool F16>
__device__ __forceinline__ void mma_16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_16(float a, float b) {
  if constexpr (F16) { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// `rows` x 80 bf16 from global (row pitch ld elements) into a padded smem tile; rows >= n_valid are zeros.
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int64_t ld, int n_valid, int rows) {
  for (int i = threadIdx.x; i < rows * 10; i += blockDim.x) {
    const int r = i / 10, c = i % 10;
    const bool ok = r < n_valid;
    cp_async16(s + r * LDS + c * 8, g + (int64_t)(ok ? r : 0) * ld + c * 8, ok);
  }
}

// NW warps per CTA, 16 q rows each: NW = 4 for the window layers (segments <= 64), 8 for the full layers, where
// the 128-row q tile halves the K/V traffic from L2 per q row.
template <int NW, bool F16>
__global__ void __launch_bounds__(32 * NW) attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                   const int4* __restrict__ tiles, int heads, float scale_log2) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  constexpr int NBUF = NW == 4 ? 1 : 2;         // window tiles (NW = 4) hold the whole segment in one K/V tile
  __nv_bfloat16* sK = sQ + 16 * NW * LDS;       // [NBUF][64][LDS]
  __nv_bfloat16* sV = sK + NBUF * kTileElems;   // [NBUF][64][LDS]
  const int4 tl = tiles[blockIdx.x];
  const int q0 = tl.x, q_len = tl.y, seg_b = tl.z, seg_e = tl.w;
  const int head = blockIdx.y;
  const int hidden = heads * HD;
  const int64_t ld = 3 * (int64_t)hidden;
  const __nv_bfloat16* gq = qkv + (int64_t)q0 * ld + head * HD;
  const __nv_bfloat16* gk = qkv + (int64_t)seg_b * ld + hidden + head * HD;
  const __nv_bfloat16* gv = gk + hidden;
  const int kv_len = seg_e - seg_b;
  const int n_kv = (kv_len + BKV - 1) / BKV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  load_tile(sQ, gq, ld, q_len, 16 * NW);
  load_tile(sK, gk, ld, min(BKV, kv_len), BKV);
  load_tile(sV, gv, ld, min(BKV, kv_len), BKV);
  cp_async_commit();

  uint32_t qf[5][4];
  float o[10][4];
#pragma unroll
  for (int i = 0; i < 10; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int j = 0; j < n_kv; ++j) {
    const int buf = NBUF == 2 ? (j & 1) : 0;
    if (NBUF == 2 && j + 1 < n_kv) {            // prefetch the next K/V tile into the other buffer
      const int nv = min(BKV, kv_len - (j + 1) * BKV);
      load_tile(sK + (buf ^ 1) * kTileElems, gk + (int64_t)(j + 1) * BKV * ld, ld, nv, BKV);
      load_tile(sV + (buf ^ 1) * kTileElems, gv + (int64_t)(j + 1) * BKV * ld, ld, nv, BKV);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) {
        const int mi = lane >> 3, r = lane & 7;
        ldsm_x4(qf[ks], sQ + (16 * warp + (mi & 1) * 8 + r) * LDS + 16 * ks + (mi >> 1) * 8);
      }
    }
    const __nv_bfloat16* k = sK + buf * kTileElems;
    const __nv_bfloat16* v = sV + buf * kTileElems;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {      // pairs of 8-wide kv tiles
        uint32_t b[4];
        const int mi = lane >> 3, r = lane & 7;
        ldsm_x4(b, k + (16 * np + (mi >> 1) * 8 + r) * LDS + 16 * ks + (mi & 1) * 8);
        mma_16<F16>(s[2 * np], qf[ks], b[0], b[1]);
        mma_16<F16>(s[2 * np + 1], qf[ks], b[2], b[3]);
      }
    }
    // mask columns beyond the segment, online softmax (rows g and g+8 of this warp's 16)
    const int kv_base = j * BKV;
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = kv_base + 8 * i + 2 * t;
      if (c >= kv_len) { s[i][0] = -INFINITY; s[i][2] = -INFINITY; }
      if (c + 1 >= kv_len) { s[i][1] = -INFINITY; s[i][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = ptx::ex2_approx((m0 - mx0) * scale_log2), c1 = ptx::ex2_approx((m1 - mx1) * scale_log2);
    m0 = mx0; m1 = mx1;
    const float ms0 = mx0 * scale_log2, ms1 = mx1 * scale_log2;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = ptx::ex2_approx(s[i][0] * scale_log2 - ms0); s[i][1] = ptx::ex2_approx(s[i][1] * scale_log2 - ms0);
      s[i][2] = ptx::ex2_approx(s[i][2] * scale_log2 - ms1); s[i][3] = ptx::ex2_approx(s[i][3] * scale_log2 - ms1);
      rs0 += s[i][0] + s[i][1];
      rs1 += s[i][2] + s[i][3];
    }
    l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < 10; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_16<F16>(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_16<F16>(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_16<F16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_16<F16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 5; ++np) {
        uint32_t b[4];
        const int mi = lane >> 3, r = lane & 7;
        ldsm_x4_trans(b, v + (16 * kk + (mi & 1) * 8 + r) * LDS + 8 * (2 * np + (mi >> 1)));
        mma_16<F16>(o[2 * np], a, b[0], b[1]);
        mma_16<F16>(o[2 * np + 1], a, b[2], b[3]);
      }
    }
    __syncthreads();   // everyone is done with buffer `buf` before iteration j+1 prefetches into it
  }
  // finalise: row sums across the quad, normalise, stage through this warp's own Q rows, 16-byte stores
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  __nv_bfloat16* so = sQ + 16 * warp * LDS;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    *reinterpret_cast<uint32_t*>(so + g * LDS + 8 * i + 2 * t) = pack_16<F16>(o[i][0] * i0, o[i][1] * i0);
    *reinterpret_cast<uint32_t*>(so + (g + 8) * LDS + 8 * i + 2 * t) = pack_16<F16>(o[i][2] * i1, o[i][3] * i1);
  }
  __syncwarp();
  for (int i = lane; i < 16 * 10; i += 32) {
    const int r = i / 10, c = i % 10;
    const int row = 16 * warp + r;
    if (row < q_len)
      *reinterpret_cast<uint4*>(out + (int64_t)(q0 + row) * hidden + head * HD + c * 8) =
          *reinterpret_cast<const uint4*>(so + r * LDS + c * 8);
  }
}

// Window layers: every segment fits one 64-row tile, so there is no online softmax; the kernel is persistent and
// software-pipelined over (window, head) items: while item i is in the tensor cores, the Q/K/V tiles of item i+1 are
// already streaming into the other shared-memory set with cp.async, so the kernel runs at memory speed instead of
// paying the load latency once per CTA.
template <bool F16>
__global__ void __launch_bounds__(128, 3) attn_window_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                             const int4* __restrict__ tiles, int n_items, int heads,
                                                             float scale_log2) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sbase = reinterpret_cast<__nv_bfloat16*>(smem_raw);       // [2 sets][Q, K, V][64][LDS]
  const int hidden = heads * HD;
  const int64_t ld = 3 * (int64_t)hidden;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;

  // the (q0, q_len, seg_begin, seg_end) record of an item is fetched two iterations ahead: its L2 latency would
  // otherwise sit in front of every prefetch and every tile (12 % of the kernel's stall samples)
  auto issue_loads = [&](const int4 tl, int item, int set) {
    const int head = item % heads;
    const int len = tl.w - tl.z;                                          // windows: q rows == kv rows == the segment
    __nv_bfloat16* s = sbase + set * 3 * kTileElems;
    const __nv_bfloat16* gq = qkv + (int64_t)tl.z * ld + head * HD;
    load_tile(s, gq, ld, len, 64);
    load_tile(s + kTileElems, gq + hidden, ld, len, 64);
    load_tile(s + 2 * kTileElems, gq + 2 * hidden, ld, len, 64);
    cp_async_commit();
  };

  int item = blockIdx.x;
  if (item >= n_items) return;
  ptx::pdl_wait();                                  // qkv (the QKV GEMM's output) is complete from here on
  int4 tl = __ldg(tiles + item / heads);
  int4 tl_next = item + (int)gridDim.x < n_items ? __ldg(tiles + (item + gridDim.x) / heads) : tl;
  issue_loads(tl, item, 0);
  for (int it = 0; item < n_items; item += gridDim.x, ++it) {
    const int set = it & 1;
    const int next = item + gridDim.x, after = next + gridDim.x;
    const int4 tl_after = after < n_items ? __ldg(tiles + after / heads) : tl_next;
    if (next < n_items) { issue_loads(tl_next, next, set ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const int head = item % heads;
    const int len = tl.w - tl.z;
    __nv_bfloat16* sQ = sbase + set * 3 * kTileElems;
    const __nv_bfloat16* k = sQ + kTileElems;
    const __nv_bfloat16* v = sQ + 2 * kTileElems;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
      uint32_t qf[4];
      ldsm_x4(qf, sQ + (16 * warp + (mi & 1) * 8 + r8) * LDS + 16 * ks + (mi >> 1) * 8);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        ldsm_x4(b, k + (16 * np + (mi >> 1) * 8 + r8) * LDS + 16 * ks + (mi & 1) * 8);
        mma_16<F16>(s[2 * np], qf, b[0], b[1]);
        mma_16<F16>(s[2 * np + 1], qf, b[2], b[3]);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = 8 * i + 2 * t;
      if (c >= len) { s[i][0] = -INFINITY; s[i][2] = -INFINITY; }
      if (c + 1 >= len) { s[i][1] = -INFINITY; s[i][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float ms0 = mx0 * scale_log2, ms1 = mx1 * scale_log2;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = ptx::ex2_approx(s[i][0] * scale_log2 - ms0); s[i][1] = ptx::ex2_approx(s[i][1] * scale_log2 - ms0);
      s[i][2] = ptx::ex2_approx(s[i][2] * scale_log2 - ms1); s[i][3] = ptx::ex2_approx(s[i][3] * scale_log2 - ms1);
      l0 += s[i][0] + s[i][1];
      l1 += s[i][2] + s[i][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float o[10][4];
#pragma unroll
    for (int i = 0; i < 10; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_16<F16>(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_16<F16>(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_16<F16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_16<F16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 5; ++np) {
        uint32_t b[4];
        ldsm_x4_trans(b, v + (16 * kk + (mi & 1) * 8 + r8) * LDS + 8 * (2 * np + (mi >> 1)));
        mma_16<F16>(o[2 * np], a, b[0], b[1]);
        mma_16<F16>(o[2 * np + 1], a, b[2], b[3]);
      }
    }
    // this warp's 16 output rows: stage over its own (already consumed) Q rows, then 16-byte stores
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    __nv_bfloat16* so = sQ + 16 * warp * LDS;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      *reinterpret_cast<uint32_t*>(so + g * LDS + 8 * i + 2 * t) = pack_16<F16>(o[i][0] * i0, o[i][1] * i0);
      *reinterpret_cast<uint32_t*>(so + (g + 8) * LDS + 8 * i + 2 * t) = pack_16<F16>(o[i][2] * i1, o[i][3] * i1);
    }
    __syncwarp();
    for (int i = lane; i < 16 * 10; i += 32) {
      const int r = i / 10, c = i % 10;
      const int row = 16 * warp + r;
      if (row < len)
        *reinterpret_cast<uint4*>(out + (int64_t)(tl.z + row) * hidden + head * HD + c * 8) =
            *reinterpret_cast<const uint4*>(so + r * LDS + c * 8);
    }
    __syncthreads();        // everyone is done with this set before the next iteration prefetches into it
    tl = tl_next;
    tl_next = tl_after;
  }
  ptx::pdl_trigger();
}

template <bool F16>
int launch_window(const void* qkv, void* out, int heads, const int32_t* tiles_dev, int n_tiles, float scale_log2, void* stream_) {
  constexpr int kSmem = 2 * 3 * kTileElems * 2;
  static std::atomic<uint64_t> attr_set{0};
  const int dev = current_device();
  if (device_needs_setup(attr_set, dev)) {
    cudaError_t e = cudaFuncSetAttribute(attn_window_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return fail(ZV_ECUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    mark_device(attr_set, dev);
  }
  const int n_sm = num_sms();
  const int n_items = n_tiles * heads;
  const int grid = n_items < 3 * n_sm ? n_items : 3 * n_sm;
  NvtxRange nvtx("zv:K3 window attention");
  KernelTimer timer(KC_ATTN_WINDOW, stream_);
  launch_pdl(attn_window_kernel<F16>, dim3((unsigned)grid), dim3(128), kSmem, static_cast<cudaStream_t>(stream_), 1,
             static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), reinterpret_cast<const int4*>(tiles_dev),
             n_items, heads, scale_log2);
  return ZV_OK;
}

template <int NW, bool F16>
int launch_attn(const void* qkv, void* out, int heads, const int32_t* tiles_dev, int n_tiles, float scale_log2,
                void* stream_, int cls) {
  static std::atomic<uint64_t> attr_set{0};
  const int dev = current_device();
  if (device_needs_setup(attr_set, dev)) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel<NW, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<NW>());
    if (e != cudaSuccess) return fail(ZV_ECUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    mark_device(attr_set, dev);
  }
  dim3 grid((unsigned)n_tiles, (unsigned)heads);
  KernelTimer timer(cls, stream_);
  attn_kernel<NW, F16><<<grid, 32 * NW, smem_bytes<NW>(), static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), reinterpret_cast<const int4*>(tiles_dev),
      heads, scale_log2);
  return ZV_OK;
}

}  // namespace

// tiles_dev: (q0, q_len, seg_begin, seg_end) work items; q_len <= 64 when !full_layer, <= 128 when full_layer.
int attention(const void* qkv, void* out, int heads, int head_dim, const int32_t* tiles_dev, int n_tiles, void* stream_,
              bool full_layer, bool f16) {
  if (head_dim != HD) return fail(ZV_EINVAL, "attention: only head_dim=80 is built (got %d)", head_dim);
  if (n_tiles <= 0) return ZV_OK;
  const float scale_log2 = (float)(1.4426950408889634 / std::sqrt((double)head_dim));
  int rc;
  if (f16)
    rc = full_layer ? launch_attn<8, true>(qkv, out, heads, tiles_dev, n_tiles, scale_log2, stream_, KC_ATTN_FULL)
                    : launch_window<true>(qkv, out, heads, tiles_dev, n_tiles, scale_log2, stream_);
  else
    rc = full_layer ? launch_attn<8, false>(qkv, out, heads, tiles_dev, n_tiles, scale_log2, stream_, KC_ATTN_FULL)
                    : launch_window<false>(qkv, out, heads, tiles_dev, n_tiles, scale_log2, stream_);
  if (rc) return rc;
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ZV_ECUDA, "attention: launch: %s", cudaGetErrorString(e));
  return ZV_OK;
}

}  // namespace zv
