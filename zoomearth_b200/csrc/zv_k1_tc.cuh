// K1 on the 5th-generation tensor cores: one resample pass of Pillow's 8 bpc bicubic resize as a banded int8 GEMM
// (tcgen05.mma kind::i8, u8 pixels x s8 coefficient digits -> s32 in TMEM: exact integer arithmetic), fed by TMA.
// Included by zv_k1.cu inside its anonymous namespace.
//
// One kernel serves both passes because a pass is "resample every input row along its contiguous axis and write the
// result transposed":
//   pass 1   input = the source image (rows y, bytes 3 x + c: three interleaved channels), output = T[3 xx + c][y]
//   pass 2   input = T (rows 3 xx + c, bytes y: one channel),                             output = U[yy][3 xx + c]
// so T is the horizontally resized image stored column by column and U is the finished (oh, ow, 3) uint8 image
// (PIL.Image.resize's result; zv_resize_u8 returns it, zv_preprocess runs k1_patchify_u8 over it).
//
// Tile = 512 input rows x 32 output bytes.  The rows enter as four A operands of 128 rows each (rows 4 m + q for
// q = 0..3, one TMA box over the input viewed as super-rows of four image rows: any row pitch that is a multiple of
// 4 bytes becomes a legal 16-byte-multiple TMA stride), so TMEM lane m of the four
// accumulator blocks holds the four rows 4 m .. 4 m + 3 of one output column: a thread packs them into one 32-bit
// word, and the 32 lanes of a warp store 128 contiguous bytes of the transposed output - no shuffles, no staging.
// B[96][128 NKB] holds, for the chunk's 32 output bytes, the three signed base-256 digits of every 22-bit tap at the
// K position (= input byte) it multiplies, zero elsewhere; it is built in shared memory from the compact tap tables
// by two otherwise idle warps while the previous chunk computes.  D[128][96] per q: digit sums; the epilogue folds
// them (a0 + 256 a1 + 65536 a2 + 2^21) >> 22 and saturates exactly like Pillow's clip8.
// A TMA box must start on a 16-byte boundary of global memory (measured with tools/micro/tc_probe.cu: any other start
// coordinate is an illegal instruction), so every box starts at its window's first byte rounded down to 16 and the 0..15
// bytes of slack move the taps inside B instead: one B variant per distinct slack.  The slack of row 4 m + q is
// ((q' pitch) + offset) mod 16 with q' = row mod 4, so there is 1 variant when the pitch is a multiple of 16 (T always),
// 2 for pitch = 8 mod 16 (a 5000-pixel row), 4 otherwise.
//
// 640 threads: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-3 = B builder (warp 2 owns the TMEM allocation),
// warps 4-19 = epilogue (lane quarter = warp & 3, output bytes 8 ((warp - 4) / 4) ..: four warps per scheduler hide the
// TMEM round trips of one another).  mbarrier pipelines: A ring
// full/empty (TMA <-> MMA), accumulator block full/empty per q (MMA <-> epilogue: block q of the next tile is
// rewritten as soon as every epilogue warp has read it), B double buffer full/empty (builder <-> MMA).
#pragma once

struct RJob {                 // one resample pass over one crop
  int64_t x_off;              // byte offset of tap-table sample `origin`, channel 0, within an input row
  int64_t in_pitch;           // bytes per input row
  uint8_t* out;               // word (output byte j, input row quad m) at out + j * out_pitch + 4 m
  int64_t out_pitch;
  int32_t row0;               // first input row
  int32_t n_rows;             // input rows (= valid bytes per output row)
  int32_t n_out;              // output samples along the axis
  int32_t ch;                 // interleaved channels: 3 (pass 1) or 1 (pass 2)
  int32_t off_b, off_k;       // int32 offsets of the bounds / tap tables in the coefficient area
  int32_t ksize;
  int32_t origin;             // table sample index of input byte x_off
  int32_t tmap;               // index of the input's tensor map
  int32_t nkb;                // 128-byte K blocks per chunk
  const uint8_t* in_base;     // the tensor map's base and extent, for the host emulation of the kernel (tests): the
  int64_t in_dim0;            // device kernel reaches the input through TMA only
  int32_t in_dim1;
  int32_t nvar;               // B variants (1, 2 or 4): distinct box-start slacks of the four row phases
  int32_t lo, hi;             // table samples [lo, hi) exist in the input; taps outside multiply nothing (Image.crop's zero
};                            // fill left / right of the image: the bytes the box holds there belong to a neighbouring row)

constexpr int kTcCols = 32;                 // output bytes per chunk
constexpr int kTcEpiWarps = 16;             // four per TMEM lane quarter, 8 output bytes each
constexpr int kTcEpiCols = kTcCols * 4 / kTcEpiWarps;
constexpr int kTcThreads = 128 + 32 * kTcEpiWarps;
constexpr int kTcN = 96;                    // 2 sub-tiles x 3 digits x 16 columns
constexpr int kTcStageBytes = 128 * 128;    // one A K block: 128 rows x 128 bytes, 128B-swizzled
constexpr int kTcBBlock = kTcN * 128;       // one B K block
constexpr int kTcMaxNkb = 4;
constexpr int kTcBarBytes = 512;

constexpr int kTcMaxBBlocks = 8;            // variants x K blocks of one B buffer (96 KB)
// B blocks = buffers x variants x K blocks
__host__ __device__ constexpr int tc_smem_bytes(int stages, int b_blocks) {
  return 1024 + stages * kTcStageBytes + b_blocks * kTcBBlock + kTcBarBytes;
}
// Two B buffers while both fit 96 KB.  (Measured: trading the second buffer for three more A stages - 8 -> 11 - changes
// nothing, 1.02 -> 1.08 ms for pass 1 of 64 images: the pipeline is not short of bytes in flight.  Alone it runs at the
// HBM rate, 5.9 TB/s; under the step's power-capped clock it is bound by the TMA unit's request rate: ~2.6 L2 requests per
// 128-byte box row that starts on a 16-byte boundary, ~590 SM clocks per 16 KB box at either clock.)
inline int tc_b_buffers(int nvar, int nkb) { return nvar * nkb <= kTcMaxBBlocks / 2 ? 2 : 1; }
inline int tc_stages(int b_blocks) {        // as many A stages as fit beside B (<= 12)
  const int room = 227 * 1024 - tc_smem_bytes(0, b_blocks);
  return std::min(12, room / kTcStageBytes);
}
// Slack of B variant v: bytes between the 16-byte aligned start of the TMA boxes of the rows with row mod nvar = v and
// the first byte of the chunk's window in those rows.
__host__ __device__ __forceinline__ int tc_slack(const RJob& j, int wb0, int v) {
  return (int)(((int64_t)v * j.in_pitch + j.x_off + wb0) & 15);
}

// signed base-256 digits of a 22-bit tap: k = d[0] + 256 d[1] + 65536 d[2], d[0], d[1] in [-128, 127]
__host__ __device__ __forceinline__ void tc_digits(int k, int (&d)[3]) {
  d[0] = ((k & 255) ^ 128) - 128;
  const int k1 = (k - d[0]) >> 8;
  d[1] = ((k1 & 255) ^ 128) - 128;
  d[2] = (k1 - d[1]) >> 8;
}
// B row of (chunk column c, digit l)
__host__ __device__ __forceinline__ int tc_b_row(int c, int l) { return (c >> 4) * 48 + 16 * l + (c & 15); }

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Instruction descriptor, kind::i8: D s32 (bits 4-5 = 2), A unsigned 8-bit (bits 7-9 = 0), B signed 8-bit (bits 10-12 = 1),
// both K-major, N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t kTcIdesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_fence_tensormap(const void* tmap) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// the chunk's first input byte (relative to x_off): first tap of its first output sample
__host__ __device__ __forceinline__ int tc_chunk_wb0(const RJob& j, const int32_t* __restrict__ coef, int chunk) {
  const int o_first = (kTcCols * chunk) / j.ch;
  return j.ch * (coef[j.off_b + 2 * o_first] - j.origin);
}

__global__ void __launch_bounds__(kTcThreads, 1) k1_resample_tc(const RJob* __restrict__ jobs, const int4* __restrict__ items,
                                                                int n_items, const int32_t* __restrict__ coef,
                                                                const CUtensorMap* __restrict__ tmaps, int stages, int nkb_max,
                                                                int nvar_max, int nbuf) {
  using namespace zv::ptx;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + stages * kTcStageBytes;                     // [nbuf][nvar_max][nkb_max][96 rows][128 B]
  const int b_buf_bytes = nvar_max * nkb_max * kTcBBlock;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + nbuf * b_buf_bytes);
  uint64_t* full = bars;                  // [stages]
  uint64_t* empty = bars + 12;            // [stages]
  uint64_t* dfull = bars + 24;            // [4]
  uint64_t* dempty = bars + 28;           // [4]
  uint64_t* bfull = bars + 32;            // [2]
  uint64_t* bempty = bars + 34;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int q = 0; q < 4; ++q) { mbar_init(dfull + q, 1); mbar_init(dempty + q, kTcEpiWarps); }
    for (int b = 0; b < 2; ++b) { mbar_init(bfull + b, 1); mbar_init(bempty + b, 1); }
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                             // the input may be the previous kernel's output (pass 1 -> pass 2, resize -> preprocess)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int4 item = __ldg(items + it);
        const RJob& j = jobs[item.x];
        const CUtensorMap* tm = tmaps + j.tmap;
        tc_fence_tensormap(tm);
        const int64_t xbase = j.x_off + tc_chunk_wb0(j, coef, item.y);
        const int nkb = j.nkb;
        for (int t = item.z; t < item.z + item.w; ++t) {
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            const int r0 = j.row0 + q + 512 * t;                   // input row of lane 0; lane m holds row r0 + 4 m
            const int32_t cy = r0 >> 2;
            const int64_t cx = ((int64_t)(r0 & 3) * j.in_pitch + xbase) & ~(int64_t)15;     // 16-byte aligned box start
            for (int kb = 0; kb < nkb; ++kb) {
              mbar_wait(empty + stage, phase ^ 1);
              mbar_arrive_expect_tx(full + stage, kTcStageBytes);
              tma_load_2d(sA + stage * kTcStageBytes, tm, full + stage, (int32_t)(cx + 128 * kb), cy);
              if (++stage == stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      pdl_trigger();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      uint32_t n_item = 0, n_tile = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n_item) {
        const int4 item = __ldg(items + it);
        const RJob& j = jobs[item.x];
        const int nkb = j.nkb, row0 = j.row0, vmask = j.nvar - 1;
        const uint32_t buf = n_item % (uint32_t)nbuf;
        mbar_wait(bfull + buf, (n_item / (uint32_t)nbuf) & 1);
        tc_fence_after();
        const uint32_t b_buf = smem_u32(sB + buf * b_buf_bytes);
        for (int t = 0; t < item.w; ++t, ++n_tile) {
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            mbar_wait(dempty + q, (n_tile & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(q * kTcN);
            const uint32_t b_addr = b_buf + (uint32_t)(((row0 + q) & vmask) * nkb_max * kTcBBlock);   // the variant of this row phase
            for (int kb = 0; kb < nkb; ++kb) {
              mbar_wait(full + stage, phase);
              tc_fence_after();
              const uint64_t da = umma_desc_k128(smem_u32(sA + stage * kTcStageBytes));
              const uint64_t db = umma_desc_k128(b_addr + kb * kTcBBlock);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma_i8(tmem_d, da + 2 * ks, db + 2 * ks, kTcIdesc, (kb | ks) != 0);
              umma_commit(empty + stage);
              if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(dfull + q);
          }
        }
        umma_commit(bempty + buf);         // every MMA that reads this B buffer has completed when this arrives
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ B builder (64 threads)
    // The taps of an item are fetched into registers BEFORE the wait for the buffer (they do not depend on it), so that with
    // a single B buffer the tensor core only waits for the zero fill + scatter, not for a global-memory round trip.
    const int tid = threadIdx.x - 64;
    constexpr int kHold = 12;                                        // (column, tap) pairs a thread holds: 32 x 23 taps / 64 threads
    uint32_t n_item = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n_item) {
      const int4 item = __ldg(items + it);
      const RJob& j = jobs[item.x];
      const uint32_t buf = n_item % (uint32_t)nbuf;
      uint8_t* B = sB + buf * b_buf_bytes;
      const int nkb = j.nkb, ch = j.ch, ksize = j.ksize, nvar = j.nvar, origin = j.origin;
      const int wb0 = tc_chunk_wb0(j, coef, item.y);
      int slack[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) slack[v] = tc_slack(j, wb0, v);
      const int n_bytes = j.n_out * ch, total = kTcCols * ksize;
      const int32_t* __restrict__ cb = coef + j.off_b;
      const int32_t* __restrict__ ck = coef + j.off_k;
      // (column c, tap t) of pair idx -> output sample o, channel cc; false when the column is past the row's end
      auto locate = [&](int idx, int& c, int& t, int& o, int& cc) -> bool {
        c = idx / ksize; t = idx - c * ksize;
        const int jb = kTcCols * item.y + c;
        o = jb / ch; cc = jb - o * ch;
        return idx < total && jb < n_bytes;
      };
      const int s_lo = j.lo, s_hi = j.hi;
      auto scatter = [&](int c, int t, int cc, int xmin, int cnt, int k) {
        if (t >= cnt || xmin + t < s_lo || xmin + t >= s_hi) return;
        int dg[3];
        tc_digits(k, dg);
        const int kwin = ch * (xmin - origin + t) + cc - wb0;         // input byte of this tap, from the window's first byte
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (v >= nvar) break;
          const int kbyte = kwin + slack[v];                         // K position (0 <= kbyte < 128 nkb by construction)
          if (kbyte < 0 || kbyte >= 128 * nkb) continue;             // (never: the host sized nkb from the same tables)
          const int kb = kbyte >> 7, kin = kbyte & 127;
          uint8_t* blk = B + (v * nkb_max + kb) * kTcBBlock + (kin & 15);
          const int ckk = kin >> 4;
#pragma unroll
          for (int l = 0; l < 3; ++l) {
            const int n = tc_b_row(c, l);
            blk[(n >> 3) * 1024 + (n & 7) * 128 + ((ckk ^ (n & 7)) << 4)] = (uint8_t)dg[l];  // 128B-swizzled K-major row
          }
        }
      };
      int hx[kHold], hc[kHold], hk[kHold];
#pragma unroll
      for (int r = 0; r < kHold; ++r) {
        int c, t, o, cc;
        hx[r] = 0; hc[r] = 0; hk[r] = 0;
        if (locate(tid + 64 * r, c, t, o, cc)) { hx[r] = __ldg(cb + 2 * o); hc[r] = __ldg(cb + 2 * o + 1); hk[r] = __ldg(ck + (int64_t)o * ksize + t); }
      }
      mbar_wait(bempty + buf, ((n_item / (uint32_t)nbuf) & 1) ^ 1);
      for (int v = 0; v < nvar; ++v) {
        uint4* z = reinterpret_cast<uint4*>(B + v * nkb_max * kTcBBlock);
        const int n16 = nkb * kTcBBlock / 16;
        for (int i = tid; i < n16; i += 64) z[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      asm volatile("bar.sync 1, 64;" ::: "memory");
#pragma unroll
      for (int r = 0; r < kHold; ++r) {
        int c, t, o, cc;
        if (locate(tid + 64 * r, c, t, o, cc)) scatter(c, t, cc, hx[r], hc[r], hk[r]);
      }
      for (int idx = tid + 64 * kHold; idx < total; idx += 64) {     // more than 24 taps per column: the rest, fetched now
        int c, t, o, cc;
        if (locate(idx, c, t, o, cc)) scatter(c, t, cc, __ldg(cb + 2 * o), __ldg(cb + 2 * o + 1), __ldg(ck + (int64_t)o * ksize + t));
      }
      tc_fence_proxy_async();              // generic-proxy stores -> visible to the tensor core
      asm volatile("bar.sync 1, 64;" ::: "memory");
      if (tid == 0) mbar_arrive(bfull + buf);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (16 warps)
    constexpr int EC = kTcEpiCols;
    const int lq = warp & 3, part = (warp - 4) >> 2;                 // TMEM lane quarter, column part
    const int m = lq * 32 + lane;
    const int c0 = part * EC;                                        // first chunk column of this thread
    const uint32_t tcol = (uint32_t)((c0 >> 4) * 48 + (c0 & 15));    // TMEM column of (c0, digit 0) within a q block
    uint32_t n_tile = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int4 item = __ldg(items + it);
      const RJob& j = jobs[item.x];
      const int n_bytes = j.n_out * j.ch;
      const int jb0 = kTcCols * item.y + c0;                         // this thread's first output byte
      const int n_rows = j.n_rows;
      const int64_t out_pitch = j.out_pitch;
      uint8_t* const out0 = j.out + (int64_t)jb0 * out_pitch;
      const bool cols_full = jb0 + EC <= n_bytes;
      const bool word_ok = ((reinterpret_cast<uintptr_t>(j.out) | (uintptr_t)out_pitch) & 3) == 0;
      for (int t = item.z; t < item.z + item.w; ++t, ++n_tile) {
        int v[3][EC];
        uint32_t word[EC];
        // software pipeline over the four accumulator blocks: the TMEM loads of block q + 1 are in flight while block q's
        // digit sums are folded (two register sets, statically alternated)
        uint32_t acc[2][3][EC];
        const uint32_t taddr0 = tmem_base + ((uint32_t)(lq * 32) << 16) + tcol;
        mbar_wait(dfull + 0, n_tile & 1);
        tc_fence_after();
        tmem_ld_x8(taddr0, acc[0][0]);
        tmem_ld_x8(taddr0 + 16, acc[0][1]);
        tmem_ld_x8(taddr0 + 32, acc[0][2]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tmem_ld_wait();                                            // block q is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dempty + q);
          if (q < 3) {
            mbar_wait(dfull + q + 1, n_tile & 1);
            tc_fence_after();
            const uint32_t ta = taddr0 + (uint32_t)((q + 1) * kTcN);
            tmem_ld_x8(ta, acc[(q + 1) & 1][0]);
            tmem_ld_x8(ta + 16, acc[(q + 1) & 1][1]);
            tmem_ld_x8(ta + 32, acc[(q + 1) & 1][2]);
          }
#pragma unroll
          for (int i = 0; i < EC; ++i) {
            const int r = finish_raw((int)acc[q & 1][0][i], (int)acc[q & 1][1][i], (int)acc[q & 1][2][i]);
            if (q < 3) v[q][i] = r; else word[i] = pack4_sat(v[0][i], v[1][i], v[2][i], r);
          }
        }
        const int row = 4 * (128 * t + m);                           // first of this thread's four input rows (relative to row0)
        if (row < n_rows) {
          uint8_t* o = out0 + row;
          // whole words when the four rows exist and the output rows are word-aligned (T and U always are; a caller's
          // zv_resize_u8 destination with an odd pitch is not), bytes otherwise
          if (cols_full && word_ok && row + 3 < n_rows) {
#pragma unroll
            for (int i = 0; i < EC; ++i, o += out_pitch) *reinterpret_cast<uint32_t*>(o) = word[i];
          } else {
            const int nb = min(4, n_rows - row);
            const bool wide = word_ok && nb == 4;
#pragma unroll 1
            for (int i = 0; i < EC && jb0 + i < n_bytes; ++i, o += out_pitch) {
              uint32_t wv = word[0];
#pragma unroll
              for (int k = 1; k < EC; ++k) wv = i == k ? word[k] : wv;
              if (wide) *reinterpret_cast<uint32_t*>(o) = wv;
              else for (int b = 0; b < nb; ++b) o[b] = (uint8_t)(wv >> (8 * b));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// zv_preprocess's last stage on the tensor-core route: the finished uint8 image U (oh, ow, 3) -> LUT normalise -> patchify
// (+ the repeated temporal frame) -> patch rows at their HF / window-order position.  One block = one 2x2 merge group.
struct PJob { const uint8_t* u; int64_t u_pitch; int64_t out_row0; int32_t lh, lw; int32_t blk0; int32_t pad_; };

// two neighbouring pixels of one channel -> one shared-memory store (element offset e is even)
template <typename OutT> __device__ __forceinline__ void store_pair(OutT* st, int e, float a, float b) {
  if constexpr (std::is_same<OutT, __nv_bfloat16>::value) *reinterpret_cast<__nv_bfloat162*>(st + e) = __floats2bfloat162_rn(a, b);
  else if constexpr (std::is_same<OutT, __half>::value) *reinterpret_cast<__half2*>(st + e) = __floats2half2_rn(a, b);
  else *reinterpret_cast<float2*>(st + e) = make_float2(a, b);
}

template <typename OutT>
__global__ void __launch_bounds__(256) k1_patchify_u8(const PJob* __restrict__ jobs, int n_jobs, int n_groups, const float* __restrict__ lut,
                                                      OutT* __restrict__ out, int row_order, int wsz) {
  // persistent: a block walks merge groups with a stride of gridDim.x; two staging buffers, one barrier per group.
  // thread = (row yl = tid / 7 of 28, segment sg = tid % 7 of the group's 84-byte rows): 12 bytes = 4 pixels x 3 channels
  // -> per channel two pixel pairs, each one shared-memory store (+ one for the repeated temporal frame)
  __shared__ __align__(16) OutT stage[2][4 * kPatchElems];
  __shared__ float s_lut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];          // uploaded by a memcpy, not by a kernel
  const int tid = threadIdx.x;
  const bool active = tid < 196;
  const int yl = tid / 7, sg = tid - yl * 7;
  const int roff = (yl / 14) * 2 * kPatchElems + (yl % 14) * 14;
  const int x0 = 4 * sg, x1 = 4 * sg + 2;
  const int e0 = roff + (x0 / 14) * kPatchElems + (x0 % 14), e1 = roff + (x1 / 14) * kPatchElems + (x1 % 14);
  zv::ptx::pdl_wait();
  __syncthreads();
  int jn = 0, buf = 0;
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x, buf ^= 1) {
    while (jn + 1 < n_jobs && __ldg(&jobs[jn + 1].blk0) <= g) ++jn;                // groups ascend: the job index only moves forward
    const PJob& j = jobs[jn];
    const int gl = g - j.blk0, lw = j.lw;
    const int my = gl / lw, mx = gl - my * lw;
    OutT* st = stage[buf];
    if (active) {
      const uint32_t* __restrict__ u = reinterpret_cast<const uint32_t*>(j.u + (int64_t)(my * 28 + yl) * j.u_pitch + mx * 84 + 12 * sg);
      const uint32_t w0 = u[0], w1 = u[1], w2 = u[2];      // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
      const float r0 = s_lut[w0 & 255], r1 = s_lut[w0 >> 24], r2 = s_lut[(w1 >> 16) & 255], r3 = s_lut[(w2 >> 8) & 255];
      const float g0 = s_lut[256 + ((w0 >> 8) & 255)], g1 = s_lut[256 + (w1 & 255)], g2 = s_lut[256 + (w1 >> 24)], g3 = s_lut[256 + ((w2 >> 16) & 255)];
      const float b0 = s_lut[512 + ((w0 >> 16) & 255)], b1 = s_lut[512 + ((w1 >> 8) & 255)], b2 = s_lut[512 + (w2 & 255)], b3 = s_lut[512 + (w2 >> 24)];
      store_pair(st, e0, r0, r1);             store_pair(st, e1, r2, r3);
      store_pair(st, e0 + 196, r0, r1);       store_pair(st, e1 + 196, r2, r3);        // + 196: the repeated temporal frame
      store_pair(st, e0 + 392, g0, g1);       store_pair(st, e1 + 392, g2, g3);
      store_pair(st, e0 + 588, g0, g1);       store_pair(st, e1 + 588, g2, g3);
      store_pair(st, e0 + 784, b0, b1);       store_pair(st, e1 + 784, b2, b3);
      store_pair(st, e0 + 980, b0, b1);       store_pair(st, e1 + 980, b2, b3);
    }
    __syncthreads();                                   // the group is staged (and the other buffer's copy-out of two groups ago is long done)
    const int pos = row_order == ZV_ORDER_WINDOW ? window_pos(my, mx, j.lh, lw, wsz) : gl;
    uint4* dst = reinterpret_cast<uint4*>(out + (j.out_row0 + 4 * (int64_t)pos) * kPatchElems);
    const uint4* srcv = reinterpret_cast<const uint4*>(st);
    constexpr int kVec = 4 * kPatchElems * (int)sizeof(OutT) / 16;
    for (int i = tid; i < kVec; i += 256) dst[i] = srcv[i];
  }
}

// ------------------------------------------------------------------------------------------------ host emulation
// The same two kernels on the CPU, over host memory, item by item and lane by lane: used by the CPU tests
// (zv_debug_k1_tc_host) to check the job descriptors, work lists, tap placement and output addressing the host code
// produces - everything except the hardware layouts (swizzle, descriptors, TMEM), which only a GPU run can check.
inline void tc_emulate_resample(const RJob* jobs, const int4* items, int n_items, const int32_t* coef) {
  std::vector<int8_t> B;
  for (int it = 0; it < n_items; ++it) {
    const int4 item = items[it];
    const RJob& j = jobs[item.x];
    const int K = 128 * j.nkb;
    B.assign((size_t)j.nvar * kTcN * K, 0);                                  // [variant][row][K]
    const int wb0 = tc_chunk_wb0(j, coef, item.y);
    const int n_bytes = j.n_out * j.ch;
    for (int c = 0; c < kTcCols; ++c) {
      const int jb = kTcCols * item.y + c;
      if (jb >= n_bytes) continue;
      const int o = jb / j.ch, cc = jb - o * j.ch;
      const int xmin = coef[j.off_b + 2 * o], cnt = coef[j.off_b + 2 * o + 1];
      for (int t = 0; t < cnt; ++t) {
        if (xmin + t < j.lo || xmin + t >= j.hi) continue;
        int dg[3];
        tc_digits(coef[j.off_k + (int64_t)o * j.ksize + t], dg);
        const int kwin = j.ch * (xmin - j.origin + t) + cc - wb0;
        for (int v = 0; v < j.nvar; ++v) {
          const int kbyte = kwin + tc_slack(j, wb0, v);
          if (kbyte < 0 || kbyte >= K) continue;
          for (int l = 0; l < 3; ++l) B[((size_t)v * kTcN + tc_b_row(c, l)) * K + kbyte] = (int8_t)dg[l];
        }
      }
    }
    std::vector<uint8_t> a((size_t)K);
    for (int t = item.z; t < item.z + item.w; ++t)
      for (int m = 0; m < 128; ++m) {
        int val[4][kTcCols];
        for (int q = 0; q < 4; ++q) {
          const int r0 = j.row0 + q + 512 * t;
          const int64_t cy = (r0 >> 2) + m, cx = ((int64_t)(r0 & 3) * j.in_pitch + j.x_off + wb0) & ~(int64_t)15;
          const int var = r0 & (j.nvar - 1);
          for (int k = 0; k < K; ++k) {                                     // one row of the TMA boxes, zero outside the tensor
            const int64_t x = cx + k;
            a[k] = (cy >= 0 && cy < j.in_dim1 && x >= 0 && x < j.in_dim0) ? j.in_base[cy * 4 * j.in_pitch + x] : 0;
          }
          for (int c = 0; c < kTcCols; ++c) {
            int acc[3];
            for (int l = 0; l < 3; ++l) {
              const int8_t* b = B.data() + ((size_t)var * kTcN + tc_b_row(c, l)) * K;
              int s = 0;
              for (int k = 0; k < K; ++k) s += (int)a[k] * (int)b[k];
              acc[l] = s;
            }
            val[q][c] = (acc[2] * 65536 + (acc[1] * 256 + (acc[0] + (1 << (kPrecisionBits - 1))))) >> kPrecisionBits;
          }
        }
        const int row = 4 * (128 * t + m);
        if (row >= j.n_rows) continue;
        for (int c = 0; c < kTcCols; ++c) {
          const int jb = kTcCols * item.y + c;
          if (jb >= n_bytes) continue;
          for (int b = 0; b < std::min(4, j.n_rows - row); ++b)
            j.out[(int64_t)jb * j.out_pitch + row + b] = (uint8_t)std::min(255, std::max(0, val[b][c]));
        }
      }
  }
}

template <typename OutT>
inline void tc_emulate_patchify(const PJob* jobs, int n_jobs, const float* lut, OutT* out, int row_order, int wsz,
                                int (*window_pos_fn)(int, int, int, int, int)) {
  for (int jn = 0; jn < n_jobs; ++jn) {
    const PJob& j = jobs[jn];
    for (int g = 0; g < j.lh * j.lw; ++g) {
      const int my = g / j.lw, mx = g - my * j.lw;
      const int pos = row_order == ZV_ORDER_WINDOW ? window_pos_fn(my, mx, j.lh, j.lw, wsz) : g;
      OutT* dst = out + (j.out_row0 + 4 * (int64_t)pos) * kPatchElems;
      for (int yl = 0; yl < 28; ++yl)
        for (int col = 0; col < 84; ++col) {
          const int xl = col / 3, ch = col - 3 * xl;
          const float v = lut[ch * 256 + j.u[(int64_t)(my * 28 + yl) * j.u_pitch + mx * 84 + col]];
          const int e = ((yl / 14) * 2 + (xl / 14)) * kPatchElems + ch * 392 + (yl % 14) * 14 + (xl % 14);
          dst[e] = (OutT)v;
          dst[e + 196] = (OutT)v;
        }
    }
  }
}
