// K1: fused crop -> Pillow-exact bicubic resize -> rescale/normalize (LUT) -> patchify-permute.
//
// Replaces, as one device pass over uint8 pixels:
//   PIL.Image.crop               reference src/eval/infer.py:72,75   (zero fill outside the image)
//   PIL.Image.resize(BICUBIC)    reference src/eval/infer.py:84, HF image_transforms.py:368
//                                (Pillow ImagingResample, 8 bpc: horizontal pass -> uint8 -> vertical pass)
//   rescale + normalize          HF image_transforms.py:89-124, 384-442   (exact 768-entry LUT)
//   patchify                     HF models/qwen2_vl/image_processing_pil_qwen2_vl.py:186-214
//
// The two resample passes cannot be merged (the intermediate is rounded to uint8), so the data flow is
//   source u8 (HWC) --hpass--> tmp u8 (rows the vertical pass needs, HWC) --vpass+LUT+permute--> patches.
// Integer arithmetic only: int32 accumulators, 22-bit fixed-point taps computed on the host in fp64
// (zv_host.cpp) - results are bit-identical to Pillow.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <type_traits>
#include <vector>

#include "zv_common.h"

namespace {

constexpr int kPrecisionBits = 22;
constexpr int kPatchElems = 1176;   // 3 * 2 * 14 * 14

struct K1Crop {
  const uint8_t* src;      // image base (device)
  int64_t pitch;           // bytes per image row
  int64_t tmp_off;         // byte offset of this crop's intermediate in the workspace
  int64_t out_row0;        // first patch row of this crop in the output
  int32_t src_h, src_w;
  int32_t x0, y0;          // crop origin in image coordinates (may be negative / extend outside)
  int32_t cw, ch;          // crop extent
  int32_t ow, oh;          // resized extent (multiples of 28)
  int32_t ksh, ksv;        // taps per output column / row
  int32_t ybox0, nrows;    // crop rows [ybox0, ybox0 + nrows) feed the vertical pass
  int32_t off_bh, off_kh, off_bv, off_kv;  // int32 offsets into the coefficient area
  int32_t hblk0, vblk0;    // first block of this crop in the hpass / vpass grids
  int32_t lh, lw;          // merge-group grid (gh/2, gw/2)
  int32_t pad_;
};

__device__ __forceinline__ int find_crop(const K1Crop* crops, int n, int blk, bool vpass) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    int b0 = vpass ? crops[mid].vblk0 : crops[mid].hblk0;
    if (b0 <= blk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ int clip8(int v) { return min(255, max(0, v)); }

// Horizontal pass.  One thread = one (tmp row, output column), all 3 channels.
__global__ void __launch_bounds__(256) k1_hpass(const K1Crop* __restrict__ crops, int n_crops,
                                                const int32_t* __restrict__ coef, uint8_t* __restrict__ tmp_base) {
  const int ci = find_crop(crops, n_crops, blockIdx.x, false);
  const K1Crop c = crops[ci];
  const int64_t item = (int64_t)(blockIdx.x - c.hblk0) * blockDim.x + threadIdx.x;
  if (item >= (int64_t)c.nrows * c.ow) return;
  const int r = (int)(item / c.ow), xx = (int)(item % c.ow);
  const int y_img = c.y0 + c.ybox0 + r;
  const int xmin = coef[c.off_bh + 2 * xx], cnt = coef[c.off_bh + 2 * xx + 1];
  const int32_t* __restrict__ k = coef + c.off_kh + (int64_t)xx * c.ksh;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  if (y_img >= 0 && y_img < c.src_h) {
    const uint8_t* __restrict__ row = c.src + (int64_t)y_img * c.pitch;
    for (int t = 0; t < cnt; ++t) {
      const int x_img = c.x0 + xmin + t;
      if (x_img >= 0 && x_img < c.src_w) {
        const int kv = __ldg(k + t);
        const uint8_t* px = row + 3 * (int64_t)x_img;
        a0 += (int)__ldg(px) * kv;
        a1 += (int)__ldg(px + 1) * kv;
        a2 += (int)__ldg(px + 2) * kv;
      }
    }
  }
  uint8_t* o = tmp_base + c.tmp_off + ((int64_t)r * c.ow + xx) * 3;
  o[0] = (uint8_t)clip8(a0 >> kPrecisionBits);
  o[1] = (uint8_t)clip8(a1 >> kPrecisionBits);
  o[2] = (uint8_t)clip8(a2 >> kPrecisionBits);
}

// Position of merge group (my, mx) in the tower's window order (closed form of argsort(window_index),
// HF modeling_qwen2_5_vl.py:411-451): windows of ws x ws merge groups, row-major over windows, row-major inside.
__device__ __forceinline__ int window_pos(int my, int mx, int lh, int lw, int ws) {
  const int wy = my / ws, wx = mx / ws;
  const int bh = min(ws, lh - wy * ws), bw = min(ws, lw - wx * ws);
  return wy * ws * lw + wx * ws * bh + (my - wy * ws) * bw + (mx - wx * ws);
}

// Vertical pass + LUT normalise + patchify.  One block = one 2x2 merge group (28 x 28 pixels, 4 patch rows);
// the block assembles the four 1176-element rows in shared memory and streams them out with 16-byte stores.
template <typename OutT>
__global__ void __launch_bounds__(256) k1_vpass(const K1Crop* __restrict__ crops, int n_crops,
                                                const int32_t* __restrict__ coef,
                                                const uint8_t* __restrict__ tmp_base, const float* __restrict__ lut,
                                                OutT* __restrict__ out, int row_order, int ws) {
  __shared__ __align__(16) OutT stage[4 * kPatchElems];
  __shared__ float s_lut[768];
  const int ci = find_crop(crops, n_crops, blockIdx.x, true);
  const K1Crop c = crops[ci];
  const int g = blockIdx.x - c.vblk0;           // merge group, raster order inside the crop
  const int my = g / c.lw, mx = g % c.lw;
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const uint8_t* __restrict__ tmp = tmp_base + c.tmp_off;
  const int64_t tpitch = (int64_t)c.ow * 3;
  for (int item = threadIdx.x; item < 28 * 84; item += blockDim.x) {
    const int yl = item / 84, col = item % 84;             // col = xl * 3 + ch
    const int xl = col / 3, ch = col % 3;
    const int yy = my * 28 + yl;
    const int ymin = coef[c.off_bv + 2 * yy] - c.ybox0, cnt = coef[c.off_bv + 2 * yy + 1];
    const int32_t* __restrict__ k = coef + c.off_kv + (int64_t)yy * c.ksv;
    const uint8_t* __restrict__ p = tmp + (int64_t)ymin * tpitch + (int64_t)(mx * 28) * 3 + col;
    int acc = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < cnt; ++t) acc += (int)__ldg(p + (int64_t)t * tpitch) * __ldg(k + t);
    const float v = s_lut[ch * 256 + clip8(acc >> kPrecisionBits)];
    const int j = (yl / 14) * 2 + (xl / 14);
    const int e = j * kPatchElems + ch * 392 + (yl % 14) * 14 + (xl % 14);
    OutT o;
    if constexpr (std::is_same<OutT, __nv_bfloat16>::value) o = __float2bfloat16_rn(v);
    else if constexpr (std::is_same<OutT, __half>::value) o = __float2half_rn(v);
    else o = v;
    stage[e] = o;            // temporal frame 0
    stage[e + 196] = o;      // temporal frame 1 = the repeated frame (HF :189-193)
  }
  __syncthreads();
  const int pos = row_order == ZV_ORDER_WINDOW ? window_pos(my, mx, c.lh, c.lw, ws) : g;
  uint4* dst = reinterpret_cast<uint4*>(out + (c.out_row0 + 4 * (int64_t)pos) * kPatchElems);
  const uint4* srcv = reinterpret_cast<const uint4*>(stage);
  constexpr int kVec = 4 * kPatchElems * (int)sizeof(OutT) / 16;
  for (int i = threadIdx.x; i < kVec; i += blockDim.x) dst[i] = srcv[i];
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct Layout {
  int64_t off_desc = 0, off_lut = 0, off_coef = 0, off_tmp = 0, bytes = 0;
  std::vector<int64_t> tmp_off;              // per crop, relative to off_tmp
  std::vector<int32_t> ybox0, nrows;
  std::map<std::pair<int32_t, int32_t>, std::pair<int32_t, int32_t>> coef_at;  // (in,out) -> (bounds off, kk off)
  std::vector<int32_t> coef;                 // concatenated int32 tables
};

// Workspace layout shared by zv_preprocess_workspace_bytes and zv_preprocess.
int build_layout(int32_t n, const int32_t* crop_box, const int32_t* resized_hw, Layout* L, bool fill) {
  L->tmp_off.resize(n); L->ybox0.resize(n); L->nrows.resize(n);
  int64_t coef_ints = 0, tmp_bytes = 0;
  for (int32_t i = 0; i < n; ++i) {
    const int32_t cw = crop_box[4 * i + 2] - crop_box[4 * i], ch = crop_box[4 * i + 3] - crop_box[4 * i + 1];
    const int32_t oh = resized_hw[2 * i], ow = resized_hw[2 * i + 1];
    if (cw <= 0 || ch <= 0 || oh <= 0 || ow <= 0 || oh % 28 || ow % 28)
      return zv::fail(ZV_EINVAL, "zv_preprocess: crop %d has extent %dx%d -> %dx%d (need >0 and multiples of 28)",
                      i, cw, ch, ow, oh);
    const std::pair<int32_t, int32_t> keys[2] = {{cw, ow}, {ch, oh}};
    for (const auto& key : keys) {
      if (L->coef_at.count(key)) continue;
      const int32_t ks = zv::resample_ksize(key.first, key.second);
      L->coef_at[key] = {(int32_t)coef_ints, (int32_t)(coef_ints + 2 * (int64_t)key.second)};
      if (fill) {
        zv::AxisCoeffs ac;
        zv::resample_coeffs(key.first, key.second, &ac);
        L->coef.insert(L->coef.end(), ac.bounds.begin(), ac.bounds.end());
        L->coef.insert(L->coef.end(), ac.kk.begin(), ac.kk.end());
      }
      coef_ints += 2 * (int64_t)key.second + (int64_t)key.second * ks;
      if (coef_ints > INT32_MAX) return zv::fail(ZV_EINVAL, "zv_preprocess: coefficient tables exceed 2^31 entries");
    }
    // rows of the crop the vertical pass reads: [first tap of row 0, last tap of the last row]
    int32_t y_first, y_last;
    if (ch == oh) { y_first = 0; y_last = ch; }
    else {
      const double scale = (double)ch / oh, fs = scale < 1.0 ? 1.0 : scale, support = 2.0 * fs;
      double c0 = 0 + (0 + 0.5) * scale, c1 = 0 + ((oh - 1) + 0.5) * scale;
      y_first = std::max<int32_t>(0, (int32_t)(c0 - support + 0.5));
      y_last = std::min<int32_t>(ch, (int32_t)(c1 + support + 0.5));
    }
    L->ybox0[i] = y_first;
    L->nrows[i] = y_last - y_first;
    L->tmp_off[i] = tmp_bytes;
    tmp_bytes += align_up((int64_t)L->nrows[i] * ow * 3, 256);
  }
  int64_t off = 0;
  L->off_desc = off; off = align_up(off + (int64_t)n * sizeof(K1Crop), 256);
  L->off_lut = off; off = align_up(off + 768 * sizeof(float), 256);
  L->off_coef = off; off = align_up(off + coef_ints * (int64_t)sizeof(int32_t), 256);
  L->off_tmp = off; off += tmp_bytes;
  L->bytes = off;
  return ZV_OK;
}

}  // namespace

extern "C" {

int64_t zv_preprocess_workspace_bytes(int32_t n, const int32_t* crop_box, const int32_t* resized_hw) {
  if (n <= 0 || !crop_box || !resized_hw) return zv::fail(ZV_EINVAL, "zv_preprocess_workspace_bytes: bad argument");
  Layout L;
  int rc = build_layout(n, crop_box, resized_hw, &L, false);
  return rc ? rc : L.bytes;
}

int zv_preprocess(const zv_cfg* cfg, int32_t n, const uint8_t* const* src_dev, const int32_t* src_hw,
                  const int64_t* src_pitch, const int32_t* crop_box, const int32_t* resized_hw,
                  const int64_t* row_off, void* out_dev, int32_t out_dtype, int32_t row_order, void* workspace_dev,
                  int64_t workspace_bytes, void* stream_) {
  zv::reset_launch_count();
  if (!cfg || n <= 0 || !src_dev || !src_hw || !src_pitch || !crop_box || !resized_hw || !out_dev || !workspace_dev)
    return zv::fail(ZV_EINVAL, "zv_preprocess: null argument");
  if (cfg->patch != 14 || cfg->merge != 2 || cfg->temporal != 2)
    return zv::fail(ZV_EINVAL, "zv_preprocess: only patch=14, merge=2, temporal=2 is built");
  if (out_dtype != ZV_F32 && out_dtype != ZV_BF16 && out_dtype != ZV_F16) return zv::fail(ZV_EINVAL, "zv_preprocess: bad out_dtype");
  if (row_order != ZV_ORDER_HF && row_order != ZV_ORDER_WINDOW) return zv::fail(ZV_EINVAL, "zv_preprocess: bad row_order");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return zv::fail(ZV_ENODEV, "zv_preprocess: no CUDA device");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Layout L;
  int rc = build_layout(n, crop_box, resized_hw, &L, true);
  if (rc) return rc;
  if (workspace_bytes < L.bytes)
    return zv::fail(ZV_ENOMEM, "zv_preprocess: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)L.bytes);

  // host image of [descriptors | LUT | coefficient tables]
  std::vector<uint8_t> host((size_t)L.off_tmp, 0);
  K1Crop* d = reinterpret_cast<K1Crop*>(host.data() + L.off_desc);
  int64_t hblk = 0, vblk = 0, row = 0;
  for (int32_t i = 0; i < n; ++i) {
    K1Crop& c = d[i];
    c.src = src_dev[i]; c.pitch = src_pitch[i];
    c.src_h = src_hw[2 * i]; c.src_w = src_hw[2 * i + 1];
    c.x0 = crop_box[4 * i]; c.y0 = crop_box[4 * i + 1];
    c.cw = crop_box[4 * i + 2] - c.x0; c.ch = crop_box[4 * i + 3] - c.y0;
    c.oh = resized_hw[2 * i]; c.ow = resized_hw[2 * i + 1];
    if (!c.src || c.pitch < 3 * (int64_t)c.src_w) return zv::fail(ZV_EINVAL, "zv_preprocess: image %d has a null pointer or short pitch", i);
    c.ksh = zv::resample_ksize(c.cw, c.ow); c.ksv = zv::resample_ksize(c.ch, c.oh);
    const auto h = L.coef_at[{c.cw, c.ow}], v = L.coef_at[{c.ch, c.oh}];
    c.off_bh = h.first; c.off_kh = h.second; c.off_bv = v.first; c.off_kv = v.second;
    c.ybox0 = L.ybox0[i]; c.nrows = L.nrows[i];
    c.tmp_off = L.off_tmp + L.tmp_off[i];
    c.lh = c.oh / 28; c.lw = c.ow / 28;
    c.out_row0 = row_off ? row_off[i] : row;
    row += (int64_t)(c.oh / 14) * (c.ow / 14);
    c.hblk0 = (int32_t)hblk; c.vblk0 = (int32_t)vblk;
    hblk += ((int64_t)c.nrows * c.ow + 255) / 256;
    vblk += (int64_t)c.lh * c.lw;
    if (hblk > INT32_MAX || vblk > INT32_MAX) return zv::fail(ZV_EINVAL, "zv_preprocess: batch too large for one launch");
  }
  zv::normalize_lut(cfg, reinterpret_cast<float*>(host.data() + L.off_lut));
  std::memcpy(host.data() + L.off_coef, L.coef.data(), L.coef.size() * sizeof(int32_t));

  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  cudaError_t e = cudaMemcpyAsync(ws, host.data(), host.size(), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return zv::fail(ZV_ECUDA, "zv_preprocess: table upload: %s", cudaGetErrorString(e));
  const K1Crop* dcrops = reinterpret_cast<const K1Crop*>(ws + L.off_desc);
  const int32_t* dcoef = reinterpret_cast<const int32_t*>(ws + L.off_coef);
  const float* dlut = reinterpret_cast<const float*>(ws + L.off_lut);
  const int wsz = cfg->window / cfg->merge / cfg->patch;
  {
    zv::KernelTimer timer(zv::KC_K1_HPASS, stream);
    k1_hpass<<<(unsigned)hblk, 256, 0, stream>>>(dcrops, n, dcoef, ws);
  }
  zv::KernelTimer timer(zv::KC_K1_VPASS, stream);
  if (out_dtype == ZV_BF16)
    k1_vpass<__nv_bfloat16><<<(unsigned)vblk, 256, 0, stream>>>(dcrops, n, dcoef, ws, dlut,
                                                                 static_cast<__nv_bfloat16*>(out_dev), row_order, wsz);
  else if (out_dtype == ZV_F16)
    k1_vpass<__half><<<(unsigned)vblk, 256, 0, stream>>>(dcrops, n, dcoef, ws, dlut, static_cast<__half*>(out_dev), row_order, wsz);
  else
    k1_vpass<float><<<(unsigned)vblk, 256, 0, stream>>>(dcrops, n, dcoef, ws, dlut, static_cast<float*>(out_dev),
                                                         row_order, wsz);
  zv::count_launch(2);
  e = cudaGetLastError();
  if (e != cudaSuccess) return zv::fail(ZV_ECUDA, "zv_preprocess: launch: %s", cudaGetErrorString(e));
  return ZV_OK;
}

}  // extern "C"
