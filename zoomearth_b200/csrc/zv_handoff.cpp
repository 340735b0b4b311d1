// LM hand-off, host side: the multimodal rotary position index the language model needs for the
// embeddings the tower produced.  Integer bookkeeping, bit-exact against the reference's own copy
//   get_rope_index   reference src/train/RL/src/open-r1-multimodal/src/open_r1/model/modeling_qwen2_vl.py:967-1114
// and, by flag, against the transformers 5.x variant (HF models/qwen2_5_vl/modeling_qwen2_5_vl.py:1024-1135), which
// differs only at padded positions (0 instead of 1) and in the length the delta is taken against (unpadded).
// Images only: the reference's zoom loop feeds no video (t = 1 everywhere), a video placeholder is an error here.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "zv_common.h"

using namespace zv;

extern "C" {

int zv_rope_index(const int64_t* input_ids, const int64_t* attention_mask, int32_t batch, int32_t seq_len,
                  const int64_t* image_grid_thw, int32_t n_images, int64_t image_token_id, int64_t video_token_id,
                  int64_t vision_start_token_id, int32_t merge, int32_t hf5_semantics, int64_t* position_ids,
                  int64_t* deltas) {
  if (!input_ids || !position_ids || !deltas || batch <= 0 || seq_len <= 0 || merge <= 0)
    return fail(ZV_EINVAL, "zv_rope_index: bad argument");
  const int64_t B = batch, L = seq_len;
  const int64_t pad = hf5_semantics ? 0 : 1;
  auto pos = [&](int d, int64_t b, int64_t l) -> int64_t& { return position_ids[(d * B + b) * L + l]; };

  if (!image_grid_thw) {
    // text-only branch (reference :1091-1112): cumsum(mask) - 1 with pads set to 1, or a plain arange
    for (int64_t b = 0; b < B; ++b) {
      int64_t run = 0, mx = INT64_MIN;
      for (int64_t l = 0; l < L; ++l) {
        int64_t v;
        if (attention_mask) {
          run += attention_mask[b * L + l] != 0;
          v = attention_mask[b * L + l] != 0 ? run - 1 : 1;
        } else {
          v = l;
        }
        for (int d = 0; d < 3; ++d) pos(d, b, l) = v;
        mx = std::max(mx, v);
      }
      deltas[b] = attention_mask ? mx + 1 - L : 0;
    }
    return ZV_OK;
  }

  int32_t image_index = 0;
  std::vector<int64_t> tok, where, p[3];
  for (int64_t b = 0; b < B; ++b) {
    tok.clear();
    where.clear();
    for (int64_t l = 0; l < L; ++l)
      if (!attention_mask || attention_mask[b * L + l] == 1) { tok.push_back(input_ids[b * L + l]); where.push_back(l); }
    const int64_t n = (int64_t)tok.size();
    // images of this sample = vision-start tokens followed by an image placeholder (reference :1034-1037)
    int64_t image_nums = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (tok[i] != vision_start_token_id) continue;
      if (i + 1 >= n) return fail(ZV_EINVAL, "zv_rope_index: sample %lld ends with a vision-start token", (long long)b);
      if (tok[i + 1] == image_token_id) ++image_nums;
      else if (tok[i + 1] == video_token_id)
        return fail(ZV_EINVAL, "zv_rope_index: sample %lld holds a video placeholder; only images are supported", (long long)b);
    }
    for (auto& v : p) v.clear();
    int64_t st = 0, last_max = -1;                  // last_max + 1 == st_idx (reference :1070)
    for (int64_t k = 0; k < image_nums; ++k) {
      int64_t ed = -1;
      for (int64_t i = st; i < n; ++i)
        if (tok[i] == image_token_id) { ed = i; break; }
      if (ed < 0) return fail(ZV_EINVAL, "zv_rope_index: sample %lld has fewer image placeholders than images", (long long)b);
      if (image_index >= n_images)
        return fail(ZV_EINVAL, "zv_rope_index: image_grid_thw has %d rows, the batch needs more", n_images);
      const int64_t t = image_grid_thw[image_index * 3], gh = image_grid_thw[image_index * 3 + 1] / merge,
                    gw = image_grid_thw[image_index * 3 + 2] / merge;
      ++image_index;
      if (t <= 0 || gh <= 0 || gw <= 0) return fail(ZV_EINVAL, "zv_rope_index: image %d has an empty grid", image_index - 1);
      const int64_t text_len = ed - st, st_idx = last_max + 1;
      for (int64_t i = 0; i < text_len; ++i)
        for (auto& v : p) v.push_back(st_idx + i);
      const int64_t base = text_len + st_idx;
      for (int64_t ti = 0; ti < t; ++ti)
        for (int64_t hi = 0; hi < gh; ++hi)
          for (int64_t wi = 0; wi < gw; ++wi) {
            p[0].push_back(base + ti);
            p[1].push_back(base + hi);
            p[2].push_back(base + wi);
          }
      last_max = base + std::max(t, std::max(gh, gw)) - 1;
      st = ed + t * gh * gw;
    }
    if (st < n) {
      const int64_t st_idx = last_max + 1, text_len = n - st;
      for (int64_t i = 0; i < text_len; ++i)
        for (auto& v : p) v.push_back(st_idx + i);
      last_max = st_idx + text_len - 1;
    }
    if ((int64_t)p[0].size() != n)
      return fail(ZV_EINVAL, "zv_rope_index: sample %lld: %lld positions for %lld tokens (placeholder count does not match image_grid_thw)",
                  (long long)b, (long long)p[0].size(), (long long)n);
    for (int d = 0; d < 3; ++d) {
      for (int64_t l = 0; l < L; ++l) pos(d, b, l) = pad;
      for (int64_t i = 0; i < n; ++i) pos(d, b, where[i]) = p[d][i];
    }
    int64_t mx = INT64_MIN;
    for (int d = 0; d < 3; ++d)
      for (int64_t i = 0; i < n; ++i) mx = std::max(mx, p[d][i]);
    if (n == 0) return fail(ZV_EINVAL, "zv_rope_index: sample %lld is fully masked", (long long)b);
    deltas[b] = mx + 1 - (hf5_semantics ? n : L);
  }
  return ZV_OK;
}

// Rows of the flattened (batch * seq_len, hidden) inputs_embeds that hold image placeholders, in row-major order:
// the k-th one receives the k-th image embedding (torch masked_scatter semantics; reference
// modeling_qwen2_vl.py:1191-1207, HF modeling_qwen2_5_vl.py:1179-1218,1301-1307).  Returns the count (>= 0).
int64_t zv_placeholder_rows(const int64_t* input_ids, int64_t n_tokens, int64_t image_token_id, int64_t* rows_out,
                            int64_t rows_cap) {
  if (!input_ids || n_tokens < 0 || (rows_cap > 0 && !rows_out)) return fail(ZV_EINVAL, "zv_placeholder_rows: bad argument");
  int64_t k = 0;
  for (int64_t i = 0; i < n_tokens; ++i)
    if (input_ids[i] == image_token_id) {
      if (k < rows_cap) rows_out[k] = i;
      ++k;
    }
  return k;
}

}  // extern "C"
