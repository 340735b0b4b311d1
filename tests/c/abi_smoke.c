/* C caller of libzoomvit.so's host entry points (no GPU needed): shows that include/zoomvit.h is a plain C ABI and
 * checks the known answers of SURVEY 8c.  Built and run by tests/test_c_abi_from_c.py with gcc. */
#include <stdio.h>
#include <string.h>

#include "zoomvit.h"

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL: " __VA_ARGS__); printf("\n"); ++fails; } } while (0)

int main(void) {
  zv_cfg cfg;
  zv_default_cfg(&cfg);
  CHECK(cfg.patch == 14 && cfg.merge == 2 && cfg.temporal == 2 && cfg.window == 112 && cfg.min_size == 512, "default cfg");

  /* cut_image boxes (reference src/eval/infer.py:41-76), 5000x5000 image */
  const double boxes[5][4] = {{1000, 1200, 2300, 2100}, {100.7, 50.2, 300.9, 260.1}, {4900, 4950, 4990, 4999},
                              {-20, -30, 100, 90}, {2000, 2000, 2600, 2300}};
  const int want[5][4] = {{1000, 1200, 2300, 2100}, {0, 0, 512, 512}, {4488, 4488, 5000, 5000}, {0, 0, 512, 512},
                          {2044, 1894, 2556, 2406}};
  for (int i = 0; i < 5; ++i) {
    int32_t out[4];
    CHECK(zv_cut_box(5000, 5000, boxes[i], 512, out) == ZV_OK, "zv_cut_box rc");
    CHECK(memcmp(out, want[i], sizeof out) == 0, "cut box %d: %d %d %d %d", i, out[0], out[1], out[2], out[3]);
  }
  /* a final crop box (min_size < 0: no cut_image rule) with right < left is Pillow's ValueError */
  const int32_t img_hw[2] = {10, 4000};
  const double bad[4] = {50, 0, 40, 10};
  int32_t crop[4], rhw[2];
  int64_t thw[3];
  zv_cfg raw = cfg;
  raw.min_size = -1;
  CHECK(zv_geometry(&raw, 1, img_hw, bad, crop, rhw, thw) == ZV_EINVAL_BOX, "bad box rc");
  CHECK(strstr(zv_last_error(), "Coordinate 'right' is less than 'left'") != NULL, "box message: %s", zv_last_error());

  /* smart_resize (HF image_processing_pil_qwen2_vl.py:57-83) */
  const long long sr[6][5] = {{5000, 5000, 1003520, 980, 980},   {5000, 5000, 12845056, 3556, 3556}, {512, 512, 12845056, 504, 504},
                              {900, 1300, 1003520, 812, 1176},   {518, 70, 12845056, 504, 56},       {30, 30, 12845056, 56, 56}};
  for (int i = 0; i < 6; ++i) {
    int32_t hw[2];
    CHECK(zv_smart_resize((int32_t)sr[i][0], (int32_t)sr[i][1], 28, 3136, sr[i][2], hw) == ZV_OK, "zv_smart_resize rc");
    CHECK(hw[0] == sr[i][3] && hw[1] == sr[i][4], "smart_resize %d: %d x %d", i, hw[0], hw[1]);
  }
  int32_t hw[2];
  CHECK(zv_smart_resize(10, 4000, 28, 3136, 12845056, hw) == ZV_EINVAL_ASPECT, "aspect ratio error code");
  CHECK(strstr(zv_last_error(), "absolute aspect ratio must be smaller than 200") != NULL, "aspect message: %s", zv_last_error());

  /* resize_image dims (reference src/eval/infer.py:78-85) */
  int32_t wh[2];
  double inv;
  CHECK(zv_resize_dims(1300, 900, 512, wh, &inv) == ZV_OK && wh[0] == 512 && wh[1] == 354, "resize_dims %d %d", wh[0], wh[1]);

  /* resize_image variants (demo.py:86-93, SFT.py:76-81 always resizes, customized_funcs.py:76-85) and SFT.py's cut_image
   * (crop -> resize to min side 512 -> centre crop); known answers from tests/golden/flows.json (the reference's own code) */
  CHECK(zv_resize_dims_ex(5000, 5000, 1024, 1, wh, &inv) == ZV_OK && wh[0] == 1024 && wh[1] == 1024, "resize_dims_ex sft %d %d", wh[0], wh[1]);
  CHECK(zv_resize_dims_ex(300, 200, 1024, 1, wh, NULL) == ZV_OK && wh[0] == 1024 && wh[1] == 682, "sft upscales: %d %d", wh[0], wh[1]);
  CHECK(zv_resize_dims_ex(300, 200, 1024, 0, wh, NULL) == ZV_OK && wh[0] == 300 && wh[1] == 200, "infer keeps small images");
  CHECK(zv_resize_dims_ex(4000, 40, 512, 2, wh, NULL) == ZV_OK && wh[0] == 3000 && wh[1] == 30, "custom min_scale: %d %d", wh[0], wh[1]);
  {
    const double bb[4] = {1000, 1200, 2300, 2100};
    int32_t box[4], rs[2], cb[4];
    CHECK(zv_cut_box_sft(5000, 5000, bb, 512, box, rs, cb) == ZV_OK, "zv_cut_box_sft rc");
    CHECK(box[0] == 1000 && box[3] == 2100 && rs[0] == 739 && rs[1] == 512 && cb[0] == 113 && cb[1] == 0 && cb[2] == 625 && cb[3] == 512,
          "cut_box_sft: resize %d x %d, centre box %d %d %d %d", rs[0], rs[1], cb[0], cb[1], cb[2], cb[3]);
  }

  /* plan: window index of grid (1, 26, 36) */
  const int64_t grid[3] = {1, 26, 36};
  zv_plan* plan = NULL;
  CHECK(zv_plan_create(&cfg, 1, grid, &plan) == ZV_OK && plan != NULL, "zv_plan_create");
  if (plan) {
    const int64_t first[20] = {0, 1, 2, 3, 18, 19, 20, 21, 36, 37, 38, 39, 54, 55, 56, 57, 4, 5, 6, 7};
    CHECK(zv_plan_num_patches(plan) == 936 && zv_plan_num_tokens(plan) == 234, "plan sizes");
    CHECK(memcmp(zv_plan_window_index(plan), first, sizeof first) == 0, "window_index[:20]");
    int32_t n = 0;
    const int32_t* cu = zv_plan_cu_window_raw(plan, &n);
    const int32_t cu8[8] = {0, 64, 128, 192, 256, 288, 352, 416};
    CHECK(n >= 8 && memcmp(cu, cu8, sizeof cu8) == 0, "cu_window_seqlens[:8]");
    zv_plan_free(plan);
  }

  /* LM hand-off: rope index of [t, <vs>, img x4, <ve>, t] with a (1, 4, 4) grid */
  const int64_t ids[8] = {5, 151652, 151655, 151655, 151655, 151655, 151653, 7};
  const int64_t g44[3] = {1, 4, 4};
  int64_t pos[24], delta[1];
  CHECK(zv_rope_index(ids, NULL, 1, 8, g44, 1, 151655, 151656, 151652, 2, 0, pos, delta) == ZV_OK, "zv_rope_index rc");
  const int64_t want_pos[24] = {0, 1, 2, 2, 2, 2, 4, 5, 0, 1, 2, 2, 3, 3, 4, 5, 0, 1, 2, 3, 2, 3, 4, 5};
  CHECK(memcmp(pos, want_pos, sizeof pos) == 0 && delta[0] == -2, "rope index / delta %lld", (long long)delta[0]);
  int64_t rows[8];
  CHECK(zv_placeholder_rows(ids, 8, 151655, rows, 8) == 4 && rows[0] == 2 && rows[3] == 5, "placeholder rows");

  printf("%s (%s)\n", fails ? "FAILED" : "ok", zv_version());
  return fails ? 1 : 0;
}
