/* pil_resample.c — CPU ORACLE (test infrastructure, not product code) for the image
 * preprocessing in front of the vision tower:
 *
 *   reference: models/clip_model.py:105-107  processor(images=image)  ->  CLIPImageProcessor
 *   (config/clip_config.yaml:7-14: resize 224 bicubic, centre crop 224, 1/255, CLIP mean/std).
 *
 * The arithmetic lives in third-party code that is not vendored in the reference: transformers'
 * CLIP image processor (4.x "slow" processor = Pillow; unpinned, requirements.txt:4) calling
 * Pillow's Image.resize(BICUBIC).  This file restates Pillow's published resampling algorithm
 * (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
 * ImagingResampleHorizontal_8bpc / Vertical_8bpc, bicubic_filter with a = -0.5) and the
 * processor's size / crop / rescale / normalise rules (image_transforms.py:
 * get_resize_output_image_size(shortest_edge), center_crop, rescale via float64, normalize in
 * float32).  Pinned by tests/test_preprocess_oracle.py against Pillow itself (bit exact on the
 * uint8 stage) and against the installed CLIPImageProcessor.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/build_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION_BITS (32 - 8 - 2)

static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

/* Pillow precompute_coeffs + normalize_coeffs_8bpc for the full-image box [0, in_size). */
static int precompute(int in_size, int out_size, int** bounds_out, int32_t** kk_out) {
  const double support_base = 2.0;
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = support_base * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  double* prekk = (double*)malloc(sizeof(double) * out_size * ksize);
  int* bounds = (int*)malloc(sizeof(int) * out_size * 2);
  int32_t* kk = (int32_t*)malloc(sizeof(int32_t) * out_size * ksize);
  for (int xx = 0; xx < out_size; xx++) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double* k = &prekk[xx * ksize];
    int x;
    for (x = 0; x < xmax; x++) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; x++)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; x++) k[x] = 0;
    bounds[xx * 2 + 0] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
  for (int i = 0; i < out_size * ksize; i++) {
    if (prekk[i] < 0) kk[i] = (int32_t)(-0.5 + prekk[i] * (1 << PRECISION_BITS));
    else kk[i] = (int32_t)(0.5 + prekk[i] * (1 << PRECISION_BITS));
  }
  free(prekk);
  *bounds_out = bounds;
  *kk_out = kk;
  return ksize;
}

static inline uint8_t clip8(int32_t v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

/* Image.resize((out_w, out_h), BICUBIC) of an RGB uint8 HWC image: horizontal pass, then vertical. */
int clm_oracle_resize_u8(const uint8_t* in, int h, int w, int out_h, int out_w, uint8_t* out) {
  if (h <= 0 || w <= 0 || out_h <= 0 || out_w <= 0) return 1;
  int *bx, *by;
  int32_t *kx, *ky;
  const uint8_t* src = in;
  uint8_t* tmp = NULL;
  if (out_w != w) {
    const int ks = precompute(w, out_w, &bx, &kx);
    tmp = (uint8_t*)malloc((size_t)h * out_w * 3);
    for (int y = 0; y < h; y++)
      for (int xx = 0; xx < out_w; xx++) {
        const int xmin = bx[xx * 2], xmax = bx[xx * 2 + 1];
        const int32_t* k = &kx[xx * ks];
        for (int c = 0; c < 3; c++) {
          int32_t ss = 1 << (PRECISION_BITS - 1);
          for (int x = 0; x < xmax; x++) ss += in[((size_t)y * w + x + xmin) * 3 + c] * k[x];
          tmp[((size_t)y * out_w + xx) * 3 + c] = clip8(ss);
        }
      }
    free(bx); free(kx);
    src = tmp;
  }
  if (out_h != h) {
    const int ks = precompute(h, out_h, &by, &ky);
    for (int yy = 0; yy < out_h; yy++) {
      const int ymin = by[yy * 2], ymax = by[yy * 2 + 1];
      const int32_t* k = &ky[yy * ks];
      for (int xx = 0; xx < out_w; xx++)
        for (int c = 0; c < 3; c++) {
          int32_t ss = 1 << (PRECISION_BITS - 1);
          for (int y = 0; y < ymax; y++) ss += src[((size_t)(y + ymin) * out_w + xx) * 3 + c] * k[y];
          out[((size_t)yy * out_w + xx) * 3 + c] = clip8(ss);
        }
    }
    free(by); free(ky);
  } else {
    memcpy(out, src, (size_t)out_h * out_w * 3);
  }
  free(tmp);
  return 0;
}

/* get_resize_output_image_size(size={"shortest_edge": s}, default_to_square=False) */
void clm_oracle_resized_shape(int h, int w, int s, int* out_h, int* out_w) {
  if (w <= h) { *out_w = s; *out_h = (int)((double)s * h / w); }
  else { *out_h = s; *out_w = (int)((double)s * w / h); }
}

/* The whole processor: resize (shortest edge = s, bicubic) -> centre crop s x s -> rescale 1/255 ->
 * normalise; pixel_values fp32 [3, s, s].  crop_u8 (may be NULL) receives the uint8 crop [s, s, 3]. */
int clm_oracle_clip_preprocess(const uint8_t* in, int h, int w, int s, const float* mean, const float* std,
                               float* pixel_values, uint8_t* crop_u8) {
  int rh, rw;
  clm_oracle_resized_shape(h, w, s, &rh, &rw);
  if (rh < s || rw < s) return 2;
  uint8_t* r = (uint8_t*)malloc((size_t)rh * rw * 3);
  int rc = clm_oracle_resize_u8(in, h, w, rh, rw, r);
  if (rc) { free(r); return rc; }
  const int top = (rh - s) / 2, left = (rw - s) / 2;
  for (int y = 0; y < s; y++)
    for (int x = 0; x < s; x++)
      for (int c = 0; c < 3; c++) {
        const uint8_t u = r[((size_t)(y + top) * rw + (x + left)) * 3 + c];
        if (crop_u8) crop_u8[((size_t)y * s + x) * 3 + c] = u;
        const float v = (float)((double)u * 0.00392156862745098);  /* rescale: float64 product, cast to float32 */
        pixel_values[((size_t)c * s + y) * s + x] = (v - mean[c]) / std[c];
      }
  free(r);
  return 0;
}
