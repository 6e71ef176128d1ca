// clm_preprocess.cu — CLIP image preprocessing on the GPU: decoded uint8 RGB
// images of arbitrary size -> pixel_values fp32 [B, 3, S, S], i.e. what
//   processor(images=image, return_tensors="pt")            reference models/clip_model.py:105-107
// produces (resize shortest edge to S with Pillow's antialiased bicubic, centre crop S x S, 1/255,
// CLIP mean / std — config/clip_config.yaml:7-14).  At >= 20k images/s per GPU the host-side
// Pillow resize (hundreds of images/s on 8 cores) would otherwise be the bottleneck.
//
// The uint8 stage is Pillow's algorithm to the bit (Resample.c): per output index a window
// [xmin, xmin+xmax) of taps with double-precision bicubic weights, normalised, rounded to 22-bit
// fixed point; a horizontal pass that rounds to uint8, then a vertical pass that rounds to uint8.
// Three kernels per batch, all stream-ordered in caller-provided workspace:
//   coef_kernel   the tap windows and fixed-point weights of both axes, in IEEE double with every
//                 rounding Pillow's x86-64 build has (explicit __dmul_rn/__dadd_rn: no FMA contraction)
//   hpass_kernel  horizontal pass for the input rows the cropped output needs -> uint8 [rows, S, 3]
//   vpass_kernel  vertical pass + rescale + normalise -> fp32 CHW
// Only the S x S crop window is ever computed (the values Pillow would crop away are skipped).
#include <math.h>

#include <vector>

#include "clm_common.cuh"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

struct ImgMeta {
  const uint8_t* data;
  long long tmp_off;   // byte offset of this image's horizontal-pass output in the workspace
  int h, w, stride;    // input rows, columns, row pitch in bytes
  int rh, rw;          // size after the shortest-edge resize
  int top, left;       // crop offset inside the resized image
  int pad;
};

__device__ __forceinline__ double bicubic_filter(double x) {
  // ((a + 2) x - (a + 3)) x x + 1   and   (((x - 5) x + 8) x - 4) a   with a = -0.5, in Pillow's order
  if (x < 0.0) x = -x;
  if (x < 1.0) {
    double t = __dmul_rn(1.5, x);
    t = __dsub_rn(t, 2.5);
    t = __dmul_rn(t, x);
    t = __dmul_rn(t, x);
    return __dadd_rn(t, 1.0);
  }
  if (x < 2.0) {
    double t = __dsub_rn(x, 5.0);
    t = __dmul_rn(t, x);
    t = __dadd_rn(t, 8.0);
    t = __dmul_rn(t, x);
    t = __dsub_rn(t, 4.0);
    return __dmul_rn(t, -0.5);
  }
  return 0.0;
}

// grid (2 axes, B), one thread per output index of the crop window.
// bounds [B][2][S][2] (first tap, tap count), coef [B][2][S][ks] 22-bit fixed point.
__global__ void coef_kernel(const ImgMeta* __restrict__ meta, int S, int ks, int* __restrict__ bounds,
                            int* __restrict__ coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int axis = blockIdx.y;  // 0 = x (horizontal), 1 = y (vertical)
  const int b = blockIdx.z;
  if (i >= S) return;
  const ImgMeta m = meta[b];
  const int in_size = axis == 0 ? m.w : m.h;
  const int out_size = axis == 0 ? m.rw : m.rh;
  const int xx = i + (axis == 0 ? m.left : m.top);
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double center = __dmul_rn(static_cast<double>(xx) + 0.5, scale);
  const double ss = __ddiv_rn(1.0, filterscale);
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
    ww = __dadd_rn(ww, bicubic_filter(arg));
  }
  int* k = coef + ((static_cast<size_t>(b) * 2 + axis) * S + i) * ks;
  for (int x = 0; x < ks; ++x) {
    int v = 0;
    if (x < xmax) {
      const double arg = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
      double w = bicubic_filter(arg);
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      const double scaled = __dmul_rn(w, static_cast<double>(1 << kPrecisionBits));
      v = w < 0 ? static_cast<int>(__dadd_rn(-0.5, scaled)) : static_cast<int>(__dadd_rn(0.5, scaled));
    }
    k[x] = v;
  }
  int* bd = bounds + ((static_cast<size_t>(b) * 2 + axis) * S + i) * 2;
  bd[0] = xmin;
  bd[1] = xmax;
}

__device__ __forceinline__ uint32_t clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<uint32_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// grid (ceil(max_rows * S / 256), B): thread = (needed input row, output column) -> 3 channels
__global__ void __launch_bounds__(256)
hpass_kernel(const ImgMeta* __restrict__ meta, int S, int ks, const int* __restrict__ bounds,
             const int* __restrict__ coef, uint8_t* __restrict__ tmp_base) {
  const int b = blockIdx.y;
  const ImgMeta m = meta[b];
  const int* by = bounds + (static_cast<size_t>(b) * 2 + 1) * S * 2;
  const int y_first = by[0];
  const int y_last = by[(S - 1) * 2] + by[(S - 1) * 2 + 1];  // one past the last input row needed
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int yr = idx / S, i = idx - yr * S;
  if (yr >= y_last - y_first) return;
  const int* bx = bounds + ((static_cast<size_t>(b) * 2 + 0) * S + i) * 2;
  const int xmin = bx[0], xmax = bx[1];
  const int* k = coef + ((static_cast<size_t>(b) * 2 + 0) * S + i) * ks;
  const uint8_t* row = m.data + static_cast<size_t>(y_first + yr) * m.stride + static_cast<size_t>(xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < xmax; ++x) {
    const int kv = __ldg(k + x);
    s0 += row[3 * x + 0] * kv;
    s1 += row[3 * x + 1] * kv;
    s2 += row[3 * x + 2] * kv;
  }
  uint8_t* o = tmp_base + m.tmp_off + (static_cast<size_t>(yr) * S + i) * 3;
  o[0] = static_cast<uint8_t>(clip8(s0));
  o[1] = static_cast<uint8_t>(clip8(s1));
  o[2] = static_cast<uint8_t>(clip8(s2));
}

// grid (ceil(S*S/256), B): thread = output pixel -> 3 channels, fp32 CHW
__global__ void __launch_bounds__(256)
vpass_kernel(const ImgMeta* __restrict__ meta, int S, int ks, const int* __restrict__ bounds,
             const int* __restrict__ coef, const uint8_t* __restrict__ tmp_base, float3 mean, float3 stdv,
             float* __restrict__ out) {
  const int b = blockIdx.y;
  const ImgMeta m = meta[b];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * S) return;
  const int yy = idx / S, xx = idx - yy * S;
  const int* by = bounds + (static_cast<size_t>(b) * 2 + 1) * S * 2;
  const int y_first = by[0];
  const int ymin = by[yy * 2], ymax = by[yy * 2 + 1];
  const int* k = coef + ((static_cast<size_t>(b) * 2 + 1) * S + yy) * ks;
  const uint8_t* col = tmp_base + m.tmp_off + (static_cast<size_t>(ymin - y_first) * S + xx) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < ymax; ++y) {
    const int kv = __ldg(k + y);
    const uint8_t* px = col + static_cast<size_t>(y) * S * 3;
    s0 += px[0] * kv;
    s1 += px[1] * kv;
    s2 += px[2] * kv;
  }
  // rescale: float64 product cast to float32 (image_transforms.rescale), then (x - mean) / std in float32
  const float v0 = __double2float_rn(__dmul_rn(static_cast<double>(clip8(s0)), 0.00392156862745098));
  const float v1 = __double2float_rn(__dmul_rn(static_cast<double>(clip8(s1)), 0.00392156862745098));
  const float v2 = __double2float_rn(__dmul_rn(static_cast<double>(clip8(s2)), 0.00392156862745098));
  float* o = out + static_cast<size_t>(b) * 3 * S * S + idx;
  o[0] = __fdiv_rn(__fsub_rn(v0, mean.x), stdv.x);
  o[static_cast<size_t>(S) * S] = __fdiv_rn(__fsub_rn(v1, mean.y), stdv.y);
  o[2 * static_cast<size_t>(S) * S] = __fdiv_rn(__fsub_rn(v2, mean.z), stdv.z);
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Plan {
  std::vector<ImgMeta> meta;
  int ks = 0;        // taps per coefficient row (max over images and axes)
  int max_rows = 0;  // upper bound of input rows any image needs
  size_t meta_off = 0, bounds_off = 0, coef_off = 0, tmp_off = 0, total = 0;
};

// Pillow: ksize = ceil(support) * 2 + 1 with support = 2 * max(scale, 1)
int ksize_for(int in_size, int out_size) {
  double scale = static_cast<double>(in_size) / out_size;
  if (scale < 1.0) scale = 1.0;
  return static_cast<int>(ceil(2.0 * scale)) * 2 + 1;
}

int make_plan(const clm_image_desc* d, int batch, int S, Plan* p) {
  p->meta.resize(batch);
  size_t tmp = 0;
  for (int b = 0; b < batch; ++b) {
    CLM_REQUIRE(d[b].data && d[b].height > 0 && d[b].width > 0 && d[b].row_stride_bytes >= 3 * d[b].width,
                "clm_preprocess_images: bad descriptor %d (data=%p %dx%d stride %d)", b,
                (const void*)d[b].data, d[b].height, d[b].width, d[b].row_stride_bytes);
    ImgMeta& m = p->meta[b];
    m.data = d[b].data;
    m.h = d[b].height; m.w = d[b].width; m.stride = d[b].row_stride_bytes;
    // get_resize_output_image_size(shortest_edge=S): the short side becomes S, the long one int(S*long/short)
    if (m.w <= m.h) { m.rw = S; m.rh = static_cast<int>(static_cast<double>(S) * m.h / m.w); }
    else { m.rh = S; m.rw = static_cast<int>(static_cast<double>(S) * m.w / m.h); }
    m.top = (m.rh - S) / 2;
    m.left = (m.rw - S) / 2;
    m.pad = 0;
    const int kx = ksize_for(m.w, m.rw), ky = ksize_for(m.h, m.rh);
    if (kx > p->ks) p->ks = kx;
    if (ky > p->ks) p->ks = ky;
    // rows the crop window can touch: S output rows spaced h/rh apart plus one support on each side
    const double sy = static_cast<double>(m.h) / m.rh;
    int rows = static_cast<int>(ceil(S * sy + 4.0 * (sy < 1.0 ? 1.0 : sy))) + 4;
    if (rows > m.h) rows = m.h;
    if (rows > p->max_rows) p->max_rows = rows;
    m.tmp_off = static_cast<long long>(tmp);
    tmp += align_up(static_cast<size_t>(rows) * S * 3, 256);
  }
  size_t off = 0;
  p->meta_off = off; off += align_up(sizeof(ImgMeta) * batch, 256);
  p->bounds_off = off; off += align_up(sizeof(int) * 2ull * batch * 2 * S, 256);
  p->coef_off = off; off += align_up(sizeof(int) * static_cast<size_t>(batch) * 2 * S * p->ks, 256);
  p->tmp_off = off; off += tmp;
  p->total = off;
  return CLM_OK;
}

}  // namespace

extern "C" size_t clm_preprocess_workspace_bytes(const clm_image_desc* descs_host, int batch, int out_size) {
  if (!descs_host || batch <= 0 || out_size <= 0) return 0;
  Plan p;
  if (make_plan(descs_host, batch, out_size, &p) != CLM_OK) return 0;
  return p.total;
}

extern "C" int clm_preprocess_images(const clm_image_desc* descs_host, int batch, int out_size,
                                     const float* mean3_host, const float* std3_host, float* pixel_values,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  CLM_REQUIRE(batch >= 0 && out_size > 0 && out_size <= 1024, "clm_preprocess_images: bad batch/out_size");
  if (batch == 0) return CLM_OK;
  CLM_REQUIRE(descs_host && mean3_host && std3_host && pixel_values && workspace,
              "clm_preprocess_images: null argument");
  const int S = out_size;
  Plan p;
  int rc = make_plan(descs_host, batch, S, &p);
  if (rc != CLM_OK) return rc;
  CLM_REQUIRE(workspace_bytes >= p.total, "clm_preprocess_images: workspace of %zu bytes, need %zu",
              workspace_bytes, p.total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  ImgMeta* meta = reinterpret_cast<ImgMeta*>(ws + p.meta_off);
  int* bounds = reinterpret_cast<int*>(ws + p.bounds_off);
  int* coef = reinterpret_cast<int*>(ws + p.coef_off);
  uint8_t* tmp = ws + p.tmp_off;
  // pageable source: the runtime stages it before returning, so `p` may die when this function returns
  CLM_CUDA_CHECK(cudaMemcpyAsync(meta, p.meta.data(), sizeof(ImgMeta) * batch, cudaMemcpyHostToDevice, s));
  double in_bytes = 0;
  for (int b = 0; b < batch; ++b) in_bytes += 3.0 * p.meta[b].h * p.meta[b].w;
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 8.0 * batch * 2 * S * (p.ks + 2), s);
    coef_kernel<<<dim3((S + 127) / 128, 2, batch), 128, 0, s>>>(meta, S, p.ks, bounds, coef);
  }
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, in_bytes + 3.0 * batch * p.max_rows * S, s);
    hpass_kernel<<<dim3((p.max_rows * S + 255) / 256, batch), 256, 0, s>>>(meta, S, p.ks, bounds, coef, tmp);
  }
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 3.0 * batch * p.max_rows * S + 12.0 * batch * S * S, s);
    vpass_kernel<<<dim3((S * S + 255) / 256, batch), 256, 0, s>>>(
        meta, S, p.ks, bounds, coef, tmp, make_float3(mean3_host[0], mean3_host[1], mean3_host[2]),
        make_float3(std3_host[0], std3_host[1], std3_host[2]), pixel_values);
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}
