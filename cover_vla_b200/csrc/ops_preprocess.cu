// Policy-side observation pre-processing on the device: replaces, per decision,
// BridgeSimplerAdapter.preprocess (INT-ACT/src/experiments/env_adapters/simpler.py:43-65): cv2.resize(frame, (224, 224),
// interpolation=cv2.INTER_LANCZOS4) on the uint8 HWC simulator frame followed by process_images (src/utils/pipeline.py:
// 34-69: x * (1/255), (x - 0.5) / 0.5 in float32).  One uint8 H2D copy per step instead of a host resize.
//
// The 8-bit INTER_LANCZOS4 path of cv::resize is integer arithmetic: per destination column / row 8 taps with int16
// coefficients (cv::interpolateLanczos4 in float / double, quantised to 11 bits with cvRound), replicated borders, an
// exact int32 horizontal pass, an exact int32 vertical pass, then (v + 2^21) >> 22 with saturation.  The coefficient
// tables depend only on the sizes: they are computed on the host exactly as OpenCV does (same libm) and cached on the
// device; the kernel evaluates the 8 x 8 taps of one output pixel per thread (all three channels).  Bit-exact against
// cv2.resize (tests/test_preprocess.py).
#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "host_common.h"
#include "ops.h"
#include "ptx.cuh"

namespace cvb {

namespace {

// cv::interpolateLanczos4 (imgproc: float x, double trigonometry, float accumulation)
void lanczos4_coeffs(float x, float* coeffs) {
  static const double s45 = 0.70710678118654752440084436210485;
  static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  const double pi = 3.1415926535897932384626433832795;
  float sum = 0;
  const double y0 = -(x + 3) * pi * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
  for (int i = 0; i < 8; i++) {
    const float y0_ = (x + 3 - i);
    if (std::fabs(y0_) >= 1e-6f) {
      const double y = -y0_ * pi * 0.25;
      coeffs[i] = static_cast<float>((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    } else {
      coeffs[i] = 1e30f;
    }
    sum += coeffs[i];
  }
  sum = 1.f / sum;
  for (int i = 0; i < 8; i++) coeffs[i] *= sum;
}

void lanczos4_tables(int src, int dst, std::vector<int>* ofs, std::vector<short>* coef) {
  const double scale = 1.0 / (static_cast<double>(dst) / src);
  ofs->resize(dst);
  coef->resize(static_cast<size_t>(dst) * 8);
  for (int d = 0; d < dst; ++d) {
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    const int s = static_cast<int>(std::floor(f));
    f -= s;
    (*ofs)[d] = s;
    float c[8];
    lanczos4_coeffs(f, c);
    for (int k = 0; k < 8; ++k) {
      const long r = std::lrintf(c[k] * 2048.0f);  // cvRound: nearest-even; saturate_cast<short>
      (*coef)[static_cast<size_t>(d) * 8 + k] = static_cast<short>(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
    }
  }
}

struct Tables {
  int* xofs = nullptr;
  int* yofs = nullptr;
  short* xcoef = nullptr;
  short* ycoef = nullptr;
};
std::mutex g_mu;
std::map<std::tuple<int, int, int, int, int>, Tables> g_tables;  // (device, H, W, dh, dw)

__global__ void __launch_bounds__(128) lanczos4_policy_image_kernel(const uint8_t* __restrict__ img, int H, int W, int dh,
                                                                    int dw, const int* __restrict__ xofs,
                                                                    const short* __restrict__ xcoef,
                                                                    const int* __restrict__ yofs,
                                                                    const short* __restrict__ ycoef,
                                                                    uint8_t* __restrict__ out_u8, float* __restrict__ out_f32) {
  pdl_wait();
  pdl_launch();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dh * dw) return;
  const int dy = idx / dw, dx = idx % dw;
  int xi[8], xa[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    xi[k] = min(max(xofs[dx] + k - 3, 0), W - 1) * 3;
    xa[k] = xcoef[dx * 8 + k];
  }
  unsigned acc[3] = {0u, 0u, 0u};  // int32 arithmetic with wrap-around, like the int accumulators of cv::resize
#pragma unroll
  for (int ky = 0; ky < 8; ++ky) {
    const int sy = min(max(yofs[dy] + ky - 3, 0), H - 1);
    const uint8_t* row = img + static_cast<long>(sy) * W * 3;
    unsigned h[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int c = 0; c < 3; ++c) h[c] += static_cast<unsigned>(static_cast<int>(row[xi[k] + c]) * xa[k]);
    }
    const int b = ycoef[dy * 8 + ky];
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += static_cast<unsigned>(static_cast<int>(h[c]) * b);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int v = (static_cast<int>(acc[c]) + (1 << 21)) >> 22;  // FixedPtCast<int, uchar, 22>
    const int u = min(max(v, 0), 255);
    if (out_u8 != nullptr) out_u8[(static_cast<long>(dy) * dw + dx) * 3 + c] = static_cast<uint8_t>(u);
    if (out_f32 != nullptr) {
      // pipeline.py: image * (1 / 255.0) -> (image - 0.5) / 0.5, float32, no contraction
      const float r = __fmul_rn(static_cast<float>(u), static_cast<float>(1 / 255.0));
      out_f32[(static_cast<long>(c) * dh + dy) * dw + dx] = __fdiv_rn(__fsub_rn(r, 0.5f), 0.5f);
    }
  }
}

}  // namespace

int preprocess_policy_image(cudaStream_t st, const uint8_t* img_hwc, int H, int W, int dh, int dw, uint8_t* out_u8_hwc,
                            float* out_f32_chw) {
  CVB_REQUIRE(img_hwc != nullptr && (out_u8_hwc != nullptr || out_f32_chw != nullptr), "null argument");
  CVB_REQUIRE(H >= 1 && W >= 1 && dh >= 1 && dw >= 1 && H <= 16384 && W <= 16384, "image sizes out of range");
  int dev = 0;
  CVB_CUDA(cudaGetDevice(&dev));
  Tables t;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, H, W, dh, dw);
    auto it = g_tables.find(key);
    if (it == g_tables.end()) {
      std::vector<int> xo, yo;
      std::vector<short> xc, yc;
      lanczos4_tables(W, dw, &xo, &xc);
      lanczos4_tables(H, dh, &yo, &yc);
      CVB_CUDA(cudaMalloc(&t.xofs, xo.size() * sizeof(int)));
      CVB_CUDA(cudaMalloc(&t.yofs, yo.size() * sizeof(int)));
      CVB_CUDA(cudaMalloc(&t.xcoef, xc.size() * sizeof(short)));
      CVB_CUDA(cudaMalloc(&t.ycoef, yc.size() * sizeof(short)));
      // synchronous copies from stack-lifetime host vectors (once per image geometry)
      CVB_CUDA(cudaMemcpy(t.xofs, xo.data(), xo.size() * sizeof(int), cudaMemcpyHostToDevice));
      CVB_CUDA(cudaMemcpy(t.yofs, yo.data(), yo.size() * sizeof(int), cudaMemcpyHostToDevice));
      CVB_CUDA(cudaMemcpy(t.xcoef, xc.data(), xc.size() * sizeof(short), cudaMemcpyHostToDevice));
      CVB_CUDA(cudaMemcpy(t.ycoef, yc.data(), yc.size() * sizeof(short), cudaMemcpyHostToDevice));
      g_tables.emplace(key, t);
    } else {
      t = it->second;
    }
  }
  const int n = dh * dw;
  CVB_TRY(launch_pdl(lanczos4_policy_image_kernel, dim3((n + 127) / 128), dim3(128), 0, st, 1, img_hwc, H, W, dh, dw, t.xofs,
                     t.xcoef, t.yofs, t.ycoef, out_u8_hwc, out_f32_chw));
  CVB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Verifier-side image transform on the device: open_clip's SigLIP transform as the reference applies it
// (bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:249-254 -> self.preprocess = open_clip image transform:
// PIL Image.resize((384, 384), BICUBIC), ToTensor (x / 255), Normalize(mean 0.5, std 0.5)).  Pillow's 8-bit resampler
// (src/libImaging/Resample.c; Pillow is unpinned in the reference, 12.2.0 in this image) is integer arithmetic: per
// output index a window [xmin, xmin + n) of normalised double-precision bicubic weights (a = -0.5, support 2 x the
// down-scale factor) quantised to 22 bits, a horizontal pass to an 8-bit intermediate (rounded, clipped), then the
// vertical pass.  Tables are built on the host as Pillow does; the kernel evaluates both passes for one output pixel
// per thread.  Bit-exact against PIL (tests/test_preprocess.py).
namespace {

constexpr int kPilPrecision = 22;

double pil_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc of Resample.c; returns ksize
int pil_tables(int in_size, int out_size, std::vector<int>* bounds, std::vector<int>* kk) {
  double scale, filterscale;
  filterscale = scale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  bounds->assign(static_cast<size_t>(out_size) * 2, 0);
  kk->assign(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < ksize; ++x) k[x] = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double w = pil_bicubic((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < ksize; ++x) {
      const double v = k[x] * (1 << kPilPrecision);
      (*kk)[static_cast<size_t>(xx) * ksize + x] = k[x] < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
    }
    (*bounds)[xx * 2] = xmin;
    (*bounds)[xx * 2 + 1] = xmax;
  }
  return ksize;
}

struct PilTables {
  int *xb = nullptr, *yb = nullptr, *xk = nullptr, *yk = nullptr;
  int xks = 0, yks = 0;
};
std::map<std::tuple<int, int, int, int, int>, PilTables> g_pil_tables;

__device__ __forceinline__ int pil_clip8(unsigned acc) {
  const int v = static_cast<int>(acc) >> kPilPrecision;
  return min(max(v, 0), 255);
}

__global__ void __launch_bounds__(128) pil_bicubic_image_kernel(const uint8_t* __restrict__ img, int H, int W, int dh, int dw,
                                                                const int* __restrict__ xb, const int* __restrict__ xk,
                                                                int xks, const int* __restrict__ yb,
                                                                const int* __restrict__ yk, int yks,
                                                                uint8_t* __restrict__ out_u8, float* __restrict__ out_f32) {
  pdl_wait();
  pdl_launch();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dh * dw) return;
  const int yy = idx / dw, xx = idx % dw;
  const int xmin = xb[xx * 2], xn = xb[xx * 2 + 1];
  const int ymin = yb[yy * 2], yn = yb[yy * 2 + 1];
  const unsigned half = 1u << (kPilPrecision - 1);
  unsigned acc[3] = {half, half, half};
  for (int y = 0; y < yn; ++y) {
    const uint8_t* row = img + (static_cast<long>(ymin + y) * W + xmin) * 3;
    unsigned h[3] = {half, half, half};
    for (int x = 0; x < xn; ++x) {
      const int kx = xk[xx * xks + x];
#pragma unroll
      for (int c = 0; c < 3; ++c) h[c] += static_cast<unsigned>(static_cast<int>(row[x * 3 + c]) * kx);
    }
    const int ky = yk[yy * yks + y];
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += static_cast<unsigned>(pil_clip8(h[c]) * ky);  // 8-bit intermediate image
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int u = pil_clip8(acc[c]);
    if (out_u8 != nullptr) out_u8[(static_cast<long>(yy) * dw + xx) * 3 + c] = static_cast<uint8_t>(u);
    if (out_f32 != nullptr)  // ToTensor: x / 255; Normalize: (x - 0.5) / 0.5 (float32)
      out_f32[(static_cast<long>(c) * dh + yy) * dw + xx] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(u), 255.0f), 0.5f), 0.5f);
  }
}

}  // namespace

int preprocess_verifier_image(cudaStream_t st, const uint8_t* img_hwc, int H, int W, int dh, int dw, uint8_t* out_u8_hwc,
                              float* out_f32_chw) {
  CVB_REQUIRE(img_hwc != nullptr && (out_u8_hwc != nullptr || out_f32_chw != nullptr), "null argument");
  CVB_REQUIRE(H >= 1 && W >= 1 && dh >= 1 && dw >= 1 && H <= 16384 && W <= 16384, "image sizes out of range");
  int dev = 0;
  CVB_CUDA(cudaGetDevice(&dev));
  PilTables t;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, H, W, dh, dw);
    auto it = g_pil_tables.find(key);
    if (it == g_pil_tables.end()) {
      std::vector<int> xb, yb, xk, yk;
      t.xks = pil_tables(W, dw, &xb, &xk);
      t.yks = pil_tables(H, dh, &yb, &yk);
      CVB_CUDA(cudaMalloc(&t.xb, xb.size() * sizeof(int)));
      CVB_CUDA(cudaMalloc(&t.yb, yb.size() * sizeof(int)));
      CVB_CUDA(cudaMalloc(&t.xk, xk.size() * sizeof(int)));
      CVB_CUDA(cudaMalloc(&t.yk, yk.size() * sizeof(int)));
      CVB_CUDA(cudaMemcpy(t.xb, xb.data(), xb.size() * sizeof(int), cudaMemcpyHostToDevice));
      CVB_CUDA(cudaMemcpy(t.yb, yb.data(), yb.size() * sizeof(int), cudaMemcpyHostToDevice));
      CVB_CUDA(cudaMemcpy(t.xk, xk.data(), xk.size() * sizeof(int), cudaMemcpyHostToDevice));
      CVB_CUDA(cudaMemcpy(t.yk, yk.data(), yk.size() * sizeof(int), cudaMemcpyHostToDevice));
      g_pil_tables.emplace(key, t);
    } else {
      t = it->second;
    }
  }
  const int n = dh * dw;
  CVB_TRY(launch_pdl(pil_bicubic_image_kernel, dim3((n + 127) / 128), dim3(128), 0, st, 1, img_hwc, H, W, dh, dw, t.xb, t.xk,
                     t.xks, t.yb, t.yk, t.yks, out_u8_hwc, out_f32_chw));
  CVB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Verifier-side frame reduction on the device: process_raw_image_to_jpg (CoVer_VLA/inference/experiments/robot/simpler/
// eval_utils.py:228-286) = tf.image.resize(frame_u8, (256, 256), BILINEAR, antialias=True) -> tf.cast(uint8).
// TensorFlow is a third-party dependency absent from /root/reference and from this image ("tensorflow" unpinned in
// CoVer_VLA/inference/pyproject.toml), so its published algorithm is restated (tensorflow/core/kernels/image/
// scale_and_translate_op.cc: ComputeSpansCore with the triangle kernel, then GatherRows, then GatherColumns):
//   scale = float(out) / float(in); inv_scale = 1 / scale; kernel_scale = max(inv_scale, 1)      (antialias)
//   sample = (x + 0.5) * inv_scale;  span = [ceil(sample - kernel_scale - 0.5), floor(sample + kernel_scale - 0.5)]
//   clamped to the image; w(src) = max(0, 1 - |src + 0.5 - sample| / kernel_scale), normalised by their float32 sum;
//   rows first into a float32 intermediate [out_h, W, 3], then columns; sequential float32 multiply-adds in source
//   order WITHOUT contraction; the cast to uint8 truncates.
// Parity: bit-exact against oracle/preprocess_oracle.tf_resize_bilinear_antialias_u8 (the same restatement in numpy
// float32); UNPINNED against TensorFlow itself (no TensorFlow here) - see DESIGN.md section 4.
namespace {

struct TfTables {
  int* starts = nullptr;   // [out]
  float* weights = nullptr;  // [out][span]
  int span = 0;
};
std::map<std::tuple<int, int, int>, TfTables> g_tf_tables;  // (device, in, out)

void tf_spans(int in_size, int out_size, std::vector<int>* starts, std::vector<float>* weights, int* span_out) {
  const float scale = static_cast<float>(out_size) / static_cast<float>(in_size);
  const float inv_scale = 1.0f / scale;
  const float kernel_scale = std::max(inv_scale, 1.0f);
  const float radius = 1.0f;  // triangle kernel
  const int span = std::min(2 * static_cast<int>(std::ceil(radius * kernel_scale)) + 1, in_size);
  starts->assign(out_size, 0);
  weights->assign(static_cast<size_t>(out_size) * span, 0.0f);
  for (int x = 0; x < out_size; ++x) {
    const float col_f = x + 0.5f;
    const float sample_f = col_f * inv_scale;
    if (sample_f < 0 || sample_f > in_size) continue;
    long span_start = static_cast<long>(std::ceil(sample_f - radius * kernel_scale - 0.5f));
    long span_end = static_cast<long>(std::floor(sample_f + radius * kernel_scale - 0.5f));
    span_start = std::min<long>(std::max<long>(span_start, 0), in_size - 1);
    span_end = std::min<long>(std::max<long>(span_end, 0), in_size - 1) + 1;
    const int n = static_cast<int>(span_end - span_start);
    float total = 0.0f;
    std::vector<float> tmp(n);
    for (int i = 0; i < n; ++i) {
      const float kernel_pos = static_cast<float>(span_start + i) + 0.5f - sample_f;
      const float v = std::fabs(kernel_pos / kernel_scale);
      const float wgt = std::max(0.0f, 1.0f - v);
      total += wgt;
      tmp[i] = wgt;
    }
    (*starts)[x] = static_cast<int>(span_start);
    if (std::fabs(total) >= 1000.0f * std::numeric_limits<float>::min()) {
      const float one_over = 1.0f / total;
      for (int i = 0; i < n && i < span; ++i) (*weights)[static_cast<size_t>(x) * span + i] = tmp[i] * one_over;
    }
  }
  *span_out = span;
}

// rows: inter[y][x][c] = sum_i float(img[ys + i][x][c]) * wy[y][i]   (sequential, no contraction)
__global__ void __launch_bounds__(256) tf_gather_rows_kernel(const uint8_t* __restrict__ img, int H, int W, int dh,
                                                             const int* __restrict__ ystart, const float* __restrict__ wy,
                                                             int span, float* __restrict__ inter) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over dh * W * 3
  if (i >= dh * W * 3) return;
  const int y = i / (W * 3), rem = i % (W * 3);
  const int ys = ystart[y];
  float acc = 0.0f;
  for (int k = 0; k < span; ++k) {
    const int sy = ys + k;
    if (sy >= H) break;
    const float w = wy[y * span + k];
    acc = __fadd_rn(acc, __fmul_rn(static_cast<float>(img[static_cast<long>(sy) * W * 3 + rem]), w));
  }
  inter[i] = acc;
}
// columns: out[y][x][c] = uint8(trunc(sum_i inter[y][xs + i][c] * wx[x][i]))
__global__ void __launch_bounds__(256) tf_gather_cols_kernel(const float* __restrict__ inter, int W, int dh, int dw,
                                                             const int* __restrict__ xstart, const float* __restrict__ wx,
                                                             int span, uint8_t* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over dh * dw * 3
  if (i >= dh * dw * 3) return;
  const int c = i % 3, x = (i / 3) % dw, y = i / (3 * dw);
  const int xs = xstart[x];
  float acc = 0.0f;
  for (int k = 0; k < span; ++k) {
    const int sx = xs + k;
    if (sx >= W) break;
    acc = __fadd_rn(acc, __fmul_rn(inter[(static_cast<long>(y) * W + sx) * 3 + c], wx[x * span + k]));
  }
  // tf.cast(float32 -> uint8): truncation toward zero (values are in [0, 255] up to rounding)
  out[i] = static_cast<uint8_t>(static_cast<int>(fminf(fmaxf(acc, 0.0f), 255.0f)));
}

int tf_get_tables(int dev, int in_size, int out_size, TfTables* t) {
  auto key = std::make_tuple(dev, in_size, out_size);
  auto it = g_tf_tables.find(key);
  if (it != g_tf_tables.end()) {
    *t = it->second;
    return 0;
  }
  std::vector<int> st;
  std::vector<float> w;
  tf_spans(in_size, out_size, &st, &w, &t->span);
  CVB_CUDA(cudaMalloc(&t->starts, st.size() * sizeof(int)));
  CVB_CUDA(cudaMalloc(&t->weights, w.size() * sizeof(float)));
  CVB_CUDA(cudaMemcpy(t->starts, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice));
  CVB_CUDA(cudaMemcpy(t->weights, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  g_tf_tables.emplace(key, *t);
  return 0;
}

}  // namespace

int resize_bilinear_antialias_u8(cudaStream_t st, const uint8_t* img_hwc, int H, int W, int dh, int dw, float* scratch_f32,
                                 uint8_t* out_u8_hwc) {
  CVB_REQUIRE(img_hwc != nullptr && out_u8_hwc != nullptr && scratch_f32 != nullptr, "null argument");
  CVB_REQUIRE(H >= 1 && W >= 1 && dh >= 1 && dw >= 1 && H <= 16384 && W <= 16384, "image sizes out of range");
  int dev = 0;
  CVB_CUDA(cudaGetDevice(&dev));
  TfTables ty, tx;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    CVB_TRY(tf_get_tables(dev, H, dh, &ty));
    CVB_TRY(tf_get_tables(dev, W, dw, &tx));
  }
  const int n1 = dh * W * 3, n2 = dh * dw * 3;
  CVB_TRY(launch_pdl(tf_gather_rows_kernel, dim3((n1 + 255) / 256), dim3(256), 0, st, 1, img_hwc, H, W, dh, ty.starts,
                     ty.weights, ty.span, scratch_f32));
  CVB_LAUNCHED();
  CVB_TRY(launch_pdl(tf_gather_cols_kernel, dim3((n2 + 255) / 256), dim3(256), 0, st, 1, scratch_f32, W, dh, dw, tx.starts,
                     tx.weights, tx.span, out_u8_hwc));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
