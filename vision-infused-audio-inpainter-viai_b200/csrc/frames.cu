// Video-frame preprocessing of the loader (Data_loaders/audio_loader.py:199-246 sample_data_new) as one kernel per stream:
// decoded uint8 frames -> cv2.resize (INTER_LINEAR, uint8 fixed point) -> BGR->RGB -> fliplr -> (v - 127) / 128 -> crop ->
// float NHWC block that the ResNet stem reads without a layout pass.  HBM-bound byte work: one thread per output pixel,
// 4 source pixels per channel, coefficients recomputed per thread in exactly OpenCV's arithmetic (bit-exact output).
// (For an exact 2x down-scale in both directions OpenCV takes its INTER_AREA fast path, (a + b + c + d + 2) >> 2: the fixed-point
// bilinear formula below yields the same bytes there -- both coefficient pairs are exactly (1024, 1024) -- so it needs no special
// case; tests/test_loader_{cpu,gpu}.py pin 512 -> 256 and 448 -> 224 against cv2.)
#include "common.cuh"
#include <math.h>
using namespace viai;

namespace {
constexpr int THREADS = 256;

// OpenCV resize.cpp, INTER_LINEAR, 8-bit: coefficient pair in 11-bit fixed point.
//   f = (float)((d + 0.5) * scale - 0.5); s = floor(f); f -= s;  columns additionally reset (s, f) at the borders,
//   rows keep f and clip the two source rows instead (resizeGeneric_Invoker).
__device__ __forceinline__ void lin_coeff(int d, double scale, int n_src, bool is_col, int& s0, int& s1, int& a0, int& a1) {
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);       // no FMA contraction: OpenCV's host arithmetic
  int s = (int)floorf(f);
  f -= (float)s;
  if (is_col) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
    s0 = s;
    s1 = min(s + 1, n_src - 1);
  } else {
    s0 = min(max(s, 0), n_src - 1);
    s1 = min(max(s + 1, 0), n_src - 1);
  }
  a0 = __float2int_rn((1.f - f) * 2048.f);
  a1 = __float2int_rn(f * 2048.f);
}

// src: (n, sh, sw, cn) uint8.  out: (n, out_h, out_w, out_c) float; channel c of the source lands in c_off + (swap_rb ? cn-1-c : c).
__global__ void __launch_bounds__(THREADS)
frames_kernel(const uint8_t* __restrict__ src, int n, int sh, int sw, int cn, int swap_rb, int dh, int dw, int flip, int crop_row,
              int crop_col, int out_h, int out_w, int out_c, int c_off, float* __restrict__ out) {
  const double sy = (double)sh / dh, sx = (double)sw / dw;
  const bool same = (sh == dh && sw == dw);
  const int64_t total = (int64_t)n * out_h * out_w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % out_w);
    const int y = (int)((i / out_w) % out_h);
    const int f = (int)(i / ((int64_t)out_w * out_h));
    const int ry = crop_row + y;                                   // row of the resized image
    const int rc = crop_col + x;
    const int rx = flip ? (dw - 1 - rc) : rc;                      // np.fliplr happens before the crop
    const uint8_t* img = src + (int64_t)f * sh * sw * cn;
    float* o = out + i * out_c + c_off;
    if (same) {                                                    // cv2.resize to the same size is a copy
      for (int c = 0; c < cn; ++c)
        o[swap_rb ? cn - 1 - c : c] = ((float)img[((int64_t)ry * sw + rx) * cn + c] - 127.f) / 128.f;
      continue;
    }
    int y0, y1, b0, b1, x0, x1, a0, a1;
    lin_coeff(ry, sy, sh, false, y0, y1, b0, b1);
    lin_coeff(rx, sx, sw, true, x0, x1, a0, a1);
    for (int c = 0; c < cn; ++c) {
      const int p00 = img[((int64_t)y0 * sw + x0) * cn + c], p01 = img[((int64_t)y0 * sw + x1) * cn + c];
      const int p10 = img[((int64_t)y1 * sw + x0) * cn + c], p11 = img[((int64_t)y1 * sw + x1) * cn + c];
      const int r0 = p00 * a0 + p01 * a1, r1 = p10 * a0 + p11 * a1;                          // HResizeLinear (int)
      const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;          // VResizeLinear<uchar,int,short>
      o[swap_rb ? cn - 1 - c : c] = ((float)min(max(v, 0), 255) - 127.f) / 128.f;
    }
  }
}
}  // namespace

extern "C" int viai_frames_preprocess(const uint8_t* src, int n_frames, int src_h, int src_w, int src_c, int swap_rb, int resize_h,
                                      int resize_w, int flip, int crop_row, int crop_col, int out_h, int out_w, int out_c, int c_off,
                                      float* out, viai_stream_t stream) {
  VIAI_REQUIRE(src && out && n_frames > 0 && src_h > 0 && src_w > 0 && (src_c == 1 || src_c == 3) && resize_h > 0 && resize_w > 0,
               "frames_preprocess: bad arguments");
  VIAI_REQUIRE(crop_row >= 0 && crop_col >= 0 && out_h > 0 && out_w > 0 && crop_row + out_h <= resize_h && crop_col + out_w <= resize_w,
               "frames_preprocess: crop (%d,%d)+(%d,%d) outside the %dx%d resized frame", crop_row, crop_col, out_h, out_w, resize_h,
               resize_w);
  VIAI_REQUIRE(c_off >= 0 && c_off + src_c <= out_c, "frames_preprocess: channels [%d,%d) outside the %d-channel block", c_off,
               c_off + src_c, out_c);
  const int64_t total = (int64_t)n_frames * out_h * out_w;
  frames_kernel<<<(int)imin64(cdiv(total, THREADS), 16 * kNumSMs), THREADS, 0, STR(stream)>>>(
      src, n_frames, src_h, src_w, src_c, swap_rb, resize_h, resize_w, flip, crop_row, crop_col, out_h, out_w, out_c, c_off, out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
