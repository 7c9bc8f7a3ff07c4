// Bilinear (align_corners=True) resize into a channel slice, channel-slice copy, pooling and small elementwise ops (NHWC fp32).
#include "common.cuh"
using namespace viai;

namespace {

constexpr int THREADS = 256;

// Same float arithmetic as ATen's upsample_bilinear2d (align_corners=True): ratio = (in-1)/(out-1) in float,
// src = ratio*dst, i0 = (int)src, lambda1 = src - i0, i1 = i0 + (i0 < in-1).
__device__ __forceinline__ float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }
__device__ __forceinline__ void src_index(float scale, int dst, int in, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * (float)dst;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  l0 = 1.f - l1;
}

// One block per output row (n, y): the row interpolation is evaluated once per block, the column interpolation once per
// thread-item; no integer division in the item loop (the previous one-item-per-thread version spent most of its time in
// 64-bit div/mod).
template <int VEC>
__global__ void __launch_bounds__(THREADS)
bilinear_fwd_kernel(const float* __restrict__ in, int N, int Hin, int Win, int C, float* __restrict__ out, int Hout,
                    int Wout, int Ctot, int coff) {
  const int cvec = C / VEC;
  const float sh = ac_scale(Hin, Hout), sw = ac_scale(Win, Wout);
  for (int row = blockIdx.x; row < N * Hout; row += gridDim.x) {
    const int n = row / Hout, y = row - n * Hout;
    int y0, y1;
    float ly0, ly1;
    src_index(sh, y, Hin, y0, y1, ly0, ly1);
    const float* r0 = in + ((int64_t)n * Hin + y0) * Win * C;
    const float* r1 = in + ((int64_t)n * Hin + y1) * Win * C;
    float* orow = out + (int64_t)row * Wout * Ctot + coff;
    const int items = Wout * cvec;
    for (int i = threadIdx.x; i < items; i += THREADS) {
      const int x = i / cvec, cv = i - x * cvec;           // 32-bit, cvec is small
      int x0, x1;
      float lx0, lx1;
      src_index(sw, x, Win, x0, x1, lx0, lx1);
      const int o0 = x0 * C + cv * VEC, o1 = x1 * C + cv * VEC;
      float* o = orow + (int64_t)x * Ctot + cv * VEC;
      if (VEC == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(r0 + o0)), bb = __ldg(reinterpret_cast<const float4*>(r0 + o1));
        const float4 c = __ldg(reinterpret_cast<const float4*>(r1 + o0)), d = __ldg(reinterpret_cast<const float4*>(r1 + o1));
        float4 r;
        r.x = ly0 * (lx0 * a.x + lx1 * bb.x) + ly1 * (lx0 * c.x + lx1 * d.x);
        r.y = ly0 * (lx0 * a.y + lx1 * bb.y) + ly1 * (lx0 * c.y + lx1 * d.y);
        r.z = ly0 * (lx0 * a.z + lx1 * bb.z) + ly1 * (lx0 * c.z + lx1 * d.z);
        r.w = ly0 * (lx0 * a.w + lx1 * bb.w) + ly1 * (lx0 * c.w + lx1 * d.w);
        *reinterpret_cast<float4*>(o) = r;
      } else {
        const float a = __ldg(r0 + o0), bb = __ldg(r0 + o1), c = __ldg(r1 + o0), d = __ldg(r1 + o1);
        *o = ly0 * (lx0 * a + lx1 * bb) + ly1 * (lx0 * c + lx1 * d);
      }
    }
  }
}

// Gather-form backward (deterministic, no atomics): for every input pixel, visit the output pixels whose
// forward footprint touches it.  Candidate range is widened by one and each candidate re-evaluates the
// forward index computation, so float rounding cannot desynchronise the two directions.
__device__ __forceinline__ void cand_range(float scale, int i, int in, int out, int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out - 1; return; }
  float inv = 1.f / scale;
  lo = (int)floorf((float)(i - 1) * inv) - 1;
  hi = (int)ceilf((float)(i + 1) * inv) + 1;
  if (lo < 0) lo = 0;
  if (hi > out - 1) hi = out - 1;
}

// One block per input row (n, iy).  The candidate output rows and their weights are the same for the whole block and are
// evaluated once (shared memory); each thread-item evaluates its candidate columns once and then only loads and accumulates.
constexpr int BIL_MAXC = 16;       // candidate rows / columns kept; wider footprints fall back to on-the-fly evaluation
template <int VEC>
__global__ void __launch_bounds__(THREADS)
bilinear_bwd_kernel(const float* __restrict__ dout, int N, int Hin, int Win, int C, float* __restrict__ din, int Hout,
                    int Wout, int Ctot, int coff) {
  __shared__ int ys_s[BIL_MAXC];
  __shared__ float wy_s[BIL_MAXC];
  __shared__ int ny_s;
  const int cvec = C / VEC;
  const float sh = ac_scale(Hin, Hout), sw = ac_scale(Win, Wout);
  for (int row = blockIdx.x; row < N * Hin; row += gridDim.x) {
    const int n = row / Hin, iy = row - n * Hin;
    __syncthreads();
    if (threadIdx.x == 0) {
      int ylo, yhi, cnt = 0;
      cand_range(sh, iy, Hin, Hout, ylo, yhi);
      for (int y = ylo; y <= yhi; ++y) {
        int y0, y1; float ly0, ly1;
        src_index(sh, y, Hin, y0, y1, ly0, ly1);
        const float wy = (y0 == iy ? ly0 : 0.f) + (y1 == iy ? ly1 : 0.f);
        if (wy != 0.f && cnt < BIL_MAXC) { ys_s[cnt] = y; wy_s[cnt] = wy; ++cnt; }
        else if (wy != 0.f) cnt = BIL_MAXC + 1;          // too many: signal the generic path
      }
      ny_s = cnt;
    }
    __syncthreads();
    const int ny = ny_s;
    const int items = Win * cvec;
    for (int i = threadIdx.x; i < items; i += THREADS) {
      const int ix = i / cvec, cv = i - ix * cvec;
      float acc[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
      int xlo, xhi;
      cand_range(sw, ix, Win, Wout, xlo, xhi);
      if (ny <= BIL_MAXC) {
        for (int x = xlo; x <= xhi; ++x) {
          int x0, x1; float lx0, lx1;
          src_index(sw, x, Win, x0, x1, lx0, lx1);
          const float wx = (x0 == ix ? lx0 : 0.f) + (x1 == ix ? lx1 : 0.f);
          if (wx == 0.f) continue;
          const float* dcol = dout + ((int64_t)n * Hout * Wout + x) * Ctot + coff + cv * VEC;
          for (int j = 0; j < ny; ++j) {
            const float w = wy_s[j] * wx;
            const float* d = dcol + (int64_t)ys_s[j] * Wout * Ctot;
            if (VEC == 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(d));
              acc[0] += w * v.x; acc[1 % VEC] += w * v.y; acc[2 % VEC] += w * v.z; acc[3 % VEC] += w * v.w;
            } else {
              acc[0] += w * __ldg(d);
            }
          }
        }
      } else {                       // generic: evaluate every (row, column) candidate
        int ylo, yhi;
        cand_range(sh, iy, Hin, Hout, ylo, yhi);
        for (int y = ylo; y <= yhi; ++y) {
          int y0, y1; float ly0, ly1;
          src_index(sh, y, Hin, y0, y1, ly0, ly1);
          const float wy = (y0 == iy ? ly0 : 0.f) + (y1 == iy ? ly1 : 0.f);
          if (wy == 0.f) continue;
          for (int x = xlo; x <= xhi; ++x) {
            int x0, x1; float lx0, lx1;
            src_index(sw, x, Win, x0, x1, lx0, lx1);
            const float wx = (x0 == ix ? lx0 : 0.f) + (x1 == ix ? lx1 : 0.f);
            if (wx == 0.f) continue;
            const float* d = dout + (((int64_t)n * Hout + y) * Wout + x) * Ctot + coff + cv * VEC;
            const float w = wy * wx;
            if (VEC == 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(d));
              acc[0] += w * v.x; acc[1 % VEC] += w * v.y; acc[2 % VEC] += w * v.z; acc[3 % VEC] += w * v.w;
            } else {
              acc[0] += w * __ldg(d);
            }
          }
        }
      }
      float* o = din + ((int64_t)row * Win + ix) * C + cv * VEC;
      if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
      else *o = acc[0];
    }
  }
}

__global__ void copy_channels_kernel(const float* __restrict__ src, int64_t rows, int Csrc, int soff, float* __restrict__ dst,
                                     int Cdst, int doff, int C) {
  const int64_t total = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / C;
    int c = (int)(i - row * C);
    dst[row * Cdst + doff + c] = __ldg(src + row * Csrc + soff + c);
  }
}

__global__ void avgpool_h_fwd_kernel(const float* __restrict__ in, int N, int H, int W, int C, int kh, float* __restrict__ out) {
  const int Ho = H / kh;
  const int64_t total = (int64_t)N * Ho * W * C;
  const float inv = 1.f / (float)kh;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t wc = i % ((int64_t)W * C);
    int64_t t = i / ((int64_t)W * C);
    int yo = (int)(t % Ho);
    int n = (int)(t / Ho);
    float s = 0.f;
    for (int k = 0; k < kh; ++k) s += __ldg(in + ((int64_t)n * H + yo * kh + k) * W * C + wc);
    out[i] = s * inv;
  }
}

__global__ void avgpool_h_bwd_kernel(const float* __restrict__ dout, int N, int H, int W, int C, int kh, float* __restrict__ din) {
  const int Ho = H / kh;
  const int64_t total = (int64_t)N * H * W * C;
  const float inv = 1.f / (float)kh;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t wc = i % ((int64_t)W * C);
    int64_t t = i / ((int64_t)W * C);
    int y = (int)(t % H);
    int n = (int)(t / H);
    int yo = y / kh;
    din[i] = yo < Ho ? __ldg(dout + ((int64_t)n * Ho + yo) * W * C + wc) * inv : 0.f;
  }
}

__global__ void maxpool_fwd_kernel(const float* __restrict__ in, int N, int H, int W, int C, float* __restrict__ out, int Ho, int Wo) {
  const int64_t total = (int64_t)N * Ho * Wo * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int xo = (int)(t % Wo); t /= Wo;
    int yo = (int)(t % Ho);
    int n = (int)(t / Ho);
    float m = -INFINITY;
    for (int r = 0; r < 3; ++r) {
      int y = yo * 2 - 1 + r;
      if (y < 0 || y >= H) continue;
      for (int s = 0; s < 3; ++s) {
        int x = xo * 2 - 1 + s;
        if (x < 0 || x >= W) continue;
        m = fmaxf(m, __ldg(in + (((int64_t)n * H + y) * W + x) * C + c));
      }
    }
    out[i] = m;
  }
}

// Gather-form backward: an input pixel collects the gradient of every window in which it is the FIRST maximum
// (row-major scan order, which is ATen's tie-breaking rule).
__global__ void maxpool_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout, int N, int H, int W, int C,
                                   float* __restrict__ din, int Ho, int Wo) {
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int x = (int)(t % W); t /= W;
    int y = (int)(t % H);
    int n = (int)(t / H);
    const float v = __ldg(in + i);
    float acc = 0.f;
    for (int yo = (y + 1) / 2 - 1; yo <= (y + 1) / 2; ++yo) {
      if (yo < 0 || yo >= Ho || yo * 2 - 1 > y || yo * 2 + 1 < y) continue;
      for (int xo = (x + 1) / 2 - 1; xo <= (x + 1) / 2; ++xo) {
        if (xo < 0 || xo >= Wo || xo * 2 - 1 > x || xo * 2 + 1 < x) continue;
        // is (y,x) the first max of window (yo,xo)?
        bool first = true;
        for (int r = 0; r < 3 && first; ++r) {
          int yy = yo * 2 - 1 + r;
          if (yy < 0 || yy >= H) continue;
          for (int s = 0; s < 3; ++s) {
            int xx = xo * 2 - 1 + s;
            if (xx < 0 || xx >= W) continue;
            float u = __ldg(in + (((int64_t)n * H + yy) * W + xx) * C + c);
            bool before = (yy < y) || (yy == y && xx < x);
            if (u > v || (before && u == v)) { first = false; break; }
          }
        }
        if (first) acc += __ldg(dout + (((int64_t)n * Ho + yo) * Wo + xo) * C + c);
      }
    }
    din[i] = acc;
  }
}

// Max-pool backward that reads the forward OUTPUT: an input element can only receive the gradient of a window whose maximum it
// equals, so almost every (element, window) pair is rejected after one load; the tie scan (first maximum in row-major order,
// ATen's rule) runs only for the elements that do equal the maximum.  4 channels per thread (C % 4 == 0).
__global__ void __launch_bounds__(256)
maxpool_bwd_out_kernel(const float4* __restrict__ in, const float4* __restrict__ out, const float4* __restrict__ dout, int N, int H,
                       int W, int C4, float4* __restrict__ din, int Ho, int Wo) {
  const int64_t total = (int64_t)N * H * W * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    int64_t t = i / C4;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    const float4 v4 = __ldg(in + i);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int yo = (y + 1) / 2 - 1; yo <= (y + 1) / 2; ++yo) {
      if (yo < 0 || yo >= Ho || yo * 2 - 1 > y || yo * 2 + 1 < y) continue;
      for (int xo = (x + 1) / 2 - 1; xo <= (x + 1) / 2; ++xo) {
        if (xo < 0 || xo >= Wo || xo * 2 - 1 > x || xo * 2 + 1 < x) continue;
        const int64_t o = (((int64_t)n * Ho + yo) * Wo + xo) * C4 + c;
        const float4 m4 = __ldg(out + o);
        const float m[4] = {m4.x, m4.y, m4.z, m4.w};
        bool hit[4];
        bool any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) { hit[k] = (v[k] == m[k]); any |= hit[k]; }
        if (!any) continue;
        // first maximum only: no element EARLIER in the window's scan order may equal the maximum as well
        for (int r = 0; r < 3; ++r) {
          const int yy = yo * 2 - 1 + r;
          if (yy < 0 || yy >= H || yy > y) continue;
          for (int q = 0; q < 3; ++q) {
            const int xx = xo * 2 - 1 + q;
            if (xx < 0 || xx >= W) continue;
            if (!(yy < y || xx < x)) continue;
            const float4 u4 = __ldg(in + (((int64_t)n * H + yy) * W + xx) * C4 + c);
            const float u[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (u[k] == m[k]) hit[k] = false;
          }
        }
        const float4 g4 = __ldg(dout + o);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (hit[k]) acc[k] += g[k];
      }
    }
    din[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a[i] * b[i];
}
__global__ void add_act_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = act_fwd(a[i] + b[i], act, 0.f);
}
__global__ void add_act_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dout, float* __restrict__ din, int64_t n, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    din[i] = (act == VIAI_ACT_RELU && !(out[i] > 0.f)) ? 0.f : dout[i];
}

inline int grid_for(int64_t total) { return (int)imin64(cdiv(total, THREADS), 16 * kNumSMs); }

}  // namespace

extern "C" int viai_bilinear_fwd(const float* in, int N, int Hin, int Win, int C, float* out, int Hout, int Wout,
                                 int Ctot, int coff, viai_stream_t stream) {
  VIAI_REQUIRE(in && out && N > 0 && Hin > 0 && Win > 0 && C > 0 && Hout > 0 && Wout > 0 && coff >= 0 && coff + C <= Ctot,
               "viai_bilinear_fwd: bad arguments");
  bool v4 = C % 4 == 0 && Ctot % 4 == 0 && coff % 4 == 0 &&
            (reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0;
  const int rows = (int)imin64((int64_t)N * Hout, 32 * kNumSMs);
  if (v4) bilinear_fwd_kernel<4><<<rows, THREADS, 0, STR(stream)>>>(in, N, Hin, Win, C, out, Hout, Wout, Ctot, coff);
  else bilinear_fwd_kernel<1><<<rows, THREADS, 0, STR(stream)>>>(in, N, Hin, Win, C, out, Hout, Wout, Ctot, coff);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_bilinear_bwd(const float* dout, int N, int Hin, int Win, int C, float* din, int Hout, int Wout,
                                 int Ctot, int coff, viai_stream_t stream) {
  VIAI_REQUIRE(dout && din && N > 0 && Hin > 0 && Win > 0 && C > 0 && Hout > 0 && Wout > 0 && coff >= 0 && coff + C <= Ctot,
               "viai_bilinear_bwd: bad arguments");
  bool v4 = C % 4 == 0 && Ctot % 4 == 0 && coff % 4 == 0 &&
            (reinterpret_cast<uintptr_t>(din) | reinterpret_cast<uintptr_t>(dout)) % 16 == 0;
  const int rows = (int)imin64((int64_t)N * Hin, 32 * kNumSMs);
  if (v4) bilinear_bwd_kernel<4><<<rows, THREADS, 0, STR(stream)>>>(dout, N, Hin, Win, C, din, Hout, Wout, Ctot, coff);
  else bilinear_bwd_kernel<1><<<rows, THREADS, 0, STR(stream)>>>(dout, N, Hin, Win, C, din, Hout, Wout, Ctot, coff);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_copy_channels(const float* src, int64_t rows, int Csrc, int soff, float* dst, int Cdst, int doff,
                                  int C, viai_stream_t stream) {
  VIAI_REQUIRE(src && dst && rows > 0 && C > 0 && soff >= 0 && doff >= 0 && soff + C <= Csrc && doff + C <= Cdst,
               "viai_copy_channels: bad arguments");
  copy_channels_kernel<<<grid_for(rows * C), THREADS, 0, STR(stream)>>>(src, rows, Csrc, soff, dst, Cdst, doff, C);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_avgpool_h_fwd(const float* in, int N, int H, int W, int C, int kh, float* out, viai_stream_t stream) {
  VIAI_REQUIRE(in && out && N > 0 && W > 0 && C > 0 && kh > 0 && H >= kh, "viai_avgpool_h_fwd: bad arguments (H=%d, kh=%d)", H, kh);
  avgpool_h_fwd_kernel<<<grid_for((int64_t)N * (H / kh) * W * C), THREADS, 0, STR(stream)>>>(in, N, H, W, C, kh, out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
extern "C" int viai_avgpool_h_bwd(const float* dout, int N, int H, int W, int C, int kh, float* din, viai_stream_t stream) {
  VIAI_REQUIRE(dout && din && N > 0 && W > 0 && C > 0 && kh > 0 && H >= kh, "viai_avgpool_h_bwd: bad arguments");
  avgpool_h_bwd_kernel<<<grid_for((int64_t)N * H * W * C), THREADS, 0, STR(stream)>>>(dout, N, H, W, C, kh, din);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
// Max-pool with a saved argmax: the forward pass stores, per output element, which of the window's 9 positions holds the FIRST
// maximum (row-major scan, strict >: ATen's tie-breaking rule) as one byte; the backward pass is then a pure gather -- an input
// element checks the <= 4 windows that contain it and takes dout where the stored position is its own -- that reads neither the
// forward input nor the forward output: 1 byte + (on a match) 4 bytes per (element, window) instead of re-scanning windows.
// Four channels per thread (C % 4 == 0).  Traffic at the ResNet stem (N x 112 x 112 x 64 -> 56 x 56): forward 1.03 N MB, backward
// 1.08 N MB (was ~2.0 N MB plus the tie scans).
__global__ void __launch_bounds__(256)
maxpool_fwd_idx_kernel(const float4* __restrict__ in, int N, int H, int W, int C4, float4* __restrict__ out, uchar4* __restrict__ idx,
                       int Ho, int Wo) {
  const int64_t total = (int64_t)N * Ho * Wo * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    int64_t t = i / C4;
    const int xo = (int)(t % Wo); t /= Wo;
    const int yo = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int a[4] = {-1, -1, -1, -1};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int y = yo * 2 - 1 + r;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int x = xo * 2 - 1 + q;
        if (x < 0 || x >= W) continue;
        const float4 u4 = __ldg(in + (((int64_t)n * H + y) * W + x) * C4 + c);
        const float u[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (a[k] < 0 || u[k] > m[k]) { m[k] = u[k]; a[k] = r * 3 + q; }
      }
    }
    out[i] = make_float4(m[0], m[1], m[2], m[3]);
    idx[i] = make_uchar4((unsigned char)a[0], (unsigned char)a[1], (unsigned char)a[2], (unsigned char)a[3]);
  }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_idx_kernel(const uchar4* __restrict__ idx, const float4* __restrict__ dout, int N, int H, int W, int C4,
                       float4* __restrict__ din, int Ho, int Wo) {
  const int64_t total = (int64_t)N * H * W * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    int64_t t = i / C4;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int yo = (y + 1) / 2 - 1 + dy;
      const int r = y - (yo * 2 - 1);
      if (yo < 0 || yo >= Ho || r < 0 || r > 2) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int xo = (x + 1) / 2 - 1 + dx;
        const int q = x - (xo * 2 - 1);
        if (xo < 0 || xo >= Wo || q < 0 || q > 2) continue;
        const int64_t o = (((int64_t)n * Ho + yo) * Wo + xo) * C4 + c;
        const uchar4 a = __ldg(idx + o);
        const unsigned char code = (unsigned char)(r * 3 + q);
        const bool h0 = a.x == code, h1 = a.y == code, h2 = a.z == code, h3 = a.w == code;
        if (h0 | h1 | h2 | h3) {
          const float4 g = __ldg(dout + o);
          acc[0] += h0 ? g.x : 0.f; acc[1] += h1 ? g.y : 0.f; acc[2] += h2 ? g.z : 0.f; acc[3] += h3 ? g.w : 0.f;
        }
      }
    }
    din[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

extern "C" int viai_maxpool3s2_fwd_idx(const float* in, int N, int H, int W, int C, float* out, unsigned char* idx, int Ho, int Wo,
                                       viai_stream_t stream) {
  VIAI_REQUIRE(in && out && idx && N > 0 && C > 0 && C % 4 == 0 && Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1,
               "viai_maxpool3s2_fwd_idx: bad arguments (C must be a multiple of 4)");
  VIAI_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 && (reinterpret_cast<uintptr_t>(idx) & 3) == 0,
               "viai_maxpool3s2_fwd_idx: pointers must be 16-byte (idx: 4-byte) aligned");
  maxpool_fwd_idx_kernel<<<grid_for((int64_t)N * Ho * Wo * (C / 4)), 256, 0, STR(stream)>>>(
      reinterpret_cast<const float4*>(in), N, H, W, C / 4, reinterpret_cast<float4*>(out), reinterpret_cast<uchar4*>(idx), Ho, Wo);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_maxpool3s2_bwd_idx(const unsigned char* idx, const float* dout, int N, int H, int W, int C, float* din, int Ho,
                                       int Wo, viai_stream_t stream) {
  VIAI_REQUIRE(idx && dout && din && N > 0 && C > 0 && C % 4 == 0 && Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1,
               "viai_maxpool3s2_bwd_idx: bad arguments (C must be a multiple of 4)");
  VIAI_REQUIRE(((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(din)) & 15) == 0 && (reinterpret_cast<uintptr_t>(idx) & 3) == 0,
               "viai_maxpool3s2_bwd_idx: pointers must be 16-byte (idx: 4-byte) aligned");
  maxpool_bwd_idx_kernel<<<grid_for((int64_t)N * H * W * (C / 4)), 256, 0, STR(stream)>>>(
      reinterpret_cast<const uchar4*>(idx), reinterpret_cast<const float4*>(dout), N, H, W, C / 4, reinterpret_cast<float4*>(din), Ho, Wo);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_maxpool3s2_fwd(const float* in, int N, int H, int W, int C, float* out, int Ho, int Wo, viai_stream_t stream) {
  VIAI_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0 && Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1,
               "viai_maxpool3s2_fwd: bad arguments");
  maxpool_fwd_kernel<<<grid_for((int64_t)N * Ho * Wo * C), THREADS, 0, STR(stream)>>>(in, N, H, W, C, out, Ho, Wo);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
extern "C" int viai_maxpool3s2_bwd(const float* in, const float* dout, int N, int H, int W, int C, float* din, int Ho,
                                   int Wo, viai_stream_t stream) {
  VIAI_REQUIRE(in && dout && din && N > 0 && H > 0 && W > 0 && C > 0 && Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1,
               "viai_maxpool3s2_bwd: bad arguments");
  maxpool_bwd_kernel<<<grid_for((int64_t)N * H * W * C), THREADS, 0, STR(stream)>>>(in, dout, N, H, W, C, din, Ho, Wo);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
extern "C" int viai_maxpool3s2_bwd_out(const float* in, const float* out, const float* dout, int N, int H, int W, int C, float* din,
                                       int Ho, int Wo, viai_stream_t stream) {
  VIAI_REQUIRE(in && out && dout && din && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && Ho == (H - 1) / 2 + 1 &&
                   Wo == (W - 1) / 2 + 1,
               "viai_maxpool3s2_bwd_out: bad arguments (C must be a multiple of 4)");
  VIAI_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) |
                 reinterpret_cast<uintptr_t>(din)) & 15) == 0,
               "viai_maxpool3s2_bwd_out: pointers must be 16-byte aligned");
  maxpool_bwd_out_kernel<<<grid_for((int64_t)N * H * W * (C / 4)), 256, 0, STR(stream)>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<const float4*>(out), reinterpret_cast<const float4*>(dout), N, H, W, C / 4,
      reinterpret_cast<float4*>(din), Ho, Wo);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
extern "C" int viai_mul(const float* a, const float* b, float* out, int64_t n, viai_stream_t stream) {
  VIAI_REQUIRE(a && b && out && n > 0, "viai_mul: bad arguments");
  mul_kernel<<<grid_for(n), THREADS, 0, STR(stream)>>>(a, b, out, n);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
extern "C" int viai_add_act(const float* a, const float* b, float* out, int64_t n, int act, viai_stream_t stream) {
  VIAI_REQUIRE(a && b && out && n > 0, "viai_add_act: bad arguments");
  add_act_kernel<<<grid_for(n), THREADS, 0, STR(stream)>>>(a, b, out, n, act);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
extern "C" int viai_add_act_bwd(const float* out, const float* dout, float* din, int64_t n, int act, viai_stream_t stream) {
  VIAI_REQUIRE(out && dout && din && n > 0, "viai_add_act_bwd: bad arguments");
  add_act_bwd_kernel<<<grid_for(n), THREADS, 0, STR(stream)>>>(out, dout, din, n, act);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
