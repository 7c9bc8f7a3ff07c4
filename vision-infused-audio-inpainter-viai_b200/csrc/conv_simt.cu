// CUDA-core fp32 implicit-GEMM convolution: forward / transposed gather and weight gradient.
// This is the exact-fp32 path of the library and the on-device validator for the tcgen05 path.
#include "common.cuh"
using namespace viai;

namespace {

__global__ void pack_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I, int R, int S,
                                   int64_t so, int64_t si, int64_t sr, int64_t ss, int flip) {
  int64_t total = (int64_t)O * R * S * I;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int i = idx % I;
    int64_t t = idx / I;
    int s = t % S; t /= S;
    int r = t % R;
    int o = t / R;
    int rr = flip ? R - 1 - r : r, sw = flip ? S - 1 - s : s;
    dst[idx] = src[o * so + i * si + rr * sr + sw * ss];
  }
}

// gather coordinate for one axis; returns false when the tap falls outside / between input samples
__device__ __forceinline__ bool gather_coord(int mode, int y, int r, int stride, int pad, int limit, int& Y) {
  if (mode == 0) {
    Y = y * stride - pad + r;
  } else {
    int t = y + pad - r;
    if (t < 0) return false;
    Y = t / stride;
    if (Y * stride != t) return false;
  }
  return Y >= 0 && Y < limit;
}

constexpr int BM = 128, BK = 16;

template <int BN, bool VEC>
__global__ void __launch_bounds__(256)
conv_gather_kernel(viai_conv_geom g, const float* __restrict__ in, const float* __restrict__ wp,
                   const float* __restrict__ bias, float* __restrict__ out) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t M = (int64_t)g.N * g.Hout * g.Wout;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = g.R * g.S * g.Cin;
  const int HWo = g.Hout * g.Wout;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // pixels this thread loads
  constexpr int NPIX = VEC ? 2 : 1;
  int pn[NPIX], py[NPIX], px[NPIX];
  bool pv[NPIX];
#pragma unroll
  for (int j = 0; j < NPIX; ++j) {
    int lm = VEC ? (tid >> 2) + 64 * j : (tid & 127);
    int64_t m = m0 + lm;
    pv[j] = m < M;
    int64_t mm = pv[j] ? m : 0;
    pn[j] = (int)(mm / HWo);
    int rem = (int)(mm - (int64_t)pn[j] * HWo);
    py[j] = rem / g.Wout;
    px[j] = rem - py[j] * g.Wout;
  }

  for (int k0 = 0; k0 < K; k0 += BK) {
    if (VEC) {
      const int kq = tid & 3;
      const int tap = k0 / g.Cin, ci0 = k0 - tap * g.Cin;
      const int r = tap / g.S, s = tap - r * g.S;
#pragma unroll
      for (int j = 0; j < NPIX; ++j) {
        int lm = (tid >> 2) + 64 * j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int Y, X;
        if (pv[j] && gather_coord(g.mode, py[j], r, g.stride_h, g.pad_h, g.Hin, Y) &&
            gather_coord(g.mode, px[j], s, g.stride_w, g.pad_w, g.Win, X)) {
          v = __ldg(reinterpret_cast<const float4*>(in + (((int64_t)pn[j] * g.Hin + Y) * g.Win + X) * g.Cin + ci0 + kq * 4));
        }
        As[kq * 4 + 0][lm] = v.x; As[kq * 4 + 1][lm] = v.y; As[kq * 4 + 2][lm] = v.z; As[kq * 4 + 3][lm] = v.w;
      }
      const int ln = tid >> 2;
      if (ln < BN) {
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + ln < g.Cout) w = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)(n0 + ln) * K + k0 + kq * 4));
        Bs[kq * 4 + 0][ln] = w.x; Bs[kq * 4 + 1][ln] = w.y; Bs[kq * 4 + 2][ln] = w.z; Bs[kq * 4 + 3][ln] = w.w;
      }
    } else {
      const int lm = tid & 127;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int kk = (tid >> 7) + 2 * e;
        int k = k0 + kk;
        float v = 0.f;
        if (k < K && pv[0]) {
          int tap = k / g.Cin, ci = k - tap * g.Cin;
          int r = tap / g.S, s = tap - r * g.S;
          int Y, X;
          if (gather_coord(g.mode, py[0], r, g.stride_h, g.pad_h, g.Hin, Y) &&
              gather_coord(g.mode, px[0], s, g.stride_w, g.pad_w, g.Win, X))
            v = __ldg(in + (((int64_t)pn[0] * g.Hin + Y) * g.Win + X) * g.Cin + ci);
        }
        As[kk][lm] = v;
      }
#pragma unroll
      for (int e = 0; e < (BN * BK) / 256; ++e) {
        int idx = tid + 256 * e;
        int kk = idx & 15, ln = idx >> 4;
        int k = k0 + kk;
        float w = 0.f;
        if (k < K && n0 + ln < g.Cout) w = __ldg(wp + (int64_t)(n0 + ln) * K + k);
        Bs[kk][ln] = w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < g.Cout) out[m * g.Cout + n] = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
    }
  }
}

template <int BA, int BB>
__global__ void __launch_bounds__(256)
wgrad_kernel(viai_conv_geom g, const float* __restrict__ U, const float* __restrict__ G, float* __restrict__ dw,
             int64_t sa, int64_t sb, int64_t sr, int64_t ss, int ksplit) {
  constexpr int TA = BA / 16, TB = BB / 16;
  __shared__ __align__(16) float Us[BK][BA + 4];
  __shared__ __align__(16) float Gs[BK][BB + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int A = g.Cout, B = g.Cin;
  const int btiles = (B + BB - 1) / BB;
  const int a0 = (blockIdx.x / btiles) * BA, b0 = (blockIdx.x % btiles) * BB;
  const int tap = blockIdx.y;
  const int r = tap / g.S, s = tap - r * g.S;
  const int64_t M = (int64_t)g.N * g.Hout * g.Wout;
  int64_t chunk = (M + ksplit - 1) / ksplit;
  chunk = (chunk + BK - 1) / BK * BK;
  const int64_t p_begin = (int64_t)blockIdx.z * chunk;
  const int64_t p_end = p_begin + chunk < M ? p_begin + chunk : M;
  const int HWo = g.Hout * g.Wout;

  float acc[TA][TB];
#pragma unroll
  for (int i = 0; i < TA; ++i)
#pragma unroll
    for (int j = 0; j < TB; ++j) acc[i][j] = 0.f;

  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
#pragma unroll
    for (int e = 0; e < (BA * BK) / 256; ++e) {
      int idx = tid + 256 * e;
      int al = idx % BA, kk = idx / BA;
      int64_t p = p0 + kk;
      float v = 0.f;
      if (p < p_end && a0 + al < A) v = __ldg(U + p * A + a0 + al);
      Us[kk][al] = v;
    }
#pragma unroll
    for (int e = 0; e < (BB * BK) / 256; ++e) {
      int idx = tid + 256 * e;
      int bl = idx % BB, kk = idx / BB;
      int64_t p = p0 + kk;
      float v = 0.f;
      if (p < p_end && b0 + bl < B) {
        int n = (int)(p / HWo);
        int rem = (int)(p - (int64_t)n * HWo);
        int y = rem / g.Wout, x = rem - y * g.Wout;
        int Y = y * g.stride_h - g.pad_h + r, X = x * g.stride_w - g.pad_w + s;
        if (Y >= 0 && Y < g.Hin && X >= 0 && X < g.Win)
          v = __ldg(G + (((int64_t)n * g.Hin + Y) * g.Win + X) * B + b0 + bl);
      }
      Gs[kk][bl] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TA], b[TB];
#pragma unroll
      for (int i = 0; i < TA; ++i) a[i] = Us[kk][ty * TA + i];
#pragma unroll
      for (int j = 0; j < TB; ++j) b[j] = Gs[kk][tx * TB + j];
#pragma unroll
      for (int i = 0; i < TA; ++i)
#pragma unroll
        for (int j = 0; j < TB; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TA; ++i) {
    int a = a0 + ty * TA + i;
    if (a >= A) continue;
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      int b = b0 + tx * TB + j;
      if (b < B) atomicAdd(dw + a * sa + b * sb + r * sr + s * ss, acc[i][j]);
    }
  }
}

// Single-channel weight gradient (A == B == 1, e.g. the WaveNet conditioning upsampler ConvTranspose2d(1,1,(3,s)),
// wavenet_vocoder/wavenet.py:153-166): dw[r][s] = sum_p U[p] * G[gather(p, r, s)].  Every thread keeps the taps in registers
// over a grid-stride range of pixels; one shuffle reduction + one atomic per tap and warp.
constexpr int C1_MAXTAP = 32;
__global__ void __launch_bounds__(256)
wgrad_c1_kernel(viai_conv_geom g, const float* __restrict__ U, const float* __restrict__ G, float* __restrict__ dw, int64_t sr,
                int64_t ss) {
  const int taps = g.R * g.S;
  const int64_t M = (int64_t)g.N * g.Hout * g.Wout;
  const int HWo = g.Hout * g.Wout;
  float acc[C1_MAXTAP];
#pragma unroll
  for (int t = 0; t < C1_MAXTAP; ++t) acc[t] = 0.f;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < M; p += (int64_t)gridDim.x * blockDim.x) {
    const float u = __ldg(U + p);
    const int n = (int)(p / HWo);
    const int rem = (int)(p - (int64_t)n * HWo);
    const int y = rem / g.Wout, x = rem - y * g.Wout;
    const float* base = G + (int64_t)n * g.Hin * g.Win;
#pragma unroll
    for (int t = 0; t < C1_MAXTAP; ++t) {
      if (t < taps) {
        const int r = t / g.S, s = t - r * g.S;
        const int Y = y * g.stride_h - g.pad_h + r, X = x * g.stride_w - g.pad_w + s;
        if (Y >= 0 && Y < g.Hin && X >= 0 && X < g.Win) acc[t] = fmaf(u, __ldg(base + (int64_t)Y * g.Win + X), acc[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < C1_MAXTAP; ++t) {
    if (t < taps) {
      float v = acc[t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(dw + (t / g.S) * sr + (t % g.S) * ss, v);
    }
  }
}

// im2col of a forward convolution, channel-major columns: out[p][c*R*S + r*S + s] = in[n, y*sh - ph + r, x*sw - pw + s, c]
// (0 outside the image and for columns >= Cin*R*S up to Kpad).  With this column order a (Cout, Kpad) GEMM result IS the
// (Cout, Cin, kh, kw) weight layout: the weight gradient of a few-channel convolution (the ResNet stem, Cin = 3 / 2) becomes a
// 1x1 weight gradient on the tensor-core kernel instead of a CUDA-core loop over 49 taps with Cin padded to 16.
__global__ void __launch_bounds__(256)
im2col_kernel(viai_conv_geom g, const float* __restrict__ in, int Kpad, float* __restrict__ out) {
  // per-k tap table (input offset relative to the patch origin, and the tap's (r, q) for the bounds test), built once per block:
  // the element loop then has no integer division by a runtime constant
  extern __shared__ int tab[];                       // [Kpad][3]: dy, dx, channel  (dy = -1: padding column)
  const int RS = g.R * g.S, K = g.Cin * RS;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    int dy = -1, dx = 0, c = 0;
    if (k < K) {
      c = k / RS;
      const int t = k - c * RS;
      dy = t / g.S;
      dx = t - dy * g.S;
    }
    tab[3 * k] = dy; tab[3 * k + 1] = dx; tab[3 * k + 2] = c;
  }
  __syncthreads();
  const int K4 = Kpad / 4;
  const int64_t M = (int64_t)g.N * g.Hout * g.Wout;
  const int HWo = g.Hout * g.Wout;
  // a thread walks the K4 float4s of a patch row in steps of the block's lanes-per-row: rows are assigned so that a warp writes
  // consecutive float4s of consecutive rows (full 128-byte lines)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M * K4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / K4;
    const int k0 = (int)(i - p * K4) * 4;
    const int n = (int)(p / HWo);
    const int rem = (int)(p - (int64_t)n * HWo);
    const int y = rem / g.Wout, x = rem - y * g.Wout;
    const int Y0 = y * g.stride_h - g.pad_h, X0 = x * g.stride_w - g.pad_w;
    const float* img = in + (int64_t)n * g.Hin * g.Win * g.Cin;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int dy = tab[3 * (k0 + j)], dx = tab[3 * (k0 + j) + 1], c = tab[3 * (k0 + j) + 2];
      const int Y = Y0 + dy, X = X0 + dx;
      v[j] = (dy >= 0 && Y >= 0 && Y < g.Hin && X >= 0 && X < g.Win) ? __ldg(img + ((int64_t)Y * g.Win + X) * g.Cin + c) : 0.f;
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

__global__ void zero_strided_kernel(float* dw, int A, int B, int R, int S, int64_t sa, int64_t sb, int64_t sr, int64_t ss) {
  int64_t total = (int64_t)A * B * R * S;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int s = idx % S; int64_t t = idx / S;
    int r = t % R; t /= R;
    int b = t % B; int a = t / B;
    dw[a * sa + b * sb + r * sr + s * ss] = 0.f;
  }
}

int check_geom(const viai_conv_geom* g) {
  VIAI_REQUIRE(g != nullptr, "conv geometry is NULL");
  VIAI_REQUIRE(g->N > 0 && g->Hin > 0 && g->Win > 0 && g->Cin > 0 && g->Hout > 0 && g->Wout > 0 && g->Cout > 0 &&
                   g->R > 0 && g->S > 0 && g->stride_h > 0 && g->stride_w > 0 && g->pad_h >= 0 && g->pad_w >= 0,
               "invalid conv geometry N=%d in=%dx%dx%d out=%dx%dx%d taps=%dx%d stride=%d,%d pad=%d,%d", g->N, g->Hin,
               g->Win, g->Cin, g->Hout, g->Wout, g->Cout, g->R, g->S, g->stride_h, g->stride_w, g->pad_h, g->pad_w);
  VIAI_REQUIRE(g->mode == 0 || g->mode == 1, "conv mode must be 0 or 1, got %d", g->mode);
  return VIAI_OK;
}

}  // namespace

extern "C" int viai_pack_weight(const float* src, float* dst, int O, int I, int R, int S, int64_t so, int64_t si,
                                int64_t sr, int64_t ss, int flip, viai_stream_t stream) {
  VIAI_REQUIRE(src && dst && O > 0 && I > 0 && R > 0 && S > 0, "viai_pack_weight: bad arguments");
  int64_t total = (int64_t)O * I * R * S;
  int blocks = (int)imin64(cdiv(total, 256), 4 * kNumSMs);
  pack_weight_kernel<<<blocks, 256, 0, STR(stream)>>>(src, dst, O, I, R, S, so, si, sr, ss, flip);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_conv2d_simt(const viai_conv_geom* g, const float* in, const float* wp, const float* bias,
                                float* out, viai_stream_t stream) {
  int rc = check_geom(g);
  if (rc) return rc;
  VIAI_REQUIRE(in && wp && out, "viai_conv2d_simt: NULL tensor");
  const int64_t M = (int64_t)g->N * g->Hout * g->Wout;
  const bool vec = (g->Cin % 16 == 0) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(wp)) % 16 == 0);
  const bool wide = g->Cout > 32;
  dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(g->Cout, wide ? 64 : 32));
  cudaStream_t st = STR(stream);
  if (wide) {
    if (vec) conv_gather_kernel<64, true><<<grid, 256, 0, st>>>(*g, in, wp, bias, out);
    else conv_gather_kernel<64, false><<<grid, 256, 0, st>>>(*g, in, wp, bias, out);
  } else {
    if (vec) conv_gather_kernel<32, true><<<grid, 256, 0, st>>>(*g, in, wp, bias, out);
    else conv_gather_kernel<32, false><<<grid, 256, 0, st>>>(*g, in, wp, bias, out);
  }
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_conv2d_wgrad_simt(const viai_conv_geom* g, const float* U, const float* G, float* dw, int64_t sa,
                                      int64_t sb, int64_t sr, int64_t ss, int accumulate, viai_stream_t stream) {
  int rc = check_geom(g);
  if (rc) return rc;
  VIAI_REQUIRE(U && G && dw, "viai_conv2d_wgrad_simt: NULL tensor");
  cudaStream_t st = STR(stream);
  const int A = g->Cout, B = g->Cin, taps = g->R * g->S;
  const int64_t M = (int64_t)g->N * g->Hout * g->Wout;
  if (!accumulate) {
    int64_t total = (int64_t)A * B * taps;
    zero_strided_kernel<<<(int)imin64(cdiv(total, 256), 2 * kNumSMs), 256, 0, st>>>(dw, A, B, g->R, g->S, sa, sb, sr, ss);
    VIAI_LAUNCHED();
  }
  if (A == 1 && B == 1 && taps <= C1_MAXTAP && M > 0) {
    wgrad_c1_kernel<<<(int)imin64(cdiv(M, 256 * 16), 4 * kNumSMs), 256, 0, st>>>(*g, U, G, dw, sr, ss);
    VIAI_LAUNCHED();
    return VIAI_OK;
  }
  int BA, BB;
  if (A > 32 && B > 32) { BA = 64; BB = 64; }
  else if (A > 32) { BA = 64; BB = 16; if (B > 16) { BA = 64; BB = 64; } }
  else if (B > 32) { BA = 16; BB = 64; if (A > 16) { BA = 64; BB = 64; } }
  else { BA = 32; BB = 32; }
  const int tiles = (int)(cdiv(A, BA) * cdiv(B, BB)) * taps;
  int ksplit = (int)cdiv(4 * kNumSMs, tiles);
  int64_t maxsplit = cdiv(M, 16 * 16);
  if (ksplit > maxsplit) ksplit = (int)maxsplit;
  if (ksplit < 1) ksplit = 1;
  dim3 grid((unsigned)(cdiv(A, BA) * cdiv(B, BB)), (unsigned)taps, (unsigned)ksplit);
  if (BA == 64 && BB == 64) wgrad_kernel<64, 64><<<grid, 256, 0, st>>>(*g, U, G, dw, sa, sb, sr, ss, ksplit);
  else if (BA == 64 && BB == 16) wgrad_kernel<64, 16><<<grid, 256, 0, st>>>(*g, U, G, dw, sa, sb, sr, ss, ksplit);
  else if (BA == 16 && BB == 64) wgrad_kernel<16, 64><<<grid, 256, 0, st>>>(*g, U, G, dw, sa, sb, sr, ss, ksplit);
  else wgrad_kernel<32, 32><<<grid, 256, 0, st>>>(*g, U, G, dw, sa, sb, sr, ss, ksplit);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_im2col(const viai_conv_geom* g, const float* in, int Kpad, float* out, viai_stream_t stream) {
  VIAI_REQUIRE(g && in && out && g->mode == 0 && Kpad % 4 == 0 && Kpad >= g->Cin * g->R * g->S, "viai_im2col: bad arguments");
  VIAI_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "viai_im2col: out must be 16-byte aligned");
  const int64_t total = (int64_t)g->N * g->Hout * g->Wout * (Kpad / 4);
  VIAI_REQUIRE(total > 0, "viai_im2col: empty tensor");
  VIAI_REQUIRE(Kpad <= 4096, "viai_im2col: Kpad %d too large", Kpad);
  im2col_kernel<<<(int)imin64(cdiv(total, 256), 32 * kNumSMs), 256, (size_t)Kpad * 3 * sizeof(int), STR(stream)>>>(*g, in, Kpad, out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
