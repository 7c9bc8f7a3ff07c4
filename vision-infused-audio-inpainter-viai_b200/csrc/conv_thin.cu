// Convolutions with a single input or a single output channel (MelEncoder.conv1 1->32, MelDiscriminator.conv1 1->64 1x4,
// MelDecoder.conv6_2 32->1, MelDiscriminator.conv4 512->1 and their gradients).  They carry <1 % of the step's FLOPs and are
// HBM-bound (their cost is reading / writing the wide tensor once), so they are plain CUDA-core kernels shaped for coalesced
// 128-bit accesses rather than GEMM tiles:
//   cout1:  one warp per output pixel, lanes stride over (tap, 4-channel vector), shuffle reduction;
//   cin1:   one thread per (output pixel, 4 output channels), the <= 12 scalar taps come from L1, weights from shared memory;
//   wgrad:  lanes over 4-channel vectors of the wide tensor, per-thread accumulators for every tap, block reduction, atomics.
#include "common.cuh"
using namespace viai;

namespace {

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

// ---- Cout == 1 ---------------------------------------------------------------------------------------------------
// wp: [R][S][Cin]
__global__ void __launch_bounds__(256) conv_cout1_kernel(viai_conv_geom g, const float* __restrict__ in, const float* __restrict__ wp,
                                                         const float* __restrict__ bias, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t M = (int64_t)g.N * g.Hout * g.Wout;
  const int c4n = g.Cin >> 2;
  const int items = g.R * g.S * c4n;
  const float b = bias ? __ldg(bias) : 0.f;
  for (int64_t m = warp0; m < M; m += nwarps) {
    const int x = (int)(m % g.Wout);
    const int64_t t = m / g.Wout;
    const int y = (int)(t % g.Hout);
    const int n = (int)(t / g.Hout);
    float acc = 0.f;
    for (int idx = lane; idx < items; idx += 32) {
      const int tap = idx / c4n, cv = idx - tap * c4n;
      const int r = tap / g.S, s = tap - r * g.S;
      int Y, X;
      if (gather_coord(g.mode, y, r, g.stride_h, g.pad_h, g.Hin, Y) && gather_coord(g.mode, x, s, g.stride_w, g.pad_w, g.Win, X)) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(in + (((int64_t)n * g.Hin + Y) * g.Win + X) * g.Cin) + cv);
        const float4 w = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)tap * g.Cin) + cv);
        acc = fmaf(a.x, w.x, acc); acc = fmaf(a.y, w.y, acc); acc = fmaf(a.z, w.z, acc); acc = fmaf(a.w, w.w, acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[m] = acc + b;
  }
}

// ---- Cin == 1 ----------------------------------------------------------------------------------------------------
// wp: [Cout][R][S]
__global__ void __launch_bounds__(256) conv_cin1_kernel(viai_conv_geom g, const float* __restrict__ in, const float* __restrict__ wp,
                                                        const float* __restrict__ bias, float* __restrict__ out) {
  extern __shared__ float wsm[];   // [tap][Cout] + bias[Cout]
  const int taps = g.R * g.S;
  for (int i = threadIdx.x; i < taps * g.Cout; i += blockDim.x) {
    const int co = i / taps, tap = i - co * taps;
    wsm[tap * g.Cout + co] = wp[i];
  }
  for (int i = threadIdx.x; i < g.Cout; i += blockDim.x) wsm[taps * g.Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int c4n = g.Cout >> 2;
  const int64_t total = (int64_t)g.N * g.Hout * g.Wout * c4n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % c4n);
    const int64_t m = i / c4n;
    const int x = (int)(m % g.Wout);
    const int64_t t = m / g.Wout;
    const int y = (int)(t % g.Hout);
    const int n = (int)(t / g.Hout);
    float4 acc = *reinterpret_cast<const float4*>(wsm + taps * g.Cout + cv * 4);
    for (int r = 0; r < g.R; ++r) {
      int Y;
      if (!gather_coord(g.mode, y, r, g.stride_h, g.pad_h, g.Hin, Y)) continue;
      for (int s = 0; s < g.S; ++s) {
        int X;
        if (!gather_coord(g.mode, x, s, g.stride_w, g.pad_w, g.Win, X)) continue;
        const float a = __ldg(in + ((int64_t)n * g.Hin + Y) * g.Win + X);
        const float4 w = *reinterpret_cast<const float4*>(wsm + (r * g.S + s) * g.Cout + cv * 4);
        acc.x = fmaf(a, w.x, acc.x); acc.y = fmaf(a, w.y, acc.y); acc.z = fmaf(a, w.z, acc.z); acc.w = fmaf(a, w.w, acc.w);
      }
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

// ---- weight gradient with A == 1 or B == 1 -------------------------------------------------------------------------
constexpr int WG_MAXTAP = 12;

__global__ void zero_dw_kernel(float* dw, int A, int B, int R, int S, int64_t sa, int64_t sb, int64_t sr, int64_t ss) {
  const int total = A * B * R * S;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int t = i;
    const int s = t % S; t /= S;
    const int r = t % R; t /= R;
    const int b = t % B;
    const int a = t / B;
    dw[a * sa + b * sb + r * sr + s * ss] = 0.f;
  }
}

// U: (N,Hout,Wout,A), G: (N,Hin,Win,B); WIDE_U: B == 1 (the wide tensor is U), else A == 1 (the wide tensor is G).
template <bool WIDE_U>
__global__ void __launch_bounds__(256) wgrad_thin_kernel(viai_conv_geom g, const float* __restrict__ U, const float* __restrict__ G,
                                                         float* __restrict__ dw, int64_t sa, int64_t sb, int64_t sr, int64_t ss,
                                                         int64_t pix_per_block) {
  extern __shared__ float red[];   // [planes][taps][C]
  const int C = WIDE_U ? g.Cout : g.Cin;
  const int c4n = C >> 2;
  const int planes = blockDim.x / c4n;
  const int cv = threadIdx.x % c4n, pl = threadIdx.x / c4n;
  const int taps = g.R * g.S;
  float4 acc[WG_MAXTAP];
#pragma unroll
  for (int t = 0; t < WG_MAXTAP; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t M = (int64_t)g.N * g.Hout * g.Wout;
  const int64_t p0 = blockIdx.x * pix_per_block, p1 = imin64(p0 + pix_per_block, M);
  if (pl < planes) {
    for (int64_t m = p0 + pl; m < p1; m += planes) {
      const int x = (int)(m % g.Wout);
      const int64_t tt = m / g.Wout;
      const int y = (int)(tt % g.Hout);
      const int n = (int)(tt / g.Hout);
      float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float us = 0.f;
      if (WIDE_U) u4 = __ldg(reinterpret_cast<const float4*>(U + m * C) + cv);
      else us = __ldg(U + m);
#pragma unroll
      for (int t = 0; t < WG_MAXTAP; ++t) {
        if (t < taps) {
          const int r = t / g.S, s = t - r * g.S;
          const int Y = y * g.stride_h - g.pad_h + r, X = x * g.stride_w - g.pad_w + s;
          if (Y >= 0 && Y < g.Hin && X >= 0 && X < g.Win) {
            const int64_t gp = ((int64_t)n * g.Hin + Y) * g.Win + X;
            if (WIDE_U) {
              const float gs = __ldg(G + gp);
              acc[t].x = fmaf(u4.x, gs, acc[t].x); acc[t].y = fmaf(u4.y, gs, acc[t].y);
              acc[t].z = fmaf(u4.z, gs, acc[t].z); acc[t].w = fmaf(u4.w, gs, acc[t].w);
            } else {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(G + gp * C) + cv);
              acc[t].x = fmaf(g4.x, us, acc[t].x); acc[t].y = fmaf(g4.y, us, acc[t].y);
              acc[t].z = fmaf(g4.z, us, acc[t].z); acc[t].w = fmaf(g4.w, us, acc[t].w);
            }
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < WG_MAXTAP; ++t)
      if (t < taps) *reinterpret_cast<float4*>(red + ((size_t)(pl * taps + t) * C) + cv * 4) = acc[t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < taps * C; i += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < planes; ++q) s += red[(size_t)q * taps * C + i];
    const int t = i / C, c = i - t * C;
    const int r = t / g.S, sx = t - r * g.S;
    const int a = WIDE_U ? c : 0, b = WIDE_U ? 0 : c;
    atomicAdd(dw + a * sa + b * sb + r * sr + sx * ss, s);
  }
}

}  // namespace

extern "C" int viai_conv2d_thin_supported(const viai_conv_geom* g) {
  if (!g) return 0;
  if (g->Cout == 1 && g->Cin % 4 == 0 && g->Cin >= 4) return 1;
  if (g->Cin == 1 && g->Cout % 4 == 0 && g->Cout >= 4 && g->R * g->S * g->Cout + g->Cout <= 12 * 1024) return 2;
  return 0;
}

// Same operands as viai_conv2d_simt: wp is the [O][R][S][I] re-layout produced by viai_pack_weight.
extern "C" int viai_conv2d_thin(const viai_conv_geom* gp, const float* in, const float* wp, const float* bias, float* out,
                                viai_stream_t stream) {
  VIAI_REQUIRE(gp && in && wp && out, "conv2d_thin: null argument");
  const viai_conv_geom& g = *gp;
  const int kind = viai_conv2d_thin_supported(gp);
  VIAI_REQUIRE(kind != 0, "conv2d_thin: unsupported geometry (Cin %d, Cout %d)", g.Cin, g.Cout);
  VIAI_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(wp) & 15) == 0,
               "conv2d_thin: pointers must be 16-byte aligned");
  const int64_t M = (int64_t)g.N * g.Hout * g.Wout;
  if (M == 0) return VIAI_OK;
  if (kind == 1) {
    const int blocks = (int)imin64(cdiv(M, 8), 16 * kNumSMs);
    conv_cout1_kernel<<<blocks, 256, 0, STR(stream)>>>(g, in, wp, bias, out);
  } else {
    const int64_t total = M * (g.Cout / 4);
    const int blocks = (int)imin64(cdiv(total, 256), 16 * kNumSMs);
    const size_t smem = sizeof(float) * (size_t)(g.R * g.S * g.Cout + g.Cout);
    conv_cin1_kernel<<<blocks, 256, smem, STR(stream)>>>(g, in, wp, bias, out);
  }
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_conv2d_wgrad_thin_supported(const viai_conv_geom* g) {
  if (!g || g->R * g->S > WG_MAXTAP) return 0;
  const int C = g->Cin == 1 ? g->Cout : (g->Cout == 1 ? g->Cin : 0);
  if (C == 0 || (g->Cin == 1 && g->Cout == 1)) return 0;
  return (C % 4 == 0 && C / 4 <= 256) ? 1 : 0;
}

// Same meaning as viai_conv2d_wgrad_simt, for A == 1 or B == 1.
extern "C" int viai_conv2d_wgrad_thin(const viai_conv_geom* gp, const float* U, const float* G, float* dw, int64_t sa, int64_t sb,
                                      int64_t sr, int64_t ss, int accumulate, viai_stream_t stream) {
  VIAI_REQUIRE(gp && U && G && dw, "conv2d_wgrad_thin: null argument");
  const viai_conv_geom& g = *gp;
  VIAI_REQUIRE(viai_conv2d_wgrad_thin_supported(gp), "conv2d_wgrad_thin: unsupported geometry (A %d, B %d)", g.Cout, g.Cin);
  VIAI_REQUIRE((reinterpret_cast<uintptr_t>(U) & 15) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0,
               "conv2d_wgrad_thin: pointers must be 16-byte aligned");
  cudaStream_t st = STR(stream);
  const int taps = g.R * g.S;
  if (!accumulate) {
    zero_dw_kernel<<<(g.Cout * g.Cin * taps + 255) / 256, 256, 0, st>>>(dw, g.Cout, g.Cin, g.R, g.S, sa, sb, sr, ss);
    VIAI_LAUNCHED();
  }
  const int64_t M = (int64_t)g.N * g.Hout * g.Wout;
  if (M == 0) return VIAI_OK;
  const bool wide_u = g.Cin == 1;
  const int C = wide_u ? g.Cout : g.Cin;
  const int c4n = C / 4, planes = 256 / c4n;
  int64_t blocks = imin64(4 * kNumSMs, cdiv(M, (int64_t)planes * 4));
  if (blocks < 1) blocks = 1;
  const int64_t ppb = cdiv(M, blocks);
  blocks = cdiv(M, ppb);
  const size_t smem = sizeof(float) * (size_t)planes * taps * C;
  static bool attr = false;
  if (!attr) {
    VIAI_CUDA(cudaFuncSetAttribute(wgrad_thin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    VIAI_CUDA(cudaFuncSetAttribute(wgrad_thin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  VIAI_REQUIRE(smem <= 64 * 1024, "conv2d_wgrad_thin: reduction buffer too large");
  if (wide_u) wgrad_thin_kernel<true><<<(unsigned)blocks, 256, smem, st>>>(g, U, G, dw, sa, sb, sr, ss, ppb);
  else wgrad_thin_kernel<false><<<(unsigned)blocks, 256, smem, st>>>(g, U, G, dw, sa, sb, sr, ss, ppb);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
