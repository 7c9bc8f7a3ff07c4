// Convolutions with a single input or a single output channel (MelEncoder.conv1 1->32, MelDiscriminator.conv1 1->64 1x4,
// MelDecoder.conv6_2 32->1, MelDiscriminator.conv4 512->1 and their gradients).  They carry <1 % of the step's FLOPs and are
// HBM-bound (their cost is reading / writing the wide tensor once), so they are plain CUDA-core kernels shaped for coalesced
// 128-bit accesses rather than GEMM tiles:
//   cout1:  input-stationary.  A CTA owns a tile of output pixels; the input patch all of its taps read is brought in ONCE by
//           TMA (32-channel slabs, 128-byte swizzle, zero fill outside the image = the padding), one thread per patch pixel
//           turns its 128-byte row into one partial dot product per filter tap, the partials are exchanged through shared
//           memory and every output pixel sums the taps that reach it.  The wide tensor is read exactly once from HBM.
//           (Geometries the tiler does not cover fall back to a warp-per-output-pixel kernel.)
//   cin1:   one thread per (4 adjacent output pixels, 4 output channels): weights come from shared memory once per tap and
//           are reused for the 4 pixels, the scalar taps from L1, the stores are full 128-bit rows of the wide tensor;
//   wgrad:  wide-tensor-stationary: lanes over 4-channel vectors of the wide tensor, every wide pixel is read once and
//           multiplied with the <= 12 scalars of the thin tensor its taps touch; per-thread accumulators for every tap,
//           block reduction, atomics.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>
using namespace viai;
using namespace viai::tc;

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


// ---- Cout == 1, TMA-tiled input-stationary kernel -------------------------------------------------------------------------
constexpr int C1_THREADS = 256;
constexpr int C1_MAXPP = 2;          // patch pixels per thread
constexpr int C1_MAXST = 4;          // slab stages

struct Cout1Params {
  CUtensorMap map;                   // (C, Win, Hin, N) fp32, box (32, PW, PH, 1), 128-byte swizzle
  viai_conv_geom g;
  const float* wp;                   // [R][S][C]
  const float* bias;
  float* out;
  int32_t TY, TX, PH, PW, tilesX, tilesY, nslab, nstage;
  uint32_t stage_bytes;
};

__host__ __device__ inline int c1_floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
// first input row / column of the patch that the output tile starting at o0 reads
__host__ __device__ inline int c1_patch_origin(int mode, int o0, int K, int stride, int pad) {
  return mode == 0 ? o0 * stride - pad : c1_floordiv(o0 + pad - (K - 1), stride);
}

template <int TAPS>
__global__ void __launch_bounds__(C1_THREADS) conv_cout1_tma_kernel(const __grid_constant__ Cout1Params p) {
  extern __shared__ __align__(1024) uint8_t c1_smem_raw[];
  uint8_t* smem = c1_smem_raw + ((1024u - (smem_u32(c1_smem_raw) & 1023u)) & 1023u);
  const viai_conv_geom& g = p.g;
  const int C = g.Cin;
  const int npix = p.PH * p.PW;
  uint8_t* stages = smem;
  float* wsm = reinterpret_cast<float*>(stages + (size_t)p.nstage * p.stage_bytes);    // [TAPS][C]
  float* part = wsm + TAPS * C;                                                          // [TAPS][npix]
  uint64_t* full = reinterpret_cast<uint64_t*>(part + TAPS * npix + ((TAPS * npix) & 1));

  int tile = blockIdx.x;
  const int tx = tile % p.tilesX; tile /= p.tilesX;
  const int ty = tile % p.tilesY;
  const int n = tile / p.tilesY;
  const int y0 = ty * p.TY, x0 = tx * p.TX;
  const int Y0 = c1_patch_origin(g.mode, y0, g.R, g.stride_h, g.pad_h);
  const int X0 = c1_patch_origin(g.mode, x0, g.S, g.stride_w, g.pad_w);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstage; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
    prefetch_tmap(&p.map);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstage && s < p.nslab; ++s) {
      mbar_expect_tx(&full[s], (uint32_t)npix * 128u);
      tma_load_4d(stages + (size_t)s * p.stage_bytes, &p.map, &full[s], s * 32, X0, Y0, n);
    }
  }
  for (int i = threadIdx.x; i < TAPS * C; i += C1_THREADS) wsm[i] = __ldg(p.wp + i);
  __syncthreads();

  float acc[C1_MAXPP][TAPS];
#pragma unroll
  for (int k = 0; k < C1_MAXPP; ++k)
#pragma unroll
    for (int t = 0; t < TAPS; ++t) acc[k][t] = 0.f;

  for (int s = 0; s < p.nslab; ++s) {
    const int st = s % p.nstage;
    mbar_wait(&full[st], (uint32_t)(s / p.nstage) & 1u);
    const uint8_t* base = stages + (size_t)st * p.stage_bytes;
    const float* wslab = wsm + s * 32;
#pragma unroll
    for (int k = 0; k < C1_MAXPP; ++k) {
      const int pix = threadIdx.x + k * C1_THREADS;
      if (pix < npix) {
        const uint8_t* row = base + (size_t)pix * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(row + ((j ^ (pix & 7)) << 4));
#pragma unroll
          for (int t = 0; t < TAPS; ++t) {
            const float4 w = *reinterpret_cast<const float4*>(wslab + t * C + j * 4);
            acc[k][t] = fmaf(v.x, w.x, acc[k][t]); acc[k][t] = fmaf(v.y, w.y, acc[k][t]);
            acc[k][t] = fmaf(v.z, w.z, acc[k][t]); acc[k][t] = fmaf(v.w, w.w, acc[k][t]);
          }
        }
      }
    }
    if (s + p.nstage < p.nslab) {
      __syncthreads();                               // every thread is done reading this stage
      if (threadIdx.x == 0) {
        mbar_expect_tx(&full[st], (uint32_t)npix * 128u);
        tma_load_4d(stages + (size_t)st * p.stage_bytes, &p.map, &full[st], (s + p.nstage) * 32, X0, Y0, n);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < C1_MAXPP; ++k) {
    const int pix = threadIdx.x + k * C1_THREADS;
    if (pix < npix) {
#pragma unroll
      for (int t = 0; t < TAPS; ++t) part[t * npix + pix] = acc[k][t];
    }
  }
  __syncthreads();
  const float b = p.bias ? __ldg(p.bias) : 0.f;
  for (int o = threadIdx.x; o < p.TY * p.TX; o += C1_THREADS) {
    const int ly = o / p.TX, lx = o - ly * p.TX;
    const int y = y0 + ly, x = x0 + lx;
    if (y >= g.Hout || x >= g.Wout) continue;
    float sum = b;
    for (int r = 0; r < g.R; ++r) {
      int Y;
      if (g.mode == 0) {
        Y = y * g.stride_h - g.pad_h + r;
      } else {
        const int t = y + g.pad_h - r;
        if (t < 0) continue;
        Y = t / g.stride_h;
        if (Y * g.stride_h != t) continue;
      }
      const int py = Y - Y0;
      if (py < 0 || py >= p.PH) continue;            // cannot happen for covered geometries; keeps the read in bounds
      for (int s = 0; s < g.S; ++s) {
        int X;
        if (g.mode == 0) {
          X = x * g.stride_w - g.pad_w + s;
        } else {
          const int t = x + g.pad_w - s;
          if (t < 0) continue;
          X = t / g.stride_w;
          if (X * g.stride_w != t) continue;
        }
        const int px = X - X0;
        if (px < 0 || px >= p.PW) continue;
        sum += part[(r * g.S + s) * npix + py * p.PW + px];
      }
    }
    p.out[((int64_t)n * g.Hout + y) * g.Wout + x] = sum;
  }
}

// extent of the input patch along one axis (max over all tiles)
inline int c1_patch_extent(int mode, int out_len, int T, int K, int stride, int pad) {
  int ext = 1;
  for (int o0 = 0; o0 < out_len; o0 += T) {
    const int last_o = o0 + T - 1;
    const int first = c1_patch_origin(mode, o0, K, stride, pad);
    const int last = mode == 0 ? last_o * stride - pad + K - 1 : c1_floordiv(last_o + pad, stride);
    if (last - first + 1 > ext) ext = last - first + 1;
  }
  return ext;
}

// Returns VIAI_OK when launched, VIAI_ERR_UNSUPPORTED when the geometry is left to the generic kernel.
int launch_cout1_tma(const viai_conv_geom& g, const float* in, const float* wp, const float* bias, float* out, cudaStream_t st) {
  const int taps = g.R * g.S;
  if (g.Cin % 32 != 0 || (taps != 9 && taps != 4)) return VIAI_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(in) & 15) != 0) return VIAI_ERR_UNSUPPORTED;
  Cout1Params p;
  memset(&p, 0, sizeof(p));
  p.g = g; p.wp = wp; p.bias = bias; p.out = out;
  p.TY = 8 * (g.mode == 1 ? g.stride_h : 1);
  p.TX = 32 * (g.mode == 1 ? g.stride_w : 1);
  p.PH = c1_patch_extent(g.mode, g.Hout, p.TY, g.R, g.stride_h, g.pad_h);
  p.PW = c1_patch_extent(g.mode, g.Wout, p.TX, g.S, g.stride_w, g.pad_w);
  const int npix = p.PH * p.PW;
  if (npix > C1_MAXPP * C1_THREADS || p.PH > 256 || p.PW > 256) return VIAI_ERR_UNSUPPORTED;
  p.tilesX = (g.Wout + p.TX - 1) / p.TX;
  p.tilesY = (g.Hout + p.TY - 1) / p.TY;
  p.nslab = g.Cin / 32;
  p.nstage = p.nslab < C1_MAXST ? p.nslab : C1_MAXST;
  p.stage_bytes = ((uint32_t)npix * 128u + 1023u) & ~1023u;
  size_t smem = 1024 + (size_t)p.nstage * p.stage_bytes + sizeof(float) * ((size_t)taps * g.Cin + (size_t)taps * npix + 2) + 8 * C1_MAXST;
  while (smem > 200 * 1024 && p.nstage > 1) { --p.nstage; smem -= p.stage_bytes; }
  if (smem > 200 * 1024) return VIAI_ERR_UNSUPPORTED;
  uint64_t dims[4] = {(uint64_t)g.Cin, (uint64_t)g.Win, (uint64_t)g.Hin, (uint64_t)g.N};
  uint64_t strides[3] = {(uint64_t)g.Cin * 4, (uint64_t)g.Win * g.Cin * 4, (uint64_t)g.Hin * g.Win * g.Cin * 4};
  uint32_t box[4] = {32, (uint32_t)p.PW, (uint32_t)p.PH, 1};
  if (encode_f32_map(&p.map, 4, in, dims, strides, box, 1)) return VIAI_ERR_CUDA;
  static bool attr = false;
  if (!attr) {
    VIAI_CUDA(cudaFuncSetAttribute(conv_cout1_tma_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VIAI_CUDA(cudaFuncSetAttribute(conv_cout1_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const int grid = g.N * p.tilesX * p.tilesY;
  if (taps == 9) conv_cout1_tma_kernel<9><<<grid, C1_THREADS, smem, st>>>(p);
  else conv_cout1_tma_kernel<4><<<grid, C1_THREADS, smem, st>>>(p);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

// ---- Cin == 1 ----------------------------------------------------------------------------------------------------
// wp: [Cout][R][S].  Eight lanes share CIN1_PX adjacent output pixels: the (tap, pixel) input scalars and their index
// arithmetic are evaluated once, then the lanes sweep the output channels 32 at a time (128-byte stores per pixel).
// STATS: the per-channel sum / sum of squares of the output (what the following BatchNorm needs; viai_channel_stats's result)
// come out of the same kernel: a lane always owns the same <= CIN1_SV channel vectors, keeps fp32 partial sums over the ~16
// pixels it produces, the partials are reduced over the lanes of a warp by shuffles, over the warps through shared memory in
// double, and one double atomic per (block, channel, moment) lands in HBM -- the separate read pass over the wide tensor
// (65 us for MelDiscriminator.conv1's 268 MB at B = 32) disappears.
// The 1x4 statistics variant is held to 80 registers (3 CTAs per SM instead of 2; 16 bytes of spill): C2 step 15.19 vs 15.26 ms.
constexpr int CIN1_PX = 4;
constexpr int CIN1_SV = 2;          // channel vectors per lane with STATS: Cout <= 64
template <int R_, int S_, bool STATS>
__global__ void __launch_bounds__(256, (STATS && R_ * S_ <= 4) ? 3 : 1) conv_cin1_kernel(viai_conv_geom g, const float* __restrict__ in, const float* __restrict__ wp,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        double* __restrict__ stat_sum, double* __restrict__ stat_sumsq) {
  extern __shared__ float wsm[];   // [tap][Cout] + bias[Cout]
  constexpr int taps = R_ * S_;
  float4 ssum[CIN1_SV], ssq[CIN1_SV];
#pragma unroll
  for (int j = 0; j < CIN1_SV; ++j) { ssum[j] = make_float4(0.f, 0.f, 0.f, 0.f); ssq[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
  for (int i = threadIdx.x; i < taps * g.Cout; i += blockDim.x) {
    const int co = i / taps, tap = i - co * taps;
    wsm[tap * g.Cout + co] = wp[i];
  }
  for (int i = threadIdx.x; i < g.Cout; i += blockDim.x) wsm[taps * g.Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int c4n = g.Cout >> 2;
  const int xg_n = (g.Wout + CIN1_PX - 1) / CIN1_PX;
  const int groups = g.N * g.Hout * xg_n;               // < 2^31 for every tensor that fits the 32-bit geometry struct
  const int l8 = threadIdx.x & 7;
  for (int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; gi < groups; gi += (gridDim.x * blockDim.x) >> 3) {
    int m = gi;
    const int xg = m % xg_n; m /= xg_n;
    const int y = m % g.Hout;
    const int n = m / g.Hout;
    const int x0 = xg * CIN1_PX;
    float a[taps][CIN1_PX];
    int Xs[S_][CIN1_PX];                                 // column of (tap s, pixel k), or -1
#pragma unroll
    for (int s = 0; s < S_; ++s)
#pragma unroll
      for (int k = 0; k < CIN1_PX; ++k) {
        int X;
        Xs[s][k] = gather_coord(g.mode, x0 + k, s, g.stride_w, g.pad_w, g.Win, X) ? X : -1;
      }
#pragma unroll
    for (int r = 0; r < R_; ++r) {
      int Y;
      const bool rv = gather_coord(g.mode, y, r, g.stride_h, g.pad_h, g.Hin, Y);
      const float* irow = in + ((int64_t)n * g.Hin + (rv ? Y : 0)) * g.Win;
#pragma unroll
      for (int s = 0; s < S_; ++s)
#pragma unroll
        for (int k = 0; k < CIN1_PX; ++k) a[r * S_ + s][k] = (rv && Xs[s][k] >= 0) ? __ldg(irow + Xs[s][k]) : 0.f;
    }
    float* obase = out + (((int64_t)n * g.Hout + y) * g.Wout + x0) * g.Cout;
    auto one_vector = [&](int cv, float4& s4, float4& q4) {
      const float4 b4 = *reinterpret_cast<const float4*>(wsm + taps * g.Cout + cv * 4);
      float4 acc[CIN1_PX];
#pragma unroll
      for (int k = 0; k < CIN1_PX; ++k) acc[k] = b4;
#pragma unroll
      for (int t = 0; t < taps; ++t) {
        const float4 w = *reinterpret_cast<const float4*>(wsm + t * g.Cout + cv * 4);
#pragma unroll
        for (int k = 0; k < CIN1_PX; ++k) {
          acc[k].x = fmaf(a[t][k], w.x, acc[k].x); acc[k].y = fmaf(a[t][k], w.y, acc[k].y);
          acc[k].z = fmaf(a[t][k], w.z, acc[k].z); acc[k].w = fmaf(a[t][k], w.w, acc[k].w);
        }
      }
#pragma unroll
      for (int k = 0; k < CIN1_PX; ++k)
        if (x0 + k < g.Wout) {
          *reinterpret_cast<float4*>(obase + (int64_t)k * g.Cout + cv * 4) = acc[k];
          if (STATS) {
            s4.x += acc[k].x; s4.y += acc[k].y; s4.z += acc[k].z; s4.w += acc[k].w;
            q4.x = fmaf(acc[k].x, acc[k].x, q4.x); q4.y = fmaf(acc[k].y, acc[k].y, q4.y);
            q4.z = fmaf(acc[k].z, acc[k].z, q4.z); q4.w = fmaf(acc[k].w, acc[k].w, q4.w);
          }
        }
    };
    if (STATS) {
#pragma unroll
      for (int j = 0; j < CIN1_SV; ++j)
        if (l8 + 8 * j < c4n) one_vector(l8 + 8 * j, ssum[j], ssq[j]);
    } else {
      float4 dummy_s = make_float4(0.f, 0.f, 0.f, 0.f), dummy_q = dummy_s;
      for (int cv = l8; cv < c4n; cv += 8) one_vector(cv, dummy_s, dummy_q);
    }
  }
  if (STATS) {
    // lanes l8, l8+8, l8+16, l8+24 of a warp own the same channels
    float v[CIN1_SV][8];
#pragma unroll
    for (int j = 0; j < CIN1_SV; ++j) {
      v[j][0] = ssum[j].x; v[j][1] = ssum[j].y; v[j][2] = ssum[j].z; v[j][3] = ssum[j].w;
      v[j][4] = ssq[j].x; v[j][5] = ssq[j].y; v[j][6] = ssq[j].z; v[j][7] = ssq[j].w;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[j][i] += __shfl_xor_sync(0xffffffffu, v[j][i], 8);
        v[j][i] += __shfl_xor_sync(0xffffffffu, v[j][i], 16);
      }
    }
    __syncthreads();                                     // every thread is done with the weights in wsm: reuse it
    float* red = wsm;                                    // [warp 8][l8 8][j CIN1_SV][8]  (4 KB <= the weight table? see host)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < CIN1_SV; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) red[((warp * 8 + lane) * CIN1_SV + j) * 8 + i] = v[j][i];
    }
    __syncthreads();
    // thread t < 2*Cout: moment m = t / Cout, channel c = t % Cout;  c = (l8 + 8 j) * 4 + e
    const int t = threadIdx.x;
    if (t < 2 * g.Cout) {
      const int m = t / g.Cout, c = t - m * g.Cout;
      const int cv = c >> 2, e = c & 3;
      const int l = cv & 7, j = cv >> 3;
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += (double)red[((w * 8 + l) * CIN1_SV + j) * 8 + m * 4 + e];
      atomicAdd((m == 0 ? stat_sum : stat_sumsq) + c, tot);
    }
  }
}

// generic tap counts: one thread per (output pixel, 4 output channels)
__global__ void __launch_bounds__(256) conv_cin1_generic_kernel(viai_conv_geom g, const float* __restrict__ in,
                                                                const float* __restrict__ wp, const float* __restrict__ bias,
                                                                float* __restrict__ out) {
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
// The loop runs over the pixels of the WIDE tensor (each read once, 128-bit); the thin tensor supplies one scalar per tap.
// R_ x S_ are compile-time (3x3, 1x4; R_ == 0: generic, up to WG_MAXTAP runtime taps) so that the tap loops unroll into
// straight-line code with the per-tap offsets folded into immediates.
template <bool WIDE_U, int R_, int S_, int UN>
__global__ void __launch_bounds__(256) wgrad_thin_kernel(viai_conv_geom g, const float* __restrict__ U, const float* __restrict__ G,
                                                         float* __restrict__ ws, int pix_per_block) {
  extern __shared__ float red[];   // [planes][taps][C]
  constexpr bool FIXED = R_ > 0;
  constexpr int NT = FIXED ? R_ * S_ : WG_MAXTAP;
  const int C = WIDE_U ? g.Cout : g.Cin;
  const int c4n = C >> 2;
  const int planes = blockDim.x / c4n;
  const int cv = threadIdx.x % c4n, pl = threadIdx.x / c4n;
  const int taps = FIXED ? NT : g.R * g.S;
  const int Sr = FIXED ? S_ : g.S;
  const int Hw = WIDE_U ? g.Hout : g.Hin, Ww = WIDE_U ? g.Wout : g.Win;       // wide tensor extent
  float4 acc[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int M = g.N * Hw * Ww;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, M);
  const float4* wide = reinterpret_cast<const float4*>(WIDE_U ? U : G);
  const float* thin = WIDE_U ? G : U;
  if (pl < planes) {
    const bool unit = g.stride_h == 1 && g.stride_w == 1;
    const int Ht = WIDE_U ? g.Hin : g.Hout, Wt = WIDE_U ? g.Win : g.Wout;       // thin tensor extent
    int m = p0 + pl;
    int x = m % Ww, tq = m / Ww;
    int y = tq % Hw, n = tq / Hw;
    // UN wide pixels (16 bytes each) in flight per thread
    while (m < p1) {
      float4 w4[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int mu = m + u * planes;
        w4[u] = mu < p1 ? __ldg(wide + (int64_t)mu * c4n + cv) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (m + u * planes < p1) {
          const float* tbase = thin + (int64_t)n * Ht * Wt;
          // WIDE_U: thin(Y, X) = (y*sh - ph + r, x*sw - pw + s);  else: thin = ((y + ph - r)/sh, (x + pw - s)/sw) when divisible
          const int by = WIDE_U ? y * g.stride_h - g.pad_h : y + g.pad_h;
          const int bx = WIDE_U ? x * g.stride_w - g.pad_w : x + g.pad_w;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            if (FIXED || t < taps) {
              const int r = FIXED ? t / S_ : t / Sr, s = FIXED ? t % S_ : t - r * Sr;
              int Y = WIDE_U ? by + r : by - r, X = WIDE_U ? bx + s : bx - s;
              bool ok = true;
              if (!WIDE_U && !unit) {
                const int ty = Y, tx = X;
                Y = ty / g.stride_h; X = tx / g.stride_w;
                ok = ty >= 0 && tx >= 0 && Y * g.stride_h == ty && X * g.stride_w == tx;
              }
              ok = ok && (unsigned)Y < (unsigned)Ht && (unsigned)X < (unsigned)Wt;
              const float v = ok ? __ldg(tbase + Y * Wt + X) : 0.f;
              acc[t].x = fmaf(w4[u].x, v, acc[t].x); acc[t].y = fmaf(w4[u].y, v, acc[t].y);
              acc[t].z = fmaf(w4[u].z, v, acc[t].z); acc[t].w = fmaf(w4[u].w, v, acc[t].w);
            }
          }
          x += planes;
          while (x >= Ww) { x -= Ww; if (++y == Hw) { y = 0; ++n; } }
        }
      }
      m += UN * planes;
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (FIXED || t < taps) *reinterpret_cast<float4*>(red + ((size_t)(pl * taps + t) * C) + cv * 4) = acc[t];
  }
  __syncthreads();
  // the block's partial sums go to its own workspace row (no atomics: hundreds of CTAs adding onto the same few hundred
  // addresses serialise in the L2 atomic units); wgrad_thin_finish_kernel adds the rows up
  for (int i = threadIdx.x; i < taps * C; i += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < planes; ++q) s += red[(size_t)q * taps * C + i];
    ws[(size_t)blockIdx.x * taps * C + i] = s;
  }
}

// dw[tap, c] (+)= sum over block rows.  Block = 32 consecutive (tap, c) elements x 8 row lanes.
__global__ void __launch_bounds__(256) wgrad_thin_finish_kernel(const float* __restrict__ ws, int nrows, int taps, int C, int S,
                                                                int wide_u, float* __restrict__ dw, int64_t sa, int64_t sb,
                                                                int64_t sr, int64_t ss, int accumulate) {
  __shared__ float part[8][33];
  const int total = taps * C;
  const int li = threadIdx.x & 31, lp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + li;
  float acc = 0.f;
  if (i < total)
    for (int q = lp; q < nrows; q += 8) acc += ws[(size_t)q * total + i];
  part[lp][li] = acc;
  __syncthreads();
  if (lp == 0 && i < total) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += part[q][li];
    const int t = i / C, c = i - t * C;
    const int r = t / S, sx = t - r * S;
    const int a = wide_u ? c : 0, b = wide_u ? 0 : c;
    float* d = dw + a * sa + b * sb + r * sr + sx * ss;
    *d = accumulate ? (*d + v) : v;
  }
}

}  // namespace

extern "C" int viai_conv2d_thin_supported(const viai_conv_geom* g) {
  if (!g) return 0;
  if (g->Cout == 1 && g->Cin % 4 == 0 && g->Cin >= 4) return 1;
  if (g->Cin == 1 && g->Cout % 4 == 0 && g->Cout >= 4 && g->R * g->S * g->Cout + g->Cout <= 12 * 1024) return 2;
  return 0;
}

extern "C" int viai_conv2d_thin_stats_supported(const viai_conv_geom* g) {
  static const bool off = [] { const char* e = getenv("VIAI_CIN1_STATS"); return e && e[0] == '0'; }();
  if (off || viai_conv2d_thin_supported(g) != 2) return 0;
  const bool shaped = (g->R == 3 && g->S == 3) || (g->R == 1 && g->S == 4);
  return (shaped && g->Cout <= 32 * CIN1_SV) ? 1 : 0;
}

static int conv2d_thin_impl(const viai_conv_geom* gp, const float* in, const float* wp, const float* bias, float* out,
                            double* stat_sum, double* stat_sumsq, viai_stream_t stream);

// Same operands as viai_conv2d_simt: wp is the [O][R][S][I] re-layout produced by viai_pack_weight.
extern "C" int viai_conv2d_thin(const viai_conv_geom* gp, const float* in, const float* wp, const float* bias, float* out,
                                viai_stream_t stream) {
  return conv2d_thin_impl(gp, in, wp, bias, out, nullptr, nullptr, stream);
}

extern "C" int viai_conv2d_thin_stats(const viai_conv_geom* gp, const float* in, const float* wp, const float* bias, float* out,
                                      double* stat_sum, double* stat_sumsq, viai_stream_t stream) {
  VIAI_REQUIRE(gp && stat_sum && stat_sumsq, "conv2d_thin_stats: null argument");
  VIAI_REQUIRE(viai_conv2d_thin_stats_supported(gp), "conv2d_thin_stats: unsupported geometry (Cin %d, Cout %d, %dx%d)", gp->Cin,
               gp->Cout, gp->R, gp->S);
  return conv2d_thin_impl(gp, in, wp, bias, out, stat_sum, stat_sumsq, stream);
}

static int conv2d_thin_impl(const viai_conv_geom* gp, const float* in, const float* wp, const float* bias, float* out,
                            double* stat_sum, double* stat_sumsq, viai_stream_t stream) {
  VIAI_REQUIRE(gp && in && wp && out, "conv2d_thin: null argument");
  const viai_conv_geom& g = *gp;
  const int kind = viai_conv2d_thin_supported(gp);
  VIAI_REQUIRE(kind != 0, "conv2d_thin: unsupported geometry (Cin %d, Cout %d)", g.Cin, g.Cout);
  VIAI_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(wp) & 15) == 0,
               "conv2d_thin: pointers must be 16-byte aligned");
  const int64_t M = (int64_t)g.N * g.Hout * g.Wout;
  if (stat_sum) {
    if (stat_sumsq == stat_sum + g.Cout) {
      VIAI_CUDA(cudaMemsetAsync(stat_sum, 0, 2 * sizeof(double) * g.Cout, STR(stream)));
    } else {
      VIAI_CUDA(cudaMemsetAsync(stat_sum, 0, sizeof(double) * g.Cout, STR(stream)));
      VIAI_CUDA(cudaMemsetAsync(stat_sumsq, 0, sizeof(double) * g.Cout, STR(stream)));
    }
  }
  if (M == 0) return VIAI_OK;
  if (kind == 1) {
    const int rc = launch_cout1_tma(g, in, wp, bias, out, STR(stream));
    if (rc != VIAI_ERR_UNSUPPORTED) return rc;
    const int blocks = (int)imin64(cdiv(M, 8), 16 * kNumSMs);
    conv_cout1_kernel<<<blocks, 256, 0, STR(stream)>>>(g, in, wp, bias, out);
  } else {
    size_t smem = sizeof(float) * (size_t)(g.R * g.S * g.Cout + g.Cout);
    const int64_t total = (int64_t)g.N * g.Hout * ((g.Wout + CIN1_PX - 1) / CIN1_PX) * 8;
    VIAI_REQUIRE(total < (int64_t)1 << 31, "conv2d_thin: tensor too large");
    const int blocks = (int)imin64(cdiv(total, 256), 16 * kNumSMs);
    if (stat_sum) {
      const size_t red = sizeof(float) * 8 * 8 * CIN1_SV * 8;       // the block reduction reuses the weight table's space
      if (smem < red) smem = red;
      if (g.R == 3 && g.S == 3) conv_cin1_kernel<3, 3, true><<<blocks, 256, smem, STR(stream)>>>(g, in, wp, bias, out, stat_sum, stat_sumsq);
      else conv_cin1_kernel<1, 4, true><<<blocks, 256, smem, STR(stream)>>>(g, in, wp, bias, out, stat_sum, stat_sumsq);
    } else if (g.R == 3 && g.S == 3) conv_cin1_kernel<3, 3, false><<<blocks, 256, smem, STR(stream)>>>(g, in, wp, bias, out, nullptr, nullptr);
    else if (g.R == 1 && g.S == 4) conv_cin1_kernel<1, 4, false><<<blocks, 256, smem, STR(stream)>>>(g, in, wp, bias, out, nullptr, nullptr);
    else {
      const int64_t tot = M * (g.Cout / 4);
      conv_cin1_generic_kernel<<<(int)imin64(cdiv(tot, 256), 16 * kNumSMs), 256, smem, STR(stream)>>>(g, in, wp, bias, out);
    }
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

namespace {
void wgrad_thin_plan(const viai_conv_geom& g, bool& wide_u, int64_t& M, int& C, int& planes, int64_t& blocks, int& ppb) {
  wide_u = g.Cin == 1;
  M = wide_u ? (int64_t)g.N * g.Hout * g.Wout : (int64_t)g.N * g.Hin * g.Win;   // pixels of the wide tensor
  C = wide_u ? g.Cout : g.Cin;
  planes = 256 / (C / 4);
  blocks = imin64(8 * kNumSMs, cdiv(M, (int64_t)planes * 4));
  if (blocks < 1) blocks = 1;
  ppb = (int)cdiv(M, blocks);
  if (ppb < 1) ppb = 1;
  blocks = M > 0 ? cdiv(M, ppb) : 0;
}
}  // namespace

extern "C" int64_t viai_wgrad_thin_workspace(const viai_conv_geom* g) {
  if (!viai_conv2d_wgrad_thin_supported(g)) return 0;
  bool wide_u; int64_t M, blocks; int C, planes, ppb;
  wgrad_thin_plan(*g, wide_u, M, C, planes, blocks, ppb);
  return (blocks > 0 ? blocks : 1) * (int64_t)g->R * g->S * C;
}

// Same meaning as viai_conv2d_wgrad_simt, for A == 1 or B == 1.  workspace: viai_wgrad_thin_workspace(g) floats.
extern "C" int viai_conv2d_wgrad_thin(const viai_conv_geom* gp, const float* U, const float* G, float* dw, int64_t sa, int64_t sb,
                                      int64_t sr, int64_t ss, int accumulate, float* workspace, viai_stream_t stream) {
  VIAI_REQUIRE(gp && U && G && dw && workspace, "conv2d_wgrad_thin: null argument");
  const viai_conv_geom& g = *gp;
  VIAI_REQUIRE(viai_conv2d_wgrad_thin_supported(gp), "conv2d_wgrad_thin: unsupported geometry (A %d, B %d)", g.Cout, g.Cin);
  VIAI_REQUIRE((reinterpret_cast<uintptr_t>(U) & 15) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0,
               "conv2d_wgrad_thin: pointers must be 16-byte aligned");
  cudaStream_t st = STR(stream);
  const int taps = g.R * g.S;
  bool wide_u; int64_t M, blocks; int C, planes, ppb;
  wgrad_thin_plan(g, wide_u, M, C, planes, blocks, ppb);
  VIAI_REQUIRE(M < (int64_t)1 << 31, "conv2d_wgrad_thin: tensor too large");
  if (M == 0 || (int64_t)g.N * g.Hout * g.Wout == 0) {
    if (!accumulate) {
      zero_dw_kernel<<<(g.Cout * g.Cin * taps + 255) / 256, 256, 0, st>>>(dw, g.Cout, g.Cin, g.R, g.S, sa, sb, sr, ss);
      VIAI_LAUNCHED();
    }
    return VIAI_OK;
  }
  const size_t smem = sizeof(float) * (size_t)planes * taps * C;
  VIAI_REQUIRE(smem <= 64 * 1024, "conv2d_wgrad_thin: reduction buffer too large");
  // wide pixels in flight per thread: VIAI_WGT_UN selects 2 (default) or 4.  Measured: 4 is SLOWER (C2 step 15.27 vs 15.19 ms):
  // the extra registers (80 -> 96 for the 3x3 kernels) cost a resident CTA, which outweighs the deeper per-thread queue
  static const int un_env = [] { const char* e = getenv("VIAI_WGT_UN"); return e ? atoi(e) : 2; }();
#define VIAI_WGT_LAUNCH_UN(WU, RR, SS, UNV)                                                                                 \
  do {                                                                                                                      \
    static bool attr_done = false;                                                                                          \
    if (!attr_done) {                                                                                                       \
      VIAI_CUDA(cudaFuncSetAttribute(wgrad_thin_kernel<WU, RR, SS, UNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); \
      attr_done = true;                                                                                                     \
    }                                                                                                                       \
    wgrad_thin_kernel<WU, RR, SS, UNV><<<(unsigned)blocks, 256, smem, st>>>(g, U, G, workspace, ppb);                       \
  } while (0)
#define VIAI_WGT_LAUNCH(WU, RR, SS)                                                                                         \
  do {                                                                                                                      \
    if (un_env == 4) VIAI_WGT_LAUNCH_UN(WU, RR, SS, 4); else VIAI_WGT_LAUNCH_UN(WU, RR, SS, 2);                             \
  } while (0)
  if (g.R == 3 && g.S == 3) { if (wide_u) VIAI_WGT_LAUNCH(true, 3, 3); else VIAI_WGT_LAUNCH(false, 3, 3); }
  else if (g.R == 1 && g.S == 4) { if (wide_u) VIAI_WGT_LAUNCH(true, 1, 4); else VIAI_WGT_LAUNCH(false, 1, 4); }
  else { if (wide_u) VIAI_WGT_LAUNCH(true, 0, 0); else VIAI_WGT_LAUNCH(false, 0, 0); }
#undef VIAI_WGT_LAUNCH
#undef VIAI_WGT_LAUNCH_UN
  VIAI_LAUNCHED();
  wgrad_thin_finish_kernel<<<(taps * C + 31) / 32, 256, 0, st>>>(workspace, (int)blocks, taps, C, g.S, wide_u ? 1 : 0, dw, sa, sb, sr, ss,
                                                               accumulate);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
