// WaveNet vocoder synthesis, FOLDED schedule (wavenet_vocoder/wavenet.py:237-364 `incremental_forward`, modules.py:162-210
// `ResidualConv1dGLU._forward`, conv.py:17-62, mixture.py:117-153): the same persistent cooperative kernel idea as
// wavenet_synth.cu (nC CTAs own fixed row slices of every layer, weights stream through a shared-memory double buffer, vectors
// cross CTAs as 64-bit {value, tag} words), but with ONE dependent exchange per layer instead of two.
//
// wavenet_synth.cu's critical path per layer is  gate rows (K = 3R + C) -> exchange h_l -> skip / residual rows (K = G/2) ->
// exchange x_{l+1}: 2L + 2 = 50 dependent L2 round trips per sample.  Here the current-time tap of layer l is folded through the
// residual 1x1 of layer l-1 at packing time (WaveNet.pack_for_synthesis_folded):
//     z_l = P'_l + M_l h_{l-1},                    M_l = r A_l Wo_{l-1}
//     P'_l = [old taps + conditioning]_l + N_l h_{l-2} + T_l x_{l-2} + const_l
// so a slot is:  wait for h_{l-1}  ->  ONE mat-vec over h_{l-1} whose rows are {skip rows l-1, residual rows l-1, M_l gate
// rows}  ->  tanh * sigmoid  ->  publish h_l;  then, while h_l travels, P'_{l+1} is evaluated from vectors that were published
// a whole slot earlier (h_{l-1}, x_{l-1}) and from taps / conditioning known since the previous sample.  The residual stream
// x_l still crosses CTAs (its consumer is T_{l+2} x_l, one slot later) but nobody waits for it.  L + 2 dependent exchanges per
// sample.  The exchanged vectors are double (h) / triple (x) buffered by slot parity, see the hazard notes at the buffers.
//
// CODE SIZE is a first-order term here: a slot is a few hundred instructions executed ONCE, so a body that does not fit the
// 32 KB instruction cache is fetched from L2 every slot.  The first version of this kernel (135 KB of SASS, mat-vec inlined five
// times and unrolled over a run-time batch) spent ~500 clocks per mat-vec iteration waiting for instructions
// (profiles/r02_wavenet_folded_phases.txt).  Hence: the batch is a template parameter, the mat-vec and the spin path of the
// tagged load are single out-of-line copies, strided loops are not unrolled, and profiling is a separate instantiation.
//
// WEIGHTS never touch shared memory: every weight is used exactly once per sample, by one thread, so each thread keeps the
// float4s of ITS slice of the next slot's two mat-vecs in registers (24 registers), loaded straight from L2 one slot ahead.
// Staging the 40 KB block per slot in shared memory (cp.async, then TMA bulk copies) made the mat-vec's shared-memory loads
// 5-10x slower while the block was landing (profiles/r02_wavenet_folded_phases.txt); only the vectors every row needs (h, x,
// old taps, conditioning) live in shared memory.
//
// Layer 0's input x_0 = fw * sample + fb is rank one in the previous output sample: every CTA rebuilds x_0 and layer 0's older
// taps locally from the last K-1 samples, and layer 0's gate rows are P'_0 + (A_0 fw) * sample.
#include "common.cuh"
#include "tc_common.cuh"
using namespace viai;
using namespace viai::tc;

namespace {

constexpr int NT = 512;            // threads per CTA
constexpr int NW = NT / 32;
constexpr int MAXB = 4;
constexpr int kSmemLimit = 220 * 1024;
__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }

struct Wn2Params {
  int L, R, G, S, C, K, O, B, T, nC;
  int layers_per_stack;
  int pairs, srows, orows, hrows;
  int K2, Kn;                          // G/2;  G/2 + R + (K-1) R + C
  const float* wl;                     // [L][nC][wpad]: see pack_for_synthesis_folded
  int64_t layer_stride, cta_stride;    // floats
  const float* wlast;                  // per CTA: skip rows of the last layer [srows][K2] + bias
  const float* first;                  // [R] weight, [R] bias
  const float* head1;                  // per CTA: [hrows][S] + [hrows] bias
  const float* head2;                  // [O][S] + [O] bias
  const float* cond;                   // (B, T, C)
  const float* uniforms;               // (T, B, O/3 + 1)
  const float* test_inputs;            // (B, Ttest) or null
  int Ttest;
  float log_scale_min;
  float* ring;                         // per layer: [ring_len_l][B][R] (layer 0's is unused)
  const int64_t* ring_off;
  unsigned long long* gbuf;            // [2][B][G/2]  gate outputs h_l, buffer l & 1
  unsigned long long* xnew;            // [3][B][R]    residual stream x_l, buffer l % 3
  unsigned long long* sbuf;            // [B][S]       relu(skips)
  unsigned long long* hbuf;            // [B][S]       relu(head1)
  float* out;
  float* logits;
};

__device__ long long g_wn2_prof[16];

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// {value, tag} in one 64-bit word: relaxed gpu-scope accesses deliver data and synchronisation in one L2 round trip.
__device__ __forceinline__ void put_tagged(unsigned long long* p, float v, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __noinline__ float get_tagged_spin(const unsigned long long* p, unsigned tag) {
  unsigned long long w;
  const long long t0 = clock64();
  do {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (clock64() - t0 > 4000000000LL) __trap();            // a protocol bug fails the launch instead of hanging the device
  } while ((unsigned)(w >> 32) != tag);
  return __uint_as_float((unsigned)w);
}
__device__ __forceinline__ float get_tagged(const unsigned long long* p, unsigned tag) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  if ((unsigned)(w >> 32) == tag) return __uint_as_float((unsigned)w);
  return get_tagged_spin(p, tag);
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

__device__ __forceinline__ float4 ldg128_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
constexpr int NWC = 2;             // float4s of the dependent mat-vec's weights a thread keeps in registers
constexpr int NWN = 4;             // ... of the independent mat-vec's

// rows x (4 K4) mat-vec for B columns, operands in shared memory (32-bit shared addresses): warp w -> row (w % rows), K slice
// (w / rows) of `kslices`.  Leaves the per-slice partial sums in part[(row * kslices + slice) * MAXB + b] after ONE barrier;
// readers add the slices with psum().  One out-of-line copy per batch size.
template <int B>
__device__ __noinline__ void matvec_part(uint32_t ws, uint32_t xs, int xstride, int rows, int K4, int kslices, float* part) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = (K4 + kslices - 1) / kslices;
#pragma unroll 1
  for (int row = warp % rows; row < rows; row += NW) {
    const int ks = (rows <= NW) ? warp / rows : 0;
    if (ks < kslices) {
      const int k0 = ks * chunk, k1 = min(K4, k0 + chunk);
      float acc[B];
#pragma unroll
      for (int b = 0; b < B; ++b) acc[b] = 0.f;
      const uint32_t wr = ws + (uint32_t)(row * K4) * 16u;
#pragma unroll 2
      for (int k = k0 + lane; k < k1; k += 32) {
        const float4 w = lds128(wr + 16u * (uint32_t)k);
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float4 v = lds128(xs + (uint32_t)(b * xstride + 4 * k) * 4u);
          acc[b] = fmaf(w.x, v.x, acc[b]); acc[b] = fmaf(w.y, v.y, acc[b]);
          acc[b] = fmaf(w.z, v.z, acc[b]); acc[b] = fmaf(w.w, v.w, acc[b]);
        }
      }
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const float s = warp_sum(acc[b]);
        if (lane == 0) part[(row * kslices + ks) * MAXB + b] = s;
      }
    }
    if (rows <= NW) break;
  }
  __syncthreads();
}
__device__ __forceinline__ int slices_of(int rows) { return rows <= NW ? NW / rows : 1; }
__device__ __forceinline__ float psum(const float* part, int row, int b, int kslices) {
  float s = 0.f;
#pragma unroll 4
  for (int ks = 0; ks < kslices; ++ks) s += part[(row * kslices + ks) * MAXB + b];
  return s;
}

// Issuer thread only: old taps of `layer` (> 0) at step t and the conditioning vector of step t -> xtail (+ b * xlen), as bulk
// copies completing on `bar`.  Returns nothing; the byte count is tap_bytes().
__device__ __noinline__ void issue_taps(const Wn2Params& p, int layer, int t, float* xtail, int xlen, uint64_t* bar) {
  const int d = 1 << (layer % p.layers_per_stack);
  const int rl = (p.K - 1) * d + 1;
  const float* ring = p.ring + p.ring_off[layer];
#pragma unroll 1
  for (int b = 0; b < p.B; ++b) {
    float* x = xtail + b * xlen;
    if (layer != 0) {
#pragma unroll 1
      for (int j = 0; j < p.K - 1; ++j) {
        const int back = (p.K - 1 - j) * d;
        const int slot = ((t - back) % rl + rl) % rl;
        bulk_load(x + j * p.R, ring + ((size_t)slot * p.B + b) * p.R, (uint32_t)p.R * 4u, bar);
      }
    }
    if (p.C > 0) bulk_load(x + (p.K - 1) * p.R, p.cond + ((size_t)b * p.T + t) * p.C, (uint32_t)p.C * 4u, bar);
  }
}
__device__ __forceinline__ uint32_t tap_bytes(const Wn2Params& p, int layer) {
  return (uint32_t)p.B * ((layer == 0 ? 0u : (uint32_t)(p.K - 1) * p.R * 4u) + (uint32_t)p.C * 4u);
}

template <int B, bool PROF>
__global__ void __launch_bounds__(NT, 1) wavenet_synth2_kernel(const __grid_constant__ Wn2Params p) {
  extern __shared__ __align__(16) float sm[];
  const int cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows1 = 2 * p.pairs, rows2 = p.srows + p.orows, rowsC = rows2 + rows1;
  const int xlen = p.Kn;                                                   // multiple of 4
  const int Hc = p.K > 1 ? p.K - 1 : 1;
  float* const xin0 = sm;                                                  // two [B][h_{l-1} | x_{l-1} | old taps | conditioning]
  float* const hst = xin0 + 2 * B * xlen;                                  // [B][K2]  h_{L-1} for the last skip rows
  float* const x0 = hst + B * p.K2;                                        // [B][R]   input of layer 0 at this step
  float* const skips = x0 + B * p.R;                                       // [srows][MAXB]
  float* const xown = skips + p.srows * MAXB;                              // [orows][MAXB]  this CTA's rows of x_{l-1}
  float* const Pn = xown + p.orows * MAXB;                                 // [rows1][MAXB]  P' of the next slot's gate rows
  const int maxrows = max(max(rowsC, p.hrows), p.O);
  float* const part = Pn + pad4(rows1) * MAXB;                             // [maxrows][NW][MAXB]
  float* const first = part + maxrows * NW * MAXB;                         // [2R]
  float* const h1w = first + 2 * p.R;
  float* const h2w = h1w + p.hrows * p.S + pad4(p.hrows);
  float* const wlast = h2w + p.O * p.S + pad4(p.O);                        // [srows][K2] + bias
  float* const vec = wlast + p.srows * p.K2 + pad4(p.srows);               // [B][S]
  float* const cur = vec + B * p.S;                                        // [MAXB]
  float* const curh = cur + MAXB;                                          // [Hc][MAXB] the last K-1 input samples
  uint64_t* const bars = reinterpret_cast<uint64_t*>(curh + pad4(Hc * MAXB));   // one mbarrier per xin buffer (taps, conditioning)
  long long* const profs = reinterpret_cast<long long*>(bars + 2);         // [16] (PROF only)
  uint32_t ph = 0;                                                         // the barriers' phase bits
  const unsigned per_sample = 2u * (unsigned)p.L + 2u;
  const float r2 = 0.70710678118654752440f;
  const size_t gstride = (size_t)B * p.K2, xstride = (size_t)B * p.R;
  const int ksC = slices_of(rowsC), ksN = slices_of(rows1), ksS = slices_of(p.srows), ksH = slices_of(p.hrows), ksO = slices_of(p.O);
  const int gate_threads = p.pairs * B, row_base = (gate_threads + 31) & ~31;
  const bool row_thread = tid >= row_base && tid < row_base + rows2 * B;

  long long tprev = 0;
  if (PROF) tprev = clock64();
#define WN2_MARK(slot)                          \
  do {                                          \
    if (PROF && cta == 0 && tid == 0) {         \
      const long long now = clock64();          \
      profs[slot] += now - tprev;               \
      tprev = now;                              \
    }                                           \
  } while (0)
  // ---- this thread's fixed slice of the two per-slot mat-vecs (weights in registers) ----
  const int K4c = p.K2 >> 2, K4n = p.Kn >> 2;
  const int rowC = warp % rowsC, chunkC = (K4c + ksC - 1) / ksC;
  const int kC0 = (warp / rowsC) * chunkC + lane, kC1 = (warp / rowsC < ksC) ? min(K4c, (warp / rowsC) * chunkC + chunkC) : 0;
  const int rowN = warp % rows1, chunkN = (K4n + ksN - 1) / ksN;
  const int kN0 = (warp / rows1) * chunkN + lane, kN1 = (warp / rows1 < ksN) ? min(K4n, (warp / rows1) * chunkN + chunkN) : 0;
  const int offN = rowsC * p.K2 + pad4(rowsC);                             // block layout: see pack_for_synthesis_folded
  const float* const blk0 = p.wl + (size_t)cta * p.cta_stride;
  float4 wc[NWC], wn[NWN];
  float bias_c = 0.f, const_n = 0.f;                                       // bC[row] of a row thread / cN[row] of a P' thread
  auto load_wc = [&](int layer) {
    const float* blk = blk0 + (size_t)layer * p.layer_stride;
#pragma unroll
    for (int i = 0; i < NWC; ++i) {
      const int k = kC0 + 32 * i;
      wc[i] = k < kC1 ? ldg128_stream(blk + ((size_t)rowC * K4c + k) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (row_thread) bias_c = __ldg(blk + rowsC * p.K2 + (tid - row_base) / B);
  };
  auto load_wn = [&](int layer) {
    const float* blk = blk0 + (size_t)layer * p.layer_stride + offN;
#pragma unroll
    for (int i = 0; i < NWN; ++i) {
      const int k = kN0 + 32 * i;
      wn[i] = k < kN1 ? ldg128_stream(blk + ((size_t)rowN * K4n + k) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < rows1 * B) const_n = __ldg(blk + rows1 * p.Kn + tid / B);
  };
  // partial sums of this thread's slice against the B columns at xs (shared address), reduced over the warp into `part`
  auto dot_regs = [&](const float4* w, int nw, int k0, int k1, int row, int ks, int kslices, uint32_t xs) {
    float acc[B];
#pragma unroll
    for (int b = 0; b < B; ++b) acc[b] = 0.f;
#pragma unroll
    for (int i = 0; i < NWN; ++i) {
      if (i < nw) {
        const int k = k0 + 32 * i;
        if (k < k1) {
#pragma unroll
          for (int b = 0; b < B; ++b) {
            const float4 v = lds128(xs + (uint32_t)(b * xlen + 4 * k) * 4u);
            acc[b] = fmaf(w[i].x, v.x, acc[b]); acc[b] = fmaf(w[i].y, v.y, acc[b]);
            acc[b] = fmaf(w[i].z, v.z, acc[b]); acc[b] = fmaf(w[i].w, v.w, acc[b]);
          }
        }
      }
    }
    if (k1 > 0) {                      // warp-uniform: this warp owns a (row, slice)
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const float s = warp_sum(acc[b]);
        if (lane == 0) part[(row * kslices + ks) * MAXB + b] = s;
      }
    }
  };

#pragma unroll 1
  for (int i = tid; i < 2 * p.R; i += NT) first[i] = p.first[i];
#pragma unroll 1
  for (int i = tid; i < p.hrows * p.S + p.hrows; i += NT) h1w[i] = p.head1[(size_t)cta * (p.hrows * p.S + pad4(p.hrows)) + i];
#pragma unroll 1
  for (int i = tid; i < p.O * p.S + p.O; i += NT) h2w[i] = p.head2[i];
#pragma unroll 1
  for (int i = tid; i < p.srows * p.K2 + p.srows; i += NT) wlast[i] = p.wlast[(size_t)cta * (p.srows * p.K2 + pad4(p.srows)) + i];
#pragma unroll 1
  for (int i = tid; i < 2 * B * xlen; i += NT) xin0[i] = 0.f;              // zero weights must not meet NaN bit patterns
  if (tid < MAXB) cur[tid] = 0.f;
  if (tid < Hc * MAXB) curh[tid] = 0.f;
  if (PROF && tid < 16) profs[tid] = 0;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  // layer 0's rank-one coefficients uc = A_0 fw (block 0), kept for the whole launch by the gate threads
  float uc_a = 0.f, uc_g = 0.f;
  if (tid < gate_threads) {
    const float* uc = blk0 + offN + rows1 * p.Kn + pad4(rows1);
    uc_a = __ldg(uc + 2 * (tid / B));
    uc_g = __ldg(uc + 2 * (tid / B) + 1);
  }

  // Old taps and conditioning vectors (needed by every row) travel as bulk (TMA) copies issued by ONE thread.
  const bool issuer = tid == NT - 32;
  auto xtail = [&](int buf) { return xin0 + buf * B * xlen + p.K2 + p.R; };
  // layer 0's old taps are rebuilt from the last input samples (zeros before t = 0) by all threads
  auto local_taps0 = [&](int t, int buf) {
#pragma unroll 1
    for (int j = 0; j < p.K - 1; ++j) {
      const int tau = t - (p.K - 1 - j);
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const float c = tau >= 0 ? curh[(tau % Hc) * MAXB + b] : 0.f;
        float* x = xtail(buf) + b * xlen + j * p.R;
#pragma unroll 1
        for (int r = tid; r < p.R; r += NT) x[r] = tau >= 0 ? fmaf(first[r], c, first[p.R + r]) : 0.f;
      }
    }
  };
  // P' of a layer's gate rows: the independent mat-vec (weights already in wn) over xin[buf]; then wn is refilled for `next`
  auto noncritical = [&](int buf, int next) {
    dot_regs(wn, NWN, kN0, kN1, rowN, warp / rows1, ksN, smem_u32(xin0 + buf * B * xlen));
    WN2_MARK(12);
    const float cn = const_n;
    if (next >= 0) load_wn(next);
    WN2_MARK(13);
    __syncthreads();
    WN2_MARK(14);
    if (tid < rows1 * B) {
      const int row = tid / B, b = tid - row * B;
      Pn[row * MAXB + b] = psum(part, row, b, ksN) + cn;
    }
  };

  __syncthreads();
  {                                                                         // prologue: P'_0 of step 0 lives in block L-1
    load_wn(p.L - 1);
    if (issuer) {
      fence_proxy_async();
      mbar_expect_tx(&bars[1], tap_bytes(p, 0));
      issue_taps(p, 0, 0, xtail(1), xlen, &bars[1]);
    }
    local_taps0(0, 1);
    mbar_wait(&bars[1], 0);
    ph ^= 2u;
    __syncthreads();
    noncritical(1, 0);
    load_wc(1);
    __syncthreads();
    if (issuer) {
      fence_proxy_async();
      mbar_expect_tx(&bars[0], tap_bytes(p, 1));
      issue_taps(p, 1, 0, xtail(0), xlen, &bars[0]);
    }
  }
  int buf = 0;
#pragma unroll 1
  for (int t = 0; t < p.T; ++t) {
    if (p.test_inputs != nullptr && t < p.Ttest) {
      if (tid < B) cur[tid] = p.test_inputs[(size_t)tid * p.Ttest + t];
      __syncthreads();
    }
    if (tid < B) curh[(t % Hc) * MAXB + tid] = cur[tid];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const float c = cur[b];
#pragma unroll 1
      for (int r = tid; r < p.R; r += NT) x0[b * p.R + r] = fmaf(first[r], c, first[p.R + r]);
    }
    const unsigned tag0 = 1u + (unsigned)t * per_sample;
#pragma unroll 1
    for (int l = 0; l < p.L; ++l) {
      const unsigned tag_h = tag0 + 2u * (unsigned)l, tag_x = tag_h + 1u;
      WN2_MARK(0);
      mbar_wait(&bars[buf], (ph >> buf) & 1u);
      ph ^= 1u << buf;
      __syncthreads();                 // the tail of xin[buf] has landed; Pn / x0 / curh of the previous phase are visible
      WN2_MARK(1);
      float* x = xin0 + buf * B * xlen;
      float bc = 0.f;
      // ---- dependent part: h_{l-1} -> {skip rows l-1, residual rows l-1, gate rows l} ----
      if (l > 0) {
        const unsigned long long* g = p.gbuf + (size_t)((l - 1) & 1) * gstride;
#pragma unroll
        for (int b = 0; b < B; ++b) {
#pragma unroll 1
          for (int k = tid; k < p.K2; k += NT) x[b * xlen + k] = get_tagged(g + b * p.K2 + k, tag_h - 2u);
        }
        __syncthreads();
        WN2_MARK(2);
        dot_regs(wc, NWC, kC0, kC1, rowC, warp / rowsC, ksC, smem_u32(x));
        bc = bias_c;
        if (l + 1 < p.L) load_wc(l + 1);                                   // next slot's slice, a whole slot ahead of its use
        __syncthreads();
        WN2_MARK(3);
      } else {
        load_wc(1);
      }
      if (tid < gate_threads) {
        const int pr = tid / B, b = tid - pr * B;
        float a = Pn[(2 * pr) * MAXB + b], g = Pn[(2 * pr + 1) * MAXB + b];
        if (l > 0) {
          a += psum(part, rows2 + 2 * pr, b, ksC);
          g += psum(part, rows2 + 2 * pr + 1, b, ksC);
        } else {
          a = fmaf(uc_a, cur[b], a);
          g = fmaf(uc_g, cur[b], g);
        }
        // h buffers alternate with the slot parity: a CTA can only be one slot ahead of the slowest reader (it needs every
        // CTA's h_{l+1} before it can publish h_{l+2}, and a CTA publishes h_{l+1} after reading h_l)
        put_tagged(p.gbuf + (size_t)(l & 1) * gstride + (size_t)b * p.K2 + cta * p.pairs + pr, tanhf(a) * (1.f / (1.f + expf(-g))), tag_h);
      } else if (l > 0 && row_thread) {
        const int i = tid - row_base;
        const int row = i / B, b = i - row * B;
        const float v = psum(part, row, b, ksC) + bc;
        if (row < p.srows) {
          skips[row * MAXB + b] = (l == 1) ? v : (skips[row * MAXB + b] + v) * r2;
        } else {
          const int j = row - p.srows, r = cta * p.orows + j;
          const float xprev = (l == 1) ? x0[b * p.R + r] : xown[j * MAXB + b];
          const float xo = (v + xprev) * r2;
          xown[j * MAXB + b] = xo;
          const int d = 1 << (l % p.layers_per_stack);
          const int rl = (p.K - 1) * d + 1;
          p.ring[p.ring_off[l] + ((size_t)(t % rl) * B + b) * p.R + r] = xo;              // taps of later samples
          // x buffers rotate over three slots: x_l is read during slot l+1 AFTER that slot's h was published, so a writer two
          // slots ahead could still race a slow reader; three slots ahead it has seen the reader's h_{l+2}
          if (l + 2 < p.L) put_tagged(p.xnew + (size_t)(l % 3) * xstride + (size_t)b * p.R + r, xo, tag_x);
        }
      }
      WN2_MARK(4);
      // ---- old taps / conditioning of the next slot's independent part into the other buffer ----
      if (l + 1 < p.L) {
        const int nl2 = (l + 2) % p.L, nt2 = t + (l + 2 >= p.L ? 1 : 0);
        if (issuer) {
          fence_proxy_async();
          mbar_expect_tx(&bars[buf ^ 1], nt2 < p.T ? tap_bytes(p, nl2) : 0u);
          if (nt2 < p.T) issue_taps(p, nl2, nt2, xtail(buf ^ 1), xlen, &bars[buf ^ 1]);
        }
        if (nl2 == 0 && nt2 < p.T) local_taps0(nt2, buf ^ 1);
      }                                // slot L-1: layer 1's taps of step t+1 may be this step's outputs -> after the fence in the head
      WN2_MARK(5);
      // ---- independent part: P' of layer (l+1) % L ----
      const int nt = t + (l + 1 == p.L ? 1 : 0);
      if (nt < p.T) {
        if (l <= 1) {
#pragma unroll
          for (int b = 0; b < B; ++b) {
#pragma unroll 1
            for (int r = tid; r < p.R; r += NT) x[b * xlen + p.K2 + r] = x0[b * p.R + r];
          }
        } else if (l + 1 < p.L) {
          const unsigned long long* xs = p.xnew + (size_t)((l - 1) % 3) * xstride;
#pragma unroll
          for (int b = 0; b < B; ++b) {
#pragma unroll 1
            for (int r = tid; r < p.R; r += NT) x[b * xlen + p.K2 + r] = get_tagged(xs + b * p.R + r, tag_x - 2u);
          }
        }
        __syncthreads();               // also: every reader of `part` above is done
        WN2_MARK(6);
        noncritical(buf, (l + 1) % p.L);
        WN2_MARK(7);
      }
      buf ^= 1;
    }
    WN2_MARK(0);
    // ---- skip rows of the last layer ----
    {
      const unsigned long long* g = p.gbuf + (size_t)((p.L - 1) & 1) * gstride;
#pragma unroll 1
      for (int i = tid; i < B * p.K2; i += NT) hst[i] = get_tagged(g + i, tag0 + 2u * (unsigned)(p.L - 1));
      __syncthreads();
      matvec_part<B>(smem_u32(wlast), smem_u32(hst), p.K2, p.srows, p.K2 >> 2, ksS, part);
      if (tid < p.srows * B) {
        const int row = tid / B, b = tid - row * B;
        const float v = psum(part, row, b, ksS) + wlast[p.srows * p.K2 + row];
        const float s = (p.L == 1) ? v : (skips[row * MAXB + b] + v) * r2;
        put_tagged(p.sbuf + (size_t)b * p.S + cta * p.srows + row, fmaxf(s, 0.f), tag0 + per_sample - 2u);
      }
    }
    WN2_MARK(8);
    // ---- output head: ReLU, 1x1 (S -> S), ReLU, 1x1 (S -> O) ----
#pragma unroll 1
    for (int i = tid; i < B * p.S; i += NT) vec[i] = get_tagged(p.sbuf + i, tag0 + per_sample - 2u);
    __syncthreads();
    WN2_MARK(9);
    matvec_part<B>(smem_u32(h1w), smem_u32(vec), p.S, p.hrows, p.S >> 2, ksH, part);
    __threadfence();                   // release: this sample's ring-buffer stores are visible before the tagged words below
    __syncthreads();                   // (every thread's fence precedes any thread's publication)
    if (tid < p.hrows * B) {
      const int row = tid / B, b = tid - row * B;
      put_tagged(p.hbuf + (size_t)b * p.S + cta * p.hrows + row, fmaxf(psum(part, row, b, ksH) + h1w[p.hrows * p.S + row], 0.f),
                 tag0 + per_sample - 1u);
    }
    __syncthreads();                   // vec is refilled below
#pragma unroll 1
    for (int i = tid; i < B * p.S; i += NT) vec[i] = get_tagged(p.hbuf + i, tag0 + per_sample - 1u);
    __threadfence();                   // acquire: every CTA's ring-buffer stores of this sample are visible from here on
    __syncthreads();
    WN2_MARK(10);
    if (t + 1 < p.T && issuer) {
      fence_proxy_async();
      mbar_expect_tx(&bars[buf], tap_bytes(p, 1));
      issue_taps(p, 1, t + 1, xtail(buf), xlen, &bars[buf]);
    }
    matvec_part<B>(smem_u32(h2w), smem_u32(vec), p.S, p.O, p.S >> 2, ksO, part);
    // ---- sample from the discretised mixture of logistics (every CTA computes the same value) ----
    if (tid < B) {
      const int b = tid, nm = p.O / 3;
      const float* u = p.uniforms + ((size_t)t * B + b) * (nm + 1);
      int arg = 0;
      float best = -INFINITY;
#pragma unroll 1
      for (int m = 0; m < nm; ++m) {
        const float v = psum(part, m, b, ksO) + h2w[p.O * p.S + m] - logf(-logf(u[m]));
        if (v > best) { best = v; arg = m; }
      }
      const float mean = psum(part, nm + arg, b, ksO) + h2w[p.O * p.S + nm + arg];
      const float ls = fmaxf(psum(part, 2 * nm + arg, b, ksO) + h2w[p.O * p.S + 2 * nm + arg], p.log_scale_min);
      const float ul = u[nm];
      float xs = mean + expf(ls) * (logf(ul) - logf(1.f - ul));
      xs = fminf(fmaxf(xs, -1.f), 1.f);
      cur[b] = xs;
      if (cta == 0) p.out[(size_t)b * p.T + t] = xs;
    }
    if (cta == 0 && p.logits != nullptr) {
#pragma unroll 1
      for (int i = tid; i < p.O * B; i += NT) {
        const int o = i / B, b = i - o * B;
        p.logits[((size_t)b * p.T + t) * p.O + o] = psum(part, o, b, ksO) + h2w[p.O * p.S + o];
      }
    }
    __syncthreads();
    WN2_MARK(11);
  }
#undef WN2_MARK
  if (PROF && cta == 0 && tid < 16) g_wn2_prof[tid] = profs[tid];
}

int64_t wn2_smem_bytes(int R, int G, int S, int C, int K, int O, int B, int nC) {
  const int pairs = (G / 2) / nC, srows = S / nC, orows = R / nC, hrows = S / nC;
  const int K2 = G / 2, Kn = K2 + R + (K - 1) * R + C;
  const int rows1 = 2 * pairs, rowsC = srows + orows + rows1;
  int maxrows = rowsC > hrows ? rowsC : hrows;
  if (O > maxrows) maxrows = O;
  const int Hc = K > 1 ? K - 1 : 1;
  int64_t f = 2 * (int64_t)B * Kn + (int64_t)B * K2 + (int64_t)B * R + srows * MAXB + orows * MAXB +
              pad4(rows1) * MAXB + (int64_t)maxrows * NW * MAXB + 2 * R + (hrows * S + pad4(hrows)) + (O * S + pad4(O)) +
              (srows * K2 + pad4(srows)) + (int64_t)B * S + MAXB + pad4(Hc * MAXB) + 4 + 32;
  return f * 4 + 64;
}

template <int B, bool PROF>
cudaError_t launch_wn2(const Wn2Params& p, size_t smem, cudaStream_t stream) {
  auto k = wavenet_synth2_kernel<B, PROF>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<Wn2Params*>(&p)};
  return cudaLaunchCooperativeKernel((void*)k, dim3(p.nC), dim3(NT), args, smem, stream);
}

}  // namespace

// Number of cooperating CTAs of the folded kernel for a configuration, 0 if unsupported (use viai_wavenet_synth then).
extern "C" int viai_wavenet2_num_ctas(int L, int R, int G, int S, int C, int K, int O, int B) {
  if (R % 4 || (G / 2) % 4 || S % 4 || C % 4 || G % 2 || O % 3 || B < 1 || B > MAXB || K < 1 || L < 3) return 0;
  int dev = 0, sms = kNumSMs;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int n = sms < 128 ? sms : 128; n >= 1; --n) {
    if ((G / 2) % n || S % n || R % n) continue;
    const int pairs = (G / 2) / n, srows = S / n, orows = R / n;
    if (pairs * B + 32 + (srows + orows) * B > NT) continue;       // the slot's post-processing is one thread per (row, column)
    const int rows1 = 2 * pairs, rowsC = srows + orows + rows1;
    if (rowsC > NW) continue;                                      // one warp per (row, K slice) of the per-slot mat-vecs,
    const int K4c = (G / 2) / 4, K4n = (G / 2 + R + (K - 1) * R + C) / 4;
    const int ksC = NW / rowsC, ksN = NW / rows1;
    if ((K4c + ksC - 1) / ksC > 32 * NWC || (K4n + ksN - 1) / ksN > 32 * NWN) continue;   // whose slice fits the register budget
    if (wn2_smem_bytes(R, G, S, C, K, O, B, n) <= kSmemLimit) return n;
  }
  return 0;
}

// Same contract as viai_wavenet_synth, with the blocks of WaveNet.pack_for_synthesis_folded: packed_layers [L][nC][...],
// last [nC][...]; gbuf holds 2 * B * (G/2) and xchg 3 * B * R 64-bit words (zero-initialised, 8-byte aligned), sbuf / hbuf
// B * S each.
extern "C" int viai_wavenet_synth2(int L, int layers_per_stack, int R, int G, int S, int C, int K, int O, int B, int T, int nC,
                                   const float* packed_layers, const float* last, const float* first, const float* head1,
                                   const float* head2, const float* cond, const float* uniforms, const float* test_inputs, int Ttest,
                                   float log_scale_min, float* ring, const int64_t* ring_off, float* gbuf, float* sbuf,
                                   float* hbuf, unsigned* xchg, float* out, float* logits, viai_stream_t stream) {
  VIAI_REQUIRE(packed_layers && last && first && head1 && head2 && cond && uniforms && ring && ring_off && gbuf && sbuf && hbuf && xchg && out,
               "wavenet_synth2: null argument");
  VIAI_REQUIRE(((reinterpret_cast<uintptr_t>(gbuf) | reinterpret_cast<uintptr_t>(sbuf) | reinterpret_cast<uintptr_t>(hbuf) |
                 reinterpret_cast<uintptr_t>(xchg)) & 7) == 0, "wavenet_synth2: exchange buffers must be 8-byte aligned");
  VIAI_REQUIRE((int64_t)T * (2 * L + 2) < 4000000000LL, "wavenet_synth2: T too large for the 32-bit stage tags");
  VIAI_REQUIRE(nC >= 1 && nC == viai_wavenet2_num_ctas(L, R, G, S, C, K, O, B), "wavenet_synth2: nC must come from viai_wavenet2_num_ctas");
  VIAI_REQUIRE(layers_per_stack >= 1 && L % layers_per_stack == 0 && T >= 0, "wavenet_synth2: bad layer configuration");
  if (T == 0) return VIAI_OK;
  Wn2Params p;
  memset(&p, 0, sizeof(p));
  p.L = L; p.R = R; p.G = G; p.S = S; p.C = C; p.K = K; p.O = O; p.B = B; p.T = T; p.nC = nC;
  p.layers_per_stack = layers_per_stack;
  p.pairs = (G / 2) / nC; p.srows = S / nC; p.orows = R / nC; p.hrows = S / nC;
  p.K2 = G / 2; p.Kn = p.K2 + R + (K - 1) * R + C;
  const int rows1 = 2 * p.pairs, rowsC = p.srows + p.orows + rows1;
  p.cta_stride = rowsC * p.K2 + pad4(rowsC) + rows1 * p.Kn + 2 * pad4(rows1);
  p.layer_stride = p.cta_stride * nC;
  p.wl = packed_layers; p.wlast = last; p.first = first; p.head1 = head1; p.head2 = head2; p.cond = cond; p.uniforms = uniforms;
  p.test_inputs = test_inputs; p.Ttest = test_inputs ? Ttest : 0; p.log_scale_min = log_scale_min;
  p.ring = ring; p.ring_off = ring_off; p.out = out; p.logits = logits;
  p.gbuf = reinterpret_cast<unsigned long long*>(gbuf); p.sbuf = reinterpret_cast<unsigned long long*>(sbuf);
  p.hbuf = reinterpret_cast<unsigned long long*>(hbuf); p.xnew = reinterpret_cast<unsigned long long*>(xchg);
  const char* pe = getenv("VIAI_WN2_PROF");
  const bool prof = pe && pe[0] == '1';
  const size_t smem = (size_t)wn2_smem_bytes(R, G, S, C, K, O, B, nC);
  cudaError_t e = cudaErrorInvalidValue;
  switch (B * 2 + (prof ? 1 : 0)) {
    case 2: e = launch_wn2<1, false>(p, smem, STR(stream)); break;
    case 3: e = launch_wn2<1, true>(p, smem, STR(stream)); break;
    case 4: e = launch_wn2<2, false>(p, smem, STR(stream)); break;
    case 5: e = launch_wn2<2, true>(p, smem, STR(stream)); break;
    case 6: e = launch_wn2<3, false>(p, smem, STR(stream)); break;
    case 7: e = launch_wn2<3, true>(p, smem, STR(stream)); break;
    case 8: e = launch_wn2<4, false>(p, smem, STR(stream)); break;
    case 9: e = launch_wn2<4, true>(p, smem, STR(stream)); break;
  }
  VIAI_CUDA(e);
  viai::g_launches.fetch_add(1, std::memory_order_relaxed);
  return VIAI_OK;
}

// Debug aid (VIAI_WN2_PROF=1 selects the profiling instantiation): clocks CTA 0 spent per phase in the last launch.  [0] rest of
// slot / loop overhead, [1] operand wait + barrier, [2] wait for h_{l-1}, [3] dependent mat-vec, [4] gate + publish, [5] prefetch
// issue, [6] x_{l-1} read, [7] independent mat-vec, [8] last skip rows, [9] wait relu(skips), [10] head 1 + wait, [11] head 2 +
// sampler.
extern "C" int viai_wavenet2_profile(long long* out16) {
  VIAI_REQUIRE(out16, "wavenet2_profile: null argument");
  VIAI_CUDA(cudaDeviceSynchronize());
  VIAI_CUDA(cudaMemcpyFromSymbol(out16, g_wn2_prof, 16 * sizeof(long long)));
  return VIAI_OK;
}
