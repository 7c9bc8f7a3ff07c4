// WaveNet vocoder synthesis (wavenet_vocoder/wavenet.py:237-364 `incremental_forward`, modules.py:162-210
// `ResidualConv1dGLU._forward`, conv.py:17-62 linearised dilated convolution, mixture.py:117-153 sampling) as ONE persistent
// cooperative kernel: the whole T-step autoregressive loop, all layers, the output head and the mixture-of-logistics sampler
// run on the device without returning to the host.
//
// The per-sample work is 2*L+2 strictly dependent mat-vec stages (L = 24 layers -> 50), ~100 MB of fp32 weights touched per
// sample: latency-bound, not roofline-bound.  Decomposition:
//   * nC CTAs (<= SM count, co-resident through a cooperative launch); every CTA owns a fixed slice of the output rows of
//     every stage: (gate/2)/nC tanh/sigmoid row PAIRS of the dilated convolution (+ the 1x1 conditioning convolution, folded
//     into the same dot product), skip/nC skip rows and res/nC residual rows, skip/nC rows of the first head layer;
//   * the weight slice of the NEXT layer streams from L2 into a shared-memory double buffer with cp.async while the current
//     layer computes and waits, the dilated taps x[t-d], x[t-2d] (known in advance) are prefetched the same way; only the
//     512-float current activation crosses CTAs on the critical path;
//   * per-layer ring buffers of 2d+1 time slots live in global memory (L2-resident) and are indexed modulo -- nothing is
//     shifted (the reference clones the whole buffer every step, conv.py:39);
//   * there is NO grid barrier: every cross-CTA vector (gate outputs, residual outputs, skip sums, head activations) travels
//     as 64-bit {value, stage tag} words written with st.release and polled with ld.acquire, so that one L2 round trip
//     delivers both the data and the synchronisation (a counter barrier costs an atomic round trip + a flag round trip per
//     stage, and there are 2*L+2 = 50 stages per sample).  The alternation of the exchanges makes single buffers safe: a CTA
//     can only overwrite a buffer after reading a later vector whose producers had all consumed the earlier one;
//   * skip accumulators stay in the owning CTA's shared memory for the whole sample; the 30-row output layer and the sampler
//     are evaluated redundantly by every CTA so that the next input needs no extra exchange.
#include "common.cuh"
using namespace viai;

namespace {

constexpr int NT = 512;            // threads per CTA
constexpr int NW = NT / 32;
constexpr int MAXB = 4;
__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }

struct WnParams {
  // dimensions
  int L, R, G, S, C, K, O, B, T, nC;
  int layers_per_stack;
  int pairs, srows, orows, hrows;      // per-CTA rows: gate pairs, skip rows, residual rows, head-1 rows
  int K1, K2;                          // stage-1 / stage-2 dot lengths: K*R + C, G/2
  // packed weights (wavenet.py pack_for_synthesis): per layer per CTA  [2*pairs][K1] [bias, padded to 4] [srows+orows][K2]
  // [bias, padded to 4]; every block starts 16-byte aligned
  const float* wl;
  int64_t layer_stride, cta_stride;    // floats
  const float* first;                  // [R] weight, [R] bias
  const float* head1;                  // per CTA: [hrows][S] + [hrows] bias
  const float* head2;                  // [O][S] + [O] bias
  const float* cond;                   // (B, T, C) upsampled conditioning
  const float* uniforms;               // (T, B, O/3 + 1)
  const float* test_inputs;            // (B, Ttest) or null
  int Ttest;
  float log_scale_min;
  // state (global, zero-initialised by the caller)
  float* ring;                         // per layer: [ring_len_l][B][R], offsets in ring_off
  const int64_t* ring_off;
  unsigned long long* gbuf;            // [B][G/2]  tagged words {value, stage tag}: gate outputs
  unsigned long long* sbuf;            // [B][S]    relu(skips)
  unsigned long long* hbuf;            // [B][S]    relu(head1)
  unsigned long long* xnew;            // [B][R]    residual output of the previous layer = newest tap of the next one
  float* out;                          // (B, T)
  float* logits;                       // (B, T, O) or null
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Tagged exchange words: low 32 bits = float value, high 32 bits = stage tag (0 = never written; buffers start zeroed).
// Value and tag share one 64-bit word, so the exchange itself needs no ordering between locations: relaxed gpu-scope accesses
// (served by L2) are enough and avoid a release fence per stage.  The plain ring-buffer stores, which other CTAs read one or
// more samples later, are ordered by ONE fence pair per sample around the last exchange (see the head stage).
__device__ __forceinline__ void put_tagged(unsigned long long* p, float v, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ float get_tagged(const unsigned long long* p, unsigned tag) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  if ((unsigned)(w >> 32) != tag) {
    const long long t0 = clock64();
    do {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
      if (clock64() - t0 > 4000000000LL) __trap();          // a protocol bug fails the launch instead of hanging the device
    } while ((unsigned)(w >> 32) != tag);
  }
  return __uint_as_float((unsigned)w);
}

// rows x K mat-vec for B batch columns: warp w -> row (w % rows), K slice (w / rows); result[row][b] in `res` (shared).
__device__ __forceinline__ void matvec(const float* __restrict__ W, const float* __restrict__ x, int xstride, int rows, int K, int B,
                                       float* part, float* res) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kslices = NW / rows > 0 ? NW / rows : 1;
  const int K4 = K >> 2;
  for (int row = warp % rows; row < rows; row += NW) {       // rows > NW: a warp takes several rows, one K slice
    const int ks = (rows <= NW) ? warp / rows : 0;
    if (ks < kslices) {
      const int chunk = (K4 + kslices - 1) / kslices;
      const int k0 = ks * chunk, k1 = min(K4, k0 + chunk);
      float acc[MAXB];
#pragma unroll
      for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
      const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)row * K);
      for (int k = k0 + lane; k < k1; k += 32) {
        const float4 w = w4[k];
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
          if (b < B) {
            const float4 v = *reinterpret_cast<const float4*>(x + (size_t)b * xstride + 4 * k);
            acc[b] = fmaf(w.x, v.x, acc[b]); acc[b] = fmaf(w.y, v.y, acc[b]);
            acc[b] = fmaf(w.z, v.z, acc[b]); acc[b] = fmaf(w.w, v.w, acc[b]);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < MAXB; ++b) {
        if (b < B) {
          const float s = warp_sum(acc[b]);
          if (lane == 0) part[(row * kslices + ks) * MAXB + b] = s;
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * B; i += NT) {
    const int row = i / B, b = i - row * B;
    float s = 0.f;
    for (int ks = 0; ks < kslices; ++ks) s += part[(row * kslices + ks) * MAXB + b];
    res[row * MAXB + b] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT, 1) wavenet_synth_kernel(const WnParams p) {
  extern __shared__ __align__(16) float sm[];
  const int cta = blockIdx.x, tid = threadIdx.x;
  const int rows1 = 2 * p.pairs, rows2 = p.srows + p.orows;
  const int wpad = rows1 * p.K1 + pad4(rows1) + rows2 * p.K2 + pad4(rows2);  // floats of one (layer, CTA) weight block
  const int xlen = (p.K1 + 3) & ~3;
  float* wbuf[2] = {sm, sm + wpad};
  float* xbuf[2] = {sm + 2 * wpad, sm + 2 * wpad + p.B * xlen};            // [B][K1]: taps oldest..newest then conditioning
  float* gsm = xbuf[1] + p.B * xlen;                                       // [B][K2]
  float* skips = gsm + p.B * p.K2;                                         // [srows][MAXB] running skip sums of this CTA
  float* part = skips + p.srows * MAXB;                                    // [max rows][kslices][MAXB]
  const int maxrows = max(max(rows1, rows2), max(p.hrows, p.O));
  float* res = part + maxrows * NW * MAXB;                                 // [max rows][MAXB]
  float* first = res + maxrows * MAXB;                                     // [2R]
  float* h1w = first + 2 * p.R;                                            // [hrows][S] + [hrows]
  float* h2w = h1w + p.hrows * p.S + pad4(p.hrows);                        // [O][S] + [O]
  float* vec = h2w + p.O * p.S + pad4(p.O);                                    // [B][S] staging of relu(skips)/relu(head1)
  float* cur = vec + p.B * p.S;                                            // [MAXB] current input sample
  const unsigned per_sample = 2u * (unsigned)p.L + 2u;             // exchanges per sample; tag = 1 + t * per_sample + index
  const float r2 = 0.70710678118654752440f;

  for (int i = tid; i < 2 * p.R; i += NT) first[i] = p.first[i];
  for (int i = tid; i < p.hrows * p.S + p.hrows; i += NT) h1w[i] = p.head1[(size_t)cta * (p.hrows * p.S + pad4(p.hrows)) + i];
  for (int i = tid; i < p.O * p.S + p.O; i += NT) h2w[i] = p.head2[i];
  if (tid < MAXB) cur[tid] = 0.f;                                          // initial input: zeros (wavenet.py:307-315)

  auto prefetch = [&](int layer, int t, int buf) {
    // weights of (layer, this CTA)
    const float* src = p.wl + (size_t)layer * p.layer_stride + (size_t)cta * p.cta_stride;
    for (int i = tid * 4; i < wpad; i += NT * 4) cp_async16(wbuf[buf] + i, src + i);
    // taps older than the current step and the conditioning vector of step t
    const int d = 1 << (layer % p.layers_per_stack);
    const int rl = (p.K - 1) * d + 1;
    float* ring = p.ring + p.ring_off[layer];
    for (int b = 0; b < p.B; ++b) {
      float* x = xbuf[buf] + b * xlen;
      for (int j = 0; j < p.K - 1; ++j) {
        const int back = (p.K - 1 - j) * d;                                 // tap j looks `back` steps into the past
        const int slot = ((t - back) % rl + rl) % rl;
        const float* s = ring + ((size_t)slot * p.B + b) * p.R;
        for (int i = tid * 4; i < p.R; i += NT * 4) cp_async16(x + j * p.R + i, s + i);
      }
      const float* c = p.cond + ((size_t)b * p.T + t) * p.C;
      for (int i = tid * 4; i < p.C; i += NT * 4) cp_async16(x + p.K * p.R + i, c + i);
    }
  };

  __syncthreads();
  if (p.T > 0) prefetch(0, 0, 0);
  int buf = 0;
  for (int t = 0; t < p.T; ++t) {
    if (p.test_inputs != nullptr && t < p.Ttest) {
      __syncthreads();
      if (tid < p.B) cur[tid] = p.test_inputs[(size_t)tid * p.Ttest + t];
      __syncthreads();
    }
    const unsigned tag0 = 1u + (unsigned)t * per_sample;
    for (int l = 0; l < p.L; ++l) {
      const unsigned tag_g = tag0 + 2u * (unsigned)l, tag_x = tag_g + 1u;   // gate exchange of layer l, residual exchange l -> l+1
      const int d = 1 << (l % p.layers_per_stack);
      const int rl = (p.K - 1) * d + 1;
      float* ring = p.ring + p.ring_off[l];
      float* x = xbuf[buf];
      // newest tap: the input of layer l at time t
      if (l == 0) {
        for (int i = tid; i < p.B * p.R; i += NT) {
          const int b = i / p.R, r = i - b * p.R;
          const float h = fmaf(first[r], cur[b], first[p.R + r]);
          x[b * xlen + (p.K - 1) * p.R + r] = h;
          if (cta == 0) ring[((size_t)(t % rl) * p.B + b) * p.R + r] = h;
        }
      } else {
        for (int i = tid; i < p.B * p.R; i += NT) {
          const int b = i / p.R, r = i - b * p.R;
          x[b * xlen + (p.K - 1) * p.R + r] = get_tagged(p.xnew + i, tag_x - 2u);     // written by layer l-1's stage 2
        }
      }
      cp_async_wait_all();
      __syncthreads();
      const float* W1 = wbuf[buf];
      const float* b1 = W1 + rows1 * p.K1;
      const float* W2 = b1 + pad4(rows1);
      const float* b2 = W2 + rows2 * p.K2;
      // ---- stage 1: dilated conv + conditioning, gated activation ----
      matvec(W1, x, xlen, rows1, p.K1, p.B, part, res);
      for (int i = tid; i < p.pairs * p.B; i += NT) {
        const int pr = i / p.B, b = i - pr * p.B;
        const float a = res[(2 * pr) * MAXB + b] + b1[2 * pr], g = res[(2 * pr + 1) * MAXB + b] + b1[2 * pr + 1];
        put_tagged(p.gbuf + (size_t)b * p.K2 + cta * p.pairs + pr, tanhf(a) * (1.f / (1.f + expf(-g))), tag_g);
      }
      // prefetch the next stage-1 operands (next layer of this step, or layer 0 of the next step) into the other buffer: issued
      // here, while the other CTAs' gate outputs are in flight, instead of on the critical path before the mat-vec
      {
        const int nl = (l + 1 == p.L) ? 0 : l + 1, nt = (l + 1 == p.L) ? t + 1 : t;
        if (nt < p.T) prefetch(nl, nt, buf ^ 1);
      }
      // ---- stage 2: skip and residual 1x1 convolutions ----
      for (int i = tid; i < p.B * p.K2; i += NT) gsm[i] = get_tagged(p.gbuf + i, tag_g);
      __syncthreads();
      matvec(W2, gsm, p.K2, rows2, p.K2, p.B, part, res);
      for (int i = tid; i < rows2 * p.B; i += NT) {
        const int row = i / p.B, b = i - row * p.B;
        const float v = res[row * MAXB + b] + b2[row];
        if (row < p.srows) {
          skips[row * MAXB + b] = (l == 0) ? v : (skips[row * MAXB + b] + v) * r2;
        } else if (l + 1 < p.L) {
          const int r = cta * p.orows + (row - p.srows);
          const float hin = x[b * xlen + (p.K - 1) * p.R + r];
          const int dn = 1 << ((l + 1) % p.layers_per_stack);
          const int rln = (p.K - 1) * dn + 1;
          const float xo = (v + hin) * r2;
          p.ring[p.ring_off[l + 1] + ((size_t)(t % rln) * p.B + b) * p.R + r] = xo;      // for the taps of later samples
          put_tagged(p.xnew + (size_t)b * p.R + r, xo, tag_x);                            // newest tap of layer l+1, now
        }
      }
      if (l + 1 == p.L) {
        for (int i = tid; i < p.srows * p.B; i += NT) {
          const int row = i / p.B, b = i - row * p.B;
          put_tagged(p.sbuf + (size_t)b * p.S + cta * p.srows + row, fmaxf(skips[row * MAXB + b], 0.f), tag0 + per_sample - 2u);
        }
      }
      __syncthreads();                 // x / gsm / res are rewritten by the next stage
      buf ^= 1;
    }
    // ---- output head: ReLU, 1x1 (S -> S), ReLU, 1x1 (S -> O) ----
    for (int i = tid; i < p.B * p.S; i += NT) vec[i] = get_tagged(p.sbuf + i, tag0 + per_sample - 2u);
    __syncthreads();
    matvec(h1w, vec, p.S, p.hrows, p.S, p.B, part, res);
    __threadfence();                   // release: this sample's ring-buffer stores are visible before the tagged words below
    __syncthreads();
    for (int i = tid; i < p.hrows * p.B; i += NT) {
      const int row = i / p.B, b = i - row * p.B;
      put_tagged(p.hbuf + (size_t)b * p.S + cta * p.hrows + row, fmaxf(res[row * MAXB + b] + h1w[p.hrows * p.S + row], 0.f),
                 tag0 + per_sample - 1u);
    }
    __syncthreads();                   // vec is refilled below
    for (int i = tid; i < p.B * p.S; i += NT) vec[i] = get_tagged(p.hbuf + i, tag0 + per_sample - 1u);
    __threadfence();                   // acquire: every CTA's ring-buffer stores of this sample are visible from here on
    __syncthreads();
    matvec(h2w, vec, p.S, p.O, p.S, p.B, part, res);
    // ---- sample from the discretised mixture of logistics (every CTA computes the same value) ----
    if (tid < p.B) {
      const int b = tid, nm = p.O / 3;
      const float* u = p.uniforms + ((size_t)t * p.B + b) * (nm + 1);
      int arg = 0;
      float best = -INFINITY;
      for (int m = 0; m < nm; ++m) {
        const float v = res[m * MAXB + b] + h2w[p.O * p.S + m] - logf(-logf(u[m]));
        if (v > best) { best = v; arg = m; }
      }
      const float mean = res[(nm + arg) * MAXB + b] + h2w[p.O * p.S + nm + arg];
      const float ls = fmaxf(res[(2 * nm + arg) * MAXB + b] + h2w[p.O * p.S + 2 * nm + arg], p.log_scale_min);
      const float ul = u[nm];
      float xs = mean + expf(ls) * (logf(ul) - logf(1.f - ul));
      xs = fminf(fmaxf(xs, -1.f), 1.f);
      cur[b] = xs;
      if (cta == 0) p.out[(size_t)b * p.T + t] = xs;
    }
    if (cta == 0 && p.logits != nullptr) {
      for (int i = tid; i < p.O * p.B; i += NT) {
        const int o = i / p.B, b = i - o * p.B;
        p.logits[((size_t)b * p.T + t) * p.O + o] = res[o * MAXB + b] + h2w[p.O * p.S + o];
      }
    }
    __syncthreads();
  }
  cp_async_wait_all();
}

}  // namespace

// Shared-memory footprint of the kernel for a configuration (bytes), or -1 if it does not fit / is unsupported.
static int64_t wn_smem_bytes(int R, int G, int S, int C, int K, int O, int B, int nC) {
  const int pairs = (G / 2) / nC, srows = S / nC, orows = R / nC, hrows = S / nC;
  const int K1 = K * R + C, K2 = G / 2;
  const int rows1 = 2 * pairs, rows2 = srows + orows;
  const int wpad = rows1 * K1 + pad4(rows1) + rows2 * K2 + pad4(rows2), xlen = (K1 + 3) & ~3;
  int maxrows = rows1 > rows2 ? rows1 : rows2;
  if (hrows > maxrows) maxrows = hrows;
  if (O > maxrows) maxrows = O;
  int64_t f = 2 * (int64_t)wpad + 2 * (int64_t)B * xlen + (int64_t)B * K2 + srows * MAXB + (int64_t)maxrows * NW * MAXB +
              (int64_t)maxrows * MAXB + 2 * R + (hrows * S + pad4(hrows)) + (O * S + pad4(O)) + (int64_t)B * S + MAXB;
  return f * 4 + 64;
}

extern "C" int viai_wavenet_num_ctas(int R, int G, int S, int C, int K, int O, int B) {
  if (R % 4 || (G / 2) % 4 || S % 4 || C % 4 || G % 2 || O % 3 || B < 1 || B > MAXB || K < 1) return 0;
  int dev = 0, sms = kNumSMs;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int n = sms < 128 ? sms : 128; n >= 1; --n) {
    if ((G / 2) % n || S % n || R % n) continue;
    if (wn_smem_bytes(R, G, S, C, K, O, B, n) <= 200 * 1024) return n;
  }
  return 0;
}

// All pointers are device pointers; ring / gbuf / sbuf / hbuf / xchg must be zero-initialised by the caller.
// packed layer weights: see viai_b200/wavenet_vocoder/wavenet.py (pack_for_synthesis).
extern "C" int viai_wavenet_synth(int L, int layers_per_stack, int R, int G, int S, int C, int K, int O, int B, int T, int nC,
                                  const float* packed_layers, const float* first, const float* head1, const float* head2,
                                  const float* cond, const float* uniforms, const float* test_inputs, int Ttest,
                                  float log_scale_min, float* ring, const int64_t* ring_off, float* gbuf, float* sbuf,
                                  float* hbuf, unsigned* xchg, float* out, float* logits, viai_stream_t stream) {
  VIAI_REQUIRE(packed_layers && first && head1 && head2 && cond && uniforms && ring && ring_off && gbuf && sbuf && hbuf && xchg && out,
               "wavenet_synth: null argument");
  VIAI_REQUIRE(((reinterpret_cast<uintptr_t>(gbuf) | reinterpret_cast<uintptr_t>(sbuf) | reinterpret_cast<uintptr_t>(hbuf) |
                 reinterpret_cast<uintptr_t>(xchg)) & 7) == 0, "wavenet_synth: exchange buffers must be 8-byte aligned");
  VIAI_REQUIRE((int64_t)T * (2 * L + 2) < 4000000000LL, "wavenet_synth: T too large for the 32-bit stage tags");
  VIAI_REQUIRE(nC >= 1 && nC == viai_wavenet_num_ctas(R, G, S, C, K, O, B), "wavenet_synth: nC must come from viai_wavenet_num_ctas");
  VIAI_REQUIRE(L >= 1 && layers_per_stack >= 1 && L % layers_per_stack == 0 && T >= 0, "wavenet_synth: bad layer configuration");
  if (T == 0) return VIAI_OK;
  WnParams p;
  memset(&p, 0, sizeof(p));
  p.L = L; p.R = R; p.G = G; p.S = S; p.C = C; p.K = K; p.O = O; p.B = B; p.T = T; p.nC = nC;
  p.layers_per_stack = layers_per_stack;
  p.pairs = (G / 2) / nC; p.srows = S / nC; p.orows = R / nC; p.hrows = S / nC;
  p.K1 = K * R + C; p.K2 = G / 2;
  const int rows1 = 2 * p.pairs, rows2 = p.srows + p.orows;
  p.cta_stride = rows1 * p.K1 + pad4(rows1) + rows2 * p.K2 + pad4(rows2);
  p.layer_stride = p.cta_stride * nC;
  p.wl = packed_layers; p.first = first; p.head1 = head1; p.head2 = head2; p.cond = cond; p.uniforms = uniforms;
  p.test_inputs = test_inputs; p.Ttest = test_inputs ? Ttest : 0; p.log_scale_min = log_scale_min;
  p.ring = ring; p.ring_off = ring_off; p.out = out; p.logits = logits;
  p.gbuf = reinterpret_cast<unsigned long long*>(gbuf); p.sbuf = reinterpret_cast<unsigned long long*>(sbuf);
  p.hbuf = reinterpret_cast<unsigned long long*>(hbuf); p.xnew = reinterpret_cast<unsigned long long*>(xchg);
  const size_t smem = (size_t)wn_smem_bytes(R, G, S, C, K, O, B, nC);
  VIAI_CUDA(cudaFuncSetAttribute(wavenet_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {&p};
  VIAI_CUDA(cudaLaunchCooperativeKernel((void*)wavenet_synth_kernel, dim3(nC), dim3(NT), args, smem, STR(stream)));
  viai::g_launches.fetch_add(1, std::memory_order_relaxed);
  return VIAI_OK;
}
