// WaveNet vocoder synthesis, folded schedule, WARP-SPECIALISED (wavenet_vocoder/wavenet.py:237-364 `incremental_forward`,
// modules.py:162-210, conv.py:17-62, mixture.py:117-153).  Same algebra and parameter blocks as wavenet_synth2.cu
// (WaveNet.pack_for_synthesis_folded):
//     z_l = P'_l + M_l h_{l-1},    P'_l = [old taps + conditioning]_l + N_l h_{l-2} + T_l x_{l-2} + const_l
// but the two halves of a slot no longer share one instruction stream.  ncu's stall sampling of wavenet_synth2 showed 41 % of
// all warp time waiting at CTA barriers for whichever warp was doing the serial work of the moment, and every weight byte
// (119 MB per sample -- the cyclic per-sample sweep defeats the 126 MB L2 completely) arriving from HBM behind a one-slot-deep
// prefetch (profiles/r02_wavenet_folded_ncu_summary.txt).  Here a CTA is 17 warps:
//   * D, 8 warps -- the DEPENDENT chain.  One warp per output unit: a (tanh, sigmoid) gate-row PAIR, a residual row or a skip
//     row.  Per slot: stage h_{l-1} (one tagged word per thread) -> one 256-thread barrier -> the warp's own dot product(s)
//     with weights read from the shared-memory stage -> lane b finishes column b (gate + publish h_l / residual update + ring
//     store + publish x_l / skip accumulation).  No CTA-wide barrier, no cross-warp reduction.
//   * I, 8 warps -- the INDEPENDENT part, one slot ahead: P'_{l+1} from h_{l-1}, x_{l-1} (both published a slot earlier), the
//     old taps and the conditioning vector; handed to the gate warps through a shared-memory mbarrier.
//   * P, 1 warp (one lane) -- the producer: streams the per-slot weight blocks through a 3-stage shared-memory ring with bulk
//     (TMA) copies, 120 KB in flight per SM, and the old taps / conditioning vectors into a 2-deep operand ring.  full / empty
//     mbarriers per stage; nothing else ever waits for memory it has not asked for three slots ago.
// Cross-CTA vectors are 64-bit {value, tag} words as before; h and x rotate over three buffers (the I group reads h_l one
// slot after the D group, so a writer must be three slots ahead before it may reuse a buffer: by then it has seen the
// reader's h_{l+2}, which that CTA's gate warps could only publish after its I group consumed h_l).
// Every wait is bounded (trap after ~2 s) so that a protocol bug fails the launch instead of hanging the device.
#include "common.cuh"
#include "tc_common.cuh"
using namespace viai;
using namespace viai::tc;

namespace {

constexpr int ND = 8, NI = 8;                      // warps of the dependent / independent group
constexpr int NDT = ND * 32, NIT = NI * 32;
constexpr int NWORK = NDT + NIT;                   // threads that take part in the per-sample barriers
constexpr int NT3 = NWORK + 32;                    // + the producer warp
constexpr int NW16 = NWORK / 32;
constexpr int MAXSTAGE = 3;           // weight stages in flight (2 when the operand buffers of a large batch need the room)
constexpr int MAXTAIL = 4;           // old-taps / conditioning buffers in flight (fewer when a large batch needs the room)
constexpr int MAXB = 4;
constexpr int NREP = 8;               // replicas of every cross-CTA vector (see put_lane)
constexpr int kSmemLimit = 226 * 1024;
__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }

struct Wn3Params {
  int L, R, G, S, C, K, O, B, T, nC;
  int layers_per_stack;
  int pairs, srows, orows, hrows;
  int K2, Kn, nstage, ntail, h2w_smem, nrep_used;
  const float* wl;
  int64_t layer_stride, cta_stride;
  const float* wlast;
  const float* first;
  const float* head1;
  const float* head2;
  const float* cond;
  const float* uniforms;
  const float* test_inputs;
  int Ttest;
  float log_scale_min;
  float* ring;
  const int64_t* ring_off;
  unsigned long long* gbuf;            // [3][NREP][B][G/2]
  unsigned long long* xnew;            // [3][NREP][B][R]
  unsigned long long* sbuf;            // [NREP][B][S]
  unsigned long long* hbuf;            // [NREP][B][S]
  float* out;
  float* logits;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Publication (scripts/microbench/exchange_bench.cu, 128 CTAs x 256-word all-to-all on B200): a plain st.relaxed.gpu of the
// tagged word costs ~2900-3000 clocks per exchange (6700 when all CTAs poll one copy), a red.max.u64 ~1800-1900 -- the
// reduction is performed at the L2 slice, and "max" is a plain overwrite here because the tag in the upper 32 bits only ever
// grows.  Every vector exists in NREP copies (a CTA reads copy cta % NREP), and the copies are written by NREP LANES of the
// publishing warp in one instruction (all lanes hold the reduced sums): serial replica stores cost ~35 clocks each.
__device__ __forceinline__ void put_max(unsigned long long* p, float v, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
// Poll until the word carries `tag`.  The loop is tight (load, compare, branch); the watchdog looks at the clock only every 256
// polls so that it does not sit between two polls of the dependent chain.
__device__ __forceinline__ float get_tagged(const unsigned long long* p, unsigned tag) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  if ((unsigned)(w >> 32) != tag) {
    long long t0 = 0;
    unsigned n = 0;
    do {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
      if ((++n & 255u) == 0u) {
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000LL) __trap();
      }
    } while ((unsigned)(w >> 32) != tag);
  }
  return __uint_as_float((unsigned)w);
}
// B tagged words `stride` apart (one per batch column): all loads are issued before the first tag is looked at, so the B columns
// cost one L2 round trip instead of B
template <int B>
__device__ __forceinline__ void get_tagged_cols(const unsigned long long* p, size_t stride, unsigned tag, float* out) {
  unsigned long long w[B];
#pragma unroll
  for (int b = 0; b < B; ++b) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w[b]) : "l"(p + (size_t)b * stride) : "memory");
#pragma unroll
  for (int b = 0; b < B; ++b) out[b] = ((unsigned)(w[b] >> 32) == tag) ? __uint_as_float((unsigned)w[b]) : get_tagged(p + (size_t)b * stride, tag);
}
// two adjacent tagged words with one 16-byte load (each 8-byte half is written atomically and carries its own tag)
__device__ __forceinline__ float2 get_tagged2(const unsigned long long* p, unsigned tag) {
  unsigned long long w0, w1;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
  if ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) {
    const long long t0 = clock64();
    do {
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
      if (clock64() - t0 > 4000000000LL) __trap();
    } while ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag);
  }
  return make_float2(__uint_as_float((unsigned)w0), __uint_as_float((unsigned)w1));
}
// ... for B columns `stride` words apart, all loads issued before the first tag is looked at
template <int B>
__device__ __forceinline__ void get_tagged2_cols(const unsigned long long* p, size_t stride, unsigned tag, float2* out) {
  unsigned long long w0[B], w1[B];
#pragma unroll
  for (int b = 0; b < B; ++b)
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[b]), "=l"(w1[b]) : "l"(p + (size_t)b * stride) : "memory");
#pragma unroll
  for (int b = 0; b < B; ++b) {
    if ((unsigned)(w0[b] >> 32) == tag && (unsigned)(w1[b] >> 32) == tag) out[b] = make_float2(__uint_as_float((unsigned)w0[b]), __uint_as_float((unsigned)w1[b]));
    else out[b] = get_tagged2(p + (size_t)b * stride, tag);
  }
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ int ld_volatile_s32(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_s32(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void bar_work() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory"); }
// group barrier on an mbarrier (count = group size): bounded, unlike a named barrier
__device__ __forceinline__ void group_sync(uint64_t* bar, uint32_t& parity) {
  mbar_arrive(bar);
  mbar_wait(bar, parity);
  parity ^= 1u;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dot product of one weight row (shared address wr, K4 float4s) with the B columns at xs (column stride xstride floats),
// slice [k0, k1) of the float4 index, this lane's share; reduced over the warp (every lane gets the sums).
template <int B>
__device__ __forceinline__ void row_dot(uint32_t wr, uint32_t xs, int xstride, int k0, int k1, float* acc) {
#pragma unroll
  for (int b = 0; b < B; ++b) acc[b] = 0.f;
#pragma unroll 2
  for (int k = k0 + (int)(threadIdx.x & 31); k < k1; k += 32) {
    const float4 w = lds128(wr + 16u * (uint32_t)k);
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const float4 v = lds128(xs + (uint32_t)(b * xstride + 4 * k) * 4u);
      acc[b] = fmaf(w.x, v.x, acc[b]); acc[b] = fmaf(w.y, v.y, acc[b]);
      acc[b] = fmaf(w.z, v.z, acc[b]); acc[b] = fmaf(w.w, v.w, acc[b]);
    }
  }
#pragma unroll
  for (int b = 0; b < B; ++b) acc[b] = warp_sum_f(acc[b]);
}

// this lane's share of one row over the float4 range [k0, k1), accumulated into acc (no reduction)
template <int B>
__device__ __forceinline__ void row_acc(uint32_t wr, uint32_t xs, int xstride, int k0, int k1, float* acc) {
#pragma unroll 2
  for (int k = k0 + (int)(threadIdx.x & 31); k < k1; k += 32) {
    const float4 w = lds128(wr + 16u * (uint32_t)k);
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const float4 v = lds128(xs + (uint32_t)(b * xstride + 4 * k) * 4u);
      acc[b] = fmaf(w.x, v.x, acc[b]); acc[b] = fmaf(w.y, v.y, acc[b]);
      acc[b] = fmaf(w.z, v.z, acc[b]); acc[b] = fmaf(w.w, v.w, acc[b]);
    }
  }
}

// two adjacent rows (the tanh / sigmoid pair) against the same columns: the x loads are shared and the reductions interleave
template <int B>
__device__ __forceinline__ void row_dot2(uint32_t wr, uint32_t row_bytes, uint32_t xs, int xstride, int k1, float* acc0, float* acc1) {
#pragma unroll
  for (int b = 0; b < B; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
#pragma unroll 2
  for (int k = (int)(threadIdx.x & 31); k < k1; k += 32) {
    const float4 w0 = lds128(wr + 16u * (uint32_t)k), w1 = lds128(wr + row_bytes + 16u * (uint32_t)k);
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const float4 v = lds128(xs + (uint32_t)(b * xstride + 4 * k) * 4u);
      acc0[b] = fmaf(w0.x, v.x, acc0[b]); acc1[b] = fmaf(w1.x, v.x, acc1[b]);
      acc0[b] = fmaf(w0.y, v.y, acc0[b]); acc1[b] = fmaf(w1.y, v.y, acc1[b]);
      acc0[b] = fmaf(w0.z, v.z, acc0[b]); acc1[b] = fmaf(w1.z, v.z, acc1[b]);
      acc0[b] = fmaf(w0.w, v.w, acc0[b]); acc1[b] = fmaf(w1.w, v.w, acc1[b]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int b = 0; b < B; ++b) {
      acc0[b] += __shfl_xor_sync(0xffffffffu, acc0[b], o);
      acc1[b] += __shfl_xor_sync(0xffffffffu, acc1[b], o);
    }
  }
}

__device__ __forceinline__ uint32_t tap_bytes(const Wn3Params& p, int layer) {
  return (uint32_t)p.B * ((layer == 0 ? 0u : (uint32_t)(p.K - 1) * p.R * 4u) + (uint32_t)p.C * 4u);
}
// producer only: old taps of `layer` (> 0) at step t and the conditioning vector of step t -> xtail (+ b * xlen)
__device__ __noinline__ void issue_taps(const Wn3Params& p, int layer, int t, float* xtail, int xlen, uint64_t* bar) {
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

struct Smem3 {                                      // offsets in floats, shared by the kernel and the host-side size check
  int wst, hx, tails, hD, Pn, partI, first, h1w, h2w, wlast, vec, res, usm, cur, curh, flags, bars, prof, total;
  int wpad, xlen, hxlen, tlen, rows1, rows2, rowsC, ksN, Hc, nstage, ntail, h2w_smem;
};
__host__ __device__ inline Smem3 smem3_layout(int R, int G, int S, int C, int K, int O, int B, int nC, int nstage, int ntail, int h2w_smem = 1) {
  Smem3 m;
  m.nstage = nstage;
  m.ntail = ntail;
  m.h2w_smem = h2w_smem;
  const int pairs = (G / 2) / nC, srows = S / nC, orows = R / nC, hrows = S / nC;
  const int K2 = G / 2;
  m.xlen = K2 + R + (K - 1) * R + C;
  m.hxlen = K2 + R;                                 // [h_{l-1} | x_{l-1}]: arrive with the slot
  m.tlen = (K - 1) * R + C;                         // [old taps | conditioning]: known long before
  m.rows1 = 2 * pairs; m.rows2 = srows + orows; m.rowsC = m.rows2 + m.rows1;
  m.wpad = m.rowsC * K2 + pad4(m.rowsC) + m.rows1 * m.xlen + 2 * pad4(m.rows1);
  m.ksN = m.rows1 <= NI ? NI / m.rows1 : 1;
  m.Hc = K > 1 ? K - 1 : 1;
  int o = 0;
  m.wst = o; o += nstage * m.wpad;
  m.hx = o; o += 2 * B * m.hxlen;
  m.tails = o; o += ntail * B * m.tlen;
  m.hD = o; o += 3 * B * K2;
  m.Pn = o; o += 2 * pad4(m.rows1) * MAXB;
  m.partI = o; o += pad4(m.rows1 * m.ksN) * MAXB;
  m.first = o; o += 2 * R;
  m.h1w = o; o += hrows * S + pad4(hrows);
  m.h2w = o; o += h2w_smem ? O * S + pad4(O) : 0;   // the last 1x1 (used once per sample) stays in L2 when the room is needed
  m.wlast = o; o += srows * K2 + pad4(srows);
  m.vec = o; o += 2 * B * S;                        // relu(skips) / relu(head 1) staged for the head
  m.res = o; o += pad4(O) * MAXB;
  m.usm = o; o += pad4(MAXB * (O / 3 + 1));          // this step's uniform draws
  m.cur = o; o += MAXB;
  m.curh = o; o += pad4(m.Hc * MAXB);
  m.flags = o; o += 4;
  m.bars = o; o += 2 * 28;                          // up to 28 mbarriers (8 bytes each)
  m.prof = o; o += 2 * 32;                          // profiling counters (PROF instantiation only)
  m.total = o;
  return m;
}

struct Smem3Config { int nstage, ntail, h2w_smem; };
// Preference: weight stages first, then tail buffers; the last 1x1 (used once per sample) moves out of shared memory to L2
// when that buys a stage or a buffer.
__host__ __device__ inline Smem3Config smem3_config(int R, int G, int S, int C, int K, int O, int B, int nC) {
  for (int ns = MAXSTAGE; ns >= 2; --ns)
    for (int nt = MAXTAIL; nt >= 2; --nt)
      for (int hs = 1; hs >= 0; --hs)
        if ((long long)smem3_layout(R, G, S, C, K, O, B, nC, ns, nt, hs).total * 4 + 64 <= kSmemLimit) return Smem3Config{ns, nt, hs};
  return Smem3Config{0, 0, 1};
}

enum { UNIT_NONE = 0, UNIT_PAIR = 1, UNIT_OUT = 2, UNIT_SKIP = 3 };

__device__ long long g_wn3_prof[32];

template <int B, bool PROF>
__global__ void __launch_bounds__(NT3, 1) wavenet_synth3_kernel(const __grid_constant__ Wn3Params p) {
  extern __shared__ __align__(16) float sm[];
  const int cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Smem3 m = smem3_layout(p.R, p.G, p.S, p.C, p.K, p.O, B, p.nC, p.nstage, p.ntail, p.h2w_smem);
  const uint32_t NSTAGE = (uint32_t)p.nstage, NTAIL = (uint32_t)p.ntail;
  const int rows1 = m.rows1, rows2 = m.rows2, rowsC = m.rowsC, xlen = m.xlen, hxlen = m.hxlen, tlen = m.tlen, wpad = m.wpad, Hc = m.Hc, ksN = m.ksN;
  float* const wst = sm + m.wst;
  float* const hx = sm + m.hx;                     // two [B][h_{l-1} | x_{l-1}]
  float* const tails = sm + m.tails;               // NTAIL x [B][old taps | conditioning]
  float* const hD = sm + m.hD;
  float* const Pn = sm + m.Pn;
  float* const partI = sm + m.partI;
  float* const first = sm + m.first;
  float* const h1w = sm + m.h1w;
  float* const h2w = sm + m.h2w;
  float* const wlast = sm + m.wlast;
  float* const vecS = sm + m.vec;
  float* const vecH = vecS + B * p.S;
  float* const res = sm + m.res;
  float* const usm = sm + m.usm;
  float* const cur = sm + m.cur;
  float* const curh = sm + m.curh;
  int* const released = reinterpret_cast<int*>(sm + m.flags);             // samples whose head (ring acquire) is done
  uint64_t* const bars = reinterpret_cast<uint64_t*>(sm + m.bars);
  uint64_t* const full = bars;                     // [NSTAGE] weight stage landed (tx)
  uint64_t* const empty = bars + MAXSTAGE;           // [NSTAGE] all 16 consumer warps are done with the stage
  uint64_t* const xfull = bars + 2 * MAXSTAGE;       // [NTAIL] old taps / conditioning landed (tx)
  uint64_t* const xempty = xfull + MAXTAIL;        // [NTAIL] the I warps are done with the tail buffer
  uint64_t* const pnfull = xempty + MAXTAIL;             // [2] P' of a slot written
  uint64_t* const pnempty = pnfull + 2;            // [2] ... and consumed by the gate lanes
  uint64_t* const dbar = pnempty + 2;              // [3] h_{l-1} staged in hD[n % 3] by the 256 D threads (n-th staging)
  uint64_t* const ibar = dbar + 3;                 // I group barrier
  uint64_t* const hbar = ibar + 1;                 // D group barrier of the head
  long long* const profs = reinterpret_cast<long long*>(sm + m.prof);
  long long tprev = 0;
#define WN3_MARK(cond, slot)                      \
  do {                                            \
    if (PROF && cta == 0 && (cond)) {             \
      const long long now = clock64();            \
      profs[slot] += now - tprev;                 \
      tprev = now;                                \
    }                                             \
  } while (0)
  const unsigned per_sample = 2u * (unsigned)p.L + 2u;
  const float r2 = 0.70710678118654752440f;
  const size_t gstride = (size_t)B * p.K2, xstride = (size_t)B * p.R, sstride = (size_t)B * p.S;   // one replica
  const int NREPB = p.nrep_used;                   // copies actually written / read (<= NREP, the buffers' stride)
  const int rep = cta % NREPB;
  const int K4c = p.K2 >> 2, K4n = xlen >> 2;
  const int offN = rowsC * p.K2 + pad4(rowsC);
  const float* const blk0 = p.wl + (size_t)cta * p.cta_stride;
  const uint32_t total_items = (uint32_t)p.T * (uint32_t)p.L;              // D slots g = 0 .. total-1; I slots v = 0 .. total-1

#pragma unroll 1
  for (int i = tid; i < 2 * p.R; i += NT3) first[i] = p.first[i];
#pragma unroll 1
  for (int i = tid; i < p.hrows * p.S + p.hrows; i += NT3) h1w[i] = p.head1[(size_t)cta * (p.hrows * p.S + pad4(p.hrows)) + i];
  if (p.h2w_smem) {
#pragma unroll 1
    for (int i = tid; i < p.O * p.S + p.O; i += NT3) h2w[i] = p.head2[i];
  }
#pragma unroll 1
  for (int i = tid; i < p.srows * p.K2 + p.srows; i += NT3) wlast[i] = p.wlast[(size_t)cta * (p.srows * p.K2 + pad4(p.srows)) + i];
#pragma unroll 1
  for (int i = tid; i < 2 * B * hxlen + (int)NTAIL * B * tlen; i += NT3) hx[i] = 0.f;   // (hx and tails are adjacent) zero weights must not meet NaNs
  if (tid < MAXB) cur[tid] = (p.test_inputs != nullptr && p.Ttest > 0 && tid < B) ? p.test_inputs[(size_t)tid * p.Ttest] : 0.f;
  if (tid < Hc * MAXB) curh[tid] = 0.f;
  if (PROF && tid < 32) profs[tid] = 0;
  if (tid == 0) {
    *released = 0;
    for (int i = 0; i < MAXSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], ND + NI); }
    for (int i = 0; i < MAXTAIL; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], NI); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pnfull[i], rows1 * B); mbar_init(&pnempty[i], p.pairs * B);   // (replica-0 lanes arrive on pnempty)
    }
    for (int i = 0; i < 3; ++i) mbar_init(&dbar[i], NDT);
    mbar_init(ibar, NIT);
    mbar_init(hbar, NDT);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid < B) curh[0 * MAXB + tid] = cur[tid];                           // cur(0) (slot 0 % Hc)
  __syncthreads();

  // ================================================= producer ====================================================
  if (warp == ND + NI) {
    if (lane != 0) return;
    fence_proxy_async();                           // the generic-proxy initialisation above precedes the bulk writes
    const uint32_t totalW = total_items + 1u, totalX = total_items;
    const uint32_t wbytes = (uint32_t)wpad * 4u;
    uint32_t sw = 0, wstg = 0, wrnd = 0, block = (uint32_t)p.L - 1u;       // weight item, its stage, sw / stages, its block
    uint32_t sx = 0, tbuf = 0, trnd = 0;                                   // tail item, its buffer, sx / buffers
    int nl = 0, nt = 0, lm = 0;                                            // ... its layer, step, layer % layers_per_stack
    int fenced = 0;
    long long t0 = clock64();
    while (sw < totalW || sx < totalX) {
      bool progress = false;
      if (sw < totalW && (wrnd == 0u || mbar_try_wait(&empty[wstg], (wrnd - 1u) & 1u))) {
        mbar_expect_tx(&full[wstg], wbytes);                               // item s carries block (s - 1) mod L
        bulk_load(wst + wstg * wpad, blk0 + (size_t)block * p.layer_stride, wbytes, &full[wstg]);
        ++sw;
        if (++wstg == NSTAGE) { wstg = 0; ++wrnd; }
        if (++block == (uint32_t)p.L) block = 0;
        progress = true;
      }
      if (sx < totalX) {
        // the taps of layer nl at step nt were written during steps <= nt - d (d = the layer's dilation): readable once the
        // head of step nt - d has been acquired, i.e. released >= nt - d + 1; they are asked for up to NTAIL slots ahead
        const int rel = ld_volatile_s32(released);
        const bool ready = nl == 0 || rel >= nt - (1 << lm) + 1;
        if (ready && (trnd == 0u || mbar_try_wait(&xempty[tbuf], (trnd - 1u) & 1u))) {
          if (nl != 0 && fenced != rel) { fence_proxy_async(); fenced = rel; }
          mbar_expect_tx(&xfull[tbuf], tap_bytes(p, nl));
          issue_taps(p, nl, nt, tails + tbuf * B * tlen, tlen, &xfull[tbuf]);
          ++sx;
          if (++tbuf == NTAIL) { tbuf = 0; ++trnd; }
          if (++lm == p.layers_per_stack) lm = 0;
          if (++nl == p.L) { nl = 0; lm = 0; ++nt; }
          progress = true;
        }
      }
      if (progress) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) __trap();
    }
    return;
  }

  // ========================================== worker warps: D and I ==============================================
  // One loop over the samples for both groups (the head is common code); inside, a warp runs either the D slots or the I
  // slots of the sample.  All per-slot indices are counters: run-time divisions (v % L, s % stages, ...) cost 100+ clocks each
  // and sat on the dependent chain.
  const bool isD = warp < ND;
  uint32_t iparity = 0, dparity = 0;
  // ---- I group: slot v computes P' of layer nl = v % L at step nt = v / L from stream item v (block (v - 1) mod L) ----
  const int wI = isD ? 0 : warp - ND, tI = tid - NDT;
  const int rowN = wI % rows1, sliceN = wI / rows1;
  // The row is split at float4 index KA0 = (K2 + R) / 4: columns [KA0, K4n) (old taps, conditioning) are known long before
  // the slot and are accumulated first; columns [0, KA0) (h_{lp-1}, x_{lp-1}) arrive together with the D group's own input.
  const int KA0 = (p.K2 + p.R) >> 2;
  const int chunkA = (K4n - KA0 + ksN - 1) / ksN, chunkB = (KA0 + ksN - 1) / ksN;
  const bool has_slice = sliceN < ksN;
  const int kA0 = KA0 + sliceN * chunkA, kA1 = has_slice ? min(K4n, kA0 + chunkA) : kA0;
  const int kB0 = sliceN * chunkB, kB1 = has_slice ? min(KA0, kB0 + chunkB) : kB0;
  uint32_t iv = 0, ist = 0, iph = 0;               // slot, its weight stage and that stage's phase
  uint32_t itb = 0, itph = 0;                      // its tail buffer and that buffer's phase
  int inl = 0, int_t = 0, ilp = -1, itq = 0;       // (layer, step) of the slot; (layer, step) of the D slot it runs beside
  unsigned itag = 1u;                              // tag0 of step itq
  uint32_t in3 = 0, inph = 0;                      // staging v - 2 of the D group: ring index and phase (valid from v = 2)
  auto islot = [&]() {
    const uint32_t v = iv, st = ist, tb = itb;
    const int nl = inl, nt = int_t, lp = ilp;
    float* x = hx + (v & 1u) * B * hxlen;          // [h | x] of this slot
    float* tl = tails + tb * B * tlen;             // [old taps | conditioning] of this slot
    WN3_MARK(tI == 0, 16);
    // ---- part A: old taps and conditioning ----
    if (nl == 0) {                                 // layer 0's old taps are rebuilt from the last input samples
#pragma unroll 1
      for (int j = 0; j < p.K - 1; ++j) {
        const int tau = nt - (p.K - 1 - j);
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float c = tau >= 0 ? curh[(tau % Hc) * MAXB + b] : 0.f;
          float* xt = tl + b * tlen + j * p.R;
#pragma unroll 1
          for (int r = tI; r < p.R; r += NIT) xt[r] = tau >= 0 ? fmaf(first[r], c, first[p.R + r]) : 0.f;
        }
      }
      fence_proxy_async();                         // these generic writes precede the bulk copies that reuse the tail later
      group_sync(ibar, iparity);
    }
    mbar_wait(&full[st], iph);
    WN3_MARK(tI == 0, 19);
    mbar_wait(&xfull[tb], itph);
    WN3_MARK(tI == 0, 20);
    const float* Wn = wst + st * wpad + offN;
    const uint32_t wrow = smem_u32(Wn) + (uint32_t)(rowN * K4n) * 16u;
    float cn = 0.f;
    if (tI < rows1 * B) cn = Wn[rows1 * xlen + tI / B];
    float acc[B];
#pragma unroll
    for (int b = 0; b < B; ++b) acc[b] = 0.f;
    row_acc<B>(wrow, smem_u32(tl) - (uint32_t)KA0 * 16u, tlen, kA0, kA1, acc);
    __syncwarp();
    if (lane == 0) mbar_arrive(&xempty[tb]);       // the tail buffer can be refilled (NTAIL slots ahead)
    WN3_MARK(tI == 0, 21);
    // ---- part B: h_{lp-1} (staged by this CTA's D group for its own slot lp) and x_{lp-1} ----
    if (v >= 1u) {
      if (lp <= 1) {                               // layers 1 and 2 see x_0 = fw * sample + fb
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float c = cur[b];
#pragma unroll 1
          for (int r = tI; r < p.R; r += NIT) x[b * hxlen + p.K2 + r] = fmaf(first[r], c, first[p.R + r]);
        }
      } else if (lp + 1 < p.L) {                   // T_{nl} x_{lp-1}: published a slot ago, two words per 16-byte load
        const unsigned long long* xs = p.xnew + ((size_t)((lp - 1) % 3) * NREP + rep) * xstride;
        const int half = p.R >> 1;
#pragma unroll 1
        for (int r2i = tI; r2i < half; r2i += NIT) {
          float2 v2[B];
          get_tagged2_cols<B>(xs + 2 * r2i, (size_t)p.R, itag + 2u * (unsigned)(lp - 1) + 1u, v2);
#pragma unroll
          for (int b = 0; b < B; ++b) *reinterpret_cast<float2*>(x + b * hxlen + p.K2 + 2 * r2i) = v2[b];
        }
      }
      if (lp >= 1 && lp + 1 < p.L) {               // N_{nl} h_{lp-1}
        // hD is a ring of THREE: the D group's residual / skip warps do not wait for P', so they start staging h_{lp+1}
        // (two stagings later) as soon as faster CTAs publish it, possibly before this copy; three stagings later they
        // cannot (that needs this CTA's own gate of slot lp+1, i.e. the P' this slot produces)
        mbar_wait(&dbar[in3], inph);
        const float* h = hD + in3 * B * p.K2;
#pragma unroll
        for (int b = 0; b < B; ++b) {
#pragma unroll 1
          for (int k = tI; k < p.K2; k += NIT) x[b * hxlen + k] = h[b * p.K2 + k];
        }
      }
    }
    WN3_MARK(tI == 0, 17);
    group_sync(ibar, iparity);
    WN3_MARK(tI == 0, 18);
    row_acc<B>(wrow, smem_u32(x), hxlen, kB0, kB1, acc);
    if (has_slice) {
#pragma unroll
      for (int b = 0; b < B; ++b) acc[b] = warp_sum_f(acc[b]);
      float sv = acc[0];
#pragma unroll
      for (int b = 1; b < B; ++b) sv = lane == b ? acc[b] : sv;
      if (lane < B) partI[(rowN * ksN + sliceN) * MAXB + lane] = sv;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    group_sync(ibar, iparity);
    WN3_MARK(tI == 0, 22);
    if (tI < rows1 * B) {
      const int row = tI / B, b = tI - row * B;
      if (v >= 2u) mbar_wait(&pnempty[v & 1u], ((v >> 1) - 1u) & 1u);
      float sum = cn;
#pragma unroll 4
      for (int ks = 0; ks < ksN; ++ks) sum += partI[(row * ksN + ks) * MAXB + b];
      Pn[(v & 1u) * pad4(rows1) * MAXB + row * MAXB + b] = sum;
      mbar_arrive(&pnfull[v & 1u]);
    }
    WN3_MARK(tI == 0, 23);
    // ---- counters of the next slot ----
    if (v >= 2u) { if (++in3 == 3u) { in3 = 0; inph ^= 1u; } }
    if (v >= 1u && nl == 0) itag += per_sample;    // (lp, tq) <- (nl, nt): tq advances when the slot's layer was 0
    ilp = nl; itq = nt;
    if (++inl == p.L) { inl = 0; ++int_t; }
    if (++ist == NSTAGE) { ist = 0; iph ^= 1u; }
    if (++itb == NTAIL) { itb = 0; itph ^= 1u; }
    ++iv;
  };

  // ---- D group: one warp per output unit ----
  int utype = UNIT_NONE, uidx = 0;
  if (isD) {
    if (warp < p.pairs) { utype = UNIT_PAIR; uidx = warp; }
    else if (warp < p.pairs + p.orows) { utype = UNIT_OUT; uidx = warp - p.pairs; }
    else if (warp < p.pairs + p.orows + p.srows) { utype = UNIT_SKIP; uidx = warp - p.pairs - p.orows; }
  }
  // block row order: [skip rows | residual rows | gate rows (a, g adjacent)]
  const int row0 = utype == UNIT_PAIR ? rows2 + 2 * uidx : utype == UNIT_OUT ? p.srows + uidx : uidx;
  float uc_a = 0.f, uc_g = 0.f;                    // layer 0's rank-one coefficients A_0 fw (block 0)
  if (utype == UNIT_PAIR) {
    const float* uc = blk0 + offN + rows1 * xlen + pad4(rows1);
    uc_a = __ldg(uc + 2 * uidx);
    uc_g = __ldg(uc + 2 * uidx + 1);
  }
  float state = 0.f;                               // lane (copy, b): x_{l-1}[own row] (residual unit) / running skip sum (skip unit)
  uint32_t dst = 1u % NSTAGE, dph = 1u / NSTAGE;   // stream item g + 1 of the current slot: stage and phase
  uint32_t dn3 = 0, dnph = 0;                      // staging counter n: ring index n % 3 and phase (n / 3) & 1
  if (isD && lane == 0) mbar_arrive(&empty[0]);    // stream item 0 (the prologue's block) has no dependent part

#pragma unroll 1
  for (int t = 0; t < p.T; ++t) {
    const unsigned tag0 = 1u + (unsigned)t * per_sample;
    if (!isD) {
      if (tI < B * (p.O / 3 + 1)) usm[tI] = __ldg(p.uniforms + (size_t)t * B * (p.O / 3 + 1) + tI);   // read by the sampler after (B)
#pragma unroll 1
      for (int l = (t == 0 ? -1 : 0); l < p.L; ++l) {   // (step 0 starts with the extra slot 0: P' of layer 0 at step 0)
        if (iv < total_items) islot();
      }
      WN3_MARK(tI == 0, 16);
    } else {
      int lm = 0;                                  // l % layers_per_stack
#pragma unroll 1
      for (int l = 0; l < p.L; ++l) {
        const uint32_t g = (uint32_t)t * (uint32_t)p.L + (uint32_t)l, st = dst;
        const unsigned tag_h = tag0 + 2u * (unsigned)l, tag_x = tag_h + 1u;
        float acc0[B], acc1[B];
        float bias = 0.f;
#pragma unroll
        for (int b = 0; b < B; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
        WN3_MARK(tid == 0, 0);
        // the weight stage (asked for three slots ago) does not depend on h_{l-1}: wait for it first
        if (l > 0) mbar_wait(&full[st], dph);
        WN3_MARK(tid == 0, 3);
        if (l > 0) {
          float* h = hD + dn3 * B * p.K2;          // n-th staging of this CTA, n = t * L + (l - 1)
          const unsigned long long* gb = p.gbuf + ((size_t)((l - 1) % 3) * NREP + rep) * gstride;
#pragma unroll 1
          for (int k = tid; k < p.K2; k += NDT) {
            float hv[B];
            get_tagged_cols<B>(gb + k, (size_t)p.K2, tag_h - 2u, hv);
#pragma unroll
            for (int b = 0; b < B; ++b) h[b * p.K2 + k] = hv[b];
          }
          WN3_MARK(tid == 0, 1);
          mbar_arrive(&dbar[dn3]);
          mbar_wait(&dbar[dn3], dnph);
          if (++dn3 == 3u) { dn3 = 0; dnph ^= 1u; }
          WN3_MARK(tid == 0, 2);
          const float* Wc = wst + st * wpad;
          if (utype == UNIT_PAIR) {
            row_dot2<B>(smem_u32(Wc) + (uint32_t)(row0 * K4c) * 16u, (uint32_t)K4c * 16u, smem_u32(h), p.K2, K4c, acc0, acc1);
          } else if (utype != UNIT_NONE) {
            row_dot<B>(smem_u32(Wc) + (uint32_t)(row0 * K4c) * 16u, smem_u32(h), p.K2, 0, K4c, acc0);
            bias = Wc[rowsC * p.K2 + row0];
          }
        }
        WN3_MARK(tid == 0, 12);
        if (++dst == NSTAGE) { dst = 0; dph ^= 1u; }
        WN3_MARK(tid == 0, 4);
        if (lane < NREPB * B) {                    // lane = (copy r_, column b): every copy lane repeats the column's arithmetic
          const int b = lane % B, r_ = lane / B;
          float a0 = acc0[0], a1 = acc1[0];
#pragma unroll
          for (int bb = 1; bb < B; ++bb) { a0 = b == bb ? acc0[bb] : a0; a1 = b == bb ? acc1[bb] : a1; }
          if (utype == UNIT_PAIR) {
            WN3_MARK(tid == 0, 13);
            if (PROF && cta == 0 && tid == 0 && !mbar_try_wait(&pnfull[g & 1u], (g >> 1) & 1u)) profs[15] += 1;   // P' not ready yet
            mbar_wait(&pnfull[g & 1u], (g >> 1) & 1u);
            WN3_MARK(tid == 0, 5);
            const float* P = Pn + (g & 1u) * pad4(rows1) * MAXB;
            float a = P[(2 * uidx) * MAXB + b], gg = P[(2 * uidx + 1) * MAXB + b];
            if (l > 0) { a += a0; gg += a1; }
            else { a = fmaf(uc_a, cur[b], a); gg = fmaf(uc_g, cur[b], gg); }
            WN3_MARK(tid == 0, 14);
            put_max(p.gbuf + ((size_t)(l % 3) * NREP + r_) * gstride + (size_t)b * p.K2 + cta * p.pairs + uidx,
                    tanhf(a) * (1.f / (1.f + expf(-gg))), tag_h);
            if (r_ == 0) mbar_arrive(&pnempty[g & 1u]);       // (arrivals have release semantics: after the publication)
            WN3_MARK(tid == 0, 6);
          } else if (utype == UNIT_OUT && l > 0) {
            const int r = cta * p.orows + uidx;
            const float xprev = (l == 1) ? fmaf(first[r], cur[b], first[p.R + r]) : state;
            const float xo = (a0 + bias + xprev) * r2;
            state = xo;
            if (r_ == 0) {
              const int rl = (p.K - 1) * (1 << lm) + 1;
              p.ring[p.ring_off[l] + ((size_t)(t % rl) * B + b) * p.R + r] = xo;          // taps of later samples
            }
            if (l + 2 < p.L) put_max(p.xnew + ((size_t)(l % 3) * NREP + r_) * xstride + (size_t)b * p.R + r, xo, tag_x);
          } else if (utype == UNIT_SKIP && l > 0) {
            const float v = a0 + bias;
            state = (l == 1) ? v : (state + v) * r2;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);    // the stage is free for the producer (after the publication, off the chain)
        if (++lm == p.layers_per_stack) lm = 0;
      }
      WN3_MARK(tid == 0, 0);
      // ---- skip rows of the last layer, relu(skips) ----
      {
        float* h = hD + dn3 * B * p.K2;
        const unsigned long long* gb = p.gbuf + ((size_t)((p.L - 1) % 3) * NREP + rep) * gstride;
#pragma unroll 1
        for (int k = tid; k < p.K2; k += NDT) {
          float hv[B];
          get_tagged_cols<B>(gb + k, (size_t)p.K2, tag0 + 2u * (unsigned)(p.L - 1), hv);
#pragma unroll
          for (int b = 0; b < B; ++b) h[b * p.K2 + k] = hv[b];
        }
        mbar_arrive(&dbar[dn3]);
        mbar_wait(&dbar[dn3], dnph);
        if (++dn3 == 3u) { dn3 = 0; dnph ^= 1u; }
        if (utype == UNIT_SKIP) {
          float acc[B];
          row_dot<B>(smem_u32(wlast) + (uint32_t)(uidx * K4c) * 16u, smem_u32(h), p.K2, 0, K4c, acc);
          if (lane < NREPB * B) {
            const int b = lane % B, r_ = lane / B;
            float a0 = acc[0];
#pragma unroll
            for (int bb = 1; bb < B; ++bb) a0 = b == bb ? acc[bb] : a0;
            const float v = a0 + wlast[p.srows * p.K2 + uidx];
            const float sk = (p.L == 1) ? v : (state + v) * r2;
            put_max(p.sbuf + (size_t)r_ * sstride + (size_t)b * p.S + cta * p.srows + uidx, fmaxf(sk, 0.f), tag0 + per_sample - 2u);
          }
        }
      }
      __threadfence();                             // release: this sample's ring stores precede the head-1 words below
      WN3_MARK(tid == 0, 7);
    }
    // ================================================ output head ==================================================
    bar_work();                                    // (A) relu(skips) of this CTA published, ring stores fenced
    WN3_MARK(tid == 0, 8);
    WN3_MARK(!isD && tI == 0, 24);
    if (isD) {                                     // head 1: relu(skips) of every CTA staged once, one warp per row of this CTA
      const unsigned long long* sb = p.sbuf + (size_t)rep * sstride;
#pragma unroll 1
      for (int k = tid; k < p.S; k += NDT) {
        float sv[B];
        get_tagged_cols<B>(sb + k, (size_t)p.S, tag0 + per_sample - 2u, sv);
#pragma unroll
        for (int b = 0; b < B; ++b) vecS[b * p.S + k] = sv[b];
      }
      group_sync(hbar, dparity);
      if (warp < p.hrows) {
        float acc[B];
        row_dot<B>(smem_u32(h1w) + (uint32_t)(warp * (p.S >> 2)) * 16u, smem_u32(vecS), p.S, 0, p.S >> 2, acc);
        if (lane < NREPB * B) {
          const int b = lane % B, r_ = lane / B;
          float sv = acc[0];
#pragma unroll
          for (int bb = 1; bb < B; ++bb) sv = b == bb ? acc[bb] : sv;
          put_max(p.hbuf + (size_t)r_ * sstride + (size_t)b * p.S + cta * p.hrows + warp, fmaxf(sv + h1w[p.hrows * p.S + warp], 0.f),
                  tag0 + per_sample - 1u);
        }
      }
    } else {                                       // the I group stages relu(head 1) of every CTA
      const unsigned long long* hb = p.hbuf + (size_t)rep * sstride;
#pragma unroll 1
      for (int k = tI; k < p.S; k += NIT) {
        float hv[B];
        get_tagged_cols<B>(hb + k, (size_t)p.S, tag0 + per_sample - 1u, hv);
#pragma unroll
        for (int b = 0; b < B; ++b) vecH[b * p.S + k] = hv[b];
      }
      __threadfence();                             // acquire: every CTA's ring stores of this sample precede the next bulk reads
    }
    bar_work();                                    // (A2)
#pragma unroll 1
    for (int o = warp; o < p.O; o += NW16) {       // head 2: every CTA, all 16 worker warps
      float acc[B];
      if (p.h2w_smem) {
        row_dot<B>(smem_u32(h2w) + (uint32_t)(o * (p.S >> 2)) * 16u, smem_u32(vecH), p.S, 0, p.S >> 2, acc);
      } else {                                     // weights straight from L2 (once per sample)
#pragma unroll
        for (int b = 0; b < B; ++b) acc[b] = 0.f;
        const float4* wg = reinterpret_cast<const float4*>(p.head2 + (size_t)o * p.S);
#pragma unroll 2
        for (int k = lane; k < (p.S >> 2); k += 32) {
          const float4 w = __ldg(wg + k);
#pragma unroll
          for (int b = 0; b < B; ++b) {
            const float4 v = *reinterpret_cast<const float4*>(vecH + b * p.S + 4 * k);
            acc[b] = fmaf(w.x, v.x, acc[b]); acc[b] = fmaf(w.y, v.y, acc[b]);
            acc[b] = fmaf(w.z, v.z, acc[b]); acc[b] = fmaf(w.w, v.w, acc[b]);
          }
        }
#pragma unroll
        for (int b = 0; b < B; ++b) acc[b] = warp_sum_f(acc[b]);
      }
      if (lane < B) {
        float sv = acc[0];
#pragma unroll
        for (int b = 1; b < B; ++b) sv = lane == b ? acc[b] : sv;
        res[o * MAXB + lane] = sv + (p.h2w_smem ? h2w[p.O * p.S + o] : __ldg(p.head2 + (size_t)p.O * p.S + o));
      }
    }
    WN3_MARK(tid == 0, 9);
    bar_work();                                    // (B)
    WN3_MARK(tid == 0, 10);
    // ---- sample from the discretised mixture of logistics (every CTA computes the same value): warp 0, lane m = mixture m
    if (warp == 0) {
      const int nm = p.O / 3;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const float* u = usm + b * (nm + 1);
        float best = -INFINITY;
        int arg = 0x7fffffff;
#pragma unroll 1
        for (int m0 = 0; m0 < nm; m0 += 32) {      // (nm = 10 in every configuration of the reference: one trip)
          const int mm = m0 + lane;
          float v = -INFINITY;
          if (mm < nm) v = res[mm * MAXB + b] - logf(-logf(u[mm]));
          if (v > best) { best = v; arg = mm; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {         // arg max, the FIRST maximum like the reference's argmax over the mixture axis
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
          if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
        }
        if (lane == 0) {
          const float mean = res[(nm + arg) * MAXB + b];
          const float ls = fmaxf(res[(2 * nm + arg) * MAXB + b], p.log_scale_min);
          const float ul = u[nm];
          float xs = mean + expf(ls) * (logf(ul) - logf(1.f - ul));
          xs = fminf(fmaxf(xs, -1.f), 1.f);
          if (cta == 0) p.out[(size_t)b * p.T + t] = xs;
          const float nxt = (p.test_inputs != nullptr && t + 1 < p.Ttest) ? p.test_inputs[(size_t)b * p.Ttest + t + 1] : xs;
          cur[b] = nxt;
          curh[((t + 1) % Hc) * MAXB + b] = nxt;
        }
      }
    }
    if (tid == 32) st_volatile_s32(released, t + 1);
    if (cta == 0 && p.logits != nullptr && isD) {
#pragma unroll 1
      for (int i = tid; i < p.O * B; i += NDT) {
        const int o = i / B, b = i - o * B;
        p.logits[((size_t)b * p.T + t) * p.O + o] = res[o * MAXB + b];
      }
    }
    bar_work();                                    // (C) next input sample in place
    WN3_MARK(tid == 0, 11);
    WN3_MARK(!isD && tI == 0, 25);
  }
  if (PROF && cta == 0 && tid < 16) g_wn3_prof[tid] = profs[tid];
  if (PROF && cta == 0 && !isD && tI < 16) g_wn3_prof[16 + tI] = profs[16 + tI];
#undef WN3_MARK
}

template <int B, bool PROF>
cudaError_t launch_wn3(const Wn3Params& p, size_t smem, cudaStream_t stream) {
  auto k = wavenet_synth3_kernel<B, PROF>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<Wn3Params*>(&p)};
  return cudaLaunchCooperativeKernel((void*)k, dim3(p.nC), dim3(NT3), args, smem, stream);
}

}  // namespace

// Number of cooperating CTAs of the warp-specialised folded kernel, 0 if the configuration is unsupported (one D warp per
// output unit, one I warp per (gate row, K slice); use viai_wavenet_synth2 / viai_wavenet_synth then).
extern "C" int viai_wavenet3_num_ctas(int L, int R, int G, int S, int C, int K, int O, int B) {
  if (R % 4 || (G / 2) % 4 || S % 4 || C % 4 || G % 2 || O % 3 || B < 1 || B > MAXB || K < 1 || L < 3) return 0;
  int dev = 0, sms = kNumSMs;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int n = sms < 128 ? sms : 128; n >= 1; --n) {
    if ((G / 2) % n || S % n || R % n) continue;
    const int pairs = (G / 2) / n, srows = S / n, orows = R / n, hrows = S / n;
    if (pairs + orows + srows > ND || 2 * pairs > NI || hrows > ND) continue;
    const Smem3Config c = smem3_config(R, G, S, C, K, O, B, n);
    if (c.nstage >= 2 && c.ntail >= 2) return n;
  }
  return 0;
}

// Same contract as viai_wavenet_synth2 (blocks of WaveNet.pack_for_synthesis_folded) except for the exchange buffers, which hold
// viai_wavenet3_replicas() copies of every vector: gbuf 3 * NREP * B * (G/2), xchg 3 * NREP * B * R, sbuf / hbuf NREP * B * S
// 64-bit words (zeroed, 8-byte aligned).
extern "C" int viai_wavenet3_replicas(void) { return NREP; }

extern "C" int viai_wavenet_synth3(int L, int layers_per_stack, int R, int G, int S, int C, int K, int O, int B, int T, int nC,
                                   const float* packed_layers, const float* last, const float* first, const float* head1,
                                   const float* head2, const float* cond, const float* uniforms, const float* test_inputs, int Ttest,
                                   float log_scale_min, float* ring, const int64_t* ring_off, float* gbuf, float* sbuf,
                                   float* hbuf, unsigned* xchg, float* out, float* logits, viai_stream_t stream) {
  VIAI_REQUIRE(packed_layers && last && first && head1 && head2 && cond && uniforms && ring && ring_off && gbuf && sbuf && hbuf && xchg && out,
               "wavenet_synth3: null argument");
  VIAI_REQUIRE(((reinterpret_cast<uintptr_t>(gbuf) | reinterpret_cast<uintptr_t>(sbuf) | reinterpret_cast<uintptr_t>(hbuf) |
                 reinterpret_cast<uintptr_t>(xchg)) & 7) == 0, "wavenet_synth3: exchange buffers must be 8-byte aligned");
  VIAI_REQUIRE((int64_t)T * (2 * L + 2) < 4000000000LL && (int64_t)T * L < 4000000000LL, "wavenet_synth3: T too large for the 32-bit tags");
  VIAI_REQUIRE(nC >= 1 && nC == viai_wavenet3_num_ctas(L, R, G, S, C, K, O, B), "wavenet_synth3: nC must come from viai_wavenet3_num_ctas");
  VIAI_REQUIRE(layers_per_stack >= 1 && L % layers_per_stack == 0 && T >= 0, "wavenet_synth3: bad layer configuration");
  if (T == 0) return VIAI_OK;
  Wn3Params p;
  memset(&p, 0, sizeof(p));
  p.L = L; p.R = R; p.G = G; p.S = S; p.C = C; p.K = K; p.O = O; p.B = B; p.T = T; p.nC = nC;
  p.layers_per_stack = layers_per_stack;
  p.pairs = (G / 2) / nC; p.srows = S / nC; p.orows = R / nC; p.hrows = S / nC;
  p.K2 = G / 2; p.Kn = p.K2 + R + (K - 1) * R + C;
  const Smem3Config cfg = smem3_config(R, G, S, C, K, O, B, nC);
  p.nstage = cfg.nstage; p.ntail = cfg.ntail; p.h2w_smem = cfg.h2w_smem;
  const Smem3 m = smem3_layout(R, G, S, C, K, O, B, nC, p.nstage, p.ntail, p.h2w_smem);
  p.cta_stride = m.wpad;
  p.layer_stride = p.cta_stride * nC;
  p.wl = packed_layers; p.wlast = last; p.first = first; p.head1 = head1; p.head2 = head2; p.cond = cond; p.uniforms = uniforms;
  p.test_inputs = test_inputs; p.Ttest = test_inputs ? Ttest : 0; p.log_scale_min = log_scale_min;
  p.ring = ring; p.ring_off = ring_off; p.out = out; p.logits = logits;
  p.gbuf = reinterpret_cast<unsigned long long*>(gbuf); p.sbuf = reinterpret_cast<unsigned long long*>(sbuf);
  p.hbuf = reinterpret_cast<unsigned long long*>(hbuf); p.xnew = reinterpret_cast<unsigned long long*>(xchg);
  const size_t smem = (size_t)m.total * 4 + 64;
  cudaError_t e = cudaErrorInvalidValue;
  {
    // copies of the exchanged vectors in use (tuning knob VIAI_WN3_NREP): with red.max publication they matter little -- measured
    // on B200, 8 / 4 / 2 / 1 copies: B = 1 22.2 / 22.3 / 21.4 / 22.0 k samples/s, B = 4 37.2 / 37.9 / 38.8 / 40.2 k
    const char* pn = getenv("VIAI_WN3_NREP");
    p.nrep_used = pn ? atoi(pn) : (B <= 2 ? NREP : 1);
    if (p.nrep_used < 1 || p.nrep_used > NREP || p.nrep_used * B > 32) p.nrep_used = B <= 2 ? NREP : 1;
  }
  const char* pe = getenv("VIAI_WN3_PROF");
  const bool prof = pe && pe[0] == '1';
  switch (B * 2 + (prof ? 1 : 0)) {
    case 2: e = launch_wn3<1, false>(p, smem, STR(stream)); break;
    case 3: e = launch_wn3<1, true>(p, smem, STR(stream)); break;
    case 4: e = launch_wn3<2, false>(p, smem, STR(stream)); break;
    case 5: e = launch_wn3<2, true>(p, smem, STR(stream)); break;
    case 6: e = launch_wn3<3, false>(p, smem, STR(stream)); break;
    case 7: e = launch_wn3<3, true>(p, smem, STR(stream)); break;
    case 8: e = launch_wn3<4, false>(p, smem, STR(stream)); break;
    case 9: e = launch_wn3<4, true>(p, smem, STR(stream)); break;
  }
  VIAI_CUDA(e);
  viai::g_launches.fetch_add(1, std::memory_order_relaxed);
  return VIAI_OK;
}

// Debug aid (VIAI_WN3_PROF=1 selects the profiling instantiation): clocks CTA 0 spent per phase in the last launch.
// [0..11] gate warp, lane 0: rest, stage h, group barrier, weight stage wait, dot products, P' wait, gate + publish, last skip
// rows + fence, barrier A, head, barrier B, sampler + barrier C.  [16..25] independent group, thread 0: rest, gathers, group
// barrier, weight stage wait, operand wait, dot product, group barrier, P' hand-over, barrier A, head.
extern "C" int viai_wavenet3_profile(long long* out32) {
  VIAI_REQUIRE(out32, "wavenet3_profile: null argument");
  VIAI_CUDA(cudaDeviceSynchronize());
  VIAI_CUDA(cudaMemcpyFromSymbol(out32, g_wn3_prof, 32 * sizeof(long long)));
  return VIAI_OK;
}
