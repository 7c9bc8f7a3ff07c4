// WaveNet vocoder synthesis on ONE 16-CTA thread-block cluster (same arithmetic and the same packed parameter blocks as
// wavenet_synth.cu with nC = 16; that kernel stays as the fallback and the on-device cross-check).
//
// Why a cluster: synthesis is a chain of 2*L+2 = 50 dependent mat-vec stages per sample, each needing the WHOLE output vector
// of the previous one.  Exchanging those vectors through global memory costs ~1.3 us per stage (store -> L2 -> polled load),
// 131 us per sample.  Inside a cluster the vectors travel through distributed shared memory: every CTA stores its slice as
// 64-bit {value, stage tag} words straight into the shared memory of all 16 CTAs (st.shared::cluster) and polls only its OWN
// shared memory, so an exchange is a few hundred nanoseconds.  The price is that only 16 SMs stream the 99 MB of fp32 weights
// every sample: measured L2 -> SM bandwidth is ~150 GB/s per SM with bulk copies (scripts/sm_l2_bandwidth_probe.cu), i.e.
// 42 us per sample -- the bound of this design.
//
//   * warp 16 = weight producer: an endless, data-independent stream of row-aligned chunks (4 gate rows of 1616 floats, or half of
//     the skip/residual block + biases; <= 25.9 KB) cp.async.bulk'ed into a ring of shared-memory slots (full / empty mbarriers);
//   * warps 0..15 = consumers: per chunk every warp owns (row, K slice), partial sums meet in shared memory; gate / residual
//     epilogues write their 16 / 32 results into all 16 CTAs;
//   * older dilated taps come from the per-layer ring buffers in global memory (L2), prefetched one stage ahead with cp.async;
//     their visibility across CTAs is ordered by one fence pair per sample around the last exchange;
//   * every spin has a watchdog: a protocol bug traps (the launch fails) instead of hanging the device.
#include "common.cuh"
#include "tc_common.cuh"
#include <cooperative_groups.h>
using namespace viai;
using namespace viai::tc;

namespace {

constexpr int CL = 16;              // CTAs of the cluster
constexpr int NTC = 512;            // consumer threads
constexpr int NWC = NTC / 32;
constexpr int NTHR = NTC + 32;      // + the producer warp
constexpr int MAXBC = 2;            // utterances per launch on this path
constexpr int MAXCHUNK = 16;
__host__ __device__ constexpr int cpad4(int n) { return (n + 3) & ~3; }

struct WcChunk {
  uint32_t off_bytes, bytes;        // inside the (layer, CTA) block
};
struct WcParams {
  int L, R, G, S, C, K, O, B, T;
  int layers_per_stack;
  int pairs, srows, orows, hrows, K1, K2, xlen;
  int nchunk, nchunk1;              // chunks per layer; the first nchunk1 hold gate rows (rows1_per_chunk each)
  int rows1_per_chunk, rows2_first; // stage-2 rows in the first of its two chunks
  int nslot;
  uint32_t slot_bytes;
  WcChunk chunk[MAXCHUNK];
  const float* wl; int64_t layer_stride, cta_stride;     // floats
  const float* first; const float* head1; const float* head2;
  const float* cond; const float* uniforms; const float* test_inputs; int Ttest; float log_scale_min;
  float* ring; const int64_t* ring_off;
  float* out; float* logits;
};

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_remote_tagged(uint32_t cluster_addr, float v, uint32_t tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(cluster_addr), "l"(w) : "memory");
}
__device__ __forceinline__ float poll_local_tagged(const unsigned long long* p, uint32_t tag) {
  const uint32_t a = smem_u32(p);
  unsigned long long w;
  asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(w) : "r"(a) : "memory");
  if ((uint32_t)(w >> 32) != tag) {
    const long long t0 = clock64();
    do {
      asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(w) : "r"(a) : "memory");
      if (clock64() - t0 > 4000000000LL) __trap();
    } while ((uint32_t)(w >> 32) != tag);
  }
  return __uint_as_float((uint32_t)w);
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTC) : "memory"); }

// After the call v[0] of lane l holds the sum over the warp's lanes of the original v[l] (31 shuffles for 32 values).
__device__ __forceinline__ void wc_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float send = up ? v[j] : v[j + off];
      const float keep = up ? v[j + off] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

__global__ void __launch_bounds__(NTHR, 1) wavenet_cluster_kernel(const WcParams p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = (int)cluster_rank();
  const int B = p.B, rows1 = 2 * p.pairs, rows2 = p.srows + p.orows, xlen = p.xlen;
  // ---- shared memory carve-up ----
  uint8_t* slots = smraw;                                                            // nslot x slot_bytes
  unsigned long long* gvec = reinterpret_cast<unsigned long long*>(slots + (size_t)p.nslot * p.slot_bytes);   // [B][K2]
  unsigned long long* xvec = gvec + B * p.K2;                                        // [B][R]
  unsigned long long* svec = xvec + B * p.R;                                         // [B][S]
  unsigned long long* hvec = svec + B * p.S;                                         // [B][S]
  uint64_t* full = reinterpret_cast<uint64_t*>(hvec + B * p.S);                      // [nslot]
  uint64_t* empty = full + p.nslot;
  float* xb0 = reinterpret_cast<float*>(empty + p.nslot);                            // [B][xlen] x 2
  float* xb1 = xb0 + B * xlen;
  float* gs = xb1 + B * xlen;                                                        // [B][K2]
  float* part = gs + B * p.K2;                                                       // [rows1][NWC][MAXBC]
  float* res = part + rows1 * NWC * MAXBC;                                             // [max(rows1, rows2, O)][MAXBC]
  const int maxrows = max(max(rows1, rows2), p.O);
  float* skips = res + maxrows * MAXBC;                                              // [srows][MAXBC]
  float* first = skips + p.srows * MAXBC;                                            // [2R]
  float* h1w = first + 2 * p.R;                                                      // [hrows][S] + bias
  float* h2w = h1w + p.hrows * p.S + cpad4(p.hrows);                                 // [O][S] + bias
  float* vec = h2w + p.O * p.S + cpad4(p.O);                                         // [B][S]
  float* cur = vec + B * p.S;                                                        // [MAXBC]
  const float r2 = 0.70710678118654752440f;
  const uint32_t per_sample = 2u * (uint32_t)p.L + 2u;

  if (tid == 0) {
    for (int i = 0; i < p.nslot; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NWC); }
    fence_barrier_init();
  }
  for (int i = tid; i < B * (p.K2 + p.R + 2 * p.S); i += NTHR) gvec[i] = 0ull;       // tag 0 = never written
  for (int i = tid; i < 2 * p.R; i += NTHR) first[i] = p.first[i];
  for (int i = tid; i < p.hrows * p.S + p.hrows; i += NTHR) h1w[i] = p.head1[(size_t)cta * (p.hrows * p.S + cpad4(p.hrows)) + i];
  for (int i = tid; i < p.O * p.S + p.O; i += NTHR) h2w[i] = p.head2[i];
  if (tid < MAXBC) cur[tid] = 0.f;
  __syncthreads();
  cluster_sync_all();                 // nobody stores into a peer before that peer has zeroed its tagged buffers

  if (warp == NWC) {
    // ===== weight producer =====
    if (lane == 0) {
      const long long nch = (long long)p.T * p.L * p.nchunk;
      int slot = 0;
      uint32_t ph = 0;
      int l = 0, k = 0;
      for (long long n = 0; n < nch; ++n) {
        mbar_wait(&empty[slot], ph ^ 1u);
        const WcChunk c = p.chunk[k];
        mbar_expect_tx(&full[slot], c.bytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wl + (size_t)l * p.layer_stride + (size_t)cta * p.cta_stride) + c.off_bytes;
        bulk_load(slots + (size_t)slot * p.slot_bytes, src, c.bytes, &full[slot]);
        if (++slot == p.nslot) { slot = 0; ph ^= 1u; }
        if (++k == p.nchunk) { k = 0; if (++l == p.L) l = 0; }
      }
    }
  } else {
    // ===== consumers =====
    int slot = 0;
    uint32_t ph = 0;
    auto prefetch_x = [&](int layer, int t, float* xb) {      // older taps + conditioning of (layer, t)
      const int d = 1 << (layer % p.layers_per_stack);
      const int rl = (p.K - 1) * d + 1;
      const float* ring = p.ring + p.ring_off[layer];
      for (int b = 0; b < B; ++b) {
        float* x = xb + b * xlen;
        for (int j = 0; j < p.K - 1; ++j) {
          const int back = (p.K - 1 - j) * d;
          const int sl = ((t - back) % rl + rl) % rl;
          const float* s = ring + ((size_t)sl * B + b) * p.R;
          for (int i = tid * 4; i < p.R; i += NTC * 4) cp16(x + j * p.R + i, s + i);
        }
        const float* c = p.cond + ((size_t)b * p.T + t) * p.C;
        for (int i = tid * 4; i < p.C; i += NTC * 4) cp16(x + p.K * p.R + i, c + i);
      }
    };
    auto release_slot = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      if (++slot == p.nslot) { slot = 0; ph ^= 1u; }
    };
    // writes `n` values (value index j -> destination word index dst_word(j)) into the tagged buffer `buf` of ALL CTAs
    auto broadcast = [&](unsigned long long* buf, int n, uint32_t tag, auto value_of, auto word_of) {
      for (int i = tid; i < n * CL; i += NTC) {
        const int dst = i / n, j = i - dst * n;
        st_remote_tagged(map_to_cta(smem_u32(buf + word_of(j)), (uint32_t)dst), value_of(j), tag);
      }
    };

    if (p.T > 0) prefetch_x(0, 0, xb0);
    int buf = 0;
    for (int t = 0; t < p.T; ++t) {
      const uint32_t tag0 = 1u + (uint32_t)t * per_sample;
      if (p.test_inputs != nullptr && t < p.Ttest) {
        consumer_sync();
        if (tid < B) cur[tid] = p.test_inputs[(size_t)tid * p.Ttest + t];
        consumer_sync();
      }
      for (int l = 0; l < p.L; ++l) {
        const uint32_t tag_g = tag0 + 2u * (uint32_t)l, tag_x = tag_g + 1u;
        const int d = 1 << (l % p.layers_per_stack);
        const int rl = (p.K - 1) * d + 1;
        float* x = buf ? xb1 : xb0;
        // newest tap
        if (l == 0) {
          float* ring = p.ring + p.ring_off[0];
          for (int i = tid; i < B * p.R; i += NTC) {
            const int b = i / p.R, r = i - b * p.R;
            const float h = fmaf(first[r], cur[b], first[p.R + r]);
            x[b * xlen + (p.K - 1) * p.R + r] = h;
            if (cta == 0) ring[((size_t)(t % rl) * B + b) * p.R + r] = h;
          }
        } else {
          for (int i = tid; i < B * p.R; i += NTC) {
            const int b = i / p.R, r = i - b * p.R;
            x[b * xlen + (p.K - 1) * p.R + r] = poll_local_tagged(xvec + i, tag_x - 2u);
          }
        }
        cp_wait_all();
        consumer_sync();
        // ---- stage 1: gate rows, chunk by chunk ----
        // thread -> one float4 column of K1 (x is read once per chunk); the 32 gate rows of the CTA are 32 independent FMA
        // chains kept in registers across the 8 chunks -- no cross-lane traffic inside the loop (profiled: per-chunk shuffle
        // reductions were a third of the kernel's time); ONE transposing reduction per layer then leaves lane r with row r.
        constexpr int RPC = 4, NCH1 = 8;             // rows per chunk, gate chunks per layer (checked by the host: rows1 == 32)
        const int K4 = p.K1 >> 2;
        float acc[MAXBC][RPC * NCH1];
#pragma unroll
        for (int b = 0; b < MAXBC; ++b)
#pragma unroll
          for (int r = 0; r < RPC * NCH1; ++r) acc[b][r] = 0.f;
#pragma unroll
        for (int k = 0; k < NCH1; ++k) {
          mbar_wait(&full[slot], ph);
          const float* W = reinterpret_cast<const float*>(slots + (size_t)slot * p.slot_bytes);
          for (int kk = tid; kk < K4; kk += NTC) {
            float4 w[RPC];
#pragma unroll
            for (int r = 0; r < RPC; ++r) w[r] = reinterpret_cast<const float4*>(W + (size_t)r * p.K1)[kk];
#pragma unroll
            for (int b = 0; b < MAXBC; ++b)
              if (b < B) {
                const float4 v = *reinterpret_cast<const float4*>(x + (size_t)b * xlen + 4 * kk);
#pragma unroll
                for (int r = 0; r < RPC; ++r) {
                  float a = acc[b][k * RPC + r];
                  a = fmaf(w[r].x, v.x, a); a = fmaf(w[r].y, v.y, a); a = fmaf(w[r].z, v.z, a); a = fmaf(w[r].w, v.w, a);
                  acc[b][k * RPC + r] = a;
                }
              }
          }
          release_slot();
        }
#pragma unroll
        for (int b = 0; b < MAXBC; ++b)
          if (b < B) {
            wc_transpose_reduce32(acc[b], lane);                       // lane r now holds the warp's sum of row r
            part[(lane * NWC + warp) * MAXBC + b] = acc[b][0];
          }
        // the chunk that follows holds the gate biases (then the first half of the stage-2 rows): it stays mapped until stage 2 ends
        mbar_wait(&full[slot], ph);
        const int slotA = slot;
        const float* blkA = reinterpret_cast<const float*>(slots + (size_t)slotA * p.slot_bytes);
        const float* b1 = blkA;                                   // [rows1] (+pad)
        const float* W2a = blkA + cpad4(rows1);                   // rows2_first x K2
        consumer_sync();
        // gate: thread (b, row) sums the row's NWC partials; the (a, b) rows of a pair are adjacent lanes
        for (int i = tid; i < rows1 * B; i += NTC) {
          const int b = i / rows1, row = i - b * rows1;
          float a = b1[row];
#pragma unroll
          for (int w = 0; w < NWC; ++w) a += part[(row * NWC + w) * MAXBC + b];
          const float g = __shfl_down_sync(0xffffffffu, a, 1);          // rows1 % 32 == 0: full warps, pairs never straddle one
          if ((row & 1) == 0) res[(row >> 1) * MAXBC + b] = tanhf(a) * (1.f / (1.f + expf(-g)));
        }
        consumer_sync();
        broadcast(gvec, p.pairs * B, tag_g, [&](int j) { return res[(j / B) * MAXBC + (j % B)]; },
                  [&](int j) { return (j % B) * p.K2 + cta * p.pairs + (j / B); });
        // next stage-1 operands (older taps, conditioning) while the gate vectors are in flight
        {
          const int nl = (l + 1 == p.L) ? 0 : l + 1, nt = (l + 1 == p.L) ? t + 1 : t;
          if (nt < p.T) prefetch_x(nl, nt, buf ? xb0 : xb1);
        }
        // ---- stage 2: skip and residual rows ----
        for (int i = tid; i < B * p.K2; i += NTC) gs[i] = poll_local_tagged(gvec + i, tag_g);
        // second stage-2 chunk
        const int slotB = (slotA + 1 == p.nslot) ? 0 : slotA + 1;
        const uint32_t phB = (slotA + 1 == p.nslot) ? (ph ^ 1u) : ph;
        mbar_wait(&full[slotB], phB);
        const float* blkB = reinterpret_cast<const float*>(slots + (size_t)slotB * p.slot_bytes);
        const float* W2b = blkB;                                  // (rows2 - rows2_first) x K2, then the stage-2 biases
        const float* b2 = blkB + (size_t)(rows2 - p.rows2_first) * p.K2;
        consumer_sync();
        for (int row = warp; row < rows2; row += NWC) {
          const float* wr = (row < p.rows2_first) ? W2a + (size_t)row * p.K2 : W2b + (size_t)(row - p.rows2_first) * p.K2;
          float acc[MAXBC];
#pragma unroll
          for (int b = 0; b < MAXBC; ++b) acc[b] = 0.f;
          for (int kk = lane; kk < (p.K2 >> 2); kk += 32) {
            const float4 w = reinterpret_cast<const float4*>(wr)[kk];
#pragma unroll
            for (int b = 0; b < MAXBC; ++b)
              if (b < B) {
                const float4 v = *reinterpret_cast<const float4*>(gs + (size_t)b * p.K2 + 4 * kk);
                acc[b] = fmaf(w.x, v.x, acc[b]); acc[b] = fmaf(w.y, v.y, acc[b]);
                acc[b] = fmaf(w.z, v.z, acc[b]); acc[b] = fmaf(w.w, v.w, acc[b]);
              }
          }
#pragma unroll
          for (int b = 0; b < MAXBC; ++b)
            if (b < B) {
              const float s = warp_sum(acc[b]);
              if (lane == 0) res[row * MAXBC + b] = s + b2[row];
            }
        }
        consumer_sync();
        // epilogue: skip accumulation; residual output -> ring (later samples) + every CTA's xvec (next layer, now)
        for (int i = tid; i < rows2 * B; i += NTC) {
          const int row = i / B, b = i - row * B;
          const float v = res[row * MAXBC + b];
          if (row < p.srows) {
            skips[row * MAXBC + b] = (l == 0) ? v : (skips[row * MAXBC + b] + v) * r2;
          } else if (l + 1 < p.L) {
            const int r = cta * p.orows + (row - p.srows);
            const float xo = (v + x[b * xlen + (p.K - 1) * p.R + r]) * r2;
            res[row * MAXBC + b] = xo;
            const int dn = 1 << ((l + 1) % p.layers_per_stack);
            const int rln = (p.K - 1) * dn + 1;
            p.ring[p.ring_off[l + 1] + ((size_t)(t % rln) * B + b) * p.R + r] = xo;
          }
        }
        consumer_sync();
        if (l + 1 < p.L) {
          broadcast(xvec, p.orows * B, tag_x, [&](int j) { return res[(p.srows + j / B) * MAXBC + (j % B)]; },
                    [&](int j) { return (j % B) * p.R + cta * p.orows + (j / B); });
        }
        // both stage-2 slots are free now
        release_slot();
        release_slot();
        buf ^= 1;
      }
      // ---- head: relu(skips) -> 1x1 (S -> S) -> relu -> 1x1 (S -> O), sampled redundantly by every CTA ----
      const uint32_t tag_s = tag0 + per_sample - 2u, tag_h = tag0 + per_sample - 1u;
      for (int i = tid; i < p.srows * B; i += NTC) res[i] = fmaxf(skips[(i / B) * MAXBC + (i % B)], 0.f);
      consumer_sync();
      broadcast(svec, p.srows * B, tag_s, [&](int j) { return res[j]; }, [&](int j) { return (j % B) * p.S + cta * p.srows + (j / B); });
      for (int i = tid; i < B * p.S; i += NTC) vec[i] = poll_local_tagged(svec + i, tag_s);
      consumer_sync();
      for (int row = warp; row < p.hrows; row += NWC) {
        float acc[MAXBC];
#pragma unroll
        for (int b = 0; b < MAXBC; ++b) acc[b] = 0.f;
        for (int kk = lane; kk < (p.S >> 2); kk += 32) {
          const float4 w = reinterpret_cast<const float4*>(h1w + (size_t)row * p.S)[kk];
#pragma unroll
          for (int b = 0; b < MAXBC; ++b)
            if (b < B) {
              const float4 v = *reinterpret_cast<const float4*>(vec + (size_t)b * p.S + 4 * kk);
              acc[b] = fmaf(w.x, v.x, acc[b]); acc[b] = fmaf(w.y, v.y, acc[b]);
              acc[b] = fmaf(w.z, v.z, acc[b]); acc[b] = fmaf(w.w, v.w, acc[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < MAXBC; ++b)
          if (b < B) {
            const float s = warp_sum(acc[b]);
            if (lane == 0) res[row * MAXBC + b] = fmaxf(s + h1w[p.hrows * p.S + row], 0.f);
          }
      }
      __threadfence();                 // release: this sample's ring-buffer stores precede the tagged words below
      consumer_sync();
      broadcast(hvec, p.hrows * B, tag_h, [&](int j) { return res[(j / B) * MAXBC + (j % B)]; },
                [&](int j) { return (j % B) * p.S + cta * p.hrows + (j / B); });
      for (int i = tid; i < B * p.S; i += NTC) vec[i] = poll_local_tagged(hvec + i, tag_h);
      __threadfence();                 // acquire: every CTA's ring-buffer stores of this sample are visible from here on
      consumer_sync();
      for (int row = warp; row < p.O; row += NWC) {
        float acc[MAXBC];
#pragma unroll
        for (int b = 0; b < MAXBC; ++b) acc[b] = 0.f;
        for (int kk = lane; kk < (p.S >> 2); kk += 32) {
          const float4 w = reinterpret_cast<const float4*>(h2w + (size_t)row * p.S)[kk];
#pragma unroll
          for (int b = 0; b < MAXBC; ++b)
            if (b < B) {
              const float4 v = *reinterpret_cast<const float4*>(vec + (size_t)b * p.S + 4 * kk);
              acc[b] = fmaf(w.x, v.x, acc[b]); acc[b] = fmaf(w.y, v.y, acc[b]);
              acc[b] = fmaf(w.z, v.z, acc[b]); acc[b] = fmaf(w.w, v.w, acc[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < MAXBC; ++b)
          if (b < B) {
            const float s = warp_sum(acc[b]);
            if (lane == 0) res[row * MAXBC + b] = s + h2w[p.O * p.S + row];
          }
      }
      consumer_sync();
      if (tid < B) {                   // mixture.py:117-153 (same expressions as wavenet_synth.cu)
        const int b = tid, nm = p.O / 3;
        const float* u = p.uniforms + ((size_t)t * B + b) * (nm + 1);
        int arg = 0;
        float best = -INFINITY;
        for (int m = 0; m < nm; ++m) {
          const float v = res[m * MAXBC + b] - logf(-logf(u[m]));
          if (v > best) { best = v; arg = m; }
        }
        const float mean = res[(nm + arg) * MAXBC + b];
        const float ls = fmaxf(res[(2 * nm + arg) * MAXBC + b], p.log_scale_min);
        const float ul = u[nm];
        float xs = mean + expf(ls) * (logf(ul) - logf(1.f - ul));
        xs = fminf(fmaxf(xs, -1.f), 1.f);
        cur[b] = xs;
        if (cta == 0) p.out[(size_t)b * p.T + t] = xs;
      }
      if (cta == 0 && p.logits != nullptr) {
        for (int i = tid; i < p.O * B; i += NTC) {
          const int o = i / B, b = i - o * B;
          p.logits[((size_t)b * p.T + t) * p.O + o] = res[o * MAXBC + b];
        }
      }
      consumer_sync();
    }
    cp_wait_all();
  }
  __syncthreads();
  cluster_sync_all();                  // no CTA leaves while a peer may still store into its shared memory
}

struct WcPlan {
  int pairs, srows, orows, hrows, K1, K2, xlen, rows1, rows2, rows1_per_chunk, nchunk1, nchunk, rows2_first, nslot;
  uint32_t slot_bytes;
  WcChunk chunk[MAXCHUNK];
  size_t smem;
};

// Returns false when the configuration does not fit this path.
bool wc_plan(int R, int G, int S, int C, int K, int O, int B, WcPlan& q) {
  if (R % 4 || (G / 2) % 4 || S % 4 || C % 4 || G % 2 || O % 3 || B < 1 || B > MAXBC || K < 1) return false;
  if ((G / 2) % CL || S % CL || R % CL) return false;
  q.pairs = (G / 2) / CL; q.srows = S / CL; q.orows = R / CL; q.hrows = S / CL;
  q.K1 = K * R + C; q.K2 = G / 2; q.xlen = (q.K1 + 3) & ~3;
  q.rows1 = 2 * q.pairs; q.rows2 = q.srows + q.orows;
  q.rows1_per_chunk = 4;
  if (q.rows1 != 32) return false;       // 8 chunks of 4 gate rows, one lane per row in the reductions (the 512-channel WaveNet)
  if ((q.K1 * 4) % 16) return false;
  q.nchunk1 = q.rows1 / q.rows1_per_chunk;
  q.nchunk = q.nchunk1 + 2;
  if (q.nchunk > MAXCHUNK) return false;
  const uint32_t row1_bytes = (uint32_t)q.K1 * 4;
  uint32_t off = 0, mx = 0;
  for (int k = 0; k < q.nchunk1; ++k) {
    q.chunk[k].off_bytes = off; q.chunk[k].bytes = row1_bytes * q.rows1_per_chunk;
    off += q.chunk[k].bytes;
  }
  q.rows2_first = q.rows2 / 2;
  // [b1 (padded)] [first rows of W2]   and   [remaining rows of W2] [b2 (padded)]
  q.chunk[q.nchunk1].off_bytes = off;
  q.chunk[q.nchunk1].bytes = (uint32_t)(cpad4(q.rows1) + q.rows2_first * q.K2) * 4;
  off += q.chunk[q.nchunk1].bytes;
  q.chunk[q.nchunk1 + 1].off_bytes = off;
  q.chunk[q.nchunk1 + 1].bytes = (uint32_t)((q.rows2 - q.rows2_first) * q.K2 + cpad4(q.rows2)) * 4;
  for (int k = 0; k < q.nchunk; ++k) {
    if (q.chunk[k].bytes % 16 || q.chunk[k].off_bytes % 16) return false;
    mx = q.chunk[k].bytes > mx ? q.chunk[k].bytes : mx;
  }
  q.slot_bytes = (mx + 127u) & ~127u;
  const int maxrows = q.rows1 > q.rows2 ? (q.rows1 > O ? q.rows1 : O) : (q.rows2 > O ? q.rows2 : O);
  const size_t fixed = (size_t)B * (q.K2 + R + 2 * S) * 8 + 64 * 8 + ((size_t)2 * B * q.xlen + (size_t)B * q.K2 + (size_t)q.rows1 * NWC * MAXBC +
                       (size_t)maxrows * MAXBC + (size_t)q.srows * MAXBC + 2 * R + (q.hrows * S + cpad4(q.hrows)) + (O * S + cpad4(O)) +
                       (size_t)B * S + MAXBC) * 4 + 256;
  for (q.nslot = 6; q.nslot >= 4; --q.nslot) {
    q.smem = (size_t)q.nslot * q.slot_bytes + fixed;
    if (q.smem <= 226 * 1024) return true;
  }
  return false;
}

}  // namespace

extern "C" int viai_wavenet_cluster_supported(int R, int G, int S, int C, int K, int O, int B) {
  WcPlan q;
  return wc_plan(R, G, S, C, K, O, B, q) ? CL : 0;
}

// Same operands as viai_wavenet_synth with parameter blocks packed for nC = 16 (viai_wavenet_cluster_supported() != 0); the
// exchange buffers live in shared memory, so only `ring` (zero-initialised) is caller-owned state.
extern "C" int viai_wavenet_synth_cluster(int L, int layers_per_stack, int R, int G, int S, int C, int K, int O, int B, int T,
                                          const float* packed_layers, const float* first, const float* head1, const float* head2,
                                          const float* cond, const float* uniforms, const float* test_inputs, int Ttest,
                                          float log_scale_min, float* ring, const int64_t* ring_off, float* out, float* logits,
                                          viai_stream_t stream) {
  VIAI_REQUIRE(packed_layers && first && head1 && head2 && cond && uniforms && ring && ring_off && out, "wavenet_synth_cluster: null argument");
  WcPlan q;
  VIAI_REQUIRE(wc_plan(R, G, S, C, K, O, B, q), "wavenet_synth_cluster: unsupported configuration");
  VIAI_REQUIRE(L >= 1 && layers_per_stack >= 1 && L % layers_per_stack == 0 && T >= 0, "wavenet_synth_cluster: bad layer configuration");
  VIAI_REQUIRE((int64_t)T * (2 * L + 2) < 4000000000LL, "wavenet_synth_cluster: T too large for the 32-bit stage tags");
  if (T == 0) return VIAI_OK;
  WcParams p;
  memset(&p, 0, sizeof(p));
  p.L = L; p.R = R; p.G = G; p.S = S; p.C = C; p.K = K; p.O = O; p.B = B; p.T = T; p.layers_per_stack = layers_per_stack;
  p.pairs = q.pairs; p.srows = q.srows; p.orows = q.orows; p.hrows = q.hrows; p.K1 = q.K1; p.K2 = q.K2; p.xlen = q.xlen;
  p.nchunk = q.nchunk; p.nchunk1 = q.nchunk1; p.rows1_per_chunk = q.rows1_per_chunk; p.rows2_first = q.rows2_first;
  p.nslot = q.nslot; p.slot_bytes = q.slot_bytes;
  for (int k = 0; k < q.nchunk; ++k) p.chunk[k] = q.chunk[k];
  p.cta_stride = (int64_t)q.rows1 * q.K1 + cpad4(q.rows1) + (int64_t)q.rows2 * q.K2 + cpad4(q.rows2);
  p.layer_stride = p.cta_stride * CL;
  p.wl = packed_layers; p.first = first; p.head1 = head1; p.head2 = head2; p.cond = cond; p.uniforms = uniforms;
  p.test_inputs = test_inputs; p.Ttest = test_inputs ? Ttest : 0; p.log_scale_min = log_scale_min;
  p.ring = ring; p.ring_off = ring_off; p.out = out; p.logits = logits;
  VIAI_CUDA(cudaFuncSetAttribute(wavenet_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.smem));
  VIAI_CUDA(cudaFuncSetAttribute(wavenet_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CL); cfg.blockDim = dim3(NTHR); cfg.dynamicSmemBytes = q.smem; cfg.stream = STR(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  VIAI_CUDA(cudaLaunchKernelEx(&cfg, wavenet_cluster_kernel, p));
  viai::g_launches.fetch_add(1, std::memory_order_relaxed);
  return VIAI_OK;
}
