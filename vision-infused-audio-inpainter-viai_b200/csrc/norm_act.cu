// Batch / instance normalisation statistics, fused normalise+activation, and its two-pass backward (NHWC fp32).
#include "common.cuh"
#include <stdlib.h>
using namespace viai;

namespace {

constexpr int THREADS = 256;

// Walk direction of the HBM passes.  Consecutive kernels of a normalisation layer sweep the same tensors; when a tensor (pair)
// is larger than the 126 MB L2, a second front-to-back sweep finds nothing of the first one (a cyclic walk through an LRU-like
// cache evicts every line just before it is needed).  Each pass here therefore starts where the previous launch ENDED
// (viai::g_sweep_end), on lines that are still resident:
//   forward : conv epilogue writes z front to back -> apply walks BACK to front (and ends on the rows the next convolution's
//             first tiles read); thin convolution -> stats (back to front) -> apply (front to back);
//   backward: the data-gradient kernel writes dz front to back -> bwd_reduce back to front -> bwd_apply front to back.
// Block (x, y) of a reversed kernel takes the rows of block (gridDim.x-1-x, gridDim.y-1-y); results are unchanged (every block
// owns the same row range as before, only WHEN it runs differs).  VIAI_NORM_WALK=0 restores all-forward sweeps (measurement).
// Measured (B200, C2 step, CUDA-graph replay, ms/step; tensors of the step are 256 / 128 / 64 / 32 / 16 / ... MiB): one box: no
// alternation 15.33, tensors >= 200 MiB 15.29, >= 100 MiB 15.26, >= 60 MiB 15.05; another box: >= 48 MiB 15.19, >= 30 MiB 15.08,
// >= 14 MiB 15.19, >= 6 MiB 15.25.  Small tensors LOSE from the reversal: they are L2-resident anyway, and block i of
// consecutive kernels lands on the same SM, hence on the L2 partition (die) that cached its rows the pass before -- the mirrored
// block order breaks that affinity.  So only tensors of at least VIAI_NORM_WALK_MB MiB (default 24) alternate.
std::atomic<int64_t> g_walk_mb{-1};        // < 0: not initialised yet
int64_t walk_min_bytes() {
  int64_t mb = g_walk_mb.load(std::memory_order_relaxed);
  if (mb < 0) {
    const char* e = getenv("VIAI_NORM_WALK_MB");
    mb = e ? (int64_t)atoll(e) : (int64_t)24;
    if (mb < 0) mb = 0;
    g_walk_mb.store(mb, std::memory_order_relaxed);
  }
  return mb << 20;
}
bool walk_alternate() {
  static const bool on = [] { const char* e = getenv("VIAI_NORM_WALK"); return !(e && e[0] == '0'); }();
  return on;
}
// -> 1: sweep back to front.  Call BEFORE the launch; after VIAI_LAUNCHED() (which records a forward sweep) call walk_done(rev).
int walk_pick(int64_t tensor_bytes) {
  return (walk_alternate() && tensor_bytes >= walk_min_bytes() && g_sweep_end.load(std::memory_order_relaxed) == 0) ? 1 : 0;
}
void walk_done(int rev) { g_sweep_end.store(rev, std::memory_order_relaxed); }

// thread -> (channel vector cv, row lane rl).  VEC channels per thread; cvec = C/VEC <= THREADS.
struct Lanes {
  int cvec, lanes;
};
__host__ __device__ inline Lanes make_lanes(int C, int VEC) {
  Lanes l;
  l.cvec = C / VEC;
  l.lanes = THREADS / l.cvec;
  if (l.lanes < 1) l.lanes = 1;
  return l;
}

// Reduces `NV` per-thread double values across the row lanes that share a channel vector, then atomically adds
// lane 0's totals to global double accumulators.
template <int NV>
__device__ __forceinline__ void block_reduce_lanes(double (&v)[NV], int cv, int rl, Lanes L, double* smem) {
  // smem: [lanes][cvec][NV]
  for (int k = 0; k < NV; ++k) smem[(rl * L.cvec + cv) * NV + k] = v[k];
  __syncthreads();
  if (rl == 0) {
    for (int k = 0; k < NV; ++k) {
      double t = 0.0;
      for (int l = 0; l < L.lanes; ++l) t += smem[(l * L.cvec + cv) * NV + k];
      v[k] = t;
    }
  }
}

template <int VEC, bool SQ>
__global__ void __launch_bounds__(THREADS)
stats_kernel(const float* __restrict__ y, int64_t rows_per_group, int C, double* __restrict__ sum, double* __restrict__ sumsq,
             int64_t rows_per_block, int rev) {
  extern __shared__ double sred[];
  const Lanes L = make_lanes(C, VEC);
  const int tid = threadIdx.x;
  const int cv = tid % L.cvec, rl = tid / L.cvec;
  const int g = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const bool active = rl < L.lanes;
  const int64_t r0 = (int64_t)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * rows_per_block;
  const int64_t r1 = imin64(r0 + rows_per_block, rows_per_group);
  // Sums and squares are accumulated in double (the square is formed in double too): the variance is later taken as
  // E[x^2] - E[x]^2, which is only safe for channels whose mean dwarfs their spread (ResNet features after residual adds)
  // when both moments are exact to ~1e-16.
  double s[VEC], q[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { s[k] = 0.0; q[k] = 0.0; }
  if (active) {
    const float* base = y + ((int64_t)g * rows_per_group) * C + cv * VEC;
#pragma unroll 4
    for (int64_t r = r0 + rl; r < r1; r += L.lanes) {
      float v[VEC];
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(base + r * C));
        v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
      } else {
        v[0] = __ldg(base + r * C);
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const double d = (double)v[k];
        s[k] += d;
        if (SQ) q[k] = fma(d, d, q[k]);
      }
    }
  }
  double v[2 * VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { v[k] = s[k]; v[VEC + k] = q[k]; }
  if (active) {
    for (int k = 0; k < 2 * VEC; ++k) sred[(rl * L.cvec + cv) * 2 * VEC + k] = v[k];
  }
  __syncthreads();
  if (active && rl == 0) {
    for (int k = 0; k < 2 * VEC; ++k) {
      double t = 0.0;
      for (int l = 0; l < L.lanes; ++l) t += sred[(l * L.cvec + cv) * 2 * VEC + k];
      v[k] = t;
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      atomicAdd(sum + (int64_t)g * C + cv * VEC + k, v[k]);
      if (SQ) atomicAdd(sumsq + (int64_t)g * C + cv * VEC + k, v[VEC + k]);
    }
  }
}

__global__ void finalize_kernel(const double* sum, const double* sumsq, int64_t cnt, int groups, int C, float eps,
                                float* mean, float* invstd, float* running_mean, float* running_var, float momentum,
                                int64_t* nbt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && nbt) *nbt += 1;
  if (i >= groups * C) return;
  double m = sum[i] / (double)cnt;
  double var = sumsq[i] / (double)cnt - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  invstd[i] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean && groups == 1) {
    double unb = cnt > 1 ? var * (double)cnt / (double)(cnt - 1) : var;
    running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * (float)m;
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unb;
  }
}

// Optional fusion of viai_norm_finalize into the forward apply pass (sum == nullptr: statistics come in as mean / invstd)
struct FusedFinalize {
  const double* sum; const double* sumsq; int64_t cnt; float eps;
  float* mean_out; float* invstd_out; float* running_mean; float* running_var; float momentum; int64_t* nbt;
};

// Elementwise kernels: a thread owns one channel vector (its per-channel constants live in registers) and walks rows of one
// statistics group -- no integer division and no parameter loads inside the loop; UNR rows are in flight per thread.
constexpr int UNR = 4;

template <int VEC>
__global__ void __launch_bounds__(THREADS)
apply_kernel(const float* __restrict__ y, int64_t rows_per_group, int C, const float* __restrict__ mean,
             const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta, int act,
             float slope, float* __restrict__ out, int64_t rows_per_block, FusedFinalize ff, int rev) {
  const Lanes L = make_lanes(C, VEC);
  const int cv = threadIdx.x % L.cvec, rl = threadIdx.x / L.cvec;
  if (rl >= L.lanes) return;
  const int g = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int64_t r0 = (int64_t)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * rows_per_block;
  const int64_t r1 = imin64(r0 + rows_per_block, rows_per_group);
  // out = act(((y - mu) * is) * ga + be): the subtraction comes first (as in the reference's batch_norm) so that channels
  // whose mean is large against their spread lose nothing to cancellation
  float mu[VEC], is[VEC], ga[VEC], be[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const int c = cv * VEC + k;
    if (ff.sum != nullptr) {
      // viai_norm_finalize fused in: every thread derives its channels' statistics from the double sums (the arithmetic of
      // finalize_kernel); the first row lane of block x = 0 publishes them for the backward pass and moves the running buffers
      const int64_t i = (int64_t)g * C + c;
      const double m = ff.sum[i] / (double)ff.cnt;
      double var = ff.sumsq[i] / (double)ff.cnt - m * m;
      if (var < 0.0) var = 0.0;
      mu[k] = (float)m;
      is[k] = (float)(1.0 / sqrt(var + (double)ff.eps));
      if (blockIdx.x == 0 && rl == 0) {            // one block per group (whichever rows it owns)
        ff.mean_out[i] = mu[k];
        ff.invstd_out[i] = is[k];
        if (ff.running_mean && gridDim.y == 1) {
          const double unb = ff.cnt > 1 ? var * (double)ff.cnt / (double)(ff.cnt - 1) : var;
          ff.running_mean[i] = (1.f - ff.momentum) * ff.running_mean[i] + ff.momentum * (float)m;
          ff.running_var[i] = (1.f - ff.momentum) * ff.running_var[i] + ff.momentum * (float)unb;
        }
        if (ff.nbt && i == 0) *ff.nbt += 1;
      }
    } else {
      mu[k] = mean ? mean[(int64_t)g * C + c] : 0.f;
      is[k] = mean ? invstd[(int64_t)g * C + c] : 1.f;
    }
    ga[k] = gamma ? gamma[c] : 1.f;
    be[k] = beta ? beta[c] : 0.f;
  }
  const int64_t off = ((int64_t)g * rows_per_group) * C + cv * VEC;
  for (int64_t r = r0 + rl; r < r1; r += (int64_t)L.lanes * UNR) {
    float v[UNR][VEC];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * L.lanes;
      if (rr < r1) {
        if (VEC == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(y + off + rr * C));
          v[u][0] = t.x; v[u][1 % VEC] = t.y; v[u][2 % VEC] = t.z; v[u][3 % VEC] = t.w;
        } else {
          v[u][0] = __ldg(y + off + rr * C);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * L.lanes;
      if (rr < r1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[u][k] = act_fwd(((v[u][k] - mu[k]) * is[k]) * ga[k] + be[k], act, slope);
        if (VEC == 4) *reinterpret_cast<float4*>(out + off + rr * C) = make_float4(v[u][0], v[u][1 % VEC], v[u][2 % VEC], v[u][3 % VEC]);
        else out[off + rr * C] = v[u][0];
      }
    }
  }
}

// g = dz * act'(pre), xhat; shared by both backward passes
__device__ __forceinline__ void bwd_terms(float dz, float yv, float mu, float is, float ga, float be, bool has_norm, int act,
                                          float slope, float& gout, float& xhat) {
  xhat = has_norm ? (yv - mu) * is : yv;
  float pre = xhat * ga + be;
  gout = dz * act_grad(pre, act, slope);
}

template <int VEC>
__global__ void __launch_bounds__(THREADS)
bwd_reduce_kernel(const float* __restrict__ dz, const float* __restrict__ y, int64_t rows_per_group, int C,
                  const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                  const float* __restrict__ beta, int act, float slope, double* __restrict__ s1, double* __restrict__ s2,
                  int64_t rows_per_block, int rev) {
  extern __shared__ double sred[];
  const Lanes L = make_lanes(C, VEC);
  const int tid = threadIdx.x;
  const int cv = tid % L.cvec, rl = tid / L.cvec;
  const int g = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const bool active = rl < L.lanes;
  const int64_t r0 = (int64_t)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * rows_per_block;
  const int64_t r1 = imin64(r0 + rows_per_block, rows_per_group);
  float a1[VEC], a2[VEC], mu[VEC], is[VEC], ga[VEC], be[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    a1[k] = 0.f; a2[k] = 0.f;
    int c = cv * VEC + k;
    bool ok = active && c < C;
    mu[k] = (ok && mean) ? mean[(int64_t)g * C + c] : 0.f;
    is[k] = (ok && mean) ? invstd[(int64_t)g * C + c] : 1.f;
    ga[k] = (ok && gamma) ? gamma[c] : 1.f;
    be[k] = (ok && beta) ? beta[c] : 0.f;
  }
  if (active) {
    const int64_t off = ((int64_t)g * rows_per_group) * C + cv * VEC;
    constexpr int UNR_R = 4;                       // rows in flight per thread
    for (int64_t r = r0 + rl; r < r1; r += (int64_t)L.lanes * UNR_R) {
      float d[UNR_R][VEC], yv[UNR_R][VEC];
#pragma unroll
      for (int u = 0; u < UNR_R; ++u) {
        const int64_t rr = r + (int64_t)u * L.lanes;
#pragma unroll
        for (int k = 0; k < VEC; ++k) { d[u][k] = 0.f; yv[u][k] = 0.f; }
        if (rr < r1) {
          if (VEC == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(dz + off + rr * C));
            const float4 w = __ldg(reinterpret_cast<const float4*>(y + off + rr * C));
            d[u][0] = t.x; d[u][1 % VEC] = t.y; d[u][2 % VEC] = t.z; d[u][3 % VEC] = t.w;
            yv[u][0] = w.x; yv[u][1 % VEC] = w.y; yv[u][2 % VEC] = w.z; yv[u][3 % VEC] = w.w;
          } else {
            d[u][0] = __ldg(dz + off + rr * C);
            yv[u][0] = __ldg(y + off + rr * C);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR_R; ++u) {
        if (r + (int64_t)u * L.lanes < r1) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            float gg, xh;
            bwd_terms(d[u][k], yv[u][k], mu[k], is[k], ga[k], be[k], mean != nullptr, act, slope, gg, xh);
            a1[k] += gg;
            a2[k] += gg * xh;
          }
        }
      }
    }
  }
  double v[2 * VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { v[k] = a1[k]; v[VEC + k] = a2[k]; }
  if (active)
    for (int k = 0; k < 2 * VEC; ++k) sred[(rl * L.cvec + cv) * 2 * VEC + k] = v[k];
  __syncthreads();
  if (active && rl == 0) {
    for (int k = 0; k < 2 * VEC; ++k) {
      double t = 0.0;
      for (int l = 0; l < L.lanes; ++l) t += sred[(l * L.cvec + cv) * 2 * VEC + k];
      v[k] = t;
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      atomicAdd(s1 + (int64_t)g * C + cv * VEC + k, v[k]);
      atomicAdd(s2 + (int64_t)g * C + cv * VEC + k, v[VEC + k]);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(THREADS)
bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ y, int64_t rows_per_group, int C,
                 const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int act, float slope, const double* __restrict__ s1,
                 const double* __restrict__ s2, float* __restrict__ dy, int64_t rows_per_block, float* __restrict__ dgamma,
                 float* __restrict__ dbeta, int accumulate, int rev) {
  if ((dgamma || dbeta) && blockIdx.x == 0 && blockIdx.y == 0) {
    // viai_fold_groups fused in: dgamma = sum over groups of s2, dbeta = of s1 (written or accumulated into the gradient bucket)
    for (int c = threadIdx.x; c < C; c += THREADS) {
      double t1 = 0.0, t2 = 0.0;
      for (int gg = 0; gg < (int)gridDim.y; ++gg) { t1 += s1[(int64_t)gg * C + c]; t2 += s2[(int64_t)gg * C + c]; }
      if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)t2;
      if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)t1;
    }
  }
  const Lanes L = make_lanes(C, VEC);
  const int cv = threadIdx.x % L.cvec, rl = threadIdx.x / L.cvec;
  if (rl >= L.lanes) return;
  const int g = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int64_t r0 = (int64_t)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * rows_per_block;
  const int64_t r1 = imin64(r0 + rows_per_block, rows_per_group);
  const float inv_cnt = 1.f / (float)rows_per_group;
  const bool has_norm = mean != nullptr;
  float mu[VEC], is[VEC], ga[VEC], be[VEC], m1[VEC], m2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const int c = cv * VEC + k;
    mu[k] = has_norm ? mean[(int64_t)g * C + c] : 0.f;
    is[k] = has_norm ? invstd[(int64_t)g * C + c] : 1.f;
    ga[k] = gamma ? gamma[c] : 1.f;
    be[k] = beta ? beta[c] : 0.f;
    m1[k] = has_norm ? (float)(s1[(int64_t)g * C + c]) * inv_cnt : 0.f;
    m2[k] = has_norm ? (float)(s2[(int64_t)g * C + c]) * inv_cnt : 0.f;
  }
  const int64_t off = ((int64_t)g * rows_per_group) * C + cv * VEC;
  for (int64_t r = r0 + rl; r < r1; r += (int64_t)L.lanes * UNR) {
    float d[UNR][VEC], yv[UNR][VEC];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * L.lanes;
      if (rr < r1) {
        if (VEC == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(dz + off + rr * C));
          const float4 w = __ldg(reinterpret_cast<const float4*>(y + off + rr * C));
          d[u][0] = t.x; d[u][1 % VEC] = t.y; d[u][2 % VEC] = t.z; d[u][3 % VEC] = t.w;
          yv[u][0] = w.x; yv[u][1 % VEC] = w.y; yv[u][2 % VEC] = w.z; yv[u][3 % VEC] = w.w;
        } else {
          d[u][0] = __ldg(dz + off + rr * C);
          yv[u][0] = __ldg(y + off + rr * C);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t rr = r + (int64_t)u * L.lanes;
      if (rr < r1) {
        float o[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          float gg, xh;
          bwd_terms(d[u][k], yv[u][k], mu[k], is[k], ga[k], be[k], has_norm, act, slope, gg, xh);
          o[k] = has_norm ? ga[k] * is[k] * (gg - m1[k] - xh * m2[k]) : gg * ga[k];
        }
        if (VEC == 4) *reinterpret_cast<float4*>(dy + off + rr * C) = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
        else dy[off + rr * C] = o[0];
      }
    }
  }
}

__global__ void fold_kernel(const double* sums, int groups, int C, float* out, int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t = 0.0;
  for (int g = 0; g < groups; ++g) t += sums[(int64_t)g * C + c];
  out[c] = (accumulate ? out[c] : 0.f) + (float)t;
}

__global__ void rsqrt_eps_kernel(const float* var, int n, float eps, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)(1.0 / sqrt((double)var[i] + (double)eps));
}

// rows handled by one block of a per-channel reduction: about four waves of blocks over the 148 SMs, never fewer than 8 rows
// per row lane
int64_t pick_rows_per_block(int64_t rows_per_group, int groups, int C, int VEC) {
  const int lanes = make_lanes(C, VEC).lanes;
  int64_t per_group = cdiv((int64_t)8 * kNumSMs, groups);
  int64_t rpb = cdiv(rows_per_group, per_group);
  // at least 32 rows per row lane: every block ends with C double atomics per moment, which dominate short tensors
  // (bias gradients of 32 000 x 512 rows: 1 143 blocks x 512 atomics took 60 us, HBM time is 11 us)
  if (rpb < (int64_t)lanes * 32) rpb = (int64_t)lanes * 32;
  return rpb;
}

// rows handled by one block of an elementwise pass: ~16 blocks per SM in total, at least UNR rows per row lane
int64_t pick_rows_per_block_elem(int64_t rows_per_group, int groups, int C, int VEC) {
  const int lanes = make_lanes(C, VEC).lanes;
  int64_t per_group = cdiv((int64_t)16 * kNumSMs, groups);
  int64_t rpb = cdiv(rows_per_group, per_group);
  if (rpb < (int64_t)lanes * UNR) rpb = (int64_t)lanes * UNR;
  return rpb;
}

int pick_vec(int C, const void* p0, const void* p1 = nullptr) {
  bool aligned = (reinterpret_cast<uintptr_t>(p0) % 16 == 0) && (p1 == nullptr || reinterpret_cast<uintptr_t>(p1) % 16 == 0);
  return (C % 4 == 0 && aligned) ? 4 : 1;
}

}  // namespace

extern "C" int viai_norm_walk_mb(int mb) {
  const int prev = (int)(walk_min_bytes() >> 20);
  if (mb >= 0) g_walk_mb.store(mb, std::memory_order_relaxed);
  return prev;
}

extern "C" int viai_channel_stats(const float* y, int64_t rows_per_group, int groups, int C, double* sum, double* sumsq,
                                  viai_stream_t stream) {
  VIAI_REQUIRE(y && sum && rows_per_group > 0 && groups > 0 && C > 0, "viai_channel_stats: bad arguments");
  cudaStream_t st = STR(stream);
  if (sumsq == sum + (size_t)groups * C) {        // one (2, groups*C) buffer: a single memset node
    VIAI_CUDA(cudaMemsetAsync(sum, 0, 2 * sizeof(double) * groups * C, st));
  } else {
    VIAI_CUDA(cudaMemsetAsync(sum, 0, sizeof(double) * groups * C, st));
    if (sumsq) VIAI_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double) * groups * C, st));
  }
  const int VEC = pick_vec(C, y);
  VIAI_REQUIRE(C / VEC <= THREADS, "viai_channel_stats: C=%d too large", C);
  const int64_t rpb = pick_rows_per_block(rows_per_group, groups, C, VEC);
  dim3 grid((unsigned)cdiv(rows_per_group, rpb), (unsigned)groups);
  size_t smem = sizeof(double) * THREADS * 2 * VEC;
  const int rev = walk_pick(rows_per_group * groups * (int64_t)C * 4);
  if (VEC == 4) {
    if (sumsq) stats_kernel<4, true><<<grid, THREADS, smem, st>>>(y, rows_per_group, C, sum, sumsq, rpb, rev);
    else stats_kernel<4, false><<<grid, THREADS, smem, st>>>(y, rows_per_group, C, sum, sumsq, rpb, rev);
  } else {
    if (sumsq) stats_kernel<1, true><<<grid, THREADS, smem, st>>>(y, rows_per_group, C, sum, sumsq, rpb, rev);
    else stats_kernel<1, false><<<grid, THREADS, smem, st>>>(y, rows_per_group, C, sum, sumsq, rpb, rev);
  }
  VIAI_LAUNCHED();
  walk_done(rev);
  return VIAI_OK;
}

extern "C" int viai_norm_finalize(const double* sum, const double* sumsq, int64_t rows_per_group, int groups, int C,
                                  float eps, float* mean, float* invstd, float* running_mean, float* running_var,
                                  float momentum, int64_t* num_batches_tracked, viai_stream_t stream) {
  VIAI_REQUIRE(sum && sumsq && mean && invstd && groups > 0 && C > 0, "viai_norm_finalize: bad arguments");
  int n = groups * C;
  finalize_kernel<<<(n + 255) / 256, 256, 0, STR(stream)>>>(sum, sumsq, rows_per_group, groups, C, eps, mean, invstd,
                                                          running_mean, running_var, momentum, num_batches_tracked);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_norm_act_fwd(const float* y, int64_t rows_per_group, int groups, int C, const float* mean,
                                 const float* invstd, const float* gamma, const float* beta, int act, float slope,
                                 float* out, viai_stream_t stream) {
  VIAI_REQUIRE(y && out && rows_per_group > 0 && groups > 0 && C > 0, "viai_norm_act_fwd: bad arguments");
  VIAI_REQUIRE((mean == nullptr) == (invstd == nullptr), "viai_norm_act_fwd: mean/invstd must both be set or NULL");
  int VEC = pick_vec(C, y, out);
  if (C / VEC > THREADS) VEC = 0;
  VIAI_REQUIRE(VEC != 0, "viai_norm_act_fwd: C=%d too large", C);
  // without statistics every row shares the same constants: treat the tensor as one group
  const int64_t rpg = mean ? rows_per_group : rows_per_group * groups;
  const int ngr = mean ? groups : 1;
  const int64_t rpb = pick_rows_per_block_elem(rpg, ngr, C, VEC);
  dim3 grid((unsigned)cdiv(rpg, rpb), (unsigned)ngr);
  FusedFinalize ff;
  memset(&ff, 0, sizeof(ff));
  const int rev = walk_pick(rpg * ngr * (int64_t)C * 4);
  if (VEC == 4) apply_kernel<4><<<grid, THREADS, 0, STR(stream)>>>(y, rpg, C, mean, invstd, gamma, beta, act, slope, out, rpb, ff, rev);
  else apply_kernel<1><<<grid, THREADS, 0, STR(stream)>>>(y, rpg, C, mean, invstd, gamma, beta, act, slope, out, rpb, ff, rev);
  VIAI_LAUNCHED();
  walk_done(rev);
  return VIAI_OK;
}

extern "C" int viai_norm_finalize_act_fwd(const float* y, int64_t rows_per_group, int groups, int C, const double* sum,
                                          const double* sumsq, float eps, const float* gamma, const float* beta, int act,
                                          float slope, float* out, float* mean, float* invstd, float* running_mean,
                                          float* running_var, float momentum, int64_t* num_batches_tracked, viai_stream_t stream) {
  VIAI_REQUIRE(y && out && sum && sumsq && mean && invstd && rows_per_group > 0 && groups > 0 && C > 0,
               "viai_norm_finalize_act_fwd: bad arguments");
  VIAI_REQUIRE(running_mean == nullptr || groups == 1, "viai_norm_finalize_act_fwd: running statistics need groups == 1");
  int VEC = pick_vec(C, y, out);
  VIAI_REQUIRE(C / VEC <= THREADS, "viai_norm_finalize_act_fwd: C=%d too large", C);
  const int64_t rpb = pick_rows_per_block_elem(rows_per_group, groups, C, VEC);
  dim3 grid((unsigned)cdiv(rows_per_group, rpb), (unsigned)groups);
  FusedFinalize ff = {sum, sumsq, rows_per_group, eps, mean, invstd, running_mean, running_var, momentum, num_batches_tracked};
  const int rev = walk_pick(rows_per_group * groups * (int64_t)C * 4);
  if (VEC == 4) apply_kernel<4><<<grid, THREADS, 0, STR(stream)>>>(y, rows_per_group, C, nullptr, nullptr, gamma, beta, act, slope, out, rpb, ff, rev);
  else apply_kernel<1><<<grid, THREADS, 0, STR(stream)>>>(y, rows_per_group, C, nullptr, nullptr, gamma, beta, act, slope, out, rpb, ff, rev);
  VIAI_LAUNCHED();
  walk_done(rev);
  return VIAI_OK;
}

extern "C" int viai_norm_act_bwd_reduce(const float* dz, const float* y, int64_t rows_per_group, int groups, int C,
                                        const float* mean, const float* invstd, const float* gamma, const float* beta,
                                        int act, float slope, double* s1, double* s2, viai_stream_t stream) {
  VIAI_REQUIRE(dz && y && s1 && s2 && rows_per_group > 0 && groups > 0 && C > 0, "viai_norm_act_bwd_reduce: bad arguments");
  cudaStream_t st = STR(stream);
  if (s2 == s1 + (size_t)groups * C) {
    VIAI_CUDA(cudaMemsetAsync(s1, 0, 2 * sizeof(double) * groups * C, st));
  } else {
    VIAI_CUDA(cudaMemsetAsync(s1, 0, sizeof(double) * groups * C, st));
    VIAI_CUDA(cudaMemsetAsync(s2, 0, sizeof(double) * groups * C, st));
  }
  const int VEC = pick_vec(C, dz, y);
  VIAI_REQUIRE(C / VEC <= THREADS, "viai_norm_act_bwd_reduce: C=%d too large", C);
  const int64_t rpb = pick_rows_per_block(rows_per_group, groups, C, VEC);
  dim3 grid((unsigned)cdiv(rows_per_group, rpb), (unsigned)groups);
  size_t smem = sizeof(double) * THREADS * 2 * VEC;
  const int rev = walk_pick(rows_per_group * groups * (int64_t)C * 4);
  if (VEC == 4) bwd_reduce_kernel<4><<<grid, THREADS, smem, st>>>(dz, y, rows_per_group, C, mean, invstd, gamma, beta, act, slope, s1, s2, rpb, rev);
  else bwd_reduce_kernel<1><<<grid, THREADS, smem, st>>>(dz, y, rows_per_group, C, mean, invstd, gamma, beta, act, slope, s1, s2, rpb, rev);
  VIAI_LAUNCHED();
  walk_done(rev);
  return VIAI_OK;
}

static int norm_act_bwd_apply_impl(const float* dz, const float* y, int64_t rows_per_group, int groups, int C,
                                   const float* mean, const float* invstd, const float* gamma, const float* beta,
                                   int act, float slope, const double* s1, const double* s2, float* dy, float* dgamma,
                                   float* dbeta, int fused_fold, int accumulate, viai_stream_t stream) {
  VIAI_REQUIRE(dz && y && dy && rows_per_group > 0 && groups > 0 && C > 0, "viai_norm_act_bwd_apply: bad arguments");
  VIAI_REQUIRE(mean == nullptr || (s1 && s2 && invstd), "viai_norm_act_bwd_apply: statistics missing");
  cudaStream_t st = STR(stream);
  const int VEC = pick_vec(C, dz, y) == 4 && pick_vec(C, dy) == 4 ? 4 : 1;
  VIAI_REQUIRE(C / VEC <= THREADS, "viai_norm_act_bwd_apply: C=%d too large", C);
  const int64_t rpg = mean ? rows_per_group : rows_per_group * groups;
  const int ngr = mean ? groups : 1;
  const int64_t rpb = pick_rows_per_block_elem(rpg, ngr, C, VEC);
  dim3 grid((unsigned)cdiv(rpg, rpb), (unsigned)ngr);
  const bool in_kernel = fused_fold && mean != nullptr && s1 && s2;      // (with statistics grid.y == groups: the kernel can fold)
  float* kg = in_kernel ? dgamma : nullptr;
  float* kb = in_kernel ? dbeta : nullptr;
  const int rev = walk_pick(rpg * ngr * (int64_t)C * 4);
  if (VEC == 4) bwd_apply_kernel<4><<<grid, THREADS, 0, st>>>(dz, y, rpg, C, mean, invstd, gamma, beta, act, slope, s1, s2, dy, rpb, kg, kb, accumulate, rev);
  else bwd_apply_kernel<1><<<grid, THREADS, 0, st>>>(dz, y, rpg, C, mean, invstd, gamma, beta, act, slope, s1, s2, dy, rpb, kg, kb, accumulate, rev);
  VIAI_LAUNCHED();
  walk_done(rev);
  if (in_kernel) return VIAI_OK;
  VIAI_REQUIRE(!fused_fold || (!dgamma && !dbeta), "viai_norm_act_bwd_apply_fold: dgamma / dbeta need the statistics path");
  if (dgamma && s2) {
    fold_kernel<<<(C + 255) / 256, 256, 0, st>>>(s2, groups, C, dgamma, 0);
    VIAI_LAUNCHED();
  }
  if (dbeta && s1) {
    fold_kernel<<<(C + 255) / 256, 256, 0, st>>>(s1, groups, C, dbeta, 0);
    VIAI_LAUNCHED();
  }
  return VIAI_OK;
}

extern "C" int viai_norm_act_bwd_apply(const float* dz, const float* y, int64_t rows_per_group, int groups, int C,
                                       const float* mean, const float* invstd, const float* gamma, const float* beta,
                                       int act, float slope, const double* s1, const double* s2, float* dy, float* dgamma,
                                       float* dbeta, viai_stream_t stream) {
  return norm_act_bwd_apply_impl(dz, y, rows_per_group, groups, C, mean, invstd, gamma, beta, act, slope, s1, s2, dy, dgamma, dbeta,
                                 0, 0, stream);
}

extern "C" int viai_norm_act_bwd_apply_fold(const float* dz, const float* y, int64_t rows_per_group, int groups, int C,
                                            const float* mean, const float* invstd, const float* gamma, const float* beta,
                                            int act, float slope, const double* s1, const double* s2, float* dy, float* dgamma,
                                            float* dbeta, int accumulate, viai_stream_t stream) {
  return norm_act_bwd_apply_impl(dz, y, rows_per_group, groups, C, mean, invstd, gamma, beta, act, slope, s1, s2, dy, dgamma, dbeta,
                                 1, accumulate, stream);
}

extern "C" int viai_fold_groups(const double* sums, int groups, int C, float* out, int accumulate, viai_stream_t stream) {
  VIAI_REQUIRE(sums && out && groups > 0 && C > 0, "viai_fold_groups: bad arguments");
  fold_kernel<<<(C + 255) / 256, 256, 0, STR(stream)>>>(sums, groups, C, out, accumulate);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_rsqrt_eps(const float* var, int n, float eps, float* out, viai_stream_t stream) {
  VIAI_REQUIRE(var && out && n > 0, "viai_rsqrt_eps: bad arguments");
  rsqrt_eps_kernel<<<(n + 255) / 256, 256, 0, STR(stream)>>>(var, n, eps, out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
