// WaveNet teacher-forced (T-parallel) training path: the pieces around the tensor-core GEMMs.
//   * shiftcat: the "linearised" operand of a dilated causal Conv1d (+ the local-conditioning features) so that the layer's
//     gate pre-activation is ONE 1x1 implicit GEMM over (B*T) pixels (K = taps*R + Cc) on the tcgen05 path,
//   * glu: tanh(a) * sigmoid(b), axpby: alpha*a + beta*b (residual / skip accumulation, EMA),
//   * dmol: per-sample negative log-likelihood of the discretized mixture of logistics and its analytic gradient,
//   * masked_sum: (losses * mask).sum() / mask.sum() (or a plain sum), sequence_mask.
// Everything here is HBM-bound elementwise / row-wise work: 128-bit accesses, grids sized in multiples of the SM count.
#include "common.cuh"
#include <math.h>
using namespace viai;

namespace {
constexpr int THREADS = 256;
constexpr int MAX_MIX = 32;
inline int grid_for(int64_t total) { return (int)imin64(cdiv(total, THREADS), 16 * kNumSMs); }

// out[b, t, k*R + r] = x[b, t - (K-1-k)*d, r] (0 before the start);  out[b, t, K*R + j] = c[b, t, j];  zero up to Kpad.
// With a dropout mask (F.dropout on the convolution input, modules.py:173) x is read as x * mask * scale.
__global__ void __launch_bounds__(THREADS)
shiftcat_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ c, const float4* __restrict__ mask, float scale, int B,
                    int T, int R4, int C4, int K, int dil, int P4, float4* __restrict__ out) {
  const int64_t total = (int64_t)B * T * P4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % P4);
    const int64_t bt = i / P4;
    const int t = (int)(bt % T);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < K * R4) {
      const int k = col / R4, r = col - k * R4;
      const int ts = t - (K - 1 - k) * dil;
      if (ts >= 0) {
        v = __ldg(x + (bt - t + ts) * R4 + r);
        if (mask) {
          const float4 m = __ldg(mask + (bt - t + ts) * R4 + r);
          v.x *= m.x * scale; v.y *= m.y * scale; v.z *= m.z * scale; v.w *= m.w * scale;
        }
      }
    } else if (col < K * R4 + C4) {
      v = __ldg(c + bt * C4 + (col - K * R4));
    }
    out[i] = v;
  }
}

// dx[b, t, r] = sum_k dout[b, t + (K-1-k)*d, k*R + r] (inside the sequence);  dc[b, t, j] = dout[b, t, K*R + j]
__global__ void __launch_bounds__(THREADS)
shiftcat_bwd_kernel(const float4* __restrict__ dout, const float4* __restrict__ mask, float scale, int B, int T, int R4, int C4,
                    int K, int dil, int P4, float4* __restrict__ dx, float4* __restrict__ dc) {
  const int W4 = R4 + (dc ? C4 : 0);
  const int64_t total = (int64_t)B * T * W4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % W4);
    const int64_t bt = i / W4;
    const int t = (int)(bt % T);
    if (col < R4) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < K; ++k) {
        const int td = t + (K - 1 - k) * dil;
        if (td < T) {
          const float4 g = __ldg(dout + (bt - t + td) * P4 + k * R4 + col);
          s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
        }
      }
      if (mask) {
        const float4 m = __ldg(mask + bt * R4 + col);
        s.x *= m.x * scale; s.y *= m.y * scale; s.z *= m.z * scale; s.w *= m.w * scale;
      }
      dx[bt * R4 + col] = s;
    } else {
      dc[bt * C4 + (col - R4)] = __ldg(dout + bt * P4 + K * R4 + (col - R4));
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// out[row, i] = tanh(y[row, i]) * sigmoid(y[row, H + i]),  H = G / 2
__global__ void __launch_bounds__(THREADS)
glu_fwd_kernel(const float4* __restrict__ y, int64_t rows, int H4, float4* __restrict__ out) {
  const int64_t total = rows * H4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / H4;
    const int col = (int)(i - row * H4);
    const float4 a = __ldg(y + row * 2 * H4 + col), b = __ldg(y + row * 2 * H4 + H4 + col);
    out[i] = make_float4(tanhf(a.x) * sigmoidf_(b.x), tanhf(a.y) * sigmoidf_(b.y), tanhf(a.z) * sigmoidf_(b.z),
                         tanhf(a.w) * sigmoidf_(b.w));
  }
}

__device__ __forceinline__ void glu_grad(float a, float b, float g, float& da, float& db) {
  const float th = tanhf(a), s = sigmoidf_(b);
  da = g * s * (1.f - th * th);
  db = g * th * s * (1.f - s);
}

__global__ void __launch_bounds__(THREADS)
glu_bwd_kernel(const float4* __restrict__ y, const float4* __restrict__ dout, int64_t rows, int H4, float4* __restrict__ dy) {
  const int64_t total = rows * H4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / H4;
    const int col = (int)(i - row * H4);
    const float4 a = __ldg(y + row * 2 * H4 + col), b = __ldg(y + row * 2 * H4 + H4 + col), g = __ldg(dout + i);
    float4 da, db;
    glu_grad(a.x, b.x, g.x, da.x, db.x);
    glu_grad(a.y, b.y, g.y, da.y, db.y);
    glu_grad(a.z, b.z, g.z, da.z, db.z);
    glu_grad(a.w, b.w, g.w, da.w, db.w);
    dy[row * 2 * H4 + col] = da;
    dy[row * 2 * H4 + H4 + col] = db;
  }
}

__global__ void __launch_bounds__(THREADS)
axpby_kernel(const float* __restrict__ a, float alpha, const float* __restrict__ b, float beta, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = b ? alpha * a[i] + beta * b[i] : alpha * a[i];
}

// ---- discretized mixture of logistics (wavenet_vocoder/mixture.py:25-105) -------------------------------------------------
__device__ __forceinline__ float softplusf_(float v) { return v > 20.f ? v : log1pf(expf(v)); }   // F.softplus (threshold 20)

struct MixTerm {
  float lp;      // log-probability of the bin under this component (before the mixture weight)
  float dm;      // d lp / d mean
  float ds;      // d lp / d log_scale (already gated by the clamp at log_scale_min)
};

__device__ __forceinline__ MixTerm mix_term(float y, float mean, float ls_raw, float half_bin, float lsm, float log_half_classes) {
  const float ls = fmaxf(ls_raw, lsm);
  const float cen = y - mean;
  const float inv = expf(-ls);
  const float plus_in = inv * (cen + half_bin), min_in = inv * (cen - half_bin), mid_in = inv * cen;
  MixTerm r;
  if (y < -0.999f) {                                   // left edge: log sigmoid(plus_in)
    r.lp = plus_in - softplusf_(plus_in);
    const float d = 1.f - sigmoidf_(plus_in);
    r.dm = -inv * d;
    r.ds = -plus_in * d;
  } else if (y > 0.999f) {                             // right edge: log(1 - sigmoid(min_in))
    r.lp = -softplusf_(min_in);
    const float d = -sigmoidf_(min_in);
    r.dm = -inv * d;
    r.ds = -min_in * d;
  } else {
    const float cp = sigmoidf_(plus_in), cm = sigmoidf_(min_in);
    const float delta = cp - cm;
    if (delta > 1e-5f) {
      r.lp = logf(fmaxf(delta, 1e-12f));
      const float dp = cp * (1.f - cp), dn = cm * (1.f - cm);
      r.dm = -inv * (dp - dn) / delta;
      r.ds = -(dp * plus_in - dn * min_in) / delta;
    } else {                                           // log pdf at the bin centre
      r.lp = mid_in - ls - 2.f * softplusf_(mid_in) - log_half_classes;
      const float d = 1.f - 2.f * sigmoidf_(mid_in);
      r.dm = -inv * d;
      r.ds = -mid_in * d - 1.f;
    }
  }
  if (ls_raw < lsm) r.ds = 0.f;                        // torch.clamp(min=): gradient passes where x >= min
  return r;
}

// one thread per (b, t) row of y_hat (3*nr_mix floats: logits | means | log scales).  dnll == nullptr: forward only.
__global__ void __launch_bounds__(THREADS)
dmol_kernel(const float* __restrict__ y_hat, const float* __restrict__ target, int64_t rows, int nm, float half_bin, float lsm,
            float log_half_classes, float* __restrict__ nll, const float* __restrict__ dnll, float* __restrict__ dy_hat) {
  for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
    const float* p = y_hat + row * 3 * nm;
    const float y = __ldg(target + row);
    float lmax = -INFINITY;
    for (int i = 0; i < nm; ++i) lmax = fmaxf(lmax, __ldg(p + i));
    float lsum = 0.f;
    for (int i = 0; i < nm; ++i) lsum += expf(__ldg(p + i) - lmax);
    const float llse = lmax + logf(lsum);              // log_softmax(logits) = logit - llse
    float lp[MAX_MIX];
    float m = -INFINITY;
    for (int i = 0; i < nm; ++i) {
      const MixTerm t = mix_term(y, __ldg(p + nm + i), __ldg(p + 2 * nm + i), half_bin, lsm, log_half_classes);
      lp[i] = t.lp + (__ldg(p + i) - llse);
      m = fmaxf(m, lp[i]);
    }
    float s = 0.f;
    for (int i = 0; i < nm; ++i) s += expf(lp[i] - m);
    const float lse = m + logf(s);
    if (nll) nll[row] = -lse;
    if (dy_hat) {
      const float g = __ldg(dnll + row);
      float* d = dy_hat + row * 3 * nm;
      for (int i = 0; i < nm; ++i) {
        const float w = expf(lp[i] - lse);             // responsibility of component i
        const float prior = expf(__ldg(p + i) - llse);
        const MixTerm t = mix_term(y, __ldg(p + nm + i), __ldg(p + 2 * nm + i), half_bin, lsm, log_half_classes);
        d[i] = g * (prior - w);
        d[nm + i] = -g * w * t.dm;
        d[2 * nm + i] = -g * w * t.ds;
      }
    }
  }
}

// sample_from_discretized_mix_logistic (mixture.py:117-153): Gumbel-max over the mixture logits with the supplied uniforms
// u[row, 0:nm], logistic sample from u[row, nm], clamped to [-1, 1].  (The synthesis kernel has the same arithmetic inline.)
__global__ void __launch_bounds__(THREADS)
dmol_sample_kernel(const float* __restrict__ y_hat, const float* __restrict__ u, int64_t rows, int nm, float lsm, float* __restrict__ out) {
  for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
    const float* p = y_hat + row * 3 * nm;
    const float* ur = u + row * (nm + 1);
    int best = 0;
    float bv = -INFINITY;
    for (int i = 0; i < nm; ++i) {
      const float v = __ldg(p + i) - logf(-logf(__ldg(ur + i)));
      if (v > bv) { bv = v; best = i; }
    }
    const float mean = __ldg(p + nm + best), ls = fmaxf(__ldg(p + 2 * nm + best), lsm);
    const float ul = __ldg(ur + nm);
    const float x = mean + expf(ls) * (logf(ul) - logf(1.f - ul));
    out[row] = fminf(fmaxf(x, -1.f), 1.f);
  }
}

// acc[0] = sum v*m, acc[1] = sum m
__global__ void __launch_bounds__(THREADS)
masked_sum_kernel(const float* __restrict__ v, const float* __restrict__ mask, int64_t n, double* acc) {
  __shared__ double sh[2][THREADS / 32];
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mask ? mask[i] : 1.f;
    s0 += (double)(v[i] * m);
    s1 += (double)m;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s0; sh[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) { t0 += sh[0][w]; t1 += sh[1][w]; }
    atomicAdd(acc, t0);
    atomicAdd(acc + 1, t1);
  }
}
__global__ void masked_sum_finish_kernel(const double* acc, int mean, float* out) {
  *out = (float)(mean ? acc[0] / acc[1] : acc[0]);
}
__global__ void __launch_bounds__(THREADS)
masked_sum_bwd_kernel(const float* __restrict__ mask, int64_t n, const double* __restrict__ acc, const float* __restrict__ gout,
                      int mean, float* __restrict__ dv) {
  const float g = mean ? (float)((double)__ldg(gout) / acc[1]) : __ldg(gout);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dv[i] = mask ? g * mask[i] : g;
}

// torch weight_norm (dim 0): w[row] = v[row] * g[row] / ||v[row]||.  One block per row.
__global__ void __launch_bounds__(128)
weight_norm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, int cols, float* __restrict__ w, float* __restrict__ nrm) {
  __shared__ float sh[4];
  const int row = blockIdx.x;
  const float* pv = v + (int64_t)row * cols;
  float s = 0.f;
  for (int k = threadIdx.x; k < cols; k += 128) s += pv[k] * pv[k];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  const float n = sqrtf(sh[0] + sh[1] + sh[2] + sh[3]);
  const float f = g[row] / n;
  for (int k = threadIdx.x; k < cols; k += 128) w[(int64_t)row * cols + k] = pv[k] * f;
  if (threadIdx.x == 0) nrm[row] = n;
}
// dg[row] = <dw, v> / n;  dv = (g / n) * dw - (g <dw, v> / n^3) * v
__global__ void __launch_bounds__(128)
weight_norm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ nrm,
                       const float* __restrict__ dw, int cols, float* __restrict__ dv, float* __restrict__ dg) {
  __shared__ float sh[4];
  const int row = blockIdx.x;
  const float* pv = v + (int64_t)row * cols;
  const float* pd = dw + (int64_t)row * cols;
  float s = 0.f;
  for (int k = threadIdx.x; k < cols; k += 128) s += pv[k] * pd[k];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  const float dot = sh[0] + sh[1] + sh[2] + sh[3];
  const float n = nrm[row], gg = g[row];
  const float a = gg / n, b = gg * dot / (n * n * n);
  for (int k = threadIdx.x; k < cols; k += 128) dv[(int64_t)row * cols + k] = a * pd[k] - b * pv[k];
  if (threadIdx.x == 0) dg[row] = dot / n;
}

__global__ void sequence_mask_kernel(const int64_t* __restrict__ lengths, int B, int T, float* __restrict__ out) {
  const int64_t total = (int64_t)B * T;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (i % T) < lengths[i / T] ? 1.f : 0.f;
}
}  // namespace

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int viai_shiftcat_fwd(const float* x, const float* c, const float* mask, float scale, int B, int T, int R, int Cc, int K,
                                 int dilation, int Kpad, float* out, viai_stream_t stream) {
  VIAI_REQUIRE(x && out && B > 0 && T > 0 && K > 0 && dilation > 0, "shiftcat_fwd: bad arguments");
  VIAI_REQUIRE(R > 0 && R % 4 == 0 && Cc >= 0 && Cc % 4 == 0 && Kpad % 4 == 0 && Kpad >= K * R + Cc && (Cc == 0 || c),
               "shiftcat_fwd: channel counts must be multiples of 4 and Kpad >= K*R + Cc (R %d Cc %d K %d Kpad %d)", R, Cc, K, Kpad);
  VIAI_REQUIRE(aligned16(x) && aligned16(out) && aligned16(c) && aligned16(mask), "shiftcat_fwd: pointers must be 16-byte aligned");
  shiftcat_fwd_kernel<<<grid_for((int64_t)B * T * (Kpad / 4)), THREADS, 0, STR(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(c), reinterpret_cast<const float4*>(mask), scale, B, T, R / 4,
      Cc / 4, K, dilation, Kpad / 4, reinterpret_cast<float4*>(out));
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_shiftcat_bwd(const float* dout, const float* mask, float scale, int B, int T, int R, int Cc, int K, int dilation,
                                 int Kpad, float* dx, float* dc, viai_stream_t stream) {
  VIAI_REQUIRE(dout && dx && B > 0 && T > 0 && K > 0 && dilation > 0, "shiftcat_bwd: bad arguments");
  VIAI_REQUIRE(R > 0 && R % 4 == 0 && Cc >= 0 && Cc % 4 == 0 && Kpad % 4 == 0 && Kpad >= K * R + Cc && (Cc > 0 || !dc),
               "shiftcat_bwd: channel counts must be multiples of 4 and Kpad >= K*R + Cc");
  VIAI_REQUIRE(aligned16(dout) && aligned16(dx) && aligned16(dc) && aligned16(mask), "shiftcat_bwd: pointers must be 16-byte aligned");
  const int W4 = R / 4 + (dc ? Cc / 4 : 0);
  shiftcat_bwd_kernel<<<grid_for((int64_t)B * T * W4), THREADS, 0, STR(stream)>>>(
      reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(mask), scale, B, T, R / 4, Cc / 4, K, dilation, Kpad / 4,
      reinterpret_cast<float4*>(dx), reinterpret_cast<float4*>(dc));
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_glu_fwd(const float* y, int64_t rows, int G, float* out, viai_stream_t stream) {
  VIAI_REQUIRE(y && out && rows > 0 && G > 0 && G % 8 == 0, "glu_fwd: gate channels must be a multiple of 8 (G %d)", G);
  VIAI_REQUIRE(aligned16(y) && aligned16(out), "glu_fwd: pointers must be 16-byte aligned");
  glu_fwd_kernel<<<grid_for(rows * (G / 8)), THREADS, 0, STR(stream)>>>(reinterpret_cast<const float4*>(y), rows, G / 8,
                                                                       reinterpret_cast<float4*>(out));
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_glu_bwd(const float* y, const float* dout, int64_t rows, int G, float* dy, viai_stream_t stream) {
  VIAI_REQUIRE(y && dout && dy && rows > 0 && G > 0 && G % 8 == 0, "glu_bwd: gate channels must be a multiple of 8 (G %d)", G);
  VIAI_REQUIRE(aligned16(y) && aligned16(dout) && aligned16(dy), "glu_bwd: pointers must be 16-byte aligned");
  glu_bwd_kernel<<<grid_for(rows * (G / 8)), THREADS, 0, STR(stream)>>>(reinterpret_cast<const float4*>(y),
                                                                       reinterpret_cast<const float4*>(dout), rows, G / 8,
                                                                       reinterpret_cast<float4*>(dy));
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_axpby(const float* a, float alpha, const float* b, float beta, float* out, int64_t n, viai_stream_t stream) {
  VIAI_REQUIRE(a && out && n > 0, "axpby: bad arguments");
  axpby_kernel<<<grid_for(n), THREADS, 0, STR(stream)>>>(a, alpha, b, beta, out, n);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_dmol_nll(const float* y_hat, const float* target, int64_t rows, int nr_mix, int num_classes, float log_scale_min,
                             float* nll, const float* dnll, float* dy_hat, viai_stream_t stream) {
  VIAI_REQUIRE(y_hat && target && rows > 0 && nr_mix > 0 && nr_mix <= MAX_MIX && num_classes > 1,
               "dmol_nll: bad arguments (nr_mix %d, at most %d)", nr_mix, MAX_MIX);
  VIAI_REQUIRE(nll || dy_hat, "dmol_nll: nothing to compute");
  VIAI_REQUIRE((dy_hat == nullptr) == (dnll == nullptr), "dmol_nll: dnll and dy_hat go together");
  dmol_kernel<<<grid_for(rows), THREADS, 0, STR(stream)>>>(y_hat, target, rows, nr_mix, 1.f / (float)(num_classes - 1), log_scale_min,
                                                          (float)log((double)(num_classes - 1) / 2.0), nll, dnll, dy_hat);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_dmol_sample(const float* y_hat, const float* uniforms, int64_t rows, int nr_mix, float log_scale_min, float* out,
                                viai_stream_t stream) {
  VIAI_REQUIRE(y_hat && uniforms && out && rows > 0 && nr_mix > 0, "dmol_sample: bad arguments");
  dmol_sample_kernel<<<grid_for(rows), THREADS, 0, STR(stream)>>>(y_hat, uniforms, rows, nr_mix, log_scale_min, out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_masked_sum_fwd(const float* v, const float* mask, int64_t n, int mean, double* acc, float* out,
                                   viai_stream_t stream) {
  VIAI_REQUIRE(v && acc && out && n > 0, "masked_sum_fwd: bad arguments");
  cudaStream_t st = STR(stream);
  VIAI_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  masked_sum_kernel<<<(int)imin64(cdiv(n, THREADS), 4 * kNumSMs), THREADS, 0, st>>>(v, mask, n, acc);
  VIAI_LAUNCHED();
  masked_sum_finish_kernel<<<1, 1, 0, st>>>(acc, mean, out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_masked_sum_bwd(const float* mask, int64_t n, int mean, const double* acc, const float* gout, float* dv,
                                   viai_stream_t stream) {
  VIAI_REQUIRE(acc && gout && dv && n > 0, "masked_sum_bwd: bad arguments");
  masked_sum_bwd_kernel<<<grid_for(n), THREADS, 0, STR(stream)>>>(mask, n, acc, gout, mean, dv);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_sequence_mask(const int64_t* lengths, int B, int T, float* out, viai_stream_t stream) {
  VIAI_REQUIRE(lengths && out && B > 0 && T > 0, "sequence_mask: bad arguments");
  sequence_mask_kernel<<<grid_for((int64_t)B * T), THREADS, 0, STR(stream)>>>(lengths, B, T, out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_weight_norm_fwd(const float* v, const float* g, int rows, int cols, float* w, float* norms, viai_stream_t stream) {
  VIAI_REQUIRE(v && g && w && norms && rows > 0 && cols > 0, "weight_norm_fwd: bad arguments");
  weight_norm_fwd_kernel<<<rows, 128, 0, STR(stream)>>>(v, g, cols, w, norms);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_weight_norm_bwd(const float* v, const float* g, const float* norms, const float* dw, int rows, int cols, float* dv,
                                    float* dg, viai_stream_t stream) {
  VIAI_REQUIRE(v && g && norms && dw && dv && dg && rows > 0 && cols > 0, "weight_norm_bwd: bad arguments");
  weight_norm_bwd_kernel<<<rows, 128, 0, STR(stream)>>>(v, g, norms, dw, cols, dv, dg);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
