// GANLoss (MSE / BCE against a scalar label), L1 loss, fused Adam, fill.
#include "common.cuh"
using namespace viai;

namespace {
constexpr int THREADS = 256;

__device__ __forceinline__ float loss_term(int kind, float p, float q, float t) {
  if (kind == 0) { float d = p - t; return d * d; }
  if (kind == 1) {  // torch BCELoss clamps log at -100
    float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
    return -(t * lp + (1.f - t) * l1p);
  }
  return fabsf(p - q);
}

__global__ void __launch_bounds__(THREADS)
loss_fwd_kernel(int kind, const float* __restrict__ p, const float* __restrict__ q, float target, int64_t n, double* acc) {
  __shared__ double sh[THREADS / 32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += (double)loss_term(kind, p[i], q ? q[i] : 0.f, target);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) t += sh[w];
    atomicAdd(acc, t);
  }
}
__global__ void loss_finish_kernel(const double* acc, int64_t n, float* out) { *out = (float)(*acc / (double)n); }

__global__ void __launch_bounds__(THREADS)
loss_bwd_kernel(int kind, const float* __restrict__ p, const float* __restrict__ q, float target, int64_t n,
                const float* __restrict__ gout, float* __restrict__ dp) {
  const float g = __ldg(gout) / (float)n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = p[i], d;
    if (kind == 0) d = 2.f * (v - target);
    else if (kind == 1) d = (v - target) / fmaxf((1.f - v) * v, 1e-12f);
    else { float e = v - q[i]; d = (e > 0.f) ? 1.f : (e < 0.f ? -1.f : 0.f); }
    dp[i] = d * g;
  }
}

__global__ void adam_tick_kernel(float* step) { *step += 1.f; }

__global__ void __launch_bounds__(THREADS)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            const float* __restrict__ lr_dev, double b1d, double b2d, float eps, const float* __restrict__ step_dev, float gscale) {
  // bias corrections in double, like torch.optim.Adam's Python-float arithmetic
  const double step = (double)__ldg(step_dev);
  const float lr = __ldg(lr_dev);
  const float b1 = (float)b1d, b2 = (float)b2d;
  const float bc2s = (float)sqrt(1.0 - pow(b2d, step));
  const float step_size = (float)((double)lr / (1.0 - pow(b1d, step)));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = m[i] * b1 + (1.f - b1) * gi;
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2s + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

__global__ void lincomb2_kernel(const float* a, float wa, const float* b, float wb, float* out) {
  *out = wa * (*a) + (b ? wb * (*b) : 0.f);
}

__global__ void fill_kernel(float* p, int64_t n, float value) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = value;
}
inline int grid_for(int64_t total) { return (int)imin64(cdiv(total, THREADS), 8 * kNumSMs); }
}  // namespace

extern "C" int viai_loss_fwd(int kind, const float* p, const float* q, float target, int64_t n, double* acc,
                             float* loss_out, viai_stream_t stream) {
  VIAI_REQUIRE(kind >= 0 && kind <= 2 && p && acc && loss_out && n > 0 && (kind != 2 || q), "viai_loss_fwd: bad arguments");
  cudaStream_t st = STR(stream);
  VIAI_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  loss_fwd_kernel<<<grid_for(n), THREADS, 0, st>>>(kind, p, q, target, n, acc);
  VIAI_LAUNCHED();
  loss_finish_kernel<<<1, 1, 0, st>>>(acc, n, loss_out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_loss_bwd(int kind, const float* p, const float* q, float target, int64_t n, const float* gout,
                             float* dp, viai_stream_t stream) {
  VIAI_REQUIRE(kind >= 0 && kind <= 2 && p && gout && dp && n > 0 && (kind != 2 || q), "viai_loss_bwd: bad arguments");
  loss_bwd_kernel<<<grid_for(n), THREADS, 0, STR(stream)>>>(kind, p, q, target, n, gout, dp);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, double beta1,
                              double beta2, double eps, float* step_dev, int tick, float grad_scale, viai_stream_t stream) {
  VIAI_REQUIRE(p && g && m && v && lr_dev && step_dev && n > 0, "viai_adam_step: bad arguments");
  cudaStream_t st = STR(stream);
  if (tick) {
    adam_tick_kernel<<<1, 1, 0, st>>>(step_dev);
    VIAI_LAUNCHED();
  }
  adam_kernel<<<grid_for(n), THREADS, 0, st>>>(p, g, m, v, n, lr_dev, beta1, beta2, (float)eps, step_dev, grad_scale);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_fill(float* p, int64_t n, float value, viai_stream_t stream) {
  VIAI_REQUIRE(p && n > 0, "viai_fill: bad arguments");
  fill_kernel<<<grid_for(n), THREADS, 0, STR(stream)>>>(p, n, value);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_lincomb2(const float* a, float wa, const float* b, float wb, float* out, viai_stream_t stream) {
  VIAI_REQUIRE(a && out, "viai_lincomb2: bad arguments");
  lincomb2_kernel<<<1, 1, 0, STR(stream)>>>(a, wa, b, wb, out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
