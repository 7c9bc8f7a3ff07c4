// Audio-visual synchronisation heads: row-wise L2 normalisation (utils/util.py:94-96), pairwise L2 distances
// (loss_functions.py:106-108) and the L2 contrastive loss built on them (loss_functions.py:111-148).
// Batches here are tens of rows of a few hundred features: latency-bound, one warp per row / pair.
#include "common.cuh"
#include <math.h>
using namespace viai;

namespace {
constexpr int THREADS = 256;
inline int warps_grid(int64_t warps) { return (int)imin64(cdiv(warps * 32, THREADS), 16 * kNumSMs); }

// y = x / max(||x||_2, eps) per row (F.normalize(p=2, dim=1))
__global__ void __launch_bounds__(THREADS)
l2norm_fwd_kernel(const float* __restrict__ x, int rows, int cols, float eps, float* __restrict__ y, float* __restrict__ nrm) {
  const int lane = threadIdx.x & 31;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += (gridDim.x * blockDim.x) >> 5) {
    const float* p = x + (int64_t)row * cols;
    float s = 0.f;
    for (int k = lane; k < cols; k += 32) s += p[k] * p[k];
    s = warp_sum(s);
    const float n = sqrtf(s), d = fmaxf(n, eps);
    for (int k = lane; k < cols; k += 32) y[(int64_t)row * cols + k] = p[k] / d;
    if (lane == 0) nrm[row] = n;
  }
}

// dx = (dy - y * <y, dy>) / n for n > eps, dy / eps below the clamp
__global__ void __launch_bounds__(THREADS)
l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ nrm, const float* __restrict__ dy, int rows, int cols,
                  float eps, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += (gridDim.x * blockDim.x) >> 5) {
    const float* py = y + (int64_t)row * cols;
    const float* pg = dy + (int64_t)row * cols;
    const float n = nrm[row];
    float dot = 0.f;
    if (n > eps) {
      for (int k = lane; k < cols; k += 32) dot += py[k] * pg[k];
      dot = warp_sum(dot);
    }
    const float d = fmaxf(n, eps);
    for (int k = lane; k < cols; k += 32) dx[(int64_t)row * cols + k] = (pg[k] - (n > eps ? py[k] * dot : 0.f)) / d;
  }
}

// scores[a][b] = || f1[a] - f2[b] ||_2
__global__ void __launch_bounds__(THREADS)
pairdist_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int n1, int n2, int F, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int total = n1 * n2;
  for (int pr = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pr < total; pr += (gridDim.x * blockDim.x) >> 5) {
    const float* pa = f1 + (int64_t)(pr / n2) * F;
    const float* pb = f2 + (int64_t)(pr % n2) * F;
    float s = 0.f;
    for (int k = lane; k < F; k += 32) { const float d = pa[k] - pb[k]; s += d * d; }
    s = warp_sum(s);
    if (lane == 0) scores[pr] = sqrtf(s);
  }
}

// df1[a][k] = sum_b ds[a][b] (f1[a][k] - f2[b][k]) / s[a][b];  df2[b][k] = -sum_a (same term).  A zero distance passes no gradient.
__global__ void __launch_bounds__(THREADS)
pairdist_bwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ scores,
                    const float* __restrict__ ds, int n1, int n2, int F, float* __restrict__ df1, float* __restrict__ df2) {
  const int total = (n1 + n2) * F;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int row = i / F, k = i - row * F;
    float acc = 0.f;
    if (row < n1) {
      const float v = f1[(int64_t)row * F + k];
      for (int b = 0; b < n2; ++b) {
        const float s = scores[row * n2 + b];
        if (s > 0.f) acc += ds[row * n2 + b] * (v - f2[(int64_t)b * F + k]) / s;
      }
      if (df1) df1[i] = acc;
    } else {
      const int b = row - n1;
      const float v = f2[(int64_t)b * F + k];
      for (int a = 0; a < n1; ++a) {
        const float s = scores[a * n2 + b];
        if (s > 0.f) acc -= ds[a * n2 + b] * (f1[(int64_t)a * F + k] - v) / s;
      }
      if (df2) df2[(int64_t)b * F + k] = acc;
    }
  }
}

// One block.  cost[a][b] = max(margin - s[a][b], 0) off the diagonal; with max_violation only the largest cost of each row
// counts (first index on ties, like torch.max on the CPU).  loss = (sum cost^2 + sum_a s[a][a]^2) / (2 B).
// dscores (optional) receives d loss / d s scaled by *gout.
__global__ void __launch_bounds__(THREADS)
contrastive_kernel(const float* __restrict__ s, int B, float margin, int max_violation, float* __restrict__ loss,
                   const float* __restrict__ gout, float* __restrict__ ds) {
  __shared__ double sh[THREADS / 32];
  const float scale = gout ? __ldg(gout) / (2.f * (float)B) : 0.f;
  double acc = 0.0;
  for (int a = threadIdx.x; a < B; a += blockDim.x) {
    const float d = s[a * B + a];
    acc += (double)d * d;
    int sel = -1;
    float best = -1.f;
    for (int b = 0; b < B; ++b) {
      const float c = (b == a) ? 0.f : fmaxf(margin - s[a * B + b], 0.f);
      if (max_violation) {
        if (c > best) { best = c; sel = b; }
      } else {
        acc += (double)c * c;
      }
    }
    if (max_violation) acc += (double)best * best;
    if (ds) {
      for (int b = 0; b < B; ++b) {
        float g;
        if (b == a) g = 2.f * d;
        else {
          const float c = fmaxf(margin - s[a * B + b], 0.f);
          g = (max_violation && b != sel) ? 0.f : -2.f * c;
        }
        ds[a * B + b] = scale * g;
      }
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    double t = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) t += sh[w];
    *loss = (float)(t / (2.0 * B));
  }
}
}  // namespace

extern "C" int viai_l2norm_fwd(const float* x, int rows, int cols, float eps, float* y, float* norms, viai_stream_t stream) {
  VIAI_REQUIRE(x && y && norms && rows > 0 && cols > 0, "l2norm_fwd: bad arguments");
  l2norm_fwd_kernel<<<warps_grid(rows), THREADS, 0, STR(stream)>>>(x, rows, cols, eps, y, norms);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_l2norm_bwd(const float* y, const float* norms, const float* dy, int rows, int cols, float eps, float* dx,
                               viai_stream_t stream) {
  VIAI_REQUIRE(y && norms && dy && dx && rows > 0 && cols > 0, "l2norm_bwd: bad arguments");
  l2norm_bwd_kernel<<<warps_grid(rows), THREADS, 0, STR(stream)>>>(y, norms, dy, rows, cols, eps, dx);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_pairdist_fwd(const float* f1, const float* f2, int n1, int n2, int F, float* scores, viai_stream_t stream) {
  VIAI_REQUIRE(f1 && f2 && scores && n1 > 0 && n2 > 0 && F > 0 && (int64_t)n1 * n2 < (1 << 26), "pairdist_fwd: bad arguments");
  pairdist_kernel<<<warps_grid((int64_t)n1 * n2), THREADS, 0, STR(stream)>>>(f1, f2, n1, n2, F, scores);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

// utils/util.py:99-121 L2retrieval, the index work: for caption i the rank of its own clip = the number of clips that are closer
// (ties: lower index first, the order a stable sort of the row would give) and the index of the closest clip.  One block per row;
// no sort of the distance matrix is needed.
__global__ void __launch_bounds__(256) retrieval_rank_kernel(const float* __restrict__ d, int n1, int n2, long long* __restrict__ rank,
                                                              long long* __restrict__ top1) {
  __shared__ int s_cnt[256];
  __shared__ float s_min[256];
  __shared__ int s_arg[256];
  const int i = blockIdx.x;
  const float* row = d + (size_t)i * n2;
  const float own = row[i];
  int cnt = 0, arg = n2;
  float mn = INFINITY;
  for (int j = threadIdx.x; j < n2; j += blockDim.x) {
    const float v = row[j];
    cnt += (v < own) || (v == own && j < i);
    if (v < mn || (v == mn && j < arg)) { mn = v; arg = j; }
  }
  s_cnt[threadIdx.x] = cnt; s_min[threadIdx.x] = mn; s_arg[threadIdx.x] = arg;
  __syncthreads();
  for (int off = blockDim.x / 2; off >= 1; off >>= 1) {
    if (threadIdx.x < off) {
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + off];
      const float v = s_min[threadIdx.x + off];
      const int a = s_arg[threadIdx.x + off];
      if (v < s_min[threadIdx.x] || (v == s_min[threadIdx.x] && a < s_arg[threadIdx.x])) { s_min[threadIdx.x] = v; s_arg[threadIdx.x] = a; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { rank[i] = s_cnt[0]; top1[i] = s_arg[0]; }
}

extern "C" int viai_retrieval_ranks(const float* dist, int n1, int n2, long long* rank, long long* top1, viai_stream_t stream) {
  VIAI_REQUIRE(dist && rank && top1 && n1 > 0 && n2 >= n1, "retrieval_ranks: bad arguments (every caption needs its own clip: n2 >= n1)");
  retrieval_rank_kernel<<<n1, 256, 0, STR(stream)>>>(dist, n1, n2, rank, top1);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_pairdist_bwd(const float* f1, const float* f2, const float* scores, const float* dscores, int n1, int n2, int F,
                                 float* df1, float* df2, viai_stream_t stream) {
  VIAI_REQUIRE(f1 && f2 && scores && dscores && (df1 || df2) && n1 > 0 && n2 > 0 && F > 0 && (int64_t)(n1 + n2) * F < (1 << 30),
               "pairdist_bwd: bad arguments");
  const int64_t total = (int64_t)(n1 + n2) * F;
  pairdist_bwd_kernel<<<(int)imin64(cdiv(total, THREADS), 16 * kNumSMs), THREADS, 0, STR(stream)>>>(f1, f2, scores, dscores, n1, n2,
                                                                                                  F, df1, df2);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_l2_contrastive(const float* scores, int B, float margin, int max_violation, float* loss, const float* gout,
                                   float* dscores, viai_stream_t stream) {
  VIAI_REQUIRE(scores && B > 0 && B <= 4096 && (loss || dscores), "l2_contrastive: bad arguments");
  VIAI_REQUIRE((gout == nullptr) == (dscores == nullptr), "l2_contrastive: gout and dscores go together");
  contrastive_kernel<<<1, THREADS, 0, STR(stream)>>>(scores, B, margin, max_violation, loss, gout, dscores);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
