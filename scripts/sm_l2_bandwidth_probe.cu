// per-SM L2 -> SM read bandwidth: G CTAs (1 per SM) each stream their own slice (L2-resident after the first pass) with LDG.128
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) rd(const float4* __restrict__ p, size_t per_cta_vec, float* out, int reps) {
  const float4* q = p + (size_t)blockIdx.x * per_cta_vec;
  float acc = 0.f;
  for (int r = 0; r < reps; ++r) {
    for (size_t i = threadIdx.x; i < per_cta_vec; i += 512 * 8) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (i + k * 512 < per_cta_vec) ? __ldcg(q + i + k * 512) : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
  }
  if (acc == 123.456f) out[blockIdx.x] = acc;
}
// same with cp.async.bulk into a 4-deep shared-memory ring of 32 KB chunks
__global__ void __launch_bounds__(128, 1) rd_bulk(const char* __restrict__ p, size_t per_cta_bytes, int reps) {
  extern __shared__ __align__(128) char sm[];
  __shared__ unsigned long long bar[4];
  const char* q = p + (size_t)blockIdx.x * per_cta_bytes;
  const unsigned CH = 32768;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[i]))); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t nch = per_cta_bytes / CH * reps;
    size_t issued = 0, done = 0;
    unsigned phase[4] = {0, 0, 0, 0};
    for (; issued < 4 && issued < nch; ++issued) {
      unsigned b = (unsigned)__cvta_generic_to_shared(&bar[issued % 4]), d = (unsigned)__cvta_generic_to_shared(sm + (issued % 4) * CH);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CH));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(q + (issued % (per_cta_bytes / CH)) * CH), "r"(CH), "r"(b) : "memory");
    }
    for (; done < nch; ++done) {
      const int s = done % 4;
      unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
      unsigned ok = 0;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(b), "r"(phase[s]));
      phase[s] ^= 1;
      if (issued < nch) {
        unsigned d = (unsigned)__cvta_generic_to_shared(sm + s * CH);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CH));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(q + (issued % (per_cta_bytes / CH)) * CH), "r"(CH), "r"(b) : "memory");
        ++issued;
      }
    }
  }
}
int main() {
  const size_t total = 99ull << 20;
  char* buf; float* out;
  cudaMalloc(&buf, total + (1 << 20)); cudaMalloc(&out, 4096);
  cudaMemset(buf, 0, total);
  cudaFuncSetAttribute(rd_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768);
  int grids[] = {8, 16, 32, 64, 148};
  for (int g : grids) {
    size_t per = (total / g) / 32768 * 32768;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rd<<<g, 512>>>((const float4*)buf, per / 16, out, 1);
    cudaEventRecord(e0); rd<<<g, 512>>>((const float4*)buf, per / 16, out, 10); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    rd_bulk<<<g, 128, 4 * 32768>>>(buf, per, 1);
    cudaEventRecord(e0); rd_bulk<<<g, 128, 4 * 32768>>>(buf, per, 10); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms2; cudaEventElapsedTime(&ms2, e0, e1);
    printf("CTAs %3d: LDG %.1f GB/s per SM (%.2f TB/s, %.1f us per 99 MB pass) | bulk %.1f GB/s per SM (%.2f TB/s, %.1f us) err=%s\n", g,
           per * 10.0 / ms / 1e6, per * 10.0 * g / ms / 1e9, ms * 100, per * 10.0 / ms2 / 1e6, per * 10.0 * g / ms2 / 1e9, ms2 * 100,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
