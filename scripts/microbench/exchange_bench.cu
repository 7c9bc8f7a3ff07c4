// Microbenchmark of the all-to-all tagged-word exchange the WaveNet synthesis kernels are built on: 128 CTAs, each publishes 2 of
// the 256 words of a vector, every CTA needs all 256 before it can publish the next one.  Prints clocks per exchange for several
// publishing / polling variants.  Build: nvcc -arch=sm_100a -o exchange_bench exchange_bench.cu ; run on one B200.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

constexpr int NCTA = 128, NTHR = 256, NWORD = 256, NBUF = 3, MAXREP = 32;

__device__ __forceinline__ void put(unsigned long long* p, float v, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void put_red(unsigned long long* p, float v, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void put_vol(unsigned long long* p, float v, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld64v(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ unsigned long long ld64(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ void ld128(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// mode bits: 1 = replicas (16), 2 = one warp polls (8 words per lane, 128-bit loads), 4 = nanosleep in the spin,
// 8 = `work` clocks of dependent ALU work between the exchange and the next publication, 16 = poll with 64 threads x 4 words
__global__ void __launch_bounds__(NTHR, 1) exch(unsigned long long* buf, int iters, int mode, int work, long long* out, float* sink) {
  __shared__ float h[NWORD];
  const int cta = blockIdx.x, tid = threadIdx.x;
  const int nrep = (mode >> 9) ? (mode >> 9) : ((mode & 1) ? 16 : 1);
  const int rep = cta % nrep;
  float acc = 0.f;
  long long t0 = 0, spin_total = 0;
  for (int it = 0; it < iters + 10; ++it) {
    if (it == 10) t0 = clock64();
    const unsigned tag = 1u + (unsigned)it;
    unsigned long long* b = buf + (size_t)(it % NBUF) * MAXREP * NWORD;
    if (mode & 256) {                                                // lane r of warp w publishes replica r of word w
      const int wd = tid >> 5, r = tid & 31;
      if (wd < 2 && r < nrep) {
        unsigned long long* q = b + (size_t)r * NWORD + (cta * 2 + wd) % NWORD;
        if (mode & 32) put_red(q, fabsf(acc) + (float)it, tag);
        else put(q, acc + (float)it, tag);
      }
    } else if (tid < 2) {
      for (int r = 0; r < nrep; ++r) {
        unsigned long long* q = b + (size_t)r * NWORD + (cta * 2 + tid) % NWORD;
        if (mode & 32) put_red(q, fabsf(acc) + (float)it, tag);
        else if (mode & 64) put_vol(q, acc + (float)it, tag);
        else put(q, acc + (float)it, tag);
      }
    }
    const unsigned long long* rb = b + (size_t)rep * NWORD;
    const long long ts = clock64();
    if (mode & 2) {
      if (tid < 32) {
        unsigned long long w[8];
        bool ok;
        do {
          ok = true;
          for (int i = 0; i < 4; ++i) ld128(rb + tid * 8 + 2 * i, w[2 * i], w[2 * i + 1]);
          for (int i = 0; i < 8; ++i) ok = ok && ((unsigned)(w[i] >> 32) == tag);
          if (!ok && (mode & 4)) __nanosleep(20);
        } while (!ok);
        for (int i = 0; i < 8; ++i) h[tid * 8 + i] = __uint_as_float((unsigned)w[i]);
      }
    } else if (mode & 16) {
      if (tid < 64) {
        unsigned long long w[4];
        bool ok;
        do {
          ok = true;
          for (int i = 0; i < 2; ++i) ld128(rb + tid * 4 + 2 * i, w[2 * i], w[2 * i + 1]);
          for (int i = 0; i < 4; ++i) ok = ok && ((unsigned)(w[i] >> 32) == tag);
        } while (!ok);
        for (int i = 0; i < 4; ++i) h[tid * 4 + i] = __uint_as_float((unsigned)w[i]);
      }
    } else {
      const int nw = 2 * gridDim.x;                                  // words actually published
      unsigned long long w = 0;
      if (tid < nw) {
        w = (mode & 128) ? ld64v(rb + tid) : ld64(rb + tid);
        while ((unsigned)(w >> 32) != tag) {
          if (mode & 4) __nanosleep(20);
          w = (mode & 128) ? ld64v(rb + tid) : ld64(rb + tid);
        }
      }
      h[tid] = __uint_as_float((unsigned)w);
    }
    if (tid == 0) spin_total += clock64() - ts;
    __syncthreads();
    acc += h[(tid * 7) & (NWORD - 1)];
    if (mode & 8) {
      const long long tw = clock64();
      while (clock64() - tw < work) acc = acc * 1.0000001f + 1e-9f;
    }
    __syncthreads();
  }
  if (tid == 0 && cta == 0) { out[0] = clock64() - t0; out[1] = spin_total; }
  sink[cta * NTHR + tid] = acc;
}

int main() {
  unsigned long long* buf;
  long long* out;
  float* sink;
  const size_t nb = (size_t)NBUF * MAXREP * NWORD * 8;
  cudaMalloc(&buf, nb);
  cudaMalloc(&out, 16);
  cudaMalloc(&sink, NCTA * NTHR * 4);
  const int iters = 4000;
  struct V { int mode, work, ncta; };
  const V vs[] = {{256 | (16 << 9), 0, 128}, {256 | 32 | (16 << 9), 0, 128}, {256 | (8 << 9), 0, 128}, {256 | 32 | (8 << 9), 0, 128},
                  {256 | (4 << 9), 0, 128}, {256 | 32 | (4 << 9), 0, 128}, {256 | (32 << 9), 0, 128}, {256 | 32 | (32 << 9), 0, 128},
                  {256 | (2 << 9), 0, 128}, {32, 0, 128}, {256 | (16 << 9), 0, 1}, {256 | 32 | (16 << 9) | 4, 0, 128}};
  for (const V& v : vs) {
    cudaMemset(buf, 0, nb);
    int it = iters, mode = v.mode, wk = v.work;
    void* args[] = {&buf, &it, &mode, &wk, &out, &sink};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)exch, dim3(v.ncta), dim3(NTHR), args, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", v.mode, cudaGetErrorString(e)); return 1; }
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("ncta %3d mode %5d (nrep=%2d red=%d stvol=%d ldvol=%d): %7.0f clocks / exchange  (thread 0 polling: %6.0f)\n", v.ncta, v.mode,
           (v.mode >> 9) ? (v.mode >> 9) : ((v.mode & 1) ? 16 : 1), (v.mode >> 5) & 1, (v.mode >> 6) & 1, (v.mode >> 7) & 1, (double)h[0] / iters, (double)h[1] / (iters + 10));
  }
  return 0;
}
