// Fused frame + window + rFFT + |.| + mel filterbank + dB + normalise (utils/audio.py:70-75 `melspectrogram`, with the lws
// framing of :90-108).  One launch turns a waveform into the normalised mel spectrogram the GAN consumes; nothing but the
// waveform is read from and nothing but the (n_mels, M) result is written to HBM.
//
// A CTA owns FR consecutive frames.  Per frame: the 8 warps gather the (zero-padded) samples with coalesced loads, multiply
// by the analysis window and run an in-place-free Stockham radix-2 FFT between two shared-memory buffers, with the twiddle
// factors staged once per CTA in shared memory; magnitudes stay in shared memory; each warp then owns n_mels/8 filterbank
// rows and walks only the non-zero span of its triangular filters.  The FR x n_mels results are staged in shared memory so
// that every global store is a full 32-byte sector of one output row.
#include "common.cuh"
using namespace viai;

namespace {

constexpr int THREADS = 256;
constexpr int FR = 8;   // frames per CTA

__global__ void __launch_bounds__(THREADS)
stft_mel_kernel(const float* __restrict__ y, int64_t T, int pad_left, int N, int logN, int hop, const float* __restrict__ window,
                const float* __restrict__ basis, const int* __restrict__ span, int n_mels, float min_level, float ref_db,
                float min_db, int M, float* __restrict__ out, float* __restrict__ mag_out) {
  extern __shared__ float sm[];
  float* re0 = sm;                 // [N]
  float* im0 = re0 + N;
  float* re1 = im0 + N;
  float* im1 = re1 + N;
  float* twr = im1 + N;            // [N/2]
  float* twi = twr + N / 2;
  float* stage = twi + N / 2;      // [n_mels][FR]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbins = N / 2 + 1;
  for (int k = tid; k < N / 2; k += THREADS) {
    float s, c;
    sincospif(-2.0f * (float)k / (float)N, &s, &c);
    twr[k] = c;
    twi[k] = s;
  }
  const int f0 = blockIdx.x * FR;
  for (int f = 0; f < FR; ++f) {
    const int m = f0 + f;
    if (m >= M) break;                                  // uniform across the CTA
    const int64_t base = (int64_t)m * hop - pad_left;
    for (int i = tid; i < N; i += THREADS) {
      const int64_t j = base + i;
      re0[i] = (j >= 0 && j < T) ? __ldg(y + j) * __ldg(window + i) : 0.f;
      im0[i] = 0.f;
    }
    __syncthreads();
    // Stockham autosort radix-2: stage s combines sub-transforms of length L = 2^s
    float *sr = re0, *si = im0, *dr = re1, *di = im1;
    for (int s = 0; s < logN; ++s) {
      const int L = 1 << s;                             // half-size of the butterflies produced by this stage
      for (int b = tid; b < N / 2; b += THREADS) {
        const int k = b & (L - 1);                      // position inside the sub-transform
        const int j = b >> s;                           // which sub-transform pair
        const int i0 = j * L + k, i1 = i0 + N / 2;
        const int tw = k * (N / 2 / L);
        const float wr = twr[tw], wi = twi[tw];
        const float ar = sr[i0], ai = si[i0], br = sr[i1], bi = si[i1];
        const float tr = br * wr - bi * wi, ti = br * wi + bi * wr;
        const int o0 = j * 2 * L + k, o1 = o0 + L;
        dr[o0] = ar + tr; di[o0] = ai + ti;
        dr[o1] = ar - tr; di[o1] = ai - ti;
      }
      __syncthreads();
      float* t;
      t = sr; sr = dr; dr = t;
      t = si; si = di; di = t;
    }
    // magnitudes into the free buffer
    for (int k = tid; k < nbins; k += THREADS) {
      const float mg = sqrtf(sr[k] * sr[k] + si[k] * si[k]);
      dr[k] = mg;
      if (mag_out) mag_out[(int64_t)k * M + m] = mg;
    }
    __syncthreads();
    for (int r = warp; r < n_mels; r += THREADS / 32) {
      const int lo = span[2 * r], hi = span[2 * r + 1];
      float acc = 0.f;
      for (int k = lo + lane; k < hi; k += 32) acc = fmaf(__ldg(basis + (int64_t)r * nbins + k), dr[k], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        float db = 20.f * log10f(fmaxf(min_level, acc)) - ref_db;
        float v = (db - min_db) / (-min_db);
        stage[r * FR + f] = fminf(fmaxf(v, 0.f), 1.f);
      }
    }
    __syncthreads();
  }
  const int nf = min(FR, M - f0);
  for (int i = tid; i < n_mels * FR; i += THREADS) {
    const int r = i / FR, f = i - r * FR;
    if (f < nf) out[(int64_t)r * M + f0 + f] = stage[i];
  }
}

}  // namespace

extern "C" int viai_stft_mel(const float* y, int64_t T, int fft_size, int hop, int pad_left, int num_frames, const float* window,
                             const float* mel_basis, const int* mel_span, int n_mels, float min_level_db, float ref_level_db,
                             float* out, float* mag_out, viai_stream_t stream) {
  VIAI_REQUIRE(y && window && mel_basis && mel_span && out, "stft_mel: null argument");
  VIAI_REQUIRE(fft_size >= 64 && fft_size <= 4096 && (fft_size & (fft_size - 1)) == 0, "stft_mel: fft_size must be a power of two in [64, 4096]");
  VIAI_REQUIRE(hop > 0 && n_mels > 0 && num_frames >= 0 && T >= 0, "stft_mel: bad sizes");
  if (num_frames == 0) return VIAI_OK;
  int logN = 0;
  while ((1 << logN) < fft_size) ++logN;
  const size_t smem = sizeof(float) * ((size_t)5 * fft_size + (size_t)n_mels * FR);
  static bool attr = false;
  if (!attr) {
    VIAI_CUDA(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  VIAI_REQUIRE(smem <= 100 * 1024, "stft_mel: shared memory");
  const float min_level = expf(min_level_db / 20.f * logf(10.f));
  const int blocks = (num_frames + FR - 1) / FR;
  stft_mel_kernel<<<blocks, THREADS, smem, STR(stream)>>>(y, T, pad_left, fft_size, logN, hop, window, mel_basis, mel_span, n_mels,
                                                         min_level, ref_level_db, min_level_db, num_frames, out, mag_out);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
