// Tensor-core implicit-GEMM convolution for sm_100a: TMA -> shared memory -> tcgen05.mma (kind::tf32, fp32 accumulate in
// TMEM) -> tcgen05.ld epilogue with fused bias and per-channel sum / sum-of-squares (the statistics of the following
// BatchNorm2d / InstanceNorm2d).
//
// One persistent CTA per SM, 7 warps: warp 0 = activation-patch TMA producer, warp 1 = weight-tile bulk-copy producer,
// warp 2 = TMEM owner + single-thread MMA issuer, warps 3..6 = epilogue (one TMEM lane quarter each).
//
// Work decomposition ("patch-resident implicit GEMM"):
//   * an M tile is a 16 x 8 patch of output pixels of one image (128 accumulator rows);
//   * for every 32-channel slab of the input, the (16+kh-1) x (8+kw-1) input patch that all filter taps of the tile read
//     is loaded ONCE by TMA into shared memory as [8 x 16-byte channel chunk][patch row][patch col][4 floats].  In that
//     layout 8 horizontally adjacent pixels of one 16-byte chunk are 128 contiguous bytes = one UMMA core matrix, so a
//     filter tap is nothing but a different start address of the same no-swizzle K-major shared-memory descriptor
//     (SBO = patch row pitch, LBO = chunk pitch).  Zero padding and ragged edges come from TMA out-of-bounds fill;
//   * strided convolutions split the input into stride_h x stride_w parity sub-grids (one strided tensor map each), so
//     every tap is again a unit-stride window of one sub-patch;
//   * the transposed gather (ConvTranspose2d forward, Conv2d data gradient) runs one launch per output parity class, each
//     a unit-stride convolution over the tap subset that class touches, writing its outputs with the class stride;
//   * the weight operand of (tap, slab, cout tile) is one contiguous pre-packed block fetched by a single cp.async.bulk.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
using namespace viai;
using namespace viai::tc;

namespace viai {
namespace tc {
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
int encode_f32_map(CUtensorMap* map, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return -1;
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]", (int)r, rank,
              (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)d[2], (unsigned long long)(rank > 3 ? d[3] : 0),
              (unsigned long long)(rank > 4 ? d[4] : 0), b[0], b[1], b[2], rank > 3 ? b[3] : 0, rank > 4 ? b[4] : 0);
    return -1;
  }
  return 0;
}
}  // namespace tc
}  // namespace viai

namespace {

constexpr int TH = 16, TW = 8;          // output pixels per M tile (TW = 8 = rows of one UMMA core matrix)
constexpr int KC = 32;                  // input channels per slab (8 chunks of 16 bytes)
constexpr int MAX_SUB = 4, MAX_TAP = 16;
// Four warpgroups: WG0 = warps 0-2 TMA patch producer / weight producer / MMA issuer (warp 3 idle), WG1 and WG2 = two
// epilogue sets (one per TMEM accumulator set: even / odd tiles, so two tiles' epilogues are in flight), WG3 = operand split.
// Registers are rebalanced with setmaxnreg after the prologue: the producers and the split warps give theirs to the epilogue.
constexpr int NTHREADS = 512;
constexpr int XF_WARP0 = 12;
constexpr int XF2_WARP0 = 8;            // warpgroup 2: second epilogue set, or (TcParams::xf2) second operand-split group
constexpr int EPI_WARP0 = 4;

struct TcSub {
  int32_t ox, oy;        // sub-grid coordinate of the patch origin relative to the tile origin
  int32_t pw, ph;        // patch extent
  uint32_t smem_off;     // byte offset of this sub-patch inside a slab stage
};
struct TcTap {
  uint32_t a_off;        // byte offset inside the slab stage of the tap's first pixel (chunk 0)
  uint32_t sbo, lbo;     // descriptor strides in bytes (patch row pitch, chunk pitch)
  uint32_t wtap;         // index of the tap in the packed weight tensor
};
struct TcParams {
  CUtensorMap mapA[MAX_SUB];
  TcSub sub[MAX_SUB];
  TcTap tap[MAX_TAP];
  int32_t nsub, ntap;
  const float* wp;
  const float* bias;
  float* out;
  double* ssum;
  double* ssq;
  int32_t stat_groups;   // 0: no statistics, 1: per channel, N: per (image, channel)
  // red_mode 0: the fused reduction is (sum x, sum x^2) of the OUTPUT (statistics of the following norm layer).
  // red_mode 1: the output is dz, the gradient w.r.t. the post-activation tensor of the layer that produced this convolution's
  //             input; the reduction is the first pass of THAT layer's norm backward: g = dz * act'(pre), (sum g, sum g*xhat),
  //             with xhat = (ry - rmean) * rinv and pre = xhat * rgamma + rbeta read per channel (groups == 1 only).
  int32_t red_mode;
  const float* ry;       // pre-normalisation tensor of the producing layer, same shape / strides as `out`
  const float* rmean; const float* rinv; const float* rgamma; const float* rbeta;
  int32_t ract; float rslope;
  int32_t N, Hv, Wv;     // virtual output grid (per image)
  int32_t tilesX, tilesY, ntilesN, ntiles;
  int32_t nchunks, BN, Cout;
  int64_t o_sn, o_sy, o_sx, o_base;   // output element strides of (n, y, x) on the virtual grid, and base offset
  uint32_t slab_bytes, slab_tx_bytes, btile_bytes;
  int32_t SA, SB;
  uint32_t idesc;
  int32_t a4d;           // 1: sub-patch loaded by 8 rank-4 TMA copies instead of one rank-5 copy
  int32_t x3;            // 1: error-compensated 3-term tf32 product (A*Bhi + A*Blo + Alo*Bhi), ~fp32 accuracy
  int32_t b_res;         // 1: the whole packed weight tensor stays resident in shared memory for the CTA's lifetime (layers whose
                         //    weights are small: the 32/64-channel decoder layers) instead of being streamed per (tile, tap, slab)
  int32_t nB;            // weight-tile slots in shared memory (= SB when streamed, ntap * nchunks when resident)
  int32_t bx3;           // 1: the same 3-term product on bf16 pairs (x = hi + lo, both bf16): kind::f16 MMAs at twice the tf32
                         //    rate; the fp32 slab is split IN PLACE into [hi: 32 x bf16 | lo: 32 x bf16] per 128-byte pixel row
  int32_t a_sw128;       // 1: activation patch stored as dense 128-byte pixel rows under the 128-byte swizzle (rank-4 TMA)
  int32_t f16;           // with bx3: the pairs are fp16 (x * a_scale = hi + lo, 22 significand bits: ~2^-21 per product, the accuracy of
                         //    tf32x3 at the bf16 MMA rate) instead of bf16 (16 bits, ~2^-17); the weights were packed scaled by 2^10
                         //    and the epilogue multiplies the accumulator by oscale = 1 / (a_scale * 2^10)
  float a_scale, oscale;
  int32_t xf2;           // 1: warpgroup 2 splits operands instead of running the second epilogue set (thin layers)
  // Shared-memory matrix descriptors relative to a stage / weight tile, built on the host: they live in the constant bank,
  // so the MMA issuer fetches them with uniform loads and issuing one MMA costs two 64-bit adds.
  uint64_t tabA[MAX_TAP * 4 * 2];   // [tap][K step][part: 0 = hi / only, 1 = lo]
  uint64_t tabB[4 * 2];             // [K step][part]
};

// Number of split-warp threads that met a value outside the fp16 range (after scaling) in an fp16-pair convolution since the
// last viai_tc_f16_overflow(reset): such values are saturated (the result stays finite but is inaccurate).
__device__ unsigned int g_f16_overflow = 0;

__device__ __forceinline__ void transpose_reduce32(float (&v)[32], int lane) {
  // After the call v[0] of lane l holds sum over lanes of the original v[l].
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      float send = up ? v[j] : v[j + off];
      float keep = up ? v[j + off] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// MODE: 0 = one tf32 product, 1 = 3-term tf32, 2 = 3-term bf16 pairs.  BRES: weights resident in shared memory.
template <int MODE, bool BRES>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzled TMA destinations need 1 KB alignment
  uint8_t* slabA = smem;
  const uint32_t a_stage = p.slab_bytes * (MODE == 1 ? 2u : 1u);     // [raw / hi slab][lo slab]
  uint8_t* tileB = slabA + (size_t)p.SA * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tileB + (size_t)p.nB * p.btile_bytes);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + p.SA;
  uint64_t* loA = emptyA + p.SA;
  uint64_t* fullB = loA + p.SA;
  uint64_t* emptyB = fullB + p.SB;
  uint64_t* tfull = emptyB + p.SB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias_s = reinterpret_cast<float*>(bars + 66);               // bias padded to ntilesN * BN floats (zeros when absent)
  const int npad = p.ntilesN * p.BN;
  float* rmu_s = bias_s + npad;                                      // red_mode 1: per-channel mean, invstd, gamma, beta
  float* ris_s = rmu_s + npad;
  float* rga_s = ris_s + npad;
  float* rbe_s = rga_s + npad;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int acc_cols = 2 * p.BN;                // two accumulator sets (MMA of tile i+1 overlaps the epilogue of tile i)
  const uint32_t tmem_cols = (acc_cols <= 32) ? 32u : (acc_cols <= 64) ? 64u : (acc_cols <= 128) ? 128u : (acc_cols <= 256) ? 256u : 512u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); mbar_init(&loA[i], 128); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
    for (int i = 0; i < p.nsub; ++i) prefetch_tmap(&p.mapA[i]);
  }
  for (int i = threadIdx.x; i < npad; i += NTHREADS) {
    const bool in = i < p.Cout;
    bias_s[i] = (p.bias != nullptr && in) ? __ldg(&p.bias[i]) : 0.f;
    if (p.red_mode == 1) {
      rmu_s[i] = (p.rmean != nullptr && in) ? __ldg(&p.rmean[i]) : 0.f;
      ris_s[i] = (p.rinv != nullptr && in) ? __ldg(&p.rinv[i]) : 1.f;
      rga_s[i] = (p.rgamma != nullptr && in) ? __ldg(&p.rgamma[i]) : 1.f;
      rbe_s[i] = (p.rbeta != nullptr && in) ? __ldg(&p.rbeta[i]) : 0.f;
    }
  }
  if (MODE == 2 && p.f16) {
    // The split warps rewrite whole stages, including the alignment gaps between sub-patches that no TMA copy ever fills:
    // zero them once so that uninitialised shared memory cannot trip the fp16 saturation counter.
    const uint32_t n16 = (uint32_t)p.SA * a_stage / 16u;
    for (uint32_t i = threadIdx.x; i < n16; i += NTHREADS) reinterpret_cast<uint4*>(slabA)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy zeros before the async-proxy (TMA) writes
  }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < EPI_WARP0) {
   reg_dealloc<56>();
   if (warp == 0) {
    // ===== activation patch producer =====
    if (lane == 0) {
      int sa = 0;
      uint32_t pha = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int rest = tile / p.ntilesN;
        const int tx = rest % p.tilesX; rest /= p.tilesX;
        const int ty = rest % p.tilesY;
        const int n = rest / p.tilesY;
        const int x0 = tx * TW, y0 = ty * TH;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(&emptyA[sa], pha ^ 1u);
          mbar_expect_tx(&fullA[sa], p.slab_tx_bytes);
          uint8_t* dst = slabA + (size_t)sa * a_stage;
          for (int s = 0; s < p.nsub; ++s) {
            const TcSub& sb = p.sub[s];
            if (p.a_sw128) {
              tma_load_4d(dst + sb.smem_off, &p.mapA[s], &fullA[sa], c * KC, x0 + sb.ox, y0 + sb.oy, n);
            } else if (!p.a4d) {
              tma_load_5d(dst + sb.smem_off, &p.mapA[s], &fullA[sa], 0, x0 + sb.ox, y0 + sb.oy, c * (KC / 4), n);
            } else {
              const uint32_t chunk_bytes = ((uint32_t)(sb.pw * sb.ph) * 16u + 127u) & ~127u;
              for (int j = 0; j < KC / 4; ++j)
                tma_load_4d(dst + sb.smem_off + j * chunk_bytes, &p.mapA[s], &fullA[sa], c * KC + j * 4, x0 + sb.ox, y0 + sb.oy, n);
            }
          }
          if (++sa == p.SA) { sa = 0; pha ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== weight tile producer =====
    if (lane == 0 && BRES) {
      // resident weights: every (slab, tap) tile is fetched once, all completing on fullB[0]
      const size_t btile_floats = p.btile_bytes / 4;
      mbar_expect_tx(&fullB[0], (uint32_t)p.nB * p.btile_bytes);
      for (int c = 0; c < p.nchunks; ++c)
        for (int t = 0; t < p.ntap; ++t)
          bulk_load(tileB + (size_t)(c * p.ntap + t) * p.btile_bytes, p.wp + ((size_t)p.tap[t].wtap * p.nchunks + c) * btile_floats,
                    p.btile_bytes, &fullB[0]);
    } else if (lane == 0) {
      int sb = 0;
      uint32_t phb = 0;
      const size_t btile_floats = p.btile_bytes / 4;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int nt = tile % p.ntilesN;
        for (int c = 0; c < p.nchunks; ++c) {
          for (int t = 0; t < p.ntap; ++t) {
            mbar_wait(&emptyB[sb], phb ^ 1u);
            mbar_expect_tx(&fullB[sb], p.btile_bytes);
            const float* src = p.wp + (((size_t)p.tap[t].wtap * p.nchunks + c) * p.ntilesN + nt) * btile_floats;
            bulk_load(tileB + (size_t)sb * p.btile_bytes, src, p.btile_bytes, &fullB[sb]);
            if (++sb == p.SB) { sb = 0; phb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===== MMA issuer =====
    // The whole warp walks the pipeline (all lanes poll the barriers), one elected lane issues.  Everything the issue loop
    // needs is either a template constant or sits in the constant bank (descriptor tables built by the host), so that the
    // per-tap work is a handful of uniform-datapath instructions: with 32-channel layers an MMA is only ~16 tensor cycles and
    // the issue loop, not the tensor pipe, bounds the kernel.
    constexpr int KK = (MODE == 2) ? KC / 16 : KC / 8;
    const bool leader = elect_one();
    const uint32_t idesc = p.idesc;
    const int ntap = p.ntap, nchunks = p.nchunks;
    const uint32_t btile16 = p.btile_bytes >> 4;
    const uint32_t tileB16 = smem_u32(tileB) >> 4;
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    int it = 0;
    if (BRES) {
      mbar_wait(&fullB[0], 0);
      tc_fence_after();
    }
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty[acc], (((uint32_t)it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
      uint32_t accumulate = 0;
      uint32_t bres16 = tileB16;                      // resident weights: tiles are stored in (slab, tap) order
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&fullA[sa], pha);
        if (MODE != 0) mbar_wait(&loA[sa], pha);
        tc_fence_after();
        const uint64_t a16 = (uint64_t)(smem_u32(slabA + (size_t)sa * a_stage) >> 4);
        for (int t = 0; t < ntap; ++t) {
          uint64_t b16;
          if (BRES) {
            b16 = bres16;
            bres16 += btile16;
          } else {
            mbar_wait(&fullB[sb], phb);
            tc_fence_after();
            b16 = tileB16 + (uint32_t)sb * btile16;
          }
          if (leader) {
            const uint64_t* ta = p.tabA + t * (KK * 2);
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) {
              if (MODE == 0) {
                mma_tf32(d_tmem, ta[kk * 2] + a16, p.tabB[kk * 2] + b16, idesc, accumulate | (uint32_t)(kk > 0));
              } else {        // small terms first, then the leading one
                const uint64_t a_hi = ta[kk * 2] + a16, a_lo = ta[kk * 2 + 1] + a16;
                const uint64_t b_hi = p.tabB[kk * 2] + b16, b_lo = p.tabB[kk * 2 + 1] + b16;
                if (MODE == 2) {
                  mma_bf16(d_tmem, a_lo, b_hi, idesc, accumulate | (uint32_t)(kk > 0));
                  mma_bf16(d_tmem, a_hi, b_lo, idesc, 1);
                  mma_bf16(d_tmem, a_hi, b_hi, idesc, 1);
                } else {
                  mma_tf32(d_tmem, a_lo, b_hi, idesc, accumulate | (uint32_t)(kk > 0));
                  mma_tf32(d_tmem, a_hi, b_lo, idesc, 1);
                  mma_tf32(d_tmem, a_hi, b_hi, idesc, 1);
                }
              }
            }
            if (!BRES) mma_commit(&emptyB[sb]);
          }
          accumulate = 1;
          if (!BRES) {
            if (++sb == p.SB) { sb = 0; phb ^= 1u; }
          }
        }
        if (leader) mma_commit(&emptyA[sa]);
        if (++sa == p.SA) { sa = 0; pha ^= 1u; }
      }
      if (leader) mma_commit(&tfull[acc]);
    }
   }
  } else if (warp >= XF_WARP0 || (MODE == 2 && p.xf2 && warp >= XF2_WARP0)) {
    reg_dealloc<80>();
    // ===== tf32x3: split every landed slab into hi = trunc_tf32(x) (in place) and lo = x - hi (second slab) =====
    if (MODE == 2) {
      // bf16 pair split, one thread per 128-byte pixel row (private to the thread, so the rewrite is in place): logical
      // 16-byte chunk j of row r sits at physical chunk j ^ (r & 7) under the 128-byte swizzle, before and after.
      // With p.xf2 (layers of <= 64 output channels, where this stage and not the tensor pipe bounds the kernel and one epilogue
      // set is idle) warpgroup 2 is a SECOND split group: the two groups take alternate slabs.
      const int tid = threadIdx.x & 127;
      const int xg = (warp >= XF_WARP0) ? 0 : 1, nxg = p.xf2 ? 2 : 1;
      int sa = 0, slab = 0;
      uint32_t pha = 0;
      const bool f16 = p.f16 != 0;
      const float a_scale = p.a_scale;
      float amax = 0.f;
      const uint32_t nrows = p.slab_bytes / 128;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int c = 0; c < p.nchunks; ++c, ++slab) {
          if (nxg == 2 && (slab & 1) != xg) {               // the other group's slab
            if (++sa == p.SA) { sa = 0; pha ^= 1u; }
            continue;
          }
          mbar_wait(&fullA[sa], pha);
          uint8_t* base = slabA + (size_t)sa * a_stage;
          for (uint32_t r = tid; r < nrows; r += 128) {
            uint8_t* row = base + (size_t)r * 128;
            const uint32_t sw = r & 7u;
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(row + ((j ^ sw) << 4));
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const float f[8] = {v[2 * m].x, v[2 * m].y, v[2 * m].z, v[2 * m].w, v[2 * m + 1].x, v[2 * m + 1].y, v[2 * m + 1].z, v[2 * m + 1].w};
              // hi and lo are both rounded to nearest (packed cvt.rn.bf16x2): hi + lo represents x to ~2^-17 |x|.  (Truncating hi
              // saves one conversion per pair but doubles the representation error; measured speed difference: none.)
              uint32_t hi[4], lo[4];
              if (!f16) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  hi[e] = pack_bf16x2(f[2 * e], f[2 * e + 1]);
                  const float h0 = __uint_as_float(hi[e] << 16), h1 = __uint_as_float(hi[e] & 0xffff0000u);
                  lo[e] = pack_bf16x2(f[2 * e] - h0, f[2 * e + 1] - h1);
                }
              } else {
                // fp16 pairs of the scaled value: hi = the value TRUNCATED to 11 significand bits (a mask on the fp32 bits, so
                // that it converts to fp16 exactly and no fp16 -> fp32 conversion is needed to form the remainder), lo = x - hi
                // (exact) rounded to fp16: |lo| < 2^-10 |x|, representation error <= 2^-21 |x|, dropped lo*lo term < 2^-20.
                // (For 32-channel layers this stage -- not the tensor pipe -- bounds the kernel: conversions are its scarce
                // resource, profiles/r02_thin_conv_ncu.txt.)  Values beyond the fp16 range saturate (finite) and are counted.
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a = f[2 * e] * a_scale, b = f[2 * e + 1] * a_scale;
                  const float ah = __uint_as_float(__float_as_uint(a) & 0xffffe000u), bh = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
                  hi[e] = pack_f16x2_sat(ah, bh);
                  lo[e] = pack_f16x2_sat(a - ah, b - bh);
                  amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
                }
              }
              *reinterpret_cast<uint4*>(row + ((m ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(row + (((4 + m) ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&loA[sa]);
          if (++sa == p.SA) { sa = 0; pha ^= 1u; }
        }
      }
      if (amax > 65504.f) atomicAdd(&g_f16_overflow, 1u);
    } else if (MODE == 1) {
      const int tid = threadIdx.x - XF_WARP0 * 32;
      int sa = 0;
      uint32_t pha = 0;
      const uint32_t n16 = p.slab_bytes / 16;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(&fullA[sa], pha);
          float4* hi = reinterpret_cast<float4*>(slabA + (size_t)sa * a_stage);
          float4* lo = reinterpret_cast<float4*>(slabA + (size_t)sa * a_stage + p.slab_bytes);
          for (uint32_t i = tid; i < n16; i += 128) {
            float4 v = hi[i], h;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            hi[i] = h;
            lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&loA[sa]);
          if (++sa == p.SA) { sa = 0; pha ^= 1u; }
        }
      }
    }
  } else {
    reg_alloc<184>();
    // ===== epilogue: TMEM -> registers -> (+bias, statistics) -> global =====
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                // accumulator row = pixel of the tile
    const int ly = row >> 3, lx = row & 7;
    const bool do_stats = p.stat_groups > 0;
    const int nj = p.BN / 32;
    const bool narrow = nj == 1;                  // 32-channel layers: statistics stay per thread until the group ends
    float rs[8], rq[8];                           // this lane's running sum / sum of squares of channel j*32 + lane
#pragma unroll
    for (int j = 0; j < 8; ++j) rs[j] = rq[j] = 0.f;
    // narrow: this thread's (= pixel row's) running sums per channel, reduced across lanes once per group;
    // wide: acc1 is the scratch for the second reduction operand of the current chunk (acc2 unused)
    float acc1[32], acc2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc1[i] = acc2[i] = 0.f;
    const float neg_slope = (p.ract == VIAI_ACT_RELU) ? 0.f : (p.ract == VIAI_ACT_LRELU) ? p.rslope : 1.f;   // act'(pre <= 0)
    const float oscale = p.oscale;                    // fp16 pairs: undoes the operand scales (a power of two)
    int cur_group = -1, cur_nt = -1;
    auto flush = [&]() {
      if (narrow) {                               // one transposing reduction per group instead of one per tile
        transpose_reduce32(acc1, lane);
        transpose_reduce32(acc2, lane);
        rs[0] = acc1[0]; rq[0] = acc2[0];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc1[i] = acc2[i] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = cur_nt * p.BN + j * 32 + lane;
        if (j < nj && ch < p.Cout) {
          atomicAdd(&p.ssum[(size_t)cur_group * p.Cout + ch], (double)rs[j]);
          atomicAdd(&p.ssq[(size_t)cur_group * p.Cout + ch], (double)rq[j]);
        }
        rs[j] = rq[j] = 0.f;
      }
    };
    // two epilogue sets, one per accumulator set (= parity of the CTA's tile counter); with p.xf2 warpgroup 1 alone serves both
    const int nset = (MODE == 2 && p.xf2) ? 1 : 2;
    const int eset = (nset == 1) ? 0 : (warp - EPI_WARP0) >> 2;
    for (int it = eset, tile = blockIdx.x + eset * gridDim.x; tile < p.ntiles; tile += nset * gridDim.x, it += nset) {
      const int nt = tile % p.ntilesN;
      int rest = tile / p.ntilesN;
      const int tx = rest % p.tilesX; rest /= p.tilesX;
      const int ty = rest % p.tilesY;
      const int n = rest / p.tilesY;
      const int y = ty * TH + ly;
      const int acc = it & 1;
      if (do_stats) {
        const int grp = (p.stat_groups > 1) ? n : 0;
        if ((grp != cur_group || nt != cur_nt) && cur_group >= 0) flush();   // finished (group, cout tile)
        cur_group = grp;
        cur_nt = nt;
      }
      mbar_wait(&tfull[acc], ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      {
        const int x = tx * TW + lx;
        const bool valid = (y < p.Hv) && (x < p.Wv);
        float* optr = p.out + p.o_base + (int64_t)n * p.o_sn + (int64_t)y * p.o_sy + (int64_t)x * p.o_sx + (int64_t)nt * p.BN;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j >= nj) break;
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN + j * 32), v);
          if (j == nj - 1) {                        // the accumulator is in registers: hand it back to the MMA warp now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          const int ch0 = nt * p.BN + j * 32;
          if (MODE == 2 && p.f16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= oscale;
          }
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + ch0);      // broadcast reads
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = b4[i];
              v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
          }
          if (valid) {
            const int nvec = min(8, (p.Cout - ch0) >> 2);       // Cout is a multiple of 4
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i < nvec) *reinterpret_cast<float4*>(optr + j * 32 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (do_stats) {
            if (p.red_mode == 1) {
              // v = dz.  g = dz * act'(pre), xhat = (y - mu) * is, pre = xhat * ga + be  (same expressions as norm_act.cu
              // bwd_terms, so that this pass and viai_norm_act_bwd_apply agree on every sign decision)
              const float4* y4 = reinterpret_cast<const float4*>(p.ry + (optr - p.out) + j * 32);
              const float4* mu4 = reinterpret_cast<const float4*>(rmu_s + ch0);
              const float4* is4 = reinterpret_cast<const float4*>(ris_s + ch0);
              const float4* ga4 = reinterpret_cast<const float4*>(rga_s + ch0);
              const float4* be4 = reinterpret_cast<const float4*>(rbe_s + ch0);
              const int nvec = min(8, (p.Cout - ch0) >> 2);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const bool on = valid && i < nvec;
                float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (on) yv = __ldg(y4 + i);
                const float4 mu = mu4[i], is = is4[i], ga = ga4[i], be = be4[i];
                float xh, pre, gg;
#define VIAI_RED1(E, YV, MU, IS, GA, BE)                                   \
                xh = (YV - MU) * IS;                                       \
                pre = xh * GA + BE;                                        \
                gg = on ? v[4 * i + E] * (pre > 0.f ? 1.f : neg_slope) : 0.f; \
                if (narrow) { acc1[4 * i + E] += gg; acc2[4 * i + E] += gg * xh; } \
                else { v[4 * i + E] = gg; acc1[4 * i + E] = gg * xh; }
                VIAI_RED1(0, yv.x, mu.x, is.x, ga.x, be.x)
                VIAI_RED1(1, yv.y, mu.y, is.y, ga.y, be.y)
                VIAI_RED1(2, yv.z, mu.z, is.z, ga.z, be.z)
                VIAI_RED1(3, yv.w, mu.w, is.w, ga.w, be.w)
#undef VIAI_RED1
              }
            } else if (narrow) {
              if (valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  acc1[i] += v[i];
                  acc2[i] = fmaf(v[i], v[i], acc2[i]);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                v[i] = valid ? v[i] : 0.f;
                acc1[i] = v[i] * v[i];
              }
            }
            if (!narrow) {
              transpose_reduce32(v, lane);
              transpose_reduce32(acc1, lane);
              rs[j] += v[0];
              rq[j] += acc1[0];
            }
          }
        }
      }
    }
    if (do_stats && cur_group >= 0) flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dst[tap][slab][cout tile][part][16-byte chunk j][row][4]; part 0 = tf32(w[o][tap][i]), part 1 (split packing only) =
// tf32(w - part 0);  i = slab*32 + j*4 + e, o = tile*BN + row
// bf16-pair packing: dst (as 16-bit words) [tap][slab][cout tile][part: hi, lo][8-channel chunk j (4)][row][8]
__global__ void pack_weight_bf16x2_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int O, int I, int R, int S,
                                          int64_t so, int64_t si, int64_t sr, int64_t ss, int flip, int BN, int nchunks, int ntilesN) {
  const int64_t total = (int64_t)R * S * nchunks * ntilesN * 2 * (KC / 8) * BN * 8;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = idx;
    const int e = t & 7; t >>= 3;
    const int row = t % BN; t /= BN;
    const int j = t % (KC / 8); t /= (KC / 8);
    const int part = t & 1; t >>= 1;
    const int nt = t % ntilesN; t /= ntilesN;
    const int c = t % nchunks; t /= nchunks;
    const int tap = (int)t;
    const int r = tap / S, s = tap % S;
    const int o = nt * BN + row, i = c * KC + j * 8 + e;
    uint16_t v = 0;
    if (o < O && i < I) {
      const int rr = flip ? R - 1 - r : r, sw = flip ? S - 1 - s : s;
      const float w = src[o * so + i * si + rr * sr + sw * ss];
      const __nv_bfloat16 hi = __float2bfloat16_rn(w);
      v = __bfloat16_as_ushort(part == 0 ? hi : __float2bfloat16_rn(w - __bfloat162float(hi)));
    }
    dst[idx] = v;
  }
}

// fp16-pair packing: the same layout as the bf16 pairs.  Operands are scaled by FIXED powers of two before the split so that
// hi keeps 11 significand bits and lo = (x - hi) the next 11 as a NORMAL fp16 number for every value that matters:
//   weights     w * 2^10:  |w| < 64 representable; full 22 bits for |w| >= 2^-13 (1.2e-4); below that lo is an fp16 subnormal
//                          and the absolute error is <= 2^-25 / 2^10 = 2.9e-11 -- against typical weights of 1e-2 .. 1e-1 that
//                          is still ~2^-28 relative to the tensor;
//   activations x * 2^3:   |x| < 8188 representable; full 22 bits for |x| >= 2^-6; absolute error below that <= 3.7e-9.
// The epilogue multiplies the accumulator by 2^-13 (exact).  Values beyond the range saturate (finite) and are counted in
// g_f16_overflow; the host side turns a non-zero count into an error where it already synchronises.
constexpr float kF16ActScale = 8.0f;
constexpr float kF16WeightScale = 1024.0f;
__global__ void pack_weight_f16x2_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int O, int I, int R, int S,
                                         int64_t so, int64_t si, int64_t sr, int64_t ss, int flip, int BN, int nchunks, int ntilesN) {
  const int64_t total = (int64_t)R * S * nchunks * ntilesN * 2 * (KC / 8) * BN * 8;
  bool ovf = false;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = idx;
    const int e = t & 7; t >>= 3;
    const int row = t % BN; t /= BN;
    const int j = t % (KC / 8); t /= (KC / 8);
    const int part = t & 1; t >>= 1;
    const int nt = t % ntilesN; t /= ntilesN;
    const int c = t % nchunks; t /= nchunks;
    const int tap = (int)t;
    const int r = tap / S, s = tap % S;
    const int o = nt * BN + row, i = c * KC + j * 8 + e;
    uint16_t v = 0;
    if (o < O && i < I) {
      const int rr = flip ? R - 1 - r : r, sw = flip ? S - 1 - s : s;
      const float w = src[o * so + i * si + rr * sr + sw * ss] * kF16WeightScale;
      const uint32_t hi = pack_f16x2_sat(w, 0.f);
      v = (uint16_t)(part == 0 ? (hi & 0xffffu) : (pack_f16x2_sat(w - f16lo_to_f32(hi), 0.f) & 0xffffu));
      ovf |= fabsf(w) > 65504.f;
    }
    dst[idx] = v;
  }
  if (ovf) atomicAdd(&g_f16_overflow, 1u);
}

__global__ void pack_weight_tc_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I, int R, int S,
                                      int64_t so, int64_t si, int64_t sr, int64_t ss, int flip, int BN, int nchunks, int ntilesN,
                                      int parts) {
  const int64_t total = (int64_t)R * S * nchunks * ntilesN * parts * (KC / 4) * BN * 4;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = idx;
    const int e = t & 3; t >>= 2;
    const int row = t % BN; t /= BN;
    const int j = t % (KC / 4); t /= (KC / 4);
    const int part = t % parts; t /= parts;
    const int nt = t % ntilesN; t /= ntilesN;
    const int c = t % nchunks; t /= nchunks;
    const int tap = (int)t;
    const int r = tap / S, s = tap % S;
    const int o = nt * BN + row, i = c * KC + j * 4 + e;
    float v = 0.f;
    if (o < O && i < I) {
      const int rr = flip ? R - 1 - r : r, sw = flip ? S - 1 - s : s;
      const float w = src[o * so + i * si + rr * sr + sw * ss];
      const float hi = to_tf32(w);
      v = part == 0 ? hi : to_tf32(w - hi);
    }
    dst[idx] = v;
  }
}

// ---- batched weight packing: every operand re-layout a training-step segment needs in ONE launch ------------------------------
// (a C2 step re-lays ~95 weight tensors between two optimizer updates: as separate ~3 us launches that is ~2 % of the step and a
// fifth of its launches).  blockIdx.y selects the tensor, blockIdx.x strides over its packed elements; descriptors travel by value.
constexpr int kMaxBatchPack = 24;
struct PackBatch {
  viai_pack_desc d[kMaxBatchPack];
  int32_t bn[kMaxBatchPack];
  int32_t n;
};

__device__ __forceinline__ float pack_src(const viai_pack_desc& q, int o, int i, int r, int s) {
  const int rr = q.flip ? q.R - 1 - r : r, sw = q.flip ? q.S - 1 - s : s;
  return q.src[o * q.so + i * q.si + rr * q.sr + sw * q.ss];
}

__global__ void __launch_bounds__(256) pack_batched_kernel(const __grid_constant__ PackBatch pb) {
  const viai_pack_desc& q = pb.d[blockIdx.y];
  const int O = q.O, I = q.I, R = q.R, S = q.S;
  if (q.kind == 0) {                                   // thin / CUDA-core operand: fp32 [O][R][S][I]
    const int64_t total = (int64_t)O * R * S * I;
    float* dst = reinterpret_cast<float*>(q.dst);
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int i = idx % I;
      int64_t t = idx / I;
      const int s = t % S; t /= S;
      const int r = t % R;
      const int o = (int)(t / R);
      dst[idx] = pack_src(q, o, i, r, s);
    }
    return;
  }
  const int BN = pb.bn[blockIdx.y], nchunks = (I + KC - 1) / KC, ntilesN = (O + BN - 1) / BN;
  if (q.kind == 1 || q.kind == 2) {                    // tf32 / tf32 pairs: the layout of pack_weight_tc_kernel
    const int parts = q.kind == 2 ? 2 : 1;
    const int64_t total = (int64_t)R * S * nchunks * ntilesN * parts * (KC / 4) * BN * 4;
    float* dst = reinterpret_cast<float*>(q.dst);
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      int64_t t = idx;
      const int e = t & 3; t >>= 2;
      const int row = t % BN; t /= BN;
      const int j = t % (KC / 4); t /= (KC / 4);
      const int part = t % parts; t /= parts;
      const int nt = t % ntilesN; t /= ntilesN;
      const int c = t % nchunks; t /= nchunks;
      const int tap = (int)t;
      const int o = nt * BN + row, i = c * KC + j * 4 + e;
      float v = 0.f;
      if (o < O && i < I) {
        const float w = pack_src(q, o, i, tap / S, tap % S);
        const float hi = to_tf32(w);
        v = part == 0 ? hi : to_tf32(w - hi);
      }
      dst[idx] = v;
    }
    return;
  }
  // 16-bit pairs (3: bf16, 4: fp16 of the scaled weight): the layout of pack_weight_bf16x2_kernel / pack_weight_f16x2_kernel
  const int64_t total = (int64_t)R * S * nchunks * ntilesN * 2 * (KC / 8) * BN * 8;
  uint16_t* dst = reinterpret_cast<uint16_t*>(q.dst);
  bool ovf = false;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = idx;
    const int e = t & 7; t >>= 3;
    const int row = t % BN; t /= BN;
    const int j = t % (KC / 8); t /= (KC / 8);
    const int part = t & 1; t >>= 1;
    const int nt = t % ntilesN; t /= ntilesN;
    const int c = t % nchunks; t /= nchunks;
    const int tap = (int)t;
    const int o = nt * BN + row, i = c * KC + j * 8 + e;
    uint16_t v = 0;
    if (o < O && i < I) {
      const float w0 = pack_src(q, o, i, tap / S, tap % S);
      if (q.kind == 3) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(w0);
        v = __bfloat16_as_ushort(part == 0 ? hi : __float2bfloat16_rn(w0 - __bfloat162float(hi)));
      } else {
        const float w = w0 * kF16WeightScale;
        const uint32_t hi = pack_f16x2_sat(w, 0.f);
        v = (uint16_t)(part == 0 ? (hi & 0xffffu) : (pack_f16x2_sat(w - f16lo_to_f32(hi), 0.f) & 0xffffu));
        ovf |= fabsf(w) > 65504.f;
      }
    }
    dst[idx] = v;
  }
  if (ovf) atomicAdd(&g_f16_overflow, 1u);
}

// Cout tile.  Capped at 128 so that two M tiles x two accumulator sets fit the 512 TMEM columns: the weight traffic of a layer
// does not depend on this width (it is K * Cout * pixels / M), only on the number of pixels that share a weight tile.
inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
inline int tc_bn(int Cout) {
  static const int cap = env_int("VIAI_TC_BNCAP", 256);      // tuning knob: N = 128 MMAs run at ~2/3 of the N = 256 rate (measured)
  int bn = ((Cout + 31) / 32) * 32;
  return bn > cap ? cap : bn;
}
inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

struct TapSpec { int suby, subx, offy, offx, wtap; };

// host copy of make_desc / make_desc_sw128 (layout: 0 = no swizzle, 2 = 128-byte swizzle; see tc_common.cuh)
inline uint64_t host_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// Builds and launches one "virtual unit-stride convolution" (see the file header).
int launch_plan(const viai_conv_geom& g, const float* in, int inH, int inW, int inC, int sub_sy, int sub_sx,
                const TapSpec* taps, int ntap, const float* wp, const float* bias, float* out, int Hv, int Wv, int64_t o_sn,
                int64_t o_sy, int64_t o_sx, int64_t o_base, int Cout, double* ssum, double* ssq, int stat_groups, int flags,
                const viai_norm_bwd_ctx* nb, cudaStream_t stream) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  if (nb != nullptr) {
    p.red_mode = 1;
    p.ry = nb->y; p.rmean = nb->mean; p.rinv = nb->invstd; p.rgamma = nb->gamma; p.rbeta = nb->beta;
    p.ract = nb->act; p.rslope = nb->slope;
  }
  VIAI_REQUIRE(ntap >= 1 && ntap <= MAX_TAP, "conv2d_tc: %d taps (max %d)", ntap, MAX_TAP);
  // default: dense 128-byte pixel rows under the 128-byte swizzle.  flags & 1: 16-byte channel chunks, no swizzle (rank-5
  // TMA); flags & 3 == 3: the same through eight rank-4 copies.  Both alternatives are kept as cross-checks of the layout.
  p.x3 = (flags & 4) ? 1 : 0;
  p.bx3 = (flags & 8) ? 1 : 0;
  p.f16 = (flags & 32) ? 1 : 0;
  VIAI_REQUIRE(!p.f16 || p.bx3, "conv2d_tc: the fp16-pair flag (32) goes with the 16-bit pair product (8)");
  if (p.f16) {
    p.a_scale = kF16ActScale;
    p.oscale = 1.0f / (kF16ActScale * kF16WeightScale);
  }
  p.a_sw128 = (flags & 1) ? 0 : 1;
  VIAI_REQUIRE(!(p.x3 && p.bx3), "conv2d_tc: VIAI_TC_X3 and VIAI_TC_BF16X3 are exclusive");
  VIAI_REQUIRE(!(p.x3 | p.bx3) || p.a_sw128, "conv2d_tc: the 3-term products need the swizzled activation layout");
  p.a4d = (!p.a_sw128 && (flags & 2)) ? 1 : 0;
  // sub-patches: one per (suby, subx) parity that occurs
  int sub_id[2][2] = {{-1, -1}, {-1, -1}};
  int mn_y[MAX_SUB], mx_y[MAX_SUB], mn_x[MAX_SUB], mx_x[MAX_SUB], sy_of[MAX_SUB], sx_of[MAX_SUB];
  int nsub = 0;
  for (int t = 0; t < ntap; ++t) {
    int& id = sub_id[taps[t].suby][taps[t].subx];
    if (id < 0) {
      id = nsub++;
      sy_of[id] = taps[t].suby; sx_of[id] = taps[t].subx;
      mn_y[id] = mx_y[id] = taps[t].offy;
      mn_x[id] = mx_x[id] = taps[t].offx;
    } else {
      mn_y[id] = taps[t].offy < mn_y[id] ? taps[t].offy : mn_y[id];
      mx_y[id] = taps[t].offy > mx_y[id] ? taps[t].offy : mx_y[id];
      mn_x[id] = taps[t].offx < mn_x[id] ? taps[t].offx : mn_x[id];
      mx_x[id] = taps[t].offx > mx_x[id] ? taps[t].offx : mx_x[id];
    }
  }
  p.nsub = nsub;
  uint32_t off = 0;
  for (int s = 0; s < nsub; ++s) {
    TcSub& sb = p.sub[s];
    sb.ox = mn_x[s]; sb.oy = mn_y[s];
    sb.pw = TW + (mx_x[s] - mn_x[s]);
    sb.ph = TH + (mx_y[s] - mn_y[s]);
    sb.smem_off = off;
    if (p.a_sw128) {
      off = (off + 1023u) & ~1023u;
      sb.smem_off = off;
      off += (uint32_t)(sb.pw * sb.ph) * (KC * 4);
    }
    const uint32_t chunk_pitch = p.a_sw128 ? 0u : p.a4d ? (((uint32_t)(sb.pw * sb.ph) * 16u + 127u) & ~127u) : (uint32_t)(sb.pw * sb.ph) * 16u;
    off += chunk_pitch * (KC / 4);
    VIAI_REQUIRE(sb.pw <= 256 && sb.ph <= 256, "conv2d_tc: patch %dx%d too large", sb.ph, sb.pw);
    const int subH = (inH - sy_of[s] + sub_sy - 1) / sub_sy, subW = (inW - sx_of[s] + sub_sx - 1) / sub_sx;
    VIAI_REQUIRE(subH >= 1 && subW >= 1, "conv2d_tc: empty parity sub-grid");
    const float* base = in + ((int64_t)sy_of[s] * inW + sx_of[s]) * inC;
    if (p.a_sw128) {
      uint64_t dims[4] = {(uint64_t)inC, (uint64_t)subW, (uint64_t)subH, (uint64_t)g.N};
      uint64_t strides[3] = {(uint64_t)sub_sx * inC * 4, (uint64_t)sub_sy * inW * inC * 4, (uint64_t)inH * inW * inC * 4};
      uint32_t box[4] = {KC, (uint32_t)sb.pw, (uint32_t)sb.ph, 1};
      if (encode_f32_map(&p.mapA[s], 4, base, dims, strides, box, 1)) return VIAI_ERR_CUDA;
    } else if (!p.a4d) {
      uint64_t dims[5] = {4, (uint64_t)subW, (uint64_t)subH, (uint64_t)(inC / 4), (uint64_t)g.N};
      uint64_t strides[4] = {(uint64_t)sub_sx * inC * 4, (uint64_t)sub_sy * inW * inC * 4, 16, (uint64_t)inH * inW * inC * 4};
      uint32_t box[5] = {4, (uint32_t)sb.pw, (uint32_t)sb.ph, KC / 4, 1};
      if (encode_f32_map(&p.mapA[s], 5, base, dims, strides, box)) return VIAI_ERR_CUDA;
    } else {
      uint64_t dims[4] = {(uint64_t)inC, (uint64_t)subW, (uint64_t)subH, (uint64_t)g.N};
      uint64_t strides[3] = {(uint64_t)sub_sx * inC * 4, (uint64_t)sub_sy * inW * inC * 4, (uint64_t)inH * inW * inC * 4};
      uint32_t box[4] = {4, (uint32_t)sb.pw, (uint32_t)sb.ph, 1};
      if (encode_f32_map(&p.mapA[s], 4, base, dims, strides, box)) return VIAI_ERR_CUDA;
    }
  }
  if (p.a_sw128) off = (off + 1023u) & ~1023u;
  p.slab_bytes = off;
  for (int s = 0; s < nsub; ++s) p.slab_tx_bytes += (uint32_t)(p.sub[s].pw * p.sub[s].ph) * (KC * 4);
  p.ntap = ntap;
  for (int t = 0; t < ntap; ++t) {
    const int s = sub_id[taps[t].suby][taps[t].subx];
    const TcSub& sb = p.sub[s];
    TcTap& tp = p.tap[t];
    const uint32_t pix_pitch = p.a_sw128 ? 128u : 16u;
    tp.a_off = sb.smem_off + (uint32_t)((taps[t].offy - sb.oy) * sb.pw + (taps[t].offx - sb.ox)) * pix_pitch;
    tp.sbo = (uint32_t)sb.pw * pix_pitch;
    tp.lbo = p.a4d ? (((uint32_t)(sb.pw * sb.ph) * 16u + 127u) & ~127u) : (uint32_t)(sb.pw * sb.ph) * 16u;
    tp.wtap = (uint32_t)taps[t].wtap;
  }
  p.wp = wp; p.bias = bias; p.out = out; p.ssum = ssum; p.ssq = ssq;
  p.stat_groups = (ssum != nullptr) ? stat_groups : 0;
  p.N = g.N; p.Hv = Hv; p.Wv = Wv;
  p.tilesX = (Wv + TW - 1) / TW; p.tilesY = (Hv + TH - 1) / TH;
  p.BN = tc_bn(Cout);
  p.ntilesN = (Cout + p.BN - 1) / p.BN;
  p.ntiles = p.N * p.tilesX * p.tilesY * p.ntilesN;
  p.nchunks = (inC + KC - 1) / KC;
  p.Cout = Cout;
  p.o_sn = o_sn; p.o_sy = o_sy; p.o_sx = o_sx; p.o_base = o_base;
  p.btile_bytes = (uint32_t)p.BN * KC * 4 * (p.x3 ? 2u : 1u);      // bf16 pairs: 2 x 2 bytes per element = the fp32 size
  // Second split group on thin layers: measured on B200 (round 2) it changes nothing (15.49 vs 15.34 ms per C2 step) -- the
  // 32-channel layers are bound by the tensor core's OPERAND FETCH from shared memory (every M=128, K=16 MMA re-reads a 4 KB A
  // tile whatever N is: 54 MMAs x 5 KB = 270 KB per 128-pixel tile, ~2 100 cycles at 128 B/clk), not by the split stage.  Kept
  // as an option (VIAI_TC_XF2=1), off by default.
  static const int xf2_env = env_int("VIAI_TC_XF2", 0);
  p.xf2 = (p.bx3 && xf2_env && p.BN <= 64) ? 1 : 0;
  p.idesc = p.f16 ? make_idesc_f16(128, p.BN, 0, 0) : p.bx3 ? make_idesc_bf16(128, p.BN, 0, 0) : make_idesc_tf32(128, p.BN, 0, 0);
  // pipeline depths under the 227 KB shared-memory limit
  const size_t fixed = 1024 /*alignment slack*/ + 66 * 8 + 5 * (size_t)(((Cout + 31) / 32) * 32 + 256) * 4;   // barriers + tmem slot + bias + 4 norm constants
  const size_t budget = 227 * 1024;
  // Activation stages: the thin layers are HBM-bound streams whose only memory-level parallelism is the TMA loads in flight
  // (a 24 KB slab per stage and CTA): with 3 stages 148 CTAs keep < 10 MB in flight, half of what ~6.5 TB/s x ~2 us needs.
  static const int sa_max = env_int("VIAI_TC_SA", 4);
  int SA = sa_max, SB = 4;
  VIAI_REQUIRE(3 * SA + 2 * SB + 4 <= 64, "conv2d_tc: barrier area");
  auto need = [&](int a, int b) { return fixed + (size_t)a * p.slab_bytes * (p.x3 ? 2 : 1) + (size_t)b * p.btile_bytes; };
  // resident weights when the whole packed tensor (one cout tile) fits beside >= 2 activation stages
  const int nB_all = ntap * p.nchunks;
  if (p.ntilesN == 1 && !(flags & 16) && (size_t)nB_all * p.btile_bytes <= 160 * 1024 && need(2, nB_all) <= budget &&
      (size_t)nB_all * p.btile_bytes < (1u << 20)) {
    p.b_res = 1;
    while (need(SA, nB_all) > budget) --SA;
    p.SA = SA; p.SB = 1; p.nB = nB_all;
  } else {
    while (need(SA, SB) > budget && SA > 3) --SA;
    while (need(SA, SB) > budget && SB > 2) --SB;
    while (need(SA, SB) > budget && SA > 2) --SA;
    while (need(SA, SB) > budget && SB > 1) --SB;
    while (need(SA, SB) > budget && SA > 1) --SA;
    VIAI_REQUIRE(need(SA, SB) <= budget, "conv2d_tc: tile does not fit in shared memory (slab %u B, weight tile %u B)", p.slab_bytes,
                 p.btile_bytes);
    p.SA = SA; p.SB = SB; p.nB = SB;
  }
  size_t smem = need(p.SA, p.nB);
  if (smem < 120 * 1024) smem = 120 * 1024;   // force one CTA per SM (each CTA may allocate up to all 512 TMEM columns)
  // descriptor tables (addresses relative to the stage / weight tile; the kernel adds the base >> 4 to the low word)
  {
    const int KK = p.bx3 ? KC / 16 : KC / 8;
    const uint32_t b_lbo = (uint32_t)p.BN * 16u, b_sbo = 128u;
    for (int t = 0; t < ntap; ++t)
      for (int kk = 0; kk < KK; ++kk)
        for (int part = 0; part < 2; ++part) {
          const TcTap& tp = p.tap[t];
          uint64_t d;
          if (p.a_sw128) d = host_desc(tp.a_off + (uint32_t)kk * 32u + (part ? (p.bx3 ? 64u : p.slab_bytes) : 0u), 16u, tp.sbo, 2);
          else d = host_desc(tp.a_off + (uint32_t)kk * 2u * tp.lbo, tp.lbo, tp.sbo, 0);
          p.tabA[(t * KK + kk) * 2 + part] = d;
        }
    for (int kk = 0; kk < KK; ++kk)
      for (int part = 0; part < 2; ++part)
        p.tabB[kk * 2 + part] = host_desc((uint32_t)kk * 2u * b_lbo + (part ? (uint32_t)p.BN * (p.bx3 ? 64u : 128u) : 0u), b_lbo, b_sbo, 0);
  }
  if (p.ntiles == 0) return VIAI_OK;
  const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
  const int mode = p.bx3 ? 2 : p.x3 ? 1 : 0;
#define VIAI_TC_LAUNCH(M, R)                                                                                            \
  do {                                                                                                                  \
    static bool attr_done = false;                                                                                      \
    if (!attr_done) {                                                                                                   \
      VIAI_CUDA(cudaFuncSetAttribute(conv_tc_kernel<M, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));    \
      attr_done = true;                                                                                                 \
    }                                                                                                                   \
    conv_tc_kernel<M, R><<<grid, NTHREADS, smem, stream>>>(p);                                                          \
  } while (0)
  if (p.b_res) { if (mode == 2) VIAI_TC_LAUNCH(2, true); else if (mode == 1) VIAI_TC_LAUNCH(1, true); else VIAI_TC_LAUNCH(0, true); }
  else { if (mode == 2) VIAI_TC_LAUNCH(2, false); else if (mode == 1) VIAI_TC_LAUNCH(1, false); else VIAI_TC_LAUNCH(0, false); }
#undef VIAI_TC_LAUNCH
  VIAI_LAUNCHED();
  return VIAI_OK;
}

}  // namespace

extern "C" int viai_tc_bn(int Cout) { return tc_bn(Cout); }

namespace {
int64_t packed_elems(int O, int I, int R, int S, int split) {
  const int BN = tc_bn(O);
  return (int64_t)R * S * ((I + KC - 1) / KC) * ((O + BN - 1) / BN) * BN * KC * (split == 1 ? 2 : 1);
}
}  // namespace

// split: 0 tf32, 1 tf32 pairs, 2 bf16 pairs, 3 fp16 pairs (same size as 2)
extern "C" int64_t viai_tc_packed_size(int O, int I, int R, int S, int split) { return packed_elems(O, I, R, S, split); }

// Threads that saturated a value in an fp16-pair convolution since the last reset (synchronises with the device).
extern "C" int viai_tc_f16_overflow(int reset, unsigned int* count) {
  VIAI_REQUIRE(count != nullptr, "tc_f16_overflow: null argument");
  VIAI_CUDA(cudaMemcpyFromSymbol(count, g_f16_overflow, sizeof(unsigned int)));
  if (reset && *count) {
    const unsigned int zero = 0;
    VIAI_CUDA(cudaMemcpyToSymbol(g_f16_overflow, &zero, sizeof(unsigned int)));
  }
  return VIAI_OK;
}

extern "C" int viai_pack_weight_tc(const float* src, float* dst, int O, int I, int R, int S, int64_t so, int64_t si, int64_t sr,
                                   int64_t ss, int flip, int split, viai_stream_t stream) {
  VIAI_REQUIRE(src && dst && O > 0 && I > 0 && R > 0 && S > 0, "pack_weight_tc: bad arguments");
  const int BN = tc_bn(O), nchunks = (I + KC - 1) / KC, ntilesN = (O + BN - 1) / BN;
  const int64_t total = packed_elems(O, I, R, S, split);
  if (split == 3) {
    const int blocks = (int)imin64(cdiv(total * 2, 256), 4096);
    pack_weight_f16x2_kernel<<<blocks, 256, 0, STR(stream)>>>(src, reinterpret_cast<uint16_t*>(dst), O, I, R, S, so, si, sr, ss, flip,
                                                             BN, nchunks, ntilesN);
    VIAI_LAUNCHED();
    return VIAI_OK;
  }
  if (split == 2) {
    const int blocks = (int)imin64(cdiv(total * 2, 256), 4096);
    pack_weight_bf16x2_kernel<<<blocks, 256, 0, STR(stream)>>>(src, reinterpret_cast<uint16_t*>(dst), O, I, R, S, so, si, sr, ss, flip,
                                                              BN, nchunks, ntilesN);
    VIAI_LAUNCHED();
    return VIAI_OK;
  }
  const int blocks = (int)imin64(cdiv(total, 256), 4096);
  pack_weight_tc_kernel<<<blocks, 256, 0, STR(stream)>>>(src, dst, O, I, R, S, so, si, sr, ss, flip, BN, nchunks, ntilesN, split ? 2 : 1);
  VIAI_LAUNCHED();
  return VIAI_OK;
}

extern "C" int viai_pack_weights_batched(const viai_pack_desc* descs, int n, viai_stream_t stream) {
  VIAI_REQUIRE(descs != nullptr && n > 0, "pack_weights_batched: bad arguments");
  for (int base = 0; base < n; base += kMaxBatchPack) {
    PackBatch pb;
    memset(&pb, 0, sizeof(pb));
    pb.n = n - base < kMaxBatchPack ? n - base : kMaxBatchPack;
    for (int k = 0; k < pb.n; ++k) {
      const viai_pack_desc& q = descs[base + k];
      VIAI_REQUIRE(q.src && q.dst && q.O > 0 && q.I > 0 && q.R > 0 && q.S > 0 && q.kind >= 0 && q.kind <= 4,
                   "pack_weights_batched: bad descriptor %d", base + k);
      pb.d[k] = q;
      pb.bn[k] = tc_bn(q.O);
    }
    pack_batched_kernel<<<dim3(2 * kNumSMs, pb.n), 256, 0, STR(stream)>>>(pb);   // the largest tensor sets the duration: give it the whole GPU
    VIAI_LAUNCHED();
  }
  return VIAI_OK;
}

extern "C" int viai_conv2d_tc_supported(const viai_conv_geom* g) {
  if (!g) return 0;
  if (g->Cin % 4 != 0 || g->Cin < 16 || g->Cout % 4 != 0 || g->Cout < 16) return 0;
  if (g->R * g->S > MAX_TAP || g->stride_h < 1 || g->stride_h > 2 || g->stride_w < 1 || g->stride_w > 2) return 0;
  if (g->mode != 0 && g->mode != 1) return 0;
  return 1;
}

static int conv2d_tc_impl(const viai_conv_geom* gp, const float* in, const float* wp_tc, const float* bias, float* out,
                          double* stat_sum, double* stat_sumsq, int stat_groups, int flags, const viai_norm_bwd_ctx* nb,
                          viai_stream_t stream) {
  VIAI_REQUIRE(gp && in && wp_tc && out, "conv2d_tc: null argument");
  const viai_conv_geom& g = *gp;
  VIAI_REQUIRE(viai_conv2d_tc_supported(gp), "conv2d_tc: unsupported geometry (Cin %d Cout %d %dx%d stride %d,%d mode %d)", g.Cin,
               g.Cout, g.R, g.S, g.stride_h, g.stride_w, g.mode);
  VIAI_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(wp_tc) & 15) == 0,
               "conv2d_tc: pointers must be 16-byte aligned");
  VIAI_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), "conv2d_tc: stat_sum and stat_sumsq go together");
  cudaStream_t st = STR(stream);
  if (stat_sum) {
    const size_t nbytes = (size_t)(stat_groups > 1 ? stat_groups : 1) * g.Cout * sizeof(double);
    if (reinterpret_cast<char*>(stat_sumsq) == reinterpret_cast<char*>(stat_sum) + nbytes) {   // one (2, n) buffer: one memset node
      VIAI_CUDA(cudaMemsetAsync(stat_sum, 0, 2 * nbytes, st));
    } else {
      VIAI_CUDA(cudaMemsetAsync(stat_sum, 0, nbytes, st));
      VIAI_CUDA(cudaMemsetAsync(stat_sumsq, 0, nbytes, st));
    }
  }
  TapSpec taps[MAX_TAP];
  const int64_t o_row = (int64_t)g.Wout * g.Cout, o_img = (int64_t)g.Hout * o_row;
  if (g.mode == 0) {
    int nt = 0;
    for (int r = 0; r < g.R; ++r)
      for (int s = 0; s < g.S; ++s) {
        const int qy = r - g.pad_h, qx = s - g.pad_w;
        taps[nt++] = TapSpec{posmod(qy, g.stride_h), posmod(qx, g.stride_w), floordiv(qy, g.stride_h), floordiv(qx, g.stride_w),
                             r * g.S + s};
      }
    return launch_plan(g, in, g.Hin, g.Win, g.Cin, g.stride_h, g.stride_w, taps, nt, wp_tc, bias, out, g.Hout, g.Wout, o_img, o_row,
                       g.Cout, 0, g.Cout, stat_sum, stat_sumsq, stat_groups, flags, nb, st);
  }
  // transposed gather: one launch per output parity class
  auto class_taps = [&](int py, int px, TapSpec* out_taps) {
    int nt = 0;
    for (int r = 0; r < g.R; ++r) {
      if (posmod(py + g.pad_h - r, g.stride_h) != 0) continue;
      for (int s = 0; s < g.S; ++s) {
        if (posmod(px + g.pad_w - s, g.stride_w) != 0) continue;
        out_taps[nt++] = TapSpec{0, 0, floordiv(py + g.pad_h - r, g.stride_h), floordiv(px + g.pad_w - s, g.stride_w), r * g.S + s};
      }
    }
    return nt;
  };
  bool any_empty = false;
  for (int py = 0; py < g.stride_h; ++py)
    for (int px = 0; px < g.stride_w; ++px)
      if (class_taps(py, px, taps) == 0) any_empty = true;
  if (any_empty) {   // classes no filter tap reaches receive no contribution
    VIAI_REQUIRE(bias == nullptr && stat_sum == nullptr, "conv2d_tc: empty parity class with bias/statistics is unsupported");
    VIAI_CUDA(cudaMemsetAsync(out, 0, (size_t)g.N * o_img * sizeof(float), st));
  }
  for (int py = 0; py < g.stride_h; ++py)
    for (int px = 0; px < g.stride_w; ++px) {
      const int Hv = (g.Hout - py + g.stride_h - 1) / g.stride_h, Wv = (g.Wout - px + g.stride_w - 1) / g.stride_w;
      if (Hv <= 0 || Wv <= 0) continue;
      const int nt = class_taps(py, px, taps);
      if (nt == 0) continue;
      int rc = launch_plan(g, in, g.Hin, g.Win, g.Cin, 1, 1, taps, nt, wp_tc, bias, out, Hv, Wv, o_img, o_row * g.stride_h,
                           (int64_t)g.Cout * g.stride_w, (int64_t)py * o_row + (int64_t)px * g.Cout, g.Cout, stat_sum, stat_sumsq,
                           stat_groups, flags, nb, st);
      if (rc != VIAI_OK) return rc;
    }
  return VIAI_OK;
}

extern "C" int viai_conv2d_tc(const viai_conv_geom* gp, const float* in, const float* wp_tc, const float* bias, float* out,
                              double* stat_sum, double* stat_sumsq, int stat_groups, int flags, viai_stream_t stream) {
  return conv2d_tc_impl(gp, in, wp_tc, bias, out, stat_sum, stat_sumsq, stat_groups, flags, nullptr, stream);
}

extern "C" int viai_conv2d_tc_bwd_reduce(const viai_conv_geom* gp, const float* in, const float* wp_tc, float* out,
                                         const viai_norm_bwd_ctx* nb, double* s1, double* s2, int flags, viai_stream_t stream) {
  VIAI_REQUIRE(nb && nb->y && s1 && s2, "conv2d_tc_bwd_reduce: null argument");
  VIAI_REQUIRE((nb->mean == nullptr) == (nb->invstd == nullptr), "conv2d_tc_bwd_reduce: mean/invstd must both be set or NULL");
  VIAI_REQUIRE((reinterpret_cast<uintptr_t>(nb->y) & 15) == 0, "conv2d_tc_bwd_reduce: y must be 16-byte aligned");
  return conv2d_tc_impl(gp, in, wp_tc, nullptr, out, s1, s2, 1, flags, nb, stream);
}
