// Tensor-core weight gradient for sm_100a (tcgen05.mma kind::tf32, fp32 accumulation in TMEM).
//
//   dW[a][b][r][s] (+)= sum_{n,y,x} U[n,y,x,a] * G[n, y*stride_h - pad_h + r, x*stride_w - pad_w + s, b]
//
// Per filter tap this is a GEMM  D_t[a][b] = U^T (A x pixels) . G_t (pixels x B)  whose reduction runs over pixels, so both
// operands are "MN-major" in shared memory: a TMA box of (32 channels, 8 columns, TH rows) lands as dense 128-byte pixel rows
// (128B swizzle with 32B atoms -- the one layout the tensor core accepts for MN-major tf32), i.e. 8 consecutive pixels of 32
// channels are the two 512-byte atoms of one K step of 8.
//   * A operand: the U tile (TH x 8 pixels, 128 channels = 4 boxes of 32 channels, LBO = box pitch);
//   * B operand: the (TH + kh - 1) x (8 + kw - 1) patch of G (32 channels) that ALL taps of the tile read, loaded once; a tap
//     is a different start row of the descriptor (as in conv_tc.cu), strides become parity sub-grids;
//   * accumulators: one 128 x 32 fp32 block per tap in TMEM (9 taps -> 288 of the 512 columns), kept for the CTA's whole
//     pixel range (split-K across CTAs), then STORED to the CTA's own slice of a packed fp32 workspace [split][tap][A][B]
//     (no atomics: measured on B200, red.global.add.v4.f32 from ~300 CTAs onto the same tile ran at ~0.5 clk per
//     instruction for the whole GPU and was 80 % of this kernel's time);
//   * a second small kernel sums the split slices and scatters into the parameter-gradient layout (overwrite or accumulate).
// Ragged tiles and zero padding need no masking: TMA fills out-of-bounds elements of either operand with zeros.
#include "common.cuh"
#include <stdlib.h>
#include "tc_common.cuh"
using namespace viai;
using namespace viai::tc;

namespace {

constexpr int TW = 8, KC = 32, BM = 128, BNW = 32;
constexpr int MAX_SUB = 4, MAX_TAP = 16;
constexpr int NTHREADS = 320;   // warp 0: TMA producer, warp 1: TMEM + MMA issuer, warps 2..5: epilogue, warps 6..9: operand rounding
constexpr int RND_WARP0 = 6;

struct WgSub {
  int32_t ox, oy, pw, ph;
  uint32_t smem_off;     // inside the G region of a stage
  int32_t c_off;         // channel offset of this box inside the work item's G channel range (channel-group mode)
};
struct WgTap {
  uint32_t col;          // first TMEM column of the tap's 128 x 32 accumulator block
  uint32_t wtap;
  uint32_t b_off;        // G channel offset of the block inside the work item's range (channel-group mode, else 0)
};
// A run = horizontally adjacent taps of one parity sub-patch: their B operands are the same patch rows shifted by one pixel
// (128 bytes) each, so ONE MMA with N = 32 * len covers the whole run by using the pixel pitch as the descriptor's stride
// between 32-channel blocks.  (Measured on B200: an M=128, N=32, K=8 MN-major MMA costs ~100 cycles whatever N is up to
// ~128 -- the A fetch dominates -- so 3 taps per instruction make the 3x3 weight gradient ~3x faster.)
struct WgRun {
  uint32_t g_off;        // byte offset (inside the G region) of the first tap's first pixel row
  uint32_t row_pitch;    // bytes between consecutive 8-pixel row segments (pw * 128)
  uint32_t col;          // first TMEM column
  uint32_t idesc;        // instruction descriptor with N = 32 * len
  uint32_t n_pitch;      // bytes between the 32-channel blocks of the B operand: one pixel (128) for adjacent taps, one box for
                         // adjacent channel slabs (channel-group mode)
};
struct WgParams {
  CUtensorMap mapU;
  CUtensorMap mapG[MAX_SUB];
  WgSub sub[MAX_SUB];
  WgTap tap[MAX_TAP];
  WgRun run[MAX_TAP];
  int32_t nsub, ntap, nrun;
  float* ws;             // [split][tap][A][B] fp32 workspace
  int64_t ws_split;      // floats per split slice
  int32_t A, B;
  int32_t b_tile_ch;     // G channels per work item: 32, or 32 * NB in channel-group mode
  int32_t N, tilesX, tilesY, TH;
  int32_t a_tiles, b_tiles, splits, tiles_per_split, npix_tiles;
  int32_t u_slices;      // 32-channel boxes of U actually loaded per stage (<= 4)
  uint32_t u_bytes, u_slice_bytes, g_bytes, g_tx_bytes, stage_bytes;
  int32_t stages;
  uint32_t idesc;
  int32_t round_rn;      // 1: warps 6..9 round every landed operand word to the nearest tf32 in place (the tensor core TRUNCATES
                         //    fp32 operands to tf32: both operands shrink, every product is ~2^-10 too small and the weight
                         //    gradient carries that as a one-sided bias; rounded operands leave a zero-mean error that averages
                         //    out over the >= 10^4 pixels of the reduction)
};


__global__ void __launch_bounds__(NTHREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = full + p.stages;
  uint64_t* done = empty + p.stages;
  uint64_t* ready = done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + p.stages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 512;

  // work item
  int item = blockIdx.x;
  const int bt = item % p.b_tiles; item /= p.b_tiles;
  const int at = item % p.a_tiles; item /= p.a_tiles;
  const int split = item;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.npix_tiles, t_begin + p.tiles_per_split);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&ready[i], 128); }
    mbar_init(done, 1);
    fence_barrier_init();
    prefetch_tmap(&p.mapU);
    for (int i = 0; i < p.nsub; ++i) prefetch_tmap(&p.mapG[i]);
  }
  // U slices that are never loaded (A < 128) must read as zeros
  if (p.u_slices < 4) {
    for (int s = 0; s < p.stages; ++s) {
      uint8_t* u = smem + (size_t)s * p.stage_bytes;
      for (uint32_t i = p.u_slices * p.u_slice_bytes + threadIdx.x * 16; i < p.u_bytes; i += NTHREADS * 16)
        *reinterpret_cast<float4*>(u + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int rest = t;
        const int tx = rest % p.tilesX; rest /= p.tilesX;
        const int ty = rest % p.tilesY;
        const int n = rest / p.tilesY;
        const int x0 = tx * TW, y0 = ty * p.TH;
        mbar_wait(&empty[st], ph ^ 1u);
        mbar_expect_tx(&full[st], p.u_slices * p.u_slice_bytes + p.g_tx_bytes);
        uint8_t* u = smem + (size_t)st * p.stage_bytes;
        uint8_t* g = u + p.u_bytes;
        for (int s = 0; s < p.u_slices; ++s)
          tma_load_4d(u + (size_t)s * p.u_slice_bytes, &p.mapU, &full[st], at * BM + s * KC, x0, y0, n);
        for (int s = 0; s < p.nsub; ++s)
          tma_load_4d(g + p.sub[s].smem_off, &p.mapG[s], &full[st], bt * p.b_tile_ch + p.sub[s].c_off, x0 + p.sub[s].ox,
                      y0 + p.sub[s].oy, n);
        if (++st == p.stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp walks the pipeline, one elected lane issues (no per-MMA lane election loops)
    const bool leader = elect_one();
    int st = 0;
    uint32_t ph = 0;
    const uint64_t a_step = 1024u >> 4;
    uint64_t* const landed = p.round_rn ? ready : full;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&landed[st], ph);
      tc_fence_after();
      const uint32_t u_base = smem_u32(smem + (size_t)st * p.stage_bytes);
      const uint32_t g_base = u_base + p.u_bytes;
      if (leader) {
        // A: MN-major, 4 blocks of 32 channels LBO apart; B: MN-major, `len` blocks of 32 channels one pixel (128 bytes)
        // apart = the taps of the run.  SBO (between K groups of 8) is unused: every MMA covers exactly one group.
        const uint64_t ad0 = make_desc_sw128x32_mn(u_base, p.u_slice_bytes, 512u);
        for (int rn = 0; rn < p.nrun; ++rn) {
          const WgRun& w = p.run[rn];
          const uint32_t d_tmem = tmem_base + w.col;
          uint64_t ad = ad0;
          uint64_t bd = make_desc_sw128x32_mn(g_base + w.g_off, w.n_pitch, 512u);
          const uint64_t b_step = w.row_pitch >> 4;
          const uint32_t idesc = w.idesc;
          mma_tf32(d_tmem, ad, bd, idesc, (t > t_begin) ? 1u : 0u);
          for (int kk = 1; kk < p.TH; ++kk) {
            ad += a_step;
            bd += b_step;
            mma_tf32(d_tmem, ad, bd, idesc, 1u);
          }
        }
        mma_commit(&empty[st]);
      }
      if (++st == p.stages) { st = 0; ph ^= 1u; }
    }
    if (leader) mma_commit(done);
  } else if (warp >= RND_WARP0) {
    // operand rounding: x -> nearest tf32 (ties away from zero) by adding half a tf32 ulp to the bit pattern; the tensor core
    // then drops the 13 low bits.  Elementwise and in place, so the swizzled layout does not matter.
    if (p.round_rn) {
      const int tid = threadIdx.x - RND_WARP0 * 32;
      const uint32_t n16 = (p.u_slices * p.u_slice_bytes) >> 4, g16 = p.g_bytes >> 4;
      int st = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&full[st], ph);
        uint4* u = reinterpret_cast<uint4*>(smem + (size_t)st * p.stage_bytes);
        uint4* g = reinterpret_cast<uint4*>(smem + (size_t)st * p.stage_bytes + p.u_bytes);
#pragma unroll 4
        for (uint32_t i = tid; i < n16; i += 128) {
          uint4 v = u[i];
          v.x += 0x1000u; v.y += 0x1000u; v.z += 0x1000u; v.w += 0x1000u;
          u[i] = v;
        }
#pragma unroll 4
        for (uint32_t i = tid; i < g16; i += 128) {
          uint4 v = g[i];
          v.x += 0x1000u; v.y += 0x1000u; v.z += 0x1000u; v.w += 0x1000u;
          g[i] = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&ready[st]);
        if (++st == p.stages) { st = 0; ph ^= 1u; }
      }
    }
  } else {
    // epilogue: store the CTA's partial D_t blocks into its slice of the workspace
    const int q = warp & 3;
    const int a = at * BM + q * 32 + lane;
    if (t_end > t_begin) {
      mbar_wait(done, 0);
      tc_fence_after();
      for (int tp = 0; tp < p.ntap; ++tp) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + p.tap[tp].col, v);
        if (a < p.A) {
          const int b0 = bt * p.b_tile_ch + (int)p.tap[tp].b_off;
          float* dst = p.ws + (size_t)split * p.ws_split + ((size_t)p.tap[tp].wtap * p.A + a) * p.B + (size_t)b0;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (b0 + i < p.B) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dw (+)= sum over split slices.  Block = 32 consecutive elements x 8 split lanes (coalesced 128-byte reads per slice).
__global__ void __launch_bounds__(256) wgrad_unpack_kernel(const float* __restrict__ ws, float* __restrict__ dw, int A, int B, int R,
                                                           int S, int64_t sa, int64_t sb, int64_t sr, int64_t ss, int accumulate,
                                                           int splits, int64_t ws_split) {
  __shared__ float part[8][33];
  const int64_t total = (int64_t)R * S * A * B;
  const int li = threadIdx.x & 31, lp = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
    const int64_t idx = base + li;
    float acc = 0.f;
    if (idx < total)
      for (int sp = lp; sp < splits; sp += 8) acc += ws[(int64_t)sp * ws_split + idx];
    part[lp][li] = acc;
    __syncthreads();
    if (lp == 0 && idx < total) {
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) v += part[q][li];
      const int b = idx % B;
      int64_t t = idx / B;
      const int a = t % A;
      const int tap = (int)(t / A);
      const int r = tap / S, s2 = tap % S;
      float* d = dw + a * sa + b * sb + r * sr + s2 * ss;
      *d = accumulate ? (*d + v) : v;
    }
    __syncthreads();
  }
}

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

}  // namespace

extern "C" int viai_conv2d_wgrad_tc_supported(const viai_conv_geom* g) {
  if (!g) return 0;
  if (g->Cin % 4 != 0 || g->Cin < 16 || g->Cout % 4 != 0 || g->Cout < 16) return 0;
  if (g->R * g->S > MAX_TAP || g->R * g->S * BNW > 512) return 0;
  if (g->stride_h < 1 || g->stride_h > 2 || g->stride_w < 1 || g->stride_w > 2) return 0;
  return 1;
}

namespace {
// Channel-group mode for 1x1 / stride-1 / unpadded convolutions (the WaveNet training GEMMs, ResNet projections): there is a
// single tap, so instead of the taps of a run the N = 32 * NB columns of one MMA are NB adjacent 32-channel slabs of G (the
// descriptor's block stride is the box pitch, exactly as for the A operand).  One work item then covers 128 x (32 NB) of dW:
// NB x fewer passes over U (the kernel is L2->SM bound here: 4.2 GB -> 1.7 GB for the 1632 -> 512 WaveNet layer at 32 000 rows).
// VIAI_WGRAD_GROUP=1 restores one slab per work item.
int wg_group(const viai_conv_geom& g) {
  static int nb_env = -1;
  if (nb_env < 0) {
    const char* e = getenv("VIAI_WGRAD_GROUP");
    nb_env = e ? atoi(e) : 4;
    if (nb_env < 1 || nb_env > 4) nb_env = 4;
  }
  const bool one = g.R == 1 && g.S == 1 && g.stride_h == 1 && g.stride_w == 1 && g.pad_h == 0 && g.pad_w == 0;
  return (one && g.Cin > BNW) ? nb_env : 1;
}
// split-K plan shared by the workspace query and the launch: the pixel range is split so that the grid is ~2 waves of the SMs
void wg_split_plan(const viai_conv_geom& g, int& TH, int& npix_tiles, int& splits, int& tiles_per_split) {
  const bool strided = g.stride_h > 1 || g.stride_w > 1;
  const int nb = wg_group(g);
  TH = (strided || nb > 1) ? 8 : 16;
  const int tilesX = (g.Wout + TW - 1) / TW, tilesY = (g.Hout + TH - 1) / TH;
  npix_tiles = g.N * tilesX * tilesY;
  const int ab = ((g.Cout + BM - 1) / BM) * ((g.Cin + BNW * nb - 1) / (BNW * nb));
  int sp = (2 * kNumSMs + ab - 1) / ab;
  if (sp > npix_tiles) sp = npix_tiles;
  if (sp < 1) sp = 1;
  tiles_per_split = (npix_tiles + sp - 1) / sp;
  splits = tiles_per_split > 0 ? (npix_tiles + tiles_per_split - 1) / tiles_per_split : 1;
}
}  // namespace

extern "C" int64_t viai_wgrad_tc_workspace(const viai_conv_geom* g) {
  if (!g) return 0;
  int TH, npix, splits, tps;
  wg_split_plan(*g, TH, npix, splits, tps);
  return (int64_t)(splits > 0 ? splits : 1) * g->R * g->S * g->Cout * g->Cin;
}

// U: (N,Hout,Wout,Cout=A), G: (N,Hin,Win,Cin=B) as in viai_conv2d_wgrad_simt.  workspace: viai_wgrad_tc_workspace(g) floats.
extern "C" int viai_conv2d_wgrad_tc(const viai_conv_geom* gp, const float* U, const float* G, float* dw, int64_t sa, int64_t sb,
                                    int64_t sr, int64_t ss, int accumulate, float* workspace, viai_stream_t stream) {
  VIAI_REQUIRE(gp && U && G && dw && workspace, "conv2d_wgrad_tc: null argument");
  const viai_conv_geom& g = *gp;
  VIAI_REQUIRE(viai_conv2d_wgrad_tc_supported(gp), "conv2d_wgrad_tc: unsupported geometry");
  VIAI_REQUIRE((reinterpret_cast<uintptr_t>(U) & 15) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
               "conv2d_wgrad_tc: pointers must be 16-byte aligned");
  cudaStream_t st = STR(stream);
  WgParams p;
  memset(&p, 0, sizeof(p));
  const int A = g.Cout, B = g.Cin;
  p.A = A; p.B = B; p.ws = workspace; p.N = g.N;
  {
    int TH, npix, splits, tps;
    wg_split_plan(g, TH, npix, splits, tps);
    p.TH = TH; p.npix_tiles = npix; p.splits = splits; p.tiles_per_split = tps;
  }
  const int nb = wg_group(g);
  p.b_tile_ch = BNW * nb;
  // taps and parity sub-patches of G
  int sub_id[2][2] = {{-1, -1}, {-1, -1}};
  int mn_y[MAX_SUB], mx_y[MAX_SUB], mn_x[MAX_SUB], mx_x[MAX_SUB], sy_of[MAX_SUB], sx_of[MAX_SUB], t_sub[MAX_TAP], t_oy[MAX_TAP], t_ox[MAX_TAP];
  int nsub = 0, ntap = 0;
  for (int r = 0; r < g.R; ++r)
    for (int s = 0; s < g.S; ++s) {
      const int qy = r - g.pad_h, qx = s - g.pad_w;
      const int py = posmod(qy, g.stride_h), px = posmod(qx, g.stride_w), oy = floordiv(qy, g.stride_h), ox = floordiv(qx, g.stride_w);
      int& id = sub_id[py][px];
      if (id < 0) {
        id = nsub++;
        sy_of[id] = py; sx_of[id] = px;
        mn_y[id] = mx_y[id] = oy; mn_x[id] = mx_x[id] = ox;
      } else {
        mn_y[id] = oy < mn_y[id] ? oy : mn_y[id]; mx_y[id] = oy > mx_y[id] ? oy : mx_y[id];
        mn_x[id] = ox < mn_x[id] ? ox : mn_x[id]; mx_x[id] = ox > mx_x[id] ? ox : mx_x[id];
      }
      t_sub[ntap] = id; t_oy[ntap] = oy; t_ox[ntap] = ox;
      p.tap[ntap].wtap = (uint32_t)(r * g.S + s);
      ++ntap;
    }
  if (nb > 1) {            // one real tap at offset (0, 0): replicate its box / accumulator block per channel slab
    nsub = ntap = nb;
    for (int i = 1; i < nb; ++i) {
      sy_of[i] = sy_of[0]; sx_of[i] = sx_of[0];
      mn_y[i] = mx_y[i] = mn_y[0]; mn_x[i] = mx_x[i] = mn_x[0];
      t_sub[i] = i; t_oy[i] = t_oy[0]; t_ox[i] = t_ox[0];
      p.tap[i].wtap = 0;
    }
  }
  p.nsub = nsub; p.ntap = ntap;
  uint32_t off = 0;
  for (int s = 0; s < nsub; ++s) {
    WgSub& sb2 = p.sub[s];
    sb2.ox = mn_x[s]; sb2.oy = mn_y[s];
    sb2.pw = TW + (mx_x[s] - mn_x[s]);
    sb2.ph = p.TH + (mx_y[s] - mn_y[s]);
    sb2.smem_off = off;
    sb2.c_off = nb > 1 ? s * BNW : 0;
    off += (uint32_t)(sb2.pw * sb2.ph) * 128u;
    p.g_tx_bytes += (uint32_t)(sb2.pw * sb2.ph) * 128u;
    off = (off + 1023u) & ~1023u;
    const int subH = (g.Hin - sy_of[s] + g.stride_h - 1) / g.stride_h, subW = (g.Win - sx_of[s] + g.stride_w - 1) / g.stride_w;
    VIAI_REQUIRE(subH >= 1 && subW >= 1, "conv2d_wgrad_tc: empty parity sub-grid");
    const float* base = G + ((int64_t)sy_of[s] * g.Win + sx_of[s]) * B;
    uint64_t dims[4] = {(uint64_t)B, (uint64_t)subW, (uint64_t)subH, (uint64_t)g.N};
    uint64_t strides[3] = {(uint64_t)g.stride_w * B * 4, (uint64_t)g.stride_h * g.Win * B * 4, (uint64_t)g.Hin * g.Win * B * 4};
    uint32_t box[4] = {KC, (uint32_t)sb2.pw, (uint32_t)sb2.ph, 1};
    if (encode_f32_map(&p.mapG[s], 4, base, dims, strides, box, 2)) return VIAI_ERR_CUDA;
  }
  p.g_bytes = off;
  if (nb > 1) {
    WgRun& rn = p.run[0];
    rn.g_off = p.sub[0].smem_off;
    rn.row_pitch = (uint32_t)p.sub[0].pw * 128u;
    rn.col = 0;
    rn.n_pitch = p.sub[1].smem_off - p.sub[0].smem_off;
    rn.idesc = make_idesc_tf32(BM, BNW * nb, 1, 1);
    for (int i = 0; i < nb; ++i) { p.tap[i].col = (uint32_t)(BNW * i); p.tap[i].b_off = (uint32_t)(BNW * i); }
    p.nrun = 1;
  } else {
    bool placed[MAX_TAP] = {false};
    int nrun = 0;
    uint32_t col = 0;
    for (int t0 = 0; t0 < ntap; ++t0) {
      if (placed[t0]) continue;
      // left-most unplaced tap of its (sub-patch, row): taps are enumerated with increasing s, hence increasing ox
      const WgSub& sb2 = p.sub[t_sub[t0]];
      WgRun& rn = p.run[nrun++];
      rn.g_off = sb2.smem_off + (uint32_t)((t_oy[t0] - sb2.oy) * sb2.pw + (t_ox[t0] - sb2.ox)) * 128u;
      rn.row_pitch = (uint32_t)sb2.pw * 128u;
      rn.n_pitch = 128u;
      rn.col = col;
      int len = 0, cur = t0;
      while (cur >= 0 && len < 8) {
        placed[cur] = true;
        p.tap[cur].col = col;
        col += BNW;
        ++len;
        int nxt = -1;
        for (int t = 0; t < ntap; ++t)
          if (!placed[t] && t_sub[t] == t_sub[cur] && t_oy[t] == t_oy[cur] && t_ox[t] == t_ox[cur] + 1) { nxt = t; break; }
        cur = nxt;
      }
      rn.idesc = make_idesc_tf32(BM, BNW * len, 1, 1);
    }
    p.nrun = nrun;
  }
  {
    uint64_t dims[4] = {(uint64_t)A, (uint64_t)g.Wout, (uint64_t)g.Hout, (uint64_t)g.N};
    uint64_t strides[3] = {(uint64_t)A * 4, (uint64_t)g.Wout * A * 4, (uint64_t)g.Hout * g.Wout * A * 4};
    uint32_t box[4] = {KC, TW, (uint32_t)p.TH, 1};
    if (encode_f32_map(&p.mapU, 4, U, dims, strides, box, 2)) return VIAI_ERR_CUDA;
  }
  p.u_slice_bytes = (uint32_t)(p.TH * TW) * 128u;
  p.u_bytes = 4 * p.u_slice_bytes;
  p.a_tiles = (A + BM - 1) / BM;
  p.b_tiles = (B + p.b_tile_ch - 1) / p.b_tile_ch;
  {
    const int rem = A - 0;   // slices needed by the widest a-tile; narrower (last) tiles read zero-filled boxes
    const int need = rem >= BM ? 4 : (rem + KC - 1) / KC;
    p.u_slices = need;
  }
  p.stage_bytes = p.u_bytes + p.g_bytes;
  p.tilesX = (g.Wout + TW - 1) / TW;
  p.tilesY = (g.Hout + p.TH - 1) / p.TH;
  const int ab = p.a_tiles * p.b_tiles;
  p.ws_split = (int64_t)g.R * g.S * A * B;
  static const int round_env = [] { const char* e = getenv("VIAI_WGRAD_ROUND"); return e ? atoi(e) : 1; }();
  p.round_rn = round_env ? 1 : 0;
  const size_t budget = 227 * 1024, fixed = 1024 + 32 * 8 + 16;
  int stages = 4;
  while (stages > 1 && fixed + (size_t)stages * p.stage_bytes > budget) --stages;
  VIAI_REQUIRE(fixed + (size_t)stages * p.stage_bytes <= budget, "conv2d_wgrad_tc: stage of %u bytes does not fit", p.stage_bytes);
  p.stages = stages;
  p.idesc = make_idesc_tf32(BM, BNW, 1, 1);
  size_t smem = fixed + (size_t)stages * p.stage_bytes;
  if (smem < 120 * 1024) smem = 120 * 1024;
  static bool attr_set = false;
  if (!attr_set) {
    VIAI_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int64_t ws_elems = (int64_t)g.R * g.S * A * B;
  if (p.npix_tiles == 0) {
    VIAI_CUDA(cudaMemsetAsync(workspace, 0, (size_t)ws_elems * sizeof(float), st));
    p.splits = 1;
  }
  const int grid = p.splits * ab;
  if (p.npix_tiles > 0) {
    wgrad_tc_kernel<<<grid, NTHREADS, smem, st>>>(p);
    VIAI_LAUNCHED();
  }
  const int blocks = (int)imin64(cdiv(ws_elems, 32), 16 * kNumSMs);
  wgrad_unpack_kernel<<<blocks, 256, 0, st>>>(workspace, dw, A, B, g.R, g.S, sa, sb, sr, ss, accumulate, p.splits, p.ws_split);
  VIAI_LAUNCHED();
  return VIAI_OK;
}
