// Depthwise 3x3 convolution + SimpleGate (forward / backward), NHWC bf16, TMA-staged tiles.
// Reference: NAFBlock.conv2 (nafnet_arch.py:96-104, groups = 2C, padding 1, bias) followed by
// SimpleGate (nafnet_arch.py:77-80): g[:, j] = v[:, j] * v[:, j + C].
//
// CUDA-core by design (<= 36 FLOP per 4 bytes moved: HBM-bound, tensor cores have nothing to contract).
// All three kernels share one structure:
//   * persistent CTAs (8 warps); a CTA is bound to one 64-channel group, so the nine taps of its two
//     channels per lane stay in registers for the whole launch, and walks a contiguous range of
//     8 x 16-pixel tiles;
//   * each tile (+1-pixel halo) is brought into shared memory by ONE 4-D TMA box per operand
//     (cp.async.bulk.tensor, zero fill outside the image = the conv's zero padding, no bounds checks),
//     double-buffered on mbarriers so the load of tile i+1 overlaps the math of tile i;
//   * warp = one tile row, lane = 2 channels: the row is swept left to right with a 3x3 register window
//     (3 conflict-free 128-byte LDS per operand per pixel), stores are 128-byte coalesced;
//   * column sums (SCA pool, bias / weight gradients) live in registers across tiles and are reduced once
//     per CTA (once per image for the pool).
#include "elementwise.cuh"

namespace {

constexpr int TH = 8, TW = 16, NWARP = 8;
constexpr int HH = TH + 2, HWD = TW + 2;
constexpr int BOX_ELEMS = HH * HWD * 64;       // one halo box: 10 x 18 pixels x 64 channels
constexpr int BOX_BYTES = BOX_ELEMS * 2;       // 23040
constexpr int DG_ELEMS = TH * TW * 64;         // interior box (no halo)
constexpr int DG_BYTES = DG_ELEMS * 2;         // 16384

// Operand tiles live in dynamic shared memory behind an aligned-up uintptr_t, which makes every pointer derived from it GENERIC to
// the compiler: the stencil loads compiled to LD.E (r01_w_ncu_bwd_elementwise_hotspots.txt: dispatch stalls, issue-bound kernels).
// SBf carries the 32-bit shared-window address instead, and lds_bf2 is an explicit ld.shared (LDS with an immediate offset).
struct SBf {
  uint32_t addr;
  __device__ __forceinline__ SBf operator+(int elems) const { return SBf{addr + 2u * (uint32_t)elems}; }
};
__device__ __forceinline__ SBf sbf(const void* generic_smem_ptr) { return SBf{smem_u32(generic_smem_ptr)}; }
__device__ __forceinline__ float2 lds_bf2(SBf p) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(p.addr));
  return OP2_TO_F32(*reinterpret_cast<const op16x2*>(&v));
}
__device__ __forceinline__ void st_bf2(bf16* p, float2 v) {
  *reinterpret_cast<op16x2*>(p) = OP2_FROM_F32(v.x, v.y);
}
__device__ __forceinline__ float2 round_bf2(float2 v) { return OP2_TO_F32(OP2_FROM_F32(v.x, v.y)); }
// Two fp32 FMAs in ONE issue slot: Blackwell's packed FFMA2 (fma.rn.f32x2).  The stencil kernels are issue-bound (r02s: ~1700
// instructions per tile row, a third of them FFMA), and every accumulation here is already a float2 = two adjacent channels.
__device__ __forceinline__ void fma2(float2& acc, float2 a, float2 b) {
  unsigned long long c = *reinterpret_cast<unsigned long long*>(&acc);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(c)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  acc = *reinterpret_cast<float2*>(&c);
}
// nine taps of two consecutive channels: w[c][tap] fp32, c = c0, c0+1
__device__ __forceinline__ void load_taps(float2 (&wt)[9], const float* __restrict__ w, int c0) {
#pragma unroll
  for (int k = 0; k < 9; ++k) wt[k] = make_float2(__ldg(w + (size_t)c0 * 9 + k), __ldg(w + (size_t)(c0 + 1) * 9 + k));
}

struct Tiles {
  int tiles_w, tiles_h, per_img, total, t0, t1;
  __device__ Tiles(int N, int H, int W) {
    tiles_w = (W + TW - 1) / TW;
    tiles_h = (H + TH - 1) / TH;
    per_img = tiles_w * tiles_h;
    total = N * per_img;
    // contiguous range per CTA, sizes differing by at most one tile (every CTA of the grid gets work: with ceil-sized chunks
    // 128 tiles on 37 CTAs left 5 CTAs idle and gave the others 4 tiles each instead of 3-4)
    t0 = (int)(((long long)total * blockIdx.x) / gridDim.x);
    t1 = (int)(((long long)total * (blockIdx.x + 1)) / gridDim.x);
  }
  __device__ void decode(int t, int& n, int& h0, int& w0) const {
    n = t / per_img;
    const int r = t - n * per_img;
    h0 = (r / tiles_w) * TH;
    w0 = (r % tiles_w) * TW;
  }
};

// ------------------------------- forward --------------------------------------
// GELU = 0: SimpleGate a * b (NAFBlock).  GELU = 1: gelu(a) * b with the exact erf GELU (Restormer GDFN,
// restormer_arch.py:97-98; its dw conv has no bias and nothing is pooled: b2 == pool == nullptr).
template <int GELU>
__global__ void __launch_bounds__(NWARP * 32)
dwgate_fwd_kernel(const __grid_constant__ CUtensorMap tmU, const float* __restrict__ w2, const float* __restrict__ b2,
                  bf16* __restrict__ g, float* __restrict__ pool, int N, int H, int W, int C) {
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full[2];
  __shared__ float2 s_pool[NWARP][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = blockIdx.y, c2 = cg * 64 + lane * 2;
  const bool chan_ok = c2 < C;
  const int ca = chan_ok ? c2 : 0, cb = C + ca;
  float2 wa[9], wb[9];
  load_taps(wa, w2, ca);
  load_taps(wb, w2, cb);
  const float2 zero2 = make_float2(0.f, 0.f);
  const float2 ba = b2 ? make_float2(__ldg(b2 + ca), __ldg(b2 + ca + 1)) : zero2;
  const float2 bb = b2 ? make_float2(__ldg(b2 + cb), __ldg(b2 + cb + 1)) : zero2;
  const Tiles T(N, H, W);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmU);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int t, int s) {
    int n, h0, w0;
    T.decode(t, n, h0, w0);
    uint8_t* dst = smem + (size_t)s * 2 * BOX_BYTES;
    mbar_arrive_expect_tx(&full[s], 2 * BOX_BYTES);
    tma_load_4d(dst, &tmU, &full[s], cg * 64, w0 - 1, h0 - 1, n);
    tma_load_4d(dst + BOX_BYTES, &tmU, &full[s], C + cg * 64, w0 - 1, h0 - 1, n);
  };
  if (threadIdx.x == 0 && T.t0 < T.t1) issue(T.t0, 0);
  float2 psum = make_float2(0.f, 0.f);
  int cur_n = -1;
  auto flush_pool = [&](int n) {  // CTA-uniform
    s_pool[warp][lane] = psum;
    __syncthreads();
    if (warp == 0 && chan_ok) {
      float2 t = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < NWARP; ++k) {
        t.x += s_pool[k][lane].x;
        t.y += s_pool[k][lane].y;
      }
      atomicAdd(pool + (size_t)n * C + c2, t.x);
      atomicAdd(pool + (size_t)n * C + c2 + 1, t.y);
    }
    psum = make_float2(0.f, 0.f);
  };
  int it = 0;
  for (int t = T.t0; t < T.t1; ++t, ++it) {
    const int s = it & 1;
    if (threadIdx.x == 0 && t + 1 < T.t1) issue(t + 1, s ^ 1);
    int n, h0, w0;
    T.decode(t, n, h0, w0);
    if (n != cur_n) {
      if (pool && cur_n >= 0) flush_pool(cur_n);
      cur_n = n;
    }
    mbar_wait(&full[s], (it >> 1) & 1);
    const SBf sA = sbf(smem + (size_t)s * 2 * BOX_BYTES) + lane * 2;
    const SBf sB = sA + BOX_ELEMS;
    const int h = h0 + warp;  // warp = tile row (TH == NWARP)
    bf16* grow = g + (((size_t)n * H + h) * W + w0) * C + ca;
    float2 A[3][3], B[3][3];
#pragma unroll
    for (int x = 0; x < HWD; ++x) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        A[r][x % 3] = lds_bf2(sA + ((warp + r) * HWD + x) * 64);
        B[r][x % 3] = lds_bf2(sB + ((warp + r) * HWD + x) * 64);
      }
      if (x >= 2) {
        float2 a = ba, b = bb;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            fma2(a, A[r][(x - 2 + d) % 3], wa[r * 3 + d]);
            fma2(b, B[r][(x - 2 + d) % 3], wb[r * 3 + d]);
          }
        const int ox = x - 2;
        if (chan_ok && h < H && w0 + ox < W) {
          if constexpr (GELU) {
            a.x = 0.5f * a.x * (1.f + erff(a.x * 0.70710678118654752f));
            a.y = 0.5f * a.y * (1.f + erff(a.y * 0.70710678118654752f));
          }
          const float2 gv = round_bf2(make_float2(a.x * b.x, a.y * b.y));
          st_bf2(grow + (size_t)ox * C, gv);
          psum.x += gv.x;
          psum.y += gv.y;
        }
      }
    }
    __syncthreads();  // stage s may be refilled by the next-but-one issue
  }
  if (pool && cur_n >= 0) flush_pool(cur_n);
}

// ------------------- plain depthwise 3x3 forward (+ row norms) ------------------
// out[px][c] = sum_{ky,kx} x[h+ky-1, w+kx-1][c] * w[c][ky][kx]   (no bias: Restormer qkv_dwconv, restormer_arch.py:110-118)
// sumsq[n][c] += sum_px out^2 for c < sq_ch: the squared L2 norms over all pixels that MDTA's F.normalize needs
// (restormer_arch.py:131-132), taken from the bf16-rounded values the attention GEMMs will read.
__global__ void __launch_bounds__(NWARP * 32)
dwconv3_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w, bf16* __restrict__ out,
                   float* __restrict__ sumsq, int sq_ch, int N, int H, int W, int CH) {
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full[2];
  __shared__ float2 s_sq[NWARP][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = blockIdx.y, c2 = cg * 64 + lane * 2;
  const bool chan_ok = c2 < CH;
  const int c0 = chan_ok ? c2 : 0;
  const bool want_sq = sumsq != nullptr && cg * 64 < sq_ch;  // CTA-uniform
  float2 wt[9];
  load_taps(wt, w, c0);
  const Tiles T(N, H, W);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int t, int s) {
    int n, h0, w0;
    T.decode(t, n, h0, w0);
    mbar_arrive_expect_tx(&full[s], BOX_BYTES);
    tma_load_4d(smem + (size_t)s * BOX_BYTES, &tmX, &full[s], cg * 64, w0 - 1, h0 - 1, n);
  };
  if (threadIdx.x == 0 && T.t0 < T.t1) issue(T.t0, 0);
  float2 sq = make_float2(0.f, 0.f);
  int cur_n = -1;
  auto flush_sq = [&](int n) {  // CTA-uniform
    s_sq[warp][lane] = sq;
    __syncthreads();
    if (warp == 0 && chan_ok) {
      float2 t = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < NWARP; ++k) {
        t.x += s_sq[k][lane].x;
        t.y += s_sq[k][lane].y;
      }
      if (c2 < sq_ch) atomicAdd(sumsq + (size_t)n * sq_ch + c2, t.x);
      if (c2 + 1 < sq_ch) atomicAdd(sumsq + (size_t)n * sq_ch + c2 + 1, t.y);
    }
    __syncthreads();
    sq = make_float2(0.f, 0.f);
  };
  int it = 0;
  for (int t = T.t0; t < T.t1; ++t, ++it) {
    const int s = it & 1;
    if (threadIdx.x == 0 && t + 1 < T.t1) issue(t + 1, s ^ 1);
    int n, h0, w0;
    T.decode(t, n, h0, w0);
    if (n != cur_n) {
      if (want_sq && cur_n >= 0) flush_sq(cur_n);
      cur_n = n;
    }
    mbar_wait(&full[s], (it >> 1) & 1);
    const SBf sA = sbf(smem + (size_t)s * BOX_BYTES) + lane * 2;
    const int h = h0 + warp;
    bf16* orow = out + (((size_t)n * H + h) * W + w0) * CH + c0;
    float2 A[3][3];
#pragma unroll
    for (int x = 0; x < HWD; ++x) {
#pragma unroll
      for (int r = 0; r < 3; ++r) A[r][x % 3] = lds_bf2(sA + ((warp + r) * HWD + x) * 64);
      if (x >= 2) {
        const int ox = x - 2;
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int d = 0; d < 3; ++d) fma2(a, A[r][(x - 2 + d) % 3], wt[r * 3 + d]);
        if (chan_ok && h < H && w0 + ox < W) {
          a = round_bf2(a);
          st_bf2(orow + (size_t)ox * CH, a);
          sq.x = fmaf(a.x, a.x, sq.x);
          sq.y = fmaf(a.y, a.y, sq.y);
        }
      }
    }
    __syncthreads();
  }
  if (want_sq && cur_n >= 0) flush_sq(cur_n);
}

// --------------------------- backward, part a ---------------------------------
// dg = dgs*s + t;  a,b = dwconv(u)+bias (recomputed);  du2_a = dg*b, du2_b = dg*a  (stored bf16);
// dW2[c][tap] += du2[c] * u[px+tap][c];  db2[c] += du2[c].
// MODE 0: NAFBlock (SimpleGate, SCA scale / shift s, t, bias).  MODE 1: Restormer GDFN (gelu(a) * b, no bias, dg = dgs;
// restormer_arch.py:97-98).  MODE 2: plain depthwise conv weight gradient (Restormer qkv_dwconv, :110-118): tmD holds the
// output gradient of all 2C channels, nothing is recomputed or stored, only dW2 (and db2) is accumulated.
// MODE 3: MODE 0 without the weight / bias gradient (18 of the 36 FMA2 per pixel and 38 accumulator registers less): the
// NAFBlock backward runs it on its critical chain and hands dW2 / db2 to a MODE 2 launch on the weight-gradient stream.
template <int MODE>
__global__ void __launch_bounds__(NWARP * 32)
dwgate_bwd_a_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmD, const float* __restrict__ s_sca,
                    const float* __restrict__ t_sca, const float* __restrict__ w2, const float* __restrict__ b2,
                    bf16* __restrict__ du2, float* __restrict__ dw2, float* __restrict__ db2, int N, int H, int W, int C,
                    const float* __restrict__ ds_sca = nullptr, const float* __restrict__ w_sca = nullptr, float inv_hw = 0.f) {
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  constexpr int STAGE_BYTES = 2 * BOX_BYTES + (MODE == 2 ? 2 : 1) * DG_BYTES;
  float2* s_red = reinterpret_cast<float2*>(smem + 2 * STAGE_BYTES);  // [NWARP][20][32]
  __shared__ uint64_t full[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = blockIdx.y, c2 = cg * 64 + lane * 2;
  const bool chan_ok = c2 < C;
  const int ca = chan_ok ? c2 : 0, cb = C + ca;
  const int C2 = 2 * C;
  float2 wa[9], wb[9], gA[9], gB[9];
  load_taps(wa, w2, ca);
  load_taps(wb, w2, cb);
#pragma unroll
  for (int k = 0; k < 9; ++k) gA[k] = gB[k] = make_float2(0.f, 0.f);
  float2 dbA = make_float2(0.f, 0.f), dbB = dbA;
  const float2 zero2 = make_float2(0.f, 0.f);
  const float2 ba = b2 ? make_float2(__ldg(b2 + ca), __ldg(b2 + ca + 1)) : zero2;
  const float2 bb = b2 ? make_float2(__ldg(b2 + cb), __ldg(b2 + cb + 1)) : zero2;
  const Tiles T(N, H, W);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmU);
    tma_prefetch_desc(&tmD);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int t, int s) {
    int n, h0, w0;
    T.decode(t, n, h0, w0);
    uint8_t* dst = smem + (size_t)s * STAGE_BYTES;
    mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
    tma_load_4d(dst, &tmU, &full[s], cg * 64, w0 - 1, h0 - 1, n);
    tma_load_4d(dst + BOX_BYTES, &tmU, &full[s], C + cg * 64, w0 - 1, h0 - 1, n);
    tma_load_4d(dst + 2 * BOX_BYTES, &tmD, &full[s], cg * 64, w0, h0, n);
    if constexpr (MODE == 2) tma_load_4d(dst + 2 * BOX_BYTES + DG_BYTES, &tmD, &full[s], C + cg * 64, w0, h0, n);
  };
  if (threadIdx.x == 0 && T.t0 < T.t1) issue(T.t0, 0);
  int it = 0;
  int t_n = -1;            // image whose SCA shift t_cur belongs to (MODE 0, in-kernel mat-vec)
  float2 t_cur = zero2;
  for (int t = T.t0; t < T.t1; ++t, ++it) {
    const int s = it & 1;
    if (threadIdx.x == 0 && t + 1 < T.t1) issue(t + 1, s ^ 1);
    int n, h0, w0;
    T.decode(t, n, h0, w0);
    float2 sv = make_float2(1.f, 1.f), tv = zero2;
    if constexpr (MODE == 0 || MODE == 3) {
      sv = make_float2(__ldg(s_sca + (size_t)n * C + ca), __ldg(s_sca + (size_t)n * C + ca + 1));
      if (t_sca) {
        tv = make_float2(__ldg(t_sca + (size_t)n * C + ca), __ldg(t_sca + (size_t)n * C + ca + 1));
      } else {
        // SCA backward folded in (nafnet_arch.py:173 backward: the pooled branch hands every pixel of image n the same shift
        // t[n][c] = (1/HW) sum_co W_sca[co][c] ds[n][co]): the CTA needs it for its 64 channels only, once per image it visits
        // (1-2 per launch), and the first tile's TMA is already in flight.  Warps split the rows of W_sca, a lane owns its
        // two channels' columns; CTA-uniform branch.
        if (n != t_n) {
          float2 a4[4] = {zero2, zero2, zero2, zero2};
          const float* dsn = ds_sca + (size_t)n * C;
          int co = warp;
          for (; co + 3 * NWARP < C; co += 4 * NWARP) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 wv = __ldg(reinterpret_cast<const float2*>(w_sca + (size_t)(co + NWARP * k) * C + ca));
              const float d = __ldg(dsn + co + NWARP * k);
              a4[k].x = fmaf(wv.x, d, a4[k].x);
              a4[k].y = fmaf(wv.y, d, a4[k].y);
            }
          }
          for (; co < C; co += NWARP) {
            const float2 wv = __ldg(reinterpret_cast<const float2*>(w_sca + (size_t)co * C + ca));
            const float d = __ldg(dsn + co);
            a4[0].x = fmaf(wv.x, d, a4[0].x);
            a4[0].y = fmaf(wv.y, d, a4[0].y);
          }
          s_red[warp * 32 + lane] = make_float2((a4[0].x + a4[1].x) + (a4[2].x + a4[3].x), (a4[0].y + a4[1].y) + (a4[2].y + a4[3].y));
          __syncthreads();
          float2 tot = zero2;
#pragma unroll
          for (int k = 0; k < NWARP; ++k) {
            const float2 e = s_red[k * 32 + lane];
            tot.x += e.x;
            tot.y += e.y;
          }
          t_cur = make_float2(tot.x * inv_hw, tot.y * inv_hw);
          t_n = n;
          __syncthreads();  // s_red is free again (the end-of-kernel reduction also uses it)
        }
        tv = t_cur;
      }
    }
    mbar_wait(&full[s], (it >> 1) & 1);
    const SBf sA = sbf(smem + (size_t)s * STAGE_BYTES) + lane * 2;
    const SBf sB = sA + BOX_ELEMS;
    const SBf sD = sA + 2 * BOX_ELEMS;
    const int h = h0 + warp;
    bf16* orow = du2 + (((size_t)n * H + h) * W + w0) * C2;
    float2 A[3][3], B[3][3];
#pragma unroll
    for (int x = 0; x < HWD; ++x) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        A[r][x % 3] = lds_bf2(sA + ((warp + r) * HWD + x) * 64);
        B[r][x % 3] = lds_bf2(sB + ((warp + r) * HWD + x) * 64);
      }
      if (x >= 2) {
        const int ox = x - 2;
        float2 a = ba, b = bb;
        if constexpr (MODE != 2) {
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              fma2(a, A[r][(x - 2 + d) % 3], wa[r * 3 + d]);
              fma2(b, B[r][(x - 2 + d) % 3], wb[r * 3 + d]);
            }
        }
        if (chan_ok && h < H && w0 + ox < W) {
          const float2 dgv = lds_bf2(sD + (warp * TW + ox) * 64);
          float2 da, db;
          if constexpr (MODE == 0 || MODE == 3) {
            const float2 dg = make_float2(fmaf(dgv.x, sv.x, tv.x), fmaf(dgv.y, sv.y, tv.y));
            da = make_float2(dg.x * b.x, dg.y * b.y);
            db = make_float2(dg.x * a.x, dg.y * a.y);
          } else if constexpr (MODE == 1) {
            // g = gelu(a) * b, gelu(a) = a * Phi(a):  dg/da = b * (Phi(a) + a * phi(a)),  dg/db = a * Phi(a)
            const float px_ = 0.5f * (1.f + erff(a.x * 0.70710678118654752f)), py_ = 0.5f * (1.f + erff(a.y * 0.70710678118654752f));
            const float qx = 0.3989422804014327f * __expf(-0.5f * a.x * a.x), qy = 0.3989422804014327f * __expf(-0.5f * a.y * a.y);
            da = make_float2(dgv.x * b.x * (px_ + a.x * qx), dgv.y * b.y * (py_ + a.y * qy));
            db = make_float2(dgv.x * a.x * px_, dgv.y * a.y * py_);
          } else {
            da = dgv;
            db = lds_bf2(sD + DG_ELEMS + (warp * TW + ox) * 64);
          }
          if constexpr (MODE != 2) {
            st_bf2(orow + (size_t)ox * C2 + ca, da);
            st_bf2(orow + (size_t)ox * C2 + cb, db);
          }
          if constexpr (MODE != 3) {
            dbA.x += da.x; dbA.y += da.y;
            dbB.x += db.x; dbB.y += db.y;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                fma2(gA[r * 3 + d], da, A[r][(x - 2 + d) % 3]);
                fma2(gB[r * 3 + d], db, B[r][(x - 2 + d) % 3]);
              }
          }
        }
      }
    }
    __syncthreads();
  }
  if constexpr (MODE != 3) {
  // CTA reduction of the 20 x 64 channel sums, then one atomic per value
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    s_red[(warp * 20 + k) * 32 + lane] = gA[k];
    s_red[(warp * 20 + 9 + k) * 32 + lane] = gB[k];
  }
  s_red[(warp * 20 + 18) * 32 + lane] = dbA;
  s_red[(warp * 20 + 19) * 32 + lane] = dbB;
  __syncthreads();
  for (int idx = threadIdx.x; idx < 20 * 32; idx += blockDim.x) {
    const int item = idx >> 5, ln = idx & 31;
    const int cc = cg * 64 + ln * 2;
    if (cc >= C) continue;
    float2 v = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < NWARP; ++k) {
      const float2 e = s_red[(k * 20 + item) * 32 + ln];
      v.x += e.x;
      v.y += e.y;
    }
    if (item < 18) {
      const int half = item / 9, tap = item % 9;
      const int c = half * C + cc;
      atomicAdd(dw2 + (size_t)c * 9 + tap, v.x);
      atomicAdd(dw2 + (size_t)(c + 1) * 9 + tap, v.y);
    } else if (db2) {
      const int c = (item - 18) * C + cc;
      atomicAdd(db2 + c, v.x);
      atomicAdd(db2 + c + 1, v.y);
    }
  }
  }
}

// --------------------------- backward, part b ---------------------------------
// du[px][c] = sum_{ky,kx} du2[h-(ky-1), w-(kx-1)][c] * w2[c][ky][kx]; colsum[c] += sum_px du.
__global__ void __launch_bounds__(NWARP * 32)
dwconv_bwd_data_kernel(const __grid_constant__ CUtensorMap tmG, const float* __restrict__ w2, bf16* __restrict__ du,
                       float* __restrict__ colsum, int N, int H, int W, int CH) {
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full[2];
  __shared__ float2 s_cs[NWARP][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = blockIdx.y, c2 = cg * 64 + lane * 2;
  const bool chan_ok = c2 < CH;
  const int c0 = chan_ok ? c2 : 0;
  float2 wt[9], wf[9];
  load_taps(wt, w2, c0);
#pragma unroll
  for (int k = 0; k < 9; ++k) wf[k] = wt[8 - k];  // window offset (r, d) pairs with tap (2-r, 2-d)
  const Tiles T(N, H, W);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmG);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int t, int s) {
    int n, h0, w0;
    T.decode(t, n, h0, w0);
    mbar_arrive_expect_tx(&full[s], BOX_BYTES);
    tma_load_4d(smem + (size_t)s * BOX_BYTES, &tmG, &full[s], cg * 64, w0 - 1, h0 - 1, n);
  };
  if (threadIdx.x == 0 && T.t0 < T.t1) issue(T.t0, 0);
  float2 cs = make_float2(0.f, 0.f);
  int it = 0;
  for (int t = T.t0; t < T.t1; ++t, ++it) {
    const int s = it & 1;
    if (threadIdx.x == 0 && t + 1 < T.t1) issue(t + 1, s ^ 1);
    int n, h0, w0;
    T.decode(t, n, h0, w0);
    mbar_wait(&full[s], (it >> 1) & 1);
    const SBf sA = sbf(smem + (size_t)s * BOX_BYTES) + lane * 2;
    const int h = h0 + warp;
    bf16* orow = du + (((size_t)n * H + h) * W + w0) * CH + c0;
    float2 A[3][3];
#pragma unroll
    for (int x = 0; x < HWD; ++x) {
#pragma unroll
      for (int r = 0; r < 3; ++r) A[r][x % 3] = lds_bf2(sA + ((warp + r) * HWD + x) * 64);
      if (x >= 2) {
        const int ox = x - 2;
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int d = 0; d < 3; ++d) fma2(a, A[r][(x - 2 + d) % 3], wf[r * 3 + d]);
        if (chan_ok && h < H && w0 + ox < W) {
          st_bf2(orow + (size_t)ox * CH, a);
          cs.x += a.x;
          cs.y += a.y;
        }
      }
    }
    __syncthreads();
  }
  s_cs[warp][lane] = cs;
  __syncthreads();
  if (colsum && warp == 0 && chan_ok) {
    float2 v = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < NWARP; ++k) {
      v.x += s_cs[k][lane].x;
      v.y += s_cs[k][lane].y;
    }
    atomicAdd(colsum + c2, v.x);
    atomicAdd(colsum + c2 + 1, v.y);
  }
}

// persistent grid: x = CTAs per channel group, y = channel groups
dim3 pick_grid(int N, int H, int W, int CHpair, int ctas_per_sm) {
  const int groups = ceil_div(CHpair, 64);
  const int tiles = N * ceil_div(H, TH) * ceil_div(W, TW);
  int slots = (dcpt_num_sms() * ctas_per_sm) / groups;
  if (slots < 1) slots = 1;
  if (slots > tiles) slots = tiles;
  return dim3(slots, groups, 1);
}

template <typename K>
int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) DCPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace

int dwgate_fwd_launch(const bf16* u, const float* w2, const float* b2, bf16* g, float* pool, int N, int H, int W, int C,
                      cudaStream_t st) {
  DCPT_CHECK_ARG(C % 8 == 0 && C >= 8 && N > 0 && H > 0 && W > 0, DCPT_E_SHAPE, "dwgate_fwd: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  CUtensorMap tmU;
  DCPT_TRY(make_tmap_nhwc(&tmU, u, N, H, W, 2 * C, HWD, HH));
  const size_t smem = 128 + (size_t)2 * 2 * BOX_BYTES;
  DCPT_TRY(set_smem(dwgate_fwd_kernel<0>, smem));
  DCPT_PROF(dcpt_prof_tag2("dwgate_fwd", (long long)N * H * W, C), 38.0 * N * H * W * C, 6.0 * N * H * W * C, st);
  DCPT_CUDA(dcpt_launch_pdl(dwgate_fwd_kernel<0>, pick_grid(N, H, W, C, 2), dim3(NWARP * 32), smem, st, tmU, w2, b2, g, pool, N, H, W, C));
  DCPT_LAUNCH_CHECK();
  return 0;
}

int dwgelu_fwd_launch(const bf16* u, const float* w2, bf16* g, int N, int H, int W, int C, cudaStream_t st) {
  DCPT_CHECK_ARG(C % 8 == 0 && C >= 8 && N > 0 && H > 0 && W > 0, DCPT_E_SHAPE, "dwgelu_fwd: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  CUtensorMap tmU;
  DCPT_TRY(make_tmap_nhwc(&tmU, u, N, H, W, 2 * C, HWD, HH));
  const size_t smem = 128 + (size_t)2 * 2 * BOX_BYTES;
  DCPT_TRY(set_smem(dwgate_fwd_kernel<1>, smem));
  DCPT_PROF("dwgelu_fwd", 60.0 * N * H * W * C, 6.0 * N * H * W * C, st);
  dwgate_fwd_kernel<1><<<pick_grid(N, H, W, C, 2), NWARP * 32, smem, st>>>(tmU, w2, nullptr, g, nullptr, N, H, W, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int dwconv3_fwd_launch(const bf16* x, const float* w, bf16* out, float* sumsq, int sq_ch, int N, int H, int W, int CH,
                       cudaStream_t st) {
  DCPT_CHECK_ARG(CH % 8 == 0 && CH >= 8 && N > 0 && H > 0 && W > 0, DCPT_E_SHAPE, "dwconv3_fwd: bad shape N=%d H=%d W=%d CH=%d", N, H, W, CH);
  CUtensorMap tmX;
  DCPT_TRY(make_tmap_nhwc(&tmX, x, N, H, W, CH, HWD, HH));
  const size_t smem = 128 + (size_t)2 * BOX_BYTES;
  DCPT_TRY(set_smem(dwconv3_fwd_kernel, smem));
  DCPT_PROF("dwconv3_fwd", 20.0 * N * H * W * CH, 4.0 * N * H * W * CH, st);
  dwconv3_fwd_kernel<<<pick_grid(N, H, W, CH, 4), NWARP * 32, smem, st>>>(tmX, w, out, sumsq, sq_ch, N, H, W, CH);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int dwgate_bwd_a_launch(const bf16* dgs, const float* s, const float* t, const bf16* u, const float* w2, const float* b2,
                        bf16* du2, float* dw2, float* db2, int N, int H, int W, int C, cudaStream_t st, const float* ds,
                        const float* w_sca) {
  DCPT_CHECK_ARG(t != nullptr || (ds != nullptr && w_sca != nullptr && C % 2 == 0), DCPT_E_ARG, "dwgate_bwd_a: need t, or ds + the SCA weight");
  DCPT_CHECK_ARG(C % 8 == 0 && C >= 8 && N > 0 && H > 0 && W > 0, DCPT_E_SHAPE, "dwgate_bwd_a: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  CUtensorMap tmU, tmD;
  DCPT_TRY(make_tmap_nhwc(&tmU, u, N, H, W, 2 * C, HWD, HH));
  DCPT_TRY(make_tmap_nhwc(&tmD, dgs, N, H, W, C, TW, TH));
  const size_t smem = 128 + (size_t)2 * (2 * BOX_BYTES + DG_BYTES) + (size_t)NWARP * 20 * 32 * sizeof(float2);
  if (dw2 == nullptr) {  // data path only (the caller takes dW2 / db2 from dwconv3_wgrad_launch on another stream)
    DCPT_TRY(set_smem(dwgate_bwd_a_kernel<3>, smem));
    DCPT_PROF(dcpt_prof_tag2("dwgate_bwd_a_data", (long long)N * H * W, C), 44.0 * N * H * W * C, 10.0 * N * H * W * C, st);
    DCPT_CUDA(dcpt_launch_pdl(dwgate_bwd_a_kernel<3>, pick_grid(N, H, W, C, 1), dim3(NWARP * 32), smem, st, tmU, tmD, s, t, w2, b2, du2, dw2, db2, N, H, W, C,
                              ds, w_sca, 1.f / (float)(H * W)));
    DCPT_LAUNCH_CHECK();
    return 0;
  }
  DCPT_TRY(set_smem(dwgate_bwd_a_kernel<0>, smem));
  DCPT_PROF(dcpt_prof_tag2("dwgate_bwd_a", (long long)N * H * W, C), 80.0 * N * H * W * C, 10.0 * N * H * W * C, st);
  DCPT_CUDA(dcpt_launch_pdl(dwgate_bwd_a_kernel<0>, pick_grid(N, H, W, C, 1), dim3(NWARP * 32), smem, st, tmU, tmD, s, t, w2, b2, du2, dw2, db2, N, H, W, C, ds, w_sca,
                            1.f / (float)(H * W)));
  DCPT_LAUNCH_CHECK();
  return 0;
}

// Restormer GDFN gate backward: dg bf16 [N,H,W,C], u bf16 [N,H,W,2C] -> du2 bf16 [N,H,W,2C]; dw2[2C][9] += (no bias).
int dwgelu_bwd_a_launch(const bf16* dg, const bf16* u, const float* w2, bf16* du2, float* dw2, int N, int H, int W, int C, cudaStream_t st) {
  DCPT_CHECK_ARG(C % 8 == 0 && C >= 8 && N > 0 && H > 0 && W > 0, DCPT_E_SHAPE, "dwgelu_bwd_a: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  CUtensorMap tmU, tmD;
  DCPT_TRY(make_tmap_nhwc(&tmU, u, N, H, W, 2 * C, HWD, HH));
  DCPT_TRY(make_tmap_nhwc(&tmD, dg, N, H, W, C, TW, TH));
  const size_t smem = 128 + (size_t)2 * (2 * BOX_BYTES + DG_BYTES) + (size_t)NWARP * 20 * 32 * sizeof(float2);
  DCPT_TRY(set_smem(dwgate_bwd_a_kernel<1>, smem));
  DCPT_PROF("dwgelu_bwd_a", 120.0 * N * H * W * C, 10.0 * N * H * W * C, st);
  dwgate_bwd_a_kernel<1><<<pick_grid(N, H, W, C, 1), NWARP * 32, smem, st>>>(tmU, tmD, nullptr, nullptr, w2, nullptr, du2, dw2, nullptr, N, H, W, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}

// Plain depthwise 3x3 weight gradient: dw[c][tap] += sum_px dy[px][c] * x[px + tap][c] for CH = 2 * Chalf channels.
int dwconv3_wgrad_launch(const bf16* dy, const bf16* x, float* dw, int N, int H, int W, int CH, cudaStream_t st, float* db) {
  DCPT_CHECK_ARG(CH % 16 == 0 && CH >= 16 && N > 0 && H > 0 && W > 0, DCPT_E_SHAPE, "dwconv3_wgrad: CH=%d must be a multiple of 16", CH);
  const int C = CH / 2;
  CUtensorMap tmU, tmD;
  DCPT_TRY(make_tmap_nhwc(&tmU, x, N, H, W, CH, HWD, HH));
  DCPT_TRY(make_tmap_nhwc(&tmD, dy, N, H, W, CH, TW, TH));
  const size_t smem = 128 + (size_t)2 * (2 * BOX_BYTES + 2 * DG_BYTES) + (size_t)NWARP * 20 * 32 * sizeof(float2);
  DCPT_TRY(set_smem(dwgate_bwd_a_kernel<2>, smem));
  DCPT_PROF("dwconv3_wgrad", 36.0 * N * H * W * CH, 4.0 * N * H * W * CH, st);
  dwgate_bwd_a_kernel<2><<<pick_grid(N, H, W, C, 1), NWARP * 32, smem, st>>>(tmU, tmD, nullptr, nullptr, dw, nullptr, nullptr, dw, db, N, H, W, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int dwconv_bwd_data_launch(const bf16* du2, const float* w2, bf16* du, float* colsum, int N, int H, int W, int C2,
                           cudaStream_t st) {
  DCPT_CHECK_ARG(C2 % 8 == 0 && C2 >= 8, DCPT_E_SHAPE, "dwconv_bwd_data: bad channel count %d", C2);
  CUtensorMap tmG;
  DCPT_TRY(make_tmap_nhwc(&tmG, du2, N, H, W, C2, HWD, HH));
  const size_t smem = 128 + (size_t)2 * BOX_BYTES;
  DCPT_TRY(set_smem(dwconv_bwd_data_kernel, smem));
  DCPT_PROF(dcpt_prof_tag2("dwconv_bwd_data", (long long)N * H * W, C2), 18.0 * N * H * W * C2, 4.0 * N * H * W * C2, st);
  DCPT_CUDA(dcpt_launch_pdl(dwconv_bwd_data_kernel, pick_grid(N, H, W, C2, 4), dim3(NWARP * 32), smem, st, tmG, w2, du, colsum, N, H, W, C2));
  DCPT_LAUNCH_CHECK();
  return 0;
}
