// GEMM interface shared by the tcgen05 kernel (gemm_sm100.cu) and the CUDA-core
// cross-check kernel (gemm_simt.cu).  All the 1x1 / 2x2-stride-2 convolutions of
// the hot path and their dgrad / wgrad run through this.
//
//   D[M,N] = sum_k A(m,k) * B(n,k)          (fp32 accumulate, bf16 operands)
//
// Operand storage:
//   K-major  (a_mn == 0): A(m,k) = A[m*lda + k]      activations [pixels, channels]
//   MN-major (a_mn == 1): A(m,k) = A[k*lda + m]      used by wgrad, where k = pixel
// likewise for B with (n,k).
#pragma once
#include "common.cuh"

enum GemmEpilogue {
  EPI_STORE = 0,     // out = acc (+bias[n]) (+resid[m,n])  -> fp32 and/or bf16
  EPI_GATE = 1,      // pair-interleaved columns: x4 (bf16, 2C) and SimpleGate sg (bf16, C)
  EPI_GATE_BWD = 2,  // acc = d(sg); reads x4, writes d(x4) (bf16, 2C)
  EPI_PIXSHUF = 3,   // pixel-shuffle(2) scatter + residual add -> fp32 (+bf16 mirror)
  EPI_ATOMIC = 4,    // split-K wgrad: red.add.f32 into out_f32
  EPI_STORE_TMA = 5, // internal: EPI_STORE with the residual tile TMA-loaded and the output tile TMA-stored (gemm_sm100.cu)
  EPI_GATE_BWD_TMA = 6,  // internal: EPI_GATE_BWD with x4 tiles TMA-loaded and d(x4) tiles TMA-stored (C % 32 == 0)
  EPI_GATE_TMA = 7,      // internal: EPI_GATE on 32-wide pair packing (PACK_PAIR32), x4 / sg tiles TMA-stored (C % 32 == 0)
  EPI_LNBWD_TMA = 8,     // internal: EPI_STORE whose accumulator is d(LN output): the LayerNorm backward runs in the epilogue (ep.lnb_*)
};

struct EpiParams {
  float* out_f32;      // nullable (STORE / PIXSHUF / ATOMIC)
  bf16* out_bf16;      // nullable (STORE / PIXSHUF mirror; GATE: x4; GATE_BWD: dx4)
  int ldo;             // row stride (elements) of out_f32 / out_bf16
  const float* bias;   // [N] nullable
  const float* resid;  // fp32 [M, ldr] nullable (STORE); PIXSHUF: same indexing as the output
  int ldr;
  bf16* out2;          // GATE: sg [M, ldo2]
  int ldo2;
  const bf16* aux;     // GATE_BWD: x4 [M, ldaux]
  int ldaux;
  int C;               // GATE / GATE_BWD: half width (x4 has 2C columns)
  int H, W, Cseg;      // PIXSHUF: input spatial dims and channels per output pixel (N = 4*Cseg)
  // Fused column reductions (TMA-tiled epilogues; gemm_tc_launch runs the separate reduction kernel otherwise):
  //   GATE_BWD: colsum[2C] += column sums of d(x4)                          (conv4's bias gradient)
  //   STORE   : with gaux (bf16 [M, ldgaux]) and rows_per_img: colsum[(m / rows_per_img) * N + n] += out[m, n] * gaux[m, n]
  //             (SCA backward: ds[img, c] = sum_px d(g*s) * g, nafnet_arch.py:116-127)
  float* colsum;
  const bf16* gaux;
  int ldgaux;
  int rows_per_img;
  long long out_batch_stride;  // ATOMIC with GemmArgs::k_per_batch: element offset of out_f32 between batches
  // Fused channel LayerNorm of the OUTPUT rows (STORE, TMA-tiled epilogue, N <= 512 so that a CTA holds whole rows in TMEM):
  // with v = acc + bias + resid the row being written to out_f32, also ln_out[m, :] = bf16((v - mean) * rstd * ln_w + ln_b) and
  // ln_stats[m] = (mean, rstd) - the LayerNorm2d forward of the CONSUMER of this tensor (nafnet_arch.py:27-35: norm2 after
  // conv3's residual add, norm1 of the next block after conv5's), which otherwise re-reads the fp32 rows in its own launch.
  const float* ln_w;
  const float* ln_b;
  bf16* ln_out;
  int ld_ln;
  float* ln_stats;
  float ln_eps;
  int ln_nocenter;  // 1: (v * rstd) * ln_w, rstd still from the variance about the mean, ln_b may be NULL (Restormer's BiasFree_LayerNorm, restormer_arch.py:26-40)
  // Fused LayerNorm BACKWARD (STORE, TMA-tiled epilogue, N <= 512; selected by lnb_x != nullptr): the accumulator is
  // dn = d(loss)/d(LN output) (a dgrad GEMM: conv4's or conv1's), and the epilogue turns it into the gradient of the LN INPUT
  // with the reference's hand-written backward (nafnet_arch.py:38-53):  g = dn * w,  xhat = (x - mean) * rstd,
  //   dx = (g - xhat * mean_c(g * xhat) - mean_c(g)) * rstd (+ dres)   -> out_f32 (+ out_bf16 mirror),
  //   lnb_dw[c] += sum_m dn * xhat,  lnb_db[c] += sum_m dn,  lnb_cs[c] += sum_m dx  (nullable)
  // - what ln_bwd_launch computes from a materialised bf16 dn, without the dn round trip and the second launch.
  const float* lnb_x;      // fp32 [M, ld_lnb]: the LayerNorm's input rows
  int ld_lnb;
  const float* lnb_stats;  // [M, 2] (mean, rstd) saved by the forward
  const float* lnb_w;      // LayerNorm weight [N]
  const float* lnb_dres;   // fp32 [M, ld_lnb] gradient of the residual branch, added to dx (nullable)
  float* lnb_dw;
  float* lnb_db;
  float* lnb_cs;
};
// true when gemm_tc_launch can fuse the LayerNorm of the output rows for this N (else the caller runs ln_fwd_launch)
inline bool gemm_ln_fusable(int N) { return N >= 8 && N % 8 == 0 && N <= 512; }

struct GemmArgs {
  int M, N, K;
  const bf16* A;
  int lda;
  int a_mn;
  const bf16* B;
  int ldb;
  int b_mn;
  int splits;  // >1 only with EPI_ATOMIC
  int epi;
  EpiParams ep;
  // Batched forms (Restormer MDTA, one problem per image in ONE launch):
  //  * K-major STORE: B is a stack of per-batch [N, K] matrices; the M tile starting at row m reads the matrix of batch
  //    m / m_per_batch, i.e. B rows offset by (m / m_per_batch) * b_rows_per_batch (m_per_batch % 128 == 0).
  //  * MN-major ATOMIC: the K axis is a concatenation of batches of k_per_batch (multiple of 64) contraction indices; every
  //    split stays inside one batch and accumulates into out_f32 + batch * ep.out_batch_stride.  `splits` = splits per batch.
  int m_per_batch, b_rows_per_batch;
  int k_per_batch;
  // EPI_ATOMIC only: launch one CTA per (tile, split) work item instead of a persistent grid of <= #SM CTAs looping over them.
  // Short-lived CTAs give thread-block slots back sooner when this GEMM runs on a low-priority stream beside a critical chain.
  int fine_grid;
};

// Geometry of the implicit-GEMM 3x3 convolution modes of the tensor-core kernel (gemm_sm100.cu).
struct ConvGeom {
  int H, W;
  int tw_log2;  // pixel patch = th x (1 << tw_log2)
  int th;
  int tiles_w, tiles_h;
  int cblocks;  // ceil(Cin / 64)
};

// Dense 3x3 convolution, stride 1, zero pad 1, NHWC bf16 (reference: BottleneckBlock.conv2,
// archs/degrad_classify_arch.py:178-188): out[px][co] = sum_{tap,ci} X[px + tap][ci] * Wp[co][tap * cin_pad + ci],
// cin_pad = Cin rounded up to 64, tap = ky*3 + kx.  The epilogue is EPI_STORE (ep.out_* indexed by linear pixel).
struct Conv3x3Args {
  const bf16* X;
  int N, H, W, Cin;
  const bf16* Wp;
  int Cout;
  EpiParams ep;
};
int conv3x3_tc_launch(const Conv3x3Args& a, cudaStream_t stream);
// G[co][tap * cin_pad + ci] += sum_px dY[px][co] * X[px + tap][ci]   (fp32, split-K over pixel patches)
int conv3x3_wgrad_tc_launch(const bf16* dY, const bf16* X, float* G, int N, int H, int W, int Cin, int Cout, cudaStream_t stream);

// Split-K factor for a wgrad-style GEMM with `tiles` output tiles and `num_kb` 64-wide k-blocks: the largest split
// whose tile count still fits ONE wave of the persistent grid (tiles * splits <= SMs).  Rounding up instead costs a
// whole second wave for a handful of tiles (measured: 152 tiles on 148 SMs ran 1.9x slower than 144).
inline int gemm_auto_splits(int tiles, int num_kb) {
  int splits = dcpt_num_sms() / (tiles > 0 ? tiles : 1);
  if (splits > num_kb) splits = num_kb;
  if (splits < 1) splits = 1;
  return splits;
}

inline GemmArgs make_gemm_args(int M, int N, int K, const bf16* A, int lda, const bf16* B, int ldb, int epi) {
  GemmArgs g = {};
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.splits = 1; g.epi = epi;
  return g;
}

int gemm_tc_launch(const GemmArgs& g, cudaStream_t stream);    // tcgen05 + TMA (product path)
int gemm_simt_launch(const GemmArgs& g, cudaStream_t stream);  // CUDA-core cross-check (tests only)
int gemm_launch(const GemmArgs& g, cudaStream_t stream);       // dispatch (DCPT_GEMM_SIMT=1 selects simt)

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Fused epilogues.  The tensor-core kernel transposes each 32x32 accumulator chunk through shared
// memory, so a lane owns 4 CONSECUTIVE columns of one row and 8 lanes cover 128 contiguous bytes
// of fp32 output: every global access below is coalesced.  All N on this path are multiples of 8.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  op16x2 p0 = OP2_FROM_F32(a, b), p1 = OP2_FROM_F32(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&p0);
  u.y = *reinterpret_cast<uint32_t*>(&p1);
  return u;
}
__device__ __forceinline__ float4 unpack4_bf16(uint2 u) {
  const float2 a = OP2_TO_F32(*reinterpret_cast<const op16x2*>(&u.x));
  const float2 b = OP2_TO_F32(*reinterpret_cast<const op16x2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// The epilogue of a 4-column piece is split into a LOAD phase (everything read from global memory) and a
// STORE phase, so the caller can issue all loads of a chunk before the first store: the output pointers are
// not provably distinct from the inputs, and a load placed after a store would wait for it (measured: 2x).
// The loads are further split into a per-COLUMN part (bias: the same for every row a lane handles) and a per-ROW
// part (residual / gate operands), so the tensor-core kernel can keep the row part of the NEXT chunk in flight
// while it finishes the current one.
struct EpiCol {
  float4 a, b;
};
struct EpiRow {
  float4 r;     // STORE / PIXSHUF: residual
  uint2 xa, xb;  // GATE_BWD: the two x4 halves (raw bf16)
};
struct EpiExtra {
  EpiCol c;
  EpiRow r;
};

// piece = 4 consecutive columns [n, n+4).  For EPI_GATE, n is the packed column `na` of the a-piece (see
// epilogue_store); its b-piece is 8 columns further.
template <int EPI>
__device__ __forceinline__ EpiCol epilogue_load_col(const EpiParams& p, int n) {
  EpiCol e;
  e.a = e.b = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (EPI == EPI_STORE) {
    if (p.bias) e.a = __ldg(reinterpret_cast<const float4*>(p.bias + n));
  } else if constexpr (EPI == EPI_GATE) {
    e.a = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    e.b = __ldg(reinterpret_cast<const float4*>(p.bias + n + 8));
  }
  return e;
}
template <int EPI>
__device__ __forceinline__ EpiRow epilogue_load_row(const EpiParams& p, int m, int n) {
  EpiRow e;
  e.r = make_float4(0.f, 0.f, 0.f, 0.f);
  e.xa = e.xb = make_uint2(0u, 0u);
  if constexpr (EPI == EPI_STORE) {
    if (p.resid) e.r = __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)m * p.ldr + n));
  } else if constexpr (EPI == EPI_GATE_BWD) {
    const bf16* x4 = p.aux + (size_t)m * p.ldaux;
    e.xa = __ldg(reinterpret_cast<const uint2*>(x4 + n));
    e.xb = __ldg(reinterpret_cast<const uint2*>(x4 + p.C + n));
  } else if constexpr (EPI == EPI_PIXSHUF) {
    if (p.resid) {
      const int w = m % p.W, t = m / p.W, h = t % p.H, img = t / p.H;
      const int q = n / p.Cseg, c = n - q * p.Cseg;
      const size_t pix = ((size_t)img * (2 * p.H) + 2 * h + (q >> 1)) * (size_t)(2 * p.W) + 2 * w + (q & 1);
      e.r = __ldg(reinterpret_cast<const float4*>(p.resid + pix * p.Cseg + c));
    }
  }
  return e;
}
template <int EPI>
__device__ __forceinline__ EpiExtra epilogue_load(const EpiParams& p, int m, int n) {
  EpiExtra e;
  e.c = epilogue_load_col<EPI>(p, n);
  e.r = epilogue_load_row<EPI>(p, m, n);
  return e;
}

// v = the 4 accumulators of the piece; for EPI_GATE, v = a-piece and vb = b-piece accumulators.
template <int EPI>
__device__ __forceinline__ void epilogue_store(const EpiParams& p, int m, int n, float4 v, float4 vb, const EpiExtra& e) {
  if constexpr (EPI == EPI_STORE) {
    v.x += e.c.a.x + e.r.r.x; v.y += e.c.a.y + e.r.r.y; v.z += e.c.a.z + e.r.r.z; v.w += e.c.a.w + e.r.r.w;
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (size_t)m * p.ldo + n) = v;
    if (p.out_bf16) *reinterpret_cast<uint2*>(p.out_bf16 + (size_t)m * p.ldo + n) = pack4_bf16(v.x, v.y, v.z, v.w);
  } else if constexpr (EPI == EPI_GATE) {
    // packed column p = 16*pp + h*8 + i  <->  channel h*C + 8*pp + i  (h = 0 first half, 1 second half);
    // (n % 16) in {0, 4}: a-piece at n, b-piece at n + 8.
    float4 a = v, b = vb;
    a.x = bf16_round(a.x + e.c.a.x); a.y = bf16_round(a.y + e.c.a.y); a.z = bf16_round(a.z + e.c.a.z); a.w = bf16_round(a.w + e.c.a.w);
    b.x = bf16_round(b.x + e.c.b.x); b.y = bf16_round(b.y + e.c.b.y); b.z = bf16_round(b.z + e.c.b.z); b.w = bf16_round(b.w + e.c.b.w);
    const int j = (n >> 4) * 8 + (n & 15);
    bf16* x4 = p.out_bf16 + (size_t)m * p.ldo;
    *reinterpret_cast<uint2*>(x4 + j) = pack4_bf16(a.x, a.y, a.z, a.w);
    *reinterpret_cast<uint2*>(x4 + p.C + j) = pack4_bf16(b.x, b.y, b.z, b.w);
    *reinterpret_cast<uint2*>(p.out2 + (size_t)m * p.ldo2 + j) = pack4_bf16(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
  } else if constexpr (EPI == EPI_GATE_BWD) {
    bf16* o = p.out_bf16 + (size_t)m * p.ldo;
    const float4 xa = unpack4_bf16(e.r.xa), xb = unpack4_bf16(e.r.xb);
    *reinterpret_cast<uint2*>(o + n) = pack4_bf16(v.x * xb.x, v.y * xb.y, v.z * xb.z, v.w * xb.w);
    *reinterpret_cast<uint2*>(o + p.C + n) = pack4_bf16(v.x * xa.x, v.y * xa.y, v.z * xa.z, v.w * xa.w);
  } else if constexpr (EPI == EPI_PIXSHUF) {
    const int w = m % p.W, t = m / p.W, h = t % p.H, img = t / p.H;
    const int q = n / p.Cseg, c = n - q * p.Cseg;
    const size_t pix = ((size_t)img * (2 * p.H) + 2 * h + (q >> 1)) * (size_t)(2 * p.W) + 2 * w + (q & 1);
    const size_t off = pix * p.Cseg + c;
    v.x += e.r.r.x; v.y += e.r.r.y; v.z += e.r.r.z; v.w += e.r.r.w;
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + off) = v;
    if (p.out_bf16) *reinterpret_cast<uint2*>(p.out_bf16 + off) = pack4_bf16(v.x, v.y, v.z, v.w);
  } else if constexpr (EPI == EPI_ATOMIC) {
    float* o = p.out_f32 + (size_t)m * p.ldo + n;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

// Row-per-thread form (CUDA-core cross-check kernel): 32 consecutive accumulator columns of row m.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& p, int m, int n0, int N, const float (&acc)[32]) {
  if constexpr (EPI == EPI_GATE) {
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int na = n0 + g * 16 + q * 4, o = g * 16 + q * 4;
        if (na < N)
          epilogue_store<EPI>(p, m, na, make_float4(acc[o], acc[o + 1], acc[o + 2], acc[o + 3]),
                              make_float4(acc[o + 8], acc[o + 9], acc[o + 10], acc[o + 11]), epilogue_load<EPI>(p, m, na));
      }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + j * 4;
      if (n < N)
        epilogue_store<EPI>(p, m, n, make_float4(acc[j * 4], acc[j * 4 + 1], acc[j * 4 + 2], acc[j * 4 + 3]),
                            make_float4(0.f, 0.f, 0.f, 0.f), epilogue_load<EPI>(p, m, n));
    }
  }
}
#endif
