// tcgen05 / TMEM / TMA GEMM for sm_100a — the tensor-core engine behind every
// pointwise (1x1) convolution, the 2x2-stride-2 down convs, the 1x1+PixelShuffle
// up convs and all their dgrad / wgrad contractions on the NAFNet hot path
// (reference: nafnet_arch.py:87-150 conv1/3/4/5, :230 downs, :238-242 ups; the
// reference runs these through cuDNN/cuBLAS fp32).
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1   : MMA issuer     (one elected lane issues tcgen05.mma, fp32 accum in TMEM)
//   warps 2-9: epilogue       (tcgen05.ld TMEM -> registers -> smem transpose -> fused epilogue -> coalesced global I/O)
// Two TMEM accumulator buffers let the epilogue of tile i overlap the main loop
// of tile i+1.  Tile = 128 x BN (BN in {64,128,256}), BK = 64 bf16 (= one 128-byte
// swizzle row).  MN-major operands (wgrad) are loaded as 64x64 boxes, which
// the UMMA descriptor addresses with LBO = 8 KiB / SBO = 1 KiB.
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "elementwise.cuh"
#include "gemm.cuh"

#ifdef DCPT_TRACE
// Debug build only (python -m dcpt_b200.build --trace -> libdcpt_sm100_trace.so): per-CTA event timeline of the GEMM kernel,
// 16 slots x (globaltimer ns, clock64) per CTA, written by one thread per event (tools/gemm_trace.py reads it).
__device__ unsigned long long* g_dcpt_trace = nullptr;
extern "C" int dcpt_debug_set_trace(void* buf) {
  return (int)cudaMemcpyToSymbol(g_dcpt_trace, &buf, sizeof(buf));
}
#define TRACE(slot)                                                       \
  do {                                                                    \
    unsigned long long* _t = g_dcpt_trace;                                \
    if (_t) {                                                             \
      _t[(blockIdx.x * 16 + (slot)) * 2] = globaltimer_ns();              \
      _t[(blockIdx.x * 16 + (slot)) * 2 + 1] = (unsigned long long)clock64(); \
    }                                                                     \
  } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

namespace {

// batched problems in one launch (GemmArgs::m_per_batch / k_per_batch); all zero = plain GEMM
struct BatchGeom {
  int m_per_batch, b_rows_per_batch;  // K-major: B row offset per batch of M rows
  int kb_per_batch, splits_per_batch; // MN-major split-K: k-blocks and splits per batch
};

struct EpiMaps {
  // EPI_STORE_TMA: fp32 output, bf16 output, fp32 residual.  EPI_GATE_BWD_TMA: o16 = d(x4), r32 = x4 (bf16 [M, 2C]).
  // EPI_GATE_TMA: o16 = x4 (bf16 [M, 2C]), o2 = sg (bf16 [M, C]).  Unused by the register-staged epilogues.
  CUtensorMap o32, o16, r32, o2;
  // bf16 outputs written as 64-column boxes (128-byte rows, 128B swizzle).  A TMA store costs ~5 cycles per box ROW whatever
  // its width (r02c/r02d traces: 2 KiB boxes of 64-byte rows drained at ~11 B/clk per SM), so narrow rows halve the store rate.
  CUtensorMap o16w;  // EPI_STORE_TMA, bf16-only output
  CUtensorMap ln;    // EPI_STORE_TMA with a fused LayerNorm of the output rows: normalised bf16 output [M, N]
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;  // 16 KiB

// EPI_STORE_TMA keeps two 4 KiB epilogue tiles per epilogue warp (residual tile in, output tile out, in place), so it
// runs with one pipeline stage fewer than the register-staged epilogues (one 4 KiB transpose tile per warp).
template <int BN, int EPI>
struct Cfg {
  static constexpr bool TMA_EPI = EPI == EPI_STORE_TMA || EPI == EPI_GATE_BWD_TMA || EPI == EPI_GATE_TMA || EPI == EPI_LNBWD_TMA;
  static constexpr int STAGES = TMA_EPI ? (BN == 256 ? 3 : (BN == 128 ? 4 : 6)) : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr uint32_t B_STAGE_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulator buffers (power of two >= 32)
  static constexpr uint32_t EPI_BYTES = 8 * (TMA_EPI ? 8192 : 4096);  // starts 1024-byte aligned (STAGE_BYTES % 1024 == 0)
  // EPI_GATE_BWD_TMA: per-CTA column sums of d(x4) (2C <= 2048 floats); EPI_STORE_TMA: row-statistics exchange of the fused
  // LayerNorm (2 slab parities x 8 warps x 32 lanes x float4)
  // EPI_LNBWD_TMA: the same exchange buffer + three per-CTA column accumulators of N <= 512 floats (dw, db, column sums of dx)
  static constexpr uint32_t COLSUM_BYTES = EPI == EPI_GATE_BWD_TMA ? 8192 : (EPI == EPI_STORE_TMA ? 8192 : (EPI == EPI_LNBWD_TMA ? 8192 + 6144 : 0));
  static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 512 /*barriers*/ + COLSUM_BYTES;
};

// UMMA shared-memory descriptor (sm_100): start addr [0,14), LBO [16,30), SBO [32,46) (all >>4),
// version=1 at [46,48), layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), a/b major (15/16),
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (DCPT_UMMA_FMT << 7) | (DCPT_UMMA_FMT << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Column sums of a 32 x 32 tile held one row per lane: 31 shuffles; afterwards v[0] of lane l = sum over the 32 rows of column l.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// CONV = 0: plain GEMM.  CONV = 1: implicit-GEMM 3x3 convolution (pad 1, stride 1), A = NHWC activation through a 4-D
// tensor map: the M tile is a TH x TW pixel patch of one image and k-block kb = (tap, 64-channel block) loads the
// patch shifted by the tap, zero-filled outside the image.  CONV = 2: its wgrad, contracting over 64-pixel patches
// (A = dY patch, B = X patch shifted by the tap of this N tile), MN-major operands.
template <int BN, int EPI, bool A_MN, bool B_MN, int CONV>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ EpiMaps em,
               int M, int N, int K, int tiles_m, int tiles_n, int splits, int kb_per_split, EpiParams ep, ConvGeom cg, BatchGeom bg,
               int pf_mode) {
  using C = Cfg<BN, EPI>;
  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 0) TRACE(0);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)C::STAGES * A_STAGE_BYTES;
  uint8_t* sEpi = smem + (size_t)C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + C::EPI_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* rfull = tempty + 2;  // [8 epilogue warps][2]: residual tile landed (EPI_STORE_TMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 16);
  float* s_colsum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  if constexpr (EPI == EPI_GATE_BWD_TMA) {
    if (ep.colsum)
      for (int i = threadIdx.x; i < 2 * ep.C; i += kThreads) s_colsum[i] = 0.f;
  }
  if constexpr (EPI == EPI_LNBWD_TMA) {
    for (int i = threadIdx.x; i < 1536; i += kThreads) s_colsum[2048 + i] = 0.f;
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (K + BK - 1) / BK;
  const int total_tiles = tiles_m * tiles_n * splits;
  // Tile schedule: tile t = (split, m_t, n_t) round-robin over the CTAs; with a fused LayerNorm over rows wider than one tile
  // (N = 512) a CTA takes whole 128-row SLABS: its tiles_n (= 2) tiles of a slab back to back, one per accumulator buffer.
  bool slab_mode = false;
  if constexpr (EPI == EPI_STORE_TMA) slab_mode = ep.ln_out != nullptr && tiles_n > 1;
  if constexpr (EPI == EPI_LNBWD_TMA) slab_mode = tiles_n > 1;
  auto tile_of = [&](int it) -> int {
    if (slab_mode) {
      const int slab = (int)blockIdx.x + (it / tiles_n) * (int)gridDim.x;
      return slab < tiles_m ? slab * tiles_n + it % tiles_n : -1;
    }
    const int t = (int)blockIdx.x + it * (int)gridDim.x;
    return t < total_tiles ? t : -1;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 8);  // one arrive per epilogue warp
    }
    for (int b = 0; b < 16; ++b) mbar_init(&rfull[b], 1);
    if constexpr (C::TMA_EPI) {
      tma_prefetch_desc(&em.o32);
      tma_prefetch_desc(&em.o16);
      tma_prefetch_desc(&em.r32);
      tma_prefetch_desc(&em.o2);
      if constexpr (EPI == EPI_STORE_TMA) {
        tma_prefetch_desc(&em.ln);
        tma_prefetch_desc(&em.o16w);
      }
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the previous kernel's tail
  pdl_sync();
  if (threadIdx.x == 0) TRACE(1);

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      constexpr int PF_DIST = 12;
      auto prefetch_kb = [&](int pk, int m_t, int n_t) {
        if constexpr (!A_MN) {
          tma_prefetch_2d(&tmA, pk * BK, m_t * BM);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) tma_prefetch_2d(&tmA, m_t * BM + c * 64, pk * BK);
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_prefetch_2d(&tmB, n_t * BN + c * 64, pk * BK);
        }
      };
      for (int it = 0, tile; (tile = tile_of(it)) >= 0; ++it) {
        const int n_t = tile % tiles_n;
        const int m_t = (tile / tiles_n) % tiles_m;
        const int sp = tile / (tiles_n * tiles_m);
        int kb0 = sp * kb_per_split;
        int kb1 = min(num_kb, kb0 + kb_per_split);
        if (bg.kb_per_batch > 0) {  // splits never straddle a batch
          const int bi = sp / bg.splits_per_batch, ls = sp - bi * bg.splits_per_batch;
          kb0 = bi * bg.kb_per_batch + ls * kb_per_split;
          kb1 = min((bi + 1) * bg.kb_per_batch, kb0 + kb_per_split);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          // L2 prefetch of the HBM-streamed operand(s), PF_DIST k-blocks ahead of the smem ring (which can hold
          // only STAGES blocks in flight: not enough to cover DRAM latency for a short-K tile).
          if constexpr (CONV == 0) {
            if (pf_mode == 1) {
              if (kb == kb0) {
                for (int pk = kb0; pk < min(kb1, kb0 + PF_DIST); ++pk) prefetch_kb(pk, m_t, n_t);
              } else if (kb + PF_DIST - 1 < kb1) {
                prefetch_kb(kb + PF_DIST - 1, m_t, n_t);
              }
            }
          }
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
          uint8_t* a_dst = sA + (size_t)stage * A_STAGE_BYTES;
          uint8_t* b_dst = sB + (size_t)stage * C::B_STAGE_BYTES;
          if constexpr (CONV == 1) {
            const int per_img = cg.tiles_h * cg.tiles_w;
            const int img = m_t / per_img, rem = m_t - img * per_img;
            const int h0 = (rem / cg.tiles_w) * cg.th, w0 = (rem % cg.tiles_w) << cg.tw_log2;
            const int tap = kb / cg.cblocks, cb = kb - tap * cg.cblocks;
            tma_load_4d(a_dst, &tmA, &full[stage], cb * 64, w0 + tap % 3 - 1, h0 + tap / 3 - 1, img);
            tma_load_2d(b_dst, &tmB, &full[stage], kb * BK, n_t * BN);
          } else if constexpr (CONV == 2) {
            const int per_img = cg.tiles_h * cg.tiles_w;
            const int img = kb / per_img, rem = kb - img * per_img;
            const int h0 = (rem / cg.tiles_w) * cg.th, w0 = (rem % cg.tiles_w) << cg.tw_log2;
            const int cin_pad = cg.cblocks * 64;
            const int tap = (n_t * BN) / cin_pad, ci0 = n_t * BN - tap * cin_pad;
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_4d(a_dst + c * 8192, &tmA, &full[stage], m_t * BM + c * 64, w0, h0, img);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_4d(b_dst + c * 8192, &tmB, &full[stage], ci0 + c * 64, w0 + tap % 3 - 1, h0 + tap / 3 - 1, img);
          } else {
            if constexpr (!A_MN) {
              tma_load_2d(a_dst, &tmA, &full[stage], kb * BK, m_t * BM);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d(a_dst + c * 8192, &tmA, &full[stage], m_t * BM + c * 64, kb * BK);
            }
            if constexpr (!B_MN) {
              const int brow = bg.m_per_batch > 0 ? (m_t * BM / bg.m_per_batch) * bg.b_rows_per_batch : 0;
              tma_load_2d(b_dst, &tmB, &full[stage], kb * BK, brow + n_t * BN);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c) tma_load_2d(b_dst + c * 8192, &tmB, &full[stage], n_t * BN + c * 64, kb * BK);
            }
          }
          if constexpr (CONV == 0) {
            if (pf_mode == 2) {  // prefetch behind the real loads: the first operands of a tile never queue behind a prefetch burst
              if (kb == kb0) {
                for (int pk = kb0 + C::STAGES; pk < min(kb1, kb0 + PF_DIST); ++pk) prefetch_kb(pk, m_t, n_t);
              } else if (kb + PF_DIST - 1 < kb1 && kb + PF_DIST - 1 >= kb0 + C::STAGES) {
                prefetch_kb(kb + PF_DIST - 1, m_t, n_t);
              }
            }
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = 0, tile; (tile = tile_of(it)) >= 0; ++it) {
        const int sp = tile / (tiles_n * tiles_m);
        int kb0 = sp * kb_per_split;
        int kb1 = min(num_kb, kb0 + kb_per_split);
        if (bg.kb_per_batch > 0) {  // splits never straddle a batch
          const int bi = sp / bg.splits_per_batch, ls = sp - bi * bg.splits_per_batch;
          kb0 = bi * bg.kb_per_batch + ls * kb_per_split;
          kb1 = min((bi + 1) * bg.kb_per_batch, kb0 + kb_per_split);
        }
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          if (it == 0 && kb == kb0) TRACE(2);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)stage * C::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 elements = 32 B inside the 128 B swizzle row; SBO = 1 KiB between 8-row groups.
            // MN-major: 16 k-rows = 2 KiB; LBO = 8 KiB between 64-wide MN chunks, SBO = 1 KiB between 8-k groups.
            const uint64_t adesc = A_MN ? make_smem_desc(a_addr + k * 2048, 8192, 1024) : make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc(b_addr + k * 2048, 8192, 1024) : make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);  // accumulator ready for the epilogue
        TRACE(it == 0 ? 3 : 4);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // -------------------------------- epilogue --------------------------------
    // Eight epilogue warps: warp w may touch TMEM lanes [32*(w%4), +32); the two warps of a lane quarter take
    // alternate 32-column chunks.  Each chunk is transposed through a private 4 KiB smem tile (float4 index
    // XOR-swizzled by row & 7: conflict-free both ways) so that a lane owns 4 consecutive columns of one row and
    // 8 lanes cover 128 contiguous bytes: all global traffic of the fused epilogues is coalesced.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int chalf = ew >> 2;
    const uint32_t stage = smem_u32(sEpi) + (uint32_t)ew * 4096u;  // float4 index * 16
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (C::TMA_EPI) {
      // ---- TMA epilogue (EPI_STORE_TMA) ----
      // Per warp and 32 x 32 chunk: the fp32 residual tile is TMA-loaded into a 128B-swizzled 4 KiB buffer (requested one
      // chunk ahead), each lane adds its accumulator row (tcgen05.ld gives lane = row) and the bias in place, and ONE
      // bulk tensor store writes the tile back (clipped at the matrix edge): no per-element address arithmetic, bounds
      // checks or uncoalesced accesses in the warp, all global traffic is asynchronous.  bf16 outputs are packed into a
      // 64B-swizzled 2 KiB tile of the same buffer.
      uint8_t* ebuf = sEpi + (size_t)ew * 8192;
      uint64_t* rf = rfull + ew * 2;
      uint32_t rphase = 0;  // bit b: parity of the next completion of rf[b]
      int ci = 0;           // chunks handled by this warp so far (buffer = ci & 1)
      const bool has_g = EPI == EPI_STORE_TMA && ep.gaux != nullptr;  // fused SCA-backward reduction (bf16 output, no residual)
      const bool has_r = EPI == EPI_GATE_BWD_TMA || (EPI == EPI_STORE_TMA && (ep.resid != nullptr || has_g));
      // the epilogue's input tile(s) of chunk n0 of row slab m0 -> 4 KiB buffer: one fp32 residual box, or the two bf16
      // x4 boxes (columns n0 and C + n0) of the SimpleGate backward
      auto issue_in = [&](uint8_t* dst, uint64_t* bar, int n0, int m0) {
        if constexpr (EPI == EPI_GATE_BWD_TMA) {
          mbar_arrive_expect_tx(bar, 4096);
          tma_load_2d(dst, &em.r32, bar, n0, m0);
          tma_load_2d(dst + 2048, &em.r32, bar, ep.C + n0, m0);
        } else if (has_g) {  // bf16 g tile into the upper half of the buffer (the bf16 output tile uses the lower half)
          mbar_arrive_expect_tx(bar, 2048);
          tma_load_2d(dst + 2048, &em.o2, bar, n0, m0);
        } else if (ep.Cseg > 0) {  // PixelShuffle scatter: residual (skip) tile of sub-pixel (i, j) = n0 / Cseg
          const int qq = n0 / ep.Cseg;
          mbar_arrive_expect_tx(bar, 4096);
          tma_load_5d(dst, &em.r32, bar, n0 - qq * ep.Cseg, qq & 1, m0 % ep.W, qq >> 1, m0 / ep.W);
        } else {
          mbar_arrive_expect_tx(bar, 4096);
          tma_load_2d(dst, &em.r32, bar, n0, m0);
        }
      };
      // output tile of chunk (n0, m0): plain [M, N] box, or (ep.Cseg > 0) the strided box of the pixel-shuffled NHWC output
      auto store_out = [&](const CUtensorMap* tmap, const void* src, int n0, int m0) {
        if (ep.Cseg > 0) {
          const int qq = n0 / ep.Cseg;
          tma_store_5d(tmap, src, n0 - qq * ep.Cseg, qq & 1, m0 % ep.W, qq >> 1, m0 / ep.W);
        } else {
          tma_store_2d(tmap, src, n0, m0);
        }
      };
      if constexpr (EPI == EPI_LNBWD_TMA) {
        // ---- LayerNorm backward in the epilogue of the dgrad GEMM that produces d(LN output) (EpiParams::lnb_*) ----
        // The CTA holds whole rows (one tile for N <= 256, the two accumulator buffers of a slab for N = 512).  Pass 1, per
        // 32 x 32 chunk: x tile TMA-loaded (one chunk ahead), g = dn * w, xhat = (x - mean) * rstd, per-row sums of g and
        // g * xhat (lane = row: thread-local), column sums dn * xhat / dn (shuffle transpose-reduce -> per-CTA shared
        // accumulators), and (g, xhat) packed as two 16-bit values back into the accumulator's own TMEM column.  The two
        // warps of a lane quarter exchange their row sums; pass 2: dx = (g - xhat * c1 - c2) * rstd + dres (dres tile
        // TMA-loaded in place, result tile TMA-stored), bf16 mirror, column sums of dx.
        float* s_dw = s_colsum + 2048;
        float* s_db = s_dw + 512;
        float* s_cs = s_db + 512;
        const bool has_dres = ep.lnb_dres != nullptr;
        const float invN = 1.f / (float)N;
        float r_g = 0.f, r_gx = 0.f;
        int par = 0;
        for (int it = 0, tile; (tile = tile_of(it)) >= 0; ++it) {
          const int n_t = tile % tiles_n;
          const int m_t = (tile / tiles_n) % tiles_m;
          const int m0 = m_t * BM + q * 32;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
          float mean = 0.f, rstd = 0.f;
          if (m0 + lane < M) {
            const float2 st = __ldg(reinterpret_cast<const float2*>(ep.lnb_stats + (size_t)(m0 + lane) * 2));
            mean = st.x;
            rstd = st.y;
          }
          bool first = true;
#pragma unroll 1
          for (int c = chalf; c < BN / 32; c += 2) {
            const int n0 = n_t * BN + c * 32;
            if (n0 >= N) break;  // warp-uniform
            const int b = ci & 1;
            if (first) {
              if (lane == 0) {
                bulk_wait_read<0>();
                mbar_arrive_expect_tx(&rf[b], 4096);
                tma_load_2d(ebuf + b * 4096, &em.r32, &rf[b], n0, m0);
              }
              mbar_wait(&tfull[acc], acc_phase);
              tc_fence_after();
              first = false;
            }
            float v[32];
            tmem_ld32(taddr + c * 32, v);
            if (lane == 0) {
              bulk_wait_read<0>();
              const int nn = n0 + 64;
              if (c + 2 < BN / 32 && nn < N) {
                mbar_arrive_expect_tx(&rf[b ^ 1], 4096);
                tma_load_2d(ebuf + (b ^ 1) * 4096, &em.r32, &rf[b ^ 1], nn, m0);
              }
            }
            float4 wv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              wv[j] = (n0 + 4 * j < N) ? __ldg(reinterpret_cast<const float4*>(ep.lnb_w + n0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            const uint32_t buf = smem_u32(ebuf + b * 4096);
            mbar_wait(&rf[b], (rphase >> b) & 1u);
            rphase ^= 1u << b;
            float4 xr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) xr[j] = lds_f4(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)));
            tmem_ld_wait();
            float pa[32];
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float xs[4] = {xr[j].x, xr[j].y, xr[j].z, xr[j].w};
              const float ws[4] = {wv[j].x, wv[j].y, wv[j].z, wv[j].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float xh = (xs[k] - mean) * rstd;
                const float dn = v[4 * j + k];
                const float g = dn * ws[k];   // columns beyond N: dn = 0 (zero-filled weights rows), w = 0
                r_g += g;
                r_gx = fmaf(g, xh, r_gx);
                pa[4 * j + k] = dn * xh;
                const op16x2 pr = OP2_FROM_F32(g, xh);
                pk[4 * j + k] = *reinterpret_cast<const uint32_t*>(&pr);
              }
            }
            tmem_st32(taddr + c * 32, *reinterpret_cast<float(*)[32]>(pk));
            const float sa = warp_colsum32(pa, lane);
            const float sb = warp_colsum32(v, lane);
            if (n0 + lane < N) {
              atomicAdd(&s_dw[n0 + lane], sa);
              atomicAdd(&s_db[n0 + lane], sb);
            }
            ++ci;
          }
          if (first) {
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
          }
          if (n_t == tiles_n - 1) {
            tmem_st_wait();
            float4* xch = reinterpret_cast<float4*>(s_colsum) + par * 256;
            xch[ew * 32 + lane] = make_float4(r_g, r_gx, 0.f, 0.f);
            tc_fence_before();
            named_bar_sync(1 + q, 64);
            tc_fence_after();
            const float4 pb = xch[(ew ^ 4) * 32 + lane];
            par ^= 1;
            const float c2 = (r_g + pb.x) * invN;    // mean_c(g)
            const float c1 = (r_gx + pb.y) * invN;   // mean_c(g * xhat)
            r_g = r_gx = 0.f;
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            for (int h = 0; h < tiles_n; ++h) {
              const int hacc = tiles_n > 1 ? h : acc;
              const uint32_t th = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hacc * BN);
              bool first2 = true;
#pragma unroll 1
              for (int c = chalf; c < BN / 32; c += 2) {
                const int n0 = h * BN + c * 32;
                if (n0 >= N) break;
                const int b = ci & 1;
                if (first2) {
                  if (lane == 0) {
                    bulk_wait_read<0>();
                    if (has_dres) {
                      mbar_arrive_expect_tx(&rf[b], 4096);
                      tma_load_2d(ebuf + b * 4096, &em.o2, &rf[b], n0, m0);
                    }
                  }
                  first2 = false;
                }
                float pkf[32];
                tmem_ld32(th + c * 32, pkf);
                if (lane == 0) {
                  bulk_wait_read<0>();
                  const int nn = n0 + 64;
                  if (has_dres && c + 2 < BN / 32 && nn < N) {
                    mbar_arrive_expect_tx(&rf[b ^ 1], 4096);
                    tma_load_2d(ebuf + (b ^ 1) * 4096, &em.o2, &rf[b ^ 1], nn, m0);
                  }
                }
                __syncwarp();
                const uint32_t buf = smem_u32(ebuf + b * 4096);
                float4 dr[8];
                if (has_dres) {
                  mbar_wait(&rf[b], (rphase >> b) & 1u);
                  rphase ^= 1u << b;
#pragma unroll
                  for (int j = 0; j < 8; ++j) dr[j] = lds_f4(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)));
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) dr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                tmem_ld_wait();
                float dx[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float ds4[4] = {dr[j].x, dr[j].y, dr[j].z, dr[j].w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 gx = OP2_TO_F32(*reinterpret_cast<const op16x2*>(&pkf[4 * j + k]));
                    dx[4 * j + k] = fmaf(gx.x - gx.y * c1 - c2, rstd, ds4[k]);
                  }
                }
                if (ep.out_f32) {
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    sts_f4(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)), make_float4(dx[4 * j], dx[4 * j + 1], dx[4 * j + 2], dx[4 * j + 3]));
                  fence_async_smem();
                  __syncwarp();
                  if (lane == 0) {
                    tma_store_2d(&em.o32, ebuf + b * 4096, n0, m0);
                    bulk_commit();
                    if (ep.out_bf16) bulk_wait_read<0>();
                  }
                  __syncwarp();
                }
                if (ep.out_bf16) {
                  // (tried: the mirror as four 16-byte st.global per lane instead of a second TMA tile + store-drain wait: slower,
                  //  50.3 vs 46.5 us per launch at C = 512 - 32 partial-sector writes per instruction)
                  __syncwarp();
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float t8[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) t8[k] = dx[8 * j + k];
                    sts_u4(buf + (uint32_t)(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)), pack8(t8));
                  }
                  fence_async_smem();
                  __syncwarp();
                  if (lane == 0) {
                    tma_store_2d(&em.o16, ebuf + b * 4096, n0, m0);
                    bulk_commit();
                  }
                }
                if (ep.lnb_cs) {
                  if (m0 + lane >= M) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) dx[j] = 0.f;   // rows beyond M carry dres = 0 but -c2 * rstd garbage: mask
                  }
                  const float sc = warp_colsum32(dx, lane);
                  if (n0 + lane < N) atomicAdd(&s_cs[n0 + lane], sc);
                }
                ++ci;
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (tiles_n > 1) {
                mbar_arrive(&tempty[0]);
                mbar_arrive(&tempty[1]);
              } else {
                mbar_arrive(&tempty[acc]);
              }
            }
          }
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      } else
      if constexpr (EPI == EPI_GATE_TMA) {
        // ---- SimpleGate forward on 32-wide pair packing: accumulator chunks (2p, 2p + 1) = (a, b) halves of channels
        // [32p, 32p + 32): x4[:, 32p..] = a, x4[:, C + 32p..] = b, sg[:, 32p..] = a * b (all bf16, rounded before the product
        // like the register-staged epilogue).  A warp takes alternate chunk PAIRS; three 2 KiB tiles per pair.
        for (int it = 0, tile; (tile = tile_of(it)) >= 0; ++it) {
          const int n_t = tile % tiles_n;
          const int m_t = (tile / tiles_n) % tiles_m;
          const int m0 = m_t * BM + q * 32;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
          mbar_wait(&tfull[acc], acc_phase);
          tc_fence_after();
#pragma unroll 1
          for (int pr = chalf; pr < BN / 64; pr += 2) {
            const int n0 = n_t * BN + pr * 64;  // packed column of the a-chunk
            if (n0 >= N) break;                 // warp-uniform
            const int j0 = n0 >> 1;             // first channel of the pair
            float va[32], vb[32];
            tmem_ld32(taddr + pr * 64, va);
            tmem_ld32(taddr + pr * 64 + 32, vb);
            if (lane == 0) bulk_wait_read<0>();  // previous stores of this warp have drained the buffers
            __syncwarp();
            tmem_ld_wait();
            const uint32_t buf = smem_u32(ebuf);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a[8], b[8];
#pragma unroll
              for (int k = 0; k < 8; k += 4) {
                const float4 ba = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + 8 * j + k));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + 32 + 8 * j + k));
                a[k] = bf16_round(va[8 * j + k] + ba.x); a[k + 1] = bf16_round(va[8 * j + k + 1] + ba.y);
                a[k + 2] = bf16_round(va[8 * j + k + 2] + ba.z); a[k + 3] = bf16_round(va[8 * j + k + 3] + ba.w);
                b[k] = bf16_round(vb[8 * j + k] + bb.x); b[k + 1] = bf16_round(vb[8 * j + k + 1] + bb.y);
                b[k + 2] = bf16_round(vb[8 * j + k + 2] + bb.z); b[k + 3] = bf16_round(vb[8 * j + k + 3] + bb.w);
              }
              float sgv[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) sgv[k] = a[k] * b[k];
              const uint32_t off = (uint32_t)(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
              if (ep.out_bf16) {  // x4 = [a | b] is only needed by the backward (inference passes NULL)
                sts_u4(buf + off, pack8(a));
                sts_u4(buf + 2048 + off, pack8(b));
              }
              sts_u4(buf + 4096 + off, pack8(sgv));
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (ep.out_bf16) {
                tma_store_2d(&em.o16, ebuf, j0, m0);
                tma_store_2d(&em.o16, ebuf + 2048, ep.C + j0, m0);
              }
              tma_store_2d(&em.o2, ebuf + 4096, j0, m0);
              bulk_commit();
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      } else {
      // fused LayerNorm of the output rows (ep.ln_out): per-lane (= per row) shifted sums over this warp's chunks
      bool ln_on = false, wide16 = false;
      if constexpr (EPI == EPI_STORE_TMA) {
        ln_on = ep.ln_out != nullptr;
        wide16 = ep.out_bf16 != nullptr && ep.out_f32 == nullptr && !has_r && ep.Cseg == 0 && !ln_on;
      }
      float ln_k = 0.f, ln_s1 = 0.f, ln_s2 = 0.f;
      int ln_n = 0, ln_par = 0;
      for (int it = 0, tile; (tile = tile_of(it)) >= 0; ++it) {
        const int n_t = tile % tiles_n;
        const int m_t = (tile / tiles_n) % tiles_m;
        const int m0 = m_t * BM + q * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
        if constexpr (EPI == EPI_STORE_TMA) {
          if (wide16) {
            // ---- bf16-only output (conv1, the conv4 / conv1 dgrads, ...): a warp takes alternate 64-column chunk PAIRS and
            // writes each as ONE 4 KiB box of 128-byte rows; two stores in flight ----
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int pr = chalf; pr < BN / 64; pr += 2) {
              const int n0 = n_t * BN + pr * 64;
              if (n0 >= N) break;  // warp-uniform
              float va[32], vb[32];
              tmem_ld32(taddr + pr * 64, va);
              tmem_ld32(taddr + pr * 64 + 32, vb);
              const int b = ci & 1;
              if (lane == 0) bulk_wait_read<1>();  // the store that read buffer b two pairs ago has drained it
              __syncwarp();
              tmem_ld_wait();
              if (ep.bias) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (n0 + 4 * j < N) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j);
                    va[4 * j] += bv.x; va[4 * j + 1] += bv.y; va[4 * j + 2] += bv.z; va[4 * j + 3] += bv.w;
                  }
                  if (n0 + 32 + 4 * j < N) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + 32) + j);
                    vb[4 * j] += bv.x; vb[4 * j + 1] += bv.y; vb[4 * j + 2] += bv.z; vb[4 * j + 3] += bv.w;
                  }
                }
              }
              const uint32_t buf = smem_u32(ebuf + b * 4096);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float t8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) t8[k] = va[8 * j + k];
                sts_u4(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)), pack8(t8));
#pragma unroll
                for (int k = 0; k < 8; ++k) t8[k] = vb[8 * j + k];
                sts_u4(buf + (uint32_t)(lane * 128 + (((4 + j) ^ (lane & 7)) << 4)), pack8(t8));
              }
              fence_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&em.o16w, ebuf + b * 4096, n0, m0);
                bulk_commit();
              }
              ++ci;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (ew == 0 && lane == 0) TRACE(it == 0 ? 6 : 8);
            if (++acc == 2) {
              acc = 0;
              acc_phase ^= 1;
            }
            continue;
          }
        }
        bool first = true;
#pragma unroll 1
        for (int c = chalf; c < BN / 32; c += 2) {
          const int n0 = n_t * BN + c * 32;
          if (n0 >= N) break;  // warp-uniform
          const int b = ci & 1;
          if (first) {
            if (lane == 0) {
              bulk_wait_read<0>();
              if (has_r) issue_in(ebuf + b * 4096, &rf[b], n0, m0);
            }
            mbar_wait(&tfull[acc], acc_phase);
            if (ew == 0 && lane == 0) TRACE(it == 0 ? 5 : 7);
            tc_fence_after();
            first = false;
          }
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          if (lane == 0) {
            bulk_wait_read<0>();  // the store that read buffer b ^ 1 (previous chunk) has drained it
            const int nn = n0 + 64;
            if (has_r && c + 2 < BN / 32 && nn < N) issue_in(ebuf + (b ^ 1) * 4096, &rf[b ^ 1], nn, m0);
          }
          __syncwarp();
          const uint32_t buf = smem_u32(ebuf + b * 4096);
          if (has_r) {
            mbar_wait(&rf[b], (rphase >> b) & 1u);
            rphase ^= 1u << b;
          }
          tmem_ld_wait();
          if constexpr (EPI == EPI_GATE_BWD_TMA) {
            // d(x4)[:, n] = d(sg) * x4[:, C + n],  d(x4)[:, C + n] = d(sg) * x4[:, n]  (nafnet_arch.py:77-80), in place
            float da32[32];  // d(x4) first half (kept for the column sums); v is overwritten with the second half
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t off = (uint32_t)(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
              float xa[8], xb[8], da[8], db[8];
              const float4 ra = lds_f4(buf + off), rb = lds_f4(buf + 2048 + off);
              unpack8(*reinterpret_cast<const uint4*>(&ra), xa);
              unpack8(*reinterpret_cast<const uint4*>(&rb), xb);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                da[k] = v[8 * j + k] * xb[k];
                db[k] = v[8 * j + k] * xa[k];
                da32[8 * j + k] = da[k];
                v[8 * j + k] = db[k];
              }
              sts_u4(buf + off, pack8(da));
              sts_u4(buf + 2048 + off, pack8(db));
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&em.o16, ebuf + b * 4096, n0, m0);
              tma_store_2d(&em.o16, ebuf + b * 4096 + 2048, ep.C + n0, m0);
              bulk_commit();
            }
            if (ep.colsum) {  // conv4 bias gradient: rows beyond M hold zeros (zero-filled operands), columns are all valid (C % 32 == 0)
              const float sa = warp_colsum32(da32, lane);
              const float sb = warp_colsum32(v, lane);
              atomicAdd(&s_colsum[n0 + lane], sa);
              atomicAdd(&s_colsum[ep.C + n0 + lane], sb);
            }
            ++ci;
            continue;
          }
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (n0 + 4 * j < N) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j);
                v[4 * j] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w;
              }
            }
          }
          if (has_r && !has_g) {
            float4 r[8];  // all eight loads in flight before the first use (the asm loads keep program order)
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = lds_f4(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] += r[j].x; v[4 * j + 1] += r[j].y; v[4 * j + 2] += r[j].z; v[4 * j + 3] += r[j].w;
            }
          }
          if constexpr (EPI == EPI_STORE_TMA) {
            if (ln_on) {  // v = the finished output row piece: statistics, and back into the accumulator columns for pass 2
              const int nv = min(32, N - n0);
              if (ln_n == 0) ln_k = v[0];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j < nv) {
                  const float d = v[j] - ln_k;
                  ln_s1 += d;
                  ln_s2 = fmaf(d, d, ln_s2);
                }
              }
              ln_n += nv;
              tmem_st32(taddr + c * 32, v);
            }
          }
          if (ep.out_f32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_f4(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              store_out(&em.o32, ebuf + b * 4096, n0, m0);
              bulk_commit();
              if (ep.out_bf16) bulk_wait_read<0>();  // the mirror below reuses the buffer
            }
            __syncwarp();
          }
          if (ep.out_bf16) {
            __syncwarp();  // bf16 row t overlays the residual rows t / 2 of other lanes: every lane has read its row
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              const uint2 lo = pack4_bf16(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]);
              const uint2 hi = pack4_bf16(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]);
              pk.x = lo.x; pk.y = lo.y; pk.z = hi.x; pk.w = hi.y;
              sts_u4(buf + (uint32_t)(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)), pk);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              store_out(&em.o16, ebuf + b * 4096, n0, m0);
              bulk_commit();
            }
            if (has_g) {  // ds[img, n] += sum_rows bf16(out) * g   (one image per 32-row slab: rows_per_img % 32 == 0)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float gv[8];
                const float4 rg = lds_f4(buf + 2048 + (uint32_t)(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)));
                unpack8(*reinterpret_cast<const uint4*>(&rg), gv);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[8 * j + k] = bf16_round(v[8 * j + k]) * gv[k];
              }
              const float sgd = warp_colsum32(v, lane);
              if (m0 < M && n0 + lane < N) atomicAdd(ep.colsum + (size_t)(m0 / ep.rows_per_img) * N + n0 + lane, sgd);
            }
          }
          ++ci;
        }
        if (first) {  // no chunk of this tile belongs to the warp (N edge): still follow the accumulator hand-shake in step
          mbar_wait(&tfull[acc], acc_phase);
          tc_fence_after();
        }
        if constexpr (EPI == EPI_STORE_TMA) {
          if (ln_on) {
            if (n_t == tiles_n - 1) {
              // ---- the CTA now holds the whole rows of this slab in TMEM: finish the statistics, normalise, store ----
              tmem_st_wait();
              if (ew == 0 && lane == 0) TRACE(11);
              const float na = (float)ln_n;
              const float mean_a = ln_n ? ln_k + ln_s1 / na : 0.f;
              const float m2_a = ln_n ? fmaxf(ln_s2 - ln_s1 * ln_s1 / na, 0.f) : 0.f;
              float4* xch = reinterpret_cast<float4*>(s_colsum) + ln_par * 256;  // [2 parities][8 warps][32 lanes]
              xch[ew * 32 + lane] = make_float4(na, mean_a, m2_a, 0.f);
              tc_fence_before();
              named_bar_sync(1 + q, 64);  // the two warps of this lane quarter
              tc_fence_after();
              const float4 pb = xch[(ew ^ 4) * 32 + lane];
              if (ew == 0 && lane == 0) TRACE(12);
              ln_par ^= 1;
              const float nt = na + pb.x;
              const float mean = (na * mean_a + pb.x * pb.y) / nt;  // Chan et al. merge of the two column halves
              const float dm = pb.y - mean_a;
              const float var = ((m2_a + pb.z) + dm * dm * (na * pb.x / nt)) / nt;
              const float rstd = 1.f / sqrtf(var + ep.ln_eps);
              const float mu = ep.ln_nocenter ? 0.f : mean;  // Restormer's BiasFree LayerNorm divides the un-centred row by sqrt(var + eps)
              if (chalf == 0 && m0 + lane < M) *reinterpret_cast<float2*>(ep.ln_stats + (size_t)(m0 + lane) * 2) = make_float2(mean, rstd);
              if (lane == 0) bulk_wait_read<0>();  // pass 1's stores have drained the staging buffers pass 2 re-partitions
              __syncwarp();
              ci = 0;
              // pass 2: 64-column pairs (a pair straddles the two partner warps' pass-1 chunks: their TMEM write-backs are
              // ordered by the fences around the named barrier above), one 4 KiB box of 128-byte rows per store
              for (int h = 0; h < tiles_n; ++h) {
                const int hacc = tiles_n > 1 ? h : acc;
                const uint32_t th = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hacc * BN);
#pragma unroll 1
                for (int pr = chalf; pr < BN / 64; pr += 2) {
                  const int n0 = h * BN + pr * 64;
                  if (n0 >= N) break;
                  float va[32], vb[32];
                  tmem_ld32(th + pr * 64, va);
                  tmem_ld32(th + pr * 64 + 32, vb);
                  const int b = ci & 1;
                  if (lane == 0) bulk_wait_read<1>();  // the store that read buffer b two pairs ago has drained it
                  __syncwarp();
                  tmem_ld_wait();
                  const uint32_t buf = smem_u32(ebuf + b * 4096);
                  // all 32 parameter loads of the pair are issued before the first shared-memory store: the stores are
                  // volatile asm, and a load placed after one waits for it (r02e: 2700 cycles per pair interleaved)
                  uint4 pk[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    float o[8];
#pragma unroll
                    for (int k = 0; k < 8; k += 4) {
                      const int col = n0 + 8 * j + k;
                      float4 wv = make_float4(0.f, 0.f, 0.f, 0.f), bv = wv;
                      if (col < N) {
                        wv = __ldg(reinterpret_cast<const float4*>(ep.ln_w + col));
                        if (ep.ln_b) bv = __ldg(reinterpret_cast<const float4*>(ep.ln_b + col));
                      }
                      const float* src = j < 4 ? &va[8 * j + k] : &vb[8 * (j - 4) + k];
                      o[k] = (src[0] - mu) * rstd * wv.x + bv.x;
                      o[k + 1] = (src[1] - mu) * rstd * wv.y + bv.y;
                      o[k + 2] = (src[2] - mu) * rstd * wv.z + bv.z;
                      o[k + 3] = (src[3] - mu) * rstd * wv.w + bv.w;
                    }
                    pk[j] = pack8(o);
                  }
#pragma unroll
                  for (int j = 0; j < 8; ++j) sts_u4(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)), pk[j]);
                  fence_async_smem();
                  __syncwarp();
                  if (lane == 0) {
                    tma_store_2d(&em.ln, ebuf + b * 4096, n0, m0);
                    bulk_commit();
                  }
                  ++ci;
                }
              }
              // (pass 1 of the next tile starts with bulk_wait_read<0>, so any ci parity is fine there)
              ln_k = ln_s1 = ln_s2 = 0.f;
              ln_n = 0;
              if (ew == 0 && lane == 0) TRACE(13);
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (tiles_n > 1) {
                  mbar_arrive(&tempty[0]);
                  mbar_arrive(&tempty[1]);
                } else {
                  mbar_arrive(&tempty[acc]);
                }
              }
            }
            if (++acc == 2) {
              acc = 0;
              acc_phase ^= 1;
            }
            continue;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (ew == 0 && lane == 0) TRACE(it == 0 ? 6 : 8);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      }
      if (lane == 0) bulk_wait_all();
      if (ew == 0 && lane == 0) TRACE(9);
    } else
    for (int it = 0, tile; (tile = tile_of(it)) >= 0; ++it) {
      const int n_t = tile % tiles_n;
      const int m_t = (tile / tiles_n) % tiles_m;
      const int m_base = m_t * BM + q * 32;
      EpiParams ept = ep;  // per-tile view of the epilogue parameters (batched split-K: output of this split's batch)
      if constexpr (EPI == EPI_ATOMIC) {
        if (bg.kb_per_batch > 0) ept.out_f32 += (size_t)((tile / (tiles_n * tiles_m)) / bg.splits_per_batch) * ep.out_batch_stride;
      }
      // row r of this warp's 32-row slab -> global output row (or -1): linear for GEMMs, patch pixel for CONV == 1
      int cv_img = 0, cv_h0 = 0, cv_w0 = 0;
      if constexpr (CONV == 1) {
        const int per_img = cg.tiles_h * cg.tiles_w;
        cv_img = m_t / per_img;
        const int rem = m_t - cv_img * per_img;
        cv_h0 = (rem / cg.tiles_w) * cg.th;
        cv_w0 = (rem % cg.tiles_w) << cg.tw_log2;
      }
      auto out_row = [&](int r) -> int {
        if constexpr (CONV == 1) {
          const int rt = q * 32 + r;
          const int h = cv_h0 + (rt >> cg.tw_log2), w = cv_w0 + (rt & ((1 << cg.tw_log2) - 1));
          return (h < cg.H && w < cg.W) ? (cv_img * cg.H + h) * cg.W + w : -1;
        } else {
          return m_base + r < M ? m_base + r : -1;
        }
      };
      // While the tensor core works on this tile, pull the epilogue's global inputs into L2.
      if constexpr (EPI == EPI_STORE && CONV == 0) {
        if (ep.resid) {
          for (int i = lane + chalf * 32; i < 32 * (BN / 32); i += 64) {
            const int r = i / (BN / 32), l = i % (BN / 32);
            if (m_base + r < M && n_t * BN + l * 32 < N) prefetch_l2(ep.resid + (size_t)(m_base + r) * ep.ldr + n_t * BN + l * 32);
          }
        }
      } else if constexpr (EPI == EPI_GATE_BWD) {
        for (int i = lane + chalf * 32; i < 32 * (BN / 64) * 2; i += 64) {
          const int r = i / (BN / 32), l = i % (BN / 32), hf = l / (BN / 64), ll = l % (BN / 64);
          const int n = n_t * BN + ll * 64;
          if (m_base + r < M && n < N) prefetch_l2(ep.aux + (size_t)(m_base + r) * ep.ldaux + hf * ep.C + n);
        }
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (EPI == EPI_GATE) {
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = chalf; c < BN / 32; c += 2) {
          const int n0 = n_t * BN + c * 32;
          if (n0 >= N) break;  // warp-uniform
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          // 4 lanes per row: lane -> (group g, quarter qq) = the a-piece / b-piece pair of 4 channels
          const int pc = lane & 3, g = pc >> 1, qq = pc & 1;
          const int na = n0 + g * 16 + qq * 4;
          const int r0 = lane >> 2;
          EpiExtra ex;
          ex.c = epilogue_load_col<EPI>(ep, na < N ? na : n0);
          ex.r = EpiRow();
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_f4(stage + (uint32_t)(lane * 8 + (j ^ (lane & 7))) * 16u, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + r0;
            const float4 a = lds_f4(stage + (uint32_t)(r * 8 + ((g * 4 + qq) ^ (r & 7))) * 16u);
            const float4 b = lds_f4(stage + (uint32_t)(r * 8 + ((g * 4 + 2 + qq) ^ (r & 7))) * 16u);
            const int m = out_row(r);
            if (m >= 0 && na < N) epilogue_store<EPI>(ep, m, na, a, b, ex);
          }
          __syncwarp();
        }
      } else {
        // a lane owns 4 consecutive columns (cc) of rows r0, r0 + 4, ...: 8 lanes cover 128 contiguous bytes.
        // The per-row global inputs (residual / gate operands) of the NEXT chunk are requested before the current
        // chunk is finished, and those of the first chunk before the accumulator is even ready: their latency
        // (the longest in the epilogue) is hidden behind the TMEM load, the transpose and the stores.
        const int cc = lane & 7, r0 = lane >> 3;
        constexpr bool kRowLoads = (EPI == EPI_STORE || EPI == EPI_GATE_BWD || EPI == EPI_PIXSHUF);
        EpiRow cur[8];
        auto load_rows = [&](int c, EpiRow (&rw)[8]) {
          const int n = n_t * BN + c * 32 + cc * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int m = out_row(it * 4 + r0);
            if (m >= 0 && n < N) rw[it] = epilogue_load_row<EPI>(ep, m, n);
            else rw[it] = EpiRow();
          }
        };
        if constexpr (kRowLoads) load_rows(chalf, cur);
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = chalf; c < BN / 32; c += 2) {
          const int n0 = n_t * BN + c * 32;
          if (n0 >= N) break;  // warp-uniform
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          const int n = n0 + cc * 4;
          EpiExtra ex;
          ex.c = epilogue_load_col<EPI>(ep, n < N ? n : n0);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_f4(stage + (uint32_t)(lane * 8 + (j ^ (lane & 7))) * 16u, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          __syncwarp();
          EpiRow nxt[8];
          if constexpr (kRowLoads) {
            if (c + 2 < BN / 32) load_rows(c + 2, nxt);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + r0;
            const float4 x = lds_f4(stage + (uint32_t)(r * 8 + (cc ^ (r & 7))) * 16u);
            const int m = out_row(r);
            ex.r = cur[it];
            if (m >= 0 && n < N) epilogue_store<EPI>(ept, m, n, x, x, ex);
          }
          __syncwarp();
          if constexpr (kRowLoads) {
#pragma unroll
            for (int it = 0; it < 8; ++it) cur[it] = nxt[it];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TRACE(10);
  if constexpr (EPI == EPI_GATE_BWD_TMA) {
    if (ep.colsum)
      for (int i = threadIdx.x; i < 2 * ep.C; i += kThreads) {
        const float sv = s_colsum[i];
        if (sv != 0.f) atomicAdd(ep.colsum + i, sv);
      }
  }
  if constexpr (EPI == EPI_LNBWD_TMA) {
    for (int i = threadIdx.x; i < N; i += kThreads) {
      const float a = s_colsum[2048 + i], b = s_colsum[2048 + 512 + i], c = s_colsum[2048 + 1024 + i];
      if (a != 0.f) atomicAdd(ep.lnb_dw + i, a);
      if (b != 0.f) atomicAdd(ep.lnb_db + i, b);
      if (ep.lnb_cs && c != 0.f) atomicAdd(ep.lnb_cs + i, c);
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------ host side -------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

// Row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box = 64 x box_rows, 128B swizzle.
int make_tmap_2d(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  DCPT_CHECK_ARG(fn != nullptr, DCPT_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  DCPT_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld % 8) == 0, DCPT_E_ALIGN,
                 "GEMM operand must be 16-byte aligned with ld %% 8 == 0 (ptr=%p ld=%lld)", ptr, ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, DCPT_TMAP_OP16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCPT_CHECK_ARG(r == CUDA_SUCCESS, DCPT_E_DRIVER, "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%d",
                 (int)r, rows, cols, ld, box_rows);
  return 0;
}

int make_tmap_epi(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int elem_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  DCPT_CHECK_ARG(fn != nullptr, DCPT_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  DCPT_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * elem_bytes) % 16 == 0, DCPT_E_ALIGN,
                 "GEMM epilogue tensor must be 16-byte aligned with a 16-byte row pitch (ptr=%p ld=%lld)", ptr, ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : DCPT_TMAP_OP16, 2, const_cast<void*>(ptr),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  elem_bytes == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCPT_CHECK_ARG(r == CUDA_SUCCESS, DCPT_E_DRIVER, "cuTensorMapEncodeTiled(epilogue) failed (%d): rows=%lld cols=%lld ld=%lld", (int)r,
                 rows, cols, ld);
  return 0;
}

int make_tmap_pixshuf(CUtensorMap* tm, const void* ptr, long long NH, int W, int Cseg, int elem_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  DCPT_CHECK_ARG(fn != nullptr, DCPT_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  const int bw = W < 32 ? W : 32, bh = 32 / bw;
  DCPT_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && Cseg % 32 == 0 && bw * bh == 32 && W % bw == 0, DCPT_E_ALIGN,
                 "pixel-shuffle epilogue tensor: need Cseg %% 32 == 0 and W a power of two >= 8 (W=%d Cseg=%d)", W, Cseg);
  const cuuint64_t e = (cuuint64_t)elem_bytes;
  cuuint64_t dims[5] = {(cuuint64_t)Cseg, 2, (cuuint64_t)W, 2, (cuuint64_t)NH};
  cuuint64_t strides[4] = {(cuuint64_t)Cseg * e, 2ull * Cseg * e, 2ull * W * Cseg * e, 4ull * W * Cseg * e};
  cuuint32_t box[5] = {32, 1, (cuuint32_t)bw, 1, (cuuint32_t)bh};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : DCPT_TMAP_OP16, 5, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  elem_bytes == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCPT_CHECK_ARG(r == CUDA_SUCCESS, DCPT_E_DRIVER, "cuTensorMapEncodeTiled(pixshuf) failed (%d): NH=%lld W=%d Cseg=%d", (int)r, NH, W, Cseg);
  return 0;
}

int make_tmap_nhwc(CUtensorMap* tm, const void* ptr, int N, int H, int W, int CH, int box_w, int box_h, int swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  DCPT_CHECK_ARG(fn != nullptr, DCPT_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  DCPT_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (CH % 8) == 0, DCPT_E_ALIGN,
                 "NHWC tensor must be 16-byte aligned with C %% 8 == 0 (ptr=%p C=%d)", ptr, CH);
  cuuint64_t dims[4] = {(cuuint64_t)CH, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)CH * 2, (cuuint64_t)W * CH * 2, (cuuint64_t)H * W * CH * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, DCPT_TMAP_OP16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCPT_CHECK_ARG(r == CUDA_SUCCESS, DCPT_E_DRIVER, "cuTensorMapEncodeTiled(4d) failed (%d): N=%d H=%d W=%d C=%d box=%dx%d", (int)r, N,
                 H, W, CH, box_w, box_h);
  return 0;
}

namespace {

// L2 prefetch policy of the TMA producer (DCPT_GEMM_PF): 0 none, 1 a burst of PF_DIST k-blocks before the first load of a tile,
// 2 the same distance but issued behind the real loads
int gemm_pf_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DCPT_GEMM_PF");
    mode = e ? atoi(e) : 0;  // r02b: 47.8 (none) vs 45.9 (burst) vs 46.4 (behind the loads) MPix/s
  }
  return mode;
}

template <int BN, int EPI, bool A_MN, bool B_MN>
int launch_cfg(const GemmArgs& g, cudaStream_t stream) {
  const ConvGeom cg0 = {};
  using C = Cfg<BN, EPI>;
  CUtensorMap tmA, tmB;
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  if constexpr (EPI == EPI_STORE_TMA) {
    if (g.ep.Cseg > 0) {  // PixelShuffle scatter (EPI_PIXSHUF routed here): 5-D maps of the [N, 2H, 2W, Cseg] output / skip
      const long long NH = (long long)(g.M / g.ep.W);
      if (g.ep.out_f32) DCPT_TRY(make_tmap_pixshuf(&em.o32, g.ep.out_f32, NH, g.ep.W, g.ep.Cseg, 4));
      if (g.ep.out_bf16) DCPT_TRY(make_tmap_pixshuf(&em.o16, g.ep.out_bf16, NH, g.ep.W, g.ep.Cseg, 2));
      if (g.ep.resid) DCPT_TRY(make_tmap_pixshuf(&em.r32, g.ep.resid, NH, g.ep.W, g.ep.Cseg, 4));
    } else {
    if (g.ep.out_f32) DCPT_TRY(make_tmap_epi(&em.o32, g.ep.out_f32, g.M, g.N, g.ep.ldo, 4));
    if (g.ep.out_bf16) DCPT_TRY(make_tmap_epi(&em.o16, g.ep.out_bf16, g.M, g.N, g.ep.ldo, 2));
    if (g.ep.resid) DCPT_TRY(make_tmap_epi(&em.r32, g.ep.resid, g.M, g.N, g.ep.ldr, 4));
    }
    if (g.ep.gaux) DCPT_TRY(make_tmap_epi(&em.o2, g.ep.gaux, g.M, g.N, g.ep.ldgaux, 2));
    if (g.ep.ln_out) {
      // (per-image B operands are fine: a 128-row slab never straddles two images, m_per_batch % 128 == 0)
      DCPT_CHECK_ARG(g.ep.Cseg == 0 && !g.ep.gaux && g.N <= 2 * BN && g.ep.ln_w && g.ep.ln_stats &&
                         ((reinterpret_cast<uintptr_t>(g.ep.ln_w) | reinterpret_cast<uintptr_t>(g.ep.ln_b)) & 15) == 0,
                     DCPT_E_ARG, "gemm: fused LayerNorm needs a plain STORE epilogue with N <= %d and 16-byte aligned ln_w / ln_b (N=%d)",
                     2 * BN, g.N);
      DCPT_CHECK_ARG(g.ep.ld_ln % 8 == 0, DCPT_E_ALIGN, "gemm: ld_ln=%d must be a multiple of 8", g.ep.ld_ln);
      DCPT_TRY(make_tmap_2d(&em.ln, g.ep.ln_out, g.M, g.N, g.ep.ld_ln, 32));  // 64-column x 32-row boxes, 128B swizzle
    }
    if (g.ep.out_bf16 && !g.ep.out_f32 && g.ep.Cseg == 0 && g.ep.ldo % 8 == 0)
      DCPT_TRY(make_tmap_2d(&em.o16w, g.ep.out_bf16, g.M, g.N, g.ep.ldo, 32));
  } else if constexpr (EPI == EPI_GATE_BWD_TMA) {
    DCPT_TRY(make_tmap_epi(&em.r32, g.ep.aux, g.M, 2 * g.ep.C, g.ep.ldaux, 2));
    DCPT_TRY(make_tmap_epi(&em.o16, g.ep.out_bf16, g.M, 2 * g.ep.C, g.ep.ldo, 2));
  } else if constexpr (EPI == EPI_GATE_TMA) {
    if (g.ep.out_bf16) DCPT_TRY(make_tmap_epi(&em.o16, g.ep.out_bf16, g.M, 2 * g.ep.C, g.ep.ldo, 2));
    DCPT_TRY(make_tmap_epi(&em.o2, g.ep.out2, g.M, g.ep.C, g.ep.ldo2, 2));
  } else if constexpr (EPI == EPI_LNBWD_TMA) {
    DCPT_CHECK_ARG(g.N <= 2 * BN && g.N <= 512 && g.ep.lnb_stats && g.ep.lnb_w && g.ep.lnb_dw && g.ep.lnb_db && (g.ep.out_f32 || g.ep.out_bf16) &&
                       !g.ep.bias && !g.ep.resid && g.m_per_batch == 0 && (reinterpret_cast<uintptr_t>(g.ep.lnb_w) & 15) == 0,
                   DCPT_E_ARG, "gemm: fused LayerNorm backward needs N <= 512, stats / weight / dw / db and no bias / residual (N=%d)", g.N);
    if (g.ep.out_f32) DCPT_TRY(make_tmap_epi(&em.o32, g.ep.out_f32, g.M, g.N, g.ep.ldo, 4));
    if (g.ep.out_bf16) DCPT_TRY(make_tmap_epi(&em.o16, g.ep.out_bf16, g.M, g.N, g.ep.ldo, 2));
    DCPT_TRY(make_tmap_epi(&em.r32, g.ep.lnb_x, g.M, g.N, g.ep.ld_lnb, 4));
    if (g.ep.lnb_dres) DCPT_TRY(make_tmap_epi(&em.o2, g.ep.lnb_dres, g.M, g.N, g.ep.ld_lnb, 4));
  }
  if (!g.a_mn) DCPT_TRY(make_tmap_2d(&tmA, g.A, g.M, g.K, g.lda, BM));
  else DCPT_TRY(make_tmap_2d(&tmA, g.A, g.K, g.M, g.lda, 64));
  if (!g.b_mn) DCPT_TRY(make_tmap_2d(&tmB, g.B, g.m_per_batch > 0 ? (long long)(g.M / g.m_per_batch) * g.b_rows_per_batch : g.N, g.K, g.ldb, BN));
  else DCPT_TRY(make_tmap_2d(&tmB, g.B, g.K, g.N, g.ldb, 64));

  const int tiles_m = ceil_div(g.M, BM), tiles_n = ceil_div(g.N, BN);
  const int num_kb = ceil_div(g.K, BK);
  int splits = g.splits < 1 ? 1 : g.splits;
  BatchGeom bg = {};
  int kbps;
  if (g.k_per_batch > 0) {  // batched split-K: `splits` per batch, none straddles a batch
    DCPT_CHECK_ARG(g.a_mn && g.b_mn && EPI == EPI_ATOMIC && g.k_per_batch % BK == 0 && g.K % g.k_per_batch == 0, DCPT_E_ARG,
                   "gemm: batched split-K needs MN-major operands, the atomic epilogue and k_per_batch %% 64 == 0");
    bg.kb_per_batch = g.k_per_batch / BK;
    if (splits > bg.kb_per_batch) splits = bg.kb_per_batch;
    kbps = ceil_div(bg.kb_per_batch, splits);
    bg.splits_per_batch = ceil_div(bg.kb_per_batch, kbps);
    splits = bg.splits_per_batch * (g.K / g.k_per_batch);
  } else {
    if (splits > num_kb) splits = num_kb;
    kbps = ceil_div(num_kb, splits);
    splits = ceil_div(num_kb, kbps);  // every split gets >= 1 k-block
  }
  if (g.m_per_batch > 0) {
    DCPT_CHECK_ARG(!g.a_mn && !g.b_mn && g.m_per_batch % BM == 0 && g.M % g.m_per_batch == 0, DCPT_E_ARG,
                   "gemm: batched B needs K-major operands and m_per_batch %% 128 == 0 (m_per_batch=%d)", g.m_per_batch);
    bg.m_per_batch = g.m_per_batch;
    bg.b_rows_per_batch = g.b_rows_per_batch;
  }
  const int total = tiles_m * tiles_n * splits;
  int grid = total < dcpt_num_sms() ? total : dcpt_num_sms();
  if (EPI == EPI_ATOMIC && g.fine_grid) grid = total;  // one work item per CTA: short-lived CTAs (wgrad on the low-priority stream)
  if ((EPI == EPI_STORE_TMA && g.ep.ln_out && tiles_n > 1) || (EPI == EPI_LNBWD_TMA && tiles_n > 1))
    grid = tiles_m < dcpt_num_sms() ? tiles_m : dcpt_num_sms();  // whole slabs per CTA

  auto kern = gemm_tc_kernel<BN, EPI, A_MN, B_MN, 0>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    DCPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    attr_set = true;
  }
  static char base_tag[48] = "";
  if (!base_tag[0]) {
    static const char* epi_names[] = {"store", "gate", "gate_bwd", "pixshuf", "atomic", "store_tma", "gate_bwd_tma", "gate_tma", "lnbwd_tma"};
    snprintf(base_tag, sizeof(base_tag), "gemm_tc<%d,%s,%s>", BN, epi_names[EPI], A_MN ? "mn" : "k");
  }
  const char* tag = base_tag;
  if (g_dcpt_prof_on && g_dcpt_prof_shapes) {
    char buf[96];
    snprintf(buf, sizeof(buf), "%s %dx%dx%d s%d", base_tag, g.M, g.N, g.K, splits);
    tag = dcpt_prof_intern(buf);
  }
  const double out_bytes = (double)g.M * g.N * ((g.ep.out_f32 ? 4.0 : 0.0) + (g.ep.out_bf16 ? 2.0 : 0.0) + (g.ep.resid ? 4.0 : 0.0) +
                                                (EPI == EPI_GATE || EPI == EPI_GATE_TMA ? 1.0 : 0.0) +
                                                (EPI == EPI_GATE_BWD || EPI == EPI_GATE_BWD_TMA ? 4.0 : 0.0));
  DCPT_PROF(tag, 2.0 * g.M * g.N * g.K, 2.0 * ((double)g.M * g.K + (double)g.N * g.K) + out_bytes, stream);
  DCPT_CUDA(dcpt_launch_pdl(kern, dim3(grid), dim3(kThreads), C::SMEM_BYTES, stream, tmA, tmB, em, g.M, g.N, g.K, tiles_m, tiles_n, splits,
                            kbps, g.ep, cg0, bg, gemm_pf_mode()));
  DCPT_LAUNCH_CHECK();
  return 0;
}

template <int EPI, bool A_MN, bool B_MN>
int launch_bn(const GemmArgs& g, cudaStream_t stream) {
  // Largest tile that N fills; small-N GEMMs (C = 64 levels) are HBM-bound and want more, smaller tiles.
  if (g.N > 128) return launch_cfg<256, EPI, A_MN, B_MN>(g, stream);
  if (g.N > 64) return launch_cfg<128, EPI, A_MN, B_MN>(g, stream);
  return launch_cfg<64, EPI, A_MN, B_MN>(g, stream);
}

ConvGeom conv_geom(int H, int W, int cin_pad, int rows_per_tile) {
  ConvGeom cg;
  cg.H = H; cg.W = W;
  cg.tw_log2 = W >= 16 ? 4 : (W >= 8 ? 3 : 2);
  cg.th = rows_per_tile >> cg.tw_log2;
  cg.tiles_w = ceil_div(W, 1 << cg.tw_log2);
  cg.tiles_h = ceil_div(H, cg.th);
  cg.cblocks = cin_pad / 64;
  return cg;
}

template <int BN>
int conv_fwd_cfg(const Conv3x3Args& a, cudaStream_t stream) {
  using C = Cfg<BN, EPI_STORE>;
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  const int cin_pad = ceil_div(a.Cin, 64) * 64;
  const ConvGeom cg = conv_geom(a.H, a.W, cin_pad, BM);
  CUtensorMap tmA, tmB;
  DCPT_TRY(make_tmap_nhwc(&tmA, a.X, a.N, a.H, a.W, a.Cin, 1 << cg.tw_log2, cg.th, 1));
  DCPT_TRY(make_tmap_2d(&tmB, a.Wp, a.Cout, 9 * cin_pad, 9 * cin_pad, BN));
  const int tiles_m = a.N * cg.tiles_h * cg.tiles_w, tiles_n = ceil_div(a.Cout, BN), num_kb = 9 * cg.cblocks;
  const int total = tiles_m * tiles_n;
  const int grid = total < dcpt_num_sms() ? total : dcpt_num_sms();
  auto kern = gemm_tc_kernel<BN, EPI_STORE, false, false, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    DCPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    attr_set = true;
  }
  const double Mpx = (double)a.N * a.H * a.W;
  DCPT_PROF(BN == 256 ? "conv3x3_tc<256>" : (BN == 128 ? "conv3x3_tc<128>" : "conv3x3_tc<64>"), 2.0 * Mpx * a.Cout * 9 * a.Cin,
            2.0 * Mpx * (a.Cin + a.Cout), stream);
  kern<<<grid, kThreads, C::SMEM_BYTES, stream>>>(tmA, tmB, em, a.N * a.H * a.W, a.Cout, num_kb * BK, tiles_m, tiles_n, 1, num_kb, a.ep, cg, BatchGeom{}, 0);
  DCPT_LAUNCH_CHECK();
  return 0;
}

template <int BN>
int conv_wgrad_cfg(const bf16* dY, const bf16* X, float* G, int N, int H, int W, int Cin, int Cout, cudaStream_t stream) {
  using C = Cfg<BN, EPI_ATOMIC>;
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  const int cin_pad = ceil_div(Cin, 64) * 64;
  const ConvGeom cg = conv_geom(H, W, cin_pad, 64);
  CUtensorMap tmA, tmB;
  DCPT_TRY(make_tmap_nhwc(&tmA, dY, N, H, W, Cout, 1 << cg.tw_log2, cg.th, 1));
  DCPT_TRY(make_tmap_nhwc(&tmB, X, N, H, W, Cin, 1 << cg.tw_log2, cg.th, 1));
  const int tiles_m = ceil_div(Cout, BM), tiles_n = 9 * cin_pad / BN, num_kb = N * cg.tiles_h * cg.tiles_w;
  int splits = gemm_auto_splits(tiles_m * tiles_n, num_kb);
  const int kbps = ceil_div(num_kb, splits);
  splits = ceil_div(num_kb, kbps);
  const int total = tiles_m * tiles_n * splits;
  const int grid = total < dcpt_num_sms() ? total : dcpt_num_sms();
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.out_f32 = G; ep.ldo = 9 * cin_pad;
  auto kern = gemm_tc_kernel<BN, EPI_ATOMIC, true, true, 2>;
  static bool attr_set = false;
  if (!attr_set) {
    DCPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    attr_set = true;
  }
  const double Mpx = (double)N * H * W;
  DCPT_PROF("conv3x3_wgrad_tc", 2.0 * Mpx * Cout * 9 * Cin, 2.0 * Mpx * (Cin + Cout), stream);
  kern<<<grid, kThreads, C::SMEM_BYTES, stream>>>(tmA, tmB, em, Cout, 9 * cin_pad, num_kb * BK, tiles_m, tiles_n, splits, kbps, ep, cg, BatchGeom{}, 0);
  DCPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int conv3x3_tc_launch(const Conv3x3Args& a, cudaStream_t stream) {
  DCPT_CHECK_ARG(a.N > 0 && a.H > 0 && a.W > 0 && a.Cin % 8 == 0 && a.Cout % 8 == 0 && a.Cin >= 8 && a.Cout >= 8, DCPT_E_SHAPE,
                 "conv3x3: need Cin, Cout multiples of 8 (N=%d H=%d W=%d Cin=%d Cout=%d)", a.N, a.H, a.W, a.Cin, a.Cout);
  DCPT_CHECK_ARG((long long)a.N * a.H * a.W < (1ll << 31) / 1024, DCPT_E_SHAPE, "conv3x3: too many pixels");
  if (a.Cout > 128) return conv_fwd_cfg<256>(a, stream);
  if (a.Cout > 64) return conv_fwd_cfg<128>(a, stream);
  return conv_fwd_cfg<64>(a, stream);
}

int conv3x3_wgrad_tc_launch(const bf16* dY, const bf16* X, float* G, int N, int H, int W, int Cin, int Cout, cudaStream_t stream) {
  DCPT_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cin % 8 == 0 && Cout % 8 == 0, DCPT_E_SHAPE, "conv3x3 wgrad: bad shape");
  const int cin_pad = ceil_div(Cin, 64) * 64;
  // one N tile must stay inside one tap: BN divides cin_pad
  if (cin_pad % 256 == 0) return conv_wgrad_cfg<256>(dY, X, G, N, H, W, Cin, Cout, stream);
  if (cin_pad % 128 == 0) return conv_wgrad_cfg<128>(dY, X, G, N, H, W, Cin, Cout, stream);
  return conv_wgrad_cfg<64>(dY, X, G, N, H, W, Cin, Cout, stream);
}

int gemm_tc_launch(const GemmArgs& g, cudaStream_t stream) {
  DCPT_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0, DCPT_E_SHAPE, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  DCPT_CHECK_ARG(g.N % 8 == 0, DCPT_E_SHAPE, "gemm: N=%d must be a multiple of 8", g.N);
  DCPT_CHECK_ARG(g.splits <= 1 || g.epi == EPI_ATOMIC, DCPT_E_ARG, "gemm: split-K needs the atomic epilogue");
  if (g.a_mn != g.b_mn) {
    dcpt_set_error("gemm: mixed operand majorness is not instantiated");
    return DCPT_E_UNSUPPORTED;
  }
  if (g.a_mn) {
    DCPT_CHECK_ARG(g.epi == EPI_ATOMIC, DCPT_E_UNSUPPORTED, "gemm: MN-major operands are only used by wgrad (atomic epilogue)");
    return launch_bn<EPI_ATOMIC, true, true>(g, stream);
  }
  switch (g.epi) {
    case EPI_STORE: {
      // TMA epilogue whenever the output / residual rows are 16-byte pitched (always true on the hot path)
      static const bool no_tma = getenv("DCPT_GEMM_NO_TMA_EPI") != nullptr;
      auto ok = [](const void* p, long long ld, int eb) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * eb) % 16 == 0); };
      const bool tma = !no_tma && (g.ep.out_f32 || g.ep.out_bf16) && ok(g.ep.out_f32, g.ep.ldo, 4) && ok(g.ep.out_bf16, g.ep.ldo, 2) &&
                       ok(g.ep.resid, g.ep.ldr, 4) && (g.ep.bias == nullptr || (reinterpret_cast<uintptr_t>(g.ep.bias) & 15) == 0);
      // fused SCA-backward reduction ds[img, c] += out * gaux: needs the TMA epilogue, a bf16-only output and whole
      // 32-row slabs per image; otherwise the separate reduction kernel runs after the GEMM
      const bool want_g = g.ep.gaux != nullptr && g.ep.colsum != nullptr;
      const bool fuse_g = want_g && tma && !g.ep.resid && !g.ep.out_f32 && g.ep.out_bf16 && g.ep.rows_per_img > 0 &&
                          g.ep.rows_per_img % 32 == 0 && ok(g.ep.gaux, g.ep.ldgaux, 2);
      if (g.ep.lnb_x) {  // the accumulator is d(LN output): LayerNorm backward in the epilogue
        DCPT_CHECK_ARG(!no_tma && gemm_ln_fusable(g.N) && ok(g.ep.out_f32, g.ep.ldo, 4) && ok(g.ep.out_bf16, g.ep.ldo, 2) &&
                           ok(g.ep.lnb_x, g.ep.ld_lnb, 4) && ok(g.ep.lnb_dres, g.ep.ld_lnb, 4) && !g.ep.ln_out && !g.ep.gaux,
                       DCPT_E_ARG, "gemm: fused LayerNorm backward needs 16-byte pitched rows and N <= 512 (N=%d)", g.N);
        return launch_bn<EPI_LNBWD_TMA, false, false>(g, stream);
      }
      DCPT_CHECK_ARG(g.ep.ln_out == nullptr || (tma && gemm_ln_fusable(g.N) && g.ep.out_f32), DCPT_E_ARG,
                     "gemm: fused LayerNorm needs the TMA-tiled fp32 STORE epilogue and N <= 512 (N=%d)", g.N);
      GemmArgs gg = g;
      gg.ep.Cseg = 0;  // PixelShuffle addressing belongs to EPI_PIXSHUF only
      if (!fuse_g) gg.ep.gaux = nullptr;
      DCPT_TRY(tma ? (launch_bn<EPI_STORE_TMA, false, false>(gg, stream)) : (launch_bn<EPI_STORE, false, false>(gg, stream)));
      if (want_g && !fuse_g) {
        DCPT_CHECK_ARG(g.ep.out_bf16 && g.ep.ldo == g.N && g.ep.ldgaux == g.N && g.ep.rows_per_img > 0 && g.M % g.ep.rows_per_img == 0,
                       DCPT_E_ARG, "gemm: the SCA-backward reduction needs a dense bf16 output and whole images");
        return sca_ds_reduce_launch(g.ep.out_bf16, g.ep.gaux, g.ep.colsum, g.M / g.ep.rows_per_img, g.ep.rows_per_img, g.N, stream);
      }
      return 0;
    }
    case EPI_GATE: return launch_bn<EPI_GATE, false, false>(g, stream);
    case EPI_GATE_TMA:  // SimpleGate forward on 32-wide pair packing (PACK_PAIR32 weights / bias)
      DCPT_CHECK_ARG(g.ep.C % 32 == 0 && g.N == 2 * g.ep.C && g.ep.bias && g.ep.out2, DCPT_E_ARG,
                     "gemm: the 32-wide gate epilogue needs C %% 32 == 0, N = 2C, bias and the gated output (C=%d N=%d)", g.ep.C, g.N);
      return launch_bn<EPI_GATE_TMA, false, false>(g, stream);
    case EPI_GATE_BWD: {
      static const bool no_tma = getenv("DCPT_GEMM_NO_TMA_EPI") != nullptr;
      if (!no_tma && g.ep.C % 32 == 0 && g.N == g.ep.C && (g.ep.ldo % 8) == 0 && (g.ep.ldaux % 8) == 0 &&
          ((reinterpret_cast<uintptr_t>(g.ep.out_bf16) | reinterpret_cast<uintptr_t>(g.ep.aux)) & 15) == 0 && g.ep.C <= 1024)
        return launch_bn<EPI_GATE_BWD_TMA, false, false>(g, stream);  // conv4's bias gradient (ep.colsum) is reduced in the epilogue
      DCPT_TRY((launch_bn<EPI_GATE_BWD, false, false>(g, stream)));
      if (g.ep.colsum) {
        DCPT_CHECK_ARG(g.ep.ldo == 2 * g.ep.C, DCPT_E_ARG, "gemm: the fused d(x4) column sums need a dense output");
        return colsum_bf16_launch(g.ep.out_bf16, g.ep.colsum, g.M, 2 * g.ep.C, stream);
      }
      return 0;
    }
    case EPI_PIXSHUF: {
      // TMA-tiled form: the 32-row slab x 32 channels of one sub-pixel is a strided 5-D box of the NHWC output
      static const bool no_tma = getenv("DCPT_GEMM_NO_TMA_EPI") != nullptr;
      const int W = g.ep.W;
      const bool pow2 = W > 0 && (W & (W - 1)) == 0;
      auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
      if (!no_tma && pow2 && W >= 8 && g.ep.Cseg % 32 == 0 && g.N == 4 * g.ep.Cseg && g.M % 32 == 0 && g.M % W == 0 && !g.ep.bias &&
          (g.ep.out_f32 || g.ep.out_bf16) && al(g.ep.out_f32) && al(g.ep.out_bf16) && al(g.ep.resid)) {
        GemmArgs gg = g;
        gg.epi = EPI_STORE;  // (only for the profile tag) the kernel keys on ep.Cseg > 0
        return launch_bn<EPI_STORE_TMA, false, false>(gg, stream);
      }
      return launch_bn<EPI_PIXSHUF, false, false>(g, stream);
    }
    case EPI_ATOMIC: return launch_bn<EPI_ATOMIC, false, false>(g, stream);
  }
  dcpt_set_error("gemm: unknown epilogue %d", g.epi);
  return DCPT_E_ARG;
}
