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
};

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
};

int gemm_tc_launch(const GemmArgs& g, cudaStream_t stream);    // tcgen05 + TMA (product path)
int gemm_simt_launch(const GemmArgs& g, cudaStream_t stream);  // CUDA-core cross-check (tests only)
int gemm_launch(const GemmArgs& g, cudaStream_t stream);       // dispatch (DCPT_GEMM_SIMT=1 selects simt)

#ifdef __CUDACC__
// Epilogue for one row `m` and 32 consecutive accumulator columns [n0, n0+32).
// Column validity is checked in groups of 8 (every N on this path is a multiple of 8).
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& p, int m, int n0, int N, const float (&acc)[32]) {
  if constexpr (EPI == EPI_STORE) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n = n0 + g * 8;
      if (n < N) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[g * 8 + i];
        if (p.bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        if (p.resid) {
          const float* r = p.resid + (size_t)m * p.ldr + n;
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(r));
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(r + 4));
          v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
          v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
        if (p.out_f32) {
          float* o = p.out_f32 + (size_t)m * p.ldo + n;
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (p.out_bf16) stg16(p.out_bf16 + (size_t)m * p.ldo + n, pack8(v));
      }
    }
  } else if constexpr (EPI == EPI_GATE) {
    // packed column p = 16*pp + h*8 + i  <->  channel h*C + 8*pp + i   (h = 0: first half, 1: second half)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int n = n0 + g * 16;
      if (n < N) {
        float a[8], b[8], s[8];
        const float4 ba0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        const float4 ba1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
        const float4 bb0 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 8));
        const float4 bb1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 12));
        const float ba[8] = {ba0.x, ba0.y, ba0.z, ba0.w, ba1.x, ba1.y, ba1.z, ba1.w};
        const float bb[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i] = bf16_round(acc[g * 16 + i] + ba[i]);
          b[i] = bf16_round(acc[g * 16 + 8 + i] + bb[i]);
          s[i] = a[i] * b[i];
        }
        const int j = (n >> 4) * 8;
        bf16* x4 = p.out_bf16 + (size_t)m * p.ldo;
        stg16(x4 + j, pack8(a));
        stg16(x4 + p.C + j, pack8(b));
        stg16(p.out2 + (size_t)m * p.ldo2 + j, pack8(s));
      }
    }
  } else if constexpr (EPI == EPI_GATE_BWD) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n = n0 + g * 8;
      if (n < N) {
        const bf16* x4 = p.aux + (size_t)m * p.ldaux;
        float a[8], b[8], da[8], db[8];
        unpack8(ldg16(x4 + n), a);
        unpack8(ldg16(x4 + p.C + n), b);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = acc[g * 8 + i];
          da[i] = d * b[i];
          db[i] = d * a[i];
        }
        bf16* o = p.out_bf16 + (size_t)m * p.ldo;
        stg16(o + n, pack8(da));
        stg16(o + p.C + n, pack8(db));
      }
    }
  } else if constexpr (EPI == EPI_PIXSHUF) {
    const int w = m % p.W;
    const int t = m / p.W;
    const int h = t % p.H;
    const int img = t / p.H;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n = n0 + g * 8;
      if (n < N) {
        const int q = n / p.Cseg;
        const int c = n - q * p.Cseg;
        const size_t pix = ((size_t)img * (2 * p.H) + 2 * h + (q >> 1)) * (size_t)(2 * p.W) + 2 * w + (q & 1);
        const size_t off = pix * p.Cseg + c;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[g * 8 + i];
        if (p.resid) {
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(p.resid + off));
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(p.resid + off + 4));
          v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
          v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
        if (p.out_f32) {
          *reinterpret_cast<float4*>(p.out_f32 + off) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(p.out_f32 + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (p.out_bf16) stg16(p.out_bf16 + off, pack8(v));
      }
    }
  } else if constexpr (EPI == EPI_ATOMIC) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int n = n0 + g * 4;
      if (n < N) {
        float* o = p.out_f32 + (size_t)m * p.ldo + n;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(acc[g * 4 + 0]), "f"(acc[g * 4 + 1]),
                     "f"(acc[g * 4 + 2]), "f"(acc[g * 4 + 3])
                     : "memory");
      }
    }
  }
}
#endif
