// CUDA-core GEMM with the same operand/epilogue contract as gemm_sm100.cu.
// NOT on the product path: it exists so tests can cross-check the tcgen05 kernel
// (and bring the rest of the pipeline up independently of it).  Selected only when
// the environment variable DCPT_GEMM_SIMT=1 is set, or through dcpt_gemm_bf16(..., impl=1).
#include <stdlib.h>

#include "elementwise.cuh"
#include "gemm.cuh"

namespace {

template <int EPI>
__global__ void gemm_simt_kernel(GemmArgs g, int k0, int k1, int chunks_n) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = (int)(idx / chunks_n);
  const int n0 = (int)(idx % chunks_n) * 32;
  if (m >= g.M) return;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int k = k0; k < k1; ++k) {
    const float a = OP_TO_F32(g.a_mn ? g.A[(size_t)k * g.lda + m] : g.A[(size_t)m * g.lda + k]);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + j;
      if (n < g.N) {
        const float b = OP_TO_F32(g.b_mn ? g.B[(size_t)k * g.ldb + n] : g.B[(size_t)n * g.ldb + k]);
        acc[j] = fmaf(a, b, acc[j]);
      }
    }
  }
  epilogue_chunk<EPI>(g.ep, m, n0, g.N, acc);
}

template <int EPI>
int launch(const GemmArgs& g, cudaStream_t stream) {
  const int chunks_n = ceil_div(g.N, 32);
  const long long total = (long long)g.M * chunks_n;
  const int splits = g.splits < 1 ? 1 : g.splits;
  const int kps = ceil_div(g.K, splits);
  DCPT_PROF("gemm_simt", 2.0 * g.M * g.N * g.K, 0.0, stream);
  for (int s = 0; s < splits; ++s) {
    const int k0 = s * kps, k1 = (k0 + kps < g.K) ? k0 + kps : g.K;
    if (k0 >= k1) break;
    gemm_simt_kernel<EPI><<<(unsigned)ceil_div_ll(total, 128), 128, 0, stream>>>(g, k0, k1, chunks_n);
    DCPT_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace

int gemm_simt_launch(const GemmArgs& g, cudaStream_t stream) {
  DCPT_CHECK_ARG(g.m_per_batch == 0 && g.k_per_batch == 0, DCPT_E_UNSUPPORTED, "gemm(simt): batched forms are tensor-core only");
  DCPT_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0 && g.N % 8 == 0, DCPT_E_SHAPE, "gemm(simt): bad shape M=%d N=%d K=%d", g.M, g.N,
                 g.K);
  DCPT_CHECK_ARG(g.splits <= 1 || g.epi == EPI_ATOMIC, DCPT_E_ARG, "gemm(simt): split-K needs the atomic epilogue");
  switch (g.epi) {
    case EPI_STORE:
      DCPT_TRY(launch<EPI_STORE>(g, stream));
      if (g.ep.gaux && g.ep.colsum)  // the tensor-core kernel reduces this in its epilogue
        return sca_ds_reduce_launch(g.ep.out_bf16, g.ep.gaux, g.ep.colsum, g.M / g.ep.rows_per_img, g.ep.rows_per_img, g.N, stream);
      return 0;
    case EPI_GATE: return launch<EPI_GATE>(g, stream);
    case EPI_GATE_BWD:
      DCPT_TRY(launch<EPI_GATE_BWD>(g, stream));
      if (g.ep.colsum) return colsum_bf16_launch(g.ep.out_bf16, g.ep.colsum, g.M, 2 * g.ep.C, stream);
      return 0;
    case EPI_PIXSHUF: return launch<EPI_PIXSHUF>(g, stream);
    case EPI_ATOMIC: return launch<EPI_ATOMIC>(g, stream);
  }
  dcpt_set_error("gemm(simt): unknown epilogue %d", g.epi);
  return DCPT_E_ARG;
}

int gemm_launch(const GemmArgs& g, cudaStream_t stream) {
  static int simt = -1;
  if (simt < 0) {
    const char* e = getenv("DCPT_GEMM_SIMT");
    simt = (e && e[0] == '1') ? 1 : 0;
  }
  return simt ? gemm_simt_launch(g, stream) : gemm_tc_launch(g, stream);
}
