// Depthwise 3x3 convolution + SimpleGate (forward / backward), NHWC bf16.
// Reference: NAFBlock.conv2 (nafnet_arch.py:96-104, groups = 2C, padding 1, bias) followed by
// SimpleGate (nafnet_arch.py:77-80): g[:, j] = v[:, j] * v[:, j + C].
// CUDA-core by design (36 FLOP/byte-pair at most; HBM/L1-bound).  Each thread owns one
// 8-channel vector (16 bytes) — for the gate, the vector j and its partner j+C — and 4
// neighbouring pixels along W, so each 16-byte load feeds up to three taps.
#include "elementwise.cuh"

namespace {

constexpr int PX = 4;  // pixels along W per thread

// weights in smem as [tap][channel] fp32, bias [channel]
__device__ __forceinline__ void load_dw_weights(float* s_w, float* s_b, const float* __restrict__ w, const float* __restrict__ b,
                                                int CH) {
  for (int i = threadIdx.x; i < CH * 9; i += blockDim.x) {
    const int c = i / 9, tap = i - c * 9;
    s_w[tap * CH + c] = w[i];
  }
  if (s_b)
    for (int i = threadIdx.x; i < CH; i += blockDim.x) s_b[i] = b[i];
}

__device__ __forceinline__ void fma8(float (&acc)[8], const float (&x)[8], const float* wp) {
  const float4 w0 = *reinterpret_cast<const float4*>(wp);
  const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
  acc[0] = fmaf(x[0], w0.x, acc[0]); acc[1] = fmaf(x[1], w0.y, acc[1]);
  acc[2] = fmaf(x[2], w0.z, acc[2]); acc[3] = fmaf(x[3], w0.w, acc[3]);
  acc[4] = fmaf(x[4], w1.x, acc[4]); acc[5] = fmaf(x[5], w1.y, acc[5]);
  acc[6] = fmaf(x[6], w1.z, acc[6]); acc[7] = fmaf(x[7], w1.w, acc[7]);
}

// ------------------------------- forward --------------------------------------
__global__ void __launch_bounds__(256)
dwgate_fwd_kernel(const bf16* __restrict__ u, const float* __restrict__ w2, const float* __restrict__ b2, bf16* __restrict__ g,
                  float* __restrict__ pool, int H, int W, int C) {
  extern __shared__ float smem[];
  const int C2 = 2 * C, CV = C / 8, WQ = (W + PX - 1) / PX;
  float* s_w = smem;             // [9][2C]
  float* s_b = s_w + 9 * C2;     // [2C]
  float* s_pool = s_b + C2;      // [C]
  load_dw_weights(s_w, s_b, w2, b2, C2);
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_pool[i] = 0.f;
  __syncthreads();

  const int n = blockIdx.y;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long items = (long long)H * WQ * CV;
  if (item < items) {
    const int cv = (int)(item % CV);
    const int wq = (int)((item / CV) % WQ);
    const int h = (int)(item / ((long long)CV * WQ));
    const int w0 = wq * PX;
    const int ca = cv * 8, cb = C + cv * 8;
    float a[PX][8], b[PX][8];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[p][i] = s_b[ca + i];
        b[p][i] = s_b[cb + i];
      }
    const bf16* un = u + (size_t)n * H * W * C2;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hh = h + r - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int col = 0; col < PX + 2; ++col) {
        const int ww = w0 + col - 1;
        if (ww < 0 || ww >= W) continue;
        const bf16* up = un + ((size_t)hh * W + ww) * C2;
        float xa[8], xb[8];
        unpack8(ldg16(up + ca), xa);
        unpack8(ldg16(up + cb), xb);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const int kx = col - p;  // tap column for output pixel p
          if (kx >= 0 && kx <= 2) {
            const float* wp = s_w + (r * 3 + kx) * C2;
            fma8(a[p], xa, wp + ca);
            fma8(b[p], xb, wp + cb);
          }
        }
      }
    }
    float psum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (w0 + p < W) {
        float gv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          gv[i] = bf16_round(a[p][i] * b[p][i]);
          psum[i] += gv[i];
        }
        stg16(g + (((size_t)n * H + h) * W + w0 + p) * C + ca, pack8(gv));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&s_pool[ca + i], psum[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float v = s_pool[i];
    if (v != 0.f) atomicAdd(pool + (size_t)n * C + i, v);
  }
}

// --------------------------- backward, part a ---------------------------------
// One thread: a fixed channel-vector pair (j, j+C) and a strided set of pixels of one image.
// dg = dgs*s + t;  a,b = dwconv(u)+bias (recomputed);  du2_a = dg*b, du2_b = dg*a;
// dW2[c][tap] += du2[c] * u[px+tap][c];  db2[c] += du2[c].
__global__ void __launch_bounds__(128)
dwgate_bwd_a_kernel(const bf16* __restrict__ dgs, const float* __restrict__ s, const float* __restrict__ t,
                    const bf16* __restrict__ u, const float* __restrict__ w2, const float* __restrict__ b2,
                    bf16* __restrict__ du2, float* __restrict__ dw2, float* __restrict__ db2, int H, int W, int C, int cvb) {
  extern __shared__ float smem[];
  const int C2 = 2 * C, CV = C / 8;
  float* s_w = smem;            // [9][2C]
  float* s_b = s_w + 9 * C2;    // [2C]
  float* s_red = s_b + C2;      // [cvb][160]
  load_dw_weights(s_w, s_b, w2, b2, C2);
  for (int i = threadIdx.x; i < cvb * 160; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();

  const int n = blockIdx.z;
  const int cvl = threadIdx.x % cvb;
  const int pl = threadIdx.x / cvb;
  const int npl = blockDim.x / cvb;
  const int cv = blockIdx.y * cvb + cvl;
  const int HW = H * W;
  float accA[9][8], accB[9][8], dbA[8], dbB[8];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) accA[k][i] = accB[k][i] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) dbA[i] = dbB[i] = 0.f;

  if (cv < CV) {
    const int ca = cv * 8, cb = C + cv * 8;
    float sv[8], tv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sv[i] = s[(size_t)n * C + ca + i];
      tv[i] = t[(size_t)n * C + ca + i];
    }
    const bf16* un = u + (size_t)n * HW * C2;
    for (int px = blockIdx.x * npl + pl; px < HW; px += gridDim.x * npl) {
      const int h = px / W, w = px - h * W;
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = s_b[ca + i];
        b[i] = s_b[cb + i];
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int hh = h + k / 3 - 1, ww = w + k % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        const bf16* up = un + ((size_t)hh * W + ww) * C2;
        float xa[8], xb[8];
        unpack8(ldg16(up + ca), xa);
        unpack8(ldg16(up + cb), xb);
        fma8(a, xa, s_w + k * C2 + ca);
        fma8(b, xb, s_w + k * C2 + cb);
      }
      float dg[8], da[8], db[8];
      unpack8(ldg16(dgs + ((size_t)n * HW + px) * C + ca), dg);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = fmaf(dg[i], sv[i], tv[i]);
        da[i] = d * b[i];
        db[i] = d * a[i];
        dbA[i] += da[i];
        dbB[i] += db[i];
      }
      bf16* op = du2 + ((size_t)n * HW + px) * C2;
      stg16(op + ca, pack8(da));
      stg16(op + cb, pack8(db));
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int hh = h + k / 3 - 1, ww = w + k % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        const bf16* up = un + ((size_t)hh * W + ww) * C2;
        float xa[8], xb[8];
        unpack8(ldg16(up + ca), xa);
        unpack8(ldg16(up + cb), xb);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          accA[k][i] = fmaf(da[i], xa[i], accA[k][i]);
          accB[k][i] = fmaf(db[i], xb[i], accB[k][i]);
        }
      }
    }
    float* red = s_red + cvl * 160;
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(&red[k * 8 + i], accA[k][i]);
        atomicAdd(&red[72 + k * 8 + i], accB[k][i]);
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&red[144 + i], dbA[i]);
      atomicAdd(&red[152 + i], dbB[i]);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < cvb * 160; idx += blockDim.x) {
    const int l = idx / 160, e = idx - l * 160;
    const int cvg = blockIdx.y * cvb + l;
    if (cvg >= CV) continue;
    const float v = s_red[idx];
    if (e < 144) {
      const int half = e / 72, k = (e % 72) / 8, i = e % 8;
      const int c = half * C + cvg * 8 + i;
      atomicAdd(dw2 + (size_t)c * 9 + k, v);
    } else {
      const int half = (e - 144) / 8, i = e % 8;
      atomicAdd(db2 + half * C + cvg * 8 + i, v);
    }
  }
}

// --------------------------- backward, part b ---------------------------------
// du[px][c] = sum_{ky,kx} du2[h-(ky-1), w-(kx-1)][c] * w2[c][ky][kx]; colsum[c] += sum_px du.
__global__ void __launch_bounds__(256)
dwconv_bwd_data_kernel(const bf16* __restrict__ du2, const float* __restrict__ w2, bf16* __restrict__ du,
                       float* __restrict__ colsum, int H, int W, int CH) {
  extern __shared__ float smem[];
  const int CV = CH / 8, WQ = (W + PX - 1) / PX;
  float* s_w = smem;            // [9][CH]
  float* s_cs = s_w + 9 * CH;   // [CH]
  load_dw_weights(s_w, nullptr, w2, nullptr, CH);
  for (int i = threadIdx.x; i < CH; i += blockDim.x) s_cs[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long items = (long long)H * WQ * CV;
  if (item < items) {
    const int cv = (int)(item % CV);
    const int wq = (int)((item / CV) % WQ);
    const int h = (int)(item / ((long long)CV * WQ));
    const int w0 = wq * PX, c0 = cv * 8;
    float acc[PX][8];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
    const bf16* dn = du2 + (size_t)n * H * W * CH;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hh = h + r - 1;  // source row; uses tap ky = 2 - r  (h = hh + ky - 1)
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int col = 0; col < PX + 2; ++col) {
        const int ww = w0 + col - 1;
        if (ww < 0 || ww >= W) continue;
        float x[8];
        unpack8(ldg16(dn + ((size_t)hh * W + ww) * CH + c0), x);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const int d = col - p;  // source col offset + 1 in [0,2]  -> tap kx = 2 - d
          if (d >= 0 && d <= 2) fma8(acc[p], x, s_w + ((2 - r) * 3 + (2 - d)) * CH + c0);
        }
      }
    }
    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (w0 + p < W) {
#pragma unroll
        for (int i = 0; i < 8; ++i) cs[i] += acc[p][i];
        stg16(du + (((size_t)n * H + h) * W + w0 + p) * CH + c0, pack8(acc[p]));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&s_cs[c0 + i], cs[i]);
  }
  __syncthreads();
  if (colsum)
    for (int i = threadIdx.x; i < CH; i += blockDim.x) {
      const float v = s_cs[i];
      if (v != 0.f) atomicAdd(colsum + i, v);
    }
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
  const size_t smem = ((size_t)9 * 2 * C + 2 * C + C) * sizeof(float);
  DCPT_TRY(set_smem(dwgate_fwd_kernel, smem));
  const long long items = (long long)H * ceil_div(W, PX) * (C / 8);
  dim3 grid((unsigned)ceil_div_ll(items, 256), N);
  DCPT_PROF("dwgate_fwd", 38.0 * N * H * W * C, 6.0 * N * H * W * C, st);
  dwgate_fwd_kernel<<<grid, 256, smem, st>>>(u, w2, b2, g, pool, H, W, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int dwgate_bwd_a_launch(const bf16* dgs, const float* s, const float* t, const bf16* u, const float* w2, const float* b2,
                        bf16* du2, float* dw2, float* db2, int N, int H, int W, int C, cudaStream_t st) {
  DCPT_CHECK_ARG(C % 8 == 0 && C >= 8 && N > 0 && H > 0 && W > 0, DCPT_E_SHAPE, "dwgate_bwd_a: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const int CV = C / 8;
  int cvb = 1;
  while (cvb < CV && cvb < 32) cvb <<= 1;  // channel vectors per block (power of two <= 32)
  const int npl = 128 / cvb;
  const int HW = H * W;
  // ~32 pixels per thread keeps the partial-sum flush small without starving the SMs.
  int gx = ceil_div(HW, npl * 32);
  if (gx < 1) gx = 1;
  const size_t smem = ((size_t)9 * 2 * C + 2 * C + (size_t)cvb * 160) * sizeof(float);
  DCPT_TRY(set_smem(dwgate_bwd_a_kernel, smem));
  dim3 grid(gx, ceil_div(CV, cvb), N);
  DCPT_PROF("dwgate_bwd_a", 80.0 * N * H * W * C, 10.0 * N * H * W * C, st);
  dwgate_bwd_a_kernel<<<grid, 128, smem, st>>>(dgs, s, t, u, w2, b2, du2, dw2, db2, H, W, C, cvb);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int dwconv_bwd_data_launch(const bf16* du2, const float* w2, bf16* du, float* colsum, int N, int H, int W, int C2,
                           cudaStream_t st) {
  DCPT_CHECK_ARG(C2 % 8 == 0 && C2 >= 8, DCPT_E_SHAPE, "dwconv_bwd_data: bad channel count %d", C2);
  const size_t smem = ((size_t)9 * C2 + C2) * sizeof(float);
  DCPT_TRY(set_smem(dwconv_bwd_data_kernel, smem));
  const long long items = (long long)H * ceil_div(W, PX) * (C2 / 8);
  dim3 grid((unsigned)ceil_div_ll(items, 256), N);
  DCPT_PROF("dwconv_bwd_data", 18.0 * N * H * W * C2, 4.0 * N * H * W * C2, st);
  dwconv_bwd_data_kernel<<<grid, 256, smem, st>>>(du2, w2, du, colsum, H, W, C2);
  DCPT_LAUNCH_CHECK();
  return 0;
}
