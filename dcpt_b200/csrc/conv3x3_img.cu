// The two 3x3 convolutions that touch the 3-channel image (reference: NAFNetBaseline.intro /
// .ending, nafnet_arch.py:202-219, :252, :271-272) and their backward.  With 3 channels on one
// side these are 27-term stencils: far too thin for tensor cores (K = 27 or N = 3) and bound by
// the HBM traffic of the `width`-channel feature map, so they are direct CUDA-core kernels.
//   img  : fp32 NCHW [N, 3, H, W]   (what the reference's callers hand to the network)
//   feat : fp32 NHWC [N*H*W, C]     (our residual-stream layout)
#include "elementwise.cuh"

namespace {

// ---- img(3) -> feat(C): intro forward, and ending dgrad (transpose_flip = 1) ----
// out[px][co] = bias[co] + sum_{ci,ky,kx} Wk[co][ci][ky][kx] * img[ci][h+ky-1][w+kx-1]
//   transpose_flip = 0: Wk[co][ci][ky][kx] = w[co][ci][ky][kx]          (w is [C][3][3][3])
//   transpose_flip = 1: Wk[co][ci][ky][kx] = w[ci][co][2-ky][2-kx]      (w is [3][C][3][3])
// Thread = (pixel, 8 output channels).
__global__ void __launch_bounds__(256)
conv3x3_img_to_feat_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                           int transpose_flip, float* __restrict__ out_f32, bf16* __restrict__ out_bf16,
                           float* __restrict__ colsum, int N, int H, int W, int C) {
  extern __shared__ float smem[];
  float* s_w = smem;           // [27][C]
  float* s_b = s_w + 27 * C;   // [C]
  float* s_cs = s_b + C;       // [C]
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) {
    const int j = i / C, co = i - j * C;  // j = ci*9 + ky*3 + kx
    const int ci = j / 9, t = j % 9;
    s_w[i] = transpose_flip ? w[((size_t)ci * C + co) * 9 + (8 - t)] : w[(size_t)co * 27 + j];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_b[i] = bias ? bias[i] : 0.f;
    s_cs[i] = 0.f;
  }
  __syncthreads();
  const int CV = C >> 3;
  const long long HW = (long long)H * W;
  const long long total = (long long)N * HW * CV;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int my_cv = -1;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int cv = (int)(idx % CV);
    const long long px = idx / CV;
    const int n = (int)(px / HW);
    const int rem = (int)(px - (long long)n * HW);
    const int h = rem / W, wq = rem - h * W;
    my_cv = cv;  // stride is a multiple of CV (host guarantees) -> cv is fixed per thread
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = s_b[cv * 8 + i];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = img + ((size_t)n * 3 + ci) * HW;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hh = h + t / 3 - 1, ww = wq + t % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        const float v = __ldg(plane + (size_t)hh * W + ww);
        const float* wp = s_w + (ci * 9 + t) * C + cv * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wp);
        const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
        acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
      }
    }
    const size_t off = (size_t)px * C + cv * 8;
    if (out_f32) {
      *reinterpret_cast<float4*>(out_f32 + off) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(out_f32 + off + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (out_bf16) stg16(out_bf16 + off, pack8(acc));
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[i] += acc[i];
  }
  if (colsum) {
    if (my_cv >= 0)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s_cs[my_cv * 8 + i], cs[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(colsum + i, s_cs[i]);
  }
}

// ---- feat(C) -> img(3): ending forward (+ bias + residual image) ----
// out[n][o][h][w] = bias[o] + resid[n][o][h][w] + sum_{c,ky,kx} w[o][c][ky][kx] * feat[(h+ky-1, w+kx-1)][c]
// One warp per output pixel: lanes split the channels, shuffle-reduce the three sums.
__global__ void __launch_bounds__(256)
conv3x3_feat_to_img_kernel(const float* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias,
                           const float* __restrict__ resid, float* __restrict__ out, int N, int H, int W, int C) {
  extern __shared__ float s_w[];  // [3][9][C]
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) {
    const int o = i / (9 * C), r = i - o * 9 * C;
    const int t = r / C, c = r - t * C;
    s_w[i] = w[((size_t)o * C + c) * 9 + t];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long HW = (long long)H * W;
  const long long total = (long long)N * HW;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long px = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); px < total; px += wstride) {
    const int n = (int)(px / HW);
    const int rem = (int)(px - (long long)n * HW);
    const int h = rem / W, wq = rem - h * W;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = wq + t % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const float* fp = feat + ((size_t)n * HW + (size_t)hh * W + ww) * C;
      for (int c = lane * 4; c < C; c += 128) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(fp + c));
        const float4 w0 = *reinterpret_cast<const float4*>(s_w + (0 * 9 + t) * C + c);
        const float4 w1 = *reinterpret_cast<const float4*>(s_w + (1 * 9 + t) * C + c);
        const float4 w2 = *reinterpret_cast<const float4*>(s_w + (2 * 9 + t) * C + c);
        a0 += x.x * w0.x + x.y * w0.y + x.z * w0.z + x.w * w0.w;
        a1 += x.x * w1.x + x.y * w1.y + x.z * w1.z + x.w * w1.w;
        a2 += x.x * w2.x + x.y * w2.y + x.z * w2.z + x.w * w2.w;
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane < 3) {
      const float a = lane == 0 ? a0 : (lane == 1 ? a1 : a2);
      const size_t o = ((size_t)n * 3 + lane) * HW + rem;
      out[o] = a + (bias ? bias[lane] : 0.f) + (resid ? resid[o] : 0.f);
    }
  }
}

// ---- wgrad of both: G[c][j] += sum_px feat[px][c] * patch_j(px), j = ci*9 + t ----
//   flip = 0: patch_j(px) = img[ci][h+ky-1][w+kx-1]   (intro:  dW[c][ci][t] = G[c][ci*9+t], feat = d(intro out))
//   flip = 1: patch_j(px) = img[ci][h-ky+1][w-kx+1]   (ending: dW[o][c][t]  = G[c][o*9+t],  feat = ending input, img = dout)
// Block = 64-pixel chunks staged in shared memory; thread t owns channel c = t % C (C <= 256) and a
// slice of the 27 patch columns.  The ending conv's bias grad (plane sums of dout) is plane_sum_kernel.
__global__ void __launch_bounds__(256)
conv3x3_small_wgrad_kernel(const float* __restrict__ feat, const float* __restrict__ img, float* __restrict__ G, int flip,
                           int N, int H, int W, int C, int chunks_per_block) {
  extern __shared__ float smem[];
  constexpr int P = 64;             // pixels per chunk
  float* s_f = smem;                // [P][C]
  float* s_p = s_f + P * C;         // [P][28]
  const long long HW = (long long)H * W;
  const long long total = (long long)N * HW;
  const int groups = blockDim.x / C > 0 ? blockDim.x / C : 1;  // column groups (C <= blockDim.x)
  const int c = threadIdx.x % C;
  const int grp = threadIdx.x / C;
  const int jper = (27 + groups - 1) / groups;
  const int j0 = grp * jper;
  float acc[27];
#pragma unroll
  for (int j = 0; j < 27; ++j) acc[j] = 0.f;
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const long long base = ((long long)blockIdx.x * chunks_per_block + ch) * P;
    if (base >= total) break;
    __syncthreads();
    for (int i = threadIdx.x; i < P * (C / 4); i += blockDim.x) {
      const int p = i / (C / 4), v = i - p * (C / 4);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base + p < total) x = __ldg(reinterpret_cast<const float4*>(feat + (size_t)(base + p) * C) + v);
      *reinterpret_cast<float4*>(s_f + p * C + v * 4) = x;
    }
    for (int i = threadIdx.x; i < P * 27; i += blockDim.x) {
      const int p = i / 27, j = i - p * 27;
      float v = 0.f;
      const long long px = base + p;
      if (px < total) {
        const int n = (int)(px / HW);
        const int rem = (int)(px - (long long)n * HW);
        const int h = rem / W, wq = rem - h * W;
        const int ci = j / 9, t = j % 9;
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const int hh = flip ? h - dy : h + dy, ww = flip ? wq - dx : wq + dx;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(img + ((size_t)n * 3 + ci) * HW + (size_t)hh * W + ww);
      }
      s_p[p * 28 + j] = v;
    }
    __syncthreads();
    if (grp < groups && threadIdx.x < groups * C) {
      for (int p = 0; p < P; ++p) {
        const float f = s_f[p * C + c];
#pragma unroll
        for (int jj = 0; jj < 27; ++jj) {
          if (jj < jper && j0 + jj < 27) acc[jj] = fmaf(f, s_p[p * 28 + j0 + jj], acc[jj]);
        }
      }
    }
  }
  if (threadIdx.x < groups * C) {
#pragma unroll
    for (int jj = 0; jj < 27; ++jj)
      if (jj < jper && j0 + jj < 27) atomicAdd(G + (size_t)c * 27 + j0 + jj, acc[jj]);
  }
}

// sum over pixels of each image plane: out[ci] += sum_{n,h,w} img[n][ci][h][w]
__global__ void plane_sum_kernel(const float* __restrict__ img, float* __restrict__ out, int N, long long HW) {
  const int ci = blockIdx.y;
  float acc = 0.f;
  const long long total = (long long)N * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, r = i - n * HW;
    acc += __ldg(img + ((size_t)n * 3 + ci) * HW + r);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + ci, acc);
}

}  // namespace

int conv3x3_img_to_feat_launch(const float* img, const float* w, const float* bias, int transpose_flip, float* out_f32,
                               bf16* out_bf16, float* colsum, int N, int H, int W, int C, cudaStream_t st) {
  DCPT_CHECK_ARG(C % 8 == 0 && C >= 8, DCPT_E_SHAPE, "conv3x3 img->feat: C=%d must be a multiple of 8", C);
  const int CV = C / 8;
  const size_t smem = ((size_t)27 * C + 2 * C) * sizeof(float);
  if (smem > 48 * 1024)
    DCPT_CUDA(cudaFuncSetAttribute(conv3x3_img_to_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)N * H * W * CV;
  long long blocks = ceil_div_ll(total, 256);
  const long long cap = (long long)dcpt_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  // keep gridDim.x*blockDim.x a multiple of CV so each thread keeps one channel vector (column sums stay per-thread)
  while ((blocks * 256) % CV != 0) ++blocks;
  DCPT_PROF("conv3x3_img_to_feat", 54.0 * N * H * W * C, (12.0 + (out_f32 ? 4.0 : 0.0) * C + (out_bf16 ? 2.0 : 0.0) * C) * N * H * W, st);
  conv3x3_img_to_feat_kernel<<<(unsigned)blocks, 256, smem, st>>>(img, w, bias, transpose_flip, out_f32, out_bf16, colsum, N, H,
                                                                 W, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int conv3x3_feat_to_img_launch(const float* feat, const float* w, const float* bias, const float* resid_img, float* out_img,
                               int N, int H, int W, int C, cudaStream_t st) {
  DCPT_CHECK_ARG(C % 4 == 0, DCPT_E_SHAPE, "conv3x3 feat->img: C=%d must be a multiple of 4", C);
  const size_t smem = (size_t)27 * C * sizeof(float);
  if (smem > 48 * 1024)
    DCPT_CUDA(cudaFuncSetAttribute(conv3x3_feat_to_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)N * H * W;
  long long blocks = ceil_div_ll(total, 8);
  const long long cap = (long long)dcpt_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  DCPT_PROF("conv3x3_feat_to_img", 54.0 * N * H * W * C, (4.0 * C + 24.0) * N * H * W, st);
  conv3x3_feat_to_img_kernel<<<(unsigned)blocks, 256, smem, st>>>(feat, w, bias, resid_img, out_img, N, H, W, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int conv3x3_small_wgrad_launch(const float* feat, const float* img, float* G, float* img_sum, int flip, int N, int H, int W,
                               int C, cudaStream_t st) {
  DCPT_CHECK_ARG(C % 4 == 0 && C <= 256, DCPT_E_SHAPE, "conv3x3 small wgrad: need C %% 4 == 0 and C <= 256 (C=%d)", C);
  const size_t smem = ((size_t)64 * C + 64 * 28) * sizeof(float);
  if (smem > 48 * 1024)
    DCPT_CUDA(cudaFuncSetAttribute(conv3x3_small_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)N * H * W;
  const long long chunks = ceil_div_ll(total, 64);
  int cpb = (int)ceil_div_ll(chunks, (long long)dcpt_num_sms() * 4);
  if (cpb < 1) cpb = 1;
  const long long blocks = ceil_div_ll(chunks, cpb);
  DCPT_PROF("conv3x3_small_wgrad", 54.0 * N * H * W * C, (4.0 * C + 12.0) * N * H * W, st);
  conv3x3_small_wgrad_kernel<<<(unsigned)blocks, 256, smem, st>>>(feat, img, G, flip, N, H, W, C, cpb);
  DCPT_LAUNCH_CHECK();
  if (img_sum) {
    dim3 grid(64, 3);
    plane_sum_kernel<<<grid, 256, 0, st>>>(img, img_sum, N, (long long)H * W);
    DCPT_LAUNCH_CHECK();
  }
  return 0;
}

// =====================================================================================================
// Tensor-core route for the two image-side 3x3 convs (used when the feature width allows it):
// the 27-term stencil of the 3-channel image is materialised once as a bf16 "patch" matrix
// P[px][32] (27 taps + 5 zero columns, 64 bytes per pixel -- small next to the 4*C bytes per pixel of the
// feature map), after which
//   intro forward          x0   = P  * Wi^T + b          (GEMM  M = pixels, N = C, K = 32)
//   intro wgrad            dWi  = dX0^T * P              (wgrad GEMM, K = pixels)
//   ending dgrad           dF   = Pd * Wd^T              (Pd = flipped patches of dout)
//   ending wgrad           dWe  = F^T * Pd
// all run on the tcgen05 engine of gemm_sm100.cu.  Patch column j = ci*9 + ky*3 + kx holds
//   flip = 0: img[ci][h+ky-1][w+kx-1]        flip = 1: img[ci][h-ky+1][w-kx+1]      (zero outside the image)
// =====================================================================================================
namespace {

__global__ void __launch_bounds__(256)
im2col3_kernel(const float* __restrict__ img, bf16* __restrict__ P, float* __restrict__ colsum, int flip, int dup, int N, int H, int W) {
  __shared__ float s_sum[32];
  if (threadIdx.x < 32) s_sum[threadIdx.x] = 0.f;
  __syncthreads();
  const long long HW = (long long)H * W, total = (long long)N * HW;
  float loc[27];
#pragma unroll
  for (int j = 0; j < 27; ++j) loc[j] = 0.f;
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < total; px += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(px / HW);
    const int rem = (int)(px - (long long)n * HW);
    const int h = rem / W, w = rem - h * W;
    float v[32];
#pragma unroll
    for (int j = 27; j < 32; ++j) v[j] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = img + ((size_t)n * 3 + ci) * HW;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const int hh = flip ? h - dy : h + dy, ww = flip ? w - dx : w + dx;
        float x = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) x = __ldg(plane + (size_t)hh * W + ww);
        x = bf16_round(x);
        v[ci * 9 + t] = x;
        loc[ci * 9 + t] += x;
      }
    }
    // dup: row = [patches | patches] (64 columns): the intro GEMM multiplies the copy by the low half of a
    // hi/lo-split weight, which restores fp32-grade weights on the image-side conv at no extra MMA cost.
    uint4* dst = reinterpret_cast<uint4*>(P + (size_t)px * (dup ? 64 : 32));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float f[8] = {v[q * 8], v[q * 8 + 1], v[q * 8 + 2], v[q * 8 + 3], v[q * 8 + 4], v[q * 8 + 5], v[q * 8 + 6], v[q * 8 + 7]};
      const uint4 pk = pack8(f);
      dst[q] = pk;
      if (dup) dst[4 + q] = pk;
    }
  }
  if (colsum) {
#pragma unroll
    for (int j = 0; j < 27; ++j) {
      const float s = warp_sum(loc[j]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum[j], s);
    }
    __syncthreads();
    if (threadIdx.x < 27) atomicAdd(colsum + threadIdx.x, s_sum[threadIdx.x]);
  }
}

// mode 0: out [C][64] = [hi | lo] split of the intro weight w[r][j] ([C][27]): hi = bf16(w), lo = bf16(w - hi)
// mode 1: out [C][32],  out[r][o*9+t] = w[o][r][t] (ending weight [3][C][9]) -- dgrad operand
__global__ void pack_w27_kernel(const float* __restrict__ w, bf16* __restrict__ out, int C, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (mode == 0) {
    if (i >= C * 64) return;
    const int r = i >> 6, jj = i & 63, j = jj & 31;
    float v = 0.f;
    if (j < 27) {
      const float wv = w[(size_t)r * 27 + j];
      const float hi = bf16_round(wv);
      v = jj < 32 ? hi : wv - hi;
    }
    out[i] = OP_FROM_F32(v);
  } else {
    if (i >= C * 32) return;
    const int r = i >> 5, j = i & 31;
    float v = 0.f;
    if (j < 27) v = w[((size_t)(j / 9) * C + r) * 9 + (j % 9)];
    out[i] = OP_FROM_F32(v);
  }
}

// mode 0 (intro):  dw[c][j] += G[c*32 + j]                      (dw is [C][27])
// mode 1 (ending): dw[o][c][t] += G[c*32 + o*9 + t]; db[o] += psum[o*9 + 4]  (centre tap = plain sum of dout[o])
__global__ void finish_w27_kernel(const float* __restrict__ G, float* __restrict__ dw, const float* __restrict__ psum,
                                  float* __restrict__ db, int C, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * 27) {
    const int c = i / 27, j = i - c * 27;
    if (mode == 0) dw[i] += G[c * 32 + j];
    else dw[((size_t)(j / 9) * C + c) * 9 + (j % 9)] += G[c * 32 + j];
  } else if (mode == 1 && i < C * 27 + 3) {
    const int o = i - C * 27;
    db[o] += psum[o * 9 + 4];
  }
}

// colsum[c] += sum_j w[o][c][t] * psum[o*9+t]   (column sums of dF = Pd * Wd^T without touching dF)
__global__ void ending_colsum_kernel(const float* __restrict__ w, const float* __restrict__ psum, float* __restrict__ colsum, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int j = 0; j < 27; ++j) acc = fmaf(w[((size_t)(j / 9) * C + c) * 9 + (j % 9)], psum[j], acc);
  colsum[c] += acc;
}

// ---- ending forward from the bf16 mirror of the last feature map: TMA tile (8x16 px + halo) x 64 channels ----
constexpr int ETH = 8, ETW = 16, EHH = ETH + 2, EHW = ETW + 2;
constexpr int EBOX_BYTES = EHH * EHW * 64 * 2;

__global__ void __launch_bounds__(256)
ending_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmF, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ resid, float* __restrict__ out, int N, int H, int W, int C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c2 = lane * 2;
  const bool chan_ok = c2 < C;
  const int c0 = chan_ok ? c2 : 0;
  float2 wt[3][9];  // taps of this lane's two channels for the 3 outputs
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int t = 0; t < 9; ++t)
      wt[o][t] = chan_ok ? make_float2(__ldg(w + ((size_t)o * C + c0) * 9 + t), __ldg(w + ((size_t)o * C + c0 + 1) * 9 + t))
                         : make_float2(0.f, 0.f);
  const float b0 = __ldg(bias), b1 = __ldg(bias + 1), b2 = __ldg(bias + 2);
  const int tiles_w = (W + ETW - 1) / ETW, tiles_h = (H + ETH - 1) / ETH, per_img = tiles_w * tiles_h, total = N * per_img;
  const int chunk = (total + gridDim.x - 1) / gridDim.x;
  const int t0 = min(total, (int)blockIdx.x * chunk), t1 = min(total, t0 + chunk);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmF);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int t, int s) {
    const int n = t / per_img, r = t - n * per_img;
    mbar_arrive_expect_tx(&full[s], EBOX_BYTES);
    tma_load_4d(smem + (size_t)s * EBOX_BYTES, &tmF, &full[s], 0, (r % tiles_w) * ETW - 1, (r / tiles_w) * ETH - 1, n);
  };
  if (threadIdx.x == 0 && t0 < t1) issue(t0, 0);
  const size_t HW = (size_t)H * W;
  int it = 0;
  for (int t = t0; t < t1; ++t, ++it) {
    const int s = it & 1;
    if (threadIdx.x == 0 && t + 1 < t1) issue(t + 1, s ^ 1);
    const int n = t / per_img, r = t - n * per_img;
    const int h = (r / tiles_w) * ETH + warp, w0 = (r % tiles_w) * ETW;
    mbar_wait(&full[s], (it >> 1) & 1);
    const bf16* sF = reinterpret_cast<const bf16*>(smem + (size_t)s * EBOX_BYTES) + lane * 2;
    float2 A[3][3];
    float keep0 = 0.f, keep1 = 0.f, keep2 = 0.f;  // lane ox keeps pixel ox's three outputs
#pragma unroll
    for (int x = 0; x < EHW; ++x) {
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
        A[rr][x % 3] = OP2_TO_F32(*reinterpret_cast<const op16x2*>(sF + ((warp + rr) * EHW + x) * 64));
      if (x >= 2) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const float2 v = A[rr][(x - 2 + d) % 3];
            const int tp = rr * 3 + d;
            a0 = fmaf(v.x, wt[0][tp].x, fmaf(v.y, wt[0][tp].y, a0));
            a1 = fmaf(v.x, wt[1][tp].x, fmaf(v.y, wt[1][tp].y, a1));
            a2 = fmaf(v.x, wt[2][tp].x, fmaf(v.y, wt[2][tp].y, a2));
          }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == x - 2) {
          keep0 = a0; keep1 = a1; keep2 = a2;
        }
      }
    }
    if (lane < ETW && h < H && w0 + lane < W) {
      const size_t o = (size_t)n * 3 * HW + (size_t)h * W + w0 + lane;
      out[o] = keep0 + b0 + (resid ? resid[o] : 0.f);
      out[o + HW] = keep1 + b1 + (resid ? resid[o + HW] : 0.f);
      out[o + 2 * HW] = keep2 + b2 + (resid ? resid[o + 2 * HW] : 0.f);
    }
    __syncthreads();
  }
}

}  // namespace

int im2col3_launch(const float* img, bf16* P, float* colsum, int flip, int dup, int N, int H, int W, cudaStream_t st) {
  const long long total = (long long)N * H * W;
  long long blocks = ceil_div_ll(total, 256);
  const long long cap = (long long)dcpt_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  DCPT_PROF("im2col3", 0.0, 76.0 * total, st);
  im2col3_kernel<<<(unsigned)blocks, 256, 0, st>>>(img, P, colsum, flip, dup, N, H, W);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int pack_w27_launch(const float* w, bf16* out, int C, int mode, cudaStream_t st) {
  DCPT_PROF("pack_w27", 0.0, 200.0 * C, st);
  pack_w27_kernel<<<ceil_div(C * 64, 256), 256, 0, st>>>(w, out, C, mode);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int finish_w27_launch(const float* G, float* dw, const float* psum, float* db, int C, int mode, cudaStream_t st) {
  DCPT_PROF("finish_w27", 0.0, 240.0 * C, st);
  finish_w27_kernel<<<ceil_div(C * 27 + 3, 256), 256, 0, st>>>(G, dw, psum, db, C, mode);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int ending_colsum_launch(const float* w, const float* psum, float* colsum, int C, cudaStream_t st) {
  DCPT_PROF("ending_colsum", 0.0, 120.0 * C, st);
  ending_colsum_kernel<<<ceil_div(C, 128), 128, 0, st>>>(w, psum, colsum, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int ending_fwd_tma_launch(const bf16* feat, const float* w, const float* bias, const float* resid_img, float* out_img, int N, int H,
                          int W, int C, cudaStream_t st) {
  DCPT_CHECK_ARG(C % 8 == 0 && C <= 64, DCPT_E_SHAPE, "ending (TMA path): C=%d must be a multiple of 8 and <= 64", C);
  CUtensorMap tmF;
  DCPT_TRY(make_tmap_nhwc(&tmF, feat, N, H, W, C, EHW, EHH));
  const size_t smem = 128 + (size_t)2 * EBOX_BYTES;
  static bool attr = false;
  if (!attr) {
    DCPT_CUDA(cudaFuncSetAttribute(ending_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int tiles = N * ceil_div(H, ETH) * ceil_div(W, ETW);
  int grid = dcpt_num_sms() * 3;
  if (grid > tiles) grid = tiles;
  DCPT_PROF("ending_fwd_tma", 54.0 * N * H * W * C, (2.0 * C + 24.0) * N * H * W, st);
  ending_fwd_tma_kernel<<<grid, 256, smem, st>>>(tmF, w, bias, resid_img, out_img, N, H, W, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}
