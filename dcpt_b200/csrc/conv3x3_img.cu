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
      out[o] = a + bias[lane] + (resid ? resid[o] : 0.f);
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
