// CUDA-core kernels of the degradation-classifier head (reference: basicsr/archs/degrad_classify_arch.py,
// PromptIR_NoImg_DC :558-641 built from BottleneckBlock :132-243 with channels-first LayerNorm :17-44).
// The head's trunk is bf16 NHWC [M = N*H*W, C]; its convolutions (1x1 and dense 3x3) run on the tcgen05 GEMM engine
// (gemm_sm100.cu, implicit-GEMM mode for the 3x3); what is here is the HBM-bound glue between them:
//   ln_act      : y = act(LN_c(x) * w + b (+ resid))      Conv2d.norm + F.relu_ (+ `out += shortcut`)   :98-103, :227-243
//   mix         : z = prev + softmax(mixing)[l] * feature                                               :633-637
//   maxpool     : y = relu(maxpool2x2(x))                 downsample_layers                             :596-602
//   meanpool_fc : logits = fc(mean_hw(x))                                                               :639-640
#include "elementwise.cuh"

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float group_sum(float v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 ld4_bf16(const bf16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = OP2_TO_F32(*reinterpret_cast<const op16x2*>(&u.x));
  const float2 b = OP2_TO_F32(*reinterpret_cast<const op16x2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4_bf16(bf16* p, float4 v) {
  op16x2 p0 = OP2_FROM_F32(v.x, v.y), p1 = OP2_FROM_F32(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&p0);
  u.y = *reinterpret_cast<uint32_t*>(&p1);
  *reinterpret_cast<uint2*>(p) = u;
}

// One row per `lpr` lanes (see layernorm.cu); NV float4 slices per lane.
template <int NV>
__global__ void __launch_bounds__(kWarps * 32)
ln_act_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                  const bf16* __restrict__ resid, bf16* __restrict__ y, float* __restrict__ stats, int M, int C, int lpr, int relu,
                  float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / lpr, sub = lane % lpr, gr = lane / lpr, nvec = C >> 2;
  const float invC = 1.f / (float)C;
  const long long row_stride = (long long)gridDim.x * kWarps * rpw;
  for (long long row = ((long long)blockIdx.x * kWarps + warp) * rpw + gr; row - gr < M; row += row_stride) {
    const bool valid = row < M;
    float4 xv[NV], rv[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = sub + i * lpr;
      xv[i] = rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid && v < nvec) {
        xv[i] = ld4_bf16(x + row * C + v * 4);
        if (resid) rv[i] = ld4_bf16(resid + row * C + v * 4);
      }
      sum += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
    }
    const float mean = group_sum(sum, lpr) * invC;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (sub + i * lpr < nvec) {
        const float a = xv[i].x - mean, bb = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
        sq += a * a + bb * bb + c * c + d * d;
      }
    }
    const float rstd = 1.f / sqrtf(group_sum(sq, lpr) * invC + eps);
    if (valid) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = sub + i * lpr;
        if (v < nvec) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + v), bv = __ldg(reinterpret_cast<const float4*>(b) + v);
          float4 o;
          o.x = (xv[i].x - mean) * rstd * wv.x + bv.x + rv[i].x;
          o.y = (xv[i].y - mean) * rstd * wv.y + bv.y + rv[i].y;
          o.z = (xv[i].z - mean) * rstd * wv.z + bv.z + rv[i].z;
          o.w = (xv[i].w - mean) * rstd * wv.w + bv.w + rv[i].w;
          if (relu) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
          }
          st4_bf16(y + row * C + v * 4, o);
        }
      }
      if (sub == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
    }
  }
}

// The incoming gradient dy and the residual-path gradient dres are fp32: LN' projects out the per-pixel mean and the
// yhat component, so rounding dy to bf16 first would be amplified by that cancellation (measured: 1.5x per block).
// g = dy * (y > 0 if relu);  dres = g;  dx = rstd * (g*w - yhat*mean_c(g*w*yhat) - mean_c(g*w));  dw += g*yhat; db += g
template <int NV>
__global__ void __launch_bounds__(kWarps * 32)
ln_act_bwd_kernel(const float* __restrict__ dy, const bf16* __restrict__ y, const bf16* __restrict__ x, const float* __restrict__ stats,
                  const float* __restrict__ w, bf16* __restrict__ dx, float* __restrict__ dres, float* __restrict__ dw,
                  float* __restrict__ db, int M, int C, int lpr, int relu) {
  extern __shared__ float s_acc[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / lpr, sub = lane % lpr, gr = lane / lpr, nvec = C >> 2;
  const float invC = 1.f / (float)C;
  float4 a_dw[NV], a_db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a_dw[i] = a_db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long row_stride = (long long)gridDim.x * kWarps * rpw;
  for (long long row = ((long long)blockIdx.x * kWarps + warp) * rpw + gr; row - gr < M; row += row_stride) {
    const bool valid = row < M;
    float2 st = make_float2(0.f, 0.f);
    if (valid) st = __ldg(reinterpret_cast<const float2*>(stats + row * 2));
    float4 g[NV], yh[NV], gw[NV];
    float sg = 0.f, sgy = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = sub + i * lpr;
      g[i] = yh[i] = gw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid && v < nvec) {
        g[i] = __ldg(reinterpret_cast<const float4*>(dy + row * C) + v);
        if (relu) {
          const float4 yo = ld4_bf16(y + row * C + v * 4);
          g[i].x = yo.x > 0.f ? g[i].x : 0.f; g[i].y = yo.y > 0.f ? g[i].y : 0.f;
          g[i].z = yo.z > 0.f ? g[i].z : 0.f; g[i].w = yo.w > 0.f ? g[i].w : 0.f;
        }
        const float4 xv = ld4_bf16(x + row * C + v * 4);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + v);
        yh[i] = make_float4((xv.x - st.x) * st.y, (xv.y - st.x) * st.y, (xv.z - st.x) * st.y, (xv.w - st.x) * st.y);
        gw[i] = make_float4(g[i].x * wv.x, g[i].y * wv.y, g[i].z * wv.z, g[i].w * wv.w);
        sg += gw[i].x + gw[i].y + gw[i].z + gw[i].w;
        sgy += gw[i].x * yh[i].x + gw[i].y * yh[i].y + gw[i].z * yh[i].z + gw[i].w * yh[i].w;
      }
    }
    const float mean_g = group_sum(sg, lpr) * invC, mean_gy = group_sum(sgy, lpr) * invC;
    if (valid) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = sub + i * lpr;
        if (v < nvec) {
          float4 o;
          o.x = st.y * (gw[i].x - yh[i].x * mean_gy - mean_g);
          o.y = st.y * (gw[i].y - yh[i].y * mean_gy - mean_g);
          o.z = st.y * (gw[i].z - yh[i].z * mean_gy - mean_g);
          o.w = st.y * (gw[i].w - yh[i].w * mean_gy - mean_g);
          st4_bf16(dx + row * C + v * 4, o);
          if (dres) *(reinterpret_cast<float4*>(dres + row * C) + v) = g[i];
          a_dw[i].x += g[i].x * yh[i].x; a_dw[i].y += g[i].y * yh[i].y; a_dw[i].z += g[i].z * yh[i].z; a_dw[i].w += g[i].w * yh[i].w;
          a_db[i].x += g[i].x; a_db[i].y += g[i].y; a_db[i].z += g[i].z; a_db[i].w += g[i].w;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = sub + i * lpr;
    if (v < nvec) {
      const int c = v * 4;
      atomicAdd(&s_acc[c + 0], a_dw[i].x); atomicAdd(&s_acc[c + 1], a_dw[i].y);
      atomicAdd(&s_acc[c + 2], a_dw[i].z); atomicAdd(&s_acc[c + 3], a_dw[i].w);
      atomicAdd(&s_acc[C + c + 0], a_db[i].x); atomicAdd(&s_acc[C + c + 1], a_db[i].y);
      atomicAdd(&s_acc[C + c + 2], a_db[i].z); atomicAdd(&s_acc[C + c + 3], a_db[i].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(dw + c, s_acc[c]);
    atomicAdd(db + c, s_acc[C + c]);
  }
}

inline int pick_lpr(int C) {
  int lpr = 2;
  while (lpr < 32 && lpr < C / 4) lpr <<= 1;
  return lpr;
}

// z = prev + mw * feat; mw is read from device memory (softmax(mixing_weights)[l] computed by the caller)
__global__ void mix_fwd_kernel(const bf16* __restrict__ prev, const float* __restrict__ feat, const float* __restrict__ mw,
                               bf16* __restrict__ z, long long nvec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const float m = __ldg(mw);
  const float4 f0 = __ldg(reinterpret_cast<const float4*>(feat) + 2 * i), f1 = __ldg(reinterpret_cast<const float4*>(feat) + 2 * i + 1);
  float v[8] = {m * f0.x, m * f0.y, m * f0.z, m * f0.w, m * f1.x, m * f1.y, m * f1.z, m * f1.w};
  if (prev) {
    float p[8];
    unpack8(ldg16(prev + i * 8), p);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += p[k];
  }
  stg16(z + i * 8, pack8(v));
}

// dfeat = mw * dz (fp32);  dmw += sum dz * feat
__global__ void __launch_bounds__(256)
mix_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ feat, const float* __restrict__ mw, float* __restrict__ dfeat,
               float* __restrict__ dmw, long long nvec) {
  const float m = __ldg(mw);
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(dz) + 2 * i), d1 = __ldg(reinterpret_cast<const float4*>(dz) + 2 * i + 1);
    const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    const float4 f0 = __ldg(reinterpret_cast<const float4*>(feat) + 2 * i), f1 = __ldg(reinterpret_cast<const float4*>(feat) + 2 * i + 1);
    acc += d[0] * f0.x + d[1] * f0.y + d[2] * f0.z + d[3] * f0.w + d[4] * f1.x + d[5] * f1.y + d[6] * f1.z + d[7] * f1.w;
    if (dfeat) {
      reinterpret_cast<float4*>(dfeat)[2 * i] = make_float4(m * d[0], m * d[1], m * d[2], m * d[3]);
      reinterpret_cast<float4*>(dfeat)[2 * i + 1] = make_float4(m * d[4], m * d[5], m * d[6], m * d[7]);
    }
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += s[k];
    atomicAdd(dmw, t);
  }
}

// y[n][h][w][c] = relu(max over the 2x2 window of x[n][2h..][2w..][c])
__global__ void maxpool2_relu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long nvec, int Ho, int Wo, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const int CV = C >> 3;
  const int cv = (int)(i % CV);
  const long long px = i / CV;
  const int wo = (int)(px % Wo);
  const long long t = px / Wo;
  const int ho = (int)(t % Ho);
  const long long n = t / Ho;
  const bf16* base = x + ((n * 2 * Ho + 2 * ho) * (2 * Wo) + 2 * wo) * (long long)C + cv * 8;
  float m[8], v[8];
  unpack8(ldg16(base), m);
  unpack8(ldg16(base + C), v);
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
  unpack8(ldg16(base + (long long)2 * Wo * C), v);
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
  unpack8(ldg16(base + (long long)2 * Wo * C + C), v);
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = fmaxf(fmaxf(m[k], v[k]), 0.f);
  stg16(y + i * 8, pack8(m));
}

// dx = dy routed to the FIRST maximal element of each window (PyTorch's MaxPool2d tie rule) when that max is > 0
__global__ void maxpool2_relu_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ dy, bf16* __restrict__ dx, long long nvec,
                                         int Ho, int Wo, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const int CV = C >> 3;
  const int cv = (int)(i % CV);
  const long long px = i / CV;
  const int wo = (int)(px % Wo);
  const long long t = px / Wo;
  const int ho = (int)(t % Ho);
  const long long n = t / Ho;
  const long long off = ((n * 2 * Ho + 2 * ho) * (2 * Wo) + 2 * wo) * (long long)C + cv * 8;
  const long long o1 = off + C, o2 = off + (long long)2 * Wo * C, o3 = o2 + C;
  float a[8], b[8], c[8], d[8];
  unpack8(ldg16(x + off), a);
  unpack8(ldg16(x + o1), b);
  unpack8(ldg16(x + o2), c);
  unpack8(ldg16(x + o3), d);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(dy) + 2 * i), g1 = __ldg(reinterpret_cast<const float4*>(dy) + 2 * i + 1);
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  float ga[8], gb[8], gc[8], gd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float m = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
    const float gg = m > 0.f ? g[k] : 0.f;
    const int which = a[k] == m ? 0 : (b[k] == m ? 1 : (c[k] == m ? 2 : 3));
    ga[k] = which == 0 ? gg : 0.f;
    gb[k] = which == 1 ? gg : 0.f;
    gc[k] = which == 2 ? gg : 0.f;
    gd[k] = which == 3 ? gg : 0.f;
  }
  stg16(dx + off, pack8(ga));
  stg16(dx + o1, pack8(gb));
  stg16(dx + o2, pack8(gc));
  stg16(dx + o3, pack8(gd));
}

// pooled[n][c] = mean_px x[n][px][c]   (one block per (n, 64-channel group); fp32 out)
__global__ void __launch_bounds__(256)
meanpool_kernel(const bf16* __restrict__ x, float* __restrict__ pooled, int HW, int C) {
  __shared__ float2 s[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.y, c2 = blockIdx.x * 64 + lane * 2;
  float2 acc = make_float2(0.f, 0.f);
  if (c2 < C)
    for (int px = warp; px < HW; px += 8) {
      const float2 v = OP2_TO_F32(__ldg(reinterpret_cast<const op16x2*>(x + ((size_t)n * HW + px) * C + c2)));
      acc.x += v.x;
      acc.y += v.y;
    }
  s[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c2 < C) {
    float2 t = make_float2(0.f, 0.f);
    for (int k = 0; k < 8; ++k) {
      t.x += s[k][lane].x;
      t.y += s[k][lane].y;
    }
    pooled[(size_t)n * C + c2] = t.x / (float)HW;
    pooled[(size_t)n * C + c2 + 1] = t.y / (float)HW;
  }
}

// logits[n][k] = b[k] + sum_c pooled[n][c] W[k][c]     (one warp per output)
__global__ void fc_fwd_kernel(const float* __restrict__ pooled, const float* __restrict__ w, const float* __restrict__ b,
                              float* __restrict__ logits, int N, int C, int K) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= N * K) return;
  const int n = gw / K, k = gw - n * K;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc = fmaf(pooled[(size_t)n * C + c], w[(size_t)k * C + c], acc);
  acc = warp_sum(acc);
  if (lane == 0) logits[gw] = acc + b[k];
}

// dW[k][c] += sum_n dl[n][k] pooled[n][c]; db[k] += sum_n dl[n][k]; dx[n][px][c] = (1/HW) sum_k dl[n][k] W[k][c]
__global__ void fc_bwd_w_kernel(const float* __restrict__ dl, const float* __restrict__ pooled, float* __restrict__ dw,
                                float* __restrict__ db, int N, int C, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * C) {
    const int k = i / C, c = i - k * C;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc = fmaf(dl[n * K + k], pooled[(size_t)n * C + c], acc);
    dw[i] += acc;
  } else if (i < K * C + K) {
    const int k = i - K * C;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc += dl[n * K + k];
    db[k] += acc;
  }
}
__global__ void meanpool_fc_bwd_x_kernel(const float* __restrict__ dl, const float* __restrict__ w, float* __restrict__ dx, int HW, int C,
                                         int K) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(dl[n * K + k], w[(size_t)k * C + c], acc);
  const float v = acc / (float)HW;
  for (int px = blockIdx.z; px < HW; px += gridDim.z) dx[((size_t)n * HW + px) * C + c] = v;
}

__global__ void add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, long long nvec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  float x[8], y[8];
  unpack8(ldg16(a + i * 8), x);
  unpack8(ldg16(b + i * 8), y);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] += y[k];
  stg16(out + i * 8, pack8(x));
}

// dense 3x3 weight [Cout][Cin][3][3] fp32 -> bf16 GEMM operands
//   fwd  : out[co][t * cin_pad + ci]  = w[co][ci][t]
//   dgrad: out[ci][t * cout_pad + co] = w[co][ci][8 - t]
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int dgrad) {
  const int R = dgrad ? Cin : Cout, Cc = dgrad ? Cout : Cin;
  const int cpad = (Cc + 63) / 64 * 64;
  const long long total = (long long)R * 9 * cpad;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int r = (int)(i / (9 * cpad));
  const int rem = (int)(i - (long long)r * 9 * cpad);
  const int t = rem / cpad, c = rem - t * cpad;
  float v = 0.f;
  if (c < Cc) v = dgrad ? w[(((size_t)c * Cin + r) * 9) + (8 - t)] : w[(((size_t)r * Cin + c) * 9) + t];
  out[i] = OP_FROM_F32(v);
}
// dw[co][ci][t] += G[co][t * cin_pad + ci]
__global__ void finish_conv3x3_kernel(const float* __restrict__ G, float* __restrict__ dw, int Cout, int Cin) {
  const int cpad = (Cin + 63) / 64 * 64;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Cin * 9) return;
  const int t = (int)(i % 9);
  const long long r = i / 9;
  const int ci = (int)(r % Cin), co = (int)(r / Cin);
  dw[i] += G[(size_t)co * 9 * cpad + (size_t)t * cpad + ci];
}

}  // namespace

// im2col of conv_embed's 7x7 stride-2 pad-3 convolution on the 3-channel image (PromptIR_DC, degrad_classify_arch.py:497-500):
// P[(n, ho, wo)][ci * 49 + ky * 7 + kx] = img[n][ci][2 ho + ky - 3][2 wo + kx - 3] (zero outside), columns 147..159 zero, so the
// conv is the GEMM P [M, 160] x W [f0, 160]^T with W = weight.view(f0, 147) zero-padded - same column order as the parameter.
__global__ void __launch_bounds__(256)
im2col7s2_kernel(const float* __restrict__ img, bf16* __restrict__ P, int N, int H, int W, int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * 20;  // 20 vectors of 8 columns per output pixel
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % 20);
    const long long px = idx / 20;
    const int wo = (int)(px % Wo), ho = (int)((px / Wo) % Ho), n = (int)(px / ((long long)Wo * Ho));
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int col = v * 8 + k;
      float x = 0.f;
      if (col < 147) {
        const int ci = col / 49, r = col - ci * 49, ky = r / 7, kx = r - ky * 7;
        const int y = 2 * ho + ky - 3, xx = 2 * wo + kx - 3;
        if (y >= 0 && y < H && xx >= 0 && xx < W) x = __ldg(img + (((size_t)n * 3 + ci) * H + y) * W + xx);
      }
      f[k] = x;
    }
    stg16(P + px * 160 + v * 8, pack8(f));
  }
}

int ln_act_fwd_launch(const bf16* x, const float* w, const float* b, const bf16* resid, bf16* y, float* stats, int M, int C, int relu,
                      float eps, cudaStream_t st) {
  DCPT_CHECK_ARG(M > 0 && C >= 8 && C % 8 == 0 && C <= 1024, DCPT_E_SHAPE, "ln_act: need C %% 8 == 0 and 8 <= C <= 1024 (C=%d)", C);
  const int lpr = pick_lpr(C), nv = ceil_div(C / 4, lpr);
  const int grid = (int)ceil_div_ll(M, kWarps * (32 / lpr));
  DCPT_PROF("ln_act_fwd", 8.0 * M * C, (resid ? 6.0 : 4.0) * M * C, st);
#define LAUNCH(NVV) ln_act_fwd_kernel<NVV><<<grid, kWarps * 32, 0, st>>>(x, w, b, resid, y, stats, M, C, lpr, relu, eps)
  if (nv <= 1) LAUNCH(1); else if (nv <= 2) LAUNCH(2); else if (nv <= 4) LAUNCH(4); else LAUNCH(8);
#undef LAUNCH
  DCPT_LAUNCH_CHECK();
  return 0;
}

int ln_act_bwd_launch(const float* dy, const bf16* y, const bf16* x, const float* stats, const float* w, bf16* dx, float* dres,
                      float* dw, float* db, int M, int C, int relu, cudaStream_t st) {
  DCPT_CHECK_ARG(M > 0 && C >= 8 && C % 8 == 0 && C <= 1024, DCPT_E_SHAPE, "ln_act bwd: need C %% 8 == 0 and 8 <= C <= 1024 (C=%d)", C);
  const int lpr = pick_lpr(C), nv = ceil_div(C / 4, lpr);
  long long grid = ceil_div_ll(M, kWarps * (32 / lpr));
  const long long cap = (long long)dcpt_num_sms() * 3;
  if (grid > cap) grid = cap;
  const size_t smem = (size_t)2 * C * sizeof(float);
  DCPT_PROF("ln_act_bwd", 16.0 * M * C, (dres ? 10.0 : 8.0) * M * C, st);
#define LAUNCH(NVV) ln_act_bwd_kernel<NVV><<<(int)grid, kWarps * 32, smem, st>>>(dy, y, x, stats, w, dx, dres, dw, db, M, C, lpr, relu)
  if (nv <= 1) LAUNCH(1); else if (nv <= 2) LAUNCH(2); else if (nv <= 4) LAUNCH(4); else LAUNCH(8);
#undef LAUNCH
  DCPT_LAUNCH_CHECK();
  return 0;
}

int mix_fwd_launch(const bf16* prev, const float* feat, const float* mw, bf16* z, long long n, cudaStream_t st) {
  DCPT_CHECK_ARG(n % 8 == 0, DCPT_E_SHAPE, "mix: element count must be a multiple of 8");
  DCPT_PROF("mix_fwd", 2.0 * n, (prev ? 8.0 : 6.0) * n, st);
  mix_fwd_kernel<<<(unsigned)ceil_div_ll(n / 8, 256), 256, 0, st>>>(prev, feat, mw, z, n / 8);
  DCPT_LAUNCH_CHECK();
  return 0;
}
int mix_bwd_launch(const float* dz, const float* feat, const float* mw, float* dfeat, float* dmw, long long n, cudaStream_t st) {
  DCPT_CHECK_ARG(n % 8 == 0, DCPT_E_SHAPE, "mix: element count must be a multiple of 8");
  long long blocks = ceil_div_ll(n / 8, 256 * 8);
  if (blocks > 4096) blocks = 4096;
  if (blocks < 1) blocks = 1;
  DCPT_PROF("mix_bwd", 3.0 * n, (dfeat ? 10.0 : 6.0) * n, st);
  mix_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(dz, feat, mw, dfeat, dmw, n / 8);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int maxpool2_relu_fwd_launch(const bf16* x, bf16* y, int N, int Ho, int Wo, int C, cudaStream_t st) {
  const long long nvec = (long long)N * Ho * Wo * (C / 8);
  DCPT_PROF("maxpool2_relu_fwd", 4.0 * nvec * 8, 10.0 * nvec * 8, st);
  maxpool2_relu_fwd_kernel<<<(unsigned)ceil_div_ll(nvec, 256), 256, 0, st>>>(x, y, nvec, Ho, Wo, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}
int maxpool2_relu_bwd_launch(const bf16* x, const float* dy, bf16* dx, int N, int Ho, int Wo, int C, cudaStream_t st) {
  const long long nvec = (long long)N * Ho * Wo * (C / 8);
  DCPT_PROF("maxpool2_relu_bwd", 8.0 * nvec * 8, 18.0 * nvec * 8, st);
  maxpool2_relu_bwd_kernel<<<(unsigned)ceil_div_ll(nvec, 256), 256, 0, st>>>(x, dy, dx, nvec, Ho, Wo, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int meanpool_fc_fwd_launch(const bf16* x, const float* w, const float* b, float* pooled, float* logits, int N, int HW, int C, int K,
                           cudaStream_t st) {
  DCPT_PROF("meanpool_fc_fwd", 2.0 * N * HW * C, 2.0 * N * HW * C, st);
  dim3 grid(ceil_div(C, 64), N);
  meanpool_kernel<<<grid, 256, 0, st>>>(x, pooled, HW, C);
  DCPT_LAUNCH_CHECK();
  fc_fwd_kernel<<<ceil_div(N * K * 32, 256), 256, 0, st>>>(pooled, w, b, logits, N, C, K);
  DCPT_LAUNCH_CHECK();
  return 0;
}
int meanpool_fc_bwd_launch(const float* dlogits, const float* pooled, const float* w, float* dw, float* db, float* dx, int N, int HW,
                           int C, int K, cudaStream_t st) {
  DCPT_PROF("meanpool_fc_bwd", 2.0 * N * HW * C, 2.0 * N * HW * C, st);
  fc_bwd_w_kernel<<<ceil_div(K * C + K, 256), 256, 0, st>>>(dlogits, pooled, dw, db, N, C, K);
  DCPT_LAUNCH_CHECK();
  dim3 grid(ceil_div(C, 128), N, HW < 32 ? HW : 32);
  meanpool_fc_bwd_x_kernel<<<grid, 128, 0, st>>>(dlogits, w, dx, HW, C, K);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int add_bf16_launch(const bf16* a, const bf16* b, bf16* out, long long n, cudaStream_t st) {
  DCPT_CHECK_ARG(n % 8 == 0, DCPT_E_SHAPE, "add: element count must be a multiple of 8");
  DCPT_PROF("add_bf16", 1.0 * n, 6.0 * n, st);
  add_bf16_kernel<<<(unsigned)ceil_div_ll(n / 8, 256), 256, 0, st>>>(a, b, out, n / 8);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int pack_conv3x3_launch(const float* w, bf16* out, int Cout, int Cin, int dgrad, cudaStream_t st) {
  const int R = dgrad ? Cin : Cout, Cc = dgrad ? Cout : Cin;
  const long long total = (long long)R * 9 * (ceil_div(Cc, 64) * 64);
  DCPT_PROF("pack_conv3x3", 0.0, 6.0 * total, st);
  pack_conv3x3_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(w, out, Cout, Cin, dgrad);
  DCPT_LAUNCH_CHECK();
  return 0;
}
int finish_conv3x3_launch(const float* G, float* dw, int Cout, int Cin, cudaStream_t st) {
  const long long total = (long long)Cout * Cin * 9;
  DCPT_PROF("finish_conv3x3", 1.0 * total, 12.0 * total, st);
  finish_conv3x3_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(G, dw, Cout, Cin);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int im2col7s2_launch(const float* img, bf16* P, int N, int H, int W, cudaStream_t st) {
  DCPT_CHECK_ARG(N > 0 && H > 0 && W > 0, DCPT_E_SHAPE, "im2col7s2: bad shape");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)N * Ho * Wo * 20;
  DCPT_PROF("im2col7s2", 0.0, 4.0 * 3 * N * H * W + 320.0 * N * Ho * Wo, st);
  long long blocks = ceil_div_ll(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  im2col7s2_kernel<<<(unsigned)blocks, 256, 0, st>>>(img, P, N, H, W, Ho, Wo);
  DCPT_LAUNCH_CHECK();
  return 0;
}
