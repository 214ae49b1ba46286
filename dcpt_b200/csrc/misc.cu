// Small CUDA-core kernels of the NAFNet hot path: simplified channel attention (SCA),
// per-image channel scaling, column sums, casts, pixel-unshuffle, weight packing and the
// wgrad finishing steps.  References: nafnet_arch.py:116-127 (sca), :173 (x * sca(x)),
// :230 (downs, 2x2 stride 2), :238-242 (ups, 1x1 + PixelShuffle(2)).
#include "elementwise.cuh"

namespace {

// ------------------------------- SCA -------------------------------------------
// s[n][co] = b[co] + (1/HW) * sum_ci W[co][ci] * pool[n][ci]           (one warp per output)
__global__ void sca_fwd_kernel(const float* __restrict__ pool, const float* __restrict__ w, const float* __restrict__ b,
                               float* __restrict__ s, int N, int C, float inv_hw) {
  pdl_sync();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= N * C) return;
  const int n = gw / C, co = gw - n * C;
  // 4 independent chains per lane, combined in a fixed order
  float a4[4] = {0.f, 0.f, 0.f, 0.f};
  int ci = lane;
  for (; ci + 96 < C; ci += 128) {
#pragma unroll
    for (int k = 0; k < 4; ++k) a4[k] = fmaf(__ldg(w + (size_t)co * C + ci + 32 * k), __ldg(pool + (size_t)n * C + ci + 32 * k), a4[k]);
  }
  for (; ci < C; ci += 32) a4[0] = fmaf(__ldg(w + (size_t)co * C + ci), __ldg(pool + (size_t)n * C + ci), a4[0]);
  float acc = warp_sum((a4[0] + a4[1]) + (a4[2] + a4[3]));
  if (lane == 0) s[gw] = b[co] + acc * inv_hw;
}

// SCA in ONE launch (nafnet_arch.py:171-173: x * sca(x)): a CTA owns (image n, 64-channel group, pixel slice).  It first computes
// its 64 values of s - every CTA of a group recomputes them (64 x C MACs, the weight rows come from L2), so no grid-wide
// dependency is needed between "s is known" and "rows are scaled" - then scales its pixels.  The per-row summation order is
// sca_fwd_kernel's (four chains per lane, fixed combine), so s is bit-identical to the two-kernel path; eight rows are in
// flight per warp.  Slice 0 writes s (saved for the backward).
__global__ void __launch_bounds__(256)
sca_scale_kernel(const float* __restrict__ pool, const float* __restrict__ w, const float* __restrict__ b, const bf16* __restrict__ g,
                 float* __restrict__ s, bf16* __restrict__ gs, int C, int HW, float inv_hw) {
  __shared__ float s_s[64];
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 64, n = blockIdx.y;
  {
    float a4[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r) a4[r][0] = a4[r][1] = a4[r][2] = a4[r][3] = 0.f;
    const float* prow = pool + (size_t)n * C;
    int ci = lane;
    for (; ci + 96 < C; ci += 128) {
      float pv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) pv[k] = __ldg(prow + ci + 32 * k);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int co = min(c0 + warp * 8 + r, C - 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) a4[r][k] = fmaf(__ldg(w + (size_t)co * C + ci + 32 * k), pv[k], a4[r][k]);
      }
    }
    for (; ci < C; ci += 32) {
      const float pv = __ldg(prow + ci);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int co = min(c0 + warp * 8 + r, C - 1);
        a4[r][0] = fmaf(__ldg(w + (size_t)co * C + ci), pv, a4[r][0]);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float acc = warp_sum((a4[r][0] + a4[r][1]) + (a4[r][2] + a4[r][3]));
      const int co = c0 + warp * 8 + r;
      if (lane == 0) {
        const float v = co < C ? b[co] + acc * inv_hw : 0.f;
        s_s[warp * 8 + r] = v;
        if (co < C && blockIdx.z == 0) s[(size_t)n * C + co] = v;
      }
    }
  }
  __syncthreads();
  // 8 threads x 16 bytes cover the group's 64 channels of one pixel; 32 pixels per pass, 4 passes in flight
  // (requesting the first rows before the mat-vec, or 4 CTAs per SM, measured no better / worse: r03b, r03c)
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c8 = c0 + sub * 8;
  if (c8 >= C) return;
  float sv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) sv[k] = s_s[sub * 8 + k];
  const int per = (HW + gridDim.z - 1) / gridDim.z;
  const int p0 = blockIdx.z * per, p1 = min(HW, p0 + per);
  const size_t img = (size_t)n * HW;
  for (int px = p0 + pl; px < p1; px += 128) {
    uint4 raw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (px + 32 * k < p1) raw[k] = ldg16(g + (img + px + 32 * k) * C + c8);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (px + 32 * k < p1) {
        float v[8];
        unpack8(raw[k], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= sv[j];
        stg16(gs + (img + px + 32 * k) * C + c8, pack8(v));
      }
  }
}

// four 16-byte vectors per thread (a block covers 1024 consecutive vectors): all loads are issued before the first use
__global__ void __launch_bounds__(256) scale_rows_kernel(const bf16* __restrict__ g, const float* __restrict__ s, bf16* __restrict__ gs,
                                                         long long nvec, int HW, int C) {
  pdl_sync();
  const int CV = C >> 3;
  const long long base = (long long)blockIdx.x * 1024 + threadIdx.x;
  uint4 raw[4];
  float4 s0[4], s1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256;
    if (i < nvec) {
      const long long px = i / CV;
      const int c = (int)(i - px * CV) * 8;
      const int n = (int)(px / HW);
      raw[k] = ldg16(g + i * 8);
      s0[k] = __ldg(reinterpret_cast<const float4*>(s + (size_t)n * C + c));
      s1[k] = __ldg(reinterpret_cast<const float4*>(s + (size_t)n * C + c + 4));
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256;
    if (i < nvec) {
      float v[8];
      unpack8(raw[k], v);
      v[0] *= s0[k].x; v[1] *= s0[k].y; v[2] *= s0[k].z; v[3] *= s0[k].w;
      v[4] *= s1[k].x; v[5] *= s1[k].y; v[6] *= s1[k].z; v[7] *= s1[k].w;
      stg16(gs + i * 8, pack8(v));
    }
  }
}

// ds[n][c] += sum_{px in image n} dgs[px][c] * g[px][c]
__global__ void __launch_bounds__(256)
sca_ds_reduce_kernel(const bf16* __restrict__ dgs, const bf16* __restrict__ g, float* __restrict__ ds, int HW, int C, int cvb) {
  extern __shared__ float s_red[];  // [cvb*8]
  for (int i = threadIdx.x; i < cvb * 8; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.z, CV = C >> 3;
  const int cvl = threadIdx.x % cvb, pl = threadIdx.x / cvb, npl = blockDim.x / cvb;
  const int cv = blockIdx.y * cvb + cvl;
  if (cv < CV) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int px = blockIdx.x * npl + pl; px < HW; px += gridDim.x * npl) {
      const size_t off = ((size_t)n * HW + px) * C + cv * 8;
      float a[8], b[8];
      unpack8(ldg16(dgs + off), a);
      unpack8(ldg16(g + off), b);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(a[i], b[i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&s_red[cvl * 8 + i], acc[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cvb * 8; i += blockDim.x) {
    const int c = blockIdx.y * cvb * 8 + i;
    if (c < C) atomicAdd(ds + (size_t)n * C + c, s_red[i]);
  }
}

// dW[co][ci] += (1/HW) sum_n ds[n][co] pool[n][ci];  db[co] += sum_n ds[n][co]
__global__ void sca_bwd_w_kernel(const float* __restrict__ ds, const float* __restrict__ pool, float* __restrict__ dw,
                                 float* __restrict__ db, int N, int C, float inv_hw) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long CC = (long long)C * C;
  if (i < CC) {
    const int co = (int)(i / C), ci = (int)(i - (long long)co * C);
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc = fmaf(__ldg(ds + (size_t)n * C + co), __ldg(pool + (size_t)n * C + ci), acc);
    dw[i] += acc * inv_hw;
  } else if (i < CC + C) {
    const int co = (int)(i - CC);
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc += ds[(size_t)n * C + co];
    db[co] += acc;
  }
}

// t[n][ci] = (1/HW) sum_co W[co][ci] ds[n][co]: block = 32 ci x 8 warps splitting co; grid (C/32, N)
__global__ void __launch_bounds__(256)
sca_bwd_t_kernel(const float* __restrict__ ds, const float* __restrict__ w, float* __restrict__ t, int C, float inv_hw) {
  __shared__ float s_part[8][32];
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci = blockIdx.x * 32 + lane, n = blockIdx.y;
  float acc = 0.f;
  if (ci < C) {
    // 4 independent chains (the loads of 4 rows of W in flight at once); the final order of additions is fixed
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    int co = warp;
    for (; co + 24 < C; co += 32) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        a4[k] = fmaf(__ldg(w + (size_t)(co + 8 * k) * C + ci), __ldg(ds + (size_t)n * C + co + 8 * k), a4[k]);
    }
    for (; co < C; co += 8) a4[0] = fmaf(__ldg(w + (size_t)co * C + ci), __ldg(ds + (size_t)n * C + co), a4[0]);
    acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  }
  s_part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && ci < C) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += s_part[k][lane];
    t[(size_t)n * C + ci] = v * inv_hw;
  }
}

// ------------------------------ TLC local pooling --------------------------------
// Test-time local converter (reference: AvgPool2d.forward, arch_util.py:339-398, non-fast path): the SCA pooling of the
// `NAFNet` variant is a k1 x k2 box mean taken from a 2-D cumulative sum, replicate-padded back to H x W.  Same algorithm
// here: fp32 inclusive prefix sums along W, then along H (in place), then the four-corner difference per pixel.
// g bf16 [N,H,W,C] -> I fp32 [N,H,W,C]; thread = (n, y, 8-channel vector), sequential over x
__global__ void tlc_rowsum_kernel(const bf16* __restrict__ g, float* __restrict__ I, long long total, int W, int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int CV = C >> 3;
  const int cv = (int)(idx % CV);
  const long long ny = idx / CV;
  const size_t base = (size_t)ny * W * C + cv * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int x = 0; x < W; ++x) {
    float v[8];
    unpack8(ldg16(g + base + (size_t)x * C), v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += v[i];
    float4* o = reinterpret_cast<float4*>(I + base + (size_t)x * C);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}
// in-place prefix sums along H; thread = (n, x, 4-channel vector)
__global__ void tlc_colsum_kernel(float* __restrict__ I, long long total, int H, int W, int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int C4 = C >> 2;
  const int c4 = (int)(idx % C4);
  const long long nx = idx / C4;
  const int x = (int)(nx % W);
  const long long n = nx / W;
  float4* p = reinterpret_cast<float4*>(I + ((size_t)n * H * W + x) * C) + c4;
  const size_t stride = (size_t)W * C / 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int y = 0; y < H; ++y) {
    const float4 v = p[(size_t)y * stride];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    p[(size_t)y * stride] = acc;
  }
}
// P[n,y,x,:] = box mean with window start (clamp(y - pt, 0, H - k1), clamp(x - pl, 0, W - k2))   (replicate padding, :391-396)
__global__ void tlc_box_kernel(const float* __restrict__ I, bf16* __restrict__ P, long long total, int H, int W, int C, int k1, int k2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int CV = C >> 3;
  const int cv = (int)(idx % CV);
  const long long px = idx / CV;
  const int x = (int)(px % W);
  const long long t = px / W;
  const int y = (int)(t % H);
  const long long n = t / H;
  const int oh = H - k1 + 1, ow = W - k2 + 1;          // size of the un-padded box-mean map
  const int pt = (H - oh) / 2, pl = (W - ow) / 2;      // replicate padding before (after = the rest)
  const int ys = min(max(y - pt, 0), oh - 1), xs = min(max(x - pl, 0), ow - 1);
  const int ye = ys + k1 - 1, xe = xs + k2 - 1;
  const float* In = I + (size_t)n * H * W * C + cv * 8;
  auto at = [&](int yy, int xx, float (&o)[8]) {
    if (yy < 0 || xx < 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
      return;
    }
    const float4* q = reinterpret_cast<const float4*>(In + ((size_t)yy * W + xx) * C);
    const float4 a = __ldg(q), b = __ldg(q + 1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  };
  float s4[8], s1[8], s2[8], s3[8], r[8];
  at(ye, xe, s4); at(ys - 1, xs - 1, s1); at(ys - 1, xe, s2); at(ye, xs - 1, s3);
  const float inv = 1.f / (float)(k1 * k2);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = (s4[i] + s1[i] - s2[i] - s3[i]) * inv;
  stg16(P + (size_t)px * C + cv * 8, pack8(r));
}
__global__ void mul_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, long long nvec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  float x[8], y[8];
  unpack8(ldg16(a + i * 8), x);
  unpack8(ldg16(b + i * 8), y);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] *= y[k];
  stg16(out + i * 8, pack8(x));
}

// ------------------------------ column sums ------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ x, float* __restrict__ out, int M, int C, int cvb) {
  extern __shared__ float s_red[];
  for (int i = threadIdx.x; i < cvb * 8; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  const int CV = C >> 3;
  const int cvl = threadIdx.x % cvb, pl = threadIdx.x / cvb, npl = blockDim.x / cvb;
  const int cv = blockIdx.y * cvb + cvl;
  if (cv < CV) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long m = (long long)blockIdx.x * npl + pl; m < M; m += (long long)gridDim.x * npl) {
      float a[8];
      unpack8(ldg16(x + (size_t)m * C + cv * 8), a);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += a[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&s_red[cvl * 8 + i], acc[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cvb * 8; i += blockDim.x) {
    const int c = blockIdx.y * cvb * 8 + i;
    if (c < C) atomicAdd(out + c, s_red[i]);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long nvec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i);
  const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  stg16(y + i * 8, pack8(v));
}

// x: fp32 [N, 2H, 2W, C]  ->  y: bf16 [N*H*W, 4C], y[m][(i*2+j)*C + c] = x[n][2h+i][2w+j][c]
__global__ void unshuffle_cast_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long nvec, int H, int W, int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nvec) return;
  const int CV = C >> 3;
  const int cv = (int)(idx % CV);
  const int q = (int)((idx / CV) & 3);
  const long long m = idx / (4 * CV);
  const int w = (int)(m % W);
  const long long tmp = m / W;
  const int h = (int)(tmp % H);
  const long long n = tmp / H;
  const size_t src = (((size_t)n * 2 * H + 2 * h + (q >> 1)) * (size_t)(2 * W) + 2 * w + (q & 1)) * C + cv * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(x + src));
  const float4 b = __ldg(reinterpret_cast<const float4*>(x + src + 4));
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  stg16(y + idx * 8, pack8(v));
}

// ------------------------------ weight packing ----------------------------------
// out is row-major bf16 [R, Cc]; (r, c) -> source element of the fp32 parameter w (and its out-channel o for row_scale).
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ row_scale, bf16* __restrict__ out,
                                   int O, int I, int mode) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)O * I;
  if (idx >= total) return;
  int o, i;  // source coordinates: w viewed as [O][I] (I = all input-side dims flattened as stored)
  switch (mode) {
    case PACK_PLAIN: {  // out[o][i]
      o = (int)(idx / I); i = (int)(idx % I);
    } break;
    case PACK_T: {  // out[i][o]
      i = (int)(idx / O); o = (int)(idx % O);
    } break;
    case PACK_PAIR: {  // out[p][i], p = 16*pp + h*8 + e  <->  o = h*C + 8*pp + e  (C = O/2)
      const int p = (int)(idx / I); i = (int)(idx % I);
      o = ((p & 15) >> 3) * (O >> 1) + (p >> 4) * 8 + (p & 7);
    } break;
    case PACK_PAIR32: {  // out[p][i], p = 64*pp + h*32 + e  <->  o = h*C + 32*pp + e  (C = O/2, C % 32 == 0)
      const int p = (int)(idx / I); i = (int)(idx % I);
      o = ((p & 63) >> 5) * (O >> 1) + (p >> 6) * 32 + (p & 31);
    } break;
    case PACK_UP: {  // out[p][i], p = q*Cseg + c'  <->  o = c'*4 + q  (Cseg = O/4)
      const int p = (int)(idx / I); i = (int)(idx % I);
      const int cseg = O >> 2;
      o = (p % cseg) * 4 + p / cseg;
    } break;
    case PACK_UP_T: {  // out[i][p]
      i = (int)(idx / O);
      const int p = (int)(idx % O);
      const int cseg = O >> 2;
      o = (p % cseg) * 4 + p / cseg;
    } break;
    case PACK_DOWN: {  // w [O][C][2][2] (I = 4C): out[o][k], k = (dy*2+dx)*C + c  <->  i = c*4 + dy*2 + dx
      o = (int)(idx / I);
      const int k = (int)(idx % I);
      const int c4 = I >> 2;
      i = (k % c4) * 4 + k / c4;
    } break;
    default: {  // PACK_DOWN_T: out[k][o]
      const int k = (int)(idx / O);
      o = (int)(idx % O);
      const int c4 = I >> 2;
      i = (k % c4) * 4 + k / c4;
    } break;
  }
  float v = w[(size_t)o * I + i];
  if (row_scale) v *= row_scale[o];
  out[idx] = OP_FROM_F32(v);
}

__global__ void pack_bias_kernel(const float* __restrict__ bias, const float* __restrict__ scale, float* __restrict__ out, int O,
                                 int mode) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= O) return;
  int o = p;
  if (mode == PACK_PAIR) o = ((p & 15) >> 3) * (O >> 1) + (p >> 4) * 8 + (p & 7);
  if (mode == PACK_PAIR32) o = ((p & 63) >> 5) * (O >> 1) + (p >> 6) * 32 + (p & 31);
  float v = bias[o];
  if (scale) v *= scale[o];
  out[p] = v;
}

// ------------------------------ wgrad finishing ----------------------------------
// Residual-scaled conv (conv3 with beta, conv5 with gamma): forward used W' = diag(scale) W, b' = scale*b.
// G = dYres^T X is the wgrad w.r.t. W' rows *before* the scale.  One warp per out-channel o:
//   dW[o][i] += scale[o] G[o][i];  dscale[o] += sum_i W[o][i] G[o][i] + b[o] S[o];  db[o] += scale[o] S[o]
__global__ void wgrad_finish_resid_kernel(const float* __restrict__ G, const float* __restrict__ w, const float* __restrict__ bias,
                                          const float* __restrict__ scale, const float* __restrict__ colsum, float* __restrict__ dw,
                                          float* __restrict__ dbias, float* __restrict__ dscale, int O, int I) {
  pdl_sync();
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (o >= O) return;
  const float sc = scale[o];
  float dot = 0.f;
  for (int i = lane; i < I; i += 32) {
    const float gv = G[(size_t)o * I + i];
    dot = fmaf(w[(size_t)o * I + i], gv, dot);
    dw[(size_t)o * I + i] += sc * gv;
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    const float S = colsum[o];
    dscale[o] += dot + bias[o] * S;
    dbias[o] += sc * S;
  }
}

__global__ void wgrad_finish_perm_kernel(const float* __restrict__ G, float* __restrict__ dw, int O, int I, int mode) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)O * I) return;
  const int p = (int)(idx / I), k = (int)(idx % I);
  int o = p, i = k;
  if (mode == FIN_UP) {  // G rows in packed order p = q*Cseg + c'
    const int cseg = O >> 2;
    o = (p % cseg) * 4 + p / cseg;
  } else if (mode == FIN_DOWN) {  // G cols k = (dy*2+dx)*C + c  ->  i = c*4 + dy*2+dx
    const int c4 = I >> 2;
    i = (k % c4) * 4 + k / c4;
  } else if (mode == FIN_CN_TO_C3) {  // G [C=O][27 = I], k = oc*9 + t  ->  dw[oc][c][t] (ending conv weight [3][C][3][3])
    const int oc = k / 9, t = k % 9;
    dw[((size_t)oc * O + p) * 9 + t] += G[idx];
    return;
  }
  dw[(size_t)o * I + i] += G[idx];
}

// out = a (+ b): fp32 copy, bf16 mirror and column sums of a residual-stream gradient.
__global__ void __launch_bounds__(256)
grad_prepare_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, bf16* __restrict__ out_bf16,
                    float* __restrict__ colsum, int M, int C, int cvb) {
  extern __shared__ float s_red[];
  for (int i = threadIdx.x; i < cvb * 8; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  const int CV = C >> 3;
  const int cvl = threadIdx.x % cvb, pl = threadIdx.x / cvb, npl = blockDim.x / cvb;
  const int cv = blockIdx.y * cvb + cvl;
  if (cv < CV) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long m = (long long)blockIdx.x * npl + pl; m < M; m += (long long)gridDim.x * npl) {
      const size_t off = (size_t)m * C + cv * 8;
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (a) {
        x0 = __ldg(reinterpret_cast<const float4*>(a + off));
        x1 = __ldg(reinterpret_cast<const float4*>(a + off + 4));
      }
      if (b) {
        const float4 y0 = __ldg(reinterpret_cast<const float4*>(b + off));
        const float4 y1 = __ldg(reinterpret_cast<const float4*>(b + off + 4));
        x0.x += y0.x; x0.y += y0.y; x0.z += y0.z; x0.w += y0.w;
        x1.x += y1.x; x1.y += y1.y; x1.z += y1.z; x1.w += y1.w;
      }
      const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      if (out) {
        *reinterpret_cast<float4*>(out + off) = x0;
        *reinterpret_cast<float4*>(out + off + 4) = x1;
      }
      if (out_bf16) stg16(out_bf16 + off, pack8(v));
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&s_red[cvl * 8 + i], acc[i]);
  }
  __syncthreads();
  if (colsum)
    for (int i = threadIdx.x; i < cvb * 8; i += blockDim.x) {
      const int c = blockIdx.y * cvb * 8 + i;
      if (c < C) atomicAdd(colsum + c, s_red[i]);
    }
}

__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

}  // namespace

int sca_fwd_launch(const float* pool, const float* w, const float* b, float* s, int N, int C, int HW, cudaStream_t st) {
  const long long threads = (long long)N * C * 32;
  DCPT_PROF("sca_fwd", 2.0 * N * C * C, 4.0 * C * C, st);
  DCPT_CUDA(dcpt_launch_pdl(sca_fwd_kernel, dim3((unsigned)ceil_div_ll(threads, 256)), dim3(256), 0, st, pool, w, b, s, N, C, 1.f / (float)HW));
  DCPT_LAUNCH_CHECK();
  return 0;
}

int sca_scale_launch(const float* pool, const float* w, const float* b, const bf16* g, float* s, bf16* gs, int N, int C, int HW,
                     cudaStream_t st) {
  DCPT_CHECK_ARG(C % 8 == 0 && C >= 8, DCPT_E_SHAPE, "sca_scale: C=%d must be a multiple of 8", C);
  const int groups = ceil_div(C, 64);
  int slices = (2 * dcpt_num_sms()) / (groups * N);  // ~2 CTAs per SM; every slice recomputes its group's 64 x C mat-vec
  if (slices < 1) slices = 1;
  if (slices > HW / 128) slices = HW / 128 > 0 ? HW / 128 : 1;
  DCPT_PROF("sca_scale", 2.0 * N * C * C * slices + (double)N * HW * C, 4.0 * N * HW * C + 4.0 * C * C, st);
  DCPT_CUDA(dcpt_launch_pdl(sca_scale_kernel, dim3(groups, N, slices), dim3(256), 0, st, pool, w, b, g, s, gs, C, HW, 1.f / (float)HW));
  DCPT_LAUNCH_CHECK();
  return 0;
}

int scale_rows_launch(const bf16* g, const float* s, bf16* gs, int N, int HW, int C, cudaStream_t st) {
  const long long nvec = (long long)N * HW * (C / 8);
  DCPT_PROF("scale_rows", (double)N * HW * C, 4.0 * N * HW * C, st);
  DCPT_CUDA(dcpt_launch_pdl(scale_rows_kernel, dim3((unsigned)ceil_div_ll(nvec, 1024)), dim3(256), 0, st, g, s, gs, nvec, HW, C));
  DCPT_LAUNCH_CHECK();
  return 0;
}

static inline int pick_cvb(int CV) {
  int cvb = 1;
  while (cvb < CV && cvb < 32) cvb <<= 1;
  return cvb;
}

int sca_ds_reduce_launch(const bf16* dgs, const bf16* g, float* ds, int N, int HW, int C, cudaStream_t st) {
  const int CV = C / 8, cvb = pick_cvb(CV), npl = 256 / cvb;
  int gx = ceil_div(HW, npl * 16);
  if (gx < 1) gx = 1;
  dim3 grid(gx, ceil_div(CV, cvb), N);
  DCPT_PROF("sca_ds_reduce", 2.0 * N * HW * C, 4.0 * N * HW * C, st);
  sca_ds_reduce_kernel<<<grid, 256, cvb * 8 * sizeof(float), st>>>(dgs, g, ds, HW, C, cvb);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int sca_bwd_launch(const float* ds, const float* pool, const float* w, float* t, float* dw, float* db, int N, int C, int HW,
                   cudaStream_t st, cudaStream_t st_w) {
  // t (needed by the gate backward) on `st`; the parameter gradients dW, db on `st_w` (the caller's weight-gradient stream,
  // already ordered after the producer of ds)
  if (st_w == nullptr) st_w = st;
  const long long total = (long long)C * C + C;
  {
    DCPT_PROF("sca_bwd_w", 2.0 * N * C * C, 8.0 * C * C, st_w);
    DCPT_CUDA(dcpt_launch_pdl(sca_bwd_w_kernel, dim3((unsigned)ceil_div_ll(total, 256)), dim3(256), 0, st_w, ds, pool, dw, db, N, C, 1.f / (float)HW));
  }
  if (t == nullptr) return 0;  // the gate backward computes its own shift (dwgate_bwd_a_launch with ds + the SCA weight)
  dim3 grid(ceil_div(C, 32), N);
  DCPT_PROF("sca_bwd_t", 2.0 * N * C * C, 4.0 * C * C, st);
  DCPT_CUDA(dcpt_launch_pdl(sca_bwd_t_kernel, grid, dim3(256), 0, st, ds, w, t, C, 1.f / (float)HW));
  DCPT_LAUNCH_CHECK();
  return 0;
}

int tlc_boxmean_launch(const bf16* g, float* I, bf16* P, int N, int H, int W, int C, int k1, int k2, cudaStream_t st) {
  DCPT_CHECK_ARG(C % 8 == 0 && k1 >= 1 && k2 >= 1 && k1 <= H && k2 <= W, DCPT_E_SHAPE, "tlc box mean: bad kernel %dx%d for %dx%d (C=%d)",
                 k1, k2, H, W, C);
  const long long t1 = (long long)N * H * (C / 8), t2 = (long long)N * W * (C / 4), t3 = (long long)N * H * W * (C / 8);
  DCPT_PROF("tlc_boxmean", 3.0 * N * H * W * C, 16.0 * N * H * W * C, st);
  tlc_rowsum_kernel<<<(unsigned)ceil_div_ll(t1, 128), 128, 0, st>>>(g, I, t1, W, C);
  tlc_colsum_kernel<<<(unsigned)ceil_div_ll(t2, 128), 128, 0, st>>>(I, t2, H, W, C);
  tlc_box_kernel<<<(unsigned)ceil_div_ll(t3, 256), 256, 0, st>>>(I, P, t3, H, W, C, k1, k2);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int mul_bf16_launch(const bf16* a, const bf16* b, bf16* out, long long n, cudaStream_t st) {
  DCPT_CHECK_ARG(n % 8 == 0, DCPT_E_SHAPE, "mul: element count %lld must be a multiple of 8", n);
  DCPT_PROF("mul_bf16", (double)n, 6.0 * n, st);
  mul_bf16_kernel<<<(unsigned)ceil_div_ll(n / 8, 256), 256, 0, st>>>(a, b, out, n / 8);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int colsum_bf16_launch(const bf16* x, float* out, int M, int C, cudaStream_t st) {
  const int CV = C / 8, cvb = pick_cvb(CV), npl = 256 / cvb;
  long long gx = ceil_div_ll(M, (long long)npl * 32);
  if (gx < 1) gx = 1;
  if (gx > 4096) gx = 4096;
  dim3 grid((unsigned)gx, ceil_div(CV, cvb));
  DCPT_PROF("colsum_bf16", (double)M * C, 2.0 * M * C, st);
  colsum_bf16_kernel<<<grid, 256, cvb * 8 * sizeof(float), st>>>(x, out, M, C, cvb);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int cast_f32_bf16_launch(const float* x, bf16* y, long long n, cudaStream_t st) {
  DCPT_CHECK_ARG(n % 8 == 0, DCPT_E_SHAPE, "cast: element count %lld must be a multiple of 8", n);
  DCPT_PROF("cast_f32_bf16", 0.0, 6.0 * n, st);
  cast_f32_bf16_kernel<<<(unsigned)ceil_div_ll(n / 8, 256), 256, 0, st>>>(x, y, n / 8);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int unshuffle_cast_launch(const float* x, bf16* y, int N, int H, int W, int C, cudaStream_t st) {
  const long long nvec = (long long)N * H * W * 4 * (C / 8);
  DCPT_PROF("unshuffle_cast", 0.0, 6.0 * nvec * 8, st);
  unshuffle_cast_kernel<<<(unsigned)ceil_div_ll(nvec, 256), 256, 0, st>>>(x, y, nvec, H, W, C);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int pack_weight_launch(const float* w, const float* row_scale, bf16* out, int O, int I, int mode, cudaStream_t st) {
  const long long total = (long long)O * I;
  DCPT_PROF("pack_weight", 0.0, 6.0 * total, st);
  pack_weight_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(w, row_scale, out, O, I, mode);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int pack_bias_launch(const float* bias, const float* scale, float* out, int O, int mode, cudaStream_t st) {
  DCPT_PROF("pack_bias", 0.0, 8.0 * O, st);
  pack_bias_kernel<<<ceil_div(O, 256), 256, 0, st>>>(bias, scale, out, O, mode);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int wgrad_finish_resid_launch(const float* G, const float* w, const float* bias, const float* scale, const float* colsum,
                              float* dw, float* dbias, float* dscale, int O, int I, cudaStream_t st) {
  DCPT_PROF("wgrad_finish_resid", 4.0 * O * I, 16.0 * O * I, st);
  DCPT_CUDA(dcpt_launch_pdl(wgrad_finish_resid_kernel, dim3(ceil_div(O * 32, 256)), dim3(256), 0, st, G, w, bias, scale, colsum, dw, dbias, dscale, O, I));
  DCPT_LAUNCH_CHECK();
  return 0;
}

int wgrad_finish_perm_launch(const float* G, float* dw, int O, int I, int mode, cudaStream_t st) {
  const long long total = (long long)O * I;
  DCPT_PROF("wgrad_finish_perm", (double)total, 12.0 * total, st);
  wgrad_finish_perm_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(G, dw, O, I, mode);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int grad_prepare_launch(const float* a, const float* b, float* out, bf16* out_bf16, float* colsum, int M, int C,
                        cudaStream_t st) {
  const int CV = C / 8, cvb = pick_cvb(CV), npl = 256 / cvb;
  long long gx = ceil_div_ll(M, (long long)npl * 16);
  if (gx < 1) gx = 1;
  if (gx > 8192) gx = 8192;
  dim3 grid((unsigned)gx, ceil_div(CV, cvb));
  DCPT_PROF("grad_prepare", (double)M * C, ((a ? 4.0 : 0.0) + (b ? 4.0 : 0.0) + (out ? 4.0 : 0.0) + (out_bf16 ? 2.0 : 0.0)) * M * C, st);
  grad_prepare_kernel<<<grid, 256, cvb * 8 * sizeof(float), st>>>(a, b, out, out_bf16, colsum, M, C, cvb);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int axpy_launch(float* dst, const float* src, int n, cudaStream_t st) {
  DCPT_PROF("axpy", (double)n, 12.0 * n, st);
  axpy_kernel<<<ceil_div(n, 256), 256, 0, st>>>(dst, src, n);
  DCPT_LAUNCH_CHECK();
  return 0;
}
