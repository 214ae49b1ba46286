// Channel LayerNorm (LayerNorm2d) forward / backward for NHWC rows.
// Reference math: basicsr/archs/nafnet_arch.py:25-53 (LayerNormFunction):
//   fwd: mu = mean_c x; var = mean_c (x-mu)^2 (biased); y = (x-mu)/sqrt(var+eps); out = w*y + b
//   bwd: g = dout*w; dx = (g - y*mean_c(g*y) - mean_c(g)) / sqrt(var+eps); dw = sum dout*y; db = sum dout
// One row (pixel) is owned by `lpr` lanes of a warp (lpr = 2..32, power of two), so the
// reductions over C are xor-shuffles; each lane keeps its NV float4 slices in registers.
// HBM-bound: fwd reads 4C and writes 2C (+8) bytes per pixel.
#include <stdlib.h>

#include "elementwise.cuh"

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float group_sum(float v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// U independent row groups per warp iteration: with narrow rows (C <= 128) a lane moves only 16-32 bytes per row, and
// one row per iteration leaves the kernel latency-bound (2.6 TB/s at C = 64); U = 4 keeps 4x the bytes in flight.
template <int NV, int U>
__global__ void __launch_bounds__(kWarps * 32, NV <= 4 ? 4 : 2)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, bf16* __restrict__ out,
              float* __restrict__ stats, int M, int C, int lpr, float eps, int center) {
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / lpr;  // rows per warp
  const int sub = lane % lpr, gr = lane / lpr;
  const int nvec = C >> 2;
  const float invC = 1.f / (float)C;
  // a warp's U row groups are CONSECUTIVE in memory (U * rpw * C * 4 contiguous bytes per warp iteration, neighbouring warps
  // adjacent): with the groups a whole grid apart every load instruction of a warp opened a different DRAM page
  const long long row_stride = rpw;
  const long long iter_stride = (long long)gridDim.x * kWarps * rpw * U;
  constexpr bool kHoist = NV <= 2;  // wide rows reload the affine parameters per row (L1 hits) to keep 4 blocks per SM
  float4 wv[NV], bv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = sub + i * lpr;
    wv[i] = (kHoist && v < nvec) ? __ldg(reinterpret_cast<const float4*>(w) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    bv[i] = (kHoist && b && v < nvec) ? __ldg(reinterpret_cast<const float4*>(b) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row0 = ((long long)blockIdx.x * kWarps + warp) * rpw * U + gr; row0 - gr < M; row0 += iter_stride) {
    float4 xv[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * row_stride;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = sub + i * lpr;
        xv[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < M && v < nvec) xv[u][i] = __ldg(reinterpret_cast<const float4*>(x + row * C) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * row_stride;
      if (row - gr >= M) break;  // warp-uniform: the whole row group is out of range
      const bool valid = row < M;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) sum += xv[u][i].x + xv[u][i].y + xv[u][i].z + xv[u][i].w;
      const float mean = group_sum(sum, lpr) * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = sub + i * lpr;
        if (v < nvec) {
          const float a = xv[u][i].x - mean, bb = xv[u][i].y - mean, c = xv[u][i].z - mean, d = xv[u][i].w - mean;
          sq += a * a + bb * bb + c * c + d * d;
        }
      }
      const float var = group_sum(sq, lpr) * invC;
      const float rstd = 1.f / sqrtf(var + eps);
      // center == 0: Restormer's BiasFree_LayerNorm (restormer_arch.py:38-40) - variance about the mean, numerator NOT centred
      const float shift = center ? mean : 0.f;
      if (valid) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = sub + i * lpr;
          if (v < nvec) {
            if constexpr (!kHoist) {
              wv[i] = __ldg(reinterpret_cast<const float4*>(w) + v);
              if (b) bv[i] = __ldg(reinterpret_cast<const float4*>(b) + v);
            }
            const float o0 = (xv[u][i].x - shift) * rstd * wv[i].x + bv[i].x;
            const float o1 = (xv[u][i].y - shift) * rstd * wv[i].y + bv[i].y;
            const float o2 = (xv[u][i].z - shift) * rstd * wv[i].z + bv[i].z;
            const float o3 = (xv[u][i].w - shift) * rstd * wv[i].w + bv[i].w;
            op16x2 p0 = OP2_FROM_F32(o0, o1), p1 = OP2_FROM_F32(o2, o3);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0);
            pk.y = *reinterpret_cast<uint32_t*>(&p1);
            *(reinterpret_cast<uint2*>(out + row * C) + v) = pk;
          }
        }
        if (stats && sub == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
      }
    }
  }
}

// NV <= 2: 16-warp blocks (few, fat blocks -> few partial-sum flushes).  NV >= 4 keeps 3 * NV float4 column
// accumulators per lane: 4-warp blocks capped at 128 registers so that 4 of them (16 warps) share an SM - with one
// 8-warp block per SM the kernel had too few bytes in flight (2.7 TB/s at C = 512).
template <int NV, int kBwdWarps, int kMinBlocks>
__global__ void __launch_bounds__(kBwdWarps * 32, kMinBlocks)
ln_bwd_kernel(const bf16* __restrict__ dn, const float* __restrict__ x, const float* __restrict__ stats,
              const float* __restrict__ w, const float* __restrict__ dres, float* __restrict__ dx, bf16* __restrict__ dx_bf16,
              float* __restrict__ dw, float* __restrict__ db, float* __restrict__ colsum, int M, int C, int lpr, int center) {
  extern __shared__ float s_acc[];  // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / lpr;
  const int sub = lane % lpr, gr = lane / lpr;
  const int nvec = C >> 2;
  const float invC = 1.f / (float)C;
  float4 a_dw[NV], a_db[NV], a_cs[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a_dw[i] = a_db[i] = a_cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long row_stride = (long long)gridDim.x * kBwdWarps * rpw;
  for (long long row = ((long long)blockIdx.x * kBwdWarps + warp) * rpw + gr; row - gr < M; row += row_stride) {
    const bool valid = row < M;
    float2 st = make_float2(0.f, 0.f);
    if (valid) st = __ldg(reinterpret_cast<const float2*>(stats + row * 2));
    const float mean = st.x, rstd = st.y;
    float4 yh[NV], g[NV], d[NV], rs[NV];
    float sg = 0.f, sgy = 0.f;
    {  // pull the warp's NEXT row towards L2 while this one is processed (the loop body is one long dependency chain)
      const long long nrow = row + row_stride;
      if (nrow < M) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = sub + i * lpr;
          if (v < nvec) {
            prefetch_l2(reinterpret_cast<const float4*>(x + nrow * C) + v);
            if (dres) prefetch_l2(reinterpret_cast<const float4*>(dres + nrow * C) + v);
            if ((v & 1) == 0) prefetch_l2(reinterpret_cast<const uint2*>(dn + nrow * C) + v);
          }
        }
      }
    }
    // every global load of the row is issued before the first dependent use: one memory round trip per row
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = sub + i * lpr;
      rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dres && valid && v < nvec) rs[i] = __ldg(reinterpret_cast<const float4*>(dres + row * C) + v);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = sub + i * lpr;
      yh[i] = g[i] = d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid && v < nvec) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + row * C) + v);
        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(dn + row * C) + v);
        const float2 d01 = OP2_TO_F32(*reinterpret_cast<const op16x2*>(&pk.x));
        const float2 d23 = OP2_TO_F32(*reinterpret_cast<const op16x2*>(&pk.y));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + v);
        d[i] = make_float4(d01.x, d01.y, d23.x, d23.y);
        yh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        g[i] = make_float4(d[i].x * wv.x, d[i].y * wv.y, d[i].z * wv.z, d[i].w * wv.w);
        sg += g[i].x + g[i].y + g[i].z + g[i].w;
        sgy += g[i].x * yh[i].x + g[i].y * yh[i].y + g[i].z * yh[i].z + g[i].w * yh[i].w;
      }
    }
    float mean_g = group_sum(sg, lpr) * invC;
    float mean_gy = group_sum(sgy, lpr) * invC;
    // center == 0: Restormer's BiasFree LN, out = x * rstd * w (restormer_arch.py:38-40).  With xh = x * rstd = yh + mean * rstd:
    //   dx = rstd * (g - yh * mean_c(g * xh)),  dw += dn * xh      (the mean is not an input of the numerator)
    const float mr = center ? 0.f : mean * rstd;
    if (!center) {
      mean_gy += mr * mean_g;
      mean_g = 0.f;
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = sub + i * lpr;
        if (v < nvec) {
          float4 o;
          o.x = rstd * (g[i].x - yh[i].x * mean_gy - mean_g);
          o.y = rstd * (g[i].y - yh[i].y * mean_gy - mean_g);
          o.z = rstd * (g[i].z - yh[i].z * mean_gy - mean_g);
          o.w = rstd * (g[i].w - yh[i].w * mean_gy - mean_g);
          o.x += rs[i].x; o.y += rs[i].y; o.z += rs[i].z; o.w += rs[i].w;
          *(reinterpret_cast<float4*>(dx + row * C) + v) = o;
          if (dx_bf16) {
            op16x2 p0 = OP2_FROM_F32(o.x, o.y), p1 = OP2_FROM_F32(o.z, o.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0);
            pk.y = *reinterpret_cast<uint32_t*>(&p1);
            *(reinterpret_cast<uint2*>(dx_bf16 + row * C) + v) = pk;
          }
          a_dw[i].x += d[i].x * (yh[i].x + mr); a_dw[i].y += d[i].y * (yh[i].y + mr);
          a_dw[i].z += d[i].z * (yh[i].z + mr); a_dw[i].w += d[i].w * (yh[i].w + mr);
          a_db[i].x += d[i].x; a_db[i].y += d[i].y; a_db[i].z += d[i].z; a_db[i].w += d[i].w;
          a_cs[i].x += o.x; a_cs[i].y += o.y; a_cs[i].z += o.z; a_cs[i].w += o.w;
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
      atomicAdd(&s_acc[2 * C + c + 0], a_cs[i].x); atomicAdd(&s_acc[2 * C + c + 1], a_cs[i].y);
      atomicAdd(&s_acc[2 * C + c + 2], a_cs[i].z); atomicAdd(&s_acc[2 * C + c + 3], a_cs[i].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (dw) atomicAdd(dw + c, s_acc[c]);
    if (db) atomicAdd(db + c, s_acc[C + c]);
    if (colsum) atomicAdd(colsum + c, s_acc[2 * C + c]);
  }
}

// ---- wide rows (C >= 512): bulk-copy pipelined backward ----
// With one warp per 2-4 KB row the register-resident kernel above keeps only ~16 rows per SM in flight and its
// per-row load -> reduce -> store chain is exposed (3.2 TB/s at C = 512, 16 K rows).  Here a CTA streams tiles of R
// consecutive rows (x, dres, dn are contiguous in memory: three 1-D bulk copies per tile, completion on an mbarrier)
// through a two-stage shared-memory ring, 8 warps x 1 row per tile, 2 CTAs per SM: ~160 KB per SM is always in flight
// and the warps read their operands from shared memory.
template <int NV>
__global__ void __launch_bounds__(256, 2)
ln_bwd_wide_kernel(const bf16* __restrict__ dn, const float* __restrict__ x, const float* __restrict__ stats,
                   const float* __restrict__ w, const float* __restrict__ dres, float* __restrict__ dx, bf16* __restrict__ dx_bf16,
                   float* __restrict__ dw, float* __restrict__ db, float* __restrict__ colsum, int M, int C, int R) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t xb = (uint32_t)R * C * 4, nb = (uint32_t)R * C * 2;
  const uint32_t stage_bytes = 2 * xb + nb;  // [x | dres | dn]
  float* s_acc = reinterpret_cast<float*>(smem + 2 * stage_bytes);  // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_acc[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_sync();
  const int tiles = (M + R - 1) / R;
  auto issue = [&](int t, int s) {
    const int r0 = t * R;
    const int valid = min(R, M - r0);
    uint8_t* dst = smem + (size_t)s * stage_bytes;
    const uint32_t vx = (uint32_t)valid * C * 4, vn = (uint32_t)valid * C * 2;
    mbar_arrive_expect_tx(&full[s], vx + (dres ? vx : 0u) + vn);
    bulk_load_1d(dst, x + (size_t)r0 * C, vx, &full[s]);
    if (dres) bulk_load_1d(dst + xb, dres + (size_t)r0 * C, vx, &full[s]);
    bulk_load_1d(dst + 2 * xb, dn + (size_t)r0 * C, vn, &full[s]);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < tiles) issue(blockIdx.x, 0);
  const float invC = 1.f / (float)C;
  float4 wv[NV], a_dw[NV], a_db[NV], a_cs[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    wv[i] = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
    a_dw[i] = a_db[i] = a_cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int it = 0;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
    const int s = it & 1;
    if (threadIdx.x == 0 && t + (int)gridDim.x < tiles) issue(t + gridDim.x, s ^ 1);
    const long long row = (long long)t * R + warp;
    const bool valid = warp < R && row < M;
    float2 st = make_float2(0.f, 0.f);
    if (valid) st = __ldg(reinterpret_cast<const float2*>(stats + row * 2));
    mbar_wait(&full[s], (it >> 1) & 1);
    if (valid) {
      const uint8_t* base = smem + (size_t)s * stage_bytes;
      const float4* sx = reinterpret_cast<const float4*>(base) + (size_t)warp * (C >> 2);
      const float4* sr = reinterpret_cast<const float4*>(base + xb) + (size_t)warp * (C >> 2);
      const uint2* sd = reinterpret_cast<const uint2*>(base + 2 * xb) + (size_t)warp * (C >> 2);
      const float mean = st.x, rstd = st.y;
      float4 yh[NV], g[NV];
      float sg = 0.f, sgy = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
        const float4 xv = sx[v];
        const uint2 pk = sd[v];
        const float2 d01 = OP2_TO_F32(*reinterpret_cast<const op16x2*>(&pk.x));
        const float2 d23 = OP2_TO_F32(*reinterpret_cast<const op16x2*>(&pk.y));
        yh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        g[i] = make_float4(d01.x * wv[i].x, d01.y * wv[i].y, d23.x * wv[i].z, d23.y * wv[i].w);
        sg += g[i].x + g[i].y + g[i].z + g[i].w;
        sgy += g[i].x * yh[i].x + g[i].y * yh[i].y + g[i].z * yh[i].z + g[i].w * yh[i].w;
        a_dw[i].x += d01.x * yh[i].x; a_dw[i].y += d01.y * yh[i].y; a_dw[i].z += d23.x * yh[i].z; a_dw[i].w += d23.y * yh[i].w;
        a_db[i].x += d01.x; a_db[i].y += d01.y; a_db[i].z += d23.x; a_db[i].w += d23.y;
      }
      const float mean_g = warp_sum(sg) * invC;
      const float mean_gy = warp_sum(sgy) * invC;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
        float4 o;
        o.x = rstd * (g[i].x - yh[i].x * mean_gy - mean_g);
        o.y = rstd * (g[i].y - yh[i].y * mean_gy - mean_g);
        o.z = rstd * (g[i].z - yh[i].z * mean_gy - mean_g);
        o.w = rstd * (g[i].w - yh[i].w * mean_gy - mean_g);
        if (dres) {
          const float4 rs = sr[v];
          o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
        }
        *(reinterpret_cast<float4*>(dx + row * C) + v) = o;
        if (dx_bf16) {
          op16x2 p0 = OP2_FROM_F32(o.x, o.y), p1 = OP2_FROM_F32(o.z, o.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&p0);
          pk.y = *reinterpret_cast<uint32_t*>(&p1);
          *(reinterpret_cast<uint2*>(dx_bf16 + row * C) + v) = pk;
        }
        a_cs[i].x += o.x; a_cs[i].y += o.y; a_cs[i].z += o.z; a_cs[i].w += o.w;
      }
    }
    __syncthreads();  // stage s is refilled by the next-but-one issue
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    atomicAdd(&s_acc[c + 0], a_dw[i].x); atomicAdd(&s_acc[c + 1], a_dw[i].y);
    atomicAdd(&s_acc[c + 2], a_dw[i].z); atomicAdd(&s_acc[c + 3], a_dw[i].w);
    atomicAdd(&s_acc[C + c + 0], a_db[i].x); atomicAdd(&s_acc[C + c + 1], a_db[i].y);
    atomicAdd(&s_acc[C + c + 2], a_db[i].z); atomicAdd(&s_acc[C + c + 3], a_db[i].w);
    atomicAdd(&s_acc[2 * C + c + 0], a_cs[i].x); atomicAdd(&s_acc[2 * C + c + 1], a_cs[i].y);
    atomicAdd(&s_acc[2 * C + c + 2], a_cs[i].z); atomicAdd(&s_acc[2 * C + c + 3], a_cs[i].w);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (dw) atomicAdd(dw + c, s_acc[c]);
    if (db) atomicAdd(db + c, s_acc[C + c]);
    if (colsum) atomicAdd(colsum + c, s_acc[2 * C + c]);
  }
}

// C = 512 (the 29 L4 blocks of NAFNet-w64: 16 K rows of 2 KiB): the one-shot kernel above needs 2048 CTAs whose warps each
// handle ONE row, and spends its time on CTA turnover and on the dependent (load -> 2 shuffle reductions -> w / b reload ->
// store) chain of that row.  Here CTAs are persistent (2 per SM), the affine parameters live in registers, and a warp loads
// its NEXT row before it reduces the current one, so the L2 / HBM latency of row i + 1 hides behind the math of row i.
// Per-row arithmetic (summation order included) is that of ln_fwd_kernel<4, 1>: results are bit-identical.
__global__ void __launch_bounds__(kWarps * 32, 2)
ln_fwd_wide_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, bf16* __restrict__ out,
                   float* __restrict__ stats, int M, float eps, int center) {
  constexpr int C = 512, NV = 4;
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invC = 1.f / (float)C;
  float4 wv[NV], bv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    wv[i] = __ldg(reinterpret_cast<const float4*>(w) + lane + i * 32);
    bv[i] = b ? __ldg(reinterpret_cast<const float4*>(b) + lane + i * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int stride = gridDim.x * kWarps;
  int row = blockIdx.x * kWarps + warp;
  float4 cur[NV], nxt[NV];
  if (row < M) {
#pragma unroll
    for (int i = 0; i < NV; ++i) cur[i] = __ldg(reinterpret_cast<const float4*>(x + (size_t)row * C) + lane + i * 32);
  }
  for (; row < M; row += stride) {
    const int rn = row + stride;
    if (rn < M) {
#pragma unroll
      for (int i = 0; i < NV; ++i) nxt[i] = __ldg(reinterpret_cast<const float4*>(x + (size_t)rn * C) + lane + i * 32);
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) sum += cur[i].x + cur[i].y + cur[i].z + cur[i].w;
    const float mean = group_sum(sum, 32) * invC;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = cur[i].x - mean, bb = cur[i].y - mean, c = cur[i].z - mean, d = cur[i].w - mean;
      sq += a * a + bb * bb + c * c + d * d;
    }
    const float var = group_sum(sq, 32) * invC;
    const float rstd = 1.f / sqrtf(var + eps);
    const float shift = center ? mean : 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float o0 = (cur[i].x - shift) * rstd * wv[i].x + bv[i].x;
      const float o1 = (cur[i].y - shift) * rstd * wv[i].y + bv[i].y;
      const float o2 = (cur[i].z - shift) * rstd * wv[i].z + bv[i].z;
      const float o3 = (cur[i].w - shift) * rstd * wv[i].w + bv[i].w;
      op16x2 p0 = OP2_FROM_F32(o0, o1), p1 = OP2_FROM_F32(o2, o3);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      *(reinterpret_cast<uint2*>(out + (size_t)row * C) + lane + i * 32) = pk;
    }
    if (stats && lane == 0) *reinterpret_cast<float2*>(stats + (size_t)row * 2) = make_float2(mean, rstd);
#pragma unroll
    for (int i = 0; i < NV; ++i) cur[i] = nxt[i];
  }
}

inline int pick_lpr(int C) {
  const int nvec = C / 4;
  int lpr = 2;
  while (lpr < 32 && lpr < nvec) lpr <<= 1;
  return lpr;
}

}  // namespace

int ln_fwd_launch(const float* x, const float* w, const float* b, bf16* n_out, float* stats, int M, int C, float eps,
                  cudaStream_t st, int center) {
  DCPT_CHECK_ARG(M > 0 && C >= 8 && C % 8 == 0 && C <= 1024, DCPT_E_SHAPE, "layernorm: need C %% 8 == 0 and 8 <= C <= 1024 (C=%d M=%d)",
                 C, M);
  const int lpr = pick_lpr(C);
  const int nv = ceil_div(C / 4, lpr);
  const int rows_per_block = kWarps * (32 / lpr);
  // U row groups per warp iteration (see the kernel); the grid is sized so that every warp gets its U groups
  const int U = nv <= 1 ? 8 : (nv <= 2 ? 4 : 1);
  long long grid = ceil_div_ll(M, (long long)rows_per_block * U);
  if (grid < 1) grid = 1;
  DCPT_PROF(dcpt_prof_tag2("ln_fwd", M, C), 8.0 * M * C, 6.0 * M * C, st);
  static const int wide_mode = getenv("DCPT_LN_FWD_WIDE") ? atoi(getenv("DCPT_LN_FWD_WIDE")) : 2;  // CTAs per SM; 0 = one-shot kernel
  if (C == 512 && M >= 2048 && wide_mode > 0) {
    int g = wide_mode * dcpt_num_sms();
    if (g > ceil_div(M, kWarps)) g = ceil_div(M, kWarps);
    DCPT_CUDA(dcpt_launch_pdl(ln_fwd_wide_kernel, dim3(g), dim3(kWarps * 32), 0, st, x, w, b, n_out, stats, M, eps, center));
    DCPT_LAUNCH_CHECK();
    return 0;
  }
#define LN_FWD(NVV, UU) DCPT_CUDA(dcpt_launch_pdl(ln_fwd_kernel<NVV, UU>, dim3((unsigned)grid), dim3(kWarps * 32), 0, st, x, w, b, n_out, stats, M, C, lpr, eps, center))
  if (nv <= 1) LN_FWD(1, 8);
  else if (nv <= 2) LN_FWD(2, 4);
  else if (nv <= 4) LN_FWD(4, 1);
  else LN_FWD(8, 1);
#undef LN_FWD
  DCPT_LAUNCH_CHECK();
  return 0;
}

int ln_bwd_launch(const bf16* dn, const float* x, const float* stats, const float* w, const float* dres, float* dx,
                  bf16* dx_bf16, float* dw, float* db, float* colsum, int M, int C, cudaStream_t st, int center) {
  DCPT_CHECK_ARG(M > 0 && C >= 8 && C % 8 == 0 && C <= 1024, DCPT_E_SHAPE, "layernorm bwd: need C %% 8 == 0 and 8 <= C <= 1024 (C=%d)", C);
  if (C == 512 && M >= 2048 && center && getenv("DCPT_LN_BWD_NARROW") == nullptr) {
    const int R = 8;  // rows per tile: 40 KB per stage
    const size_t smem = 128 + (size_t)2 * R * C * 10 + (size_t)3 * C * sizeof(float);
    const int tiles = ceil_div(M, R);
    int grid = 2 * dcpt_num_sms();
    if (grid > tiles) grid = tiles;
    DCPT_PROF(dcpt_prof_tag2("ln_bwd", M, C), 20.0 * M * C, (2.0 + 4.0 + (dres ? 4.0 : 0.0) + 4.0 + (dx_bf16 ? 2.0 : 0.0)) * M * C, st);
    if (C == 512) {
      static bool once = false;
      if (!once) { DCPT_CUDA(cudaFuncSetAttribute(ln_bwd_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); once = true; }
      DCPT_CUDA(dcpt_launch_pdl(ln_bwd_wide_kernel<4>, dim3(grid), dim3(256), smem, st, dn, x, stats, w, dres, dx, dx_bf16, dw, db, colsum, M, C, R));
    }
    DCPT_LAUNCH_CHECK();
    return 0;
  }
  const int lpr = pick_lpr(C);
  const int nv = ceil_div(C / 4, lpr);
  const int warps = nv <= 2 ? 16 : 4;
  const int rows_per_block = warps * (32 / lpr);
  long long grid = ceil_div_ll(M, rows_per_block);
  const long long cap = (long long)dcpt_num_sms() * (nv <= 1 ? 2 : (nv <= 2 ? 1 : (nv <= 4 ? 4 : 2)));  // resident blocks / SM
  if (grid > cap) grid = cap;
  const size_t smem = (size_t)3 * C * sizeof(float);
  DCPT_PROF(dcpt_prof_tag2("ln_bwd", M, C), 20.0 * M * C, (2.0 + 4.0 + (dres ? 4.0 : 0.0) + 4.0 + (dx_bf16 ? 2.0 : 0.0)) * M * C, st);
#define LN_BWD(NVV, WW, MB) \
  DCPT_CUDA(dcpt_launch_pdl(ln_bwd_kernel<NVV, WW, MB>, dim3((unsigned)grid), dim3(WW * 32), smem, st, dn, x, stats, w, dres, dx, dx_bf16, dw, db, colsum, M, C, lpr, center))
  if (nv <= 1) LN_BWD(1, 16, 2);
  else if (nv <= 2) LN_BWD(2, 16, 1);
  else if (nv <= 4) LN_BWD(4, 4, 4);
  else LN_BWD(8, 4, 2);
#undef LN_BWD
  DCPT_LAUNCH_CHECK();
  return 0;
}
