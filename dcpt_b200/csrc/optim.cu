// Fused parameter update of the training step: clip_grad_norm_ + Adam / AdamW + EMA in two passes
// over the parameters instead of the reference's ~10 eager passes and its 664-tensor Python EMA loop.
//
// Reference path: SRModel.optimize_parameters (basicsr/models/sr_model.py:164-174): `clip_grad_norm_` (optional) ->
// `optimizer_g.step()` (torch.optim.Adam / AdamW from BaseModel.get_optimizer, base_model.py:120-139) -> `model_ema`
// (base_model.py:86-95: ema.mul_(decay).add_(param, alpha=1-decay) per parameter).
//
// Kernels (both HBM-bound, multi-tensor: one launch covers every parameter tensor through a chunk table):
//   grad_sqsum_kernel     : per-chunk sum of squares of the gradients -> partial[chunk]          (4 B / parameter)
//   grad_norm_final_kernel: deterministic reduction of the partials   -> total_norm             (tiny)
//   adam_ema_kernel       : p, m, v, ema updated in place; reads g                              (36 B / parameter with EMA)
// A chunk is <= CHUNK consecutive elements of ONE tensor (16-byte vector path when all five pointers of the tensor are
// 16-byte aligned, which torch allocations and the flat gradient buffer's 256-byte-aligned views are).
#include <math.h>

#include <new>
#include <vector>

#include "../../include/dcpt_ops.h"
#include "common.cuh"

namespace {

constexpr int CHUNK = 8192;  // elements per CTA: 256 threads x 8 float4
constexpr int THREADS = 256;

struct TensorPtrs {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* ema;
};
struct Chunk {
  long long start;
  int tensor;
  int count;
};

struct StepArgs {
  float decay_mul;   // AdamW: 1 - lr * wd (applied to p first); Adam: 1
  float l2;          // Adam: wd (g += wd * p); AdamW: 0
  float w1;          // 1 - beta1
  float beta2, w2;   // beta2, 1 - beta2
  float inv_bc2_sqrt_den;  // bias_correction2 ** 0.5 (a divisor, as torch divides)
  float eps;
  float neg_step_size;  // -(lr / bias_correction1)
  float max_norm;       // <= 0: no clipping
  float ema_decay, ema_w;  // ema = ema * decay + p * (1 - decay); decay <= 0: no EMA
};

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < THREADS / 32 ? s_red[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;  // valid in warp 0
}

__global__ void __launch_bounds__(THREADS) grad_sqsum_kernel(const TensorPtrs* __restrict__ T, const Chunk* __restrict__ Cn,
                                                             float* __restrict__ partial) {
  __shared__ float s_red[THREADS / 32];
  const Chunk c = Cn[blockIdx.x];
  const float* g = T[c.tensor].g + c.start;
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = c.count >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int i = threadIdx.x; i < n4; i += THREADS) {
      const float4 x = __ldg(g4 + i);
      acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < c.count; i += THREADS) acc += g[i] * g[i];
  } else {
    for (int i = threadIdx.x; i < c.count; i += THREADS) acc += g[i] * g[i];
  }
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// one CTA, fixed summation order (double accumulation): total_norm = sqrt(sum of squares)
__global__ void __launch_bounds__(THREADS) grad_norm_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ norm,
                                                                  float* __restrict__ norm_out) {
  __shared__ double s_d[THREADS];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += THREADS) acc += (double)partial[i];
  s_d[threadIdx.x] = acc;
  __syncthreads();
  for (int o = THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_d[threadIdx.x] += s_d[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float t = (float)sqrt(s_d[0]);
    *norm = t;
    if (norm_out) *norm_out = t;
  }
}

// torch.optim.adam._single_tensor_adam, element by element, in its operation order (fp32):
//   AdamW: p *= 1 - lr*wd            Adam: g += wd * p
//   m.lerp_(g, 1 - b1)  ->  m + (1-b1) * (g - m)
//   v = v * b2 + (1-b2) * g * g
//   denom = sqrt(v) / sqrt(bc2) + eps;  p += -(lr / bc1) * (m / denom)
// preceded by clip_grad_norm_'s g *= min(1, max_norm / (total_norm + 1e-6)) and followed by model_ema's
// ema = ema * decay + (1 - decay) * p.
__device__ __forceinline__ void update1(float& p, float g, float& m, float& v, float& e, const StepArgs& a, float clip, bool has_ema) {
  g *= clip;
  p *= a.decay_mul;
  g = fmaf(a.l2, p, g);
  m = m + a.w1 * (g - m);
  v = v * a.beta2 + a.w2 * g * g;
  const float denom = sqrtf(v) / a.inv_bc2_sqrt_den + a.eps;
  p = p + a.neg_step_size * (m / denom);
  if (has_ema) e = e * a.ema_decay + a.ema_w * p;
}

__global__ void __launch_bounds__(THREADS) adam_ema_kernel(const TensorPtrs* __restrict__ T, const Chunk* __restrict__ Cn,
                                                           const float* __restrict__ norm, const StepArgs a) {
  const Chunk c = Cn[blockIdx.x];
  const TensorPtrs t = T[c.tensor];
  float clip = 1.f;
  if (a.max_norm > 0.f) clip = fminf(a.max_norm / (*norm + 1e-6f), 1.f);
  const bool has_ema = a.ema_decay > 0.f && t.ema != nullptr;
  float* p = t.p + c.start;
  const float* g = t.g + c.start;
  float* m = t.m + c.start;
  float* v = t.v + c.start;
  float* e = has_ema ? t.ema + c.start : nullptr;
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(e);
  int done = 0;
  if ((al & 15) == 0) {
    const int n4 = c.count >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    float4* e4 = reinterpret_cast<float4*>(e);
    // two vectors in flight per thread: 10 independent 16-byte loads before the first dependent store
    for (int i = threadIdx.x; i < n4; i += 2 * THREADS) {
      const int j = i + THREADS;
      const bool two = j < n4;
      float4 P0 = p4[i], G0 = __ldg(g4 + i), M0 = m4[i], V0 = v4[i], E0 = has_ema ? e4[i] : make_float4(0, 0, 0, 0);
      float4 P1 = P0, G1 = G0, M1 = M0, V1 = V0, E1 = E0;
      if (two) {
        P1 = p4[j]; G1 = __ldg(g4 + j); M1 = m4[j]; V1 = v4[j];
        if (has_ema) E1 = e4[j];
      }
      update1(P0.x, G0.x, M0.x, V0.x, E0.x, a, clip, has_ema);
      update1(P0.y, G0.y, M0.y, V0.y, E0.y, a, clip, has_ema);
      update1(P0.z, G0.z, M0.z, V0.z, E0.z, a, clip, has_ema);
      update1(P0.w, G0.w, M0.w, V0.w, E0.w, a, clip, has_ema);
      p4[i] = P0; m4[i] = M0; v4[i] = V0;
      if (has_ema) e4[i] = E0;
      if (two) {
        update1(P1.x, G1.x, M1.x, V1.x, E1.x, a, clip, has_ema);
        update1(P1.y, G1.y, M1.y, V1.y, E1.y, a, clip, has_ema);
        update1(P1.z, G1.z, M1.z, V1.z, E1.z, a, clip, has_ema);
        update1(P1.w, G1.w, M1.w, V1.w, E1.w, a, clip, has_ema);
        p4[j] = P1; m4[j] = M1; v4[j] = V1;
        if (has_ema) e4[j] = E1;
      }
    }
    done = n4 << 2;
  }
  for (int i = done + threadIdx.x; i < c.count; i += THREADS) {
    float P = p[i], M = m[i], V = v[i], E = has_ema ? e[i] : 0.f;
    update1(P, g[i], M, V, E, a, clip, has_ema);
    p[i] = P; m[i] = M; v[i] = V;
    if (has_ema) e[i] = E;
  }
}

// Order-independent 64-bit fingerprint of the bound PARAMETER tensors: sum over all elements of bits(p_i) * odd(position) mod 2^64.
// Any single changed bit changes the sum; used by the engines to notice parameter writes that bypass torch's version counter
// (`p.data.mul_()` in the reference's model_ema, base_model.py:86-95) before re-using the packed bf16 operand cache.
__global__ void __launch_bounds__(THREADS) param_hash_kernel(const TensorPtrs* __restrict__ T, const Chunk* __restrict__ Cn,
                                                             unsigned long long* __restrict__ out) {
  __shared__ unsigned long long s_h[THREADS / 32];
  const Chunk c = Cn[blockIdx.x];
  const uint32_t* p = reinterpret_cast<const uint32_t*>(T[c.tensor].p + c.start);
  const unsigned long long base = ((unsigned long long)c.tensor << 40) + (unsigned long long)c.start;
  unsigned long long h = 0;
  for (int i = threadIdx.x; i < c.count; i += THREADS)
    h += (unsigned long long)p[i] * (((base + (unsigned long long)i) * 0x9E3779B97F4A7C15ull) | 1ull);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  if ((threadIdx.x & 31) == 0) s_h[threadIdx.x >> 5] = h;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < THREADS / 32; ++w) t += s_h[w];
    atomicAdd(out, t);
  }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

struct dcpt_optim_plan {
  std::vector<long long> numels;
  std::vector<Chunk> chunks;
  size_t off_tensors, off_chunks, off_partial, off_norm, total;
};

extern "C" {

dcpt_optim_plan* dcpt_optim_create(const long long* host_numels, int n_tensors) {
  if (!host_numels || n_tensors <= 0) {
    dcpt_set_error("optim_create: need at least one tensor");
    return nullptr;
  }
  dcpt_optim_plan* pl = new (std::nothrow) dcpt_optim_plan();
  if (!pl) return nullptr;
  for (int t = 0; t < n_tensors; ++t) {
    if (host_numels[t] < 0) {
      dcpt_set_error("optim_create: negative numel at %d", t);
      delete pl;
      return nullptr;
    }
    pl->numels.push_back(host_numels[t]);
    for (long long s = 0; s < host_numels[t]; s += CHUNK) {
      const long long rest = host_numels[t] - s;
      pl->chunks.push_back(Chunk{s, t, (int)(rest < CHUNK ? rest : CHUNK)});
    }
  }
  size_t o = 0;
  pl->off_tensors = o; o = align_up(o + sizeof(TensorPtrs) * n_tensors, 256);
  pl->off_chunks = o;  o = align_up(o + sizeof(Chunk) * pl->chunks.size(), 256);
  pl->off_partial = o; o = align_up(o + sizeof(float) * pl->chunks.size(), 256);
  pl->off_norm = o;    o = align_up(o + sizeof(float), 256);
  pl->total = o;
  return pl;
}

void dcpt_optim_destroy(dcpt_optim_plan* plan) { delete plan; }

size_t dcpt_optim_workspace_bytes(const dcpt_optim_plan* plan) { return plan ? plan->total : 0; }

long long dcpt_optim_num_chunks(const dcpt_optim_plan* plan) { return plan ? (long long)plan->chunks.size() : 0; }

int dcpt_optim_bind(const dcpt_optim_plan* plan, void* workspace, float* const* host_params, const float* const* host_grads,
                    float* const* host_exp_avg, float* const* host_exp_avg_sq, float* const* host_ema, dcpt_stream_t stream) {
  DCPT_CHECK_ARG(plan && workspace && host_params && host_grads && host_exp_avg && host_exp_avg_sq, DCPT_E_ARG, "optim_bind: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = (int)plan->numels.size();
  std::vector<TensorPtrs> tp(n);
  for (int t = 0; t < n; ++t) {
    tp[t] = TensorPtrs{host_params[t], host_grads[t], host_exp_avg[t], host_exp_avg_sq[t], host_ema ? host_ema[t] : nullptr};
    DCPT_CHECK_ARG(plan->numels[t] == 0 || (tp[t].p && tp[t].g && tp[t].m && tp[t].v), DCPT_E_ARG, "optim_bind: null pointer for tensor %d", t);
    DCPT_CHECK_ARG(((reinterpret_cast<uintptr_t>(tp[t].p) | reinterpret_cast<uintptr_t>(tp[t].g) | reinterpret_cast<uintptr_t>(tp[t].m) |
                     reinterpret_cast<uintptr_t>(tp[t].v) | reinterpret_cast<uintptr_t>(tp[t].ema)) & 3) == 0,
                   DCPT_E_ALIGN, "optim_bind: tensor %d is not 4-byte aligned", t);
  }
  char* w = static_cast<char*>(workspace);
  // pageable host memory: the copies are staged by the runtime before the call returns (the vectors may die afterwards)
  DCPT_CUDA(cudaMemcpyAsync(w + plan->off_tensors, tp.data(), sizeof(TensorPtrs) * n, cudaMemcpyHostToDevice, st));
  DCPT_CUDA(cudaMemcpyAsync(w + plan->off_chunks, plan->chunks.data(), sizeof(Chunk) * plan->chunks.size(), cudaMemcpyHostToDevice, st));
  return 0;
}

int dcpt_optim_grad_norm(const dcpt_optim_plan* plan, void* workspace, float* total_norm, dcpt_stream_t stream) {
  DCPT_CHECK_ARG(plan && workspace, DCPT_E_ARG, "optim_grad_norm: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  const int nc = (int)plan->chunks.size();
  long long total = 0;
  for (long long v : plan->numels) total += v;
  float* partial = reinterpret_cast<float*>(w + plan->off_partial);
  float* norm = reinterpret_cast<float*>(w + plan->off_norm);
  if (nc > 0) {
    DCPT_PROF("optim_grad_sqsum", 2.0 * total, 4.0 * total, st);
    grad_sqsum_kernel<<<nc, THREADS, 0, st>>>(reinterpret_cast<const TensorPtrs*>(w + plan->off_tensors),
                                              reinterpret_cast<const Chunk*>(w + plan->off_chunks), partial);
    DCPT_LAUNCH_CHECK();
  }
  grad_norm_final_kernel<<<1, THREADS, 0, st>>>(partial, nc, norm, total_norm);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int dcpt_optim_set_norm(const dcpt_optim_plan* plan, void* workspace, const float* total_norm, dcpt_stream_t stream) {
  DCPT_CHECK_ARG(plan && workspace && total_norm, DCPT_E_ARG, "optim_set_norm: null argument");
  DCPT_CUDA(cudaMemcpyAsync(static_cast<char*>(workspace) + plan->off_norm, total_norm, sizeof(float), cudaMemcpyDeviceToDevice,
                            static_cast<cudaStream_t>(stream)));
  return 0;
}

int dcpt_optim_param_hash(const dcpt_optim_plan* plan, void* workspace, unsigned long long* hash, dcpt_stream_t stream) {
  DCPT_CHECK_ARG(plan && workspace && hash, DCPT_E_ARG, "optim_param_hash: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  const int nc = (int)plan->chunks.size();
  DCPT_CUDA(cudaMemsetAsync(hash, 0, sizeof(unsigned long long), st));
  if (nc == 0) return 0;
  long long total = 0;
  for (long long v : plan->numels) total += v;
  DCPT_PROF("optim_param_hash", 0.0, 4.0 * total, st);
  param_hash_kernel<<<nc, THREADS, 0, st>>>(reinterpret_cast<const TensorPtrs*>(w + plan->off_tensors),
                                            reinterpret_cast<const Chunk*>(w + plan->off_chunks), hash);
  DCPT_LAUNCH_CHECK();
  return 0;
}

int dcpt_optim_step(const dcpt_optim_plan* plan, void* workspace, int decoupled_weight_decay, double lr, double beta1, double beta2,
                    double eps, double weight_decay, long long step, double max_norm, double ema_decay, dcpt_stream_t stream) {
  DCPT_CHECK_ARG(plan && workspace, DCPT_E_ARG, "optim_step: null argument");
  DCPT_CHECK_ARG(step >= 1, DCPT_E_ARG, "optim_step: step = %lld must be >= 1 (the value AFTER torch's step += 1)", step);
  DCPT_CHECK_ARG(lr >= 0 && eps >= 0 && beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1 && weight_decay >= 0, DCPT_E_ARG,
                 "optim_step: invalid hyper-parameters (lr %g betas %g %g eps %g wd %g)", lr, beta1, beta2, eps, weight_decay);
  DCPT_CHECK_ARG(ema_decay < 1, DCPT_E_ARG, "optim_step: ema_decay = %g must be < 1", ema_decay);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  const int nc = (int)plan->chunks.size();
  if (nc == 0) return 0;
  // scalars exactly as torch computes them (Python doubles), then narrowed to the fp32 the tensor ops run in
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  StepArgs a;
  a.decay_mul = (decoupled_weight_decay && weight_decay != 0) ? (float)(1.0 - lr * weight_decay) : 1.f;
  a.l2 = (!decoupled_weight_decay && weight_decay != 0) ? (float)weight_decay : 0.f;
  a.w1 = (float)(1.0 - beta1);
  a.beta2 = (float)beta2;
  a.w2 = (float)(1.0 - beta2);
  a.inv_bc2_sqrt_den = (float)sqrt(bc2);
  a.eps = (float)eps;
  a.neg_step_size = (float)(-(lr / bc1));
  a.max_norm = max_norm > 0 ? (float)max_norm : 0.f;
  a.ema_decay = ema_decay > 0 ? (float)ema_decay : 0.f;
  a.ema_w = ema_decay > 0 ? (float)(1.0 - ema_decay) : 0.f;
  long long total = 0;
  for (long long v : plan->numels) total += v;
  DCPT_PROF("optim_adam_ema", 20.0 * total, (ema_decay > 0 ? 36.0 : 28.0) * total, st);
  adam_ema_kernel<<<nc, THREADS, 0, st>>>(reinterpret_cast<const TensorPtrs*>(w + plan->off_tensors),
                                          reinterpret_cast<const Chunk*>(w + plan->off_chunks),
                                          reinterpret_cast<const float*>(w + plan->off_norm), a);
  DCPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
