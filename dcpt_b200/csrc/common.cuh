// Common device/host helpers for the dcpt_b200 sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "prof.h"

// The 16-bit tensor-core OPERAND type.  Default build: bfloat16 (the fast path the bench measures).  -DDCPT_OPERAND_FP16
// (python -m dcpt_b200.build --fp16 -> libdcpt_sm100_fp16.so, selected with DCPT_OPERAND=fp16) stores every branch tensor and
// packed weight as IEEE half instead: 11 significand bits instead of 8, i.e. 8x smaller operand rounding, which is what the
// north-star's 1e-3 parity bar needs on deep stacks (DESIGN.md numerics).  Same layouts, same kernels, same tcgen05 kind::f16
// instruction (its descriptor carries the operand format); the type keeps the historical name `bf16` throughout the sources.
#ifdef DCPT_OPERAND_FP16
#include <cuda_fp16.h>
typedef __half bf16;
typedef __half2 op16x2;
#define OP_FROM_F32(x) __float2half_rn(x)
#define OP_TO_F32(x) __half2float(x)
#define OP2_FROM_F32(a, b) __floats2half2_rn((a), (b))
#define OP2_TO_F32(v) __half22float2(v)
#define DCPT_TMAP_OP16 CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define DCPT_UMMA_FMT 0u /* kind::f16 a_format / b_format: F16 */
#define DCPT_OPERAND_ID 1
#else
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 op16x2;
#define OP_FROM_F32(x) __float2bfloat16_rn(x)
#define OP_TO_F32(x) __bfloat162float(x)
#define OP2_FROM_F32(a, b) __floats2bfloat162_rn((a), (b))
#define OP2_TO_F32(v) __bfloat1622float2(v)
#define DCPT_TMAP_OP16 CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define DCPT_UMMA_FMT 1u /* kind::f16 a_format / b_format: BF16 */
#define DCPT_OPERAND_ID 0
#endif

// ----------------------------------------------------------------------------
// error plumbing (no C++ exception ever crosses the C ABI; see include/dcpt_ops.h)
// ----------------------------------------------------------------------------
#define DCPT_OK 0
#define DCPT_E_ARG (-1)
#define DCPT_E_SHAPE (-2)
#define DCPT_E_ALIGN (-3)
#define DCPT_E_DRIVER (-4)
#define DCPT_E_UNSUPPORTED (-5)

void dcpt_set_error(const char* fmt, ...);

#define DCPT_CHECK_ARG(cond, code, ...)     \
  do {                                      \
    if (!(cond)) {                          \
      dcpt_set_error(__VA_ARGS__);          \
      return (code);                        \
    }                                       \
  } while (0)

#define DCPT_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      dcpt_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

#define DCPT_TRY(expr)        \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

#define DCPT_LAUNCH_CHECK() DCPT_CUDA(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int dcpt_num_sms();

// Programmatic dependent launch (PDL): a kernel launched through dcpt_launch_pdl may become resident while its predecessor
// in the stream is still draining; it must call pdl_sync() (device) before its first access to global memory that another
// kernel produces or still reads.  The launch latency and the prologue (barrier init, TMEM allocation, descriptor
// prefetch) of kernel i+1 then overlap the tail of kernel i.  Opt-in with DCPT_PDL=1 (default: plain stream order).
bool dcpt_pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t dcpt_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = dcpt_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// Host-side TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point; gemm_sm100.cu).
// 2-D: row-major bf16 [rows, cols], box 64 x box_rows, 128-byte swizzle (UMMA operand tiles).
int make_tmap_2d(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_rows);
// Epilogue tiles: row-major [rows, cols] of fp32 (elem_bytes 4, 128-byte swizzle) or bf16 (elem_bytes 2, 64-byte swizzle),
// box = 32 columns x 32 rows.
int make_tmap_epi(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int elem_bytes);
// PixelShuffle(2) view of an NHWC tensor [NH, 2, W, 2, Cseg] (= [N, 2H, 2W, Cseg]) as a 5-D map (c, j, w, i, nh); box = 32 channels x
// 1 x bw x 1 x bh with bw * bh = 32: the 32 GEMM rows (consecutive input pixels) of one (i, j) sub-pixel.
int make_tmap_pixshuf(CUtensorMap* tm, const void* ptr, long long NH, int W, int Cseg, int elem_bytes);
// 4-D: bf16 NHWC [N, H, W, CH], box 64 x box_w x box_h x 1 (stencil tiles / implicit-GEMM conv operands, zero-filled halo).
int make_tmap_nhwc(CUtensorMap* tm, const void* ptr, int N, int H, int W, int CH, int box_w, int box_h, int swizzle128 = 0);

// ----------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------
#ifdef __CUDACC__

// let the next kernel in the stream start launching, then wait until every kernel this one depends on has completed and
// its memory is visible (no-op when the kernel was launched without the PDL attribute)
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 8 bf16 <-> 8 floats (one 16-byte vector)
struct __align__(16) bf16x8 {
  op16x2 v[4];
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const op16x2* p = reinterpret_cast<const op16x2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = OP2_TO_F32(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  op16x2* p = reinterpret_cast<op16x2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = OP2_FROM_F32(f[2 * i], f[2 * i + 1]);
  return u;
}

__device__ __forceinline__ float bf16_round(float x) { return OP_TO_F32(OP_FROM_F32(x)); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// explicit shared-space 16-byte accesses (a pointer derived from an aligned-up uintptr_t is generic to the compiler)
__device__ __forceinline__ void sts_f4(uint32_t saddr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}

// ---------------------------- mbarrier ---------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must trap (-> cudaErrorLaunchFailure), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0 && globaltimer_ns() - t0 > 4000000000ull) __trap();
  }
}

// ------------------------------- TMA -----------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load global -> shared; c0 = coordinate along the contiguous dim, c1 = row.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tiled store shared -> global (bulk async group); out-of-range parts of the box are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// 5-D forms (PixelShuffle scatter: the 32-row slab of a GEMM tile is a strided box of the NHWC output)
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before a bulk store reads them
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts_u4(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 1-D bulk copy global -> shared (contiguous bytes; size and both addresses multiples of 16), completion on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 prefetch of a 2-D box (no shared-memory destination, no barrier).
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
// 4-D tiled load (NHWC activation tile with halo): coordinates (c, w, h, n); out-of-range -> zero fill.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ----------------------------- tcgen05 ---------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same shape as tmem_ld32 (thread t of the warp writes lane base_lane + t, 32 consecutive fp32 columns)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&r)[32]) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(r);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]),
        "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]),
        "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]),
        "r"(u[30]), "r"(u[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// named barrier among `nthreads` threads (whole warps) of the CTA
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#endif  // __CUDACC__
