// Host-side orchestration of the NAFNet hot path and the C ABI (include/dcpt_ops.h).
//
// NAFBlock (reference basicsr/archs/nafnet_arch.py:83-186) is executed as
//   LN1 -> [GEMM conv1 +bias] -> dw3x3+SimpleGate(+pool) -> SCA -> g*s -> [GEMM conv3 (beta folded) +bias +inp = y]
//   -> LN2 -> [GEMM conv4 +bias, SimpleGate epilogue] -> [GEMM conv5 (gamma folded) +bias +y = out]
// with fp32 residual stream (x, y, out), bf16 branch tensors, fp32 accumulation everywhere.
// beta / gamma are folded into the packed bf16 weights (W' = diag(beta) W), so the residual
// update is a plain epilogue add; their gradients are recovered from the wgrad GEMM of the
// *unscaled* weight (see wgrad_finish_resid_kernel in misc.cu).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/dcpt_ops.h"
#include "elementwise.cuh"
#include "gemm.cuh"

// ------------------------------------------------------------------------------------
// error state / device info
// ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void dcpt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- profiling (prof.h) ----
bool g_dcpt_prof_on = false;
bool g_dcpt_prof_shapes = false;
long long g_dcpt_launches = 0;
namespace {
struct ProfRec { const char* tag; double flops, bytes; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
const char* dcpt_prof_intern(const char* s) {
  static std::map<std::string, int> pool;
  return pool.emplace(s, 0).first->first.c_str();
}
void dcpt_prof_begin(const char* tag, double flops, double bytes, cudaStream_t st) {
  ProfRec r{tag, flops, bytes, prof_event(), prof_event()};
  cudaEventRecord(r.e0, st);
  g_prof_recs.push_back(r);
}
void dcpt_prof_end(cudaStream_t st) { if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().e1, st); }

bool dcpt_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DCPT_PDL");
    on = (e && e[0] == '1') ? 1 : 0;  // opt-in: measured neutral inside the CUDA graph (24.65 vs 24.50 ms / step), see DESIGN.md
  }
  return on != 0;
}

int dcpt_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  return sms;
}

namespace {

// bump allocator over a caller-provided arena (base == nullptr: size computation only)
struct Arena {
  char* base;
  size_t off;
  explicit Arena(void* b) : base(static_cast<char*>(b)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
  size_t size() const { return (off + 255) & ~(size_t)255; }
};

enum {  // NAFBlock parameter indices (named_parameters() order)
  P_BETA = 0, P_GAMMA, P_C1W, P_C1B, P_C2W, P_C2B, P_C3W, P_C3B, P_SCAW, P_SCAB, P_C4W, P_C4B, P_C5W, P_C5B,
  P_N1W, P_N1B, P_N2W, P_N2B
};

struct BlockPacked {
  bf16 *w1, *w1t, *w3b, *w3bt, *w4p, *w4t, *w5g, *w5gt, *wsca;
  float *b3b, *b4p, *b5g;
  BlockPacked(Arena& a, int C) {
    const size_t cc = (size_t)C * C;
    w1 = a.take<bf16>(2 * cc); w1t = a.take<bf16>(2 * cc);
    w3b = a.take<bf16>(cc); w3bt = a.take<bf16>(cc);
    w4p = a.take<bf16>(2 * cc); w4t = a.take<bf16>(2 * cc);
    w5g = a.take<bf16>(cc); w5gt = a.take<bf16>(cc);
    b3b = a.take<float>(C); b4p = a.take<float>(2 * C); b5g = a.take<float>(C);
    wsca = a.take<bf16>(cc);  // SCA matrix as a GEMM operand (TLC: per-pixel pooled vectors)
  }
};

struct BlockSaved {
  bf16 *n1, *u, *g, *gs, *n2, *x4, *sg;
  float *stats1, *stats2, *pool, *s, *y;
  BlockSaved(Arena& a, int N, int H, int W, int C) {
    const size_t M = (size_t)N * H * W;
    n1 = a.take<bf16>(M * C); stats1 = a.take<float>(M * 2);
    u = a.take<bf16>(M * 2 * C); g = a.take<bf16>(M * C);
    pool = a.take<float>((size_t)N * C); s = a.take<float>((size_t)N * C);
    gs = a.take<bf16>(M * C); y = a.take<float>(M * C);
    n2 = a.take<bf16>(M * C); stats2 = a.take<float>(M * 2);
    x4 = a.take<bf16>(M * 2 * C); sg = a.take<bf16>(M * C);
  }
};

// Accumulators a block backward adds into (split-K wgrad scratch for conv5 / conv3, column sums, the SCA ds): they must
// start at zero.  The network backward hands every block its own slice of ONE region cleared by ONE memset per pass
// (6 memset nodes per block on the captured graph otherwise); zeros == nullptr: the block clears wk's own buffers itself.
struct BlockZeros {
  float *G5, *G3, *Sy, *ds;
  static size_t floats(int N, int C) { return 2 * ((size_t)C * C + 64) + (size_t)C + 64 + (size_t)N * C + 64; }
  BlockZeros(float* base, int N, int C) {
    auto up = [](size_t v) { return (v + 63) / 64 * 64; };
    G5 = base; base += up((size_t)C * C);
    G3 = base; base += up((size_t)C * C);
    Sy = base; base += up((size_t)C);
    ds = base;
  }
};

struct BlockWork {
  float *G, *dy, *Sy, *ds, *t;
  bf16 *dx4, *dn, *dyT, *dgs, *du2, *du;
  BlockWork(Arena& a, int N, int H, int W, int C) {
    const size_t M = (size_t)N * H * W;
    G = a.take<float>((size_t)C * C);
    dx4 = a.take<bf16>(M * 2 * C); dn = a.take<bf16>(M * C);
    dy = a.take<float>(M * C); dyT = a.take<bf16>(M * C); Sy = a.take<float>(C);
    dgs = a.take<bf16>(M * C); ds = a.take<float>((size_t)N * C); t = a.take<float>((size_t)N * C);
    du2 = a.take<bf16>(M * 2 * C); du = a.take<bf16>(M * 2 * C);
  }
};

GemmArgs gemm_args(int M, int N, int K, const bf16* A, int lda, const bf16* B, int ldb, int epi) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.splits = 1; g.epi = epi;
  return g;
}

// wgrad: out[O, I] += dY[Mpx, O]^T * X[Mpx, I]   (both operands MN-major, split-K over pixels, fp32 atomics)
int wgrad_gemm(const bf16* dY, int O, const bf16* X, int I, float* out, int Mpx, cudaStream_t st, int ldx = 0) {
  GemmArgs g = gemm_args(O, I, Mpx, dY, O, X, ldx > 0 ? ldx : I, EPI_ATOMIC);
  g.a_mn = 1; g.b_mn = 1;
  const int bn = I > 128 ? 256 : (I > 64 ? 128 : 64);
  const int tiles = ceil_div(O, 128) * ceil_div(I, bn);
  const int num_kb = ceil_div(Mpx, 64);
  g.splits = gemm_auto_splits(tiles, num_kb);
  {  // experiment knobs: DCPT_WGRAD_SPLIT_MULT=k / DCPT_WGRAD_SPLIT_DIV=k (k x, 1/k x the one-wave split factor), DCPT_WGRAD_FINE=1 (one
     // work item per CTA).  Same-box A/B (r03k, ms per step): default 21.17-21.25; x2 21.43; x2 fine 21.6; x4 fine 23.0; /2 21.4;
     // /3 21.8 - the one-wave split with persistent CTAs is the optimum on either side.
    static int mult = -1, fine = -1;
    if (mult < 0) {
      const char* e = getenv("DCPT_WGRAD_SPLIT_MULT");
      mult = e ? atoi(e) : 1;
      if (mult < 1) mult = 1;
      const char* f = getenv("DCPT_WGRAD_FINE");
      fine = (f && f[0] == '1') ? 1 : 0;
    }
    g.splits *= mult;
    {
      static int dv = -1;
      if (dv < 0) {
        const char* e = getenv("DCPT_WGRAD_SPLIT_DIV");
        dv = e ? atoi(e) : 1;
        if (dv < 1) dv = 1;
      }
      g.splits = g.splits / dv > 0 ? g.splits / dv : 1;
    }
    if (g.splits > num_kb) g.splits = num_kb;
    g.fine_grid = fine;
  }
  g.ep.out_f32 = out; g.ep.ldo = I;
  return gemm_launch(g, st);
}

int nafblock_pack_impl(const float* const* P, BlockPacked& pk, int C, cudaStream_t st) {
  DCPT_TRY(pack_weight_launch(P[P_C1W], nullptr, pk.w1, 2 * C, C, PACK_PLAIN, st));
  DCPT_TRY(pack_weight_launch(P[P_C1W], nullptr, pk.w1t, 2 * C, C, PACK_T, st));
  DCPT_TRY(pack_weight_launch(P[P_C3W], P[P_BETA], pk.w3b, C, C, PACK_PLAIN, st));
  DCPT_TRY(pack_weight_launch(P[P_C3W], P[P_BETA], pk.w3bt, C, C, PACK_T, st));
  // conv4's output columns are interleaved so that the SimpleGate halves of a channel land in the same accumulator tile:
  // 32-wide pairs (TMA-tiled epilogue) when C % 32 == 0, else 8-wide pairs (register-staged epilogue)
  const int pair_mode = C % 32 == 0 ? PACK_PAIR32 : PACK_PAIR;
  DCPT_TRY(pack_weight_launch(P[P_C4W], nullptr, pk.w4p, 2 * C, C, pair_mode, st));
  DCPT_TRY(pack_weight_launch(P[P_C4W], nullptr, pk.w4t, 2 * C, C, PACK_T, st));
  DCPT_TRY(pack_weight_launch(P[P_C5W], P[P_GAMMA], pk.w5g, C, C, PACK_PLAIN, st));
  DCPT_TRY(pack_weight_launch(P[P_C5W], P[P_GAMMA], pk.w5gt, C, C, PACK_T, st));
  DCPT_TRY(pack_bias_launch(P[P_C3B], P[P_BETA], pk.b3b, C, PACK_PLAIN, st));
  DCPT_TRY(pack_bias_launch(P[P_C4B], nullptr, pk.b4p, 2 * C, pair_mode, st));
  DCPT_TRY(pack_bias_launch(P[P_C5B], P[P_GAMMA], pk.b5g, C, PACK_PLAIN, st));
  DCPT_TRY(pack_weight_launch(P[P_SCAW], nullptr, pk.wsca, C, C, PACK_PLAIN, st));
  return 0;
}

// A LayerNorm2d forward fused into the epilogue of the GEMM that PRODUCES its input rows (gemm.cuh, EpiParams::ln_*): the
// consumer's affine parameters and where its normalised bf16 rows / (mean, rstd) go.  n == nullptr: not fused.
struct LnFuse {
  const float* w;
  const float* b;
  bf16* n;
  float* stats;
};
constexpr float kLnEps = 1e-6f;  // LayerNorm2d default (nafnet_arch.py:57)

// SCA mat-vec and the row scaling in one launch (sca_scale_launch); DCPT_SCA_FUSE=0 keeps the two kernels
bool sca_fused() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DCPT_SCA_FUSE");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// DCPT_DWGATE_SPLIT=1: depthwise-conv weight gradient on the side stream instead of inside the gate backward.  OFF by default:
// measured 22.01 vs 21.38 ms per step (r03f, same box) - the chain kernel gets faster (122 instead of 174 registers, half the
// FMAs) but the extra pass over u and du2 (66 MB per C = 512 block) does not fit into the idle SM time the weight-gradient
// stream lives on; that budget (~4.5 ms per step) is already spent by the wgrad GEMMs.
bool dwgate_split() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DCPT_DWGATE_SPLIT");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on != 0;
}

bool ln_fuse_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DCPT_LN_FUSE");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
// can the GEMM writing a [*, C] fp32 tensor also run its consumer's LayerNorm?
bool ln_fusable(int C) { return ln_fuse_enabled() && gemm_ln_fusable(C) && C % 8 == 0; }

// ... and its backward into the epilogue of the dgrad GEMM that produces d(LN output) (DCPT_LNB_FUSE=0: separate ln_bwd kernel)
bool lnb_fusable(int C) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DCPT_LNB_FUSE");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  // C = 64: one 32-column chunk per epilogue warp and tile - the two passes cannot overlap anything and the fused launch is
  // slower than GEMM + ln_bwd (r02q: 268 vs 258 us at 1 M pixels); from C = 128 on it wins (121 vs 151, 68 vs 88, 46 vs 55 us)
  return on != 0 && gemm_ln_fusable(C) && C >= 128;
}

void set_ln_epilogue(GemmArgs& g, const LnFuse& f, int C) {
  g.ep.ln_w = f.w; g.ep.ln_b = f.b; g.ep.ln_out = f.n; g.ep.ld_ln = C; g.ep.ln_stats = f.stats; g.ep.ln_eps = kLnEps;
}

// ln1_done: the producer of `x` already wrote sv.n1 / sv.stats1 (fused LayerNorm epilogue); next: the LayerNorm of the
// consumer of `out` (the following block's norm1), fused into conv5's epilogue when supported.
int nafblock_fwd_impl(const float* const* P, const BlockPacked& pk, const float* x, float* out, bf16* out_bf16,
                      const BlockSaved& sv, int N, int H, int W, int C, cudaStream_t st, int tlc_kh = 0, int tlc_kw = 0,
                      bool ln1_done = false, const LnFuse* next = nullptr, bool pool_zeroed = false, bool keep = true) {
  const int HW = H * W, M = N * HW;
  constexpr float eps = kLnEps;
  // norm1 -> conv1 (+bias)
  if (!ln1_done) DCPT_TRY(ln_fwd_launch(x, P[P_N1W], P[P_N1B], sv.n1, sv.stats1, M, C, eps, st));
  {
    GemmArgs g = gemm_args(M, 2 * C, C, sv.n1, C, pk.w1, C, EPI_STORE);
    g.ep.out_bf16 = sv.u; g.ep.ldo = 2 * C; g.ep.bias = P[P_C1B];
    DCPT_TRY(gemm_launch(g, st));
  }
  // conv2 (dw 3x3) + SimpleGate, global average pool partial sums
  if (!pool_zeroed) DCPT_CUDA(cudaMemsetAsync(sv.pool, 0, (size_t)N * C * sizeof(float), st));
  DCPT_TRY(dwgate_fwd_launch(sv.u, P[P_C2W], P[P_C2B], sv.g, sv.pool, N, H, W, C, st));
  if (tlc_kh > 0 && (tlc_kh < H || tlc_kw < W)) {
    // TLC inference (NAFNet / Local_Base, arch_util.py:339-398): the pooled vector is a per-pixel box mean, so SCA becomes
    // a [pixels, C] x [C, C] GEMM and a per-element product.  Scratch: y (fp32 integral image), n1 (box means), u (s map).
    const int k1 = tlc_kh < H ? tlc_kh : H, k2 = tlc_kw < W ? tlc_kw : W;
    DCPT_TRY(tlc_boxmean_launch(sv.g, sv.y, sv.n1, N, H, W, C, k1, k2, st));
    GemmArgs g = gemm_args(M, C, C, sv.n1, C, pk.wsca, C, EPI_STORE);
    g.ep.out_bf16 = sv.u; g.ep.ldo = C; g.ep.bias = P[P_SCAB];
    DCPT_TRY(gemm_launch(g, st));
    DCPT_TRY(mul_bf16_launch(sv.g, sv.u, sv.gs, (long long)M * C, st));
  } else {
    // sca, x * sca(x)
    if (sca_fused()) {
      DCPT_TRY(sca_scale_launch(sv.pool, P[P_SCAW], P[P_SCAB], sv.g, sv.s, sv.gs, N, C, HW, st));
    } else {
      DCPT_TRY(sca_fwd_launch(sv.pool, P[P_SCAW], P[P_SCAB], sv.s, N, C, HW, st));
      DCPT_TRY(scale_rows_launch(sv.g, sv.s, sv.gs, N, HW, C, st));
    }
  }
  // conv3, y = inp + x*beta
  {
    GemmArgs g = gemm_args(M, C, C, sv.gs, C, pk.w3b, C, EPI_STORE);
    g.ep.out_f32 = sv.y; g.ep.ldo = C; g.ep.bias = pk.b3b; g.ep.resid = x; g.ep.ldr = C;
    if (ln_fusable(C)) set_ln_epilogue(g, LnFuse{P[P_N2W], P[P_N2B], sv.n2, sv.stats2}, C);  // norm2 in conv3's epilogue
    DCPT_TRY(gemm_launch(g, st));
  }
  // norm2 -> conv4 -> SimpleGate
  if (!ln_fusable(C)) DCPT_TRY(ln_fwd_launch(sv.y, P[P_N2W], P[P_N2B], sv.n2, sv.stats2, M, C, eps, st));
  {
    GemmArgs g = gemm_args(M, 2 * C, C, sv.n2, C, pk.w4p, C, C % 32 == 0 ? EPI_GATE_TMA : EPI_GATE);
    // (x4, conv4's output before the gate, is read by the backward only: an inference pass does not store it)
    g.ep.out_bf16 = (keep || C % 32 != 0) ? sv.x4 : nullptr; g.ep.ldo = 2 * C; g.ep.out2 = sv.sg; g.ep.ldo2 = C; g.ep.C = C; g.ep.bias = pk.b4p;
    DCPT_TRY(gemm_launch(g, st));
  }
  // conv5, out = y + x*gamma
  {
    GemmArgs g = gemm_args(M, C, C, sv.sg, C, pk.w5g, C, EPI_STORE);
    g.ep.out_f32 = out; g.ep.out_bf16 = out_bf16; g.ep.ldo = C; g.ep.bias = pk.b5g; g.ep.resid = sv.y; g.ep.ldr = C;
    if (next && next->n) set_ln_epilogue(g, *next, C);  // the next block's norm1
    DCPT_TRY(gemm_launch(g, st));
  }
  return 0;
}

// Side stream for the weight-gradient work of a block backward.  The four wgrad GEMMs and their finishing kernels only
// produce parameter gradients: nothing on the activation-gradient chain waits for them, and every kernel of that chain is
// short and latency-bound at C = 512 (tensor pipe ~33 % busy, ncu), so they are forked onto a second stream and run in the
// chain's idle slots.  Fork / join are events on the caller's stream (capturable: the side stream joins the CUDA graph);
// the side stream is joined before the block returns, so the ABI contract ("everything is ordered on `stream`") holds.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // 0-3, 5: forks; 4: join; 6, 7: lagged joins
  bool pending[2] = {false, false};  // a lagged join event (ev[6 + parity]) has been recorded and not waited for yet
  bool ok = false, tried = false;
  // created on the first (eager, un-captured) use; owned by a plan (one per engine: two engines on two streams never share
  // events) or, for the plan-less block ABI, by the calling thread
  void ensure() {
    if (tried) return;
    tried = true;
    const char* e = getenv("DCPT_WGRAD_STREAM");
    if (!(e && e[0] == '0') && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess) {
      ok = true;
      for (auto& v : ev) ok = ok && cudaEventCreateWithFlags(&v, cudaEventDisableTiming) == cudaSuccess;
    }
  }
  ~SideStream() {
    for (auto& v : ev)
      if (v) cudaEventDestroy(v);
    if (s) cudaStreamDestroy(s);
  }
};
SideStream& thread_side_stream() {
  static thread_local SideStream ss;
  return ss;
}

// Network backward: the side stream may lag ONE block behind the chain.  Consecutive blocks use two BlockWork arenas in turn
// (parity), a block waits at its START for the side work of the block before last (the previous user of its arena) instead
// of waiting at its END for its own - whose last wgrad GEMM was forked only one kernel earlier and, running at the lower
// priority, is rarely done by then.  DCPT_SIDE_LAG=0: join at the end of every block.
bool side_lag_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DCPT_SIDE_LAG");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
int side_join_all(SideStream& ss, cudaStream_t st) {
  for (int k = 0; k < 2; ++k)
    if (ss.pending[k]) {
      DCPT_CUDA(cudaStreamWaitEvent(st, ss.ev[6 + k], 0));
      ss.pending[k] = false;
    }
  return 0;
}

int nafblock_bwd_impl(const float* const* P, const BlockPacked& pk, const BlockSaved& sv, const float* x, const float* dout,
                      const bf16* doutT, const float* Sout, float* dx, bf16* dxT, float* Sx, float* const* G, const BlockWork& wk,
                      int N, int H, int W, int C, cudaStream_t st, const BlockZeros* zeros = nullptr, SideStream* side = nullptr,
                      int lag_parity = -1) {
  const int HW = H * W, M = N * HW;
  const size_t cc = (size_t)C * C;
  float* const G5 = zeros ? zeros->G5 : wk.G;
  float* const G3 = zeros ? zeros->G3 : wk.G;
  float* const Sy = zeros ? zeros->Sy : wk.Sy;
  float* const ds = zeros ? zeros->ds : wk.ds;
  SideStream& ss = side ? *side : thread_side_stream();
  ss.ensure();
  const bool fork = ss.ok && !g_dcpt_prof_on;   // per-kernel profiling keeps everything on one stream
  cudaStream_t sw = fork ? ss.s : st;           // stream of the weight-gradient work
  const bool lag = fork && lag_parity >= 0;
  if (lag && ss.pending[lag_parity]) {           // the side work that last read this arena (two blocks ago) must be done
    DCPT_CUDA(cudaStreamWaitEvent(st, ss.ev[6 + lag_parity], 0));
    ss.pending[lag_parity] = false;
  }
  auto fork_here = [&](int i) -> int {          // side stream waits for everything issued on `st` so far
    if (fork) {
      DCPT_CUDA(cudaEventRecord(ss.ev[i], st));
      DCPT_CUDA(cudaStreamWaitEvent(sw, ss.ev[i], 0));
    }
    return 0;
  };
  // ---- conv5 (gamma folded): wgrad, dgamma, dbias; dgrad fused with SimpleGate backward ----
  DCPT_TRY(fork_here(0));
  if (!zeros) DCPT_CUDA(cudaMemsetAsync(G5, 0, cc * sizeof(float), sw));
  DCPT_TRY(wgrad_gemm(doutT, C, sv.sg, C, G5, M, sw));
  DCPT_TRY(wgrad_finish_resid_launch(G5, P[P_C5W], P[P_C5B], P[P_GAMMA], Sout, G[P_C5W], G[P_C5B], G[P_GAMMA], C, C, sw));
  {
    GemmArgs g = gemm_args(M, C, C, doutT, C, pk.w5gt, C, EPI_GATE_BWD);
    g.ep.out_bf16 = wk.dx4; g.ep.ldo = 2 * C; g.ep.aux = sv.x4; g.ep.ldaux = 2 * C; g.ep.C = C;
    g.ep.colsum = G[P_C4B];  // conv4's bias gradient = column sums of d(x4), reduced in the GEMM epilogue
    DCPT_TRY(gemm_launch(g, st));
  }
  // ---- conv4: wgrad, dgrad ----
  DCPT_TRY(fork_here(1));
  DCPT_TRY(wgrad_gemm(wk.dx4, 2 * C, sv.n2, C, G[P_C4W], M, sw));
  const bool fuse_lnb = lnb_fusable(C);  // LayerNorm backward in the dgrad GEMM's epilogue (no dn round trip, one launch less)
  if (!zeros) DCPT_CUDA(cudaMemsetAsync(Sy, 0, C * sizeof(float), st));
  {
    GemmArgs g = gemm_args(M, C, 2 * C, wk.dx4, 2 * C, pk.w4t, 2 * C, EPI_STORE);
    if (fuse_lnb) {  // ---- conv4 dgrad + norm2 backward + residual: dy = dout + LN'(dn2) ----
      g.ep.out_f32 = wk.dy; g.ep.out_bf16 = wk.dyT; g.ep.ldo = C;
      g.ep.lnb_x = sv.y; g.ep.ld_lnb = C; g.ep.lnb_stats = sv.stats2; g.ep.lnb_w = P[P_N2W]; g.ep.lnb_dres = dout;
      g.ep.lnb_dw = G[P_N2W]; g.ep.lnb_db = G[P_N2B]; g.ep.lnb_cs = Sy;
    } else {
      g.ep.out_bf16 = wk.dn; g.ep.ldo = C;
    }
    DCPT_TRY(gemm_launch(g, st));
  }
  if (!fuse_lnb) DCPT_TRY(ln_bwd_launch(wk.dn, sv.y, sv.stats2, P[P_N2W], dout, wk.dy, wk.dyT, G[P_N2W], G[P_N2B], Sy, M, C, st));
  // ---- conv3 (beta folded) ----
  DCPT_TRY(fork_here(2));
  if (!zeros) DCPT_CUDA(cudaMemsetAsync(G3, 0, cc * sizeof(float), sw));
  DCPT_TRY(wgrad_gemm(wk.dyT, C, sv.gs, C, G3, M, sw));
  DCPT_TRY(wgrad_finish_resid_launch(G3, P[P_C3W], P[P_C3B], P[P_BETA], Sy, G[P_C3W], G[P_C3B], G[P_BETA], C, C, sw));
  if (!zeros) DCPT_CUDA(cudaMemsetAsync(ds, 0, (size_t)N * C * sizeof(float), st));
  {  // dgrad, with the SCA backward's ds[n, c] = sum_px d(g*s) * g reduced in the GEMM epilogue
    GemmArgs g = gemm_args(M, C, C, wk.dyT, C, pk.w3bt, C, EPI_STORE);
    g.ep.out_bf16 = wk.dgs; g.ep.ldo = C;
    g.ep.gaux = sv.g; g.ep.ldgaux = C; g.ep.colsum = ds; g.ep.rows_per_img = HW;
    DCPT_TRY(gemm_launch(g, st));
  }
  // ---- SCA backward ----
  DCPT_TRY(fork_here(5));
  const bool t_folded = sca_fused();  // the transposed SCA mat-vec runs inside the gate backward (no sca_bwd_t launch on the chain)
  DCPT_TRY(sca_bwd_launch(ds, sv.pool, P[P_SCAW], t_folded ? nullptr : wk.t, G[P_SCAW], G[P_SCAB], N, C, HW, st, sw));
  // ---- SimpleGate + depthwise conv backward ----
  // the depthwise conv's weight / bias gradient leaves the chain: the gate backward writes du2 only, and dW2 / db2 come from
  // (du2, u) on the weight-gradient stream (below, after the next fork)
  const bool dw_split = fork && dwgate_split();
  DCPT_TRY(dwgate_bwd_a_launch(wk.dgs, sv.s, t_folded ? nullptr : wk.t, sv.u, P[P_C2W], P[P_C2B], wk.du2, dw_split ? nullptr : G[P_C2W],
                               dw_split ? nullptr : G[P_C2B], N, H, W, C, st, ds, P[P_SCAW]));
  DCPT_TRY(dwconv_bwd_data_launch(wk.du2, P[P_C2W], wk.du, G[P_C1B], N, H, W, 2 * C, st));
  // ---- conv1 ----
  DCPT_TRY(fork_here(3));
  if (dw_split) DCPT_TRY(dwconv3_wgrad_launch(wk.du2, sv.u, G[P_C2W], N, H, W, 2 * C, sw, G[P_C2B]));
  DCPT_TRY(wgrad_gemm(wk.du, 2 * C, sv.n1, C, G[P_C1W], M, sw));
  {
    GemmArgs g = gemm_args(M, C, 2 * C, wk.du, 2 * C, pk.w1t, 2 * C, EPI_STORE);
    if (fuse_lnb) {  // ---- conv1 dgrad + norm1 backward + residual: dx = dy + LN'(dn1) ----
      g.ep.out_f32 = dx; g.ep.out_bf16 = dxT; g.ep.ldo = C;
      g.ep.lnb_x = x; g.ep.ld_lnb = C; g.ep.lnb_stats = sv.stats1; g.ep.lnb_w = P[P_N1W]; g.ep.lnb_dres = wk.dy;
      g.ep.lnb_dw = G[P_N1W]; g.ep.lnb_db = G[P_N1B]; g.ep.lnb_cs = Sx;
    } else {
      g.ep.out_bf16 = wk.dn; g.ep.ldo = C;
    }
    DCPT_TRY(gemm_launch(g, st));
  }
  if (!fuse_lnb) DCPT_TRY(ln_bwd_launch(wk.dn, x, sv.stats1, P[P_N1W], wk.dy, dx, dxT, G[P_N1W], G[P_N1B], Sx, M, C, st));
  if (lag) {   // the block after next waits for this (it reuses this block's arena)
    DCPT_CUDA(cudaEventRecord(ss.ev[6 + lag_parity], sw));
    ss.pending[lag_parity] = true;
  } else if (fork) {  // join: the block's buffers (dx4, dyT, du, G, Sy) are reused by the next block
    DCPT_CUDA(cudaEventRecord(ss.ev[4], sw));
    DCPT_CUDA(cudaStreamWaitEvent(st, ss.ev[4], 0));
  }
  return 0;
}

int check_ptrs16(const void* const* ptrs, int n, const char* what);

int check_block_shape(int N, int H, int W, int C) {
  DCPT_CHECK_ARG(N > 0 && H > 0 && W > 0, DCPT_E_SHAPE, "nafblock: bad spatial shape N=%d H=%d W=%d", N, H, W);
  DCPT_CHECK_ARG(C >= 8 && C % 8 == 0 && C <= 1024, DCPT_E_SHAPE, "nafblock: C=%d must be a multiple of 8 in [8, 1024]", C);
  DCPT_CHECK_ARG((long long)N * H * W * 2 * C < (1ll << 31), DCPT_E_SHAPE, "nafblock: tensor too large for 32-bit pixel index");
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------
// NAFNetBaseline plan
// ------------------------------------------------------------------------------------
struct dcpt_nafnet_plan {
  int img_channel, width, middle_blk_num;
  std::vector<int> tlc_kh, tlc_kw;  // per resolution level; empty = global SCA pooling (NAFNetBaseline)
  std::vector<int> enc, dec;
  struct ParamInfo { int dims[4]; long long numel; };
  std::vector<ParamInfo> params;
  struct Blk { int pidx; int C; int level; };  // level: resolution level (0 = full res)
  std::vector<std::vector<Blk>> enc_blks, dec_blks;
  std::vector<Blk> mid_blks;
  std::vector<int> up_pidx, down_pidx;  // ups.i.0.weight ; downs.i.weight (bias = +1)
  int mid_C;
  // which block of decoder level i delivers host_feats[i] / receives host_dfeats[i] (-1 = the level's last block).  DCPT's
  // one-dot rule hooks `decoder{i}.0`, the FIRST block (degradation_classification_pretrain_model.py:64-67).
  // Optional CUDA event recorded by dcpt_nafnet_bwd on its stream as soon as the gradients of the deepest encoder level, the
  // middle blocks, the up convs, every decoder and the ending conv are final (i.e. after the backward of encoders.{n_enc-1}.*):
  // a data-parallel caller starts the all-reduce of that contiguous 90 % slice of the flat gradient buffer on a second
  // stream while the shallower levels are still being differentiated (base_model.py:107-118: what DDP's buckets do).
  mutable SideStream side;  // weight-gradient side stream of this plan's backward (nafblock_bwd_impl)
  // false: the next forward is inference - tensors only the backward reads (conv4's pre-gate output) are not stored
  mutable bool keep_activations = true;
  void* split_event = nullptr;
  bool split_event_external = true;  // inside a stream capture: external-event node (waited on from outside the graph) or plain
  std::vector<int> hook_blk;
  int hook_block(int i) const { return (i < (int)hook_blk.size() && hook_blk[i] >= 0) ? hook_blk[i] : dec[i] - 1; }
};

namespace {

void add_param(dcpt_nafnet_plan* p, int d0, int d1 = 1, int d2 = 1, int d3 = 1) {
  dcpt_nafnet_plan::ParamInfo pi;
  pi.dims[0] = d0; pi.dims[1] = d1; pi.dims[2] = d2; pi.dims[3] = d3;
  pi.numel = (long long)d0 * d1 * d2 * d3;
  p->params.push_back(pi);
}

int add_block_params(dcpt_nafnet_plan* p, int c) {
  const int idx = (int)p->params.size();
  add_param(p, 1, c, 1, 1); add_param(p, 1, c, 1, 1);                // beta, gamma
  add_param(p, 2 * c, c, 1, 1); add_param(p, 2 * c);                 // conv1
  add_param(p, 2 * c, 1, 3, 3); add_param(p, 2 * c);                 // conv2
  add_param(p, c, c, 1, 1); add_param(p, c);                         // conv3
  add_param(p, c, c, 1, 1); add_param(p, c);                         // sca.1
  add_param(p, 2 * c, c, 1, 1); add_param(p, 2 * c);                 // conv4
  add_param(p, c, c, 1, 1); add_param(p, c);                         // conv5
  add_param(p, c); add_param(p, c); add_param(p, c); add_param(p, c);  // norm1, norm2
  return idx;
}

// Device-side layout of everything the forward keeps for the backward.
struct NetSaved {
  float* x0;                                   // intro output
  bf16* P;                                     // intro patch matrix [M0, 32] (im2col of the input image)
  bf16* xlast_bf16;                            // bf16 mirror of the last decoder output (ending conv operand)
  std::vector<std::vector<BlockSaved>> enc_sv, dec_sv;
  std::vector<std::vector<float*>> enc_out, dec_out;  // block outputs (fp32)
  std::vector<BlockSaved> mid_sv;
  std::vector<float*> mid_out;
  std::vector<bf16*> xu;                       // unshuffled down-conv inputs (bf16 [M/4, 4C])
  std::vector<float*> xd;                      // down-conv outputs
  std::vector<bf16*> up_in;                    // bf16 mirror of each up-conv input
  std::vector<float*> xup;                     // up-conv outputs (after skip add)
  NetSaved(const dcpt_nafnet_plan* p, Arena& a, int N, int H, int W) {
    int C = p->width, h = H, w = W;
    x0 = a.take<float>((size_t)N * h * w * C);
    P = a.take<bf16>((size_t)N * h * w * 64);
    xlast_bf16 = a.take<bf16>((size_t)N * h * w * C);
    const int ne = (int)p->enc.size(), nd = (int)p->dec.size();
    enc_sv.resize(ne); enc_out.resize(ne);
    for (int i = 0; i < ne; ++i) {
      for (int j = 0; j < p->enc[i]; ++j) {
        enc_sv[i].emplace_back(a, N, h, w, C);
        enc_out[i].push_back(a.take<float>((size_t)N * h * w * C));
      }
      xu.push_back(a.take<bf16>((size_t)N * h * w * C));
      h /= 2; w /= 2; C *= 2;
      xd.push_back(a.take<float>((size_t)N * h * w * C));
    }
    for (int j = 0; j < p->middle_blk_num; ++j) {
      mid_sv.emplace_back(a, N, h, w, C);
      mid_out.push_back(a.take<float>((size_t)N * h * w * C));
    }
    dec_sv.resize(nd); dec_out.resize(nd);
    for (int i = 0; i < nd; ++i) {
      up_in.push_back(a.take<bf16>((size_t)N * h * w * C));
      h *= 2; w *= 2; C /= 2;
      xup.push_back(a.take<float>((size_t)N * h * w * C));
      for (int j = 0; j < p->dec[i]; ++j) {
        dec_sv[i].emplace_back(a, N, h, w, C);
        dec_out[i].push_back(a.take<float>((size_t)N * h * w * C));
      }
    }
    // the SCA pooling sums of ALL blocks in one contiguous region, cleared by one memset per forward (they are accumulated
    // with atomics by dwgate_fwd); the per-block `pool` slots carved above stay unused
    size_t total = 0;
    auto count = [&](std::vector<BlockSaved>& v, int c) { total += v.size() * (((size_t)N * c + 63) / 64 * 64); };
    C = p->width;
    for (int i = 0; i < ne; ++i) { count(enc_sv[i], C); C *= 2; }
    count(mid_sv, C);
    for (int i = 0; i < nd; ++i) { C /= 2; count(dec_sv[i], C); }
    pools = a.take<float>(total);
    pool_floats = total;
    float* q = pools;
    auto assign = [&](std::vector<BlockSaved>& v, int c) {
      for (auto& b : v) { b.pool = q; if (q) q += ((size_t)N * c + 63) / 64 * 64; }
    };
    C = p->width;
    for (int i = 0; i < ne; ++i) { assign(enc_sv[i], C); C *= 2; }
    assign(mid_sv, C);
    for (int i = 0; i < nd; ++i) { C /= 2; assign(dec_sv[i], C); }
  }
  float* pools = nullptr;
  size_t pool_floats = 0;
};

struct NetPacked {
  std::vector<std::vector<BlockPacked>> enc_pk, dec_pk;
  std::vector<BlockPacked> mid_pk;
  std::vector<bf16*> down, down_t, up, up_t;
  bf16 *wi_p, *wd_p;  // intro weight [C][32] and ending weight as dgrad operand [C][32] (27 taps + zero pad)
  NetPacked(const dcpt_nafnet_plan* p, Arena& a) {
    int C = p->width;
    wi_p = a.take<bf16>((size_t)C * 64);
    wd_p = a.take<bf16>((size_t)C * 32);
    const int ne = (int)p->enc.size(), nd = (int)p->dec.size();
    enc_pk.resize(ne); dec_pk.resize(nd);
    for (int i = 0; i < ne; ++i) {
      for (int j = 0; j < p->enc[i]; ++j) enc_pk[i].emplace_back(a, C);
      down.push_back(a.take<bf16>((size_t)2 * C * 4 * C));
      down_t.push_back(a.take<bf16>((size_t)2 * C * 4 * C));
      C *= 2;
    }
    for (int j = 0; j < p->middle_blk_num; ++j) mid_pk.emplace_back(a, C);
    for (int i = 0; i < nd; ++i) {
      up.push_back(a.take<bf16>((size_t)2 * C * C));
      up_t.push_back(a.take<bf16>((size_t)2 * C * C));
      C /= 2;
      for (int j = 0; j < p->dec[i]; ++j) dec_pk[i].emplace_back(a, C);
    }
  }
};

struct NetWork {
  char* blk_base;   // BlockWork arena (sized for the largest level)
  char* blk_base2;  // second arena: consecutive blocks alternate, so the weight-gradient stream may lag one block behind
  size_t blk_bytes;
  float* dxa; float* dxb; bf16* dta; bf16* dtb; float* Sa; float* Sb;  // ping-pong gradient stream
  std::vector<float*> dskip;                                          // per encoder level
  bf16* dconv;      // unshuffled up-conv output gradient
  float* G;         // wgrad scratch for up/down/ending convs
  bf16* Pd;         // flipped patch matrix of dout [M0, 32]
  float* psum;      // column sums of Pd [32]
  NetWork(const dcpt_nafnet_plan* p, Arena& a, int N, int H, int W) {
    int C = p->width, h = H, w = W;
    size_t max_act = 0, max_blk = 0, max_g = (size_t)32 * C;
    const int ne = (int)p->enc.size();
    std::vector<size_t> lvl_act;
    int maxC = C;
    for (int i = 0; i <= ne; ++i) {
      const size_t act = (size_t)N * h * w * C;
      lvl_act.push_back(act);
      if (act > max_act) max_act = act;
      Arena probe(nullptr);
      BlockWork bw(probe, N, h, w, C);
      (void)bw;
      if (probe.size() > max_blk) max_blk = probe.size();
      if (i < ne) {
        const size_t g = (size_t)2 * C * 4 * C;  // down weight [2C, 4C]; up weight of the mirrored level has the same count
        if (g > max_g) max_g = g;
        h /= 2; w /= 2; C *= 2;
        if (C > maxC) maxC = C;
      }
    }
    blk_bytes = max_blk;
    blk_base = a.take<char>(max_blk);
    blk_base2 = a.take<char>(max_blk);
    dxa = a.take<float>(max_act); dxb = a.take<float>(max_act);
    dta = a.take<bf16>(max_act); dtb = a.take<bf16>(max_act);
    Sa = a.take<float>(maxC); Sb = a.take<float>(maxC);
    for (int i = 0; i < ne; ++i) dskip.push_back(a.take<float>(lvl_act[i]));
    dconv = a.take<bf16>(max_act);
    G = a.take<float>(max_g);
    Pd = a.take<bf16>((size_t)N * H * W * 32);
    psum = a.take<float>(32);
    // per-block accumulators (BlockZeros) + each block's column-sum output Sx, one region cleared once per backward
    zero_floats = 0;
    int Cz = p->width;
    auto add = [&](int nblk, int c) { zero_floats += (size_t)nblk * (BlockZeros::floats(N, c) + ((size_t)c + 63) / 64 * 64); };
    for (int i = 0; i < ne; ++i) { add(p->enc[i], Cz); Cz *= 2; }
    add(p->middle_blk_num, Cz);
    for (size_t i = 0; i < p->dec.size(); ++i) { Cz /= 2; add(p->dec[i], Cz); }
    zeros = a.take<float>(zero_floats);
  }
  float* zeros = nullptr;
  size_t zero_floats = 0;
};

int check_ptrs16(const void* const* ptrs, int n, const char* what) {
  DCPT_CHECK_ARG(ptrs != nullptr, DCPT_E_ARG, "%s: null pointer table", what);
  for (int i = 0; i < n; ++i)
    DCPT_CHECK_ARG(ptrs[i] != nullptr && (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, DCPT_E_ALIGN,
                   "%s[%d] = %p must be non-null and 16-byte aligned", what, i, ptrs[i]);
  return 0;
}

int check_net_shape(const dcpt_nafnet_plan* p, int N, int H, int W) {
  const int f = 1 << (int)p->enc.size();
  DCPT_CHECK_ARG(N > 0 && H > 0 && W > 0 && H % f == 0 && W % f == 0, DCPT_E_SHAPE,
                 "nafnet: H=%d W=%d must be positive multiples of %d (pad with SRModel.pre_test's window_size)", H, W, f);
  DCPT_CHECK_ARG(p->img_channel == 3, DCPT_E_UNSUPPORTED, "nafnet: img_channel=%d (only 3 is built)", p->img_channel);
  DCPT_CHECK_ARG((long long)N * H * W * 2 * p->width < (1ll << 31), DCPT_E_SHAPE, "nafnet: batch too large for 32-bit pixel index");
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
extern "C" {

int dcpt_abi_version(void) { return DCPT_ABI_VERSION; }

int dcpt_operand_dtype(void) { return DCPT_OPERAND_ID; }

long long dcpt_launch_count(void) { return g_dcpt_launches; }

int dcpt_prof_enable(int on) {
  g_dcpt_prof_on = on != 0;
  g_dcpt_prof_shapes = on == 2;
  return 0;
}

// Synchronises the device, aggregates the recorded launches by tag and writes one line per tag:
//   tag <TAB> launches <TAB> total_ms <TAB> total_flops <TAB> total_bytes
// Returns the number of bytes written (truncated to cap), clears the records.
long long dcpt_prof_dump(char* buf, long long cap) {
  cudaDeviceSynchronize();
  struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      Agg& a = agg[r.tag];
      a.n++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    }
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  g_prof_recs.clear();
  (void)cudaGetLastError();
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\t%.6e\t%.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops,
             kv.second.bytes);
    out += line;
  }
  long long n = (long long)out.size() < cap - 1 ? (long long)out.size() : cap - 1;
  if (buf && cap > 0) { memcpy(buf, out.data(), n); buf[n] = 0; }
  return n;
}
const char* dcpt_last_error(void) { return g_err; }

int dcpt_layernorm2d_fwd(const float* x, const float* weight, const float* bias, void* out_bf16, float* stats, int M, int C,
                         float eps, dcpt_stream_t stream) {
  return ln_fwd_launch(x, weight, bias, static_cast<bf16*>(out_bf16), stats, M, C, eps, static_cast<cudaStream_t>(stream));
}

int dcpt_layernorm2d_bwd(const void* dn_bf16, const float* x, const float* stats, const float* weight, const float* dres,
                         float* dx, void* dx_bf16, float* dweight, float* dbias, float* colsum, int M, int C,
                         dcpt_stream_t stream) {
  return ln_bwd_launch(static_cast<const bf16*>(dn_bf16), x, stats, weight, dres, dx, static_cast<bf16*>(dx_bf16), dweight, dbias,
                       colsum, M, C, static_cast<cudaStream_t>(stream));
}

int dcpt_gemm_ex(const dcpt_gemm_desc* d, int impl, dcpt_stream_t stream) {
  DCPT_CHECK_ARG(d != nullptr, DCPT_E_ARG, "gemm_ex: null descriptor");
  GemmArgs g = gemm_args(d->M, d->N, d->K, static_cast<const bf16*>(d->A), d->lda, static_cast<const bf16*>(d->B), d->ldb,
                         d->epilogue);
  g.a_mn = d->a_mn; g.b_mn = d->b_mn; g.splits = d->splits < 1 ? 1 : d->splits;
  if (d->splits == 0 && d->epilogue == EPI_ATOMIC) {  // auto split-K: fill the machine
    const int bn = d->N > 128 ? 256 : (d->N > 64 ? 128 : 64);
    const int tiles = ceil_div(d->M, 128) * ceil_div(d->N, bn);
    g.splits = gemm_auto_splits(tiles, ceil_div(d->K, 64));
  }
  g.ep.out_f32 = d->out_f32; g.ep.out_bf16 = static_cast<bf16*>(d->out_bf16); g.ep.ldo = d->ldo;
  g.ep.bias = d->bias; g.ep.resid = d->resid; g.ep.ldr = d->ldr;
  g.ep.out2 = static_cast<bf16*>(d->out2_bf16); g.ep.ldo2 = d->ldo2;
  g.ep.aux = static_cast<const bf16*>(d->aux_bf16); g.ep.ldaux = d->ldaux;
  g.ep.C = d->C; g.ep.H = d->H; g.ep.W = d->W; g.ep.Cseg = d->Cseg;
  g.ep.ln_w = d->ln_weight; g.ep.ln_b = d->ln_bias; g.ep.ln_out = static_cast<bf16*>(d->ln_out); g.ep.ld_ln = d->ld_ln;
  g.ep.ln_stats = d->ln_stats; g.ep.ln_eps = d->ln_eps; g.ep.ln_nocenter = d->ln_nocenter;
  g.ep.lnb_x = d->lnb_x; g.ep.ld_lnb = d->ld_lnb; g.ep.lnb_stats = d->lnb_stats; g.ep.lnb_w = d->lnb_weight; g.ep.lnb_dres = d->lnb_dres;
  g.ep.lnb_dw = d->lnb_dweight; g.ep.lnb_db = d->lnb_dbias; g.ep.lnb_cs = d->lnb_colsum;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return impl == 1 ? gemm_simt_launch(g, st) : gemm_tc_launch(g, st);
}

int dcpt_gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int N, int K, float* out_f32,
                   void* out_bf16, int ldo, const float* bias, const float* resid, int splits, int accumulate, int impl,
                   dcpt_stream_t stream) {
  dcpt_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = M; d.N = N; d.K = K; d.A = A; d.lda = lda; d.a_mn = a_mn; d.B = B; d.ldb = ldb; d.b_mn = b_mn;
  d.splits = splits; d.epilogue = (splits > 1 || accumulate) ? EPI_ATOMIC : EPI_STORE;
  d.out_f32 = out_f32; d.out_bf16 = out_bf16; d.ldo = ldo; d.bias = bias; d.resid = resid; d.ldr = ldo;
  return dcpt_gemm_ex(&d, impl, stream);
}

int dcpt_dwconv3x3_gate_fwd(const void* u_bf16, const float* weight, const float* bias, void* g_bf16, float* pool, int N, int H,
                            int W, int C, dcpt_stream_t stream) {
  return dwgate_fwd_launch(static_cast<const bf16*>(u_bf16), weight, bias, static_cast<bf16*>(g_bf16), pool, N, H, W, C,
                           static_cast<cudaStream_t>(stream));
}

// ------------------------------- NAFBlock ---------------------------------------
size_t dcpt_nafblock_packed_bytes(int C) {
  Arena a(nullptr);
  BlockPacked pk(a, C);
  (void)pk;
  return a.size();
}
size_t dcpt_nafblock_saved_bytes(int N, int H, int W, int C) {
  Arena a(nullptr);
  BlockSaved sv(a, N, H, W, C);
  (void)sv;
  return a.size();
}
size_t dcpt_nafblock_workspace_bytes(int N, int H, int W, int C) {
  Arena a(nullptr);
  BlockWork wk(a, N, H, W, C);
  (void)wk;
  return a.size();
}

int dcpt_nafblock_pack(const float* const* host_params, void* packed, int C, dcpt_stream_t stream) {
  DCPT_TRY(check_block_shape(1, 1, 1, C));
  DCPT_TRY(check_ptrs16(reinterpret_cast<const void* const*>(host_params), DCPT_NAFBLOCK_NPARAMS, "params"));
  Arena a(packed);
  BlockPacked pk(a, C);
  return nafblock_pack_impl(host_params, pk, C, static_cast<cudaStream_t>(stream));
}

int dcpt_nafblock_fwd(const float* const* host_params, const void* packed, const float* x, float* out, void* out_bf16, void* saved,
                      int N, int H, int W, int C, dcpt_stream_t stream) {
  DCPT_TRY(check_block_shape(N, H, W, C));
  DCPT_TRY(check_ptrs16(reinterpret_cast<const void* const*>(host_params), DCPT_NAFBLOCK_NPARAMS, "params"));
  Arena ap(const_cast<void*>(packed));
  BlockPacked pk(ap, C);
  Arena as(saved);
  BlockSaved sv(as, N, H, W, C);
  return nafblock_fwd_impl(host_params, pk, x, out, static_cast<bf16*>(out_bf16), sv, N, H, W, C, static_cast<cudaStream_t>(stream));
}

int dcpt_nafblock_bwd(const float* const* host_params, const void* packed, const void* saved, const float* x, const float* dout,
                      const void* dout_bf16, const float* dout_colsum, float* dx, void* dx_bf16, float* dx_colsum,
                      float* const* host_grads, void* workspace, int N, int H, int W, int C, dcpt_stream_t stream) {
  DCPT_TRY(check_block_shape(N, H, W, C));
  DCPT_TRY(check_ptrs16(reinterpret_cast<const void* const*>(host_params), DCPT_NAFBLOCK_NPARAMS, "params"));
  DCPT_TRY(check_ptrs16(reinterpret_cast<const void* const*>(host_grads), DCPT_NAFBLOCK_NPARAMS, "grads"));
  Arena ap(const_cast<void*>(packed));
  BlockPacked pk(ap, C);
  Arena as(const_cast<void*>(saved));
  BlockSaved sv(as, N, H, W, C);
  Arena aw(workspace);
  BlockWork wk(aw, N, H, W, C);
  return nafblock_bwd_impl(host_params, pk, sv, x, dout, static_cast<const bf16*>(dout_bf16), dout_colsum, dx,
                           static_cast<bf16*>(dx_bf16), dx_colsum, host_grads, wk, N, H, W, C, static_cast<cudaStream_t>(stream));
}

// ------------------------------- NAFNetBaseline ---------------------------------
dcpt_nafnet_plan* dcpt_nafnet_create(int img_channel, int width, int middle_blk_num, const int* enc_blk_nums, int n_enc,
                                     const int* dec_blk_nums, int n_dec) {
  if (img_channel <= 0 || width < 8 || width % 8 != 0 || middle_blk_num < 0 || n_enc < 0 || n_dec != n_enc) {
    dcpt_set_error("nafnet_create: need width %% 8 == 0 (>= 8) and len(enc_blk_nums) == len(dec_blk_nums) (width=%d n_enc=%d n_dec=%d)",
                   width, n_enc, n_dec);
    return nullptr;
  }
  if ((long long)width << n_enc > 1024) {
    dcpt_set_error("nafnet_create: bottleneck width %lld exceeds 1024", (long long)width << n_enc);
    return nullptr;
  }
  dcpt_nafnet_plan* p = new dcpt_nafnet_plan();
  p->img_channel = img_channel; p->width = width; p->middle_blk_num = middle_blk_num;
  p->enc.assign(enc_blk_nums, enc_blk_nums + n_enc);
  p->dec.assign(dec_blk_nums, dec_blk_nums + n_dec);
  // named_parameters() order of the reference module (nafnet_arch.py:202-248)
  add_param(p, width, img_channel, 3, 3); add_param(p, width);       // intro
  add_param(p, img_channel, width, 3, 3); add_param(p, img_channel); // ending
  int C = width;
  p->enc_blks.resize(n_enc);
  for (int i = 0; i < n_enc; ++i) {
    for (int j = 0; j < p->enc[i]; ++j) p->enc_blks[i].push_back({add_block_params(p, C), C, i});
    C *= 2;
  }
  p->mid_C = C;
  for (int j = 0; j < middle_blk_num; ++j) p->mid_blks.push_back({add_block_params(p, C), C, n_enc});
  for (int i = 0; i < n_dec; ++i) {
    p->up_pidx.push_back((int)p->params.size());
    add_param(p, 2 * C, C, 1, 1);
    C /= 2;
  }
  C = width;
  for (int i = 0; i < n_enc; ++i) {
    p->down_pidx.push_back((int)p->params.size());
    add_param(p, 2 * C, C, 2, 2); add_param(p, 2 * C);
    C *= 2;
  }
  p->dec_blks.resize(n_dec);
  for (int i = 0; i < n_dec; ++i) {
    C /= 2;
    for (int j = 0; j < p->dec[i]; ++j) p->dec_blks[i].push_back({add_block_params(p, C), C, n_enc - 1 - i});
  }
  return p;
}

void dcpt_nafnet_destroy(dcpt_nafnet_plan* plan) { delete plan; }

int dcpt_nafnet_set_tlc(dcpt_nafnet_plan* plan, const int* kh, const int* kw, int n_levels) {
  DCPT_CHECK_ARG(plan != nullptr && n_levels >= 0 && (n_levels == 0 || (kh && kw)), DCPT_E_ARG, "nafnet_set_tlc: bad arguments");
  plan->tlc_kh.assign(kh, kh + n_levels);
  plan->tlc_kw.assign(kw, kw + n_levels);
  for (int i = 0; i < n_levels; ++i)
    DCPT_CHECK_ARG(kh[i] >= 1 && kw[i] >= 1, DCPT_E_ARG, "nafnet_set_tlc: kernel of level %d must be positive", i);
  return 0;
}
int dcpt_nafnet_set_hook_blocks(dcpt_nafnet_plan* plan, const int* block_idx, int n_levels) {
  DCPT_CHECK_ARG(plan != nullptr && n_levels >= 0 && n_levels <= (int)plan->dec.size() && (n_levels == 0 || block_idx), DCPT_E_ARG,
                 "nafnet_set_hook_blocks: bad arguments");
  for (int i = 0; i < n_levels; ++i)
    DCPT_CHECK_ARG(block_idx[i] < plan->dec[i] || plan->dec[i] == 0, DCPT_E_ARG, "nafnet_set_hook_blocks: decoder%d has %d blocks, got index %d", i,
                   plan->dec[i], block_idx[i]);
  plan->hook_blk.assign(block_idx, block_idx + n_levels);
  return 0;
}
int dcpt_nafnet_set_keep_activations(const dcpt_nafnet_plan* plan, int keep) {
  DCPT_CHECK_ARG(plan != nullptr, DCPT_E_ARG, "nafnet_set_keep_activations: null plan");
  plan->keep_activations = keep != 0;
  return 0;
}

int dcpt_nafnet_set_bwd_split_event(dcpt_nafnet_plan* plan, void* cuda_event, int external) {
  DCPT_CHECK_ARG(plan != nullptr, DCPT_E_ARG, "nafnet_set_bwd_split_event: null plan");
  plan->split_event = cuda_event;
  plan->split_event_external = external != 0;
  return 0;
}
int dcpt_nafnet_num_params(const dcpt_nafnet_plan* plan) { return (int)plan->params.size(); }
long long dcpt_nafnet_param_shape(const dcpt_nafnet_plan* plan, int i, int dims[4]) {
  if (i < 0 || i >= (int)plan->params.size()) return -1;
  for (int k = 0; k < 4; ++k) dims[k] = plan->params[i].dims[k];
  return plan->params[i].numel;
}
size_t dcpt_nafnet_packed_bytes(const dcpt_nafnet_plan* plan) {
  Arena a(nullptr);
  NetPacked pk(plan, a);
  return a.size();
}
size_t dcpt_nafnet_saved_bytes(const dcpt_nafnet_plan* plan, int N, int H, int W) {
  Arena a(nullptr);
  NetSaved sv(plan, a, N, H, W);
  return a.size();
}
size_t dcpt_nafnet_workspace_bytes(const dcpt_nafnet_plan* plan, int N, int H, int W) {
  Arena a(nullptr);
  NetWork wk(plan, a, N, H, W);
  return a.size();
}

int dcpt_nafnet_pack(const dcpt_nafnet_plan* p, const float* const* P, void* packed, dcpt_stream_t stream) {
  DCPT_TRY(check_ptrs16(reinterpret_cast<const void* const*>(P), (int)p->params.size(), "params"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a(packed);
  NetPacked pk(p, a);
  const int ne = (int)p->enc.size(), nd = (int)p->dec.size();
  int C = p->width;
  DCPT_TRY(pack_w27_launch(P[0], pk.wi_p, C, 0, st));
  DCPT_TRY(pack_w27_launch(P[2], pk.wd_p, C, 1, st));
  for (int i = 0; i < ne; ++i) {
    for (int j = 0; j < p->enc[i]; ++j) DCPT_TRY(nafblock_pack_impl(P + p->enc_blks[i][j].pidx, pk.enc_pk[i][j], C, st));
    DCPT_TRY(pack_weight_launch(P[p->down_pidx[i]], nullptr, pk.down[i], 2 * C, 4 * C, PACK_DOWN, st));
    DCPT_TRY(pack_weight_launch(P[p->down_pidx[i]], nullptr, pk.down_t[i], 2 * C, 4 * C, PACK_DOWN_T, st));
    C *= 2;
  }
  for (int j = 0; j < p->middle_blk_num; ++j) DCPT_TRY(nafblock_pack_impl(P + p->mid_blks[j].pidx, pk.mid_pk[j], C, st));
  for (int i = 0; i < nd; ++i) {
    DCPT_TRY(pack_weight_launch(P[p->up_pidx[i]], nullptr, pk.up[i], 2 * C, C, PACK_UP, st));
    DCPT_TRY(pack_weight_launch(P[p->up_pidx[i]], nullptr, pk.up_t[i], 2 * C, C, PACK_UP_T, st));
    C /= 2;
    for (int j = 0; j < p->dec[i]; ++j) DCPT_TRY(nafblock_pack_impl(P + p->dec_blks[i][j].pidx, pk.dec_pk[i][j], C, st));
  }
  return 0;
}

int dcpt_nafnet_fwd(const dcpt_nafnet_plan* p, const float* const* P, const void* packed, const float* inp, float* out, void* saved,
                    float* const* host_feats, int hook, int N, int H, int W, dcpt_stream_t stream) {
  DCPT_TRY(check_net_shape(p, N, H, W));
  DCPT_CHECK_ARG(hook || out != nullptr, DCPT_E_ARG, "nafnet_fwd: out is NULL but hook == 0");
  DCPT_TRY(check_ptrs16(reinterpret_cast<const void* const*>(P), (int)p->params.size(), "params"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena ap(const_cast<void*>(packed));
  NetPacked pk(p, ap);
  Arena as(saved);
  NetSaved sv(p, as, N, H, W);
  const int ne = (int)p->enc.size(), nd = (int)p->dec.size();
  int C = p->width, h = H, w = W;
  auto tkh = [&](int lvl) { return lvl < (int)p->tlc_kh.size() ? p->tlc_kh[lvl] : 0; };
  auto tkw = [&](int lvl) { return lvl < (int)p->tlc_kw.size() ? p->tlc_kw[lvl] : 0; };

  // The norm1 of a block is computed by the epilogue of the GEMM that writes the block's input (intro conv, down conv, the
  // previous block's conv5) whenever that GEMM holds whole rows (C <= 512): `fused` says the next block finds n1 / stats1 ready.
  auto ln1_of = [&](const dcpt_nafnet_plan::Blk& b, const BlockSaved& bsv) -> LnFuse {
    if (!ln_fusable(b.C)) return LnFuse{nullptr, nullptr, nullptr, nullptr};
    return LnFuse{P[b.pidx + P_N1W], P[b.pidx + P_N1B], bsv.n1, bsv.stats1};
  };
  const LnFuse no_ln = {nullptr, nullptr, nullptr, nullptr};
  // first block after the encoder level i's down conv (or after the intro conv for i = -1)
  auto first_after = [&](int i) -> LnFuse {
    for (int k = i + 1; k < ne; ++k) {
      if (p->enc[k] > 0) return k == i + 1 ? ln1_of(p->enc_blks[k][0], sv.enc_sv[k][0]) : no_ln;
      return no_ln;  // an empty level: its down conv consumes the tensor
    }
    return (i + 1 == ne && p->middle_blk_num > 0) ? ln1_of(p->mid_blks[0], sv.mid_sv[0]) : no_ln;
  };
  bool fused = false;
  DCPT_CUDA(cudaMemsetAsync(sv.pools, 0, sv.pool_floats * sizeof(float), st));  // every block's SCA pooling sums
  // intro (nafnet_arch.py:252)
  {  // x0 = im2col(inp) * Wi^T + b on the tensor cores
    DCPT_TRY(im2col3_launch(inp, sv.P, nullptr, 0, 1, N, h, w, st));
    GemmArgs g = gemm_args(N * h * w, C, 64, sv.P, 64, pk.wi_p, 64, EPI_STORE);  // [P | P] x [W_hi | W_lo]
    g.ep.out_f32 = sv.x0; g.ep.ldo = C; g.ep.bias = P[1];
    const LnFuse f = first_after(-1);
    fused = f.n != nullptr;
    if (fused) set_ln_epilogue(g, f, C);
    DCPT_TRY(gemm_launch(g, st));
  }
  const float* x = sv.x0;
  // encoders + downs (:256-259)
  for (int i = 0; i < ne; ++i) {
    for (int j = 0; j < p->enc[i]; ++j) {
      const LnFuse nx = j + 1 < p->enc[i] ? ln1_of(p->enc_blks[i][j + 1], sv.enc_sv[i][j + 1]) : no_ln;
      DCPT_TRY(nafblock_fwd_impl(P + p->enc_blks[i][j].pidx, pk.enc_pk[i][j], x, sv.enc_out[i][j], nullptr, sv.enc_sv[i][j], N, h,
                                 w, C, st, tkh(i), tkw(i), fused, &nx, true, p->keep_activations));
      fused = nx.n != nullptr;
      x = sv.enc_out[i][j];
    }
    DCPT_TRY(unshuffle_cast_launch(x, sv.xu[i], N, h / 2, w / 2, C, st));
    h /= 2; w /= 2;
    GemmArgs g = gemm_args(N * h * w, 2 * C, 4 * C, sv.xu[i], 4 * C, pk.down[i], 4 * C, EPI_STORE);
    g.ep.out_f32 = sv.xd[i]; g.ep.ldo = 2 * C; g.ep.bias = P[p->down_pidx[i] + 1];
    if (i == ne - 1 && p->middle_blk_num == 0 && nd > 0) g.ep.out_bf16 = sv.up_in[0];
    const LnFuse f = first_after(i);
    fused = f.n != nullptr;
    if (fused) set_ln_epilogue(g, f, 2 * C);
    DCPT_TRY(gemm_launch(g, st));
    C *= 2;
    x = sv.xd[i];
  }
  // middle (:261)
  for (int j = 0; j < p->middle_blk_num; ++j) {
    bf16* mirror = (j == p->middle_blk_num - 1 && nd > 0) ? sv.up_in[0] : nullptr;
    const LnFuse nx = j + 1 < p->middle_blk_num ? ln1_of(p->mid_blks[j + 1], sv.mid_sv[j + 1]) : no_ln;
    DCPT_TRY(nafblock_fwd_impl(P + p->mid_blks[j].pidx, pk.mid_pk[j], x, sv.mid_out[j], mirror, sv.mid_sv[j], N, h, w, C, st, tkh(ne),
                               tkw(ne), fused, &nx, true, p->keep_activations));
    fused = nx.n != nullptr;
    x = sv.mid_out[j];
  }
  if (ne == 0 && p->middle_blk_num == 0 && nd > 0) { dcpt_set_error("nafnet_fwd: degenerate network"); return DCPT_E_UNSUPPORTED; }
  // ups + skip + decoders (:263-267)
  for (int i = 0; i < nd; ++i) {
    const float* skip = p->enc[ne - 1 - i] > 0 ? sv.enc_out[ne - 1 - i].back() : (ne - 1 - i > 0 ? sv.xd[ne - 2 - i] : sv.x0);
    GemmArgs g = gemm_args(N * h * w, 2 * C, C, sv.up_in[i], C, pk.up[i], C, EPI_PIXSHUF);
    g.ep.out_f32 = sv.xup[i]; g.ep.resid = skip; g.ep.H = h; g.ep.W = w; g.ep.Cseg = C / 2;
    if (p->dec[i] == 0 && i + 1 < nd) g.ep.out_bf16 = sv.up_in[i + 1];
    DCPT_TRY(gemm_launch(g, st));
    h *= 2; w *= 2; C /= 2;
    x = sv.xup[i];
    fused = false;  // the up conv's pixel-shuffle epilogue scatters rows: the first decoder block runs its own norm1
    for (int j = 0; j < p->dec[i]; ++j) {
      bf16* mirror = (j == p->dec[i] - 1) ? (i + 1 < nd ? sv.up_in[i + 1] : sv.xlast_bf16) : nullptr;
      const LnFuse nx = j + 1 < p->dec[i] ? ln1_of(p->dec_blks[i][j + 1], sv.dec_sv[i][j + 1]) : no_ln;
      DCPT_TRY(nafblock_fwd_impl(P + p->dec_blks[i][j].pidx, pk.dec_pk[i][j], x, sv.dec_out[i][j], mirror, sv.dec_sv[i][j], N, h, w,
                                 C, st, tkh(ne - 1 - i), tkw(ne - 1 - i), fused, &nx, true, p->keep_activations));
      fused = nx.n != nullptr;
      x = sv.dec_out[i][j];
      if (host_feats && host_feats[i] && j == p->hook_block(i))
        DCPT_CUDA(cudaMemcpyAsync(host_feats[i], x, (size_t)N * h * w * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (host_feats && host_feats[i] && p->dec[i] == 0)
      DCPT_CUDA(cudaMemcpyAsync(host_feats[i], x, (size_t)N * h * w * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  // ending + global residual (:271-272)
  if (!hook) {
    const bool mirrored = nd > 0 && p->dec[nd - 1] > 0 && C <= 64;  // the last block wrote sv.xlast_bf16
    if (mirrored) DCPT_TRY(ending_fwd_tma_launch(sv.xlast_bf16, P[2], P[3], inp, out, N, h, w, C, st));
    else DCPT_TRY(conv3x3_feat_to_img_launch(x, P[2], P[3], inp, out, N, h, w, C, st));
  }
  return 0;
}

int dcpt_nafnet_bwd(const dcpt_nafnet_plan* p, const float* const* P, const void* packed, const void* saved, const float* inp,
                    const float* dout, const float* const* host_dfeats, float* const* G, void* workspace, int N, int H, int W,
                    dcpt_stream_t stream) {
  DCPT_TRY(check_net_shape(p, N, H, W));
  DCPT_CHECK_ARG(p->tlc_kh.empty(), DCPT_E_UNSUPPORTED, "nafnet_bwd: the TLC variant is a test-time converter (inference only)");
  DCPT_TRY(check_ptrs16(reinterpret_cast<const void* const*>(P), (int)p->params.size(), "params"));
  DCPT_TRY(check_ptrs16(reinterpret_cast<const void* const*>(G), (int)p->params.size(), "grads"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena ap(const_cast<void*>(packed));
  NetPacked pk(p, ap);
  Arena as(const_cast<void*>(saved));
  NetSaved sv(p, as, N, H, W);
  Arena aw(workspace);
  NetWork wk(p, aw, N, H, W);
  const int ne = (int)p->enc.size(), nd = (int)p->dec.size();

  // Gradient stream: d(loss)/d(current residual tensor) as fp32, its bf16 mirror (GEMM operand) and its
  // column sums (bias / beta / gamma gradients).  Two slots, ping-pong.
  struct Slot { float* f; bf16* t; float* s; };
  Slot slot[2] = {{wk.dxa, wk.dta, wk.Sa}, {wk.dxb, wk.dtb, wk.Sb}};
  float* const Sorig[2] = {wk.Sa, wk.Sb};
  int ci = 0;
  bool have = false;
#define CUR slot[ci]
#define NXT slot[ci ^ 1]

  int C = p->width, h = H, w = W;  // the last decoder level runs at full resolution (n_enc == n_dec)
  const float* x_last = nd > 0 ? (p->dec[nd - 1] > 0 ? sv.dec_out[nd - 1].back() : sv.xup[nd - 1])
                               : (p->middle_blk_num > 0 ? sv.mid_out.back() : sv.x0);
  if (dout) {
    // ending conv: wgrad, bias grad, dgrad (nafnet_arch.py:271)
    const bool mirrored = nd > 0 && p->dec[nd - 1] > 0 && C <= 64;
    if (mirrored) {
      // Pd = flipped patches of dout;  dWe = F^T Pd (wgrad GEMM),  dF = Pd Wd^T (GEMM),  colsum(dF) = Wd * colsum(Pd)
      const int M0 = N * h * w;
      DCPT_CUDA(cudaMemsetAsync(wk.psum, 0, 32 * sizeof(float), st));
      DCPT_TRY(im2col3_launch(dout, wk.Pd, wk.psum, 1, 0, N, h, w, st));
      DCPT_CUDA(cudaMemsetAsync(wk.G, 0, (size_t)32 * C * sizeof(float), st));
      DCPT_TRY(wgrad_gemm(sv.xlast_bf16, C, wk.Pd, 32, wk.G, M0, st));
      DCPT_TRY(finish_w27_launch(wk.G, G[2], wk.psum, G[3], C, 1, st));
      GemmArgs g = gemm_args(M0, C, 32, wk.Pd, 32, pk.wd_p, 32, EPI_STORE);
      g.ep.out_f32 = CUR.f; g.ep.out_bf16 = CUR.t; g.ep.ldo = C;
      DCPT_TRY(gemm_launch(g, st));
      DCPT_CUDA(cudaMemsetAsync(CUR.s, 0, C * sizeof(float), st));
      DCPT_TRY(ending_colsum_launch(P[2], wk.psum, CUR.s, C, st));
    } else {
      DCPT_CUDA(cudaMemsetAsync(wk.G, 0, (size_t)27 * C * sizeof(float), st));
      DCPT_TRY(conv3x3_small_wgrad_launch(x_last, dout, wk.G, G[3], 1, N, h, w, C, st));
      DCPT_TRY(wgrad_finish_perm_launch(wk.G, G[2], C, 27, FIN_CN_TO_C3, st));
      DCPT_CUDA(cudaMemsetAsync(CUR.s, 0, C * sizeof(float), st));
      DCPT_TRY(conv3x3_img_to_feat_launch(dout, P[2], nullptr, 1, CUR.f, CUR.t, CUR.s, N, h, w, C, st));
    }
    have = true;
  }
  p->side.pending[0] = p->side.pending[1] = false;
  // one memset for every block's accumulators (BlockZeros) and column-sum outputs
  DCPT_CUDA(cudaMemsetAsync(wk.zeros, 0, wk.zero_floats * sizeof(float), st));
  float* zcur = wk.zeros;
  int blk_i = 0;
  auto block_bwd = [&](const dcpt_nafnet_plan::Blk& b, const BlockPacked& bpk, const BlockSaved& bsv, const float* x, int hh,
                       int ww) -> int {
    DCPT_CHECK_ARG(have, DCPT_E_ARG, "nafnet_bwd: no gradient reaches a block (dout and dfeats all NULL?)");
    const int parity = side_lag_enabled() ? (blk_i++ & 1) : -1;
    Arena a(parity == 1 ? wk.blk_base2 : wk.blk_base);
    BlockWork bw(a, N, hh, ww, b.C);
    const BlockZeros bz(zcur, N, b.C);
    zcur += BlockZeros::floats(N, b.C);
    NXT.s = zcur;  // this block's Sx (column sums of dx): its own pre-cleared slice, read by the next block / conv backward
    zcur += ((size_t)b.C + 63) / 64 * 64;
    DCPT_TRY(nafblock_bwd_impl(P + b.pidx, bpk, bsv, x, CUR.f, CUR.t, CUR.s, NXT.f, NXT.t, NXT.s, G + b.pidx, bw, N, hh, ww, b.C, st, &bz, &p->side,
                               parity));
    ci ^= 1;
    return 0;
  };

  // ---------------- decoders (reverse) ----------------
  for (int i = nd - 1; i >= 0; --i) {
    const float* ext = host_dfeats ? host_dfeats[i] : nullptr;
    auto inject = [&]() -> int {  // gradient injected by the DCPT head where the hooked feature was taken
      NXT.s = Sorig[ci ^ 1];  // (a block backward may have pointed the slot at its own pre-cleared slice)
      DCPT_CUDA(cudaMemsetAsync(NXT.s, 0, C * sizeof(float), st));
      DCPT_TRY(grad_prepare_launch(have ? CUR.f : nullptr, ext, NXT.f, NXT.t, NXT.s, N * h * w, C, st));
      ci ^= 1;
      have = true;
      return 0;
    };
    if (ext && p->dec[i] == 0) DCPT_TRY(inject());
    for (int j = p->dec[i] - 1; j >= 0; --j) {
      if (ext && j == p->hook_block(i)) DCPT_TRY(inject());  // output of block j of this level
      if (!have) continue;  // hook pass: the blocks behind the hooked one of the last level receive no gradient (stay zero)
      DCPT_TRY(block_bwd(p->dec_blks[i][j], pk.dec_pk[i][j], sv.dec_sv[i][j], j > 0 ? sv.dec_out[i][j - 1] : sv.xup[i], h, w));
    }
    DCPT_CHECK_ARG(have, DCPT_E_ARG, "nafnet_bwd: no gradient at decoder level %d", i);
    // CUR = d(xup_i): it is also the gradient of the encoder skip (x = up(x) + enc_skip, :264-265)
    const int lvl = ne - 1 - i;
    DCPT_CUDA(cudaMemcpyAsync(wk.dskip[lvl], CUR.f, (size_t)N * h * w * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // up conv backward (1x1, no bias, + PixelShuffle(2))
    const int Cin = 2 * C, hin = h / 2, win = w / 2, Min = N * hin * win;
    DCPT_TRY(unshuffle_cast_launch(CUR.f, wk.dconv, N, hin, win, C, st));  // [Min, 2*Cin] in packed column order
    DCPT_CUDA(cudaMemsetAsync(wk.G, 0, (size_t)2 * Cin * Cin * sizeof(float), st));
    DCPT_TRY(wgrad_gemm(wk.dconv, 2 * Cin, sv.up_in[i], Cin, wk.G, Min, st));
    DCPT_TRY(wgrad_finish_perm_launch(wk.G, G[p->up_pidx[i]], 2 * Cin, Cin, FIN_UP, st));
    {
      GemmArgs g = gemm_args(Min, Cin, 2 * Cin, wk.dconv, 2 * Cin, pk.up_t[i], 2 * Cin, EPI_STORE);
      g.ep.out_f32 = NXT.f; g.ep.out_bf16 = NXT.t; g.ep.ldo = Cin;
      DCPT_TRY(gemm_launch(g, st));
      NXT.s = Sorig[ci ^ 1];
      DCPT_CUDA(cudaMemsetAsync(NXT.s, 0, Cin * sizeof(float), st));
      DCPT_TRY(grad_prepare_launch(NXT.f, nullptr, nullptr, nullptr, NXT.s, Min, Cin, st));
      ci ^= 1;
    }
    h = hin; w = win; C = Cin;
  }
  // ---------------- middle (reverse) ----------------
  for (int j = p->middle_blk_num - 1; j >= 0; --j)
    DCPT_TRY(block_bwd(p->mid_blks[j], pk.mid_pk[j], sv.mid_sv[j], j > 0 ? sv.mid_out[j - 1] : (ne > 0 ? sv.xd[ne - 1] : sv.x0), h, w));
  // ---------------- encoders (reverse) ----------------
  for (int i = ne - 1; i >= 0; --i) {
    DCPT_CHECK_ARG(have, DCPT_E_ARG, "nafnet_bwd: no gradient at encoder level %d", i);
    // CUR = d(xd_i) [N, h, w, C]; down conv (2x2 stride 2, C/2 -> C, bias) backward (:230, :259)
    const int Cl = C / 2, hl = h * 2, wl = w * 2, Mh = N * h * w;
    DCPT_TRY(axpy_launch(G[p->down_pidx[i] + 1], CUR.s, C, st));
    DCPT_CUDA(cudaMemsetAsync(wk.G, 0, (size_t)C * 4 * Cl * sizeof(float), st));
    DCPT_TRY(wgrad_gemm(CUR.t, C, sv.xu[i], 4 * Cl, wk.G, Mh, st));
    DCPT_TRY(wgrad_finish_perm_launch(wk.G, G[p->down_pidx[i]], C, 4 * Cl, FIN_DOWN, st));
    {
      GemmArgs g = gemm_args(Mh, 4 * Cl, C, CUR.t, C, pk.down_t[i], C, EPI_PIXSHUF);
      g.ep.out_f32 = NXT.f; g.ep.out_bf16 = NXT.t; g.ep.resid = wk.dskip[i];  // + gradient of the decoder skip
      g.ep.H = h; g.ep.W = w; g.ep.Cseg = Cl;
      DCPT_TRY(gemm_launch(g, st));
      NXT.s = Sorig[ci ^ 1];
      DCPT_CUDA(cudaMemsetAsync(NXT.s, 0, Cl * sizeof(float), st));
      DCPT_TRY(grad_prepare_launch(NXT.f, nullptr, nullptr, nullptr, NXT.s, N * hl * wl, Cl, st));
      ci ^= 1;
    }
    h = hl; w = wl; C = Cl;
    for (int j = p->enc[i] - 1; j >= 0; --j)
      DCPT_TRY(block_bwd(p->enc_blks[i][j], pk.enc_pk[i][j], sv.enc_sv[i][j],
                         j > 0 ? sv.enc_out[i][j - 1] : (i > 0 ? sv.xd[i - 1] : sv.x0), h, w));
    if (i == ne - 1 && p->split_event) {  // deepest level done: its slice of the gradient buffer is final (see split_event)
      DCPT_TRY(side_join_all(p->side, st));  // ... once the lagging weight-gradient work of its last blocks has landed
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      DCPT_CUDA(cudaStreamIsCapturing(st, &cap));
      DCPT_CUDA(cudaEventRecordWithFlags(static_cast<cudaEvent_t>(p->split_event), st,
                                         (cap == cudaStreamCaptureStatusActive && p->split_event_external) ? cudaEventRecordExternal
                                                                                                           : cudaEventRecordDefault));
    }
  }
  // ---------------- intro: wgrad + bias grad (the input image needs no gradient) ----------------
  DCPT_CHECK_ARG(have, DCPT_E_ARG, "nafnet_bwd: no gradient reached the intro conv");
  DCPT_CUDA(cudaMemsetAsync(wk.G, 0, (size_t)32 * C * sizeof(float), st));
  DCPT_TRY(wgrad_gemm(CUR.t, C, sv.P, 32, wk.G, N * h * w, st, 64));  // dWi = dX0^T * im2col(inp)
  DCPT_TRY(finish_w27_launch(wk.G, G[0], nullptr, nullptr, C, 0, st));
  DCPT_TRY(axpy_launch(G[1], CUR.s, C, st));
  DCPT_TRY(side_join_all(p->side, st));  // everything is ordered on `stream` when the call returns
#undef CUR
#undef NXT
  return 0;
}

// ------------------------------- DC head building blocks -------------------------
#define ST(s) static_cast<cudaStream_t>(s)
#define BF(p) static_cast<bf16*>(p)
#define CBF(p) static_cast<const bf16*>(p)

int dcpt_pack_matrix(const float* w, void* out_bf16, int O, int I, int transpose, dcpt_stream_t stream) {
  return pack_weight_launch(w, nullptr, BF(out_bf16), O, I, transpose ? PACK_T : PACK_PLAIN, ST(stream));
}
size_t dcpt_conv3x3_packed_elems(int Cout, int Cin, int dgrad) {
  const int R = dgrad ? Cin : Cout, Cc = dgrad ? Cout : Cin;
  return (size_t)R * 9 * (ceil_div(Cc, 64) * 64);
}
int dcpt_conv3x3_pack(const float* w, void* out_bf16, int Cout, int Cin, int dgrad, dcpt_stream_t stream) {
  return pack_conv3x3_launch(w, BF(out_bf16), Cout, Cin, dgrad, ST(stream));
}
int dcpt_conv3x3_fwd(const void* x_bf16, const void* w_packed, void* out_bf16, float* out_f32, int N, int H, int W, int Cin, int Cout,
                     dcpt_stream_t stream) {
  Conv3x3Args a;
  memset(&a, 0, sizeof(a));
  a.X = CBF(x_bf16); a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Wp = CBF(w_packed); a.Cout = Cout;
  a.ep.out_bf16 = BF(out_bf16); a.ep.out_f32 = out_f32; a.ep.ldo = Cout;
  return conv3x3_tc_launch(a, ST(stream));
}
int dcpt_conv3x3_wgrad(const void* dy_bf16, const void* x_bf16, float* scratch, float* dw, int N, int H, int W, int Cin, int Cout,
                       dcpt_stream_t stream) {
  const size_t n = (size_t)Cout * 9 * (ceil_div(Cin, 64) * 64);
  DCPT_CUDA(cudaMemsetAsync(scratch, 0, n * sizeof(float), ST(stream)));
  DCPT_TRY(conv3x3_wgrad_tc_launch(CBF(dy_bf16), CBF(x_bf16), scratch, N, H, W, Cin, Cout, ST(stream)));
  return finish_conv3x3_launch(scratch, dw, Cout, Cin, ST(stream));
}
int dcpt_ln_act_fwd(const void* x_bf16, const float* weight, const float* bias, const void* resid_bf16, void* y_bf16, float* stats,
                    int M, int C, int relu, float eps, dcpt_stream_t stream) {
  return ln_act_fwd_launch(CBF(x_bf16), weight, bias, CBF(resid_bf16), BF(y_bf16), stats, M, C, relu, eps, ST(stream));
}
int dcpt_ln_act_bwd(const float* dy, const void* y_bf16, const void* x_bf16, const float* stats, const float* weight,
                    void* dx_bf16, float* dres, float* dweight, float* dbias, int M, int C, int relu, dcpt_stream_t stream) {
  return ln_act_bwd_launch(dy, CBF(y_bf16), CBF(x_bf16), stats, weight, BF(dx_bf16), dres, dweight, dbias, M, C, relu, ST(stream));
}
int dcpt_mix_fwd(const void* prev_bf16, const float* feat, const float* mw, void* z_bf16, long long n, dcpt_stream_t stream) {
  return mix_fwd_launch(CBF(prev_bf16), feat, mw, BF(z_bf16), n, ST(stream));
}
int dcpt_mix_bwd(const float* dz, const float* feat, const float* mw, float* dfeat, float* dmw, long long n, dcpt_stream_t stream) {
  return mix_bwd_launch(dz, feat, mw, dfeat, dmw, n, ST(stream));
}
int dcpt_maxpool2_relu_fwd(const void* x_bf16, void* y_bf16, int N, int Ho, int Wo, int C, dcpt_stream_t stream) {
  return maxpool2_relu_fwd_launch(CBF(x_bf16), BF(y_bf16), N, Ho, Wo, C, ST(stream));
}
int dcpt_maxpool2_relu_bwd(const void* x_bf16, const float* dy, void* dx_bf16, int N, int Ho, int Wo, int C, dcpt_stream_t stream) {
  return maxpool2_relu_bwd_launch(CBF(x_bf16), dy, BF(dx_bf16), N, Ho, Wo, C, ST(stream));
}
int dcpt_meanpool_fc_fwd(const void* x_bf16, const float* weight, const float* bias, float* pooled, float* logits, int N, int HW, int C,
                         int K, dcpt_stream_t stream) {
  return meanpool_fc_fwd_launch(CBF(x_bf16), weight, bias, pooled, logits, N, HW, C, K, ST(stream));
}
int dcpt_meanpool_fc_bwd(const float* dlogits, const float* pooled, const float* weight, float* dweight, float* dbias, float* dx,
                         int N, int HW, int C, int K, dcpt_stream_t stream) {
  return meanpool_fc_bwd_launch(dlogits, pooled, weight, dweight, dbias, dx, N, HW, C, K, ST(stream));
}
int dcpt_im2col7x7s2(const float* img, void* patches_bf16, int N, int H, int W, dcpt_stream_t stream) {
  return im2col7s2_launch(img, BF(patches_bf16), N, H, W, ST(stream));
}
int dcpt_add_bf16(const void* a, const void* b, void* out, long long n, dcpt_stream_t stream) {
  return add_bf16_launch(CBF(a), CBF(b), BF(out), n, ST(stream));
}
#undef ST
#undef BF
#undef CBF

}  // extern "C"
