/* dcpt_ops.h — C ABI of the B200 (sm_100a) hot-path library `libdcpt_sm100.so`.
 *
 * This is the drop-in boundary for the image-restoration forward/backward hot path of
 * MILab-PKU/dcpt (basicsr/archs).  The reference has no native code on this path — every op is an
 * eager PyTorch call — so each entry point cites the *Python* interface it replaces; the packaging
 * mirrors the reference's own native-op pattern (basicsr/ops/layernorm/layernorm.py:7-67:
 * autograd.Function over a compiled extension).  INTEGRATION.md shows the ctypes binding a
 * maintainer adds under basicsr/ops/.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless its name says `host_`.
 *   - activations are NHWC ("channels_last"): row-major [M = N*H*W pixels, C channels], C % 8 == 0,
 *     16-byte aligned.  fp32 = residual stream / parameters / parameter gradients / statistics,
 *     bf16 = branch-internal activations (tensor-core operands).
 *   - the caller owns every buffer (outputs, saved-for-backward arenas, workspaces); the library
 *     never allocates device memory, never synchronises, and launches everything on `stream`
 *     (CUDA-graph capturable).
 *   - return value: 0 on success, negative DCPT_E_* for argument errors, positive = cudaError_t.
 *     dcpt_last_error() returns a thread-local message.  No C++ exception crosses this boundary.
 *   - parameter gradients are ACCUMULATED (+=) into the caller's fp32 buffers, like autograd.
 */
#ifndef DCPT_OPS_H_
#define DCPT_OPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCPT_ABI_VERSION 1

#define DCPT_E_ARG (-1)
#define DCPT_E_SHAPE (-2)
#define DCPT_E_ALIGN (-3)
#define DCPT_E_DRIVER (-4)
#define DCPT_E_UNSUPPORTED (-5)

typedef void* dcpt_stream_t; /* cudaStream_t */

int dcpt_abi_version(void);
/* Storage type of the 16-bit tensor-core operands (every `*_bf16` argument and the packed weights) this library was built for:
 * 0 = bfloat16 (libdcpt_sm100.so, the fast path), 1 = IEEE half (libdcpt_sm100_fp16.so, built with -DDCPT_OPERAND_FP16: the
 * parity mode - 8x smaller operand rounding; callers pass fp16 tensors wherever the header says bf16). */
int dcpt_operand_dtype(void);
const char* dcpt_last_error(void);

/* Instrumentation for bench.py (no reference counterpart; the reference's SRModel.nondist_profile,
 * basicsr/models/sr_model.py:520-568, times whole forwards with CUDA events).
 * dcpt_launch_count: kernels launched by this library so far (process-wide).
 * dcpt_prof_enable(1): record CUDA events around every launch; dcpt_prof_dump aggregates them by kernel tag
 * into "tag\tlaunches\ttotal_ms\tflops\tbytes" lines (synchronises the device), returns bytes written. */
long long dcpt_launch_count(void);
int dcpt_prof_enable(int on);
long long dcpt_prof_dump(char* host_buf, long long cap);

/* ------------------------------------------------------------------------------------------
 * Standalone ops (used by the tests and by callers that keep their own block structure).
 * ------------------------------------------------------------------------------------------ */

/* LayerNorm2d.forward — basicsr/archs/nafnet_arch.py:27-35 (LayerNormFunction.forward), :56-64.
 * x fp32 [M,C] -> out bf16 [M,C]; stats fp32 [M,2] = (mean, 1/sqrt(var+eps)) saved for backward. */
int dcpt_layernorm2d_fwd(const float* x, const float* weight, const float* bias, void* out_bf16, float* stats, int M,
                         int C, float eps, dcpt_stream_t stream);

/* LayerNormFunction.backward — nafnet_arch.py:38-53.  dx = dres + LN'(dn) (dres nullable);
 * dx_bf16 (nullable) is a bf16 mirror of dx; dweight/dbias/colsum (nullable) are accumulated:
 * dweight += sum_m dn*yhat, dbias += sum_m dn, colsum += sum_m dx. */
int dcpt_layernorm2d_bwd(const void* dn_bf16, const float* x, const float* stats, const float* weight, const float* dres,
                         float* dx, void* dx_bf16, float* dweight, float* dbias, float* colsum, int M, int C,
                         dcpt_stream_t stream);

/* Pointwise-convolution GEMM: D[M,N] = A[M,K] * B[N,K]^T, bf16 operands, fp32 accumulate
 * (nn.Conv2d(k=1) at nafnet_arch.py:87-95,105-113,133-150 seen as [pixels,Cin]x[Cout,Cin]^T).
 *   a_mn/b_mn = 1: operand stored [K, M] / [K, N] (wgrad, where K = pixels).
 *   out_f32 / out_bf16 nullable; bias[N] / resid fp32 [M,ldo] nullable; splits > 1 or accumulate = 1
 *   selects the split-K path that atomically ADDS into out_f32 (splits = 0: chosen to fill the SMs).
 *   impl: 0 = tcgen05/TMA (product), 1 = CUDA-core cross-check (tests only). */
int dcpt_gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int N, int K, float* out_f32,
                   void* out_bf16, int ldo, const float* bias, const float* resid, int splits, int accumulate, int impl,
                   dcpt_stream_t stream);

/* Full-control variant used by the tests to exercise every fused epilogue of the GEMM engine.
 * epilogue: 0 STORE (bias/resid -> fp32 and/or bf16), 1 GATE (conv4 -> x4 + SimpleGate, nafnet_arch.py:180-181),
 * 2 GATE_BWD (SimpleGate backward), 3 PIXSHUF (PixelShuffle(2) scatter + skip add, nafnet_arch.py:238-242,264-265),
 * 4 ATOMIC (split-K wgrad accumulate), 7 GATE on 32-wide pair packing (packed column 64p + 32h + i <-> channel h*C + 32p + i,
 * C % 32 == 0; what the NAFBlock path uses: x4 / SimpleGate tiles leave through TMA bulk stores).  STORE and GATE_BWD
 * pick their TMA-tiled epilogue automatically when rows are 16-byte pitched (and C % 32 == 0). */
typedef struct dcpt_gemm_desc {
  int M, N, K;
  const void* A; int lda; int a_mn;
  const void* B; int ldb; int b_mn;
  int splits;
  int epilogue;
  float* out_f32; void* out_bf16; int ldo;
  const float* bias;
  const float* resid; int ldr;
  void* out2_bf16; int ldo2;
  const void* aux_bf16; int ldaux;
  int C;
  int H, W, Cseg;
  /* STORE only, optional (ln_out != NULL; needs out_f32, 16-byte pitched rows and N <= 512): the channel LayerNorm of the
   * finished output rows v = acc + bias + resid (LayerNorm2d forward, nafnet_arch.py:27-35; the norm that CONSUMES this tensor)
   * computed in the same epilogue: ln_out[m,:] = bf16((v - mean) * rstd * ln_weight + ln_bias), ln_stats[m] = (mean, rstd). */
  const float* ln_weight; const float* ln_bias;
  void* ln_out; int ld_ln;
  float* ln_stats; float ln_eps;
  /* STORE only, optional (lnb_x != NULL; N <= 512, no bias / resid): the GEMM result is dn = d(loss)/d(LN output) and the epilogue
   * applies the reference's hand-written LayerNorm backward (LayerNormFunction.backward, nafnet_arch.py:38-53):
   * out_f32 / out_bf16 = (g - xhat * mean_c(g * xhat) - mean_c(g)) * rstd + lnb_dres, g = dn * lnb_weight,
   * xhat = (lnb_x - mean) * rstd with (mean, rstd) = lnb_stats[m]; lnb_dweight[c] += sum_m dn * xhat, lnb_dbias[c] += sum_m dn,
   * lnb_colsum[c] += sum_m out (nullable).  lnb_x / lnb_dres are fp32 [M, ld_lnb]. */
  const float* lnb_x; int ld_lnb;
  const float* lnb_stats; const float* lnb_weight; const float* lnb_dres;
  float* lnb_dweight; float* lnb_dbias; float* lnb_colsum;
  /* with ln_out: 1 = Restormer's BiasFree_LayerNorm (restormer_arch.py:26-40): ln_out = bf16(v * rstd * ln_weight), rstd still from
   * the variance about the mean, ln_bias may be NULL; 0 = the centred form above. */
  int ln_nocenter;
} dcpt_gemm_desc;
int dcpt_gemm_ex(const dcpt_gemm_desc* desc, int impl, dcpt_stream_t stream);

/* conv2 (depthwise 3x3, pad 1, bias) + SimpleGate — nafnet_arch.py:96-104,:171-172,:77-80.
 * u bf16 [N,H,W,2C] -> g bf16 [N,H,W,C]; pool fp32 [N,C] += sum_px g (must be zeroed by the caller). */
int dcpt_dwconv3x3_gate_fwd(const void* u_bf16, const float* weight, const float* bias, void* g_bf16, float* pool, int N,
                            int H, int W, int C, dcpt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * NAFBlock — basicsr/archs/nafnet_arch.py:83-186 (NAFBlock.__init__/forward) and its autograd.
 * ------------------------------------------------------------------------------------------ */

/* The 18 parameters in the reference's named_parameters() order:
 *  0 beta  1 gamma  2 conv1.weight  3 conv1.bias  4 conv2.weight  5 conv2.bias  6 conv3.weight
 *  7 conv3.bias  8 sca.1.weight  9 sca.1.bias  10 conv4.weight  11 conv4.bias  12 conv5.weight
 * 13 conv5.bias  14 norm1.weight  15 norm1.bias  16 norm2.weight  17 norm2.bias                 */
#define DCPT_NAFBLOCK_NPARAMS 18

size_t dcpt_nafblock_packed_bytes(int C);                     /* bf16 weight cache              */
size_t dcpt_nafblock_saved_bytes(int N, int H, int W, int C); /* saved-for-backward arena       */
size_t dcpt_nafblock_workspace_bytes(int N, int H, int W, int C); /* backward scratch            */

/* Refresh the bf16 operand cache from the fp32 parameters (call after load_state_dict / optimizer step). */
int dcpt_nafblock_pack(const float* const* host_params, void* packed, int C, dcpt_stream_t stream);

/* x fp32 [N,H,W,C] -> out fp32 [N,H,W,C]; out_bf16 (nullable) mirrors out for a following GEMM. */
int dcpt_nafblock_fwd(const float* const* host_params, const void* packed, const float* x, float* out, void* out_bf16,
                      void* saved, int N, int H, int W, int C, dcpt_stream_t stream);

/* dout fp32 + its bf16 mirror + its column sums (fp32 [C]) -> dx fp32, dx_bf16 mirror (nullable),
 * dx_colsum (fp32 [C], accumulated; nullable); host_grads[i] accumulates d(param i). */
int dcpt_nafblock_bwd(const float* const* host_params, const void* packed, const void* saved, const float* x,
                      const float* dout, const void* dout_bf16, const float* dout_colsum, float* dx, void* dx_bf16,
                      float* dx_colsum, float* const* host_grads, void* workspace, int N, int H, int W, int C,
                      dcpt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * NAFNetBaseline — basicsr/archs/nafnet_arch.py:189-274.
 * Parameters / gradients are arrays of device pointers in named_parameters() order
 * (intro.weight, intro.bias, ending.weight, ending.bias, encoders.*, middle_blks.*, ups.*.0.weight,
 *  downs.*.{weight,bias}, decoder{i}.*), see dcpt_nafnet_num_params().
 * ------------------------------------------------------------------------------------------ */
typedef struct dcpt_nafnet_plan dcpt_nafnet_plan;

dcpt_nafnet_plan* dcpt_nafnet_create(int img_channel, int width, int middle_blk_num, const int* enc_blk_nums, int n_enc,
                                     const int* dec_blk_nums, int n_dec);
void dcpt_nafnet_destroy(dcpt_nafnet_plan* plan);
/* `NAFNet` = Local_Base + NAFNetBaseline (nafnet_arch.py:277-288): test-time local converter.  kh / kw[l] = SCA pooling kernel
 * of resolution level l (0 = full resolution .. n_enc = bottleneck), as fixed by Local_Base.convert on the train_size dummy
 * input (arch_util.py:341-347, 450-455).  A level whose map fits inside its kernel keeps the global mean (:352-353); otherwise
 * the pooled vector becomes a per-pixel replicate-padded box mean (:379-396).  Inference only; n_levels = 0 switches it off. */
int dcpt_nafnet_set_tlc(dcpt_nafnet_plan* plan, const int* kh, const int* kw, int n_levels);
/* Which block of decoder level i (0-based, < dec_blk_nums[i]) delivers host_feats[i] in dcpt_nafnet_fwd and receives
 * host_dfeats[i] in dcpt_nafnet_bwd; -1 (default) = the level's last block, i.e. the output of `decoder{i}`.  The reference's
 * DCPTModel hooks the modules whose name contains hook_names and has exactly one dot (models/
 * degradation_classification_pretrain_model.py:64-67): `decoder{i}.0`, the FIRST block, for an unwrapped NAFNetBaseline. */
int dcpt_nafnet_set_hook_blocks(dcpt_nafnet_plan* plan, const int* block_idx, int n_levels);
/* Data-parallel overlap: cuda_event (a cudaEvent_t, or NULL to turn it off) is recorded by dcpt_nafnet_bwd on its stream once
 * the gradients of encoders.{n_enc-1}.*, middle_blks.*, ups.*, decoder*.* and ending.* are final, so the caller can all-reduce
 * that slice on another stream while the shallower encoder levels are still being differentiated (what DDP's bucketed
 * all-reduce does in the reference, models/base_model.py:107-118).  Inside a stream capture the record is an external-event
 * node when external != 0 (a stream OUTSIDE the captured graph waits on it), else an ordinary captured dependency (the
 * waiting stream - e.g. the one NCCL is captured on - belongs to the same capture). */
int dcpt_nafnet_set_bwd_split_event(dcpt_nafnet_plan* plan, void* cuda_event, int external);
int dcpt_nafnet_num_params(const dcpt_nafnet_plan* plan);
/* shape of parameter i as up to 4 dims (unused dims = 1); returns number of elements */
long long dcpt_nafnet_param_shape(const dcpt_nafnet_plan* plan, int i, int dims[4]);
size_t dcpt_nafnet_packed_bytes(const dcpt_nafnet_plan* plan);
size_t dcpt_nafnet_saved_bytes(const dcpt_nafnet_plan* plan, int N, int H, int W);
size_t dcpt_nafnet_workspace_bytes(const dcpt_nafnet_plan* plan, int N, int H, int W);

int dcpt_nafnet_pack(const dcpt_nafnet_plan* plan, const float* const* host_params, void* packed, dcpt_stream_t stream);

/* inp fp32 NCHW [N,3,H,W] -> out fp32 NCHW [N,3,H,W] (out = ending(...) + inp).
 * hook != 0: skip `ending` (nafnet_arch.py:269-274), `out` may be NULL.
 * host_feats (nullable): n_dec device pointers receiving the decoder-level outputs as fp32 NHWC
 * (what DCPT's forward hooks capture, degradation_classification_pretrain_model.py:60-72). */
/* keep = 0: the following dcpt_nafnet_fwd calls are inference passes (torch.no_grad(): SRModel.test, sr_model.py:176-185) - tensors
 * that only dcpt_nafnet_bwd reads (conv4's output before the SimpleGate, 2C channels per pixel and block) are not written, so `saved`
 * must not be handed to dcpt_nafnet_bwd afterwards.  keep = 1 (default) restores the training forward.  Baked into a CUDA graph at
 * capture time like every other argument. */
int dcpt_nafnet_set_keep_activations(const dcpt_nafnet_plan* plan, int keep);
int dcpt_nafnet_fwd(const dcpt_nafnet_plan* plan, const float* const* host_params, const void* packed, const float* inp,
                    float* out, void* saved, float* const* host_feats, int hook, int N, int H, int W,
                    dcpt_stream_t stream);

/* dout fp32 NCHW (NULL when the forward ran with hook != 0); host_dfeats (nullable): gradients flowing
 * into the decoder-level outputs (fp32 NHWC, entries may be NULL).  Accumulates into host_grads. */
int dcpt_nafnet_bwd(const dcpt_nafnet_plan* plan, const float* const* host_params, const void* packed, const void* saved,
                    const float* inp, const float* dout, const float* const* host_dfeats, float* const* host_grads,
                    void* workspace, int N, int H, int W, dcpt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Degradation-classifier head building blocks — basicsr/archs/degrad_classify_arch.py
 * (PromptIR_NoImg_DC :558-641, BottleneckBlock :132-243, channels-first LayerNorm :17-44).
 * Trunk tensors are bf16 NHWC [M = N*H*W, C]; 1x1 convs use dcpt_gemm_bf16 / dcpt_gemm_ex.
 * ------------------------------------------------------------------------------------------ */

/* fp32 [O, I] parameter -> bf16 GEMM operand ([O, I], or [I, O] when transpose != 0). */
int dcpt_pack_matrix(const float* w, void* out_bf16, int O, int I, int transpose, dcpt_stream_t stream);

/* Dense 3x3 conv, stride 1, pad 1, no bias (BottleneckBlock.conv2, :178-188) as an implicit GEMM on tcgen05.
 * dcpt_conv3x3_pack: weight [Cout, Cin, 3, 3] -> bf16 operand; dgrad = 0: forward operand [Cout, 9*pad64(Cin)],
 * dgrad = 1: the flipped/transposed operand [Cin, 9*pad64(Cout)] that turns dcpt_conv3x3_fwd into the data gradient
 * (call it with x = dy and Cin/Cout swapped).  dcpt_conv3x3_packed_elems gives the operand size in elements.
 * dcpt_conv3x3_wgrad: dweight[Cout, Cin, 3, 3] += dy^T (*) x; scratch = fp32 [Cout * 9 * pad64(Cin)]. */
size_t dcpt_conv3x3_packed_elems(int Cout, int Cin, int dgrad);
int dcpt_conv3x3_pack(const float* w, void* out_bf16, int Cout, int Cin, int dgrad, dcpt_stream_t stream);
int dcpt_conv3x3_fwd(const void* x_bf16, const void* w_packed, void* out_bf16, float* out_f32, int N, int H, int W, int Cin,
                     int Cout, dcpt_stream_t stream);
int dcpt_conv3x3_wgrad(const void* dy_bf16, const void* x_bf16, float* scratch, float* dweight, int N, int H, int W, int Cin,
                       int Cout, dcpt_stream_t stream);

/* y = act(LN_c(x) * weight + bias (+ resid)), act = ReLU when relu != 0: `Conv2d.norm` + `F.relu_` (+ `out += shortcut`)
 * (:98-103, :227-243; LayerNorm channels_first :39-44, eps 1e-6).  stats fp32 [M, 2] = (mean, rstd).
 * Backward: g = dy * (y > 0); dx = LN'(g) (bf16: the next GEMM's operand), dres = g (fp32, nullable); dweight / dbias
 * accumulated.  dy and dres are fp32: LN' cancels the per-pixel mean, which would amplify a bf16 rounding of dy. */
int dcpt_ln_act_fwd(const void* x_bf16, const float* weight, const float* bias, const void* resid_bf16, void* y_bf16, float* stats,
                    int M, int C, int relu, float eps, dcpt_stream_t stream);
int dcpt_ln_act_bwd(const float* dy, const void* y_bf16, const void* x_bf16, const float* stats, const float* weight,
                    void* dx_bf16, float* dres, float* dweight, float* dbias, int M, int C, int relu, dcpt_stream_t stream);

/* z = prev + mw[0] * feat (prev nullable; feat fp32, n elements): `lq_feats + mixing_weights[i] * feature` (:637).
 * Backward: dfeat = mw * dz (nullable), dmw[0] += sum dz * feat. */
int dcpt_mix_fwd(const void* prev_bf16, const float* feat, const float* mw, void* z_bf16, long long n, dcpt_stream_t stream);
int dcpt_mix_bwd(const float* dz, const float* feat, const float* mw, float* dfeat, float* dmw, long long n, dcpt_stream_t stream);

/* y = relu(maxpool2x2(x)): `nn.MaxPool2d(2, 2)`, `nn.ReLU()` of downsample_layers (:596-602). x is [N, 2Ho, 2Wo, C]. */
int dcpt_maxpool2_relu_fwd(const void* x_bf16, void* y_bf16, int N, int Ho, int Wo, int C, dcpt_stream_t stream);
int dcpt_maxpool2_relu_bwd(const void* x_bf16, const float* dy, void* dx_bf16, int N, int Ho, int Wo, int C, dcpt_stream_t stream);

/* logits = fc(mean over H,W of x): `.mean(dim=[-1, -2])` + `self.fc` (:639-640). pooled fp32 [N, C], logits fp32 [N, K]. */
int dcpt_meanpool_fc_fwd(const void* x_bf16, const float* weight, const float* bias, float* pooled, float* logits, int N, int HW,
                         int C, int K, dcpt_stream_t stream);
int dcpt_meanpool_fc_bwd(const float* dlogits, const float* pooled, const float* weight, float* dweight, float* dbias, float* dx,
                         int N, int HW, int C, int K, dcpt_stream_t stream);

/* out = a + b (bf16, n elements): sums the two gradient paths of a residual block. */
/* Patch matrix of `conv_embed` = Conv2d(3, f0, 7, stride 2, padding 3) of PromptIR_DC (degrad_classify_arch.py:497-500):
 * img fp32 NCHW [N,3,H,W] -> patches bf16 [N*Ho*Wo, 160], Ho = (H-1)/2+1, column ci*49 + ky*7 + kx (= the layout of
 * weight.view(f0, 147)), columns 147..159 zero; the convolution is then dcpt_gemm_bf16 against the weight padded to 160. */
int dcpt_im2col7x7s2(const float* img, void* patches_bf16, int N, int H, int W, dcpt_stream_t stream);
int dcpt_add_bf16(const void* a, const void* b, void* out, long long n, dcpt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Restormer — basicsr/archs/restormer_arch.py (TransformerBlock :148-159 = MDTA :103-145 + GDFN :75-100,
 * Downsample/Upsample :175-202, Restormer.forward :376-422).  Forward (inference) path.
 * ------------------------------------------------------------------------------------------ */

/* Channel LayerNorm of the Restormer blocks (restormer_arch.py:26-72) on NHWC rows: x fp32 [M,C] -> out bf16 [M,C].
 * bias == NULL and center == 0: BiasFree_LayerNorm (:38-40), x / sqrt(var + 1e-6) * weight with var about the mean;
 * center != 0: WithBias_LayerNorm (:56-59).  stats (nullable) receives (mean, rstd) per row. */
int dcpt_layernorm_rows_fwd(const float* x, const float* weight, const float* bias, void* out_bf16, float* stats, int M, int C,
                            float eps, int center, dcpt_stream_t stream);

/* Depthwise 3x3, stride 1, zero pad 1, no bias (Attention.qkv_dwconv, restormer_arch.py:110-118) on bf16 NHWC
 * [N,H,W,CH]; weight fp32 [CH,1,3,3].  sumsq (nullable) fp32 [N, sq_ch] += sum over pixels of out^2 for the first
 * sq_ch channels: the squared norms F.normalize(q / k, dim=-1) divides by (:131-132). */
int dcpt_dwconv3x3_fwd(const void* x_bf16, const float* weight, void* out_bf16, float* sumsq, int sq_ch, int N, int H, int W, int CH,
                       dcpt_stream_t stream);

/* FeedForward gate (restormer_arch.py:97-98): g = gelu(dw3x3(u)[:, :C]) * dw3x3(u)[:, C:], exact erf GELU, no bias.
 * u bf16 [N,H,W,2C], weight fp32 [2C,1,3,3], g bf16 [N,H,W,C]. */
int dcpt_dwconv3x3_gelu_gate_fwd(const void* u_bf16, const float* weight, void* g_bf16, int N, int H, int W, int C,
                                 dcpt_stream_t stream);

typedef struct dcpt_restormer_plan dcpt_restormer_plan;

/* Restormer.__init__ (restormer_arch.py:236-250); num_blocks and heads have 4 entries.  inp/out channels must be 3,
 * scale 1, dual_pixel_task False.  Parameters travel in the module's named_parameters() order. */
dcpt_restormer_plan* dcpt_restormer_create(int inp_channels, int out_channels, int dim, const int* num_blocks,
                                           int num_refinement_blocks, const int* heads, double ffn_expansion_factor, int bias,
                                           int ln_with_bias);
void dcpt_restormer_destroy(dcpt_restormer_plan* plan);
/* Activation of the transposed attention map of every block of the plan: 0 (default) relu(attn) as in the fork's Restormer
 * (restormer_arch.py:135-136: the softmax is commented out), 1 attn.softmax(dim=-1) as in the PromptIR transformer blocks
 * (promptir_arch.py:108-149, softmax at :140) and the original Restormer.  Forward and backward. */
int dcpt_restormer_set_attention(dcpt_restormer_plan* plan, int softmax);
int dcpt_restormer_num_params(const dcpt_restormer_plan* plan);
long long dcpt_restormer_param_shape(const dcpt_restormer_plan* plan, int i, int dims[4]);
size_t dcpt_restormer_packed_bytes(const dcpt_restormer_plan* plan);
size_t dcpt_restormer_workspace_bytes(const dcpt_restormer_plan* plan, int N, int H, int W);
int dcpt_restormer_pack(const dcpt_restormer_plan* plan, const float* const* host_params, void* packed, dcpt_stream_t stream);

/* Restormer.forward(inp_img, hook): inp fp32 NCHW [N,3,H,W] (H, W multiples of 8) -> out fp32 NCHW = output(...) + inp.
 * hook != 0: stop after decoder_level1 (:403), `out` may be NULL.  host_feats (nullable): 3 device pointers receiving
 * decoder_level3 / 2 / 1 outputs as fp32 NHWC ([N,H/4,W/4,4dim], [N,H/2,W/2,2dim], [N,H,W,2dim]) - what DCPT's hooks on
 * `decoder_level{k}.body` capture (degradation_classification_pretrain_model.py:60-68). */
int dcpt_restormer_fwd(const dcpt_restormer_plan* plan, const float* const* host_params, const void* packed, const float* inp,
                       float* out, void* workspace, float* const* host_feats, int hook, int N, int H, int W,
                       dcpt_stream_t stream);

/* One TransformerBlock (restormer_arch.py:156-159) of the plan, in place on x fp32 NHWC [N,H,W,d] where d is the width of
 * block j of `stage` (0..7 = encoder_level1, 2, 3, latent, decoder_level3, 2, 1, refinement).  `workspace` must be at least
 * dcpt_restormer_workspace_bytes(plan, N, H << l, W << l), l = the stage's resolution level (0,1,2,3,2,1,0,0).  Used by the
 * parity tests against the reference's TransformerBlock. */
int dcpt_restormer_block_fwd(const dcpt_restormer_plan* plan, int stage, int j, const float* const* host_params, const void* packed,
                             float* x, void* workspace, int N, int H, int W, dcpt_stream_t stream);

/* Training form of one TransformerBlock: forward that keeps its intermediates in `saved`
 * (dcpt_restormer_block_saved_bytes), and the backward through MDTA (F.normalize, ReLU attention, temperature,
 * qkv / qkv_dwconv / project_out) and GDFN (GELU gate, dwconv, project_in / project_out) and both LayerNorms -
 * what autograd does for restormer_arch.py:156-159.  x, xout, dout, dx: fp32 NHWC [N,H,W,d].  host_grads: one device
 * pointer per plan parameter (named_parameters() order; only the block's own entries are touched, accumulated +=). */
size_t dcpt_restormer_block_saved_bytes(const dcpt_restormer_plan* plan, int stage, int j, int N, int H, int W);
size_t dcpt_restormer_block_workspace_bytes(const dcpt_restormer_plan* plan, int stage, int j, int N, int H, int W);
int dcpt_restormer_block_fwd_train(const dcpt_restormer_plan* plan, int stage, int j, const float* const* host_params,
                                   const void* packed, const float* x, float* xout, void* saved, int N, int H, int W,
                                   dcpt_stream_t stream);
int dcpt_restormer_block_bwd(const dcpt_restormer_plan* plan, int stage, int j, const float* const* host_params, const void* packed,
                             const void* saved, const float* x, const float* dout, float* dx, float* const* host_grads,
                             void* workspace, int N, int H, int W, dcpt_stream_t stream);

/* Training path of the whole Restormer (what autograd does for Restormer.forward, restormer_arch.py:376-422):
 * dcpt_restormer_fwd_train keeps every block's intermediates in `saved` (dcpt_restormer_saved_bytes; `workspace` as for
 * dcpt_restormer_fwd), dcpt_restormer_bwd turns dout = d(loss)/d(out) (fp32 NCHW) into the gradients of all parameters
 * (accumulated += into host_grads, named_parameters() order; workspace: dcpt_restormer_bwd_workspace_bytes).
 * DCPT pretraining (models/degradation_classification_pretrain_model.py:60-68, 133-169): host_feats (may be NULL) is a
 * HOST array of 3 device pointers (entries may be NULL) receiving the outputs of decoder_level3, 2, 1 as fp32 NHWC
 * ([N,H/4,W/4,4dim], [N,H/2,W/2,2dim], [N,H,W,2dim]) - what the reference's forward hooks collect; hook != 0 stops after
 * decoder_level1 (restormer_arch.py:403; `out` may then be NULL).  dcpt_restormer_bwd takes the classifier's gradients
 * w.r.t. those features in dfeats (same layout; NULL or NULL entries = none) and adds them where the features were
 * taken; dout == NULL is the backward of a hook pass (refinement / output conv gradients are left untouched). */
size_t dcpt_restormer_saved_bytes(const dcpt_restormer_plan* plan, int N, int H, int W);
size_t dcpt_restormer_bwd_workspace_bytes(const dcpt_restormer_plan* plan, int N, int H, int W);
int dcpt_restormer_fwd_train(const dcpt_restormer_plan* plan, const float* const* host_params, const void* packed, const float* inp,
                             float* out, void* saved, void* workspace, float* const* host_feats, int hook, int N, int H, int W,
                             dcpt_stream_t stream);
int dcpt_restormer_bwd(const dcpt_restormer_plan* plan, const float* const* host_params, const void* packed, const void* saved,
                       const float* inp, const float* dout, const float* const* dfeats, float* const* host_grads, void* workspace,
                       int N, int H, int W, dcpt_stream_t stream);

/* PromptIR (basicsr/archs/promptir_arch.py:267-518), inference: the Restormer trunk with softmax attention (:140), three
 * PromptGenBlocks (:238-263: global mean -> Linear -> softmax -> weighted sum of the prompt components -> bilinear resize
 * -> 3x3 conv) concatenated to the decoder stream, a "noise" TransformerBlock and a 1x1 reduce conv after each (:478-505).
 * The plan is a dcpt_restormer_plan (destroy / num_params / param_shape as above; parameters in PromptIR.named_parameters()
 * order, `prompt_param` [1,5,D,S,S] reported as [5*D, S, S, 1]); dim must be 48 (the prompt widths are literals in the
 * reference), LayerNorm "WithBias" = ln_with_bias 1 is the reference's default.  Replaces PromptIR.forward(inp_img) (:465-518)
 * as called by SRModel.test (sr_model.py:176-185) for options/all_in_one/test/test_PromptIR_5d.yml. */
dcpt_restormer_plan* dcpt_promptir_create(int inp_channels, int out_channels, int dim, const int* num_blocks,
                                          int num_refinement_blocks, const int* heads, double ffn_expansion_factor, int bias,
                                          int ln_with_bias);
size_t dcpt_promptir_packed_bytes(const dcpt_restormer_plan* plan);
size_t dcpt_promptir_workspace_bytes(const dcpt_restormer_plan* plan, int N, int H, int W);
int dcpt_promptir_pack(const dcpt_restormer_plan* plan, const float* const* host_params, void* packed, dcpt_stream_t stream);
int dcpt_promptir_fwd(const dcpt_restormer_plan* plan, const float* const* host_params, const void* packed, const float* inp,
                      float* out, void* workspace, int N, int H, int W, dcpt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Parameter update of the training step — SURVEY.md §8(f) row 1.  Replaces, in SRModel.optimize_parameters
 * (basicsr/models/sr_model.py:164-174): torch.nn.utils.clip_grad_norm_(net_g.parameters(), grad_clip) (:166-167),
 * optimizer_g.step() (:169; torch.optim.Adam / AdamW built by BaseModel.get_optimizer, base_model.py:120-139) and
 * BaseModel.model_ema (base_model.py:86-95: ema.mul_(decay).add_(param, alpha=1-decay) for each of the 664 tensors).
 * All tensors fp32; one plan covers a list of n_tensors parameter tensors (a param group) through a chunk table, so the
 * whole update is two multi-tensor launches + one tiny reduction regardless of the number of tensors.
 *   dcpt_optim_create          host_numels[n_tensors] (HOST array).  NULL on error.
 *   dcpt_optim_workspace_bytes device workspace (pointer table, chunk table, per-chunk partial sums, total_norm).
 *   dcpt_optim_bind            uploads the pointer tables (HOST arrays of DEVICE pointers: params, grads, exp_avg,
 *                              exp_avg_sq, ema; host_ema may be NULL) into the workspace.  H2D copies: not graph-capturable;
 *                              call again whenever a pointer changes (e.g. autograd handed out new .grad tensors).
 *   dcpt_optim_grad_norm       total_norm = ||all gradients||_2 (deterministic two-level reduction) kept in the workspace
 *                              for dcpt_optim_step and, if total_norm != NULL, written there (clip_grad_norm_'s return value).
 *   dcpt_optim_step            g' = g * min(1, max_norm / (total_norm + 1e-6)) when max_norm > 0 (needs a preceding
 *                              dcpt_optim_grad_norm; gradients themselves are NOT rewritten), then torch's Adam
 *                              (decoupled_weight_decay == 0: g' += wd * p) or AdamW (p *= 1 - lr * wd) update in torch's
 *                              operation order with bias corrections for `step` (the 1-based count of this update), then,
 *                              when ema_decay > 0 and an ema pointer was bound, ema = ema * decay + (1 - decay) * p.
 *                              amsgrad / maximize are not supported.
 *   dcpt_optim_set_norm        overrides the workspace's total_norm with *total_norm (DEVICE pointer): clip_grad_norm_ over
 *                              SEVERAL plans (param groups / step counts) = sqrt(sum of the per-plan norms squared), which
 *                              the caller combines and hands to each plan before its dcpt_optim_step.
 *   dcpt_optim_param_hash      *hash (DEVICE u64) = order-independent fingerprint of the bound PARAMETER tensors' bits; lets the
 *                              host notice writes that bypass torch's version counter (BaseModel.model_ema's
 *                              `ema.data.mul_().add_()`, base_model.py:86-95) before re-using a packed bf16 operand cache. */
typedef struct dcpt_optim_plan dcpt_optim_plan;
dcpt_optim_plan* dcpt_optim_create(const long long* host_numels, int n_tensors);
void dcpt_optim_destroy(dcpt_optim_plan* plan);
size_t dcpt_optim_workspace_bytes(const dcpt_optim_plan* plan);
long long dcpt_optim_num_chunks(const dcpt_optim_plan* plan);
int dcpt_optim_bind(const dcpt_optim_plan* plan, void* workspace, float* const* host_params, const float* const* host_grads,
                    float* const* host_exp_avg, float* const* host_exp_avg_sq, float* const* host_ema, dcpt_stream_t stream);
int dcpt_optim_grad_norm(const dcpt_optim_plan* plan, void* workspace, float* total_norm, dcpt_stream_t stream);
int dcpt_optim_set_norm(const dcpt_optim_plan* plan, void* workspace, const float* total_norm, dcpt_stream_t stream);
int dcpt_optim_param_hash(const dcpt_optim_plan* plan, void* workspace, unsigned long long* hash, dcpt_stream_t stream);
int dcpt_optim_step(const dcpt_optim_plan* plan, void* workspace, int decoupled_weight_decay, double lr, double beta1, double beta2,
                    double eps, double weight_decay, long long step, double max_norm, double ema_decay, dcpt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DCPT_OPS_H_ */
