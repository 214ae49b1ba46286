// Launchers for the CUDA-core (HBM-bound) kernels of the NAFNet hot path.
// Layout everywhere: NHWC, i.e. row-major [M = N*H*W pixels, C channels].
// fp32 = residual stream / statistics / parameters / gradients of parameters,
// bf16 = branch-internal activations that feed the tensor-core GEMMs.
#pragma once
#include "common.cuh"

// LayerNorm2d over channels (reference: nafnet_arch.py:25-64).  b and stats are nullable; center = 0 is Restormer's
// BiasFree_LayerNorm (restormer_arch.py:38-40: x / sqrt(var + eps) * w with the variance taken about the mean).
int ln_fwd_launch(const float* x, const float* w, const float* b, bf16* n_out, float* stats, int M, int C, float eps,
                  cudaStream_t st, int center = 1);
// dx = dres + LNbwd(dn); also dx mirror in bf16, sum_m dn*yhat -> dw, sum_m dn -> db, sum_m dx -> colsum (all +=).
int ln_bwd_launch(const bf16* dn, const float* x, const float* stats, const float* w, const float* dres, float* dx,
                  bf16* dx_bf16, float* dw, float* db, float* colsum, int M, int C, cudaStream_t st, int center = 1);

// depthwise 3x3 (+bias, zero pad) on 2C channels followed by SimpleGate; pool[n,c] += sum_px g.
int dwgate_fwd_launch(const bf16* u, const float* w2, const float* b2, bf16* g, float* pool, int N, int H, int W, int C,
                      cudaStream_t st);
// Restormer GDFN gate (restormer_arch.py:97-98): g = gelu(dw3x3(u)[:, :C]) * dw3x3(u)[:, C:], no bias; w2 is [2C][9].
int dwgelu_fwd_launch(const bf16* u, const float* w2, bf16* g, int N, int H, int W, int C, cudaStream_t st);
// plain depthwise 3x3 (no bias) on CH channels; sumsq[n][c] += sum_px out^2 for c < sq_ch (nullable).
int dwconv3_fwd_launch(const bf16* x, const float* w, bf16* out, float* sumsq, int sq_ch, int N, int H, int W, int CH,
                       cudaStream_t st);
// backward part a: dg = dgs*s + t; du2 = SimpleGate'(dg); dW2 += du2 (*) u; db2 += sum du2.
// t = the SCA backward's per-image shift (sca_bwd_launch), or nullptr with ds [N, C] + the SCA weight [C, C]: the kernel then does
// that transposed mat-vec itself for the channels it owns
int dwgate_bwd_a_launch(const bf16* dgs, const float* s, const float* t, const bf16* u, const float* w2, const float* b2,
                        bf16* du2, float* dw2, float* db2, int N, int H, int W, int C, cudaStream_t st, const float* ds = nullptr,
                        const float* w_sca = nullptr);
int dwgelu_bwd_a_launch(const bf16* dg, const bf16* u, const float* w2, bf16* du2, float* dw2, int N, int H, int W, int C, cudaStream_t st);
int dwconv3_wgrad_launch(const bf16* dy, const bf16* x, float* dw, int N, int H, int W, int CH, cudaStream_t st, float* db = nullptr);
// backward part b: du = dwconv^T(du2); colsum[c] += sum_px du.
int dwconv_bwd_data_launch(const bf16* du2, const float* w2, bf16* du, float* colsum, int N, int H, int W, int C2,
                           cudaStream_t st);

// Simplified channel attention (nafnet_arch.py:116-127).
int sca_fwd_launch(const float* pool, const float* w, const float* b, float* s, int N, int C, int HW, cudaStream_t st);
int scale_rows_launch(const bf16* g, const float* s, bf16* gs, int N, int HW, int C, cudaStream_t st);
// both in one launch: s as sca_fwd_launch (bit-identical), gs = g * s
int sca_scale_launch(const float* pool, const float* w, const float* b, const bf16* g, float* s, bf16* gs, int N, int C, int HW,
                     cudaStream_t st);
int sca_ds_reduce_launch(const bf16* dgs, const bf16* g, float* ds, int N, int HW, int C, cudaStream_t st);
int sca_bwd_launch(const float* ds, const float* pool, const float* w, float* t, float* dw, float* db, int N, int C, int HW,
                   cudaStream_t st, cudaStream_t st_w = nullptr);

// TLC (test-time local converter, arch_util.py:339-398): P = replicate-padded k1 x k2 box mean of g; I = fp32 scratch [N,H,W,C]
int tlc_boxmean_launch(const bf16* g, float* I, bf16* P, int N, int H, int W, int C, int k1, int k2, cudaStream_t st);
int mul_bf16_launch(const bf16* a, const bf16* b, bf16* out, long long n, cudaStream_t st);
int colsum_bf16_launch(const bf16* x, float* out, int M, int C, cudaStream_t st);
// out = a + b (either nullable -> treated as 0): fp32 (nullable), bf16 mirror (nullable), colsum += column sums (nullable).
int grad_prepare_launch(const float* a, const float* b, float* out, bf16* out_bf16, float* colsum, int M, int C,
                        cudaStream_t st);
int axpy_launch(float* dst, const float* src, int n, cudaStream_t st);
int cast_f32_bf16_launch(const float* x, bf16* y, long long n, cudaStream_t st);
// fp32 NHWC [N, 2H, 2W, C] -> bf16 [N*H*W, 4C], column = (i*2+j)*C + c for sub-pixel (i, j).
int unshuffle_cast_launch(const float* x, bf16* y, int N, int H, int W, int C, cudaStream_t st);

// Weight packing (fp32 parameter -> bf16 GEMM operand), see nafnet.cu for the modes.
enum PackMode { PACK_PLAIN = 0, PACK_T = 1, PACK_PAIR = 2, PACK_UP = 3, PACK_UP_T = 4, PACK_DOWN = 5, PACK_DOWN_T = 6, PACK_PAIR32 = 7 };
int pack_weight_launch(const float* w, const float* row_scale, bf16* out, int O, int I, int mode, cudaStream_t st);
// small fp32 vectors: out[p] = bias[src(p)] * scale[src(p)] (mode PACK_PLAIN or PACK_PAIR)
int pack_bias_launch(const float* bias, const float* scale, float* out, int O, int mode, cudaStream_t st);

// wgrad finishing: scratch G (fp32, from the split-K GEMM) -> parameter gradients (+=).
enum FinishMode { FIN_RESID = 0, FIN_UP = 1, FIN_DOWN = 2, FIN_C3_TO_CN = 3, FIN_CN_TO_C3 = 4 };
int wgrad_finish_resid_launch(const float* G, const float* w, const float* bias, const float* scale, const float* colsum,
                              float* dw, float* dbias, float* dscale, int O, int I, cudaStream_t st);
int wgrad_finish_perm_launch(const float* G, float* dw, int O, int I, int mode, cudaStream_t st);

// 3x3 convs touching the 3-channel image (intro / ending), direct CUDA-core kernels.
// img: fp32 NCHW [N,3,H,W]; feat: fp32 NHWC [N*H*W, C].
int conv3x3_img_to_feat_launch(const float* img, const float* w, const float* bias, int transpose_flip, float* out_f32,
                               bf16* out_bf16, float* colsum, int N, int H, int W, int C, cudaStream_t st);
int conv3x3_feat_to_img_launch(const float* feat, const float* w, const float* bias, const float* resid_img, float* out_img,
                               int N, int H, int W, int C, cudaStream_t st);
// G[c][j] (+)= sum_px feat[px][c] * patch_j(img)[px], j = ci*9 + ky*3 + kx (27 columns); flip mirrors the taps.
int conv3x3_small_wgrad_launch(const float* feat, const float* img, float* G, float* img_sum, int flip, int N, int H, int W,
                               int C, cudaStream_t st);

// Tensor-core route for the image-side 3x3 convs (conv3x3_img.cu, second half): bf16 patch matrix P [pixels, 32].
int im2col3_launch(const float* img, bf16* P, float* colsum32, int flip, int dup, int N, int H, int W, cudaStream_t st);
int pack_w27_launch(const float* w, bf16* out, int C, int mode, cudaStream_t st);  // mode 0 intro [C][27], 1 ending [3][C][9]
int finish_w27_launch(const float* G, float* dw, const float* psum, float* db, int C, int mode, cudaStream_t st);
int ending_colsum_launch(const float* w, const float* psum, float* colsum, int C, cudaStream_t st);
int ending_fwd_tma_launch(const bf16* feat, const float* w, const float* bias, const float* resid_img, float* out_img, int N, int H,
                          int W, int C, cudaStream_t st);

// Degradation-classifier head glue kernels (dchead.cu), bf16 NHWC trunk.
int ln_act_fwd_launch(const bf16* x, const float* w, const float* b, const bf16* resid, bf16* y, float* stats, int M, int C, int relu,
                      float eps, cudaStream_t st);
int ln_act_bwd_launch(const float* dy, const bf16* y, const bf16* x, const float* stats, const float* w, bf16* dx, float* dres,
                      float* dw, float* db, int M, int C, int relu, cudaStream_t st);
int mix_fwd_launch(const bf16* prev, const float* feat, const float* mw, bf16* z, long long n, cudaStream_t st);
int mix_bwd_launch(const float* dz, const float* feat, const float* mw, float* dfeat, float* dmw, long long n, cudaStream_t st);
int maxpool2_relu_fwd_launch(const bf16* x, bf16* y, int N, int Ho, int Wo, int C, cudaStream_t st);
int maxpool2_relu_bwd_launch(const bf16* x, const float* dy, bf16* dx, int N, int Ho, int Wo, int C, cudaStream_t st);
int meanpool_fc_fwd_launch(const bf16* x, const float* w, const float* b, float* pooled, float* logits, int N, int HW, int C, int K,
                           cudaStream_t st);
int meanpool_fc_bwd_launch(const float* dlogits, const float* pooled, const float* w, float* dw, float* db, float* dx, int N, int HW,
                           int C, int K, cudaStream_t st);
int add_bf16_launch(const bf16* a, const bf16* b, bf16* out, long long n, cudaStream_t st);
int pack_conv3x3_launch(const float* w, bf16* out, int Cout, int Cin, int dgrad, cudaStream_t st);
int finish_conv3x3_launch(const float* G, float* dw, int Cout, int Cin, cudaStream_t st);
int im2col7s2_launch(const float* img, bf16* P, int N, int H, int W, cudaStream_t st);  // PromptIR_DC conv_embed patches [M, 160]
