// Memory-bound network kernels that sit between the implicit-GEMM convolutions
// (declarations; definitions in nn_kernels.cu).
#pragma once
#include "common.cuh"

// Stem: conv 7x7 stride 2 pad 3 (+bias, ReLU), PyTorch-LineInference.ipynb:L233-L236.
// in: u8 [N][H][W] (cin = 1, value * in_scale) or fp32 [N][H][W][3] (cin = 3, value * in_scale)
// w: fp32 [49*cin][64] (tap-major), out: bf16 NHWC [N][H/2][W/2][64]
int stem_conv_launch(bdrc_ctx* ctx, const void* in, int cin, float in_scale, int N, int H, int W,
                     const float* w_packed, const float* bias, bf16* out, cudaStream_t s);

// MaxPool 3x3 stride 2 pad 1 (ipynb:L237-L238), bf16 NHWC, C % 8 == 0
int maxpool3x3s2_launch(bdrc_ctx* ctx, const bf16* in, int N, int H, int W, int C, bf16* out,
                        cudaStream_t s);

// Depthwise 3x3, dilation d, pad d, no bias (ipynb:L411-L420, L445, L455).  w: fp32 [9][C]
int depthwise3x3_launch(bdrc_ctx* ctx, const bf16* in, int in_ld, int in_coff, int N, int H, int W,
                        int C, int dil, const float* w9c, bf16* out, int out_ld, int out_coff,
                        cudaStream_t s);

// Bilinear upsample x`f`, align_corners=True (ipynb:L448-L452), bf16 NHWC -> bf16 NHWC slice
int upsample_bilinear_ac_launch(bdrc_ctx* ctx, const bf16* in, int N, int H, int W, int C, int f,
                                bf16* out, int out_ld, int out_coff, cudaStream_t s);

// Head upsample x4 align_corners=True: fp32 NHWC low-res [N][h][w][C] -> fp32 NCHW [N][C][4h][4w]
int upsample_logits_launch(bdrc_ctx* ctx, const float* in, int N, int h, int w, int C, float* out,
                           cudaStream_t s);

// ASPP image-pooling branch folded into a per-tile bias of the projection conv
// (ipynb:L421-L440; SURVEY a-2 note):
//   g = mean_hw(x) [Cin]; h = relu(Wp g + bp) [Cmid]; bias_eff = b_proj + Wtail h  [Cout_pad]
int aspp_pool_bias_launch(bdrc_ctx* ctx, const bf16* x, int N, int HW, int Cin, const float* Wp,
                          const float* bp, int Cmid, const float* Wtail, const float* bproj,
                          int Cout, int Cout_pad, float* bias_eff, const float* mean_in, cudaStream_t s);
// mean_in (optional): precomputed g [N][Cin] (aspp_dw3_launch produces it while the map is in shared memory)

// Easter2 squeeze-excite over time: data = relu(res + d * sigmoid(W2 relu(W1 mean_T(d) + b1) + b2));
// res rows have stride res_ld, out rows stride out_ld (channel slices of wider buffers).
// two short parallel launches (nn_kernels.cu: per-chunk column sums + the gate by the line's last CTA, then the
// elementwise apply).  scratch: se_fused_scratch_floats(L, C) fp32 words of the caller's workspace whose ticket
// counters se_fused_reset zeroes once per forward.
int se_fused_launch(bdrc_ctx* ctx, const bf16* res, int res_ld, const bf16* d, int L, int T, int C, const float* W1,
                    const float* b1, int Cmid, const float* W2, const float* b2, bf16* out, int out_ld, float* scratch,
                    int pdl, cudaStream_t s);
size_t se_fused_scratch_floats(int L, int C);
int se_fused_reset(bdrc_ctx* ctx, float* scratch, int L, int C, cudaStream_t s);
int se_fused_chunks(void);
// Stem + max-pool fused on the tensor cores (stem_tc.cu): u8 tiles [N][512][512] (3 identical channels
// folded) -> bf16 NHWC [N][128][128][64].  w_bf16: [64 cout][64 k] with k = 8 * kernel_row + kernel_col.
int stem_tc_pool_launch(bdrc_ctx* ctx, const uint8_t* tiles, int N, const bf16* w_bf16, const float* bias, bf16* out,
                        cudaStream_t s);

// The three dilated depthwise 3x3 ASPP branches (32x32 maps) in one launch: x [N][32][32][C] -> out[i] (same shape),
// rate rates[i], weights w9c[i] fp32 [9][C].
int aspp_dw3_launch(bdrc_ctx* ctx, const bf16* x, int N, int C, const int rates[3], const float* const w9c[3],
                    bf16* const out[3], float* mean_out, cudaStream_t s);

// Decoder tail in one kernel (dec_tail.cu): depthwise 3x3 (C channels, pad 1) -> 1x1 conv C -> 256 + bias + ReLU -> head
// 1x1 256 -> head_C + bias.  x: bf16 NHWC [N][H][W][C]; pw: the packed 1x1 weights; head_w_f32: [head_C][256] fp32 holding
// the bf16-rounded head weights; out: fp32 NHWC [N][H][W][head_C].
int dec_tail_supported(int C, int Cout, int head_C, int H, int W);
int dec_tail_launch(bdrc_ctx* ctx, const bf16* x, int N, int H, int W, int C, const float* dw_w9c, const ConvWeights& pw,
                    const float* bias_host, const float* head_w_host, const float* head_b_host, int head_C, float* out,
                    cudaStream_t s);
